#!/bin/bash
# one optimisation iteration on a B200: parity suite, bench lines (nyuv2, sunrgbd), in-situ timeline
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -4
for cfg in nyuv2 sunrgbd; do
  timeout 300 python bench.py --config $cfg --steps 300 --warmup 5 --no-e2e --no-cpu-baseline --no-api --no-extra 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$cfg', round(d['value']), 'frames/s', round(d['ms_per_step']*1e3,1),'us path', round(d['roofline_path']['frac'],3), 'kernel', round(d['roofline']['frac'],3), d['quality'], d['clocks']['sm_mhz'], d['clocks']['reasons'])"
done
NPB_LIB_PATH=$PWD/build/timeline/libnicr_panoptic_b200.so timeout 300 python scripts/probes/timeline.py --config nyuv2 2>&1 | tail -11
