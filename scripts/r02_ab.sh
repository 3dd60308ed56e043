#!/bin/bash
# A/B on one B200: bench lines of build/base (an older revision, scripts/build_rev.sh) and the tree, interleaved
set -u
CFGS=${CFGS:-"nyuv2 sunrgbd"}
for rep in 1 2; do
for cfg in $CFGS; do
  for lib in base new; do
    if [ $lib = base ]; then export NPB_LIB_PATH=$PWD/build/base/libnicr_panoptic_b200.so; PIPE=--no-pipeline; else unset NPB_LIB_PATH; PIPE=""; fi
    timeout 300 python bench.py --config $cfg --steps 300 --warmup 5 --no-e2e --no-cpu-baseline --no-api --no-extra $PIPE 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$cfg $lib', round(d['value']), 'frames/s', round(d['ms_per_step']*1e3,1),'us path', round(d['roofline_path']['frac'],3), 'kernel', round(d['roofline']['frac'],3), round(d['roofline']['kernel_ms']*1e3,1), 'us', d['clocks']['sm_mhz'], d['clocks']['reasons'])"
  done
done
done
