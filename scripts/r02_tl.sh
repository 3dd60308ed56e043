#!/bin/bash
set -u
python -m pytest tests -q -m gpu -x 2>&1 | tail -4
for pdl in 0 1; do
  NPB_NO_PDL=$pdl python bench.py --config nyuv2 --steps 300 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('nyuv2 NO_PDL=$pdl', round(d['value']), round(d['ms_per_step']*1e3,1),'us', d['quality'])"
done
python bench.py --config sunrgbd --steps 300 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('sunrgbd', round(d['value']), round(d['ms_per_step']*1e3,1),'us', d['quality'], d['clocks'])"
export NPB_LIB_PATH=$PWD/build/timeline/libnicr_panoptic_b200.so
python scripts/probes/timeline.py --config nyuv2 2>&1 | tail -12
