#!/bin/bash
# A/B of whole steps on one B200: bench lines with the in-tree library and with other builds of
# the same ABI (build/<name>/libnicr_panoptic_b200.so), interleaved
set -u
VARIANTS=${VARIANTS:-"tree v3"}
CFGS=${CFGS:-"nyuv2 sunrgbd"}
for rep in 1 2; do
for cfg in $CFGS; do
  for v in $VARIANTS; do
    if [ $v = tree ]; then unset NPB_LIB_PATH; else export NPB_LIB_PATH=$PWD/build/$v/libnicr_panoptic_b200.so; fi
    timeout 300 python bench.py --config $cfg --steps 300 --warmup 5 --no-e2e --no-cpu-baseline --no-api --no-extra 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$cfg $v', round(d['value']), 'frames/s', round(d['ms_per_step']*1e3,1),'us path', round(d['roofline_path']['frac'],3), 'kernel', round(d['roofline']['frac'],3), d['clocks']['sm_mhz'], d['clocks']['reasons'])"
  done
done
done
