#!/bin/bash
# A/B of the evaluation pixel pass on one B200: eval-only workload with the in-tree library and
# with instrumented / re-parameterised builds (build/<name>/libnicr_panoptic_b200.so), interleaved
set -u
VARIANTS=${VARIANTS:-"tree v5 v6"}
for rep in 1 2; do
for v in $VARIANTS; do
  if [ $v = tree ]; then unset NPB_LIB_PATH; else export NPB_LIB_PATH=$PWD/build/$v/libnicr_panoptic_b200.so; fi
  timeout 300 python scripts/bench_eval.py --frames 25600 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('eval $v', round(d['value']), round(d['roofline']['frac'],3), d['quality'])"
done
done
