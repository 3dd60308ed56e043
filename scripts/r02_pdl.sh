#!/bin/bash
set -u
mkdir -p gpurun_out
python scripts/probes/pdl_debug.py 2>&1 | grep -E "build info|bad iterations|iter .* diffs" | head -12
NPB_NO_PDL=1 python scripts/probes/pdl_debug.py 2>&1 | grep -E "build info|bad iterations|iter .* diffs" | head -12
NPB_LIB_PATH=$PWD/build/r01/libnicr_panoptic_b200.so python scripts/probes/pdl_debug.py 2>&1 | grep -E "build info|bad iterations|iter .* diffs" | head -12
python -m pytest tests -q -m gpu -x 2>&1 | tail -5
for pdl in 0 1; do
  NPB_NO_PDL=$pdl python bench.py --config nyuv2 --steps 300 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('nyuv2 NO_PDL=$pdl', round(d['value']), round(d['ms_per_step']*1e3,1),'us')"
done
