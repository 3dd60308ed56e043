#!/bin/bash
# state of the tree on one B200: parity suite, PDL race probe, bench lines of the two main shapes
set -u
mkdir -p gpurun_out
O=gpurun_out/r02c
python -m pytest tests -q -m gpu -x 2>&1 | tail -6 > ${O}_tests.log; tail -3 ${O}_tests.log
python scripts/probes/pdl_debug.py 2>&1 | grep -E "build info|bad iterations|iter .* diffs" | head -8
for cfg in nyuv2 sunrgbd; do
  python bench.py --config $cfg --steps 300 --warmup 5 --no-e2e --no-cpu-baseline > ${O}_${cfg}.json 2>${O}_${cfg}.err
  python - <<PY
import json
try:
    d = json.loads(open('${O}_${cfg}.json').read().strip().splitlines()[-1])
    print('$cfg', round(d['value']), 'frames/s', round(d['ms_per_step']*1e3, 1), 'us/step', 'path frac', round(d['roofline_path']['frac'], 3), 'kernel frac', round(d['roofline']['frac'], 3), d['clocks'])
except Exception as e:
    print('$cfg failed', e)
PY
done
KERNELS='regex:npb'
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 40 --csv \
    --log-file ${O}_launches_nyuv2.csv python bench.py --config nyuv2 --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-graph \
    > ${O}_ncu_launches.log 2>&1
tail -16 ${O}_launches_nyuv2.csv | cut -d, -f5,9,15 | cut -c1-150
