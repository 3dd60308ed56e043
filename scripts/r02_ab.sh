#!/bin/bash
# A/B of the forward + evaluation chain on one B200 (run under gpurun): GPU parity suite with the
# new library, then bench lines for the round-1 library (build/r01, same ABI) and the new one.
set -u
mkdir -p gpurun_out
OUT=gpurun_out/r02a
python -m pytest tests -q -m gpu -x 2>&1 | tail -15 > ${OUT}_tests.log
tail -3 ${OUT}_tests.log
for cfg in nyuv2 sunrgbd; do
  for lib in old new; do
    if [ $lib = old ]; then export NPB_LIB_PATH=$PWD/build/r01/libnicr_panoptic_b200.so; else unset NPB_LIB_PATH; fi
    python bench.py --config $cfg --steps 300 --warmup 3 --no-e2e --no-cpu-baseline > ${OUT}_${cfg}_${lib}.json 2>${OUT}_${cfg}_${lib}.err
    python - <<PY
import json
try:
    d = json.loads(open('${OUT}_${cfg}_${lib}.json').read().strip().splitlines()[-1])
    print('$cfg $lib', round(d['value']), 'frames/s', round(d['ms_per_step']*1e3, 1), 'us/step', 'path frac', round(d['roofline_path']['frac'], 3), 'kernel frac', round(d['roofline']['frac'], 3), d['clocks'])
except Exception as e:
    print('$cfg $lib failed', e)
PY
  done
done
unset NPB_LIB_PATH
KERNELS='regex:group_pixels|pair_count|nms_candidates|select_centers|finalize_instances|match_frames|accumulate_frames|write_panoptic'
ncu --metrics gpu__time_duration.sum --clock-control none -k "$KERNELS" -s 40 -c 60 --csv \
    --log-file ${OUT}_launches_nyuv2.csv python bench.py --config nyuv2 --steps 50 --warmup 5 --no-e2e --no-cpu-baseline \
    > ${OUT}_ncu_launches.log 2>&1
tail -12 ${OUT}_launches_nyuv2.csv | cut -d, -f5,9,15
