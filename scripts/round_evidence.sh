#!/bin/bash
# Round evidence on one B200 (run under gpurun): GPU parity suite, smoke, the driver's two bench
# commands, the ncu launch list of the bench command and `--set full` captures of the dominant kernel
# (grouping) and of the evaluation pixel pass on the headline workload.  Numbers printed under ncu
# are never bench values.
set -u
R=${ROUND:-r02}
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -x 2>&1 | tail -3 > gpurun_out/${R}_tests.log
# the same suite against the -DNPB_DEBUG build (bounds / overflow asserts of the shared-memory tables)
if [ -f build/debug/libnicr_panoptic_b200.so ]; then
  NPB_LIB_PATH=$PWD/build/debug/libnicr_panoptic_b200.so python -m pytest tests -q -m gpu -x 2>&1 | tail -3 > gpurun_out/${R}_tests_debug_build.log
  NPB_LIB_PATH=$PWD/build/debug/libnicr_panoptic_b200.so python -c "from nicr_mt_scene_analysis_b200 import _lib; print(_lib.lib().npb_build_info().decode())" >> gpurun_out/${R}_tests_debug_build.log 2>&1
fi
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${R}_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/${R}_smoke.log
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/${R}_bench_reference.json 2>gpurun_out/${R}_bench_reference.err
python bench.py --steps 20 --warmup 3 > gpurun_out/${R}_bench_steps20.json 2>gpurun_out/${R}_bench.err
python bench.py > gpurun_out/${R}_bench_default.json 2>>gpurun_out/${R}_bench.err
KERNELS='regex:group_pixels|pair_count|nms_candidates|select_centers|finalize_instances|match_frames|accumulate_frames|write_panoptic'
ncu --metrics gpu__time_duration.sum --clock-control none -k "$KERNELS" -s 40 -c 32 --csv \
    --log-file gpurun_out/${R}_launches_nyuv2.csv python bench.py --steps 50 --warmup 5 --no-e2e --no-cpu-baseline --no-api --no-extra \
    > gpurun_out/${R}_ncu_launches.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k "$KERNELS" -s 40 -c 16 --csv \
    --log-file gpurun_out/${R}_launches_sunrgbd.csv python bench.py --config sunrgbd --steps 30 --warmup 5 --no-e2e --no-cpu-baseline --no-api --no-extra \
    > gpurun_out/${R}_ncu_launches_sunrgbd.log 2>&1
for cfg in nyuv2 sunrgbd; do
  ncu --set full --clock-control none --import-source on -k regex:group_pixels -s 6 -c 1 -f \
      -o gpurun_out/${R}_group_pixels_$cfg python bench.py --config $cfg --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-api --no-extra --no-graph > gpurun_out/${R}_ncu_group_$cfg.log 2>&1
  ncu -i gpurun_out/${R}_group_pixels_$cfg.ncu-rep --page raw --csv > gpurun_out/${R}_group_pixels_${cfg}_ncu_raw.csv 2>/dev/null
done
ncu --set full --clock-control none --import-source on -k regex:pair_count -s 6 -c 1 -f \
    -o gpurun_out/${R}_pair_count_nyuv2 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-api --no-extra --no-graph > gpurun_out/${R}_ncu_pair.log 2>&1
ncu -i gpurun_out/${R}_pair_count_nyuv2.ncu-rep --page raw --csv > gpurun_out/${R}_pair_count_nyuv2_ncu_raw.csv 2>/dev/null
# evaluation-only workload (BASELINE configs[4]): bench line + capture of the stand-alone pixel pass
python scripts/bench_eval.py > gpurun_out/${R}_bench_eval.json 2>gpurun_out/${R}_bench_eval.err
ncu --set full --clock-control none --import-source on -k regex:pair_count -s 6 -c 1 -f \
    -o gpurun_out/${R}_pair_count_eval python scripts/bench_eval.py --frames 2560 > gpurun_out/${R}_ncu_pair_eval.log 2>&1
ncu -i gpurun_out/${R}_pair_count_eval.ncu-rep --page raw --csv > gpurun_out/${R}_pair_count_eval_ncu_raw.csv 2>/dev/null
cut -c1-260 gpurun_out/${R}_bench_eval.json
tail -2 gpurun_out/${R}_tests.log; tail -1 gpurun_out/${R}_smoke.log; cut -c1-330 gpurun_out/${R}_bench_default.json; cut -c1-200 gpurun_out/${R}_bench_reference.json
tail -9 gpurun_out/${R}_launches_nyuv2.csv | cut -d, -f5,9,15 | cut -c1-120
