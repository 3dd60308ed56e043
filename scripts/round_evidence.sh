#!/bin/bash
# Round-end evidence on one B200 (run under gpurun): GPU parity suite, default bench line, the
# ncu launch list of the same bench command and one `--set full` capture of the stand-alone
# confusion-matrix kernel.  Numbers printed under ncu are never bench values.
set -u
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -x 2>&1 | tail -3 > gpurun_out/r01_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r01_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r01_smoke.log
python bench.py --steps 200 --warmup 3 > gpurun_out/r01_bench.log 2>gpurun_out/r01_bench.err
KERNELS='regex:group_pixels|pair_count|nms_candidates|select_centers|finalize_instances|match_frames|accumulate_frames|write_panoptic'
ncu --metrics gpu__time_duration.sum --clock-control none -k "$KERNELS" -s 35 -c 140 --csv \
    --log-file gpurun_out/r01_launches.csv python bench.py --steps 50 --warmup 5 --no-e2e --no-cpu-baseline \
    > gpurun_out/r01_ncu_launches.log 2>&1
python scripts/bench_miou.py > gpurun_out/r01_miou.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:confmat_stream -s 4 -c 1 \
    -o gpurun_out/r01_confmat python scripts/bench_miou.py --warmup 3 --reps 3 > gpurun_out/r01_ncu_confmat.log 2>&1
tail -2 gpurun_out/r01_tests.log; tail -1 gpurun_out/r01_smoke.log; cut -c1-400 gpurun_out/r01_bench.log
