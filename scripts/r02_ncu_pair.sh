#!/bin/bash
# ncu --set full capture of the evaluation pixel pass: stand-alone (eval-only workload) and fused (nyuv2 step)
set -u
mkdir -p gpurun_out
python scripts/bench_eval.py --frames 5120 > gpurun_out/r02_eval_plain.log 2>&1; tail -1 gpurun_out/r02_eval_plain.log | cut -c1-300
ncu --set full --clock-control none --import-source on -k regex:pair_count -s 6 -c 1 -f \
    -o gpurun_out/r02_pair_eval python scripts/bench_eval.py --frames 2560 > gpurun_out/r02_ncu_pair_eval.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pair_count -s 6 -c 1 -f \
    -o gpurun_out/r02_pair_fused python bench.py --config nyuv2 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-api --no-extra --no-graph > gpurun_out/r02_ncu_pair_fused.log 2>&1
ls -la gpurun_out/*.ncu-rep
