#!/bin/bash
# in-situ timeline of the replayed step (library built with -DNPB_TIMELINE) + ncu launch list
set -u
mkdir -p gpurun_out
NPB_LIB_PATH=$PWD/build/timeline/libnicr_panoptic_b200.so python scripts/probes/timeline.py --config nyuv2 2>&1 | tail -14
KERNELS='regex:group_pixels|pair_count|nms_candidates|select_centers|finalize_instances|match_frames|accumulate_frames|write_panoptic'
ncu --metrics gpu__time_duration.sum --clock-control none -k "$KERNELS" -s 40 -c 24 --csv \
    --log-file gpurun_out/r02_launches_nyuv2.csv python bench.py --config nyuv2 --steps 30 --warmup 5 --no-e2e --no-cpu-baseline \
    > gpurun_out/r02_ncu_launches.log 2>&1
tail -12 gpurun_out/r02_launches_nyuv2.csv | cut -d, -f5,9,15 | cut -c1-160
