#!/bin/bash
# one optimisation iteration on a B200: parity suite, A/B bench lines, in-situ timeline
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -2
CFGS="${CFGS:-nyuv2}" bash scripts/r02_ab.sh
NPB_LIB_PATH=$PWD/build/timeline/libnicr_panoptic_b200.so timeout 300 python scripts/probes/timeline.py --config nyuv2 2>&1 | tail -11
${EXTRA:-true}
