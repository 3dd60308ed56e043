#!/bin/bash
# build the CUDA library of another git revision (same ABI) for A/B runs: build/<name>/libnicr_panoptic_b200.so
# usage: scripts/build_rev.sh <rev> <name>
set -eu
REV=$1; NAME=$2
TMP=$(mktemp -d)
git archive "$REV" nicr-multitask-scene-analysis_b200/csrc include | tar -x -C "$TMP"
mkdir -p build/$NAME
/usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -shared -Xcompiler -fPIC \
  -I "$TMP/include" -I "$TMP/nicr-multitask-scene-analysis_b200/csrc" "$TMP"/nicr-multitask-scene-analysis_b200/csrc/*.cu \
  -o build/$NAME/libnicr_panoptic_b200.so
rm -rf "$TMP"
echo build/$NAME/libnicr_panoptic_b200.so
