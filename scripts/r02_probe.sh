#!/bin/bash
set -u
mkdir -p gpurun_out
python scripts/probes/graph_gap.py nyuv2 2>&1 | tail -8
grep -c "label" gpurun_out/step_graph.dot; grep -o 'label="[^"]*"' gpurun_out/step_graph.dot | cut -c1-200 | head -30
