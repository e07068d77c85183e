#!/bin/bash
# round 2, call y (1 GPU): the fused SpMMV instance by itself (no neighbour), per kernel variant
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for c in "dp 8 50 0" "dp 8 50 17" "dp 8 50 6" "dp 8 50 2" "dp 4 50 0" "sp 8 50 0" "sp 4 50 0" "dp 2 50 0" "sp 16 50 0" "dp 16 50 0"; do timeout 300 python scripts/mmv_fused_probe.py $c 2>&1 | tail -1 | tee -a gpurun_out/r02y_probe.txt; done
