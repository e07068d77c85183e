#!/bin/bash
# round 2, call D (1 GPU): fused instance probe after the 16-byte-row variant choice
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for c in "sp 4 50 0" "dp 2 50 0" "dp 4 50 0" "sp 8 50 0" "sp 2 50 0" "hp 8 50 0" "hp 16 50 0"; do timeout 300 python scripts/mmv_fused_probe.py $c 2>&1 | tail -1 | tee -a gpurun_out/r02D_probe.txt; done
