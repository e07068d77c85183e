#!/bin/bash
# round 2, call r (1 GPU): SpMMV variants with a register budget (MINB); the fused instance with its boundary pass out of line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python scripts/tune_mmv3.py 2>&1 | tee gpurun_out/r02r_tune_mmv3.txt
for c in "sp 8" "dp 8" "dp 4"; do timeout 300 python scripts/mmv_fused_probe.py $c 50 2>&1 | tail -1 | tee -a gpurun_out/r02r_probe.txt; done
