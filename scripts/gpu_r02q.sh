#!/bin/bash
# round 2, call q (1 GPU): what does the FUSED instance of the row-major SpMMV kernel cost by itself (no neighbour)?  timing, then ncu
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for c in "sp 8" "dp 8" "dp 4"; do timeout 300 python scripts/mmv_fused_probe.py $c 50 2>&1 | tail -1 | tee -a gpurun_out/r02q_probe.txt; done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:stream_mmv -s 5 -c 2 -o /tmp/r02q_mmv python scripts/mmv_fused_probe.py sp 8 1 > gpurun_out/r02q_ncu.log 2>&1
ncu -i /tmp/r02q_mmv.ncu-rep --page raw --csv > gpurun_out/r02q_mmv_raw.csv 2>/dev/null
ncu -i /tmp/r02q_mmv.ncu-rep --page details > gpurun_out/r02q_mmv_details.txt 2>/dev/null
ls -la gpurun_out/ | grep r02q
