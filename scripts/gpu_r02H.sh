#!/bin/bash
# round 2, call H (1 GPU): ncu of the sp SpMV kernel and its fused instance (no neighbour) — why is the instance 25 % slower?
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_scs32_stream -s 18 -c 4 -o /tmp/r02H_sp python scripts/spmv_fused_probe.py 10 sp > gpurun_out/r02H_ncu.log 2>&1
ncu -i /tmp/r02H_sp.ncu-rep --page raw --csv > gpurun_out/r02H_sp_raw.csv 2>/dev/null
ncu -i /tmp/r02H_sp.ncu-rep --page details > gpurun_out/r02H_sp_details.txt 2>/dev/null
ncu -i /tmp/r02H_sp.ncu-rep --page source --csv > gpurun_out/r02H_sp_source.csv 2>/dev/null
ls -la gpurun_out | grep r02H; tail -3 gpurun_out/r02H_ncu.log
