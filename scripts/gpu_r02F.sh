#!/bin/bash
# round 2, call G (1 GPU): the fused SpMV instance against the plain kernel, previous commit's library and this tree's
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for lib in libuspmv_b200_prev.so libuspmv_b200.so; do
  USPMV_B200_LIB=$PWD/ultimate-spmv_b200/lib/$lib timeout 300 python scripts/spmv_fused_probe.py 2>&1 | tail -1 | tee -a gpurun_out/r02G_spmv_fused_probe.txt
done
