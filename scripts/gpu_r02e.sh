#!/bin/bash
# ncu --set full of the SELL-32 SpMV kernel, sp and hp, round-1 loop (libold) against the lean loop (libnew); summaries only
cd "$(dirname "$0")/.."
mkdir -p gpurun_out /tmp/ncu
L=$PWD/ultimate-spmv_b200/lib/ab
for lib in old new; do for vt in sp hp; do
  USPMV_B200_LIB=$L/lib$lib.so timeout 600 ncu --set full --clock-control none -k regex:k_scs32_stream --launch-skip 3 --launch-count 1 \
      -o /tmp/ncu/r02e_${lib}_${vt} -f python scripts/ab_one.py $vt > /tmp/ncu/log_${lib}_${vt}.txt 2>&1
  echo "$lib $vt rc=$?"
  python scripts/ncu_summary.py /tmp/ncu/r02e_${lib}_${vt}.ncu-rep > gpurun_out/r02e_${lib}_${vt}_ncu_summary.txt 2>&1
done; done
ls -la /tmp/ncu/
