# usage: bash scripts/gpu_ncu_spmv.sh <tag> [extra bench args]   (one GPU; run under gpurun)
TAG=${1:-r01}; shift
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e $*"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_launches.log 2>&1
$CMD > gpurun_out/${TAG}_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_scs32_stream|k_scs_spmv' -s 3 -c 2 -f -o gpurun_out/${TAG}_spmv $CMD > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -2 gpurun_out/${TAG}_plain.log; tail -3 gpurun_out/${TAG}_ncu_full.log; ls -la gpurun_out | tail -8
