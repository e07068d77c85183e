# final single-GPU evidence for round 1: bench line, ncu launch list of the same command, full captures of the three streamed kernels
# (the .ncu-rep files are ~40 MB each with sources: summarised on the box, only text / CSV comes back)
mkdir -p gpurun_out
summarise() { rep=$1; tag=$2
  python scripts/ncu_summary.py $rep > gpurun_out/${tag}_ncu_summary.txt 2>&1
  ncu -i $rep --page details --csv > gpurun_out/${tag}_details.csv 2>/dev/null
  ncu -i $rep --page source --csv > gpurun_out/${tag}_source.csv 2>/dev/null
  rm -f $rep; }
python bench.py > gpurun_out/r01p_bench_n1.json 2> gpurun_out/r01p_bench_n1.err
python bench.py --impl reference --steps 50 --warmup 5 > gpurun_out/r01p_bench_ref.json 2>> gpurun_out/r01p_bench_n1.err
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/r01p_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01p_launches.csv $CMD > gpurun_out/r01p_ncu_launches.log 2>&1
$CMD > gpurun_out/r01p_plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'k_scs32_stream' -s 3 -c 1 -f -o /tmp/r01p_spmv $CMD > gpurun_out/r01p_ncu_full.log 2>&1
summarise /tmp/r01p_spmv.ncu-rep r01p_spmv
python scripts/one_case.py spmmv dp 8 rowwise > gpurun_out/r01p_mmv_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'k_scs32_stream_mmv' -s 2 -c 1 -f -o /tmp/r01p_mmv python scripts/one_case.py spmmv dp 8 rowwise > gpurun_out/r01p_mmv_ncu.log 2>&1
summarise /tmp/r01p_mmv.ncu-rep r01p_mmv_dp8row
python scripts/one_powerlaw.py 22 16384 > gpurun_out/r01p_pl_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'k_scs32_stream_ap' -s 2 -c 1 -f -o /tmp/r01p_ap python scripts/one_powerlaw.py 22 16384 > gpurun_out/r01p_pl_ncu.log 2>&1
summarise /tmp/r01p_ap.ncu-rep r01p_powerlaw_ap
ls -la gpurun_out/r01p_* | awk '{print $5, $9}'
