# final single-GPU evidence for the current tree: full GPU test-suite, smoke, bench line + reference arm, ncu launch list of the same
# command, and one full ncu capture of the fused AP kernel on the 2^25-row power-law matrix (config 4, one GPU)
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -x -q -m gpu ) > gpurun_out/r01v_pytest_gpu.log 2>&1; tail -4 gpurun_out/r01v_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py > gpurun_out/r01v_bench_n1.json 2> gpurun_out/r01v_bench_n1.err; cut -c1-400 gpurun_out/r01v_bench_n1.json
python bench.py --impl reference --steps 50 --warmup 5 > gpurun_out/r01v_bench_ref.json 2>> gpurun_out/r01v_bench_n1.err; cut -c1-200 gpurun_out/r01v_bench_ref.json
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/r01v_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01v_launches.csv $CMD > gpurun_out/r01v_ncu_launches.log 2>&1
python scripts/one_powerlaw.py 25 16384 > gpurun_out/r01v_pl25_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'k_scs32_stream_ap' -s 2 -c 1 -f -o /tmp/r01v_ap python scripts/one_powerlaw.py 25 16384 > gpurun_out/r01v_pl25_ncu.log 2>&1
python scripts/ncu_summary.py /tmp/r01v_ap.ncu-rep > gpurun_out/r01v_powerlaw_32M_ap_ncu_summary.txt 2>&1
ncu -i /tmp/r01v_ap.ncu-rep --page details --csv > gpurun_out/r01v_powerlaw_32M_ap_details.csv 2>/dev/null
tail -12 gpurun_out/r01v_powerlaw_32M_ap_ncu_summary.txt
