# final check of the tree as committed: GPU test-suite, smoke, bench line, reference arm, launch list of the bench command
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -x -q -m gpu ) > gpurun_out/r01z_pytest_gpu.log 2>&1; tail -5 gpurun_out/r01z_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py > gpurun_out/r01z_bench_n1.json 2> gpurun_out/r01z_bench_n1.err; cut -c1-330 gpurun_out/r01z_bench_n1.json
python bench.py --impl reference --steps 30 --warmup 3 > gpurun_out/r01z_bench_ref.json 2>> gpurun_out/r01z_bench_n1.err; cut -c1-200 gpurun_out/r01z_bench_ref.json
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/r01z_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01z_launches.csv $CMD > gpurun_out/r01z_ncu_launches.log 2>&1
grep -c k_scs32_stream gpurun_out/r01z_launches.csv
