# config sweep after the AP-segment / multi-GPU changes (one GPU; run under gpurun)
mkdir -p gpurun_out
python scripts/bench_configs.py cusparse 2>&1 | tail -6
python scripts/bench_configs.py spmmv 2>&1 | tail -9
PL_ROWS=4194304 PL_SIGMA=16384 python scripts/bench_configs.py ap 2>&1 | tail -5
python bench.py --bvs 4 --layout rowwise --steps 200 --warmup 10 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_n1_bvs4.json; cat gpurun_out/bench_n1_bvs4.json
python bench.py --solve --steps 500 --warmup 10 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_n1_solve.json; cat gpurun_out/bench_n1_solve.json
PL_ROWS=33554432 PL_SIGMA=16384 timeout 900 python scripts/bench_configs.py ap 2>&1 | tail -5
