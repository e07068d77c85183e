#!/bin/bash
# round 2, call b: full GPU test-suite (incl. full-size parity against oracle/_ref), the driver's bench command, config-4 probe
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
free -g > gpurun_out/r02b_host.txt; nproc >> gpurun_out/r02b_host.txt
( time timeout 1800 python -m pytest tests -m gpu -q ) > gpurun_out/r02b_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02b_pytest_gpu.log
tail -5 gpurun_out/r02b_pytest_gpu.log
( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/r02b_bench_n1.json 2> gpurun_out/r02b_bench_n1.err
echo "bench rc=$?"; tail -c 1500 gpurun_out/r02b_bench_n1.err
head -c 8000 gpurun_out/r02b_bench_n1.json
( time timeout 900 python scripts/config4_probe.py gpurun_out/r02b_config4_probe.json 25 512,16384 4,8,16 ) > gpurun_out/r02b_config4_probe.log 2>&1
echo "probe rc=$?"; tail -20 gpurun_out/r02b_config4_probe.log
