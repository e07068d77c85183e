#!/bin/bash
# round 2, call a: full GPU test-suite (incl. the two-ranks-on-one-GPU protocol tests), the driver's bench command, launch list
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader > gpurun_out/r02a_gpus.txt
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r02a_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02a_pytest_gpu.log
tail -5 gpurun_out/r02a_pytest_gpu.log
( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/r02a_bench_n1.json 2> gpurun_out/r02a_bench_n1.err
echo "bench rc=$?"; tail -c 1500 gpurun_out/r02a_bench_n1.err
head -c 6000 gpurun_out/r02a_bench_n1.json
