#!/bin/bash
# round 2, call f: restored SELL-32 loop + fused exchange for SpMMV / wide chunks: parity (one GPU, two ranks on it) + timings
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -q ) > gpurun_out/r02f_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02f_pytest_gpu.log
grep -E "passed|failed|FAILED|rc=" gpurun_out/r02f_pytest_gpu.log | tail -8
( timeout 900 python scripts/bench_configs.py formats spmmv ) > gpurun_out/r02f_configs.log 2>&1; cp gpurun_out/configs.json gpurun_out/r02f_configs_formats_spmmv.json 2>/dev/null
tail -24 gpurun_out/r02f_configs.log
