#!/bin/bash
# round 2, call v (2 GPUs): fused SpMMV step with the coalesced push published after the first chunk; distributed tests on real peers
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 scripts/dist_probe_mmv.py dp:4 dp:8 dp:8:17 dp:8:6 sp:8 sp:4 dp:2 2>&1 | grep "^{" | tee gpurun_out/r02v_dist_probe_mmv.txt
timeout 1200 python -m pytest tests/test_gpu_dist_runtime.py tests/test_gpu_cli.py -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r02v_pytest.log
