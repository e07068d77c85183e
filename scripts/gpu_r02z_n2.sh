#!/bin/bash
# round 2, call A (2 GPUs): fused SpMMV step = coalesced segment push published at once + inline main loop + out-of-line boundary
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 scripts/dist_probe_mmv.py dp:4 dp:8 sp:8 sp:4 dp:2 sp:16 2>&1 | grep "^{" | tee gpurun_out/r02A_dist_probe_mmv.txt
for c in "dp 8 50 0" "dp 4 50 0" "sp 8 50 0" "sp 4 50 0" "dp 2 50 0" "sp 16 50 0"; do timeout 300 python scripts/mmv_fused_probe.py $c 2>&1 | tail -1 | tee -a gpurun_out/r02A_probe.txt; done
