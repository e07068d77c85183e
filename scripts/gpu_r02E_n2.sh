#!/bin/bash
# round 2, call E (2 GPUs): boundary chunks in the middle of the item stream + early acknowledgement — tests, then the headline at N = 2 per position
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_dist_runtime.py -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r02E_pytest_n2.log
run() { out=$1; shift; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 5 --no-other-configs --no-e2e --steady-steps 2000 "$@" > gpurun_out/$out.json 2> gpurun_out/$out.err; echo "$out rc=$?"; }
for pos in 100 50 25 75 100 50; do run r02E_n2_pos${pos}_$RANDOM --set fused_boundary_pos=$pos; done
run r02E_n2_sp_pos100 --vt sp --set fused_boundary_pos=100
run r02E_n2_sp_pos50 --vt sp --set fused_boundary_pos=50
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r02E_n2_*.json')):
    for line in open(f):
        if line.startswith('{'):
            d = json.loads(line)
            print(f.split('/')[-1][:-5], 'value %.1f step %.4f steady %.4f kernel %.4f valid %s' % (d['value'], d['ms_per_step'], d['steady_state']['ms_per_step'], d['roofline']['kernel_ms'], d['validated']))
PY
