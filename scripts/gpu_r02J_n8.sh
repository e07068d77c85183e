#!/bin/bash
# round 2, call J (8 GPUs): the driver's weak-scaling command with the acknowledgement at kernel end (previous library, pos 100) and early (pos 50)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L | wc -l > gpurun_out/r02J_gpus.txt
run() { out=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 --no-other-configs --no-e2e --no-cpu-baseline --steady-steps 2000 "$@" > gpurun_out/$out.json 2> gpurun_out/$out.err; echo "$out rc=$?"; }
USPMV_B200_LIB=$PWD/ultimate-spmv_b200/lib/libuspmv_b200_prev.so run r02J_n8_prevlib
run r02J_n8_pos50 --set fused_boundary_pos=50
run r02J_n8_pos100 --set fused_boundary_pos=100
run r02J_n8_pos50_b --set fused_boundary_pos=50
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r02J_n8_*.json')):
    for line in open(f):
        if line.startswith('{'):
            d = json.loads(line)
            print(f.split('/')[-1][:-5], 'value %.1f step %.4f steady %.4f kernel %.4f valid %s' % (d['value'], d['ms_per_step'], d['steady_state']['ms_per_step'], d['roofline']['kernel_ms'], d['validated']))
PY
