#!/bin/bash
# round 2, call t (2 GPUs): distributed row-major SpMMV — fused one-kernel step (mmv_fused_rowwise=1) against the multi-kernel overlap (0)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { out=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 5 --no-other-configs --no-e2e --steady-steps 300 "$@" > gpurun_out/$out.json 2> gpurun_out/$out.err; echo "$out rc=$?"; }
for f in 1 0; do
  run r02t_n2_bvs4_dp_fused$f --bvs 4 --layout rowwise --set mmv_fused_rowwise=$f
  run r02t_n2_bvs8_dp_fused$f --bvs 8 --layout rowwise --set mmv_fused_rowwise=$f
  run r02t_n2_bvs8_sp_fused$f --bvs 8 --layout rowwise --vt sp --set mmv_fused_rowwise=$f
  run r02t_n2_bvs4_sp_fused$f --bvs 4 --layout rowwise --vt sp --set mmv_fused_rowwise=$f
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r02t_n2_*.json')):
    for line in open(f):
        if line.startswith('{'):
            d = json.loads(line)
            print(f.split('/')[-1][:-5], 'value %.1f step %.4f steady %.4f kernel %.4f valid %s' % (d['value'], d['ms_per_step'], d['steady_state']['ms_per_step'], d['roofline']['kernel_ms'], d['validated']))
PY
