#!/bin/bash
# round 2, call I (2 GPUs): early acknowledgement (boundary chunks mid-stream) against the previous commit's library; sp / hp: fused vs multi-kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for lib in libuspmv_b200_prev.so libuspmv_b200.so; do
  USPMV_B200_LIB=$PWD/ultimate-spmv_b200/lib/$lib timeout 300 python scripts/spmv_fused_probe.py 2>&1 | tail -1 | tee -a gpurun_out/r02I_spmv_fused_probe.txt
done
run() { out=$1; shift; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 5 --no-other-configs --no-e2e --no-cpu-baseline --steady-steps 2000 "$@" > gpurun_out/$out.json 2> gpurun_out/$out.err; echo "$out rc=$?"; }
USPMV_B200_LIB=$PWD/ultimate-spmv_b200/lib/libuspmv_b200_prev.so run r02I_n2_dp_prevlib_a
run r02I_n2_dp_pos50_a --set fused_boundary_pos=50
run r02I_n2_dp_pos100_a --set fused_boundary_pos=100
USPMV_B200_LIB=$PWD/ultimate-spmv_b200/lib/libuspmv_b200_prev.so run r02I_n2_dp_prevlib_b
run r02I_n2_dp_pos50_b --set fused_boundary_pos=50
run r02I_n2_dp_pos25 --set fused_boundary_pos=25
run r02I_n2_sp_fused --vt sp
run r02I_n2_sp_mode1 --vt sp --p2p-mode 1
run r02I_n2_hp_fused --vt hp
run r02I_n2_hp_mode1 --vt hp --p2p-mode 1
run r02I_n2_dp_mode1 --p2p-mode 1
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r02I_n2_*.json')):
    for line in open(f):
        if line.startswith('{'):
            d = json.loads(line)
            print(f.split('/')[-1][:-5], 'value %.1f step %.4f steady %.4f kernel %.4f valid %s' % (d['value'], d['ms_per_step'], d['steady_state']['ms_per_step'], d['roofline']['kernel_ms'], d['validated']))
PY
