#!/bin/bash
# round 2, call C (2 GPUs): distributed tests on real peers, the default bench at N = 2, SpMMV lines, the step decomposition
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_dist_runtime.py tests/test_gpu_cli.py -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r02C_pytest_n2.log
run() { out=$1; shift; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 "$@" > gpurun_out/$out.json 2> gpurun_out/$out.err; echo "$out rc=$?"; }
run r02C_bench_n2
run r02C_n2_bvs4_dp --steps 50 --warmup 5 --no-other-configs --no-e2e --steady-steps 300 --bvs 4 --layout rowwise
run r02C_n2_bvs8_dp --steps 50 --warmup 5 --no-other-configs --no-e2e --steady-steps 300 --bvs 8 --layout rowwise
run r02C_n2_bvs8_sp --steps 50 --warmup 5 --no-other-configs --no-e2e --steady-steps 300 --bvs 8 --layout rowwise --vt sp
run r02C_n2_bvs4_sp --steps 50 --warmup 5 --no-other-configs --no-e2e --steady-steps 300 --bvs 4 --layout rowwise --vt sp
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 scripts/dist_probe_mmv.py dp:8 sp:8 dp:4 2>&1 | grep "^{" | tee gpurun_out/r02C_dist_probe_mmv.txt
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r02C_*.json')):
    for line in open(f):
        if line.startswith('{'):
            d = json.loads(line)
            print(f.split('/')[-1][:-5], 'value %.1f step %.4f steady %s kernel %.4f valid %s' % (d['value'], d['ms_per_step'], (d.get('steady_state') or {}).get('ms_per_step'), d['roofline']['kernel_ms'], d['validated']))
            for o in d.get('other_configs', []):
                print('   ', o.get('config'), '| value', o.get('value'), 'step', o.get('ms_per_step'), 'valid', o.get('validated'))
PY
