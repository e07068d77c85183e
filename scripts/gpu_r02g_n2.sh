#!/bin/bash
# round 2, call g (2 GPUs): the two-rank tests over real NVLink (log kept under profiles/), bench.py at N = 2 with validation,
# the config-5 strong-scaling slab in other_configs, SpMMV with the fused exchange
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/r02g_gpus.txt
( time timeout 1200 python -m pytest tests/test_gpu_dist_runtime.py tests/test_gpu_cli.py tests/test_gpu_ap_dist.py -m gpu -q -rs ) > gpurun_out/r02g_pytest_2gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02g_pytest_2gpu.log
grep -E "passed|failed|FAILED|SKIPPED|rc=" gpurun_out/r02g_pytest_2gpu.log | tail -12
run() { out=$1; shift; ( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 "$@" ) > gpurun_out/$out.json 2> gpurun_out/$out.err; echo "$out rc=$?"; tail -c 400 gpurun_out/$out.err; }
run r02g_bench_n2 --steps 20 --warmup 5
run r02g_bench_n2_bvs4 --steps 50 --warmup 5 --bvs 4 --layout rowwise --no-other-configs
run r02g_bench_n2_c64 --steps 50 --warmup 5 --C 64 --no-other-configs
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r02g_bench_n2*.json')):
    for line in open(f):
        if line.startswith('{'):
            d = json.loads(line)
            print(f.split('/')[-1], 'value %.1f ms %.4f kernel_ms %.4f valid %s err %.2e exch_err %s steady %s e2e %s' % (
                d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['validated'], d['max_rel_err'], d['exchange_errors'],
                d['steady_state'] and round(d['steady_state']['ms_per_step'], 4), d['e2e'] and round(d['e2e']['value'], 1)))
            for o in d.get('other_configs', []):
                print('   ', o['config'][:110], '| %.1f GF %.3f ms frac %.3f valid %s' % (o['value'], o['ms_per_step'], o['roofline']['frac'], o['validated']))
PY
