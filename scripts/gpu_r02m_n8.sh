#!/bin/bash
# round 2, call m (8 GPUs): the driver's scaling commands at N = 8, 4, 2 (validated y on every rank, config 5 in other_configs),
# config 4 (adaptive precision, seg-nnz) over 8 GPUs at sigma = 512
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/r02m_gpus.txt
run() { n=$1; out=$2; shift 2; ( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n "$@" ) > gpurun_out/$out.json 2> gpurun_out/$out.err; echo "$out rc=$?"; grep -E "^real" gpurun_out/$out.err; }
run 8 r02m_bench_n8 --steps 20 --warmup 5
run 4 r02m_bench_n4 --steps 20 --warmup 5
run 2 r02m_bench_n2 --steps 20 --warmup 5
run 8 r02m_bench_n8_cfg4 --steps 20 --warmup 5 --workload powerlaw_25 --ap "ap[dp_sp_hp]" --C 32 --sigma 512 --no-other-configs
run 8 r02m_bench_n8_bvs4 --steps 50 --warmup 5 --bvs 4 --layout rowwise --no-other-configs
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r02m_bench_*.json')):
    for line in open(f):
        if line.startswith('{'):
            d = json.loads(line)
            print(f.split('/')[-1][:-5], 'N=%d value %.1f step %.4f steady %s kernel %.4f valid %s err %.2e exch_err %s e2e %s' % (
                d['n_gpus'], d['value'], d['ms_per_step'], d['steady_state'] and round(d['steady_state']['ms_per_step'], 4), d['roofline']['kernel_ms'],
                d['validated'], d['max_rel_err'], d['exchange_errors'], d['e2e'] and round(d['e2e']['value'], 1)))
            for o in d.get('other_configs', []):
                print('   ', o['config'][:110], '| %.1f GF %.3f ms frac %.3f valid %s' % (o['value'], o['ms_per_step'], o['roofline']['frac'], o['validated']))
PY
