# usage: bash scripts/gpu_bench_multi.sh N [extra args]  (run under gpurun --gpus N)
N=${1:-2}; shift
TAG=n${N}$(echo "$*" | tr -d ' -')
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 200 --warmup 20 $* > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
echo "rc=$?"; grep -v "^\*\*\*\|OMP_NUM\|^$" gpurun_out/bench_${TAG}.err | tail -8; python -c "
import json,sys
for l in open('gpurun_out/bench_${TAG}.json'):
    try: d=json.loads(l)
    except Exception: continue
    print({k:d[k] for k in ('value','n_gpus','ms_per_step','gbs','gpu_launches')}, d['roofline']['kernel_ms'], d['e2e'] and d['e2e']['value'], d['config']['partition'])
"
