# usage: bash scripts/gpu_r01l_multi.sh N   (run under gpurun --gpus N): default step, solve loop, SpMMV bvs 4 / 8
N=${1:-2}
mkdir -p gpurun_out
run() { tag=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N "$@" 2> gpurun_out/r01l_n${N}_${tag}.err | tail -1 > gpurun_out/r01l_n${N}_${tag}.json; python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r01l_n${N}_${tag}.json"))
    print("${tag}", "N=", d["n_gpus"], "ms/step", round(d["ms_per_step"], 4), "GFLOP/s", round(d["value"], 1), "kernel_ms", round(d["roofline"]["kernel_ms"], 4), "e2e", d["e2e"] and round(d["e2e"]["value"], 1))
except Exception as e:
    print("${tag} FAILED", e); print(open("gpurun_out/r01l_n${N}_${tag}.err").read()[-1500:])
PY
}
run default --steps 2000 --warmup 50
run solve --solve --steps 2000 --warmup 50
run bvs4row --bvs 4 --layout rowwise --steps 500 --warmup 20
run bvs8row --bvs 8 --layout rowwise --steps 300 --warmup 20
run bvs4col --bvs 4 --layout colwise --steps 500 --warmup 20
