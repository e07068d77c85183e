# 8-GPU batch: default weak scaling, solve loop, SpMMV bvs 4, config 5 strong scaling, config 4 (2^25 rows, seg-nnz)
N=8
mkdir -p gpurun_out
run() { tag=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29631 bench.py --gpus $N "$@" 2> gpurun_out/r01o_n${N}_${tag}.err | tail -1 > gpurun_out/r01o_n${N}_${tag}.json; python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r01o_n${N}_${tag}.json"))
    print("${tag}", "N=", d["n_gpus"], "ms/step", round(d["ms_per_step"], 4), "GFLOP/s", round(d["value"], 1), "kernel_ms", round(d["roofline"]["kernel_ms"], 4), "e2e", d["e2e"] and round(d["e2e"]["value"], 1))
except Exception as e:
    print("${tag} FAILED", e); print(open("gpurun_out/r01o_n${N}_${tag}.err").read()[-1500:])
PY
}
run default --steps 2000 --warmup 50
run solve --solve --steps 2000 --warmup 50
run bvs4row --bvs 4 --layout rowwise --steps 500 --warmup 20 --no-e2e
run cfg5 --workload stencil27_512 --strong --steps 500 --warmup 20 --no-e2e
run cfg4 --workload powerlaw_25 --ap "ap[dp_sp_hp]" --sigma 16384 --steps 50 --warmup 5
