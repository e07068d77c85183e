# usage: bash scripts/gpu_r01n.sh N LOG2ROWS  (run under gpurun --gpus N): config 4 through bench.py
N=${1:-1}; L=${2:-22}
mkdir -p gpurun_out
if [ "$N" = "1" ]; then LAUNCH="python"; else LAUNCH="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29621"; fi
$LAUNCH bench.py --gpus $N --workload powerlaw_$L --ap "ap[dp_sp_hp]" --sigma 16384 --steps 50 --warmup 5 2> gpurun_out/r01n_n${N}_pl$L.err | tail -1 > gpurun_out/r01n_n${N}_pl$L.json
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r01n_n${N}_pl$L.json"))
    print("N=", d["n_gpus"], "ms/step", round(d["ms_per_step"], 4), "GFLOP/s", round(d["value"], 1), "kernel_ms", round(d["roofline"]["kernel_ms"], 4), d["config"]["partition"][:120], "halo", d["config"]["halo_elements_per_gpu"], "nnz/gpu", d["config"]["nnz_per_gpu"])
except Exception as e:
    print("FAILED", e); print(open("gpurun_out/r01n_n${N}_pl$L.err").read()[-2500:])
PY
