# usage: bash scripts/gpu_r01r_push.sh N LOG2ROWS [CASES] [TESTS]  (under gpurun --gpus N): [2-rank runtime tests,] push-kernel sweep
N=${1:-2}; L=${2:-23}; CASES=${3:-0:0,1:2,1:4,1:8,2:2,2:4,2:8}; TESTS=${4:-1}
mkdir -p gpurun_out
[ "$TESTS" = "1" ] && timeout 600 python -m pytest tests/test_gpu_dist_runtime.py -x -q -m gpu 2>&1 | tail -5
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    scripts/tune_push.py --log2 $L --cases $CASES --out gpurun_out/r01r_tune_push_n${N}_pl$L.json 2> gpurun_out/r01r_tune_push_n${N}.err | tail -12
tail -5 gpurun_out/r01r_tune_push_n${N}.err
