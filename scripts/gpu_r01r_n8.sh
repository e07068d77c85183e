# under gpurun --gpus 8: push sweep on config 4 (2^25 rows), then config 4 through bench.py with the library defaults
mkdir -p gpurun_out
L="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 400 $L --master-port 29533 scripts/tune_push.py --log2 25 --cases 1:6,1:8,3:4,3:8 --out gpurun_out/r01r_tune_push_n8_pl25b.json 2> gpurun_out/r01r_tune_push_n8b.err | tail -8
timeout 400 $L --master-port 29541 bench.py --gpus 8 --workload powerlaw_25 --ap "ap[dp_sp_hp]" --sigma 16384 --steps 50 --warmup 5 2> gpurun_out/r01r_n8_cfg4.err | tail -1 > gpurun_out/r01r_n8_cfg4.json
python -c "
import json; d=json.load(open('gpurun_out/r01r_n8_cfg4.json')); print('bench cfg4 N=8 ms/step', d['ms_per_step'], 'GFLOP/s', d['value'], 'kernel_ms', d['roofline']['kernel_ms'])" || tail -20 gpurun_out/r01r_n8_cfg4.err
