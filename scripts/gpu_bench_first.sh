mkdir -p gpurun_out
python bench.py --steps 200 --warmup 20 > gpurun_out/bench_s1.json 2> gpurun_out/bench_s1.err; tail -3 gpurun_out/bench_s1.err; cat gpurun_out/bench_s1.json
python bench.py --steps 200 --warmup 20 --sigma 512 --no-cpu-baseline > gpurun_out/bench_s512.json 2> gpurun_out/bench_s512.err; tail -3 gpurun_out/bench_s512.err; cat gpurun_out/bench_s512.json
python bench.py --impl reference --steps 50 --warmup 5 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -3 gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json
nproc; lscpu | grep "Model name"
