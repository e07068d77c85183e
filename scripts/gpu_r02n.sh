#!/bin/bash
# round 2, call n (1 GPU): per-context options, shared-memory sigma-window sort: full suite, build timings, the driver's bench command
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -q ) > gpurun_out/r02n_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02n_pytest_gpu.log
grep -E "passed|failed|FAILED|rc=" gpurun_out/r02n_pytest_gpu.log | tail -8
python - > gpurun_out/r02n_build_times.txt 2>&1 <<'PY'
import importlib, sys, os, time, torch
sys.path.insert(0, os.getcwd())
pkg = importlib.import_module("ultimate-spmv_b200"); eng = pkg.engine
def build(mtx, C, sigma, vt="dp"):
    torch.cuda.synchronize(); t0 = time.time()
    s = eng.convert_to_scs(mtx, C, sigma, vt)
    torch.cuda.synchronize(); dt = time.time() - t0
    ne = s.n_elements
    del s
    return dt, ne
mtx = eng.MtxData.stencil(7, 256, 256, 256)
build(mtx, 32, 1)
for sigma in (1, 64, 512, 16384, 1 << 20):
    dt, ne = build(mtx, 32, sigma)
    print(f"convert_to_scs 7-pt 256^3 (16.8 M rows, 117 M nnz) C=32 sigma={sigma}: {dt * 1e3:.1f} ms (n_elements {ne})", flush=True)
del mtx
n = 1 << 25
mtx = eng.MtxData.powerlaw(n)
coos = eng.partition_precisions(mtx, "ap[dp_sp_hp]", 1.0, 1e-2)
del mtx
for sigma in (512, 16384):
    dt, ne = build(coos[0], 32, sigma)
    print(f"convert_to_scs power-law 2^25 rows, dp part ({coos[0].nnz} nnz) C=32 sigma={sigma}: {dt * 1e3:.1f} ms (n_elements {ne})", flush=True)
PY
cat gpurun_out/r02n_build_times.txt
( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 --no-banded ) > gpurun_out/r02n_bench_n1.json 2> gpurun_out/r02n_bench_n1.err
echo "bench rc=$?"; tail -c 300 gpurun_out/r02n_bench_n1.err
python - <<'PY'
import json
for line in open('gpurun_out/r02n_bench_n1.json'):
    if line.startswith('{'):
        d=json.loads(line)
        print('headline %.1f GF %.4f ms (steady %.4f) kernel %.4f frac %.3f valid %s gpu_base %s e2e %.1f' % (d['value'], d['ms_per_step'], d['steady_state']['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['validated'], d['gpu_baseline'] and round(d['gpu_baseline'].get('value'),1), d['e2e']['value']))
        for o in d['other_configs']:
            print(o['config'][:100], '| %.1f GF step %.4f kernel %.4f ms frac %.3f valid %s build %.1f' % (o['value'],o['ms_per_step'],o['roofline']['kernel_ms'],o['roofline']['frac'],o['validated'],o['build_s']))
PY
