#!/bin/bash
# round 2, call h (1 GPU): full GPU test-suite + the driver's bench command on the current tree
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -q ) > gpurun_out/r02h_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02h_pytest_gpu.log
grep -E "passed|failed|FAILED|rc=" gpurun_out/r02h_pytest_gpu.log | tail -8
( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 --no-banded ) > gpurun_out/r02h_bench_n1.json 2> gpurun_out/r02h_bench_n1.err
echo "bench rc=$?"; tail -c 400 gpurun_out/r02h_bench_n1.err
python - <<'PY'
import json
for line in open('gpurun_out/r02h_bench_n1.json'):
    if line.startswith('{'):
        d=json.loads(line)
        print('headline %.1f GF %.4f ms (steady %.4f) kernel %.4f frac %.3f valid %s gpu_base %s e2e %.1f' % (d['value'], d['ms_per_step'], d['steady_state']['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['validated'], d['gpu_baseline'] and round(d['gpu_baseline'].get('value'),1), d['e2e']['value']))
        for o in d['other_configs']:
            print(o['config'][:100], '| %.1f GF step %.4f kernel %.4f ms frac %.3f valid %s' % (o['value'],o['ms_per_step'],o['roofline']['kernel_ms'],o['roofline']['frac'],o['validated']))
PY
