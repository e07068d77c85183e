#!/bin/bash
# round 2, call c: lean streaming loop (uniform producer, register headers, templated slot counts): parity + timings
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -q -x ) > gpurun_out/r02c_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02c_pytest_gpu.log
tail -4 gpurun_out/r02c_pytest_gpu.log
( timeout 900 python scripts/bench_configs.py formats spmmv ) > gpurun_out/r02c_configs.log 2>&1; cp gpurun_out/configs.json gpurun_out/r02c_configs_formats_spmmv.json 2>/dev/null
cat gpurun_out/r02c_configs.log | tail -40
( time timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-banded ) > gpurun_out/r02c_bench_n1.json 2> gpurun_out/r02c_bench_n1.err
echo "bench rc=$?"; tail -c 600 gpurun_out/r02c_bench_n1.err
python - <<'PY'
import json
for line in open('gpurun_out/r02c_bench_n1.json'):
    if line.startswith('{'):
        d=json.loads(line)
        print('headline %.1f GF %.4f ms frac %.3f valid %s gpu_base %s e2e %.1f' % (d['value'], d['ms_per_step'], d['roofline']['frac'], d['validated'], d['gpu_baseline'] and d['gpu_baseline'].get('value'), d['e2e']['value']))
        for o in d['other_configs']:
            print(o['config'][:100], '| %.1f GF %.3f ms frac %.3f valid %s' % (o['value'],o['ms_per_step'],o['roofline']['frac'],o['validated']))
PY
( time AP_VARIANTS=1,2,3,4 timeout 600 python scripts/config4_probe.py gpurun_out/r02c_config4_probe.json 25 512 "" ) > gpurun_out/r02c_config4_probe.log 2>&1
echo "probe rc=$?"; tail -12 gpurun_out/r02c_config4_probe.log
