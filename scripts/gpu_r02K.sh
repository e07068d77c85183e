#!/bin/bash
# round 2, call K (1 GPU): full GPU suite + default bench after the SpMMV changes
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/r02K_pytest_gpu.log
( time timeout 900 python bench.py > gpurun_out/r02K_bench_n1.json ) 2> gpurun_out/r02K_bench_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
for line in open('gpurun_out/r02K_bench_n1.json'):
    if line.startswith('{'):
        d = json.loads(line)
        print('headline %.1f GF %.4f ms kernel %.4f frac %.3f valid %s gpu_base %s cusparse %s e2e %.1f' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['validated'], (d.get('gpu_baseline') or {}).get('value'), ((d.get('gpu_baseline') or {}).get('cusparse') or {}).get('value'), d['e2e']['value']))
        for o in d.get('other_configs', []):
            print(o.get('config'), '| step', o.get('ms_per_step'), 'frac', (o.get('roofline') or {}).get('frac'), 'valid', o.get('validated'))
PY
tail -n 4 gpurun_out/r02K_bench_n1.err
