#!/bin/bash
# round 2, call s (1 GPU): new SpMMV defaults (register budget) — fused instance probe, SpMMV / distributed tests, config 3 lines
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for c in "sp 8" "dp 8" "dp 4" "sp 4"; do timeout 300 python scripts/mmv_fused_probe.py $c 50 2>&1 | tail -1 | tee -a gpurun_out/r02s_probe.txt; done
timeout 1500 python -m pytest tests -m gpu -x -q -k "spmmv or mmv or block or dist_runtime" 2>&1 | tail -4 | tee gpurun_out/r02s_pytest.log
timeout 900 python bench.py --steps 50 --warmup 5 --no-config4 --no-e2e --no-cpu-baseline --no-gpu-baseline > gpurun_out/r02s_bench.json 2> gpurun_out/r02s_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
for line in open('gpurun_out/r02s_bench.json'):
    if line.startswith('{'):
        d = json.loads(line)
        for o in d.get('other_configs', []):
            print(o.get('config'), '| step', o.get('ms_per_step'), 'frac', (o.get('roofline') or {}).get('frac'), 'valid', o.get('validated'))
PY
