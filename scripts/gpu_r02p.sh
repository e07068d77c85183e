#!/bin/bash
# round 2, call p (1 GPU): cuSPARSE baseline test + headline line with the cuSPARSE sliced-ELL / CSR numbers beside the reference's own kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_cusparse_baseline.py -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/r02p_pytest.log
timeout 900 python bench.py --steps 50 --warmup 5 --no-other-configs > gpurun_out/r02p_bench_scs.json 2> gpurun_out/r02p_bench_scs.err; echo "scs rc=$?"
timeout 900 python bench.py --steps 50 --warmup 5 --no-other-configs --C 1 --sigma 1 > gpurun_out/r02p_bench_crs.json 2> gpurun_out/r02p_bench_crs.err; echo "crs rc=$?"
timeout 900 python bench.py --steps 50 --warmup 5 --no-other-configs --vt sp > gpurun_out/r02p_bench_scs_sp.json 2> gpurun_out/r02p_bench_scs_sp.err; echo "sp rc=$?"
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r02p_bench_*.json')):
    for line in open(f):
        if line.startswith('{'):
            d = json.loads(line); g = d.get('gpu_baseline') or {}
            print(f.split('/')[-1], 'ours %.1f | ref kernel %s | cusparse %s' % (d['value'], g.get('value'), json.dumps(g.get('cusparse'))))
PY
tail -n 3 gpurun_out/r02p_bench_*.err
