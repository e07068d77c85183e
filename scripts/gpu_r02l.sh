#!/bin/bash
# round 2, call l (1 GPU): fp16 pair kernel — parity + timing; full suite
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -q ) > gpurun_out/r02l_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02l_pytest_gpu.log
grep -E "passed|failed|FAILED|rc=" gpurun_out/r02l_pytest_gpu.log | tail -8
( timeout 900 python scripts/bench_configs.py formats ) > gpurun_out/r02l_configs.log 2>&1
grep -E " hp | sp " gpurun_out/r02l_configs.log
python - <<'PY'
import importlib, sys, os, torch
sys.path.insert(0, os.getcwd())
pkg = importlib.import_module("ultimate-spmv_b200"); eng, capi = pkg.engine, pkg.capi
def timeit(fn, n=100):
    for _ in range(10): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
for pts, n in ((7, 256), (27, 160)):
    mtx = eng.MtxData.stencil(pts, n, n, n)
    scs = eng.convert_to_scs(mtx, 32, 1, "hp"); eng.permute_scs_cols(scs); del mtx
    x = torch.full((scs.n_rows_padded,), 0.5, dtype=torch.float16, device="cuda"); y = torch.zeros_like(x)
    nb = scs.n_elements * 6 + 8 * scs.n_chunks + 4 * scs.n_rows_padded
    for opt in (1, 0):
        capi.set_option("pair_hp", opt)
        us = timeit(lambda: eng.spmv(scs, x, y))
        print(f"hp {pts}-pt {n}^3 C32 pair_hp={opt}: {us:.1f} us  {nb / us / 1e3:.0f} GB/s ({nb / us / 1e3 / 6458.1:.2f} of measured peak)", flush=True)
    capi.set_option("pair_hp", 1)
    del scs, x, y
PY
