"""Column-banded plan: quick correctness + timing probe.  python scripts/banded_probe.py [log2 rows] [bands ...]"""
import importlib, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
pkg = importlib.import_module("ultimate-spmv_b200"); eng, mats = pkg.engine, pkg.matrices
L = int(sys.argv[1]) if len(sys.argv) > 1 else 16
bands = [int(v) for v in sys.argv[2:]] or [1, 4]
n = 1 << L; slab = 1 << 20
parts = [mats.powerlaw_coo(n, n * 15, row0=r0, row1=min(n, r0 + slab)) for r0 in range(0, n, slab)]
I = np.concatenate([p[2] + r0 for p, r0 in zip(parts, range(0, n, slab))]).astype(np.int32)
J = np.concatenate([p[3] for p in parts]); V = np.concatenate([p[4] for p in parts]); del parts
mtx = eng.MtxData.from_host(n, n, I, J, V)
x = torch.rand(n, dtype=torch.float64, device="cuda") - 0.5
ref = None
for ap in (None, "ap[dp_sp_hp]"):
    for K in bands:
        t0 = time.time(); plan = eng.BandedPlan(mtx, 32, 16384, "dp", ap=ap, t1=1.0, t2=1e-2, n_bands=K); tb = time.time() - t0
        y = torch.zeros(plan.n_rows_padded, dtype=torch.float64, device="cuda")
        for _ in range(3): plan.spmv(x, y)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): plan.spmv(x, y)
        e1.record(); torch.cuda.synchronize()
        yu = y[torch.from_numpy(plan.old_to_new.astype(np.int64)).cuda()]
        if ref is None: ref = yu.clone()
        err = float((yu - ref).abs().max() / ref.abs().max())
        print(f"{ap or 'dp'} K={plan.n_bands}: {e0.elapsed_time(e1) / 10 * 1e3:.0f} us per SpMV, slots {plan.n_elements} ({plan.n_elements / plan.nnz:.2f} per nnz), "
              f"build {tb:.2f} s, max rel diff to first {err:.2e}", flush=True)
        del plan, y
