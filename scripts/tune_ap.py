"""Power-law matrix, fused ap[dp_sp_hp] SpMV: ring/register variants x segment length (run under gpurun)."""
import importlib, json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
pkg = importlib.import_module("ultimate-spmv_b200"); eng, capi, mats = pkg.engine, pkg.capi, pkg.matrices
n = 1 << int(os.environ.get("PL_LOG2", "22")); SIG = int(os.environ.get("PL_SIGMA", "16384")); slab = 1 << 20
parts = [mats.powerlaw_coo(n, n * 15, row0=r0, row1=min(n, r0 + slab)) for r0 in range(0, n, slab)]
I = np.concatenate([p[2] + r0 for p, r0 in zip(parts, range(0, n, slab))]).astype(np.int32)
J = np.concatenate([p[3] for p in parts]); V = np.concatenate([p[4] for p in parts]); del parts
mtx = eng.MtxData.from_host(n, n, I, J, V); nnz = len(I); del I, J, V
def timeit(fn, k=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k * 1e3
res = {}
for mode in os.environ.get("PL_MODES", "ap[dp_sp_hp],ap[dp_sp]").split(","):
    coos = eng.partition_precisions(mtx, mode, 1.0, 1e-2)
    used = [k for k in range(3) if coos[k] is not None]
    vts = ("dp", "sp", "hp")
    P = [None] * 3
    P[used[0]] = eng.convert_to_scs(coos[used[0]], 32, SIG, vts[used[0]])
    perm = P[used[0]].export().old_to_new
    for k in used[1:]:
        P[k] = eng.convert_to_scs(coos[k], 32, SIG, vts[k], fixed_permutation=perm)
    n_pad = P[used[0]].n_rows_padded
    x = torch.full((max(n_pad, n),), 1.0, dtype=torch.float64, device="cuda"); y = torch.zeros(n_pad, dtype=torch.float64, device="cuda")
    print(mode, "n_elements", [p.n_elements if p is not None else 0 for p in P], "nnz", nnz, flush=True)
    for gran in [int(v) for v in os.environ.get("PL_L2", "0").split(",")]:
        if gran:
            capi.set_option("l2_fetch_granularity", gran)
            print(f" cudaLimitMaxL2FetchGranularity = {gran}", flush=True)
        for var in [int(v) for v in os.environ.get("PL_VARIANTS", "0,1,2,3,4").split(",")]:
            line = f"  variant {var}: "
            for split in [int(v) for v in os.environ.get("PL_SPLITS", "0,64,128,256,512").split(",")]:
                capi.set_option("ap_variant", var); capi.set_option("split_long_chunks", split)
                us = timeit(lambda: eng.ap_spmv(mode, P[0], P[1], P[2], x, y))
                res[f"{mode}|v{var}|split{split}|l2gran{gran}"] = us
                line += f"split{split}={us:.0f} "
            print(line, flush=True)
    del coos, P
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", f"tune_ap_{n}.json"), "w"), indent=1)
