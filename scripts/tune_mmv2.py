"""Row-major SpMMV sweep over the stage size / ring depth / warps per CTA variants (run under gpurun).
A far-row L1::no_allocate gather policy was also measured here (round 1): slower in every case, removed."""
import importlib, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
pkg = importlib.import_module("ultimate-spmv_b200"); eng, capi = pkg.engine, pkg.capi
N = int(os.environ.get("GRID", "256"))
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
TD = {"dp": torch.float64, "sp": torch.float32}
res = {}
cases = [("dp", 8), ("dp", 4), ("sp", 8), ("sp", 16), ("dp", 16), ("dp", 2)]
cur = None
for vt, bvs in cases:
    if cur != vt:
        mtx = eng.MtxData.stencil(7, N, N, N); scs = eng.convert_to_scs(mtx, 32, 1, vt); eng.permute_scs_cols(scs); del mtx; cur = vt
    ld = scs.n_rows_padded
    X = torch.full((ld * bvs,), 1.0, dtype=TD[vt], device="cuda"); Y = torch.zeros_like(X)
    line = f"spmmv {vt} b{bvs} rowwise: "
    for var in (1, 2, 6, 8, 9, 10, 11, 12, 13, 14):
        capi.set_option("mmv_variant", var)
        us = timeit(lambda: eng.spmmv(scs, X, Y, bvs, ld, "rowwise"))
        res[f"{vt}|b{bvs}|v{var}"] = us
        line += f"v{var}={us:.0f} "
    print(line, flush=True); del X, Y
    print("   best:", min((v, k) for k, v in res.items() if k.startswith(f"{vt}|b{bvs}|")), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "tune_mmv2.json"), "w"), indent=1)
