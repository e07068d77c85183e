"""Row-major SpMMV: the 8-warp kernel with a minimum of resident CTAs per SM in __launch_bounds__ (variants 15-20) against the
variants without one (1, 2, 6) — ptxas keeps the gathers in flight only when it is given a register budget (run under gpurun)."""
import importlib, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
pkg = importlib.import_module("ultimate-spmv_b200"); eng, capi = pkg.engine, pkg.capi
N = int(os.environ.get("GRID", "256"))
def timeit(fn, n=30):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
TD = {"dp": torch.float64, "sp": torch.float32, "hp": torch.float16}
res = {}
cases = [("sp", 8), ("sp", 4), ("sp", 16), ("sp", 2), ("dp", 8), ("dp", 4), ("dp", 2), ("dp", 16), ("hp", 16), ("hp", 8)]
cur = None
for vt, bvs in cases:
    if cur != vt:
        mtx = eng.MtxData.stencil(7, N, N, N); scs = eng.convert_to_scs(mtx, 32, 1, vt); eng.permute_scs_cols(scs); del mtx; cur = vt
    ld = scs.n_rows_padded
    X = torch.rand((ld * bvs,), dtype=torch.float32, device="cuda").to(TD[vt]); Y = torch.zeros_like(X)
    line = f"spmmv {vt} b{bvs} rowwise: "
    ref = None
    for var in (0, 1, 2, 6, 15, 16, 17, 18, 19, 20):
        capi.set_option("mmv_variant", var)
        us = timeit(lambda: eng.spmmv(scs, X, Y, bvs, ld, "rowwise"))
        if ref is None: ref = Y.clone()
        same = bool(torch.equal(Y.view(torch.uint8), ref.view(torch.uint8)))
        res[f"{vt}|b{bvs}|v{var}"] = us
        line += f"v{var}={us:.0f}{'' if same else '(!)'} "
    print(line, flush=True); del X, Y
    print("   best:", min((v, k) for k, v in res.items() if k.startswith(f"{vt}|b{bvs}|")), flush=True)
capi.set_option("mmv_variant", 0)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "r02r_tune_mmv3.json"), "w"), indent=1)
