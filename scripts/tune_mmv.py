"""Sweep of the streamed SpMMV variants x CTAs/SM (run under gpurun)."""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
pkg = importlib.import_module("ultimate-spmv_b200"); eng, capi = pkg.engine, pkg.capi
N = 256
def timeit(fn, n=30):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
TD = {"dp": torch.float64, "sp": torch.float32}
for vt in ("dp", "sp"):
    mtx = eng.MtxData.stencil(7, N, N, N); scs = eng.convert_to_scs(mtx, 32, 1, vt); eng.permute_scs_cols(scs); del mtx
    ld = scs.n_rows_padded
    x = torch.full((ld,), 1.0, dtype=TD[vt], device="cuda"); y = torch.zeros_like(x)
    print(f"spmv {vt}: {timeit(lambda: eng.spmv(scs, x, y)):.0f} us", flush=True)
    for bvs in (4, 8):
        for layout in ("rowwise", "colwise"):
            X = torch.full((ld * bvs,), 1.0, dtype=TD[vt], device="cuda"); Y = torch.zeros_like(X)
            line = f"spmmv {vt} b{bvs} {layout}: "
            for var in (1, 2, 5, 6, 7, 8):
                for bps in (0,):
                    capi.set_option("mmv_variant", var); capi.set_option("mmv_blocks_per_sm", bps)
                    line += f"v{var}/b{bps}={timeit(lambda: eng.spmmv(scs, X, Y, bvs, ld, layout)):.0f} "
            print(line, flush=True); del X, Y
