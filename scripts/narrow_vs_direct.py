"""C = 16 / 8 on the 256^3 Laplacian: narrow-chunk streamed kernel vs direct kernel (option scs_stream_wide)."""
import importlib, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
pkg = importlib.import_module("ultimate-spmv_b200"); eng, capi = pkg.engine, pkg.capi
TD = {"dp": torch.float64, "sp": torch.float32, "hp": torch.float16}
def timeit(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
for C in (16,):
    for vt in ("dp", "sp", "hp"):
        mtx = eng.MtxData.stencil(7, 256, 256, 256)
        scs = eng.convert_to_scs(mtx, C, 1, vt); eng.permute_scs_cols(scs); del mtx
        x = torch.full((scs.n_rows_padded,), 0.5, dtype=TD[vt], device="cuda"); y = torch.zeros_like(x)
        out = []
        for wide in (1, 0):
            capi.set_option("scs_stream_wide", wide)
            out.append(timeit(lambda: eng.spmv(scs, x, y)))
        print(f"C={C} {vt}: streamed {out[0]:.1f} us, direct {out[1]:.1f} us", flush=True)
        del scs, x, y
