"""A/B timing of SELL-32 SpMV (7-point 256^3) for one build of the library (USPMV_B200_LIB) and a list of stream variants."""
import importlib, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
pkg = importlib.import_module("ultimate-spmv_b200"); eng, capi = pkg.engine, pkg.capi
tag = sys.argv[1]
variants = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0]
TD = {"dp": torch.float64, "sp": torch.float32, "hp": torch.float16}
def timeit(fn, n=100):
    for _ in range(10): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
for vt in ("dp", "sp", "hp"):
    mtx = eng.MtxData.stencil(7, 256, 256, 256)
    scs = eng.convert_to_scs(mtx, 32, 1, vt); eng.permute_scs_cols(scs); del mtx
    x = torch.full((scs.n_rows_padded,), 0.5, dtype=TD[vt], device="cuda"); y = torch.zeros_like(x)
    out = []
    for v in variants:
        try:
            capi.set_option("stream_variant", v)
            out.append(f"v{v}: {timeit(lambda: eng.spmv(scs, x, y)):.1f}")
        except Exception as e:
            out.append(f"v{v}: {e}")
    capi.set_option("stream_variant", 0)
    print(f"{tag:26s} {vt}  " + "  ".join(out) + " us", flush=True)
    del scs, x, y
