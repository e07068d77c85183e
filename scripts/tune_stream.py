"""GPU tuning sweep for the C=32 SpMV kernels (run under gpurun): times every stream variant x CTAs/SM."""
import importlib, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("ultimate-spmv_b200")
eng, capi = pkg.engine, pkg.capi
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
sigma = int(sys.argv[2]) if len(sys.argv) > 2 else 1
vt = sys.argv[3] if len(sys.argv) > 3 else "dp"
pts = int(sys.argv[4]) if len(sys.argv) > 4 else 7
r = eng.SingleGpuSpmv(eng.default_context(0), pts, n, 32, sigma, vt)
vsize = {"dp": 8, "sp": 4, "hp": 2}[vt]
nbytes = r.n_elements * (vsize + 4) + r.n_chunks * 8 + vsize * (r.n_cols_local + r.n_rows_padded)
y_ref = None
res = []
for name, variant, bps in [("direct", -1, 0)] + [(f"v{v}", v, b) for v in (0, 10, 11, 12) for b in (2, 3, 4)]:
    if variant < 0:
        capi.set_option("scs_stream", 0)
    else:
        capi.set_option("scs_stream", 1); capi.set_option("stream_variant", variant); capi.set_option("stream_blocks_per_sm", bps)
    try:
        for _ in range(5): r.step()
        torch.cuda.synchronize()
        ms = r.time_kernel(50)
    except Exception as e:
        print(name, bps, "FAILED", e); continue
    y = r.y.clone()
    ok = True if y_ref is None else bool(torch.equal(y, y_ref))
    if y_ref is None: y_ref = y
    print(f"{name:7s} bps={bps} {ms*1e3:8.1f} us  {nbytes/ms/1e6:8.1f} GB/s  same_y={ok}", flush=True)
    res.append({"name": name, "bps": bps, "us": ms * 1e3, "gbs": nbytes / ms / 1e6, "same_y": ok})
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", f"tune_stream_{pts}pt_{n}_{sigma}_{vt}.json"), "w"), indent=1)
