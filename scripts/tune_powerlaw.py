"""Power-law matrix: plain dp SpMV over the stream variants (deeper rings help lone warps on very long chunks)."""
import importlib, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
pkg = importlib.import_module("ultimate-spmv_b200"); eng, capi, mats = pkg.engine, pkg.capi, pkg.matrices
n = 1 << 22; slab = 1 << 20
parts = [mats.powerlaw_coo(n, n * 15, row0=r0, row1=min(n, r0 + slab)) for r0 in range(0, n, slab)]
I = np.concatenate([p[2] + r0 for p, r0 in zip(parts, range(0, n, slab))]).astype(np.int32)
J = np.concatenate([p[3] for p in parts]); V = np.concatenate([p[4] for p in parts]); del parts
mtx = eng.MtxData.from_host(n, n, I, J, V); nnz = len(I)
def timeit(fn, k=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k * 1e3
for sigma in (512, 16384):
    scs = eng.convert_to_scs(mtx, 32, sigma, "dp")
    x = torch.full((scs.n_rows_padded,), 1.0, dtype=torch.float64, device="cuda"); y = torch.zeros_like(x)
    nb = scs.n_elements * 12 + 8 * scs.n_chunks + 16 * scs.n_rows_padded
    line = f"sigma {sigma} beta {scs.nnz/scs.n_elements:.3f} bytes {nb/1e9:.2f} GB ideal {nb/6458.1e3:.0f} us: "
    for var, bps in ((0, 2), (1, 3), (2, 2), (6, 4), (7, 2), (8, 1)):
        capi.set_option("stream_variant", var); capi.set_option("stream_blocks_per_sm", bps)
        line += f"v{var}/b{bps}={timeit(lambda: eng.spmv(scs, x, y)):.0f} "
    print(line, flush=True)
