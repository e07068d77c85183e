"""Single-GPU timing of the other BASELINE.json configs (run under gpurun): SpMMV (config 3), CRS, other C, hp, and
adaptive precision on the power-law matrix (config 4).  Prints one line per case and writes gpurun_out/configs.json."""
import importlib, json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
pkg = importlib.import_module("ultimate-spmv_b200"); eng, capi, mats = pkg.engine, pkg.capi, pkg.matrices
which = sys.argv[1:] or ["spmmv", "formats", "ap"]
N = int(os.environ.get("GRID", "256"))
PEAK = 6458.1
res = []
ctx = eng.default_context(0)

def timeit(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e-3

def report(name, sec, nbytes, flops):
    r = {"case": name, "us": sec * 1e6, "gbs": nbytes / sec / 1e9, "frac_of_measured_peak": nbytes / sec / 1e9 / PEAK, "gflops": flops / sec / 1e9}
    res.append(r)
    print(f"{name:48s} {r['us']:9.1f} us {r['gbs']:8.1f} GB/s ({r['frac_of_measured_peak']:.2f} of measured) {r['gflops']:9.1f} GFLOP/s", flush=True)

TD = {"dp": torch.float64, "sp": torch.float32, "hp": torch.float16}
VS = {"dp": 8, "sp": 4, "hp": 2}

if "spmmv" in which:
    for vt in ("dp", "sp"):
        mtx = eng.MtxData.stencil(7, N, N, N)
        scs = eng.convert_to_scs(mtx, 32, 1, vt); eng.permute_scs_cols(scs); del mtx
        for bvs in (4, 8):
            for layout in ("rowwise", "colwise"):
                ld = scs.n_rows_padded
                X = torch.full((ld * bvs,), 5.0, dtype=TD[vt], device="cuda"); Y = torch.zeros_like(X)
                sec = timeit(lambda: eng.spmmv(scs, X, Y, bvs, ld, layout))
                nb = scs.n_elements * (VS[vt] + 4) + 8 * scs.n_chunks + 2 * bvs * VS[vt] * ld
                report(f"spmmv 7pt{N} C32 {vt} bvs{bvs} {layout}", sec, nb, 2.0 * scs.nnz * bvs)
                del X, Y
        del scs

if "formats" in which:
    for (C, sigma, vt) in ((1, 1, "dp"), (16, 1, "dp"), (8, 1, "dp"), (16, 1, "sp"), (16, 1, "hp"), (64, 1, "dp"), (128, 1, "dp"), (64, 1, "sp"), (128, 1, "sp"), (64, 1, "hp"), (32, 1, "sp"), (32, 1, "hp"), (1, 1, "sp"), (32, 512, "dp")):
        mtx = eng.MtxData.stencil(7, N, N, N)
        t0 = time.time(); scs = eng.convert_to_scs(mtx, C, sigma, vt); eng.permute_scs_cols(scs); torch.cuda.synchronize(); tb = time.time() - t0
        del mtx
        x = torch.full((scs.n_rows_padded,), 5.0 if vt != "hp" else 0.5, dtype=TD[vt], device="cuda"); y = torch.zeros_like(x)
        sec = timeit(lambda: eng.spmv(scs, x, y))
        nb = scs.n_elements * (VS[vt] + 4) + 8 * scs.n_chunks + 2 * VS[vt] * scs.n_rows_padded
        report(f"spmv 7pt{N} C{C} s{sigma} {vt} (build {tb*1e3:.0f} ms)", sec, nb, 2.0 * scs.nnz)
        if C == 1 and vt == "dp":  # the raw-array entry point uspmv_csr_gpu on caller-owned arrays of exactly nnz elements
            e = scs.export()
            nnz = int(e.chunk_ptrs[scs.n_rows])
            rp, ci, va = (torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in (e.chunk_ptrs[: scs.n_rows + 1], e.col_idxs[:nnz], e.values[:nnz]))
            del e
            sec = timeit(lambda: eng.uspmv_csr_gpu(scs.n_rows, rp, ci, va, x, y))
            report(f"uspmv_csr_gpu (caller-owned arrays) 7pt{N} dp", sec, nb, 2.0 * scs.nnz)
            del rp, ci, va
        del scs, x, y

if "ap" in which:
    n = int(os.environ.get("PL_ROWS", str(1 << 22))); target = n * 15
    t0 = time.time()
    parts = []
    slab = 1 << 20
    for r0 in range(0, n, slab):
        parts.append(mats.powerlaw_coo(n, target, row0=r0, row1=min(n, r0 + slab)))
    I = np.concatenate([p[2] + r0 for p, r0 in zip(parts, range(0, n, slab))]).astype(np.int32)
    J = np.concatenate([p[3] for p in parts]); V = np.concatenate([p[4] for p in parts]); del parts
    print(f"power-law matrix: {n} rows, {len(I)} nnz generated on the host in {time.time()-t0:.1f} s", flush=True)
    mtx = eng.MtxData.from_host(n, n, I, J, V); nnz = len(I); del I, J, V
    SIG = int(os.environ.get("PL_SIGMA", "512"))
    for mode in ("ap[dp_sp_hp]", "ap[dp_sp]", None):
        if mode is None:
            t0 = time.time(); scs = eng.convert_to_scs(mtx, 32, SIG, "dp"); torch.cuda.synchronize(); print(f"plain build {time.time()-t0:.2f} s")
            x = torch.full((scs.n_rows_padded,), 1.0, dtype=torch.float64, device="cuda"); y = torch.zeros_like(x)
            sec = timeit(lambda: eng.spmv_unpermuted(scs, x, y) if False else eng.spmv(scs, x, y), 20)
            nb = scs.n_elements * 12 + 8 * scs.n_chunks + 16 * scs.n_rows_padded
            report(f"powerlaw {n} plain dp C32 s{SIG} (beta {scs.nnz/scs.n_elements:.3f})", sec, nb, 2.0 * nnz)
            continue
        t0 = time.time()
        coos = eng.partition_precisions(mtx, mode, 1.0, 1e-2)
        used = [k for k in range(3) if coos[k] is not None]
        vts = ("dp", "sp", "hp")
        P = [None] * 3
        P[used[0]] = eng.convert_to_scs(coos[used[0]], 32, SIG, vts[used[0]])
        perm = P[used[0]].export().old_to_new
        for k in used[1:]:
            P[k] = eng.convert_to_scs(coos[k], 32, SIG, vts[k], fixed_permutation=perm)
        torch.cuda.synchronize(); tb = time.time() - t0
        n_pad = P[used[0]].n_rows_padded
        x = torch.full((max(n_pad, n),), 1.0, dtype=torch.float64, device="cuda"); y = torch.zeros(n_pad, dtype=torch.float64, device="cuda")
        sec = timeit(lambda: eng.ap_spmv(mode, P[0], P[1], P[2], x, y), 20)
        nb = sum(P[k].n_elements * ((8, 4, 2)[k] + 4) + 8 * P[k].n_chunks for k in used) + 16 * n_pad
        frac = [coos[k].nnz / nnz if coos[k] is not None else 0 for k in range(3)]
        report(f"powerlaw {n} {mode} C32 s{SIG} split {frac[0]:.2f}/{frac[1]:.2f}/{frac[2]:.2f} (build {tb:.1f} s)", sec, nb, 2.0 * nnz)
        del coos, P

if "cusparse" in which:
    # second GPU baseline (the reference's cuSPARSE comparison mode, utilities.hpp:3380-3550): cusparseSpMV on the same matrix in
    # CSR, reached through torch's sparse CSR tensor (library code, not part of the product path)
    for vt in ("dp", "sp"):
        mtx = eng.MtxData.stencil(7, N, N, N)
        I, J, V = mtx.to_host()
        n = mtx.n_rows
        crow = np.zeros(n + 1, np.int64); np.add.at(crow, I.astype(np.int64) + 1, 1); crow = np.cumsum(crow)
        A = torch.sparse_csr_tensor(torch.from_numpy(crow.astype(np.int32)).cuda(), torch.from_numpy(J).cuda(),
                                    torch.from_numpy(V.astype({"dp": np.float64, "sp": np.float32}[vt])).cuda(), size=(n, n))
        x = torch.full((n,), 5.0, dtype=TD[vt], device="cuda")
        sec = timeit(lambda: torch.mv(A, x), 30)
        nb = len(I) * (VS[vt] + 4) + 4 * (n + 1) + 2 * VS[vt] * n
        report(f"cuSPARSE CSR SpMV (torch.mv) 7pt{N} {vt}", sec, nb, 2.0 * len(I))
        scs = eng.convert_to_scs(mtx, 32, 1, vt); eng.permute_scs_cols(scs); del mtx
        xs = torch.full((scs.n_rows_padded,), 5.0, dtype=TD[vt], device="cuda"); ys = torch.zeros_like(xs)
        sec = timeit(lambda: eng.spmv(scs, xs, ys), 30)
        nb = scs.n_elements * (VS[vt] + 4) + 8 * scs.n_chunks + 2 * VS[vt] * scs.n_rows_padded
        report(f"this library SELL-32 SpMV 7pt{N} {vt}", sec, nb, 2.0 * scs.nnz)
        del A, x, scs, xs, ys, I, J, V

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "configs_" + "_".join(which) + ".json"), "w"), indent=1)
