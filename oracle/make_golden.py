"""TEST INFRASTRUCTURE ONLY.  Generates tests/golden/*.npz by running the UNMODIFIED reference
(oracle/_ref, compiled from /root/reference by oracle/Makefile) in the dev container.

    python -m oracle.make_golden

Outputs (committed; /root/reference does not exist on the GPU box):
  tests/golden/matrices.npz      COO (after read_mtx: symmetric expansion + stable row sort,
                                 utilities.hpp:2148-2309) of the ten matrices/*.mtx fixtures
  tests/golden/ref_scs.npz       for (matrix, C, sigma, vt): chunk_ptrs, chunk_lengths, old_to_new, SHA-256 of
                                 col_idxs / values (after permute_scs_cols), and y = A x in user order for a seeded x
  tests/golden/ref_stdsort.npz   old_to_new of C=1, sigma=n builds for tie-heavy / adversarial count vectors, which
                                 pins libstdc++ 13.3's std::sort tie order (introsort + heapsort fallback)
  tests/golden/ref_dist.npz      seg_work_sharing_arr and collect_local_needed_heri results (P = 2, 3, 4)
  tests/golden/ref_ap.npz        adaptive-precision y (interface.hpp kernels) for dp_sp / dp_hp / sp_hp / dp_sp_hp
  tests/golden/ref_ingest.npz    raw file-order entries of the ten matrices + values after the reference's equilibrate_matrix
"""
from __future__ import annotations

import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.bindings import Ref, RefIface, build  # noqa: E402

REF_MAT = "/root/reference/matrices"
OUT = os.path.join(ROOT, "tests", "golden")
NAMES = ["FDM-2d-16", "bcsstk13", "impcol_e", "matrix1", "matrix1int", "matrix1ones", "matrix_band_klein", "myBigMat", "myMat", "mySymmMat"]
NPT = {"dp": np.float64, "sp": np.float32, "hp": np.float16}


def sha(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), dtype=np.uint8).copy()


def seeded_x(n, seed=1234):
    return np.random.default_rng(seed).uniform(-1.0, 1.0, n)


def killer(n):
    k = n // 2
    a = [0] * n
    for i in range(1, k + 1):
        if i % 2 == 1:
            a[i - 1] = i
            a[i] = k + i
        a[k + i - 1] = 2 * i
    return np.array(a)


def ingest_fixtures(ref):
    """tests/golden/ref_ingest.npz: per matrix the raw file-order entries (what the device ingest receives; parsed with the
    product's text parser) and the values after the reference's equilibrate_matrix applied to its own read_mtx result."""
    import importlib
    mats = importlib.import_module("ultimate-spmv_b200.matrices")
    g = {}
    for name in NAMES:
        path = os.path.join(REF_MAT, name + ".mtx")
        n, nc, I, J, V, sym = mats.read_mtx_entries(path)
        g[f"{name}__n"] = np.int64(n)
        g[f"{name}__sym"] = np.int64(1 if sym else 0)
        g[f"{name}__I"], g[f"{name}__J"], g[f"{name}__V"] = I, J, V
        rn, rc, rI, rJ, rV = ref.read_mtx(path)
        g[f"{name}__equil"] = ref.equilibrate(rn, rc, rI, rJ, rV)
    np.savez_compressed(os.path.join(OUT, "ref_ingest.npz"), **g)


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "ingest":  # only the ingest / equilibrate fixtures
        build("all")
        ingest_fixtures(Ref("col"))
        return
    build("all")
    os.makedirs(OUT, exist_ok=True)
    ref = Ref("col")
    iface = RefIface()

    # ---- matrices ------------------------------------------------------------------------------
    mats = {}
    store = {}
    for name in NAMES:
        n, nc, I, J, V = ref.read_mtx(os.path.join(REF_MAT, name + ".mtx"))
        assert n == nc
        mats[name] = (n, I, J, V)
        store[f"{name}__n"] = np.int64(n)
        store[f"{name}__I"] = I
        store[f"{name}__J"] = J
        store[f"{name}__V"] = V
    np.savez_compressed(os.path.join(OUT, "matrices.npz"), **store)

    # ---- convert_to_scs + permute_scs_cols + SpMV ------------------------------------------------
    g = {}
    cases = []
    for name in NAMES:
        n, I, J, V = mats[name]
        for vt in ("dp", "sp", "hp"):
            for (C, sigma) in [(1, 1), (4, 1), (4, 8), (8, 32), (32, 1), (32, 512), (16, 64), (10, 30), (64, 64), (128, 256)]:
                if vt == "hp" and name == "bcsstk13":
                    continue  # values overflow fp16
                if name == "bcsstk13" and (C, sigma) not in [(1, 1), (32, 1), (32, 512), (16, 64), (128, 256)]:
                    continue
                s = ref.convert_to_scs(n, n, I, J, V, C, sigma, vt, permute_cols=True)
                x = seeded_x(n).astype(NPT[vt])
                xp = np.zeros(s.n_rows_padded, NPT[vt])
                xp[s.old_to_new] = x
                yp = ref.spmv_scs(s, xp, adv=False)
                y = yp[s.old_to_new]
                key = f"{name}|{C}|{sigma}|{vt}"
                cases.append(key)
                g[key + "|dims"] = np.array([s.n_rows_padded, s.n_chunks, s.n_elements, s.nnz], np.int64)
                g[key + "|chunk_ptrs"] = s.chunk_ptrs
                g[key + "|chunk_lengths"] = s.chunk_lengths
                g[key + "|old_to_new"] = s.old_to_new
                g[key + "|col_sha"] = sha(s.col_idxs)
                g[key + "|val_sha"] = sha(s.values)
                g[key + "|y"] = y
    g["cases"] = np.array(cases)
    np.savez_compressed(os.path.join(OUT, "ref_scs.npz"), **g)

    # ---- std::sort tie order -------------------------------------------------------------------------
    st = {}
    rng = np.random.default_rng(7)
    pats = {}
    for n in (17, 33, 100, 512, 1000, 4096):
        pats[f"ties3_{n}"] = rng.integers(0, 3, n)
        pats[f"const_{n}"] = np.full(n, 7)
        pats[f"rand_{n}"] = rng.integers(0, 50, n)
        pats[f"organ_{n}"] = np.minimum(np.arange(n), np.arange(n)[::-1])
        k = killer(n)
        pats[f"killer_{n}"] = k.max() - k
    for name, cnt in pats.items():
        n = len(cnt)
        I = np.repeat(np.arange(n), cnt).astype(np.int32)
        if len(I) == 0:
            continue
        J = np.zeros(len(I), np.int32)
        s = ref.convert_to_scs(n, n, I, J, np.ones(len(I)), 1, n, "dp")
        st[name + "|cnt"] = cnt.astype(np.int32)
        st[name + "|old_to_new"] = s.old_to_new
    np.savez_compressed(os.path.join(OUT, "ref_stdsort.npz"), **st)

    # ---- row partitioning + halo bookkeeping ------------------------------------------------------------
    d = {}
    dcases = []
    for name in ("FDM-2d-16", "bcsstk13", "impcol_e", "matrix1"):
        n, I, J, V = mats[name]
        for P in (2, 3, 4):
            for method in ("seg-rows", "seg-nnz"):
                wsa = ref.seg_work_sharing_arr(method, n, I, P)
                key = f"{name}|{P}|{method}"
                d[key + "|wsa"] = wsa
                if name == "bcsstk13" and P != 4:
                    continue
                for (C, sigma) in ((1, 1), (8, 16)):
                    for r in range(P):
                        lI, lJ, lV = ref.seg_mtx(n, I, J, V, wsa, r)
                        n_loc = int(wsa[r + 1] - wsa[r])
                        h = ref.build_handle(n_loc, n, lI, lJ, lV, C, sigma, "dp")
                        s0 = ref.export(h, "dp")
                        need, cum = ref.collect_halo(h, wsa, r)
                        ref.lib.ref_scs_permute_cols(h, s0.old_to_new.ctypes.data_as(__import__("ctypes").c_void_p))
                        s1 = ref.export(h, "dp")
                        ref.lib.ref_scs_free(h)
                        k2 = f"{key}|{C}|{sigma}|{r}"
                        dcases.append(k2)
                        d[k2 + "|col_idxs"] = s1.col_idxs
                        d[k2 + "|old_to_new"] = s1.old_to_new
                        d[k2 + "|recv_cumsum"] = cum
                        d[k2 + "|need_flat"] = np.concatenate(need) if sum(len(a) for a in need) else np.zeros(0, np.int32)
                        d[k2 + "|need_ptr"] = np.cumsum([0] + [len(a) for a in need]).astype(np.int32)
    d["cases"] = np.array(dcases)
    np.savez_compressed(os.path.join(OUT, "ref_dist.npz"), **d)

    # ---- adaptive precision (interface.hpp kernels, real std::abs thresholds restated in the caller) ------
    a = {}
    acases = []
    for name in ("FDM-2d-16", "impcol_e", "matrix1", "myBigMat"):
        n, I, J, V = mats[name]
        absv = np.abs(V)
        t1 = float(np.quantile(absv, 0.66))
        t2 = float(np.quantile(absv, 0.33))
        if t1 <= t2:
            t1 = t2 * 2 + 1e-3
        for mode, mname in enumerate(("ap[dp_sp]", "ap[dp_hp]", "ap[sp_hp]", "ap[dp_sp_hp]")):
            for (C, sigma) in ((1, 1), (4, 8), (32, 64)):
                if mode == 3:
                    part = np.where(absv >= t1, 0, np.where((absv <= t1) & (absv >= t2), 1, 2))
                else:
                    hi, lo = (1, 2) if mode == 2 else (0, 1 if mode == 0 else 2)
                    part = np.where(absv >= t1, hi, lo)
                used = ((0, 1), (0, 2), (1, 2), (0, 1, 2))[mode]
                first = used[0]
                sel = part == first
                s_first = ref.convert_to_scs(n, n, I[sel], J[sel], V[sel], C, sigma, ("dp", "sp", "hp")[first])
                parts = [None, None, None]
                parts[first] = s_first
                if (s_first.old_to_new >= n).any():
                    # a real row landed on a padding position: the reference's fixed_permutation branch then zeroes
                    # its count and the fill overruns the chunk (utilities.hpp:1919-1922,2029) -> undefined behaviour
                    continue
                for p in used[1:]:
                    sel = part == p
                    parts[p] = ref.convert_to_scs(n, n, I[sel], J[sel], V[sel], C, sigma, ("dp", "sp", "hp")[p], fixed_perm=s_first.old_to_new)
                x = seeded_x(n)
                xp = np.zeros(s_first.n_rows_padded)
                xp[:n] = x  # AP oracle works in unpermuted-COLUMN space (hard part 6)
                yp = iface.ap_scs(mode, parts[0], parts[1], parts[2], xp, xp.astype(np.float32))
                key = f"{name}|{mname}|{C}|{sigma}"
                acases.append(key)
                a[key + "|t"] = np.array([t1, t2])
                a[key + "|part"] = part.astype(np.int8)
                a[key + "|perm"] = s_first.old_to_new
                a[key + "|y"] = yp[s_first.old_to_new]
    a["cases"] = np.array(acases)
    np.savez_compressed(os.path.join(OUT, "ref_ap.npz"), **a)
    ingest_fixtures(ref)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
