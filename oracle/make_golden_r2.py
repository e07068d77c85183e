"""TEST INFRASTRUCTURE ONLY.  Round-2 additions to tests/golden/, produced by running the UNMODIFIED reference (oracle/_ref):

    python -m oracle.make_golden_r2

  tests/golden/ref_xinit.npz    init_std_vec_with_ptr_or_value(..., random) (utilities.hpp:880-981): default-seeded mt19937 +
                                uniform_real_distribution<double>(min, max) + the padding rule, dp / sp, colwise / rowwise block vectors
  tests/golden/ref_seg_mtx.npz  seg_mtx_struct + localize_row_idx (mpi_funcs.hpp:636-674,862-877) for every rank of P = 2, 3 (seg-rows and
                                seg-nnz work_sharing_arr from the reference's seg_work_sharing_arr) on the matrices/*.mtx fixtures
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.bindings import Ref  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
XINIT_CASES = [(-2.5, 7.25, 1000, 1024, 1), (-2.5, 7.25, 1000, 1056, 3), (0.001, 9834.5, 37, 64, 4), (-1.0, 6.0, 20000, 20032, 2)]


def main():
    refs = {"col": Ref("col"), "row": Ref("row")}
    g = {}
    for lay, ref in refs.items():
        for vt in ("dp", "sp"):
            for k, (lo, hi, n_rows, n_pad, bvs) in enumerate(XINIT_CASES):
                g[f"{lay}|{vt}|{k}"] = ref.random_x(vt, lo, hi, n_pad * bvs, n_rows, n_pad, bvs)
    g["cases"] = np.array(XINIT_CASES, dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, "ref_xinit.npz"), **g)

    z = np.load(os.path.join(OUT, "matrices.npz"))
    names = sorted({k.split("__")[0] for k in z.files})
    ref = refs["col"]
    s = {}
    for name in names:
        n = int(z[f"{name}__n"])
        I, J, V = z[f"{name}__I"], z[f"{name}__J"], z[f"{name}__V"]
        for P in (2, 3):
            if n < P:
                continue
            for seg in ("seg-rows", "seg-nnz"):
                wsa = ref.seg_work_sharing_arr(seg, n, I, P)
                # the reference looks the slab up with a linear search for the ROW IDS wsa[r] / wsa[r+1] (get_index): only defined
                # when those rows hold an element
                rows = set(I.tolist())
                if any(int(w) not in rows for w in wsa[:-1]) or np.any(np.diff(wsa) <= 0):
                    continue
                s[f"{name}|{P}|{seg}|wsa"] = wsa
                for r in range(P):
                    lI, lJ, lV = ref.seg_mtx(n, I, J, V, wsa, r)
                    s[f"{name}|{P}|{seg}|{r}|I"], s[f"{name}|{P}|{seg}|{r}|J"], s[f"{name}|{P}|{seg}|{r}|V"] = lI, lJ, lV
    np.savez_compressed(os.path.join(OUT, "ref_seg_mtx.npz"), **s)
    print(f"ref_xinit.npz: {len(g) - 1} vectors; ref_seg_mtx.npz: {len(s)} arrays")


if __name__ == "__main__":
    main()
