"""CPU: host-side logic of the product library that needs no GPU — uspmv_seg_work_sharing_arr (seg_work_sharing_arr,
mpi_funcs.hpp:424-622) — against the fixtures produced by the unmodified reference (tests/golden/ref_dist.npz) and the oracle."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import GOLDEN, load_matrix


def _wsa(capi, method, n, I, P):
    I = np.ascontiguousarray(I, np.int32)
    out = np.zeros(P + 1, np.int32)
    capi.call("uspmv_seg_work_sharing_arr", capi.SEG_NNZ if method == "seg-nnz" else capi.SEG_ROWS, int(n), int(len(I)),
              I.ctypes.data_as(C.c_void_p), int(P), out.ctypes.data_as(C.c_void_p))
    return out


def test_work_sharing_arr_matches_reference_fixture(pkg):
    capi = pkg.capi
    z = np.load(os.path.join(GOLDEN, "ref_dist.npz"))
    keys = [k[:-4] for k in z.files if k.endswith("|wsa")]
    assert len(keys) >= 10
    for key in keys:
        name, P, method = key.split("|")
        n, nc, I, J, V = load_matrix(name)
        assert np.array_equal(_wsa(capi, method, n, I, int(P)), z[key + "|wsa"]), key


@pytest.mark.parametrize("method", ["seg-rows", "seg-nnz"])
def test_work_sharing_arr_matches_oracle_on_uneven_rows(pkg, orc, method):
    capi = pkg.capi
    rng = np.random.default_rng(5)
    for n, P in ((1, 1), (7, 3), (100, 8), (1000, 7), (4096, 16)):
        if P > n:
            continue
        cnt = rng.integers(0, 9, n)
        cnt[rng.integers(0, n, max(1, n // 50))] += rng.integers(50, 400)
        cnt[-1] = max(cnt[-1], 1)  # the reference's partitioner expects the last row to exist in the COO
        cnt[0] = max(cnt[0], 1)
        I = np.repeat(np.arange(n), cnt).astype(np.int32)
        got = _wsa(capi, method, n, I, P)
        ref = orc.seg_work_sharing_arr(method, n, I, P)
        assert np.array_equal(got, ref), (method, n, P)
        assert got[0] == 0 and got[-1] == n and np.all(np.diff(got) >= 0)
