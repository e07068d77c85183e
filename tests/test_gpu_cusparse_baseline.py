"""The reference's second GPU baseline — its USE_CUSPARSE comparison mode (code/utilities.hpp:3380-3550: cusparseSpMV on CSR, or on
cuSPARSE's sliced-ELL created directly from the SELL-C-sigma arrays) — restated in oracle/cusparse_driver.cu as measurement
infrastructure.  Here: the device-built SELL-C-sigma structure is accepted by cusparseCreateSlicedEll as it stands (the same claim the
reference makes for its own structure) and cuSPARSE's product agrees with this library's kernel and with the oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cusp():
    from oracle import bindings
    if not bindings.cusparse_available():
        pytest.skip("oracle/_ref/libuspmv_cusparse.so not built")
    return bindings.CuSparse()


@pytest.mark.parametrize("vt", ["dp", "sp"])
@pytest.mark.parametrize("kind,C,sigma", [("sell", 32, 1), ("sell", 32, 64), ("sell", 16, 1), ("csr", 1, 1)])
def test_cusparse_product_agrees(eng, orc, cusp, vt, kind, C, sigma):
    import torch as t
    n = 24
    mtx = eng.MtxData.stencil(27, n, n, n)
    scs = eng.convert_to_scs(mtx, C, sigma, vt)
    eng.permute_scs_cols(scs)
    s = scs.export()
    npt = {"dp": np.float64, "sp": np.float32}[vt]
    x = np.random.default_rng(5).uniform(-1.0, 1.0, s.n_rows_padded).astype(npt)
    xd = t.from_numpy(x).cuda()
    yd = t.zeros(s.n_rows_padded, dtype=xd.dtype, device="cuda")
    eng.spmv(scs, xd, yd)
    t.cuda.synchronize()
    ours = yd.cpu().numpy()[: s.n_rows]
    y, ms = cusp.spmv(kind, vt, s.n_rows, s.n_rows_padded, s.nnz, C, s.chunk_ptrs, s.col_idxs, s.values, x, warmup=1, steps=3)
    assert ms > 0
    scale = 27.0 * {"dp": 2.0 ** -52, "sp": 2.0 ** -23}[vt] * 26.0
    assert np.max(np.abs(y - ours)) <= 8 * scale, (kind, C, sigma, vt, float(np.max(np.abs(y - ours))))
