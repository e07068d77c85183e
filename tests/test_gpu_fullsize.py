"""GPU parity at BASELINE.json's FULL size (configs[1]: 7-point Laplacian 256^3, 16.8 M rows, 117 M nnz), where the CPU oracle would
take minutes: size-independent properties instead — known-answer products, the format invariants of SURVEY.md section 8, equality of
differently built formats (sigma, CRS, SpMMV columns), linearity, and precision tolerances."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

N = 256
NR = N ** 3
NNZ = 7 * NR - 6 * N * N           # 117 047 296: every face drops one neighbour per boundary row
NEL_C32 = 117_178_368              # SURVEY.md section 8 [probe]: n_elements for C = 32, sigma = 1


@pytest.fixture(scope="module")
def t():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    if torch.cuda.get_device_properties(0).total_memory < 40e9:
        pytest.skip("needs a large-memory GPU")
    return torch


@pytest.fixture(scope="module")
def lap(eng, t):
    mtx = eng.MtxData.stencil(7, N, N, N)
    scs = eng.convert_to_scs(mtx, 32, 1, "dp")
    eng.permute_scs_cols(scs)
    return mtx, scs


def test_full_size_structure_invariants(eng, lap):
    mtx, scs = lap
    assert (mtx.n_rows, mtx.nnz) == (NR, NNZ)
    assert (scs.n_rows_padded, scs.n_chunks, scs.n_elements, scs.nnz) == (NR, NR // 32, NEL_C32, NNZ)
    g = scs.export()
    cl = g.chunk_lengths.astype(np.int64)
    assert g.chunk_ptrs[0] == 0 and g.chunk_ptrs[-1] == NEL_C32
    assert np.array_equal(np.diff(g.chunk_ptrs.astype(np.int64)), cl * 32)          # chunk_ptrs = running sum of len * C
    assert cl.min() >= 4 and cl.max() == 7
    assert np.array_equal(g.old_to_new, np.arange(NR, dtype=np.int32))               # sigma = 1: identity
    assert int((g.values != 0).sum()) == NNZ and float(g.values.sum()) == 6.0 * NR - (NNZ - NR)
    assert g.col_idxs.min() == 0 and g.col_idxs.max() == NR - 1


def test_full_size_known_answer_row_sums(eng, lap, t):
    """A * 1 = number of missing neighbours per row (diag 6, six -1): 0 inside, 1 on faces, 2 on edges, 3 at corners — exact in fp64."""
    _, scs = lap
    x = t.ones(NR, dtype=t.float64, device="cuda")
    y = t.empty(NR, dtype=t.float64, device="cuda")
    eng.spmv(scs, x, y)
    cnt = t.bincount(y.to(t.int64), minlength=4).cpu().numpy()
    m = N - 2
    assert cnt.tolist() == [m ** 3, 6 * m * m, 12 * m, 8]
    assert float(y.sum()) == 7.0 * NR - NNZ
    # coordinates: y[(z*N + yy)*N + xx] counts the coordinates on the boundary
    idx = t.arange(NR, device="cuda")
    xx, yy, zz = idx % N, (idx // N) % N, idx // (N * N)
    on = lambda c: ((c == 0) | (c == N - 1)).to(t.float64)
    assert t.equal(y, on(xx) + on(yy) + on(zz))


def test_full_size_linearity_and_formats(eng, lap, t):
    mtx, scs = lap
    g = t.Generator(device="cuda").manual_seed(7)
    x1 = t.rand(NR, dtype=t.float64, device="cuda", generator=g) - 0.5
    x2 = t.rand(NR, dtype=t.float64, device="cuda", generator=g) - 0.5
    y1, y2, y12 = (t.empty(NR, dtype=t.float64, device="cuda") for _ in range(3))
    eng.spmv(scs, x1, y1)
    eng.spmv(scs, x2, y2)
    eng.spmv(scs, 2.0 * x1 - 3.0 * x2, y12)
    scale = 12.0 * (2.0 * 0.5 + 3.0 * 0.5)  # sum |a_ij| * max|x| per row
    assert float((y12 - (2.0 * y1 - 3.0 * y2)).abs().max()) <= 1e-12 * scale
    # same product through sigma = 512 (rows scrambled inside 512-row windows, y written back in user order): the per-row
    # summation order is the storage order in both, so the results are bit-identical
    s512 = eng.convert_to_scs(mtx, 32, 512, "dp")
    yu = t.empty(NR, dtype=t.float64, device="cuda")
    eng.spmv_unpermuted(s512, x1, yu)
    assert t.equal(yu, y1)
    del s512
    # CRS (C = 1, sigma = 1) through the streamed CRS kernel: sequential per row -> bit-identical as well
    crs = eng.convert_to_scs(mtx, 1, 1, "dp")
    eng.permute_scs_cols(crs)
    yc = t.empty(NR, dtype=t.float64, device="cuda")
    eng.spmv(crs, x1, yc)
    assert t.equal(yc, y1)
    del crs
    # C = 128 / 64 through the wide-chunk streamed kernel, C = 16 / 8 through the narrow-chunk one (same per-row FMA order)
    for Cw in (128, 64, 16, 8):
        sw = eng.convert_to_scs(mtx, Cw, 1, "dp")
        eng.permute_scs_cols(sw)
        yc.zero_()
        eng.spmv(sw, x1, yc)
        assert t.equal(yc, y1), Cw
        del sw
    # SpMMV: every column of the block product equals the SpMV of that column (same FMA order), both layouts
    for layout in ("rowwise", "colwise"):
        bvs = 4
        X = t.empty(NR * bvs, dtype=t.float64, device="cuda")
        cols = [x1, x2, x1 + x2, x1 - x2]
        for v, c in enumerate(cols):
            if layout == "rowwise":
                X.view(NR, bvs)[:, v] = c
            else:
                X[v * NR:(v + 1) * NR] = c
        Y = t.empty_like(X)
        eng.spmmv(scs, X, Y, bvs, NR, layout)
        for v, c in enumerate(cols):
            eng.spmv(scs, c.contiguous(), yc)
            got = Y.view(NR, bvs)[:, v] if layout == "rowwise" else Y[v * NR:(v + 1) * NR]
            assert t.equal(got, yc), (layout, v)


@pytest.mark.parametrize("vt,tol", [("sp", 1e-5), ("hp", 1e-2)])
def test_full_size_lower_precisions(eng, lap, t, vt, tol):
    mtx, scs = lap
    dt = {"sp": t.float32, "hp": t.float16}[vt]
    g = t.Generator(device="cuda").manual_seed(3)
    x = (t.rand(NR, dtype=t.float64, device="cuda", generator=g) + 0.5).to(dt)      # exactly representable in the narrow type
    y64 = t.empty(NR, dtype=t.float64, device="cuda")
    eng.spmv(scs, x.to(t.float64), y64)
    s = eng.convert_to_scs(mtx, 32, 1, vt)
    eng.permute_scs_cols(s)
    y = t.empty(NR, dtype=dt, device="cuda")
    eng.spmv(s, x, y)
    scale = 12.0 * 1.5
    assert float((y.to(t.float64) - y64).abs().max()) <= tol * scale
