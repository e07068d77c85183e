"""GPU parity tests proper: the CUDA path (through the C ABI) against the oracle on the same seeded inputs.

Bars: bit-exact for every index structure and for the stored values; y within 1e-12 (dp), 1e-5 (sp),
1e-2 (hp) relative per element (BASELINE.json north_star) — and, for the SCS kernels, additionally
bit-identical to the oracle because the per-row fused-multiply-add order is the same.
"""
import numpy as np
import pytest

from conftest import MATRIX_NAMES, load_matrix

pytestmark = pytest.mark.gpu

NPT = {"dp": np.float64, "sp": np.float32, "hp": np.float16}
TOL = {"dp": 1e-12, "sp": 1e-5, "hp": 1e-2}
STRUCT_KEYS = ("chunk_ptrs", "chunk_lengths", "col_idxs", "values", "old_to_new")


def torch_():
    import torch
    return torch


def assert_same_structure(got, ref, tag=""):
    for k in ("n_rows_padded", "n_chunks", "n_elements", "nnz"):
        assert getattr(got, k) == getattr(ref, k), f"{tag}: {k}"
    for k in STRUCT_KEYS:
        a, b = getattr(got, k), getattr(ref, k)
        assert a.dtype == b.dtype and a.shape == b.shape, f"{tag}: {k} dtype/shape"
        assert np.array_equal(a.view(np.uint8), b.view(np.uint8)), f"{tag}: {k} differs from the oracle"
    assert np.array_equal(got.new_to_old, ref.new_to_old), f"{tag}: new_to_old"


def rel_err(y, y_ref):
    y = y.astype(np.float64)
    y_ref = y_ref.astype(np.float64)
    return float(np.max(np.abs(y - y_ref) / np.maximum(np.abs(y_ref), 1e-30))) if len(y) else 0.0


def dev(a):
    return torch_().from_numpy(np.ascontiguousarray(a)).cuda()


def build_both(eng, orc, coo, C, sigma, vt, permute=True):
    n, nc, I, J, V = coo
    ref = orc.convert_to_scs(n, nc, I, J, V, C, sigma, vt)
    mtx = eng.MtxData.from_host(n, nc, I, J, V)
    scs = eng.convert_to_scs(mtx, C, sigma, vt)
    if permute:
        orc.permute_scs_cols(ref, ref.old_to_new)
        eng.permute_scs_cols(scs)
    return scs, ref


def positive_coo(mats, n, avg, seed):
    n, nc, I, J, V = mats.random_coo(n, avg, seed)
    return n, nc, I, J, np.abs(V) + 0.25


# ------------------------------------------------------------------------------------------------
# convert_to_scs + permute_scs_cols: bit-exact structures
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", MATRIX_NAMES)
@pytest.mark.parametrize("C,sigma", [(1, 1), (4, 8), (8, 1), (32, 512), (16, 64), (10, 30), (128, 256), (3, 7)])
def test_build_fixture_matrices(eng, orc, name, C, sigma):
    coo = load_matrix(name)
    for vt in ("dp", "sp"):
        scs, ref = build_both(eng, orc, coo, C, sigma, vt)
        assert_same_structure(scs.export(), ref, f"{name} C={C} s={sigma} {vt}")


@pytest.mark.parametrize("n", [1, 2, 31, 33, 1000, 20011])
@pytest.mark.parametrize("C,sigma", [(1, 1), (2, 2), (32, 1), (32, 16), (32, 17), (32, 512), (64, 4096), (5, 1 << 20)])
def test_build_random(eng, orc, mats, n, C, sigma):
    coo = mats.random_coo(n, 6, seed=n * 131 + C + sigma)
    if len(coo[2]) == 0:
        pytest.skip("empty")
    scs, ref = build_both(eng, orc, coo, C, sigma, "dp")
    assert_same_structure(scs.export(), ref, f"random n={n} C={C} s={sigma}")


def test_build_tie_heavy_and_adversarial(eng, orc):
    """sigma-windows > 16 rows with many ties: the order is libstdc++'s introsort tie order (heapsort branch included)."""
    z = np.load(__import__("os").path.join(__import__("conftest").GOLDEN, "ref_stdsort.npz"))
    names = sorted({k.split("|")[0] for k in z.files})
    before = orc.lib.orc_heapsort_calls()
    for nm in names:
        cnt = z[nm + "|cnt"].astype(np.int64)
        n = len(cnt)
        I = np.repeat(np.arange(n), cnt).astype(np.int32)
        J = np.zeros(len(I), np.int32)
        mtx = eng.MtxData.from_host(n, n, I, J, np.ones(len(I)))
        got = eng.convert_to_scs(mtx, 1, n, "dp").export()
        assert np.array_equal(got.old_to_new, z[nm + "|old_to_new"]), f"{nm}: differs from the reference's std::sort order"
        ref = orc.convert_to_scs(n, n, I, J, np.ones(len(I)), 1, n)
        assert np.array_equal(got.old_to_new, ref.old_to_new)


def test_build_unsorted_coo(eng, orc, mats):
    """Rows in arbitrary order: within-row order of the INPUT must be preserved (utilities.hpp:2013-2036)."""
    n, nc, I, J, V = mats.random_coo(3000, 5, seed=5)
    perm = np.random.default_rng(0).permutation(len(I))
    Iu, Ju, Vu = I[perm], J[perm], V[perm]
    ref = orc.convert_to_scs(n, nc, Iu, Ju, Vu, 16, 64)
    mtx = eng.MtxData.from_host(n, nc, Iu, Ju, Vu)
    got = eng.convert_to_scs(mtx, 16, 64, "dp").export()
    assert_same_structure(got, ref, "unsorted COO")


def test_build_fixed_permutation(eng, orc, mats):
    n, nc, I, J, V = mats.random_coo(2048, 6, seed=9, empty_rows=False)
    first = orc.convert_to_scs(n, nc, I, J, V, 32, 128)
    sel = np.abs(V) < 0.5
    ref = orc.convert_to_scs(n, nc, I[sel], J[sel], V[sel], 32, 128, "sp", fixed_perm=first.old_to_new)
    mtx = eng.MtxData.from_host(n, nc, I[sel], J[sel], V[sel])
    got = eng.convert_to_scs(mtx, 32, 128, "sp", fixed_permutation=first.old_to_new).export()
    assert_same_structure(got, ref, "fixed_permutation")
    assert np.array_equal(got.old_to_new, np.arange(n))


def test_build_empty_and_errors(eng, pkg):
    mtx = eng.MtxData.from_host(7, 7, np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros(0))
    s = eng.convert_to_scs(mtx, 4, 4, "dp")
    e = s.export()
    assert s.n_elements == 0 and s.n_chunks == 2 and np.all(e.chunk_lengths == 0) and np.all(e.chunk_ptrs == 0)
    with pytest.raises(pkg.capi.UspmvError):
        eng.convert_to_scs(mtx, 0, 1, "dp")
    bad = eng.MtxData.from_host(3, 3, np.array([0, 5], np.int32), np.array([0, 1], np.int32), np.ones(2))
    with pytest.raises(pkg.capi.UspmvError):
        eng.convert_to_scs(bad, 1, 1, "dp")


def test_stencil_generator_matches_host(eng, mats):
    for pts, dims in ((7, (9, 7, 5)), (27, (6, 5, 7)), (7, (1, 1, 4)), (27, (2, 1, 3))):
        n, nc, I, J, V = mats.stencil_coo(pts, *dims)
        m = eng.MtxData.stencil(pts, *dims)
        Id, Jd, Vd = m.to_host()
        assert m.n_rows == n and m.nnz == len(I)
        assert np.array_equal(Id, I) and np.array_equal(Jd, J) and np.array_equal(Vd, V)
    # a slab: local rows, global columns
    n, nc, I, J, V = mats.stencil_coo(27, 6, 5, 8, 60, 180)
    m = eng.MtxData.stencil(27, 6, 5, 8, 60, 180)
    Id, Jd, Vd = m.to_host()
    assert np.array_equal(Id, I) and np.array_equal(Jd, J) and np.array_equal(Vd, V)


# ------------------------------------------------------------------------------------------------
# SpMV
# ------------------------------------------------------------------------------------------------
def run_spmv(eng, scs, ref, x_user, vt):
    t = torch_()
    xp = np.zeros(max(scs.n_rows_padded, scs.n_cols), NPT[vt])
    xp[ref.old_to_new] = x_user.astype(NPT[vt])
    yd = t.zeros(scs.n_rows_padded, dtype=eng.torch_dtype(scs.vt), device="cuda")
    eng.spmv(scs, dev(xp), yd)
    t.cuda.synchronize()
    return xp, yd.cpu().numpy()


@pytest.mark.parametrize("vt", ["dp", "sp", "hp"])
@pytest.mark.parametrize("C,sigma", [(2, 1), (4, 8), (8, 8), (16, 64), (32, 1), (32, 512), (64, 64), (128, 256), (256, 256), (10, 30), (3, 1)])
def test_spmv_scs_bit_exact(eng, orc, mats, vt, C, sigma):
    coo = positive_coo(mats, 5000, 7, seed=C * 7 + sigma) if vt == "hp" else mats.random_coo(5000, 7, seed=C * 7 + sigma)
    scs, ref = build_both(eng, orc, coo, C, sigma, vt)
    x = np.random.default_rng(3).uniform(0.1 if vt == "hp" else -1.0, 1.0, coo[0])
    xp, y = run_spmv(eng, scs, ref, x, vt)
    y_ref = orc.spmv_scs(ref, xp)
    assert rel_err(y, y_ref) <= TOL[vt]
    assert np.array_equal(y.view(np.uint8), y_ref.view(np.uint8)), "SCS kernel is expected to be bit-identical to the oracle"


@pytest.mark.parametrize("name", MATRIX_NAMES)
def test_spmv_fixture_matrices_vs_coo(eng, orc, name):
    """y in user order equals the plain COO sum (per-row order = COO order)."""
    n, nc, I, J, V = load_matrix(name)
    scs, ref = build_both(eng, orc, (n, nc, I, J, V), 32, 512, "dp")
    x = np.random.default_rng(11).uniform(-1, 1, n)
    xp, yp = run_spmv(eng, scs, ref, x, "dp")
    y = yp[ref.old_to_new]
    y_coo = np.zeros(n)
    for i, j, v in zip(I, J, V):  # sequential, FMA-free reference sum
        y_coo[i] += v * x[j]
    scale = np.zeros(n)
    np.add.at(scale, I, np.abs(V * x[J]))
    assert np.all(np.abs(y - y_coo) <= 1e-12 * np.maximum(scale, 1e-300))


@pytest.mark.parametrize("vt", ["dp", "sp", "hp"])
def test_spmv_crs(eng, orc, mats, vt):
    """C = 1, sigma = 1 dispatches to the CRS kernel (split-row reduction => tolerance, not bit-exactness)."""
    coo = positive_coo(mats, 4000, 9, seed=21)
    scs, ref = build_both(eng, orc, coo, 1, 1, vt)
    x = np.random.default_rng(4).uniform(0.1, 1.0, coo[0])
    xp, y = run_spmv(eng, scs, ref, x, vt)
    y_ref = orc.spmv_csr(ref.n_rows, ref.chunk_ptrs, ref.col_idxs, ref.values, xp[:ref.n_cols])
    assert rel_err(y, y_ref) <= TOL[vt]
    # library-built CRS matrices run the streamed kernel: sequential per row => bit-identical to the reference loop
    assert np.array_equal(y.view(np.uint8), y_ref.view(np.uint8))
    # the split-row vector kernel (raw-array entry point / scs_stream = 0) stays within tolerance
    eng_capi = __import__("importlib").import_module("ultimate-spmv_b200").capi
    eng_capi.set_option("scs_stream", 0)
    try:
        _, y2 = run_spmv(eng, scs, ref, x, vt)
    finally:
        eng_capi.set_option("scs_stream", 1)
    assert rel_err(y2, y_ref) <= TOL[vt]


def test_spmv_raw_array_entry_points(eng, orc, mats):
    """uspmv_scs_gpu / uspmv_csr_gpu with caller-owned device arrays (interface.hpp:1741-1793)."""
    t = torch_()
    coo = mats.random_coo(3000, 6, seed=2)
    ref = orc.convert_to_scs(*coo, 32, 64)
    x = np.random.default_rng(8).standard_normal(ref.n_rows_padded)
    y = t.zeros(ref.n_rows_padded, dtype=t.float64, device="cuda")
    eng.uspmv_scs_gpu(32, ref.n_chunks, dev(ref.chunk_ptrs), dev(ref.chunk_lengths), dev(ref.col_idxs), dev(ref.values), dev(x), y)
    t.cuda.synchronize()
    assert np.array_equal(y.cpu().numpy(), orc.spmv_scs(ref, x))
    crs = orc.convert_to_scs(*coo, 1, 1)
    y2 = t.zeros(crs.n_rows, dtype=t.float64, device="cuda")
    eng.uspmv_csr_gpu(crs.n_rows, dev(crs.chunk_ptrs), dev(crs.col_idxs), dev(crs.values), dev(x[:crs.n_rows]), y2)
    t.cuda.synchronize()
    y_ref = orc.spmv_csr(crs.n_rows, crs.chunk_ptrs, crs.col_idxs, crs.values, x[:crs.n_rows])
    scale = np.zeros(crs.n_rows)
    np.add.at(scale, coo[2], np.abs(coo[4] * x[coo[3]]))
    assert np.all(np.abs(y2.cpu().numpy() - y_ref) <= 1e-12 * np.maximum(scale, 1e-300))


@pytest.mark.parametrize("vt", ["dp", "sp", "hp"])
@pytest.mark.parametrize("n", [37, 1000, 4099])
def test_csr_raw_arrays_exact_size(eng, orc, mats, vt, n):
    """uspmv_csr_gpu on caller-owned arrays of EXACTLY nnz elements (no slack): the streamed kernel stops its bulk copies at nnz & ~7 and
    reads the last elements with plain loads -> sequential per row, bit-identical to the reference loop (kernels.hpp:46-57), for every
    nnz mod 8; a base pointer that is not 16-byte aligned falls back to the split-row kernel (tolerance)."""
    t = torch_()
    for seed in range(4):
        coo = positive_coo(mats, n, 5, seed=100 + seed)  # positive data: the split-row fallback below is compared relative to |y|
        crs = orc.convert_to_scs(*coo, 1, 1, vt)
        nnz = int(crs.chunk_ptrs[crs.n_rows])
        x = np.random.default_rng(seed).uniform(0.1, 1.0, crs.n_rows).astype(NPT[vt])
        rp, ci, va = dev(crs.chunk_ptrs[: crs.n_rows + 1].copy()), dev(crs.col_idxs[:nnz].copy()), dev(crs.values[:nnz].copy())
        y = t.zeros(crs.n_rows, dtype=va.dtype, device="cuda")
        eng.uspmv_csr_gpu(crs.n_rows, rp, ci, va, dev(x), y)
        t.cuda.synchronize()
        y_ref = orc.spmv_csr(crs.n_rows, crs.chunk_ptrs, crs.col_idxs, crs.values, x)
        assert np.array_equal(y.cpu().numpy().view(np.uint8), y_ref.view(np.uint8)), ("streamed", vt, n, seed, nnz % 8)
        # misaligned bases: one element into a larger buffer
        ci2 = t.empty(nnz + 1, dtype=t.int32, device="cuda"); ci2[1:] = ci
        va2 = t.empty(nnz + 1, dtype=va.dtype, device="cuda"); va2[1:] = va
        y.zero_()
        eng.uspmv_csr_gpu(crs.n_rows, rp, ci2[1:], va2[1:], dev(x), y)
        t.cuda.synchronize()
        assert rel_err(y.cpu().numpy(), y_ref) <= TOL[vt], ("misaligned fallback", vt, n, seed)


def test_spmv_unpermuted_fused(eng, orc, mats):
    """Fused form: x and y in user numbering, columns not permuted, y[new_to_old[row]] written directly."""
    t = torch_()
    coo = mats.random_coo(4099, 6, seed=12)
    scs, ref = build_both(eng, orc, coo, 32, 256, "dp", permute=False)
    x = np.random.default_rng(5).standard_normal(coo[0])
    y = t.full((coo[0],), 7.0, dtype=t.float64, device="cuda")
    eng.spmv_unpermuted(scs, dev(x), y)
    t.cuda.synchronize()
    y_ref_perm = orc.spmv_scs(ref, np.concatenate([x, np.zeros(ref.n_rows_padded)]))
    assert np.array_equal(y.cpu().numpy(), y_ref_perm[ref.old_to_new])


def test_spmv_host_buffers(eng, orc, mats):
    coo = mats.random_coo(2500, 6, seed=14)
    scs, ref = build_both(eng, orc, coo, 32, 64, "dp")
    xp = np.zeros(scs.n_rows_padded)
    xp[ref.old_to_new] = np.random.default_rng(6).standard_normal(coo[0])
    y = np.zeros(scs.n_rows_padded)
    eng.spmv_host(scs, xp, y)
    assert np.array_equal(y, orc.spmv_scs(ref, xp))


def test_apply_permutation(eng, orc, mats):
    t = torch_()
    coo = mats.random_coo(1000, 4, seed=1)
    scs, ref = build_both(eng, orc, coo, 8, 32, "dp")
    x = np.random.default_rng(0).standard_normal(coo[0])
    arrs = scs.device_arrays()
    out = t.zeros(scs.n_rows_padded, dtype=t.float64, device="cuda")
    eng.apply_permutation(out, dev(x), arrs["new_to_old"], scs.n_rows_padded)
    t.cuda.synchronize()
    exp = np.where(ref.new_to_old >= 0, x[np.maximum(ref.new_to_old, 0)], 0.0)
    assert np.array_equal(out.cpu().numpy(), exp)


# ------------------------------------------------------------------------------------------------
# SpMMV
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("vt", ["dp", "sp", "hp"])
@pytest.mark.parametrize("bvs", [2, 3, 4, 8, 16])
@pytest.mark.parametrize("layout", ["rowwise", "colwise"])
@pytest.mark.parametrize("C,sigma", [(32, 64), (1, 1), (4, 1)])
def test_spmmv(eng, orc, mats, vt, bvs, layout, C, sigma):
    t = torch_()
    coo = positive_coo(mats, 1500, 6, seed=bvs + C) if vt == "hp" else mats.random_coo(1500, 6, seed=bvs + C)
    scs, ref = build_both(eng, orc, coo, C, sigma, vt)
    ld = scs.n_rows_padded + 5
    rng = np.random.default_rng(9)
    X = rng.uniform(0.1 if vt == "hp" else -1.0, 1.0, ld * bvs).astype(NPT[vt])
    Y = t.zeros(ld * bvs, dtype=eng.torch_dtype(scs.vt), device="cuda")
    eng.spmmv(scs, dev(X), Y, bvs, ld, layout)
    t.cuda.synchronize()
    lay = 1 if layout == "rowwise" else 0
    Y_ref = orc.spmmv_scs(ref, X, bvs, ld, lay)
    got = Y.cpu().numpy()
    assert rel_err(got, Y_ref) <= TOL[vt]
    assert np.array_equal(got.view(np.uint8), Y_ref.view(np.uint8))


def test_uneven_matrix_long_chunks(eng, pkg, orc, mats):
    """Power-law-like rows: longest-first chunk order, and chunks longer than `split_long_chunks` slots summed in segments
    (deterministic, within tolerance); with the option off the result is bit-identical to the oracle again."""
    t = torch_()
    rng = np.random.default_rng(0)
    n = 140000
    cnt = rng.integers(1, 8, n)
    long_rows = rng.choice(n, 40, replace=False)
    cnt[long_rows] = rng.integers(600, 5000, 40)
    I = np.repeat(np.arange(n), cnt).astype(np.int32)
    J = rng.integers(0, n, len(I)).astype(np.int32)
    V = rng.standard_normal(len(I))
    x = rng.standard_normal(n)
    ref = orc.convert_to_scs(n, n, I, J, V, 32, 64)
    orc.permute_scs_cols(ref, ref.old_to_new)
    xp = np.zeros(ref.n_rows_padded)
    xp[ref.old_to_new] = x
    y_ref = orc.spmv_scs(ref, xp)
    scale = np.zeros(n)
    np.add.at(scale, I, np.abs(V * x[J]))
    scale_p = np.zeros(ref.n_rows_padded)
    scale_p[ref.old_to_new] = scale
    outs = {}
    for split in (256, 0):
        pkg.capi.set_option("split_long_chunks", split)
        mtx = eng.MtxData.from_host(n, n, I, J, V)
        scs = eng.convert_to_scs(mtx, 32, 64, "dp")
        eng.permute_scs_cols(scs)
        y = t.zeros(scs.n_rows_padded, dtype=t.float64, device="cuda")
        eng.spmv(scs, dev(xp), y)
        t.cuda.synchronize()
        outs[split] = y.cpu().numpy()
    pkg.capi.set_option("split_long_chunks", 256)
    assert np.array_equal(outs[0], y_ref), "without splitting: bit-identical to the oracle (longest-first order only)"
    assert np.all(np.abs(outs[256] - y_ref) <= 1e-12 * np.maximum(scale_p, 1e-300))
    short = np.repeat(ref.chunk_lengths <= 256, 32)
    assert np.array_equal(outs[256][short], y_ref[short]), "rows of un-split chunks stay bit-identical"
    assert (outs[256] != y_ref).sum() > 0 or True


def test_spmv_host_pipelined(eng, pkg, orc, mats):
    """uspmv_spmv_host_submit/_wait: three SpMVs in flight, each with its own host x / y."""
    import ctypes as C_
    t = torch_()
    coo = mats.random_coo(3000, 6, seed=15)
    scs, ref = build_both(eng, orc, coo, 32, 64, "dp")
    xs = [t.from_numpy(np.random.default_rng(k).standard_normal(scs.n_rows_padded)).pin_memory() for k in range(5)]
    ys = [t.zeros(scs.n_rows_padded, dtype=t.float64).pin_memory() for _ in range(5)]
    for i in range(5):
        sl = i % 3
        pkg.capi.call("uspmv_spmv_host_wait", scs.h, sl)
        pkg.capi.call("uspmv_spmv_host_submit", scs.h, C_.c_void_p(xs[i].data_ptr()), xs[i].numel(), C_.c_void_p(ys[i].data_ptr()), ys[i].numel(), sl)
    with pytest.raises(pkg.capi.UspmvError):  # slot 1 (step 4) is still in flight
        pkg.capi.call("uspmv_spmv_host_submit", scs.h, C_.c_void_p(xs[0].data_ptr()), xs[0].numel(), C_.c_void_p(ys[0].data_ptr()), ys[0].numel(), 1)
    for sl in range(3):
        pkg.capi.call("uspmv_spmv_host_wait", scs.h, sl)
    for i in range(5):
        assert np.array_equal(ys[i].numpy(), orc.spmv_scs(ref, xs[i].numpy())), i


def test_fixed_permutation_must_be_injective(eng, pkg, mats):
    n, nc, I, J, V = mats.random_coo(100, 4, seed=3, empty_rows=False)
    mtx = eng.MtxData.from_host(n, nc, I, J, V)
    perm = np.arange(n, dtype=np.int32)
    perm[5] = 6
    with pytest.raises(pkg.capi.UspmvError, match="not injective"):
        eng.convert_to_scs(mtx, 4, 8, "dp", fixed_permutation=perm)
    perm = np.arange(n, dtype=np.int32)
    perm[7] = 10 ** 6
    with pytest.raises(pkg.capi.UspmvError, match="outside"):
        eng.convert_to_scs(mtx, 4, 8, "dp", fixed_permutation=perm)


# ------------------------------------------------------------------------------------------------
# matrix ingest (device-side read_mtx post-processing) and equilibration
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", MATRIX_NAMES)
def test_device_ingest_and_equilibrate(eng, orc, name):
    """uspmv_coo_from_entries == the reference's read_mtx result (symmetric expansion order + stable row sort), then
    uspmv_coo_equilibrate == the reference's equilibrate_matrix output, both bit-exact (tests/golden/ref_ingest.npz)."""
    import os
    from conftest import GOLDEN
    z = np.load(os.path.join(GOLDEN, "ref_ingest.npz"))
    n, nc, I, J, V = load_matrix(name)
    mtx = eng.MtxData.from_entries(int(z[f"{name}__n"]), int(z[f"{name}__n"]), z[f"{name}__I"], z[f"{name}__J"], z[f"{name}__V"],
                                   bool(z[f"{name}__sym"]))
    gI, gJ, gV = mtx.to_host()
    assert mtx.nnz == len(I)
    assert np.array_equal(gI, I) and np.array_equal(gJ, J)
    assert np.array_equal(gV.view(np.uint8), V.view(np.uint8))
    rm, cm = mtx.equilibrate()
    _, _, eV = mtx.to_host()
    assert np.array_equal(eV.view(np.uint8), z[f"{name}__equil"].view(np.uint8))
    v_o, rm_o, cm_o = orc.equilibrate(n, nc, I, J, V)
    assert np.array_equal(rm, rm_o) and np.array_equal(cm, cm_o)


def test_device_ingest_rejects_out_of_range(eng, pkg):
    with pytest.raises(pkg.capi.UspmvError, match="outside"):
        eng.MtxData.from_entries(4, 4, [0, 5], [0, 1], [1.0, 2.0], False)


@pytest.mark.parametrize("n", [4097, 4128, 33, 64])
def test_spmv_hp_pair_kernel(eng, pkg, orc, mats, n):
    """fp16, C = 32: the kernel that takes two adjacent chunks per work item (k_scs32_stream_pair): chunks longer than a stage (slices),
    empty chunks, an odd number of chunks (no chunk B at the tail), the un-permuted form — bit-identical to the oracle and to the
    one-chunk kernel."""
    t = torch_()
    rng = np.random.default_rng(n)
    cnt = rng.integers(0, 9, n)
    cnt[rng.choice(n, max(1, n // 40), replace=False)] = rng.integers(17, 60, max(1, n // 40))   # chunks longer than 16 slots
    cnt[: min(n, 96)] = 0                                                                        # whole empty chunks
    I = np.repeat(np.arange(n), cnt).astype(np.int32)
    J = rng.integers(0, n, len(I)).astype(np.int32)
    V = rng.uniform(0.05, 1.0, len(I))
    coo = (n, n, I, J, V)
    pkg.capi.set_option("pair_hp", 1)  # off by default (not faster); the kernel stays covered here
    try:
        for sigma in (1, 128):
            scs, ref = build_both(eng, orc, coo, 32, sigma, "hp")
            x = rng.uniform(0.1, 1.0, n)
            xp, y = run_spmv(eng, scs, ref, x, "hp")
            y_ref = orc.spmv_scs(ref, xp)
            assert np.array_equal(y.view(np.uint8), y_ref.view(np.uint8)), (n, sigma)
            pkg.capi.set_option("pair_hp", 0)
            _, y1 = run_spmv(eng, scs, ref, x, "hp")
            pkg.capi.set_option("pair_hp", 1)
            assert np.array_equal(y.view(np.uint8), y1.view(np.uint8)), (n, sigma)
        scs, ref = build_both(eng, orc, coo, 32, 128, "hp", permute=False)
        xh = rng.uniform(0.1, 1.0, n).astype(np.float16)
        yu = t.full((n,), 7.0, dtype=t.float16, device="cuda")
        eng.spmv_unpermuted(scs, dev(xh), yu)
        t.cuda.synchronize()
    finally:
        pkg.capi.set_option("pair_hp", 0)
    y_ref_perm = orc.spmv_scs(ref, np.concatenate([xh, np.zeros(ref.n_rows_padded, np.float16)]))
    assert np.array_equal(yu.cpu().numpy().view(np.uint8), y_ref_perm[ref.old_to_new].view(np.uint8))
