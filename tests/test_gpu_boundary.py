"""GPU: the rest of the drop-in boundary — reference-signature dispatch, raw-array entry points, slab extraction, vector helpers.

  * uspmv_coo_seg_mtx                  == the reference's seg_mtx_struct + localize_row_idx (fixtures from oracle/_ref)
  * uspmv_scs_from_arrays              a ScsData built by the reference / the oracle runs unchanged on the GPU
  * uspmv_block_spmv_gpu, uspmv_scs_ap_gpu   raw device arrays per call, like the reference's kernels
  * uspmv_apply_strided_permutation, uspmv_apply_permutation_block, uspmv_generate_inv_perm
  * include/uspmv_harness_adapter.hpp  run THROUGH the reference's std::function typedefs (oracle/_ref/adapter_check_*)"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, MATRIX_NAMES, ROOT, load_matrix

pytestmark = pytest.mark.gpu
vp = C.c_void_p


def _t():
    import torch
    return torch


def dev(a):
    return _t().from_numpy(np.ascontiguousarray(a)).cuda()


def same_bits(a, b):
    """Bit-identical, except that any NaN equals any NaN (fp16 overflow of a matrix with huge entries: inf - inf; the payload of a
    NaN is not part of the contract)."""
    a, b = np.asarray(a), np.asarray(b)
    nan = np.isnan(a)
    if not np.array_equal(nan, np.isnan(b)):
        return False
    w = {2: np.uint16, 4: np.uint32, 8: np.uint64}[a.dtype.itemsize]
    return np.array_equal(a.view(w)[~nan], b.view(w)[~nan])


def test_seg_mtx_equals_reference_fixture(eng, pkg):
    z = np.load(os.path.join(GOLDEN, "ref_seg_mtx.npz"))
    keys = [k for k in z.files if k.endswith("|wsa")]
    assert len(keys) >= 20
    for k in keys:
        name, P, seg, _ = k.split("|")
        P = int(P)
        wsa = np.ascontiguousarray(z[k], np.int32)
        n, _, I, J, V = load_matrix(name)
        total = eng.MtxData.from_host(n, n, I, J, V)
        for r in range(P):
            h, nd = vp(), C.c_long(0)
            pkg.capi.call("uspmv_coo_seg_mtx", total.h, wsa.ctypes.data_as(vp), r, P, C.byref(h), C.byref(nd))
            loc = eng.MtxData(h, total.ctx)
            gI, gJ, gV = loc.to_host()
            assert np.array_equal(gI, z[f"{name}|{P}|{seg}|{r}|I"]) and np.array_equal(gJ, z[f"{name}|{P}|{seg}|{r}|J"]), (k, r)
            assert np.array_equal(gV, z[f"{name}|{P}|{seg}|{r}|V"]), (k, r)
            assert loc.n_rows == int(wsa[r + 1] - wsa[r]) and loc.n_cols == n
            assert nd.value == len(np.unique(gI))  # what the reference stores as the local n_rows (mpi_funcs.hpp:770)


def test_seg_mtx_rejects_unsorted_and_bad_ranges(eng, pkg):
    n, _, I, J, V = load_matrix("impcol_e")
    total = eng.MtxData.from_host(n, n, I[::-1].copy(), J[::-1].copy(), V[::-1].copy())
    wsa = np.array([0, n // 2, n], np.int32)
    h = vp()
    with pytest.raises(pkg.capi.UspmvError):
        pkg.capi.call("uspmv_coo_seg_mtx", total.h, wsa.ctypes.data_as(vp), 0, 2, C.byref(h), None)
    total = eng.MtxData.from_host(n, n, I, J, V)
    bad = np.array([0, n + 5, n], np.int32)
    with pytest.raises(pkg.capi.UspmvError):
        pkg.capi.call("uspmv_coo_seg_mtx", total.h, bad.ctypes.data_as(vp), 0, 2, C.byref(h), None)


@pytest.mark.parametrize("name", ["bcsstk13", "impcol_e", "FDM-2d-16"])
@pytest.mark.parametrize("vt", ["dp", "sp", "hp"])
def test_adopted_reference_built_matrix_runs_unchanged(eng, orc, pkg, name, vt):
    """uspmv_scs_from_arrays: arrays built on the host (here by the oracle, which is pinned to the reference's convert_to_scs) become a
    device handle; y is bit-identical to the oracle's kernel on the same arrays — for SpMV and SpMMV."""
    t = _t()
    n, _, I, J, V = load_matrix(name)
    for Cc, sigma in ((32, 64), (1, 1), (16, 16)):
        ref = orc.convert_to_scs(n, n, I, J, V, Cc, sigma, vt)
        orc.permute_scs_cols(ref, ref.old_to_new)
        h = vp()
        pkg.capi.call("uspmv_scs_from_arrays", eng.default_context().h, eng.VT_CODE[vt], Cc, sigma, n, n, ref.n_chunks, ref.chunk_ptrs.ctypes.data_as(vp),
                      ref.chunk_lengths.ctypes.data_as(vp), ref.col_idxs.ctypes.data_as(vp), ref.values.ctypes.data_as(vp),
                      ref.old_to_new.ctypes.data_as(vp), 0, C.byref(h))
        scs = eng.ScsData(h, eng.default_context())
        scs.vt = eng.VT_CODE[vt]
        assert (scs.n_chunks, scs.n_elements, scs.n_rows_padded) == (ref.n_chunks, ref.n_elements, ref.n_rows_padded)
        g = scs.export()
        assert np.array_equal(g.col_idxs, ref.col_idxs) and np.array_equal(g.old_to_new, ref.old_to_new)
        npt = ref.values.dtype
        x = np.random.default_rng(3).uniform(-1, 1, ref.n_rows_padded).astype(npt)
        y_ref = orc.spmv_scs(ref, x)
        yd = t.zeros(ref.n_rows_padded, dtype=eng.torch_dtype(scs.vt), device="cuda")
        eng.spmv(scs, dev(x), yd)
        assert same_bits(yd.cpu().numpy(), y_ref), (name, vt, Cc, sigma)


@pytest.mark.parametrize("vt", ["dp", "sp"])
@pytest.mark.parametrize("layout", ["rowwise", "colwise"])
@pytest.mark.parametrize("Cc,sigma", [(32, 64), (1, 1), (8, 8)])
def test_block_spmv_on_raw_device_arrays(eng, orc, pkg, vt, layout, Cc, sigma):
    t = _t()
    n, _, I, J, V = load_matrix("bcsstk13")
    ref = orc.convert_to_scs(n, n, I, J, V, Cc, sigma, vt)
    orc.permute_scs_cols(ref, ref.old_to_new)
    npt = ref.values.dtype
    for bvs in (2, 3, 8):
        ld = ref.n_rows_padded
        X = np.random.default_rng(bvs).uniform(-1, 1, ld * bvs).astype(npt)
        Y_ref = orc.spmmv_scs(ref, X, bvs, ld, 1 if layout == "rowwise" else 0)
        cp, cl, ci, v = dev(ref.chunk_ptrs), dev(ref.chunk_lengths), dev(ref.col_idxs), dev(ref.values)
        Xd = dev(X)
        Yd = t.zeros_like(Xd)
        pkg.capi.call("uspmv_block_spmv_gpu", eng.default_context().h, eng.VT_CODE[vt], Cc, ref.n_chunks, vp(cp.data_ptr()),
                      vp(cl.data_ptr()) if Cc > 1 else None, vp(ci.data_ptr()), vp(v.data_ptr()), vp(Xd.data_ptr()), vp(Yd.data_ptr()), bvs, ld,
                      eng.LAYOUT[layout], None)
        t.cuda.synchronize()
        assert np.array_equal(Yd.cpu().numpy().view(np.uint8), Y_ref.view(np.uint8)), (vt, layout, Cc, bvs)


@pytest.mark.parametrize("mode", ["ap[dp_sp]", "ap[dp_hp]", "ap[sp_hp]", "ap[dp_sp_hp]"])
@pytest.mark.parametrize("Cc,sigma", [(32, 128), (1, 1), (4, 8)])
def test_ap_on_raw_device_arrays_equals_reference(eng, orc, refs, pkg, mats, mode, Cc, sigma):
    """uspmv_scs_ap_gpu with the raw arrays of every part (the reference's MultiPrecKernelArgs style): bit-identical to the reference's
    interface.hpp kernels on the same arrays."""
    t = _t()
    used = {"ap[dp_sp]": (0, 1), "ap[dp_hp]": (0, 2), "ap[sp_hp]": (1, 2), "ap[dp_sp_hp]": (0, 1, 2)}[mode]
    vts = ("dp", "sp", "hp")
    n = 4096
    _, _, I, J, V = mats.random_coo(n, 8, seed=Cc, empty_rows=False)
    V = np.sign(V) * 10.0 ** np.random.default_rng(1).uniform(-3, 1, len(V))
    part, _ = orc.partition_precisions(mode, I, J, V, 0.5, 0.01)
    parts, perm = [None] * 3, None
    for k, p in enumerate(used):
        sel = part == p
        parts[p] = orc.convert_to_scs(n, n, I[sel], J[sel], V[sel], Cc, sigma, vts[p], fixed_perm=perm)
        if k == 0:
            perm = parts[p].old_to_new
    x = np.random.default_rng(5).uniform(-1, 1, n)
    y_ref = refs.iface.ap_scs(mode, parts[0], parts[1], parts[2], x, x.astype(np.float32))
    keep, arr = [], (vp * 12)()
    for p in range(3):
        if parts[p] is None:
            continue
        for q, a in enumerate((parts[p].chunk_ptrs, parts[p].chunk_lengths, parts[p].col_idxs, parts[p].values)):
            d = dev(a)
            keep.append(d)
            arr[4 * p + q] = d.data_ptr()
        if Cc == 1:
            arr[4 * p + 1] = None  # CRS: the row-length array is optional (unused by the reference's CRS kernels)
    dt = t.float32 if mode == "ap[sp_hp]" else t.float64
    xd = dev(x).to(dt)
    yd = t.zeros(len(y_ref), dtype=dt, device="cuda")
    pkg.capi.call("uspmv_scs_ap_gpu", eng.default_context().h, eng.AP_MODE[mode], Cc, parts[used[0]].n_chunks, arr, vp(xd.data_ptr()), vp(yd.data_ptr()), None)
    t.cuda.synchronize()
    assert np.array_equal(yd.cpu().numpy().view(np.uint8), y_ref.view(np.uint8)), (mode, Cc)


def test_strided_and_block_permutations(eng, pkg):
    """apply_strided_permutation (utilities.hpp:1784-1799) literally, one vector of a row-major block vector at a time the way the
    harness calls it, against numpy; the block form (all vectors in one launch) must give the same block vector, both layouts."""
    t = _t()
    ctx = eng.default_context()
    rng = np.random.default_rng(0)
    n, bvs = 1000, 4
    perm = rng.permutation(n).astype(np.int32)
    pd = dev(perm)
    for npt in (np.float64, np.float32, np.float16):
        vt = {np.float64: 0, np.float32: 1, np.float16: 2}[npt]
        X = rng.uniform(-1, 1, n * bvs).astype(npt)
        Xd = dev(X)
        out = t.full_like(Xd, 7.0)
        for v in range(bvs):  # out[i*bvs + v] = in[perm[i]*bvs + v]
            es = Xd.element_size()
            pkg.capi.call("uspmv_apply_strided_permutation", ctx.h, vp(out.data_ptr() + v * es), vp(Xd.data_ptr() + v * es), vp(pd.data_ptr()), n, bvs, vt, None)
        t.cuda.synchronize()
        want = X.reshape(n, bvs)[perm].reshape(-1)
        assert np.array_equal(out.cpu().numpy(), want), npt
        out2 = t.zeros_like(Xd)
        eng.apply_permutation_block(out2, Xd, vp(pd.data_ptr()), n, bvs, n, "rowwise")
        assert np.array_equal(out2.cpu().numpy(), want), npt
        # column-major: out[i + v*ld] = in[perm[i] + v*ld], ld > n (padding between the vectors is left alone)
        ld = n + 24
        Xc = rng.uniform(-1, 1, ld * bvs).astype(npt)
        Xcd = dev(Xc)
        out3 = t.full_like(Xcd, 9.0)
        eng.apply_permutation_block(out3, Xcd, vp(pd.data_ptr()), n, bvs, ld, "colwise")
        w3 = np.full(ld * bvs, 9.0, npt)
        for v in range(bvs):
            w3[v * ld: v * ld + n] = Xc[v * ld: v * ld + n][perm]
        assert np.array_equal(out3.cpu().numpy(), w3), npt
        # a stride-1 strided permutation is apply_permutation
        o4 = t.zeros(n, dtype=Xd.dtype, device="cuda")
        pkg.capi.call("uspmv_apply_strided_permutation", ctx.h, vp(o4.data_ptr()), vp(Xd.data_ptr()), vp(pd.data_ptr()), n, 1, vt, None)
        o5 = t.zeros(n, dtype=Xd.dtype, device="cuda")
        eng.apply_permutation(o5, Xd, vp(pd.data_ptr()), n)
        assert t.equal(o4, o5)


def test_generate_inv_perm(eng, pkg):
    t = _t()
    ctx = eng.default_context()
    n = 5000
    perm = np.random.default_rng(2).permutation(n).astype(np.int32)
    pd = dev(perm)
    inv = t.full((n,), -5, dtype=t.int32, device="cuda")
    pkg.capi.call("uspmv_generate_inv_perm", ctx.h, vp(pd.data_ptr()), vp(inv.data_ptr()), n, n, None)
    want = np.empty(n, np.int32)
    want[perm] = np.arange(n, dtype=np.int32)   # inv_perm[perm[i]] = i (utilities.hpp:1755-1766)
    assert np.array_equal(inv.cpu().numpy(), want)
    # the structure's own pair: new_to_old is the inverse of old_to_new on the real rows
    nn, _, I, J, V = load_matrix("bcsstk13")
    scs = eng.convert_to_scs(eng.MtxData.from_host(nn, nn, I, J, V), 32, 512, "dp")
    g = scs.export()
    o2n = dev(g.old_to_new)
    inv2 = t.full((scs.n_rows_padded,), -1, dtype=t.int32, device="cuda")
    pkg.capi.call("uspmv_generate_inv_perm", ctx.h, vp(o2n.data_ptr()), vp(inv2.data_ptr()), nn, scs.n_rows_padded, None)
    assert np.array_equal(inv2.cpu().numpy(), g.new_to_old)
    bad = dev(np.array([0, 1, n + 3], np.int32))
    with pytest.raises(pkg.capi.UspmvError):
        pkg.capi.call("uspmv_generate_inv_perm", ctx.h, vp(bad.data_ptr()), vp(inv.data_ptr()), 3, n, None)


@pytest.mark.parametrize("layout", ["col", "row"])
def test_harness_adapter_through_the_reference_typedefs(eng, layout):
    """oracle/_ref/adapter_check_*: the launchers of include/uspmv_harness_adapter.hpp called through SpmvKernel::OnePrecFuncPtr /
    MultiPrecFuncPtr objects of the reference's own classes_structs.hpp, on matrices built by the reference's own convert_to_scs, against
    the reference's own host kernels (host arrays, device arrays + device scalars, CRS, block vectors, ap[dp_sp])."""
    exe = os.path.join(ROOT, "oracle", "_ref", f"adapter_check_{layout}")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/adapter_check not built (needs /root/reference)")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count(" OK ") == 5 and "FAILED" not in r.stdout, r.stdout


def test_options_are_per_context(eng, pkg, orc, mats):
    """Two contexts in one process can run different kernel variants (uspmv_ctx_set_option); uspmv_set_option still reaches every live
    context; the results are identical either way (the variants are bit-identical)."""
    t = _t()
    capi = pkg.capi
    a, b = eng.Context(0), eng.Context(0)

    def get(ctx, name):
        v = C.c_long(-99)
        capi.call("uspmv_ctx_get_option", ctx.h if ctx else None, name.encode(), C.byref(v))
        return v.value
    assert get(a, "scs_stream") == 1 and get(b, "scs_stream") == 1
    capi.call("uspmv_ctx_set_option", a.h, b"scs_stream", 0)          # context a: the direct-load kernel
    capi.call("uspmv_ctx_set_option", b.h, b"stream_variant", 6)      # context b: another instantiation of the streamed kernel
    assert (get(a, "scs_stream"), get(b, "scs_stream"), get(None, "scs_stream")) == (0, 1, 1)
    assert (get(a, "stream_variant"), get(b, "stream_variant")) == (0, 6)
    coo = mats.random_coo(5000, 7, seed=2)
    ref = orc.convert_to_scs(*coo, 32, 64, "dp")
    orc.permute_scs_cols(ref, ref.old_to_new)
    x = np.random.default_rng(1).uniform(-1, 1, ref.n_rows_padded)
    y_ref = orc.spmv_scs(ref, x)
    launches = []
    for ctx in (a, b):
        scs = eng.convert_to_scs(eng.MtxData.from_host(*coo, ctx=ctx), 32, 64, "dp")
        eng.permute_scs_cols(scs)
        yd = t.zeros(ref.n_rows_padded, dtype=t.float64, device="cuda")
        eng.spmv(scs, dev(x), yd)
        t.cuda.synchronize()
        assert np.array_equal(yd.cpu().numpy(), y_ref)
        launches.append(scs)
    capi.set_option("stream_variant", 0)                               # process-wide: defaults and every live context
    assert (get(a, "stream_variant"), get(b, "stream_variant"), get(None, "stream_variant")) == (0, 0, 0)
    assert get(a, "scs_stream") == 0                                   # a context-level choice survives an unrelated global change
    capi.call("uspmv_ctx_set_option", a.h, b"scs_stream", 1)
