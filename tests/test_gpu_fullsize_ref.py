"""GPU parity at BASELINE.json's FULL sizes against the UNMODIFIED reference compiled into oracle/_ref (not against properties):

  * config 2: 7-point Laplacian 256^3, C = 32, sigma in {1, 512} (sigma = 512 on a stencil is all ties: the libstdc++ introsort tie order
    is exactly what a device builder can get wrong) — every array of the structure equal to the reference's convert_to_scs +
    permute_scs_cols (utilities.hpp:1842-2104, 1802-1831), y bit-equal to spmv_omp_scs_adv (kernels.hpp:265-301);
  * config 3: SpMMV block_vec_size 4 / 8, row-major, dp / sp — bit-equal to block_spmv_omp_scs_general (kernels.hpp:306-398);
  * config 5: one z-slab of the 27-point 512^3 matrix (the slab of an interior rank) with the reference's own slab order
    convert_to_scs -> collect_local_needed_heri -> permute_scs_cols (main.cpp:1271-1308; mpi_funcs.hpp:242-415): structure, halo
    numbering, need lists and y with a filled halo;
  * config 4: the device power-law generator at 2^22 rows, all four adaptive-precision modes: partition, the three structures and y
    against the oracle and the reference's interface.hpp kernels (interface.hpp:1434-1733).
The reference runs on the host CPU here (seconds per case); the tests need oracle/_ref (prebuilt, travels with the repository)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

N = 256
NR = N ** 3
FIELDS = ("chunk_ptrs", "chunk_lengths", "col_idxs", "old_to_new")


def _t():
    import torch
    return torch


@pytest.fixture(scope="module")
def big_gpu():
    t = _t()
    if not t.cuda.is_available() or t.cuda.get_device_properties(0).total_memory < 40e9:
        pytest.skip("needs a large-memory GPU")
    return t


@pytest.fixture(scope="module")
def lap_host(orc):
    """The 256^3 7-point COO from the oracle's generator (the same order as the device generator, checked below)."""
    return orc.stencil_coo(7, N, N, N)


def _same_struct(got, ref, what):
    for k in FIELDS:
        assert np.array_equal(getattr(got, k), getattr(ref, k)), (what, k)
    assert np.array_equal(got.values.view(np.uint8), ref.values.view(np.uint8)), (what, "values")
    assert (got.n_elements, got.n_chunks, got.n_rows_padded) == (ref.n_elements, ref.n_chunks, ref.n_rows_padded), what


def test_256_device_generator_equals_oracle_generator(eng, big_gpu, lap_host):
    nr, nc, I, J, V = lap_host
    mtx = eng.MtxData.stencil(7, N, N, N)
    gI, gJ, gV = mtx.to_host()
    assert (mtx.n_rows, mtx.nnz) == (nr, len(I))
    assert np.array_equal(gI, I) and np.array_equal(gJ, J) and np.array_equal(gV, V)


@pytest.mark.parametrize("sigma", [1, 512])
def test_256_structure_and_y_equal_the_reference(eng, refs, big_gpu, lap_host, sigma):
    t = big_gpu
    nr, nc, I, J, V = lap_host
    ref = refs.col.convert_to_scs(nr, nc, I, J, V, 32, sigma, "dp", permute_cols=True)
    mtx = eng.MtxData.stencil(7, N, N, N)
    scs = eng.convert_to_scs(mtx, 32, sigma, "dp")
    eng.permute_scs_cols(scs)
    del mtx
    got = scs.export()
    _same_struct(got, ref, f"sigma={sigma}")
    if sigma > 1:
        assert not np.array_equal(ref.old_to_new, np.arange(nr, dtype=np.int32))  # the tie order really is a permutation here
    # y: random x in user order -> permuted like the harness (main.cpp:86-93) -> reference kernel on the host, device kernel
    x = np.random.default_rng(11).uniform(-1.0, 1.0, nr)
    xp = np.zeros(ref.n_rows_padded)
    xp[ref.old_to_new] = x
    y_ref = refs.col.spmv_scs(ref, xp, adv=True)
    xd = t.from_numpy(xp).cuda()
    yd = t.full((scs.n_rows_padded,), float("nan"), dtype=t.float64, device="cuda")
    eng.spmv(scs, xd, yd)
    t.cuda.synchronize()
    assert np.array_equal(yd.cpu().numpy().view(np.uint64), y_ref.view(np.uint64)), f"sigma={sigma}: y differs from spmv_omp_scs_adv"


@pytest.mark.parametrize("vt", ["dp", "sp"])
@pytest.mark.parametrize("bvs", [4, 8])
def test_256_spmmv_rowwise_equals_block_spmv_general(eng, refs, big_gpu, lap_host, vt, bvs):
    t = big_gpu
    nr, nc, I, J, V = lap_host
    ref = refs.row.convert_to_scs(nr, nc, I, J, V, 32, 1, vt, permute_cols=True)
    mtx = eng.MtxData.stencil(7, N, N, N)
    scs = eng.convert_to_scs(mtx, 32, 1, vt)
    eng.permute_scs_cols(scs)
    del mtx
    npt = {"dp": np.float64, "sp": np.float32}[vt]
    X = np.random.default_rng(bvs).uniform(-1.0, 1.0, nr * bvs).astype(npt)   # X[col * bvs + v], sigma = 1: no permutation
    Y_ref = refs.row.spmmv_scs(ref, X, bvs, nr)
    Xd = t.from_numpy(X).cuda()
    Yd = t.full((nr * bvs,), float("nan"), dtype=Xd.dtype, device="cuda")
    eng.spmmv(scs, Xd, Yd, bvs, nr, "rowwise")
    t.cuda.synchronize()
    got = Yd.cpu().numpy()
    assert np.array_equal(got.view(np.uint8), Y_ref.view(np.uint8)), (vt, bvs)


def _slab_thickness():
    try:
        import psutil
        avail = psutil.virtual_memory().available / 2 ** 30
    except Exception:
        avail = 32
    return 64 if avail >= 110 else (32 if avail >= 56 else 16)


def test_512_27pt_slab_with_halo_equals_the_reference(eng, refs, orc, pkg, big_gpu):
    """One z-slab of BASELINE config 5 (27-point, 512^3) as an interior rank of a seg-rows partition sees it, in the reference's
    order of operations, replicated bit for bit (strict_reference_halo: the padding's column 0 is a halo element of rank 0)."""
    t = big_gpu
    import ctypes as C
    capi, d = pkg.capi, pkg.dist
    n, th = 512, _slab_thickness()
    P = n // th
    rank = P // 2 - 1 if P > 2 else 0
    rows = n * n * th
    wsa = (np.arange(P + 1, dtype=np.int64) * rows).astype(np.int32) if P * rows < 2 ** 31 else None
    assert wsa is not None
    row0, row1 = int(wsa[rank]), int(wsa[rank + 1])
    nr, nc, I, J, V = orc.stencil_coo(27, n, n, n, row0, row1)
    # reference: convert_to_scs on the slab (global columns) -> collect_local_needed_heri -> permute_scs_cols
    h = refs.col.build_handle(nr, nc, I, J, V, 32, 1, "dp")
    need_ref, cum_ref = refs.col.collect_halo(h, wsa, rank)
    s0 = refs.col.export(h, "dp")
    refs.col.lib.ref_scs_permute_cols(h, s0.old_to_new.ctypes.data_as(C.c_void_p))
    ref = refs.col.export(h, "dp")
    refs.col.lib.ref_scs_free(h)
    del I, J, V, s0
    capi.set_option("strict_reference_halo", 1)
    try:
        mtx = eng.MtxData.stencil(27, n, n, n, row0, row1)
        scs = eng.convert_to_scs(mtx, 32, 1, "dp")
        del mtx
        plan = d.HaloPlan(scs, wsa, rank, P)
        eng.permute_scs_cols(scs)
    finally:
        capi.set_option("strict_reference_halo", 0)
    assert np.array_equal(plan.recv_cumsum, cum_ref)
    for p in range(P):
        assert np.array_equal(plan.need_lists[p], need_ref[p]), p
    got = scs.export()
    _same_struct(got, ref, "27-point slab")
    n_halo = int(cum_ref[-1])
    assert n_halo >= 2 * n * n  # both z-neighbours' faces (+ the spurious padding element, if the slab had padding)
    # y with a filled halo: local part random, halo random; sigma = 1 so x needs no permutation
    rng = np.random.default_rng(5)
    xv = rng.uniform(-1.0, 1.0, nr + max(n_halo, ref.n_rows_padded - nr))
    y_ref = refs.col.spmv_scs(ref, xv, adv=True)
    xd = t.from_numpy(xv).cuda()
    yd = t.full((scs.n_rows_padded,), float("nan"), dtype=t.float64, device="cuda")
    eng.spmv(scs, xd, yd)
    t.cuda.synchronize()
    assert np.array_equal(yd.cpu().numpy().view(np.uint64), y_ref.view(np.uint64))
    # the interior / boundary classification used for the overlap covers exactly the chunks that read a halo column
    ni, nb = C.c_long(0), C.c_long(0)
    capi.call("uspmv_scs_split_chunks", scs.h, C.byref(ni), C.byref(nb))
    touches = (ref.col_idxs >= nr)
    chunk_of = np.repeat(np.arange(ref.n_chunks), np.diff(ref.chunk_ptrs.astype(np.int64)))
    assert int(nb.value) == len(np.unique(chunk_of[touches])) and int(ni.value) + int(nb.value) == ref.n_chunks


LOG2_AP = 22
MODES = ("ap[dp_sp]", "ap[dp_hp]", "ap[sp_hp]", "ap[dp_sp_hp]")
USED = {"ap[dp_sp]": (0, 1), "ap[dp_hp]": (0, 2), "ap[sp_hp]": (1, 2), "ap[dp_sp_hp]": (0, 1, 2)}
VTS = ("dp", "sp", "hp")
NPT = {"dp": np.float64, "sp": np.float32, "hp": np.float16}


@pytest.fixture(scope="module")
def powerlaw_host(eng, big_gpu):
    n = 1 << LOG2_AP
    mtx = eng.MtxData.powerlaw(n)
    I, J, V = mtx.to_host()
    return n, I, J, V


def test_powerlaw_device_generator_equals_host_generator(eng, mats, big_gpu):
    """uspmv_coo_powerlaw == matrices.powerlaw_coo (numpy): rows / columns exactly; values to the last bit of pow() (the two pow
    implementations may differ by one ulp)."""
    n = 1 << 16
    mtx = eng.MtxData.powerlaw(n, d_min=mats.POWERLAW_D_MIN_CONFIG4)
    gI, gJ, gV = mtx.to_host()
    _, _, I, J, V = mats.powerlaw_coo(n, 0, d_min=mats.POWERLAW_D_MIN_CONFIG4)
    assert np.array_equal(gI, I) and np.array_equal(gJ, J)
    assert np.all(np.abs(gV - V) <= 1e-15 * np.abs(V))   # CUDA pow: <= 2 ulp, glibc pow: < 1 ulp
    # any row range generates independently (what every rank does at N > 1)
    a, b = n // 3, n // 3 + 5000
    part = eng.MtxData.powerlaw(n, a, b, d_min=mats.POWERLAW_D_MIN_CONFIG4)
    pI, pJ, pV = part.to_host()
    sel = (gI >= a) & (gI < b)
    assert np.array_equal(pI, gI[sel] - a) and np.array_equal(pJ, gJ[sel]) and np.array_equal(pV, gV[sel])


@pytest.mark.parametrize("mode", MODES)
def test_powerlaw_ap_modes_equal_oracle_and_reference(eng, orc, refs, pkg, big_gpu, powerlaw_host, mode):
    """BASELINE config 4 at 2^22 rows (t1 = 1.0, t2 = 1e-2, C = 32, sigma = 512): partition and structures bit-exact against the oracle
    (pinned to the reference, tests/test_oracle_pinning.py), y against the reference's interface.hpp AP kernels — bit-identical with
    long chunks summed sequentially, within 1e-12 (1e-5 for sp_hp) with the default segmented sums."""
    t = big_gpu
    n, I, J, V = powerlaw_host
    t1, t2 = 1.0, 1e-2
    assert 14.0 < len(I) / n < 16.0  # the density BASELINE asks for (5.0e8 elements at 2^25 rows)
    part, counts = orc.partition_precisions(mode, I, J, V, t1, t2)
    mtx = eng.MtxData.from_host(n, n, I, J, V)
    coos = eng.partition_precisions(mtx, mode, t1, t2)
    del mtx
    used = USED[mode]
    ref_parts, dev_parts = [None] * 3, [None] * 3
    perm = None
    for k, p in enumerate(used):
        sel = part == p
        assert coos[p].nnz == counts[p] == int(sel.sum())
        ref_parts[p] = orc.convert_to_scs(n, n, I[sel], J[sel], V[sel], 32, 512, VTS[p], fixed_perm=perm)
        dev_parts[p] = eng.convert_to_scs(coos[p], 32, 512, VTS[p], fixed_permutation=perm)
        g = dev_parts[p].export()
        if k == 0:
            perm = ref_parts[p].old_to_new
            assert np.array_equal(g.old_to_new, perm)
        for f in ("chunk_ptrs", "chunk_lengths", "col_idxs"):
            assert np.array_equal(getattr(g, f), getattr(ref_parts[p], f)), (mode, p, f)
        assert np.array_equal(g.values.view(np.uint8), ref_parts[p].values.view(np.uint8)), (mode, p)
        del g
    del coos
    x = np.random.default_rng(9).uniform(-1.0, 1.0, n)
    x32 = x.astype(np.float32)
    y_ref = refs.iface.ap_scs(mode, ref_parts[0], ref_parts[1], ref_parts[2], x, x32)
    dt = t.float32 if mode == "ap[sp_hp]" else t.float64
    xd = t.from_numpy(x32 if mode == "ap[sp_hp]" else x).cuda()
    yd = t.full((ref_parts[used[0]].n_rows_padded,), float("nan"), dtype=dt, device="cuda")
    pkg.capi.set_option("split_long_chunks", 0)
    try:
        eng.ap_spmv(mode, dev_parts[0], dev_parts[1], dev_parts[2], xd, yd)
        t.cuda.synchronize()
        y_seq = yd.cpu().numpy()
    finally:
        pkg.capi.set_option("split_long_chunks", 256)
    assert np.array_equal(y_seq.view(np.uint8), y_ref.view(np.uint8)), f"{mode}: y differs from the reference's interface.hpp kernel"
    # default plan (long chunks cut into 256-slot segments, partial sums added in slot order): tolerance
    yd.fill_(float("nan"))
    eng.ap_spmv(mode, dev_parts[0], dev_parts[1], dev_parts[2], xd, yd)  # the plan is keyed on the option and rebuilt
    t.cuda.synchronize()
    scale = np.zeros(n)
    np.add.at(scale, I, np.abs(V * x[J]))
    sp = np.zeros(len(y_ref))
    sp[perm] = scale
    tol = 1e-5 if mode == "ap[sp_hp]" else 1e-12
    assert np.all(np.abs(yd.cpu().numpy().astype(np.float64) - y_ref.astype(np.float64)) <= tol * np.maximum(sp, 1e-300)), mode
