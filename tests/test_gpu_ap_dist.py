"""GPU parity: adaptive precision (partition + fused kernels) and the row-partitioned path (halo discovery,
pack, interior/boundary split) — every rank of a P-rank run is emulated serially on one GPU, the exchange is
a device-to-device copy; the real N > 1 exchange is covered by tests/test_dist_cpu.py (gloo) and bench.py."""
import numpy as np
import pytest

from conftest import load_matrix

pytestmark = pytest.mark.gpu

MODES = ("ap[dp_sp]", "ap[dp_hp]", "ap[sp_hp]", "ap[dp_sp_hp]")
USED = {"ap[dp_sp]": (0, 1), "ap[dp_hp]": (0, 2), "ap[sp_hp]": (1, 2), "ap[dp_sp_hp]": (0, 1, 2)}
VTS = ("dp", "sp", "hp")
NPT = {"dp": np.float64, "sp": np.float32, "hp": np.float16}


def torch_():
    import torch
    return torch


def dev(a):
    return torch_().from_numpy(np.ascontiguousarray(a)).cuda()


def ap_matrix(mats, n=4096, seed=3):
    rng = np.random.default_rng(seed)
    _, _, I, J, V = mats.random_coo(n, 8, seed=seed, empty_rows=False)
    V = np.sign(V) * 10.0 ** rng.uniform(-3, 1, len(V))
    return n, n, I, J, V


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("equil", [False, True])
def test_partition_precisions(eng, orc, mats, mode, equil):
    n, nc, I, J, V = ap_matrix(mats)
    t1, t2 = 0.5, 0.01
    rm = cm = None
    if equil:
        rm, cm = orc.largest_elems(n, nc, I, J, V)
        rm[rm == 0] = 1.0
        cm[cm == 0] = 1.0
    part, counts = orc.partition_precisions(mode, I, J, V, t1, t2, rm, cm)
    mtx = eng.MtxData.from_host(n, nc, I, J, V)
    got = eng.partition_precisions(mtx, mode, t1, t2, rm, cm)
    for p in range(3):
        if p not in USED[mode]:
            assert got[p] is None
            continue
        gI, gJ, gV = got[p].to_host(NPT[VTS[p]])
        sel = part == p
        assert np.array_equal(gI, I[sel]) and np.array_equal(gJ, J[sel])
        assert np.array_equal(gV.view(np.uint8), V[sel].astype(NPT[VTS[p]]).view(np.uint8))
        assert got[p].nnz == counts[p]


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("C,sigma", [(1, 1), (4, 8), (32, 128), (64, 64)])
def test_ap_spmv_fused(eng, orc, mats, mode, C, sigma):
    t = torch_()
    n, nc, I, J, V = ap_matrix(mats, seed=C)
    part, _ = orc.partition_precisions(mode, I, J, V, 0.5, 0.01)
    mtx = eng.MtxData.from_host(n, nc, I, J, V)
    coos = eng.partition_precisions(mtx, mode, 0.5, 0.01)
    used = USED[mode]
    ref_parts, dev_parts = [None] * 3, [None] * 3
    sel = part == used[0]
    ref_parts[used[0]] = orc.convert_to_scs(n, nc, I[sel], J[sel], V[sel], C, sigma, VTS[used[0]])
    dev_parts[used[0]] = eng.convert_to_scs(coos[used[0]], C, sigma, VTS[used[0]])
    perm = ref_parts[used[0]].old_to_new
    assert np.array_equal(dev_parts[used[0]].export().old_to_new, perm)
    for p in used[1:]:
        sel = part == p
        ref_parts[p] = orc.convert_to_scs(n, nc, I[sel], J[sel], V[sel], C, sigma, VTS[p], fixed_perm=perm)
        dev_parts[p] = eng.convert_to_scs(coos[p], C, sigma, VTS[p], fixed_permutation=perm)
        g = dev_parts[p].export()
        for k in ("chunk_ptrs", "chunk_lengths", "col_idxs"):
            assert np.array_equal(getattr(g, k), getattr(ref_parts[p], k)), (mode, p, k)
        assert np.array_equal(g.values.view(np.uint8), ref_parts[p].values.view(np.uint8))
    x = np.random.default_rng(5).uniform(-1, 1, n)
    y_ref = orc.ap_scs(mode, ref_parts[0], ref_parts[1], ref_parts[2], x, x.astype(np.float32))
    n_pad = ref_parts[used[0]].n_rows_padded
    if mode == "ap[sp_hp]":
        xd, yd = dev(x.astype(np.float32)), t.zeros(n_pad, dtype=t.float32, device="cuda")
    else:
        xd, yd = dev(x), t.zeros(n_pad, dtype=t.float64, device="cuda")
    eng.ap_spmv(mode, dev_parts[0], dev_parts[1], dev_parts[2], xd, yd)
    t.cuda.synchronize()
    y = yd.cpu().numpy()
    tol = 1e-5 if mode == "ap[sp_hp]" else 1e-12
    scale = np.zeros(n)
    np.add.at(scale, I, np.abs(V * x[J]))
    sp_ = np.zeros(n_pad)
    sp_[perm] = scale
    assert np.all(np.abs(y.astype(np.float64) - y_ref.astype(np.float64)) <= tol * np.maximum(sp_, 1e-300))
    assert np.array_equal(y.view(np.uint8), y_ref.view(np.uint8)), "fused AP kernel is expected to be bit-identical to the oracle"


@pytest.mark.parametrize("mode", MODES)
def test_ap_spmv_uneven_long_chunks(eng, pkg, orc, mode):
    """Power-law-like rows: chunks whose parts hold more than `split_long_chunks` slots are summed in per-part segments
    (deterministic, within tolerance); un-split chunks and the option-off run stay bit-identical to the oracle."""
    t = torch_()
    rng = np.random.default_rng(11)
    n = 140000
    cnt = rng.integers(1, 8, n)
    long_rows = rng.choice(n, 40, replace=False)
    cnt[long_rows] = rng.integers(600, 3000, 40)
    I = np.repeat(np.arange(n), cnt).astype(np.int32)
    J = rng.integers(0, n, len(I)).astype(np.int32)
    V = np.sign(rng.standard_normal(len(I))) * 10.0 ** rng.uniform(-3, 1, len(I))
    C, sigma = 32, 128
    part, _ = orc.partition_precisions(mode, I, J, V, 0.5, 0.01)
    used = USED[mode]
    ref_parts = [None] * 3
    sel = part == used[0]
    ref_parts[used[0]] = orc.convert_to_scs(n, n, I[sel], J[sel], V[sel], C, sigma, VTS[used[0]])
    perm = ref_parts[used[0]].old_to_new
    for p in used[1:]:
        sel = part == p
        ref_parts[p] = orc.convert_to_scs(n, n, I[sel], J[sel], V[sel], C, sigma, VTS[p], fixed_perm=perm)
    x = rng.uniform(-1, 1, n)
    y_ref = orc.ap_scs(mode, ref_parts[0], ref_parts[1], ref_parts[2], x, x.astype(np.float32))
    n_pad = ref_parts[used[0]].n_rows_padded
    tot = sum(ref_parts[p].chunk_lengths[: n_pad // 32].astype(np.int64) for p in used)
    assert tot.max() > 256
    scale = np.zeros(n)
    np.add.at(scale, I, np.abs(V * x[J]))
    sp_ = np.zeros(n_pad)
    sp_[perm] = scale
    tol = 1e-5 if mode == "ap[sp_hp]" else 1e-12
    mtx = eng.MtxData.from_host(n, n, I, J, V)
    coos = eng.partition_precisions(mtx, mode, 0.5, 0.01)
    dev_parts = [None] * 3
    dev_parts[used[0]] = eng.convert_to_scs(coos[used[0]], C, sigma, VTS[used[0]])
    for p in used[1:]:
        dev_parts[p] = eng.convert_to_scs(coos[p], C, sigma, VTS[p], fixed_permutation=perm)
    outs = {}
    try:
        for split in (256, 0, 64):
            pkg.capi.set_option("split_long_chunks", split)
            if mode == "ap[sp_hp]":
                xd, yd = dev(x.astype(np.float32)), t.zeros(n_pad, dtype=t.float32, device="cuda")
            else:
                xd, yd = dev(x), t.zeros(n_pad, dtype=t.float64, device="cuda")
            for _ in range(2):  # second call re-uses the cached plan
                eng.ap_spmv(mode, dev_parts[0], dev_parts[1], dev_parts[2], xd, yd)
            t.cuda.synchronize()
            outs[split] = yd.cpu().numpy()
    finally:
        pkg.capi.set_option("split_long_chunks", 256)
    assert np.array_equal(outs[0].view(np.uint8), y_ref.view(np.uint8)), "without segments: bit-identical to the oracle"
    for split in (256, 64):
        assert np.all(np.abs(outs[split].astype(np.float64) - y_ref.astype(np.float64)) <= tol * np.maximum(sp_, 1e-300)), split
        short = np.repeat(tot <= split, 32)
        assert np.array_equal(outs[split][short].view(np.uint8), y_ref[short].view(np.uint8)), "rows of un-split chunks stay bit-identical"


def test_ap_golden_fixtures(eng):
    """y of the reference's interface.hpp AP kernels (tests/golden/ref_ap.npz, generated from the real reference)."""
    import os
    from conftest import GOLDEN
    t = torch_()
    z = np.load(os.path.join(GOLDEN, "ref_ap.npz"))
    for key in z["cases"]:
        name, mode, C, sigma = key.split("|")
        C, sigma = int(C), int(sigma)
        n, nc, I, J, V = load_matrix(name)
        t1, t2 = z[key + "|t"]
        mtx = eng.MtxData.from_host(n, nc, I, J, V)
        coos = eng.partition_precisions(mtx, mode, t1, t2)
        used = USED[mode]
        parts = [None] * 3
        parts[used[0]] = eng.convert_to_scs(coos[used[0]], C, sigma, VTS[used[0]])
        perm = parts[used[0]].export().old_to_new
        assert np.array_equal(perm, z[key + "|perm"]), key
        for p in used[1:]:
            parts[p] = eng.convert_to_scs(coos[p], C, sigma, VTS[p], fixed_permutation=perm)
        x = np.random.default_rng(1234).uniform(-1.0, 1.0, n)
        n_pad = parts[used[0]].n_rows_padded
        xp = np.zeros(max(n_pad, n))
        xp[:n] = x
        sp_mode = mode == "ap[sp_hp]"
        xd = dev(xp.astype(np.float32) if sp_mode else xp)
        yd = t.zeros(n_pad, dtype=t.float32 if sp_mode else t.float64, device="cuda")
        eng.ap_spmv(mode, parts[0], parts[1], parts[2], xd, yd)
        t.cuda.synchronize()
        y = yd.cpu().numpy()[perm]
        assert np.array_equal(y.view(np.uint8), z[key + "|y"].view(np.uint8)), key


# ------------------------------------------------------------------------------------------------
# row partitioning + halo
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("method", ["seg-rows", "seg-nnz"])
@pytest.mark.parametrize("P", [2, 3, 5])
def test_seg_work_sharing_arr(eng, orc, mats, method, P):
    for name in ("bcsstk13", "FDM-2d-16", "impcol_e"):
        n, nc, I, J, V = load_matrix(name)
        assert np.array_equal(eng.seg_work_sharing_arr(method, n, I, P), orc.seg_work_sharing_arr(method, n, I, P)), (name, P, method)


def _rank_slab(I, J, V, wsa, r):
    sel = (I >= wsa[r]) & (I < wsa[r + 1])
    return (I[sel] - wsa[r]).astype(np.int32), J[sel], V[sel]


def _pad_mask(ref, lI, n_loc):
    """True for the padding slots of an SCS structure (slot j >= stored elements of the row at that position)."""
    cnt_row = np.bincount(lI, minlength=n_loc)
    cnt_pos = np.zeros(ref.n_rows_padded, np.int64)
    cnt_pos[ref.old_to_new] = cnt_row
    mask = np.zeros(ref.n_elements, bool)
    C = ref.C
    for c in range(ref.n_chunks):
        L = int(ref.chunk_lengths[c])
        if L == 0:
            continue
        j = np.repeat(np.arange(L), C)
        lane = np.tile(np.arange(C), L)
        mask[ref.chunk_ptrs[c]: ref.chunk_ptrs[c] + L * C] = j >= cnt_pos[c * C + lane]
    return mask


@pytest.mark.parametrize("strict", [True, False])
@pytest.mark.parametrize("P", [2, 4])
@pytest.mark.parametrize("C,sigma", [(1, 1), (8, 16), (32, 64)])
def test_halo_plan_pack_and_distributed_spmv(eng, pkg, orc, mats, P, C, sigma, strict):
    """P ranks emulated on one GPU: bit-exact halo renumbering / need lists, pack, interior+boundary == full SpMV,
    and the assembled distributed result equals the single-rank result."""
    t = torch_()
    dist = pkg.dist
    n, nc, I, J, V = mats.random_coo(3001, 7, seed=P + C, empty_rows=False)
    x_glob = np.random.default_rng(2).standard_normal(n)
    wsa = eng.seg_work_sharing_arr("seg-nnz", n, I, P)
    ranks = []
    # strict: padding's column 0 is GLOBAL column 0 -> a halo element from rank 0 on ranks > 0 (the reference);
    # default: padding stays local (column 0 of the rank)
    pkg.capi.set_option("strict_reference_halo", 1 if strict else 0)
    for r in range(P):
        lI, lJ, lV = _rank_slab(I, J, V, wsa, r)
        n_loc = int(wsa[r + 1] - wsa[r])
        ref = orc.convert_to_scs(n_loc, n, lI, lJ, lV, C, sigma)
        if not strict:
            ref.col_idxs[_pad_mask(ref, lI, n_loc)] = wsa[r]
        need_ref, cum_ref = orc.collect_halo(ref.col_idxs, wsa, r)
        orc.permute_scs_cols(ref, ref.old_to_new)
        mtx = eng.MtxData.from_host(n_loc, n, lI, lJ, lV)
        scs = eng.convert_to_scs(mtx, C, sigma, "dp")
        plan = dist.HaloPlan(scs, wsa, r, P)
        eng.permute_scs_cols(scs)
        got = scs.export()
        assert np.array_equal(got.col_idxs, ref.col_idxs), (r, "halo-renumbered + permuted columns")
        assert np.array_equal(plan.recv_cumsum, cum_ref)
        for a, b in zip(plan.need_lists, need_ref):
            assert np.array_equal(a, b)
        ranks.append(dict(scs=scs, plan=plan, ref=ref, n_loc=n_loc))
    pkg.capi.set_option("strict_reference_halo", 0)
    # comm schedule = transpose of the need lists (collect_comm_idxs)
    for r in range(P):
        ranks[r]["plan"].set_send([ranks[q]["plan"].need_lists[r] for q in range(P)])
    # vectors in permuted space, pack, "exchange" by device copies
    for r in range(P):
        d = ranks[r]
        s, ref = d["scs"], d["ref"]
        xl = np.zeros(d["n_loc"] + max(s.n_rows_padded - s.n_rows, d["plan"].n_halo))
        xl[ref.old_to_new] = x_glob[wsa[r]:wsa[r + 1]]
        d["x"] = dev(xl)
        d["send"] = t.zeros(max(d["plan"].n_send, 1), dtype=t.float64, device="cuda")
        d["plan"].pack(d["x"], d["send"])
    t.cuda.synchronize()
    for r in range(P):
        d = ranks[r]
        exp = np.concatenate([x_glob[wsa[r]:wsa[r + 1]][lst] for lst in d["plan"].send_lists]) if d["plan"].n_send else np.zeros(0)
        assert np.array_equal(d["send"].cpu().numpy()[:d["plan"].n_send], exp), "pack kernel"
    for r in range(P):
        d = ranks[r]
        for p in range(P):
            cnt = int(d["plan"].recv_cumsum[p + 1] - d["plan"].recv_cumsum[p])
            if cnt:
                src = ranks[p]
                o = int(src["plan"].send_ptr[r])
                d["x"][d["n_loc"] + int(d["plan"].recv_cumsum[p]): d["n_loc"] + int(d["plan"].recv_cumsum[p + 1])] = src["send"][o:o + cnt]
    y_glob = np.zeros(n)
    import ctypes as C_
    for r in range(P):
        d = ranks[r]
        s, ref = d["scs"], d["ref"]
        y_full = t.zeros(s.n_rows_padded, dtype=t.float64, device="cuda")
        eng.spmv(s, d["x"], y_full)
        ni, nb = C_.c_long(0), C_.c_long(0)
        pkg.capi.call("uspmv_scs_split_chunks", s.h, C_.byref(ni), C_.byref(nb))
        assert ni.value + nb.value == s.n_chunks
        y_parts = t.full((s.n_rows_padded,), float("nan"), dtype=t.float64, device="cuda")
        pkg.capi.call("uspmv_spmv_part", s.h, 1, C_.c_void_p(d["x"].data_ptr()), C_.c_void_p(y_parts.data_ptr()), None)
        pkg.capi.call("uspmv_spmv_part", s.h, 2, C_.c_void_p(d["x"].data_ptr()), C_.c_void_p(y_parts.data_ptr()), None)
        t.cuda.synchronize()
        if C == 1 and sigma == 1:  # full launch = CRS kernel (split-row reduction), parts = sequential SCS kernel
            assert t.allclose(y_full, y_parts, rtol=1e-12, atol=1e-13)
        else:
            assert t.equal(y_full, y_parts), "interior + boundary launches must equal the full launch"
        y_ref = orc.spmv_scs(ref, d["x"].cpu().numpy()) if C > 1 or sigma > 1 else None
        if y_ref is not None:
            assert np.array_equal(y_full.cpu().numpy(), y_ref)
        y_glob[wsa[r]:wsa[r + 1]] = y_full.cpu().numpy()[ref.old_to_new]
    # against the plain COO product
    y_coo = np.zeros(n)
    np.add.at(y_coo, I, V * x_glob[J])
    scale = np.zeros(n)
    np.add.at(scale, I, np.abs(V * x_glob[J]))
    assert np.all(np.abs(y_glob - y_coo) <= 1e-12 * np.maximum(scale, 1e-300))
