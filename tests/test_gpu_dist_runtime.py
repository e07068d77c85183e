"""GPU runtime tests of the one-process-per-GPU path (dist.py + uspmv_p2p_*): world_size 1 in-process (fused kernel
with no peers) and, when the box has >= 2 GPUs, a real 2-rank NVLink run for every exchange mode."""
import importlib
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# (block_vec_size, exchange mode): 2 = ONE fused kernel (C = 32, block_vec_size 2 / 4 / 8 / 16), 1 = push / wait kernels next to the
# interior kernel, 0 = exchange first; block_vec_size 3 has no streamed kernel and falls back to exchange + one full SpMMV
BLOCK_CASES = ((4, 2), (8, 2), (4, 1), (4, 0), (3, 1), (3, 2))


def _init(rank, world, port, same_device):
    """same_device: every rank on cuda:0 (the driver's one-GPU test box).  The ranks are still separate processes whose arenas are
    exchanged through CUDA IPC, so pushes, flag waits and acknowledgements run against a real peer; the GPU time-slices between the
    two contexts (a rank spinning on a flag is preempted so that its neighbour can raise it).  NCCL refuses two ranks on one device,
    so the plumbing (all_gather_object / barrier) runs over gloo there."""
    import torch
    import torch.distributed as dist
    dev = 0 if same_device else rank
    torch.cuda.set_device(dev)
    if same_device:
        dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    else:
        dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world, device_id=torch.device("cuda", dev))
    return dev


def _run_rank(rank, world, port, out_dir, C=32, same_device=False):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    dev = _init(rank, world, port, same_device)
    try:
        pkg = importlib.import_module("ultimate-spmv_b200")
        eng, capi, d = pkg.engine, pkg.capi, pkg.dist
        n = 48
        results = {}
        # pv > 0: the large-halo push kernels (tiled direct stores / shared memory + bulk copy / 4096-element tiles) forced onto this
        # small halo, here with the permuted-x gather (perm != NULL); the un-permuted form is covered by the AP test below
        cases = (("p2p", 2, 0), ("p2p", 1, 0), ("p2p", 0, 0), ("nccl", True, 0), ("nccl", False, 0), ("p2p", 1, 1), ("p2p", 0, 2), ("p2p", 1, 3))
        if same_device:
            cases = tuple(c for c in cases if c[0] == "p2p")
        for halo, mode, pv in cases:
            capi.set_option("push_variant", pv if pv else -1)
            capi.set_option("push_min_elements", 0 if pv else 1 << 20)
            r = d.DistributedSpmv(eng.default_context(dev), 27, n, C, 64, "dp", rank, world, overlap=mode, halo=halo)
            # x = global row index pattern so that halo values are distinguishable
            rows = torch.arange(rank * n ** 3, (rank + 1) * n ** 3, device="cuda", dtype=torch.float64)
            xs = torch.sin(rows * 0.37) + 1.5
            perm = torch.from_numpy(r.scs.export().old_to_new.astype(np.int64)).cuda()
            r.x[: r.scs.n_rows][perm] = xs
            for _ in range(3):  # several steps: epochs, acks and buffer reuse
                r.y.zero_()
                r.step()
            torch.cuda.synchronize()
            if r.p2p is not None:
                err, ep = r.p2p.status()
                assert err == 0 and ep == 3
            results[f"{halo}{mode}pv{pv}"] = r.y[: r.scs.n_rows_padded][perm].cpu().numpy()
            # the runner's own checked step (what bench.py runs before timing): NaN-poisoned halo, y against the stencil formula
            assert r.validate() <= 1e-12, (halo, mode, pv)
            r.close()  # collective teardown: sync, barrier, unmap the neighbours' arenas, barrier, free
            del r
        capi.set_option("push_variant", -1)
        capi.set_option("push_min_elements", 1 << 20)
        # narrow value types: mode 2 defaults to the push / wait kernels next to the single-GPU interior kernel (the fused instance is
        # issue-bound there); "fused_narrow" keeps the fused kernel reachable — both against the stencil formula, solve buffers included
        for vt, tol in (("sp", 1e-5), ("hp", 1e-2)):
            for narrow in (0, 1):
                capi.set_option("fused_narrow", narrow)
                r = d.DistributedSpmv(eng.default_context(dev), 27, n, C, 64, vt, rank, world, overlap=2, halo="p2p")
                assert r.validate(steps=3) <= tol, (vt, narrow)
                r.close()
                del r
        capi.set_option("fused_narrow", 0)
        base = results["p2p2pv0"]
        for k, v in results.items():
            assert np.array_equal(v, base), k
        np.save(os.path.join(out_dir, f"y{rank}.npy"), base)

        # ---- block vectors (bulkvec exchange): all bvs vectors of the halo rows in one push, both layouts, with / without overlap
        rows = torch.arange(rank * n ** 3, (rank + 1) * n ** 3, device="cuda", dtype=torch.float64)
        for layout in ("rowwise", "colwise"):
            for bvs, mode in BLOCK_CASES:
                r = d.DistributedSpmv(eng.default_context(dev), 27, n, C, 64, "dp", rank, world, overlap=mode, halo="p2p", bvs=bvs, layout=layout)
                perm = torch.from_numpy(r.scs.export().old_to_new.astype(np.int64)).cuda()
                nl, ld = r.scs.n_rows, r.vec_length
                r.x.zero_()
                for v in range(bvs):
                    xs = torch.sin(rows * (0.37 + 0.11 * v)) + 1.5
                    if layout == "rowwise":
                        r.x[:nl * bvs].view(nl, bvs)[perm, v] = xs
                    else:
                        r.x[v * ld: v * ld + nl][perm] = xs
                torch.cuda.synchronize()
                dist.barrier()  # a fast neighbour must not push its halo into x before the zero_() above has run
                for _ in range(3):
                    r.y.zero_()
                    r.step()
                torch.cuda.synchronize()
                err, ep = r.p2p.status()
                assert err == 0 and ep == 3
                out = np.zeros((bvs, nl))
                for v in range(bvs):
                    if layout == "rowwise":
                        out[v] = r.y[: r.scs.n_rows_padded * bvs].view(-1, bvs)[perm, v].cpu().numpy()
                    else:
                        out[v] = r.y[v * ld: v * ld + r.scs.n_rows_padded][perm].cpu().numpy()
                np.save(os.path.join(out_dir, f"Y{rank}_{layout}_{bvs}_{mode}.npy"), out)
                assert r.validate() <= 1e-12, (layout, bvs, mode)
                r.close()
                del r
        # row-major block vectors through the push / wait kernels next to the interior kernel as well (mode 2 defaults to the fused step)
        capi.set_option("mmv_fused_rowwise", 0)
        r = d.DistributedSpmv(eng.default_context(dev), 27, n, C, 64, "dp", rank, world, overlap=2, halo="p2p", bvs=4, layout="rowwise")
        assert r.validate() <= 1e-12
        r.close()
        del r
        capi.set_option("mmv_fused_rowwise", 1)

        # ---- device-resident solve loop: 3 x { exchange ; SpMV ; swap } over the two arena buffers, every overlap mode
        for mode in (2, 1, 0):
            r = d.DistributedSpmv(eng.default_context(dev), 27, n, C, 64, "dp", rank, world, overlap=mode, halo="p2p", n_buf=2)
            perm = torch.from_numpy(r.scs.export().old_to_new.astype(np.int64)).cuda()
            for b in r.p2p.bufs:
                b.zero_()
            r.p2p.bufs[0][: r.scs.n_rows][perm] = (torch.sin(rows * 0.37) + 1.5) / 64.0
            torch.cuda.synchronize()
            dist.barrier()
            xf = r.solve(3)
            torch.cuda.synchronize()
            err, ep = r.p2p.status()
            assert err == 0 and ep == 3
            np.save(os.path.join(out_dir, f"solve{rank}_{mode}.npy"), xf[: r.scs.n_rows][perm].cpu().numpy())
            # ---- host-buffer calls pipelined over the same two buffers (uspmv_p2p_spmv_host_submit / _wait): 5 calls with different x,
            #      two in flight; every y must equal the device-resident SpMV of that x
            nl = r.scs.n_rows
            xs_h = [((torch.sin(rows * (0.21 + 0.05 * k)) + 1.5).cpu()) for k in range(5)]
            xh = [torch.empty(nl, dtype=torch.float64).pin_memory() for _ in range(5)]
            yh = [torch.zeros(r.scs.n_rows_padded, dtype=torch.float64).pin_memory() for _ in range(5)]
            pc = perm.cpu()
            for k in range(5):
                xh[k][pc] = xs_h[k]
            torch.cuda.synchronize()
            dist.barrier()
            for k in range(5):
                capi.call("uspmv_p2p_spmv_host_wait", r.p2p.h, k & 1)
                capi.call("uspmv_p2p_spmv_host_submit", r.p2p.h, r.scs.h, xh[k].data_ptr(), yh[k].data_ptr(), k & 1)
            for sl in (0, 1):
                capi.call("uspmv_p2p_spmv_host_wait", r.p2p.h, sl)
            with pytest.raises(capi.UspmvError):  # the slots alternate with the arena buffers
                capi.call("uspmv_p2p_spmv_host_submit", r.p2p.h, r.scs.h, xh[0].data_ptr(), yh[0].data_ptr(), 0)  # call 5 belongs to slot 1
            err, ep2 = r.p2p.status()
            assert err == 0 and ep2 == ep + 5
            np.save(os.path.join(out_dir, f"hosty{rank}_{mode}.npy"), np.stack([yh[k][: r.scs.n_rows_padded][pc].numpy() for k in range(5)]))
            r.close()
            del r
        dist.barrier()
    finally:
        dist.destroy_process_group()


def _check(out_dir, world, mats):
    n = 48
    nr, nc, I, J, V = mats.stencil_coo(27, n, n, n * world)
    x = np.sin(np.arange(nr) * 0.37) + 1.5
    y = np.concatenate([np.load(os.path.join(out_dir, f"y{r}.npy")) for r in range(world)])
    y_ref = np.zeros(nr)
    np.add.at(y_ref, I, V * x[J])
    scale = np.zeros(nr)
    np.add.at(scale, I, np.abs(V * x[J]))
    assert np.all(np.abs(y - y_ref) <= 1e-12 * scale)
    # block vectors
    for layout in ("rowwise", "colwise"):
        for bvs, mode in BLOCK_CASES:
            Y = np.concatenate([np.load(os.path.join(out_dir, f"Y{r}_{layout}_{bvs}_{mode}.npy")) for r in range(world)], axis=1)
            for v in range(bvs):
                xv = np.sin(np.arange(nr) * (0.37 + 0.11 * v)) + 1.5
                ref = np.zeros(nr)
                np.add.at(ref, I, V * xv[J])
                sc = np.zeros(nr)
                np.add.at(sc, I, np.abs(V * xv[J]))
                assert np.all(np.abs(Y[v] - ref) <= 1e-12 * sc), (layout, bvs, mode, v)
    # solve loop: x3 = A^3 x0
    import scipy.sparse as sp
    A = sp.csr_matrix((V, (I, J)), shape=(nr, nr))
    x3 = (np.sin(np.arange(nr) * 0.37) + 1.5) / 64.0
    for _ in range(3):
        x3 = A @ x3
    for mode in (2, 1, 0):
        got = np.concatenate([np.load(os.path.join(out_dir, f"solve{r}_{mode}.npy")) for r in range(world)])
        assert np.max(np.abs(got - x3)) <= 1e-11 * np.max(np.abs(x3)), mode
        hy = np.concatenate([np.load(os.path.join(out_dir, f"hosty{r}_{mode}.npy")) for r in range(world)], axis=1)
        for k in range(5):
            xk = np.sin(np.arange(nr) * (0.21 + 0.05 * k)) + 1.5
            ref = A @ xk
            assert np.max(np.abs(hy[k] - ref)) <= 1e-12 * np.max(np.abs(ref)), (mode, k)


@pytest.mark.parametrize("C", [32, 16, 64])
def test_world_size_1_fused_kernel(eng, mats, tmp_path, C):
    """C = 32 / 64: fused kernel (SELL-32 / wide-chunk); C = 16: the P2P step falls back to the push / wait / ack kernels around the
    narrow-chunk SpMV kernel."""
    port = 29700 + os.getpid() % 200 + C
    _run_rank(0, 1, port, str(tmp_path), C)
    _check(str(tmp_path), 1, mats)


@pytest.mark.parametrize("C", [32, 16, 64])
def test_two_ranks_over_nvlink(eng, mats, tmp_path, C):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    port = 29900 + os.getpid() % 100 + C
    mp.spawn(_run_rank, args=(2, port, str(tmp_path), C), nprocs=2, join=True)
    _check(str(tmp_path), 2, mats)


@pytest.mark.parametrize("C", [32, 16, 64])
def test_two_ranks_on_one_gpu(eng, mats, tmp_path, C):
    """The same two-rank run with both processes on cuda:0 (see _init): every exchange mode — fused kernel, push / wait / ack kernels,
    exchange-then-SpMV, the tiled and bulk-copy push kernels, block vectors, the two-buffer solve loop and the pipelined host-buffer
    calls — against a real peer, on a box with ONE GPU."""
    import torch.multiprocessing as mp
    port = 29800 + os.getpid() % 100 + C
    mp.spawn(_run_rank, args=(2, port, str(tmp_path), C, True), nprocs=2, join=True)
    _check(str(tmp_path), 2, mats)


# ---- adaptive precision, row-partitioned with seg-nnz (BASELINE config 4) --------------------------------------------------------
def _ap_matrix(n=6000, seed=5):
    rng = np.random.default_rng(seed)
    cnt = rng.integers(1, 12, n)
    cnt[rng.choice(n, 12, replace=False)] = rng.integers(200, 700, 12)       # a few long rows: uneven nnz per rank
    I = np.repeat(np.arange(n), cnt).astype(np.int32)
    J = rng.integers(0, n, len(I)).astype(np.int32)
    V = np.sign(rng.standard_normal(len(I))) * 10.0 ** rng.uniform(-3, 1, len(I))
    return n, I, J, V, cnt


def _run_ap_rank(rank, world, port, out_dir, same_device=False):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    dev = _init(rank, world, port, same_device)
    try:
        pkg = importlib.import_module("ultimate-spmv_b200")
        eng, d = pkg.engine, pkg.dist
        n, I, J, V, cnt = _ap_matrix()
        wsa = d.seg_nnz_from_row_counts(cnt, world)
        sel = (I >= wsa[rank]) & (I < wsa[rank + 1])
        n_loc = int(wsa[rank + 1] - wsa[rank])
        x_glob = np.sin(np.arange(n) * 0.37) + 1.5
        for mode in ("ap[dp_sp_hp]", "ap[dp_sp]", "ap[sp_hp]"):
            r = d.DistributedApSpmv(eng.default_context(dev), wsa, (n_loc, n, (I[sel] - wsa[rank]).astype(np.int32), J[sel], V[sel]), mode, 0.5, 0.01,
                                    32, 64, rank, world, keep_coo=True)
            r.x.zero_()
            r.x[:n_loc] = torch.from_numpy(x_glob[wsa[rank]:wsa[rank + 1]]).to(r.x.dtype).cuda()   # original local row order
            torch.cuda.synchronize()
            dist.barrier()
            for _ in range(3):
                r.y.zero_()
                r.step()
            torch.cuda.synchronize()
            err, ep = r.p2p.status()
            assert err == 0 and ep == 3
            # the large-halo push kernels (tiled direct stores / shared memory + bulk copy over NVLink), forced onto this small halo
            y0 = r.y.clone()
            for variant in (0, 1, 2):
                pkg.capi.set_option("push_variant", variant)
                pkg.capi.set_option("push_min_elements", 0)
                r.x[n_loc:] = 0
                r.y.zero_()
                torch.cuda.synchronize()
                dist.barrier()
                r.step()
                torch.cuda.synchronize()
                assert torch.equal(r.y, y0), (mode, variant)
            pkg.capi.set_option("push_variant", -1)
            pkg.capi.set_option("push_min_elements", 1 << 20)
            err, ep = r.p2p.status()
            assert err == 0 and ep == 6
            y = r.y.cpu().numpy()[r.old_to_new]
            np.save(os.path.join(out_dir, f"ap{rank}_{mode}.npy"), y.astype(np.float64))
            assert r.validate() <= (1e-5 if mode == "ap[sp_hp]" else 1e-12), mode
            r.close()
            del r
        np.save(os.path.join(out_dir, f"apwsa{rank}.npy"), wsa)
        dist.barrier()
    finally:
        dist.destroy_process_group()


def _check_ap(out_dir, world):
    n, I, J, V, cnt = _ap_matrix()
    x = np.sin(np.arange(n) * 0.37) + 1.5
    a = np.abs(V)
    for mode in ("ap[dp_sp_hp]", "ap[dp_sp]", "ap[sp_hp]"):
        # values as stored by partition_precisions (interface.hpp:938-964): dp / sp / hp by abs(v) against t1 = 0.5, t2 = 0.01
        if mode == "ap[dp_sp_hp]":
            Vs = np.where(a >= 0.5, V, np.where(a >= 0.01, V.astype(np.float32).astype(np.float64), V.astype(np.float16).astype(np.float64)))
            xs, tol = x, 1e-12
        elif mode == "ap[dp_sp]":
            Vs = np.where(a >= 0.5, V, V.astype(np.float32).astype(np.float64))
            xs, tol = x, 1e-12
        else:
            Vs = np.where(a >= 0.5, V.astype(np.float32).astype(np.float64), V.astype(np.float16).astype(np.float64))
            xs, tol = x.astype(np.float32).astype(np.float64), 1e-5
        y = np.concatenate([np.load(os.path.join(out_dir, f"ap{r}_{mode}.npy")) for r in range(world)])
        ref = np.zeros(n)
        np.add.at(ref, I, Vs * xs[J])
        sc = np.zeros(n)
        np.add.at(sc, I, np.abs(Vs * xs[J]))
        assert len(y) == n and np.all(np.abs(y - ref) <= tol * sc), mode
    wsa = np.load(os.path.join(out_dir, "apwsa0.npy"))
    assert wsa[0] == 0 and wsa[-1] == n and np.all(np.diff(wsa) > 0)


def test_ap_world_size_1(eng, tmp_path):
    _run_ap_rank(0, 1, 29300 + os.getpid() % 200, str(tmp_path))
    _check_ap(str(tmp_path), 1)


def test_ap_two_ranks_seg_nnz(eng, tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    mp.spawn(_run_ap_rank, args=(2, 29350 + os.getpid() % 100, str(tmp_path)), nprocs=2, join=True)
    _check_ap(str(tmp_path), 2)


def test_ap_two_ranks_on_one_gpu(eng, tmp_path):
    import torch.multiprocessing as mp
    mp.spawn(_run_ap_rank, args=(2, 29250 + os.getpid() % 100, str(tmp_path), True), nprocs=2, join=True)
    _check_ap(str(tmp_path), 2)
