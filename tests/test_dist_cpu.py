"""N > 1 host-side logic on CPU: two gloo ranks run the comm-schedule transpose (collect_comm_idxs) and the
in-place halo exchange of ultimate-spmv_b200/dist.py on CPU tensors; the per-rank arithmetic is the oracle's, so
this checks partitioning + halo numbering + exchange plumbing end to end against the single-rank product."""
import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, method, C, sigma, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import types
        # dist.py needs capi only for the GPU classes; give it a stub so the pure host logic is importable here
        pkg_name = "ultimate-spmv_b200"
        pkg = importlib.import_module(pkg_name)
        try:
            d = importlib.import_module(pkg_name + ".dist")
        except ImportError:
            stub = types.ModuleType(pkg_name + ".capi")
            stub.call, stub.vp, stub.COLWISE, stub.lib = None, None, 0, None
            sys.modules[pkg_name + ".capi"] = stub
            d = importlib.import_module(pkg_name + ".dist")
        from oracle.bindings import Oracle
        orc = Oracle()
        n, nc, I, J, V = pkg.matrices.random_coo(1500, 6, seed=11, empty_rows=False)
        x_glob = np.random.default_rng(0).standard_normal(n)
        wsa = orc.seg_work_sharing_arr(method, n, I, world)
        sel = (I >= wsa[rank]) & (I < wsa[rank + 1])
        n_loc = int(wsa[rank + 1] - wsa[rank])
        s = orc.convert_to_scs(n_loc, n, (I[sel] - wsa[rank]).astype(np.int32), J[sel], V[sel], C, sigma)
        need, cum = orc.collect_halo(s.col_idxs, wsa, rank)
        orc.permute_scs_cols(s, s.old_to_new)
        send_lists = d.comm_schedule(need, rank, world)
        send_ptr = np.cumsum([0] + [len(a) for a in send_lists])
        n_halo = int(cum[-1])
        x = np.zeros(n_loc + max(s.n_rows_padded - n_loc, n_halo))
        x[s.old_to_new] = x_glob[wsa[rank]:wsa[rank + 1]]
        # pack_send_buf: buf = x[perm[send_idx]]  (classes_structs.hpp:813-831)
        flat = np.concatenate(send_lists) if send_ptr[-1] else np.zeros(0, np.int32)
        sendbuf = torch.from_numpy(x[s.old_to_new[flat]].copy()) if len(flat) else torch.zeros(1, dtype=torch.float64)
        xt = torch.from_numpy(x)
        ex = d.HaloExchange(rank, world, n_loc, cum, send_ptr)
        d.HaloExchange.finish(ex.begin(xt, sendbuf))
        y = orc.spmv_scs(s, xt.numpy())[s.old_to_new]
        np.save(os.path.join(out_dir, f"y{rank}.npy"), y)
        np.save(os.path.join(out_dir, f"wsa{rank}.npy"), wsa)
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("method,C,sigma", [("seg-rows", 8, 16), ("seg-nnz", 32, 64)])
def test_two_rank_gloo_halo_exchange(tmp_path, method, C, sigma):
    import torch.multiprocessing as mp
    world = 2
    port = 29500 + (os.getpid() % 2000) + (0 if method == "seg-rows" else 1)
    mp.spawn(_worker, args=(world, port, method, C, sigma, str(tmp_path)), nprocs=world, join=True)
    sys.path.insert(0, ROOT)
    pkg = importlib.import_module("ultimate-spmv_b200")
    n, nc, I, J, V = pkg.matrices.random_coo(1500, 6, seed=11, empty_rows=False)
    x_glob = np.random.default_rng(0).standard_normal(n)
    y = np.concatenate([np.load(tmp_path / f"y{r}.npy") for r in range(world)])
    y_coo = np.zeros(n)
    np.add.at(y_coo, I, V * x_glob[J])
    scale = np.zeros(n)
    np.add.at(scale, I, np.abs(V * x_glob[J]))
    assert len(y) == n
    assert np.all(np.abs(y - y_coo) <= 1e-12 * np.maximum(scale, 1e-300))


def test_seg_nnz_from_row_counts_matches_the_oracle():
    """dist.seg_nnz_from_row_counts (used when no rank holds the whole COO) == seg_work_sharing_arr on the row array."""
    sys.path.insert(0, ROOT)
    pkg = importlib.import_module("ultimate-spmv_b200")
    from oracle.bindings import Oracle
    orc = Oracle()
    rng = np.random.default_rng(0)
    for _ in range(200):
        n = int(rng.integers(5, 400))
        P = int(rng.integers(1, min(n, 9)))
        cnt = rng.integers(0, 9, n)
        cnt[rng.integers(0, n)] += int(rng.integers(1, 50))
        I = np.repeat(np.arange(n), cnt).astype(np.int32)
        assert np.array_equal(pkg.dist.seg_nnz_from_row_counts(cnt, P), orc.seg_work_sharing_arr("seg-nnz", n, I, P)), (n, P)
    # the reference's own probe (SURVEY.md section 8 a'-11): bcsstk13, P = 4 -> 0 724 1183 1600 2003
    z = np.load(os.path.join(ROOT, "tests", "golden", "matrices.npz"))
    cnt = np.bincount(z["bcsstk13__I"], minlength=int(z["bcsstk13__n"]))
    assert pkg.dist.seg_nnz_from_row_counts(cnt, 4).tolist() == [0, 724, 1183, 1600, 2003]
