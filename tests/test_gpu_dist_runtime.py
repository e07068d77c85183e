"""GPU runtime tests of the one-process-per-GPU path (dist.py + uspmv_p2p_*): world_size 1 in-process (fused kernel
with no peers) and, when the box has >= 2 GPUs, a real 2-rank NVLink run for every exchange mode."""
import importlib
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run_rank(rank, world, port, out_dir, C=32):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        pkg = importlib.import_module("ultimate-spmv_b200")
        eng, capi, d = pkg.engine, pkg.capi, pkg.dist
        n = 48
        results = {}
        for halo, mode in (("p2p", 2), ("p2p", 1), ("p2p", 0), ("nccl", True), ("nccl", False)):
            r = d.DistributedSpmv(eng.default_context(rank), 27, n, C, 64, "dp", rank, world, overlap=mode, halo=halo)
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
            results[f"{halo}{mode}"] = r.y[: r.scs.n_rows_padded][perm].cpu().numpy()
            del r
        base = results["p2p2"]
        for k, v in results.items():
            assert np.array_equal(v, base), k
        np.save(os.path.join(out_dir, f"y{rank}.npy"), base)
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


@pytest.mark.parametrize("C", [32, 16])
def test_world_size_1_fused_kernel(eng, mats, tmp_path, C):
    """C = 32: fused kernel; C = 16: the P2P step falls back to the push / wait / ack kernels around the direct SpMV kernel."""
    port = 29700 + os.getpid() % 200 + C
    _run_rank(0, 1, port, str(tmp_path), C)
    _check(str(tmp_path), 1, mats)


@pytest.mark.parametrize("C", [32, 16])
def test_two_ranks_over_nvlink(eng, mats, tmp_path, C):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    port = 29900 + os.getpid() % 100 + C
    mp.spawn(_run_rank, args=(2, port, str(tmp_path), C), nprocs=2, join=True)
    _check(str(tmp_path), 2, mats)
