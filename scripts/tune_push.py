"""N >= 2 GPUs (torchrun): the large-halo push kernels of the NVLink exchange on the power-law matrix (BASELINE config 4).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
        scripts/tune_push.py --log2 23 --out gpurun_out/tune_push_n2.json

Per push variant (0 = one store per thread, 1 = tiled direct stores, 2 = tiled shared memory + bulk copy) and CTAs per SM:
exchange alone and the full distributed AP step, CUDA events, max over ranks; and a known-answer check of the received halo."""
import argparse
import ctypes as C
import importlib
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ap = argparse.ArgumentParser()
ap.add_argument("--log2", type=int, default=23)
ap.add_argument("--mode", default="ap[dp_sp_hp]")
ap.add_argument("--steps", type=int, default=30)
ap.add_argument("--out", default=None)
ap.add_argument("--cases", default="0:0,1:2,1:4,1:8,2:2,2:4,2:8", help="push_variant:ctas_per_sm, comma separated")
args = ap.parse_args()
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
pkg = importlib.import_module("ultimate-spmv_b200")
eng, capi = pkg.engine, pkg.capi
import bench  # noqa: E402

r, wsa = bench.build_powerlaw_ap(pkg, eng.default_context(lr), args.log2, args.mode, 32, 16384, rank, world)
vp = C.c_void_p
main = torch.cuda.current_stream()


opt = capi.set_option


def exchange():
    capi.call("uspmv_p2p_exchange", r.p2p.h, 0, vp(main.cuda_stream), vp(r.comm_stream.cuda_stream))


def timeit(fn, n):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / n], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


# known answer: x_local[i] = global row id; the halo tail must then hold the global column ids of the need lists
n_loc = r.n_rows
expect = np.concatenate([wsa[p] + r.plan.need_lists[p].astype(np.int64) for p in range(world)]).astype(np.float64)
res = {"world": world, "log2_rows": args.log2, "mode": args.mode, "n_local": n_loc, "n_halo": r.n_halo,
       "n_send": int(r.plan.send_ptr[-1]) if r.plan.send_ptr is not None else None, "cases": []}
r.time_kernel(3)
kt = torch.tensor([r.time_kernel(args.steps)], device="cuda")
kts = [torch.zeros_like(kt) for _ in range(world)]
dist.all_gather(kts, kt)
res["local_kernel_ms_per_rank"] = [float(t.item()) for t in kts]
for variant, per_sm in [tuple(int(v) for v in c.split(":")) for c in args.cases.split(",")]:
    opt("push_variant", variant)
    opt("push_ctas_per_sm", per_sm)
    r.x.zero_()
    r.x[:n_loc] = torch.arange(int(wsa[rank]), int(wsa[rank + 1]), device="cuda").to(r.x.dtype)
    torch.cuda.synchronize()
    dist.barrier()
    exchange()
    torch.cuda.synchronize()
    dist.barrier()
    got = r.x[n_loc:n_loc + r.n_halo].cpu().numpy().astype(np.float64)
    bad = np.flatnonzero(got != expect)
    ok = len(bad) == 0
    okt = torch.tensor([len(bad), int(bad[0]) if len(bad) else -1, int(bad[-1]) if len(bad) else -1], device="cuda")
    okl = [torch.zeros_like(okt) for _ in range(world)]
    dist.all_gather(okl, okt)
    if rank == 0 and any(int(o[0]) for o in okl):
        print("MISMATCH variant", variant, per_sm, "per rank [count, first, last]:", [o.tolist() for o in okl],
              "recv_cumsum", r.plan.recv_cumsum.tolist(), flush=True)
    if not ok:
        b = int(bad[0])
        print(f"rank {rank} variant {variant}: first bad halo slot {b}: got {got[b:b+4]} expect {expect[b:b+4]}", flush=True)
    dist.barrier()  # nobody may start the next exchange (it overwrites the halo) before every rank has checked its own
    r.x.fill_(1.0)
    t_ex = timeit(exchange, args.steps)
    t_step = timeit(r.step, args.steps)
    err, ep = r.p2p.status()
    nnz = torch.tensor([float(r.nnz)], device="cuda", dtype=torch.float64)
    dist.all_reduce(nnz)
    case = {"gflops_step": 2 * float(nnz.item()) / (t_step * 1e-3) / 1e9, "push_variant": variant, "ctas_per_sm": per_sm, "halo_ok": ok, "exchange_ms": t_ex, "step_ms": t_step, "err": err,
            "push_gbs_out": r.plan.send_ptr[-1] * r.x.element_size() / (t_ex * 1e-3) / 1e9}
    res["cases"].append(case)
    if rank == 0:
        print(json.dumps(case), flush=True)
oks = torch.tensor([int(all(c["halo_ok"] for c in res["cases"]))], device="cuda")
dist.all_reduce(oks, op=dist.ReduceOp.MIN)
res["halo_ok_all_ranks"] = bool(oks.item())
if rank == 0:
    print(json.dumps({k: v for k, v in res.items() if k != "cases"}), flush=True)
    if args.out:
        with open(args.out, "w") as f:
            json.dump(res, f, indent=1)
dist.barrier()
dist.destroy_process_group()
