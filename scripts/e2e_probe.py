"""N >= 2 GPUs (torchrun): host-buffer (e2e) distributed SpMV — one step at a time vs pipelined, with and without binding the pinned
staging buffers to the GPU's NUMA node.  Prints the PCIe / NUMA topology first."""
import importlib, json, os, subprocess, sys
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
pkg = importlib.import_module("ultimate-spmv_b200"); eng, d = pkg.engine, pkg.dist
if rank == 0:
    print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout, flush=True)
    print(subprocess.run("lscpu | grep -i -E 'numa|model name|socket|^CPU\\(s\\)'", shell=True, capture_output=True, text=True).stdout, flush=True)
full = os.sched_getaffinity(0)
r = d.DistributedSpmv(eng.default_context(lr), 7, 256, 32, 1, "dp", rank, world, halo="p2p", n_buf=2)
def barrier():
    dist.barrier(); torch.cuda.synchronize()
for _ in range(5): r.step()
barrier()
y_ref = r.y.clone()
res = {}
nnz = torch.tensor([float(r.nnz)], device="cuda", dtype=torch.float64); dist.all_reduce(nnz)
for bind in (False, True):
    os.sched_setaffinity(0, full)
    b = d.bind_to_gpu_numa_node(lr) if bind else None
    info = [None] * world
    dist.all_gather_object(info, None if b is None else (b["node"], b["cpus"], b["bdf"]))
    if rank == 0 and bind: print("numa binding per rank:", info, flush=True)
    for pipe in (False, True):
        sec = r.time_e2e(10, barrier, pipelined=pipe)
        t = torch.tensor([sec], device="cuda", dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX)
        same = bool(torch.equal(r.e2e_y[: r.n_rows].cuda(), y_ref[: r.n_rows]))
        ok = torch.tensor([int(same)], device="cuda"); dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        res[f"bind{int(bind)}_pipe{int(pipe)}"] = {"ms_per_step": float(t.item()) * 1e3, "gflops": 2 * float(nnz.item()) / float(t.item()) / 1e9, "y_same": bool(ok.item())}
        if rank == 0: print(f"bind={bind} pipelined={pipe}: {res[f'bind{int(bind)}_pipe{int(pipe)}']}", flush=True)
os.sched_setaffinity(0, full)
err, ep = r.p2p.status()
if rank == 0:
    print("p2p status", err, ep)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", f"r01t_e2e_probe_n{world}.json"), "w"), indent=1)
dist.barrier(); dist.destroy_process_group()
