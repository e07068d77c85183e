"""2+ GPU probe (torchrun): times interior / boundary / full / p2p step pieces with CUDA events."""
import importlib, os, sys, ctypes as C
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
pkg = importlib.import_module("ultimate-spmv_b200"); eng, capi = pkg.engine, pkg.capi
r = pkg.dist.DistributedSpmv(eng.default_context(lr), 7, 256, 32, 1, "dp", rank, world, halo="p2p")
vp = C.c_void_p
def timeit(fn, n=100):
    for _ in range(5): fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
s = torch.cuda.current_stream()
full = timeit(lambda: eng.spmv(r.scs, r.x, r.y))
inte = timeit(lambda: capi.call("uspmv_spmv_part", r.scs.h, 1, vp(r.x.data_ptr()), vp(r.y.data_ptr()), vp(s.cuda_stream)))
bnd = timeit(lambda: capi.call("uspmv_spmv_part", r.scs.h, 2, vp(r.x.data_ptr()), vp(r.y.data_ptr()), vp(s.cuda_stream)))
y_ref = r.y.clone()
res = {}
for mode in (2, 1, 0):
    capi.call("uspmv_p2p_set_overlap", r.p2p.h, mode)
    r.y.zero_()
    res[mode] = timeit(r.step)
    res[f"same{mode}"] = bool(torch.equal(r.y, y_ref))
err, ep = r.p2p.status()
print(f"rank {rank}: full {full:.1f} us | interior {inte:.1f} ({r.n_interior_chunks} chunks) | boundary {bnd:.1f} ({r.n_boundary_chunks} chunks) | p2p step fused {res[2]:.1f} | multi-kernel overlap {res[1]:.1f} | no-overlap {res[0]:.1f} | y same {res['same2']},{res['same1']},{res['same0']} | err {err} epoch {ep}", flush=True)
dist.barrier(); dist.destroy_process_group()
