"""One rank, no neighbour: the distributed SpMV step (the FUSED instance of k_scs32_stream, nothing to exchange) against the plain
single-GPU kernel on the same matrix — what the fused instance itself costs, per value type.
usage: python scripts/spmv_fused_probe.py [steps] [dp,sp,hp]      (USPMV_B200_LIB selects the library)"""
import importlib, os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 100
vts = sys.argv[2].split(",") if len(sys.argv) > 2 else ["dp", "sp", "hp"]
torch.cuda.set_device(0)
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:29534", rank=0, world_size=1)
pkg = importlib.import_module("ultimate-spmv_b200")
eng, capi, d = pkg.engine, pkg.capi, pkg.dist


def timeit(fn):
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps * 1e3


out = {"lib": os.path.basename(capi.LIB_PATH)}
for vt in vts:
    r = d.DistributedSpmv(eng.default_context(0), 7, 256, 32, 1, vt, 0, 1, overlap=2)
    r.x.uniform_(-1, 1)
    plain = timeit(lambda: eng.spmv(r.scs, r.x, r.y))
    y0 = r.y.clone()
    fused = timeit(r.step)
    same = bool(torch.equal(r.y.view(torch.uint8), y0.view(torch.uint8)))
    plain2 = timeit(lambda: eng.spmv(r.scs, r.x, r.y))
    out[vt] = {"plain_us": round(plain, 1), "fused_us": round(fused, 1), "plain_again_us": round(plain2, 1), "y_same": same}
    r.close()
print(json.dumps(out))
dist.destroy_process_group()
