"""One rank, no neighbour: the distributed SpMMV step through (a) the plain streamed kernel and (b) the FUSED instance of the same kernel
(option mmv_fused_rowwise = 2 forces it although nothing is exchanged) — isolates what the fused instance itself costs.
usage: python scripts/mmv_fused_probe.py [vt] [bvs] [steps] [mmv_variant]       (run it under ncu -k regex:stream_mmv for the counters)"""
import importlib, os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

vt = sys.argv[1] if len(sys.argv) > 1 else "sp"
bvs = int(sys.argv[2]) if len(sys.argv) > 2 else 8
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 50
variant = int(sys.argv[4]) if len(sys.argv) > 4 else 0
torch.cuda.set_device(0)
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:29533", rank=0, world_size=1)
pkg = importlib.import_module("ultimate-spmv_b200")
eng, capi, d = pkg.engine, pkg.capi, pkg.dist
r = d.DistributedSpmv(eng.default_context(0), 7, 256, 32, 1, vt, 0, 1, overlap=2, bvs=bvs, layout="rowwise")
r.x.uniform_(-1, 1)
capi.set_option("mmv_variant", variant)
out = {}
for name, force in (("plain", 0), ("fused", 2), ("plain_again", 0)):
    capi.set_option("mmv_fused_rowwise", force)
    for _ in range(5):
        r.step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        r.step()
    e1.record()
    torch.cuda.synchronize()
    out[name] = e0.elapsed_time(e1) / steps
    out[name + "_ysum"] = float(r.y.double().sum())
print(json.dumps({"vt": vt, "bvs": bvs, "variant": variant, "ms_per_step": out}))
r.close()
dist.destroy_process_group()
