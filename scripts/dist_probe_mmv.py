"""2-GPU probe (torchrun) of the distributed SpMMV step: where does the time over the local kernel go?
Per (vt, bvs): local kernel (full / interior part / boundary part), the halo exchange alone (push + wait + ack), and the step in
its three modes (0 exchange first, 1 multi-kernel overlap, 2 fused one-kernel step).  CUDA events, max over ranks."""
import importlib, os, sys, json, ctypes as C
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
pkg = importlib.import_module("ultimate-spmv_b200"); eng, capi = pkg.engine, pkg.capi
call = capi.call
vp = C.c_void_p
def timeit(fn, n=100):
    for _ in range(5): fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / n * 1e3], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)
cases = [a.split(":") for a in sys.argv[1:]] or [["dp", "4"], ["dp", "8"], ["sp", "8"]]
for case in cases:
    vt, bvs = case[0], int(case[1])
    variant = int(case[2]) if len(case) > 2 else 0   # "dp:8:17": force SpMMV kernel variant 17 (plain and fused instance)
    capi.set_option("mmv_variant", variant)
    for k, v in (("mmv_fused_rowwise", 1),):
        capi.set_option(k, v)
    r = pkg.dist.DistributedSpmv(eng.default_context(lr), 7, 256, 32, 1, vt, rank, world, overlap=2, bvs=bvs, layout="rowwise")
    r.x.uniform_(-1, 1)
    s, cs = torch.cuda.current_stream(), r.comm_stream
    X, Y = vp(r.x.data_ptr()), vp(r.y.data_ptr())
    out = {"vt": vt, "bvs": bvs, "variant": variant, "interior_chunks": r.n_interior_chunks, "boundary_chunks": r.n_boundary_chunks, "n_send": int(r.plan.n_send)}
    out["kernel_full"] = timeit(lambda: call("uspmv_spmmv", r.scs.h, X, Y, bvs, r.vec_length, r.layout, vp(s.cuda_stream)))
    out["kernel_interior"] = timeit(lambda: call("uspmv_spmmv_part", r.scs.h, 1, X, Y, bvs, r.vec_length, r.layout, vp(s.cuda_stream)))
    out["kernel_boundary"] = timeit(lambda: call("uspmv_spmmv_part", r.scs.h, 2, X, Y, bvs, r.vec_length, r.layout, vp(s.cuda_stream)))
    out["exchange_only"] = timeit(lambda: call("uspmv_p2p_exchange", r.p2p.h, 0, vp(s.cuda_stream), vp(cs.cuda_stream)))
    for mode in (0, 1, 2):
        call("uspmv_p2p_set_overlap", r.p2p.h, mode)
        out[f"step_mode{mode}"] = timeit(r.step)
    # the local kernel once more, LAST: is "interior slower than full" an effect of the order of measurement (clocks under load)?
    out["kernel_full_again"] = timeit(lambda: call("uspmv_spmmv", r.scs.h, X, Y, bvs, r.vec_length, r.layout, vp(s.cuda_stream)))
    err, ep = r.p2p.status()
    out["err"] = err
    if rank == 0:
        print(json.dumps(out), flush=True)
    r.close()
dist.barrier(); dist.destroy_process_group()
