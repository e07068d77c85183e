"""Config 4 (power-law 2^LOG2 rows, ap[dp_sp_hp], C = 32, sigma = 512) — a few launches of the fused AP kernel, for ncu / timing sweeps.
python scripts/ap_one.py [log2_rows] [split_long_chunks,...]"""
import importlib, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
pkg = importlib.import_module("ultimate-spmv_b200"); eng, capi = pkg.engine, pkg.capi
LOG2 = int(sys.argv[1]) if len(sys.argv) > 1 else 25
LS = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else []
n = 1 << LOG2
mode = "ap[dp_sp_hp]"
mtx = eng.MtxData.powerlaw(n)
coos = eng.partition_precisions(mtx, mode, 1.0, 1e-2)
del mtx
parts = [None] * 3
parts[0] = eng.convert_to_scs(coos[0], 32, 512, "dp")
perm = pkg.validate.device_int_tensor(parts[0].device_arrays()["old_to_new"].value, n, torch.device("cuda")).cpu().numpy()
parts[1] = eng.convert_to_scs(coos[1], 32, 512, "sp", fixed_permutation=perm)
parts[2] = eng.convert_to_scs(coos[2], 32, 512, "hp", fixed_permutation=perm)
del coos
x = torch.rand(n, dtype=torch.float64, device="cuda") + 0.5
y = torch.zeros(parts[0].n_rows_padded, dtype=torch.float64, device="cuda")
def timeit(reps=10):
    for _ in range(3): eng.ap_spmv(mode, parts[0], parts[1], parts[2], x, y)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): eng.ap_spmv(mode, parts[0], parts[1], parts[2], x, y)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
if not LS:
    for _ in range(4): eng.ap_spmv(mode, parts[0], parts[1], parts[2], x, y)
    torch.cuda.synchronize()
else:
    for L in LS:
        capi.set_option("split_long_chunks", L)
        print(json.dumps({"split_long_chunks": L, "ms": timeit()}), flush=True)
