"""Runs ONE kernel configuration a few times (for ncu): python scripts/one_case.py spmmv dp 8 rowwise [grid]"""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
pkg = importlib.import_module("ultimate-spmv_b200"); eng = pkg.engine
kind, vt, bvs, layout = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4]
N = int(sys.argv[5]) if len(sys.argv) > 5 else 256
C = int(sys.argv[6]) if len(sys.argv) > 6 else 32
TD = {"dp": torch.float64, "sp": torch.float32, "hp": torch.float16}
mtx = eng.MtxData.stencil(7, N, N, N)
scs = eng.convert_to_scs(mtx, C, 1, vt); eng.permute_scs_cols(scs); del mtx
ld = scs.n_rows_padded
X = torch.full((ld * bvs,), 1.0, dtype=TD[vt], device="cuda"); Y = torch.zeros_like(X)
for _ in range(4):
    if kind == "spmmv": eng.spmmv(scs, X, Y, bvs, ld, layout)
    else: eng.spmv(scs, X, Y)
torch.cuda.synchronize()
print("done", kind, vt, bvs, layout, N, C)
