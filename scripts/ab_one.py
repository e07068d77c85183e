"""One precision, one library build, a few launches (for ncu): python scripts/ab_one.py <vt> [stream_variant]"""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
pkg = importlib.import_module("ultimate-spmv_b200"); eng, capi = pkg.engine, pkg.capi
vt = sys.argv[1]
if len(sys.argv) > 2: capi.set_option("stream_variant", int(sys.argv[2]))
TD = {"dp": torch.float64, "sp": torch.float32, "hp": torch.float16}
mtx = eng.MtxData.stencil(7, 256, 256, 256)
scs = eng.convert_to_scs(mtx, 32, 1, vt); eng.permute_scs_cols(scs); del mtx
x = torch.full((scs.n_rows_padded,), 0.5, dtype=TD[vt], device="cuda"); y = torch.zeros_like(x)
for _ in range(6): eng.spmv(scs, x, y)
torch.cuda.synchronize()
