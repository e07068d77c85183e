"""Power-law matrix (config 4 shape), a few plain-dp and ap[dp_sp_hp] SpMVs (for ncu): python scripts/one_powerlaw.py [log2 rows] [sigma]"""
import importlib, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
pkg = importlib.import_module("ultimate-spmv_b200"); eng, capi, mats = pkg.engine, pkg.capi, pkg.matrices
n = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 22); SIG = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
slab = 1 << 20
parts = [mats.powerlaw_coo(n, n * 15, row0=r0, row1=min(n, r0 + slab)) for r0 in range(0, n, slab)]
I = np.concatenate([p[2] + r0 for p, r0 in zip(parts, range(0, n, slab))]).astype(np.int32)
J = np.concatenate([p[3] for p in parts]); V = np.concatenate([p[4] for p in parts]); del parts
mtx = eng.MtxData.from_host(n, n, I, J, V); nnz = len(I); del I, J, V
scs = eng.convert_to_scs(mtx, 32, SIG, "dp")
x = torch.full((scs.n_rows_padded,), 1.0, dtype=torch.float64, device="cuda"); y = torch.zeros_like(x)
for _ in range(4): eng.spmv(scs, x, y)
torch.cuda.synchronize()
print(f"plain dp: nnz {nnz} n_elements {scs.n_elements} n_chunks {scs.n_chunks}")
del scs
mode = "ap[dp_sp_hp]"
coos = eng.partition_precisions(mtx, mode, 1.0, 1e-2)
P = [None] * 3
P[0] = eng.convert_to_scs(coos[0], 32, SIG, "dp")
perm = P[0].export().old_to_new
for k, vt in ((1, "sp"), (2, "hp")):
    P[k] = eng.convert_to_scs(coos[k], 32, SIG, vt, fixed_permutation=perm)
n_pad = P[0].n_rows_padded
x = torch.full((max(n_pad, n),), 1.0, dtype=torch.float64, device="cuda"); y = torch.zeros(n_pad, dtype=torch.float64, device="cuda")
for _ in range(4): eng.ap_spmv(mode, P[0], P[1], P[2], x, y)
torch.cuda.synchronize()
print("ap: n_elements", [p.n_elements for p in P])
