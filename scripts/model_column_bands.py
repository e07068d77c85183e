"""CPU model for DESIGN.md section 9 item 1: how much SELL-32-sigma padding does a column-banded execution plan cost on the config-4
power-law matrix?  Rows keep the sigma-sorted order of the WHOLE matrix (one permutation, so y is accumulated in place); every band is its
own SELL-32 structure.  Prints stored slots (x nnz) for K = 1, 2, 4, 8, 16 bands and the resulting bytes per SpMV."""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
mats = importlib.import_module("ultimate-spmv_b200.matrices") if False else None
sys.path.insert(0, os.path.join(ROOT, "ultimate-spmv_b200"))
import matrices as mats  # noqa: E402  (host-only generator, no CUDA needed)

LOG2 = int(sys.argv[1]) if len(sys.argv) > 1 else 20
n = 1 << LOG2
C, SIGMA = 32, 16384
_, _, I, J, V = mats.powerlaw_coo(n, n * 15)
nnz = len(I)
cnt = np.bincount(I, minlength=n)
# sigma-window sort by total row length (descending); ties do not matter for slot counts
perm = np.concatenate([w0 + np.argsort(-cnt[w0:w0 + SIGMA], kind="stable") for w0 in range(0, n, SIGMA)])
pos = np.empty(n, np.int64); pos[perm] = np.arange(n)          # row -> position
def slots(row_counts_by_pos):
    m = row_counts_by_pos.reshape(-1, C).max(axis=1)
    return int(m.sum()) * C
base = slots(cnt[perm])
print(f"power-law 2^{LOG2}: nnz {nnz}, un-banded slots {base} (beta {nnz / base:.3f})")
near = np.abs(J.astype(np.int64) - I.astype(np.int64)) <= 1024
print(f"  entries within +-1024 of the diagonal: {near.mean():.3f}")
for K in (2, 4, 8, 16):
    w = (n + K - 1) // K
    band = J // w
    tot = 0
    for b in range(K):
        cb = np.bincount(I[band == b], minlength=n)
        tot += slots(cb[perm])
    # bytes: dp matrix (12 B/slot) + K passes over y (16 B/row) + x once
    by = tot * 12 + K * 16 * n + 8 * n
    by0 = base * 12 + 16 * n
    print(f"  K = {K:2d} column bands: slots {tot} = {tot / base:.2f} x un-banded (beta {nnz / tot:.3f}); dp bytes per SpMV {by / 1e6:.0f} MB vs {by0 / 1e6:.0f} MB "
          f"algorithmic un-banded; un-banded kernel moves ~{(base * 12 + 0.63 * nnz * 128) / 1e6:.0f} MB when x misses L2")

# adaptive precision ap[dp_sp_hp] (t1 = 1.0, t2 = 1e-2): every (band, part) is its own SELL-32 structure on the same permutation
a = np.abs(V)
part = np.where(a >= 1.0, 0, np.where(a >= 1e-2, 1, 2))
bytes_per = (12, 8, 6)
def ap_bytes(K):
    w = (n + K - 1) // K
    band = J // w
    tot_b, tot_s = 0, 0
    for b in range(K):
        for p in range(3):
            cb = np.bincount(I[(band == b) & (part == p)], minlength=n)
            s = slots(cb[perm])
            tot_s += s
            tot_b += s * bytes_per[p]
    return tot_s, tot_b + K * 16 * n + 8 * n
s1, b1 = ap_bytes(1)
print(f"ap[dp_sp_hp]: un-banded slots {s1} ({s1 / nnz:.2f} per nnz), algorithmic {b1 / 1e6:.0f} MB, moved when x misses L2 ~{(b1 + 0.63 * nnz * 128) / 1e6:.0f} MB")
for K in (4, 8):
    sk, bk = ap_bytes(K)
    print(f"  K = {K} bands: slots {sk} = {sk / s1:.2f} x, bytes per SpMV {bk / 1e6:.0f} MB")
