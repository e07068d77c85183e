"""uspmv_csr_gpu on cudaMalloc'd arrays of exactly nnz elements, every nnz mod 8, against the COO sum.  Written to run under
compute-sanitizer --tool memcheck (any read past the caller's arrays by the bulk copies or the tail loads would be reported); the tool is
closed on the shared GPU pool, so there it only checks the results."""
import ctypes as C, importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
pkg = importlib.import_module("ultimate-spmv_b200"); eng, capi = pkg.engine, pkg.capi
ctx = eng.default_context(0)
vp = C.c_void_p
def dmalloc(a):
    p = vp()
    capi.call("uspmv_malloc", ctx.h, max(a.nbytes, 16), C.byref(p))
    capi.call("uspmv_memcpy_h2d", ctx.h, p, a.ctypes.data_as(vp), a.nbytes, None)
    return p
rng = np.random.default_rng(1)
worst = 0.0
for n in (5, 40, 333, 1000):
    for extra in range(8):
        cnt = rng.integers(0, 9, n); cnt[-1] += extra
        rp = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int32); nnz = int(rp[-1])
        ci = rng.integers(0, n, nnz).astype(np.int32); va = rng.uniform(-1, 1, nnz); x = rng.uniform(-1, 1, n)
        d_rp, d_ci, d_va, d_x = dmalloc(rp), dmalloc(ci), dmalloc(va), dmalloc(x)
        d_y = dmalloc(np.zeros(n))
        capi.call("uspmv_csr_gpu", ctx.h, capi.F64, n, d_rp, d_ci, d_va, d_x, d_y, None)
        capi.call("uspmv_ctx_sync", ctx.h)
        y = np.zeros(n)
        capi.call("uspmv_memcpy_d2h", ctx.h, y.ctypes.data_as(vp), d_y, y.nbytes, None)
        ref = np.zeros(n); np.add.at(ref, np.repeat(np.arange(n), cnt), va * x[ci])
        worst = max(worst, float(np.max(np.abs(y - ref))))
        for p in (d_rp, d_ci, d_va, d_x, d_y): capi.call("uspmv_free", ctx.h, p)
print("max |y - ref| =", worst)
assert worst < 1e-12
