"""TEST INFRASTRUCTURE ONLY: numpy/ctypes bindings for oracle/liboracle.so (the C restatement) and
oracle/_ref/*.so (the unmodified reference compiled by oracle/Makefile).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from types import SimpleNamespace

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
VT = {"dp": 0, "sp": 1, "hp": 2, np.float64: 0, np.float32: 1, np.float16: 2}
NPT = {0: np.float64, 1: np.float32, 2: np.float16}
MODE = {"ap[dp_sp]": 0, "ap[dp_hp]": 1, "ap[sp_hp]": 2, "ap[dp_sp_hp]": 3}

_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")


def _vt(v):
    if isinstance(v, str):
        return VT[v]
    return VT[np.dtype(v).type]


def _p(a):
    """void* of a numpy array (or NULL)."""
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def build(which="all"):
    subprocess.run(["make", "-s", "-C", HERE, which], check=True)


# ------------------------------------------------------------------------------------------------
# C restatement
# ------------------------------------------------------------------------------------------------
class Oracle:
    def __init__(self):
        path = os.path.join(HERE, "liboracle.so")
        if not os.path.exists(path):
            build("oracle")
        self.lib = L = C.CDLL(path)
        L.orc_scs_structure.restype = C.c_long
        L.orc_collect_halo.restype = C.c_long
        L.orc_stencil_coo.restype = C.c_long
        L.orc_ingest_entries.restype = C.c_long

    def stencil_coo(self, points, nx, ny, nz, row0=0, row1=None):
        row1 = nx * ny * nz if row1 is None else row1
        args = (C.c_int(points), C.c_long(nx), C.c_long(ny), C.c_long(nz), C.c_long(row0), C.c_long(row1))
        nnz = self.lib.orc_stencil_coo(*args, None, None, None)
        I, J, V = np.empty(nnz, np.int32), np.empty(nnz, np.int32), np.empty(nnz, np.float64)
        self.lib.orc_stencil_coo(*args, _p(I), _p(J), _p(V))
        return row1 - row0, nx * ny * nz, I, J, V

    # -- std::sort restatement ---------------------------------------------------------------
    def sort_window(self, cnt):
        cnt = np.ascontiguousarray(cnt, dtype=np.int64).copy()
        idx = np.arange(len(cnt), dtype=np.int64)
        self.lib.orc_sort_window(_p(idx), _p(cnt), C.c_long(len(cnt)))
        return idx, cnt

    # -- convert_to_scs -----------------------------------------------------------------------
    def convert_to_scs(self, n_rows, n_cols, I, J, vals, Cc, sigma, vt="dp", fixed_perm=None):
        I, J = _i32(I), _i32(J)
        vals = np.ascontiguousarray(vals, dtype=np.float64)
        nnz = len(I)
        n_chunks = (n_rows + Cc - 1) // Cc
        n_pad = n_chunks * Cc
        cp = np.zeros(n_chunks + 1, np.int32)
        cl = np.zeros(n_chunks, np.int32)
        o2n = np.zeros(n_rows, np.int32)
        n2o = np.zeros(n_pad, np.int32)
        fp = None if fixed_perm is None else _i32(fixed_perm)
        ne = self.lib.orc_scs_structure(C.c_long(n_rows), C.c_long(nnz), _p(I), C.c_long(Cc), C.c_long(sigma), _p(fp),
                                        _p(cp), _p(cl), _p(o2n), _p(n2o))
        if ne == -2:
            raise ValueError("fixed_permutation maps a non-empty row onto a padding position (undefined behaviour in the reference)")
        if ne < 0:
            raise OverflowError("chunk_ptrs exceed int32")
        vtc = _vt(vt)
        ci = np.zeros(ne, np.int32)
        v = np.zeros(ne, NPT[vtc])
        row_to_pos = fp if fp is not None else o2n
        self.lib.orc_scs_fill(C.c_long(nnz), _p(I), _p(J), _p(vals), C.c_long(Cc), C.c_long(n_pad), _p(cp), _p(row_to_pos),
                              C.c_int(vtc), C.c_long(ne), _p(ci), _p(v))
        if fp is not None:  # the struct's own permutation is the identity in fixed_permutation mode
            o2n = np.arange(n_rows, dtype=np.int32)
            n2o = np.full(n_pad, -1, np.int32)
            n2o[:n_rows] = o2n
        return SimpleNamespace(C=Cc, sigma=sigma, n_rows=n_rows, n_cols=n_cols, n_rows_padded=n_pad, n_chunks=n_chunks,
                               n_elements=int(ne), nnz=nnz, chunk_ptrs=cp, chunk_lengths=cl, col_idxs=ci, values=v,
                               old_to_new=o2n, new_to_old=n2o)

    def permute_scs_cols(self, scs, perm):
        self.lib.orc_permute_scs_cols(C.c_long(scs.n_elements), C.c_long(scs.n_rows), _p(scs.col_idxs), _p(_i32(perm)))

    # -- kernels ------------------------------------------------------------------------------
    def spmv_scs(self, scs, x):
        vtc = _vt(scs.values.dtype)
        x = np.ascontiguousarray(x, dtype=NPT[vtc])
        y = np.zeros(scs.n_rows_padded, NPT[vtc])
        self.lib.orc_spmv_scs(C.c_int(vtc), C.c_long(scs.C), C.c_long(scs.n_chunks), _p(scs.chunk_ptrs), _p(scs.chunk_lengths),
                              _p(scs.col_idxs), _p(scs.values), _p(x), _p(y))
        return y

    def spmv_csr(self, n_rows, rp, ci, vals, x):
        vtc = _vt(vals.dtype)
        x = np.ascontiguousarray(x, dtype=NPT[vtc])
        y = np.zeros(n_rows, NPT[vtc])
        self.lib.orc_spmv_csr(C.c_int(vtc), C.c_long(n_rows), _p(_i32(rp)), _p(_i32(ci)), _p(vals), _p(x), _p(y))
        return y

    def spmmv_scs(self, scs, X, bvs, vec_length, layout):
        """X flat; layout 0 colwise (X[col + v*vec_length]) / 1 rowwise (X[col*bvs + v])."""
        vtc = _vt(scs.values.dtype)
        X = np.ascontiguousarray(X, dtype=NPT[vtc])
        Y = np.zeros(len(X), NPT[vtc])
        self.lib.orc_spmmv_scs(C.c_int(vtc), C.c_long(scs.C), C.c_long(scs.n_chunks), _p(scs.chunk_ptrs), _p(scs.chunk_lengths),
                               _p(scs.col_idxs), _p(scs.values), _p(X), _p(Y), C.c_int(bvs), C.c_long(vec_length), C.c_int(layout))
        return Y

    def ap_scs(self, mode, dp, sp, hp, dp_x, sp_x):
        m = MODE[mode] if isinstance(mode, str) else mode
        first = dp if dp is not None else sp
        y = np.zeros(first.n_rows_padded, np.float32 if m == 2 else np.float64)

        def parts(s):
            if s is None:
                return (None, None, None, None)
            return (_p(s.chunk_ptrs), _p(s.chunk_lengths), _p(s.col_idxs), _p(s.values))
        dpx = None if dp_x is None else np.ascontiguousarray(dp_x, np.float64)
        spx = None if sp_x is None else np.ascontiguousarray(sp_x, np.float32)
        self.lib.orc_ap_scs(C.c_int(m), C.c_long(first.C), C.c_long(first.n_chunks), *parts(dp), *parts(sp), *parts(hp),
                            _p(dpx), _p(spx), _p(y))
        return y

    # -- partitioning -------------------------------------------------------------------------
    def partition_precisions(self, mode, I, J, vals, t1, t2=0.0, rowmax=None, colmax=None):
        m = MODE[mode] if isinstance(mode, str) else mode
        I, J = _i32(I), _i32(J)
        vals = np.ascontiguousarray(vals, np.float64)
        part = np.zeros(len(I), np.int8)
        counts = np.zeros(3, np.int64)
        self.lib.orc_partition_precisions(C.c_int(m), C.c_long(len(I)), _p(I), _p(J), _p(vals), C.c_double(t1), C.c_double(t2),
                                          _p(rowmax), _p(colmax), _p(part), _p(counts))
        return part, counts

    def largest_elems(self, n_rows, n_cols, I, J, vals):
        I, J = _i32(I), _i32(J)
        vals = np.ascontiguousarray(vals, np.float64)
        rm, cm = np.zeros(n_rows), np.zeros(n_cols)
        self.lib.orc_largest_elems(C.c_long(len(I)), _p(I), _p(J), _p(vals), _p(rm), _p(cm))
        return rm, cm

    def equilibrate(self, n_rows, n_cols, I, J, vals):
        """equilibrate_matrix (utilities.hpp:2668-2684).  Returns (scaled values, rowmax, colmax)."""
        I, J = _i32(I), _i32(J)
        v = np.ascontiguousarray(vals, np.float64).copy()
        rm, cm = np.zeros(n_rows), np.zeros(n_cols)
        self.lib.orc_equilibrate(C.c_long(len(I)), _p(I), _p(J), _p(v), _p(rm), _p(cm))
        return v, rm, cm

    def ingest_entries(self, n_rows, I, J, vals, symmetric):
        """read_mtx post-processing: symmetric expansion + stable row sort of the file-order entries."""
        I, J = _i32(I), _i32(J)
        vals = np.ascontiguousarray(vals, np.float64)
        nz = len(I)
        Io, Jo, Vo = np.zeros(2 * nz + 1, np.int32), np.zeros(2 * nz + 1, np.int32), np.zeros(2 * nz + 1)
        n = self.lib.orc_ingest_entries(C.c_long(nz), _p(I), _p(J), _p(vals), C.c_int(1 if symmetric else 0), C.c_int(n_rows), _p(Io), _p(Jo), _p(Vo))
        return Io[:n].copy(), Jo[:n].copy(), Vo[:n].copy()

    def seg_work_sharing_arr(self, method, n_rows, I, P):
        I = _i32(I)
        wsa = np.zeros(P + 1, np.int32)
        self.lib.orc_seg_work_sharing_arr(C.c_int(1 if method in (1, "seg-nnz", "seg_nnz") else 0), C.c_long(n_rows), C.c_long(len(I)),
                                          _p(I), C.c_int(P), _p(wsa))
        return wsa

    def collect_halo(self, col_idxs, wsa, rank):
        """Rewrites col_idxs in place.  Returns (need lists per owner, recv_counts_cumsum)."""
        P = len(wsa) - 1
        wsa = _i32(wsa)
        cap = max(1, int(wsa[-1]))
        flat = np.zeros(cap, np.int32)
        ptr = np.zeros(P + 1, np.int32)
        cum = np.zeros(P + 1, np.int32)
        n = self.lib.orc_collect_halo(C.c_long(len(col_idxs)), _p(col_idxs), _p(wsa), C.c_int(rank), C.c_int(P), _p(flat),
                                      C.c_long(cap), _p(ptr), _p(cum))
        assert n >= 0
        return [flat[ptr[p]:ptr[p + 1]].copy() for p in range(P)], cum


# ------------------------------------------------------------------------------------------------
# the real reference (only where oracle/_ref/*.so exist)
# ------------------------------------------------------------------------------------------------
def ref_available():
    return all(os.path.exists(os.path.join(HERE, "_ref", f)) for f in
               ("libuspmv_ref_col.so", "libuspmv_ref_row.so", "libuspmv_ref_iface.so"))


class Ref:
    """layout: 'col' or 'row' — the block-vector layout is a compile-time switch in the reference."""

    def __init__(self, layout="col"):
        self.lib = L = C.CDLL(os.path.join(HERE, "_ref", f"libuspmv_ref_{layout}.so"))
        L.ref_scs_build.restype = C.c_void_p
        L.ref_scs_collect_halo.restype = C.c_long
        L.ref_partition_dpsp.restype = C.c_long
        L.ref_read_mtx.restype = C.c_long
        L.ref_seg_mtx.restype = C.c_long
        self.layout = L.ref_layout()

    def omp_threads(self):
        return int(self.lib.ref_omp_max_threads())

    def build_handle(self, n_rows, n_cols, I, J, vals, Cc, sigma, vt="dp", fixed_perm=None):
        I, J = _i32(I), _i32(J)
        vals = np.ascontiguousarray(vals, np.float64)
        fp = None if fixed_perm is None else _i32(fixed_perm)
        h = self.lib.ref_scs_build(C.c_int(_vt(vt)), C.c_long(n_rows), C.c_long(n_cols), C.c_long(len(I)), _p(I), _p(J), _p(vals),
                                   C.c_long(Cc), C.c_long(sigma), _p(fp))
        assert h
        return C.c_void_p(h)

    def export(self, h, vt="dp"):
        d = np.zeros(8, np.int64)
        self.lib.ref_scs_dims(h, _p(d))
        Cc, sigma, n_rows, n_cols, n_pad, n_chunks, ne, nnz = (int(v) for v in d)
        cp = np.zeros(n_chunks + 1, np.int32)
        cl = np.zeros(n_chunks, np.int32)
        ci = np.zeros(ne, np.int32)
        v = np.zeros(ne, NPT[_vt(vt)])
        o2n = np.zeros(n_rows, np.int32)
        n2o = np.zeros(n_rows, np.int32)
        self.lib.ref_scs_copy(h, _p(cp), _p(cl), _p(ci), _p(v), _p(o2n), _p(n2o))
        return SimpleNamespace(C=Cc, sigma=sigma, n_rows=n_rows, n_cols=n_cols, n_rows_padded=n_pad, n_chunks=n_chunks,
                               n_elements=ne, nnz=nnz, chunk_ptrs=cp, chunk_lengths=cl, col_idxs=ci, values=v,
                               old_to_new=o2n, new_to_old=n2o)

    def convert_to_scs(self, n_rows, n_cols, I, J, vals, Cc, sigma, vt="dp", fixed_perm=None, permute_cols=False):
        h = self.build_handle(n_rows, n_cols, I, J, vals, Cc, sigma, vt, fixed_perm)
        if permute_cols:
            s0 = self.export(h, vt)
            self.lib.ref_scs_permute_cols(h, _p(s0.old_to_new))
        s = self.export(h, vt)
        self.lib.ref_scs_free(h)
        return s

    def collect_halo(self, h, wsa, rank):
        P = len(wsa) - 1
        wsa = _i32(wsa)
        cap = max(1, int(wsa[-1]))
        flat = np.zeros(cap, np.int32)
        ptr = np.zeros(P + 1, np.int32)
        cum = np.zeros(P + 1, np.int32)
        n = self.lib.ref_scs_collect_halo(h, _p(wsa), C.c_int(rank), C.c_int(P), _p(flat), C.c_long(cap), _p(ptr), _p(cum))
        assert 0 <= n <= cap
        return [flat[ptr[p]:ptr[p + 1]].copy() for p in range(P)], cum

    def spmv_scs(self, scs, x, adv=False):
        vtc = _vt(scs.values.dtype)
        x = np.ascontiguousarray(x, NPT[vtc]).copy()
        y = np.zeros(scs.n_rows_padded, NPT[vtc])
        self.lib.ref_spmv_scs(C.c_int(vtc), C.c_int(int(adv)), C.c_long(scs.C), C.c_long(scs.n_chunks), _p(scs.chunk_ptrs),
                              _p(scs.chunk_lengths), _p(scs.col_idxs), _p(scs.values), _p(x), _p(y))
        return y

    def spmv_scs_raw(self, vtc, adv, Cc, n_chunks, cp, cl, ci, v, x, y):
        """No copies: used by the timed CPU baseline."""
        self.lib.ref_spmv_scs(C.c_int(vtc), C.c_int(int(adv)), C.c_long(Cc), C.c_long(n_chunks), _p(cp), _p(cl), _p(ci), _p(v), _p(x), _p(y))

    def spmv_csr(self, n_rows, rp, ci, vals, x):
        vtc = _vt(vals.dtype)
        x = np.ascontiguousarray(x, NPT[vtc]).copy()
        y = np.zeros(n_rows, NPT[vtc])
        self.lib.ref_spmv_csr(C.c_int(vtc), C.c_long(n_rows), _p(_i32(rp)), _p(_i32(ci)), _p(vals), _p(x), _p(y))
        return y

    def spmmv_scs(self, scs, X, bvs, vec_length):
        vtc = _vt(scs.values.dtype)
        X = np.ascontiguousarray(X, NPT[vtc]).copy()
        Y = np.zeros(len(X), NPT[vtc])
        self.lib.ref_spmmv_scs(C.c_int(vtc), C.c_long(scs.C), C.c_long(scs.n_chunks), _p(scs.chunk_ptrs), _p(scs.chunk_lengths),
                               _p(scs.col_idxs), _p(scs.values), _p(X), _p(Y), C.c_int(bvs), C.c_int(vec_length))
        return Y

    def ap_scs_dpsp(self, dp, sp, dp_x, sp_x):
        y = np.zeros(dp.n_rows_padded, np.float64)
        dpx = np.ascontiguousarray(dp_x, np.float64).copy()
        spx = np.ascontiguousarray(sp_x, np.float32).copy()
        self.lib.ref_ap_scs_dpsp(C.c_long(dp.C), C.c_long(dp.n_chunks), _p(dp.chunk_ptrs), _p(dp.chunk_lengths), _p(dp.col_idxs), _p(dp.values),
                                 _p(sp.chunk_ptrs), _p(sp.chunk_lengths), _p(sp.col_idxs), _p(sp.values), _p(dpx), _p(spx), _p(y))
        return y

    def partition_dpsp(self, n_rows, n_cols, I, J, vals, t1):
        I, J = _i32(I), _i32(J)
        vals = np.ascontiguousarray(vals, np.float64)
        n = len(I)
        dI, dJ, dV = np.zeros(n, np.int32), np.zeros(n, np.int32), np.zeros(n, np.float64)
        sI, sJ, sV = np.zeros(n, np.int32), np.zeros(n, np.int32), np.zeros(n, np.float32)
        nsp = C.c_long(0)
        nd = self.lib.ref_partition_dpsp(C.c_long(n_rows), C.c_long(n_cols), C.c_long(n), _p(I), _p(J), _p(vals), C.c_double(t1),
                                         _p(dI), _p(dJ), _p(dV), _p(sI), _p(sJ), _p(sV), C.byref(nsp))
        ns = nsp.value
        return (dI[:nd], dJ[:nd], dV[:nd]), (sI[:ns], sJ[:ns], sV[:ns])

    def equilibrate(self, n_rows, n_cols, I, J, vals):
        I, J = _i32(I), _i32(J)
        v = np.ascontiguousarray(vals, np.float64).copy()
        self.lib.ref_equilibrate(C.c_long(n_rows), C.c_long(n_cols), C.c_long(len(I)), _p(I), _p(J), _p(v))
        return v

    def read_mtx(self, path):
        nr, nc = C.c_long(0), C.c_long(0)
        nnz = self.lib.ref_read_mtx(path.encode(), C.byref(nr), C.byref(nc), None, None, None)
        I, J, V = np.zeros(nnz, np.int32), np.zeros(nnz, np.int32), np.zeros(nnz, np.float64)
        self.lib.ref_read_mtx(path.encode(), C.byref(nr), C.byref(nc), _p(I), _p(J), _p(V))
        return nr.value, nc.value, I, J, V

    def seg_work_sharing_arr(self, method, n_rows, I, P):
        I = _i32(I)
        wsa = np.zeros(P + 2, np.int32)
        self.lib.ref_seg_work_sharing_arr(C.c_int(1 if method in (1, "seg-nnz", "seg_nnz") else 0), C.c_long(n_rows), C.c_long(len(I)),
                                          _p(I), C.c_int(P), _p(wsa))
        return wsa[:P + 1].copy()

    def random_x(self, vt, vmin, vmax, n_x, n_rows, n_rows_padded, bvs=1):
        """init_std_vec_with_ptr_or_value(..., '1') of this library's block-vector layout (utilities.hpp:914-981)."""
        x = np.zeros(n_x, NPT[_vt(vt)])
        self.lib.ref_random_x(C.c_int(_vt(vt)), C.c_double(vmin), C.c_double(vmax), C.c_long(n_x), C.c_int(n_rows), C.c_int(n_rows_padded),
                              C.c_int(bvs), _p(x))
        return x

    def seg_mtx(self, n_rows, I, J, vals, wsa, rank):
        I, J = _i32(I), _i32(J)
        vals = np.ascontiguousarray(vals, np.float64)
        n = len(I)
        lI, lJ, lV = np.zeros(n, np.int32), np.zeros(n, np.int32), np.zeros(n, np.float64)
        k = self.lib.ref_seg_mtx(C.c_long(n_rows), C.c_long(n), _p(I), _p(J), _p(vals), _p(_i32(wsa)), C.c_int(rank), _p(lI), _p(lJ), _p(lV))
        return lI[:k].copy(), lJ[:k].copy(), lV[:k].copy()


class RefIface:
    """The reference's library header (interface.hpp): AP kernels including hp."""

    def __init__(self):
        self.lib = L = C.CDLL(os.path.join(HERE, "_ref", "libuspmv_ref_iface.so"))
        L.iface_convert_f64.restype = C.c_long

    def ap_scs(self, mode, dp, sp, hp, dp_x, sp_x):
        m = MODE[mode] if isinstance(mode, str) else mode
        first = dp if dp is not None else sp
        y = np.zeros(first.n_rows_padded, np.float32 if m == 2 else np.float64)

        def parts(s):
            if s is None:
                return (None, None, None, None)
            return (_p(s.chunk_ptrs), _p(s.chunk_lengths), _p(s.col_idxs), _p(s.values))
        dpx = None if dp_x is None else np.ascontiguousarray(dp_x, np.float64).copy()
        spx = None if sp_x is None else np.ascontiguousarray(sp_x, np.float32).copy()
        self.lib.iface_ap_scs(C.c_int(m), C.c_long(first.C), C.c_long(first.n_chunks), *parts(dp), *parts(sp), *parts(hp),
                              _p(dpx), _p(spx), _p(y))
        return y

    def convert_f64(self, n_rows, n_cols, I, J, vals, Cc, sigma):
        I, J = _i32(I), _i32(J)
        vals = np.ascontiguousarray(vals, np.float64)
        n_chunks = (n_rows + Cc - 1) // Cc
        cap = max(1, 4 * len(I) + n_chunks * Cc)
        while True:
            cp, cl = np.zeros(n_chunks + 1, np.int32), np.zeros(n_chunks, np.int32)
            ci, v, o2n = np.zeros(cap, np.int32), np.zeros(cap, np.float64), np.zeros(n_rows, np.int32)
            ne = self.lib.iface_convert_f64(C.c_long(n_rows), C.c_long(n_cols), C.c_long(len(I)), _p(I), _p(J), _p(vals), C.c_long(Cc),
                                            C.c_long(sigma), _p(cp), _p(cl), _p(ci), _p(v), _p(o2n), C.c_long(cap))
            if ne >= 0:
                return SimpleNamespace(chunk_ptrs=cp, chunk_lengths=cl, col_idxs=ci[:ne], values=v[:ne], old_to_new=o2n, n_elements=int(ne))
            cap = -ne


# ------------------------------------------------------------------------------------------------
# the reference's own CUDA kernels, recompiled for sm_100 (oracle/ref_gpu_driver.cu) — bench.py's gpu_baseline
# ------------------------------------------------------------------------------------------------
def ref_gpu_available():
    return os.path.exists(os.path.join(HERE, "_ref", "libuspmv_ref_gpu.so"))


class RefGpu:
    """spmv_gpu_scs_adv / spmv_gpu_scs / spmv_gpu_csr (code/kernels.hpp:579-775) through the reference's own launchers."""

    KERNELS = {"scs_adv": 0, "scs": 1, "csr": 2}

    def __init__(self):
        self.lib = L = C.CDLL(os.path.join(HERE, "_ref", "libuspmv_ref_gpu.so"))
        L.refgpu_last_error.restype = C.c_char_p
        L.refgpu_spmv.argtypes = [C.c_int, C.c_int, C.c_long, C.c_long, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_long, C.c_void_p,
                                  C.c_long, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_double)]
        self.threads_per_block = int(L.refgpu_threads_per_block())

    def spmv(self, kernel, vt, C_, n_chunks, chunk_ptrs, chunk_lengths, col_idxs, values, x, warmup=3, steps=20):
        """Returns (y, ms per launch)."""
        y = np.zeros(int(n_chunks) * int(C_), NPT[_vt(vt)])
        ms = C.c_double(0.0)
        rc = self.lib.refgpu_spmv(self.KERNELS[kernel], _vt(vt), int(C_), int(n_chunks), _p(_i32(chunk_ptrs)),
                                  _p(_i32(chunk_lengths)) if chunk_lengths is not None else None, _p(_i32(col_idxs)), _p(values), len(col_idxs),
                                  _p(x), len(x), _p(y), int(warmup), int(steps), C.byref(ms))
        if rc:
            raise RuntimeError(self.lib.refgpu_last_error().decode())
        return y, float(ms.value)


# ------------------------------------------------------------------------------------------------
# cuSPARSE, the way the reference's USE_CUSPARSE comparison mode calls it (oracle/cusparse_driver.cu)
# ------------------------------------------------------------------------------------------------
def cusparse_available():
    return os.path.exists(os.path.join(HERE, "_ref", "libuspmv_cusparse.so"))


class CuSparse:
    def __init__(self):
        self.lib = L = C.CDLL(os.path.join(HERE, "_ref", "libuspmv_cusparse.so"))
        L.cusp_last_error.restype = C.c_char_p
        L.cusp_spmv.argtypes = [C.c_int, C.c_int, C.c_long, C.c_long, C.c_long, C.c_long, C.c_int, C.c_void_p, C.c_long, C.c_void_p, C.c_void_p,
                                C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_double)]

    def spmv(self, kind, vt, n_rows, n_cols, nnz, C_, ptrs, cols, vals, x, warmup=3, steps=20):
        """kind 'csr' (ptrs = row_ptrs) or 'sell' (ptrs = chunk_ptrs, slice size C_).  Returns (y[n_rows], ms per call)."""
        ptrs, cols = _i32(ptrs), _i32(cols)
        y = np.zeros(int(n_rows), NPT[_vt(vt)])
        ms = C.c_double(0.0)
        rc = self.lib.cusp_spmv(1 if kind == "sell" else 0, _vt(vt), int(n_rows), int(n_cols), int(nnz), len(cols), int(C_), _p(ptrs), len(ptrs),
                                _p(cols), _p(vals), _p(x), _p(y), int(warmup), int(steps), C.byref(ms))
        if rc:
            raise RuntimeError(self.lib.cusp_last_error().decode())
        return y, float(ms.value)
