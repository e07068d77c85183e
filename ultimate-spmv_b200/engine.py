"""Host-side mirror of the reference's library interface (code/interface.hpp, API_doc.md) on top of the
C ABI.  Same names and argument meaning as the reference:

    MtxData, ScsData, convert_to_scs, permute_scs_cols, apply_permutation, partition_precisions,
    uspmv_scs_gpu, uspmv_csr_gpu, execute_uspmv, SpmvKernel

PyTorch is plumbing only: device vectors are torch tensors whose data_ptr() is handed to the C ABI, and
streams are torch streams.  All arithmetic runs in libuspmv_b200.so.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import capi
from .capi import call, vp

VT_CODE = {"dp": capi.F64, "sp": capi.F32, "hp": capi.F16}
NP_OF = {capi.F64: np.float64, capi.F32: np.float32, capi.F16: np.float16}
AP_MODE = {"ap[dp_sp]": capi.AP_DP_SP, "ap[dp_hp]": capi.AP_DP_HP, "ap[sp_hp]": capi.AP_SP_HP, "ap[dp_sp_hp]": capi.AP_DP_SP_HP}
LAYOUT = {"colwise": capi.COLWISE, "rowwise": capi.ROWWISE}


def _torch():
    import torch
    return torch


def torch_dtype(vt: int):
    t = _torch()
    return {capi.F64: t.float64, capi.F32: t.float32, capi.F16: t.float16}[vt]


def vt_code(v) -> int:
    if isinstance(v, str):
        return VT_CODE[v]
    if isinstance(v, int):
        return v
    return {np.float64: capi.F64, np.float32: capi.F32, np.float16: capi.F16}[np.dtype(v).type]


def _hp(a):
    return None if a is None else a.ctypes.data_as(vp)


def _dp(t):
    return None if t is None else vp(t.data_ptr())


def _stream():
    return vp(_torch().cuda.current_stream().cuda_stream)


class Context:
    """One GPU (the reference's `cudaSetDevice(rank % ndev)`, main.cpp:1838-1842)."""

    def __init__(self, device: int = 0):
        h = vp()
        call("uspmv_ctx_create", int(device), C.byref(h))
        self.h = h
        self.device = int(device)

    def sync(self):
        call("uspmv_ctx_sync", self.h)

    def __del__(self):
        if getattr(self, "h", None) and capi is not None:
            capi.lib.uspmv_ctx_destroy(self.h)
            self.h = None


_default_ctx: dict[int, Context] = {}


def default_context(device: int | None = None) -> Context:
    if device is None:
        t = _torch()
        device = t.cuda.current_device() if t.cuda.is_available() else 0
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]


class MtxData:
    """COO matrix (interface.hpp:16-56), resident on the device."""

    def __init__(self, handle, ctx: Context):
        self.h, self.ctx = handle, ctx
        d = (C.c_long * 3)()
        call("uspmv_coo_dims", self.h, d)
        self.n_rows, self.n_cols, self.nnz = int(d[0]), int(d[1]), int(d[2])

    @classmethod
    def from_host(cls, n_rows, n_cols, I, J, values, ctx: Context | None = None):
        ctx = ctx or default_context()
        I = np.ascontiguousarray(I, np.int32)
        J = np.ascontiguousarray(J, np.int32)
        values = np.ascontiguousarray(values)
        if values.dtype not in (np.float64, np.float32, np.float16):
            values = values.astype(np.float64)
        h = vp()
        call("uspmv_coo_from_host", ctx.h, int(n_rows), int(n_cols), len(I), _hp(I), _hp(J), _hp(values), vt_code(values.dtype), C.byref(h))
        return cls(h, ctx)

    @classmethod
    def from_entries(cls, n_rows, n_cols, I, J, values, symmetric, ctx: Context | None = None):
        """read_mtx's post-processing on the device (utilities.hpp:2214-2290): file-order entries -> symmetric expansion -> stable
        sort by row.  See matrices.read_mtx_entries for the text parser."""
        ctx = ctx or default_context()
        I = np.ascontiguousarray(I, np.int32)
        J = np.ascontiguousarray(J, np.int32)
        values = np.ascontiguousarray(values, np.float64)
        h = vp()
        call("uspmv_coo_from_entries", ctx.h, int(n_rows), int(n_cols), len(I), _hp(I), _hp(J), _hp(values), 1 if symmetric else 0, C.byref(h))
        return cls(h, ctx)

    def equilibrate(self):
        """equilibrate_matrix (utilities.hpp:2668-2684) in place; returns (rowmax, colmax) for partition_precisions."""
        rm, cm = np.zeros(self.n_rows), np.zeros(self.n_cols)
        call("uspmv_coo_equilibrate", self.h, _hp(rm), _hp(cm))
        return rm, cm

    @classmethod
    def stencil(cls, points, nx, ny, nz, row0=0, row1=None, ctx: Context | None = None):
        ctx = ctx or default_context()
        if row1 is None:
            row1 = nx * ny * nz
        h = vp()
        call("uspmv_coo_stencil", ctx.h, int(points), int(nx), int(ny), int(nz), int(row0), int(row1), C.byref(h))
        return cls(h, ctx)

    @classmethod
    def powerlaw(cls, n, row0=0, row1=None, d_min=None, alpha=2.2, max_deg=4096, seed=0x5EED, ctx: Context | None = None):
        """BASELINE config 4's power-law matrix, rows [row0, row1) generated on the device (uspmv_coo_powerlaw)."""
        from .matrices import POWERLAW_D_MIN_CONFIG4
        ctx = ctx or default_context()
        h = vp()
        call("uspmv_coo_powerlaw", ctx.h, int(n), int(row0), int(n if row1 is None else row1), float(d_min or POWERLAW_D_MIN_CONFIG4),
             float(alpha), int(max_deg), int(seed), C.byref(h))
        return cls(h, ctx)

    def device_arrays(self):
        ptrs = [vp() for _ in range(3)]
        call("uspmv_coo_device_arrays", self.h, *[C.byref(p) for p in ptrs])
        return dict(zip(("I", "J", "values"), ptrs))

    def to_host(self, mt=np.float64):
        I = np.zeros(self.nnz, np.int32)
        J = np.zeros(self.nnz, np.int32)
        V = np.zeros(self.nnz, mt)
        call("uspmv_coo_export", self.h, _hp(I), _hp(J), _hp(V))
        return I, J, V

    def __del__(self):
        if getattr(self, "h", None) and capi is not None:
            capi.lib.uspmv_coo_destroy(self.h)
            self.h = None


class ScsData:
    """SELL-C-sigma matrix (interface.hpp:58-80), resident on the device."""

    def __init__(self, handle, ctx: Context):
        self.h, self.ctx = handle, ctx
        d = (C.c_long * 8)()
        call("uspmv_scs_dims", self.h, d)
        (self.C, self.sigma, self.n_rows, self.n_cols, self.n_rows_padded, self.n_chunks, self.n_elements, self.nnz) = (int(v) for v in d)
        self.vt = None

    def export(self):
        """Host copies of every array (for bit-exact comparison / filling a C++ ScsData)."""
        from types import SimpleNamespace
        cp = np.zeros(self.n_chunks + 1, np.int32)
        cl = np.zeros(self.n_chunks, np.int32)
        ci = np.zeros(self.n_elements, np.int32)
        v = np.zeros(self.n_elements, NP_OF[self.vt])
        o2n = np.zeros(self.n_rows, np.int32)
        n2o = np.zeros(self.n_rows_padded, np.int32)
        call("uspmv_scs_export", self.h, _hp(cp), _hp(cl), _hp(ci), _hp(v), _hp(o2n), _hp(n2o))
        return SimpleNamespace(C=self.C, sigma=self.sigma, n_rows=self.n_rows, n_cols=self.n_cols, n_rows_padded=self.n_rows_padded,
                               n_chunks=self.n_chunks, n_elements=self.n_elements, nnz=self.nnz, chunk_ptrs=cp, chunk_lengths=cl,
                               col_idxs=ci, values=v, old_to_new=o2n, new_to_old=n2o)

    def device_arrays(self):
        ptrs = [vp() for _ in range(6)]
        call("uspmv_scs_device_arrays", self.h, *[C.byref(p) for p in ptrs])
        return dict(zip(("chunk_ptrs", "chunk_lengths", "col_idxs", "values", "old_to_new", "new_to_old"), ptrs))

    def __del__(self):
        if getattr(self, "h", None) and capi is not None:
            capi.lib.uspmv_scs_destroy(self.h)
            self.h = None


def convert_to_scs(mtx: MtxData, C_: int, sigma: int, value_type="dp", fixed_permutation=None) -> ScsData:
    """convert_to_scs(local_mtx, C, sigma, scs, fixed_permutation) — utilities.hpp:1842-2104."""
    h = vp()
    fp = None if fixed_permutation is None else np.ascontiguousarray(fixed_permutation, np.int32)
    if fp is not None and len(fp) < mtx.n_rows:
        raise ValueError("fixed_permutation must have n_rows entries")
    vt = vt_code(value_type)
    call("uspmv_scs_build", mtx.ctx.h, mtx.h, int(C_), int(sigma), vt, _hp(fp), C.byref(h))
    s = ScsData(h, mtx.ctx)
    s.vt = vt
    return s


def permute_scs_cols(scs: ScsData, perm=None) -> None:
    """permute_scs_cols(scs, perm) — utilities.hpp:1802-1831; perm=None uses scs.old_to_new_idx (main.cpp:1308)."""
    p = None if perm is None else np.ascontiguousarray(perm, np.int32)
    call("uspmv_scs_permute_cols", scs.h, _hp(p))


def apply_permutation(out, vec, perm_dev_ptr, n: int, ctx: Context | None = None) -> None:
    """apply_permutation(permuted_vec, vec_to_permute, perm, n) — utilities.hpp:1768-1782 (device tensors)."""
    ctx = ctx or default_context()
    vt = {8: capi.F64, 4: capi.F32, 2: capi.F16}[out.element_size()]
    call("uspmv_apply_permutation", ctx.h, _dp(out), _dp(vec), perm_dev_ptr, int(n), vt, _stream())


def apply_permutation_block(out, vec, perm_dev_ptr, n, bvs, ld, layout, ctx: Context | None = None) -> None:
    """apply_strided_permutation (utilities.hpp:1784-1799) for whole block rows."""
    ctx = ctx or default_context()
    vt = {8: capi.F64, 4: capi.F32, 2: capi.F16}[out.element_size()]
    lay = LAYOUT[layout] if isinstance(layout, str) else layout
    call("uspmv_apply_permutation_block", ctx.h, _dp(out), _dp(vec), perm_dev_ptr, int(n), vt, int(bvs), int(ld), lay, _stream())


def uspmv_scs_gpu(C_, n_chunks, chunk_ptrs, chunk_lengths, col_idxs, values, x, y, ctx: Context | None = None):
    """uspmv_scs_gpu(C, n_chunks, chunk_ptrs, chunk_lengths, col_idxs, values, x, y) — interface.hpp:1766-1793.
    All arrays are device torch tensors."""
    ctx = ctx or default_context()
    vt = {8: capi.F64, 4: capi.F32, 2: capi.F16}[values.element_size()]
    call("uspmv_scs_gpu", ctx.h, vt, int(C_), int(n_chunks), _dp(chunk_ptrs), _dp(chunk_lengths), _dp(col_idxs), _dp(values), _dp(x), _dp(y), _stream())


def uspmv_csr_gpu(n_rows, row_ptrs, col_idxs, values, x, y, ctx: Context | None = None):
    """uspmv_csr_gpu(num_rows, row_ptrs, ..., col_idxs, values, x, y) — interface.hpp:1741-1760."""
    ctx = ctx or default_context()
    vt = {8: capi.F64, 4: capi.F32, 2: capi.F16}[values.element_size()]
    call("uspmv_csr_gpu", ctx.h, vt, int(n_rows), _dp(row_ptrs), _dp(col_idxs), _dp(values), _dp(x), _dp(y), _stream())


def spmv(scs: ScsData, x, y) -> None:
    """SpmvKernel::execute for one precision (classes_structs.hpp:997-1035): y (n_rows_padded, permuted order)."""
    call("uspmv_spmv", scs.h, _dp(x), _dp(y), _stream())


def spmv_unpermuted(scs: ScsData, x, y) -> None:
    call("uspmv_spmv_unpermuted", scs.h, _dp(x), _dp(y), _stream())


def spmmv(scs: ScsData, X, Y, block_vec_size: int, vec_length: int, layout="rowwise") -> None:
    """block_spmv_* (kernels.hpp:68-154,306-398)."""
    lay = LAYOUT[layout] if isinstance(layout, str) else layout
    call("uspmv_spmmv", scs.h, _dp(X), _dp(Y), int(block_vec_size), int(vec_length), lay, _stream())


def spmv_host(scs: ScsData, x_host: np.ndarray, y_host: np.ndarray) -> None:
    """Reference-facing host-buffer call: H2D(x) + kernel + D2H(y) + sync."""
    call("uspmv_spmv_host", scs.h, _hp(x_host), len(x_host), _hp(y_host), len(y_host))


def partition_precisions(mtx: MtxData, ap_value_type: str, ap_threshold_1: float, ap_threshold_2: float = 0.0,
                         largest_row_elems=None, largest_col_elems=None):
    """partition_precisions — interface.hpp:690-978.  Returns (dp, sp, hp) MtxData (None where unused)."""
    mode = AP_MODE[ap_value_type]
    rm = None if largest_row_elems is None else np.ascontiguousarray(largest_row_elems, np.float64)
    cm = None if largest_col_elems is None else np.ascontiguousarray(largest_col_elems, np.float64)
    hs = [vp(), vp(), vp()]
    call("uspmv_partition_precisions", mtx.ctx.h, mtx.h, mode, float(ap_threshold_1), float(ap_threshold_2), _hp(rm), _hp(cm),
         C.byref(hs[0]), C.byref(hs[1]), C.byref(hs[2]))
    return tuple(MtxData(h, mtx.ctx) if h.value else None for h in hs)


def ap_spmv(ap_value_type, dp: ScsData | None, sp: ScsData | None, hp: ScsData | None, x, y) -> None:
    """uspmv_{scs,csr}_ap* (interface.hpp:1129-1733) as one fused pass."""
    mode = AP_MODE[ap_value_type] if isinstance(ap_value_type, str) else ap_value_type
    call("uspmv_ap_spmv", mode, dp.h if dp else None, sp.h if sp else None, hp.h if hp else None, _dp(x), _dp(y), _stream())


class BandedPlan:
    """Column-banded execution plan (EXPERIMENTAL; uspmv_banded_*): for matrices whose x does not fit the L2.  x in the original
    column numbering, y (n_rows_padded) in the permuted row order given by `old_to_new`."""

    def __init__(self, mtx: MtxData, C_: int, sigma: int, vt: str = "dp", ap: str | None = None, t1: float = 0.0, t2: float = 0.0,
                 n_bands: int = 0):
        h = vp()
        call("uspmv_banded_build", mtx.ctx.h, mtx.h, int(C_), int(sigma), VT_CODE[vt], AP_MODE[ap] if ap else -1, float(t1), float(t2),
             int(n_bands), C.byref(h))
        self.h, self.ctx = h, mtx.ctx
        d = (C.c_long * 8)()
        call("uspmv_banded_dims", self.h, d)
        (self.n_bands, self.band_width, self.n_rows, self.n_cols, self.n_rows_padded, self.nnz, self.n_elements, self.vt) = [int(v) for v in d]
        self.old_to_new = np.zeros(max(self.n_rows, 1), np.int32)
        call("uspmv_banded_perm", self.h, _hp(self.old_to_new))
        self.old_to_new = self.old_to_new[: self.n_rows]

    def spmv(self, x, y) -> None:
        call("uspmv_banded_spmv", self.h, _dp(x), _dp(y), _stream())

    def __del__(self):
        if getattr(self, "h", None) and capi is not None:
            capi.lib.uspmv_banded_destroy(self.h)
            self.h = None


def execute_uspmv(scs: ScsData, x, y, ap=None) -> None:
    """execute_uspmv (interface.hpp:1871-2187): SCS kernels iff C > 1 or sigma > 1, else CRS; AP by ap_value_type.
    `ap` = (ap_value_type, dp, sp, hp) for adaptive precision."""
    if ap is not None:
        ap_spmv(ap[0], ap[1], ap[2], ap[3], x, y)
    else:
        spmv(scs, x, y)


def seg_work_sharing_arr(seg_method: str, n_rows: int, I: np.ndarray, comm_size: int) -> np.ndarray:
    """seg_work_sharing_arr — mpi_funcs.hpp:424-622 ('seg-rows' / 'seg-nnz')."""
    I = np.ascontiguousarray(I, np.int32)
    wsa = np.zeros(comm_size + 1, np.int32)
    methods = {"seg-nnz": capi.SEG_NNZ, "seg-rows": capi.SEG_ROWS}
    key = seg_method.replace("_", "-")
    if key not in methods:  # seg-metis (mpi_funcs.hpp:494-600) is out of scope; typos must not fall back silently
        raise ValueError(f"seg_work_sharing_arr: unknown seg_method {seg_method!r} (supported: seg-rows, seg-nnz)")
    m = methods[key]
    call("uspmv_seg_work_sharing_arr", m, int(n_rows), len(I), _hp(I), int(comm_size), _hp(wsa))
    return wsa


def seg_mtx_struct(total_mtx: MtxData, work_sharing_arr, loop_rank: int):
    """seg_mtx_struct + localize_row_idx — mpi_funcs.hpp:636-674,862-877: the slab of `loop_rank` (local row ids, global columns) of a
    row-sorted device COO.  Returns (local MtxData, number of distinct rows = the reference's local n_rows, mpi_funcs.hpp:770)."""
    wsa = np.ascontiguousarray(work_sharing_arr, np.int32)
    h, nd = vp(), C.c_long(0)
    call("uspmv_coo_seg_mtx", total_mtx.h, _hp(wsa), int(loop_rank), len(wsa) - 1, C.byref(h), C.byref(nd))
    return MtxData(h, total_mtx.ctx), int(nd.value)


def generate_inv_perm(perm_dev_ptr, inv_perm_dev_ptr, perm_len: int, inv_len: int | None = None, ctx: Context | None = None) -> None:
    """generate_inv_perm(perm, inv_perm, perm_len) — utilities.hpp:1755-1766 (device arrays)."""
    ctx = ctx or default_context()
    call("uspmv_generate_inv_perm", ctx.h, perm_dev_ptr, inv_perm_dev_ptr, int(perm_len), int(perm_len if inv_len is None else inv_len), _stream())


def random_init(vmin: float, vmax: float, n: int, value_type="dp", n_rows: int = -1, vec_length: int = 0, block_vec_size: int = 1,
                layout="colwise") -> np.ndarray:
    """random_init + the padding rule of init_std_vec_with_ptr_or_value — utilities.hpp:880-981 (host logic; bit-identical x)."""
    vt = vt_code(value_type)
    out = np.zeros(int(n), NP_OF[vt])
    lay = LAYOUT[layout] if isinstance(layout, str) else layout
    call("uspmv_random_init_host", float(vmin), float(vmax), int(n), vt, _hp(out), int(n_rows), int(vec_length), int(block_vec_size), lay)
    return out


@dataclass
class SpmvKernel:
    """Harness-side kernel object (classes_structs.hpp:280-1166): picks the kernel from the format, owns x/y
    handling in permuted space and exposes execute()/swap_local_vectors()."""
    scs: ScsData
    block_vec_size: int = 1
    layout: str = "colwise"
    vec_length: int = 0
    ap: tuple | None = None
    n_calls: int = field(default=0)

    def execute(self, x, y):
        self.n_calls += 1
        if self.ap is not None:
            ap_spmv(self.ap[0], self.ap[1], self.ap[2], self.ap[3], x, y)
        elif self.block_vec_size > 1:
            spmmv(self.scs, x, y, self.block_vec_size, self.vec_length or self.scs.n_rows_padded, self.layout)
        else:
            spmv(self.scs, x, y)

    @staticmethod
    def swap_local_vectors(x, y):
        """classes_structs.hpp:1130-1165: y becomes the next x."""
        return y, x


def spmv_host_ptr(scs: ScsData, x_ptr: int, x_len: int, y_ptr: int, y_len: int) -> None:
    """uspmv_spmv_host on raw host addresses (e.g. pinned torch tensors)."""
    call("uspmv_spmv_host", scs.h, vp(x_ptr), int(x_len), vp(y_ptr), int(y_len))


class SingleGpuSpmv:
    """bench.py's N = 1 runner: stencil matrix generated and converted on the device, the harness' bench-mode
    protocol (bench_spmv, main.cpp:380-527): x = 5.0 in permuted space, repeated execute()."""

    kernel_name = "k_scs32_stream"

    def __init__(self, ctx: Context, points: int, n: int, C_: int, sigma: int, vt: str):
        t = _torch()
        self.ctx, self.points, self.n = ctx, points, n
        mtx = MtxData.stencil(points, n, n, n, ctx=ctx)
        if vt != "dp":  # the stencil generator emits doubles; convert_to_scs narrows (MT -> VT)
            pass
        self.scs = convert_to_scs(mtx, C_, sigma, vt)
        permute_scs_cols(self.scs)
        del mtx
        s = self.scs
        self.nnz, self.n_elements, self.n_chunks = s.nnz, s.n_elements, s.n_chunks
        self.n_rows, self.n_rows_padded, self.n_cols_local, self.n_halo = s.n_rows, s.n_rows_padded, s.n_cols, 0
        dt = torch_dtype(s.vt)
        self.x = t.full((max(s.n_rows_padded, s.n_cols),), 5.0, dtype=dt, device=f"cuda:{ctx.device}")
        self.y = t.zeros(s.n_rows_padded, dtype=dt, device=f"cuda:{ctx.device}")
        self.e2e_h2d_bytes = self.x.numel() * self.x.element_size()
        self.e2e_d2h_bytes = self.y.numel() * self.y.element_size()
        if s.C == 1 and s.sigma == 1:
            self.kernel_name = "k_csr_spmv"
        elif s.C != 32:
            self.kernel_name = "k_scs_spmv"

    def step(self):
        spmv(self.scs, self.x, self.y)

    def validate(self, steps: int = 2) -> float:
        """Checked steps: x = f(row) (permuted like the harness does, main.cpp:86-93), y against the stencil formula.  Returns
        max |y - ref| / sum|a||x|; x is restored to the timing default 5.0."""
        from . import validate as V
        t = _torch()
        s = self.scs
        x_loc, y_ref, scale = V.stencil_product(self.points, self.n, self.n, self.n, 0, s.n_rows, 0, self.x.device, self.x.dtype)
        perm = V.device_int_tensor(s.device_arrays()["old_to_new"].value, s.n_rows, self.x.device).long()
        self.x[: s.n_rows][perm] = x_loc.to(self.x.dtype)
        for _ in range(steps):
            self.y.fill_(float("nan"))
            self.step()
        t.cuda.synchronize()
        worst = V.max_rel_err(self.y[: s.n_rows_padded][perm], y_ref, scale)
        self._vref = (x_loc, y_ref, scale, perm)  # the host-buffer path (time_e2e) is checked against the same product
        self.x.fill_(5.0)
        self.y.zero_()
        t.cuda.synchronize()
        return worst

    def close(self):
        pass

    def time_kernel(self, steps: int) -> float:
        """Average device time of the SpMV kernel (ms), CUDA events on the launching stream."""
        t = _torch()
        e0, e1 = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
        t.cuda.synchronize()
        e0.record()
        for _ in range(steps):
            spmv(self.scs, self.x, self.y)
        e1.record()
        t.cuda.synchronize()
        return e0.elapsed_time(e1) / steps

    def time_e2e(self, steps: int, barrier, pipelined: bool = True) -> float:
        """Seconds per step through the host-buffer C-ABI call with pinned host x / y.  pipelined: up to three SpMVs in
        flight (uspmv_spmv_host_submit / _wait), each still copying its own x in and its own y out."""
        import time
        t = _torch()
        nslots = 3 if pipelined else 1
        xh = [t.full((self.x.numel(),), 5.0, dtype=self.x.dtype).pin_memory() for _ in range(nslots)]
        yh = [t.zeros(self.y.numel(), dtype=self.y.dtype).pin_memory() for _ in range(nslots)]
        vref = getattr(self, "_vref", None)
        if vref is not None:  # host x = the validation vector (permuted), so the y that comes back can be checked
            xp = t.zeros_like(self.x)
            xp[: self.scs.n_rows][vref[3]] = vref[0].to(self.x.dtype)
            for b in xh:
                b.copy_(xp)
            del xp

        def run(k):
            if not pipelined:
                for _ in range(k):
                    spmv_host_ptr(self.scs, xh[0].data_ptr(), xh[0].numel(), yh[0].data_ptr(), yh[0].numel())
                return
            for i in range(k):
                sl = i % nslots
                call("uspmv_spmv_host_wait", self.scs.h, sl)
                call("uspmv_spmv_host_submit", self.scs.h, vp(xh[sl].data_ptr()), xh[sl].numel(), vp(yh[sl].data_ptr()), yh[sl].numel(), sl)
            for sl in range(nslots):
                call("uspmv_spmv_host_wait", self.scs.h, sl)
        run(3)
        barrier()
        t0 = time.perf_counter()
        run(steps)
        barrier()
        dt = (time.perf_counter() - t0) / steps
        self.e2e_checksum = float(yh[0][: self.n_rows].double().sum())
        if vref is not None:
            from . import validate as V
            self.e2e_max_rel_err = max(V.max_rel_err(b.to(self.x.device)[: self.scs.n_rows_padded][vref[3]], vref[1], vref[2]) for b in yh)
        return dt
