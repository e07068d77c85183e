"""Result validation of a (distributed) SpMV / SpMMV / AP step against an INDEPENDENT reference — what the reference harness does in
solve mode after gather_results (code/main.cpp:528-631,968-990; write_results.hpp:160-440, there against MKL).

The reference product is formed analytically from the stencil definition (or, for generated COO matrices, by a torch index_add over
the COO triplets) with x = f(GLOBAL row), so it needs neither the SELL-C-sigma structure nor the halo exchange it checks: a push
that never lands, lands at the wrong index or lands in the wrong epoch changes y.  The halo tail of x is poisoned with NaN first.
torch is plumbing here (device tensors); the SpMV under test is the library's.
"""
from __future__ import annotations

import numpy as np

TOL = {8: 1e-12, 4: 1e-5, 2: 1e-2}  # BASELINE.json north_star: relative, per element, against sum |a||x|


def x_of(g, v=0):
    """The validation vector: x_v[g] = sin((0.37 + 0.11 v) g) + 1.5 (g = GLOBAL row index, float64 tensor)."""
    import torch
    return torch.sin(g * (0.37 + 0.11 * v)) + 1.5


def stencil_product(points, nx, ny, nz, row0, row1, v, device, x_dtype=None):
    """(x_local, y_ref, scale) for rows [row0, row1) of the 7- / 27-point stencil matrix (diagonal points-1, off-diagonals -1,
    Dirichlet; uspmv_coo_stencil) and x = x_of(., v).  scale = sum_j |a_ij x_j|.  x_dtype: the precision x is STORED in (the
    product uses the rounded x, so a narrow-precision run is judged on its arithmetic, not on the rounding of its input)."""
    import torch
    g = torch.arange(row0, row1, device=device, dtype=torch.int64)
    ix, iy, iz = g % nx, (g // nx) % ny, g // (nx * ny)
    gf = g.to(torch.float64)

    def xs(gg):
        x = x_of(gg, v)
        return x.to(x_dtype).to(torch.float64) if x_dtype is not None and x_dtype != torch.float64 else x
    x_loc = xs(gf)
    y = float(points - 1) * x_loc
    scale = float(points - 1) * x_loc.abs()
    for dz in (-1, 0, 1):
        for dy in (-1, 0, 1):
            for dx in (-1, 0, 1):
                nzd = (dx != 0) + (dy != 0) + (dz != 0)
                if nzd == 0 or (points == 7 and nzd > 1):
                    continue
                ok = torch.ones_like(g, dtype=torch.bool)
                if dx:
                    ok &= (ix + dx >= 0) & (ix + dx < nx)
                if dy:
                    ok &= (iy + dy >= 0) & (iy + dy < ny)
                if dz:
                    ok &= (iz + dz >= 0) & (iz + dz < nz)
                xn = xs(gf + float((dz * ny + dy) * nx + dx))
                xn = torch.where(ok, xn, torch.zeros_like(xn))
                y -= xn
                scale += xn.abs()
    return x_loc, y, scale


def max_rel_err(y, y_ref, scale):
    """max |y - y_ref| / scale; NaN / inf anywhere in y -> inf."""
    import torch
    d = (y.to(torch.float64) - y_ref).abs() / scale.clamp_min(1e-300)
    if not bool(torch.isfinite(y.to(torch.float64)).all()):
        return float("inf")
    return float(d.max()) if d.numel() else 0.0


def device_int_tensor(ptr, n, device):
    """Zero-copy int32 torch view of library-owned device memory."""
    import torch

    class _A:
        pass
    a = _A()
    a.__cuda_array_interface__ = {"shape": (int(n),), "typestr": np.dtype(np.int32).str, "data": (int(ptr), False), "version": 3}
    return torch.as_tensor(a, device=device)


def to_half_rn(v):
    """double -> fp16 with ONE rounding (what static_cast<_Float16>(double) and __double2half do; torch's .half() on a double tensor
    goes through float and double-rounds).  Among the double-rounded value and its two fp16 neighbours the one nearest to v wins; a v
    exactly on an fp16 tie is exact in fp32, so the second rounding of the detour already applied ties-to-even correctly."""
    import torch
    c = v.float().half()
    bits = c.view(torch.int16)
    best, best_err = c, (c.double() - v).abs()
    for delta in (1, -1):
        cand = (bits + delta).view(torch.float16)
        err = (cand.double() - v).abs()
        take = (err < best_err) & torch.isfinite(cand)
        best = torch.where(take, cand, best)
        best_err = torch.where(take, err, best_err)
    return best


def coo_ap_product(mtx, mode, t1, t2, x_dtype, device):
    """(y_ref, scale) per local row for the adaptive-precision product of a device COO (engine.MtxData holding doubles) with
    x[j] = x_of(GLOBAL column j) stored in x_dtype: values as partition_precisions stores them (interface.hpp:938-964: dp / fp32 / fp16
    by |v| against t1, t2), products and sums in fp64.  mode None: plain dp."""
    import torch

    class _A:
        pass
    d = mtx.device_arrays()
    I = device_int_tensor(d["I"].value, mtx.nnz, device).long()
    J = device_int_tensor(d["J"].value, mtx.nnz, device)
    a_ = _A()
    a_.__cuda_array_interface__ = {"shape": (int(mtx.nnz),), "typestr": np.dtype(np.float64).str, "data": (int(d["values"].value), False), "version": 3}
    vals = torch.as_tensor(a_, device=device)
    a = vals.abs()
    if mode == "ap[dp_sp_hp]":
        vs = torch.where(a >= t1, vals, torch.where(a >= t2, vals.float().double(), to_half_rn(vals).double()))
    elif mode == "ap[dp_sp]":
        vs = torch.where(a >= t1, vals, vals.float().double())
    elif mode == "ap[dp_hp]":
        vs = torch.where(a >= t1, vals, to_half_rn(vals).double())
    elif mode == "ap[sp_hp]":
        vs = torch.where(a >= t1, vals.float().double(), to_half_rn(vals).double())
    else:
        vs = vals
    del a
    prod = vs * x_of(J.double()).to(x_dtype).double()
    del vs
    ref = torch.zeros(mtx.n_rows, dtype=torch.float64, device=device).index_add_(0, I, prod)
    scale = torch.zeros(mtx.n_rows, dtype=torch.float64, device=device).index_add_(0, I, prod.abs_())
    return ref, scale
