"""Host-side matrix ingest: Matrix Market reader and the synthetic generators named in BASELINE.json.

read_mtx mirrors the reference's read_mtx (code/utilities.hpp:2148-2309 + code/mmio.h:138-263):
real / integer / pattern coordinate files, general or symmetric; pattern values are 0.01; a symmetric file
is expanded to general with (i,j) immediately followed by (j,i) (utilities.hpp:2213-2267); the result is
STABLE-sorted by row (utilities.hpp:2278, sort_perm :2139-2146).  Only square matrices are accepted.

The generators produce the same COO order the device generator (uspmv_coo_stencil) produces, so host- and
device-built matrices are bit-identical.
"""
from __future__ import annotations

import numpy as np


def read_mtx_entries(path: str):
    """Text parsing only: (n_rows, n_cols, I, J, values, symmetric) in FILE order, 0-based — the input of the device-side ingest
    (engine.MtxData.from_entries -> uspmv_coo_from_entries), which does the symmetric expansion and the stable row sort."""
    with open(path, "r") as f:
        banner = f.readline().strip().split()
        if len(banner) < 5 or banner[0] != "%%MatrixMarket":
            raise ValueError("mm_read_unsymetric: Could not process Matrix Market banner")
        obj, fmt, field, symm = (s.lower() for s in banner[1:5])
        if obj != "matrix" or fmt != "coordinate":
            raise ValueError("The matrix market file provided is not supported: matrix has to be sparse")
        if field not in ("real", "integer", "pattern"):
            raise ValueError("The matrix market file provided is not supported: matrix has to be real or pattern")
        if symm not in ("general", "symmetric"):
            raise ValueError("The matrix market file provided is not supported: matrix has to be either general or symmetric")
        line = f.readline()
        while line.startswith("%") or not line.strip():
            line = f.readline()
        M, N, nz = (int(t) for t in line.split()[:3])
        if M != N:
            raise ValueError("Matrix not square. Currently only square matrices are supported")
        data = np.loadtxt(f, ndmin=2, dtype=np.float64) if nz else np.zeros((0, 3))
    if data.shape[0] != nz:
        raise ValueError("premature EOF in matrix market file")
    I = np.ascontiguousarray(data[:, 0].astype(np.int32) - 1)
    J = np.ascontiguousarray(data[:, 1].astype(np.int32) - 1)
    V = np.full(nz, 0.01) if field == "pattern" else np.ascontiguousarray(data[:, 2].astype(np.float64))
    return M, N, I, J, V, symm == "symmetric"


def read_mtx(path: str):
    """Returns (n_rows, n_cols, I, J, values) with int32 indices and float64 values."""
    with open(path, "r") as f:
        banner = f.readline().strip().split()
        if len(banner) < 5 or banner[0] != "%%MatrixMarket":
            raise ValueError("mm_read_unsymetric: Could not process Matrix Market banner")
        obj, fmt, field, symm = (s.lower() for s in banner[1:5])
        if obj != "matrix" or fmt != "coordinate":
            raise ValueError("The matrix market file provided is not supported: matrix has to be sparse")
        if field not in ("real", "integer", "pattern"):
            raise ValueError("The matrix market file provided is not supported: matrix has to be real or pattern")
        if symm not in ("general", "symmetric"):
            raise ValueError("The matrix market file provided is not supported: matrix has to be either general or symmetric")
        line = f.readline()
        while line.startswith("%") or not line.strip():
            line = f.readline()
        M, N, nz = (int(t) for t in line.split()[:3])
        if M != N:
            raise ValueError("Matrix not square. Currently only square matrices are supported")
        data = np.loadtxt(f, ndmin=2, dtype=np.float64) if nz else np.zeros((0, 3))
    if data.shape[0] != nz:
        raise ValueError("premature EOF in matrix market file")
    I = data[:, 0].astype(np.int32) - 1
    J = data[:, 1].astype(np.int32) - 1
    V = np.full(nz, 0.01) if field == "pattern" else data[:, 2].astype(np.float64)
    if symm == "symmetric":
        off = I != J
        reps = 1 + off.astype(np.int64)
        idx = np.repeat(np.arange(nz), reps)
        # second copy of an off-diagonal entry is the transposed one
        second = np.zeros(len(idx), bool)
        second[1:] = idx[1:] == idx[:-1]
        I2 = np.where(second, J[idx], I[idx]).astype(np.int32)
        J2 = np.where(second, I[idx], J[idx]).astype(np.int32)
        I, J, V = I2, J2, V[idx]
    order = np.argsort(I, kind="stable")
    return M, N, np.ascontiguousarray(I[order]), np.ascontiguousarray(J[order]), np.ascontiguousarray(V[order])


def stencil_coo(points: int, nx: int, ny: int, nz: int, row0: int = 0, row1: int | None = None):
    """3-D 7- or 27-point stencil, row = (z*ny + y)*nx + x, columns ascending, Dirichlet boundaries,
    diagonal = points-1, off-diagonal = -1.  Rows [row0,row1) as a slab with
    LOCAL row ids and GLOBAL column ids (localize_row_idx, mpi_funcs.hpp:862-877)."""
    assert points in (7, 27)
    n = nx * ny * nz
    row1 = n if row1 is None else row1
    g = np.arange(row0, row1, dtype=np.int64)
    x, y, z = g % nx, (g // nx) % ny, g // (nx * ny)
    Is, Js, Vs, Ks = [], [], [], []
    k = 0
    for dz in (-1, 0, 1):
        for dy in (-1, 0, 1):
            for dx in (-1, 0, 1):
                nzd = (dx != 0) + (dy != 0) + (dz != 0)
                if points == 7 and nzd > 1:
                    continue
                ok = (x + dx >= 0) & (x + dx < nx) & (y + dy >= 0) & (y + dy < ny) & (z + dz >= 0) & (z + dz < nz)
                rows = (g[ok] - row0)
                cols = ((z[ok] + dz) * ny + (y[ok] + dy)) * nx + (x[ok] + dx)
                Is.append(rows)
                Js.append(cols)
                Vs.append(np.full(len(rows), float(points - 1) if nzd == 0 else -1.0))
                Ks.append(np.full(len(rows), k, np.int64))
                k += 1
    I = np.concatenate(Is)
    J = np.concatenate(Js)
    V = np.concatenate(Vs)
    K = np.concatenate(Ks)
    order = np.lexsort((K, I))  # by row, then by neighbour ordinal (= ascending column)
    return (row1 - row0, n, np.ascontiguousarray(I[order].astype(np.int32)), np.ascontiguousarray(J[order].astype(np.int32)),
            np.ascontiguousarray(V[order]))


def _splitmix64(z: np.ndarray) -> np.ndarray:
    z = (z + np.uint64(0x9E3779B97F4A7C15)).astype(np.uint64)
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def _u01(h: np.ndarray) -> np.ndarray:
    return (h >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


POWERLAW_D_MIN_CONFIG4 = 3.28  # calibrated: ~14.9 stored elements per row after the clamp and de-duplication (5.0e8 at 2^25 rows)


def powerlaw_coo(n: int, target_nnz: int, alpha: float = 2.2, max_deg: int = 4096, seed: int = 0x5EED, row0: int = 0, row1: int | None = None,
                 d_min: float | None = None):
    """Irregular power-law matrix of BASELINE.json config 4: row degree
    d = clamp(floor(d_min * (1-u)^(-1/(alpha-1))), 1, max_deg); columns: half within +-1024 of the diagonal,
    half uniform, de-duplicated, ascending; values sign * 10^w, w ~ U(-4, 2).  All randomness is
    splitmix64(seed, row, k), so any row range can be generated independently.  d_min overrides the density derived from
    target_nnz.  Device twin: uspmv_coo_powerlaw (same matrix)."""
    row1 = n if row1 is None else row1
    with np.errstate(over="ignore"):
        rows = np.arange(row0, row1, dtype=np.uint64)
        u = _u01(_splitmix64(np.uint64(seed) ^ (rows * np.uint64(0xD1342543DE82EF95))))
        # mean of floor(d_min*(1-u)^(-1/(a-1))) ~ d_min*(a-1)/(a-2); pick d_min for the target density
        if d_min is None:
            d_min = max(1.0, (target_nnz / n) * (alpha - 2.0) / (alpha - 1.0))
        deg = np.clip(np.floor(d_min * (1.0 - u) ** (-1.0 / (alpha - 1.0))), 1, max_deg).astype(np.int64)
        deg = np.minimum(deg, n)
        ptr = np.concatenate(([0], np.cumsum(deg)))
        total = int(ptr[-1])
        r_of = np.repeat(np.arange(row1 - row0, dtype=np.int64), deg)
        k_of = np.arange(total, dtype=np.int64) - ptr[r_of]
        g_of = (r_of + row0).astype(np.uint64)
        h = _splitmix64(_splitmix64(np.uint64(seed) + g_of) ^ (k_of.astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15)))
        local = (k_of % 2) == 0
        near = (g_of.astype(np.int64) + (h % np.uint64(2049)).astype(np.int64) - 1024) % n
        far = (h % np.uint64(n)).astype(np.int64)
        col = np.where(local, near, far)
        h2 = _splitmix64(h)
        w = _u01(h2) * 6.0 - 4.0
        sign = np.where((h2 & np.uint64(1)) == 0, 1.0, -1.0)
        val = sign * np.power(10.0, w)
    # de-duplicate within a row (keep the first k), ascending columns
    key = r_of * np.int64(n) + col
    order = np.argsort(key, kind="stable")
    key_s = key[order]
    keep = np.ones(total, bool)
    keep[1:] = key_s[1:] != key_s[:-1]
    sel = order[keep]
    return (row1 - row0, n, np.ascontiguousarray(r_of[sel].astype(np.int32)), np.ascontiguousarray(col[sel].astype(np.int32)),
            np.ascontiguousarray(val[sel]))


def random_coo(n: int, avg: int, seed: int = 0, empty_rows: bool = True):
    """Small random test matrix: row-sorted, unsorted columns, possibly empty rows."""
    rng = np.random.default_rng(seed)
    cnt = rng.integers(0 if empty_rows else 1, 2 * avg + 1, n)
    I = np.repeat(np.arange(n), cnt).astype(np.int32)
    J = rng.integers(0, n, len(I)).astype(np.int32)
    V = rng.standard_normal(len(I))
    return n, n, I, J, V
