// Device-side MtxData (COO) handling and SELL-C-sigma construction.
//
// Replaces, bit-exactly, the reference's serial host routines
//   convert_to_scs     code/utilities.hpp:1842-2104  (library twin code/interface.hpp:401-656)
//   permute_scs_cols   code/utilities.hpp:1802-1831
// Pipeline (all on the GPU, int32 indices per rank like the reference's IT=int):
//   [stable radix sort by row if the COO is not row-sorted]  -> row_ptr (boundary detection)
//   -> per-row counts incl. zero-count padding rows
//   -> one thread per sigma-window runs libstdc++-13's std::sort algorithm (introsort with the
//      comparator `a.count > b.count`, utilities.hpp:1936-1940) so that the UNSTABLE tie order of
//      the reference is reproduced exactly
//   -> chunk_lengths = max count per chunk, chunk_ptrs = exclusive scan of len*C (64-bit, overflow checked)
//   -> old_to_new / new_to_old
//   -> gather-fill: one thread per padded row writes its slots j = 0..len-1 of the chunk, taking the
//      row's COO elements in input order (utilities.hpp:2013-2036) and padding with value 0 / column 0
//      (utilities.hpp:1991-2002).  Writes are coalesced across the C lanes of a chunk.
#include "common.cuh"

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <vector>

using namespace uspmv;

namespace {

constexpr int TPB = 256;
inline unsigned blocks_for(long n, int tpb = TPB) { return (unsigned)((n + tpb - 1) / tpb); }

// ---------------------------------------------------------------------------------------------
// small kernels
// ---------------------------------------------------------------------------------------------
__global__ void k_check_sorted(const int *__restrict__ I, long nnz, long n_rows, int *flags) {
    long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i >= nnz) return;
    int r = I[i];
    if (r < 0 || r >= n_rows) atomicOr(&flags[1], 1);
    if (i > 0 && I[i - 1] > r) atomicOr(&flags[0], 1);
}

__global__ void k_check_cols(const int *__restrict__ J, long nnz, int *flags) {
    long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i < nnz && J[i] < 0) atomicOr(&flags[1], 1);
}

__global__ void k_iota(int *p, long n) {
    long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i < n) p[i] = (int)i;
}

// row_ptr[r] = first s with Isorted[s] >= r, for r in [0, n_ptr)
__global__ void k_row_ptr(const int *__restrict__ Is, long nnz, long n_ptr, int *__restrict__ row_ptr) {
    long s = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (s > nnz) return;
    long lo = (s == 0) ? 0 : (long)Is[s - 1] + 1;
    long hi = (s == nnz) ? n_ptr - 1 : (long)Is[s];
    for (long r = lo; r <= hi; ++r) row_ptr[r] = (int)s;
}

struct RC {
    int idx;  // original row
    int cnt;  // stored elements in that row
};
static_assert(sizeof(RC) == 8, "RC must be one 64-bit word");

__global__ void k_init_rc(const int *__restrict__ row_ptr, long n_rows, long n_pad, RC *__restrict__ rc) {
    long r = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (r >= n_pad) return;
    RC v;
    v.idx = (int)r;
    v.cnt = (r < n_rows) ? row_ptr[r + 1] - row_ptr[r] : 0;
    rc[r] = v;
}

// fixed_permutation mode (utilities.hpp:1911-1928): position p holds the row i with fixed_perm[i] == p
__global__ void k_scatter_fixed(const int *__restrict__ row_ptr, const int *__restrict__ fixed_perm, long n_rows, long n_pad,
                                RC *__restrict__ rc, int *flags) {
    long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i >= n_rows) return;
    int p = fixed_perm[i];
    if (p < 0 || p >= n_pad) { atomicOr(&flags[1], 1); return; }
    if (atomicExch(&rc[p].idx, (int)i) != -1) { atomicOr(&flags[1], 2); return; }  // two rows mapped to one position
    rc[p].cnt = row_ptr[i + 1] - row_ptr[i];
}

__global__ void k_fill_rc_empty(long n_pad, RC *rc) {
    long r = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (r < n_pad) { RC v; v.idx = -1; v.cnt = 0; rc[r] = v; }
}

// ---------------------------------------------------------------------------------------------
// libstdc++ 13 std::sort on RC with comparator "longer rows first" — one thread per sigma window.
// Algorithm (published in GCC's bits/stl_algo.h:1848-1951 and bits/stl_heap.h): __introsort_loop with
// depth budget 2*floor(log2 n) and threshold 16, median-of-three pivot moved to `first`,
// __unguarded_partition, heapsort when the budget is spent, then __final_insertion_sort.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ bool longer(const RC &a, const RC &b) { return a.cnt > b.cnt; }
__device__ __forceinline__ void swp(RC *a, RC *b) { RC t = *a; *a = *b; *b = t; }

__device__ void d_push_heap(RC *first, long hole, long top, RC value) {
    long parent = (hole - 1) / 2;
    while (hole > top && longer(first[parent], value)) {
        first[hole] = first[parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    first[hole] = value;
}

__device__ void d_adjust_heap(RC *first, long hole, long len, RC value) {
    const long top = hole;
    long child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (longer(first[child], first[child - 1])) child--;
        first[hole] = first[child];
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        first[hole] = first[child - 1];
        hole = child - 1;
    }
    d_push_heap(first, hole, top, value);
}

__device__ void d_heapsort(RC *first, long n) {
    if (n >= 2) {
        long parent = (n - 2) / 2;
        for (;;) {
            RC v = first[parent];
            d_adjust_heap(first, parent, n, v);
            if (parent == 0) break;
            parent--;
        }
    }
    long last = n;
    while (last > 1) {
        --last;
        RC v = first[last];
        first[last] = first[0];
        d_adjust_heap(first, 0, last, v);
    }
}

__device__ __forceinline__ void d_median_to_first(RC *result, RC *a, RC *b, RC *c) {
    if (longer(*a, *b)) {
        if (longer(*b, *c)) swp(result, b);
        else if (longer(*a, *c)) swp(result, c);
        else swp(result, a);
    } else if (longer(*a, *c)) swp(result, a);
    else if (longer(*b, *c)) swp(result, c);
    else swp(result, b);
}

__device__ __forceinline__ RC *d_unguarded_partition(RC *first, RC *last, RC *pivot_p) {
    const RC pivot = *pivot_p;  // the pivot slot (window[0]) is never written during the partition
    for (;;) {
        while (longer(*first, pivot)) ++first;
        --last;
        while (longer(pivot, *last)) --last;
        if (!(first < last)) return first;
        swp(first, last);
        ++first;
    }
}

__device__ __forceinline__ void d_unguarded_linear_insert(RC *last) {
    RC val = *last;
    RC *next = last - 1;
    while (longer(val, *next)) {
        *last = *next;
        last = next;
        --next;
    }
    *last = val;
}

__device__ void d_insertion_sort(RC *first, RC *last) {
    if (first == last) return;
    for (RC *i = first + 1; i != last; ++i) {
        if (longer(*i, *first)) {
            RC val = *i;
            for (RC *p = i; p != first; --p) *p = *(p - 1);
            *first = val;
        } else
            d_unguarded_linear_insert(i);
    }
}

__device__ void d_std_sort(RC *first, long n) {
    if (n <= 1) return;
    long lg = 0;
    for (long t = n; t > 1; t >>= 1) ++lg;
    // __introsort_loop, recursion on the right part turned into an explicit stack (depth <= 2*lg)
    struct Frame { RC *first, *last; long depth; };
    Frame stack[130];
    int sp = 0;
    stack[sp++] = Frame{first, first + n, 2 * lg};
    while (sp > 0) {
        Frame f = stack[--sp];
        while (f.last - f.first > 16) {
            if (f.depth == 0) {
                d_heapsort(f.first, f.last - f.first);
                break;
            }
            --f.depth;
            RC *mid = f.first + (f.last - f.first) / 2;
            d_median_to_first(f.first, f.first + 1, mid, f.last - 1);
            RC *cut = d_unguarded_partition(f.first + 1, f.last, f.first);
            stack[sp++] = Frame{cut, f.last, f.depth};
            f.last = cut;
        }
    }
    if (n > 16) {
        d_insertion_sort(first, first + 16);
        for (RC *i = first + 16; i != first + n; ++i) d_unguarded_linear_insert(i);
    } else
        d_insertion_sort(first, first + n);
}

__global__ void k_sort_windows(RC *rc, long n_pad, long sigma, long n_windows) {
    long w = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (w >= n_windows) return;
    long begin = w * sigma;
    long end = begin + sigma < n_pad ? begin + sigma : n_pad;
    d_std_sort(rc + begin, end - begin);
}

// The same sort with the window staged in SHARED memory: one CTA per window loads its {row, count} records with coalesced accesses,
// one thread runs the (inherently sequential, data-dependent) introsort restatement on the shared copy — ~30-cycle instead of
// ~500-cycle dependent accesses — and the CTA writes the window back.  Bit-exact by construction (same code, same order of
// comparisons and swaps).  sigma = 16 384 at 2^25 rows: seconds -> tens of milliseconds (round 1: 1.1-4.2 s).
__global__ void __launch_bounds__(64) k_sort_windows_smem(RC *rc, long n_pad, long sigma, long n_windows) {
    extern __shared__ __align__(8) unsigned char win_raw[];
    RC *win = reinterpret_cast<RC *>(win_raw);
    for (long w = blockIdx.x; w < n_windows; w += gridDim.x) {
        const long begin = w * sigma;
        const long n = begin + sigma < n_pad ? sigma : n_pad - begin;
        for (long i = threadIdx.x; i < n; i += blockDim.x) win[i] = rc[begin + i];
        __syncthreads();
        if (threadIdx.x == 0) d_std_sort(win, n);
        __syncthreads();
        for (long i = threadIdx.x; i < n; i += blockDim.x) rc[begin + i] = win[i];
        __syncthreads();
    }
}

// chunk_lengths[c] = max count in chunk; len64[c] = len*C
__global__ void k_chunk_lengths(const RC *__restrict__ rc, long n_chunks, int C, int *__restrict__ chunk_lengths,
                                long long *__restrict__ len64) {
    long c = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (c >= n_chunks) return;
    int mx = 0;
    for (int i = 0; i < C; ++i) mx = max(mx, rc[c * C + i].cnt);
    chunk_lengths[c] = mx;
    len64[c] = (long long)mx * C;
}

__global__ void k_narrow_ptrs(const long long *__restrict__ ptr64, long n, int *__restrict__ ptr32) {
    long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i < n) ptr32[i] = (int)ptr64[i];
}

__global__ void k_perms(const RC *__restrict__ rc, long n_rows, long n_pad, bool identity, int *__restrict__ old_to_new,
                        int *__restrict__ new_to_old, int *__restrict__ row_lengths) {
    long p = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (p >= n_pad) return;
    row_lengths[p] = rc[p].cnt;
    if (identity) {  // fixed_permutation mode: the struct's own permutation is the identity
        if (p < n_rows) old_to_new[p] = (int)p;
        new_to_old[p] = p < n_rows ? (int)p : -1;
        return;
    }
    int old = rc[p].idx;
    if (old >= 0 && old < n_rows) {
        old_to_new[old] = (int)p;
        new_to_old[p] = old;
    } else
        new_to_old[p] = -1;
}

// adopted matrices: pass 0 clears new_to_old and sets row_lengths to the chunk length (padding is unknown: every slot counts),
// pass 1 scatters the inverse permutation
__global__ void k_adopt_perms(const int *__restrict__ old_to_new, long n_rows, long n_pad, const int *__restrict__ chunk_lengths, int C,
                              int *__restrict__ new_to_old, int *__restrict__ row_lengths, int pass) {
    long p = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (p >= n_pad) return;
    if (pass == 0) {
        new_to_old[p] = -1;
        row_lengths[p] = chunk_lengths[p / C];
    } else if (p < n_rows) {
        const int q = old_to_new[p];
        if (q >= 0 && q < n_pad) new_to_old[q] = (int)p;
    }
}

template <typename T> struct Cvt;
template <> struct Cvt<double> {
    __device__ static double from(double v) { return v; }
    __device__ static double from(float v) { return (double)v; }
    __device__ static double from(__half v) { return (double)__half2float(v); }
};
template <> struct Cvt<float> {
    __device__ static float from(double v) { return (float)v; }
    __device__ static float from(float v) { return v; }
    __device__ static float from(__half v) { return __half2float(v); }
};
template <> struct Cvt<__half> {
    __device__ static __half from(double v) { return __double2half(v); }  // single rounding, like static_cast<_Float16>(double)
    __device__ static __half from(float v) { return __float2half_rn(v); }
    __device__ static __half from(__half v) { return v; }
};

// one thread per padded row position; lanes of a chunk are adjacent threads => coalesced stores
template <typename MT, typename VT>
__global__ void k_fill(const RC *__restrict__ rc, const int *__restrict__ row_ptr, const int *__restrict__ order,
                       const int *__restrict__ J, const MT *__restrict__ vals, const int *__restrict__ chunk_ptrs,
                       const int *__restrict__ chunk_lengths, long n_pad, int C, int *__restrict__ col_idxs,
                       VT *__restrict__ values) {
    long p = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (p >= n_pad) return;
    long c = p / C;
    int lane = (int)(p - c * C);
    int len = chunk_lengths[c];
    long e = (long)chunk_ptrs[c] + lane;
    RC r = rc[p];
    int src0 = (r.idx >= 0 && r.cnt > 0) ? row_ptr[r.idx] : 0;
    for (int j = 0; j < len; ++j, e += C) {
        int col = 0;
        VT v = Cvt<VT>::from(MT(0.0));
        if (j < r.cnt) {
            int s = src0 + j;
            if (order) s = order[s];
            col = J[s];
            v = Cvt<VT>::from(vals[s]);
        }
        col_idxs[e] = col;
        values[e] = v;
    }
}

__global__ void k_permute_cols(int *__restrict__ col_idxs, long n_elements, int n_rows, const int *__restrict__ perm) {
    long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i >= n_elements) return;
    int c = col_idxs[i];
    if (c < n_rows) col_idxs[i] = perm[c];
}

// ---- stencil generator ------------------------------------------------------------------------
__device__ __forceinline__ int axis_cnt(long v, long n) { return 1 + (v > 0) + (v < n - 1); }

__global__ void k_stencil_counts(int points, long nx, long ny, long nz, long row0, long n_local, long long *cnt) {
    long r = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (r >= n_local) return;
    long g = row0 + r;
    long x = g % nx, y = (g / nx) % ny, z = g / (nx * ny);
    int cx = axis_cnt(x, nx), cy = axis_cnt(y, ny), cz = axis_cnt(z, nz);
    cnt[r] = points == 7 ? (cx + cy + cz - 2) : (long long)cx * cy * cz;
}

__global__ void k_stencil_fill(int points, long nx, long ny, long nz, long row0, long n_local, const long long *__restrict__ ptr,
                               int *__restrict__ I, int *__restrict__ J, double *__restrict__ V) {
    long r = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (r >= n_local) return;
    long g = row0 + r;
    long x = g % nx, y = (g / nx) % ny, z = g / (nx * ny);
    long o = ptr[r];
    for (int dz = -1; dz <= 1; ++dz) {
        long zz = z + dz;
        if (zz < 0 || zz >= nz) continue;
        for (int dy = -1; dy <= 1; ++dy) {
            long yy = y + dy;
            if (yy < 0 || yy >= ny) continue;
            for (int dx = -1; dx <= 1; ++dx) {
                long xx = x + dx;
                if (xx < 0 || xx >= nx) continue;
                int nzd = (dx != 0) + (dy != 0) + (dz != 0);
                if (points == 7 && nzd > 1) continue;
                I[o] = (int)r;
                J[o] = (int)((zz * ny + yy) * nx + xx);
                V[o] = nzd == 0 ? (double)(points - 1) : -1.0;
                ++o;
            }
        }
    }
}

// ---- power-law generator (BASELINE.json config 4) ---------------------------------------------
// Row degree d = clamp(floor(d_min (1-u)^(-1/(alpha-1))), 1, max_deg); element k of row g: h = splitmix64(splitmix64(seed + g) ^
// k*golden); even k: column within +-1024 of the diagonal (wrapping), odd k: uniform over [0, n); value sign * 10^w, w ~ U(-4, 2).
// All randomness is a pure function of (seed, g, k), so any row range is generated independently (every rank its own rows) and
// the host generator matrices.powerlaw_coo produces the same matrix.
__device__ __forceinline__ unsigned long long splitmix64(unsigned long long z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__device__ __forceinline__ double u01(unsigned long long h) { return (double)(h >> 11) * (1.0 / 9007199254740992.0); }

__global__ void k_powerlaw_deg(long n, long row0, long n_local, double d_min, double alpha, int max_deg, unsigned long long seed,
                               long long *__restrict__ deg) {
    long r = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (r >= n_local) return;
    const unsigned long long g = (unsigned long long)(row0 + r);
    const double u = u01(splitmix64(seed ^ (g * 0xD1342543DE82EF95ull)));
    double d = floor(d_min * pow(1.0 - u, -1.0 / (alpha - 1.0)));
    d = fmin(fmax(d, 1.0), (double)max_deg);
    long long di = (long long)d;
    if (di > n) di = n;
    deg[r] = di;
}

// one warp per row: raw (un-deduplicated) elements, key = local row << 32 | column; generation order k ascending
__global__ void k_powerlaw_raw(long n, long row0, long n_local, unsigned long long seed, const long long *__restrict__ ptr,
                               unsigned long long *__restrict__ keys, double *__restrict__ vals) {
    const long w = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= n_local) return;
    const unsigned long long g = (unsigned long long)(row0 + w);
    const long long b = ptr[w], e = ptr[w + 1];
    const unsigned long long hg = splitmix64(seed + g);
    for (long long k = lane; k < e - b; k += 32) {
        const unsigned long long h = splitmix64(hg ^ ((unsigned long long)k * 0x9E3779B97F4A7C15ull));
        long long col;
        if ((k & 1) == 0) {
            col = ((long long)g + (long long)(h % 2049ull) - 1024) % n;
            if (col < 0) col += n;  // numpy's % is a floored modulo
        } else
            col = (long long)(h % (unsigned long long)n);
        const unsigned long long h2 = splitmix64(h);
        const double wexp = __dsub_rn(__dmul_rn(u01(h2), 6.0), 4.0);  // no FMA contraction: the host generator rounds twice
        const double sign = (h2 & 1ull) == 0 ? 1.0 : -1.0;
        keys[b + k] = ((unsigned long long)w << 32) | (unsigned long long)col;
        vals[b + k] = sign * pow(10.0, wexp);
    }
}

__global__ void k_flag_first(const unsigned long long *__restrict__ keys, long long total, int *__restrict__ flag) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < total) flag[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1 : 0;
    else if (i == total) flag[i] = 0;
}

__global__ void k_compact_keys(const unsigned long long *__restrict__ keys, const double *__restrict__ vals, const int *__restrict__ flag,
                               const int *__restrict__ pos, long long total, int *__restrict__ I, int *__restrict__ J, double *__restrict__ V) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= total || !flag[i]) return;
    const int o = pos[i];
    I[o] = (int)(keys[i] >> 32);
    J[o] = (int)(keys[i] & 0xffffffffull);
    V[o] = vals[i];
}

// ---- slab extraction (seg_mtx_struct + localize_row_idx, mpi_funcs.hpp:636-674,862-877) ---------------------------------
// lower_bound of two row ids in the row-sorted I array, one thread each
__global__ void k_row_bounds(const int *__restrict__ I, long nnz, int r0, int r1, long long *__restrict__ out2) {
    const int t = threadIdx.x;
    if (t > 1) return;
    const int key = t == 0 ? r0 : r1;
    long lo = 0, hi = nnz;
    while (lo < hi) {
        const long mid = (lo + hi) >> 1;
        if (I[mid] < key) lo = mid + 1; else hi = mid;
    }
    out2[t] = lo;
}
__global__ void k_localize_rows(const int *__restrict__ I, long n, int first_row, int *__restrict__ out, int *__restrict__ new_row_flag) {
    long k = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (k >= n) return;
    out[k] = I[k] - first_row;
    new_row_flag[k] = (k == 0 || I[k] != I[k - 1]) ? 1 : 0;
}

// ---- host helpers --------------------------------------------------------------------------------
void exclusive_scan_i64(const long long *in, long long *out, long n, cudaStream_t st) {
    size_t bytes = 0;
    USPMV_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, (int)n, st));
    DevBuf<unsigned char> tmp(bytes);
    USPMV_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, bytes, in, out, (int)n, st));
    g_launches.fetch_add(2);
    USPMV_CUDA(cudaStreamSynchronize(st));
}

template <typename MT, typename VT>
void launch_fill(const uspmv_coo *coo, uspmv_scs *s, const RC *rc, const int *row_ptr, const int *order) {
    k_fill<MT, VT><<<blocks_for(s->n_rows_padded), TPB>>>(rc, row_ptr, order, coo->J.p, reinterpret_cast<const MT *>(coo->values.p),
                                                        s->chunk_ptrs.p, s->chunk_lengths.p, s->n_rows_padded, (int)s->C,
                                                        s->col_idxs.p, reinterpret_cast<VT *>(s->values.p));
    USPMV_LAUNCH_CHECK();
}

template <typename MT>
void dispatch_fill_vt(const uspmv_coo *coo, uspmv_scs *s, const RC *rc, const int *row_ptr, const int *order) {
    switch (s->vt) {
    case USPMV_F64: launch_fill<MT, double>(coo, s, rc, row_ptr, order); break;
    case USPMV_F32: launch_fill<MT, float>(coo, s, rc, row_ptr, order); break;
    default: launch_fill<MT, __half>(coo, s, rc, row_ptr, order); break;
    }
}

// Longest-first chunk order for matrices with very uneven chunk lengths (power-law rows): the kernels give one warp a whole
// chunk (that keeps the per-row summation order of the reference), so a few 4096-slot chunks at the end of an interleaved
// assignment would leave the GPU idle behind them.
void build_balanced_order(uspmv_scs *s) {
    const long nc = s->n_chunks;
    if (nc < 4096 || s->n_elements == 0) return;
    std::vector<int> len(nc);
    USPMV_CUDA(cudaMemcpy(len.data(), s->chunk_lengths.p, nc * sizeof(int), cudaMemcpyDeviceToHost));
    int mx = 0;
    for (int v : len) mx = std::max(mx, v);
    const double mean = (double)s->n_elements / (double)(nc * s->C);
    if (mx <= 64 || mx < 8.0 * mean) return;
    DevBuf<int> iota(nc), keys_out(nc);
    s->balanced_order.alloc(nc);
    k_iota<<<blocks_for(nc), TPB>>>(iota.p, nc);
    USPMV_LAUNCH_CHECK();
    size_t bytes = 0;
    USPMV_CUDA(cub::DeviceRadixSort::SortPairsDescending(nullptr, bytes, s->chunk_lengths.p, keys_out.p, iota.p, s->balanced_order.p, (int)nc));
    DevBuf<unsigned char> tmp(bytes);
    USPMV_CUDA(cub::DeviceRadixSort::SortPairsDescending(tmp.p, bytes, s->chunk_lengths.p, keys_out.p, iota.p, s->balanced_order.p, (int)nc));
    g_launches.fetch_add(8);
    USPMV_CUDA(cudaDeviceSynchronize());

    // virtual items: long chunks cut into segments of <= L slots, all items longest first
    const int L = options().split_long_chunks;
    if (s->C != 32 || L <= 0 || mx <= L) return;
    std::vector<int> order(nc);
    USPMV_CUDA(cudaMemcpy(order.data(), s->balanced_order.p, nc * sizeof(int), cudaMemcpyDeviceToHost));
    std::vector<int4> items;
    std::vector<int> split_chunk, split_ptr{0};
    items.reserve(nc + nc / 8);
    for (long k = 0; k < nc; ++k) {
        const int c = order[k], len_c = len[c];
        if (len_c > L) {
            for (int j0 = 0; j0 < len_c; j0 += L) items.push_back(make_int4(c, j0, std::min(L, len_c - j0), (int)(split_ptr.back() + j0 / L)));
            split_chunk.push_back(c);
            split_ptr.push_back(split_ptr.back() + (len_c + L - 1) / L);
        } else
            items.push_back(make_int4(c, 0, len_c, -1));
    }
    std::stable_sort(items.begin(), items.end(), [](const int4 &a, const int4 &b) { return a.z > b.z; });
    s->n_vitems = (long)items.size();
    s->n_split = (long)split_chunk.size();
    s->vitems.alloc(items.size());
    s->split_chunk.alloc(split_chunk.size());
    s->split_ptr.alloc(split_ptr.size());
    s->partials.alloc((size_t)split_ptr.back() * 32 * vt_size(s->vt));
    USPMV_CUDA(cudaMemcpy(s->vitems.p, items.data(), items.size() * sizeof(int4), cudaMemcpyHostToDevice));
    USPMV_CUDA(cudaMemcpy(s->split_chunk.p, split_chunk.data(), split_chunk.size() * sizeof(int), cudaMemcpyHostToDevice));
    USPMV_CUDA(cudaMemcpy(s->split_ptr.p, split_ptr.data(), split_ptr.size() * sizeof(int), cudaMemcpyHostToDevice));
}

// ---- matrix ingest (read_mtx, utilities.hpp:2148-2309): symmetric expansion and the stable sort by row, on the device ----
__global__ void k_sym_count(const int *__restrict__ I, const int *__restrict__ J, long nz, int *__restrict__ cnt) {
    long k = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (k < nz) cnt[k] = 1 + (I[k] != J[k]);
    else if (k == nz) cnt[k] = 0;
}
// entry k -> (i, j, v) at pos[k], immediately followed by (j, i, v) when off-diagonal (utilities.hpp:2237-2251)
__global__ void k_sym_expand(const int *__restrict__ I, const int *__restrict__ J, const double *__restrict__ V, long nz,
                             const int *__restrict__ pos, int *__restrict__ Io, int *__restrict__ Jo, double *__restrict__ Vo) {
    long k = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (k >= nz) return;
    const int i = I[k], j = J[k], p = pos[k];
    const double v = V[k];
    Io[p] = i; Jo[p] = j; Vo[p] = v;
    if (i != j) { Io[p + 1] = j; Jo[p + 1] = i; Vo[p + 1] = v; }
}
__global__ void k_gather_coo(const int *__restrict__ order, const int *__restrict__ J, const double *__restrict__ V, long nnz,
                             int *__restrict__ Jo, double *__restrict__ Vo) {
    long k = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (k >= nnz) return;
    const int o = order[k];
    Jo[k] = J[o];
    Vo[k] = V[o];
}

// ---- equilibration (equilibrate_matrix, utilities.hpp:2605-2684): |v| >= 0, so its IEEE bit pattern orders like an integer ----
__global__ void k_absmax_by(const int *__restrict__ idx, const double *__restrict__ V, long nnz, unsigned long long *__restrict__ mx) {
    long k = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (k >= nnz) return;
    const double a = fabs(V[k]);
    if (a == a) atomicMax(&mx[idx[k]], (unsigned long long)__double_as_longlong(a));
}
__global__ void k_scale_by(const int *__restrict__ idx, double *__restrict__ V, long nnz, const double *__restrict__ mx) {
    long k = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (k < nnz) V[k] = V[k] / mx[idx[k]];
}

void check_flags(DevBuf<int> &flags, int out[2]) {
    USPMV_CUDA(cudaMemcpy(out, flags.p, 2 * sizeof(int), cudaMemcpyDeviceToHost));
}

}  // namespace

extern "C" {

// ---------------------------------------------------------------------------------------------
// COO
// ---------------------------------------------------------------------------------------------
static void coo_common(uspmv_ctx *ctx, long n_rows, long n_cols, long nnz, int mt, uspmv_coo *c) {
    if (!ctx) fail("coo: ctx is NULL");
    if (n_rows < 0 || n_cols < 0 || nnz < 0) fail("coo: negative dimension");
    if (n_rows > INT32_MAX - 1024 || nnz > INT32_MAX) fail("coo: n_rows/nnz exceed the reference's int index type (per rank)");
    vt_size(mt);
    USPMV_CUDA(cudaSetDevice(ctx->device));
    c->ctx = ctx;
    c->n_rows = n_rows;
    c->n_cols = n_cols;
    c->nnz = nnz;
    c->mt = mt;
    c->I.alloc(nnz);
    c->J.alloc(nnz);
    c->values.alloc(nnz * vt_size(mt));
}

int uspmv_coo_from_host(uspmv_ctx *ctx, long n_rows, long n_cols, long nnz, const int *I_h, const int *J_h, const void *values_h,
                        int mt, uspmv_coo **out) {
    return guarded([&] {
        OptScope opt_scope(ctx_of(ctx));  // this context's options govern everything below
        if (!out) fail("uspmv_coo_from_host: out is NULL");
        if (nnz > 0 && (!I_h || !J_h || !values_h)) fail("uspmv_coo_from_host: NULL array");
        auto c = new uspmv_coo();
        try {
            coo_common(ctx, n_rows, n_cols, nnz, mt, c);
            if (nnz) {
                USPMV_CUDA(cudaMemcpy(c->I.p, I_h, nnz * sizeof(int), cudaMemcpyHostToDevice));
                USPMV_CUDA(cudaMemcpy(c->J.p, J_h, nnz * sizeof(int), cudaMemcpyHostToDevice));
                USPMV_CUDA(cudaMemcpy(c->values.p, values_h, nnz * vt_size(mt), cudaMemcpyHostToDevice));
            }
        } catch (...) { delete c; throw; }
        *out = c;
    });
}

int uspmv_coo_from_device(uspmv_ctx *ctx, long n_rows, long n_cols, long nnz, const int *I_d, const int *J_d, const void *values_d,
                          int mt, uspmv_coo **out) {
    return guarded([&] {
        OptScope opt_scope(ctx_of(ctx));  // this context's options govern everything below
        if (!out) fail("uspmv_coo_from_device: out is NULL");
        if (nnz > 0 && (!I_d || !J_d || !values_d)) fail("uspmv_coo_from_device: NULL array");
        auto c = new uspmv_coo();
        try {
            coo_common(ctx, n_rows, n_cols, nnz, mt, c);
            if (nnz) {
                USPMV_CUDA(cudaMemcpy(c->I.p, I_d, nnz * sizeof(int), cudaMemcpyDeviceToDevice));
                USPMV_CUDA(cudaMemcpy(c->J.p, J_d, nnz * sizeof(int), cudaMemcpyDeviceToDevice));
                USPMV_CUDA(cudaMemcpy(c->values.p, values_d, nnz * vt_size(mt), cudaMemcpyDeviceToDevice));
            }
        } catch (...) { delete c; throw; }
        *out = c;
    });
}

int uspmv_coo_stencil(uspmv_ctx *ctx, int points, long nx, long ny, long nz, long row0, long row1, uspmv_coo **out) {
    return guarded([&] {
        OptScope opt_scope(ctx_of(ctx));  // this context's options govern everything below
        if (!ctx || !out) fail("uspmv_coo_stencil: NULL argument");
        if (points != 7 && points != 27) fail("uspmv_coo_stencil: points must be 7 or 27");
        long n = nx * ny * nz;
        if (nx <= 0 || ny <= 0 || nz <= 0 || row0 < 0 || row1 > n || row0 > row1) fail("uspmv_coo_stencil: bad grid / row range");
        if (n > INT32_MAX) fail("uspmv_coo_stencil: grid exceeds int column indices");
        USPMV_CUDA(cudaSetDevice(ctx->device));
        long n_local = row1 - row0;
        DevBuf<long long> cnt(n_local + 1), ptr(n_local + 1);
        USPMV_CUDA(cudaMemset(cnt.p, 0, (n_local + 1) * sizeof(long long)));
        if (n_local) {
            k_stencil_counts<<<blocks_for(n_local), TPB>>>(points, nx, ny, nz, row0, n_local, cnt.p);
            USPMV_LAUNCH_CHECK();
        }
        exclusive_scan_i64(cnt.p, ptr.p, n_local + 1, 0);
        long long nnz = 0;
        USPMV_CUDA(cudaMemcpy(&nnz, ptr.p + n_local, sizeof(long long), cudaMemcpyDeviceToHost));
        auto c = new uspmv_coo();
        try {
            coo_common(ctx, n_local, n, (long)nnz, USPMV_F64, c);
            if (n_local) {
                k_stencil_fill<<<blocks_for(n_local), TPB>>>(points, nx, ny, nz, row0, n_local, ptr.p, c->I.p, c->J.p,
                                                            reinterpret_cast<double *>(c->values.p));
                USPMV_LAUNCH_CHECK();
            }
            USPMV_CUDA(cudaDeviceSynchronize());
        } catch (...) { delete c; throw; }
        *out = c;
    });
}

/* BASELINE.json config 4's irregular matrix, rows [row0, row1) generated on the device (local row ids, global columns; ascending
 * de-duplicated columns per row, the first occurrence of a duplicate kept).  d_min sets the density (about 3.3 gives ~14.9 stored
 * elements per row = 5.0e8 at 2^25 rows).  Same matrix as matrices.powerlaw_coo (host). */
int uspmv_coo_powerlaw(uspmv_ctx *ctx, long n, long row0, long row1, double d_min, double alpha, int max_deg, unsigned long seed,
                       uspmv_coo **out) {
    return guarded([&] {
        OptScope opt_scope(ctx_of(ctx));  // this context's options govern everything below
        if (!ctx || !out) fail("uspmv_coo_powerlaw: NULL argument");
        if (n <= 0 || n > INT32_MAX - 1024 || row0 < 0 || row1 > n || row0 > row1) fail("uspmv_coo_powerlaw: bad size / row range");
        if (!(alpha > 1.0) || !(d_min > 0.0) || max_deg < 1) fail("uspmv_coo_powerlaw: need alpha > 1, d_min > 0, max_deg >= 1");
        USPMV_CUDA(cudaSetDevice(ctx->device));
        const long n_local = row1 - row0;
        DevBuf<long long> deg(n_local + 1), ptr(n_local + 1);
        USPMV_CUDA(cudaMemset(deg.p, 0, (n_local + 1) * sizeof(long long)));
        if (n_local) {
            k_powerlaw_deg<<<blocks_for(n_local), TPB>>>(n, row0, n_local, d_min, alpha, max_deg, (unsigned long long)seed, deg.p);
            USPMV_LAUNCH_CHECK();
        }
        exclusive_scan_i64(deg.p, ptr.p, n_local + 1, 0);
        long long total = 0;
        USPMV_CUDA(cudaMemcpy(&total, ptr.p + n_local, sizeof(long long), cudaMemcpyDeviceToHost));
        if (total > INT32_MAX - 1024) fail("uspmv_coo_powerlaw: %lld elements exceed the reference's int index type (per rank)", total);
        auto c = new uspmv_coo();
        try {
            long nnz = 0;
            DevBuf<unsigned long long> keys(total), keys_s(total);
            DevBuf<double> vals(total), vals_s(total);
            DevBuf<int> flag(total + 1), pos(total + 1);
            if (total) {
                k_powerlaw_raw<<<blocks_for(n_local * 32), TPB>>>(n, row0, n_local, (unsigned long long)seed, ptr.p, keys.p, vals.p);
                USPMV_LAUNCH_CHECK();
                int col_bits = 1, row_bits = 1;
                while (col_bits < 32 && (1L << col_bits) < n) ++col_bits;
                while (row_bits < 31 && (1L << row_bits) < n_local) ++row_bits;
                size_t bytes = 0;  // LSD radix sort is stable: equal (row, column) keys keep their generation order k
                USPMV_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, keys.p, keys_s.p, vals.p, vals_s.p, (int)total, 0, 32 + row_bits));
                DevBuf<unsigned char> tmp(bytes);
                USPMV_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, bytes, keys.p, keys_s.p, vals.p, vals_s.p, (int)total, 0, 32 + row_bits));
                g_launches.fetch_add(8);
                (void)col_bits;
                k_flag_first<<<blocks_for(total + 1), TPB>>>(keys_s.p, total, flag.p);
                USPMV_LAUNCH_CHECK();
                size_t b2 = 0;
                USPMV_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, b2, flag.p, pos.p, (int)(total + 1)));
                DevBuf<unsigned char> tmp2(b2);
                USPMV_CUDA(cub::DeviceScan::ExclusiveSum(tmp2.p, b2, flag.p, pos.p, (int)(total + 1)));
                g_launches.fetch_add(2);
                int tot = 0;
                USPMV_CUDA(cudaMemcpy(&tot, pos.p + total, sizeof(int), cudaMemcpyDeviceToHost));
                nnz = tot;
            }
            coo_common(ctx, n_local, n, nnz, USPMV_F64, c);
            if (nnz) {
                k_compact_keys<<<blocks_for(total), TPB>>>(keys_s.p, vals_s.p, flag.p, pos.p, total, c->I.p, c->J.p,
                                                          reinterpret_cast<double *>(c->values.p));
                USPMV_LAUNCH_CHECK();
            }
            USPMV_CUDA(cudaDeviceSynchronize());
        } catch (...) { delete c; throw; }
        *out = c;
    });
}

/* seg_mtx_struct + localize_row_idx (mpi_funcs.hpp:636-674,862-877) on the device: the rows [wsa[rank], wsa[rank+1]) of a ROW-SORTED
 * COO as a new COO with process-local row ids and GLOBAL columns (input order kept).  n_distinct_rows (optional) receives the number
 * of distinct rows present, which is what the reference stores as local n_rows (mpi_funcs.hpp:770); the returned COO has
 * n_rows = wsa[rank+1] - wsa[rank] (the two differ only when the slab contains empty rows, where the reference's n_rows no longer
 * covers its own row ids).  The reference makes rows local by subtracting the FIRST STORED row id and finds the slab with a linear
 * search for the row ids wsa[rank] / wsa[rank+1] (get_index, utilities.hpp:855-878: -1 when such a row is empty); here local row =
 * row - wsa[rank], identical whenever the reference is defined. */
int uspmv_coo_seg_mtx(const uspmv_coo *total, const int *wsa_h, int rank, int P, uspmv_coo **out, long *n_distinct_rows) {
    return guarded([&] {
        OptScope opt_scope(ctx_of(total));  // this context's options govern everything below
        if (!total || !wsa_h || !out) fail("uspmv_coo_seg_mtx: NULL argument");
        if (P < 1 || rank < 0 || rank >= P) fail("uspmv_coo_seg_mtx: bad rank / comm_size");
        const int r0 = wsa_h[rank], r1 = wsa_h[rank + 1];
        if (r0 < 0 || r1 < r0 || r1 > total->n_rows) fail("uspmv_coo_seg_mtx: work_sharing_arr [%d, %d) outside the matrix (%ld rows)", r0, r1, total->n_rows);
        USPMV_CUDA(cudaSetDevice(total->ctx->device));
        const long nnz = total->nnz;
        if (nnz) {
            DevBuf<int> flags(2);
            USPMV_CUDA(cudaMemset(flags.p, 0, 2 * sizeof(int)));
            k_check_sorted<<<blocks_for(nnz), TPB>>>(total->I.p, nnz, total->n_rows, flags.p);
            USPMV_LAUNCH_CHECK();
            int hf[2];
            check_flags(flags, hf);
            if (hf[0] || hf[1]) fail("uspmv_coo_seg_mtx: the COO matrix must be sorted by row (read_mtx / uspmv_coo_from_entries order)");
        }
        long long b[2] = {0, 0};
        if (nnz) {
            DevBuf<long long> bd(2);
            k_row_bounds<<<1, 32>>>(total->I.p, nnz, r0, r1, bd.p);
            USPMV_LAUNCH_CHECK();
            USPMV_CUDA(cudaMemcpy(b, bd.p, 2 * sizeof(long long), cudaMemcpyDeviceToHost));
        }
        const long lo = (long)b[0], n = (long)(b[1] - b[0]);
        const size_t es = vt_size(total->mt);
        auto c = new uspmv_coo();
        try {
            coo_common(total->ctx, r1 - r0, total->n_cols, n, total->mt, c);
            long distinct = 0;
            if (n) {
                DevBuf<int> flag(n + 1), pos(n + 1);
                USPMV_CUDA(cudaMemset(flag.p + n, 0, sizeof(int)));
                k_localize_rows<<<blocks_for(n), TPB>>>(total->I.p + lo, n, r0, c->I.p, flag.p);
                USPMV_LAUNCH_CHECK();
                USPMV_CUDA(cudaMemcpy(c->J.p, total->J.p + lo, n * sizeof(int), cudaMemcpyDeviceToDevice));
                USPMV_CUDA(cudaMemcpy(c->values.p, total->values.p + (size_t)lo * es, n * es, cudaMemcpyDeviceToDevice));
                size_t bytes = 0;
                USPMV_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, bytes, flag.p, pos.p, (int)(n + 1)));
                DevBuf<unsigned char> tmp(bytes);
                USPMV_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, bytes, flag.p, pos.p, (int)(n + 1)));
                g_launches.fetch_add(2);
                int tot = 0;
                USPMV_CUDA(cudaMemcpy(&tot, pos.p + n, sizeof(int), cudaMemcpyDeviceToHost));
                distinct = tot;
            }
            if (n_distinct_rows) *n_distinct_rows = distinct;
            USPMV_CUDA(cudaDeviceSynchronize());
        } catch (...) { delete c; throw; }
        *out = c;
    });
}

/* raw device pointers of the COO arrays (any out pointer may be NULL); values are `mt`-typed */
int uspmv_coo_device_arrays(const uspmv_coo *coo, const int **I_d, const int **J_d, const void **values_d) {
    return guarded([&] {
        OptScope opt_scope(ctx_of(coo));  // this context's options govern everything below
        if (!coo) fail("uspmv_coo_device_arrays: coo is NULL");
        if (I_d) *I_d = coo->I.p;
        if (J_d) *J_d = coo->J.p;
        if (values_d) *values_d = coo->values.p;
    });
}

int uspmv_coo_dims(const uspmv_coo *coo, long out3[3]) {
    return guarded([&] {
        OptScope opt_scope(ctx_of(coo));  // this context's options govern everything below
        if (!coo || !out3) fail("uspmv_coo_dims: NULL argument");
        out3[0] = coo->n_rows; out3[1] = coo->n_cols; out3[2] = coo->nnz;
    });
}

/* read_mtx's post-processing on the device (utilities.hpp:2214-2290): the nz (row, col, value) entries of a Matrix Market file in
 * file order, 0-based; symmetric != 0 expands every off-diagonal entry (i,j) into (i,j),(j,i); then a STABLE sort by row. */
int uspmv_coo_from_entries(uspmv_ctx *ctx, long n_rows, long n_cols, long nz, const int *I_h, const int *J_h, const double *V_h,
                           int symmetric, uspmv_coo **out) {
    return guarded([&] {
        OptScope opt_scope(ctx_of(ctx));  // this context's options govern everything below
        if (!out) fail("uspmv_coo_from_entries: out is NULL");
        if (nz > 0 && (!I_h || !J_h || !V_h)) fail("uspmv_coo_from_entries: NULL array");
        if (!ctx) fail("uspmv_coo_from_entries: ctx is NULL");
        if (n_rows < 0 || n_cols < 0 || nz < 0 || nz > INT32_MAX / 2) fail("uspmv_coo_from_entries: bad dimensions");
        USPMV_CUDA(cudaSetDevice(ctx->device));
        for (long k = 0; k < nz; ++k)
            if (I_h[k] < 0 || I_h[k] >= n_rows || J_h[k] < 0 || J_h[k] >= n_cols || (symmetric && (J_h[k] >= n_rows || I_h[k] >= n_cols)))
                fail("uspmv_coo_from_entries: entry %ld (%d, %d) outside the %ld x %ld matrix", k, I_h[k], J_h[k], n_rows, n_cols);
        DevBuf<int> I0(nz), J0(nz);
        DevBuf<double> V0(nz);
        if (nz) {
            USPMV_CUDA(cudaMemcpy(I0.p, I_h, nz * sizeof(int), cudaMemcpyHostToDevice));
            USPMV_CUDA(cudaMemcpy(J0.p, J_h, nz * sizeof(int), cudaMemcpyHostToDevice));
            USPMV_CUDA(cudaMemcpy(V0.p, V_h, nz * sizeof(double), cudaMemcpyHostToDevice));
        }
        long nnz = nz;
        DevBuf<int> I1, J1;
        DevBuf<double> V1;
        const int *Iu = I0.p, *Ju = J0.p;
        const double *Vu = V0.p;
        if (symmetric && nz) {
            DevBuf<int> cnt(nz + 1), pos(nz + 1);
            k_sym_count<<<blocks_for(nz + 1), TPB>>>(I0.p, J0.p, nz, cnt.p);
            USPMV_LAUNCH_CHECK();
            size_t bytes = 0;
            USPMV_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, bytes, cnt.p, pos.p, (int)(nz + 1)));
            DevBuf<unsigned char> tmp(bytes);
            USPMV_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, bytes, cnt.p, pos.p, (int)(nz + 1)));
            g_launches.fetch_add(2);
            int tot = 0;
            USPMV_CUDA(cudaMemcpy(&tot, pos.p + nz, sizeof(int), cudaMemcpyDeviceToHost));
            nnz = tot;
            I1.alloc(nnz); J1.alloc(nnz); V1.alloc(nnz);
            k_sym_expand<<<blocks_for(nz), TPB>>>(I0.p, J0.p, V0.p, nz, pos.p, I1.p, J1.p, V1.p);
            USPMV_LAUNCH_CHECK();
            Iu = I1.p; Ju = J1.p; Vu = V1.p;
        }
        auto c = new uspmv_coo();
        try {
            coo_common(ctx, n_rows, n_cols, nnz, USPMV_F64, c);
            if (nnz) {
                DevBuf<int> iota(nnz), order(nnz);
                k_iota<<<blocks_for(nnz), TPB>>>(iota.p, nnz);
                USPMV_LAUNCH_CHECK();
                int end_bit = 1;
                while (end_bit < 31 && (1L << end_bit) < n_rows) ++end_bit;
                size_t bytes = 0;  // LSD radix sort: stable, i.e. std::stable_sort by row (sort_perm, utilities.hpp:2139-2146)
                USPMV_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, Iu, c->I.p, iota.p, order.p, (int)nnz, 0, end_bit));
                DevBuf<unsigned char> tmp(bytes);
                USPMV_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, bytes, Iu, c->I.p, iota.p, order.p, (int)nnz, 0, end_bit));
                g_launches.fetch_add(8);
                k_gather_coo<<<blocks_for(nnz), TPB>>>(order.p, Ju, Vu, nnz, c->J.p, reinterpret_cast<double *>(c->values.p));
                USPMV_LAUNCH_CHECK();
                USPMV_CUDA(cudaDeviceSynchronize());
            }
        } catch (...) { delete c; throw; }
        *out = c;
    });
}

/* equilibrate_matrix (utilities.hpp:2668-2684) in place on a dp COO: every value divided by the largest |value| of its row, then
 * by the largest |scaled value| of its column.  rowmax_h / colmax_h (optional, n_rows / n_cols doubles) receive the two maxima —
 * what the harness hands to partition_precisions in AP mode (main.cpp:1143-1153). */
int uspmv_coo_equilibrate(uspmv_coo *coo, double *rowmax_h, double *colmax_h) {
    return guarded([&] {
        OptScope opt_scope(ctx_of(coo));  // this context's options govern everything below
        if (!coo) fail("uspmv_coo_equilibrate: coo is NULL");
        if (coo->mt != USPMV_F64) fail("uspmv_coo_equilibrate: the COO matrix must hold doubles");
        USPMV_CUDA(cudaSetDevice(coo->ctx->device));
        const long nnz = coo->nnz;
        DevBuf<unsigned long long> rm(coo->n_rows > 0 ? coo->n_rows : 1), cm(coo->n_cols > 0 ? coo->n_cols : 1);
        USPMV_CUDA(cudaMemset(rm.p, 0, rm.n * 8));
        USPMV_CUDA(cudaMemset(cm.p, 0, cm.n * 8));
        double *V = reinterpret_cast<double *>(coo->values.p);
        if (nnz) {
            k_absmax_by<<<blocks_for(nnz), TPB>>>(coo->I.p, V, nnz, rm.p);
            USPMV_LAUNCH_CHECK();
            k_scale_by<<<blocks_for(nnz), TPB>>>(coo->I.p, V, nnz, reinterpret_cast<const double *>(rm.p));
            USPMV_LAUNCH_CHECK();
            k_absmax_by<<<blocks_for(nnz), TPB>>>(coo->J.p, V, nnz, cm.p);
            USPMV_LAUNCH_CHECK();
            k_scale_by<<<blocks_for(nnz), TPB>>>(coo->J.p, V, nnz, reinterpret_cast<const double *>(cm.p));
            USPMV_LAUNCH_CHECK();
        }
        if (rowmax_h && coo->n_rows) USPMV_CUDA(cudaMemcpy(rowmax_h, rm.p, coo->n_rows * 8, cudaMemcpyDeviceToHost));
        if (colmax_h && coo->n_cols) USPMV_CUDA(cudaMemcpy(colmax_h, cm.p, coo->n_cols * 8, cudaMemcpyDeviceToHost));
        USPMV_CUDA(cudaDeviceSynchronize());
    });
}

int uspmv_coo_export(const uspmv_coo *coo, int *I_h, int *J_h, void *values_h) {
    return guarded([&] {
        OptScope opt_scope(ctx_of(coo));  // this context's options govern everything below
        if (!coo) fail("uspmv_coo_export: coo is NULL");
        USPMV_CUDA(cudaSetDevice(coo->ctx->device));
        if (I_h && coo->nnz) USPMV_CUDA(cudaMemcpy(I_h, coo->I.p, coo->nnz * sizeof(int), cudaMemcpyDeviceToHost));
        if (J_h && coo->nnz) USPMV_CUDA(cudaMemcpy(J_h, coo->J.p, coo->nnz * sizeof(int), cudaMemcpyDeviceToHost));
        if (values_h && coo->nnz) USPMV_CUDA(cudaMemcpy(values_h, coo->values.p, coo->nnz * vt_size(coo->mt), cudaMemcpyDeviceToHost));
    });
}

void uspmv_coo_destroy(uspmv_coo *coo) { delete coo; }

// ---------------------------------------------------------------------------------------------
// convert_to_scs
// ---------------------------------------------------------------------------------------------
int uspmv_scs_build(uspmv_ctx *ctx, const uspmv_coo *coo, long C, long sigma, int vt, const int *fixed_perm_h, uspmv_scs **out) {
    return guarded([&] {
        OptScope opt_scope(ctx_of(ctx));  // this context's options govern everything below
        if (!ctx || !coo || !out) fail("uspmv_scs_build: NULL argument");
        if (C < 1 || sigma < 1) fail("uspmv_scs_build: C and sigma must be >= 1 (got C=%ld sigma=%ld)", C, sigma);
        vt_size(vt);
        USPMV_CUDA(cudaSetDevice(ctx->device));
        const long n_rows = coo->n_rows, nnz = coo->nnz;
        const long n_chunks = (n_rows + C - 1) / C;
        const long n_pad = n_chunks * C;
        if (n_pad > INT32_MAX - 1024) fail("uspmv_scs_build: no. of padded rows exceeds the index type");

        DevBuf<int> flags(2);
        USPMV_CUDA(cudaMemset(flags.p, 0, 2 * sizeof(int)));
        int hflags[2] = {0, 0};
        if (nnz) {
            k_check_sorted<<<blocks_for(nnz), TPB>>>(coo->I.p, nnz, n_rows, flags.p);
            USPMV_LAUNCH_CHECK();
            k_check_cols<<<blocks_for(nnz), TPB>>>(coo->J.p, nnz, flags.p);
            USPMV_LAUNCH_CHECK();
            check_flags(flags, hflags);
            if (hflags[1]) fail("uspmv_scs_build: row index outside [0, n_rows) or negative column index");
        }

        // stable sort by row only if needed; `order` maps sorted position -> input position
        DevBuf<int> Isorted, order;
        const int *Is = coo->I.p;
        const int *ord = nullptr;
        if (hflags[0]) {
            DevBuf<int> iota(nnz);
            Isorted.alloc(nnz);
            order.alloc(nnz);
            k_iota<<<blocks_for(nnz), TPB>>>(iota.p, nnz);
            USPMV_LAUNCH_CHECK();
            int end_bit = 1;
            while (end_bit < 31 && (1L << end_bit) < n_rows) ++end_bit;
            size_t bytes = 0;
            USPMV_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, coo->I.p, Isorted.p, iota.p, order.p, (int)nnz, 0, end_bit));
            DevBuf<unsigned char> tmp(bytes);
            USPMV_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, bytes, coo->I.p, Isorted.p, iota.p, order.p, (int)nnz, 0, end_bit));
            g_launches.fetch_add(8);
            USPMV_CUDA(cudaDeviceSynchronize());
            Is = Isorted.p;
            ord = order.p;
        }

        DevBuf<int> row_ptr(n_pad + 1);
        k_row_ptr<<<blocks_for(nnz + 1), TPB>>>(Is, nnz, n_rows + 1, row_ptr.p);
        USPMV_LAUNCH_CHECK();

        DevBuf<RC> rc(n_pad + 1);
        DevBuf<int> fixed_perm_d;
        if (fixed_perm_h) {
            fixed_perm_d.alloc(n_rows);
            USPMV_CUDA(cudaMemcpy(fixed_perm_d.p, fixed_perm_h, n_rows * sizeof(int), cudaMemcpyHostToDevice));
            if (n_pad) {
                k_fill_rc_empty<<<blocks_for(n_pad), TPB>>>(n_pad, rc.p);
                USPMV_LAUNCH_CHECK();
            }
            if (n_rows) {
                k_scatter_fixed<<<blocks_for(n_rows), TPB>>>(row_ptr.p, fixed_perm_d.p, n_rows, n_pad, rc.p, flags.p);
                USPMV_LAUNCH_CHECK();
                check_flags(flags, hflags);
                if (hflags[1] & 2) fail("uspmv_scs_build: fixed_permutation is not injective (two rows map to the same position)");
                if (hflags[1]) fail("uspmv_scs_build: fixed_permutation entry outside [0, n_rows_padded)");
            }
        } else if (n_pad) {
            k_init_rc<<<blocks_for(n_pad), TPB>>>(row_ptr.p, n_rows, n_pad, rc.p);
            USPMV_LAUNCH_CHECK();
            if (sigma > 1) {
                long n_windows = (n_pad + sigma - 1) / sigma;
                const size_t win_bytes = (size_t)std::min<long>(sigma, n_pad) * sizeof(RC);
                if (sigma >= 64 && win_bytes <= 200 * 1024) {  // window staged in shared memory, one CTA per window
                    static bool configured_on[uspmv::MAX_DEVICES] = {};
                    const int dev = uspmv::current_device();
                    if (!configured_on[dev]) {
                        USPMV_CUDA(cudaFuncSetAttribute(k_sort_windows_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
                        configured_on[dev] = true;
                    }
                    const long grid = std::min<long>(n_windows, 32L * sm_count(dev));
                    k_sort_windows_smem<<<(unsigned)grid, 64, win_bytes>>>(rc.p, n_pad, sigma, n_windows);
                } else  // tiny windows (insertion-sort regime) or windows beyond the shared-memory size: one thread per window
                    k_sort_windows<<<blocks_for(n_windows, 64), 64>>>(rc.p, n_pad, sigma, n_windows);
                USPMV_LAUNCH_CHECK();
            }
        }

        auto s = new uspmv_scs();
        try {
            s->ctx = ctx;
            s->C = C; s->sigma = sigma; s->n_rows = n_rows; s->n_cols = coo->n_cols;
            s->n_rows_padded = n_pad; s->n_chunks = n_chunks; s->nnz = nnz; s->vt = vt;
            s->x_min_len = coo->n_cols;
            s->chunk_ptrs.alloc(n_chunks + 1);
            s->chunk_lengths.alloc(n_chunks);
            s->old_to_new.alloc(n_rows);
            s->new_to_old.alloc(n_pad);
            s->row_lengths.alloc(n_pad);
            DevBuf<long long> len64(n_chunks + 1), ptr64(n_chunks + 1);
            USPMV_CUDA(cudaMemset(len64.p, 0, (n_chunks + 1) * sizeof(long long)));
            if (n_chunks) {
                k_chunk_lengths<<<blocks_for(n_chunks), TPB>>>(rc.p, n_chunks, (int)C, s->chunk_lengths.p, len64.p);
                USPMV_LAUNCH_CHECK();
            }
            exclusive_scan_i64(len64.p, ptr64.p, n_chunks + 1, 0);
            long long n_el = 0;
            USPMV_CUDA(cudaMemcpy(&n_el, ptr64.p + n_chunks, sizeof(long long), cudaMemcpyDeviceToHost));
            if (n_el > INT32_MAX) fail("uspmv_scs_build: chunk_ptrs exceed index type (n_elements = %lld)", n_el);
            s->n_elements = (long)n_el;
            k_narrow_ptrs<<<blocks_for(n_chunks + 1), TPB>>>(ptr64.p, n_chunks + 1, s->chunk_ptrs.p);
            USPMV_LAUNCH_CHECK();
            if (n_pad) {
                k_perms<<<blocks_for(n_pad), TPB>>>(rc.p, n_rows, n_pad, fixed_perm_h != nullptr, s->old_to_new.p, s->new_to_old.p, s->row_lengths.p);
                USPMV_LAUNCH_CHECK();
            }
            // 8 elements of slack: the streamed CRS kernel rounds its bulk copies up to 16-byte multiples
            s->col_idxs.alloc(s->n_elements + 8);
            s->values.alloc((s->n_elements + 8) * vt_size(vt));
            USPMV_CUDA(cudaMemset(s->col_idxs.p + s->n_elements, 0, 8 * sizeof(int)));
            USPMV_CUDA(cudaMemset(s->values.p + s->n_elements * vt_size(vt), 0, 8 * vt_size(vt)));
            if (n_pad && s->n_elements) {
                switch (coo->mt) {
                case USPMV_F64: dispatch_fill_vt<double>(coo, s, rc.p, row_ptr.p, ord); break;
                case USPMV_F32: dispatch_fill_vt<float>(coo, s, rc.p, row_ptr.p, ord); break;
                default: dispatch_fill_vt<__half>(coo, s, rc.p, row_ptr.p, ord); break;
                }
            }
            USPMV_CUDA(cudaDeviceSynchronize());
            build_balanced_order(s);
        } catch (...) { delete s; throw; }
        *out = s;
    });
}

/* Adopt a SELL-C-sigma matrix that somebody else built (e.g. the reference's own convert_to_scs on the host, or a file): the arrays
 * are COPIED into a device-resident handle that every kernel entry point accepts.  Host or device pointers (`on_device`).  old_to_new
 * may be NULL (identity).  chunk_ptrs has n_chunks + 1 entries, n_elements = chunk_ptrs[n_chunks]. */
int uspmv_scs_from_arrays(uspmv_ctx *ctx, int vt, long C, long sigma, long n_rows, long n_cols, long n_chunks, const int *chunk_ptrs,
                          const int *chunk_lengths, const int *col_idxs, const void *values, const int *old_to_new, int on_device,
                          uspmv_scs **out) {
    return guarded([&] {
        OptScope opt_scope(ctx_of(ctx));  // this context's options govern everything below
        if (!ctx || !out) fail("uspmv_scs_from_arrays: NULL argument");
        if (C < 1 || sigma < 1 || n_rows < 0 || n_chunks != (n_rows + C - 1) / C) fail("uspmv_scs_from_arrays: inconsistent C / n_rows / n_chunks");
        if (n_chunks > 0 && (!chunk_ptrs || !chunk_lengths)) fail("uspmv_scs_from_arrays: NULL chunk array");
        const size_t es = vt_size(vt);
        USPMV_CUDA(cudaSetDevice(ctx->device));
        const cudaMemcpyKind kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
        int n_el = 0;
        if (n_chunks) {
            if (on_device) USPMV_CUDA(cudaMemcpy(&n_el, chunk_ptrs + n_chunks, sizeof(int), cudaMemcpyDeviceToHost));
            else n_el = chunk_ptrs[n_chunks];
        }
        if (n_el < 0) fail("uspmv_scs_from_arrays: negative n_elements");
        if (n_el > 0 && (!col_idxs || !values)) fail("uspmv_scs_from_arrays: NULL element array");
        auto s = new uspmv_scs();
        try {
            s->ctx = ctx; s->C = C; s->sigma = sigma; s->n_rows = n_rows; s->n_cols = n_cols; s->n_rows_padded = n_chunks * C;
            s->n_chunks = n_chunks; s->n_elements = n_el; s->vt = vt;
            s->cols_permuted = true;  // whatever numbering the arrays carry is final
            s->x_min_len = n_cols;
            s->chunk_ptrs.alloc(n_chunks + 1);
            s->chunk_lengths.alloc(n_chunks);
            s->old_to_new.alloc(n_rows);
            s->new_to_old.alloc(s->n_rows_padded);
            s->row_lengths.alloc(s->n_rows_padded);
            s->col_idxs.alloc((size_t)n_el + 8);
            s->values.alloc(((size_t)n_el + 8) * es);
            USPMV_CUDA(cudaMemset(s->col_idxs.p + n_el, 0, 8 * sizeof(int)));
            USPMV_CUDA(cudaMemset(s->values.p + (size_t)n_el * es, 0, 8 * es));
            if (n_chunks) {
                USPMV_CUDA(cudaMemcpy(s->chunk_ptrs.p, chunk_ptrs, (n_chunks + 1) * sizeof(int), kind));
                USPMV_CUDA(cudaMemcpy(s->chunk_lengths.p, chunk_lengths, n_chunks * sizeof(int), kind));
            } else
                USPMV_CUDA(cudaMemset(s->chunk_ptrs.p, 0, sizeof(int)));
            if (n_el) {
                USPMV_CUDA(cudaMemcpy(s->col_idxs.p, col_idxs, (size_t)n_el * sizeof(int), kind));
                USPMV_CUDA(cudaMemcpy(s->values.p, values, (size_t)n_el * es, kind));
            }
            if (n_rows) {
                if (old_to_new) USPMV_CUDA(cudaMemcpy(s->old_to_new.p, old_to_new, n_rows * sizeof(int), kind));
                else {
                    k_iota<<<blocks_for(n_rows), TPB>>>(s->old_to_new.p, n_rows);
                    USPMV_LAUNCH_CHECK();
                }
            }
            if (s->n_rows_padded) {
                k_adopt_perms<<<blocks_for(s->n_rows_padded), TPB>>>(s->old_to_new.p, n_rows, s->n_rows_padded, s->chunk_lengths.p, (int)C,
                                                                    s->new_to_old.p, s->row_lengths.p, 0);
                USPMV_LAUNCH_CHECK();
                k_adopt_perms<<<blocks_for(s->n_rows_padded), TPB>>>(s->old_to_new.p, n_rows, s->n_rows_padded, s->chunk_lengths.p, (int)C,
                                                                    s->new_to_old.p, s->row_lengths.p, 1);
                USPMV_LAUNCH_CHECK();
            }
            // nnz is not recoverable without the row lengths; the stored-element count stands in (only used for reporting)
            s->nnz = n_el;
            USPMV_CUDA(cudaDeviceSynchronize());
            build_balanced_order(s);
        } catch (...) { delete s; throw; }
        *out = s;
    });
}

int uspmv_scs_dims(const uspmv_scs *s, long o[8]) {
    return guarded([&] {
        OptScope opt_scope(ctx_of(s));  // this context's options govern everything below
        if (!s || !o) fail("uspmv_scs_dims: NULL argument");
        o[0] = s->C; o[1] = s->sigma; o[2] = s->n_rows; o[3] = s->n_cols;
        o[4] = s->n_rows_padded; o[5] = s->n_chunks; o[6] = s->n_elements; o[7] = s->nnz;
    });
}

int uspmv_scs_export(const uspmv_scs *s, int *chunk_ptrs_h, int *chunk_lengths_h, int *col_idxs_h, void *values_h, int *old_to_new_h,
                     int *new_to_old_h) {
    return guarded([&] {
        OptScope opt_scope(ctx_of(s));  // this context's options govern everything below
        if (!s) fail("uspmv_scs_export: scs is NULL");
        USPMV_CUDA(cudaSetDevice(s->ctx->device));
        auto d2h = [](void *dst, const void *src, size_t bytes) {
            if (dst && bytes) USPMV_CUDA(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost));
        };
        d2h(chunk_ptrs_h, s->chunk_ptrs.p, (s->n_chunks + 1) * sizeof(int));
        d2h(chunk_lengths_h, s->chunk_lengths.p, s->n_chunks * sizeof(int));
        d2h(col_idxs_h, s->col_idxs.p, s->n_elements * sizeof(int));
        d2h(values_h, s->values.p, s->n_elements * vt_size(s->vt));
        d2h(old_to_new_h, s->old_to_new.p, s->n_rows * sizeof(int));
        d2h(new_to_old_h, s->new_to_old.p, s->n_rows_padded * sizeof(int));
    });
}

int uspmv_scs_permute_cols(uspmv_scs *s, const int *perm_h) {
    return guarded([&] {
        OptScope opt_scope(ctx_of(s));  // this context's options govern everything below
        if (!s) fail("uspmv_scs_permute_cols: scs is NULL");
        USPMV_CUDA(cudaSetDevice(s->ctx->device));
        DevBuf<int> perm_d;
        const int *perm = s->old_to_new.p;
        if (perm_h) {
            perm_d.alloc(s->n_rows);
            USPMV_CUDA(cudaMemcpy(perm_d.p, perm_h, s->n_rows * sizeof(int), cudaMemcpyHostToDevice));
            perm = perm_d.p;
        }
        if (s->n_elements) {
            k_permute_cols<<<blocks_for(s->n_elements), TPB>>>(s->col_idxs.p, s->n_elements, (int)s->n_rows, perm);
            USPMV_LAUNCH_CHECK();
        }
        USPMV_CUDA(cudaDeviceSynchronize());
        s->cols_permuted = true;
    });
}

int uspmv_scs_device_arrays(const uspmv_scs *s, const int **chunk_ptrs_d, const int **chunk_lengths_d, const int **col_idxs_d,
                            const void **values_d, const int **old_to_new_d, const int **new_to_old_d) {
    return guarded([&] {
        OptScope opt_scope(ctx_of(s));  // this context's options govern everything below
        if (!s) fail("uspmv_scs_device_arrays: scs is NULL");
        if (chunk_ptrs_d) *chunk_ptrs_d = s->chunk_ptrs.p;
        if (chunk_lengths_d) *chunk_lengths_d = s->chunk_lengths.p;
        if (col_idxs_d) *col_idxs_d = s->col_idxs.p;
        if (values_d) *values_d = s->values.p;
        if (old_to_new_d) *old_to_new_d = s->old_to_new.p;
        if (new_to_old_d) *new_to_old_d = s->new_to_old.p;
    });
}

void uspmv_scs_destroy(uspmv_scs *s) { delete s; }

}  // extern "C"
