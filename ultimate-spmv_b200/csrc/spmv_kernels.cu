// SpMV / SpMMV kernels for sm_100a and their C-ABI launchers.
//
// Replaces the reference's kernels
//   spmv_gpu_scs / spmv_gpu_scs_adv / scs_impl_gpu<C>   code/kernels.hpp:579-775
//   spmv_gpu_csr                                         code/kernels.hpp:631-680
//   block_spmv_gpu_*_launcher (stubs in the reference)   code/kernels.hpp:777-844
//   CPU semantics of the block kernels                   code/kernels.hpp:68-154,306-398
//
// Design (HBM-bound, CUDA cores — see DESIGN.md):
//   * SCS: one thread per padded row, lanes of a chunk are adjacent threads, so for C >= 32/sizeof-ratio
//     every `j` step of a warp is one fully coalesced 32*sizeof(VT) + 128 B request.  The value and
//     column streams are read exactly once with streaming (evict-first, no L1 allocate) loads, all
//     loads of an unrolled group of U slots are issued before the first use (memory-level
//     parallelism), x is gathered through the read-only path so neighbouring rows share L1/L2 lines.
//   * The per-row accumulation order is j = 0..len-1 with one fused multiply-add per slot — the same
//     order and rounding as the reference's CPU loop (kernels.hpp:242-246 built with -O3 on an FMA
//     machine), so dp/sp/hp results are bit-identical to the oracle, padding slots included.
//   * hp accumulates in fp16 with product+sum formed in fp32 and rounded once per step, which is what
//     g++ -std=c++23 emits for `_Float16 tmp += v * x` on x86 without AVX512-FP16.
#include "common.cuh"
#include "scs_stream.cuh"

#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

using namespace uspmv;

namespace {

constexpr int TPB = 256;

// ---- per-precision arithmetic --------------------------------------------------------------------
template <typename VT> struct Arith;
template <> struct Arith<double> {
    using acc_t = double;
    __device__ static __forceinline__ acc_t zero() { return 0.0; }
    __device__ static __forceinline__ acc_t mad(double v, double x, acc_t a) { return fma(v, x, a); }
    __device__ static __forceinline__ double out(acc_t a) { return a; }
    __device__ static __forceinline__ acc_t add(acc_t a, acc_t b) { return a + b; }
};
template <> struct Arith<float> {
    using acc_t = float;
    __device__ static __forceinline__ acc_t zero() { return 0.0f; }
    __device__ static __forceinline__ acc_t mad(float v, float x, acc_t a) { return fmaf(v, x, a); }
    __device__ static __forceinline__ float out(acc_t a) { return a; }
    __device__ static __forceinline__ acc_t add(acc_t a, acc_t b) { return a + b; }
};
template <> struct Arith<__half> {
    using acc_t = __half;
    __device__ static __forceinline__ acc_t zero() { return __float2half_rn(0.0f); }
    __device__ static __forceinline__ acc_t mad(__half v, __half x, acc_t a) {
        return __float2half_rn(fmaf(__half2float(v), __half2float(x), __half2float(a)));
    }
    __device__ static __forceinline__ __half out(acc_t a) { return a; }
    __device__ static __forceinline__ acc_t add(acc_t a, acc_t b) { return __float2half_rn(__half2float(a) + __half2float(b)); }
};

// streaming loads for the matrix (read once), read-only cached loads for x
template <typename T> __device__ __forceinline__ T ld_stream(const T *p) { return __ldcs(p); }
template <typename T> __device__ __forceinline__ T ld_x(const T *p) { return __ldg(p); }

// ---------------------------------------------------------------------------------------------
// SCS SpMV.  CT > 0: compile-time chunk height; CT == 0: runtime C (any C > 0, like spmv_gpu_scs).
// UNPERM: fused un-permuted write y[new_to_old[row]] (columns must be in original numbering).
// ---------------------------------------------------------------------------------------------
template <typename VT, int CT, int U, bool UNPERM>
__global__ void __launch_bounds__(TPB)
k_scs_spmv(long n_pad, int Crt, const int *__restrict__ chunk_list, int chunk_offset, const int *__restrict__ chunk_ptrs,
           const int *__restrict__ chunk_lengths, const int *__restrict__ col_idxs, const VT *__restrict__ values,
           const VT *__restrict__ x, VT *__restrict__ y, const int *__restrict__ new_to_old) {
    using A = Arith<VT>;
    long row = blockIdx.x * (long)TPB + threadIdx.x;  // n_pad = (#work items) * C
    if (row >= n_pad) return;
    const int C = CT > 0 ? CT : Crt;
    const long item = row / C;
    const int lane = (int)(row - item * C);
    const long c = chunk_list ? chunk_list[item] : item + chunk_offset;
    row = c * C + lane;
    const int len = chunk_lengths[c];
    long e = (long)chunk_ptrs[c] + lane;
    typename A::acc_t acc = A::zero();
    for (int j = 0; j < len; j += U) {
        int col[U];
        VT v[U], xv[U];
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (j + u < len) {
                col[u] = ld_stream(col_idxs + e + (long)u * C);
                v[u] = ld_stream(values + e + (long)u * C);
            }
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (j + u < len) xv[u] = ld_x(x + col[u]);
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (j + u < len) acc = A::mad(v[u], xv[u], acc);
        e += (long)U * C;
    }
    if (UNPERM) {
        const int o = new_to_old[row];
        if (o >= 0) y[o] = A::out(acc);
    } else
        y[row] = A::out(acc);
}

// ---------------------------------------------------------------------------------------------
// CRS SpMV (C = 1, sigma = 1): T adjacent lanes share one row, rows of a warp are contiguous in
// memory so the value/column streams stay coalesced; partial sums are combined by a shuffle tree.
// (Summation order differs from the sequential CPU loop => compared within tolerance, not bit-exact.)
// ---------------------------------------------------------------------------------------------
template <typename VT, int T, int R>
__global__ void __launch_bounds__(TPB)
k_csr_spmv(long n_rows, const int *__restrict__ row_ptrs, const int *__restrict__ col_idxs, const VT *__restrict__ values,
           const VT *__restrict__ x, VT *__restrict__ y) {
    // a group of T lanes handles R ADJACENT rows (their elements are contiguous); the loads of all R rows are issued
    // before any reduction so that R*ceil(len/T) independent load chains are in flight per lane
    using A = Arith<VT>;
    const long gt = blockIdx.x * (long)TPB + threadIdx.x;
    const long row0 = (gt / T) * R;
    const int t = (int)(gt % T);
    typename A::acc_t acc[R];
    int beg[R + 1];
#pragma unroll
    for (int r = 0; r <= R; ++r) beg[r] = row0 + r <= n_rows ? row_ptrs[row0 + r] : (row0 < n_rows ? row_ptrs[n_rows] : 0);
#pragma unroll
    for (int r = 0; r < R; ++r) {
        acc[r] = A::zero();
        if (row0 + r < n_rows)
            for (int j = beg[r] + t; j < beg[r + 1]; j += T) acc[r] = A::mad(ld_stream(values + j), ld_x(x + ld_stream(col_idxs + j)), acc[r]);
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
#pragma unroll
        for (int off = T / 2; off > 0; off >>= 1) {
            typename A::acc_t other = __shfl_down_sync(0xffffffffu, acc[r], off, T);
            acc[r] = A::add(acc[r], other);
        }
        if (row0 + r < n_rows && t == 0) y[row0 + r] = A::out(acc[r]);
    }
}

// ---------------------------------------------------------------------------------------------
// SpMMV: one thread per padded row, BVS accumulators in registers.
//   rowwise: X[col*bvs + v]  — the bvs values of a gathered row are contiguous => vector loads
//   colwise: X[col + v*ld]
// BT > 0: compile-time block width; BT == 0: runtime bvs <= 16.
// ---------------------------------------------------------------------------------------------
// vectorised access to one row of a row-major block vector (BT*sizeof(VT) bytes, naturally aligned)
template <typename VT, int BT, typename W>
__device__ __forceinline__ void load_row_as(const VT *__restrict__ src, VT *xv) {
    constexpr int PER = sizeof(W) / sizeof(VT);
    const W *p = reinterpret_cast<const W *>(src);
#pragma unroll
    for (int k = 0; k < BT / PER; ++k) {
        alignas(sizeof(W)) VT tmp[PER];
        *reinterpret_cast<W *>(tmp) = __ldg(p + k);
#pragma unroll
        for (int m = 0; m < PER; ++m) xv[k * PER + m] = tmp[m];
    }
}

template <typename VT, int BT, typename W>
__device__ __forceinline__ void store_row_as(VT *__restrict__ dst, const VT *yv) {
    constexpr int PER = sizeof(W) / sizeof(VT);
    W *p = reinterpret_cast<W *>(dst);
#pragma unroll
    for (int k = 0; k < BT / PER; ++k) {
        alignas(sizeof(W)) VT tmp[PER];
#pragma unroll
        for (int m = 0; m < PER; ++m) tmp[m] = yv[k * PER + m];
        p[k] = *reinterpret_cast<W *>(tmp);
    }
}

template <typename VT, int BT>
__device__ __forceinline__ void load_block_row(const VT *__restrict__ X, long col, int bvs, VT *xv) {
    constexpr int BYTES = BT * (int)sizeof(VT);
    if constexpr (BT > 0 && BYTES % 16 == 0) load_row_as<VT, BT, int4>(X + col * BT, xv);
    else if constexpr (BT > 0 && BYTES % 8 == 0) load_row_as<VT, BT, int2>(X + col * BT, xv);
    else if constexpr (BT > 0 && BYTES % 4 == 0) load_row_as<VT, BT, int>(X + col * BT, xv);
    else {
        const int n = BT > 0 ? BT : bvs;
#pragma unroll
        for (int v = 0; v < (BT > 0 ? BT : 16); ++v)
            if (v < n) xv[v] = ld_x(X + col * n + v);
    }
}

template <typename VT, int BT>
__device__ __forceinline__ void store_block_row(VT *__restrict__ Y, long row, int bvs, const VT *yv) {
    constexpr int BYTES = BT * (int)sizeof(VT);
    if constexpr (BT > 0 && BYTES % 16 == 0) store_row_as<VT, BT, int4>(Y + row * BT, yv);
    else if constexpr (BT > 0 && BYTES % 8 == 0) store_row_as<VT, BT, int2>(Y + row * BT, yv);
    else if constexpr (BT > 0 && BYTES % 4 == 0) store_row_as<VT, BT, int>(Y + row * BT, yv);
    else {
        const int n = BT > 0 ? BT : bvs;
#pragma unroll
        for (int v = 0; v < (BT > 0 ? BT : 16); ++v)
            if (v < n) Y[row * n + v] = yv[v];
    }
}

template <typename VT, int BT, int LAYOUT>
__global__ void __launch_bounds__(TPB)
k_scs_spmmv(long n_pad, int C, const int *__restrict__ chunk_ptrs, const int *__restrict__ chunk_lengths,
            const int *__restrict__ col_idxs, const VT *__restrict__ values, const VT *__restrict__ X, VT *__restrict__ Y, int bvs_rt,
            long ld) {
    using A = Arith<VT>;
    constexpr int NB = BT > 0 ? BT : 16;
    const int bvs = BT > 0 ? BT : bvs_rt;
    const long row = blockIdx.x * (long)TPB + threadIdx.x;
    if (row >= n_pad) return;
    const long c = row / C;
    const int lane = (int)(row - c * C);
    const int len = chunk_lengths[c];
    long e = (long)chunk_ptrs[c] + lane;
    typename A::acc_t acc[NB];
#pragma unroll
    for (int v = 0; v < NB; ++v) acc[v] = A::zero();
    for (int j = 0; j < len; ++j, e += C) {
        const int col = ld_stream(col_idxs + e);
        const VT val = ld_stream(values + e);
        VT xv[NB];
        if (LAYOUT == USPMV_ROWWISE) {
            load_block_row<VT, BT>(X, (long)col, bvs, xv);
        } else {
#pragma unroll
            for (int v = 0; v < NB; ++v)
                if (v < bvs) xv[v] = ld_x(X + col + v * ld);
        }
#pragma unroll
        for (int v = 0; v < NB; ++v)
            if (v < bvs) acc[v] = A::mad(val, xv[v], acc[v]);
    }
    if (LAYOUT == USPMV_ROWWISE) {
        VT yv[NB];
#pragma unroll
        for (int v = 0; v < NB; ++v) yv[v] = A::out(acc[v]);
        store_block_row<VT, BT>(Y, row, bvs, yv);
    } else {
#pragma unroll
        for (int v = 0; v < NB; ++v)
            if (v < bvs) Y[row + v * ld] = A::out(acc[v]);
    }
}

// a chunk is "boundary" if any of its slots references a halo column (>= n_local)
__global__ void k_flag_boundary_chunks(long n_chunks, int C, int n_local, const int *__restrict__ chunk_ptrs,
                                       const int *__restrict__ chunk_lengths, const int *__restrict__ col_idxs, int *__restrict__ flag) {
    const long w = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= n_chunks) return;
    const long beg = chunk_ptrs[w], end = beg + (long)chunk_lengths[w] * C;
    int hit = 0;
    for (long e = beg + lane; e < end; e += 32) hit |= col_idxs[e] >= n_local;
    hit = __any_sync(0xffffffffu, hit);
    if (lane == 0) flag[w] = hit;
}

// measured best SpMMV variant (scripts/tune_mmv2.py on B200, 256^3 7-pt, profiles/r01m_tune_spmmv_variants.json): 1 = lane per row /
// 8 warps, 2 = T lanes per row ("wide") / 8 warps, 6 = wide / 24 warps per CTA
inline int mmv_default_variant(size_t vsize, int bvs, bool rowwise) {
    // measured on the 256^3 7-point matrix, B200 (profiles/r02r_tune_mmv3.txt); 15-20 = 8-warp CTAs with a register budget
    if (!rowwise) return 1;
    const long row_bytes = (long)vsize * bvs;
    if (row_bytes < 32) return (vsize == 2 && bvs == 8) ? 20 : 1;  // one 128-bit load per row: nothing to widen (hp bvs 8: 254 vs 270 us)
    if (vsize == 8 && (bvs == 8 || bvs == 4)) return 16;           // dp bvs 8: 551 vs 574 us (variant 6), dp bvs 4: 375 vs 378
    if (vsize == 4 && bvs == 8) return 15;                         // sp bvs 8: 323 vs 333 us
    if (vsize == 2 && bvs == 16) return 19;                        // hp bvs 16: 418 vs 453 us
    return 2;
}

// ---- launch helpers ----------------------------------------------------------------------------
inline unsigned blocks_for(long n) { return (unsigned)((n + TPB - 1) / TPB); }

// ---- C = 32: bulk-copy (TMA) streamed kernel, see scs_stream.cuh -----------------------------------------
template <typename VT, bool UNPERM, int LMAX, int D, int WARPS>
void launch_stream_v(long n_chunks, const int *list, int off, const int *cp, const int *cl, const int *ci, const VT *v, const VT *x, VT *y,
                     const int *n2o, cudaStream_t st, int bps) {
    using R = stream::WarpRing<VT, LMAX, D>;
    auto kern = stream::k_scs32_stream<VT, Arith<VT>, LMAX, D, WARPS, UNPERM, false>;
    constexpr int smem = WARPS * R::BYTES_ALIGNED;
    static bool configured_on[uspmv::MAX_DEVICES] = {};
    bool &configured = configured_on[uspmv::current_device()];
    if (!configured) {
        USPMV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = true;
    }
    int dev = 0;
    USPMV_CUDA(cudaGetDevice(&dev));
    long grid = (long)sm_count(dev) * bps;
    const long need = (n_chunks + WARPS - 1) / WARPS;
    if (grid > need) grid = need;
    kern<<<(unsigned)grid, WARPS * 32, smem, st>>>(n_chunks, list, off, cp, cl, ci, v, x, y, n2o, stream::FusedArgs{});
}

// C = 64 / 128: the wide-chunk streamed kernel (scs_stream.cuh, k_scsw_stream)
template <typename VT, bool UNPERM, int H>
void launch_stream_wide(long n_chunks, const int *list, int off, const int *cp, const int *cl, const int *ci, const VT *v, const VT *x, VT *y,
                        const int *n2o, cudaStream_t st) {
    constexpr int D = 2, WARPS = 16;
    using R = stream::WarpRing<VT, 8, D>;
    auto kern = stream::k_scsw_stream<VT, Arith<VT>, H, D, WARPS, UNPERM>;
    constexpr int smem = WARPS * R::BYTES_ALIGNED;
    static bool configured_on[uspmv::MAX_DEVICES] = {};
    bool &configured = configured_on[uspmv::current_device()];
    if (!configured) {
        USPMV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = true;
    }
    int dev = 0;
    USPMV_CUDA(cudaGetDevice(&dev));
    long grid = (long)sm_count(dev) * options().stream_blocks_per_sm;
    const long need = (n_chunks + WARPS - 1) / WARPS;
    if (grid > need) grid = need;
    kern<<<(unsigned)grid, WARPS * 32, smem, st>>>((int)n_chunks, list, off, cp, cl, ci, v, x, y, n2o, stream::FusedArgs{});
}

// C = 32, fp16: two adjacent chunks per work item, pieces of up to 16 slots (scs_stream.cuh, k_scs32_stream_pair); contiguous ranges only
template <typename VT, bool UNPERM>
void launch_stream_pair(long n_chunks, int off, const int *cp, const int *ci, const VT *v, const VT *x, VT *y, const int *n2o, cudaStream_t st) {
    constexpr int LMAX = 16, D = 2, WARPS = 16;
    using R = stream::WarpRing<VT, LMAX, D>;
    auto kern = stream::k_scs32_stream_pair<VT, Arith<VT>, LMAX, D, WARPS, UNPERM>;
    constexpr int smem = WARPS * R::BYTES_ALIGNED;
    static bool configured_on[uspmv::MAX_DEVICES] = {};
    const int dev = uspmv::current_device();
    if (!configured_on[dev]) {
        USPMV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured_on[dev] = true;
    }
    long grid = (long)sm_count(dev) * options().stream_blocks_per_sm;
    const long need = ((n_chunks + 1) / 2 + WARPS - 1) / WARPS;
    if (grid > need) grid = need;
    kern<<<(unsigned)grid, WARPS * 32, smem, st>>>((int)n_chunks, off, cp, ci, v, x, y, n2o);
}

// C = 16 / 8: the narrow-chunk streamed kernel (scs_stream.cuh, k_scsn_stream); contiguous chunk ranges only
template <typename VT, bool UNPERM, int G, int WARPS = 16>
void launch_stream_narrow(long n_chunks, int off, const int *cp, const int *cl, const int *ci, const VT *v, const VT *x, VT *y, const int *n2o,
                          cudaStream_t st) {
    constexpr int D = 2;
    using R = stream::NarrowRing<VT, G, D>;
    auto kern = stream::k_scsn_stream<VT, Arith<VT>, G, D, WARPS, UNPERM>;
    constexpr int smem = WARPS * R::BYTES_ALIGNED;
    static bool configured_on[uspmv::MAX_DEVICES] = {};
    bool &configured = configured_on[uspmv::current_device()];
    if (!configured) {
        USPMV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = true;
    }
    int dev = 0;
    USPMV_CUDA(cudaGetDevice(&dev));
    long grid = (long)sm_count(dev) * options().stream_blocks_per_sm;
    const long n_items = (n_chunks + G - 1) / G;
    const long need = (n_items + WARPS - 1) / WARPS;
    if (grid > need) grid = need;
    kern<<<(unsigned)grid, WARPS * 32, smem, st>>>((int)n_chunks, off, cp, cl, ci, v, x, y, n2o);
}

template <typename VT, bool UNPERM, int LMAX, int D, int WARPS>
void launch_stream_pf(long n_chunks, const int *list, int off, const int *cp, const int *cl, const int *ci, const VT *v, const VT *x, VT *y,
                      const int *n2o, cudaStream_t st, int bps_req) {
    using R = stream::WarpRing<VT, LMAX, D>;
    auto kern = stream::k_scs32_stream_pf<VT, Arith<VT>, LMAX, D, WARPS, UNPERM>;
    constexpr int smem = WARPS * R::BYTES_ALIGNED;
    static bool configured_on[uspmv::MAX_DEVICES] = {};
    bool &configured = configured_on[uspmv::current_device()];
    static int bps_max_on[uspmv::MAX_DEVICES];
    int &bps_max = bps_max_on[uspmv::current_device()];
    if (!configured) {
        USPMV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        USPMV_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps_max, kern, WARPS * 32, smem));
        if (bps_max < 1) bps_max = 1;
        configured = true;
    }
    int dev = 0;
    USPMV_CUDA(cudaGetDevice(&dev));
    long grid = (long)sm_count(dev) * std::min(bps_req > 0 ? bps_req : bps_max, bps_max);
    const long need = (n_chunks + WARPS - 1) / WARPS;
    if (grid > need) grid = need;
    kern<<<(unsigned)grid, WARPS * 32, smem, st>>>((int)n_chunks, list, off, cp, cl, ci, v, x, y, n2o);
}

template <typename VT, bool UNPERM>
void launch_stream(long n_chunks, const int *list, int off, const int *cp, const int *cl, const int *ci, const VT *v, const VT *x, VT *y,
                   const int *n2o, cudaStream_t st) {
    const Options &c = options();
    switch (c.stream_variant) {
    case 1: launch_stream_v<VT, UNPERM, 8, 3, 8>(n_chunks, list, off, cp, cl, ci, v, x, y, n2o, st, c.stream_blocks_per_sm); break;
    case 2: launch_stream_v<VT, UNPERM, 8, 4, 8>(n_chunks, list, off, cp, cl, ci, v, x, y, n2o, st, c.stream_blocks_per_sm); break;
    case 3: launch_stream_v<VT, UNPERM, 4, 4, 16>(n_chunks, list, off, cp, cl, ci, v, x, y, n2o, st, c.stream_blocks_per_sm); break;
    case 4: launch_stream_v<VT, UNPERM, 4, 3, 16>(n_chunks, list, off, cp, cl, ci, v, x, y, n2o, st, c.stream_blocks_per_sm); break;
    case 5: launch_stream_v<VT, UNPERM, 4, 2, 16>(n_chunks, list, off, cp, cl, ci, v, x, y, n2o, st, c.stream_blocks_per_sm); break;
    case 6: launch_stream_v<VT, UNPERM, 8, 2, 8>(n_chunks, list, off, cp, cl, ci, v, x, y, n2o, st, c.stream_blocks_per_sm); break;
    case 7: launch_stream_v<VT, UNPERM, 16, 2, 8>(n_chunks, list, off, cp, cl, ci, v, x, y, n2o, st, c.stream_blocks_per_sm); break;
    case 8: launch_stream_v<VT, UNPERM, 8, 2, 32>(n_chunks, list, off, cp, cl, ci, v, x, y, n2o, st, c.stream_blocks_per_sm); break;
    case 9: launch_stream_v<VT, UNPERM, 2, 4, 16>(n_chunks, list, off, cp, cl, ci, v, x, y, n2o, st, c.stream_blocks_per_sm); break;
    case 13: launch_stream_v<VT, UNPERM, 8, 3, 16>(n_chunks, list, off, cp, cl, ci, v, x, y, n2o, st, c.stream_blocks_per_sm); break;
    case 14: launch_stream_v<VT, UNPERM, 8, 4, 16>(n_chunks, list, off, cp, cl, ci, v, x, y, n2o, st, c.stream_blocks_per_sm); break;
    case 10: launch_stream_pf<VT, UNPERM, 8, 3, 8>(n_chunks, list, off, cp, cl, ci, v, x, y, n2o, st, c.stream_blocks_per_sm); break;
    case 11: launch_stream_pf<VT, UNPERM, 8, 4, 8>(n_chunks, list, off, cp, cl, ci, v, x, y, n2o, st, c.stream_blocks_per_sm); break;
    case 12: launch_stream_pf<VT, UNPERM, 4, 4, 8>(n_chunks, list, off, cp, cl, ci, v, x, y, n2o, st, c.stream_blocks_per_sm); break;
    default: launch_stream_v<VT, UNPERM, 8, 2, 16>(n_chunks, list, off, cp, cl, ci, v, x, y, n2o, st, c.stream_blocks_per_sm); break;
    }
}

}  // namespace

namespace uspmv {
// One-launch distributed SpMV (C = 32): push + interior + wait + boundary + ack, see scs_stream.cuh.
template <typename VT>
static void launch_fused_t(const uspmv_scs *s, const void *x, void *y, const stream::FusedArgs &fa, cudaStream_t st) {
    constexpr int LMAX = 8, D = 2, WARPS = 16;
    using R = stream::WarpRing<VT, LMAX, D>;
    auto kern = stream::k_scs32_stream<VT, Arith<VT>, LMAX, D, WARPS, false, true>;
    constexpr int smem = WARPS * R::BYTES_ALIGNED;
    static bool configured_on[uspmv::MAX_DEVICES] = {};
    bool &configured = configured_on[uspmv::current_device()];
    if (!configured) {
        USPMV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = true;
    }
    int dev = 0;
    USPMV_CUDA(cudaGetDevice(&dev));
    // every CTA must be resident at once (warps spin on peer flags): never more than 2 CTAs per SM
    const long grid = (long)sm_count(dev) * 2;
    kern<<<(unsigned)grid, WARPS * 32, smem, st>>>(fa.n_int + fa.n_bnd, nullptr, 0, s->chunk_ptrs.p, s->chunk_lengths.p, s->col_idxs.p,
                                                 reinterpret_cast<const VT *>(s->values.p), static_cast<const VT *>(x), static_cast<VT *>(y),
                                                 nullptr, fa);
    USPMV_LAUNCH_CHECK();
}

// C = 64 / 128: the wide-chunk kernel with the fused exchange
template <typename VT, int H>
static void launch_fused_wide_t(const uspmv_scs *s, const void *x, void *y, const stream::FusedArgs &fa, cudaStream_t st) {
    constexpr int D = 2, WARPS = 16;
    using R = stream::WarpRing<VT, 8, D>;
    auto kern = stream::k_scsw_stream<VT, Arith<VT>, H, D, WARPS, false, true>;
    constexpr int smem = WARPS * R::BYTES_ALIGNED;
    static bool configured_on[uspmv::MAX_DEVICES] = {};
    const int dev = uspmv::current_device();
    if (!configured_on[dev]) {
        USPMV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured_on[dev] = true;
    }
    const long grid = (long)sm_count(dev) * 2;  // 2 CTAs of 16 warps per SM, all resident (warps spin on peer flags)
    kern<<<(unsigned)grid, WARPS * 32, smem, st>>>((int)(fa.n_int + fa.n_bnd), nullptr, 0, s->chunk_ptrs.p, s->chunk_lengths.p, s->col_idxs.p,
                                                 reinterpret_cast<const VT *>(s->values.p), static_cast<const VT *>(x), static_cast<VT *>(y), nullptr, fa);
    USPMV_LAUNCH_CHECK();
}

bool scs_fused_supported(const uspmv_scs *s) { return options().scs_stream && (s->C == 32 || ((s->C == 64 || s->C == 128) && options().scs_stream_wide)); }

void launch_scs32_fused(const uspmv_scs *s, const void *x, void *y, const stream::FusedArgs &fa, cudaStream_t st) {
    if (!scs_fused_supported(s)) fail("fused halo exchange kernel needs C = 32, 64 or 128 (got %ld)", s->C);
#define USPMV_FUSED_VT(VT)                                                    \
    do {                                                                      \
        if (s->C == 32) launch_fused_t<VT>(s, x, y, fa, st);                  \
        else if (s->C == 64) launch_fused_wide_t<VT, 2>(s, x, y, fa, st);     \
        else launch_fused_wide_t<VT, 4>(s, x, y, fa, st);                     \
    } while (0)
    switch (s->vt) {
    case USPMV_F64: USPMV_FUSED_VT(double); break;
    case USPMV_F32: USPMV_FUSED_VT(float); break;
    default: USPMV_FUSED_VT(__half);
    }
#undef USPMV_FUSED_VT
}
}  // namespace uspmv

namespace {

// uneven C = 32 matrices: segment work items + ordered partial reduction (scs_stream.cuh)
template <typename VT>
void launch_split(const uspmv_scs *s, const void *x, void *y, cudaStream_t st) {
    constexpr int LMAX = 8, D = 4, WARPS = 8;
    using R = stream::WarpRing<VT, LMAX, D>;
    auto kern = stream::k_scs32_stream_split<VT, Arith<VT>, LMAX, D, WARPS>;
    constexpr int smem = WARPS * R::BYTES_ALIGNED;
    static bool configured_on[uspmv::MAX_DEVICES] = {};
    bool &configured = configured_on[uspmv::current_device()];
    static int bps_on[uspmv::MAX_DEVICES];
    int &bps = bps_on[uspmv::current_device()];
    if (!configured) {
        USPMV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        USPMV_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, kern, WARPS * 32, smem));
        if (bps < 1) bps = 1;
        configured = true;
    }
    int dev = 0;
    USPMV_CUDA(cudaGetDevice(&dev));
    long grid = (long)sm_count(dev) * bps;
    const long need = (s->n_vitems + WARPS - 1) / WARPS;
    if (grid > need) grid = need;
    VT *partial = reinterpret_cast<VT *>(s->partials.p);
    kern<<<(unsigned)grid, WARPS * 32, smem, st>>>(s->n_vitems, s->vitems.p, s->chunk_ptrs.p, s->col_idxs.p, reinterpret_cast<const VT *>(s->values.p),
                                                  static_cast<const VT *>(x), static_cast<VT *>(y), partial);
    USPMV_LAUNCH_CHECK();
    if (s->n_split) {
        stream::k_reduce_partials<VT, Arith<VT>><<<(unsigned)((s->n_split * 32 + 255) / 256), 256, 0, st>>>(s->n_split, s->split_chunk.p, s->split_ptr.p,
                                                                                                          partial, static_cast<VT *>(y));
        USPMV_LAUNCH_CHECK();
    }
}

template <typename VT, bool UNPERM>
void launch_scs(long C, long n_chunks, const int *cp, const int *cl, const int *ci, const void *vals, const void *x, void *y,
                const int *n2o, cudaStream_t st, const int *list = nullptr, int off = 0) {
    // n_chunks = number of work items: all chunks, or the length of `list`
    const long n_pad = n_chunks * C;
    if (n_pad == 0) return;
    const VT *v = static_cast<const VT *>(vals);
    const VT *xx = static_cast<const VT *>(x);
    VT *yy = static_cast<VT *>(y);
    if constexpr (sizeof(VT) == 2) {
        // fp16: pieces of two adjacent chunks (contiguous chunk ranges; chunk lists — longest-first order, interior / boundary
        // subsets — keep the one-chunk kernel)
        if (C == 32 && options().scs_stream && options().pair_hp && !list && options().stream_variant == 0) {
            launch_stream_pair<VT, UNPERM>(n_chunks, off, cp, ci, v, xx, yy, n2o, st);
            USPMV_LAUNCH_CHECK();
            return;
        }
    }
    if (C == 32 && options().scs_stream) {
        launch_stream<VT, UNPERM>(n_chunks, list, off, cp, cl, ci, v, xx, yy, n2o, st);
        USPMV_LAUNCH_CHECK();
        return;
    }
    // narrow chunks: G adjacent chunks per warp need a contiguous chunk range that starts at a multiple of G
    // Measured at 256^3 (scripts/narrow_vs_direct.py): C = 16 streamed / direct: dp 340 / 342 us, sp 203 / 215, hp 204 / 230; C = 8 (G = 4):
    // 410 / 349, 255 / 228, 281 / 234 — the producer of the narrow kernel issues 2 G bulk copies per piece and its instruction count,
    // not bandwidth, bounds it.  So: C = 16 in sp / hp only; everything else narrow stays on the direct kernel.
    if (C == 16 && (sizeof(VT) < 8 || options().narrow_dp) && options().scs_stream && options().scs_stream_wide && !list && off % 2 == 0) {
        // fp64: 12 warps per CTA (<= 85 registers, nothing spilled); 16 warps / 64 registers spill the pending metadata loads
        if (sizeof(VT) == 8) launch_stream_narrow<VT, UNPERM, 2, 12>(n_chunks, off, cp, cl, ci, v, xx, yy, n2o, st);
        else launch_stream_narrow<VT, UNPERM, 2>(n_chunks, off, cp, cl, ci, v, xx, yy, n2o, st);
        USPMV_LAUNCH_CHECK();
        return;
    }
    if ((C == 64 || C == 128) && options().scs_stream && options().scs_stream_wide) {
        if (C == 64) launch_stream_wide<VT, UNPERM, 2>(n_chunks, list, off, cp, cl, ci, v, xx, yy, n2o, st);
        else launch_stream_wide<VT, UNPERM, 4>(n_chunks, list, off, cp, cl, ci, v, xx, yy, n2o, st);
        USPMV_LAUNCH_CHECK();
        return;
    }
    const unsigned g = blocks_for(n_pad);
#define USPMV_SCS_CASE(CC)                                                                                              \
    case CC: k_scs_spmv<VT, CC, 8, UNPERM><<<g, TPB, 0, st>>>(n_pad, (int)C, list, off, cp, cl, ci, v, xx, yy, n2o); break;
    switch (C) {
        USPMV_SCS_CASE(1)
        USPMV_SCS_CASE(2)
        USPMV_SCS_CASE(4)
        USPMV_SCS_CASE(8)
        USPMV_SCS_CASE(16)
        USPMV_SCS_CASE(32)
        USPMV_SCS_CASE(64)
        USPMV_SCS_CASE(128)
        USPMV_SCS_CASE(256)
    default: k_scs_spmv<VT, 0, 8, UNPERM><<<g, TPB, 0, st>>>(n_pad, (int)C, list, off, cp, cl, ci, v, xx, yy, n2o);
    }
#undef USPMV_SCS_CASE
    USPMV_LAUNCH_CHECK();
}

template <typename VT>
void launch_csr(long n_rows, long nnz_hint, const int *rp, const int *ci, const void *vals, const void *x, void *y, cudaStream_t st) {
    if (n_rows == 0) return;
    const VT *v = static_cast<const VT *>(vals);
    const VT *xx = static_cast<const VT *>(x);
    VT *yy = static_cast<VT *>(y);
    const double avg = nnz_hint >= 0 ? (double)nnz_hint / (double)n_rows : 8.0;
    int T = 2;
    while (T < 32 && T < avg) T *= 2;
    constexpr int R = 4;
    const unsigned g = blocks_for((n_rows + R - 1) / R * T);
    switch (T) {
    case 2: k_csr_spmv<VT, 2, R><<<g, TPB, 0, st>>>(n_rows, rp, ci, v, xx, yy); break;
    case 4: k_csr_spmv<VT, 4, R><<<g, TPB, 0, st>>>(n_rows, rp, ci, v, xx, yy); break;
    case 8: k_csr_spmv<VT, 8, R><<<g, TPB, 0, st>>>(n_rows, rp, ci, v, xx, yy); break;
    case 16: k_csr_spmv<VT, 16, R><<<g, TPB, 0, st>>>(n_rows, rp, ci, v, xx, yy); break;
    default: k_csr_spmv<VT, 32, R><<<g, TPB, 0, st>>>(n_rows, rp, ci, v, xx, yy); break;
    }
    USPMV_LAUNCH_CHECK();
}

// CRS of a library-built matrix (arrays carry the 8-element slack): streamed, sequential per row => bit-identical
template <typename VT>
void launch_csr_stream(long n_rows, const int *rp, const int *ci, const void *vals, const void *x, void *y, cudaStream_t st) {
    if (n_rows == 0) return;
    // tiles of 256 (fp64) / 384 (fp32) / 512 (fp16) elements = 3 KB per stage for every type (2 CTAs x 16 warps stay resident):
    // the bulk-copy engine serves roughly one copy per 46 cycles per SM, so narrow types want more elements per copy
    constexpr int LMAX = sizeof(VT) == 8 ? 8 : (sizeof(VT) == 4 ? 12 : 16), D = 2, WARPS = 16;
    using R = stream::WarpRing<VT, LMAX, D>;
    auto kern = stream::k_csr_stream<VT, Arith<VT>, LMAX, D, WARPS>;
    constexpr int smem = WARPS * R::BYTES_ALIGNED;
    static bool configured_on[uspmv::MAX_DEVICES] = {};
    bool &configured = configured_on[uspmv::current_device()];
    if (!configured) {
        USPMV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = true;
    }
    int dev = 0;
    USPMV_CUDA(cudaGetDevice(&dev));
    long grid = (long)sm_count(dev) * 2;
    const long need = ((n_rows + 31) / 32 + WARPS - 1) / WARPS;
    if (grid > need) grid = need;
    kern<<<(unsigned)grid, WARPS * 32, smem, st>>>(n_rows, rp, ci, static_cast<const VT *>(vals), static_cast<const VT *>(x), static_cast<VT *>(y));
    USPMV_LAUNCH_CHECK();
}

// C = 32 and bvs in {2,4,8,16}: streamed kernel
// chunk subset of a launch: n items, item k -> chunk list[k] (list != NULL) or k + off
struct ChunkSel {
    long n;
    const int *list;
    int off;
};

// the arrays of a SELL-C-sigma matrix, library-built (view_of) or caller-owned (the raw-array entry points)
struct ScsView {
    long C, n_chunks, n_rows_padded;
    const int *cp, *cl, *ci;
    const void *vals;
};

template <typename VT, int BVS, bool ROWWISE, int LMAX, int WARPS, bool WIDE, int D = 2, int MINB = 0>
void launch_spmmv_stream_v(const ScsView &s, const VT *X, VT *Y, long ld, cudaStream_t st, const ChunkSel &sel) {
    using R = stream::WarpRing<VT, LMAX, D>;
    auto kern = stream::k_scs32_stream_mmv<VT, Arith<VT>, LMAX, D, WARPS, BVS, ROWWISE, WIDE, false, MINB>;
    constexpr int smem = WARPS * R::BYTES_ALIGNED;
    static bool configured_on[uspmv::MAX_DEVICES] = {};
    bool &configured = configured_on[uspmv::current_device()];
    static int blocks_per_sm_on[uspmv::MAX_DEVICES];
    int &blocks_per_sm = blocks_per_sm_on[uspmv::current_device()];
    if (!configured) {
        USPMV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        USPMV_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, kern, WARPS * 32, smem));
        if (blocks_per_sm < 1) blocks_per_sm = 1;
        configured = true;
    }
    int dev = 0;
    USPMV_CUDA(cudaGetDevice(&dev));
    const int bps = options().mmv_blocks_per_sm > 0 ? std::min(options().mmv_blocks_per_sm, blocks_per_sm) : blocks_per_sm;
    long grid = (long)sm_count(dev) * bps;
    const long need = (sel.n + WARPS - 1) / WARPS;
    if (grid > need) grid = need;
    if (grid < 1) return;
    kern<<<(unsigned)grid, WARPS * 32, smem, st>>>(sel.n, sel.list, sel.off, s.cp, s.cl, s.ci,
                                                  static_cast<const VT *>(s.vals), X, Y, ld, stream::FusedArgs{});
}

// ONE kernel per distributed SpMMV (push + interior + wait + boundary + ack): every CTA must be resident, so the grid is exactly
// SMs x resident CTAs per SM
template <typename VT, int BVS, bool ROWWISE, int LMAX, int WARPS, bool WIDE, int MINB = 0>
void launch_spmmv_fused_v(const ScsView &s, const VT *X, VT *Y, long ld, cudaStream_t st, const stream::FusedArgs &fa) {
    constexpr int D = 2;
    using R = stream::WarpRing<VT, LMAX, D>;
    auto kern = stream::k_scs32_stream_mmv<VT, Arith<VT>, LMAX, D, WARPS, BVS, ROWWISE, WIDE, true, MINB>;
    constexpr int smem = WARPS * R::BYTES_ALIGNED;
    static bool configured_on[uspmv::MAX_DEVICES] = {};
    static int bps_on[uspmv::MAX_DEVICES];
    const int dev = uspmv::current_device();
    if (!configured_on[dev]) {
        USPMV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        USPMV_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps_on[dev], kern, WARPS * 32, smem));
        if (bps_on[dev] < 1) bps_on[dev] = 1;
        configured_on[dev] = true;
    }
    const int bps = options().mmv_blocks_per_sm > 0 ? std::min(options().mmv_blocks_per_sm, bps_on[dev]) : bps_on[dev];
    const long grid = (long)sm_count(dev) * bps;
    kern<<<(unsigned)grid, WARPS * 32, smem, st>>>(fa.n_int + fa.n_bnd, nullptr, 0, s.cp, s.cl, s.ci, static_cast<const VT *>(s.vals), X, Y, ld, fa);
}

template <typename VT, int BVS, bool ROWWISE>
void launch_spmmv_fused(const ScsView &s, const VT *X, VT *Y, long ld, cudaStream_t st, const stream::FusedArgs &fa) {
    int v = options().mmv_variant;  // 0: the tuned instantiation per (precision, block_vec_size, layout); else forced (A/B runs)
    if (v == 0) {
        v = mmv_default_variant(sizeof(VT), BVS, ROWWISE);
        // dp, block_vec_size 8: at the 80-register bound of variants 16 / 6 the FUSED instance loses 12-75 % to its single-GPU twin
        // (ptxas), with 128 registers it does not: 610 vs 603 us (profiles/r02w_probe.txt)
        if (ROWWISE && sizeof(VT) == 8 && BVS == 8) v = 17;
        // (16-byte rows, sp bvs 4: the fused instance of variant 1 is 13 % slower than its twin and stating its 48 registers — variant
        // 20 — does not help: profiles/r02D_probe.txt; the step still equals the multi-kernel one, 282 vs 280 us)
    }
    switch (v) {
    case 2: launch_spmmv_fused_v<VT, BVS, ROWWISE, 8, 8, true>(s, X, Y, ld, st, fa); break;
    case 6: launch_spmmv_fused_v<VT, BVS, ROWWISE, 8, 24, true>(s, X, Y, ld, st, fa); break;
    case 17: launch_spmmv_fused_v<VT, BVS, ROWWISE, 8, 8, true, 2>(s, X, Y, ld, st, fa); break;
    case 15: launch_spmmv_fused_v<VT, BVS, ROWWISE, 8, 8, true, 4>(s, X, Y, ld, st, fa); break;
    case 16: launch_spmmv_fused_v<VT, BVS, ROWWISE, 8, 8, true, 3>(s, X, Y, ld, st, fa); break;
    case 19: launch_spmmv_fused_v<VT, BVS, ROWWISE, 8, 8, false, 3>(s, X, Y, ld, st, fa); break;
    case 20: launch_spmmv_fused_v<VT, BVS, ROWWISE, 8, 8, true, 5>(s, X, Y, ld, st, fa); break;
    default: launch_spmmv_fused_v<VT, BVS, ROWWISE, 8, 8, false>(s, X, Y, ld, st, fa); break;
    }
}

// variant: 0 = tuned default per (precision, bvs, layout); 1..4 force (wide body?, slots per stage)
template <typename VT, int BVS, bool ROWWISE>
void launch_spmmv_stream(const ScsView &s, const VT *X, VT *Y, long ld, cudaStream_t st, const ChunkSel &sel) {
    int v = options().mmv_variant;
    if (v == 0) v = mmv_default_variant(sizeof(VT), BVS, ROWWISE);
    switch (v) {
    case 2: launch_spmmv_stream_v<VT, BVS, ROWWISE, 8, 8, true>(s, X, Y, ld, st, sel); break;
    case 3: launch_spmmv_stream_v<VT, BVS, ROWWISE, 4, 8, false>(s, X, Y, ld, st, sel); break;
    case 4: launch_spmmv_stream_v<VT, BVS, ROWWISE, 4, 8, true>(s, X, Y, ld, st, sel); break;
    case 5: launch_spmmv_stream_v<VT, BVS, ROWWISE, 8, 16, true>(s, X, Y, ld, st, sel); break;
    case 6: launch_spmmv_stream_v<VT, BVS, ROWWISE, 8, 24, true>(s, X, Y, ld, st, sel); break;
    case 7: launch_spmmv_stream_v<VT, BVS, ROWWISE, 8, 16, false>(s, X, Y, ld, st, sel); break;
    case 8: launch_spmmv_stream_v<VT, BVS, ROWWISE, 8, 24, false>(s, X, Y, ld, st, sel); break;
    case 9: launch_spmmv_stream_v<VT, BVS, ROWWISE, 4, 24, true>(s, X, Y, ld, st, sel); break;    // half-size stages: twice the L1 left for X
    case 10: launch_spmmv_stream_v<VT, BVS, ROWWISE, 4, 24, false>(s, X, Y, ld, st, sel); break;
    case 11: launch_spmmv_stream_v<VT, BVS, ROWWISE, 4, 24, true, 3>(s, X, Y, ld, st, sel); break;
    case 12: launch_spmmv_stream_v<VT, BVS, ROWWISE, 4, 16, true, 3>(s, X, Y, ld, st, sel); break;
    case 13: launch_spmmv_stream_v<VT, BVS, ROWWISE, 4, 32, true>(s, X, Y, ld, st, sel); break;
    case 14: launch_spmmv_stream_v<VT, BVS, ROWWISE, 4, 32, false>(s, X, Y, ld, st, sel); break;
    // 8-warp CTAs with a minimum of resident CTAs per SM: ptxas then keeps the gathers of a piece in flight (see k_scs32_stream_mmv)
    case 15: launch_spmmv_stream_v<VT, BVS, ROWWISE, 8, 8, true, 2, 4>(s, X, Y, ld, st, sel); break;   // <= 64 registers
    case 16: launch_spmmv_stream_v<VT, BVS, ROWWISE, 8, 8, true, 2, 3>(s, X, Y, ld, st, sel); break;   // <= 80
    case 17: launch_spmmv_stream_v<VT, BVS, ROWWISE, 8, 8, true, 2, 2>(s, X, Y, ld, st, sel); break;   // <= 128
    case 18: launch_spmmv_stream_v<VT, BVS, ROWWISE, 8, 8, false, 2, 4>(s, X, Y, ld, st, sel); break;
    case 19: launch_spmmv_stream_v<VT, BVS, ROWWISE, 8, 8, false, 2, 3>(s, X, Y, ld, st, sel); break;
    case 20: launch_spmmv_stream_v<VT, BVS, ROWWISE, 8, 8, true, 2, 5>(s, X, Y, ld, st, sel); break;   // <= 48, but told so
    default: launch_spmmv_stream_v<VT, BVS, ROWWISE, 8, 8, false>(s, X, Y, ld, st, sel); break;
    }
}

inline bool spmmv_streamed(long C, int bvs) {
    return C == 32 && options().scs_stream && (bvs == 2 || bvs == 4 || bvs == 8 || bvs == 16);
}

template <typename VT, int LAYOUT>
void launch_spmmv_l(const ScsView &s, const void *X, void *Y, int bvs, long ld, cudaStream_t st, const ChunkSel *subset = nullptr) {
    const long n_pad = s.n_rows_padded;
    if (n_pad == 0) return;
    if (subset && !spmmv_streamed(s.C, bvs)) fail("uspmv_spmmv_part: chunk subsets need the streamed kernel (C = 32, block_vec_size 2/4/8/16)");
    const ChunkSel sel = subset ? *subset : ChunkSel{s.n_chunks, nullptr, 0};
    const VT *v = static_cast<const VT *>(s.vals);
    const VT *xx = static_cast<const VT *>(X);
    VT *yy = static_cast<VT *>(Y);
    // the bulk copies of the streamed kernel need 16-byte aligned array bases (always true for library-built matrices)
    const bool aligned = reinterpret_cast<uintptr_t>(s.ci) % 16 == 0 && reinterpret_cast<uintptr_t>(s.vals) % 16 == 0;
    if (subset && !aligned) fail("uspmv_spmmv_part: misaligned matrix arrays");
    if (spmmv_streamed(s.C, bvs) && aligned) {
        constexpr bool RW = LAYOUT == USPMV_ROWWISE;
        switch (bvs) {
        case 2: launch_spmmv_stream<VT, 2, RW>(s, xx, yy, ld, st, sel); break;
        case 4: launch_spmmv_stream<VT, 4, RW>(s, xx, yy, ld, st, sel); break;
        case 8: launch_spmmv_stream<VT, 8, RW>(s, xx, yy, ld, st, sel); break;
        default: launch_spmmv_stream<VT, 16, RW>(s, xx, yy, ld, st, sel);
        }
        USPMV_LAUNCH_CHECK();
        return;
    }
    const unsigned g = blocks_for(n_pad);
    const int C = (int)s.C;
#define USPMV_MMV_CASE(BB)                                                                                                      \
    case BB: k_scs_spmmv<VT, BB, LAYOUT><<<g, TPB, 0, st>>>(n_pad, C, s.cp, s.cl, s.ci, v, xx, yy, \
                                                           bvs, ld); break;
    switch (bvs) {
        USPMV_MMV_CASE(2)
        USPMV_MMV_CASE(4)
        USPMV_MMV_CASE(8)
        USPMV_MMV_CASE(16)
    default: k_scs_spmmv<VT, 0, LAYOUT><<<g, TPB, 0, st>>>(n_pad, C, s.cp, s.cl, s.ci, v, xx, yy, bvs, ld);
    }
#undef USPMV_MMV_CASE
    USPMV_LAUNCH_CHECK();
}

template <typename VT>
void launch_spmmv(const ScsView &s, const void *X, void *Y, int bvs, long ld, int layout, cudaStream_t st, const ChunkSel *subset = nullptr) {
    if (layout == USPMV_ROWWISE) launch_spmmv_l<VT, USPMV_ROWWISE>(s, X, Y, bvs, ld, st, subset);
    else launch_spmmv_l<VT, USPMV_COLWISE>(s, X, Y, bvs, ld, st, subset);
}

inline ScsView view_of(const uspmv_scs *s) {
    return ScsView{s->C, s->n_chunks, s->n_rows_padded, s->chunk_ptrs.p, s->chunk_lengths.p, s->col_idxs.p, s->values.p};
}
inline bool spmmv_streamed(const uspmv_scs *s, int bvs) { return spmmv_streamed(s->C, bvs); }

}  // namespace

namespace uspmv {
bool spmmv_fused_supported(const uspmv_scs *s, int bvs) { return spmmv_streamed(s->C, bvs); }

// One-launch distributed SpMMV over the arena's block vector (C = 32, block_vec_size 2 / 4 / 8 / 16)
void launch_spmmv_fused_any(const uspmv_scs *s, const void *X, void *Y, int bvs, long ld, int layout, const stream::FusedArgs &fa, cudaStream_t st) {
    if (!spmmv_fused_supported(s, bvs)) fail("fused SpMMV needs C = 32 and block_vec_size 2 / 4 / 8 / 16");
    const ScsView v = view_of(s);
#define USPMV_FUSED_MMV(VT, RW)                                                                                             \
    switch (bvs) {                                                                                                          \
    case 2: launch_spmmv_fused<VT, 2, RW>(v, static_cast<const VT *>(X), static_cast<VT *>(Y), ld, st, fa); break;          \
    case 4: launch_spmmv_fused<VT, 4, RW>(v, static_cast<const VT *>(X), static_cast<VT *>(Y), ld, st, fa); break;          \
    case 8: launch_spmmv_fused<VT, 8, RW>(v, static_cast<const VT *>(X), static_cast<VT *>(Y), ld, st, fa); break;          \
    default: launch_spmmv_fused<VT, 16, RW>(v, static_cast<const VT *>(X), static_cast<VT *>(Y), ld, st, fa);               \
    }
#define USPMV_FUSED_MMV_L(VT)                                        \
    do {                                                             \
        if (layout == USPMV_ROWWISE) { USPMV_FUSED_MMV(VT, true) }   \
        else { USPMV_FUSED_MMV(VT, false) }                          \
    } while (0)
    switch (s->vt) {
    case USPMV_F64: USPMV_FUSED_MMV_L(double); break;
    case USPMV_F32: USPMV_FUSED_MMV_L(float); break;
    default: USPMV_FUSED_MMV_L(__half);
    }
#undef USPMV_FUSED_MMV_L
#undef USPMV_FUSED_MMV
    USPMV_LAUNCH_CHECK();
}
}  // namespace uspmv

namespace {

// ---- permutation kernels -----------------------------------------------------------------------
template <typename VT>
__global__ void k_apply_perm(VT *__restrict__ out, const VT *__restrict__ in, const int *__restrict__ perm, long n) {
    long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    int p = perm[i];
    out[i] = p >= 0 ? in[p] : VT(0.0);
}

__global__ void k_row_lengths(const int *__restrict__ row_ptrs, long n, int *__restrict__ len) {
    long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i < n) len[i] = row_ptrs[i + 1] - row_ptrs[i];
}

template <typename VT>
__global__ void k_apply_perm_strided(VT *__restrict__ out, const VT *__restrict__ in, const int *__restrict__ perm, long n, long stride) {
    long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i < n) out[i * stride] = in[(long)perm[i] * stride];
}

__global__ void k_inv_perm(const int *__restrict__ perm, int *__restrict__ inv, long n) {
    long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i < n) inv[perm[i]] = (int)i;
}

__global__ void k_check_perm_range(const int *__restrict__ perm, long n, long limit, int *flag) {
    long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i < n && (perm[i] < 0 || perm[i] >= limit)) atomicOr(flag, 1);
}

template <typename VT>
__global__ void k_apply_perm_block(VT *__restrict__ out, const VT *__restrict__ in, const int *__restrict__ perm, long n, int bvs,
                                   long ld, int layout) {
    long t = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (t >= n * bvs) return;
    long i, v;
    if (layout == USPMV_ROWWISE) { i = t / bvs; v = t % bvs; }
    else { v = t / n; i = t % n; }
    int p = perm[i];
    VT val = VT(0.0);
    if (p >= 0) val = layout == USPMV_ROWWISE ? in[(long)p * bvs + v] : in[(long)p + v * ld];
    if (layout == USPMV_ROWWISE) out[i * bvs + v] = val;
    else out[i + v * ld] = val;
}

}  // namespace

extern "C" {

int uspmv_scs_gpu(uspmv_ctx *ctx, int vt, long C, long n_chunks, const int *cp, const int *cl, const int *ci, const void *vals,
                  const void *x, void *y, void *stream) {
    return guarded([&] {
        OptScope opt_scope(ctx_of(ctx));  // this context's options govern everything below
        if (!ctx) fail("uspmv_scs_gpu: ctx is NULL");
        if (C < 1) fail("uspmv_scs_gpu: C must be >= 1");
        use_device(ctx);
        cudaStream_t st = as_stream(stream);
        switch (vt) {
        case USPMV_F64: launch_scs<double, false>(C, n_chunks, cp, cl, ci, vals, x, y, nullptr, st); break;
        case USPMV_F32: launch_scs<float, false>(C, n_chunks, cp, cl, ci, vals, x, y, nullptr, st); break;
        case USPMV_F16: launch_scs<__half, false>(C, n_chunks, cp, cl, ci, vals, x, y, nullptr, st); break;
        default: fail("uspmv_scs_gpu: invalid value type %d", vt);
        }
    });
}

int uspmv_csr_gpu(uspmv_ctx *ctx, int vt, long n_rows, const int *rp, const int *ci, const void *vals, const void *x, void *y,
                  void *stream) {
    return guarded([&] {
        OptScope opt_scope(ctx_of(ctx));  // this context's options govern everything below
        if (!ctx) fail("uspmv_csr_gpu: ctx is NULL");
        use_device(ctx);
        cudaStream_t st = as_stream(stream);
        // caller-owned arrays: the streamed kernel (sequential per row, bit-identical to kernels.hpp:46-57) needs 16-byte aligned
        // bases for its bulk copies; otherwise the split-row vector kernel
        const bool streamed = options().scs_stream && reinterpret_cast<uintptr_t>(ci) % 16 == 0 && reinterpret_cast<uintptr_t>(vals) % 16 == 0;
        switch (vt) {
        case USPMV_F64: streamed ? launch_csr_stream<double>(n_rows, rp, ci, vals, x, y, st) : launch_csr<double>(n_rows, -1, rp, ci, vals, x, y, st); break;
        case USPMV_F32: streamed ? launch_csr_stream<float>(n_rows, rp, ci, vals, x, y, st) : launch_csr<float>(n_rows, -1, rp, ci, vals, x, y, st); break;
        case USPMV_F16: streamed ? launch_csr_stream<__half>(n_rows, rp, ci, vals, x, y, st) : launch_csr<__half>(n_rows, -1, rp, ci, vals, x, y, st); break;
        default: fail("uspmv_csr_gpu: invalid value type %d", vt);
        }
    });
}

int uspmv_spmv(const uspmv_scs *s, const void *x, void *y, void *stream) {
    return guarded([&] {
        OptScope opt_scope(ctx_of(s));  // this context's options govern everything below
        if (!s) fail("uspmv_spmv: scs is NULL");
        if (s->n_rows_padded && (!x || !y)) fail("uspmv_spmv: NULL vector");
        use_device(s->ctx);
        cudaStream_t st = as_stream(stream);
        const bool crs = (s->C == 1 && s->sigma == 1);  // execute_uspmv's rule, interface.hpp:1911
        if (s->n_vitems > 0 && options().scs_stream && options().split_long_chunks > 0) {
            switch (s->vt) {
            case USPMV_F64: launch_split<double>(s, x, y, st); break;
            case USPMV_F32: launch_split<float>(s, x, y, st); break;
            default: launch_split<__half>(s, x, y, st);
            }
            return;
        }
        switch (s->vt) {
        case USPMV_F64:
            if (crs && options().scs_stream) launch_csr_stream<double>(s->n_rows, s->chunk_ptrs.p, s->col_idxs.p, s->values.p, x, y, st);
            else if (crs) launch_csr<double>(s->n_rows, s->nnz, s->chunk_ptrs.p, s->col_idxs.p, s->values.p, x, y, st);
            else launch_scs<double, false>(s->C, s->n_chunks, s->chunk_ptrs.p, s->chunk_lengths.p, s->col_idxs.p, s->values.p, x, y, nullptr, st, s->balanced_order.p);
            break;
        case USPMV_F32:
            if (crs && options().scs_stream) launch_csr_stream<float>(s->n_rows, s->chunk_ptrs.p, s->col_idxs.p, s->values.p, x, y, st);
            else if (crs) launch_csr<float>(s->n_rows, s->nnz, s->chunk_ptrs.p, s->col_idxs.p, s->values.p, x, y, st);
            else launch_scs<float, false>(s->C, s->n_chunks, s->chunk_ptrs.p, s->chunk_lengths.p, s->col_idxs.p, s->values.p, x, y, nullptr, st, s->balanced_order.p);
            break;
        default:
            if (crs && options().scs_stream) launch_csr_stream<__half>(s->n_rows, s->chunk_ptrs.p, s->col_idxs.p, s->values.p, x, y, st);
            else if (crs) launch_csr<__half>(s->n_rows, s->nnz, s->chunk_ptrs.p, s->col_idxs.p, s->values.p, x, y, st);
            else launch_scs<__half, false>(s->C, s->n_chunks, s->chunk_ptrs.p, s->chunk_lengths.p, s->col_idxs.p, s->values.p, x, y, nullptr, st, s->balanced_order.p);
        }
    });
}

int uspmv_scs_split_chunks(uspmv_scs *s, long *n_interior, long *n_boundary) {
    return guarded([&] {
        OptScope opt_scope(ctx_of(s));  // this context's options govern everything below
        if (!s) fail("uspmv_scs_split_chunks: scs is NULL");
        USPMV_CUDA(cudaSetDevice(s->ctx->device));
        const long nc = s->n_chunks;
        DevBuf<int> flag(nc + 1), pos(nc + 1);
        USPMV_CUDA(cudaMemset(flag.p, 0, (nc + 1) * sizeof(int)));
        if (nc) {
            k_flag_boundary_chunks<<<blocks_for(nc * 32), TPB>>>(nc, (int)s->C, (int)s->n_rows, s->chunk_ptrs.p, s->chunk_lengths.p, s->col_idxs.p, flag.p);
            USPMV_LAUNCH_CHECK();
        }
        std::vector<int> fh(nc + 1);
        USPMV_CUDA(cudaMemcpy(fh.data(), flag.p, (nc + 1) * sizeof(int), cudaMemcpyDeviceToHost));
        std::vector<int> in, bd;
        for (long c = 0; c < nc; ++c) (fh[c] ? bd : in).push_back((int)c);  // order preserved: classify, do not reorder
        s->interior_chunks.alloc(in.size());
        s->boundary_chunks.alloc(bd.size());
        if (!in.empty()) USPMV_CUDA(cudaMemcpy(s->interior_chunks.p, in.data(), in.size() * sizeof(int), cudaMemcpyHostToDevice));
        if (!bd.empty()) USPMV_CUDA(cudaMemcpy(s->boundary_chunks.p, bd.data(), bd.size() * sizeof(int), cudaMemcpyHostToDevice));
        s->chunks_split = true;
        // a contiguous class (typical for slab partitions) needs no index list: chunk = k + offset
        auto contiguous = [](const std::vector<int> &v) { return !v.empty() && (long)v.back() - v.front() + 1 == (long)v.size(); };
        s->interior_contig = contiguous(in); s->interior_off = in.empty() ? 0 : in.front();
        s->boundary_contig = contiguous(bd); s->boundary_off = bd.empty() ? 0 : bd.front();
        if (n_interior) *n_interior = (long)in.size();
        if (n_boundary) *n_boundary = (long)bd.size();
    });
}

int uspmv_spmv_part(const uspmv_scs *s, int which, const void *x, void *y, void *stream) {
    return guarded([&] {
        OptScope opt_scope(ctx_of(s));  // this context's options govern everything below
        if (!s) fail("uspmv_spmv_part: scs is NULL");
        if (which == 0) {
            if (uspmv_spmv(s, x, y, stream)) throw Error(uspmv_last_error());
            return;
        }
        if (which != 1 && which != 2) fail("uspmv_spmv_part: which must be 0 (all), 1 (interior) or 2 (boundary)");
        if (!s->chunks_split) fail("uspmv_spmv_part: call uspmv_scs_split_chunks first");
        use_device(s->ctx);
        const DevBuf<int> &l = which == 1 ? s->interior_chunks : s->boundary_chunks;
        if (l.n == 0) return;
        const bool contig = which == 1 ? s->interior_contig : s->boundary_contig;
        const int off = contig ? (which == 1 ? s->interior_off : s->boundary_off) : 0;
        const int *lp = contig ? nullptr : l.p;
        cudaStream_t st = as_stream(stream);
        switch (s->vt) {
        case USPMV_F64: launch_scs<double, false>(s->C, (long)l.n, s->chunk_ptrs.p, s->chunk_lengths.p, s->col_idxs.p, s->values.p, x, y, nullptr, st, lp, off); break;
        case USPMV_F32: launch_scs<float, false>(s->C, (long)l.n, s->chunk_ptrs.p, s->chunk_lengths.p, s->col_idxs.p, s->values.p, x, y, nullptr, st, lp, off); break;
        default: launch_scs<__half, false>(s->C, (long)l.n, s->chunk_ptrs.p, s->chunk_lengths.p, s->col_idxs.p, s->values.p, x, y, nullptr, st, lp, off);
        }
    });
}

int uspmv_spmv_unpermuted(const uspmv_scs *s, const void *x, void *y, void *stream) {
    return guarded([&] {
        OptScope opt_scope(ctx_of(s));  // this context's options govern everything below
        if (!s) fail("uspmv_spmv_unpermuted: scs is NULL");
        if (s->cols_permuted) fail("uspmv_spmv_unpermuted: columns were already permuted (permute_scs_cols); use uspmv_spmv");
        use_device(s->ctx);
        cudaStream_t st = as_stream(stream);
        switch (s->vt) {
        case USPMV_F64: launch_scs<double, true>(s->C, s->n_chunks, s->chunk_ptrs.p, s->chunk_lengths.p, s->col_idxs.p, s->values.p, x, y, s->new_to_old.p, st, s->balanced_order.p); break;
        case USPMV_F32: launch_scs<float, true>(s->C, s->n_chunks, s->chunk_ptrs.p, s->chunk_lengths.p, s->col_idxs.p, s->values.p, x, y, s->new_to_old.p, st, s->balanced_order.p); break;
        default: launch_scs<__half, true>(s->C, s->n_chunks, s->chunk_ptrs.p, s->chunk_lengths.p, s->col_idxs.p, s->values.p, x, y, s->new_to_old.p, st, s->balanced_order.p);
        }
    });
}

int uspmv_spmmv(const uspmv_scs *s, const void *X, void *Y, int bvs, long vec_length, int layout, void *stream) {
    return guarded([&] {
        OptScope opt_scope(ctx_of(s));  // this context's options govern everything below
        if (!s) fail("uspmv_spmmv: scs is NULL");
        if (bvs < 1 || bvs > 16) fail("uspmv_spmmv: block_vec_size must be in [1,16] (got %d)", bvs);
        if (layout != USPMV_COLWISE && layout != USPMV_ROWWISE) fail("uspmv_spmmv: invalid layout %d", layout);
        if (layout == USPMV_COLWISE && vec_length < s->n_rows_padded) fail("uspmv_spmmv: vec_length %ld < n_rows_padded %ld", vec_length, s->n_rows_padded);
        use_device(s->ctx);
        cudaStream_t st = as_stream(stream);
        switch (s->vt) {
        case USPMV_F64: launch_spmmv<double>(view_of(s), X, Y, bvs, vec_length, layout, st); break;
        case USPMV_F32: launch_spmmv<float>(view_of(s), X, Y, bvs, vec_length, layout, st); break;
        default: launch_spmmv<__half>(view_of(s), X, Y, bvs, vec_length, layout, st);
        }
    });
}

/* SpMMV over the interior (1) or boundary (2) chunks only (uspmv_scs_split_chunks); 0 = all.  Subsets need the streamed
 * kernel; uspmv_spmmv_part_supported tells the caller whether to overlap or to run one full SpMMV after the exchange. */
int uspmv_spmmv_part_supported(const uspmv_scs *s, int bvs) { return s && s->chunks_split && spmmv_streamed(s, bvs) ? 1 : 0; }

int uspmv_spmmv_part(const uspmv_scs *s, int which, const void *X, void *Y, int bvs, long vec_length, int layout, void *stream) {
    return guarded([&] {
        OptScope opt_scope(ctx_of(s));  // this context's options govern everything below
        if (!s) fail("uspmv_spmmv_part: scs is NULL");
        if (which == 0) {
            if (uspmv_spmmv(s, X, Y, bvs, vec_length, layout, stream)) throw Error(uspmv_last_error());
            return;
        }
        if (which != 1 && which != 2) fail("uspmv_spmmv_part: which must be 0 (all), 1 (interior) or 2 (boundary)");
        if (!s->chunks_split) fail("uspmv_spmmv_part: call uspmv_scs_split_chunks first");
        use_device(s->ctx);
        if (bvs < 1 || bvs > 16) fail("uspmv_spmmv_part: block_vec_size must be in [1,16] (got %d)", bvs);
        if (layout != USPMV_COLWISE && layout != USPMV_ROWWISE) fail("uspmv_spmmv_part: invalid layout %d", layout);
        if (layout == USPMV_COLWISE && vec_length < s->n_rows_padded) fail("uspmv_spmmv_part: vec_length %ld < n_rows_padded %ld", vec_length, s->n_rows_padded);
        const DevBuf<int> &l = which == 1 ? s->interior_chunks : s->boundary_chunks;
        if (l.n == 0) return;
        const bool contig = which == 1 ? s->interior_contig : s->boundary_contig;
        const ChunkSel sel{(long)l.n, contig ? nullptr : l.p, contig ? (which == 1 ? s->interior_off : s->boundary_off) : 0};
        cudaStream_t st = as_stream(stream);
        switch (s->vt) {
        case USPMV_F64: launch_spmmv<double>(view_of(s), X, Y, bvs, vec_length, layout, st, &sel); break;
        case USPMV_F32: launch_spmmv<float>(view_of(s), X, Y, bvs, vec_length, layout, st, &sel); break;
        default: launch_spmmv<__half>(view_of(s), X, Y, bvs, vec_length, layout, st, &sel);
        }
    });
}

/* block_spmv_{scs,csr} on caller-owned DEVICE arrays (the reference's block kernels take raw arrays per call, kernels.hpp:68-154,
 * 306-398; its GPU launchers are stubs, :777-844).  C == 1 is CRS (chunk_ptrs = row_ptrs, chunk_lengths may be NULL and is then not
 * read).  Y has n_chunks * C block rows.  The streamed kernel needs 16-byte aligned col_idxs / values (cudaMalloc gives 256). */
int uspmv_block_spmv_gpu(uspmv_ctx *ctx, int vt, long C, long n_chunks, const int *cp, const int *cl, const int *ci, const void *vals,
                         const void *X, void *Y, int bvs, long vec_length, int layout, void *stream) {
    return guarded([&] {
        OptScope opt_scope(ctx_of(ctx));  // this context's options govern everything below
        if (!ctx) fail("uspmv_block_spmv_gpu: ctx is NULL");
        if (C < 1) fail("uspmv_block_spmv_gpu: C must be >= 1");
        if (bvs < 1 || bvs > 16) fail("uspmv_block_spmv_gpu: block_vec_size must be in [1,16] (got %d)", bvs);
        if (layout != USPMV_COLWISE && layout != USPMV_ROWWISE) fail("uspmv_block_spmv_gpu: invalid layout %d", layout);
        if (n_chunks > 0 && (!cp || !ci || !vals || !X || !Y)) fail("uspmv_block_spmv_gpu: NULL array");
        if (C > 1 && n_chunks > 0 && !cl) fail("uspmv_block_spmv_gpu: chunk_lengths is NULL");
        if (layout == USPMV_COLWISE && vec_length < n_chunks * C) fail("uspmv_block_spmv_gpu: vec_length %ld < n_chunks * C = %ld", vec_length, n_chunks * C);
        USPMV_CUDA(cudaSetDevice(ctx->device));
        cudaStream_t st = as_stream(stream);
        DevBuf<int> cl_tmp;
        if (!cl && n_chunks > 0) {  // CRS without a length array: lengths = row_ptrs[r + 1] - row_ptrs[r]
            cl_tmp.alloc(n_chunks);
            k_row_lengths<<<blocks_for(n_chunks), TPB, 0, st>>>(cp, n_chunks, cl_tmp.p);
            USPMV_LAUNCH_CHECK();
            cl = cl_tmp.p;
        }
        const ScsView v{C, n_chunks, n_chunks * C, cp, cl, ci, vals};
        switch (vt) {
        case USPMV_F64: launch_spmmv<double>(v, X, Y, bvs, vec_length, layout, st); break;
        case USPMV_F32: launch_spmmv<float>(v, X, Y, bvs, vec_length, layout, st); break;
        case USPMV_F16: launch_spmmv<__half>(v, X, Y, bvs, vec_length, layout, st); break;
        default: fail("uspmv_block_spmv_gpu: invalid value type %d", vt);
        }
        if (cl_tmp.p) USPMV_CUDA(cudaStreamSynchronize(st));  // the temporary is freed on return
    });
}

int uspmv_spmv_host(const uspmv_scs *s_, const void *x_h, long x_len, void *y_h, long y_len) {
    return guarded([&] {
        OptScope opt_scope(ctx_of(s_));  // this context's options govern everything below
        uspmv_scs *s = const_cast<uspmv_scs *>(s_);
        if (!s || !x_h || !y_h) fail("uspmv_spmv_host: NULL argument");
        if (y_len < s->n_rows_padded) fail("uspmv_spmv_host: y_len %ld < n_rows_padded %ld", y_len, s->n_rows_padded);
        if (x_len < s->x_min_len) fail("uspmv_spmv_host: x_len %ld < %ld (the matrix references columns up to there)", x_len, s->x_min_len);
        USPMV_CUDA(cudaSetDevice(s->ctx->device));
        const size_t es = vt_size(s->vt);
        if (s->h2d_stage_x.n < (size_t)x_len * es) s->h2d_stage_x.alloc((size_t)x_len * es);
        if (s->d2h_stage_y.n < (size_t)s->n_rows_padded * es) s->d2h_stage_y.alloc((size_t)s->n_rows_padded * es);
        USPMV_CUDA(cudaMemcpyAsync(s->h2d_stage_x.p, x_h, (size_t)x_len * es, cudaMemcpyHostToDevice, 0));
        if (uspmv_spmv(s, s->h2d_stage_x.p, s->d2h_stage_y.p, nullptr)) throw Error(uspmv_last_error());
        USPMV_CUDA(cudaMemcpyAsync(y_h, s->d2h_stage_y.p, (size_t)s->n_rows_padded * es, cudaMemcpyDeviceToHost, 0));
        USPMV_CUDA(cudaStreamSynchronize(0));
    });
}

// Pipelined host-buffer SpMV: submit() enqueues H2D(x) on a copy stream, the kernel on a compute stream and D2H(y) on a
// second copy stream; up to HOST_SLOTS calls are in flight, so the PCIe transfers of one SpMV overlap the kernel and the
// opposite-direction transfer of its neighbours.  x_h / y_h should be pinned (uspmv_host_alloc) for the copies to be async.
int uspmv_spmv_host_submit(const uspmv_scs *s_, const void *x_h, long x_len, void *y_h, long y_len, int slot) {
    return guarded([&] {
        OptScope opt_scope(ctx_of(s_));  // this context's options govern everything below
        uspmv_scs *s = const_cast<uspmv_scs *>(s_);
        if (!s || !x_h || !y_h) fail("uspmv_spmv_host_submit: NULL argument");
        if (slot < 0 || slot >= uspmv_scs::HOST_SLOTS) fail("uspmv_spmv_host_submit: slot must be in [0,%d)", uspmv_scs::HOST_SLOTS);
        if (y_len < s->n_rows_padded) fail("uspmv_spmv_host_submit: y_len %ld < n_rows_padded %ld", y_len, s->n_rows_padded);
        if (x_len < s->x_min_len) fail("uspmv_spmv_host_submit: x_len %ld < %ld (the matrix references columns up to there)", x_len, s->x_min_len);
        if (s->slot_busy[slot]) fail("uspmv_spmv_host_submit: slot %d is still in flight (call uspmv_spmv_host_wait)", slot);
        USPMV_CUDA(cudaSetDevice(s->ctx->device));
        const size_t es = vt_size(s->vt);
        if (!s->s_run) {
            USPMV_CUDA(cudaStreamCreateWithFlags(&s->s_h2d, cudaStreamNonBlocking));
            USPMV_CUDA(cudaStreamCreateWithFlags(&s->s_run, cudaStreamNonBlocking));
            USPMV_CUDA(cudaStreamCreateWithFlags(&s->s_d2h, cudaStreamNonBlocking));
            for (int k = 0; k < uspmv_scs::HOST_SLOTS; ++k) {
                USPMV_CUDA(cudaEventCreateWithFlags(&s->ev_x[k], cudaEventDisableTiming));
                USPMV_CUDA(cudaEventCreateWithFlags(&s->ev_y[k], cudaEventDisableTiming));
                USPMV_CUDA(cudaEventCreateWithFlags(&s->ev_done[k], cudaEventDisableTiming));
            }
        }
        if (s->slot_x[slot].n < (size_t)x_len * es) s->slot_x[slot].alloc((size_t)x_len * es);
        if (s->slot_y[slot].n < (size_t)s->n_rows_padded * es) s->slot_y[slot].alloc((size_t)s->n_rows_padded * es);
        USPMV_CUDA(cudaMemcpyAsync(s->slot_x[slot].p, x_h, (size_t)x_len * es, cudaMemcpyHostToDevice, s->s_h2d));
        USPMV_CUDA(cudaEventRecord(s->ev_x[slot], s->s_h2d));
        USPMV_CUDA(cudaStreamWaitEvent(s->s_run, s->ev_x[slot], 0));
        if (uspmv_spmv(s, s->slot_x[slot].p, s->slot_y[slot].p, s->s_run)) throw Error(uspmv_last_error());
        USPMV_CUDA(cudaEventRecord(s->ev_y[slot], s->s_run));
        USPMV_CUDA(cudaStreamWaitEvent(s->s_d2h, s->ev_y[slot], 0));
        USPMV_CUDA(cudaMemcpyAsync(y_h, s->slot_y[slot].p, (size_t)s->n_rows_padded * es, cudaMemcpyDeviceToHost, s->s_d2h));
        USPMV_CUDA(cudaEventRecord(s->ev_done[slot], s->s_d2h));
        s->slot_busy[slot] = true;
    });
}

int uspmv_spmv_host_wait(const uspmv_scs *s_, int slot) {
    return guarded([&] {
        OptScope opt_scope(ctx_of(s_));  // this context's options govern everything below
        uspmv_scs *s = const_cast<uspmv_scs *>(s_);
        if (!s) fail("uspmv_spmv_host_wait: NULL argument");
        if (slot < 0 || slot >= uspmv_scs::HOST_SLOTS) fail("uspmv_spmv_host_wait: slot must be in [0,%d)", uspmv_scs::HOST_SLOTS);
        if (!s->slot_busy[slot]) return;
        USPMV_CUDA(cudaEventSynchronize(s->ev_done[slot]));
        s->slot_busy[slot] = false;
    });
}

int uspmv_apply_permutation(uspmv_ctx *ctx, void *out, const void *in, const int *perm, long n, int vt, void *stream) {
    return guarded([&] {
        OptScope opt_scope(ctx_of(ctx));  // this context's options govern everything below
        if (!ctx) fail("uspmv_apply_permutation: ctx is NULL");
        if (n == 0) return;
        cudaStream_t st = as_stream(stream);
        const unsigned g = blocks_for(n);
        switch (vt) {
        case USPMV_F64: k_apply_perm<double><<<g, TPB, 0, st>>>((double *)out, (const double *)in, perm, n); break;
        case USPMV_F32: k_apply_perm<float><<<g, TPB, 0, st>>>((float *)out, (const float *)in, perm, n); break;
        case USPMV_F16: k_apply_perm<__half><<<g, TPB, 0, st>>>((__half *)out, (const __half *)in, perm, n); break;
        default: fail("uspmv_apply_permutation: invalid value type %d", vt);
        }
        USPMV_LAUNCH_CHECK();
    });
}

int uspmv_apply_permutation_block(uspmv_ctx *ctx, void *out, const void *in, const int *perm, long n, int vt, int bvs, long ld,
                                  int layout, void *stream) {
    return guarded([&] {
        OptScope opt_scope(ctx_of(ctx));  // this context's options govern everything below
        if (!ctx) fail("uspmv_apply_permutation_block: ctx is NULL");
        if (n == 0 || bvs == 0) return;
        cudaStream_t st = as_stream(stream);
        const unsigned g = blocks_for(n * bvs);
        switch (vt) {
        case USPMV_F64: k_apply_perm_block<double><<<g, TPB, 0, st>>>((double *)out, (const double *)in, perm, n, bvs, ld, layout); break;
        case USPMV_F32: k_apply_perm_block<float><<<g, TPB, 0, st>>>((float *)out, (const float *)in, perm, n, bvs, ld, layout); break;
        case USPMV_F16: k_apply_perm_block<__half><<<g, TPB, 0, st>>>((__half *)out, (const __half *)in, perm, n, bvs, ld, layout); break;
        default: fail("uspmv_apply_permutation_block: invalid value type %d", vt);
        }
        USPMV_LAUNCH_CHECK();
    });
}

/* apply_strided_permutation (utilities.hpp:1784-1799), literally: out[i * stride] = in[perm[i] * stride], i < n — ONE element per
 * row; the harness calls it once per vector of a row-major block vector with the base pointers offset by the vector index. */
int uspmv_apply_strided_permutation(uspmv_ctx *ctx, void *out, const void *in, const int *perm, long n, long stride, int vt, void *stream) {
    return guarded([&] {
        OptScope opt_scope(ctx_of(ctx));  // this context's options govern everything below
        if (!ctx) fail("uspmv_apply_strided_permutation: ctx is NULL");
        if (stride < 1) fail("uspmv_apply_strided_permutation: stride must be >= 1");
        if (n == 0) return;
        cudaStream_t st = as_stream(stream);
        const unsigned g = blocks_for(n);
        switch (vt) {
        case USPMV_F64: k_apply_perm_strided<double><<<g, TPB, 0, st>>>((double *)out, (const double *)in, perm, n, stride); break;
        case USPMV_F32: k_apply_perm_strided<float><<<g, TPB, 0, st>>>((float *)out, (const float *)in, perm, n, stride); break;
        case USPMV_F16: k_apply_perm_strided<__half><<<g, TPB, 0, st>>>((__half *)out, (const __half *)in, perm, n, stride); break;
        default: fail("uspmv_apply_strided_permutation: invalid value type %d", vt);
        }
        USPMV_LAUNCH_CHECK();
    });
}

/* generate_inv_perm (utilities.hpp:1755-1766): inv_perm[perm[i]] = i, i < perm_len (device arrays).  Entries of perm outside
 * [0, inv_len) are an error here (the reference writes out of bounds); positions of inv_perm that no perm[i] names are left alone. */
int uspmv_generate_inv_perm(uspmv_ctx *ctx, const int *perm_d, int *inv_perm_d, long perm_len, long inv_len, void *stream) {
    return guarded([&] {
        OptScope opt_scope(ctx_of(ctx));  // this context's options govern everything below
        if (!ctx) fail("uspmv_generate_inv_perm: ctx is NULL");
        if (perm_len == 0) return;
        if (!perm_d || !inv_perm_d) fail("uspmv_generate_inv_perm: NULL array");
        USPMV_CUDA(cudaSetDevice(ctx->device));
        cudaStream_t st = as_stream(stream);
        DevBuf<int> flag(1);
        USPMV_CUDA(cudaMemsetAsync(flag.p, 0, sizeof(int), st));
        k_check_perm_range<<<blocks_for(perm_len), TPB, 0, st>>>(perm_d, perm_len, inv_len, flag.p);
        USPMV_LAUNCH_CHECK();
        int bad = 0;
        USPMV_CUDA(cudaMemcpyAsync(&bad, flag.p, sizeof(int), cudaMemcpyDeviceToHost, st));
        USPMV_CUDA(cudaStreamSynchronize(st));
        if (bad) fail("uspmv_generate_inv_perm: perm has an entry outside [0, %ld)", inv_len);
        k_inv_perm<<<blocks_for(perm_len), TPB, 0, st>>>(perm_d, inv_perm_d, perm_len);
        USPMV_LAUNCH_CHECK();
    });
}

}  // extern "C"
