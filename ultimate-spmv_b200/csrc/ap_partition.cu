// Adaptive precision: device-side partition_precisions and the fused dp/sp/hp SpMV kernels.
//
// Replaces
//   partition_precisions           code/interface.hpp:690-978, code/utilities.hpp:2810-3123
//   uspmv_scs_ap{dpsp,dphp,sphp,dpsphp}_cpu / uspmv_csr_ap*   code/interface.hpp:1129-1733
//   spmv_gpu_ap_scs / spmv_gpu_scs_ap_adv / spmv_gpu_ap_csr   code/ap_kernels.hpp:637-953 (dp+sp only)
//
// Arithmetic (canonical = the library kernels, SURVEY.md §8a' item 9): every part accumulates in its own
// fp64 register with one FMA per slot — the narrower value is widened to fp64 BEFORE the multiply — and
// y = dp + sp + hp in that association.  ap[sp_hp] multiplies in fp32 against the fp32 x (float*float,
// half*float), accumulates those products in fp64 and rounds y to fp32 (interface.hpp:1620-1645).
// One launch reads all parts: the parts share C, n_chunks and the row permutation of the first part.
#include "common.cuh"
#include "scs_stream.cuh"

#include <cub/device/device_scan.cuh>

using namespace uspmv;

namespace {

constexpr int TPB = 256;
inline unsigned blocks_for(long n) { return (unsigned)((n + TPB - 1) / TPB); }

template <typename T> __device__ __forceinline__ double to_f64(T v);
template <> __device__ __forceinline__ double to_f64<double>(double v) { return v; }
template <> __device__ __forceinline__ double to_f64<float>(float v) { return (double)v; }
template <> __device__ __forceinline__ double to_f64<__half>(__half v) { return (double)__half2float(v); }

// ---- partition ---------------------------------------------------------------------------------------
// part id per element: 0 dp, 1 sp, 2 hp.  Thresholds compared in fp64 on abs(value) (interface.hpp:938-964).
template <typename MT>
__global__ void k_classify(const MT *__restrict__ vals, const int *__restrict__ I, const int *__restrict__ J, long nnz, int mode, double t1,
                           double t2, const double *__restrict__ rowmax, const double *__restrict__ colmax, int *__restrict__ f0,
                           int *__restrict__ f1, int *__restrict__ f2) {
    long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i >= nnz) return;
    const double a = fabs(to_f64(vals[i]));
    double th1 = t1, th2 = t2;
    if (rowmax) {
        const double s = colmax[J[i]] * rowmax[I[i]];
        th1 = t1 / s;
        th2 = t2 / s;
    }
    int p;
    if (mode == USPMV_AP_DP_SP_HP) p = (a >= th1) ? 0 : ((a <= th1 && a >= th2) ? 1 : 2);
    else {
        const int hi = (mode == USPMV_AP_SP_HP) ? 1 : 0;
        const int lo = (mode == USPMV_AP_DP_SP) ? 1 : 2;
        p = (a >= th1) ? hi : lo;
    }
    f0[i] = p == 0;
    f1[i] = p == 1;
    f2[i] = p == 2;
}

template <typename MT, typename OT> __device__ __forceinline__ OT narrow(MT v);
template <> __device__ __forceinline__ double narrow<double, double>(double v) { return v; }
template <> __device__ __forceinline__ float narrow<double, float>(double v) { return (float)v; }
template <> __device__ __forceinline__ __half narrow<double, __half>(double v) { return __double2half(v); }
template <> __device__ __forceinline__ double narrow<float, double>(float v) { return (double)v; }
template <> __device__ __forceinline__ float narrow<float, float>(float v) { return v; }
template <> __device__ __forceinline__ __half narrow<float, __half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ double narrow<__half, double>(__half v) { return (double)__half2float(v); }
template <> __device__ __forceinline__ float narrow<__half, float>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ __half narrow<__half, __half>(__half v) { return v; }

// stable compaction: element i goes to position pos[i] of its part (pos = exclusive scan of the part's flags)
template <typename MT, typename OT>
__global__ void k_scatter_part(const MT *__restrict__ vals, const int *__restrict__ I, const int *__restrict__ J, long nnz,
                               const int *__restrict__ flag, const int *__restrict__ pos, int *__restrict__ oI, int *__restrict__ oJ,
                               OT *__restrict__ oV) {
    long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i >= nnz || !flag[i]) return;
    const int p = pos[i];
    oI[p] = I[i];
    oJ[p] = J[i];
    oV[p] = narrow<MT, OT>(vals[i]);
}

// ---- fused AP SpMV, one thread per padded row ------------------------------------------------------------
struct PartArgs {
    const int *cp, *cl, *ci;
    const void *v;
};

template <typename VT, int U>
__device__ __forceinline__ double part_sum_f64(const PartArgs &p, long c, int lane, int C, const double *__restrict__ x) {
    double acc = 0.0;
    const int len = p.cl[c];
    long e = (long)p.cp[c] + lane;
    const VT *vals = static_cast<const VT *>(p.v);
    for (int j = 0; j < len; j += U) {
        int col[U];
        VT v[U];
        double xv[U];
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (j + u < len) {
                col[u] = __ldcs(p.ci + e + (long)u * C);
                v[u] = __ldcs(vals + e + (long)u * C);
            }
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (j + u < len) xv[u] = __ldg(x + col[u]);
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (j + u < len) acc = fma(to_f64(v[u]), xv[u], acc);
        e += (long)U * C;
    }
    return acc;
}

// sp_hp: fp32 products (not contracted with the fp64 accumulation), fp64 accumulators
template <typename VT, int U>
__device__ __forceinline__ double part_sum_f32prod(const PartArgs &p, long c, int lane, int C, const float *__restrict__ x) {
    double acc = 0.0;
    const int len = p.cl[c];
    long e = (long)p.cp[c] + lane;
    const VT *vals = static_cast<const VT *>(p.v);
    for (int j = 0; j < len; j += U) {
        int col[U];
        VT v[U];
        float xv[U];
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (j + u < len) {
                col[u] = __ldcs(p.ci + e + (long)u * C);
                v[u] = __ldcs(vals + e + (long)u * C);
            }
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (j + u < len) xv[u] = __ldg(x + col[u]);
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (j + u < len) {
                float vf;
                if constexpr (sizeof(VT) == 4) vf = v[u];
                else vf = __half2float(v[u]);
                acc += (double)__fmul_rn(vf, xv[u]);
            }
        e += (long)U * C;
    }
    return acc;
}

template <int MODE>
__global__ void __launch_bounds__(TPB)
k_ap_spmv(long n_pad, int C, PartArgs dp, PartArgs sp, PartArgs hp, const void *__restrict__ x, void *__restrict__ y) {
    const long row = blockIdx.x * (long)TPB + threadIdx.x;
    if (row >= n_pad) return;
    const long c = row / C;
    const int lane = (int)(row - c * C);
    if constexpr (MODE == USPMV_AP_SP_HP) {
        const float *xf = static_cast<const float *>(x);
        const double s = part_sum_f32prod<float, 4>(sp, c, lane, C, xf);
        const double h = part_sum_f32prod<__half, 4>(hp, c, lane, C, xf);
        static_cast<float *>(y)[row] = (float)(s + h);
    } else {
        const double *xd = static_cast<const double *>(x);
        double r = part_sum_f64<double, 4>(dp, c, lane, C, xd);
        if constexpr (MODE == USPMV_AP_DP_SP || MODE == USPMV_AP_DP_SP_HP) r = r + part_sum_f64<float, 4>(sp, c, lane, C, xd);
        if constexpr (MODE == USPMV_AP_DP_HP || MODE == USPMV_AP_DP_SP_HP) r = r + part_sum_f64<__half, 4>(hp, c, lane, C, xd);
        static_cast<double *>(y)[row] = r;
    }
}

template <int MODE>
void launch_ap_stream(long n_chunks, const int *order, const stream::ApPart &a, const stream::ApPart &b, const stream::ApPart &c, const void *x,
                      void *y, cudaStream_t st) {
    constexpr int LMAX = 8, D = 2, WARPS = 8;
    using R = stream::WarpRing<double, LMAX, D>;
    auto kern = stream::k_scs32_stream_ap<MODE, LMAX, D, WARPS>;
    constexpr int smem = WARPS * R::BYTES_ALIGNED;
    static bool configured = false;
    static int bps = 1;
    if (!configured) {
        USPMV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        USPMV_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, kern, WARPS * 32, smem));
        if (bps < 1) bps = 1;
        configured = true;
    }
    int dev = 0;
    USPMV_CUDA(cudaGetDevice(&dev));
    long grid = (long)sm_count(dev) * bps;
    const long need = (n_chunks + WARPS - 1) / WARPS;
    if (grid > need) grid = need;
    kern<<<(unsigned)grid, WARPS * 32, smem, st>>>(n_chunks, order, a, b, c, x, y);
}

void exclusive_scan_i32(const int *in, int *out, long n) {
    size_t bytes = 0;
    USPMV_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, (int)n));
    DevBuf<unsigned char> tmp(bytes);
    USPMV_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, bytes, in, out, (int)n));
    g_launches.fetch_add(2);
}

template <typename MT, typename OT>
uspmv_coo *make_part(const uspmv_coo *coo, const int *flag, int vt_out) {
    const long nnz = coo->nnz;
    DevBuf<int> pos(nnz + 1);
    long cnt = 0;
    if (nnz) {
        exclusive_scan_i32(flag, pos.p, nnz);
        int last_pos = 0, last_flag = 0;
        USPMV_CUDA(cudaMemcpy(&last_pos, pos.p + nnz - 1, sizeof(int), cudaMemcpyDeviceToHost));
        USPMV_CUDA(cudaMemcpy(&last_flag, flag + nnz - 1, sizeof(int), cudaMemcpyDeviceToHost));
        cnt = (long)last_pos + last_flag;
    }
    auto out = new uspmv_coo();
    try {
        out->ctx = coo->ctx;
        out->n_rows = coo->n_rows;
        out->n_cols = coo->n_cols;
        out->nnz = cnt;
        out->mt = vt_out;
        out->I.alloc(cnt);
        out->J.alloc(cnt);
        out->values.alloc(cnt * sizeof(OT));
        if (nnz && cnt) {
            k_scatter_part<MT, OT><<<blocks_for(nnz), TPB>>>(reinterpret_cast<const MT *>(coo->values.p), coo->I.p, coo->J.p, nnz, flag, pos.p,
                                                             out->I.p, out->J.p, reinterpret_cast<OT *>(out->values.p));
            USPMV_LAUNCH_CHECK();
        }
        USPMV_CUDA(cudaDeviceSynchronize());
    } catch (...) { delete out; throw; }
    return out;
}

template <typename MT>
void partition_typed(const uspmv_coo *coo, int mode, double t1, double t2, const double *rowmax_d, const double *colmax_d, uspmv_coo **dp,
                     uspmv_coo **sp, uspmv_coo **hp) {
    const long nnz = coo->nnz;
    DevBuf<int> f0(nnz + 1), f1(nnz + 1), f2(nnz + 1);
    if (nnz) {
        k_classify<MT><<<blocks_for(nnz), TPB>>>(reinterpret_cast<const MT *>(coo->values.p), coo->I.p, coo->J.p, nnz, mode, t1, t2, rowmax_d,
                                                 colmax_d, f0.p, f1.p, f2.p);
        USPMV_LAUNCH_CHECK();
    }
    const bool use_dp = mode != USPMV_AP_SP_HP, use_sp = mode != USPMV_AP_DP_HP, use_hp = mode != USPMV_AP_DP_SP;
    uspmv_coo *a = nullptr, *b = nullptr, *c = nullptr;
    try {
        if (use_dp) a = make_part<MT, double>(coo, f0.p, USPMV_F64);
        if (use_sp) b = make_part<MT, float>(coo, f1.p, USPMV_F32);
        if (use_hp) c = make_part<MT, __half>(coo, f2.p, USPMV_F16);
    } catch (...) { delete a; delete b; delete c; throw; }
    long tot = (a ? a->nnz : 0) + (b ? b->nnz : 0) + (c ? c->nnz : 0);
    if (tot != nnz) {  // the reference's "Elements have been lost" check (interface.hpp:971-975)
        delete a; delete b; delete c;
        fail("partition_precisions: %ld elements have been lost when separating the matrix", nnz - tot);
    }
    if (dp) *dp = a; else delete a;
    if (sp) *sp = b; else delete b;
    if (hp) *hp = c; else delete c;
}

}  // namespace

extern "C" {

int uspmv_partition_precisions(uspmv_ctx *ctx, const uspmv_coo *coo, int ap_mode, double t1, double t2, const double *rowmax_h,
                               const double *colmax_h, uspmv_coo **dp, uspmv_coo **sp, uspmv_coo **hp) {
    return guarded([&] {
        if (!ctx || !coo) fail("uspmv_partition_precisions: NULL argument");
        if (ap_mode < 0 || ap_mode > 3) fail("uspmv_partition_precisions: invalid ap mode %d", ap_mode);
        if (ap_mode == USPMV_AP_DP_SP_HP && !(t1 > t2)) fail("uspmv_partition_precisions: thresholds must satisfy t1 > t2 (utilities.hpp:1514-1517)");
        if ((rowmax_h == nullptr) != (colmax_h == nullptr)) fail("uspmv_partition_precisions: rowmax and colmax must be given together");
        USPMV_CUDA(cudaSetDevice(ctx->device));
        if (dp) *dp = nullptr;
        if (sp) *sp = nullptr;
        if (hp) *hp = nullptr;
        DevBuf<double> rm, cm;
        if (rowmax_h) {
            rm.alloc(coo->n_rows);
            cm.alloc(coo->n_cols);
            USPMV_CUDA(cudaMemcpy(rm.p, rowmax_h, coo->n_rows * sizeof(double), cudaMemcpyHostToDevice));
            USPMV_CUDA(cudaMemcpy(cm.p, colmax_h, coo->n_cols * sizeof(double), cudaMemcpyHostToDevice));
        }
        switch (coo->mt) {
        case USPMV_F64: partition_typed<double>(coo, ap_mode, t1, t2, rm.p, cm.p, dp, sp, hp); break;
        case USPMV_F32: partition_typed<float>(coo, ap_mode, t1, t2, rm.p, cm.p, dp, sp, hp); break;
        default: partition_typed<__half>(coo, ap_mode, t1, t2, rm.p, cm.p, dp, sp, hp);
        }
    });
}

int uspmv_ap_spmv(int ap_mode, const uspmv_scs *dp, const uspmv_scs *sp, const uspmv_scs *hp, const void *x, void *y, void *stream) {
    return guarded([&] {
        const bool use_dp = ap_mode != USPMV_AP_SP_HP, use_sp = ap_mode != USPMV_AP_DP_HP, use_hp = ap_mode != USPMV_AP_DP_SP;
        if (ap_mode < 0 || ap_mode > 3) fail("uspmv_ap_spmv: invalid ap mode %d", ap_mode);
        if ((use_dp && !dp) || (use_sp && !sp) || (use_hp && !hp)) fail("uspmv_ap_spmv: a matrix part required by the mode is NULL");
        const uspmv_scs *first = use_dp ? dp : sp;
        auto check = [&](const uspmv_scs *s, int vt, const char *nm) {
            if (!s) return;
            if (s->vt != vt) fail("uspmv_ap_spmv: %s part has the wrong value type", nm);
            if (s->C != first->C || s->n_chunks != first->n_chunks) fail("uspmv_ap_spmv: %s part does not share C / n_chunks with the first part", nm);
        };
        check(use_dp ? dp : nullptr, USPMV_F64, "dp");
        check(use_sp ? sp : nullptr, USPMV_F32, "sp");
        check(use_hp ? hp : nullptr, USPMV_F16, "hp");
        const long n_pad = first->n_rows_padded;
        if (n_pad == 0) return;
        auto args = [](const uspmv_scs *s) {
            PartArgs a{nullptr, nullptr, nullptr, nullptr};
            if (s) a = PartArgs{s->chunk_ptrs.p, s->chunk_lengths.p, s->col_idxs.p, s->values.p};
            return a;
        };
        const PartArgs a = args(use_dp ? dp : nullptr), b = args(use_sp ? sp : nullptr), c = args(use_hp ? hp : nullptr);
        cudaStream_t st = as_stream(stream);
        if (first->C == 32 && options().scs_stream) {  // streamed one-pass kernel
            const stream::ApPart sa{a.cp, a.cl, a.ci, a.v}, sb{b.cp, b.cl, b.ci, b.v}, sc{c.cp, c.cl, c.ci, c.v};
            const int *order = first->balanced_order.p;
            switch (ap_mode) {
            case USPMV_AP_DP_SP: launch_ap_stream<USPMV_AP_DP_SP>(first->n_chunks, order, sa, sb, sc, x, y, st); break;
            case USPMV_AP_DP_HP: launch_ap_stream<USPMV_AP_DP_HP>(first->n_chunks, order, sa, sb, sc, x, y, st); break;
            case USPMV_AP_SP_HP: launch_ap_stream<USPMV_AP_SP_HP>(first->n_chunks, order, sa, sb, sc, x, y, st); break;
            default: launch_ap_stream<USPMV_AP_DP_SP_HP>(first->n_chunks, order, sa, sb, sc, x, y, st);
            }
            USPMV_LAUNCH_CHECK();
            return;
        }
        const unsigned g = blocks_for(n_pad);
        const int C = (int)first->C;
        switch (ap_mode) {
        case USPMV_AP_DP_SP: k_ap_spmv<USPMV_AP_DP_SP><<<g, TPB, 0, st>>>(n_pad, C, a, b, c, x, y); break;
        case USPMV_AP_DP_HP: k_ap_spmv<USPMV_AP_DP_HP><<<g, TPB, 0, st>>>(n_pad, C, a, b, c, x, y); break;
        case USPMV_AP_SP_HP: k_ap_spmv<USPMV_AP_SP_HP><<<g, TPB, 0, st>>>(n_pad, C, a, b, c, x, y); break;
        default: k_ap_spmv<USPMV_AP_DP_SP_HP><<<g, TPB, 0, st>>>(n_pad, C, a, b, c, x, y);
        }
        USPMV_LAUNCH_CHECK();
    });
}

}  // extern "C"
