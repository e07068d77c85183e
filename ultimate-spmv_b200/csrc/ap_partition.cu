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

#include <algorithm>
#include <vector>

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

__global__ void k_diff_lengths(const int *__restrict__ ptrs, long n, int *__restrict__ len) {
    long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i < n) len[i] = ptrs[i + 1] - ptrs[i];
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

// Work items for very uneven matrices (power-law rows): a chunk whose parts together hold more than L slots is cut, part by
// part, into segments of <= L slots; every segment is its own work item (k_scs32_stream_ap writes its sum to a partial buffer,
// k_reduce_partials_ap adds a part's segments in slot order).  All items longest first.  Without this one warp walks a 4096-slot
// chunk alone (~0.4 ms) while the rest of the GPU idles (ncu r01j: L1 busy 53 % of active but 27 % of elapsed cycles).
const uspmv_scs::ApPlan *ap_plan_for(int mode, const uspmv_scs *first, const uspmv_scs *o1, const uspmv_scs *o2) {
    const int L = options().split_long_chunks;
    std::lock_guard<std::mutex> guard(first->ap_plan_mutex);  // built on first use, possibly from several threads
    uspmv_scs::ApPlan *pl = first->ap_plan.get();
    const long ne1 = o1 ? o1->n_elements : -1, ne2 = o2 ? o2->n_elements : -1;
    // keyed on the other parts' GENERATION ids: an address can be reused by a different matrix after uspmv_scs_destroy
    const unsigned long long g1 = o1 ? o1->generation : 0, g2 = o2 ? o2->generation : 0;
    if (pl && pl->mode == mode && pl->other[0] == g1 && pl->other[1] == g2 && pl->other_ne[0] == ne1 && pl->other_ne[1] == ne2 && pl->seg_slots == L)
        return pl;
    first->ap_plan.reset(new uspmv_scs::ApPlan());
    pl = first->ap_plan.get();
    pl->mode = mode; pl->other[0] = g1; pl->other[1] = g2; pl->other_ne[0] = ne1; pl->other_ne[1] = ne2; pl->seg_slots = L;
    const long nc = first->n_chunks;
    if (L <= 0 || nc < 4096) return pl;
    // parts in dp, sp, hp order; `first` is the dp part (or sp for sp_hp)
    const uspmv_scs *part[3] = {nullptr, nullptr, nullptr};
    if (mode == USPMV_AP_SP_HP) { part[1] = first; part[2] = o2; }
    else { part[0] = first; part[1] = o1; part[2] = o2; }
    std::vector<int> len[3];
    std::vector<long> tot(nc, 0);
    long slots_all = 0;
    for (int q = 0; q < 3; ++q) {
        if (!part[q]) continue;
        len[q].resize(nc);
        USPMV_CUDA(cudaMemcpy(len[q].data(), part[q]->chunk_lengths.p, nc * sizeof(int), cudaMemcpyDeviceToHost));
        for (long c = 0; c < nc; ++c) { tot[c] += len[q][c]; slots_all += len[q][c]; }
    }
    long mx = 0;
    for (long c = 0; c < nc; ++c) mx = std::max(mx, tot[c]);
    const double mean = (double)slots_all / (double)nc;
    if (mx <= L || mx < 8.0 * mean) return pl;
    std::vector<int4> items;
    std::vector<int> split_chunk, split_ptr{0};
    std::vector<unsigned char> seg_part;
    items.reserve(nc + nc / 8);
    for (long c = 0; c < nc; ++c) {
        if (tot[c] <= L) { items.push_back(make_int4((int)c, 0, (int)tot[c], -1)); continue; }
        for (int q = 0; q < 3; ++q) {
            if (!part[q]) continue;
            for (int j0 = 0; j0 < len[q][c]; j0 += L) {
                const int pslot = (int)seg_part.size();
                if (pslot >= (1 << 28)) fail("uspmv_ap_spmv: too many chunk segments");
                items.push_back(make_int4((int)c, j0, std::min(L, len[q][c] - j0), (pslot << 2) | q));
                seg_part.push_back((unsigned char)q);
            }
        }
        split_chunk.push_back((int)c);
        split_ptr.push_back((int)seg_part.size());
    }
    std::stable_sort(items.begin(), items.end(), [](const int4 &a, const int4 &b) { return a.z > b.z; });
    pl->n_items = (long)items.size();
    pl->n_split = (long)split_chunk.size();
    pl->items.alloc(items.size());
    pl->split_chunk.alloc(split_chunk.size());
    pl->split_ptr.alloc(split_ptr.size());
    pl->seg_part.alloc(seg_part.size());
    pl->partials.alloc(seg_part.size() * 32);
    USPMV_CUDA(cudaMemcpy(pl->items.p, items.data(), items.size() * sizeof(int4), cudaMemcpyHostToDevice));
    USPMV_CUDA(cudaMemcpy(pl->split_chunk.p, split_chunk.data(), split_chunk.size() * sizeof(int), cudaMemcpyHostToDevice));
    USPMV_CUDA(cudaMemcpy(pl->split_ptr.p, split_ptr.data(), split_ptr.size() * sizeof(int), cudaMemcpyHostToDevice));
    USPMV_CUDA(cudaMemcpy(pl->seg_part.p, seg_part.data(), seg_part.size(), cudaMemcpyHostToDevice));
    pl->use = true;
    return pl;
}

template <int MODE, int D, int MINB>
void launch_ap_stream_v(long n_chunks, const int *order, const uspmv_scs::ApPlan *pl, const stream::ApPart &a, const stream::ApPart &b,
                        const stream::ApPart &c, const void *x, void *y, cudaStream_t st) {
    constexpr int LMAX = 8, WARPS = 8;
    using R = stream::WarpRing<double, LMAX, D>;
    auto kern = stream::k_scs32_stream_ap<MODE, LMAX, D, WARPS, MINB>;
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
    const bool split = pl && pl->use;
    const long n_items = split ? pl->n_items : n_chunks;
    long grid = (long)sm_count(dev) * bps;
    const long need = (n_items + WARPS - 1) / WARPS;
    if (grid > need) grid = need;
    kern<<<(unsigned)grid, WARPS * 32, smem, st>>>(n_items, order, split ? pl->items.p : nullptr, a, b, c, x, y, split ? pl->partials.p : nullptr);
    if (split && pl->n_split) {
        USPMV_LAUNCH_CHECK();
        stream::k_reduce_partials_ap<MODE><<<(unsigned)((pl->n_split * 32 + 255) / 256), 256, 0, st>>>(pl->n_split, pl->split_chunk.p, pl->split_ptr.p,
                                                                                                     pl->seg_part.p, pl->partials.p, y);
    }
}

// ap_variant: ring depth x register cap (1 = uncapped ~80 registers / 24 warps per SM, 4 = 64 registers / 32 warps per SM)
template <int MODE>
void launch_ap_stream(long n_chunks, const int *order, const uspmv_scs::ApPlan *pl, const stream::ApPart &a, const stream::ApPart &b,
                      const stream::ApPart &c, const void *x, void *y, cudaStream_t st) {
    switch (options().ap_variant) {
    case 1: launch_ap_stream_v<MODE, 4, 1>(n_chunks, order, pl, a, b, c, x, y, st); break;
    case 2: launch_ap_stream_v<MODE, 2, 4>(n_chunks, order, pl, a, b, c, x, y, st); break;
    case 3: launch_ap_stream_v<MODE, 3, 3>(n_chunks, order, pl, a, b, c, x, y, st); break;
    case 4: launch_ap_stream_v<MODE, 2, 1>(n_chunks, order, pl, a, b, c, x, y, st); break;
    default: launch_ap_stream_v<MODE, 2, 3>(n_chunks, order, pl, a, b, c, x, y, st); break;  // <= 80 registers: 24 warps per SM
    }
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
        OptScope opt_scope(ctx_of(ctx));  // this context's options govern everything below
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
        OptScope opt_scope(ctx_of(dp ? dp : sp));  // this context's options govern everything below
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
        use_device(first->ctx);
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
            const uspmv_scs::ApPlan *pl = ap_plan_for(ap_mode, first, use_dp && use_sp ? sp : nullptr, use_hp ? hp : nullptr);
            switch (ap_mode) {
            case USPMV_AP_DP_SP: launch_ap_stream<USPMV_AP_DP_SP>(first->n_chunks, order, pl, sa, sb, sc, x, y, st); break;
            case USPMV_AP_DP_HP: launch_ap_stream<USPMV_AP_DP_HP>(first->n_chunks, order, pl, sa, sb, sc, x, y, st); break;
            case USPMV_AP_SP_HP: launch_ap_stream<USPMV_AP_SP_HP>(first->n_chunks, order, pl, sa, sb, sc, x, y, st); break;
            default: launch_ap_stream<USPMV_AP_DP_SP_HP>(first->n_chunks, order, pl, sa, sb, sc, x, y, st);
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

/* uspmv_scs_ap{dpsp,dphp,sphp,dpsphp}_gpu / uspmv_csr_ap*_gpu on caller-owned DEVICE arrays — the reference's AP kernels take the raw
 * arrays of every precision part per call (interface.hpp:1129-1733; GPU: ap_kernels.hpp:637-953, MultiPrecKernelArgs
 * classes_structs.hpp:238-261).  arrays[4 * p + {0,1,2,3}] = chunk_ptrs, chunk_lengths, col_idxs, values of part p (0 dp, 1 sp, 2 hp;
 * parts the mode does not use are ignored).  All parts share C and n_chunks.  C == 1 is CRS (chunk_ptrs = row_ptrs; a NULL
 * chunk_lengths is derived).  One fused pass; very long chunks are NOT cut into segments here (no per-matrix plan without a handle). */
int uspmv_scs_ap_gpu(uspmv_ctx *ctx, int ap_mode, long C, long n_chunks, const void *const *arrays, const void *x, void *y, void *stream) {
    return guarded([&] {
        OptScope opt_scope(ctx_of(ctx));  // this context's options govern everything below
        if (!ctx || !arrays) fail("uspmv_scs_ap_gpu: NULL argument");
        if (ap_mode < 0 || ap_mode > 3) fail("uspmv_scs_ap_gpu: invalid ap mode %d", ap_mode);
        if (C < 1) fail("uspmv_scs_ap_gpu: C must be >= 1");
        const long n_pad = n_chunks * C;
        if (n_pad == 0) return;
        if (!x || !y) fail("uspmv_scs_ap_gpu: NULL vector");
        USPMV_CUDA(cudaSetDevice(ctx->device));
        cudaStream_t st = as_stream(stream);
        const bool use[3] = {ap_mode != USPMV_AP_SP_HP, ap_mode != USPMV_AP_DP_HP, ap_mode != USPMV_AP_DP_SP};
        PartArgs pa[3];
        DevBuf<int> cl_tmp[3];
        bool aligned = true, derived = false;
        for (int p = 0; p < 3; ++p) {
            pa[p] = PartArgs{nullptr, nullptr, nullptr, nullptr};
            if (!use[p]) continue;
            const int *cp = static_cast<const int *>(arrays[4 * p]), *cl = static_cast<const int *>(arrays[4 * p + 1]);
            const int *ci = static_cast<const int *>(arrays[4 * p + 2]);
            const void *v = arrays[4 * p + 3];
            if (!cp || !ci || !v) fail("uspmv_scs_ap_gpu: an array of part %d (required by the mode) is NULL", p);
            if (!cl) {
                if (C != 1) fail("uspmv_scs_ap_gpu: chunk_lengths of part %d is NULL", p);
                cl_tmp[p].alloc(n_chunks);
                k_diff_lengths<<<blocks_for(n_chunks), TPB, 0, st>>>(cp, n_chunks, cl_tmp[p].p);
                USPMV_LAUNCH_CHECK();
                cl = cl_tmp[p].p;
                derived = true;
            }
            pa[p] = PartArgs{cp, cl, ci, v};
            aligned = aligned && reinterpret_cast<uintptr_t>(ci) % 16 == 0 && reinterpret_cast<uintptr_t>(v) % 16 == 0;
        }
        if (C == 32 && options().scs_stream && aligned) {
            const stream::ApPart sa{pa[0].cp, pa[0].cl, pa[0].ci, pa[0].v}, sb{pa[1].cp, pa[1].cl, pa[1].ci, pa[1].v}, sc{pa[2].cp, pa[2].cl, pa[2].ci, pa[2].v};
            switch (ap_mode) {
            case USPMV_AP_DP_SP: launch_ap_stream<USPMV_AP_DP_SP>(n_chunks, nullptr, nullptr, sa, sb, sc, x, y, st); break;
            case USPMV_AP_DP_HP: launch_ap_stream<USPMV_AP_DP_HP>(n_chunks, nullptr, nullptr, sa, sb, sc, x, y, st); break;
            case USPMV_AP_SP_HP: launch_ap_stream<USPMV_AP_SP_HP>(n_chunks, nullptr, nullptr, sa, sb, sc, x, y, st); break;
            default: launch_ap_stream<USPMV_AP_DP_SP_HP>(n_chunks, nullptr, nullptr, sa, sb, sc, x, y, st);
            }
            USPMV_LAUNCH_CHECK();
        } else {
            const unsigned g = blocks_for(n_pad);
            switch (ap_mode) {
            case USPMV_AP_DP_SP: k_ap_spmv<USPMV_AP_DP_SP><<<g, TPB, 0, st>>>(n_pad, (int)C, pa[0], pa[1], pa[2], x, y); break;
            case USPMV_AP_DP_HP: k_ap_spmv<USPMV_AP_DP_HP><<<g, TPB, 0, st>>>(n_pad, (int)C, pa[0], pa[1], pa[2], x, y); break;
            case USPMV_AP_SP_HP: k_ap_spmv<USPMV_AP_SP_HP><<<g, TPB, 0, st>>>(n_pad, (int)C, pa[0], pa[1], pa[2], x, y); break;
            default: k_ap_spmv<USPMV_AP_DP_SP_HP><<<g, TPB, 0, st>>>(n_pad, (int)C, pa[0], pa[1], pa[2], x, y);
            }
            USPMV_LAUNCH_CHECK();
        }
        if (derived) USPMV_CUDA(cudaStreamSynchronize(st));  // the temporaries are freed on return
    });
}

}  // extern "C"
