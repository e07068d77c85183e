// Column-banded execution plan for matrices whose x does not fit the L2 (EXPERIMENTAL, off unless asked for).
//
// Why: on the 2^25-row power-law matrix of BASELINE config 4 every missing 8-byte x gather costs a 128-byte DRAM fill — ncu: 38 GB
// moved per SpMV for 7.5 GB of algorithmic bytes.  Cutting the COLUMNS into K bands whose x slice (n_cols / K
// values) stays L2-resident turns those fills into L2 hits; the price is one SELL-C-sigma structure per band (more padding: 1.7 x the
// slots at K = 8, scripts/model_column_bands.py) and one pass over y per band.
//
// What: nothing but a composition of the library's own, separately tested pieces —
//   build : sigma-sort permutation of the WHOLE matrix (uspmv_scs_build, then dropped)  ->  per band: order-preserving split of the COO
//           by column range  ->  [partition_precisions]  ->  uspmv_scs_build(..., fixed_permutation = that permutation)
//           (the reference's own mechanism for structures that must share a row order, main.cpp:1175-1219);
//   spmv  : band 0 straight into y, band b > 0 into a scratch vector and y += scratch (k_axpy).
// x stays in the ORIGINAL numbering (bands are ranges of original columns, like the AP structs, main.cpp:1308-1332), y comes out in
// the permuted row order of the plan (uspmv_banded_perm gives old_to_new).  The band sums are added in band order, so y differs
// from the sequential row sum in the last bits only (1e-12 / 1e-5 / 1e-2 relative to sum |a||x|, tested); the un-banded kernels
// remain the bit-identical default.
#include "common.cuh"

#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <vector>

using namespace uspmv;

struct uspmv_banded {
    uspmv_ctx *ctx = nullptr;
    int n_bands = 0, ap_mode = -1, vt = USPMV_F64;  // ap_mode < 0: one precision
    long n_rows = 0, n_cols = 0, n_rows_padded = 0, nnz = 0, n_elements = 0, band_width = 0;
    std::vector<uspmv_scs *> parts;  // [band * 3 + part]; one precision: part 0 only
    std::vector<int> old_to_new;     // row permutation shared by every band (host copy)
    DevBuf<unsigned char> scratch;   // n_rows_padded values of the y type
};

const uspmv_ctx *ctx_of(const uspmv_banded *b) { return b ? b->ctx : nullptr; }

namespace {

constexpr int TPB = 256;
inline unsigned blocks_for(long n) { return (unsigned)((n + TPB - 1) / TPB); }

__global__ void k_band_flag(const int *__restrict__ J, long nnz, int lo, int hi, int *__restrict__ flag) {
    const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i < nnz) flag[i] = J[i] >= lo && J[i] < hi;
}

template <int ES>
__global__ void k_band_scatter(const int *__restrict__ I, const int *__restrict__ J, const unsigned char *__restrict__ V, long nnz,
                               const int *__restrict__ flag, const int *__restrict__ pos, int *__restrict__ Io, int *__restrict__ Jo,
                               unsigned char *__restrict__ Vo) {
    const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i >= nnz || !flag[i]) return;
    const long k = pos[i];
    Io[k] = I[i];
    Jo[k] = J[i];
#pragma unroll
    for (int b = 0; b < ES; ++b) Vo[k * ES + b] = V[i * ES + b];
}

template <typename T>
__global__ void k_axpy(T *__restrict__ y, const T *__restrict__ t, long n);
template <>
__global__ void k_axpy<double>(double *__restrict__ y, const double *__restrict__ t, long n) {
    const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i < n) y[i] += t[i];
}
template <>
__global__ void k_axpy<float>(float *__restrict__ y, const float *__restrict__ t, long n) {
    const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i < n) y[i] += t[i];
}
template <>
__global__ void k_axpy<__half>(__half *__restrict__ y, const __half *__restrict__ t, long n) {
    const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i < n) y[i] = __float2half_rn(__half2float(y[i]) + __half2float(t[i]));
}

// entries of `coo` with lo <= column < hi, input order preserved (so the per-row order of every band is the COO order)
uspmv_coo *split_band(const uspmv_coo *coo, int lo, int hi) {
    const long nnz = coo->nnz;
    const size_t es = vt_size(coo->mt);
    DevBuf<int> flag(nnz + 1), pos(nnz + 1);
    long cnt = 0;
    if (nnz) {
        k_band_flag<<<blocks_for(nnz), TPB>>>(coo->J.p, nnz, lo, hi, flag.p);
        USPMV_LAUNCH_CHECK();
        size_t bytes = 0;
        USPMV_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, bytes, flag.p, pos.p, (int)nnz));
        DevBuf<unsigned char> tmp(bytes);
        USPMV_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, bytes, flag.p, pos.p, (int)nnz));
        g_launches.fetch_add(2);
        int last_pos = 0, last_flag = 0;
        USPMV_CUDA(cudaMemcpy(&last_pos, pos.p + nnz - 1, sizeof(int), cudaMemcpyDeviceToHost));
        USPMV_CUDA(cudaMemcpy(&last_flag, flag.p + nnz - 1, sizeof(int), cudaMemcpyDeviceToHost));
        cnt = (long)last_pos + last_flag;
    }
    auto out = new uspmv_coo();
    try {
        out->ctx = coo->ctx;
        out->n_rows = coo->n_rows;
        out->n_cols = coo->n_cols;
        out->nnz = cnt;
        out->mt = coo->mt;
        out->I.alloc(std::max<long>(cnt, 1));
        out->J.alloc(std::max<long>(cnt, 1));
        out->values.alloc(std::max<size_t>((size_t)cnt * es, 16));
        if (nnz && cnt) {
            const unsigned g = blocks_for(nnz);
            switch (es) {
            case 8: k_band_scatter<8><<<g, TPB>>>(coo->I.p, coo->J.p, coo->values.p, nnz, flag.p, pos.p, out->I.p, out->J.p, out->values.p); break;
            case 4: k_band_scatter<4><<<g, TPB>>>(coo->I.p, coo->J.p, coo->values.p, nnz, flag.p, pos.p, out->I.p, out->J.p, out->values.p); break;
            default: k_band_scatter<2><<<g, TPB>>>(coo->I.p, coo->J.p, coo->values.p, nnz, flag.p, pos.p, out->I.p, out->J.p, out->values.p);
            }
            USPMV_LAUNCH_CHECK();
        }
        USPMV_CUDA(cudaDeviceSynchronize());
    } catch (...) { delete out; throw; }
    return out;
}

void ck(int rc) {
    if (rc) throw Error(uspmv_last_error());
}

}  // namespace

extern "C" {

void uspmv_banded_destroy(uspmv_banded *b) {
    if (!b) return;
    for (uspmv_scs *s : b->parts)
        if (s) uspmv_scs_destroy(s);
    delete b;
}

/* Column-banded plan (see the header of this file).  ap_mode < 0: one precision `vt`; else USPMV_AP_* with thresholds t1 / t2 (the
 * value types of the parts are fixed by the mode).  n_bands == 0 picks the number of bands so that one band of x is about 32 MB. */
int uspmv_banded_build(uspmv_ctx *ctx, const uspmv_coo *coo, long C, long sigma, int vt, int ap_mode, double t1, double t2, int n_bands,
                       uspmv_banded **out) {
    return guarded([&] {
        OptScope opt_scope(ctx_of(ctx));  // this context's options govern everything below
        if (!ctx || !coo || !out) fail("uspmv_banded_build: NULL argument");
        if (ap_mode > 3) fail("uspmv_banded_build: invalid ap mode %d", ap_mode);
        if (n_bands < 0 || n_bands > 256) fail("uspmv_banded_build: n_bands must be in [0,256]");
        if (coo->n_cols < 1) fail("uspmv_banded_build: matrix has no columns");
        USPMV_CUDA(cudaSetDevice(ctx->device));
        const bool ap = ap_mode >= 0;
        const int xvt = ap ? (ap_mode == USPMV_AP_SP_HP ? USPMV_F32 : USPMV_F64) : vt;
        if (n_bands == 0) {
            const double x_mb = (double)coo->n_cols * (double)vt_size(xvt) / (1024.0 * 1024.0);
            n_bands = (int)std::min(256.0, std::max(1.0, std::ceil(x_mb / 32.0)));
        }
        n_bands = (int)std::min<long>(n_bands, coo->n_cols);
        auto b = new uspmv_banded();
        try {
            b->ctx = ctx; b->n_bands = n_bands; b->ap_mode = ap_mode; b->vt = xvt;
            b->n_rows = coo->n_rows; b->n_cols = coo->n_cols; b->nnz = coo->nnz;
            b->band_width = (coo->n_cols + n_bands - 1) / n_bands;
            b->parts.assign((size_t)n_bands * 3, nullptr);
            const int first = ap && ap_mode == USPMV_AP_SP_HP ? 1 : 0;
            const int vts[3] = {USPMV_F64, USPMV_F32, USPMV_F16};
            // (1) the row order: sigma-sort of the WHOLE matrix (AP: of its first part, as in the un-banded AP path)
            {
                uspmv_coo *pc[3] = {nullptr, nullptr, nullptr};
                const uspmv_coo *src = coo;
                if (ap) {
                    ck(uspmv_partition_precisions(ctx, coo, ap_mode, t1, t2, nullptr, nullptr, &pc[0], &pc[1], &pc[2]));
                    src = pc[first];
                }
                uspmv_scs *full = nullptr;
                const int rc = uspmv_scs_build(ctx, src, C, sigma, ap ? vts[first] : vt, nullptr, &full);
                for (uspmv_coo *q : pc)
                    if (q) uspmv_coo_destroy(q);
                ck(rc);
                b->n_rows_padded = full->n_rows_padded;
                b->old_to_new.resize(std::max<long>(coo->n_rows, 1));
                const int rc2 = uspmv_scs_export(full, nullptr, nullptr, nullptr, nullptr, b->old_to_new.data(), nullptr);
                uspmv_scs_destroy(full);
                ck(rc2);
            }
            // (2) one structure (AP: one per part) per column band, all on that row order
            for (int k = 0; k < n_bands; ++k) {
                const long lo = (long)k * b->band_width, hi = std::min<long>(lo + b->band_width, coo->n_cols);
                uspmv_coo *band = split_band(coo, (int)lo, (int)hi);
                uspmv_coo *pc[3] = {nullptr, nullptr, nullptr};
                try {
                    if (!ap) {
                        ck(uspmv_scs_build(ctx, band, C, sigma, vt, b->old_to_new.data(), &b->parts[(size_t)k * 3]));
                        b->n_elements += b->parts[(size_t)k * 3]->n_elements;
                    } else {
                        ck(uspmv_partition_precisions(ctx, band, ap_mode, t1, t2, nullptr, nullptr, &pc[0], &pc[1], &pc[2]));
                        for (int p = 0; p < 3; ++p)
                            if (pc[p]) {
                                ck(uspmv_scs_build(ctx, pc[p], C, sigma, vts[p], b->old_to_new.data(), &b->parts[(size_t)k * 3 + p]));
                                b->n_elements += b->parts[(size_t)k * 3 + p]->n_elements;
                            }
                    }
                } catch (...) {
                    for (uspmv_coo *q : pc)
                        if (q) uspmv_coo_destroy(q);
                    uspmv_coo_destroy(band);
                    throw;
                }
                for (uspmv_coo *q : pc)
                    if (q) uspmv_coo_destroy(q);
                uspmv_coo_destroy(band);
            }
            if (n_bands > 1) b->scratch.alloc((size_t)std::max<long>(b->n_rows_padded, 1) * vt_size(xvt));
        } catch (...) { uspmv_banded_destroy(b); throw; }
        *out = b;
    });
}

/* out8 = n_bands, band_width, n_rows, n_cols, n_rows_padded, nnz, n_elements (all bands and parts), value type of x / y */
int uspmv_banded_dims(const uspmv_banded *b, long out8[8]) {
    return guarded([&] {
        OptScope opt_scope(ctx_of(b));  // this context's options govern everything below
        if (!b || !out8) fail("uspmv_banded_dims: NULL argument");
        out8[0] = b->n_bands; out8[1] = b->band_width; out8[2] = b->n_rows; out8[3] = b->n_cols; out8[4] = b->n_rows_padded;
        out8[5] = b->nnz; out8[6] = b->n_elements; out8[7] = b->vt;
    });
}

/* old_to_new of the plan (n_rows ints): y_user[i] = y[old_to_new[i]] */
int uspmv_banded_perm(const uspmv_banded *b, int *old_to_new_h) {
    return guarded([&] {
        OptScope opt_scope(ctx_of(b));  // this context's options govern everything below
        if (!b || !old_to_new_h) fail("uspmv_banded_perm: NULL argument");
        std::copy(b->old_to_new.begin(), b->old_to_new.begin() + b->n_rows, old_to_new_h);
    });
}

/* y (n_rows_padded values, permuted row order) = A x, x in the original column numbering (n_cols values) */
int uspmv_banded_spmv(const uspmv_banded *b, const void *x, void *y, void *stream) {
    return guarded([&] {
        OptScope opt_scope(ctx_of(b));  // this context's options govern everything below
        if (!b || !x || !y) fail("uspmv_banded_spmv: NULL argument");
        cudaStream_t st = as_stream(stream);
        const long n = b->n_rows_padded;
        for (int k = 0; k < b->n_bands; ++k) {
            void *dst = k == 0 ? y : static_cast<void *>(b->scratch.p);
            uspmv_scs *const *p = &b->parts[(size_t)k * 3];
            if (b->ap_mode < 0) ck(uspmv_spmv(p[0], x, dst, stream));
            else ck(uspmv_ap_spmv(b->ap_mode, p[0], p[1], p[2], x, dst, stream));
            if (k == 0 || n == 0) continue;
            switch (b->vt) {
            case USPMV_F64: k_axpy<double><<<blocks_for(n), TPB, 0, st>>>(static_cast<double *>(y), reinterpret_cast<const double *>(b->scratch.p), n); break;
            case USPMV_F32: k_axpy<float><<<blocks_for(n), TPB, 0, st>>>(static_cast<float *>(y), reinterpret_cast<const float *>(b->scratch.p), n); break;
            default: k_axpy<__half><<<blocks_for(n), TPB, 0, st>>>(static_cast<__half *>(y), reinterpret_cast<const __half *>(b->scratch.p), n);
            }
            USPMV_LAUNCH_CHECK();
        }
    });
}

}  // extern "C"
