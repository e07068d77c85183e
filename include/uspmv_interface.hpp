// uspmv_interface.hpp — C++ shim that re-creates the reference's library header (code/interface.hpp, API_doc.md)
// on top of the C ABI of libuspmv_b200.so.  An application that today does
//
//     #include "interface.hpp"
//     convert_to_scs(&mtx, C, sigma, &scs);  permute_scs_cols(&scs, scs.old_to_new_idx.data());
//     uspmv_scs_gpu<<<...>>>(...)            // or execute_uspmv(...)
//
// switches the include to this header and links -luspmv_b200: same struct fields (interface.hpp:16-80), same function
// names and argument order.  Differences, all deliberate:
//   * construction (convert_to_scs, permute_scs_cols, partition_precisions) runs on the GPU; the std::vector members of
//     ScsData are filled by exporting the device arrays (bit-exact), and `ScsData::device` keeps the device-resident
//     matrix so that kernels do not need another host->device copy;
//   * the kernels are host-callable functions taking DEVICE pointers (the reference's are __global__ templates that
//     the caller launches with <<<grid, 128>>>, interface.hpp:1741-1867);
//   * errors throw std::runtime_error with the library's message instead of printf + exit (interface.hpp:791-803).
// No CUDA headers are needed to compile against this file.
#ifndef USPMV_INTERFACE_HPP
#define USPMV_INTERFACE_HPP

#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

#include "uspmv_b200.h"

using ST = long;  // mmio.h:21

namespace uspmv_detail {
inline void check(int rc) {
    if (rc) throw std::runtime_error(uspmv_last_error());
}
template <typename VT> struct vt_of;
template <> struct vt_of<double> { static constexpr int value = USPMV_F64; };
template <> struct vt_of<float> { static constexpr int value = USPMV_F32; };
#if defined(__FLT16_MAX__)
template <> struct vt_of<_Float16> { static constexpr int value = USPMV_F16; };
#endif
struct uspmv_half_bits { unsigned short bits; };  // opaque fp16 storage for compilers without _Float16
template <> struct vt_of<uspmv_half_bits> { static constexpr int value = USPMV_F16; };

inline uspmv_ctx *default_ctx(int device = 0) {
    static uspmv_ctx *ctx = nullptr;
    if (!ctx) check(uspmv_ctx_create(device, &ctx));
    return ctx;
}
struct coo_deleter { void operator()(uspmv_coo *p) const { uspmv_coo_destroy(p); } };
struct scs_deleter { void operator()(uspmv_scs *p) const { uspmv_scs_destroy(p); } };
}  // namespace uspmv_detail

// interface.hpp:16-56
template <typename VT, typename IT>
struct MtxData {
    ST n_rows{};
    ST n_cols{};
    ST nnz{};
    bool is_sorted{};
    bool is_symmetric{};
    std::vector<IT> I;
    std::vector<IT> J;
    std::vector<VT> values;
};

// interface.hpp:58-80
template <typename VT, typename IT>
struct ScsData {
    ST C{};
    ST sigma{};
    ST n_rows{};
    ST n_cols{};
    ST n_rows_padded{};
    ST n_chunks{};
    ST n_elements{};  // No. of nz + padding.
    ST nnz{};         // No. of nz only.
    std::vector<IT> chunk_ptrs;     // Chunk start offsets into col_idxs & values.
    std::vector<IT> chunk_lengths;  // Length of one row in a chunk.
    std::vector<IT> col_idxs;
    std::vector<VT> values;
    std::vector<IT> old_to_new_idx;
    IT *new_to_old_idx{};  // raw pointer like the reference (utilities.hpp:2060-2069); owned by new_to_old_storage here
    std::vector<IT> new_to_old_storage;
    std::shared_ptr<uspmv_scs> device;  // device-resident twin (extension)
};

namespace uspmv_detail {
template <typename VT, typename IT>
void export_into(ScsData<VT, IT> *scs) {
    static_assert(std::is_same<IT, int>::value, "IT must be int (the reference's index type)");
    long d[8];
    check(uspmv_scs_dims(scs->device.get(), d));
    scs->C = d[0]; scs->sigma = d[1]; scs->n_rows = d[2]; scs->n_cols = d[3];
    scs->n_rows_padded = d[4]; scs->n_chunks = d[5]; scs->n_elements = d[6]; scs->nnz = d[7];
    // the reference allocates sigma entries of slack on every array (utilities.hpp:1944-1945,1974,1984-1985)
    scs->chunk_ptrs.assign(scs->n_chunks + 1 + scs->sigma, 0);
    scs->chunk_lengths.assign(scs->n_chunks + scs->sigma, 0);
    scs->col_idxs.assign(scs->n_elements + scs->sigma, 0);
    scs->values.assign(scs->n_elements + scs->sigma, VT{});
    scs->old_to_new_idx.assign(scs->n_rows + scs->sigma, 0);
    scs->new_to_old_storage.assign(scs->n_rows_padded + scs->sigma, 0);
    check(uspmv_scs_export(scs->device.get(), scs->chunk_ptrs.data(), scs->chunk_lengths.data(), scs->col_idxs.data(), scs->values.data(),
                           scs->old_to_new_idx.data(), scs->new_to_old_storage.data()));
    for (auto &v : scs->new_to_old_storage)
        if (v < 0) v = 0;  // positions no real row maps to are uninitialised in the reference; 0 is a safe index
    scs->new_to_old_idx = scs->new_to_old_storage.data();
}
}  // namespace uspmv_detail

// convert_to_scs — interface.hpp:401-656 / utilities.hpp:1842-2104
template <typename MT, typename VT, typename IT>
void convert_to_scs(const MtxData<MT, IT> *local_mtx, ST C, ST sigma, ScsData<VT, IT> *scs, int *fixed_permutation = NULL) {
    using namespace uspmv_detail;
    uspmv_ctx *ctx = default_ctx();
    uspmv_coo *coo_raw = nullptr;
    check(uspmv_coo_from_host(ctx, local_mtx->n_rows, local_mtx->n_cols, local_mtx->nnz, local_mtx->I.data(), local_mtx->J.data(),
                              local_mtx->values.data(), vt_of<MT>::value, &coo_raw));
    std::unique_ptr<uspmv_coo, coo_deleter> coo(coo_raw);
    uspmv_scs *s = nullptr;
    check(uspmv_scs_build(ctx, coo.get(), C, sigma, vt_of<VT>::value, fixed_permutation, &s));
    scs->device = std::shared_ptr<uspmv_scs>(s, scs_deleter());
    export_into(scs);
}

// permute_scs_cols — interface.hpp:659-688 / utilities.hpp:1802-1831
template <typename VT, typename IT>
void permute_scs_cols(ScsData<VT, IT> *scs, IT *perm) {
    using namespace uspmv_detail;
    if (!scs->device) throw std::runtime_error("permute_scs_cols: ScsData was not built by convert_to_scs");
    check(uspmv_scs_permute_cols(scs->device.get(), perm));
    check(uspmv_scs_export(scs->device.get(), nullptr, nullptr, scs->col_idxs.data(), nullptr, nullptr, nullptr));
}

// apply_permutation — interface.hpp:379-392 (HOST vectors; the gather itself runs on the device)
template <typename VT, typename IT>
void apply_permutation(VT *permuted_vec, VT *vec_to_permute, IT *perm, int num_elems_to_permute) {
    using namespace uspmv_detail;
    uspmv_ctx *ctx = default_ctx();
    const long n = num_elems_to_permute;
    if (n <= 0) return;
    long src_len = 0;
    for (long i = 0; i < n; ++i) src_len = perm[i] + 1 > src_len ? perm[i] + 1 : src_len;
    void *d_in = nullptr, *d_out = nullptr, *d_perm = nullptr;
    check(uspmv_malloc(ctx, src_len * sizeof(VT), &d_in));
    check(uspmv_malloc(ctx, n * sizeof(VT), &d_out));
    check(uspmv_malloc(ctx, n * sizeof(IT), &d_perm));
    check(uspmv_memcpy_h2d(ctx, d_in, vec_to_permute, src_len * sizeof(VT), nullptr));
    check(uspmv_memcpy_h2d(ctx, d_perm, perm, n * sizeof(IT), nullptr));
    check(uspmv_apply_permutation(ctx, d_out, d_in, static_cast<const int *>(d_perm), n, vt_of<VT>::value, nullptr));
    check(uspmv_memcpy_d2h(ctx, permuted_vec, d_out, n * sizeof(VT), nullptr));
    uspmv_free(ctx, d_in); uspmv_free(ctx, d_out); uspmv_free(ctx, d_perm);
}

// partition_precisions — interface.hpp:690-978.  The reference compares `char*` with string literals (dead branches);
// here ap_value_type is compared as a string, which is the evident intent.
template <typename VT, typename IT, typename HT = uspmv_detail::uspmv_half_bits>
void partition_precisions(MtxData<VT, IT> *local_mtx, MtxData<double, int> *dp_local_mtx, MtxData<float, int> *sp_local_mtx,
                          MtxData<HT, int> *hp_local_mtx, std::vector<VT> *largest_row_elems, std::vector<VT> *largest_col_elems,
                          double ap_threshold_1, double ap_threshold_2, const char *ap_value_type, bool is_equilibrated) {
    using namespace uspmv_detail;
    const std::string t(ap_value_type ? ap_value_type : "");
    int mode = t == "ap[dp_sp]" ? USPMV_AP_DP_SP : t == "ap[dp_hp]" ? USPMV_AP_DP_HP : t == "ap[sp_hp]" ? USPMV_AP_SP_HP
               : t == "ap[dp_sp_hp]" ? USPMV_AP_DP_SP_HP : -1;
    if (mode < 0) throw std::runtime_error("partition_precisions: unknown ap_value_type '" + t + "'");
    uspmv_ctx *ctx = default_ctx();
    uspmv_coo *coo_raw = nullptr;
    check(uspmv_coo_from_host(ctx, local_mtx->n_rows, local_mtx->n_cols, local_mtx->nnz, local_mtx->I.data(), local_mtx->J.data(),
                              local_mtx->values.data(), vt_of<VT>::value, &coo_raw));
    std::unique_ptr<uspmv_coo, coo_deleter> coo(coo_raw);
    std::vector<double> rm, cm;
    if (is_equilibrated) {
        rm.assign(largest_row_elems->begin(), largest_row_elems->end());
        cm.assign(largest_col_elems->begin(), largest_col_elems->end());
    }
    uspmv_coo *parts[3] = {nullptr, nullptr, nullptr};
    check(uspmv_partition_precisions(ctx, coo.get(), mode, ap_threshold_1, ap_threshold_2, is_equilibrated ? rm.data() : nullptr,
                                     is_equilibrated ? cm.data() : nullptr, &parts[0], &parts[1], &parts[2]));
    auto pull = [&](uspmv_coo *p, auto *out) {
        out->is_sorted = local_mtx->is_sorted; out->is_symmetric = local_mtx->is_symmetric;
        out->n_rows = local_mtx->n_rows; out->n_cols = local_mtx->n_cols; out->nnz = 0;
        out->I.clear(); out->J.clear(); out->values.clear();
        if (!p) return;
        std::unique_ptr<uspmv_coo, coo_deleter> guard(p);
        long d[3];
        check(uspmv_coo_dims(p, d));
        out->nnz = d[2];
        out->I.resize(d[2]); out->J.resize(d[2]); out->values.resize(d[2]);
        check(uspmv_coo_export(p, out->I.data(), out->J.data(), out->values.data()));
    };
    pull(parts[0], dp_local_mtx);
    pull(parts[1], sp_local_mtx);
    if (hp_local_mtx) pull(parts[2], hp_local_mtx);
    else if (parts[2]) uspmv_coo_destroy(parts[2]);
}

// uspmv_scs_gpu — interface.hpp:1766-1793 (x, y and the matrix arrays are DEVICE pointers; y has n_chunks*C entries)
template <typename VT, typename IT>
void uspmv_scs_gpu(const ST C, const ST n_chunks, const IT *chunk_ptrs, const IT *chunk_lengths, const IT *col_idxs, const VT *values,
                   const VT *x, VT *y, void *stream = nullptr) {
    using namespace uspmv_detail;
    check(::uspmv_scs_gpu(default_ctx(), vt_of<VT>::value, C, n_chunks, chunk_ptrs, chunk_lengths, col_idxs, values, x, y, stream));
}

// uspmv_csr_gpu — interface.hpp:1741-1760
template <typename VT, typename IT>
void uspmv_csr_gpu(const ST num_rows, const IT *row_ptrs, const IT *row_lengths, const IT *col_idxs, const VT *values, const VT *x, VT *y,
                   void *stream = nullptr) {
    using namespace uspmv_detail;
    (void)row_lengths;
    check(::uspmv_csr_gpu(default_ctx(), vt_of<VT>::value, num_rows, row_ptrs, col_idxs, values, x, y, stream));
}

// execute_uspmv — interface.hpp:1871-2187: SELL-C-sigma kernels iff C > 1 or sigma > 1, else CRS; with ap_value_type the
// fused adaptive-precision kernel over the dp/sp/hp parts.  x / y are DEVICE pointers.
template <typename VT, typename IT, typename HT = uspmv_detail::uspmv_half_bits>
void execute_uspmv(const ScsData<VT, IT> *scs, const VT *x, VT *y, const ScsData<double, IT> *dp = nullptr,
                   const ScsData<float, IT> *sp = nullptr, const ScsData<HT, IT> *hp = nullptr, const void *ap_x = nullptr,
                   void *ap_y = nullptr, const char *ap_value_type = nullptr, void *stream = nullptr) {
    using namespace uspmv_detail;
    const std::string t(ap_value_type ? ap_value_type : "");
    if (t.empty() || t == "dp" || t == "sp" || t == "hp") {
        check(uspmv_spmv(scs->device.get(), x, y, stream));
        return;
    }
    int mode = t == "ap[dp_sp]" ? USPMV_AP_DP_SP : t == "ap[dp_hp]" ? USPMV_AP_DP_HP : t == "ap[sp_hp]" ? USPMV_AP_SP_HP
               : t == "ap[dp_sp_hp]" ? USPMV_AP_DP_SP_HP : -1;
    if (mode < 0) throw std::runtime_error("execute_uspmv: unknown ap_value_type '" + t + "'");
    check(uspmv_ap_spmv(mode, dp ? dp->device.get() : nullptr, sp ? sp->device.get() : nullptr, hp ? hp->device.get() : nullptr, ap_x, ap_y,
                        stream));
}

#endif  // USPMV_INTERFACE_HPP
