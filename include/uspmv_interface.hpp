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

#include <algorithm>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

#include "uspmv_detail.hpp"



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
template <typename VT, typename IT>
void register_twin(const ScsData<VT, IT> *scs) {
    long mx = -1;
    for (ST e = 0; e < scs->n_elements; ++e) mx = scs->col_idxs[e] > mx ? scs->col_idxs[e] : mx;
    std::lock_guard<std::mutex> g(twins_mutex());
    twin &t = twins()[scs->values.data()];
    t.dev = scs->device; t.adopted.reset();
    t.n_chunks = scs->n_chunks; t.n_elements = scs->n_elements; t.x_len = std::max<long>(mx + 1, scs->n_cols);  // x has n_cols entries
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
    register_twin(scs);
}

// permute_scs_cols — interface.hpp:659-688 / utilities.hpp:1802-1831
template <typename VT, typename IT>
void permute_scs_cols(ScsData<VT, IT> *scs, IT *perm) {
    using namespace uspmv_detail;
    if (!scs->device) throw std::runtime_error("permute_scs_cols: ScsData was not built by convert_to_scs");
    check(uspmv_scs_permute_cols(scs->device.get(), perm));
    check(uspmv_scs_export(scs->device.get(), nullptr, nullptr, scs->col_idxs.data(), nullptr, nullptr, nullptr));
    register_twin(scs);
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

// apply_strided_permutation — utilities.hpp:1784-1799 (HOST vectors): permuted_vec[i*stride] = vec_to_permute[perm[i]*stride]
template <typename VT, typename IT>
void apply_strided_permutation(VT *permuted_vec, VT *vec_to_permute, IT *perm, int num_elems_to_permute, int stride) {
    using namespace uspmv_detail;
    uspmv_ctx *ctx = default_ctx();
    const long n = num_elems_to_permute;
    if (n <= 0) return;
    long src_rows = 0;
    for (long i = 0; i < n; ++i) src_rows = perm[i] + 1 > src_rows ? perm[i] + 1 : src_rows;
    const long src_len = (src_rows - 1) * stride + 1, dst_len = (n - 1) * stride + 1;
    void *d_in = nullptr, *d_out = nullptr, *d_perm = nullptr;
    check(uspmv_malloc(ctx, src_len * sizeof(VT), &d_in));
    check(uspmv_malloc(ctx, dst_len * sizeof(VT), &d_out));
    check(uspmv_malloc(ctx, n * sizeof(IT), &d_perm));
    check(uspmv_memcpy_h2d(ctx, d_in, vec_to_permute, src_len * sizeof(VT), nullptr));
    check(uspmv_memcpy_h2d(ctx, d_out, permuted_vec, dst_len * sizeof(VT), nullptr));  // the elements between the strides are kept
    check(uspmv_memcpy_h2d(ctx, d_perm, perm, n * sizeof(IT), nullptr));
    check(uspmv_apply_strided_permutation(ctx, d_out, d_in, static_cast<const int *>(d_perm), n, stride, vt_of<VT>::value, nullptr));
    check(uspmv_memcpy_d2h(ctx, permuted_vec, d_out, dst_len * sizeof(VT), nullptr));
    uspmv_free(ctx, d_in); uspmv_free(ctx, d_out); uspmv_free(ctx, d_perm);
}

// generate_inv_perm — utilities.hpp:1755-1766 (HOST arrays; tiny, done in place on the host like the reference)
template <typename IT>
void generate_inv_perm(int *perm, int *inv_perm, int perm_len) {
    for (int i = 0; i < perm_len; ++i) inv_perm[perm[i]] = i;
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

// ---------------------------------------------------------------------------------------------------------------------------
// execute_uspmv with the reference's OWN signatures (interface.hpp:1871-1910): a pointer bundle of 8 entries per precision and
// `char *ap_value_type`.  The reference has three spellings, selected by its USE_AP / HAVE_HALF_MATH macros; all three exist here as
// overloads, so a call site compiles unchanged whichever way the reference was configured.
//   * arrays on the HOST (what a user of the reference's CPU library passes): the matrix runs from its device twin — registered by
//     convert_to_scs, or adopted from the raw arrays on first use — and x / y are staged through the device (uspmv_spmv_host);
//   * arrays on the DEVICE (what the reference's GPU launchers get, classes_structs.hpp:213-261): the raw-array kernels directly.
// Kernel choice as in the reference (interface.hpp:1911): SELL-C-sigma kernels iff C > 1 (CHUNK_SIZE / SIGMA are macros there; with
// C == 1 the arrays ARE the CRS arrays), else CRS.  ap_value_type is compared as a string (the reference compares pointers).
// ---------------------------------------------------------------------------------------------------------------------------


// (1) the reference built without USE_AP
template <typename VT, typename IT>
void execute_uspmv(const ST *C, const ST *n_chunks, const IT *chunk_ptrs, const IT *chunk_lengths, const IT *col_idxs, const VT *values, VT *x, VT *y,
                   char *ap_value_type) {
    (void)ap_value_type;
    uspmv_detail::execute_one_prec<VT, IT>(C, n_chunks, chunk_ptrs, chunk_lengths, col_idxs, values, x, y);
}

// (2) USE_AP and HAVE_HALF_MATH: dp, sp and hp bundles (HT: _Float16 where the compiler has it, else the opaque 16-bit storage type)
template <typename VT, typename IT, typename HT>
void execute_uspmv(const ST *C, const ST *n_chunks, const IT *chunk_ptrs, const IT *chunk_lengths, const IT *col_idxs, const VT *values, VT *x, VT *y,
                   const ST *dp_C, const ST *dp_n_chunks, const IT *dp_chunk_ptrs, const IT *dp_chunk_lengths, const IT *dp_col_idxs,
                   const double *dp_values, double *dp_x, double *dp_y,
                   const ST *sp_C, const ST *sp_n_chunks, const IT *sp_chunk_ptrs, const IT *sp_chunk_lengths, const IT *sp_col_idxs,
                   const float *sp_values, float *sp_x, float *sp_y,
                   const ST *hp_C, const ST *hp_n_chunks, const IT *hp_chunk_ptrs, const IT *hp_chunk_lengths, const IT *hp_col_idxs,
                   const HT *hp_values, HT *hp_x, HT *hp_y, char *ap_value_type) {
    (void)hp_C; (void)hp_n_chunks; (void)hp_x; (void)hp_y;
    const int mode = uspmv_detail::ap_mode_of(ap_value_type);
    if (mode < 0) {
        uspmv_detail::execute_one_prec<VT, IT>(C, n_chunks, chunk_ptrs, chunk_lengths, col_idxs, values, x, y);
        return;
    }
    uspmv_detail::execute_ap<IT, HT>(mode, dp_C, dp_n_chunks, dp_chunk_ptrs, dp_chunk_lengths, dp_col_idxs, dp_values, dp_x, dp_y, sp_C, sp_n_chunks,
                                     sp_chunk_ptrs, sp_chunk_lengths, sp_col_idxs, sp_values, sp_x, sp_y, hp_chunk_ptrs, hp_chunk_lengths, hp_col_idxs,
                                     hp_values);
}

// (3) USE_AP without HAVE_HALF_MATH: dp and sp bundles only
template <typename VT, typename IT>
void execute_uspmv(const ST *C, const ST *n_chunks, const IT *chunk_ptrs, const IT *chunk_lengths, const IT *col_idxs, const VT *values, VT *x, VT *y,
                   const ST *dp_C, const ST *dp_n_chunks, const IT *dp_chunk_ptrs, const IT *dp_chunk_lengths, const IT *dp_col_idxs,
                   const double *dp_values, double *dp_x, double *dp_y,
                   const ST *sp_C, const ST *sp_n_chunks, const IT *sp_chunk_ptrs, const IT *sp_chunk_lengths, const IT *sp_col_idxs,
                   const float *sp_values, float *sp_x, float *sp_y, char *ap_value_type) {
    const int mode = uspmv_detail::ap_mode_of(ap_value_type);
    if (mode < 0) {
        uspmv_detail::execute_one_prec<VT, IT>(C, n_chunks, chunk_ptrs, chunk_lengths, col_idxs, values, x, y);
        return;
    }
    if (mode != USPMV_AP_DP_SP) throw std::runtime_error("execute_uspmv: this overload (no hp bundle) only serves ap[dp_sp]");
    uspmv_detail::execute_ap<IT, uspmv_detail::uspmv_half_bits>(mode, dp_C, dp_n_chunks, dp_chunk_ptrs, dp_chunk_lengths, dp_col_idxs, dp_values, dp_x,
                                                               dp_y, sp_C, sp_n_chunks, sp_chunk_ptrs, sp_chunk_lengths, sp_col_idxs, sp_values, sp_x,
                                                               sp_y, nullptr, nullptr, nullptr, nullptr);
}

#endif  // USPMV_INTERFACE_HPP
