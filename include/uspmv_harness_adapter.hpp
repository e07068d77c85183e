// uspmv_harness_adapter.hpp — launchers with EXACTLY the two std::function signatures through which the reference harness calls every
// kernel (code/classes_structs.hpp:283-333):
//
//   SpmvKernel<VT,IT>::OnePrecFuncPtr    (bool warmup_flag, const ST *C, const ST *n_chunks, const IT *chunk_ptrs, const IT *chunk_lengths,
//                                         const IT *col_idxs, const VT *values, VT *x, VT *y, int *block_vec_size, int *vec_length,
//                                         [const ST n_thread_blocks,]  const int *my_rank)
//   SpmvKernel<VT,IT>::MultiPrecFuncPtr  (bool, dp bundle (8), sp bundle (8), [hp bundle (8),] [const ST n_thread_blocks,] const int *)
//
// `n_thread_blocks` exists in the typedefs only under __CUDACC__, the hp bundle only under HAVE_HALF_MATH; the launchers below follow
// the SAME two preprocessor conditions, so they are assignable to the typedefs however the harness is configured:
//
//     #include "classes_structs.hpp"              // the reference harness
//     #include "uspmv_harness_adapter.hpp"
//     one_prec_kernel_func_ptr   = uspmv_b200::spmv_scs_launcher<VT, IT>;        // instead of spmv_gpu_scs_adv_launcher<VT, IT>
//     multi_prec_kernel_func_ptr = uspmv_b200::spmv_ap_scs_launcher<IT>;         // instead of spmv_gpu_ap_scs_launcher<IT>
//
// (the assignments in SpmvKernel's constructor, classes_structs.hpp:435-688).  Pointer kinds are detected per call: under nvcc the
// harness passes DEVICE arrays and DEVICE scalars C / n_chunks (utilities.hpp:3739-3811) and the raw-array kernels run directly; in a
// host build it passes host arrays, which run from a device twin adopted on first use.  Like the reference's launchers the call has
// completed when it returns (classes_structs.hpp:1032-1034 synchronises after every launch).
// Depends on the C ABI only; defines nothing the harness already defines.  tests: oracle/adapter_check.cpp compiles this header
// against the real classes_structs.hpp typedefs and runs the launchers through them.
#ifndef USPMV_HARNESS_ADAPTER_HPP
#define USPMV_HARNESS_ADAPTER_HPP

#include "uspmv_detail.hpp"

#ifdef __CUDACC__
#define USPMV_ADAPTER_NTB const ST n_thread_blocks,
#define USPMV_ADAPTER_NTB_UNUSED (void)n_thread_blocks;
#else
#define USPMV_ADAPTER_NTB
#define USPMV_ADAPTER_NTB_UNUSED
#endif

namespace uspmv_b200 {

#if defined(ROWWISE_BLOCK_VECTOR_LAYOUT)
constexpr int harness_block_vector_layout = USPMV_ROWWISE;  // Makefile:26-31: the layout is a compile-time switch of the harness
#else
constexpr int harness_block_vector_layout = USPMV_COLWISE;
#endif

namespace adapter_detail {
template <typename VT, typename IT>
void one_prec(const ST *C, const ST *n_chunks, const IT *chunk_ptrs, const IT *chunk_lengths, const IT *col_idxs, const VT *values, VT *x, VT *y,
              const int *block_vec_size, const int *vec_length) {
    using namespace uspmv_detail;
    const int bvs = block_vec_size ? *block_vec_size : 1;
    if (bvs <= 1) {
        execute_one_prec<VT, IT>(C, n_chunks, chunk_ptrs, chunk_lengths, col_idxs, values, x, y);
        return;
    }
    // block vectors (kernels.hpp:68-154,306-398): X[col + v * vec_length] (colwise) or X[col * bvs + v] (rowwise)
    const long c = scalar(C), nc = scalar(n_chunks);
    const long ld = vec_length ? *vec_length : nc * c;
    if (is_device(values)) {
        check(uspmv_block_spmv_gpu(default_ctx(), vt_of<VT>::value, c, nc, chunk_ptrs, chunk_lengths, col_idxs, values, x, y, bvs, ld,
                                   harness_block_vector_layout, nullptr));
        check(uspmv_ctx_sync(default_ctx()));
        return;
    }
    long x_rows = 0;
    std::shared_ptr<uspmv_scs> d = twin_of<VT, IT>(c, nc, chunk_ptrs, chunk_lengths, col_idxs, values, &x_rows);
    const bool rowwise = harness_block_vector_layout == USPMV_ROWWISE;
    const size_t n_x = rowwise ? (size_t)std::max(x_rows, ld) * bvs : (size_t)ld * bvs;
    const size_t n_y = rowwise ? (size_t)nc * c * bvs : (size_t)ld * bvs;
    dev_vec xd(n_x * sizeof(VT)), yd(n_y * sizeof(VT));
    check(uspmv_memcpy_h2d(default_ctx(), xd.p, x, n_x * sizeof(VT), nullptr));
    check(uspmv_spmmv(d.get(), xd.p, yd.p, bvs, ld, harness_block_vector_layout, nullptr));
    check(uspmv_memcpy_d2h(default_ctx(), y, yd.p, n_y * sizeof(VT), nullptr));
}
}  // namespace adapter_detail

// ---- OnePrecFuncPtr: SELL-C-sigma, CRS and the block-vector forms (one body: the kernel is chosen from *C and *block_vec_size) ----------
#define USPMV_ADAPTER_ONE_PREC(NAME)                                                                                                        \
    template <typename VT, typename IT>                                                                                                     \
    void NAME(bool warmup_flag, const ST *C, const ST *n_chunks, const IT *chunk_ptrs, const IT *chunk_lengths, const IT *col_idxs,         \
              const VT *values, VT *x, VT *y, int *block_vec_size, int *vec_length, USPMV_ADAPTER_NTB const int *my_rank) {               \
        (void)warmup_flag; (void)my_rank; USPMV_ADAPTER_NTB_UNUSED                                                                        \
        adapter_detail::one_prec<VT, IT>(C, n_chunks, chunk_ptrs, chunk_lengths, col_idxs, values, x, y, block_vec_size, vec_length);      \
    }
USPMV_ADAPTER_ONE_PREC(spmv_scs_launcher)        // replaces spmv_gpu_scs_launcher / spmv_gpu_scs_adv_launcher (kernels.hpp:610-628,757-775)
USPMV_ADAPTER_ONE_PREC(spmv_csr_launcher)        // replaces spmv_gpu_csr_launcher (kernels.hpp:661-680)
USPMV_ADAPTER_ONE_PREC(block_spmv_scs_launcher)  // replaces block_spmv_gpu_scs_launcher (a stub in the reference, kernels.hpp:811-844)
USPMV_ADAPTER_ONE_PREC(block_spmv_csr_launcher)  // replaces block_spmv_gpu_csr_launcher (a stub in the reference, kernels.hpp:777-809)
#undef USPMV_ADAPTER_ONE_PREC

// ---- MultiPrecFuncPtr: adaptive precision, 2-way (dp + sp) or — under HAVE_HALF_MATH — the bundle of three ------------------------------
// The harness stores the AP mode in its config, not in the call; the launcher takes it from which parts hold elements: a part whose
// n_chunks scalar is 0 / NULL is unused (dp+sp -> ap[dp_sp], dp+hp -> ap[dp_hp], sp+hp -> ap[sp_hp], all three -> ap[dp_sp_hp]), or
// from set_ap_mode() when the caller wants to be explicit.
inline int &forced_ap_mode() { static int m = -1; return m; }
inline void set_ap_mode(int uspmv_ap_mode) { forced_ap_mode() = uspmv_ap_mode; }

#ifdef HAVE_HALF_MATH
#if defined(__CUDACC__)
using harness_half_t = __half;
#else
using harness_half_t = _Float16;
#endif
#define USPMV_ADAPTER_HP_PARAMS                                                                                                      \
    const ST *hp_C, const ST *hp_n_chunks, const IT *hp_chunk_ptrs, const IT *hp_chunk_lengths, const IT *hp_col_idxs,               \
        const harness_half_t *hp_values, harness_half_t *hp_x, harness_half_t *hp_y,
#else
#define USPMV_ADAPTER_HP_PARAMS
#endif

#define USPMV_ADAPTER_MULTI_PREC(NAME)                                                                                                      \
    template <typename IT>                                                                                                                  \
    void NAME(bool warmup_flag, const ST *dp_C, const ST *dp_n_chunks, const IT *dp_chunk_ptrs, const IT *dp_chunk_lengths,                 \
              const IT *dp_col_idxs, const double *dp_values, double *dp_x, double *dp_y, const ST *sp_C, const ST *sp_n_chunks,            \
              const IT *sp_chunk_ptrs, const IT *sp_chunk_lengths, const IT *sp_col_idxs, const float *sp_values, float *sp_x, float *sp_y, \
              USPMV_ADAPTER_HP_PARAMS USPMV_ADAPTER_NTB const int *my_rank) {                                                             \
        (void)warmup_flag; (void)my_rank; USPMV_ADAPTER_NTB_UNUSED                                                                        \
        adapter_multi_body(dp_C, dp_n_chunks, dp_chunk_ptrs, dp_chunk_lengths, dp_col_idxs, dp_values, dp_x, dp_y, sp_C, sp_n_chunks,      \
                           sp_chunk_ptrs, sp_chunk_lengths, sp_col_idxs, sp_values, sp_x, sp_y USPMV_ADAPTER_HP_ARGS);                     \
    }

namespace adapter_detail {
inline bool part_used(const ST *n_chunks, const void *values) { return n_chunks && values && uspmv_detail::scalar(n_chunks) > 0; }
}  // namespace adapter_detail

#ifdef HAVE_HALF_MATH
#define USPMV_ADAPTER_HP_ARGS , hp_C, hp_n_chunks, hp_chunk_ptrs, hp_chunk_lengths, hp_col_idxs, hp_values, hp_x, hp_y
template <typename IT>
void adapter_multi_body(const ST *dp_C, const ST *dp_n_chunks, const IT *dp_cp, const IT *dp_cl, const IT *dp_ci, const double *dp_v, double *dp_x,
                        double *dp_y, const ST *sp_C, const ST *sp_n_chunks, const IT *sp_cp, const IT *sp_cl, const IT *sp_ci, const float *sp_v,
                        float *sp_x, float *sp_y, const ST *hp_C, const ST *hp_n_chunks, const IT *hp_cp, const IT *hp_cl, const IT *hp_ci,
                        const harness_half_t *hp_v, harness_half_t *hp_x, harness_half_t *hp_y) {
    (void)hp_C; (void)hp_x; (void)hp_y;
    using namespace adapter_detail;
    int mode = forced_ap_mode();
    if (mode < 0) {
        const bool d = part_used(dp_n_chunks, dp_v), s = part_used(sp_n_chunks, sp_v), h = part_used(hp_n_chunks, hp_v);
        mode = (d && s && h) ? USPMV_AP_DP_SP_HP : (d && s) ? USPMV_AP_DP_SP : (d && h) ? USPMV_AP_DP_HP : (s && h) ? USPMV_AP_SP_HP : -1;
        if (mode < 0) throw std::runtime_error("uspmv_b200 adaptive-precision launcher: fewer than two precision parts in use");
    }
    // uspmv_half_bits is layout-compatible 16-bit storage; the device kernels read IEEE binary16 either way
    uspmv_detail::execute_ap<IT, uspmv_detail::uspmv_half_bits>(mode, dp_C, dp_n_chunks, dp_cp, dp_cl, dp_ci, dp_v, dp_x, dp_y, sp_C, sp_n_chunks, sp_cp,
                                                               sp_cl, sp_ci, sp_v, sp_x, sp_y, hp_cp, hp_cl, hp_ci,
                                                               reinterpret_cast<const uspmv_detail::uspmv_half_bits *>(hp_v));
}
#else
#define USPMV_ADAPTER_HP_ARGS
template <typename IT>
void adapter_multi_body(const ST *dp_C, const ST *dp_n_chunks, const IT *dp_cp, const IT *dp_cl, const IT *dp_ci, const double *dp_v, double *dp_x,
                        double *dp_y, const ST *sp_C, const ST *sp_n_chunks, const IT *sp_cp, const IT *sp_cl, const IT *sp_ci, const float *sp_v,
                        float *sp_x, float *sp_y) {
    uspmv_detail::execute_ap<IT, uspmv_detail::uspmv_half_bits>(USPMV_AP_DP_SP, dp_C, dp_n_chunks, dp_cp, dp_cl, dp_ci, dp_v, dp_x, dp_y, sp_C,
                                                               sp_n_chunks, sp_cp, sp_cl, sp_ci, sp_v, sp_x, sp_y, nullptr, nullptr, nullptr, nullptr);
}
#endif

USPMV_ADAPTER_MULTI_PREC(spmv_ap_scs_launcher)  // replaces spmv_gpu_ap_scs_launcher / spmv_gpu_scs_ap_adv_launcher (ap_kernels.hpp:757-953)
USPMV_ADAPTER_MULTI_PREC(spmv_ap_csr_launcher)  // replaces spmv_gpu_ap_csr_launcher (ap_kernels.hpp:637-720)
#undef USPMV_ADAPTER_MULTI_PREC

}  // namespace uspmv_b200

#endif  // USPMV_HARNESS_ADAPTER_HPP
