// uspmv_detail.hpp — plumbing shared by the two C++ faces of libuspmv_b200.so:
//   include/uspmv_interface.hpp        the reference's LIBRARY header re-created (code/interface.hpp: MtxData, ScsData, convert_to_scs, ...)
//   include/uspmv_harness_adapter.hpp  launchers with the reference HARNESS' std::function signatures (classes_structs.hpp:283-333)
// It depends on the C ABI only and defines nothing outside namespace uspmv_detail (the harness has its own MtxData / ScsData, so the
// adapter must not drag the library shim's names in).  No CUDA headers needed.
#ifndef USPMV_DETAIL_HPP
#define USPMV_DETAIL_HPP

#include <algorithm>
#include <map>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>

#include "uspmv_b200.h"

using ST = long;  // mmio.h:21, classes_structs.hpp:31

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

// Device twins of HOST matrix arrays, keyed by the address of the host `values` array.  convert_to_scs / permute_scs_cols register
// the ScsData they fill; a host array nobody registered (a ScsData built by other code, e.g. the reference's own convert_to_scs) is
// adopted on first use (uspmv_scs_from_arrays).  This is what lets execute_uspmv keep the reference's pointer-bundle signature
// (interface.hpp:1871-1910) while the matrix lives on the GPU.
struct twin {
    std::weak_ptr<uspmv_scs> dev;       // registered by convert_to_scs (owned by the ScsData)
    std::shared_ptr<uspmv_scs> adopted; // adopted arrays are owned by the registry
    long n_chunks = 0, n_elements = 0, x_len = 0;
};
inline std::map<const void *, twin> &twins() { static std::map<const void *, twin> m; return m; }
inline std::mutex &twins_mutex() { static std::mutex m; return m; }
inline bool is_device(const void *p) {
    int d = 0;
    check(uspmv_pointer_is_device(p, &d));
    return d != 0;
}
inline long scalar(const ST *p) {  // C / n_chunks: host scalars in the library, DEVICE scalars in the nvcc harness (utilities.hpp:3739-3749)
    if (!is_device(p)) return *p;
    static std::map<const ST *, long> cache;
    static std::mutex m;
    std::lock_guard<std::mutex> g(m);
    auto it = cache.find(p);
    if (it != cache.end()) return it->second;
    ST v = 0;
    check(uspmv_memcpy_d2h(default_ctx(), &v, p, sizeof(ST), nullptr));
    cache[p] = v;
    return v;
}

// device twin of a host matrix given as raw arrays: registered, or adopted now
template <typename VT, typename IT>
std::shared_ptr<uspmv_scs> twin_of(long C, long n_chunks, const IT *chunk_ptrs, const IT *chunk_lengths, const IT *col_idxs, const VT *values,
                                   long *x_len) {
    std::lock_guard<std::mutex> g(twins_mutex());
    auto it = twins().find(values);
    const long n_el = n_chunks > 0 ? chunk_ptrs[n_chunks] : 0;
    if (it != twins().end() && it->second.n_chunks == n_chunks && it->second.n_elements == n_el) {
        std::shared_ptr<uspmv_scs> d = it->second.adopted ? it->second.adopted : it->second.dev.lock();
        if (d) { *x_len = it->second.x_len; return d; }
    }
    long mx = -1;
    for (long e = 0; e < n_el; ++e) mx = col_idxs[e] > mx ? col_idxs[e] : mx;
    uspmv_scs *raw = nullptr;
    check(uspmv_scs_from_arrays(default_ctx(), vt_of<VT>::value, C, 1, n_chunks * C, mx + 1, n_chunks, chunk_ptrs, chunk_lengths, col_idxs, values,
                                nullptr, 0, &raw));
    twin &t = twins()[values];
    t.adopted = std::shared_ptr<uspmv_scs>(raw, scs_deleter());
    t.dev.reset(); t.n_chunks = n_chunks; t.n_elements = n_el; t.x_len = mx + 1;
    *x_len = t.x_len;
    return t.adopted;
}

inline int ap_mode_of(const char *ap_value_type) {
    const std::string t(ap_value_type ? ap_value_type : "");
    return t == "ap[dp_sp]" ? USPMV_AP_DP_SP : t == "ap[dp_hp]" ? USPMV_AP_DP_HP : t == "ap[sp_hp]" ? USPMV_AP_SP_HP
           : t == "ap[dp_sp_hp]" ? USPMV_AP_DP_SP_HP : -1;
}
template <typename VT, typename IT>
void execute_one_prec(const ST *C, const ST *n_chunks, const IT *chunk_ptrs, const IT *chunk_lengths, const IT *col_idxs, const VT *values, VT *x,
                      VT *y) {
    const long c = scalar(C), nc = scalar(n_chunks);
    if (is_device(values)) {
        if (c > 1) check(::uspmv_scs_gpu(default_ctx(), vt_of<VT>::value, c, nc, chunk_ptrs, chunk_lengths, col_idxs, values, x, y, nullptr));
        else check(::uspmv_csr_gpu(default_ctx(), vt_of<VT>::value, nc, chunk_ptrs, col_idxs, values, x, y, nullptr));
        check(uspmv_ctx_sync(default_ctx()));  // the reference's kernels have completed when the call returns
        return;
    }
    long x_len = 0;
    std::shared_ptr<uspmv_scs> d = twin_of<VT, IT>(c, nc, chunk_ptrs, chunk_lengths, col_idxs, values, &x_len);
    check(uspmv_spmv_host(d.get(), x, x_len, y, nc * c));
}
struct dev_vec {  // scratch device vector for the host-array AP path
    void *p = nullptr;
    explicit dev_vec(size_t bytes) { check(uspmv_malloc(default_ctx(), bytes, &p)); }
    ~dev_vec() { uspmv_free(default_ctx(), p); }
};
template <typename IT, typename HT>
void execute_ap(int mode, const ST *dp_C, const ST *dp_n_chunks, const IT *dp_cp, const IT *dp_cl, const IT *dp_ci, const double *dp_v, double *dp_x,
                double *dp_y, const ST *sp_C, const ST *sp_n_chunks, const IT *sp_cp, const IT *sp_cl, const IT *sp_ci, const float *sp_v, float *sp_x,
                float *sp_y, const IT *hp_cp, const IT *hp_cl, const IT *hp_ci, const HT *hp_v) {
    const bool sphp = mode == USPMV_AP_SP_HP;
    const long c = scalar(sphp ? sp_C : dp_C), nc = scalar(sphp ? sp_n_chunks : dp_n_chunks);
    const bool use[3] = {mode != USPMV_AP_SP_HP, mode != USPMV_AP_DP_HP, mode != USPMV_AP_DP_SP};
    const void *first_vals = sphp ? static_cast<const void *>(sp_v) : static_cast<const void *>(dp_v);
    void *x = sphp ? static_cast<void *>(sp_x) : static_cast<void *>(dp_x);
    void *y = sphp ? static_cast<void *>(sp_y) : static_cast<void *>(dp_y);
    const size_t xs = sphp ? 4 : 8;
    if (is_device(first_vals)) {
        const void *arr[12] = {dp_cp, dp_cl, dp_ci, dp_v, sp_cp, sp_cl, sp_ci, sp_v, hp_cp, hp_cl, hp_ci, hp_v};
        check(uspmv_scs_ap_gpu(default_ctx(), mode, c, nc, arr, x, y, nullptr));
        check(uspmv_ctx_sync(default_ctx()));
        return;
    }
    long xl[3] = {0, 0, 0};
    std::shared_ptr<uspmv_scs> d[3];
    if (use[0]) d[0] = twin_of<double, IT>(c, nc, dp_cp, dp_cl, dp_ci, dp_v, &xl[0]);
    if (use[1]) d[1] = twin_of<float, IT>(c, nc, sp_cp, sp_cl, sp_ci, sp_v, &xl[1]);
    if (use[2]) d[2] = twin_of<HT, IT>(c, nc, hp_cp, hp_cl, hp_ci, hp_v, &xl[2]);
    const long x_len = std::max(xl[0], std::max(xl[1], xl[2]));
    dev_vec xd(x_len * xs), yd(nc * c * xs);
    check(uspmv_memcpy_h2d(default_ctx(), xd.p, x, x_len * xs, nullptr));
    check(uspmv_ap_spmv(mode, d[0].get(), d[1].get(), d[2].get(), xd.p, yd.p, nullptr));
    check(uspmv_memcpy_d2h(default_ctx(), y, yd.p, nc * c * xs, nullptr));
}
}  // namespace uspmv_detail

#endif  // USPMV_DETAIL_HPP
