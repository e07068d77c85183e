/* TEST INFRASTRUCTURE ONLY (oracle/).
 *
 * Thin extern "C" driver around the UNMODIFIED reference headers.  It is compiled by
 * oracle/Makefile with `-I/root/reference/code` (the reference sources are never
 * copied into this repository) into oracle/_ref/libuspmv_ref_{col,row}.so.  It is used
 *   (a) by tests/ to pin the C restatement in oracle/uspmv_oracle.c against the real
 *       reference (same libstdc++ std::sort, same loops), and to generate the golden
 *       fixtures in tests/golden/ (oracle/make_golden.py);
 *   (b) by bench.py's `cpu_baseline` leg and `--impl reference` arm, which time the
 *       reference's own OpenMP kernels on the GPU box's host cores.
 * Nothing in the product path (ultimate-spmv_b200/, include/) links or loads this.
 *
 * Reference entry points wrapped (all in /root/reference/code):
 *   convert_to_scs            utilities.hpp:1842-2104
 *   permute_scs_cols          utilities.hpp:1802-1831
 *   apply_permutation         utilities.hpp:1768-1782
 *   spmv_omp_csr              kernels.hpp:22-63
 *   spmv_omp_scs              kernels.hpp:159-211
 *   spmv_omp_scs_adv          kernels.hpp:265-301
 *   block_spmv_omp_csr        kernels.hpp:68-154
 *   block_spmv_omp_scs_general kernels.hpp:306-398   (layout fixed at compile time)
 *   spmv_omp_scs_ap_adv       ap_kernels.hpp:90-142  (dp+sp)
 *   spmv_omp_csr_apdpsp       ap_kernels.hpp:144-223
 *   partition_precisions      utilities.hpp:2810-3123 (2-way dp/sp only: the harness
 *                             version exits on any hp element, utilities.hpp:2941-2944)
 *   read_mtx                  utilities.hpp:2148-2309
 *   equilibrate_matrix        utilities.hpp:2668-2684
 *   seg_work_sharing_arr      mpi_funcs.hpp:424-622
 *   seg_mtx_struct            mpi_funcs.hpp:636-674
 *   localize_row_idx          mpi_funcs.hpp:862-877
 *   collect_local_needed_heri mpi_funcs.hpp:242-415
 *   init_std_vec_with_ptr_or_value / random_init  utilities.hpp:880-981
 */
#include "mmio.h"
#include "utilities.hpp"
#include "kernels.hpp"
#include "ap_kernels.hpp"
#include "mpi_funcs.hpp"

#include <cstring>
#include <vector>
#include <string>

#ifdef ROWWISE_BLOCK_VECTOR_LAYOUT
#define REF_LAYOUT 1
#else
#define REF_LAYOUT 0
#endif

namespace {

enum { VT_F64 = 0, VT_F32 = 1, VT_F16 = 2 };

struct RefScs {
    int vt;
    ScsData<double, int> d;
    ScsData<float, int> f;
#ifdef HAVE_HALF_MATH
    ScsData<_Float16, int> h;
#endif
};

template <typename VT>
void fill_mtx(MtxData<VT, int> &m, long n_rows, long n_cols, long nnz, const int *I, const int *J,
              const double *vals) {
    m.n_rows = n_rows;
    m.n_cols = n_cols;
    m.nnz = nnz;
    m.is_sorted = true;
    m.is_symmetric = false;
    m.I.assign(I, I + nnz);
    m.J.assign(J, J + nnz);
    m.values.resize(nnz);
    for (long i = 0; i < nnz; ++i) m.values[i] = static_cast<VT>(vals[i]);  // MtxData::copy semantics
}

template <typename VT>
void build(ScsData<VT, int> &s, long n_rows, long n_cols, long nnz, const int *I, const int *J,
           const double *vals, long C, long sigma, const int *fixed_perm) {
    MtxData<VT, int> m;
    fill_mtx(m, n_rows, n_cols, nnz, I, J, vals);
    convert_to_scs<VT, VT, int>(&m, C, sigma, &s, const_cast<int *>(fixed_perm));
}

template <typename VT>
void copy_out(ScsData<VT, int> &s, int *chunk_ptrs, int *chunk_lengths, int *col_idxs, void *values,
              int *old_to_new, int *new_to_old) {
    if (chunk_ptrs) std::memcpy(chunk_ptrs, s.chunk_ptrs.data(), sizeof(int) * (s.n_chunks + 1));
    if (chunk_lengths) std::memcpy(chunk_lengths, s.chunk_lengths.data(), sizeof(int) * s.n_chunks);
    if (col_idxs) std::memcpy(col_idxs, s.col_idxs.data(), sizeof(int) * s.n_elements);
    if (values) std::memcpy(values, s.values.data(), sizeof(VT) * s.n_elements);
    if (old_to_new) std::memcpy(old_to_new, s.old_to_new_idx.data(), sizeof(int) * s.n_rows);
    // new_to_old_idx is an uninitialised `new int[n_rows + sigma]`; only entries that are the image
    // of a real row are defined (utilities.hpp:2060-2069).  We copy n_rows entries verbatim.
    if (new_to_old) std::memcpy(new_to_old, s.new_to_old_idx, sizeof(int) * s.n_rows);
}

template <typename VT>
void dims(ScsData<VT, int> &s, long *o) {
    o[0] = s.C; o[1] = s.sigma; o[2] = s.n_rows; o[3] = s.n_cols;
    o[4] = s.n_rows_padded; o[5] = s.n_chunks; o[6] = s.n_elements; o[7] = s.nnz;
}

}  // namespace

extern "C" {

int ref_layout(void) { return REF_LAYOUT; }
int ref_have_half(void) {
#ifdef HAVE_HALF_MATH
    return 1;
#else
    return 0;
#endif
}

void *ref_scs_build(int vt, long n_rows, long n_cols, long nnz, const int *I, const int *J,
                    const double *vals, long C, long sigma, const int *fixed_perm) {
    RefScs *r = new RefScs();
    r->vt = vt;
    if (vt == VT_F64) build(r->d, n_rows, n_cols, nnz, I, J, vals, C, sigma, fixed_perm);
    else if (vt == VT_F32) build(r->f, n_rows, n_cols, nnz, I, J, vals, C, sigma, fixed_perm);
#ifdef HAVE_HALF_MATH
    else if (vt == VT_F16) build(r->h, n_rows, n_cols, nnz, I, J, vals, C, sigma, fixed_perm);
#endif
    else { delete r; return nullptr; }
    return r;
}

void ref_scs_dims(void *h, long *out8) {
    RefScs *r = static_cast<RefScs *>(h);
    if (r->vt == VT_F64) dims(r->d, out8);
    else if (r->vt == VT_F32) dims(r->f, out8);
#ifdef HAVE_HALF_MATH
    else dims(r->h, out8);
#endif
}

void ref_scs_copy(void *h, int *chunk_ptrs, int *chunk_lengths, int *col_idxs, void *values,
                  int *old_to_new, int *new_to_old) {
    RefScs *r = static_cast<RefScs *>(h);
    if (r->vt == VT_F64) copy_out(r->d, chunk_ptrs, chunk_lengths, col_idxs, values, old_to_new, new_to_old);
    else if (r->vt == VT_F32) copy_out(r->f, chunk_ptrs, chunk_lengths, col_idxs, values, old_to_new, new_to_old);
#ifdef HAVE_HALF_MATH
    else copy_out(r->h, chunk_ptrs, chunk_lengths, col_idxs, values, old_to_new, new_to_old);
#endif
}

void ref_scs_permute_cols(void *h, const int *perm) {
    RefScs *r = static_cast<RefScs *>(h);
    if (r->vt == VT_F64) permute_scs_cols(&r->d, const_cast<int *>(perm));
    else if (r->vt == VT_F32) permute_scs_cols(&r->f, const_cast<int *>(perm));
#ifdef HAVE_HALF_MATH
    else permute_scs_cols(&r->h, const_cast<int *>(perm));
#endif
}

/* collect_local_needed_heri on the handle's col_idxs (in place).  recv_idxs_flat receives the P need
 * lists back to back, recv_idxs_ptr[P+1] their offsets, recv_counts_cumsum[P+1] as the reference fills it. */
long ref_scs_collect_halo(void *h, const int *work_sharing_arr, int my_rank, int comm_size,
                          int *recv_idxs_flat, long recv_cap, int *recv_idxs_ptr, int *recv_counts_cumsum) {
    RefScs *r = static_cast<RefScs *>(h);
    std::vector<std::vector<int>> recv_idxs(comm_size);
    std::vector<int> cumsum(comm_size + 2, 0);
    if (r->vt == VT_F64)
        collect_local_needed_heri<double, int>("dp", &recv_idxs, &cumsum, &r->d, work_sharing_arr, my_rank, comm_size);
    else if (r->vt == VT_F32)
        collect_local_needed_heri<float, int>("sp", &recv_idxs, &cumsum, &r->f, work_sharing_arr, my_rank, comm_size);
#ifdef HAVE_HALF_MATH
    else
        collect_local_needed_heri<_Float16, int>("hp", &recv_idxs, &cumsum, &r->h, work_sharing_arr, my_rank, comm_size);
#endif
    long tot = 0;
    for (int p = 0; p < comm_size; ++p) {
        recv_idxs_ptr[p] = (int)tot;
        for (int v : recv_idxs[p]) {
            if (tot < recv_cap) recv_idxs_flat[tot] = v;
            ++tot;
        }
    }
    recv_idxs_ptr[comm_size] = (int)tot;
    for (int p = 0; p <= comm_size; ++p) recv_counts_cumsum[p] = cumsum[p];
    return tot;
}

void ref_scs_free(void *h) { delete static_cast<RefScs *>(h); }

/* ---- kernels on raw arrays ------------------------------------------------------------------- */

void ref_spmv_scs(int vt, int adv, long C, long n_chunks, const int *cp, const int *cl, const int *ci,
                  const void *vals, void *x, void *y) {
    int bvs = 1, vl = 0, rank = 0;
    if (vt == VT_F64) {
        if (adv) spmv_omp_scs_adv<double, int>(false, &C, &n_chunks, cp, cl, ci, (const double *)vals, (double *)x, (double *)y, &bvs, &vl, &rank);
        else spmv_omp_scs<double, int>(false, &C, &n_chunks, cp, cl, ci, (const double *)vals, (double *)x, (double *)y, &bvs, &vl, &rank);
    } else if (vt == VT_F32) {
        if (adv) spmv_omp_scs_adv<float, int>(false, &C, &n_chunks, cp, cl, ci, (const float *)vals, (float *)x, (float *)y, &bvs, &vl, &rank);
        else spmv_omp_scs<float, int>(false, &C, &n_chunks, cp, cl, ci, (const float *)vals, (float *)x, (float *)y, &bvs, &vl, &rank);
    }
#ifdef HAVE_HALF_MATH
    else {
        if (adv) spmv_omp_scs_adv<_Float16, int>(false, &C, &n_chunks, cp, cl, ci, (const _Float16 *)vals, (_Float16 *)x, (_Float16 *)y, &bvs, &vl, &rank);
        else spmv_omp_scs<_Float16, int>(false, &C, &n_chunks, cp, cl, ci, (const _Float16 *)vals, (_Float16 *)x, (_Float16 *)y, &bvs, &vl, &rank);
    }
#endif
}

void ref_spmv_csr(int vt, long n_rows, const int *rp, const int *ci, const void *vals, void *x, void *y) {
    long C = 1;
    int bvs = 1, vl = 0, rank = 0;
    if (vt == VT_F64) spmv_omp_csr<double, int>(false, &C, &n_rows, rp, nullptr, ci, (const double *)vals, (double *)x, (double *)y, &bvs, &vl, &rank);
    else if (vt == VT_F32) spmv_omp_csr<float, int>(false, &C, &n_rows, rp, nullptr, ci, (const float *)vals, (float *)x, (float *)y, &bvs, &vl, &rank);
#ifdef HAVE_HALF_MATH
    else spmv_omp_csr<_Float16, int>(false, &C, &n_rows, rp, nullptr, ci, (const _Float16 *)vals, (_Float16 *)x, (_Float16 *)y, &bvs, &vl, &rank);
#endif
}

/* SpMMV; the block-vector layout is the one this TU was compiled for (ref_layout()). */
void ref_spmmv_scs(int vt, long C, long n_chunks, const int *cp, const int *cl, const int *ci, const void *vals,
                   void *X, void *Y, int bvs, int vec_length) {
    int rank = 0;
    if (vt == VT_F64) block_spmv_omp_scs_general<double, int>(false, &C, &n_chunks, cp, cl, ci, (const double *)vals, (double *)X, (double *)Y, &bvs, &vec_length, &rank);
    else if (vt == VT_F32) block_spmv_omp_scs_general<float, int>(false, &C, &n_chunks, cp, cl, ci, (const float *)vals, (float *)X, (float *)Y, &bvs, &vec_length, &rank);
#ifdef HAVE_HALF_MATH
    else block_spmv_omp_scs_general<_Float16, int>(false, &C, &n_chunks, cp, cl, ci, (const _Float16 *)vals, (_Float16 *)X, (_Float16 *)Y, &bvs, &vec_length, &rank);
#endif
}

void ref_spmmv_csr(int vt, long n_rows, const int *rp, const int *ci, const void *vals, void *X, void *Y,
                   int bvs, int vec_length) {
    long C = 1;
    int rank = 0;
    if (vt == VT_F64) block_spmv_omp_csr<double, int>(false, &C, &n_rows, rp, nullptr, ci, (const double *)vals, (double *)X, (double *)Y, &bvs, &vec_length, &rank);
    else if (vt == VT_F32) block_spmv_omp_csr<float, int>(false, &C, &n_rows, rp, nullptr, ci, (const float *)vals, (float *)X, (float *)Y, &bvs, &vec_length, &rank);
#ifdef HAVE_HALF_MATH
    else block_spmv_omp_csr<_Float16, int>(false, &C, &n_rows, rp, nullptr, ci, (const _Float16 *)vals, (_Float16 *)X, (_Float16 *)Y, &bvs, &vec_length, &rank);
#endif
}

/* dp+sp adaptive precision, SCS (templated-C kernel scs_ap_impl_cpu) */
void ref_ap_scs_dpsp(long C, long n_chunks, const int *dcp, const int *dcl, const int *dci, const double *dv,
                     const int *scp, const int *scl, const int *sci, const float *sv, double *dp_x, float *sp_x,
                     double *dp_y) {
    int rank = 0;
    std::vector<float> sp_y(n_chunks * C);
#ifdef HAVE_HALF_MATH
    spmv_omp_scs_ap_adv<int>(false, &C, &n_chunks, dcp, dcl, dci, dv, dp_x, dp_y, &C, &n_chunks, scp, scl, sci, sv, sp_x, sp_y.data(),
                             &C, &n_chunks, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, &rank);
#else
    spmv_omp_scs_ap_adv<int>(false, &C, &n_chunks, dcp, dcl, dci, dv, dp_x, dp_y, &C, &n_chunks, scp, scl, sci, sv, sp_x, sp_y.data(), &rank);
#endif
}

/* 2-way dp/sp split through the harness' partition_precisions.  Returns dp count; arrays sized nnz. */
long ref_partition_dpsp(long n_rows, long n_cols, long nnz, const int *I, const int *J, const double *vals, double t1,
                        int *dI, int *dJ, double *dV, int *sI, int *sJ, float *sV, long *n_sp) {
    Config cfg;
    cfg.value_type = "ap[dp_sp]";
    cfg.ap_threshold_1 = t1;
    cfg.equilibrate = 0;
    MtxData<double, int> m, dp;
    MtxData<float, int> sp;
    fill_mtx(m, n_rows, n_cols, nnz, I, J, vals);
    std::vector<double> rowmax, colmax;
#ifdef HAVE_HALF_MATH
    MtxData<_Float16, int> hp;
    partition_precisions<double, int>(&cfg, &m, &dp, &sp, &hp, &rowmax, &colmax, 0);
#else
    partition_precisions<double, int>(&cfg, &m, &dp, &sp, &rowmax, &colmax, 0);
#endif
    std::memcpy(dI, dp.I.data(), sizeof(int) * dp.nnz);
    std::memcpy(dJ, dp.J.data(), sizeof(int) * dp.nnz);
    std::memcpy(dV, dp.values.data(), sizeof(double) * dp.nnz);
    std::memcpy(sI, sp.I.data(), sizeof(int) * sp.nnz);
    std::memcpy(sJ, sp.J.data(), sizeof(int) * sp.nnz);
    std::memcpy(sV, sp.values.data(), sizeof(float) * sp.nnz);
    *n_sp = sp.nnz;
    return dp.nnz;
}

/* read_mtx: returns nnz (after symmetric expansion + stable row sort); call twice (first with NULLs). */
long ref_read_mtx(const char *path, long *n_rows, long *n_cols, int *I, int *J, double *vals) {
    static MtxData<double, int> cache;
    static std::string cached_path;
    if (cached_path != path) {
        Config cfg;
        cfg.matrix_file_name = path;
        cache = MtxData<double, int>();
        read_mtx(cfg, &cache, 0);
        cached_path = path;
    }
    *n_rows = cache.n_rows;
    *n_cols = cache.n_cols;
    if (I) std::memcpy(I, cache.I.data(), sizeof(int) * cache.nnz);
    if (J) std::memcpy(J, cache.J.data(), sizeof(int) * cache.nnz);
    if (vals) std::memcpy(vals, cache.values.data(), sizeof(double) * cache.nnz);
    return cache.nnz;
}

/* equilibrate_matrix (utilities.hpp:2668-2684) on a square dp COO, values scaled in place */
void ref_equilibrate(long n_rows, long n_cols, long nnz, const int *I, const int *J, double *vals) {
    MtxData<double, int> m;
    m.n_rows = n_rows; m.n_cols = n_cols; m.nnz = nnz;
    m.I.assign(I, I + nnz);
    m.J.assign(J, J + nnz);
    m.values.assign(vals, vals + nnz);
    equilibrate_matrix<double, int>(&m);
    std::memcpy(vals, m.values.data(), sizeof(double) * nnz);
}

/* seg_method: 0 = seg-rows, 1 = seg-nnz.  wsa has P+1 entries. */
void ref_seg_work_sharing_arr(int seg_method, long n_rows, long nnz, const int *I, int P, int *wsa) {
    Config cfg;
    cfg.seg_method = seg_method ? "seg-nnz" : "seg-rows";
    MtxData<double, int> m;
    m.n_rows = n_rows;
    m.n_cols = n_rows;
    m.nnz = nnz;
    m.I.assign(I, I + nnz);
    seg_work_sharing_arr<double, int>(&cfg, &m, wsa, P, 0);
}

/* seg_mtx_struct + localize_row_idx for one rank; returns local nnz; outputs sized >= nnz. */
long ref_seg_mtx(long n_rows, long nnz, const int *I, const int *J, const double *vals, const int *wsa, int rank,
                 int *lI, int *lJ, double *lV) {
    MtxData<double, int> m;
    fill_mtx(m, n_rows, n_rows, nnz, I, J, vals);
    std::vector<int> li, lj;
    std::vector<double> lv;
    seg_mtx_struct<double, int>(&m, &li, &lj, &lv, wsa, rank);
    MtxData<double, int> loc;
    loc.nnz = (long)li.size();
    loc.I = li;
    loc.J = lj;
    loc.values = lv;
    localize_row_idx<double, int>(&loc);
    std::memcpy(lI, loc.I.data(), sizeof(int) * loc.nnz);
    std::memcpy(lJ, loc.J.data(), sizeof(int) * loc.nnz);
    std::memcpy(lV, loc.values.data(), sizeof(double) * loc.nnz);
    return loc.nnz;
}

/* init_std_vec_with_ptr_or_value with random numbers (utilities.hpp:914-981 -> random_init :880-912): default-seeded mt19937,
 * uniform_real_distribution<double>(matrix_min, matrix_max) over all n_x values, then the padding rule of this TU's block-vector
 * layout.  vt: 0 double, 1 float. */
void ref_random_x(int vt, double vmin, double vmax, long n_x, int n_rows, int n_rows_padded, int bvs, void *out) {
    Config cfg;
    cfg.matrix_min = vmin;
    cfg.matrix_max = vmax;
    cfg.block_vec_size = bvs;
    if (vt == VT_F64) {
        std::vector<double> x(n_x);
        init_std_vec_with_ptr_or_value<double>(&cfg, x, n_x, n_rows, n_rows_padded, 5.0, '1');
        std::memcpy(out, x.data(), sizeof(double) * n_x);
    } else {
        std::vector<float> x(n_x);
        init_std_vec_with_ptr_or_value<float>(&cfg, x, n_x, n_rows, n_rows_padded, 5.0f, '1');
        std::memcpy(out, x.data(), sizeof(float) * n_x);
    }
}

int ref_omp_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

}  // extern "C"
