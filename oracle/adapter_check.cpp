/* TEST INFRASTRUCTURE ONLY (oracle/).
 *
 * Compiles include/uspmv_harness_adapter.hpp against the REAL typedefs of the reference harness — SpmvKernel<VT,IT>::OnePrecFuncPtr /
 * MultiPrecFuncPtr, code/classes_structs.hpp:283-333, reached through the unmodified utilities.hpp — and runs the launchers THROUGH
 * those std::function objects, exactly where the harness would (SpmvKernel::execute_one_prec / execute_two_prec,
 * classes_structs.hpp:997-1115), against the reference's own host kernels:
 *     spmv_omp_scs            kernels.hpp:159-211      (bit-equal y, host arrays and device arrays)
 *     spmv_omp_csr            kernels.hpp:22-63        (omp simd partial sums: within 1e-12 of sum|a||x|)
 *     block_spmv_omp_scs_general  kernels.hpp:306-398  (block vectors in the layout this TU is compiled for)
 *     spmv_omp_scs_ap_adv     ap_kernels.hpp:90-142    (dp + sp; the harness kernel multiplies the sp part with the FLOAT copy of x,
 *                                                       the library kernel with the double one: compared within 1e-6 of sum|a||x|)
 * Built by oracle/Makefile into oracle/_ref/adapter_check_{col,row} (needs /root/reference; the binaries travel to the GPU box).
 * Exit code 0 = every check passed; prints one line per check.  `adapter_check --compile-only` returns 0 without touching the GPU
 * (the assignability of the launchers to the typedefs is a COMPILE-time fact). */
#include "mmio.h"
#include "utilities.hpp"
#include "kernels.hpp"
#include "ap_kernels.hpp"

#include "uspmv_harness_adapter.hpp"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <random>
#include <vector>

using K = SpmvKernel<double, int>;
using KF = SpmvKernel<float, int>;

// the compile-time part: every launcher converts to the harness' function-pointer typedefs
static K::OnePrecFuncPtr f_scs = uspmv_b200::spmv_scs_launcher<double, int>;
static K::OnePrecFuncPtr f_csr = uspmv_b200::spmv_csr_launcher<double, int>;
static K::OnePrecFuncPtr f_bscs = uspmv_b200::block_spmv_scs_launcher<double, int>;
static K::OnePrecFuncPtr f_bcsr = uspmv_b200::block_spmv_csr_launcher<double, int>;
static KF::OnePrecFuncPtr f_scs_sp = uspmv_b200::spmv_scs_launcher<float, int>;
static K::MultiPrecFuncPtr g_ap_scs = uspmv_b200::spmv_ap_scs_launcher<int>;
static K::MultiPrecFuncPtr g_ap_csr = uspmv_b200::spmv_ap_csr_launcher<int>;

static int fails = 0;
static void report(const char *what, bool ok, double err) {
    std::printf("adapter_check: %-58s %s (max err %.3e)\n", what, ok ? "OK" : "FAILED", err);
    if (!ok) ++fails;
}

template <typename T>
static T *to_device(const std::vector<T> &v) {
    void *d = nullptr;
    uspmv_detail::check(uspmv_malloc(uspmv_detail::default_ctx(), (v.size() + 1) * sizeof(T), &d));
    uspmv_detail::check(uspmv_memcpy_h2d(uspmv_detail::default_ctx(), d, v.data(), v.size() * sizeof(T), nullptr));
    return static_cast<T *>(d);
}

int main(int argc, char **argv) {
    if (argc > 1 && !std::strcmp(argv[1], "--compile-only")) {
        std::printf("adapter_check: launchers are assignable to OnePrecFuncPtr / MultiPrecFuncPtr (%s block vectors%s)\n",
#ifdef ROWWISE_BLOCK_VECTOR_LAYOUT
                    "rowwise",
#else
                    "colwise",
#endif
#ifdef HAVE_HALF_MATH
                    ", HAVE_HALF_MATH"
#else
                    ""
#endif
        );
        return (f_scs && f_csr && f_bscs && f_bcsr && f_scs_sp && g_ap_scs && g_ap_csr) ? 0 : 1;
    }
    // ---- a matrix with rows of very different lengths, through the REFERENCE's own convert_to_scs --------------------------------
    const int n = 3000;
    std::mt19937 rng(7);
    MtxData<double, int> m;
    m.n_rows = m.n_cols = n; m.is_sorted = true; m.is_symmetric = false;
    for (int i = 0; i < n; ++i) {
        const int cnt = 1 + (int)(rng() % 9) + ((i % 97) == 0 ? 60 : 0);
        for (int k = 0; k < cnt; ++k) {
            m.I.push_back(i); m.J.push_back((int)(rng() % n));
            const double mag = std::pow(10.0, -3.0 + 4.0 * (double)(rng() % 1000) / 1000.0);
            m.values.push_back((rng() & 1) ? mag : -mag);
        }
    }
    m.nnz = (long)m.values.size();
    ST C = 32, sigma = 64;
    ScsData<double, int> scs;
    convert_to_scs<double, double, int>(&m, C, sigma, &scs);
    permute_scs_cols(&scs, &scs.old_to_new_idx[0]);
    ST n_chunks = scs.n_chunks;
    const long n_pad = scs.n_rows_padded;
    std::vector<double> x(n_pad), y_ref(n_pad, 0.0), y(n_pad, -1.0);
    for (long i = 0; i < n_pad; ++i) x[i] = std::sin(0.37 * i) + 1.5;
    int bvs = 1, vec_length = (int)n_pad, rank = 0;
    spmv_omp_scs<double, int>(false, &C, &n_chunks, scs.chunk_ptrs.data(), scs.chunk_lengths.data(), scs.col_idxs.data(), scs.values.data(), x.data(),
                              y_ref.data(), &bvs, &vec_length, &rank);
    auto max_diff = [&](const std::vector<double> &a, const std::vector<double> &b) {
        double d = 0.0;
        for (size_t i = 0; i < a.size(); ++i) d = std::fmax(d, std::fabs(a[i] - b[i]));
        return d;
    };
    // (1) host arrays through OnePrecFuncPtr (a host build of the harness)
    f_scs(false, &C, &n_chunks, scs.chunk_ptrs.data(), scs.chunk_lengths.data(), scs.col_idxs.data(), scs.values.data(), x.data(), y.data(), &bvs,
          &vec_length, &rank);
    report("OnePrecFuncPtr  SELL-32-64 dp, host arrays", std::memcmp(y.data(), y_ref.data(), n_pad * 8) == 0, max_diff(y, y_ref));
    // (2) device arrays + device scalars (what the nvcc harness passes, utilities.hpp:3739-3811)
    {
        std::vector<ST> sc{C}, snc{n_chunks};
        ST *C_d = to_device(sc), *nc_d = to_device(snc);
        std::vector<int> cp(scs.chunk_ptrs.begin(), scs.chunk_ptrs.begin() + n_chunks + 1), cl(scs.chunk_lengths.begin(), scs.chunk_lengths.begin() + n_chunks);
        std::vector<int> ci(scs.col_idxs.begin(), scs.col_idxs.begin() + scs.n_elements);
        std::vector<double> v(scs.values.begin(), scs.values.begin() + scs.n_elements);
        int *cp_d = to_device(cp), *cl_d = to_device(cl), *ci_d = to_device(ci);
        double *v_d = to_device(v), *x_d = to_device(x), *y_d = to_device(y);
        std::fill(y.begin(), y.end(), -1.0);
        f_scs(false, C_d, nc_d, cp_d, cl_d, ci_d, v_d, x_d, y_d, &bvs, &vec_length, &rank);
        uspmv_detail::check(uspmv_memcpy_d2h(uspmv_detail::default_ctx(), y.data(), y_d, n_pad * 8, nullptr));
        report("OnePrecFuncPtr  SELL-32-64 dp, device arrays + device scalars", std::memcmp(y.data(), y_ref.data(), n_pad * 8) == 0, max_diff(y, y_ref));
    }
    // (3) CRS
    {
        ScsData<double, int> crs;
        ST one = 1;
        convert_to_scs<double, double, int>(&m, 1, 1, &crs);
        ST nr = crs.n_chunks;
        std::vector<double> yr(n, 0.0), yc(n, -1.0);
        spmv_omp_csr<double, int>(false, &one, &nr, crs.chunk_ptrs.data(), crs.chunk_lengths.data(), crs.col_idxs.data(), crs.values.data(), x.data(), yr.data(),
                                  &bvs, &vec_length, &rank);
        f_csr(false, &one, &nr, crs.chunk_ptrs.data(), crs.chunk_lengths.data(), crs.col_idxs.data(), crs.values.data(), x.data(), yc.data(), &bvs, &vec_length,
              &rank);
        // the reference's CRS loop carries `#pragma omp simd simdlen(SIMD_LENGTH)` (kernels.hpp:49): its row sums are 4 interleaved
        // partial sums, not the sequential sum -> compared within 1e-12 of sum |a||x| (the north-star tolerance), not bit for bit
        std::vector<double> sc(n, 0.0);
        for (long k = 0; k < m.nnz; ++k) sc[m.I[k]] += std::fabs(m.values[k] * x[m.J[k]]);
        bool ok = true;
        for (int i = 0; i < n; ++i) ok = ok && std::fabs(yc[i] - yr[i]) <= 1e-12 * std::fmax(sc[i], 1e-300);
        report("OnePrecFuncPtr  CRS dp, host arrays (1e-12 of sum|a||x|)", ok, max_diff(yc, yr));
    }
    // (4) block vectors, in the layout this TU is compiled for
    {
        int b = 4, ld = (int)n_pad;
        std::vector<double> X((size_t)ld * b), Yr((size_t)ld * b, 0.0), Y((size_t)ld * b, -1.0);
        for (size_t i = 0; i < X.size(); ++i) X[i] = std::sin(0.11 * (double)i) + 0.25;
        block_spmv_omp_scs_general<double, int>(false, &C, &n_chunks, scs.chunk_ptrs.data(), scs.chunk_lengths.data(), scs.col_idxs.data(), scs.values.data(),
                                                X.data(), Yr.data(), &b, &ld, &rank);
        f_bscs(false, &C, &n_chunks, scs.chunk_ptrs.data(), scs.chunk_lengths.data(), scs.col_idxs.data(), scs.values.data(), X.data(), Y.data(), &b, &ld, &rank);
        report("OnePrecFuncPtr  SpMMV block_vec_size 4 dp, host arrays", std::memcmp(Y.data(), Yr.data(), Y.size() * 8) == 0, max_diff(Y, Yr));
    }
    // (5) adaptive precision dp + sp through MultiPrecFuncPtr, split by the reference's partition_precisions
    {
        Config cfg;
        cfg.value_type = "ap[dp_sp]";
        cfg.ap_threshold_1 = 0.5;
        cfg.equilibrate = 0;
        MtxData<double, int> dpm;
        MtxData<float, int> spm;
        std::vector<double> rowmax, colmax;
#ifdef HAVE_HALF_MATH
        MtxData<_Float16, int> hpm;
        partition_precisions<double, int>(&cfg, &m, &dpm, &spm, &hpm, &rowmax, &colmax, 0);
#else
        partition_precisions<double, int>(&cfg, &m, &dpm, &spm, &rowmax, &colmax, 0);
#endif
        ScsData<double, int> dps;
        ScsData<float, int> sps;
        convert_to_scs<double, double, int>(&dpm, C, sigma, &dps);
        convert_to_scs<float, float, int>(&spm, C, sigma, &sps, &dps.old_to_new_idx[0]);
        ST nc = dps.n_chunks;
        std::vector<double> xa(dps.n_rows_padded), ya_ref(dps.n_rows_padded, 0.0), ya(dps.n_rows_padded, -1.0);
        std::vector<float> xs(dps.n_rows_padded), ys(dps.n_rows_padded, 0.f);
        for (size_t i = 0; i < xa.size(); ++i) { xa[i] = std::sin(0.37 * i) + 1.5; xs[i] = (float)xa[i]; }
#ifdef HAVE_HALF_MATH
        ST zero = 0;
        spmv_omp_scs_ap_adv<int>(false, &C, &nc, dps.chunk_ptrs.data(), dps.chunk_lengths.data(), dps.col_idxs.data(), dps.values.data(), xa.data(),
                                 ya_ref.data(), &C, &nc, sps.chunk_ptrs.data(), sps.chunk_lengths.data(), sps.col_idxs.data(), sps.values.data(), xs.data(),
                                 ys.data(), &C, &nc, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, &rank);
        g_ap_scs(false, &C, &nc, dps.chunk_ptrs.data(), dps.chunk_lengths.data(), dps.col_idxs.data(), dps.values.data(), xa.data(), ya.data(), &C, &nc,
                 sps.chunk_ptrs.data(), sps.chunk_lengths.data(), sps.col_idxs.data(), sps.values.data(), xs.data(), ys.data(), &C, &zero, nullptr, nullptr,
                 nullptr, nullptr, nullptr, nullptr, &rank);
#else
        spmv_omp_scs_ap_adv<int>(false, &C, &nc, dps.chunk_ptrs.data(), dps.chunk_lengths.data(), dps.col_idxs.data(), dps.values.data(), xa.data(),
                                 ya_ref.data(), &C, &nc, sps.chunk_ptrs.data(), sps.chunk_lengths.data(), sps.col_idxs.data(), sps.values.data(), xs.data(),
                                 ys.data(), &rank);
        g_ap_scs(false, &C, &nc, dps.chunk_ptrs.data(), dps.chunk_lengths.data(), dps.col_idxs.data(), dps.values.data(), xa.data(), ya.data(), &C, &nc,
                 sps.chunk_ptrs.data(), sps.chunk_lengths.data(), sps.col_idxs.data(), sps.values.data(), xs.data(), ys.data(), &rank);
#endif
        // per-row scale sum |a||x| in the dp part's row order
        std::vector<double> scale(dps.n_rows_padded, 0.0);
        for (long k = 0; k < m.nnz; ++k) scale[dps.old_to_new_idx[m.I[k]]] += std::fabs(m.values[k] * xa[m.J[k]]);
        double worst = 0.0;
        bool ok = true;
        for (size_t i = 0; i < ya.size(); ++i) {
            const double d = std::fabs(ya[i] - ya_ref[i]);
            worst = std::fmax(worst, d);
            if (!(d <= 1e-6 * std::fmax(scale[i], 1e-300) + 1e-300)) ok = false;
        }
        report("MultiPrecFuncPtr ap[dp_sp] SELL-32-64, host arrays", ok, worst);
    }
    return fails ? 1 : 0;
}
