// Library use exactly as the reference documents it (API_doc.md), through include/uspmv_interface.hpp:
//   MtxData -> convert_to_scs -> permute_scs_cols -> apply_permutation(x) -> uspmv_scs_gpu -> apply_permutation(y)
// Prints the max abs difference to a host COO product and exits non-zero if it is not tiny.
#include <cmath>
#include <cstdio>
#include <vector>

#include "../../include/uspmv_interface.hpp"

int main() {
    const int n = 1000;
    MtxData<double, int> mtx;  // tridiagonal test matrix, row-sorted COO
    mtx.n_rows = mtx.n_cols = n;
    mtx.is_sorted = true;
    for (int i = 0; i < n; ++i)
        for (int j = i - 1; j <= i + 1; ++j)
            if (j >= 0 && j < n && !(i % 7 == 3 && j != i)) {  // rows with 1, 2 and 3 entries: sigma sorting has something to do
                mtx.I.push_back(i); mtx.J.push_back(j); mtx.values.push_back(i == j ? 2.0 + i * 1e-3 : -1.0);
            }
    mtx.nnz = (ST)mtx.values.size();
    ScsData<double, int> scs;
    convert_to_scs<double, double, int>(&mtx, 32, 128, &scs);
    permute_scs_cols(&scs, scs.old_to_new_idx.data());

    std::vector<double> x(scs.n_rows_padded, 0.0), xp(scs.n_rows_padded, 0.0), yp(scs.n_rows_padded, 0.0), y(n, 0.0);
    for (int i = 0; i < n; ++i) x[i] = std::sin(0.1 * i);
    apply_permutation(xp.data(), x.data(), scs.new_to_old_idx, n);

    uspmv_ctx *ctx = uspmv_detail::default_ctx();
    void *xd, *yd;
    uspmv_detail::check(uspmv_malloc(ctx, xp.size() * 8, &xd));
    uspmv_detail::check(uspmv_malloc(ctx, yp.size() * 8, &yd));
    uspmv_detail::check(uspmv_memcpy_h2d(ctx, xd, xp.data(), xp.size() * 8, nullptr));
    const int *cp, *cl, *ci; const void *vals;
    uspmv_detail::check(uspmv_scs_device_arrays(scs.device.get(), &cp, &cl, &ci, &vals, nullptr, nullptr));
    uspmv_scs_gpu<double, int>(scs.C, scs.n_chunks, cp, cl, ci, static_cast<const double *>(vals), static_cast<const double *>(xd), static_cast<double *>(yd));
    uspmv_detail::check(uspmv_memcpy_d2h(ctx, yp.data(), yd, yp.size() * 8, nullptr));
    apply_permutation(y.data(), yp.data(), scs.old_to_new_idx.data(), n);

    double max_diff = 0.0;
    std::vector<double> yr(n, 0.0);
    for (ST k = 0; k < mtx.nnz; ++k) yr[mtx.I[k]] += mtx.values[k] * x[mtx.J[k]];
    for (int i = 0; i < n; ++i) max_diff = std::fmax(max_diff, std::fabs(y[i] - yr[i]));
    std::printf("example_interface: n=%d nnz=%ld n_chunks=%ld n_elements=%ld max|y - y_coo| = %.3e\n", n, mtx.nnz, scs.n_chunks, scs.n_elements, max_diff);
    if (!(max_diff < 1e-12)) return 1;

    // ---- adaptive precision exactly as the harness wires it (main.cpp:1170-1221): partition_precisions, the first part sorted,
    //      the others built with fixed_permutation = first.old_to_new_idx, then one execute_uspmv call over all parts ----
    using half_t = uspmv_detail::uspmv_half_bits;
    MtxData<double, int> dp_m;
    MtxData<float, int> sp_m;
    MtxData<half_t, int> hp_m;
    std::vector<double> none;
    partition_precisions<double, int>(&mtx, &dp_m, &sp_m, &hp_m, &none, &none, /*t1*/ 2.5, /*t2*/ 1.5, "ap[dp_sp_hp]", false);
    if (dp_m.nnz + sp_m.nnz + hp_m.nnz != mtx.nnz) return 2;
    ScsData<double, int> dp_s;
    ScsData<float, int> sp_s;
    ScsData<half_t, int> hp_s;
    convert_to_scs<double, double, int>(&dp_m, 32, 128, &dp_s);
    convert_to_scs<float, float, int>(&sp_m, 32, 128, &sp_s, dp_s.old_to_new_idx.data());
    convert_to_scs<half_t, half_t, int>(&hp_m, 32, 128, &hp_s, dp_s.old_to_new_idx.data());
    // AP structs keep original column numbering, so x is used un-permuted; y comes out in the dp part's row order
    std::vector<double> xa(dp_s.n_rows_padded, 0.0), ya(dp_s.n_rows_padded, 0.0);
    for (int i = 0; i < n; ++i) xa[i] = x[i];
    uspmv_detail::check(uspmv_memcpy_h2d(ctx, xd, xa.data(), xa.size() * 8, nullptr));
    execute_uspmv<double, int, half_t>(&dp_s, nullptr, nullptr, &dp_s, &sp_s, &hp_s, xd, yd, "ap[dp_sp_hp]");
    uspmv_detail::check(uspmv_memcpy_d2h(ctx, ya.data(), yd, ya.size() * 8, nullptr));
    double ap_diff = 0.0;
    for (int i = 0; i < n; ++i) ap_diff = std::fmax(ap_diff, std::fabs(ya[dp_s.old_to_new_idx[i]] - yr[i]));
    std::printf("example_interface: ap[dp_sp_hp] split %ld / %ld / %ld, max|y - y_coo| = %.3e (fp32/fp16 storage of the small entries)\n",
                dp_m.nnz, sp_m.nnz, hp_m.nnz, ap_diff);
    if (!(ap_diff < 5e-3)) return 3;

    // ---- execute_uspmv with the reference's OWN signature (interface.hpp:1871-1910): pointer bundles of HOST arrays, scalars by pointer,
    //      `char *ap_value_type` — the call a user of the reference's library makes, unchanged ----
    char dp_type[] = "dp", ap_type[] = "ap[dp_sp_hp]";
    ST Cs = scs.C, ncs = scs.n_chunks;
    std::vector<double> y2(scs.n_rows_padded, -1.0);
    execute_uspmv<double, int>(&Cs, &ncs, scs.chunk_ptrs.data(), scs.chunk_lengths.data(), scs.col_idxs.data(), scs.values.data(), xp.data(), y2.data(),
                               dp_type);
    double d2 = 0.0;
    for (long i = 0; i < scs.n_rows_padded; ++i) d2 = std::fmax(d2, std::fabs(y2[i] - yp[i]));
    std::printf("example_interface: execute_uspmv(pointer bundle, host arrays): max|y - y(uspmv_scs_gpu)| = %.3e\n", d2);
    if (d2 != 0.0) return 4;
    ST Cd = dp_s.C, ncd = dp_s.n_chunks;
    std::vector<double> y3(dp_s.n_rows_padded, -1.0);
    std::vector<float> xs(xa.begin(), xa.end()), ysp(dp_s.n_rows_padded, 0.f);
    std::vector<half_t> xh(dp_s.n_rows_padded), yh(dp_s.n_rows_padded);
    execute_uspmv<double, int, half_t>(&Cd, &ncd, dp_s.chunk_ptrs.data(), dp_s.chunk_lengths.data(), dp_s.col_idxs.data(), dp_s.values.data(), xa.data(),
                                       y3.data(),
                                       &Cd, &ncd, dp_s.chunk_ptrs.data(), dp_s.chunk_lengths.data(), dp_s.col_idxs.data(), dp_s.values.data(), xa.data(), y3.data(),
                                       &Cd, &ncd, sp_s.chunk_ptrs.data(), sp_s.chunk_lengths.data(), sp_s.col_idxs.data(), sp_s.values.data(), xs.data(), ysp.data(),
                                       &Cd, &ncd, hp_s.chunk_ptrs.data(), hp_s.chunk_lengths.data(), hp_s.col_idxs.data(), hp_s.values.data(), xh.data(), yh.data(),
                                       ap_type);
    double d3 = 0.0;
    for (long i = 0; i < dp_s.n_rows_padded; ++i) d3 = std::fmax(d3, std::fabs(y3[i] - ya[i]));
    std::printf("example_interface: execute_uspmv(dp + sp + hp bundles, host arrays): max|y - y(device call)| = %.3e\n", d3);
    return d3 == 0.0 ? 0 : 5;
}
