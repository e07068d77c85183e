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
    return max_diff < 1e-12 ? 0 : 1;
}
