// Solve-mode report of the `uspmv` harness clone: the same file names and table layout as the reference's write_result_to_file
// (code/write_results.hpp:160-440) — spmv_mkl_compare_{dp,sp,hp,ap}.txt, a two-line run description, then either one summary row
// (verbose 0) or one row per result element (verbose 1) — so that scripts written against the reference's output keep working.
// The "mkl" columns hold the host-side COO product in double that stands in for the reference's MKL validator
// (write_results.hpp:442-556 needs MKL, which is neither in this image nor on the path).
#pragma once
#include <cmath>
#include <fstream>
#include <iomanip>
#include <string>
#include <vector>

struct ResultReport {
    std::string matrix_file_name, kernel_format, value_type, block_vec_layout, seg_method;
    long chunk_size = 1, sigma = 1, n_blocks = 0;
    int threads_per_block = 256, ranks = 1, verbose = 0;
    unsigned long revisions = 1;
    double beta = 0.0;
};

// ref / got: block_vec_size vectors of n_rows entries each, vector v at [v * n_rows, (v + 1) * n_rows).
// Returns the largest relative difference over the elements whose reference value is not zero (the reference divides by the MKL value
// and would print inf there, write_results.hpp:362-372).
inline double write_result_to_file(const ResultReport &c, const std::vector<double> &ref, const std::vector<double> &got, long n_rows) {
    const bool is_ap = c.value_type.rfind("ap[", 0) == 0;
    std::ofstream f("spmv_mkl_compare_" + (is_ap ? std::string("ap") : c.value_type) + ".txt", std::ios::app);
    f << c.matrix_file_name << " with ";
    if (c.ranks > 1) f << c.ranks << " MPI processes, and ";
    f << c.n_blocks << " block(s), and " << c.threads_per_block << " thread(s) per block" << std::endl;
    f << "kernel: " << c.kernel_format;
    if (c.kernel_format == "scs") f << ", C: " << c.chunk_size << ", sigma: " << c.sigma << std::fixed << std::setprecision(8) << ", beta: " << c.beta;
    f << ", block_vec_layout: " << c.block_vec_layout << ", data_type: " << c.value_type << ", revisions: " << c.revisions;
    if (c.ranks > 1) f << ", seg_method: " << c.seg_method << ", MPI_mode: bulkvec";
    f << std::endl;

    const long total = (long)got.size();
    const int digits = total > 0 ? (int)std::log10((double)total) + 1 : 1;
    int width = c.verbose ? 24 : 18;
    if (c.verbose) {
        f << std::left << std::setw(digits + 8) << "vec idx:" << std::setw(digits + 8) << "row idx:" << std::setw(width) << "mkl results:" << std::setw(width)
          << "uspmv results:" << std::setw(width) << "rel. diff(%):" << std::setw(width) << "abs. diff:" << std::endl;
        f << std::left << std::setw(digits + 8) << "--------" << std::setw(digits + 8) << "--------" << std::setw(width) << "-----------" << std::setw(width)
          << "------------" << std::setw(width) << "------------" << std::setw(width) << "---------" << std::endl;
    } else {
        f << std::left << std::setw(width - 2) << "mkl rel. elem:" << std::setw(width) << "uspmv rel. elem:" << std::setw(width) << "MAX rel. diff(%):"
          << std::setw(width - 1) << "mkl abs. elem:" << std::setw(width) << "uspmv abs. elem:" << std::setw(width) << "MAX abs. diff:" << std::setw(width + 4)
          << "||mkl - uspmv||_2" << std::setw(width + 4) << "||mkl - uspmv||/||mkl||_2" << std::endl;
        f << std::left << std::setw(width - 2) << "-------------" << std::setw(width) << "---------------" << std::setw(width) << "----------------"
          << std::setw(width - 1) << "-------------" << std::setw(width) << "---------------" << std::setw(width) << "-------------" << std::setw(width + 4)
          << "-----------------" << std::setw(width + 4) << "-------------------------" << std::endl;
    }
    double max_rel = 0.0, max_abs = 0.0, rel_ref = 0.0, rel_got = 0.0, abs_ref = 0.0, abs_got = 0.0, dist2 = 0.0, mag2 = 0.0;
    for (long i = 0; i < total; ++i) {
        const double r = ref[i], g = got[i];
        const double ad = std::fabs(r - g);
        const double rd = r != 0.0 ? std::fabs((r - g) / r) : 0.0;
        dist2 += (r - g) * (r - g);
        mag2 += r * r;
        if (rd > max_rel || std::isnan(rd)) { max_rel = rd; rel_ref = r; rel_got = g; }
        if (ad > max_abs || std::isnan(ad)) { max_abs = ad; abs_ref = r; abs_got = g; }
        if (c.verbose) {
            f << std::left << std::setw(digits + 8) << (n_rows ? i / n_rows : 0) << std::setw(digits + 8) << (n_rows ? i % n_rows : i) << std::setprecision(16)
              << std::scientific << std::setw(width) << r << std::setw(width) << g << std::setw(width) << 100 * rd << std::setw(width) << ad;
            if (rd > .01 || std::isinf(rd) || std::isnan(rd)) f << std::setw(width) << "ERROR";
            else if (rd > .0001) f << std::setw(width) << "WARNING";
            f << std::endl;
        }
    }
    if (!c.verbose) {
        const double dist = std::sqrt(dist2), mag = std::sqrt(mag2);
        f << std::scientific << std::left << std::setw(width) << rel_ref << std::setw(width) << rel_got << std::setw(width) << 100 * max_rel << std::setw(width)
          << abs_ref << std::setw(width) << abs_got << std::setw(width) << max_abs << std::setw(width + 6) << dist << std::setw(width + 6)
          << (mag > 0 ? dist / mag : 0.0);
        if (max_rel > .01 || std::isnan(max_rel) || std::isinf(max_rel) || std::isnan(max_abs) || std::isinf(max_abs)) f << std::setw(width) << "ERROR";
        else if (max_rel > .0001) f << std::setw(width) << "WARNING";
        f << std::endl;
    }
    f << "\n";
    return max_rel;
}
