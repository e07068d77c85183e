/* TEST INFRASTRUCTURE ONLY (oracle/).
 *
 * extern "C" driver around the reference's LIBRARY header code/interface.hpp (API_doc.md).  That
 * header does not compile as shipped (a stray token on line 532; it also relies on the includer to
 * define ST, CHUNK_SIZE and SIGMA).  oracle/Makefile therefore pipes it through `sed 532d` into a
 * temporary directory at build time (never into this repository) and passes that directory with -I;
 * the patched file is deleted after the compile.  Only the templates that are instantiated below are
 * used; the string-compared dispatchers (partition_precisions, execute_uspmv: `char* == "literal"`,
 * interface.hpp:756,1914) are dead at run time and are NOT used as an oracle.
 *
 * Entry points wrapped (all in /root/reference/code/interface.hpp):
 *   convert_to_scs  :401-656     permute_scs_cols :659-688 (not instantiable: vector<IT,IT> at :669)
 *   uspmv_csr_cpu   :988-1018    uspmv_scs_cpu    :1023-1053
 *   uspmv_csr_ap{dpsp,dphp,sphp,dpsphp}_cpu :1129-1429
 *   uspmv_scs_ap{dpsp,dphp,sphp,dpsphp}_cpu :1434-1733
 */
#include <vector>
#include <iostream>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <utility>

using ST = long;
static long CHUNK_SIZE = 1;
static long SIGMA = 1;

#include "interface_patched.hpp"

extern "C" {

/* mode: 0 = dp_sp, 1 = dp_hp, 2 = sp_hp, 3 = dp_sp_hp.  Unused parts may be NULL.
 * y is double for modes 0,1,3 and float for mode 2 (interface.hpp:1644). */
void iface_ap_scs(int mode, long C, long n_chunks,
                  const int *dcp, const int *dcl, const int *dci, const double *dv,
                  const int *scp, const int *scl, const int *sci, const float *sv,
                  const int *hcp, const int *hcl, const int *hci, const void *hv,
                  double *dp_x, float *sp_x, void *y) {
    const _Float16 *hvv = static_cast<const _Float16 *>(hv);
    double *dy = static_cast<double *>(y);
    float *sy = static_cast<float *>(y);
    switch (mode) {
    case 0:
        uspmv_scs_apdpsp_cpu<double, int>(&C, &n_chunks, dcp, dcl, dci, dv, dp_x, dy, &C, &n_chunks, scp, scl, sci, sv, sp_x, nullptr,
                                          &C, &n_chunks, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
        break;
    case 1:
        uspmv_scs_apdphp_cpu<double, int>(&C, &n_chunks, dcp, dcl, dci, dv, dp_x, dy, &C, &n_chunks, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr,
                                          &C, &n_chunks, hcp, hcl, hci, hvv, nullptr, nullptr);
        break;
    case 2:
        uspmv_scs_apsphp_cpu<double, int>(&C, &n_chunks, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, &C, &n_chunks, scp, scl, sci, sv, sp_x, sy,
                                          &C, &n_chunks, hcp, hcl, hci, hvv, nullptr, nullptr);
        break;
    default:
        uspmv_scs_apdpsphp_cpu<double, int>(&C, &n_chunks, dcp, dcl, dci, dv, dp_x, dy, &C, &n_chunks, scp, scl, sci, sv, sp_x, nullptr,
                                            &C, &n_chunks, hcp, hcl, hci, hvv, nullptr, nullptr);
    }
}

void iface_ap_csr(int mode, long n_rows,
                  const int *drp, const int *dci, const double *dv,
                  const int *srp, const int *sci, const float *sv,
                  const int *hrp, const int *hci, const void *hv,
                  double *dp_x, float *sp_x, void *y) {
    long C = 1;
    const _Float16 *hvv = static_cast<const _Float16 *>(hv);
    double *dy = static_cast<double *>(y);
    float *sy = static_cast<float *>(y);
    switch (mode) {
    case 0:
        uspmv_csr_apdpsp_cpu<int>(&C, &n_rows, drp, nullptr, dci, dv, dp_x, dy, &C, &n_rows, srp, nullptr, sci, sv, sp_x, nullptr,
                                  &C, &n_rows, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
        break;
    case 1:
        uspmv_csr_apdphp_cpu<int>(&C, &n_rows, drp, nullptr, dci, dv, dp_x, dy, &C, &n_rows, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr,
                                  &C, &n_rows, hrp, nullptr, hci, hvv, nullptr, nullptr);
        break;
    case 2:
        uspmv_csr_apsphp_cpu<int>(&C, &n_rows, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, &C, &n_rows, srp, nullptr, sci, sv, sp_x, sy,
                                  &C, &n_rows, hrp, nullptr, hci, hvv, nullptr, nullptr);
        break;
    default:
        uspmv_csr_apdpsphp_cpu<int>(&C, &n_rows, drp, nullptr, dci, dv, dp_x, dy, &C, &n_rows, srp, nullptr, sci, sv, sp_x, nullptr,
                                    &C, &n_rows, hrp, nullptr, hci, hvv, nullptr, nullptr);
    }
}

/* Library SCS kernel on raw arrays (dp only; the harness twin is covered by ref_driver.cpp). */
void iface_scs_f64(int C, int n_chunks, const int *cp, const int *cl, const int *ci, const double *v, double *x, double *y) {
    uspmv_scs_cpu<double, double, int>(C, n_chunks, cp, cl, ci, v, x, y);
}

/* Library convert_to_scs, to confirm it agrees with the harness copy. */
long iface_convert_f64(long n_rows, long n_cols, long nnz, const int *I, const int *J, const double *vals, long C, long sigma,
                       int *chunk_ptrs, int *chunk_lengths, int *col_idxs, double *values, int *old_to_new, long cap) {
    MtxData<double, int> m;
    m.n_rows = n_rows; m.n_cols = n_cols; m.nnz = nnz; m.is_sorted = true; m.is_symmetric = false;
    m.I.assign(I, I + nnz); m.J.assign(J, J + nnz); m.values.assign(vals, vals + nnz);
    ScsData<double, int> s;
    convert_to_scs<double, double, int>(&m, C, sigma, &s);
    if (s.n_elements > cap) return -s.n_elements;
    std::memcpy(chunk_ptrs, s.chunk_ptrs.data(), sizeof(int) * (s.n_chunks + 1));
    std::memcpy(chunk_lengths, s.chunk_lengths.data(), sizeof(int) * s.n_chunks);
    std::memcpy(col_idxs, s.col_idxs.data(), sizeof(int) * s.n_elements);
    std::memcpy(values, s.values.data(), sizeof(double) * s.n_elements);
    std::memcpy(old_to_new, s.old_to_new_idx.data(), sizeof(int) * s.n_rows);
    return s.n_elements;
}

}  // extern "C"
