// MEASUREMENT INFRASTRUCTURE ONLY (never linked into the product): the second GPU baseline the reference itself offers — its
// USE_CUSPARSE comparison mode (code/utilities.hpp:3380-3550): cusparseSpMV with CUSPARSE_SPMV_ALG_DEFAULT on the matrix in CSR
// (kernel_format crs) or, for scs, in cuSPARSE's sliced-ELL format created DIRECTLY from the SELL-C-sigma arrays
// (cusparseCreateSlicedEll(rows, cols, nnz, n_elements, C, chunk_ptrs, col_idxs, values), utilities.hpp:3443-3457) — same calls, same
// arguments.  Host arrays in, y out, average time per cusparseSpMV from CUDA events.  Built by oracle/Makefile into
// oracle/_ref/libuspmv_cusparse.so; bench.py reports it inside `gpu_baseline`.
#include <cuda_runtime.h>
#include <cusparse.h>
#include <cstdio>

static char g_err[512];
#define CK(call)                                                                                                  \
    do {                                                                                                          \
        cudaError_t e__ = (call);                                                                                 \
        if (e__ != cudaSuccess) { std::snprintf(g_err, sizeof g_err, "%s at line %d", cudaGetErrorString(e__), __LINE__); return 1; } \
    } while (0)
#define CS(call)                                                                                                  \
    do {                                                                                                          \
        cusparseStatus_t s__ = (call);                                                                            \
        if (s__ != CUSPARSE_STATUS_SUCCESS) { std::snprintf(g_err, sizeof g_err, "cusparse: %s at line %d", cusparseGetErrorString(s__), __LINE__); return 1; } \
    } while (0)

extern "C" {
const char *cusp_last_error(void) { return g_err; }

/* kind: 0 = CSR (ptrs = row_ptrs[n_rows + 1], n_elements = nnz), 1 = sliced ELL (ptrs = chunk_ptrs[n_chunks + 1], slice size C).
 * vt: 0 = double, 1 = float.  x has n_cols entries, y_out n_rows (SELL: n_chunks * C are allocated). */
int cusp_spmv(int kind, int vt, long n_rows, long n_cols, long nnz, long n_elements, int C, const int *ptrs_h, long n_ptrs, const int *cols_h,
              const void *vals_h, const void *x_h, void *y_h, int warmup, int steps, double *ms_per_call) {
    g_err[0] = 0;
    const size_t es = vt == 0 ? 8 : 4;
    const cudaDataType dt = vt == 0 ? CUDA_R_64F : CUDA_R_32F;
    const long y_len = kind == 1 ? (n_ptrs - 1) * (long)C : n_rows;
    int *ptrs = nullptr, *cols = nullptr;
    void *vals = nullptr, *x = nullptr, *y = nullptr, *buf = nullptr;
    CK(cudaMalloc(&ptrs, n_ptrs * sizeof(int)));
    CK(cudaMalloc(&cols, (n_elements > 0 ? n_elements : 1) * sizeof(int)));
    CK(cudaMalloc(&vals, (n_elements > 0 ? n_elements : 1) * es));
    CK(cudaMalloc(&x, n_cols * es));
    CK(cudaMalloc(&y, y_len * es));
    CK(cudaMemcpy(ptrs, ptrs_h, n_ptrs * sizeof(int), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(cols, cols_h, n_elements * sizeof(int), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(vals, vals_h, n_elements * es, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(x, x_h, n_cols * es, cudaMemcpyHostToDevice));
    CK(cudaMemset(y, 0, y_len * es));
    cusparseHandle_t h = nullptr;
    cusparseSpMatDescr_t A;
    cusparseDnVecDescr_t vx, vy;
    CS(cusparseCreate(&h));
    if (kind == 0)
        CS(cusparseCreateCsr(&A, n_rows, n_cols, nnz, ptrs, cols, vals, CUSPARSE_INDEX_32I, CUSPARSE_INDEX_32I, CUSPARSE_INDEX_BASE_ZERO, dt));
    else
        CS(cusparseCreateSlicedEll(&A, n_rows, n_cols, nnz, n_elements, C, ptrs, cols, vals, CUSPARSE_INDEX_32I, CUSPARSE_INDEX_32I,
                                   CUSPARSE_INDEX_BASE_ZERO, dt));
    CS(cusparseCreateDnVec(&vx, n_cols, x, dt));
    CS(cusparseCreateDnVec(&vy, n_rows, y, dt));
    const double alpha_d = 1.0, beta_d = 0.0;
    const float alpha_f = 1.0f, beta_f = 0.0f;
    const void *alpha = vt == 0 ? (const void *)&alpha_d : (const void *)&alpha_f;
    const void *beta = vt == 0 ? (const void *)&beta_d : (const void *)&beta_f;
    size_t bytes = 0;
    CS(cusparseSpMV_bufferSize(h, CUSPARSE_OPERATION_NON_TRANSPOSE, alpha, A, vx, beta, vy, dt, CUSPARSE_SPMV_ALG_DEFAULT, &bytes));
    CK(cudaMalloc(&buf, bytes ? bytes : 16));
    for (int i = 0; i < warmup; ++i)
        CS(cusparseSpMV(h, CUSPARSE_OPERATION_NON_TRANSPOSE, alpha, A, vx, beta, vy, dt, CUSPARSE_SPMV_ALG_DEFAULT, buf));
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0, 0));
    for (int i = 0; i < steps; ++i)
        CS(cusparseSpMV(h, CUSPARSE_OPERATION_NON_TRANSPOSE, alpha, A, vx, beta, vy, dt, CUSPARSE_SPMV_ALG_DEFAULT, buf));
    CK(cudaEventRecord(e1, 0));
    CK(cudaEventSynchronize(e1));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    *ms_per_call = steps > 0 ? (double)ms / steps : 0.0;
    if (y_h) CK(cudaMemcpy(y_h, y, n_rows * es, cudaMemcpyDeviceToHost));
    cusparseDestroySpMat(A); cusparseDestroyDnVec(vx); cusparseDestroyDnVec(vy); cusparseDestroy(h);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(ptrs); cudaFree(cols); cudaFree(vals); cudaFree(x); cudaFree(y); cudaFree(buf);
    return 0;
}
}
