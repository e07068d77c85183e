// TEST / MEASUREMENT INFRASTRUCTURE ONLY (never linked into the product): the reference's OWN CUDA kernels
//   spmv_gpu_scs_adv / scs_impl_gpu<C>   code/kernels.hpp:685-775
//   spmv_gpu_scs                         code/kernels.hpp:579-608
//   spmv_gpu_csr                         code/kernels.hpp:631-659
// compiled UNMODIFIED from the sources where they lie (#include of $(REF)/code/kernels.hpp; oracle/Makefile builds
// oracle/_ref/libuspmv_ref_gpu.so with nvcc -gencode arch=compute_100,code=sm_100 and the reference's THREADS_PER_BLOCK=128,
// config.mk:20), launched through the reference's own launchers with the reference's own grid size
// (n_thread_blocks = ceil(n_rows_padded / THREADS_PER_BLOCK), utilities.hpp:3739-3749), C and n_chunks as DEVICE pointers like the
// harness does.  bench.py reports it as `gpu_baseline` next to `cpu_baseline`: the recompiled-kernel baseline this engine replaces.
#include <cuda_runtime.h>
#include <cstdio>

using ST = long;  // mmio.h:21, classes_structs.hpp:31
#ifndef THREADS_PER_BLOCK
#define THREADS_PER_BLOCK 128
#endif
#include "kernels.hpp"

#define CK(call)                                                                                        \
    do {                                                                                                \
        cudaError_t e__ = (call);                                                                       \
        if (e__ != cudaSuccess) {                                                                       \
            std::snprintf(g_err, sizeof g_err, "%s at %s:%d", cudaGetErrorString(e__), __FILE__, __LINE__); \
            return 1;                                                                                   \
        }                                                                                               \
    } while (0)

static char g_err[512];

template <typename VT>
static int run(int kernel, long C, long n_chunks, const int *cp_h, const int *cl_h, const int *ci_h, const void *vals_h, long n_elements,
               const void *x_h, long x_len, void *y_h, int warmup, int steps, double *ms_per_launch) {
    const long n_pad = n_chunks * C;
    ST *C_d = nullptr, *nc_d = nullptr;
    int *cp = nullptr, *cl = nullptr, *ci = nullptr;
    VT *v = nullptr, *x = nullptr, *y = nullptr;
    CK(cudaMalloc(&C_d, sizeof(ST)));
    CK(cudaMalloc(&nc_d, sizeof(ST)));
    CK(cudaMalloc(&cp, (n_chunks + 1) * sizeof(int)));
    CK(cudaMalloc(&cl, (n_chunks > 0 ? n_chunks : 1) * sizeof(int)));
    CK(cudaMalloc(&ci, (n_elements > 0 ? n_elements : 1) * sizeof(int)));
    CK(cudaMalloc(&v, (n_elements > 0 ? n_elements : 1) * sizeof(VT)));
    CK(cudaMalloc(&x, x_len * sizeof(VT)));
    CK(cudaMalloc(&y, (n_pad > 0 ? n_pad : 1) * sizeof(VT)));
    CK(cudaMemcpy(C_d, &C, sizeof(ST), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(nc_d, &n_chunks, sizeof(ST), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(cp, cp_h, (n_chunks + 1) * sizeof(int), cudaMemcpyHostToDevice));
    if (cl_h) CK(cudaMemcpy(cl, cl_h, n_chunks * sizeof(int), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ci, ci_h, n_elements * sizeof(int), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(v, vals_h, n_elements * sizeof(VT), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(x, x_h, x_len * sizeof(VT), cudaMemcpyHostToDevice));
    CK(cudaMemset(y, 0, (n_pad > 0 ? n_pad : 1) * sizeof(VT)));
    const ST n_thread_blocks = (n_pad + THREADS_PER_BLOCK - 1) / THREADS_PER_BLOCK;
    int bvs = 1, vec_length = (int)x_len;
    auto launch = [&] {
        if (kernel == 0) spmv_gpu_scs_adv_launcher<VT, int>(false, C_d, nc_d, cp, cl, ci, v, x, y, &bvs, &vec_length, n_thread_blocks);
        else if (kernel == 1) spmv_gpu_scs_launcher<VT, int>(false, C_d, nc_d, cp, cl, ci, v, x, y, &bvs, &vec_length, n_thread_blocks);
        else spmv_gpu_csr_launcher<VT, int>(false, C_d, nc_d, cp, cl, ci, v, x, y, &bvs, &vec_length, n_thread_blocks);
    };
    for (int i = 0; i < warmup; ++i) launch();
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0, 0));
    for (int i = 0; i < steps; ++i) launch();
    CK(cudaEventRecord(e1, 0));
    CK(cudaEventSynchronize(e1));
    CK(cudaGetLastError());
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    *ms_per_launch = steps > 0 ? (double)ms / steps : 0.0;
    if (y_h) CK(cudaMemcpy(y_h, y, n_pad * sizeof(VT), cudaMemcpyDeviceToHost));
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(C_d); cudaFree(nc_d); cudaFree(cp); cudaFree(cl); cudaFree(ci); cudaFree(v); cudaFree(x); cudaFree(y);
    return 0;
}

extern "C" {
const char *refgpu_last_error(void) { return g_err; }
int refgpu_threads_per_block(void) { return THREADS_PER_BLOCK; }
/* kernel: 0 = spmv_gpu_scs_adv (templated C), 1 = spmv_gpu_scs (run-time C), 2 = spmv_gpu_csr (cp_h = row_ptrs, n_chunks = n_rows, C = 1).
 * vt: 0 = double, 1 = float.  Host arrays in, y (n_chunks * C values) out, average launch time in ms (CUDA events). */
int refgpu_spmv(int kernel, int vt, long C, long n_chunks, const int *cp_h, const int *cl_h, const int *ci_h, const void *vals_h,
                long n_elements, const void *x_h, long x_len, void *y_h, int warmup, int steps, double *ms_per_launch) {
    g_err[0] = 0;
    if (vt == 0) return run<double>(kernel, C, n_chunks, cp_h, cl_h, ci_h, vals_h, n_elements, x_h, x_len, y_h, warmup, steps, ms_per_launch);
    if (vt == 1) return run<float>(kernel, C, n_chunks, cp_h, cl_h, ci_h, vals_h, n_elements, x_h, x_len, y_h, warmup, steps, ms_per_launch);
    std::snprintf(g_err, sizeof g_err, "refgpu_spmv: value type %d not built (the reference's GPU half path needs HAVE_HALF_MATH + cuda_fp16)", vt);
    return 1;
}
}
