// L2 -> SM read bandwidth ceiling on this GPU: every SM re-reads an L2-resident buffer with 128-bit read-only loads.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l2bw l2bw.cu ; ./l2bw
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(1024, 2) k_read(const int4 *__restrict__ p, size_t n, int reps, int4 *out) {
    int4 acc = make_int4(0, 0, 0, 0);
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (int r = 0; r < reps; ++r)
        for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += 4 * stride) {
            int4 a = __ldg(p + i), b = i + stride < n ? __ldg(p + i + stride) : a, c = i + 2 * stride < n ? __ldg(p + i + 2 * stride) : a,
                 d = i + 3 * stride < n ? __ldg(p + i + 3 * stride) : a;
            acc.x ^= a.x ^ b.x ^ c.x ^ d.x; acc.y ^= a.y ^ b.y ^ c.y ^ d.y; acc.z ^= a.z ^ b.z ^ c.z ^ d.z; acc.w ^= a.w ^ b.w ^ c.w ^ d.w;
        }
    if (acc.x == 0x12345678) *out = acc;
}
int main() {
    int sm = 0;
    cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, 0);
    int4 *out; cudaMalloc(&out, 64);
    for (size_t mb : {8, 16, 32, 64, 96, 256, 2048}) {
        size_t n = mb * 1024 * 1024 / 16;
        int4 *p; cudaMalloc(&p, n * 16); cudaMemset(p, 1, n * 16);
        int reps = mb >= 256 ? 4 : 40;
        k_read<<<sm * 2, 1024>>>(p, n, 2, out);
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        k_read<<<sm * 2, 1024>>>(p, n, reps, out);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("buffer %5zu MB: %.0f GB/s\n", mb, (double)n * 16 * reps / ms / 1e6);
        cudaFree(p);
    }
    return 0;
}
