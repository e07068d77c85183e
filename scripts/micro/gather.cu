// Random 8-byte gather ceiling on this GPU (what bounds SpMV on the power-law matrix of BASELINE config 4, whose x does not fit L2):
// n index loads (streamed, coalesced) + n gathers x[idx] from a table of S MB, U independent gathers in flight per thread.
// mode 0: every lane of a warp hits a random line (32 L1 wavefronts per warp load); mode 1: the 32 lanes of a warp share one random
// 256-byte block (1-2 wavefronts) -- separates the L1 wavefront cost from the L2/DRAM sector cost.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather gather.cu ; ./gather
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int U>
__global__ void __launch_bounds__(256) k_gather(const int *__restrict__ idx, const double *__restrict__ x, long n, double *out) {
    double acc = 0.0;
    const long T = (long)gridDim.x * blockDim.x;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += U * T) {
        int c[U];
#pragma unroll
        for (int u = 0; u < U; ++u) c[u] = i + u * T < n ? __ldcs(idx + i + u * T) : 0;
#pragma unroll
        for (int u = 0; u < U; ++u) acc += __ldg(x + c[u]);
    }
    if (acc == 0.123) *out = acc;
}
__global__ void k_fill(int *idx, long n, long m, int mode) {
    long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t z = (mode ? (uint64_t)(i >> 5) : (uint64_t)i) * 0x9E3779B97F4A7C15ull + 0x5EED;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; z ^= z >> 31;
    long c = (long)(z % (uint64_t)m);
    if (mode) c = (c & ~31L) + (i & 31);
    idx[i] = (int)c;
}
template <int U> float run(const int *idx, const double *x, long n, double *out, int ctas) {
    k_gather<U><<<ctas, 256>>>(idx, x, n, out);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    for (int r = 0; r < 3; ++r) k_gather<U><<<ctas, 256>>>(idx, x, n, out);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms / 3;
}
int main() {
    int sm = 0;
    cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, 0);
    const long n = 1L << 28;  // 268 M gathers per launch (1 GB of indices: streamed, larger than L2)
    int *idx; cudaMalloc(&idx, n * 4);
    double *out; cudaMalloc(&out, 64);
    size_t g0 = 0;
    cudaDeviceGetLimit(&g0, cudaLimitMaxL2FetchGranularity);
    printf("default cudaLimitMaxL2FetchGranularity = %zu B\n", g0);
    for (size_t gran : {(size_t)0, (size_t)128, (size_t)64, (size_t)32}) {
      if (gran) {
          cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran);
          size_t g1 = 0;
          cudaDeviceGetLimit(&g1, cudaLimitMaxL2FetchGranularity);
          printf("---- cudaLimitMaxL2FetchGranularity set to %zu (%s), reads back %zu\n", gran, cudaGetErrorString(e), g1);
      }
    for (int mode = 0; mode < (gran ? 1 : 2); ++mode)
        for (long mb : {32, 96, 134, 268, 1024}) {
            if (gran && mb != 268 && mb != 1024) continue;
            const long m = mb * 1024 * 1024 / 8;
            double *x; cudaMalloc(&x, m * 8); cudaMemset(x, 0, m * 8);
            k_fill<<<(unsigned)((n + 255) / 256), 256>>>(idx, n, m, mode);
            for (int per_sm : {4, 8}) {
                float a = run<4>(idx, x, n, out, sm * per_sm), b = run<8>(idx, x, n, out, sm * per_sm), c = run<16>(idx, x, n, out, sm * per_sm);
                printf("%s table %4ld MB, %d CTAs/SM: U=4 %.1f  U=8 %.1f  U=16 %.1f G gathers/s\n", mode ? "warp-coalesced" : "lane-random   ", mb, per_sm,
                       n / a / 1e6, n / b / 1e6, n / c / 1e6);
            }
            cudaFree(x);
        }
    }
    return 0;
}
