// Which load instruction moves the least data per random 8-byte gather?  Same loop as gather.cu (table >> L2), one variant per PTX
// load form.  ncu of the fused AP kernel on the 2^25-row power-law matrix showed ~128 B of DRAM traffic and ~3.5 L2 sectors per gathered
// element with ld.global.nc (LDG.E.64.CONSTANT).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather2 gather2.cu ; ./gather2
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int V> __device__ __forceinline__ double ld(const double *p) {
    double v;
    if (V == 0) return __ldg(p);
    else if (V == 1) asm volatile("ld.global.f64 %0, [%1];" : "=d"(v) : "l"(p));
    else if (V == 2) asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(v) : "l"(p));
    else if (V == 3) asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
    else if (V == 4) asm volatile("ld.global.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
    else if (V == 5) asm volatile("ld.global.cs.f64 %0, [%1];" : "=d"(v) : "l"(p));
    else if (V == 6) asm volatile("ld.global.L1::evict_first.f64 %0, [%1];" : "=d"(v) : "l"(p));
    else if (V == 7) asm volatile("ld.global.relaxed.gpu.f64 %0, [%1];" : "=d"(v) : "l"(p));
    else if (V == 8) asm volatile("ld.global.nc.L2::64B.f64 %0, [%1];" : "=d"(v) : "l"(p));
    else if (V == 9) asm volatile("ld.global.L2::64B.f64 %0, [%1];" : "=d"(v) : "l"(p));
    else if (V == 10) asm volatile("ld.global.nc.L2::128B.f64 %0, [%1];" : "=d"(v) : "l"(p));
    else if (V == 11) asm volatile("ld.global.nc.L1::no_allocate.L2::64B.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
template <int V, int U>
__global__ void __launch_bounds__(256) k_gather(const int *__restrict__ idx, const double *x, long n, double *out) {
    double acc = 0.0;
    const long T = (long)gridDim.x * blockDim.x;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += U * T) {
        int c[U];
#pragma unroll
        for (int u = 0; u < U; ++u) c[u] = i + u * T < n ? __ldcs(idx + i + u * T) : 0;
        double v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = ld<V>(x + c[u]);
#pragma unroll
        for (int u = 0; u < U; ++u) acc += v[u];
    }
    if (acc == 0.123) *out = acc;
}
__global__ void k_fill(int *idx, long n, long m) {
    long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t z = (uint64_t)i * 0x9E3779B97F4A7C15ull + 0x5EED;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; z ^= z >> 31;
    idx[i] = (int)(z % (uint64_t)m);
}
template <int V> float run(const int *idx, const double *x, long n, double *out, int ctas) {
    k_gather<V, 8><<<ctas, 256>>>(idx, x, n, out);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    for (int r = 0; r < 3; ++r) k_gather<V, 8><<<ctas, 256>>>(idx, x, n, out);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms / 3;
}
int main() {
    int sm = 0;
    cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, 0);
    const long n = 1L << 28;
    int *idx; cudaMalloc(&idx, n * 4);
    double *out; cudaMalloc(&out, 64);
    const char *names[12] = {"ld.global.nc (__ldg)", "ld.global", "ld.global.cg", "ld.global.nc.L1::no_allocate", "ld.global.L1::no_allocate", "ld.global.cs",
                             "ld.global.L1::evict_first", "ld.global.relaxed.gpu", "ld.global.nc.L2::64B", "ld.global.L2::64B", "ld.global.nc.L2::128B",
                             "ld.global.nc.L1::no_allocate.L2::64B"};
    for (long mb : {268, 1024}) {
        const long m = mb * 1024 * 1024 / 8;
        double *x; cudaMalloc(&x, m * 8); cudaMemset(x, 0, m * 8);
        k_fill<<<(unsigned)((n + 255) / 256), 256>>>(idx, n, m);
        float t[12] = {run<0>(idx, x, n, out, sm * 8), run<1>(idx, x, n, out, sm * 8), run<2>(idx, x, n, out, sm * 8), run<3>(idx, x, n, out, sm * 8),
                       run<4>(idx, x, n, out, sm * 8), run<5>(idx, x, n, out, sm * 8), run<6>(idx, x, n, out, sm * 8), run<7>(idx, x, n, out, sm * 8),
                       run<8>(idx, x, n, out, sm * 8), run<9>(idx, x, n, out, sm * 8), run<10>(idx, x, n, out, sm * 8), run<11>(idx, x, n, out, sm * 8)};
        for (int v = 0; v < 12; ++v) printf("table %4ld MB  %-32s %.1f G gathers/s\n", mb, names[v], n / t[v] / 1e6);
        cudaFree(x);
    }
    return 0;
}
