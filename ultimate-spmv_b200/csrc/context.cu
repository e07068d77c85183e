// Context, error state and device-memory helpers of the C ABI (include/uspmv_b200.h).
#include "common.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

namespace uspmv {
static thread_local std::string t_error;
std::atomic<long> g_launches{0};
void set_error(const std::string &msg) { t_error = msg; }
static thread_local const Options *t_active = nullptr;
static std::mutex g_ctx_mutex;
static std::vector<uspmv_ctx *> g_live_ctxs;
const Options &options() { return t_active ? *t_active : default_options(); }
OptScope::OptScope(const uspmv_ctx *ctx) : prev(t_active) { if (ctx) t_active = &ctx->opts; }
OptScope::~OptScope() { t_active = prev; }
Options &default_options() {
    static Options o = [] {
        Options c;
        if (const char *e = std::getenv("USPMV_SCS_KERNEL")) c.scs_stream = std::strcmp(e, "direct") != 0;
        if (const char *e = std::getenv("USPMV_STREAM_VARIANT")) c.stream_variant = std::atoi(e);
        if (const char *e = std::getenv("USPMV_STREAM_BPS")) c.stream_blocks_per_sm = std::max(1, std::atoi(e));
        if (const char *e = std::getenv("USPMV_STRICT_REFERENCE_HALO")) c.strict_reference_halo = std::atoi(e) != 0;
        if (const char *e = std::getenv("USPMV_L2_FETCH")) c.l2_fetch_granularity = std::atoi(e);
        if (const char *e = std::getenv("USPMV_PUSH_VARIANT")) c.push_variant = std::atoi(e);
        if (const char *e = std::getenv("USPMV_PUSH_CTAS_PER_SM")) c.push_ctas_per_sm = std::max(0, std::atoi(e));
        return c;
    }();
    return o;
}
}  // namespace uspmv

using namespace uspmv;

extern "C" {

const char *uspmv_last_error(void) { return t_error.c_str(); }
int uspmv_version(void) { return 100; }
long uspmv_kernel_launches(void) { return g_launches.load(); }

static void apply_option(Options &c, const char *name, long value, int device) {
        if (!std::strcmp(name, "scs_stream")) c.scs_stream = value != 0;
        else if (!std::strcmp(name, "scs_stream_wide")) c.scs_stream_wide = value != 0;
        else if (!std::strcmp(name, "narrow_dp")) c.narrow_dp = value != 0;
        else if (!std::strcmp(name, "pair_hp")) c.pair_hp = value != 0;
        else if (!std::strcmp(name, "mmv_fused_rowwise")) c.mmv_fused_rowwise = (int)value;
        else if (!std::strcmp(name, "mmv_push_first")) c.mmv_push_first = value != 0;
        else if (!std::strcmp(name, "fused_narrow")) c.fused_narrow = value != 0;
        else if (!std::strcmp(name, "stream_variant")) c.stream_variant = (int)value;
        else if (!std::strcmp(name, "stream_blocks_per_sm")) c.stream_blocks_per_sm = (int)std::max(1L, value);
        else if (!std::strcmp(name, "mmv_variant")) c.mmv_variant = (int)std::max(0L, value);
        else if (!std::strcmp(name, "mmv_blocks_per_sm")) c.mmv_blocks_per_sm = (int)std::max(0L, value);
        else if (!std::strcmp(name, "ap_variant")) c.ap_variant = (int)std::max(0L, value);
        else if (!std::strcmp(name, "split_long_chunks")) c.split_long_chunks = (int)std::max(0L, value);
        else if (!std::strcmp(name, "strict_reference_halo")) c.strict_reference_halo = value != 0;
        else if (!std::strcmp(name, "push_variant")) c.push_variant = (int)value;
        else if (!std::strcmp(name, "push_ctas_per_sm")) c.push_ctas_per_sm = (int)std::max(0L, value);
        else if (!std::strcmp(name, "push_min_elements")) c.push_min_elements = std::max(0L, value);
        else if (!std::strcmp(name, "l2_fetch_granularity")) {
            if (value != 0 && value != 32 && value != 64 && value != 128) fail("uspmv_set_option: l2_fetch_granularity must be 0, 32, 64 or 128");
            c.l2_fetch_granularity = (int)value;
            if (value) {
                if (device >= 0) USPMV_CUDA(cudaSetDevice(device));
                USPMV_CUDA(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)value));  // a device limit: the context's (or current) device
            }
        }
        else fail("uspmv_set_option: unknown option '%s'", name);
}

/* process-wide: the defaults new contexts start from AND every live context */
int uspmv_set_option(const char *name, long value) {
    return guarded([&] {
        if (!name) fail("uspmv_set_option: name is NULL");
        apply_option(default_options(), name, value, -1);
        std::lock_guard<std::mutex> g(g_ctx_mutex);
        for (uspmv_ctx *c : g_live_ctxs) apply_option(c->opts, name, value, -1);
    });
}

/* one context only: two contexts (e.g. two GPUs, or two solver instances on one GPU) can run different kernel variants */
int uspmv_ctx_set_option(uspmv_ctx *ctx, const char *name, long value) {
    return guarded([&] {
        if (!ctx || !name) fail("uspmv_ctx_set_option: NULL argument");
        apply_option(ctx->opts, name, value, ctx->device);
    });
}

int uspmv_ctx_get_option(const uspmv_ctx *ctx, const char *name, long *out) {
    return guarded([&] {
        if (!name || !out) fail("uspmv_ctx_get_option: NULL argument");
        const Options &c = ctx ? ctx->opts : default_options();
        if (!std::strcmp(name, "scs_stream")) *out = c.scs_stream;
        else if (!std::strcmp(name, "scs_stream_wide")) *out = c.scs_stream_wide;
        else if (!std::strcmp(name, "stream_variant")) *out = c.stream_variant;
        else if (!std::strcmp(name, "stream_blocks_per_sm")) *out = c.stream_blocks_per_sm;
        else if (!std::strcmp(name, "mmv_variant")) *out = c.mmv_variant;
        else if (!std::strcmp(name, "ap_variant")) *out = c.ap_variant;
        else if (!std::strcmp(name, "split_long_chunks")) *out = c.split_long_chunks;
        else if (!std::strcmp(name, "strict_reference_halo")) *out = c.strict_reference_halo;
        else if (!std::strcmp(name, "push_variant")) *out = c.push_variant;
        else if (!std::strcmp(name, "pair_hp")) *out = c.pair_hp;
        else fail("uspmv_ctx_get_option: unknown option '%s'", name);
    });
}

/* a second CUDA stream for the halo exchange next to the SpMV stream (host code without CUDA headers: the `uspmv` harness) */
int uspmv_stream_create(uspmv_ctx *ctx, void **out) {
    return guarded([&] {
        if (!ctx || !out) fail("uspmv_stream_create: NULL argument");
        USPMV_CUDA(cudaSetDevice(ctx->device));
        cudaStream_t st;
        USPMV_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        *out = st;
    });
}
int uspmv_stream_destroy(uspmv_ctx *ctx, void *stream) {
    return guarded([&] {
        if (!ctx) fail("uspmv_stream_destroy: NULL argument");
        if (stream) USPMV_CUDA(cudaStreamDestroy(static_cast<cudaStream_t>(stream)));
    });
}

int uspmv_device_count(int *out) {
    return guarded([&] {
        if (!out) fail("uspmv_device_count: out is NULL");
        int ndev = 0;
        cudaError_t e = cudaGetDeviceCount(&ndev);
        if (e != cudaSuccess || ndev == 0)
            fail("uspmv_device_count: no CUDA device available (%s); this library has no CPU fallback",
                 e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
        *out = ndev;
    });
}

int uspmv_ctx_create(int device, uspmv_ctx **out) {
    return guarded([&] {
        if (!out) fail("uspmv_ctx_create: out is NULL");
        int ndev = 0;
        cudaError_t e = cudaGetDeviceCount(&ndev);
        if (e != cudaSuccess || ndev == 0)
            fail("uspmv_ctx_create: no CUDA device available (%s); this library has no CPU fallback",
                 e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
        if (device < 0 || device >= ndev) fail("uspmv_ctx_create: device %d out of range [0,%d)", device, ndev);
        USPMV_CUDA(cudaSetDevice(device));
        cudaDeviceProp prop;
        USPMV_CUDA(cudaGetDeviceProperties(&prop, device));
        if (prop.major != 10)
            fail("uspmv_ctx_create: device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, prop.major,
                 prop.minor);
        if (default_options().l2_fetch_granularity)
            USPMV_CUDA(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)default_options().l2_fetch_granularity));
        auto *ctx = new uspmv_ctx();
        ctx->device = device;
        ctx->n_sm = prop.multiProcessorCount;
        ctx->opts = default_options();
        {
            std::lock_guard<std::mutex> g(g_ctx_mutex);
            g_live_ctxs.push_back(ctx);
        }
        *out = ctx;
    });
}

void uspmv_ctx_destroy(uspmv_ctx *ctx) {
    if (!ctx) return;
    {
        std::lock_guard<std::mutex> g(g_ctx_mutex);
        g_live_ctxs.erase(std::remove(g_live_ctxs.begin(), g_live_ctxs.end(), ctx), g_live_ctxs.end());
    }
    delete ctx;
}

int uspmv_ctx_sync(uspmv_ctx *ctx) {
    return guarded([&] {
        if (!ctx) fail("uspmv_ctx_sync: ctx is NULL");
        USPMV_CUDA(cudaSetDevice(ctx->device));
        USPMV_CUDA(cudaDeviceSynchronize());
    });
}

int uspmv_malloc(uspmv_ctx *ctx, size_t bytes, void **out_d) {
    return guarded([&] {
        if (!ctx || !out_d) fail("uspmv_malloc: NULL argument");
        USPMV_CUDA(cudaSetDevice(ctx->device));
        *out_d = nullptr;
        if (bytes) USPMV_CUDA(cudaMalloc(out_d, bytes));
    });
}

int uspmv_free(uspmv_ctx *ctx, void *ptr_d) {
    return guarded([&] {
        if (!ctx) fail("uspmv_free: ctx is NULL");
        if (ptr_d) USPMV_CUDA(cudaFree(ptr_d));
    });
}

int uspmv_memcpy_h2d(uspmv_ctx *ctx, void *dst_d, const void *src_h, size_t bytes, void *stream) {
    return guarded([&] {
        if (!ctx) fail("uspmv_memcpy_h2d: ctx is NULL");
        USPMV_CUDA(cudaMemcpyAsync(dst_d, src_h, bytes, cudaMemcpyHostToDevice, as_stream(stream)));
        USPMV_CUDA(cudaStreamSynchronize(as_stream(stream)));
    });
}

int uspmv_memcpy_d2h(uspmv_ctx *ctx, void *dst_h, const void *src_d, size_t bytes, void *stream) {
    return guarded([&] {
        if (!ctx) fail("uspmv_memcpy_d2h: ctx is NULL");
        USPMV_CUDA(cudaMemcpyAsync(dst_h, src_d, bytes, cudaMemcpyDeviceToHost, as_stream(stream)));
        USPMV_CUDA(cudaStreamSynchronize(as_stream(stream)));
    });
}

int uspmv_memset(uspmv_ctx *ctx, void *dst_d, int byte, size_t bytes, void *stream) {
    return guarded([&] {
        if (!ctx) fail("uspmv_memset: ctx is NULL");
        USPMV_CUDA(cudaMemsetAsync(dst_d, byte, bytes, as_stream(stream)));
    });
}

/* 1 if ptr is device (or managed) memory of some CUDA device, 0 for ordinary / pinned host memory and NULL */
int uspmv_pointer_is_device(const void *ptr, int *out) {
    return guarded([&] {
        if (!out) fail("uspmv_pointer_is_device: out is NULL");
        *out = 0;
        if (!ptr) return;
        cudaPointerAttributes a;
        cudaError_t e = cudaPointerGetAttributes(&a, ptr);
        if (e != cudaSuccess) { cudaGetLastError(); return; }
        *out = (a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged) ? 1 : 0;
    });
}

/* random_init (utilities.hpp:880-912) + the padding rule of init_std_vec_with_ptr_or_value (:914-981), HOST logic: a DEFAULT-seeded
 * std::mt19937 and uniform_real_distribution<double>(min, max) drawn sequentially over all n values, each converted to the vector's
 * type; then, per vector of vec_length values (colwise) / over the first n_rows * bvs values (rowwise), everything past the real rows
 * is zeroed.  libstdc++'s distribution draws two 32-bit words per double: u = (a0 + a1 2^32) / 2^64 rounded to double (1.0 is replaced
 * by the largest double below it), value = (max - min) * u + min. */
int uspmv_random_init_host(double vmin, double vmax, long n, int vt, void *out_h, long n_rows, long vec_length, int bvs, int layout) {
    return guarded([&] {
        if (n < 0 || (n > 0 && !out_h)) fail("uspmv_random_init_host: NULL output");
        vt_size(vt);
        // mt19937 (32-bit Mersenne twister, seed 5489) restated: the standard fixes its output sequence
        unsigned int mt[624];
        mt[0] = 5489u;
        for (int i = 1; i < 624; ++i) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (unsigned int)i;
        int idx = 624;
        auto next = [&]() -> unsigned int {
            if (idx >= 624) {
                for (int k = 0; k < 624; ++k) {
                    const unsigned int yv = (mt[k] & 0x80000000u) | (mt[(k + 1) % 624] & 0x7fffffffu);
                    mt[k] = mt[(k + 397) % 624] ^ (yv >> 1) ^ ((yv & 1u) ? 0x9908b0dfu : 0u);
                }
                idx = 0;
            }
            unsigned int yv = mt[idx++];
            yv ^= yv >> 11;
            yv ^= (yv << 7) & 0x9d2c5680u;
            yv ^= (yv << 15) & 0xefc60000u;
            yv ^= yv >> 18;
            return yv;
        };
        for (long i = 0; i < n; ++i) {
            const unsigned long long a0 = next(), a1 = next();
            const long double sum = (long double)a0 + (long double)a1 * 4294967296.0L;
            double u = (double)(sum / 18446744073709551616.0L);
            if (u >= 1.0) u = 0.99999999999999988897769753748434595763683319091796875;  // nextafter(1, 0)
            // libstdc++: `return (__aurng() * (__p.b() - __p.a())) + __p.a();` — g++ -O3 on an FMA machine (the reference builds with
            // -march=native) contracts it into one fused multiply-add, so the reference's x is the FMA result
            const double v = std::fma(u, vmax - vmin, vmin);
            switch (vt) {
            case USPMV_F64: static_cast<double *>(out_h)[i] = v; break;
            case USPMV_F32: static_cast<float *>(out_h)[i] = (float)v; break;
            default: {
                // double -> fp16, one rounding (static_cast<_Float16>(double))
                const __half h = __double2half(v);
                static_cast<__half *>(out_h)[i] = h;
            }
            }
        }
        if (n_rows < 0 || vec_length <= 0) return;  // no padding rule requested
        auto zero = [&](long i) {
            switch (vt) {
            case USPMV_F64: static_cast<double *>(out_h)[i] = 0.0; break;
            case USPMV_F32: static_cast<float *>(out_h)[i] = 0.0f; break;
            default: static_cast<unsigned short *>(out_h)[i] = 0;
            }
        };
        if (layout == USPMV_ROWWISE) {
            for (long i = n_rows * (long)bvs; i < n; ++i) zero(i);
        } else {
            for (long i = 0; i < n; ++i)
                if (i % vec_length >= n_rows) zero(i);
        }
    });
}

int uspmv_host_alloc(size_t bytes, void **out_h) {
    return guarded([&] {
        if (!out_h) fail("uspmv_host_alloc: out is NULL");
        USPMV_CUDA(cudaMallocHost(out_h, bytes ? bytes : 1));
    });
}

int uspmv_host_free(void *ptr_h) {
    return guarded([&] {
        if (ptr_h) USPMV_CUDA(cudaFreeHost(ptr_h));
    });
}

}  // extern "C"
