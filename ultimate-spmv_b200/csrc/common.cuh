// Shared internals of libuspmv_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>

#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <memory>
#include <mutex>
#include <string>

#include "../../include/uspmv_b200.h"

namespace uspmv {

// ---- error plumbing: C++ exceptions inside, int + message at the C boundary ---------------------
void set_error(const std::string &msg);
extern std::atomic<long> g_launches;

struct Error : std::runtime_error {
    using std::runtime_error::runtime_error;
};

[[noreturn]] inline void fail(const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    throw Error(buf);
}

#define USPMV_CUDA(call)                                                                          \
    do {                                                                                          \
        cudaError_t e__ = (call);                                                                 \
        if (e__ != cudaSuccess)                                                                   \
            ::uspmv::fail("CUDA error %s at %s:%d: %s", cudaGetErrorName(e__), __FILE__, __LINE__, \
                          cudaGetErrorString(e__));                                               \
    } while (0)

// every kernel launch goes through this so launches are counted and launch errors are not lost
#define USPMV_LAUNCH_CHECK()                                  \
    do {                                                      \
        ::uspmv::g_launches.fetch_add(1, std::memory_order_relaxed); \
        USPMV_CUDA(cudaGetLastError());                       \
    } while (0)

template <typename F>
inline int guarded(F &&f) noexcept {
    try {
        f();
        return 0;
    } catch (const std::exception &e) {
        set_error(e.what());
        return 1;
    } catch (...) {
        set_error("unknown error");
        return 2;
    }
}

inline size_t vt_size(int vt) {
    switch (vt) {
    case USPMV_F64: return 8;
    case USPMV_F32: return 4;
    case USPMV_F16: return 2;
    }
    fail("invalid value type %d", vt);
}

inline cudaStream_t as_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }

// RAII device buffer
template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    DevBuf() = default;
    explicit DevBuf(size_t count) { alloc(count); }
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    DevBuf(DevBuf &&o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
    DevBuf &operator=(DevBuf &&o) noexcept {
        if (this != &o) { release(); p = o.p; n = o.n; o.p = nullptr; o.n = 0; }
        return *this;
    }
    ~DevBuf() { release(); }
    void alloc(size_t count) {
        release();
        n = count;
        if (count) USPMV_CUDA(cudaMalloc(&p, count * sizeof(T)));
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
    }
};

// run-time knobs (uspmv_set_option / environment), defined in context.cu
struct Options {
    bool scs_stream = true;      // C = 32: bulk-copy streamed kernel (false: direct-load kernel)
    bool scs_stream_wide = true; // C = 64 / 128: wide-chunk streamed kernel (false: direct-load kernel)
    bool narrow_dp = true;       // C = 16 in fp64 through the narrow-chunk streamed kernel (12 warps per CTA)
    bool fused_narrow = false;       // distributed SELL-32 SpMV in sp / hp through the fused kernel too (default: multi-kernel overlap, see halo.cu)
    bool mmv_push_first = true;      // distributed SpMMV, multi-kernel overlap: launch the push before the interior kernel (see halo.cu)
    int mmv_fused_rowwise = 1;       // distributed SpMMV with row-major block vectors: 1 fused one-kernel step (default), 0 push / wait kernels next to the
                                     // interior kernel, 2 fused even without a neighbour (profiling the instance)
    bool pair_hp = false;        // C = 32 in fp16: two adjacent chunks per work item (k_scs32_stream_pair) — bit-identical, measured
                                 // no faster than the one-chunk kernel (7-pt 256^3: 154 vs 149 us, profiles/r02l_*), so off by default
    int stream_variant = 0;      // (slots per piece, ring depth, warps per CTA) instantiation
    int stream_blocks_per_sm = 2;
    int mmv_variant = 0;         // SpMMV streamed kernel: 0 = tuned default, 1..4 force a variant (see spmv_kernels.cu)
    int mmv_blocks_per_sm = 0;   // SpMMV streamed kernel: CTAs (8 warps) per SM, 0 = as many as fit
    int ap_variant = 0;          // fused adaptive-precision streamed kernel: ring depth / register cap instantiation
    int split_long_chunks = 256; // C = 32, uneven matrices: chunks longer than this many slots are summed in segments (0 = never)
    bool strict_reference_halo = false;  // true: padding slots (column 0) become a halo element on ranks > 0, like the reference
    int push_variant = -1;       // large single-vector halos: -1 = tuned default, 0 = one store per thread, 1 = tiles of 2048 elements,
                                 // direct peer stores, 2 = tiles gathered into shared memory and written to the peer by the bulk-copy
                                 // engine, 3 = like 1 with tiles of 4096 elements
    int push_ctas_per_sm = 0;    // CTAs per SM of the large-halo push kernels (0 = default)
    int l2_fetch_granularity = 0; // cudaLimitMaxL2FetchGranularity (32 / 64 / 128 bytes; 0 = leave the driver default)
    long push_min_elements = 1L << 20;  // halos with at least this many elements to send use the tiled push kernels
};
// Options are per CONTEXT (two contexts in one process can differ): every context owns a copy, taken from the process defaults when it
// is created.  uspmv_set_option changes the defaults AND every live context (the process-wide knob tests and benchmarks use),
// uspmv_ctx_set_option one context only.  Inside the library `options()` is the option block of the context whose entry point is
// executing on this thread (OptScope at the top of every entry point that takes a handle), else the process defaults.
Options &default_options();
const Options &options();
struct OptScope {
    const Options *prev;
    explicit OptScope(const uspmv_ctx *ctx);
    ~OptScope();
    OptScope(const OptScope &) = delete;
    OptScope &operator=(const OptScope &) = delete;
};

// per-device once-flags of the launchers (cudaFuncSetAttribute and occupancy are per device; a process may hold contexts on several)
constexpr int MAX_DEVICES = 64;
inline int current_device() {
    int dev = 0;
    USPMV_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= MAX_DEVICES) fail("device ordinal %d not supported (max %d)", dev, MAX_DEVICES);
    return dev;
}
// compute entry points run on the context's device whatever the caller's current device is
inline void use_device(const uspmv_ctx *ctx);

inline int sm_count(int device) {
    static int cached[64] = {0};
    if (device >= 0 && device < 64 && cached[device]) return cached[device];
    int n = 0;
    USPMV_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device));
    if (device >= 0 && device < 64) cached[device] = n;
    return n;
}

}  // namespace uspmv

// ---- opaque handle definitions ------------------------------------------------------------------
struct uspmv_ctx {
    int device = 0;
    int n_sm = 0;
    uspmv::Options opts;  // this context's kernel-selection knobs
};
namespace uspmv {
inline void use_device(const uspmv_ctx *ctx) {
    int dev = -1;
    USPMV_CUDA(cudaGetDevice(&dev));
    if (dev != ctx->device) USPMV_CUDA(cudaSetDevice(ctx->device));
}
inline unsigned long long next_generation() {
    static std::atomic<unsigned long long> g{1};
    return g.fetch_add(1);
}
}  // namespace uspmv

struct uspmv_coo {
    uspmv_ctx *ctx = nullptr;
    long n_rows = 0, n_cols = 0, nnz = 0;
    int mt = USPMV_F64;
    uspmv::DevBuf<int> I, J;
    uspmv::DevBuf<unsigned char> values;  // nnz * vt_size(mt) bytes
};

struct uspmv_scs {
    uspmv_ctx *ctx = nullptr;
    const unsigned long long generation = uspmv::next_generation();  // identity of this handle (an address can be reused after destroy)
    long x_min_len = 0;  // smallest x the kernels may be given: n_cols as built, n_local + n_halo after halo renumbering
    long C = 1, sigma = 1, n_rows = 0, n_cols = 0, n_rows_padded = 0, n_chunks = 0, n_elements = 0, nnz = 0;
    int vt = USPMV_F64;
    bool cols_permuted = false;
    uspmv::DevBuf<int> chunk_ptrs;     // n_chunks + 1
    uspmv::DevBuf<int> chunk_lengths;  // n_chunks
    uspmv::DevBuf<int> col_idxs;       // n_elements
    uspmv::DevBuf<unsigned char> values;
    uspmv::DevBuf<int> old_to_new;  // n_rows
    uspmv::DevBuf<int> new_to_old;  // n_rows_padded, -1 where no real row lands
    uspmv::DevBuf<int> row_lengths; // n_rows_padded: stored elements of the row at each (permuted) position
    uspmv::DevBuf<unsigned char> h2d_stage_x, d2h_stage_y;  // device staging for the host-buffer call
    // pipelined host-buffer calls: per slot a device x / y pair, three streams and events
    static constexpr int HOST_SLOTS = 3;
    uspmv::DevBuf<unsigned char> slot_x[HOST_SLOTS], slot_y[HOST_SLOTS];
    cudaStream_t s_h2d = nullptr, s_run = nullptr, s_d2h = nullptr;
    cudaEvent_t ev_x[HOST_SLOTS] = {}, ev_y[HOST_SLOTS] = {}, ev_done[HOST_SLOTS] = {};
    bool slot_busy[HOST_SLOTS] = {};
    ~uspmv_scs() {
        for (int k = 0; k < HOST_SLOTS; ++k) {
            if (ev_x[k]) cudaEventDestroy(ev_x[k]);
            if (ev_y[k]) cudaEventDestroy(ev_y[k]);
            if (ev_done[k]) cudaEventDestroy(ev_done[k]);
        }
        if (s_h2d) cudaStreamDestroy(s_h2d);
        if (s_run) cudaStreamDestroy(s_run);
        if (s_d2h) cudaStreamDestroy(s_d2h);
    }
    // chunk ids sorted by length (longest first, ties in chunk order); only built when lengths are very uneven, so that
    // one-warp-per-chunk kernels stay load balanced (longest-processing-time-first over the persistent warps)
    uspmv::DevBuf<int> balanced_order;
    // C = 32 only: work items {chunk, first slot, slots, partial slot} with long chunks cut into segments (see scs_stream.cuh)
    uspmv::DevBuf<int4> vitems;
    uspmv::DevBuf<int> split_chunk, split_ptr;
    uspmv::DevBuf<unsigned char> partials;
    long n_vitems = 0, n_split = 0;
    bool chunks_split = false;
    // adaptive precision, C = 32, very uneven matrices: work items of the fused kernel (built lazily by the first uspmv_ap_spmv
    // call on this part as the FIRST part, keyed on the other parts), see k_scs32_stream_ap
    struct ApPlan {
        unsigned long long other[2] = {0, 0};  // generation ids of the other parts
        long other_ne[2] = {-1, -1};
        int mode = -1, seg_slots = 0;
        bool use = false;
        long n_items = 0, n_split = 0;
        uspmv::DevBuf<int4> items;
        uspmv::DevBuf<int> split_chunk, split_ptr;
        uspmv::DevBuf<unsigned char> seg_part;
        uspmv::DevBuf<double> partials;
    };
    mutable std::unique_ptr<ApPlan> ap_plan;
    mutable std::mutex ap_plan_mutex;
    uspmv::DevBuf<int> interior_chunks, boundary_chunks;  // chunk ids without / with halo columns (order kept)
    bool interior_contig = false, boundary_contig = false;
    int interior_off = 0, boundary_off = 0;
};

// the context behind any handle (null-safe), for OptScope
inline const uspmv_ctx *ctx_of(const uspmv_ctx *c) { return c; }
inline const uspmv_ctx *ctx_of(const uspmv_scs *s) { return s ? s->ctx : nullptr; }
inline const uspmv_ctx *ctx_of(const uspmv_coo *c) { return c ? c->ctx : nullptr; }
const uspmv_ctx *ctx_of(const uspmv_halo *h);    // halo.cu
const uspmv_ctx *ctx_of(const uspmv_p2p *p);     // halo.cu
struct uspmv_banded;
const uspmv_ctx *ctx_of(const uspmv_banded *b);  // column_bands.cu

