// SELL-32-sigma SpMV for sm_100a with the matrix streamed through shared memory by the bulk-copy (TMA) engine.
//
// Why: a thread-per-row kernel that loads values/columns straight into registers is latency-bound on B200
// (ncu r01a: 45 % occupancy, 90 % "no eligible warp", 2.4 TB/s) because the bytes in flight are tied to
// registers x resident threads and every thread walks a 3-deep dependent chain (chunk meta -> col/val -> x).
// Here each warp owns a private ring of D shared-memory stages that lane 0 keeps full with
// `cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes` copies (SASS: UBLKCP), so the matrix
// stream (the 84 % of the traffic that is perfectly contiguous: a chunk's values and column indices are
// `len*32` consecutive elements) is decoupled from the register file; the lanes only do shared-memory reads,
// the x gathers and the fused multiply-adds.  In-flight bytes per SM = warps x D x piece size (~170 KB).
//
// Work split: warp gw takes chunks gw, gw + W, gw + 2W, ... (W = warps in the grid), so all SMs advance as one
// front over the rows and the x lines shared by neighbouring chunks are hit in L1/L2 (measured DRAM traffic ==
// algorithmic bytes).  A chunk of length len is cut into pieces of <= LMAX slots (contiguous because SELL
// chunks are slot-major); piece headers {ns, chunk, first/last} live next to the stage.  Per-row accumulation
// order is j = 0..len-1 with one FMA per slot: bit-identical to the direct kernel and to the oracle.
#pragma once
#include <cstdint>
#include <type_traits>

namespace uspmv {
namespace stream {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
// global -> shared bulk copy (bytes % 16 == 0, both addresses 16 B aligned), completion on an mbarrier
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(dst_smem)),
        "l"(__cvta_generic_to_global(src_gmem)), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}

struct PieceHdr {
    int ns;     // slots in this piece (0 with flags != 0: empty chunk; flags == 0 && ns == 0: end of stream)
    int flags;  // bit0 first piece of the chunk, bit1 last piece, bit2 valid
    int chunk;
    int pad;
};

template <typename VT, int LMAX, int D>
struct WarpRing {
    static constexpr int STAGE_ELEMS = LMAX * 32;
    static constexpr int VAL_BYTES = STAGE_ELEMS * (int)sizeof(VT);
    static constexpr int COL_BYTES = STAGE_ELEMS * 4;
    static constexpr int STAGE_BYTES = VAL_BYTES + COL_BYTES;
    static constexpr int BYTES = D * STAGE_BYTES + D * 8 + D * (int)sizeof(PieceHdr);  // stages + mbarriers + headers
    static constexpr int BYTES_ALIGNED = (BYTES + 127) / 128 * 128;
};

// ---- fused halo exchange (one process per GPU, NVLink peer memory) -------------------------------------------
// With FUSED the SAME kernel (a) gathers this rank's send elements and stores them straight into the neighbours'
// x tails, then raises their `arrived` epoch flags, (b) runs the interior chunks, (c) waits for its own `arrived`
// flags and runs the boundary chunks (the only ones that read halo columns), (d) the last warp acknowledges
// consumption to the senders and bumps the epoch.  One launch per distributed SpMV, transfer overlapped with
// the interior math, no collective call and no host round trip.  Flags live in the IPC arena (halo.cu).
struct FusedArgs {
    long n_int, n_bnd;            // interior items first, then boundary items
    const int *int_list, *bnd_list;
    int int_off, bnd_off;
    int P, my_rank;
    long n_send;
    const int *send_ptr, *is_receiver, *is_sender, *send_idx, *perm;
    const unsigned long long *peer_x0, *peer_arrived, *peer_acked;  // peer_x0[q]: start of q's (block) vector
    const long *peer_base, *peer_ld;  // per peer q: row offset of OUR rows in q's vector (q.n_local + q.recv_cumsum[me]); q's vec_length
    int bvs, layout;              // block vectors: all bvs values of a halo row travel in the same push (bulkvec, classes_structs.hpp:909-970)
    long ld;                      // own vec_length (column-major block vectors)
    unsigned int *acked, *arrived, *epoch, *error, *counters;  // counters[0] push warps done, [1] warps finished
    int y_rows;                   // rows >= y_rows are not stored (solve loop: y is the other x buffer, see uspmv_p2p_spmv_buf)
};

__device__ __forceinline__ unsigned int ld_flag(const unsigned int *p) {
    unsigned int v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long gtimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// lane p (< P, selected) spins until flags[p] >= target; bounded (10 s) so that a lost peer cannot hang the GPU
__device__ __forceinline__ void warp_wait_flags(const unsigned int *flags, const int *select, int P, unsigned int target, int lane,
                                                unsigned int *error) {
    if (lane < P && select[lane]) {
        const unsigned long long t0 = gtimer_ns();
        while (ld_flag(flags + lane) < target) {
            if (gtimer_ns() - t0 > 10000000000ull) {
                atomicExch(error, 1u);
                break;
            }
            __nanosleep(100);
        }
    }
    __syncwarp();
}
template <typename T> __device__ __forceinline__ T ld_x_coherent(const T *p) { return __ldcg(p); }  // L2 (coherent with peer stores)

// (a) of the fused step: warp pw handles elements [32 pw, 32 pw + 32) of this rank's send list (x bvs for block vectors: row-major
// element (i, v) of peer q lands at q.X[(base_q + i) * bvs + v], column-major at q.X[base_q + i + v * ld_q]); the last push warp of the
// grid raises the neighbours' `arrived` flags.  ROWWISE_LAYOUT value = 1 (USPMV_ROWWISE).
// Stores only: returns the number of 32-unit groups this warp pushed (0: it took no part).  The caller must follow up with
// fused_push_signal — right away (fused_push) or after some interior work, when the stores have long been acknowledged and the
// system-scope fence is cheap (SELL-32 SpMV kernel).
// work units of the push.  Single vectors and column-major block vectors: one element.  Row-major block vectors: one SEGMENT of a
// halo row — 16 bytes when the row is a multiple of 16 bytes (rows are then 16-byte aligned on both sides, the buffers being
// 256-byte aligned), else one element; consecutive units are consecutive in the PEER's vector, so a warp-wide store covers 512
// contiguous bytes of NVLink traffic (a lane-per-row copy wrote 16-byte pieces a row apart and gave a quarter of the warps the whole
// push to do; profiles/r02u_dist_probe_mmv.txt).
template <typename VT>
__device__ __forceinline__ int fused_push_seg_bytes(const FusedArgs &fa) {
    const int row_bytes = fa.bvs * (int)sizeof(VT);
    return (fa.bvs > 1 && fa.layout == 1 && row_bytes % 16 == 0) ? 16 : (int)sizeof(VT);
}
template <typename VT>
__device__ __forceinline__ long fused_push_units(const FusedArgs &fa) {
    return fa.n_send * (long)(fa.bvs * (int)sizeof(VT) / fused_push_seg_bytes<VT>(fa));
}

template <typename VT>
__device__ __forceinline__ unsigned int fused_push_stores(const FusedArgs &fa, const VT *__restrict__ x, const long gw, const long W, const int lane,
                                                          const unsigned int epoch_e) {
    const bool by_row = fa.bvs > 1 && fa.layout == 1;
    const int seg = fused_push_seg_bytes<VT>(fa);
    const int spr = fa.bvs * (int)sizeof(VT) / seg;  // units per halo row (1 for single vectors)
    const long units = fa.n_send * (long)spr;
    const long n_push_warps = (units + 31) / 32;
    if (gw >= n_push_warps) return 0;
    warp_wait_flags(fa.acked, fa.is_receiver, fa.P, epoch_e - 1u, lane, fa.error);  // receivers consumed step e-1
    unsigned int mine = 0;  // 32-unit groups this warp pushed
    for (long pw = gw; pw < n_push_warps; pw += W, ++mine) {
        const long t = pw * 32 + lane;
        if (t < units) {
            long i, v;  // halo row, unit within the row (row-major) / vector (column-major)
            if (by_row) { i = t / spr; v = t - i * spr; }
            else { v = t / fa.n_send; i = t - v * fa.n_send; }
            int q = 0;
            while (i >= fa.send_ptr[q + 1]) ++q;
            const long src = fa.perm ? fa.perm[fa.send_idx[i]] : fa.send_idx[i];
            const long k = fa.peer_base[q] + (i - fa.send_ptr[q]);
            VT *dst = reinterpret_cast<VT *>(fa.peer_x0[q]);
            if (!by_row) dst[k + v * fa.peer_ld[q]] = x[src + v * fa.ld];  // bvs == 1: v == 0
            else if (seg == 16) reinterpret_cast<int4 *>(dst + k * fa.bvs)[v] = reinterpret_cast<const int4 *>(x + src * fa.bvs)[v];
            else dst[k * fa.bvs + v] = x[src * fa.bvs + v];
        }
    }
    return mine;
}

// ONE system-scope fence per warp (every push of this warp is ordered before its counter update), then the warp that completes the
// count raises the neighbours' `arrived` flags
__device__ __forceinline__ void fused_push_signal(const FusedArgs &fa, const long units, const unsigned int mine, const int lane,
                                                  const unsigned int epoch_e) {
    const unsigned int n_push_warps = (unsigned int)((units + 31) / 32);
    __threadfence_system();
    __syncwarp();
    unsigned int last = 0;
    if (lane == 0) last = (atomicAdd(&fa.counters[0], mine) + mine == n_push_warps);
    last = __shfl_sync(0xffffffffu, last, 0);
    if (last) {
        __threadfence_system();
        if (lane < fa.P && fa.is_receiver[lane]) {
            volatile unsigned int *f = reinterpret_cast<volatile unsigned int *>(fa.peer_arrived[lane]);
            *f = epoch_e;
        }
        __threadfence_system();
    }
}

template <typename VT>
__device__ __forceinline__ void fused_push(const FusedArgs &fa, const VT *__restrict__ x, const long gw, const long W, const int lane,
                                           const unsigned int epoch_e) {
    const unsigned int mine = fused_push_stores<VT>(fa, x, gw, W, lane, epoch_e);
    if (mine) fused_push_signal(fa, fused_push_units<VT>(fa), mine, lane, epoch_e);
}

// (d) of the fused step: the last warp of the grid acknowledges consumption to the senders and closes the epoch
__device__ __forceinline__ void fused_finish(const FusedArgs &fa, const long W, const int lane, const unsigned int epoch_e) {
    __syncwarp();
    unsigned int last = 0;
    if (lane == 0) {
        __threadfence();
        last = (atomicAdd(&fa.counters[1], 1u) == (unsigned int)(W - 1));
    }
    last = __shfl_sync(0xffffffffu, last, 0);
    if (last) {
        if (fa.n_bnd == 0) warp_wait_flags(fa.arrived, fa.is_sender, fa.P, epoch_e, lane, fa.error);  // keep epochs in step
        if (lane < fa.P && fa.is_sender[lane]) {
            volatile unsigned int *f = reinterpret_cast<volatile unsigned int *>(fa.peer_acked[lane]);
            *f = epoch_e;
        }
        __threadfence_system();
        if (lane == 0) {
            fa.counters[0] = 0;
            fa.counters[1] = 0;
            *reinterpret_cast<volatile unsigned int *>(fa.epoch) = epoch_e;
            __threadfence();
        }
    }
}

// ---- per-piece arithmetic ("bodies") ----------------------------------------------------------------------------
// A body owns the accumulators of the 32 rows of the current chunk (one row per lane):
//   begin_chunk()                       reset
//   piece(ns, sv, sc)                   consume ns slots: sv/sc point at this lane's value / column of slot 0 (stride 32)
//   end_chunk(chunk)                    write the rows of `chunk`
// SpMV: y = A x.  COHERENT: x through L2 only (ld.global.cg) because a peer GPU wrote part of it during this kernel.
template <typename VT, typename A, int LMAX, bool UNPERM, bool COHERENT, bool BOUNDED = false>
struct SpmvBody {
    const VT *__restrict__ x;
    VT *__restrict__ y;
    const int *__restrict__ new_to_old;
    int lane;
    typename A::acc_t acc;
    int y_rows = 0;  // BOUNDED: positions >= y_rows are not stored (y aliases the next x, whose tail holds the halo)
    __device__ __forceinline__ void begin_chunk(int = 0, int = 0) { acc = A::zero(); }
    // Slot count at run time (predicated loads): measured FASTER than a switch over template-constant slot counts for the narrow
    // types on B200 (7-point 256^3, profiles/r02d_ab_stream.log: sp 162 vs 194 us, hp 150 vs 172 us, dp equal) although it executes
    // more instructions — ncu (profiles/r02e_*): the kernel is latency-, not issue-limited once the predication is gone
    // (IPC 3.0 -> 2.1, "no eligible" 24 % -> 49 % for hp), so the shorter instruction stream only exposes the gather latency.
    __device__ __forceinline__ void piece(const int ns, const VT *sv, const int *sc) {
        VT v[LMAX], xv[LMAX];
        int col[LMAX];
#pragma unroll
        for (int j = 0; j < LMAX; ++j)
            if (j < ns) col[j] = sc[j * 32];
#pragma unroll
        for (int j = 0; j < LMAX; ++j)
            if (j < ns) xv[j] = COHERENT ? __ldcg(x + col[j]) : __ldg(x + col[j]);
#pragma unroll
        for (int j = 0; j < LMAX; ++j)
            if (j < ns) v[j] = sv[j * 32];
#pragma unroll
        for (int j = 0; j < LMAX; ++j)
            if (j < ns) acc = A::mad(v[j], xv[j], acc);
    }
    __device__ __forceinline__ void end_chunk(const int chunk) {
        const long row = (long)chunk * 32 + lane;
        if (UNPERM) {
            const int o = new_to_old[row];
            if (o >= 0) y[o] = A::out(acc);
        } else if (!BOUNDED || chunk * 32 + lane < y_rows)
            y[row] = A::out(acc);
    }
};

// Body of the fused distributed SELL-32 SpMV (k_scs32_stream<..., FUSED>).  Interior and boundary chunks run through ONE stream
// (boundary items follow the interior ones in the item list and carry header flag 8): the bulk copies of the first boundary chunks
// are in flight while the last interior chunks are summed, instead of a cold ring restart after the flag wait.  The first boundary
// chunk of a warp waits for the neighbours' `arrived` flags; boundary pieces gather x through L2 (ld.global.cg: a peer wrote the halo
// during this kernel).  The warp's push is SIGNALLED (fence.sys + counter) after its first chunk, when its peer stores have long been
// acknowledged — or before it waits for anybody else's, whichever comes first.
template <typename VT, typename A, int LMAX>
struct SpmvBodyFusedStream {
    const VT *__restrict__ x;
    VT *__restrict__ y;
    int lane, y_rows;
    const FusedArgs *fa;
    unsigned int epoch_e, mine;
    bool sig_pending, halo_ready, bnd;
    typename A::acc_t acc;
    __device__ __forceinline__ void signal() {
        fused_push_signal(*fa, fused_push_units<VT>(*fa), mine, lane, epoch_e);
        sig_pending = false;
    }
    __device__ __forceinline__ void begin_chunk(int = 0, const int flags = 0) {
        acc = A::zero();
        bnd = (flags & 8) != 0;
        if (bnd && !halo_ready) {
            if (sig_pending) signal();  // never wait for a neighbour while one's own push is unpublished
            warp_wait_flags(fa->arrived, fa->is_sender, fa->P, epoch_e, lane, fa->error);
            halo_ready = true;
        }
    }
    __device__ __forceinline__ void piece(const int ns, const VT *sv, const int *sc) {
        VT xv[LMAX];
        int col[LMAX];
#pragma unroll
        for (int j = 0; j < LMAX; ++j)
            if (j < ns) col[j] = sc[j * 32];
        if (bnd) {  // warp-uniform: only the gathers differ between interior and boundary pieces
#pragma unroll
            for (int j = 0; j < LMAX; ++j)
                if (j < ns) xv[j] = __ldcg(x + col[j]);
        } else {
#pragma unroll
            for (int j = 0; j < LMAX; ++j)
                if (j < ns) xv[j] = __ldg(x + col[j]);
        }
#pragma unroll
        for (int j = 0; j < LMAX; ++j)
            if (j < ns) acc = A::mad(sv[j * 32], xv[j], acc);  // values read from the stage right before use (64-register budget)
    }
    __device__ __forceinline__ void end_chunk(const int chunk) {
        if (chunk * 32 + lane < y_rows) y[(long)chunk * 32 + lane] = A::out(acc);
        if (sig_pending) signal();
    }
};

// SpMMV: Y = A X with BVS right-hand sides.  ROWWISE: X[col*BVS + v] (one vector load per gathered row); else X[col + v*ld].
// Slots are consumed SB at a time so that SB*BVS gathered values are in flight per lane without blowing the register file.
template <typename VT, typename A, int LMAX, int BVS, bool ROWWISE, bool COHERENT = false>
struct SpmmvBody {
    static constexpr int SB = (BVS * (int)sizeof(VT) >= 64) ? 2 : (BVS * (int)sizeof(VT) >= 32 ? 4 : 8);
    const VT *__restrict__ X;
    VT *__restrict__ Y;
    long ld;
    int lane;
    typename A::acc_t acc[BVS];
    __device__ __forceinline__ void begin_chunk(int = 0, int = 0) {
#pragma unroll
        for (int v = 0; v < BVS; ++v) acc[v] = A::zero();
    }
    __device__ __forceinline__ void load_row(const long col, VT *xv) const {
        if constexpr (ROWWISE) {
            constexpr int BYTES = BVS * (int)sizeof(VT);
            if constexpr (BYTES % 16 == 0) {
                const int4 *p = reinterpret_cast<const int4 *>(X + col * BVS);
#pragma unroll
                for (int k = 0; k < BYTES / 16; ++k) reinterpret_cast<int4 *>(xv)[k] = COHERENT ? __ldcg(p + k) : __ldg(p + k);
            } else if constexpr (BYTES % 8 == 0) {
                const int2 *p = reinterpret_cast<const int2 *>(X + col * BVS);
#pragma unroll
                for (int k = 0; k < BYTES / 8; ++k) reinterpret_cast<int2 *>(xv)[k] = COHERENT ? __ldcg(p + k) : __ldg(p + k);
            } else {
#pragma unroll
                for (int v = 0; v < BVS; ++v) xv[v] = COHERENT ? __ldcg(X + col * BVS + v) : __ldg(X + col * BVS + v);
            }
        } else {
#pragma unroll
            for (int v = 0; v < BVS; ++v) xv[v] = COHERENT ? __ldcg(X + col + v * ld) : __ldg(X + col + v * ld);
        }
    }
    // The slot count is a template constant: with a run-time count every gather became a predicated load into a scratch register
    // followed by a predicated move (ncu r01e: 40 % of the stall samples sat on those moves), which serialises the loads.
    template <int NS>
    __device__ __forceinline__ void piece_n(const VT *sv, const int *sc) {
#pragma unroll
        for (int j0 = 0; j0 < NS; j0 += SB) {
            alignas(16) VT xv[SB][BVS];
            VT v[SB];
#pragma unroll
            for (int u = 0; u < SB; ++u)
                if (j0 + u < NS) load_row((long)sc[(j0 + u) * 32], xv[u]);
#pragma unroll
            for (int u = 0; u < SB; ++u)
                if (j0 + u < NS) v[u] = sv[(j0 + u) * 32];
#pragma unroll
            for (int u = 0; u < SB; ++u)
                if (j0 + u < NS) {
#pragma unroll
                    for (int w = 0; w < BVS; ++w) acc[w] = A::mad(v[u], xv[u][w], acc[w]);
                }
        }
    }
    // run-time slot count: kept for column-major block vectors (scalar gathers per vector; the templated form raised the register
    // count there and was 35-75 % slower, measured)
    __device__ __forceinline__ void piece_rt(const int ns, const VT *sv, const int *sc) {
#pragma unroll
        for (int j0 = 0; j0 < LMAX; j0 += SB) {
            if (j0 < ns) {
                alignas(16) VT xv[SB][BVS];
                VT v[SB];
#pragma unroll
                for (int u = 0; u < SB; ++u)
                    if (j0 + u < ns) load_row((long)sc[(j0 + u) * 32], xv[u]);
#pragma unroll
                for (int u = 0; u < SB; ++u)
                    if (j0 + u < ns) v[u] = sv[(j0 + u) * 32];
#pragma unroll
                for (int u = 0; u < SB; ++u)
                    if (j0 + u < ns) {
#pragma unroll
                        for (int w = 0; w < BVS; ++w) acc[w] = A::mad(v[u], xv[u][w], acc[w]);
                    }
            }
        }
    }
    __device__ __forceinline__ void piece(const int ns, const VT *sv, const int *sc) {
        static_assert(LMAX <= 8, "piece() dispatches on up to 8 slots");
        if constexpr (!ROWWISE) {
            piece_rt(ns, sv, sc);
            return;
        }
        switch (ns) {
        case 1: piece_n<1>(sv, sc); break;
        case 2: piece_n<(LMAX >= 2 ? 2 : 1)>(sv, sc); break;
        case 3: piece_n<(LMAX >= 3 ? 3 : 1)>(sv, sc); break;
        case 4: piece_n<(LMAX >= 4 ? 4 : 1)>(sv, sc); break;
        case 5: piece_n<(LMAX >= 5 ? 5 : 1)>(sv, sc); break;
        case 6: piece_n<(LMAX >= 6 ? 6 : 1)>(sv, sc); break;
        case 7: piece_n<(LMAX >= 7 ? 7 : 1)>(sv, sc); break;
        default: piece_n<LMAX>(sv, sc); break;
        }
    }
    __device__ __forceinline__ void end_chunk(const int chunk) {
        const long row = (long)chunk * 32 + lane;
        if constexpr (ROWWISE) {
            alignas(16) VT yv[BVS];
#pragma unroll
            for (int v = 0; v < BVS; ++v) yv[v] = A::out(acc[v]);
            constexpr int BYTES = BVS * (int)sizeof(VT);
            if constexpr (BYTES % 16 == 0) {
#pragma unroll
                for (int k = 0; k < BYTES / 16; ++k) reinterpret_cast<int4 *>(Y + row * BVS)[k] = reinterpret_cast<const int4 *>(yv)[k];
            } else if constexpr (BYTES % 8 == 0) {
#pragma unroll
                for (int k = 0; k < BYTES / 8; ++k) reinterpret_cast<int2 *>(Y + row * BVS)[k] = reinterpret_cast<const int2 *>(yv)[k];
            } else {
#pragma unroll
                for (int v = 0; v < BVS; ++v) Y[row * BVS + v] = yv[v];
            }
        } else {
#pragma unroll
            for (int v = 0; v < BVS; ++v) Y[row + v * ld] = A::out(acc[v]);
        }
    }
};

// Row-major SpMMV with wide block rows (BVS*sizeof(VT) = 32..128 B): T = bytes/16 adjacent lanes fetch ONE gathered row with
// one 128-bit load each, so every warp-wide load touches whole 128-byte lines (a lane-per-row int4 load touches 32 different
// lines per instruction and is L1-wavefront bound: measured 0.47 of peak for dp bvs 8).  Lane l owns the 16-byte slice
// (l % T) of the rows l/T + (32/T)*k, k = 0..T-1 of the chunk; per (row, vector) the FMA order is unchanged (slot order).
template <typename VT, typename A, int LMAX, int BVS, bool COHERENT = false>
struct SpmmvBodyRowWide {
    static constexpr int PER = 16 / (int)sizeof(VT);            // values per 128-bit load
    static constexpr int T = BVS * (int)sizeof(VT) / 16;        // lanes per row == rows per lane
    static constexpr int RPI = 32 / T;                          // rows covered by one warp-wide load
    static constexpr int SB = (T >= 8) ? 1 : (T == 4 ? 2 : 4);  // slots in flight per lane (SB*T 128-bit loads)
    const VT *__restrict__ X;
    VT *__restrict__ Y;
    int lane;
    typename A::acc_t acc[T][PER];
    __device__ __forceinline__ void begin_chunk(int = 0, int = 0) {
#pragma unroll
        for (int k = 0; k < T; ++k)
#pragma unroll
            for (int m = 0; m < PER; ++m) acc[k][m] = A::zero();
    }
    // sv / sc arrive offset by `lane`; rebase to the chunk's lane 0
    template <int NS>  // slot count as a template constant, see SpmmvBody::piece_n
    __device__ __forceinline__ void piece_n(const VT *v0, const int *c0, const int r0, const int part) {
#pragma unroll
        for (int j0 = 0; j0 < NS; j0 += SB) {
            alignas(16) VT xv[SB][T][PER];
            VT v[SB][T];
#pragma unroll
            for (int u = 0; u < SB; ++u)
                if (j0 + u < NS) {
#pragma unroll
                    for (int k = 0; k < T; ++k) {
                        const long col = c0[(j0 + u) * 32 + r0 + RPI * k];
                        const int4 *src = reinterpret_cast<const int4 *>(X + col * BVS) + part;
                        *reinterpret_cast<int4 *>(xv[u][k]) = COHERENT ? __ldcg(src) : __ldg(src);
                    }
                }
#pragma unroll
            for (int u = 0; u < SB; ++u)
                if (j0 + u < NS) {
#pragma unroll
                    for (int k = 0; k < T; ++k) v[u][k] = v0[(j0 + u) * 32 + r0 + RPI * k];
                }
#pragma unroll
            for (int u = 0; u < SB; ++u)
                if (j0 + u < NS) {
#pragma unroll
                    for (int k = 0; k < T; ++k)
#pragma unroll
                        for (int m = 0; m < PER; ++m) acc[k][m] = A::mad(v[u][k], xv[u][k][m], acc[k][m]);
                }
        }
    }
    __device__ __forceinline__ void piece(const int ns, const VT *sv, const int *sc) {
        static_assert(LMAX <= 8, "piece() dispatches on up to 8 slots");
        const VT *v0 = sv - lane;
        const int *c0 = sc - lane;
        const int r0 = lane / T, part = lane % T;
        switch (ns) {
        case 1: piece_n<1>(v0, c0, r0, part); break;
        case 2: piece_n<(LMAX >= 2 ? 2 : 1)>(v0, c0, r0, part); break;
        case 3: piece_n<(LMAX >= 3 ? 3 : 1)>(v0, c0, r0, part); break;
        case 4: piece_n<(LMAX >= 4 ? 4 : 1)>(v0, c0, r0, part); break;
        case 5: piece_n<(LMAX >= 5 ? 5 : 1)>(v0, c0, r0, part); break;
        case 6: piece_n<(LMAX >= 6 ? 6 : 1)>(v0, c0, r0, part); break;
        case 7: piece_n<(LMAX >= 7 ? 7 : 1)>(v0, c0, r0, part); break;
        default: piece_n<LMAX>(v0, c0, r0, part); break;
        }
    }
    __device__ __forceinline__ void end_chunk(const int chunk) {
        const int r0 = lane / T, part = lane % T;
#pragma unroll
        for (int k = 0; k < T; ++k) {
            alignas(16) VT yv[PER];
#pragma unroll
            for (int m = 0; m < PER; ++m) yv[m] = A::out(acc[k][m]);
            const long row = (long)chunk * 32 + r0 + RPI * k;
            reinterpret_cast<int4 *>(Y + row * BVS)[part] = *reinterpret_cast<const int4 *>(yv);
        }
    }
};

// The streaming loop of one warp over work items first, first + W, ... < n_items (item k -> chunk list[k] or k + off).
template <typename VT, int LMAX, int D, typename Body>
__device__ __forceinline__ void stream_items(unsigned char *base, uint64_t *bars, PieceHdr *hdrs, uint32_t &phase_bits, const int W,
                                             const int first, const int lane, const int n_items, const int *__restrict__ chunk_list,
                                             const int chunk_offset, const int *__restrict__ chunk_ptrs,
                                             const int *__restrict__ chunk_lengths, const int *__restrict__ col_idxs,
                                             const VT *__restrict__ values, Body &body, const uint64_t pol, const int n_split = -1,
                                             const int *__restrict__ chunk_list2 = nullptr, const int chunk_offset2 = 0) {
    using R = WarpRing<VT, LMAX, D>;
    // n_split >= 0: items [n_split, n_items) come from a SECOND list (chunk_list2 / chunk_offset2) and carry header flag 8 — the
    // boundary chunks of the fused distributed step, appended to the interior ones so that the ring never drains between them
    auto item_chunk = [&](int k) -> int {  // 32-bit index math throughout
        if (n_split >= 0 && k >= n_split) return chunk_list2 ? chunk_list2[k - n_split] : k - n_split + chunk_offset2;
        return chunk_list ? chunk_list[k] : k + chunk_offset;
    };

    // ---- producer state (meaningful in lane 0 only) --------------------------------------------------
    // Three-deep metadata lookahead so that no load issued by lane 0 is consumed in the same piece:
    //   cur  = (pchunk, plen, pcs)  item being cut into pieces
    //   nxt  = (nchunk, nlen, ncs)  item pc + W, loaded when cur became current
    //   n2chunk                     chunk id of item pc + 2W (only needed with a chunk list)
    int pc = first;  // work item of the next piece (n_items + 2W < 2^31: one item per 32 rows)
    int pj = 0, plen = 0, pcs = 0, pchunk = 0, nlen = 0, ncs = 0, nchunk = 0, n2chunk = 0;
    if (lane == 0) {
        if (pc < n_items) {
            pchunk = item_chunk(pc);
            plen = chunk_lengths[pchunk];
            pcs = chunk_ptrs[pchunk];
        }
        if (pc + W < n_items) {
            nchunk = item_chunk(pc + W);
            nlen = chunk_lengths[nchunk];
            ncs = chunk_ptrs[nchunk];
        }
        if (pc + 2 * W < n_items) n2chunk = item_chunk(pc + 2 * W);
    }
    auto issue = [&](int s) {  // lane 0: fill stage s with the next piece of this warp's stream
        PieceHdr h;
        if (pc >= n_items) {
            h.ns = 0; h.flags = 0; h.chunk = 0; h.pad = 0;
            hdrs[s] = h;
            return;
        }
        const int ns = min(LMAX, plen - pj);
        h.ns = ns;
        h.flags = 4 | (pj == 0 ? 1 : 0) | (pj + ns >= plen ? 2 : 0) | ((n_split >= 0 && pc >= n_split) ? 8 : 0);
        h.chunk = pchunk;
        h.pad = 0;
        hdrs[s] = h;
        if (ns > 0) {
            const int e0 = pcs + pj * 32;  // < n_elements < 2^31
            const uint32_t vb = (uint32_t)ns * 32u * (uint32_t)sizeof(VT), cb = (uint32_t)ns * 128u;
            unsigned char *st = base + s * R::STAGE_BYTES;
            mbar_expect_tx(&bars[s], vb + cb);
            bulk_g2s(st, values + e0, vb, &bars[s], pol);
            bulk_g2s(st + R::VAL_BYTES, col_idxs + e0, cb, &bars[s], pol);
        }
        pj += ns;
        if (pj >= plen) {  // advance: nxt -> cur, start the loads for the new nxt
            pc += W;
            pj = 0;
            pchunk = nchunk; plen = nlen; pcs = ncs;
            nchunk = n2chunk;
            if (pc + W < n_items) {
                nlen = chunk_lengths[nchunk];
                ncs = chunk_ptrs[nchunk];
            }
            if (pc + 2 * W < n_items) n2chunk = item_chunk(pc + 2 * W);
        }
    };

    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < D; ++s) issue(s);
    }
    __syncwarp();

    // ---- consumer -----------------------------------------------------------------------------------------
    body.begin_chunk();
    for (int s = 0;; s = (s + 1 == D) ? 0 : s + 1) {
        const PieceHdr h = hdrs[s];
        if (h.flags == 0) break;
        if (h.flags & 1) body.begin_chunk(h.chunk, h.flags);
        if (h.ns > 0) {
            mbar_wait(&bars[s], (phase_bits >> s) & 1u);
            phase_bits ^= (1u << s);
            const VT *sv = reinterpret_cast<const VT *>(base + s * R::STAGE_BYTES) + lane;
            const int *sc = reinterpret_cast<const int *>(base + s * R::STAGE_BYTES + R::VAL_BYTES) + lane;
            body.piece(h.ns, sv, sc);
        }
        if (h.flags & 2) body.end_chunk(h.chunk);
        __syncwarp();  // every lane is done with stage s (data and header) before it is refilled
        if (lane == 0) issue(s);
        __syncwarp();
    }
}

// The same loop in "lean" form (used by the wide-chunk kernel, where it is the faster one: C = 64 hp 145 -> 131 us).
//
// Round-2 form.  SASS of the round-1 loop: 266 instructions per piece, of which 28 were the arithmetic of a 7-slot piece — lane 0 ran
// the producer (header stores to shared memory, address arithmetic, three-deep metadata lookahead) inside a divergent branch while 31
// lanes waited, every lane then re-read the header, and the run-time slot count predicated every load.  Now
//   * the producer STATE is warp-uniform: every lane keeps the same {item, slot, length, pointer} registers and runs the same few
//     integer instructions; only the mbarrier.expect_tx and the two bulk copies are issued by one elected lane.  No divergent
//     region, no header in shared memory, one __syncwarp per piece instead of two;
//   * the piece header {slots, flags, chunk} of each ring stage lives in registers (the stage loop is unrolled over the D stages, so
//     stage offsets are immediates);
//   * the metadata of the next item is loaded (by all lanes, one broadcast transaction) when the current one becomes current, i.e. a
//     whole chunk before it is needed;
//   * the bodies take the slot count as a template constant (piece_n<NS>).
// Per-row accumulation order is unchanged (one FMA per slot in slot order): results stay bit-identical.
struct PieceReg {
    int ns, flags, chunk;
};

template <typename VT, int LMAX, int D, typename Body, int H = 1>  // H: a slot holds 32 * H elements (wide chunks, C = 32 * H)
__device__ __forceinline__ void stream_items_u(unsigned char *base, uint64_t *bars, PieceHdr *, uint32_t &phase_bits, const int W, const int first,
                                             const int lane, const int n_items, const int *__restrict__ chunk_list, const int chunk_offset,
                                             const int *__restrict__ chunk_ptrs, const int *__restrict__ chunk_lengths,
                                             const int *__restrict__ col_idxs, const VT *__restrict__ values, Body &body, const uint64_t pol) {
    using R = WarpRing<VT, LMAX * H, D>;
    if (first >= n_items) return;  // warp-uniform
    auto item_chunk = [&](int k) -> int { return chunk_list ? __ldg(chunk_list + k) : k + chunk_offset; };  // 32-bit index math throughout

    // ---- producer state, identical in every lane --------------------------------------------------------------------------
    int pc = first, pj = 0;                       // work item / first slot of the next piece
    int pchunk = item_chunk(pc);
    int plen = __ldg(chunk_lengths + pchunk), pcs = __ldg(chunk_ptrs + pchunk);
    int nchunk = 0, nlen = 0, ncs = 0;            // item pc + W, requested when item pc became current
    if (pc + W < n_items) {
        nchunk = item_chunk(pc + W);
        nlen = __ldg(chunk_lengths + nchunk);
        ncs = __ldg(chunk_ptrs + nchunk);
    }
    PieceReg hdr[D];
    auto issue = [&](const int s) {  // fill ring stage s (compile-time constant) with the next piece of this warp's stream
        if (pc >= n_items) {
            hdr[s].ns = 0; hdr[s].flags = 0; hdr[s].chunk = 0;
            return;
        }
        const int ns = min(LMAX, plen - pj);
        hdr[s].ns = ns;
        hdr[s].flags = 4 | (pj == 0 ? 1 : 0) | (pj + ns >= plen ? 2 : 0);
        hdr[s].chunk = pchunk;
        if (ns > 0 && lane == 0) {
            const int e0 = pcs + pj * (32 * H);  // < n_elements < 2^31
            const uint32_t vb = (uint32_t)ns * (32u * H) * (uint32_t)sizeof(VT), cb = (uint32_t)ns * (128u * H);
            unsigned char *st = base + s * R::STAGE_BYTES;
            mbar_expect_tx(&bars[s], vb + cb);
            bulk_g2s(st, values + e0, vb, &bars[s], pol);
            bulk_g2s(st + R::VAL_BYTES, col_idxs + e0, cb, &bars[s], pol);
        }
        pj += ns;
        if (pj >= plen) {  // next item becomes current; request the metadata of the one after it
            pc += W;
            pj = 0;
            pchunk = nchunk; plen = nlen; pcs = ncs;
            if (pc + W < n_items) {
                nchunk = item_chunk(pc + W);
                nlen = __ldg(chunk_lengths + nchunk);
                ncs = __ldg(chunk_ptrs + nchunk);
            }
        }
    };
#pragma unroll
    for (int s = 0; s < D; ++s) issue(s);

    // ---- consumer --------------------------------------------------------------------------------------------------------
    body.begin_chunk();
    bool more = true;
    while (more) {
#pragma unroll
        for (int s = 0; s < D; ++s) {
            const PieceReg h = hdr[s];
            if (h.flags == 0) { more = false; break; }
            if (h.flags & 1) body.begin_chunk(h.chunk);
            if (h.ns > 0) {
                mbar_wait(&bars[s], (phase_bits >> s) & 1u);
                phase_bits ^= (1u << s);
                const VT *sv = reinterpret_cast<const VT *>(base + s * R::STAGE_BYTES) + lane;
                const int *sc = reinterpret_cast<const int *>(base + s * R::STAGE_BYTES + R::VAL_BYTES) + lane;
                body.piece(h.ns, sv, sc);
            }
            if (h.flags & 2) body.end_chunk(h.chunk);
            __syncwarp();  // every lane has read stage s before it is refilled
            issue(s);
        }
    }
}

// FUSED: the fused distributed step, interior and boundary chunks in ONE item stream (SpmvBodyFusedStream).
// Known cost of the fused instance (one rank, nothing to exchange; profiles/r02G_spmv_fused_probe.txt, ncu
// r02H_spmv_sp_fused_instance_ncu_summary.txt): it executes 270 instructions per chunk against 204 of the single-GPU loop (flag
// decoding, the warp-uniform gather branch, a recomputed store address, the two-list producer).  dp is HBM-bound and loses 2 %
// (250 vs 245 us); sp / hp are ISSUE-bound and lose 25-30 % (203 vs 163, 196 vs 150 us), which is why halo.cu gives them the push /
// wait kernels next to the single-GPU interior kernel instead.  Tried in round 2 and dropped (profiles/r02J_early_ack.md):
//   * the boundary chunks in the MIDDLE of the stream with the acknowledgement right behind them, so that neighbours may drift half
//     a kernel apart: the decoupling is worth 2-3 us per step at N = 2 and N = 8, the extra code in the instance costs the same;
//   * the step in three PHASES around the unchanged single-GPU loop (push | first chunk | publish | interior | boundary + ack |
//     interior): ptxas spills 120-320 bytes inside the hot loop at the 64-register bound, boundary pass inlined or not.
template <typename VT, typename A, int LMAX, int D, int WARPS, bool UNPERM, bool FUSED>
__global__ void __launch_bounds__(WARPS * 32, (1024 / (WARPS * 32)) > 0 ? (1024 / (WARPS * 32)) : 1)  // <= 64 registers: 32 warps/SM
k_scs32_stream(long n_items, const int *__restrict__ chunk_list, int chunk_offset, const int *__restrict__ chunk_ptrs,
               const int *__restrict__ chunk_lengths, const int *__restrict__ col_idxs, const VT *__restrict__ values,
               const VT *__restrict__ x, VT *__restrict__ y, const int *__restrict__ new_to_old, const __grid_constant__ FusedArgs fa) {
    // (__grid_constant__: the fused body keeps a POINTER to the argument block; without it taking the address copies the struct to the stack)
    using R = WarpRing<VT, LMAX, D>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char *base = smem_raw + (size_t)warp * R::BYTES_ALIGNED;
    uint64_t *bars = reinterpret_cast<uint64_t *>(base + D * R::STAGE_BYTES);
    PieceHdr *hdrs = reinterpret_cast<PieceHdr *>(base + D * R::STAGE_BYTES + D * 8);
    const long W = (long)gridDim.x * WARPS;
    const long gw = (long)blockIdx.x * WARPS + warp;

    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < D; ++s) mbar_init(&bars[s], 1);
        fence_barrier_init();
    }
    __syncwarp();
    const uint64_t pol = policy_evict_first();
    uint32_t phase_bits = 0;  // one parity bit per stage, flipped after every completed wait

    if constexpr (!FUSED) {
        SpmvBody<VT, A, LMAX, UNPERM, false> body{x, y, new_to_old, lane, A::zero()};
        stream_items<VT, LMAX, D>(base, bars, hdrs, phase_bits, (int)W, (int)gw, lane, (int)n_items, chunk_list, chunk_offset,
                                  chunk_ptrs, chunk_lengths, col_idxs, values, body, pol);
    } else {
        // (a) store this rank's elements into the neighbours' x tails (published after this warp's first chunk, see the body)
        const unsigned int epoch_e = ld_flag(fa.epoch) + 1u;
        const unsigned int mine = fused_push_stores<VT>(fa, x, gw, W, lane, epoch_e);
        // (b) + (c) interior chunks, then — in the same stream — the boundary chunks once the neighbours' elements have landed
        SpmvBodyFusedStream<VT, A, LMAX> body;
        body.x = x; body.y = y; body.lane = lane; body.y_rows = fa.y_rows; body.fa = &fa; body.epoch_e = epoch_e; body.mine = mine;
        body.sig_pending = mine > 0; body.halo_ready = false; body.bnd = false; body.acc = A::zero();
        stream_items<VT, LMAX, D>(base, bars, hdrs, phase_bits, (int)W, (int)gw, lane, (int)(fa.n_int + fa.n_bnd), fa.int_list, fa.int_off, chunk_ptrs,
                                  chunk_lengths, col_idxs, values, body, pol, (int)fa.n_int, fa.bnd_list, fa.bnd_off);
        if (body.sig_pending) body.signal();  // a warp without any chunk still publishes its push
        // (d) the last warp of the grid acknowledges consumption to the senders and closes the epoch
        fused_finish(fa, W, lane, epoch_e);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Wide chunks: C = 32 * H (H = 2, 4: C = 64, 128).  A chunk of C rows is still ONE contiguous range (slot j holds C consecutive
// elements), so the same per-warp ring applies: a stage keeps 8 / H slots of C elements (the stage size of the C = 32 kernel),
// lane l owns the H rows l, l + 32, ... of the chunk with one accumulator each and walks its slots in order (bit-identical to the
// direct kernel and to the oracle).  8 gathers in flight per lane like the C = 32 kernel.
// ---------------------------------------------------------------------------------------------------------------------
// Body of the wide-chunk kernel: lane l owns the H rows l, l + 32, ... of the chunk; a piece of NS slots is NS * H slot-rows of 32
// elements, slot-row r = j * H + h holds slot j of the rows h * 32 + lane.
template <typename VT, typename A, int H, bool UNPERM, bool COHERENT = false, bool BOUNDED = false>
struct WideBody {
    static constexpr int C = 32 * H;
    const VT *__restrict__ x;
    VT *__restrict__ y;
    const int *__restrict__ new_to_old;
    int lane;
    int y_rows = 0;  // BOUNDED: positions >= y_rows are not stored (y aliases the next x, whose tail holds the halo)
    typename A::acc_t acc[H];
    __device__ __forceinline__ void begin_chunk(int = 0, int = 0) {
#pragma unroll
        for (int h = 0; h < H; ++h) acc[h] = A::zero();
    }
    template <int NS>
    __device__ __forceinline__ void piece_n(const VT *sv, const int *sc) {
        constexpr int NR = NS * H;
        VT xv[NR];
        int col[NR];
#pragma unroll
        for (int r = 0; r < NR; ++r) col[r] = sc[r * 32];
#pragma unroll
        for (int r = 0; r < NR; ++r) xv[r] = COHERENT ? __ldcg(x + col[r]) : __ldg(x + col[r]);
#pragma unroll
        for (int r = 0; r < NR; ++r) acc[r % H] = A::mad(sv[r * 32], xv[r], acc[r % H]);
    }
    __device__ __forceinline__ void piece(const int ns, const VT *sv, const int *sc) {
        static_assert(8 % H == 0 && H <= 8, "a stage holds 8 slot-rows");
        switch (ns) {
        case 0: break;
        case 1: piece_n<1>(sv, sc); break;
        case 2: piece_n<(8 / H >= 2 ? 2 : 1)>(sv, sc); break;
        case 3: piece_n<(8 / H >= 3 ? 3 : 1)>(sv, sc); break;
        default: piece_n<(8 / H >= 4 ? 4 : 1)>(sv, sc); break;
        }
    }
    __device__ __forceinline__ void end_chunk(const int chunk) {
#pragma unroll
        for (int h = 0; h < H; ++h) {
            const long row = (long)chunk * C + h * 32 + lane;
            if (UNPERM) {
                const int o = new_to_old[row];
                if (o >= 0) y[o] = A::out(acc[h]);
            } else if (!BOUNDED || row < y_rows)
                y[row] = A::out(acc[h]);
        }
    }
};

template <typename VT, typename A, int H, int D, int WARPS, bool UNPERM, bool FUSED = false>
__global__ void __launch_bounds__(WARPS * 32, (1024 / (WARPS * 32)) > 0 ? (1024 / (WARPS * 32)) : 1)
k_scsw_stream(int n_items, const int *__restrict__ chunk_list, int chunk_offset, const int *__restrict__ chunk_ptrs,
              const int *__restrict__ chunk_lengths, const int *__restrict__ col_idxs, const VT *__restrict__ values,
              const VT *__restrict__ x, VT *__restrict__ y, const int *__restrict__ new_to_old, const FusedArgs fa) {
    static_assert(H == 2 || H == 4, "C = 64 or 128");
    constexpr int LS = 8;         // slot-rows (32 elements each) per stage
    constexpr int LW = LS / H;    // slots of a wide chunk per piece
    using R = WarpRing<VT, LS, D>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char *base = smem_raw + (size_t)warp * R::BYTES_ALIGNED;
    uint64_t *bars = reinterpret_cast<uint64_t *>(base + D * R::STAGE_BYTES);
    const int W = (int)gridDim.x * WARPS;
    const int first = (int)blockIdx.x * WARPS + warp;
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < D; ++s) mbar_init(&bars[s], 1);
        fence_barrier_init();
    }
    __syncwarp();
    const uint64_t pol = policy_evict_first();
    uint32_t phase_bits = 0;
    if constexpr (!FUSED) {
        WideBody<VT, A, H, UNPERM> body;
        body.x = x; body.y = y; body.new_to_old = new_to_old; body.lane = lane;
        stream_items_u<VT, LW, D, WideBody<VT, A, H, UNPERM>, H>(base, bars, nullptr, phase_bits, W, first, lane, n_items, chunk_list, chunk_offset,
                                                               chunk_ptrs, chunk_lengths, col_idxs, values, body, pol);
    } else {
        // ONE kernel per distributed SpMV for C = 64 / 128 (see k_scs32_stream): push, interior chunks, wait, boundary chunks, acknowledge
        const unsigned int epoch_e = ld_flag(fa.epoch) + 1u;
        fused_push<VT>(fa, x, first, W, lane, epoch_e);
        {
            using B = WideBody<VT, A, H, UNPERM, false, true>;
            B body;
            body.x = x; body.y = y; body.new_to_old = new_to_old; body.lane = lane; body.y_rows = fa.y_rows;
            stream_items_u<VT, LW, D, B, H>(base, bars, nullptr, phase_bits, W, first, lane, (int)fa.n_int, fa.int_list, fa.int_off, chunk_ptrs,
                                            chunk_lengths, col_idxs, values, body, pol);
        }
        if (first < fa.n_bnd) {
            warp_wait_flags(fa.arrived, fa.is_sender, fa.P, epoch_e, lane, fa.error);
            using B = WideBody<VT, A, H, UNPERM, true, true>;
            B body;
            body.x = x; body.y = y; body.new_to_old = new_to_old; body.lane = lane; body.y_rows = fa.y_rows;
            stream_items_u<VT, LW, D, B, H>(base, bars, nullptr, phase_bits, W, first, lane, (int)fa.n_bnd, fa.bnd_list, fa.bnd_off, chunk_ptrs,
                                            chunk_lengths, col_idxs, values, body, pol);
        }
        fused_finish(fa, W, lane, epoch_e);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Narrow value types, short chunks: TWO consecutive SELL-32 chunks per work item ("pair" kernel, used for fp16).
// ncu on the fp16 instance of k_scs32_stream (profiles/r02e_*): 240 warp instructions per 7-slot piece at IPC 3.0 — a piece carries
// only 1.3 KB, so the per-piece cost (producer, barrier wait, header) is paid per 1.3 KB and each lane has at most 7 gathers in flight.
// Chunks 2k and 2k + 1 are adjacent in memory (chunk_ptrs[c + 1] = chunk_ptrs[c] + len * 32), so their len_A + len_B slots form ONE
// contiguous run that is cut into pieces of <= LMAX (16) slots: one pair of bulk copies, one barrier wait and one producer step now
// serve up to 14-16 slots, and up to 16 gathers are in flight per lane.  Inside a piece slots [0, a) belong to chunk A, [a, ns) to
// chunk B; each lane runs ONE accumulator chain in slot order and switches from A's sum to B's at slot a — per row the FMA order is
// the storage order, bit-identical to the one-chunk kernel and the oracle.  Lean loop form (warp-uniform producer, see stream_items_u).
// ---------------------------------------------------------------------------------------------------------------------
template <typename VT, typename A, int LMAX, int D, int WARPS, bool UNPERM>
__global__ void __launch_bounds__(WARPS * 32, (1024 / (WARPS * 32)) > 0 ? (1024 / (WARPS * 32)) : 1)
k_scs32_stream_pair(int n_chunks, int chunk_offset, const int *__restrict__ chunk_ptrs, const int *__restrict__ col_idxs,
                    const VT *__restrict__ values, const VT *__restrict__ x, VT *__restrict__ y, const int *__restrict__ new_to_old) {
    using R = WarpRing<VT, LMAX, D>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char *base = smem_raw + (size_t)warp * R::BYTES_ALIGNED;
    uint64_t *bars = reinterpret_cast<uint64_t *>(base + D * R::STAGE_BYTES);
    const int W = (int)gridDim.x * WARPS;
    const int first = (int)blockIdx.x * WARPS + warp;
    const int n_items = (n_chunks + 1) / 2;
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < D; ++s) mbar_init(&bars[s], 1);
        fence_barrier_init();
    }
    __syncwarp();
    if (first >= n_items) return;
    const uint64_t pol = policy_evict_first();

    // ---- producer (warp-uniform state): item k = chunks 2k, 2k + 1 of the range; lengths from the pointer differences --------------
    auto load_item = [&](const int k, int &cs, int &la, int &lb) {
        const int c = chunk_offset + 2 * k;
        const int p0 = __ldg(chunk_ptrs + c), p1 = __ldg(chunk_ptrs + c + 1);
        const int p2 = (2 * k + 1 < n_chunks) ? __ldg(chunk_ptrs + c + 2) : p1;
        cs = p0; la = (p1 - p0) >> 5; lb = (p2 - p1) >> 5;
    };
    int pc = first, pj = 0, pcs, pla, plb, ncs = 0, nla = 0, nlb = 0;
    load_item(pc, pcs, pla, plb);
    if (pc + W < n_items) load_item(pc + W, ncs, nla, nlb);
    struct PairReg { int ns, a, flags, item; };  // flags: 1 A begins, 2 A ends, 4 B begins, 8 B ends, 16 valid
    PairReg hdr[D];
    auto issue = [&](const int s) {
        if (pc >= n_items) {
            hdr[s].ns = 0; hdr[s].a = 0; hdr[s].flags = 0; hdr[s].item = 0;
            return;
        }
        const int tot = pla + plb;
        const int ns = min(LMAX, tot - pj);
        const int a = max(0, min(ns, pla - pj));            // slots of chunk A in this piece
        int f = 16;
        if (pj == 0) f |= 1;                                 // A begins with the item
        if (pj <= pla && pj + ns >= pla) f |= 2 | 4;         // A's last slot (or an empty A) lies in this piece: A ends, B begins
        if (pj + ns >= tot) f |= 8;                          // B ends
        hdr[s].ns = ns; hdr[s].a = a; hdr[s].flags = f; hdr[s].item = pc;
        if (ns > 0 && lane == 0) {
            const int e0 = pcs + pj * 32;
            const uint32_t vb = (uint32_t)ns * 32u * (uint32_t)sizeof(VT), cb = (uint32_t)ns * 128u;
            unsigned char *st = base + s * R::STAGE_BYTES;
            mbar_expect_tx(&bars[s], vb + cb);
            bulk_g2s(st, values + e0, vb, &bars[s], pol);
            bulk_g2s(st + R::VAL_BYTES, col_idxs + e0, cb, &bars[s], pol);
        }
        pj += ns;
        if (pj >= tot) {
            pc += W;
            pj = 0;
            pcs = ncs; pla = nla; plb = nlb;
            if (pc + W < n_items) load_item(pc + W, ncs, nla, nlb);
        }
    };
#pragma unroll
    for (int s = 0; s < D; ++s) issue(s);

    // ---- consumer ------------------------------------------------------------------------------------------------------------------
    uint32_t phase_bits = 0;
    typename A::acc_t accA = A::zero(), accB = A::zero();
    auto store = [&](const int chunk, const typename A::acc_t acc) {
        if (chunk >= chunk_offset + n_chunks) return;  // the odd tail has no chunk B
        const long row = (long)chunk * 32 + lane;
        if (UNPERM) {
            const int o = new_to_old[row];
            if (o >= 0) y[o] = A::out(acc);
        } else
            y[row] = A::out(acc);
    };
    bool running = true;
    while (running) {
#pragma unroll
        for (int s = 0; s < D; ++s) {
            const PairReg h = hdr[s];
            if (h.flags == 0) { running = false; break; }
            if (h.flags & 1) accA = A::zero();
            if (h.flags & 4) accB = A::zero();
            if (h.ns > 0) {
                mbar_wait(&bars[s], (phase_bits >> s) & 1u);
                phase_bits ^= (1u << s);
                const VT *sv = reinterpret_cast<const VT *>(base + s * R::STAGE_BYTES) + lane;
                const int *sc = reinterpret_cast<const int *>(base + s * R::STAGE_BYTES + R::VAL_BYTES) + lane;
                int col[LMAX];
                VT xv[LMAX];
#pragma unroll
                for (int j = 0; j < LMAX; ++j)
                    if (j < h.ns) col[j] = sc[j * 32];
#pragma unroll
                for (int j = 0; j < LMAX; ++j)
                    if (j < h.ns) xv[j] = __ldg(x + col[j]);
                // one chain in slot order; at slot a it leaves chunk A's sum and continues chunk B's
                typename A::acc_t acc = h.a > 0 ? accA : accB;
#pragma unroll
                for (int j = 0; j < LMAX; ++j)
                    if (j < h.ns) {
                        if (j == h.a && j > 0) { accA = acc; acc = accB; }
                        acc = A::mad(sv[j * 32], xv[j], acc);
                    }
                if (h.a == h.ns) accA = acc; else accB = acc;
            }
            const int cA = chunk_offset + 2 * h.item;
            if (h.flags & 2) store(cA, accA);
            if (h.flags & 8) store(cA + 1, accB);
            __syncwarp();  // every lane has read stage s before it is refilled
            issue(s);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Narrow chunks: C = 32 / G (G = 2, 4: C = 16 — a half-warp —, 8).  A warp takes G ADJACENT chunks at once (work item k = chunks
// G k ... G k + G - 1 of a contiguous chunk range): lane l belongs to sub-chunk g = l / C, so the 32 rows of an item are the 32
// consecutive padded rows 32 k + l.  Every piece covers the same slot window [pj, pj + 8) of all G sub-chunks; sub-chunk g
// contributes ns_g = clamp(len_g - pj, 0, 8) slots, fetched by its own pair of bulk copies into region g of the stage (regions
// are one slot-row apart modulo the bank window, so the two half-warps never meet in a shared-memory bank).  Per row the slots
// are consumed in order: bit-identical to the direct kernel and the oracle.
// ---------------------------------------------------------------------------------------------------------------------
template <typename VT, int G, int D>
struct NarrowRing {  // per-warp ring of the narrow-chunk kernel: G regions per array, one slot-row of padding BETWEEN regions
    static constexpr int LMAX = 8, C = 32 / G;
    static constexpr int RV = (LMAX + 1) * C * (int)sizeof(VT), RC = (LMAX + 1) * C * 4;  // region strides
    static constexpr int VAL_BYTES = G * LMAX * C * (int)sizeof(VT) + (G - 1) * C * (int)sizeof(VT);
    static constexpr int COL_BYTES = G * LMAX * C * 4 + (G - 1) * C * 4;
    static constexpr int STAGE_BYTES = VAL_BYTES + COL_BYTES;
    static constexpr int META_BYTES = 4 * G * 4;  // producer state of lane 0: {len, ptr} of the current item (+ spare)
    static constexpr int BYTES = D * STAGE_BYTES + D * 8 + D * (int)sizeof(PieceHdr) + META_BYTES;
    static constexpr int BYTES_ALIGNED = (BYTES + 127) / 128 * 128;
};

template <typename VT, typename A, int G, int D, int WARPS, bool UNPERM>
__global__ void __launch_bounds__(WARPS * 32, (1024 / (WARPS * 32)) > 0 ? (1024 / (WARPS * 32)) : 1)
k_scsn_stream(int n_chunks, int chunk_offset, const int *__restrict__ chunk_ptrs, const int *__restrict__ chunk_lengths,
              const int *__restrict__ col_idxs, const VT *__restrict__ values, const VT *__restrict__ x, VT *__restrict__ y,
              const int *__restrict__ new_to_old) {
    using R = NarrowRing<VT, G, D>;
    constexpr int LMAX = R::LMAX, C = R::C, RV = R::RV, RC = R::RC;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char *base = smem_raw + (size_t)warp * R::BYTES_ALIGNED;
    uint64_t *bars = reinterpret_cast<uint64_t *>(base + D * R::STAGE_BYTES);
    PieceHdr *hdrs = reinterpret_cast<PieceHdr *>(base + D * R::STAGE_BYTES + D * 8);
    int *meta = reinterpret_cast<int *>(base + D * R::STAGE_BYTES + D * 8 + D * (int)sizeof(PieceHdr));
    int *plen = meta, *pcs = meta + G;  // current item: shared memory, touched by lane 0 only
    int nlen[G], ncs[G];                // next item: registers, so that its loads stay in flight until the item becomes current
    const int W = (int)gridDim.x * WARPS;
    const int first = (int)blockIdx.x * WARPS + warp;
    const int n_items = (n_chunks + G - 1) / G;
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < D; ++s) mbar_init(&bars[s], 1);
        fence_barrier_init();
    }
    __syncwarp();
    const uint64_t pol = policy_evict_first();
    uint32_t phase_bits = 0;

    // ---- producer (lane 0): current item + one item of metadata lookahead ---------------------------------------------
    int pc = first, pj = 0;
    auto load_meta = [&](int item, int *len, int *cs) {
#pragma unroll
        for (int g = 0; g < G; ++g) {
            const int c = item * G + g;
            const bool ok = c < n_chunks;
            len[g] = ok ? chunk_lengths[chunk_offset + c] : 0;
            cs[g] = ok ? chunk_ptrs[chunk_offset + c] : 0;
        }
    };
#pragma unroll
    for (int g = 0; g < G; ++g) { nlen[g] = 0; ncs[g] = 0; }
    if (lane == 0) {
#pragma unroll
        for (int g = 0; g < 2 * G; ++g) meta[g] = 0;
        if (pc < n_items) load_meta(pc, plen, pcs);
        if (pc + W < n_items) load_meta(pc + W, nlen, ncs);
    }
    auto issue = [&](int s) {
        PieceHdr h;
        if (pc >= n_items) {
            h.ns = 0; h.flags = 0; h.chunk = 0; h.pad = 0;
            hdrs[s] = h;
            return;
        }
        int maxlen = 0;
#pragma unroll
        for (int g = 0; g < G; ++g) maxlen = max(maxlen, plen[g]);
        unsigned int packed = 0, bytes = 0;
#pragma unroll
        for (int g = 0; g < G; ++g) {
            const int n = min(LMAX, max(0, plen[g] - pj));
            packed |= (unsigned int)n << (8 * g);
            bytes += (unsigned int)(n * C) * ((unsigned int)sizeof(VT) + 4u);
        }
        const int ns = min(LMAX, maxlen - pj);
        h.ns = ns;
        h.flags = 4 | (pj == 0 ? 1 : 0) | (pj + ns >= maxlen ? 2 : 0);
        h.chunk = pc;
        h.pad = (int)packed;
        hdrs[s] = h;
        if (bytes > 0) {
            unsigned char *st = base + s * R::STAGE_BYTES;
            mbar_expect_tx(&bars[s], bytes);
#pragma unroll
            for (int g = 0; g < G; ++g) {
                const int n = (int)((packed >> (8 * g)) & 0xffu);
                if (n > 0) {
                    const int e0 = pcs[g] + pj * C;
                    bulk_g2s(st + g * RV, values + e0, (uint32_t)(n * C) * (uint32_t)sizeof(VT), &bars[s], pol);
                    bulk_g2s(st + R::VAL_BYTES + g * RC, col_idxs + e0, (uint32_t)(n * C) * 4u, &bars[s], pol);
                }
            }
        }
        pj += ns;
        if (pj >= maxlen) {
            pc += W;
            pj = 0;
#pragma unroll
            for (int g = 0; g < G; ++g) { plen[g] = nlen[g]; pcs[g] = ncs[g]; }
            if (pc + W < n_items) load_meta(pc + W, nlen, ncs);
        }
    };
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < D; ++s) issue(s);
    }
    __syncwarp();

    // ---- consumer ----------------------------------------------------------------------------------------------------
    const int g = lane / C, l = lane % C;
    typename A::acc_t acc = A::zero();
    for (int s = 0;; s = (s + 1 == D) ? 0 : s + 1) {
        const PieceHdr hd = hdrs[s];
        if (hd.flags == 0) break;
        if (hd.flags & 1) acc = A::zero();
        if (hd.pad != 0) {  // some sub-chunk has slots in this piece
            mbar_wait(&bars[s], (phase_bits >> s) & 1u);
            phase_bits ^= (1u << s);
            const int my_ns = ((unsigned int)hd.pad >> (8 * g)) & 0xff;
            const VT *sv = reinterpret_cast<const VT *>(base + s * R::STAGE_BYTES + g * RV) + l;
            const int *sc = reinterpret_cast<const int *>(base + s * R::STAGE_BYTES + R::VAL_BYTES + g * RC) + l;
            VT v[LMAX], xv[LMAX];
            int col[LMAX];
#pragma unroll
            for (int j = 0; j < LMAX; ++j)
                if (j < my_ns) col[j] = sc[j * C];
#pragma unroll
            for (int j = 0; j < LMAX; ++j)
                if (j < my_ns) xv[j] = __ldg(x + col[j]);
#pragma unroll
            for (int j = 0; j < LMAX; ++j)
                if (j < my_ns) v[j] = sv[j * C];
#pragma unroll
            for (int j = 0; j < LMAX; ++j)
                if (j < my_ns) acc = A::mad(v[j], xv[j], acc);
        }
        if (hd.flags & 2) {
            const long row = ((long)chunk_offset + (long)hd.chunk * G) * C + lane;  // the item's 32 consecutive padded rows
            if (row < ((long)chunk_offset + n_chunks) * C) {
                if (UNPERM) {
                    const int o = new_to_old[row];
                    if (o >= 0) y[o] = A::out(acc);
                } else
                    y[row] = A::out(acc);
            }
        }
        __syncwarp();
        if (lane == 0) issue(s);
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Variant with SOFTWARE-PIPELINED x gathers (D >= 3).  ncu on k_scs32_stream (profiles/r01i_*): 46 % of the stall samples
// sit on the first FMA of a piece, i.e. warps wait for the x gathers they have just issued.  Here the gathers of piece p+1
// are issued BEFORE the FMAs of piece p, so a piece's gathers have one whole piece-time to come back:
//   stage s   : piece p    (gathered x already in registers)  -> FMAs now
//   stage s+1 : piece p+1  (data landed)                       -> read columns, issue gathers now
//   stage s+2 : piece p+2                                      -> bulk copy in flight
// Same producer, same per-row FMA order (bit-identical); costs a second set of x registers (~80 regs, 24 warps/SM).
// ---------------------------------------------------------------------------------------------------------------------
template <typename VT, typename A, int LMAX, int D, int WARPS, bool UNPERM>
__global__ void __launch_bounds__(WARPS * 32, 768 / (WARPS * 32))
k_scs32_stream_pf(int n_items, const int *__restrict__ chunk_list, int chunk_offset, const int *__restrict__ chunk_ptrs,
                  const int *__restrict__ chunk_lengths, const int *__restrict__ col_idxs, const VT *__restrict__ values,
                  const VT *__restrict__ x, VT *__restrict__ y, const int *__restrict__ new_to_old) {
    static_assert(D >= 3, "the prefetching consumer needs three stages");
    using R = WarpRing<VT, LMAX, D>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char *base = smem_raw + (size_t)warp * R::BYTES_ALIGNED;
    uint64_t *bars = reinterpret_cast<uint64_t *>(base + D * R::STAGE_BYTES);
    PieceHdr *hdrs = reinterpret_cast<PieceHdr *>(base + D * R::STAGE_BYTES + D * 8);
    const int W = (int)gridDim.x * WARPS;
    const int first = (int)blockIdx.x * WARPS + warp;
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < D; ++s) mbar_init(&bars[s], 1);
        fence_barrier_init();
    }
    __syncwarp();
    const uint64_t pol = policy_evict_first();
    auto item_chunk = [&](int k) -> int { return chunk_list ? chunk_list[k] : k + chunk_offset; };

    int pc = first;
    int pj = 0, plen = 0, pcs = 0, pchunk = 0, nlen = 0, ncs = 0, nchunk = 0, n2chunk = 0;
    if (lane == 0) {
        if (pc < n_items) { pchunk = item_chunk(pc); plen = chunk_lengths[pchunk]; pcs = chunk_ptrs[pchunk]; }
        if (pc + W < n_items) { nchunk = item_chunk(pc + W); nlen = chunk_lengths[nchunk]; ncs = chunk_ptrs[nchunk]; }
        if (pc + 2 * W < n_items) n2chunk = item_chunk(pc + 2 * W);
    }
    auto issue = [&](int s) {
        PieceHdr h;
        if (pc >= n_items) {
            h.ns = 0; h.flags = 0; h.chunk = 0; h.pad = 0;
            hdrs[s] = h;
            return;
        }
        const int ns = min(LMAX, plen - pj);
        h.ns = ns;
        h.flags = 4 | (pj == 0 ? 1 : 0) | (pj + ns >= plen ? 2 : 0);
        h.chunk = pchunk;
        h.pad = 0;
        hdrs[s] = h;
        if (ns > 0) {
            const int e0 = pcs + pj * 32;
            const uint32_t vb = (uint32_t)ns * 32u * (uint32_t)sizeof(VT), cb = (uint32_t)ns * 128u;
            unsigned char *st = base + s * R::STAGE_BYTES;
            mbar_expect_tx(&bars[s], vb + cb);
            bulk_g2s(st, values + e0, vb, &bars[s], pol);
            bulk_g2s(st + R::VAL_BYTES, col_idxs + e0, cb, &bars[s], pol);
        }
        pj += ns;
        if (pj >= plen) {
            pc += W;
            pj = 0;
            pchunk = nchunk; plen = nlen; pcs = ncs;
            nchunk = n2chunk;
            if (pc + W < n_items) { nlen = chunk_lengths[nchunk]; ncs = chunk_ptrs[nchunk]; }
            if (pc + 2 * W < n_items) n2chunk = item_chunk(pc + 2 * W);
        }
    };
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < D; ++s) issue(s);
    }
    __syncwarp();

    uint32_t phase_bits = 0;
    auto gather = [&](int s, int ns, VT *xo) {  // wait for stage s, read its columns, issue the x gathers
        mbar_wait(&bars[s], (phase_bits >> s) & 1u);
        phase_bits ^= (1u << s);
        const int *sc = reinterpret_cast<const int *>(base + s * R::STAGE_BYTES + R::VAL_BYTES) + lane;
        int col[LMAX];
#pragma unroll
        for (int j = 0; j < LMAX; ++j)
            if (j < ns) col[j] = sc[j * 32];
#pragma unroll
        for (int j = 0; j < LMAX; ++j)
            if (j < ns) xo[j] = __ldg(x + col[j]);
    };

    typename A::acc_t acc = A::zero();
    int s = 0;
    PieceHdr hc = hdrs[0];
    if (hc.flags == 0) return;
    VT xv[LMAX];
    if (hc.ns > 0) gather(0, hc.ns, xv);
    for (;;) {
        const int sn = (s + 1 == D) ? 0 : s + 1;
        const PieceHdr hn = hdrs[sn];
        VT xn[LMAX];
        if (hn.flags != 0 && hn.ns > 0) gather(sn, hn.ns, xn);
        if (hc.flags & 1) acc = A::zero();
        if (hc.ns > 0) {
            const VT *sv = reinterpret_cast<const VT *>(base + s * R::STAGE_BYTES) + lane;
            VT v[LMAX];
#pragma unroll
            for (int j = 0; j < LMAX; ++j)
                if (j < hc.ns) v[j] = sv[j * 32];
#pragma unroll
            for (int j = 0; j < LMAX; ++j)
                if (j < hc.ns) acc = A::mad(v[j], xv[j], acc);
        }
        if (hc.flags & 2) {
            const long row = (long)hc.chunk * 32 + lane;
            if (UNPERM) {
                const int o = new_to_old[row];
                if (o >= 0) y[o] = A::out(acc);
            } else
                y[row] = A::out(acc);
        }
        __syncwarp();
        if (lane == 0) issue(s);
        __syncwarp();
        if (hn.flags == 0) break;
        hc = hn;
#pragma unroll
        for (int j = 0; j < LMAX; ++j) xv[j] = xn[j];
        s = sn;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// "Virtual items" for matrices with very uneven chunk lengths (power-law rows).  One warp per chunk keeps the reference's
// summation order but a lone warp streams only ~3 GB/s, so a 4096-slot chunk alone takes ~1 ms.  Chunks longer than a
// threshold are therefore cut into slot segments {chunk, first slot, slots, partial slot}; each segment is an independent
// work item whose 32 row sums go to a partial buffer, and k_reduce_partials adds a chunk's partials in segment order
// (deterministic; differs from the strictly sequential sum in the last bits for those rows only).  Items are ordered
// longest first.  Kept separate from stream_items so that the main kernel's register allocation is untouched.
// ---------------------------------------------------------------------------------------------------------------------
template <typename VT, typename A, int LMAX, int D, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, (1024 / (WARPS * 32)) > 0 ? (1024 / (WARPS * 32)) : 1)
k_scs32_stream_split(long n_items, const int4 *__restrict__ vitems, const int *__restrict__ chunk_ptrs, const int *__restrict__ col_idxs,
                     const VT *__restrict__ values, const VT *__restrict__ x, VT *__restrict__ y, VT *__restrict__ partial) {
    using R = WarpRing<VT, LMAX, D>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char *base = smem_raw + (size_t)warp * R::BYTES_ALIGNED;
    uint64_t *bars = reinterpret_cast<uint64_t *>(base + D * R::STAGE_BYTES);
    PieceHdr *hdrs = reinterpret_cast<PieceHdr *>(base + D * R::STAGE_BYTES + D * 8);
    const long W = (long)gridDim.x * WARPS;
    const long gw = (long)blockIdx.x * WARPS + warp;
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < D; ++s) mbar_init(&bars[s], 1);
        fence_barrier_init();
    }
    __syncwarp();
    const uint64_t pol = policy_evict_first();

    long pc = gw;
    int pj = 0, pcs = 0, ncs = 0;
    int4 cur = make_int4(0, 0, 0, -1), nxt = cur, nx2 = cur;  // {chunk, first slot, slots, partial slot}
    if (lane == 0) {
        if (pc < n_items) { cur = vitems[pc]; pcs = chunk_ptrs[cur.x]; }
        if (pc + W < n_items) { nxt = vitems[pc + W]; ncs = chunk_ptrs[nxt.x]; }
        if (pc + 2 * W < n_items) nx2 = vitems[pc + 2 * W];
    }
    auto issue = [&](int s) {
        PieceHdr h;
        if (pc >= n_items) {
            h.ns = 0; h.flags = 0; h.chunk = 0; h.pad = 0;
            hdrs[s] = h;
            return;
        }
        const int ns = min(LMAX, cur.z - pj);
        h.ns = ns;
        h.flags = 4 | (pj == 0 ? 1 : 0) | (pj + ns >= cur.z ? 2 : 0);
        h.chunk = cur.x;
        h.pad = cur.w;
        hdrs[s] = h;
        if (ns > 0) {
            const long e0 = (long)pcs + (long)(cur.y + pj) * 32;
            const uint32_t vb = (uint32_t)ns * 32u * (uint32_t)sizeof(VT), cb = (uint32_t)ns * 128u;
            unsigned char *st = base + s * R::STAGE_BYTES;
            mbar_expect_tx(&bars[s], vb + cb);
            bulk_g2s(st, values + e0, vb, &bars[s], pol);
            bulk_g2s(st + R::VAL_BYTES, col_idxs + e0, cb, &bars[s], pol);
        }
        pj += ns;
        if (pj >= cur.z) {
            pc += W;
            pj = 0;
            cur = nxt; pcs = ncs;
            nxt = nx2;
            if (pc + W < n_items) ncs = chunk_ptrs[nxt.x];
            if (pc + 2 * W < n_items) nx2 = vitems[pc + 2 * W];
        }
    };
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < D; ++s) issue(s);
    }
    __syncwarp();

    uint32_t phase_bits = 0;
    SpmvBody<VT, A, LMAX, false, false> body{x, y, nullptr, lane, A::zero()};
    for (int s = 0;; s = (s + 1 == D) ? 0 : s + 1) {
        const PieceHdr h = hdrs[s];
        if (h.flags == 0) break;
        if (h.flags & 1) body.begin_chunk();
        if (h.ns > 0) {
            mbar_wait(&bars[s], (phase_bits >> s) & 1u);
            phase_bits ^= (1u << s);
            const VT *sv = reinterpret_cast<const VT *>(base + s * R::STAGE_BYTES) + lane;
            const int *sc = reinterpret_cast<const int *>(base + s * R::STAGE_BYTES + R::VAL_BYTES) + lane;
            body.piece(h.ns, sv, sc);
        }
        if (h.flags & 2) {
            if (h.pad < 0) body.end_chunk(h.chunk);
            else partial[(long)h.pad * 32 + lane] = A::out(body.acc);
        }
        __syncwarp();
        if (lane == 0) issue(s);
        __syncwarp();
    }
}

// y[chunk rows] = partial[s0] + partial[s0+1] + ... (in that order) for every split chunk; one warp per chunk
template <typename VT, typename A>
__global__ void k_reduce_partials(long n_split, const int *__restrict__ split_chunk, const int *__restrict__ split_ptr,
                                  const VT *__restrict__ partial, VT *__restrict__ y) {
    const long w = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= n_split) return;
    const int s0 = split_ptr[w], s1 = split_ptr[w + 1];
    typename A::acc_t acc = partial[(long)s0 * 32 + lane];
    for (int s = s0 + 1; s < s1; ++s) acc = A::add(acc, partial[(long)s * 32 + lane]);
    y[(long)split_chunk[w] * 32 + lane] = A::out(acc);
}

// ---------------------------------------------------------------------------------------------------------------------
// CRS (C = 1, sigma = 1) through the same per-warp ring.  A warp takes blocks of 32 consecutive rows; their elements are
// one contiguous range [row_ptrs[r0], row_ptrs[r0+32]) which is streamed in tiles of LMAX*32 elements (tile starts aligned to
// 8 elements so that every bulk copy is 16-byte aligned for fp64 / fp32 / fp16; copies stop at nnz & ~7, the last < 8 elements of
// the matrix are read with plain loads, so caller-owned arrays without slack are fine).
// Lane l walks the part of ITS row that lies inside the current tile sequentially, so the per-row summation order is the
// reference's (kernels.hpp:46-57): bit-identical to the sequential loop, unlike the split-row vector kernel.
// ---------------------------------------------------------------------------------------------------------------------
template <typename VT, typename A, int LMAX, int D, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, (1024 / (WARPS * 32)) > 0 ? (1024 / (WARPS * 32)) : 1)
k_csr_stream(long n_rows, const int *__restrict__ row_ptrs, const int *__restrict__ col_idxs, const VT *__restrict__ values,
             const VT *__restrict__ x, VT *__restrict__ y) {
    using R = WarpRing<VT, LMAX, D>;
    constexpr int TILE = LMAX * 32;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char *base = smem_raw + (size_t)warp * R::BYTES_ALIGNED;
    uint64_t *bars = reinterpret_cast<uint64_t *>(base + D * R::STAGE_BYTES);
    const long W = (long)gridDim.x * WARPS;
    const long gw = (long)blockIdx.x * WARPS + warp;
    const long n_blocks = (n_rows + 31) / 32;
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < D; ++s) mbar_init(&bars[s], 1);
        fence_barrier_init();
    }
    __syncwarp();
    const uint64_t pol = policy_evict_first();

    const int nnz8 = __ldg(row_ptrs + n_rows) & ~7;  // elements below this index are fetched by (16-byte granular) bulk copies
    if (gw >= n_blocks) return;
    auto block_range = [&](long rb, int &b0, int &b1) {
        const long r0 = rb * 32, r1 = min(r0 + 32, n_rows);
        b0 = __ldg(row_ptrs + r0);
        b1 = __ldg(row_ptrs + r1);
    };
    // ---- producer: tiles of the blocks gw, gw + W, ...; warp-uniform state, only the copies are elected (see stream_items) ---------
    long pb = gw;
    int pb0 = 0, pb1 = 0, nb0 = 0, nb1 = 0, pt = 0;  // current block's element range, next block's, next tile start
    block_range(pb, pb0, pb1);
    pt = pb0 & ~7;
    if (pb + W < n_blocks) block_range(pb + W, nb0, nb1);
    struct TileReg { int ns, flags, chunk, pad; };
    TileReg hdr[D];
    auto issue = [&](const int s) {
        if (pb >= n_blocks) {
            hdr[s].ns = 0; hdr[s].flags = 0; hdr[s].chunk = 0; hdr[s].pad = 0;
            return;
        }
        int n = 0;
        if (pt < pb1 && pt < nnz8) {
            const int end8 = min((pb1 + 7) & ~7, nnz8);  // never read past the arrays: the last (< 8) elements go through global loads
            n = min(TILE, end8 - pt);
            if (lane == 0) {
                unsigned char *st = base + s * R::STAGE_BYTES;
                const uint32_t vb = (uint32_t)n * (uint32_t)sizeof(VT), cb = (uint32_t)n * 4u;
                mbar_expect_tx(&bars[s], vb + cb);
                bulk_g2s(st, values + pt, vb, &bars[s], pol);
                bulk_g2s(st + R::VAL_BYTES, col_idxs + pt, cb, &bars[s], pol);
            }
        }
        const bool first = pt == (pb0 & ~7);
        const bool last = pt + n >= pb1 || pt + n >= nnz8;
        hdr[s].ns = n;                 // elements in this tile (0: the block has no elements at all)
        hdr[s].flags = 4 | (first ? 1 : 0) | (last ? 2 : 0);
        hdr[s].chunk = (int)pb;        // row block
        hdr[s].pad = pt;               // first element of the tile
        pt += n;
        if (last) {
            pb += W;
            pb0 = nb0; pb1 = nb1; pt = pb0 & ~7;
            if (pb + W < n_blocks) block_range(pb + W, nb0, nb1);
        }
    };
#pragma unroll
    for (int s = 0; s < D; ++s) issue(s);

    // ---- consumer -----------------------------------------------------------------------------------------------------
    uint32_t phase_bits = 0;
    typename A::acc_t acc = A::zero();
    int beg = 0, end = 0;
    bool running = true;
    while (running) {
#pragma unroll
        for (int s = 0; s < D; ++s) {
            const TileReg h = hdr[s];
            if (h.flags == 0) { running = false; break; }
            if (h.flags & 1) {
                acc = A::zero();
                const long r = (long)h.chunk * 32 + lane;
                beg = __ldg(row_ptrs + min(r, n_rows));
                end = __ldg(row_ptrs + min(r + 1, n_rows));
            }
            if (h.ns > 0) {
                mbar_wait(&bars[s], (phase_bits >> s) & 1u);
                phase_bits ^= (1u << s);
                const VT *sv = reinterpret_cast<const VT *>(base + s * R::STAGE_BYTES);
                const int *sc = reinterpret_cast<const int *>(base + s * R::STAGE_BYTES + R::VAL_BYTES);
                const int lo = max(beg, h.pad) - h.pad, hi = min(end, h.pad + h.ns) - h.pad;
                constexpr int G = sizeof(VT) == 8 ? 4 : 8;  // gathers in flight per lane (fp64: 4, the 64-register budget)
                int j = lo;
                for (; j + G <= hi; j += G) {  // full groups: no predication
                    int col[G];
                    VT xv[G];
#pragma unroll
                    for (int u = 0; u < G; ++u) col[u] = sc[j + u];
#pragma unroll
                    for (int u = 0; u < G; ++u) xv[u] = __ldg(x + col[u]);
#pragma unroll
                    for (int u = 0; u < G; ++u) acc = A::mad(sv[j + u], xv[u], acc);
                }
                if (j < hi) {  // remainder of 1 .. G-1 elements
                    int col[G - 1];
                    VT xv[G - 1];
#pragma unroll
                    for (int u = 0; u < G - 1; ++u)
                        if (j + u < hi) col[u] = sc[j + u];
#pragma unroll
                    for (int u = 0; u < G - 1; ++u)
                        if (j + u < hi) xv[u] = __ldg(x + col[u]);
#pragma unroll
                    for (int u = 0; u < G - 1; ++u)
                        if (j + u < hi) acc = A::mad(sv[j + u], xv[u], acc);
                }
            }
            if (h.flags & 2) {
                // the final nnz % 8 elements of the matrix are not covered by a bulk copy (caller-owned arrays have no slack):
                // the rows that own them finish through global loads, in order
                for (int e = max(beg, h.pad + h.ns); e < end; ++e) acc = A::mad(values[e], __ldg(x + col_idxs[e]), acc);
                const long r = (long)h.chunk * 32 + lane;
                if (r < n_rows) y[r] = A::out(acc);
            }
            __syncwarp();
            issue(s);
        }
    }
}

// One warp's pass over a chunk subset with the SpMMV body: items first, first + W, ... < n (COH: gathers through L2 only, for chunks
// that read a halo a peer GPU wrote during this kernel).
template <typename VT, typename A, int LMAX, int D, int BVS, bool ROWWISE, bool USE_WIDE, bool COH>
__device__ __forceinline__ void mmv_run(unsigned char *base, uint64_t *bars, PieceHdr *hdrs, uint32_t &phase_bits, const int W, const int first,
                                        const int lane, const int n, const int *__restrict__ list, const int off,
                                        const int *__restrict__ chunk_ptrs, const int *__restrict__ chunk_lengths,
                                        const int *__restrict__ col_idxs, const VT *__restrict__ values, const VT *__restrict__ X,
                                        VT *__restrict__ Y, const long ld, const uint64_t pol) {
    using Body = std::conditional_t<USE_WIDE, SpmmvBodyRowWide<VT, A, LMAX, BVS, COH>, SpmmvBody<VT, A, LMAX, BVS, ROWWISE, COH>>;
    Body body;
    body.X = X; body.Y = Y; body.lane = lane;
    if constexpr (!USE_WIDE) body.ld = ld;
    stream_items<VT, LMAX, D>(base, bars, hdrs, phase_bits, W, first, lane, n, list, off, chunk_ptrs, chunk_lengths, col_idxs, values, body, pol);
}

// The boundary pass of the fused distributed step as an OUT-OF-LINE function, because ptxas is fragile here (B200, 256^3 7-point
// matrix, one rank without a neighbour so that only the instance is measured; profiles/r02q_mmv_fused_instance.md):
//   * a second copy of the unrolled body inlined after the main loop changed the schedule of the FIRST one (sp, block_vec_size 8:
//     one gather in flight per lane instead of two, same instruction count and traffic, 334 -> 419 us);
//   * publishing the push from the main loop's end_chunk (an out-of-line call taken once) cost dp block_vec_size 8 at its
//     80-register bound a factor 1.8;
//   * the warp's first chunk as a call before an inlined main loop: sp block_vec_size 8 +14 %;
//   * the main loop itself out of line: dp block_vec_size 4 +20 %, dp 8 +47 %.
// What ships: the push stores inline at the start, the main loop inline (exactly the single-GPU kernel's), the publish right behind it
// (published at once after the stores, every step cost 25-35 us more: r02z vs r02A_dist_probe_mmv.txt), the boundary out of line.
template <typename VT, typename A, int LMAX, int D, int BVS, bool ROWWISE, bool USE_WIDE, bool COH>
__device__ __noinline__ uint32_t mmv_run_outofline(unsigned char *base, uint64_t *bars, PieceHdr *hdrs, uint32_t phase_bits, const int W,
                                                   const int gw, const int lane, const int n, const int *list, const int off,
                                                   const int *chunk_ptrs, const int *chunk_lengths, const int *col_idxs, const VT *values,
                                                   const VT *X, VT *Y, const long ld, const uint64_t pol) {
    mmv_run<VT, A, LMAX, D, BVS, ROWWISE, USE_WIDE, COH>(base, bars, hdrs, phase_bits, W, gw, lane, n, list, off, chunk_ptrs, chunk_lengths,
                                                         col_idxs, values, X, Y, ld, pol);
    return phase_bits;
}

// SELL-32 SpMMV through the same per-warp bulk-copy ring (smaller stages: the block vectors want the L1 capacity).
// MINB = minimum resident CTAs per SM handed to __launch_bounds__.  It is NOT cosmetic: with no minimum (MINB = 0) ptxas aims for the
// smallest register count and re-serialises the gathers the source issues together (sp, block_vec_size 8: 48 registers, ONE or two
// 128-bit gathers in flight per lane instead of eight); with a minimum it takes the registers the bound allows and keeps them in flight.
template <typename VT, typename A, int LMAX, int D, int WARPS, int BVS, bool ROWWISE, bool WIDE, bool FUSED = false, int MINB = 0>
__global__ void __launch_bounds__(WARPS * 32, MINB)
k_scs32_stream_mmv(long n_items, const int *__restrict__ chunk_list, int chunk_offset, const int *__restrict__ chunk_ptrs,
                   const int *__restrict__ chunk_lengths, const int *__restrict__ col_idxs, const VT *__restrict__ values,
                   const VT *__restrict__ X, VT *__restrict__ Y, long ld, const __grid_constant__ FusedArgs fa) {
    using R = WarpRing<VT, LMAX, D>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char *base = smem_raw + (size_t)warp * R::BYTES_ALIGNED;
    uint64_t *bars = reinterpret_cast<uint64_t *>(base + D * R::STAGE_BYTES);
    PieceHdr *hdrs = reinterpret_cast<PieceHdr *>(base + D * R::STAGE_BYTES + D * 8);
    const long W = (long)gridDim.x * WARPS;
    const long gw = (long)blockIdx.x * WARPS + warp;
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < D; ++s) mbar_init(&bars[s], 1);
        fence_barrier_init();
    }
    __syncwarp();
    const uint64_t pol = policy_evict_first();
    uint32_t phase_bits = 0;
    constexpr int ROW_BYTES = BVS * (int)sizeof(VT);
    constexpr bool USE_WIDE = WIDE && ROWWISE && ROW_BYTES >= 32 && ROW_BYTES <= 128 && (ROW_BYTES & (ROW_BYTES - 1)) == 0;
    if constexpr (!FUSED) {
        mmv_run<VT, A, LMAX, D, BVS, ROWWISE, USE_WIDE, false>(base, bars, hdrs, phase_bits, (int)W, (int)gw, lane, (int)n_items, chunk_list,
                                                               chunk_offset, chunk_ptrs, chunk_lengths, col_idxs, values, X, Y, ld, pol);
    } else {
        // ONE kernel per distributed SpMMV (see k_scs32_stream): (a) store all block_vec_size values of the halo rows into the neighbours'
        // vectors — 16-byte segments spread over ALL warps, so each warp issues at most a few coalesced stores; (b) the interior
        // chunks; (c) publish the push (system-scope fence + counter; the last pushing warp raises the neighbours' `arrived` flags);
        // (d) wait for the own halo, boundary chunks with L2-coherent gathers; (e) acknowledge.
        const unsigned int epoch_e = ld_flag(fa.epoch) + 1u;
        fused_push_stores<VT>(fa, X, gw, W, lane, epoch_e);
        mmv_run<VT, A, LMAX, D, BVS, ROWWISE, USE_WIDE, false>(base, bars, hdrs, phase_bits, (int)W, (int)gw, lane, (int)fa.n_int, fa.int_list, fa.int_off,
                                                               chunk_ptrs, chunk_lengths, col_idxs, values, X, Y, ld, pol);
        {   // publish the push now that its stores have long landed (recomputing the warp's group count keeps the loop's registers free)
            const long n_push_warps = (fused_push_units<VT>(fa) + 31) / 32;
            if (gw < n_push_warps) fused_push_signal(fa, fused_push_units<VT>(fa), (unsigned int)((n_push_warps - gw + W - 1) / W), lane, epoch_e);
        }
        if (gw < fa.n_bnd) {
            warp_wait_flags(fa.arrived, fa.is_sender, fa.P, epoch_e, lane, fa.error);
            phase_bits = mmv_run_outofline<VT, A, LMAX, D, BVS, ROWWISE, USE_WIDE, true>(base, bars, hdrs, phase_bits, (int)W, (int)gw, lane,
                                                                                        (int)fa.n_bnd, fa.bnd_list, fa.bnd_off, chunk_ptrs,
                                                                                        chunk_lengths, col_idxs, values, X, Y, ld, pol);
        }
        fused_finish(fa, W, lane, epoch_e);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Adaptive precision, C = 32: ONE pass over the dp / sp / hp parts of every chunk through the same per-warp ring.
// The parts share C, n_chunks and the row permutation; each has its own chunk_ptrs / chunk_lengths / col_idxs / values.
// For a chunk the producer emits the pieces of the dp part, then sp, then hp; the header's `pad` field carries the part.
// Arithmetic = the reference's library kernels (interface.hpp:1434-1733): one fp64 accumulator per part, the narrower value
// widened before the FMA, y = dp + sp + hp; MODE 2 (sp_hp): fp32 products against the fp32 x, fp64 accumulators, fp32 y.
// ---------------------------------------------------------------------------------------------------------------------
struct ApPart {
    const int *cp, *cl, *ci;
    const void *v;
};

template <int MODE, int LMAX, int D, int WARPS, int MINB = 1>
__global__ void __launch_bounds__(WARPS * 32, MINB)
k_scs32_stream_ap(long n_items, const int *__restrict__ order, const int4 *__restrict__ items, ApPart p0, ApPart p1, ApPart p2,
                  const void *__restrict__ xv_, void *__restrict__ yv_, double *__restrict__ partial) {
    using R = WarpRing<double, LMAX, D>;  // stages sized for the widest part
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char *base = smem_raw + (size_t)warp * R::BYTES_ALIGNED;
    uint64_t *bars = reinterpret_cast<uint64_t *>(base + D * R::STAGE_BYTES);
    const long W = (long)gridDim.x * WARPS;
    const long gw = (long)blockIdx.x * WARPS + warp;
    constexpr bool USE0 = MODE != 2, USE1 = MODE != 1, USE2 = MODE != 0;  // dp, sp, hp parts in use
    const ApPart parts[3] = {p0, p1, p2};

    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < D; ++s) mbar_init(&bars[s], 1);
        fence_barrier_init();
    }
    __syncwarp();
    const uint64_t pol = policy_evict_first();

    // ---- producer: items gw, gw + W, ...; per item the parts in order dp, sp, hp.  The state is WARP-UNIFORM (every lane keeps the
    //      same registers and runs the same integer instructions, see stream_items); only expect_tx + the bulk copies are elected ----
    if (gw >= n_items) return;
    long pc = gw;
    int pchunk = 0, nchunk = 0;
    int plen[3] = {0, 0, 0}, pcs[3] = {0, 0, 0}, nlen[3] = {0, 0, 0}, ncs[3] = {0, 0, 0};
    int pp = 0, pj = 0;        // current part and slot inside it
    bool pfirst = true;        // no piece of the current chunk emitted yet
    auto load_meta = [&](int chunk, int *len, int *cs) {
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            const bool use = q == 0 ? USE0 : (q == 1 ? USE1 : USE2);
            len[q] = use ? __ldg(parts[q].cl + chunk) : 0;
            cs[q] = use ? __ldg(parts[q].cp + chunk) : 0;
        }
    };
    // With `items` (very uneven matrices) a work item is either a whole chunk (code < 0) or ONE slot segment of ONE part of a long
    // chunk, {chunk, first slot, slots, code = partial slot << 2 | part}; its sum goes to partial[] and k_reduce_partials_ap adds the
    // segments of a part in slot order.  A segment looks to the producer like a chunk whose other parts are empty.
    // Lookahead: the DESCRIPTOR of item pc + 2W is fetched one advance before its chunk metadata (lengths / pointers of up to three
    // parts, which depend on the chunk id) is requested, and that metadata is first used one advance later again.
    int pcode = -1, ncode = -1;
    int4 n2it = make_int4(0, 0, 0, -1);
    auto fetch_desc = [&](long k) -> int4 {
        if (items) return __ldg(items + k);
        return make_int4(order ? __ldg(order + k) : (int)k, 0, 0, -1);
    };
    auto load_item = [&](const int4 it, int &chunk, int *len, int *cs, int &code) {
        chunk = it.x;
        code = it.w;
        if (code < 0) load_meta(chunk, len, cs);
        else {
            const int part = code & 3;
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                len[q] = q == part ? it.z : 0;
                cs[q] = q == part ? __ldg(parts[q].cp + chunk) + it.y * 32 : 0;
            }
        }
    };
    load_item(fetch_desc(pc), pchunk, plen, pcs, pcode);
    if (pc + W < n_items) load_item(fetch_desc(pc + W), nchunk, nlen, ncs, ncode);
    if (pc + 2 * W < n_items) n2it = fetch_desc(pc + 2 * W);
    struct ApPieceReg { int ns, flags, chunk, pad; };
    ApPieceReg hdr[D];
    auto issue = [&](const int s) {
        hdr[s].ns = 0; hdr[s].flags = 0; hdr[s].chunk = 0; hdr[s].pad = 0;
        if (pc >= n_items) return;
        while (pp < 3 && pj >= plen[pp]) { ++pp; pj = 0; }  // skip exhausted / empty parts
        int ns = 0;
        if (pp < 3) {
            ns = min(LMAX, plen[pp] - pj);
            if (lane == 0) {
                const long e0 = (long)pcs[pp] + (long)pj * 32;
                const uint32_t vsz = pp == 0 ? 8u : (pp == 1 ? 4u : 2u);
                const uint32_t vb = (uint32_t)ns * 32u * vsz, cb = (uint32_t)ns * 128u;
                unsigned char *st = base + s * R::STAGE_BYTES;
                mbar_expect_tx(&bars[s], vb + cb);
                bulk_g2s(st, static_cast<const unsigned char *>(parts[pp].v) + e0 * vsz, vb, &bars[s], pol);
                bulk_g2s(st + R::VAL_BYTES, parts[pp].ci + e0, cb, &bars[s], pol);
            }
            hdr[s].pad = pp | (pcode < 0 ? 0 : ((pcode >> 2) + 1) << 2);
            pj += ns;
        }
        // is anything left in this chunk after this piece?
        bool more = false;
#pragma unroll
        for (int q = 0; q < 3; ++q) more |= (q > pp && plen[q] > 0) || (q == pp && pj < plen[q]);
        hdr[s].ns = ns;
        hdr[s].flags = 4 | (pfirst ? 1 : 0) | (more ? 0 : 2);
        hdr[s].chunk = pchunk;
        pfirst = false;
        if (!more) {  // next chunk
            pc += W;
            pp = 0; pj = 0; pfirst = true;
            pchunk = nchunk;
            pcode = ncode;
#pragma unroll
            for (int q = 0; q < 3; ++q) { plen[q] = nlen[q]; pcs[q] = ncs[q]; }
            if (pc + W < n_items) load_item(n2it, nchunk, nlen, ncs, ncode);
            if (pc + 2 * W < n_items) n2it = fetch_desc(pc + 2 * W);
        }
    };
#pragma unroll
    for (int s = 0; s < D; ++s) issue(s);

    // ---- consumer: one piece = NS slots of ONE part; the slot count is a template constant (see SpmvBody::piece_n) ----------------
    uint32_t phase_bits = 0;
    double acc[3] = {0.0, 0.0, 0.0};
    auto piece_n = [&](auto ns_tag, const int part, const unsigned char *sv, const int *sc) {
        constexpr int NS = decltype(ns_tag)::value;
        int col[NS];
#pragma unroll
        for (int j = 0; j < NS; ++j) col[j] = sc[j * 32];
        if constexpr (MODE == 2) {
            const float *x = static_cast<const float *>(xv_);
            float xf[NS];
#pragma unroll
            for (int j = 0; j < NS; ++j) xf[j] = __ldg(x + col[j]);
            if (part == 1) {
                const float *v = reinterpret_cast<const float *>(sv) + lane;
#pragma unroll
                for (int j = 0; j < NS; ++j) acc[1] += (double)__fmul_rn(v[j * 32], xf[j]);
            } else {
                const __half *v = reinterpret_cast<const __half *>(sv) + lane;
#pragma unroll
                for (int j = 0; j < NS; ++j) acc[2] += (double)__fmul_rn(__half2float(v[j * 32]), xf[j]);
            }
        } else {
            const double *x = static_cast<const double *>(xv_);
            double xd[NS];
#pragma unroll
            for (int j = 0; j < NS; ++j) xd[j] = __ldg(x + col[j]);
            if (part == 0) {
                const double *v = reinterpret_cast<const double *>(sv) + lane;
#pragma unroll
                for (int j = 0; j < NS; ++j) acc[0] = fma(v[j * 32], xd[j], acc[0]);
            } else if (part == 1) {
                const float *v = reinterpret_cast<const float *>(sv) + lane;
#pragma unroll
                for (int j = 0; j < NS; ++j) acc[1] = fma((double)v[j * 32], xd[j], acc[1]);
            } else {
                const __half *v = reinterpret_cast<const __half *>(sv) + lane;
#pragma unroll
                for (int j = 0; j < NS; ++j) acc[2] = fma((double)__half2float(v[j * 32]), xd[j], acc[2]);
            }
        }
    };
    static_assert(LMAX == 8, "the AP consumer dispatches on 1..8 slots");
    bool running = true;
    while (running) {
#pragma unroll
        for (int s = 0; s < D; ++s) {
            const ApPieceReg h = hdr[s];
            if (h.flags == 0) { running = false; break; }
            if (h.flags & 1) { acc[0] = 0.0; acc[1] = 0.0; acc[2] = 0.0; }
            const int part = h.pad & 3, pslot1 = h.pad >> 2;  // pslot1 > 0: segment of a split chunk -> partial[pslot1 - 1]
            if (h.ns > 0) {
                mbar_wait(&bars[s], (phase_bits >> s) & 1u);
                phase_bits ^= (1u << s);
                const unsigned char *sv = base + s * R::STAGE_BYTES;
                const int *sc = reinterpret_cast<const int *>(base + s * R::STAGE_BYTES + R::VAL_BYTES) + lane;
                switch (h.ns) {
                case 1: piece_n(std::integral_constant<int, 1>{}, part, sv, sc); break;
                case 2: piece_n(std::integral_constant<int, 2>{}, part, sv, sc); break;
                case 3: piece_n(std::integral_constant<int, 3>{}, part, sv, sc); break;
                case 4: piece_n(std::integral_constant<int, 4>{}, part, sv, sc); break;
                case 5: piece_n(std::integral_constant<int, 5>{}, part, sv, sc); break;
                case 6: piece_n(std::integral_constant<int, 6>{}, part, sv, sc); break;
                case 7: piece_n(std::integral_constant<int, 7>{}, part, sv, sc); break;
                default: piece_n(std::integral_constant<int, 8>{}, part, sv, sc); break;
                }
            }
            if ((h.flags & 2) && pslot1 > 0) {
                partial[(long)(pslot1 - 1) * 32 + lane] = part == 0 ? acc[0] : (part == 1 ? acc[1] : acc[2]);
            } else if (h.flags & 2) {
                const long row = (long)h.chunk * 32 + lane;
                if constexpr (MODE == 2) static_cast<float *>(yv_)[row] = (float)(acc[1] + acc[2]);
                else if constexpr (MODE == 0) static_cast<double *>(yv_)[row] = acc[0] + acc[1];
                else if constexpr (MODE == 1) static_cast<double *>(yv_)[row] = acc[0] + acc[2];
                else static_cast<double *>(yv_)[row] = acc[0] + acc[1] + acc[2];
            }
            __syncwarp();  // every lane has read stage s before it is refilled
            issue(s);
        }
    }
}

// y[rows of a split chunk] from its segment sums: per part the segments are added in slot order, then dp + sp + hp as in the
// un-split path; one warp per split chunk
template <int MODE>
__global__ void k_reduce_partials_ap(long n_split, const int *__restrict__ split_chunk, const int *__restrict__ split_ptr,
                                     const unsigned char *__restrict__ seg_part, const double *__restrict__ partial, void *__restrict__ yv_) {
    const long w = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= n_split) return;
    double acc[3] = {0.0, 0.0, 0.0};
    bool any[3] = {false, false, false};
    for (int s = split_ptr[w]; s < split_ptr[w + 1]; ++s) {
        const int q = seg_part[s];
        const double v = partial[(long)s * 32 + lane];
#pragma unroll
        for (int k = 0; k < 3; ++k)
            if (q == k) { acc[k] = any[k] ? acc[k] + v : v; any[k] = true; }
    }
    const long row = (long)split_chunk[w] * 32 + lane;
    if constexpr (MODE == 2) static_cast<float *>(yv_)[row] = (float)(acc[1] + acc[2]);
    else if constexpr (MODE == 0) static_cast<double *>(yv_)[row] = acc[0] + acc[1];
    else if constexpr (MODE == 1) static_cast<double *>(yv_)[row] = acc[0] + acc[2];
    else static_cast<double *>(yv_)[row] = acc[0] + acc[1] + acc[2];
}

}  // namespace stream
}  // namespace uspmv
