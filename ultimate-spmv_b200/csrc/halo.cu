// Row partitioning and halo bookkeeping for the one-process-per-GPU mode.
//
// Replaces
//   seg_work_sharing_arr        code/mpi_funcs.hpp:424-622   (host logic, rank 0 in the reference)
//   collect_local_needed_heri   code/mpi_funcs.hpp:242-415   (serial hash-set scan -> device kernels here)
//   pack_send_buf / pack_d_send_buf  code/classes_structs.hpp:786-831, code/kernels.hpp:554-577
//
// Halo discovery on the device, bit-exact with the reference's first-seen-order numbering:
//   1. every slot (padding included, storage order) with a non-local column does atomicMin(first[col], slot);
//   2. the distinct remote columns are compacted and sorted by (owner rank, first slot) — that IS the
//      reference's order: grouped by owner ascending, first-seen order inside an owner (SURVEY.md §8a' 12);
//   3. halo id = n_local + rank in that order; columns are rewritten in place; the need list sent to owner j
//      holds owner-local indices col - wsa[j].
#include "common.cuh"
#include "scs_stream.cuh"

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <cstring>
#include <vector>

using namespace uspmv;

namespace uspmv {  // spmv_kernels.cu
void launch_scs32_fused(const uspmv_scs *s, const void *x, void *y, const stream::FusedArgs &fa, cudaStream_t st);
bool scs_fused_supported(const uspmv_scs *s);
bool spmmv_fused_supported(const uspmv_scs *s, int bvs);
void launch_spmmv_fused_any(const uspmv_scs *s, const void *X, void *Y, int bvs, long ld, int layout, const stream::FusedArgs &fa, cudaStream_t st);
}

struct uspmv_halo {
    uspmv_ctx *ctx = nullptr;
    int rank = 0, P = 1;
    long n_local = 0, n_halo = 0;
    std::vector<int> recv_cumsum;  // P+1
    std::vector<int> need_flat;    // n_halo owner-local indices, grouped by owner
    std::vector<int> send_ptr;     // P+1
    long n_send = 0;
    DevBuf<int> send_idx;          // owner-local row ids requested by the peers (concatenated)
    const int *perm_d = nullptr;   // old_to_new of the local matrix (x lives in permuted order)
};

namespace {

constexpr int TPB = 256;
inline unsigned blocks_for(long n) { return (unsigned)((n + TPB - 1) / TPB); }

__global__ void k_fill_int(int *p, long n, int v) {
    long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// one thread per (permuted) row position: slot j of the row is padding iff j >= row_lengths[p]
__global__ void k_mark_remote(const int *__restrict__ chunk_ptrs, const int *__restrict__ chunk_lengths, const int *__restrict__ row_lengths,
                              const int *__restrict__ col_idxs, long n_pad, int C, int lo, int hi, int n_glob, bool strict,
                              int slot_base, int *__restrict__ first) {
    long p = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (p >= n_pad) return;
    const long c = p / C;
    const int lane = (int)(p - c * C), len = chunk_lengths[c], rl = row_lengths[p];
    long e = (long)chunk_ptrs[c] + lane;
    for (int j = 0; j < len; ++j, e += C) {
        if (j >= rl && !strict) continue;  // padding is not a remote element unless the reference is replicated
        const int col = col_idxs[e];
        if (col >= lo && col < hi) continue;
        if (col < 0 || col >= n_glob) continue;  // no owner: the reference leaves such a column untouched
        atomicMin(&first[col], slot_base + (int)e);  // slot_base: storage order continues over the parts of an AP matrix
    }
}

__global__ void k_flag_seen(const int *__restrict__ first, long n_glob, int *__restrict__ flag) {
    long c = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (c < n_glob) flag[c] = first[c] != INT32_MAX;
}

__global__ void k_compact_remote(const int *__restrict__ first, const int *__restrict__ flag, const int *__restrict__ pos, long n_glob,
                                 const int *__restrict__ wsa, int P, unsigned long long *__restrict__ keys, int *__restrict__ cols) {
    long c = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (c >= n_glob || !flag[c]) return;
    int owner = 0;
    while (owner + 1 < P && (int)c >= wsa[owner + 1]) ++owner;
    const int k = pos[c];
    keys[k] = ((unsigned long long)owner << 32) | (unsigned int)first[c];
    cols[k] = (int)c;
}

// after the sort: k-th distinct remote column gets halo id n_local + k
__global__ void k_assign_halo_ids(const unsigned long long *__restrict__ keys, const int *__restrict__ cols, long n_halo, int n_local,
                                  const int *__restrict__ wsa, int *__restrict__ remap, int *__restrict__ need_flat,
                                  int *__restrict__ owner_of_k) {
    long k = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (k >= n_halo) return;
    const int owner = (int)(keys[k] >> 32);
    const int col = cols[k];
    remap[col] = n_local + (int)k;
    need_flat[k] = col - wsa[owner];
    owner_of_k[k] = owner;
}

__global__ void k_rewrite_cols(const int *__restrict__ chunk_ptrs, const int *__restrict__ chunk_lengths, const int *__restrict__ row_lengths,
                               int *__restrict__ col_idxs, long n_pad, int C, int lo, int hi, int n_glob, bool strict,
                               const int *__restrict__ remap) {
    long p = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (p >= n_pad) return;
    const long c = p / C;
    const int lane = (int)(p - c * C), len = chunk_lengths[c], rl = row_lengths[p];
    long e = (long)chunk_ptrs[c] + lane;
    for (int j = 0; j < len; ++j, e += C) {
        if (j >= rl && !strict) { col_idxs[e] = 0; continue; }  // padding -> local column 0 (value is 0)
        const int col = col_idxs[e];
        if (col >= lo && col < hi) col_idxs[e] = col - lo;
        else if (col >= 0 && col < n_glob) col_idxs[e] = remap[col];
    }
}

template <typename VT>
__global__ void k_pack(const int *__restrict__ send_idx, const int *__restrict__ perm, long n_send, const VT *__restrict__ x,
                       VT *__restrict__ buf, int bvs, long ld, int layout) {
    long t = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (t >= n_send * bvs) return;
    long i, v;
    if (layout == USPMV_ROWWISE) { i = t / bvs; v = t % bvs; }
    else { v = t / n_send; i = t % n_send; }
    const long src = perm ? perm[send_idx[i]] : send_idx[i];
    // buffer layout: rowwise -> [i][v]; colwise -> [v][i] (one contiguous run per vector, like bulkvec messages)
    if (layout == USPMV_ROWWISE) buf[i * bvs + v] = x[src * bvs + v];
    else buf[v * n_send + i] = x[src + v * ld];
}

}  // namespace

extern "C" {

int uspmv_seg_work_sharing_arr(int seg_method, long n_rows, long nnz, const int *I, int P, int *wsa) {
    return guarded([&] {
        if (!I || !wsa) fail("uspmv_seg_work_sharing_arr: NULL argument");
        if (P < 1) fail("uspmv_seg_work_sharing_arr: comm_size must be >= 1");
        if (n_rows < P) fail("seg_work_sharing_arr ERROR: total_mtx->n_rows < comm_size.");  // mpi_funcs.hpp:442-444
        if (nnz < 1) fail("uspmv_seg_work_sharing_arr: empty matrix");
        for (int s = 0; s <= P; ++s) wsa[s] = 0;
        if (seg_method == USPMV_SEG_ROWS) {
            const int per = (int)(n_rows / P);
            for (int s = 1; s <= P; ++s) wsa[s] = s * per;
            wsa[P] = I[nnz - 1] + 1;
        } else if (seg_method == USPMV_SEG_NNZ) {
            const int per = (int)(nnz / P);
            int seg = 1, local = 0;
            for (long g = 0; g < nnz; ++g) {
                if (local == per) {
                    if (seg <= P) wsa[seg] = I[g] + 1;
                    ++seg;
                    local = 0;
                    continue;
                }
                ++local;
            }
            wsa[P] = I[nnz - 1] + 1;
        } else
            fail("uspmv_seg_work_sharing_arr: unknown seg_method %d (seg-metis is out of scope)", seg_method);
        if (wsa[P - 1] == wsa[P])  // "last process gets no work" fix-up, mpi_funcs.hpp:602-606
            for (int r = 1; r < P; ++r) wsa[r] -= 1;
        for (int i = 1; i <= P; ++i)
            if (wsa[i] < wsa[i - 1]) fail("seg_work_sharing_arr ERROR: flaw in work_sharing_arr, work_sharing_arr[i] < work_sharing_arr[i-1].");
    });
}

/* Halo discovery over n_parts matrices that share the rows of one rank (the dp / sp / hp parts of an adaptive-precision matrix; one
 * part = collect_local_needed_heri, mpi_funcs.hpp:242-415): ONE numbering of the remote columns for all parts, first-seen order of a
 * scan that walks part 0's storage, then part 1's, ...  x_permuted = 0: the x vector stays in the original row order (AP structs keep
 * the original column numbering, main.cpp:1308-1332), so the elements sent to the neighbours are x[send_idx], not x[perm[send_idx]]. */
int uspmv_halo_plan_create_multi(uspmv_scs **parts, int n_parts, const int *wsa_h, int rank, int P, int x_permuted, uspmv_halo **out) {
    return guarded([&] {
        OptScope opt_scope(ctx_of(parts && n_parts > 0 ? parts[0] : nullptr));  // this context's options govern everything below
        if (!parts || !wsa_h || !out || n_parts < 1 || n_parts > 3) fail("uspmv_halo_plan_create: NULL argument or bad part count");
        if (P < 1 || rank < 0 || rank >= P) fail("uspmv_halo_plan_create: bad rank/comm_size");
        uspmv_scs *s0 = parts[0];
        long slots_total = 0;
        for (int q = 0; q < n_parts; ++q) {
            if (!parts[q]) fail("uspmv_halo_plan_create: part %d is NULL", q);
            if (parts[q]->cols_permuted) fail("uspmv_halo_plan_create: call before permute_scs_cols (main.cpp:1271-1308 order)");
            if (parts[q]->n_rows != s0->n_rows || parts[q]->C != s0->C || parts[q]->n_rows_padded != s0->n_rows_padded)
                fail("uspmv_halo_plan_create: the parts do not share n_rows / C");
            slots_total += parts[q]->n_elements;
        }
        // first-seen order is tracked as a 32-bit storage position over all parts; a single rank has no remote column at all
        if (P > 1 && slots_total > INT32_MAX - 1024) fail("uspmv_halo_plan_create: more than 2^31 stored elements on one rank");
        USPMV_CUDA(cudaSetDevice(s0->ctx->device));
        const int lo = wsa_h[rank], hi = wsa_h[rank + 1], n_glob = wsa_h[P];
        const long n_local = hi - lo;
        if (n_local != s0->n_rows) fail("uspmv_halo_plan_create: work_sharing_arr gives %ld local rows but the matrix has %ld", n_local, s0->n_rows);
        auto h = new uspmv_halo();
        try {
            h->ctx = s0->ctx; h->rank = rank; h->P = P; h->n_local = n_local;
            h->recv_cumsum.assign(P + 1, 0);
            h->send_ptr.assign(P + 1, 0);
            h->perm_d = x_permuted ? s0->old_to_new.p : nullptr;
            DevBuf<int> wsa_d(P + 1), first(n_glob > 0 ? n_glob : 1), flag(n_glob + 1), pos(n_glob + 1);
            USPMV_CUDA(cudaMemcpy(wsa_d.p, wsa_h, (P + 1) * sizeof(int), cudaMemcpyHostToDevice));
            if (n_glob) {
                k_fill_int<<<blocks_for(n_glob), TPB>>>(first.p, n_glob, INT32_MAX);
                USPMV_LAUNCH_CHECK();
            }
            const bool strict = options().strict_reference_halo;
            const long n_pad = s0->n_rows_padded;
            long slot_base = 0;
            for (int q = 0; q < n_parts && P > 1; ++q) {
                uspmv_scs *s = parts[q];
                if (s->n_elements) {
                    k_mark_remote<<<blocks_for(n_pad), TPB>>>(s->chunk_ptrs.p, s->chunk_lengths.p, s->row_lengths.p, s->col_idxs.p, n_pad, (int)s->C,
                                                             lo, hi, n_glob, strict, (int)slot_base, first.p);
                    USPMV_LAUNCH_CHECK();
                }
                slot_base += s->n_elements;
            }
            long n_halo = 0;
            if (n_glob && P > 1) {
                k_flag_seen<<<blocks_for(n_glob), TPB>>>(first.p, n_glob, flag.p);
                USPMV_LAUNCH_CHECK();
                USPMV_CUDA(cudaMemset(flag.p + n_glob, 0, sizeof(int)));
                size_t bytes = 0;
                USPMV_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, bytes, flag.p, pos.p, n_glob + 1));
                DevBuf<unsigned char> tmp(bytes);
                USPMV_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, bytes, flag.p, pos.p, n_glob + 1));
                g_launches.fetch_add(2);
                int tot = 0;
                USPMV_CUDA(cudaMemcpy(&tot, pos.p + n_glob, sizeof(int), cudaMemcpyDeviceToHost));
                n_halo = tot;
            }
            h->n_halo = n_halo;
            h->need_flat.assign(n_halo, 0);
            if (n_halo) {
                DevBuf<unsigned long long> keys(n_halo), keys_s(n_halo);
                DevBuf<int> cols(n_halo), cols_s(n_halo), need_d(n_halo), owner_d(n_halo);
                k_compact_remote<<<blocks_for(n_glob), TPB>>>(first.p, flag.p, pos.p, n_glob, wsa_d.p, P, keys.p, cols.p);
                USPMV_LAUNCH_CHECK();
                size_t bytes = 0;
                USPMV_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, keys.p, keys_s.p, cols.p, cols_s.p, (int)n_halo));
                DevBuf<unsigned char> tmp(bytes);
                USPMV_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, bytes, keys.p, keys_s.p, cols.p, cols_s.p, (int)n_halo));
                g_launches.fetch_add(8);
                // `first` is reused as the column -> halo id map
                k_assign_halo_ids<<<blocks_for(n_halo), TPB>>>(keys_s.p, cols_s.p, n_halo, (int)n_local, wsa_d.p, first.p, need_d.p, owner_d.p);
                USPMV_LAUNCH_CHECK();
                std::vector<int> owner_h(n_halo);
                USPMV_CUDA(cudaMemcpy(h->need_flat.data(), need_d.p, n_halo * sizeof(int), cudaMemcpyDeviceToHost));
                USPMV_CUDA(cudaMemcpy(owner_h.data(), owner_d.p, n_halo * sizeof(int), cudaMemcpyDeviceToHost));
                for (long k = 0; k < n_halo; ++k) h->recv_cumsum[owner_h[k] + 1]++;
                for (int p = 0; p < P; ++p) h->recv_cumsum[p + 1] += h->recv_cumsum[p];
            }
            for (int q = 0; q < n_parts; ++q) {
                uspmv_scs *s = parts[q];
                if (!s->n_elements) continue;
                k_rewrite_cols<<<blocks_for(n_pad), TPB>>>(s->chunk_ptrs.p, s->chunk_lengths.p, s->row_lengths.p, s->col_idxs.p, n_pad, (int)s->C,
                                                          lo, hi, n_glob, strict, first.p);
                USPMV_LAUNCH_CHECK();
            }
            for (int q = 0; q < n_parts; ++q) parts[q]->x_min_len = n_local + n_halo;  // columns are local / halo ids from here on
            USPMV_CUDA(cudaDeviceSynchronize());
        } catch (...) { delete h; throw; }
        *out = h;
    });
}

int uspmv_halo_plan_create(uspmv_scs *s, const int *wsa_h, int rank, int P, uspmv_halo **out) {
    if (!s) {
        set_error("uspmv_halo_plan_create: NULL argument");
        return 1;
    }
    return uspmv_halo_plan_create_multi(&s, 1, wsa_h, rank, P, 1, out);
}

int uspmv_halo_plan_counts(const uspmv_halo *h, int *recv_counts_cumsum_h, long *n_halo) {
    return guarded([&] {
        OptScope opt_scope(ctx_of(h));  // this context's options govern everything below
        if (!h) fail("uspmv_halo_plan_counts: plan is NULL");
        if (recv_counts_cumsum_h)
            for (int p = 0; p <= h->P; ++p) recv_counts_cumsum_h[p] = h->recv_cumsum[p];
        if (n_halo) *n_halo = h->n_halo;
    });
}

int uspmv_halo_plan_need(const uspmv_halo *h, int *need_flat_h, int *need_ptr_h) {
    return guarded([&] {
        OptScope opt_scope(ctx_of(h));  // this context's options govern everything below
        if (!h) fail("uspmv_halo_plan_need: plan is NULL");
        if (need_flat_h)
            for (long k = 0; k < h->n_halo; ++k) need_flat_h[k] = h->need_flat[k];
        if (need_ptr_h)
            for (int p = 0; p <= h->P; ++p) need_ptr_h[p] = h->recv_cumsum[p];
    });
}

int uspmv_halo_plan_set_send(uspmv_halo *h, const int *send_flat_h, const int *send_ptr_h) {
    return guarded([&] {
        OptScope opt_scope(ctx_of(h));  // this context's options govern everything below
        if (!h || !send_ptr_h) fail("uspmv_halo_plan_set_send: NULL argument");
        USPMV_CUDA(cudaSetDevice(h->ctx->device));
        for (int p = 0; p <= h->P; ++p) h->send_ptr[p] = send_ptr_h[p];
        h->n_send = send_ptr_h[h->P];
        for (long i = 0; i < h->n_send; ++i)
            if (send_flat_h[i] < 0 || send_flat_h[i] >= h->n_local) fail("uspmv_halo_plan_set_send: send index %d outside the local rows", send_flat_h[i]);
        h->send_idx.alloc(h->n_send);
        if (h->n_send) USPMV_CUDA(cudaMemcpy(h->send_idx.p, send_flat_h, h->n_send * sizeof(int), cudaMemcpyHostToDevice));
    });
}

int uspmv_halo_pack(const uspmv_halo *h, const void *x, void *sendbuf, int vt, int bvs, long vec_length, int layout, void *stream) {
    return guarded([&] {
        OptScope opt_scope(ctx_of(h));  // this context's options govern everything below
        if (!h) fail("uspmv_halo_pack: plan is NULL");
        if (h->n_send == 0) return;
        if (!x || !sendbuf) fail("uspmv_halo_pack: NULL buffer");
        if (bvs < 1) fail("uspmv_halo_pack: block_vec_size must be >= 1");
        cudaStream_t st = as_stream(stream);
        const unsigned g = blocks_for(h->n_send * bvs);
        switch (vt) {
        case USPMV_F64: k_pack<double><<<g, TPB, 0, st>>>(h->send_idx.p, h->perm_d, h->n_send, (const double *)x, (double *)sendbuf, bvs, vec_length, layout); break;
        case USPMV_F32: k_pack<float><<<g, TPB, 0, st>>>(h->send_idx.p, h->perm_d, h->n_send, (const float *)x, (float *)sendbuf, bvs, vec_length, layout); break;
        case USPMV_F16: k_pack<__half><<<g, TPB, 0, st>>>(h->send_idx.p, h->perm_d, h->n_send, (const __half *)x, (__half *)sendbuf, bvs, vec_length, layout); break;
        default: fail("uspmv_halo_pack: invalid value type %d", vt);
        }
        USPMV_LAUNCH_CHECK();
    });
}

void uspmv_halo_destroy(uspmv_halo *h) { delete h; }

}  // extern "C"

// =====================================================================================================
// NVLink peer-to-peer halo exchange (one process per GPU, CUDA IPC): the pack kernel stores straight into
// the neighbours' x tails and signals with epoch flags — no NCCL call, no host round trip in the SpMV loop.
//
//   arena (one cudaMalloc, one IPC handle per rank):  [ x storage | arrived[P] | acked[P] | epoch | error ]
//   step e (= epoch + 1) on rank r
//     comm stream : k_p2p_push   wait acked[q] >= e-1 (receiver q consumed the previous halo), gather
//                                x[perm[send_idx]] and store into peer q's x tail, fence.sys, last CTA sets
//                                peer_q.arrived[r] = e
//                   k_p2p_wait   spin until arrived[p] >= e for every sender p
//     main stream : interior SpMV  ||  (comm stream)  ... then boundary SpMV, then
//                   k_p2p_ack    peer_p.acked[r] = e for every sender p; epoch = e
// Spins are bounded (globaltimer, 10 s) and raise the arena's error word instead of hanging the GPU.
// Replaces MPI_Isend/Irecv/Waitall of init/finalize_halo_exchange (classes_structs.hpp:857-995).
// =====================================================================================================
struct uspmv_p2p {
    uspmv_halo *plan = nullptr;
    int vt = USPMV_F64;
    int n_buf = 1, bvs = 1, layout = USPMV_COLWISE;
    long vec_length = 0;             // elements per vector: n_local + max(padding, halo)
    long x_len = 0;                  // elements per buffer = bvs * vec_length
    size_t x_bytes = 0, arena_bytes = 0;
    unsigned char *arena = nullptr;  // local: [ buffer 0 | ... | buffer n_buf-1 | arrived[P] | acked[P] | epoch | error ]
    unsigned int *arrived = nullptr, *acked = nullptr, *epoch = nullptr, *error = nullptr;
    std::vector<unsigned char *> peer_arena;  // P entries (NULL for self / unused)
    DevBuf<unsigned long long> peer_x_dst;    // [buffer][peer q]: device address where OUR elements for q start (bvs = 1)
    DevBuf<unsigned long long> peer_x0;       // [buffer][peer q]: start of q's buffer
    DevBuf<long> peer_base, peer_ld;          // per peer q: element offset of our elements in q's vectors; q's vec_length
    DevBuf<unsigned long long> peer_arrived;  // per peer q: address of q.arrived[rank]
    DevBuf<unsigned long long> peer_acked;    // per peer p: address of p.acked[rank]
    DevBuf<int> send_ptr_d, is_sender_d, is_receiver_d;
    DevBuf<unsigned int> block_counter;
    DevBuf<unsigned char> scratch_y;          // solve loop on the multi-kernel paths: y is staged here, n_rows are copied
    cudaEvent_t ev_main = nullptr, ev_comm = nullptr;
    bool connected = false;
    int mode = 2;  // 0: exchange, then one full SpMV; 1: push/wait kernels next to the interior kernel; 2: ONE fused kernel (C = 32)
    DevBuf<unsigned int> fused_counters;
    // pipelined host-buffer steps (uspmv_p2p_spmv_host_submit / _wait): two slots = the two arena buffers
    cudaStream_t s_h2d = nullptr, s_run = nullptr, s_d2h = nullptr, s_comm = nullptr;
    cudaEvent_t ev_x[2] = {nullptr, nullptr}, ev_y[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr};
    DevBuf<unsigned char> slot_y[2];
    bool slot_busy[2] = {false, false};
    unsigned int *err_h = nullptr;   // pinned: the arena's error word as of the end of each slot's step
    long host_submits = 0;
    long n_push_tiles = 0, n_push_tiles_4k = 0;  // tiles (2048 / 4096 elements) of the large-halo push kernels (k_p2p_push_tiles)
    unsigned char *buffer(int b) const { return arena + (size_t)b * x_bytes; }
};

const uspmv_ctx *ctx_of(const uspmv_halo *h) { return h ? h->ctx : nullptr; }
const uspmv_ctx *ctx_of(const uspmv_p2p *p) { return p && p->plan ? p->plan->ctx : nullptr; }

namespace {

constexpr int USPMV_PUSH_DEFAULT_VARIANT = 1;

__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ unsigned int ld_volatile_sys(const unsigned int *p) {
    unsigned int v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__device__ bool spin_until_ge(const unsigned int *flag, unsigned int target, unsigned int *error) {
    const unsigned long long t0 = globaltimer_ns();
    while (ld_volatile_sys(flag) < target) {
        if (globaltimer_ns() - t0 > 10000000000ull) {
            atomicExch(error, 1u);
            return false;
        }
        __nanosleep(200);
    }
    return true;
}

// Gathers x[perm[send_idx]] (all bvs vectors of a block vector) and stores it straight into the neighbours' vectors over NVLink.
// Row-major block vectors: element (i, v) of peer q lands at q.X[(base_q + i) * bvs + v]; column-major: q.X[base_q + i + v * ld_q]
// (the reference sends one strided / per-vector message per neighbour instead, classes_structs.hpp:909-970).
template <typename VT>
__global__ void __launch_bounds__(256)
k_p2p_push(int P, int my_rank, const int *__restrict__ send_ptr, const int *__restrict__ is_receiver, const int *__restrict__ send_idx,
           const int *__restrict__ perm, const VT *__restrict__ x, const unsigned long long *__restrict__ peer_x0,
           const long *__restrict__ peer_base, const long *__restrict__ peer_ld, int bvs, int layout, long ld,
           const unsigned long long *__restrict__ peer_arrived, const unsigned int *acked, const unsigned int *epoch, unsigned int *error,
           unsigned int *block_counter) {
    __shared__ unsigned int s_last;
    const unsigned int e = *epoch + 1u;
    // (1) receivers must have consumed the halo of step e-1 before it is overwritten
    if (threadIdx.x < P && is_receiver[threadIdx.x]) spin_until_ge(acked + threadIdx.x, e - 1u, error);
    __syncthreads();
    // (2) gather + store over NVLink
    const long n_send = send_ptr[P];
    const long total = n_send * bvs;
    for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < total; t += (long)gridDim.x * blockDim.x) {
        long i, v;
        if (layout == USPMV_ROWWISE) { i = t / bvs; v = t - i * bvs; }
        else { v = t / n_send; i = t - v * n_send; }
        int q = 0;
        while (i >= send_ptr[q + 1]) ++q;
        const long src = perm ? perm[send_idx[i]] : send_idx[i];  // AP: x stays in the original row order
        const long k = peer_base[q] + (i - send_ptr[q]);
        VT *dst = reinterpret_cast<VT *>(peer_x0[q]);
        if (layout == USPMV_ROWWISE) dst[k * bvs + v] = x[src * bvs + v];
        else dst[k + v * peer_ld[q]] = x[src + v * ld];
    }
    // (3) publish: all stores of all CTAs are system-visible before the flag
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(block_counter, 1u) == gridDim.x - 1);
    __syncthreads();
    if (s_last) {
        __threadfence_system();
        if (threadIdx.x < P && is_receiver[threadIdx.x]) {
            volatile unsigned int *f = reinterpret_cast<volatile unsigned int *>(peer_arrived[threadIdx.x]);
            *f = e;
        }
        if (threadIdx.x == 0) *block_counter = 0;
        __threadfence_system();
    }
}

// ---- large single-vector halos ------------------------------------------------------------------------------------------------
// Power-law matrices exchange ~100 MB per rank and SpMV (BASELINE config 4 at 8 GPUs: 14.7 M elements per rank); one dependent
// idx -> x -> peer-store chain per thread leaves NVLink mostly idle.  Here the send list is cut into tiles of PUSH_TILE elements
// that never cross a peer and start, in the PEER's vector, at multiples of PUSH_TILE (so every tile but the edges of a peer's
// range is 16-byte aligned at the destination); a CTA gathers a tile with PUSH_TILE/256 independent loads per thread and
//   MODE 1: stores it straight to the peer (8 or 16 stores in flight per thread; the default — measured fastest at 2, 4 and 8 GPUs);
//   MODE 2: lands it in shared memory and lets the bulk-copy engine write the 16-byte-aligned middle of the tile over NVLink
//           (cp.async.bulk.global.shared::cta, double-buffered), the unaligned edge elements go by plain stores.
inline long push_tile_count(const std::vector<int> &send_ptr, const std::vector<long> &peer_base, int P, long PUSH_TILE) {
    long t = 0;
    for (int q = 0; q < P; ++q) {
        const long nq = send_ptr[q + 1] - send_ptr[q];
        if (nq > 0) t += (peer_base[q] + nq - 1) / PUSH_TILE - peer_base[q] / PUSH_TILE + 1;
    }
    return t;
}

template <typename VT, int MODE, int PUSH_TILE>
__global__ void __launch_bounds__(256)
k_p2p_push_tiles(int P, const int *__restrict__ send_ptr, const int *__restrict__ is_receiver, const int *__restrict__ send_idx,
                 const int *__restrict__ perm, const VT *__restrict__ x, const unsigned long long *__restrict__ peer_x0,
                 const long *__restrict__ peer_base, const unsigned long long *__restrict__ peer_arrived, const unsigned int *acked,
                 const unsigned int *epoch, unsigned int *error, unsigned int *block_counter) {
    constexpr int U = PUSH_TILE / 256;
    constexpr long A = 16 / (long)sizeof(VT);  // elements per 16 bytes
    extern __shared__ __align__(128) unsigned char push_smem[];  // MODE 2: two tiles
    __shared__ int s_tile_ptr[257];
    __shared__ unsigned int s_last;
    const int tid = threadIdx.x;
    const unsigned int e = *epoch + 1u;
    if (tid == 0) {
        int t = 0;
        s_tile_ptr[0] = 0;
        for (int q = 0; q < P; ++q) {
            const long nq = send_ptr[q + 1] - send_ptr[q], b = peer_base[q];
            if (nq > 0) t += (int)((b + nq - 1) / PUSH_TILE - b / PUSH_TILE + 1);
            s_tile_ptr[q + 1] = t;
        }
    }
    if (tid < P && is_receiver[tid]) spin_until_ge(acked + tid, e - 1u, error);
    __syncthreads();
    const int total_tiles = s_tile_ptr[P];
    int q = 0, it = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
        while (t >= s_tile_ptr[q + 1]) ++q;
        const long base = peer_base[q];
        const int s0 = send_ptr[q], nq = send_ptr[q + 1] - s0;
        const long k0 = (base / PUSH_TILE + (t - s_tile_ptr[q])) * PUSH_TILE;  // destination index of the tile's first slot
        const long k_lo = k0 > base ? k0 : base;
        const long k_hi = k0 + PUSH_TILE < base + nq ? k0 + PUSH_TILE : base + nq;
        VT *dst = reinterpret_cast<VT *>(peer_x0[q]);
        VT *buf = reinterpret_cast<VT *>(push_smem) + (size_t)(it & 1) * PUSH_TILE;
        long ka = (k_lo + A - 1) / A * A, kb = k_hi / A * A;
        if (kb < ka) kb = ka;
        if (MODE == 2) {
            // the bulk copy issued two tiles ago has finished READING this buffer
            if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            __syncthreads();
        }
        int src[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long k = k0 + tid + u * 256;
            src[u] = (k >= k_lo && k < k_hi) ? send_idx[s0 + (k - base)] : -1;
        }
        if (perm) {
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (src[u] >= 0) src[u] = perm[src[u]];
        }
        VT v[U];
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (src[u] >= 0) v[u] = x[src[u]];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long k = k0 + tid + u * 256;
            if (src[u] < 0) continue;
            if (MODE == 2 && k >= ka && k < kb) buf[k - k0] = v[u];
            else dst[k] = v[u];
        }
        if (MODE == 2) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the bulk-copy engine
            __syncthreads();
            if (tid == 0) {
                if (kb > ka) {
                    const unsigned int s_addr = (unsigned int)__cvta_generic_to_shared(buf + (ka - k0));
                    const unsigned int bytes = (unsigned int)((kb - ka) * (long)sizeof(VT));
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + ka), "r"(s_addr), "r"(bytes)
                                 : "memory");
                }
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        }
    }
    if (MODE == 2 && tid == 0) {
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // every bulk write of this CTA has been performed
        asm volatile("fence.proxy.async;" ::: "memory");
    }
    __threadfence_system();
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(block_counter, 1u) == gridDim.x - 1);
    __syncthreads();
    if (s_last) {
        __threadfence_system();
        if (tid < P && is_receiver[tid]) {
            volatile unsigned int *f = reinterpret_cast<volatile unsigned int *>(peer_arrived[tid]);
            *f = e;
        }
        if (tid == 0) *block_counter = 0;
        __threadfence_system();
    }
}

__global__ void k_p2p_wait(int P, const int *__restrict__ is_sender, const unsigned int *arrived, const unsigned int *epoch, unsigned int *error) {
    const unsigned int e = *epoch + 1u;
    if (threadIdx.x < P && is_sender[threadIdx.x]) spin_until_ge(arrived + threadIdx.x, e, error);
    __threadfence_system();
}

__global__ void k_p2p_ack(int P, const int *__restrict__ is_sender, const unsigned long long *__restrict__ peer_acked, unsigned int *epoch) {
    const unsigned int e = *epoch + 1u;
    __syncthreads();
    if (threadIdx.x < P && is_sender[threadIdx.x]) {
        volatile unsigned int *f = reinterpret_cast<volatile unsigned int *>(peer_acked[threadIdx.x]);
        *f = e;
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) *epoch = e;
}

void launch_push(uspmv_p2p *p, const void *x, int buf, cudaStream_t comm) {
    uspmv_halo *h = p->plan;
    const int P = h->P;
    const long total = h->n_send * p->bvs;
    const unsigned long long *x0 = p->peer_x0.p + (size_t)buf * P;
    // large single-vector halos: tiled gather with many loads in flight per thread (see k_p2p_push_tiles)
    int variant = options().push_variant;
    if (variant < 0) variant = USPMV_PUSH_DEFAULT_VARIANT;
    if (p->bvs == 1 && total >= options().push_min_elements && variant != 0 && p->n_push_tiles > 0) {
        const int per_sm = options().push_ctas_per_sm > 0 ? options().push_ctas_per_sm : 8;
        const long tiles = variant == 3 ? p->n_push_tiles_4k : p->n_push_tiles;
        long g = std::min<long>(tiles, (long)per_sm * sm_count(h->ctx->device));
#define USPMV_PUSH_T(T, M, TILE, SH)                                                                                                     \
    k_p2p_push_tiles<T, M, TILE><<<(unsigned)g, 256, SH, comm>>>(P, p->send_ptr_d.p, p->is_receiver_d.p, h->send_idx.p, h->perm_d,        \
                                                                 (const T *)x, x0, p->peer_base.p, p->peer_arrived.p, p->acked, p->epoch, \
                                                                 p->error, p->block_counter.p)
#define USPMV_PUSH_TV(T)                                                        \
    do {                                                                        \
        if (variant == 2) USPMV_PUSH_T(T, 2, 2048, 2 * 2048 * sizeof(T));       \
        else if (variant == 3) USPMV_PUSH_T(T, 1, 4096, 0);                     \
        else USPMV_PUSH_T(T, 1, 2048, 0);                                       \
    } while (0)
        switch (p->vt) {
        case USPMV_F64: USPMV_PUSH_TV(double); break;
        case USPMV_F32: USPMV_PUSH_TV(float); break;
        default: USPMV_PUSH_TV(__half);
        }
#undef USPMV_PUSH_TV
#undef USPMV_PUSH_T
        USPMV_LAUNCH_CHECK();
        return;
    }
    // small halos (stencil faces): a few CTAs next to the interior kernel
    long g = (total + 2047) / 2048;
    if (g < 1) g = 1;
    const long cap = total >= (1L << 20) ? 4L * sm_count(h->ctx->device) : 64;
    if (g > cap) g = cap;
#define USPMV_PUSH(T)                                                                                                                     \
    k_p2p_push<T><<<(unsigned)g, 256, 0, comm>>>(P, h->rank, p->send_ptr_d.p, p->is_receiver_d.p, h->send_idx.p, h->perm_d, (const T *)x, x0, \
                                                 p->peer_base.p, p->peer_ld.p, p->bvs, p->layout, p->vec_length, p->peer_arrived.p, p->acked,  \
                                                 p->epoch, p->error, p->block_counter.p)
    switch (p->vt) {
    case USPMV_F64: USPMV_PUSH(double); break;
    case USPMV_F32: USPMV_PUSH(float); break;
    default: USPMV_PUSH(__half);
    }
#undef USPMV_PUSH
    USPMV_LAUNCH_CHECK();
}

}  // namespace

extern "C" {

/* Arena with n_buf buffers of bvs vectors of vec_length elements each (bvs = 1, n_buf = 1: uspmv_p2p_create).
 * x_d receives the n_buf buffer addresses.  Two buffers make the solve loop's swap a pointer swap (SpmvKernel::swap_local_vectors,
 * classes_structs.hpp:1130-1165): step k reads buffer k & 1 and writes y into the other one. */
int uspmv_p2p_create_ex(uspmv_halo *plan, int vt, long vec_length, int bvs, int layout, int n_buf, uspmv_p2p **out, void *ipc_handle64,
                        void **x_d) {
    return guarded([&] {
        OptScope opt_scope(ctx_of(plan));  // this context's options govern everything below
        if (!plan || !out || !ipc_handle64 || !x_d) fail("uspmv_p2p_create: NULL argument");
        if (plan->P > 256) fail("uspmv_p2p_create: at most 256 ranks");
        if (bvs < 1 || bvs > 16) fail("uspmv_p2p_create: block_vec_size must be in [1,16] (got %d)", bvs);
        if (n_buf < 1 || n_buf > 2) fail("uspmv_p2p_create: 1 or 2 buffers");
        if (layout != USPMV_COLWISE && layout != USPMV_ROWWISE) fail("uspmv_p2p_create: invalid layout %d", layout);
        if (vec_length < plan->n_local + plan->n_halo)
            fail("uspmv_p2p_create: vec_length %ld < n_local + n_halo = %ld", vec_length, plan->n_local + plan->n_halo);
        USPMV_CUDA(cudaSetDevice(plan->ctx->device));
        static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
        auto p = new uspmv_p2p();
        try {
            p->plan = plan; p->vt = vt; p->vec_length = vec_length; p->bvs = bvs; p->layout = layout; p->n_buf = n_buf;
            p->x_len = vec_length * bvs;
            const size_t es = vt_size(vt);
            p->x_bytes = ((size_t)p->x_len * es + 255) / 256 * 256;
            const int P = plan->P;
            p->arena_bytes = (size_t)n_buf * p->x_bytes + (2 * (size_t)P + 2) * sizeof(unsigned int) + 256;
            USPMV_CUDA(cudaMalloc(&p->arena, p->arena_bytes));
            USPMV_CUDA(cudaMemset(p->arena, 0, p->arena_bytes));
            p->arrived = reinterpret_cast<unsigned int *>(p->arena + (size_t)n_buf * p->x_bytes);
            p->acked = p->arrived + P;
            p->epoch = p->acked + P;
            p->error = p->epoch + 1;
            p->block_counter.alloc(1);
            USPMV_CUDA(cudaMemset(p->block_counter.p, 0, sizeof(unsigned int)));
            p->fused_counters.alloc(4);
            USPMV_CUDA(cudaMemset(p->fused_counters.p, 0, 4 * sizeof(unsigned int)));
            cudaIpcMemHandle_t h;
            USPMV_CUDA(cudaIpcGetMemHandle(&h, p->arena));
            std::memcpy(ipc_handle64, &h, 64);
            USPMV_CUDA(cudaEventCreateWithFlags(&p->ev_main, cudaEventDisableTiming));
            USPMV_CUDA(cudaEventCreateWithFlags(&p->ev_comm, cudaEventDisableTiming));
            USPMV_CUDA(cudaDeviceSynchronize());
        } catch (...) { if (p->arena) cudaFree(p->arena); delete p; throw; }
        for (int b = 0; b < n_buf; ++b) x_d[b] = p->buffer(b);
        *out = p;
    });
}

int uspmv_p2p_create(uspmv_halo *plan, int vt, long x_len, uspmv_p2p **out, void *ipc_handle64, void **x_d) {
    return uspmv_p2p_create_ex(plan, vt, x_len, 1, USPMV_COLWISE, 1, out, ipc_handle64, x_d);
}

/* all_handles: P x 64 bytes (IPC handle of every rank's arena); peer_x_bytes[q]: size of ONE buffer of q in bytes (every rank uses
 * the same number of buffers); peer_halo_base[q]: element offset in q's vectors where this rank's elements start
 * (q.n_local + q.recv_cumsum[rank]); peer_vec_length[q]: q's vec_length (column-major block vectors; may be NULL for bvs = 1). */
int uspmv_p2p_connect_ex(uspmv_p2p *p, const void *all_handles, const long *peer_x_bytes, const long *peer_halo_base,
                         const long *peer_vec_length) {
    return guarded([&] {
        OptScope opt_scope(ctx_of(p));  // this context's options govern everything below
        if (!p || !all_handles || !peer_x_bytes || !peer_halo_base) fail("uspmv_p2p_connect: NULL argument");
        if (!peer_vec_length && p->bvs > 1 && p->layout == USPMV_COLWISE) fail("uspmv_p2p_connect: column-major block vectors need peer_vec_length");
        uspmv_halo *h = p->plan;
        USPMV_CUDA(cudaSetDevice(h->ctx->device));
        const int P = h->P, me = h->rank, nb = p->n_buf;
        const size_t es = vt_size(p->vt);
        p->peer_arena.assign(P, nullptr);
        std::vector<unsigned long long> dst((size_t)nb * P, 0), x0((size_t)nb * P, 0), arr(P, 0), ack(P, 0);
        std::vector<long> base(P, 0), pld(P, 0);
        std::vector<int> is_s(P, 0), is_r(P, 0);
        for (int q = 0; q < P; ++q) {
            is_s[q] = h->recv_cumsum[q + 1] > h->recv_cumsum[q];
            is_r[q] = h->send_ptr[q + 1] > h->send_ptr[q];
            if (q == me || !(is_s[q] || is_r[q])) continue;
            cudaIpcMemHandle_t ih;
            std::memcpy(&ih, static_cast<const unsigned char *>(all_handles) + (size_t)q * 64, 64);
            void *ptr = nullptr;
            USPMV_CUDA(cudaIpcOpenMemHandle(&ptr, ih, cudaIpcMemLazyEnablePeerAccess));
            p->peer_arena[q] = static_cast<unsigned char *>(ptr);
            unsigned int *q_arrived = reinterpret_cast<unsigned int *>(p->peer_arena[q] + (size_t)nb * peer_x_bytes[q]);
            unsigned int *q_acked = q_arrived + P;
            for (int b = 0; b < nb; ++b) {
                unsigned char *qb = p->peer_arena[q] + (size_t)b * peer_x_bytes[q];
                x0[(size_t)b * P + q] = reinterpret_cast<unsigned long long>(qb);
                dst[(size_t)b * P + q] = reinterpret_cast<unsigned long long>(qb + (size_t)peer_halo_base[q] * es);
            }
            base[q] = peer_halo_base[q];
            pld[q] = peer_vec_length ? peer_vec_length[q] : 0;
            arr[q] = reinterpret_cast<unsigned long long>(q_arrived + me);
            ack[q] = reinterpret_cast<unsigned long long>(q_acked + me);
        }
        p->peer_x_dst.alloc((size_t)nb * P); p->peer_x0.alloc((size_t)nb * P); p->peer_base.alloc(P); p->peer_ld.alloc(P);
        p->peer_arrived.alloc(P); p->peer_acked.alloc(P);
        p->send_ptr_d.alloc(P + 1); p->is_sender_d.alloc(P); p->is_receiver_d.alloc(P);
        USPMV_CUDA(cudaMemcpy(p->peer_x_dst.p, dst.data(), dst.size() * 8, cudaMemcpyHostToDevice));
        USPMV_CUDA(cudaMemcpy(p->peer_x0.p, x0.data(), x0.size() * 8, cudaMemcpyHostToDevice));
        USPMV_CUDA(cudaMemcpy(p->peer_base.p, base.data(), P * sizeof(long), cudaMemcpyHostToDevice));
        USPMV_CUDA(cudaMemcpy(p->peer_ld.p, pld.data(), P * sizeof(long), cudaMemcpyHostToDevice));
        USPMV_CUDA(cudaMemcpy(p->peer_arrived.p, arr.data(), P * 8, cudaMemcpyHostToDevice));
        USPMV_CUDA(cudaMemcpy(p->peer_acked.p, ack.data(), P * 8, cudaMemcpyHostToDevice));
        USPMV_CUDA(cudaMemcpy(p->send_ptr_d.p, h->send_ptr.data(), (P + 1) * sizeof(int), cudaMemcpyHostToDevice));
        USPMV_CUDA(cudaMemcpy(p->is_sender_d.p, is_s.data(), P * sizeof(int), cudaMemcpyHostToDevice));
        USPMV_CUDA(cudaMemcpy(p->is_receiver_d.p, is_r.data(), P * sizeof(int), cudaMemcpyHostToDevice));
        p->n_push_tiles = push_tile_count(h->send_ptr, base, P, 2048);
        p->n_push_tiles_4k = push_tile_count(h->send_ptr, base, P, 4096);
        p->connected = true;
    });
}

int uspmv_p2p_connect(uspmv_p2p *p, const void *all_handles, const long *peer_x_bytes, const long *peer_halo_base) {
    return uspmv_p2p_connect_ex(p, all_handles, peer_x_bytes, peer_halo_base, nullptr);
}

/* One distributed SpMV step (comm_halos = 1): x = buffer x_buf of the arena.  y goes to y_d, or — y_d NULL, y_buf >= 0 — into
 * buffer y_buf (rows < n_rows only, the tail of that buffer is the next step's halo), which makes `rev` x { SpMV ; swap } a
 * device-resident loop with no copy (solve mode, main.cpp:528-631).  Overlap modes as set by uspmv_p2p_set_overlap. */
int uspmv_p2p_spmv_buf(uspmv_p2p *p, const uspmv_scs *scs, int x_buf, int y_buf, void *y_d, void *stream, void *comm_stream) {
    return guarded([&] {
        OptScope opt_scope(ctx_of(p));  // this context's options govern everything below
        if (!p || !scs) fail("uspmv_p2p_spmv: NULL argument");
        if (!p->connected) fail("uspmv_p2p_spmv: call uspmv_p2p_connect first");
        if (!scs->chunks_split) fail("uspmv_p2p_spmv: call uspmv_scs_split_chunks first");
        if (p->bvs != 1) fail("uspmv_p2p_spmv: the arena was created for block vectors; use uspmv_p2p_spmmv");
        if (!p->plan->perm_d) fail("uspmv_p2p_spmv: the halo plan was built for an un-permuted x (adaptive precision); use uspmv_p2p_ap_spmv");
        if (x_buf < 0 || x_buf >= p->n_buf) fail("uspmv_p2p_spmv: x buffer %d out of range", x_buf);
        const bool to_buf = y_d == nullptr;
        if (to_buf && (y_buf < 0 || y_buf >= p->n_buf || y_buf == x_buf)) fail("uspmv_p2p_spmv: y needs a device pointer or another buffer of the arena");
        if (to_buf && scs->n_rows_padded > p->x_len) fail("uspmv_p2p_spmv: buffer shorter than n_rows_padded");
        uspmv_halo *h = p->plan;
        cudaStream_t main = as_stream(stream), comm = as_stream(comm_stream);
        const int P = h->P;
        const void *x = p->buffer(x_buf);
        // The fused one-kernel step where it wins.  C = 32 in sp / hp does not: the fused instance of the kernel executes 32 % more
        // instructions per chunk than the single-GPU loop and the narrow types are issue-bound (scs_stream.cuh, k_scs32_stream), so
        // they take the push / wait kernels next to the unchanged interior kernel — N = 2, 256^3 slab per rank: sp 0.175 vs 0.210 ms,
        // hp 0.163 vs 0.203 ms; dp the other way round, 0.253 (fused) vs 0.276 (profiles/r02I_n2_*.json).  "fused_narrow" = 1 forces it.
        const bool fused_pays = scs->C != 32 || p->vt == USPMV_F64 || options().fused_narrow;
        if (p->mode == 2 && fused_pays && scs_fused_supported(scs) && P <= 32) {
            stream::FusedArgs fa{};
            fa.n_int = (long)scs->interior_chunks.n;
            fa.n_bnd = (long)scs->boundary_chunks.n;
            fa.int_list = scs->interior_contig ? nullptr : scs->interior_chunks.p;
            fa.bnd_list = scs->boundary_contig ? nullptr : scs->boundary_chunks.p;
            fa.int_off = scs->interior_off;
            fa.bnd_off = scs->boundary_off;
            fa.P = P; fa.my_rank = h->rank; fa.n_send = h->n_send;
            fa.send_ptr = p->send_ptr_d.p; fa.is_receiver = p->is_receiver_d.p; fa.is_sender = p->is_sender_d.p;
            fa.send_idx = h->send_idx.p; fa.perm = h->perm_d;
            fa.peer_x0 = p->peer_x0.p + (size_t)x_buf * P; fa.peer_base = p->peer_base.p; fa.peer_ld = p->peer_ld.p;
            fa.bvs = 1; fa.layout = USPMV_COLWISE; fa.ld = p->vec_length;
            fa.peer_arrived = p->peer_arrived.p; fa.peer_acked = p->peer_acked.p;
            fa.acked = p->acked; fa.arrived = p->arrived; fa.epoch = p->epoch; fa.error = p->error; fa.counters = p->fused_counters.p;
            fa.y_rows = (int)(to_buf ? scs->n_rows : scs->n_rows_padded);
            launch_scs32_fused(scs, x, to_buf ? (void *)p->buffer(y_buf) : y_d, fa, main);
            return;
        }
        void *y = y_d;
        if (to_buf) {  // the multi-kernel paths store every padded row: stage y, then copy the real rows
            const size_t need = (size_t)scs->n_rows_padded * vt_size(p->vt);
            if (p->scratch_y.n < need) p->scratch_y.alloc(need);
            y = p->scratch_y.p;
        }
        const bool overlap = p->mode != 0;
        USPMV_CUDA(cudaEventRecord(p->ev_main, main));
        // the persistent interior kernel is launched FIRST so that its CTAs get their SM slots; the (small) push /
        // wait kernels then run next to it in the leftover registers instead of delaying some of its CTAs
        if (overlap && uspmv_spmv_part(scs, 1, x, y, stream)) throw Error(uspmv_last_error());
        USPMV_CUDA(cudaStreamWaitEvent(comm, p->ev_main, 0));
        launch_push(p, x, x_buf, comm);
        k_p2p_wait<<<1, 256, 0, comm>>>(P, p->is_sender_d.p, p->arrived, p->epoch, p->error);
        USPMV_LAUNCH_CHECK();
        USPMV_CUDA(cudaEventRecord(p->ev_comm, comm));
        USPMV_CUDA(cudaStreamWaitEvent(main, p->ev_comm, 0));
        if (overlap) {
            if (uspmv_spmv_part(scs, 2, x, y, stream)) throw Error(uspmv_last_error());
        } else {
            if (uspmv_spmv(scs, x, y, stream)) throw Error(uspmv_last_error());
        }
        k_p2p_ack<<<1, 256, 0, main>>>(P, p->is_sender_d.p, p->peer_acked.p, p->epoch);
        USPMV_LAUNCH_CHECK();
        if (to_buf) USPMV_CUDA(cudaMemcpyAsync(p->buffer(y_buf), y, (size_t)scs->n_rows * vt_size(p->vt), cudaMemcpyDeviceToDevice, main));
    });
}

int uspmv_p2p_spmv(uspmv_p2p *p, const uspmv_scs *scs, void *y_d, void *stream, void *comm_stream) {
    if (!y_d) {
        set_error("uspmv_p2p_spmv: y is NULL");
        return 1;
    }
    return uspmv_p2p_spmv_buf(p, scs, 0, -1, y_d, stream, comm_stream);
}

/* One distributed SpMMV step over the arena's block vector (buffer x_buf): every neighbour receives all block_vec_size vectors
 * of its halo rows in ONE push (the reference's bulkvec mode, classes_structs.hpp:909-923,962-970), overlapped with the interior
 * chunks when the streamed kernel applies (C = 32, block_vec_size 2/4/8/16); otherwise exchange first, then one full SpMMV. */
int uspmv_p2p_spmmv(uspmv_p2p *p, const uspmv_scs *scs, int x_buf, void *Y_d, void *stream, void *comm_stream) {
    return guarded([&] {
        OptScope opt_scope(ctx_of(p));  // this context's options govern everything below
        if (!p || !scs || !Y_d) fail("uspmv_p2p_spmmv: NULL argument");
        if (!p->connected) fail("uspmv_p2p_spmmv: call uspmv_p2p_connect first");
        if (!scs->chunks_split) fail("uspmv_p2p_spmmv: call uspmv_scs_split_chunks first");
        if (x_buf < 0 || x_buf >= p->n_buf) fail("uspmv_p2p_spmmv: x buffer %d out of range", x_buf);
        uspmv_halo *h = p->plan;
        cudaStream_t main = as_stream(stream), comm = as_stream(comm_stream);
        const int P = h->P;
        const void *X = p->buffer(x_buf);
        if (h->n_send == 0 && h->n_halo == 0 && options().mmv_fused_rowwise < 2) {  // no neighbour at all (one rank): the tuned single-GPU kernel, then close the epoch
            if (uspmv_spmmv(scs, X, Y_d, p->bvs, p->vec_length, p->layout, stream)) throw Error(uspmv_last_error());
            k_p2p_ack<<<1, 256, 0, main>>>(P, p->is_sender_d.p, p->peer_acked.p, p->epoch);
            USPMV_LAUNCH_CHECK();
            return;
        }
        // ONE fused kernel (push, interior, publish, wait, boundary, ack) by default for both layouts.  N = 2, 256^3 slab per rank, B200
        // (profiles/r02A_dist_probe_mmv.txt; us per step: fused / push + wait kernels next to the interior kernel / local kernel):
        // row-major dp bvs 4 418 / 440 / 365, dp bvs 8 686 / 708 / 586, sp bvs 8 363 / 419 / 325, sp bvs 16 663 / 674 / 595,
        // sp bvs 4 282 / 280 / 229, dp bvs 2 314 / 312 / 295; column-major dp bvs 4 483 / 503.  "mmv_fused_rowwise" = 0 selects the
        // multi-kernel overlap for row-major block vectors (A/B runs, tests).
        const bool fused_wanted = p->mode == 2 && (p->layout == USPMV_COLWISE || options().mmv_fused_rowwise);
        if (fused_wanted && spmmv_fused_supported(scs, p->bvs) && P <= 32) {
            stream::FusedArgs fa{};
            fa.n_int = (long)scs->interior_chunks.n;
            fa.n_bnd = (long)scs->boundary_chunks.n;
            fa.int_list = scs->interior_contig ? nullptr : scs->interior_chunks.p;
            fa.bnd_list = scs->boundary_contig ? nullptr : scs->boundary_chunks.p;
            fa.int_off = scs->interior_off;
            fa.bnd_off = scs->boundary_off;
            fa.P = P; fa.my_rank = h->rank; fa.n_send = h->n_send;
            fa.send_ptr = p->send_ptr_d.p; fa.is_receiver = p->is_receiver_d.p; fa.is_sender = p->is_sender_d.p;
            fa.send_idx = h->send_idx.p; fa.perm = h->perm_d;
            fa.peer_x0 = p->peer_x0.p + (size_t)x_buf * P; fa.peer_base = p->peer_base.p; fa.peer_ld = p->peer_ld.p;
            fa.bvs = p->bvs; fa.layout = p->layout; fa.ld = p->vec_length;
            fa.peer_arrived = p->peer_arrived.p; fa.peer_acked = p->peer_acked.p;
            fa.acked = p->acked; fa.arrived = p->arrived; fa.epoch = p->epoch; fa.error = p->error; fa.counters = p->fused_counters.p;
            fa.y_rows = (int)scs->n_rows_padded;
            launch_spmmv_fused_any(scs, X, Y_d, p->bvs, p->vec_length, p->layout, fa, main);
            return;
        }
        const bool overlap = p->mode != 0 && uspmv_spmmv_part_supported(scs, p->bvs);
        USPMV_CUDA(cudaEventRecord(p->ev_main, main));
        // Launch order.  The SpMMV kernels hold ~80 registers x 24 warps per SM: launched first, the persistent interior kernel leaves no
        // room for the push CTAs, which then run only when it retires — no overlap at all, and the neighbours get their halo one
        // kernel late.  So the (short) push goes FIRST and the interior CTAs fill the SMs behind it ("mmv_push_first", default on).
        const bool push_first = options().mmv_push_first;
        if (overlap && !push_first && uspmv_spmmv_part(scs, 1, X, Y_d, p->bvs, p->vec_length, p->layout, stream)) throw Error(uspmv_last_error());
        USPMV_CUDA(cudaStreamWaitEvent(comm, p->ev_main, 0));
        launch_push(p, X, x_buf, comm);
        if (overlap && push_first && uspmv_spmmv_part(scs, 1, X, Y_d, p->bvs, p->vec_length, p->layout, stream)) throw Error(uspmv_last_error());
        k_p2p_wait<<<1, 256, 0, comm>>>(P, p->is_sender_d.p, p->arrived, p->epoch, p->error);
        USPMV_LAUNCH_CHECK();
        USPMV_CUDA(cudaEventRecord(p->ev_comm, comm));
        USPMV_CUDA(cudaStreamWaitEvent(main, p->ev_comm, 0));
        if (uspmv_spmmv_part(scs, overlap ? 2 : 0, X, Y_d, p->bvs, p->vec_length, p->layout, stream)) throw Error(uspmv_last_error());
        k_p2p_ack<<<1, 256, 0, main>>>(P, p->is_sender_d.p, p->peer_acked.p, p->epoch);
        USPMV_LAUNCH_CHECK();
    });
}

/* One distributed adaptive-precision SpMV (BASELINE config 4 with seg_nnz partitioning; the reference has no AP + MPI path,
 * utilities.hpp:1445-1451): push the halo of x (original row order) to the neighbours, wait for the own halo, one fused pass over
 * the dp / sp / hp parts, acknowledge.  The plan must come from uspmv_halo_plan_create_multi over the same parts. */
int uspmv_p2p_ap_spmv(uspmv_p2p *p, int ap_mode, const uspmv_scs *dp, const uspmv_scs *sp, const uspmv_scs *hp, void *y_d, void *stream,
                      void *comm_stream) {
    return guarded([&] {
        OptScope opt_scope(ctx_of(p));  // this context's options govern everything below
        if (!p || !y_d) fail("uspmv_p2p_ap_spmv: NULL argument");
        if (!p->connected) fail("uspmv_p2p_ap_spmv: call uspmv_p2p_connect first");
        if (p->bvs != 1) fail("uspmv_p2p_ap_spmv: the arena was created for block vectors");
        if (p->plan->perm_d) fail("uspmv_p2p_ap_spmv: the halo plan was built for a permuted x (use uspmv_halo_plan_create_multi with x_permuted = 0)");
        if (p->vt != (ap_mode == USPMV_AP_SP_HP ? USPMV_F32 : USPMV_F64)) fail("uspmv_p2p_ap_spmv: arena value type does not match the AP mode's x");
        uspmv_halo *h = p->plan;
        cudaStream_t main = as_stream(stream), comm = as_stream(comm_stream);
        const int P = h->P;
        const void *x = p->buffer(0);
        USPMV_CUDA(cudaEventRecord(p->ev_main, main));
        USPMV_CUDA(cudaStreamWaitEvent(comm, p->ev_main, 0));
        launch_push(p, x, 0, comm);
        k_p2p_wait<<<1, 256, 0, comm>>>(P, p->is_sender_d.p, p->arrived, p->epoch, p->error);
        USPMV_LAUNCH_CHECK();
        USPMV_CUDA(cudaEventRecord(p->ev_comm, comm));
        USPMV_CUDA(cudaStreamWaitEvent(main, p->ev_comm, 0));
        if (uspmv_ap_spmv(ap_mode, dp, sp, hp, x, y_d, stream)) throw Error(uspmv_last_error());
        k_p2p_ack<<<1, 256, 0, main>>>(P, p->is_sender_d.p, p->peer_acked.p, p->epoch);
        USPMV_LAUNCH_CHECK();
    });
}

/* The halo exchange alone (init_halo_exchange + finalize_halo_exchange, classes_structs.hpp:857-995): push the halo rows of
 * buffer x_buf into the neighbours' vectors, wait for the own halo, acknowledge.  After it x_buf holds local rows + halo. */
int uspmv_p2p_exchange(uspmv_p2p *p, int x_buf, void *stream, void *comm_stream) {
    return guarded([&] {
        OptScope opt_scope(ctx_of(p));  // this context's options govern everything below
        if (!p) fail("uspmv_p2p_exchange: NULL argument");
        if (!p->connected) fail("uspmv_p2p_exchange: call uspmv_p2p_connect first");
        if (x_buf < 0 || x_buf >= p->n_buf) fail("uspmv_p2p_exchange: x buffer %d out of range", x_buf);
        uspmv_halo *h = p->plan;
        cudaStream_t main = as_stream(stream), comm = as_stream(comm_stream);
        USPMV_CUDA(cudaEventRecord(p->ev_main, main));
        USPMV_CUDA(cudaStreamWaitEvent(comm, p->ev_main, 0));
        launch_push(p, p->buffer(x_buf), x_buf, comm);
        k_p2p_wait<<<1, 256, 0, comm>>>(h->P, p->is_sender_d.p, p->arrived, p->epoch, p->error);
        USPMV_LAUNCH_CHECK();
        USPMV_CUDA(cudaEventRecord(p->ev_comm, comm));
        USPMV_CUDA(cudaStreamWaitEvent(main, p->ev_comm, 0));
        k_p2p_ack<<<1, 256, 0, main>>>(h->P, p->is_sender_d.p, p->peer_acked.p, p->epoch);
        USPMV_LAUNCH_CHECK();
    });
}

/* Host-buffer distributed SpMV, pipelined (the e2e path of bench.py at N > 1; single GPU: uspmv_spmv_host_submit): the arena needs
 * two buffers; call k (slot = k & 1 on EVERY rank — the buffers alternate like in the solve loop) copies this rank's x rows from
 * pinned host memory into buffer `slot`, runs the distributed SpMV (exchange included) and copies y (n_rows_padded entries, permuted
 * order) back, on three streams, so the PCIe transfers of call k overlap the kernel and the opposite-direction copy of call k +- 1. */
int uspmv_p2p_spmv_host_submit(uspmv_p2p *p, const uspmv_scs *scs, const void *x_h, void *y_h, int slot) {
    return guarded([&] {
        OptScope opt_scope(ctx_of(p));  // this context's options govern everything below
        if (!p || !scs || !x_h || !y_h) fail("uspmv_p2p_spmv_host_submit: NULL argument");
        if (!p->connected) fail("uspmv_p2p_spmv_host_submit: call uspmv_p2p_connect first");
        if (p->n_buf != 2 || p->bvs != 1) fail("uspmv_p2p_spmv_host_submit: needs an arena with two single-vector buffers (uspmv_p2p_create_ex, n_buf = 2)");
        if (slot != (int)(p->host_submits & 1)) fail("uspmv_p2p_spmv_host_submit: slot must alternate 0, 1, 0, ... on every rank (expected %d)", (int)(p->host_submits & 1));
        if (p->slot_busy[slot]) fail("uspmv_p2p_spmv_host_submit: slot %d is still in flight (call uspmv_p2p_spmv_host_wait)", slot);
        uspmv_halo *h = p->plan;
        USPMV_CUDA(cudaSetDevice(h->ctx->device));
        const size_t es = vt_size(p->vt);
        if (!p->s_run) {
            USPMV_CUDA(cudaStreamCreateWithFlags(&p->s_h2d, cudaStreamNonBlocking));
            USPMV_CUDA(cudaStreamCreateWithFlags(&p->s_run, cudaStreamNonBlocking));
            USPMV_CUDA(cudaStreamCreateWithFlags(&p->s_d2h, cudaStreamNonBlocking));
            USPMV_CUDA(cudaStreamCreateWithFlags(&p->s_comm, cudaStreamNonBlocking));
            USPMV_CUDA(cudaMallocHost(&p->err_h, 2 * sizeof(unsigned int)));
            p->err_h[0] = p->err_h[1] = 0;
            for (int k = 0; k < 2; ++k) {
                USPMV_CUDA(cudaEventCreateWithFlags(&p->ev_x[k], cudaEventDisableTiming));
                USPMV_CUDA(cudaEventCreateWithFlags(&p->ev_y[k], cudaEventDisableTiming));
                USPMV_CUDA(cudaEventCreateWithFlags(&p->ev_done[k], cudaEventDisableTiming));
            }
        }
        const size_t y_bytes = (size_t)scs->n_rows_padded * es;
        if (p->slot_y[slot].n < y_bytes) p->slot_y[slot].alloc(y_bytes);
        // buffer `slot` was last read by the SpMV of two calls ago (ev_y[slot]); its D2H (ev_done[slot]) was waited for by the caller
        if (p->host_submits >= 2) USPMV_CUDA(cudaStreamWaitEvent(p->s_h2d, p->ev_y[slot], 0));
        USPMV_CUDA(cudaMemcpyAsync(p->buffer(slot), x_h, (size_t)h->n_local * es, cudaMemcpyHostToDevice, p->s_h2d));
        USPMV_CUDA(cudaEventRecord(p->ev_x[slot], p->s_h2d));
        USPMV_CUDA(cudaStreamWaitEvent(p->s_run, p->ev_x[slot], 0));
        if (uspmv_p2p_spmv_buf(p, scs, slot, -1, p->slot_y[slot].p, p->s_run, p->s_comm)) throw Error(uspmv_last_error());
        USPMV_CUDA(cudaEventRecord(p->ev_y[slot], p->s_run));
        USPMV_CUDA(cudaStreamWaitEvent(p->s_d2h, p->ev_y[slot], 0));
        USPMV_CUDA(cudaMemcpyAsync(y_h, p->slot_y[slot].p, y_bytes, cudaMemcpyDeviceToHost, p->s_d2h));
        USPMV_CUDA(cudaMemcpyAsync(p->err_h + slot, p->error, sizeof(unsigned int), cudaMemcpyDeviceToHost, p->s_d2h));  // checked by _wait
        USPMV_CUDA(cudaEventRecord(p->ev_done[slot], p->s_d2h));
        p->slot_busy[slot] = true;
        ++p->host_submits;
    });
}

int uspmv_p2p_spmv_host_wait(uspmv_p2p *p, int slot) {
    return guarded([&] {
        OptScope opt_scope(ctx_of(p));  // this context's options govern everything below
        if (!p) fail("uspmv_p2p_spmv_host_wait: NULL argument");
        if (slot < 0 || slot > 1) fail("uspmv_p2p_spmv_host_wait: slot must be 0 or 1");
        if (!p->slot_busy[slot]) return;
        USPMV_CUDA(cudaEventSynchronize(p->ev_done[slot]));
        p->slot_busy[slot] = false;
        if (p->err_h && p->err_h[slot])
            fail("uspmv_p2p_spmv_host_wait: halo exchange error word = %u (a bounded flag wait timed out; y of slot %d is not valid)", p->err_h[slot], slot);
    });
}

int uspmv_p2p_set_overlap(uspmv_p2p *p, int overlap) {
    return guarded([&] {
        OptScope opt_scope(ctx_of(p));  // this context's options govern everything below
        if (!p) fail("uspmv_p2p_set_overlap: NULL argument");
        if (overlap < 0 || overlap > 2) fail("uspmv_p2p_set_overlap: mode must be 0, 1 or 2");
        p->mode = overlap;
    });
}

/* Waits for everything queued on the device and FAILS if a bounded spin of any step timed out (a peer never signalled): the
 * kernels only raise the arena's error word and carry on, so this is where a lost push becomes a return code. */
int uspmv_p2p_sync(uspmv_p2p *p) {
    return guarded([&] {
        OptScope opt_scope(ctx_of(p));  // this context's options govern everything below
        if (!p) fail("uspmv_p2p_sync: NULL argument");
        USPMV_CUDA(cudaSetDevice(p->plan->ctx->device));
        USPMV_CUDA(cudaDeviceSynchronize());
        unsigned int v[2] = {0, 0};
        USPMV_CUDA(cudaMemcpy(v, p->epoch, 2 * sizeof(unsigned int), cudaMemcpyDeviceToHost));
        if (v[1]) fail("uspmv_p2p_sync: halo exchange error word = %u at epoch %u (a bounded flag wait timed out: a neighbour never pushed / acknowledged)", v[1], v[0]);
    });
}

/* Teardown is COLLECTIVE (freeing CUDA-IPC-exported memory while an importer still has it mapped, or while a neighbour's last
 * kernel still writes its acknowledgement into it, is undefined behaviour):
 *     uspmv_p2p_sync  ->  barrier over all ranks  ->  uspmv_p2p_disconnect  ->  barrier  ->  uspmv_p2p_destroy
 * disconnect waits for this rank's device and closes the neighbours' imported arenas; after it the handle only frees. */
int uspmv_p2p_disconnect(uspmv_p2p *p) {
    return guarded([&] {
        OptScope opt_scope(ctx_of(p));  // this context's options govern everything below
        if (!p) fail("uspmv_p2p_disconnect: NULL argument");
        USPMV_CUDA(cudaSetDevice(p->plan->ctx->device));
        USPMV_CUDA(cudaDeviceSynchronize());
        for (unsigned char *&q : p->peer_arena)
            if (q) { USPMV_CUDA(cudaIpcCloseMemHandle(q)); q = nullptr; }
        p->connected = false;
    });
}

/* 0 = fine; 1 = a bounded spin timed out (a peer never signalled) */
int uspmv_p2p_status(uspmv_p2p *p, int *error_flag, long *epoch) {
    return guarded([&] {
        OptScope opt_scope(ctx_of(p));  // this context's options govern everything below
        if (!p) fail("uspmv_p2p_status: NULL argument");
        unsigned int v[2] = {0, 0};
        USPMV_CUDA(cudaMemcpy(v, p->epoch, 2 * sizeof(unsigned int), cudaMemcpyDeviceToHost));
        if (epoch) *epoch = v[0];
        if (error_flag) *error_flag = (int)v[1];
    });
}

/* Collective, see uspmv_p2p_disconnect.  (Still closes the imported handles itself when disconnect was skipped, e.g. on an error
 * path — the safe order is the caller's job.) */
void uspmv_p2p_destroy(uspmv_p2p *p) {
    if (!p) return;
    for (unsigned char *q : p->peer_arena)
        if (q) cudaIpcCloseMemHandle(q);
    for (cudaStream_t st : {p->s_h2d, p->s_run, p->s_d2h, p->s_comm})
        if (st) cudaStreamDestroy(st);
    for (int k = 0; k < 2; ++k)
        for (cudaEvent_t ev : {p->ev_x[k], p->ev_y[k], p->ev_done[k]})
            if (ev) cudaEventDestroy(ev);
    if (p->err_h) cudaFreeHost(p->err_h);
    if (p->ev_main) cudaEventDestroy(p->ev_main);
    if (p->ev_comm) cudaEventDestroy(p->ev_comm);
    if (p->arena) cudaFree(p->arena);
    delete p;
}

}  // extern "C"
