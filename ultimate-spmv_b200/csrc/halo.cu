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

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <vector>

using namespace uspmv;

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

__global__ void k_mark_remote(const int *__restrict__ col_idxs, long n_elements, int lo, int hi, int n_glob, int *__restrict__ first) {
    long e = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (e >= n_elements) return;
    const int col = col_idxs[e];
    if (col >= lo && col < hi) return;
    if (col < 0 || col >= n_glob) return;  // no owner: the reference leaves such a column untouched
    atomicMin(&first[col], (int)e);
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

__global__ void k_rewrite_cols(int *__restrict__ col_idxs, long n_elements, int lo, int hi, int n_glob, const int *__restrict__ remap) {
    long e = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (e >= n_elements) return;
    const int col = col_idxs[e];
    if (col >= lo && col < hi) col_idxs[e] = col - lo;
    else if (col >= 0 && col < n_glob) col_idxs[e] = remap[col];
}

template <typename VT>
__global__ void k_pack(const int *__restrict__ send_idx, const int *__restrict__ perm, long n_send, const VT *__restrict__ x,
                       VT *__restrict__ buf, int bvs, long ld, int layout) {
    long t = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (t >= n_send * bvs) return;
    long i, v;
    if (layout == USPMV_ROWWISE) { i = t / bvs; v = t % bvs; }
    else { v = t / n_send; i = t % n_send; }
    const long src = perm[send_idx[i]];
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

int uspmv_halo_plan_create(uspmv_scs *s, const int *wsa_h, int rank, int P, uspmv_halo **out) {
    return guarded([&] {
        if (!s || !wsa_h || !out) fail("uspmv_halo_plan_create: NULL argument");
        if (P < 1 || rank < 0 || rank >= P) fail("uspmv_halo_plan_create: bad rank/comm_size");
        if (s->cols_permuted) fail("uspmv_halo_plan_create: call before permute_scs_cols (main.cpp:1271-1308 order)");
        USPMV_CUDA(cudaSetDevice(s->ctx->device));
        const int lo = wsa_h[rank], hi = wsa_h[rank + 1], n_glob = wsa_h[P];
        const long n_local = hi - lo;
        if (n_local != s->n_rows) fail("uspmv_halo_plan_create: work_sharing_arr gives %ld local rows but the matrix has %ld", n_local, s->n_rows);
        auto h = new uspmv_halo();
        try {
            h->ctx = s->ctx; h->rank = rank; h->P = P; h->n_local = n_local;
            h->recv_cumsum.assign(P + 1, 0);
            h->send_ptr.assign(P + 1, 0);
            h->perm_d = s->old_to_new.p;
            const long ne = s->n_elements;
            DevBuf<int> wsa_d(P + 1), first(n_glob > 0 ? n_glob : 1), flag(n_glob + 1), pos(n_glob + 1);
            USPMV_CUDA(cudaMemcpy(wsa_d.p, wsa_h, (P + 1) * sizeof(int), cudaMemcpyHostToDevice));
            if (n_glob) {
                k_fill_int<<<blocks_for(n_glob), TPB>>>(first.p, n_glob, INT32_MAX);
                USPMV_LAUNCH_CHECK();
            }
            if (ne) {
                k_mark_remote<<<blocks_for(ne), TPB>>>(s->col_idxs.p, ne, lo, hi, n_glob, first.p);
                USPMV_LAUNCH_CHECK();
            }
            long n_halo = 0;
            if (n_glob) {
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
            if (ne) {
                k_rewrite_cols<<<blocks_for(ne), TPB>>>(s->col_idxs.p, ne, lo, hi, n_glob, first.p);
                USPMV_LAUNCH_CHECK();
            }
            USPMV_CUDA(cudaDeviceSynchronize());
        } catch (...) { delete h; throw; }
        *out = h;
    });
}

int uspmv_halo_plan_counts(const uspmv_halo *h, int *recv_counts_cumsum_h, long *n_halo) {
    return guarded([&] {
        if (!h) fail("uspmv_halo_plan_counts: plan is NULL");
        if (recv_counts_cumsum_h)
            for (int p = 0; p <= h->P; ++p) recv_counts_cumsum_h[p] = h->recv_cumsum[p];
        if (n_halo) *n_halo = h->n_halo;
    });
}

int uspmv_halo_plan_need(const uspmv_halo *h, int *need_flat_h, int *need_ptr_h) {
    return guarded([&] {
        if (!h) fail("uspmv_halo_plan_need: plan is NULL");
        if (need_flat_h)
            for (long k = 0; k < h->n_halo; ++k) need_flat_h[k] = h->need_flat[k];
        if (need_ptr_h)
            for (int p = 0; p <= h->P; ++p) need_ptr_h[p] = h->recv_cumsum[p];
    });
}

int uspmv_halo_plan_set_send(uspmv_halo *h, const int *send_flat_h, const int *send_ptr_h) {
    return guarded([&] {
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
