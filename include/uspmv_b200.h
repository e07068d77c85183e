/* uspmv_b200 — C ABI of the B200-native SELL-C-sigma SpMV/SpMMV engine.
 *
 * Drop-in boundary for the compute path of RRZE-HPC/Ultimate-SpMV.  The reference has no C ABI: its
 * "interface" is a header of C++ templates (code/interface.hpp, API_doc.md) plus two std::function
 * signatures inside the harness (code/classes_structs.hpp:283-333).  Every entry point below names the
 * reference interface it replaces (file:line, paths relative to the reference checkout).  The C++ shim
 * include/uspmv_interface.hpp re-creates the reference's template API (MtxData, ScsData,
 * convert_to_scs, uspmv_scs_gpu, ...) on top of these calls; INTEGRATION.md shows the binding.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; uspmv_last_error() gives the message
 *     (the reference prints and exit()s instead: classes_structs.hpp:33-41, kernels.hpp:295-299);
 *   - ST = long, IT = int as in the reference (mmio.h:21, classes_structs.hpp:31);
 *   - "_h" pointers are host memory, "_d" pointers are device memory of the context's GPU;
 *   - `stream` is a cudaStream_t passed as void* (NULL = the legacy default stream);
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails with an error.
 */
#ifndef USPMV_B200_H
#define USPMV_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct uspmv_ctx uspmv_ctx; /* one GPU (one rank)                                         */
typedef struct uspmv_coo uspmv_coo; /* device-resident MtxData (interface.hpp:16-56)              */
typedef struct uspmv_scs uspmv_scs; /* device-resident ScsData (interface.hpp:58-80)              */
typedef struct uspmv_halo uspmv_halo; /* per-rank halo plan (ContextData, classes_structs.hpp:156) */

/* -dp / -sp / -hp (utilities.hpp:1190-1260) */
enum { USPMV_F64 = 0, USPMV_F32 = 1, USPMV_F16 = 2 };
/* BLOCK_VECTOR_LAYOUT colwise / rowwise (Makefile:17-31; kernels.hpp:352,358) */
enum { USPMV_COLWISE = 0, USPMV_ROWWISE = 1 };
/* -ap[dp_sp] / -ap[dp_hp] / -ap[sp_hp] / -ap[dp_sp_hp] (utilities.hpp:1262-1330) */
enum { USPMV_AP_DP_SP = 0, USPMV_AP_DP_HP = 1, USPMV_AP_SP_HP = 2, USPMV_AP_DP_SP_HP = 3 };
/* -seg_rows / -seg_nnz (mpi_funcs.hpp:446-493) */
enum { USPMV_SEG_ROWS = 0, USPMV_SEG_NNZ = 1 };

const char *uspmv_last_error(void);
int uspmv_version(void);
/* number of this library's kernel launches since load (bench.py's "gpu_launches" claim) */
long uspmv_kernel_launches(void);
/* Kernel-selection knobs (no reference counterpart; THREADS_PER_BLOCK is the closest, config.mk:20):
 *   "scs_stream" 0/1          C = 32: bulk-copy (TMA) streamed kernel (default 1) or the direct-load kernel
 *   "stream_variant" 0..5     (slots per piece, ring depth, warps per CTA) instantiation
 *   "stream_blocks_per_sm"    persistent CTAs per SM
 *   "strict_reference_halo"   0 (default): padding slots stay local (column 0 of the rank); 1: replicate the reference,
 *                             where padding's column 0 is GLOBAL column 0 and so becomes one spurious halo element
 *                             received from rank 0 on every rank > 0 (utilities.hpp:1991-2002, mpi_funcs.hpp:279-283)
 * The complete table (SpMMV / AP variants, fused vs multi-kernel exchange, push kernels) with defaults: INTEGRATION.md section 7.
 * No knob changes a result bit. */
int uspmv_set_option(const char *name, long value);
/* The knobs are per CONTEXT: uspmv_set_option changes the process defaults AND every live context; uspmv_ctx_set_option changes one
 * context only, so two contexts in one process (two GPUs, two solver instances) can run different variants.  get reads them back
 * (ctx == NULL: the process defaults). */
int uspmv_ctx_set_option(uspmv_ctx *ctx, const char *name, long value);
int uspmv_ctx_get_option(const uspmv_ctx *ctx, const char *name, long *out);

/* ---- context and device memory --------------------------------------------------------------- */
/* cudaGetDeviceCount + cudaSetDevice(rank % ndev) in the reference: main.cpp:1838-1842 */
int uspmv_device_count(int *out_ndev);
int uspmv_ctx_create(int device, uspmv_ctx **out);
void uspmv_ctx_destroy(uspmv_ctx *ctx);
int uspmv_ctx_sync(uspmv_ctx *ctx);
/* a second stream for the halo exchange (cudaStream_t as void *; the harness' per-peer cudaStreams, classes_structs.hpp:857-995) */
int uspmv_stream_create(uspmv_ctx *ctx, void **out_stream);
int uspmv_stream_destroy(uspmv_ctx *ctx, void *stream);
/* cudaMalloc/cudaMemcpy staging of assign_spmv_kernel_gpu_data (utilities.hpp:3302-3815, 3720-3811) */
int uspmv_malloc(uspmv_ctx *ctx, size_t bytes, void **out_d);
int uspmv_free(uspmv_ctx *ctx, void *ptr_d);
int uspmv_memcpy_h2d(uspmv_ctx *ctx, void *dst_d, const void *src_h, size_t bytes, void *stream);
int uspmv_memcpy_d2h(uspmv_ctx *ctx, void *dst_h, const void *src_d, size_t bytes, void *stream);
int uspmv_memset(uspmv_ctx *ctx, void *dst_d, int byte, size_t bytes, void *stream);
/* pinned host staging (the harness uses pageable std::vector; pinned is what makes e2e copies fast) */
int uspmv_host_alloc(size_t bytes, void **out_h);
int uspmv_host_free(void *ptr_h);

/* ---- MtxData (COO) ---------------------------------------------------------------------------- */
/* MtxData filled by the caller (interface.hpp:16-56).  Values are `mt`-typed (USPMV_F64/F32/F16).
 * Rows need not be sorted; within-row order of the input is preserved (utilities.hpp:2013-2036). */
int uspmv_coo_from_host(uspmv_ctx *ctx, long n_rows, long n_cols, long nnz, const int *I_h, const int *J_h,
                        const void *values_h, int mt, uspmv_coo **out);
/* same, arrays already on the device (copied) */
int uspmv_coo_from_device(uspmv_ctx *ctx, long n_rows, long n_cols, long nnz, const int *I_d, const int *J_d,
                          const void *values_d, int mt, uspmv_coo **out);
/* Synthetic stencil generated on the device (BASELINE.json configs 2/3/5): rows [row0, row1) of the
 * nx*ny*nz grid, row = (z*ny + y)*nx + x, columns ascending, Dirichlet; points = 7 or 27; diagonal =
 * points-1, off-diagonals -1; dp values.  I is made slab-local (I - row0), J stays global
 * (localize_row_idx, mpi_funcs.hpp:862-877). */
int uspmv_coo_stencil(uspmv_ctx *ctx, int points, long nx, long ny, long nz, long row0, long row1, uspmv_coo **out);
/* Synthetic irregular power-law matrix generated on the device (BASELINE.json config 4): n x n, rows
 * [row0, row1) with local row ids and global columns; row degree clamp(floor(d_min (1-u)^(-1/(alpha-1))), 1, max_deg), half of a
 * row's columns within +-1024 of the diagonal, half uniform, de-duplicated and ascending; values sign * 10^w, w ~ U(-4, 2); all
 * randomness splitmix64(seed, row, k), so every rank can generate its own rows.  (The reference reads such matrices from .mtx files,
 * utilities.hpp:2148-2309; a 5e8-element file is not practical on the benchmark box.) */
int uspmv_coo_powerlaw(uspmv_ctx *ctx, long n, long row0, long row1, double d_min, double alpha, int max_deg, unsigned long seed,
                       uspmv_coo **out);
/* raw device pointers of the COO arrays (MtxData::I / J / values, interface.hpp:16-56); any out pointer may be NULL */
int uspmv_coo_device_arrays(const uspmv_coo *coo, const int **I_d, const int **J_d, const void **values_d);
/* read_mtx's post-processing (utilities.hpp:2214-2290) on the device: entries in file order (0-based), symmetric != 0 expands
 * (i,j) into (i,j),(j,i) for i != j, then a stable sort by row.  Text parsing stays with the caller. */
int uspmv_coo_from_entries(uspmv_ctx *ctx, long n_rows, long n_cols, long nz, const int *I_h, const int *J_h, const double *values_h,
                           int symmetric, uspmv_coo **out);
/* equilibrate_matrix (utilities.hpp:2605-2684), in place on a dp COO; rowmax_h / colmax_h (optional) receive the row maxima and
 * the column maxima of the row-scaled matrix (-equilibrate; main.cpp:1118-1153). */
int uspmv_coo_equilibrate(uspmv_coo *coo, double *rowmax_h, double *colmax_h);
int uspmv_coo_dims(const uspmv_coo *coo, long out3[3]); /* n_rows, n_cols, nnz */
int uspmv_coo_export(const uspmv_coo *coo, int *I_h, int *J_h, void *values_h);
void uspmv_coo_destroy(uspmv_coo *coo);

/* ---- SELL-C-sigma construction ------------------------------------------------------------------ */
/* convert_to_scs (utilities.hpp:1842-2104; interface.hpp:401-656), built on the device, bit-exact:
 * sigma-window std::sort order (libstdc++ 13 introsort tie order), chunk_ptrs/chunk_lengths,
 * padding (value 0, column 0), COO fill order, old_to_new / new_to_old.  fixed_perm_h (may be NULL)
 * is the reference's `fixed_permutation` (n_rows ints).  CRS is C = 1, sigma = 1.  vt = value type of
 * the stored matrix (MT -> VT conversion of utilities.hpp:2033 is a single rounding). */
int uspmv_scs_build(uspmv_ctx *ctx, const uspmv_coo *coo, long C, long sigma, int vt, const int *fixed_perm_h,
                    uspmv_scs **out);
/* Adopt a SELL-C-sigma matrix built elsewhere — e.g. a ScsData filled by the reference's own host convert_to_scs
 * (interface.hpp:401-656) — into a device-resident handle: the arrays (host, or device when on_device != 0) are copied.
 * chunk_ptrs has n_chunks + 1 entries; n_elements = chunk_ptrs[n_chunks]; old_to_new may be NULL (identity). */
int uspmv_scs_from_arrays(uspmv_ctx *ctx, int vt, long C, long sigma, long n_rows, long n_cols, long n_chunks, const int *chunk_ptrs,
                          const int *chunk_lengths, const int *col_idxs, const void *values, const int *old_to_new, int on_device,
                          uspmv_scs **out);
/* out8 = C, sigma, n_rows, n_cols, n_rows_padded, n_chunks, n_elements, nnz (ScsData scalars) */
int uspmv_scs_dims(const uspmv_scs *scs, long out8[8]);
/* Copy the arrays to the host (any pointer may be NULL).  Sizes: chunk_ptrs n_chunks+1, chunk_lengths
 * n_chunks, col_idxs / values n_elements, old_to_new n_rows, new_to_old n_rows_padded.  Positions of
 * new_to_old that no real row maps to are uninitialised in the reference (utilities.hpp:2060-2066);
 * here they hold -1. */
int uspmv_scs_export(const uspmv_scs *scs, int *chunk_ptrs_h, int *chunk_lengths_h, int *col_idxs_h, void *values_h,
                     int *old_to_new_h, int *new_to_old_h);
/* permute_scs_cols (utilities.hpp:1802-1831).  perm_h == NULL uses the struct's own old_to_new_idx
 * (main.cpp:1308).  Columns >= n_rows (halo) are left alone; padding slots are permuted too. */
int uspmv_scs_permute_cols(uspmv_scs *scs, const int *perm_h);
/* Raw device pointers, for callers that keep the reference's kernel-argument style
 * (OnePrecKernelArgs, classes_structs.hpp:213-236).  Any out pointer may be NULL. */
int uspmv_scs_device_arrays(const uspmv_scs *scs, const int **chunk_ptrs_d, const int **chunk_lengths_d,
                            const int **col_idxs_d, const void **values_d, const int **old_to_new_d,
                            const int **new_to_old_d);
void uspmv_scs_destroy(uspmv_scs *scs);

/* ---- vector permutations ------------------------------------------------------------------------ */
/* apply_permutation (utilities.hpp:1768-1782): out[i] = in[perm[i]], i < n; perm[i] < 0 gives 0. */
int uspmv_apply_permutation(uspmv_ctx *ctx, void *out_d, const void *in_d, const int *perm_d, long n, int vt, void *stream);
/* apply_strided_permutation (utilities.hpp:1784-1799) generalised to whole block rows:
 * rowwise: out[i*bvs + v] = in[perm[i]*bvs + v]; colwise: out[i + v*ld] = in[perm[i] + v*ld]. */
int uspmv_apply_permutation_block(uspmv_ctx *ctx, void *out_d, const void *in_d, const int *perm_d, long n, int vt,
                                  int bvs, long ld, int layout, void *stream);

/* apply_strided_permutation (utilities.hpp:1784-1799), literally: out[i * stride] = in[perm[i] * stride], i < n. */
int uspmv_apply_strided_permutation(uspmv_ctx *ctx, void *out_d, const void *in_d, const int *perm_d, long n, long stride, int vt,
                                    void *stream);
/* generate_inv_perm (utilities.hpp:1755-1766): inv_perm[perm[i]] = i, i < perm_len; inv_perm has inv_len entries (an entry of perm
 * outside [0, inv_len) is an error here, an out-of-bounds write in the reference). */
int uspmv_generate_inv_perm(uspmv_ctx *ctx, const int *perm_d, int *inv_perm_d, long perm_len, long inv_len, void *stream);
/* random_init (utilities.hpp:880-912) + the padding rule of init_std_vec_with_ptr_or_value (:914-981) — HOST logic, no GPU needed:
 * a DEFAULT-seeded std::mt19937 feeding uniform_real_distribution<double>(vmin, vmax), drawn sequentially over all n values of the
 * (block) vector, converted to `vt`; then everything past the real rows is zeroed (colwise: positions >= n_rows of every
 * vec_length-long vector; rowwise: everything from n_rows * bvs on).  n_rows < 0 skips the padding rule. */
int uspmv_random_init_host(double vmin, double vmax, long n, int vt, void *out_h, long n_rows, long vec_length, int bvs, int layout);
/* 1 if ptr is device (or managed) memory, 0 for host memory / NULL — lets the C++ shim accept either kind of array */
int uspmv_pointer_is_device(const void *ptr, int *out);

/* ---- kernels ------------------------------------------------------------------------------------- */
/* uspmv_scs_gpu / uspmv_csr_gpu (interface.hpp:1741-1867) == spmv_gpu_scs[_adv] / spmv_gpu_csr
 * (kernels.hpp:579-775): raw device arrays, y has n_chunks*C entries in permuted order. */
int uspmv_scs_gpu(uspmv_ctx *ctx, int vt, long C, long n_chunks, const int *chunk_ptrs_d, const int *chunk_lengths_d,
                  const int *col_idxs_d, const void *values_d, const void *x_d, void *y_d, void *stream);
int uspmv_csr_gpu(uspmv_ctx *ctx, int vt, long n_rows, const int *row_ptrs_d, const int *col_idxs_d, const void *values_d,
                  const void *x_d, void *y_d, void *stream);
/* SpmvKernel::execute on a built matrix (classes_structs.hpp:997-1035,1117): CRS kernel iff C == 1 and
 * sigma == 1 (execute_uspmv's rule, interface.hpp:1911), else the SCS kernel. */
int uspmv_spmv(const uspmv_scs *scs, const void *x_d, void *y_d, void *stream);
/* Fused form (SURVEY §7 hard part 4): gathers x in ORIGINAL numbering through columns that have NOT
 * been permuted and writes y[new_to_old[row]] — no separate apply_permutation passes. */
int uspmv_spmv_unpermuted(const uspmv_scs *scs, const void *x_d, void *y_d, void *stream);
/* block_spmv_{csr,scs} (kernels.hpp:68-154,306-398; GPU launchers are stubs in the reference,
 * kernels.hpp:777-844).  bvs = block_vec_size, vec_length = n_local + per_vector_padding
 * (classes_structs.hpp:1024), layout = USPMV_COLWISE / USPMV_ROWWISE. */
int uspmv_spmmv(const uspmv_scs *scs, const void *X_d, void *Y_d, int bvs, long vec_length, int layout, void *stream);
/* The same on caller-owned DEVICE arrays, like every kernel of the reference (raw arrays per call, kernels.hpp:68-154,306-398).
 * C == 1 is CRS (chunk_ptrs = row_ptrs; chunk_lengths may be NULL).  Y has n_chunks * C block rows. */
int uspmv_block_spmv_gpu(uspmv_ctx *ctx, int vt, long C, long n_chunks, const int *chunk_ptrs_d, const int *chunk_lengths_d,
                         const int *col_idxs_d, const void *values_d, const void *X_d, void *Y_d, int bvs, long vec_length, int layout,
                         void *stream);
/* Host-buffer call (the reference-facing path measured as "e2e"): copies x to the device, runs the
 * kernel, copies y (n_rows_padded entries) back, synchronises. */
int uspmv_spmv_host(const uspmv_scs *scs, const void *x_h, long x_len, void *y_h, long y_len);
/* Pipelined form of the same call for back-to-back SpMVs with different host vectors (up to 3 slots in flight): submit
 * enqueues H2D(x) -> kernel -> D2H(y) on three streams and returns; wait blocks until y_h of that slot is complete.
 * The transfers of one SpMV overlap the kernel and the opposite-direction transfer of its neighbours (PCIe is full duplex).
 * x_h / y_h must stay valid until wait; pin them with uspmv_host_alloc. */
int uspmv_spmv_host_submit(const uspmv_scs *scs, const void *x_h, long x_len, void *y_h, long y_len, int slot);
int uspmv_spmv_host_wait(const uspmv_scs *scs, int slot);

/* ---- adaptive precision --------------------------------------------------------------------------- */
/* partition_precisions (interface.hpp:690-978; utilities.hpp:2810-3123): order-preserving split of a
 * COO matrix by abs(value) against t1 (and t2 for the 3-way split).  rowmax_h/colmax_h (NULL = not
 * equilibrated) divide the thresholds (interface.hpp:908-936).  Outputs that the mode does not use are
 * set to NULL.  dp part holds doubles, sp floats, hp halves. */
int uspmv_partition_precisions(uspmv_ctx *ctx, const uspmv_coo *coo, int ap_mode, double t1, double t2,
                               const double *rowmax_h, const double *colmax_h, uspmv_coo **dp, uspmv_coo **sp,
                               uspmv_coo **hp);
/* uspmv_scs_ap{dpsp,dphp,sphp,dpsphp} / uspmv_csr_ap* (interface.hpp:1129-1733; ap_kernels.hpp:24-953)
 * as ONE fused pass.  Parts not used by the mode are NULL.  Modes 0,1,3: dp_x (double) in, y double.
 * Mode 2 (sp_hp): sp_x (float) in, y float (interface.hpp:1620-1645). */
int uspmv_ap_spmv(int ap_mode, const uspmv_scs *dp, const uspmv_scs *sp, const uspmv_scs *hp, const void *x_d, void *y_d,
                  void *stream);

/* The same on caller-owned DEVICE arrays — uspmv_scs_ap{dpsp,dphp,sphp,dpsphp}_gpu / uspmv_csr_ap*_gpu in the reference's argument
 * style (raw arrays of every precision part per call: interface.hpp:1129-1733, ap_kernels.hpp:637-953, MultiPrecKernelArgs
 * classes_structs.hpp:238-261).  arrays[4 p + 0..3] = chunk_ptrs, chunk_lengths, col_idxs, values of part p (0 dp, 1 sp, 2 hp); parts
 * the mode does not use are ignored.  The parts share C and n_chunks.  C == 1 is CRS (a NULL chunk_lengths is derived). */
int uspmv_scs_ap_gpu(uspmv_ctx *ctx, int ap_mode, long C, long n_chunks, const void *const *arrays12, const void *x_d, void *y_d,
                     void *stream);

/* ---- column-banded execution plan (EXPERIMENTAL) ------------------------------------------------------ */
/* For matrices whose x does not fit the L2 (BASELINE config 4 on one GPU: every missing 8-byte gather costs a 128-byte DRAM fill):
 * the columns are cut into n_bands ranges, every band becomes its own SELL-C-sigma structure (AP: one per precision part) built with
 * the reference's fixed_permutation mechanism (utilities.hpp:1911-1928; main.cpp:1175-1219) on the sigma-sorted row order of the
 * whole matrix, and y = sum over bands, band by band, so that one band of x stays cache resident.  x in the original column
 * numbering, y (n_rows_padded values) in the permuted row order (old_to_new from uspmv_banded_perm).  ap_mode < 0: one precision
 * `vt`; otherwise USPMV_AP_* with thresholds t1 / t2.  n_bands == 0: about 32 MB of x per band.  Within 1e-12 / 1e-5 / 1e-2 of
 * the un-banded result (band sums are added in band order); the un-banded kernels stay the bit-identical default. */
typedef struct uspmv_banded uspmv_banded;
int uspmv_banded_build(uspmv_ctx *ctx, const uspmv_coo *coo, long C, long sigma, int vt, int ap_mode, double t1, double t2, int n_bands,
                       uspmv_banded **out);
/* out8 = n_bands, band_width, n_rows, n_cols, n_rows_padded, nnz, n_elements over all bands and parts, value type of x / y */
int uspmv_banded_dims(const uspmv_banded *plan, long out8[8]);
int uspmv_banded_perm(const uspmv_banded *plan, int *old_to_new_h);
int uspmv_banded_spmv(const uspmv_banded *plan, const void *x_d, void *y_d, void *stream);
void uspmv_banded_destroy(uspmv_banded *plan);

/* ---- row partitioning and halo exchange (one rank per GPU) ------------------------------------------ */
/* seg_work_sharing_arr (mpi_funcs.hpp:424-622): wsa_h has P+1 entries.  I_h is the row array of the
 * row-sorted global COO. */
int uspmv_seg_work_sharing_arr(int seg_method, long n_rows, long nnz, const int *I_h, int P, int *wsa_h);
/* seg_mtx_struct + localize_row_idx (mpi_funcs.hpp:636-674,862-877) on the device: rows [wsa[rank], wsa[rank+1]) of a ROW-SORTED
 * COO as a new COO with process-local row ids and GLOBAL columns, input order kept.  n_distinct_rows (optional) = number of distinct
 * rows present — what the reference stores as the local n_rows (mpi_funcs.hpp:770); the returned COO has n_rows = wsa[rank+1] -
 * wsa[rank] (different only when the slab holds empty rows, where the reference's n_rows no longer covers its own row ids). */
int uspmv_coo_seg_mtx(const uspmv_coo *total, const int *wsa_h, int rank, int P, uspmv_coo **out, long *n_distinct_rows);
/* collect_local_needed_heri (mpi_funcs.hpp:242-415) on the device: rewrites the matrix' global
 * columns to local/halo numbering (first-seen order, grouped by owner) and records the need lists. */
int uspmv_halo_plan_create(uspmv_scs *scs, const int *wsa_h, int rank, int P, uspmv_halo **out);
/* recv_counts_cumsum (P+1 ints) and the per-owner need lists, flattened; need_ptr has P+1 entries. */
/* the same over the dp / sp / hp parts of an adaptive-precision matrix (one shared numbering of the remote columns);
 * x_permuted = 0: x stays in the original row order, as the AP kernels expect (main.cpp:1308-1332) */
int uspmv_halo_plan_create_multi(uspmv_scs **parts, int n_parts, const int *wsa_h, int rank, int P, int x_permuted, uspmv_halo **out);
int uspmv_halo_plan_counts(const uspmv_halo *plan, int *recv_counts_cumsum_h, long *n_halo);
int uspmv_halo_plan_need(const uspmv_halo *plan, int *need_flat_h, int *need_ptr_h);
/* collect_comm_idxs (mpi_funcs.hpp:117-172): install what every peer needs from this rank
 * (send_ptr P+1 entries, owner-local row ids). */
int uspmv_halo_plan_set_send(uspmv_halo *plan, const int *send_flat_h, const int *send_ptr_h);
/* Overlap support (the reference has none, main.cpp:464-468: begin -> finish -> execute): classify the chunks of a
 * halo-renumbered matrix into interior (no column >= n_rows) and boundary ones WITHOUT reordering them, and run the
 * SpMV on one class: which = 0 all, 1 interior, 2 boundary.  Rows of the other class are not written. */
int uspmv_scs_split_chunks(uspmv_scs *scs, long *n_interior, long *n_boundary);
int uspmv_spmv_part(const uspmv_scs *scs, int which, const void *x_d, void *y_d, void *stream);
/* pack_send_buf / pack_d_send_buf (classes_structs.hpp:786-831; kernels.hpp:554-577) for ALL peers in one
 * launch: buf[send_ptr[p] + i] = x[perm[send_idx[p][i]]] (block vectors: bvs values per index). */
/* SpMMV over the interior (1) / boundary (2) chunks only (0 = all); subsets need the streamed kernel — ask first. */
int uspmv_spmmv_part_supported(const uspmv_scs *scs, int block_vec_size);
int uspmv_spmmv_part(const uspmv_scs *scs, int which, const void *X_d, void *Y_d, int block_vec_size, long vec_length, int layout,
                     void *stream);
int uspmv_halo_pack(const uspmv_halo *plan, const void *x_d, void *sendbuf_d, int vt, int bvs, long vec_length, int layout,
                    void *stream);
void uspmv_halo_destroy(uspmv_halo *plan);

/* NVLink peer-to-peer halo exchange for one-process-per-GPU runs (replaces MPI_Isend/Irecv/Waitall in
 * init/finalize_halo_exchange, classes_structs.hpp:857-995, which assume CUDA-aware MPI).  Every rank allocates an
 * "arena" (its x vector + epoch flags), shares it through a 64-byte CUDA IPC handle, and the pack kernel of the
 * neighbours stores their elements straight into the tail of x.  The caller all-gathers the handles (any transport)
 * and passes, per peer q, the size of q's x region and the element offset where this rank's data starts in q's x
 * (q.n_local + q.recv_counts_cumsum[rank]). */
typedef struct uspmv_p2p uspmv_p2p;
int uspmv_p2p_create(uspmv_halo *plan, int vt, long x_len, uspmv_p2p **out, void *ipc_handle64, void **x_d);
int uspmv_p2p_connect(uspmv_p2p *p2p, const void *all_handles, const long *peer_x_bytes, const long *peer_halo_base);
/* one SpMV incl. halo exchange, overlapped: push + wait on comm_stream, interior chunks on stream, then boundary */
int uspmv_p2p_spmv(uspmv_p2p *p2p, const uspmv_scs *scs, void *y_d, void *stream, void *comm_stream);
/* mode 2 (default, C = 32): ONE fused kernel per SpMV — push to the peers, interior chunks, wait for the own halo, boundary
 * chunks, acknowledge; mode 1: separate push/wait kernels on comm_stream next to the interior kernel; mode 0: exchange,
 * then one full SpMV (the reference's begin -> finish -> execute order, main.cpp:464-468) */
/* distributed adaptive-precision SpMV: halo push + wait, one fused pass over the parts, acknowledge (x = the arena's buffer 0,
 * original row order; plan from uspmv_halo_plan_create_multi).  New behaviour: the reference refuses AP with MPI
 * (utilities.hpp:1445-1451); BASELINE config 4 defines it as seg_nnz partitioning + per-rank partition_precisions. */
int uspmv_p2p_ap_spmv(uspmv_p2p *p2p, int ap_mode, const uspmv_scs *dp, const uspmv_scs *sp, const uspmv_scs *hp, void *y_d,
                      void *stream, void *comm_stream);
int uspmv_p2p_set_overlap(uspmv_p2p *p2p, int overlap);
/* Generalised arena: n_buf (1 or 2) buffers, each holding block_vec_size vectors of vec_length elements in `layout`.
 * x_d receives the n_buf buffer addresses.  peer_vec_length[q] (connect) is q's vec_length, needed for column-major block
 * vectors.  Replaces the per-vector / strided-datatype messages of the reference's multivec / bulkvec modes
 * (classes_structs.hpp:909-970, mpi_funcs.hpp:1003-1059) with ONE push per neighbour. */
int uspmv_p2p_create_ex(uspmv_halo *plan, int vt, long vec_length, int block_vec_size, int layout, int n_buf, uspmv_p2p **out,
                        void *ipc_handle64, void **x_d);
int uspmv_p2p_connect_ex(uspmv_p2p *p2p, const void *all_handles, const long *peer_x_bytes, const long *peer_halo_base,
                         const long *peer_vec_length);
/* SpMV that reads x from buffer x_buf and writes y to y_d, or (y_d NULL) into buffer y_buf — rows < n_rows only, the tail of
 * that buffer receives the next step's halo.  With two buffers `rev` x { SpMV ; swap } (solve mode, main.cpp:528-631;
 * SpmvKernel::swap_local_vectors, classes_structs.hpp:1130-1165) stays on the device with no copy. */
int uspmv_p2p_spmv_buf(uspmv_p2p *p2p, const uspmv_scs *scs, int x_buf, int y_buf, void *y_d, void *stream, void *comm_stream);
/* SpMMV over the arena's block vector (buffer x_buf) incl. the exchange of all block_vec_size vectors of the halo rows,
 * overlapped with the interior chunks when the streamed kernel applies. */
int uspmv_p2p_spmmv(uspmv_p2p *p2p, const uspmv_scs *scs, int x_buf, void *Y_d, void *stream, void *comm_stream);
/* The halo exchange alone — init_halo_exchange + finalize_halo_exchange (classes_structs.hpp:857-995): push the halo rows of
 * buffer x_buf into the neighbours' vectors over NVLink, wait for the own halo, acknowledge. */
int uspmv_p2p_exchange(uspmv_p2p *p2p, int x_buf, void *stream, void *comm_stream);
/* Host-buffer distributed SpMV, pipelined over the two buffers of the arena (n_buf = 2): call k uses slot k & 1 on every rank;
 * x_h holds this rank's n_local x values (permuted order), y_h receives n_rows_padded values; both should be pinned.  H2D, the
 * exchange + SpMV and D2H of neighbouring calls overlap.  (The harness copies x / y once around the whole loop,
 * main.cpp:727-767; this is the per-call host-buffer form an application with host-resident vectors would use.) */
int uspmv_p2p_spmv_host_submit(uspmv_p2p *p2p, const uspmv_scs *scs, const void *x_h, void *y_h, int slot);
int uspmv_p2p_spmv_host_wait(uspmv_p2p *p2p, int slot);
int uspmv_p2p_status(uspmv_p2p *p2p, int *error_flag, long *epoch);
/* cudaDeviceSynchronize + FAIL when a bounded flag wait of any step timed out (the kernels only raise the arena's error word);
 * uspmv_p2p_spmv_host_wait reports the same condition for its slot.  The reference would sit in MPI_Waitall instead
 * (classes_structs.hpp:926-995). */
int uspmv_p2p_sync(uspmv_p2p *p2p);
/* Teardown is COLLECTIVE (the arena is CUDA-IPC-exported; neighbours write acknowledgements into it after this rank's last step):
 *   uspmv_p2p_sync -> barrier over all ranks -> uspmv_p2p_disconnect (closes the imported arenas) -> barrier -> uspmv_p2p_destroy.
 * The harness does this after its bench loop (MPI_Barrier + MPI_Finalize in the reference, main.cpp:1895-1899). */
int uspmv_p2p_disconnect(uspmv_p2p *p2p);
void uspmv_p2p_destroy(uspmv_p2p *p2p);

#ifdef __cplusplus
}
#endif
#endif /* USPMV_B200_H */
