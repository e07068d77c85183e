/* TEST INFRASTRUCTURE ONLY — CPU restatement ("oracle") of the Ultimate-SpMV hot path.
 *
 * Plain C restatement of what the reference computes on the SELL-C-sigma path, written from the
 * reference's behaviour (file:line cited per function; all paths relative to /root/reference/code).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * this library; nothing under ultimate-spmv_b200/ or include/ links, loads or calls it.
 *
 * PARITY PIN: every function here is checked (tests/test_oracle_pinning.py, all CPU)
 *   - against the reference itself, compiled unmodified from /root/reference into oracle/_ref/ by
 *     oracle/Makefile (bit-exact for all index structures; bit-exact y for SCS kernels), and
 *   - against the reference's own unit-test goldens (test_suite/test_data/M_big.cpp, M1.cpp),
 *     transcribed into tests/golden/ref_testsuite_goldens.json, and
 *   - against tests/golden/ npz fixtures generated from oracle/_ref by oracle/make_golden.py.
 *
 * THIRD-PARTY ALGORITHM: the sigma-window row ordering is decided by libstdc++'s std::sort
 * (GCC 13.3.0, bits/stl_algo.h:1848-1951 + bits/stl_heap.h), which the reference calls at
 * utilities.hpp:1936 / interface.hpp:493 and does not pin.  orc_sort_window() restates that
 * published algorithm (introsort: median-of-3 quicksort down to 16-element runs, heapsort when the
 * depth budget 2*floor(log2 n) is exhausted, then one guarded + unguarded insertion sort pass).
 *
 * Arithmetic convention: the reference is built with g++ -O3 on an FMA machine, where GCC contracts
 * `acc += a*b` into a fused multiply-add (-ffp-contract=fast is GCC's default for C++).  The
 * restatement therefore calls fma()/fmaf() explicitly; tests assert bit-equality with oracle/_ref.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef _Float16 f16;
enum { VT_F64 = 0, VT_F32 = 1, VT_F16 = 2 };

typedef struct {
    long idx; /* original row */
    long cnt; /* number of stored elements in that row */
} rowcnt_t;

/* comparator of utilities.hpp:1938-1940: "sort longer rows first" */
static inline int longer(const rowcnt_t *a, const rowcnt_t *b) { return a->cnt > b->cnt; }

static inline void swap_rc(rowcnt_t *a, rowcnt_t *b) {
    rowcnt_t t = *a;
    *a = *b;
    *b = t;
}

/* ---- libstdc++ heap primitives (stl_heap.h: __push_heap, __adjust_heap, __make_heap, __pop_heap) ---- */
static void push_heap_rc(rowcnt_t *first, long hole, long top, rowcnt_t value) {
    long parent = (hole - 1) / 2;
    while (hole > top && longer(&first[parent], &value)) {
        first[hole] = first[parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    first[hole] = value;
}

static void adjust_heap_rc(rowcnt_t *first, long hole, long len, rowcnt_t value) {
    const long top = hole;
    long child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (longer(&first[child], &first[child - 1])) child--;
        first[hole] = first[child];
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        first[hole] = first[child - 1];
        hole = child - 1;
    }
    push_heap_rc(first, hole, top, value);
}

static long g_heapsort_calls = 0;
/* test hook: how often the depth budget ran out (proves the heapsort branch is exercised) */
long orc_heapsort_calls(void) { return g_heapsort_calls; }

static void heapsort_rc(rowcnt_t *first, long n) {
    ++g_heapsort_calls;
    /* __partial_sort(first, last, last): __heap_select degenerates to __make_heap, then __sort_heap */
    if (n >= 2) {
        long parent = (n - 2) / 2;
        for (;;) {
            rowcnt_t v = first[parent];
            adjust_heap_rc(first, parent, n, v);
            if (parent == 0) break;
            parent--;
        }
    }
    long last = n;
    while (last > 1) {
        --last;
        rowcnt_t v = first[last];
        first[last] = first[0];
        adjust_heap_rc(first, 0, last, v);
    }
}

/* stl_algo.h: __move_median_to_first(result, a, b, c) */
static void median_to_first(rowcnt_t *result, rowcnt_t *a, rowcnt_t *b, rowcnt_t *c) {
    if (longer(a, b)) {
        if (longer(b, c)) swap_rc(result, b);
        else if (longer(a, c)) swap_rc(result, c);
        else swap_rc(result, a);
    } else if (longer(a, c)) swap_rc(result, a);
    else if (longer(b, c)) swap_rc(result, c);
    else swap_rc(result, b);
}

/* stl_algo.h: __unguarded_partition(first, last, pivot) */
static rowcnt_t *unguarded_partition(rowcnt_t *first, rowcnt_t *last, rowcnt_t *pivot) {
    for (;;) {
        while (longer(first, pivot)) ++first;
        --last;
        while (longer(pivot, last)) --last;
        if (!(first < last)) return first;
        swap_rc(first, last);
        ++first;
    }
}

static void introsort_loop(rowcnt_t *first, rowcnt_t *last, long depth_limit) {
    while (last - first > 16) {
        if (depth_limit == 0) {
            heapsort_rc(first, last - first);
            return;
        }
        --depth_limit;
        rowcnt_t *mid = first + (last - first) / 2;
        median_to_first(first, first + 1, mid, last - 1);
        rowcnt_t *cut = unguarded_partition(first + 1, last, first);
        introsort_loop(cut, last, depth_limit);
        last = cut;
    }
}

static void unguarded_linear_insert(rowcnt_t *last) {
    rowcnt_t val = *last;
    rowcnt_t *next = last - 1;
    while (longer(&val, next)) {
        *last = *next;
        last = next;
        --next;
    }
    *last = val;
}

static void insertion_sort(rowcnt_t *first, rowcnt_t *last) {
    if (first == last) return;
    for (rowcnt_t *i = first + 1; i != last; ++i) {
        if (longer(i, first)) {
            rowcnt_t val = *i;
            memmove(first + 1, first, (size_t)(i - first) * sizeof(rowcnt_t));
            *first = val;
        } else
            unguarded_linear_insert(i);
    }
}

static void std_sort_rc(rowcnt_t *first, long n) {
    if (n <= 0) return;
    long lg = 0;
    for (long t = n; t > 1; t >>= 1) ++lg; /* std::__lg */
    introsort_loop(first, first + n, 2 * lg);
    if (n > 16) {
        insertion_sort(first, first + 16);
        for (rowcnt_t *i = first + 16; i != first + n; ++i) unguarded_linear_insert(i);
    } else
        insertion_sort(first, first + n);
}

/* Exposed for direct tests of the std::sort restatement.  idx/cnt are permuted in place. */
void orc_sort_window(long *idx, long *cnt, long n) {
    rowcnt_t *a = (rowcnt_t *)malloc(sizeof(rowcnt_t) * (size_t)(n > 0 ? n : 1));
    for (long i = 0; i < n; ++i) { a[i].idx = idx[i]; a[i].cnt = cnt[i]; }
    std_sort_rc(a, n);
    for (long i = 0; i < n; ++i) { idx[i] = a[i].idx; cnt[i] = a[i].cnt; }
    free(a);
}

/* ---------------------------------------------------------------------------------------------
 * convert_to_scs, structural half  (utilities.hpp:1842-1982, 2060-2069; interface.hpp:401-560)
 *
 *   chunk_ptrs[n_chunks+1], chunk_lengths[n_chunks], old_to_new[n_rows],
 *   new_to_old[n_rows_padded]: the reference leaves positions that no real row maps to
 *   uninitialised (`new int[]`, utilities.hpp:2060-2066); the restatement writes -1 there.
 *   With fixed_perm the struct's own permutation is the identity (utilities.hpp:1911-1928,1976-1982).
 * Returns n_elements, -1 on int overflow of chunk_ptrs, or -2 when fixed_perm sends a non-empty row to a
 * padding position (>= n_rows): the reference then zeroes that count and its fill loop overruns the chunk
 * (utilities.hpp:1919-1922,2029) — undefined behaviour that no oracle can pin.
 * --------------------------------------------------------------------------------------------- */
long orc_scs_structure(long n_rows, long nnz, const int *I, long C, long sigma, const int *fixed_perm,
                       int *chunk_ptrs, int *chunk_lengths, int *old_to_new, int *new_to_old) {
    const long n_chunks = (n_rows + C - 1) / C;
    const long n_pad = n_chunks * C;
    rowcnt_t *rc = (rowcnt_t *)calloc((size_t)(n_pad + sigma + 1), sizeof(rowcnt_t));
    for (long i = 0; i < n_pad; ++i) rc[i].idx = i;
    for (long i = 0; i < nnz; ++i) ++rc[I[i]].cnt;

    if (fixed_perm) {
        for (long i = 0; i < n_rows; ++i)
            if (fixed_perm[i] >= n_rows && rc[i].cnt > 0) { free(rc); return -2; }
        rowcnt_t *tmp = (rowcnt_t *)calloc((size_t)(n_pad + 1), sizeof(rowcnt_t));
        for (long i = 0; i < n_pad; ++i) {
            if (i < n_rows) {
                tmp[i].idx = rc[i].idx;
                tmp[fixed_perm[i]].cnt = rc[i].cnt;
            } else {
                tmp[i].idx = rc[i].idx;
                tmp[i].cnt = rc[i].cnt;
            }
        }
        memcpy(rc, tmp, sizeof(rowcnt_t) * (size_t)n_pad);
        free(tmp);
    } else {
        for (long i = 0; i < n_pad; i += sigma) {
            long end = (i + sigma) < n_pad ? i + sigma : n_pad;
            std_sort_rc(rc + i, end - i);
        }
    }

    long cur = 0;
    int overflow = 0;
    for (long c = 0; c < n_chunks; ++c) {
        long mx = rc[c * C].cnt;
        for (long i = 1; i < C; ++i)
            if (rc[c * C + i].cnt > mx) mx = rc[c * C + i].cnt;
        chunk_lengths[c] = (int)mx;
        chunk_ptrs[c] = (int)cur;
        cur += mx * C;
        if (cur > INT32_MAX) overflow = 1;
    }
    chunk_ptrs[n_chunks] = (int)cur;

    for (long i = 0; i < n_pad; ++i) new_to_old[i] = -1;
    for (long i = 0; i < n_pad; ++i) {
        long old = rc[i].idx;
        if (old < n_rows) old_to_new[old] = (int)i;
    }
    for (long i = 0; i < n_rows; ++i) new_to_old[old_to_new[i]] = (int)i;
    free(rc);
    return overflow ? -1 : cur;
}

/* convert_to_scs, fill half (utilities.hpp:1984-2036): padding value 0 / column 0, then COO order.
 * row_to_pos = old_to_new (sorted mode) or fixed_perm.  `vals` are doubles, narrowed with a single
 * rounding like MtxData::copy's static_cast (classes_structs.hpp:1278). */
void orc_scs_fill(long nnz, const int *I, const int *J, const double *vals, long C, long n_rows_padded,
                  const int *chunk_ptrs, const int *row_to_pos, int vt, long n_elements, int *col_idxs,
                  void *values) {
    int *fillcnt = (int *)calloc((size_t)n_rows_padded + 1, sizeof(int));
    for (long i = 0; i < n_elements; ++i) col_idxs[i] = 0;
    if (vt == VT_F64) memset(values, 0, sizeof(double) * (size_t)n_elements);
    else if (vt == VT_F32) memset(values, 0, sizeof(float) * (size_t)n_elements);
    else memset(values, 0, sizeof(f16) * (size_t)n_elements);
    for (long i = 0; i < nnz; ++i) {
        long row = row_to_pos[I[i]];
        long idx = (long)chunk_ptrs[row / C] + (long)fillcnt[row] * C + row % C;
        col_idxs[idx] = J[i];
        if (vt == VT_F64) ((double *)values)[idx] = vals[i];
        else if (vt == VT_F32) ((float *)values)[idx] = (float)vals[i];
        else ((f16 *)values)[idx] = (f16)vals[i];
        fillcnt[row]++;
    }
    free(fillcnt);
}

/* permute_scs_cols (utilities.hpp:1802-1831): every slot, padding included; halo columns untouched */
void orc_permute_scs_cols(long n_elements, long n_rows, int *col_idxs, const int *perm) {
    for (long i = 0; i < n_elements; ++i)
        if (col_idxs[i] < n_rows) col_idxs[i] = perm[col_idxs[i]];
}

/* ---------------------------------------------------------------------------------------------
 * SpMV / SpMMV kernels
 * --------------------------------------------------------------------------------------------- */

/* spmv_omp_scs / scs_impl_cpu (kernels.hpp:159-258): j outer, lane inner, y in permuted order,
 * n_chunks*C outputs.  hp accumulates in _Float16 with the product+sum formed in float and rounded
 * once per step (g++ -std=c++23 => -fexcess-precision=standard on x86 without AVX512-FP16). */
void orc_spmv_scs(int vt, long C, long n_chunks, const int *cp, const int *cl, const int *ci, const void *vals,
                  const void *x, void *y) {
    for (long c = 0; c < n_chunks; ++c) {
        long cs = cp[c];
        for (long i = 0; i < C; ++i) {
            if (vt == VT_F64) {
                double t = 0.0;
                for (long j = 0; j < cl[c]; ++j) {
                    long e = cs + j * C + i;
                    t = fma(((const double *)vals)[e], ((const double *)x)[ci[e]], t);
                }
                ((double *)y)[c * C + i] = t;
            } else if (vt == VT_F32) {
                float t = 0.0f;
                for (long j = 0; j < cl[c]; ++j) {
                    long e = cs + j * C + i;
                    t = fmaf(((const float *)vals)[e], ((const float *)x)[ci[e]], t);
                }
                ((float *)y)[c * C + i] = t;
            } else {
                f16 t = (f16)0.0f;
                for (long j = 0; j < cl[c]; ++j) {
                    long e = cs + j * C + i;
                    t = (f16)fmaf((float)((const f16 *)vals)[e], (float)((const f16 *)x)[ci[e]], (float)t);
                }
                ((f16 *)y)[c * C + i] = t;
            }
        }
    }
}

/* spmv_omp_csr (kernels.hpp:22-63): sequential in j */
void orc_spmv_csr(int vt, long n_rows, const int *rp, const int *ci, const void *vals, const void *x, void *y) {
    for (long r = 0; r < n_rows; ++r) {
        if (vt == VT_F64) {
            double t = 0.0;
            for (long j = rp[r]; j < rp[r + 1]; ++j) t = fma(((const double *)vals)[j], ((const double *)x)[ci[j]], t);
            ((double *)y)[r] = t;
        } else if (vt == VT_F32) {
            float t = 0.0f;
            for (long j = rp[r]; j < rp[r + 1]; ++j) t = fmaf(((const float *)vals)[j], ((const float *)x)[ci[j]], t);
            ((float *)y)[r] = t;
        } else {
            f16 t = (f16)0.0f;
            for (long j = rp[r]; j < rp[r + 1]; ++j)
                t = (f16)fmaf((float)((const f16 *)vals)[j], (float)((const f16 *)x)[ci[j]], (float)t);
            ((f16 *)y)[r] = t;
        }
    }
}

/* block_spmv_omp_scs_general (kernels.hpp:306-398).  layout 0 = colwise X[col + v*vec_length],
 * 1 = rowwise X[col*bvs + v] (kernels.hpp:352,358,372,378).  block_spmv_omp_scs_adv is NOT restated:
 * its accumulator is never zeroed (kernels.hpp:434). */
void orc_spmmv_scs(int vt, long C, long n_chunks, const int *cp, const int *cl, const int *ci, const void *vals,
                   const void *X, void *Y, int bvs, long vec_length, int layout) {
    for (long c = 0; c < n_chunks; ++c) {
        long cs = cp[c];
        for (long i = 0; i < C; ++i)
            for (int v = 0; v < bvs; ++v) {
                long yo = layout ? (c * C + i) * bvs + v : (c * C + i) + v * vec_length;
                if (vt == VT_F64) {
                    double t = 0.0;
                    for (long j = 0; j < cl[c]; ++j) {
                        long e = cs + j * C + i;
                        long xo = layout ? (long)ci[e] * bvs + v : (long)ci[e] + v * vec_length;
                        t = fma(((const double *)vals)[e], ((const double *)X)[xo], t);
                    }
                    ((double *)Y)[yo] = t;
                } else if (vt == VT_F32) {
                    float t = 0.0f;
                    for (long j = 0; j < cl[c]; ++j) {
                        long e = cs + j * C + i;
                        long xo = layout ? (long)ci[e] * bvs + v : (long)ci[e] + v * vec_length;
                        t = fmaf(((const float *)vals)[e], ((const float *)X)[xo], t);
                    }
                    ((float *)Y)[yo] = t;
                } else {
                    f16 t = (f16)0.0f;
                    for (long j = 0; j < cl[c]; ++j) {
                        long e = cs + j * C + i;
                        long xo = layout ? (long)ci[e] * bvs + v : (long)ci[e] + v * vec_length;
                        t = (f16)fmaf((float)((const f16 *)vals)[e], (float)((const f16 *)X)[xo], (float)t);
                    }
                    ((f16 *)Y)[yo] = t;
                }
            }
    }
}

/* block_spmv_omp_csr (kernels.hpp:68-154) */
void orc_spmmv_csr(int vt, long n_rows, const int *rp, const int *ci, const void *vals, const void *X, void *Y,
                   int bvs, long vec_length, int layout) {
    orc_spmmv_scs(vt, 1, n_rows, rp, NULL, ci, vals, X, Y, bvs, vec_length, layout);
}

/* ---------------------------------------------------------------------------------------------
 * Adaptive precision (canonical = interface.hpp:1434-1733 for SCS, :1129-1429 for CRS, and
 * ap_kernels.hpp:24-82 for the harness' dp+sp).  mode: 0 dp_sp, 1 dp_hp, 2 sp_hp, 3 dp_sp_hp.
 * Every part has its own chunk_ptrs/chunk_lengths/col_idxs/values and shares C and n_chunks.
 * Modes 0,1,3: products and accumulators are fp64 against dp_x, y(double) = dp + sp + hp in that
 * association.  Mode 2: products are fp32 (float*float, _Float16*float) against sp_x, accumulators
 * fp64, y(float) = (float)(sp + hp)  (interface.hpp:1620-1645).
 * --------------------------------------------------------------------------------------------- */
void orc_ap_scs(int mode, long C, long n_chunks,
                const int *dcp, const int *dcl, const int *dci, const double *dv,
                const int *scp, const int *scl, const int *sci, const float *sv,
                const int *hcp, const int *hcl, const int *hci, const void *hv_,
                const double *dp_x, const float *sp_x, void *y) {
    const f16 *hv = (const f16 *)hv_;
    const int use_dp = (mode != 2), use_sp = (mode != 1), use_hp = (mode != 0);
    for (long c = 0; c < n_chunks; ++c)
        for (long i = 0; i < C; ++i) {
            double dt = 0.0, st = 0.0, ht = 0.0;
            if (use_dp)
                for (long j = 0; j < dcl[c]; ++j) {
                    long e = (long)dcp[c] + j * C + i;
                    dt = fma(dv[e], dp_x[dci[e]], dt);
                }
            if (use_sp)
                for (long j = 0; j < scl[c]; ++j) {
                    long e = (long)scp[c] + j * C + i;
                    if (mode == 2) st += (double)(sv[e] * sp_x[sci[e]]);
                    else st = fma((double)sv[e], dp_x[sci[e]], st);
                }
            if (use_hp)
                for (long j = 0; j < hcl[c]; ++j) {
                    long e = (long)hcp[c] + j * C + i;
                    if (mode == 2) ht += (double)((float)hv[e] * sp_x[hci[e]]);
                    else ht = fma((double)hv[e], dp_x[hci[e]], ht);
                }
            if (mode == 2) ((float *)y)[c * C + i] = (float)(st + ht);
            else if (mode == 0) ((double *)y)[c * C + i] = dt + st;
            else if (mode == 1) ((double *)y)[c * C + i] = dt + ht;
            else ((double *)y)[c * C + i] = dt + st + ht;
        }
}

/* ---------------------------------------------------------------------------------------------
 * partition_precisions (interface.hpp:690-978; harness twin utilities.hpp:2810-3123).
 * part[i] = 0 dp, 1 sp, 2 hp.  2-way: abs(v) >= t1 -> high part else low part.  3-way: abs(v) >= t1
 * -> dp; t2 <= abs(v) <= t1 -> sp; else hp (interface.hpp:938-964).  With rowmax/colmax
 * (equilibrated) thresholds are divided by colmax[J]*rowmax[I] (interface.hpp:908-936).
 * Order inside each part is input order, so the part ids are the whole result.
 * --------------------------------------------------------------------------------------------- */
void orc_partition_precisions(int mode, long nnz, const int *I, const int *J, const double *vals, double t1, double t2,
                              const double *rowmax, const double *colmax, int8_t *part, long *counts3) {
    counts3[0] = counts3[1] = counts3[2] = 0;
    const int hi = (mode == 2) ? 1 : 0;
    const int lo = (mode == 0) ? 1 : 2;
    for (long i = 0; i < nnz; ++i) {
        double a = fabs(vals[i]);
        double th1 = t1, th2 = t2;
        if (rowmax && colmax) {
            double s = colmax[J[i]] * rowmax[I[i]];
            th1 = t1 / s;
            th2 = t2 / s;
        }
        int p;
        if (mode == 3) p = (a >= th1) ? 0 : ((a <= th1 && a >= th2) ? 1 : 2);
        else p = (a >= th1) ? hi : lo;
        part[i] = (int8_t)p;
        counts3[p]++;
    }
}

/* extract_largest_{row,col}_elems (utilities.hpp:2605-2643); outputs must be zero-initialised */
void orc_largest_elems(long nnz, const int *I, const int *J, const double *vals, double *rowmax, double *colmax) {
    for (long i = 0; i < nnz; ++i) {
        double a = fabs(vals[i]);
        if (a > rowmax[I[i]]) rowmax[I[i]] = a;
        if (a > colmax[J[i]]) colmax[J[i]] = a;
    }
}

/* equilibrate_matrix (utilities.hpp:2668-2684): rows scaled by their largest |value| (:2646-2654), then columns by the largest
 * |row-scaled value| (:2657-2665).  rowmax / colmax (n_rows / n_cols, zero-initialised) return the two sets of maxima. */
void orc_equilibrate(long nnz, const int *I, const int *J, double *vals, double *rowmax, double *colmax) {
    for (long i = 0; i < nnz; ++i) {
        double a = fabs(vals[i]);
        if (a > rowmax[I[i]]) rowmax[I[i]] = a;
    }
    for (long i = 0; i < nnz; ++i) vals[i] = vals[i] / rowmax[I[i]];
    for (long i = 0; i < nnz; ++i) {
        double a = fabs(vals[i]);
        if (a > colmax[J[i]]) colmax[J[i]] = a;
    }
    for (long i = 0; i < nnz; ++i) vals[i] = vals[i] / colmax[J[i]];
}

/* read_mtx post-processing (utilities.hpp:2214-2290): entries in file order; symmetric files are expanded (i,j) -> (i,j),(j,i) for
 * i != j (:2237-2251), then a stable sort by row (:2278, sort_perm :2139-2146).  Outputs need 2*nz capacity; returns nnz. */
long orc_ingest_entries(long nz, const int *I, const int *J, const double *V, int symmetric, int n_rows, int *Io, int *Jo, double *Vo) {
    long nnz = 0;
    int *ti = (int *)malloc(sizeof(int) * (size_t)(2 * nz + 1)), *tj = (int *)malloc(sizeof(int) * (size_t)(2 * nz + 1));
    double *tv = (double *)malloc(sizeof(double) * (size_t)(2 * nz + 1));
    for (long k = 0; k < nz; ++k) {
        ti[nnz] = I[k]; tj[nnz] = J[k]; tv[nnz] = V[k]; ++nnz;
        if (symmetric && I[k] != J[k]) { ti[nnz] = J[k]; tj[nnz] = I[k]; tv[nnz] = V[k]; ++nnz; }
    }
    /* stable counting sort by row == std::stable_sort with the row comparator */
    long *start = (long *)calloc((size_t)n_rows + 1, sizeof(long));
    for (long k = 0; k < nnz; ++k) start[ti[k] + 1]++;
    for (int r = 0; r < n_rows; ++r) start[r + 1] += start[r];
    for (long k = 0; k < nnz; ++k) {
        long p = start[ti[k]]++;
        Io[p] = ti[k]; Jo[p] = tj[k]; Vo[p] = tv[k];
    }
    free(ti); free(tj); free(tv); free(start);
    return nnz;
}

/* ---------------------------------------------------------------------------------------------
 * Row partitioning + halo bookkeeping
 * --------------------------------------------------------------------------------------------- */

/* seg_work_sharing_arr (mpi_funcs.hpp:424-622), seg_method 0 = seg-rows, 1 = seg-nnz */
void orc_seg_work_sharing_arr(int seg_method, long n_rows, long nnz, const int *I, int P, int *wsa) {
    for (int s = 0; s <= P; ++s) wsa[s] = 0;
    if (seg_method == 0) {
        int per = (int)(n_rows / P);
        for (int s = 1; s <= P; ++s) wsa[s] = s * per;
        wsa[P] = I[nnz - 1] + 1;
    } else {
        int per = (int)(nnz / P);
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
    }
    if (wsa[P - 1] == wsa[P])
        for (int r = 1; r < P; ++r) wsa[r] -= 1;
}

/* collect_local_needed_heri (mpi_funcs.hpp:242-415).  col_idxs (global columns) is rewritten in place:
 * local columns -> col - wsa[rank]; distinct remote columns -> n_local + (first-seen order grouped by
 * owner, lower owners first).  need_flat/need_ptr: per-owner lists of owner-local indices.
 * recv_counts_cumsum[P+1].  Returns the number of halo elements, or -needed if need_cap is too small. */
long orc_collect_halo(long n_elements, int *col_idxs, const int *wsa, int rank, int P, int *need_flat, long need_cap,
                      int *need_ptr, int *recv_counts_cumsum) {
    const int lo = wsa[rank], hi = wsa[rank + 1];
    const int n_local = hi - lo;
    const int n_glob = wsa[P];
    /* first-seen rank of each remote column, -1 = unseen (dense map over global columns) */
    int *seen = (int *)malloc(sizeof(int) * (size_t)(n_glob > 0 ? n_glob : 1));
    for (int i = 0; i < n_glob; ++i) seen[i] = -1;
    int *cnt = (int *)calloc((size_t)P + 1, sizeof(int));
    int *owner_of = (int *)malloc(sizeof(int) * (size_t)(n_glob > 0 ? n_glob : 1));
    for (int p = 0; p < P; ++p)
        for (int c = wsa[p]; c < wsa[p + 1]; ++c) owner_of[c] = p;
    /* pass 1: count per owner in first-seen order, remember the within-owner ordinal */
    for (long i = 0; i < n_elements; ++i) {
        int col = col_idxs[i];
        if (col >= lo && col <= hi - 1) continue;
        if (col < 0 || col >= n_glob) continue; /* no owner: the reference leaves such a column untouched */
        if (seen[col] < 0) seen[col] = cnt[owner_of[col]]++;
    }
    long total = 0;
    for (int p = 0; p < P; ++p) { need_ptr[p] = (int)total; total += cnt[p]; }
    need_ptr[P] = (int)total;
    for (int p = 0; p <= P; ++p) recv_counts_cumsum[p] = need_ptr[p];
    if (total > need_cap) { free(seen); free(cnt); free(owner_of); return -total; }
    /* pass 2: rewrite */
    for (long i = 0; i < n_elements; ++i) {
        int col = col_idxs[i];
        if (col >= lo && col <= hi - 1) { col_idxs[i] = col - lo; continue; }
        if (col < 0 || col >= n_glob) continue;
        int p = owner_of[col];
        int slot = need_ptr[p] + seen[col];
        need_flat[slot] = col - wsa[p];
        col_idxs[i] = n_local + slot;
    }
    free(seen); free(cnt); free(owner_of);
    return total;
}

/* pack_send_buf (classes_structs.hpp:813-831, kernels.hpp:554-577): buf[i] = x[perm[send_idx[i]] + off] */
void orc_pack_f64(long n, const int *send_idxs, const int *perm, const double *x, double *buf) {
    for (long i = 0; i < n; ++i) buf[i] = x[perm[send_idxs[i]]];
}

/* Synthetic stencil COO (BASELINE.json configs 2/3/5): rows [row0,row1) of an nx*ny*nz
 * grid, row = (z*ny + y)*nx + x, columns ascending, Dirichlet, diagonal = points-1, off-diagonals -1; local row
 * ids, global column ids.  Call with I == NULL to get the count.  Used to feed the reference's own builder in
 * bench.py's CPU legs; the product has its own device generator (uspmv_coo_stencil). */
long orc_stencil_coo(int points, long nx, long ny, long nz, long row0, long row1, int *I, int *J, double *V) {
    long o = 0;
    for (long g = row0; g < row1; ++g) {
        long x = g % nx, y = (g / nx) % ny, z = g / (nx * ny);
        for (int dz = -1; dz <= 1; ++dz) {
            long zz = z + dz;
            if (zz < 0 || zz >= nz) continue;
            for (int dy = -1; dy <= 1; ++dy) {
                long yy = y + dy;
                if (yy < 0 || yy >= ny) continue;
                for (int dx = -1; dx <= 1; ++dx) {
                    long xx = x + dx;
                    if (xx < 0 || xx >= nx) continue;
                    int nzd = (dx != 0) + (dy != 0) + (dz != 0);
                    if (points == 7 && nzd > 1) continue;
                    if (I) {
                        I[o] = (int)(g - row0);
                        J[o] = (int)((zz * ny + yy) * nx + xx);
                        V[o] = nzd == 0 ? (double)(points - 1) : -1.0;
                    }
                    ++o;
                }
            }
        }
    }
    return o;
}
