// uspmv — harness CLI clone of the reference's benchmark driver (code/main.cpp) on top of libuspmv_b200.so.
//
//   ./uspmv <matrix.mtx | gen:laplace7:<n> | gen:stencil27:<n>> <crs|csr|scs> [options]
//
// Same positional arguments, flags (dash and underscore spellings), defaults, rejections and report file
// (spmv_bench.txt, write_results.hpp:43-157) as the reference (utilities.hpp:983-1545, classes_structs.hpp:47-153).
// Host side is plain C++ (no CUDA headers); everything numerical goes through the C ABI (include/uspmv_b200.h).
// Protocol of bench mode (main.cpp:380-527): x = 5.0 (or random / matrix mean) in permuted space, 100 warm-up SpMVs,
// then n_iter = 2,4,8,... until one loop takes >= bench_time; Gflops = 2*nnz*block_vec_size / (t / n_iter) / 1e9.
// Solve mode (main.cpp:528-631): `rev` x { SpMV ; y becomes the next x }, result un-permuted and compared with a
// host-side COO product (the reference needs MKL for this step, write_results.hpp:442-556).
// Multi-GPU: `-gpus N` (or USPMV_NUM_GPUS=N) stands in for `mpirun -n N ./uspmv ...`: the process forks one rank per GPU; the ranks
// partition the rows (-seg_rows / -seg_nnz, identical work_sharing_arr), build their slab's SELL-C-sigma, discover the halo on the
// device, trade need lists and CUDA-IPC handles through a shared-memory segment (what the reference does with MPI_Alltoallv /
// MPI_Bcast at setup, mpi_funcs.hpp:117-232,1061-1124) and then run the SpMV loop with the NVLink peer-to-peer exchange
// (uspmv_p2p_*): no MPI, no NCCL, no Python.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <new>
#include <numeric>
#include <random>
#include <sstream>
#include <string>
#include <vector>

#include <pthread.h>
#include <signal.h>
#include <sys/mman.h>
#include <sys/wait.h>
#include <unistd.h>

#include "../../include/uspmv_b200.h"
#include "result_report.hpp"

using ST = long;

namespace {

struct Config {  // classes_structs.hpp:47-153
    long chunk_size = 1, sigma = 1;
    char random_init_x = '0';
    unsigned long n_repetitions = 1;
    int validate_result = 1, verbose = 0, block_vec_size = 1, comm_halos = 1, ba_synch = 1, par_pack = 0, no_pack = 0, print_comm_vol = 0,
        equilibrate = 0, n_gpus = 1;
    char mode = 'b';
    double bench_time = 5.0, ap_threshold_1 = 0.0, ap_threshold_2 = 0.0, dropout = 0, dropout_threshold = 0.0;
    double matrix_min = 1.0, matrix_mean = 1.0, matrix_max = 1.0;
    std::string matrix_file_name, seg_method = "seg-rows", value_type = "dp", kernel_format = "scs", block_vec_layout = "colwise";
    std::string output_filename_bench = "spmv_bench.txt";
};

[[noreturn]] void die(const char *msg) {
    fprintf(stderr, "%s\n", msg);
    exit(1);
}
void ck(int rc) {
    if (rc) {
        fprintf(stderr, "uspmv ERROR: %s\n", uspmv_last_error());
        exit(1);
    }
}

void usage(const char *argv0, const Config &c) {
    fprintf(stderr,
            "Usage: %s <martix-market-filename | gen:laplace7:<n> | gen:stencil27:<n>> <kernel-format> [options]\n "
            "options [defaults] (description): \n"
            "-block_vec_size [%i] (int: width of block vectors for SpMMV) \n"
            "-block_vec_layout [%s] (colwise/rowwise: block vector layout; a compile-time switch in the reference) \n"
            "-c [%li] (int: chunk size (required for scs)) \n"
            "-s [%li] (int: sigma (required for scs)) \n"
            "-rev [%li] (int: number of back-to-back revisions to perform) \n"
            "-rand_x [%c] (0/1: random x vector option) \n"
            "-dp / sp / hp / ap[dp_sp] / ap[dp_hp] / ap[sp_hp] / ap[dp_sp_hp] [%s] (numerical precision of matrix data) \n"
            "-gpus [1] (ranks = GPUs of this node, replaces `mpirun -n`; env USPMV_NUM_GPUS) \n"
            "-seg_metis / seg_nnz / seg_rows [%s] (global matrix partitioning over the ranks) \n"
            "-validate [%i] (0/1: check result against a host COO product in solve mode) \n"
            "-verbose [%i] (0/1: verbose validation of results) \n"
            "-mode [%c] ('s'/'b': either in solve mode or bench mode) \n"
            "-bench_time [%g] (float: minimum number of seconds for SpMV benchmark) \n"
            "-ba_synch [%i] -comm_halos [%i] -par_pack [%i] -no_pack [%i] (accepted for compatibility; single process) \n"
            "-equilibrate [%i] (0/1: normalize rows of matrix) \n"
            "--------------------------- Adaptive Precision Options --------------------------- \n"
            "-ap_threshold_1 [%f] (float: threshold for two-way matrix partitioning for adaptive precision `-ap`) \n"
            "-ap_threshold_2 [%f] (float: threshold for three-way matrix partitioning for adaptive precision `-ap`) \n"
            "-dropout[%f] (0/1: enable dropout of elements below theh designated threshold) \n"
            "-dropout_threshold [%f] (float: remove matrix elements below this range) \n\n",
            argv0, c.block_vec_size, c.block_vec_layout.c_str(), c.chunk_size, c.sigma, (long)c.n_repetitions, c.random_init_x,
            c.value_type.c_str(), c.seg_method.c_str(), c.validate_result, c.verbose, c.mode, c.bench_time, c.ba_synch, c.comm_halos,
            c.par_pack, c.no_pack, c.equilibrate, c.ap_threshold_1, c.ap_threshold_2, c.dropout, c.dropout_threshold);
}

bool is_ap(const std::string &v) { return v.rfind("ap[", 0) == 0; }

void parse_cli(int argc, char **argv, Config &c) {  // utilities.hpp:1047-1545
    if (argc < 3) { usage(argv[0], c); exit(1); }
    c.matrix_file_name = argv[1];
    c.kernel_format = argv[2];
    if (const char *e = getenv("USPMV_NUM_GPUS")) c.n_gpus = std::max(1, std::min(16, atoi(e)));
    auto need = [&](int &i) -> const char * {
        if (i + 1 >= argc) { fprintf(stderr, "ERROR: missing value for %s\n", argv[i]); usage(argv[0], c); exit(1); }
        return argv[++i];
    };
    auto bad = [&](const char *m) { fprintf(stderr, "%s\n", m); usage(argv[0], c); exit(1); };
    for (int i = 3; i < argc; ++i) {
        std::string a = argv[i];
        std::replace(a.begin() + 1, a.end(), '-', '_');  // "-bench-time" == "-bench_time"
        if (a == "-c") { c.chunk_size = atoi(need(i)); if (c.chunk_size < 1) bad("ERROR: chunk size must be >= 1."); }
        else if (a == "-s") { c.sigma = atoi(need(i)); if (c.sigma < 1) bad("ERROR: sigma must be >= 1."); }
        else if (a == "-block_vec_size") { c.block_vec_size = atoi(need(i)); if (c.block_vec_size < 1) bad("ERROR: block_vec_size must be >= 1."); }
        else if (a == "-block_vec_layout") { c.block_vec_layout = need(i); if (c.block_vec_layout != "colwise" && c.block_vec_layout != "rowwise") bad("ERROR: block_vec_layout must be colwise or rowwise."); }
        else if (a == "-bench_time") { c.bench_time = atof(need(i)); if (c.bench_time < 0) bad("ERROR: bench_time must be > 0."); }
        else if (a == "-rev") { long r = atol(need(i)); if (r < 1) bad("ERROR: revisions must be >= 1."); c.n_repetitions = r; }
        else if (a == "-verbose") { c.verbose = atoi(need(i)); if (c.verbose != 0 && c.verbose != 1) bad("ERROR: Only validation verbosity levels 0 and 1 are supported."); }
        else if (a == "-validate") { c.validate_result = atoi(need(i)); if (c.validate_result != 0 && c.validate_result != 1) bad("ERROR: You can only choose to validate result (1, i.e. yes) or not (0, i.e. no)."); }
        else if (a == "-mode") { c.mode = need(i)[0]; if (c.mode != 'b' && c.mode != 's') bad("ERROR: Only bench (b) and solve (s) modes are supported."); }
        else if (a == "-rand_x") { c.random_init_x = need(i)[0]; if (c.random_init_x != '0' && c.random_init_x != '1' && c.random_init_x != 'm') bad("ERROR: You can only choose to initialize x randomly (1), with the default value (0) or the matrix mean (m)."); }
        else if (a == "-comm_halos") c.comm_halos = atoi(need(i));
        else if (a == "-ba_synch") c.ba_synch = atoi(need(i));
        else if (a == "-par_pack") c.par_pack = atoi(need(i));
        else if (a == "-no_pack") c.no_pack = atoi(need(i));
        else if (a == "-print_comm_vol") c.print_comm_vol = atoi(need(i));
        else if (a == "-ap_threshold_1" || a == "-apt1") { c.ap_threshold_1 = atof(need(i)); if (c.ap_threshold_1 < 0) bad("ERROR: ap_threshold_1 must be nonnegative."); }
        else if (a == "-ap_threshold_2" || a == "-apt2") { c.ap_threshold_2 = atof(need(i)); if (c.ap_threshold_2 < 0) bad("ERROR: ap_threshold_2 must be nonnegative."); }
        else if (a == "-dropout" || a == "-do") c.dropout = atof(need(i));
        else if (a == "-dropout_threshold" || a == "-dt") c.dropout_threshold = atof(need(i));
        else if (a == "-equilibrate") c.equilibrate = atoi(need(i));
        else if (a == "-dp" || a == "-sp" || a == "-hp") c.value_type = a.substr(1);
        else if (a == "-ap[dp_sp]" || a == "-ap[sp_hp]" || a == "-ap[dp_hp]" || a == "-ap[dp_sp_hp]") c.value_type = a.substr(1);
        else if (a == "-seg_rows") c.seg_method = "seg-rows";
        else if (a == "-seg_nnz") c.seg_method = "seg-nnz";
        else if (a == "-seg_metis") c.seg_method = "seg-metis";
        else if (a == "-gpus") { c.n_gpus = atoi(need(i)); if (c.n_gpus < 1 || c.n_gpus > 16) bad("ERROR: -gpus must be in [1,16]."); }
        else { fprintf(stderr, "ERROR: unknown argument: %s\n", argv[i]); usage(argv[0], c); exit(1); }
    }
    // sanity checks, utilities.hpp:1371-1545
    if (c.block_vec_layout == "rowwise" && c.block_vec_size == 1)
        die("ERROR: Row-wise block vector layout selected, but block vector width is 1.\n Please choose colwise block vector layout if using SpMV.");
    if (c.block_vec_size > 1 && is_ap(c.value_type)) die("ERROR: SpMMV is not yet implemented for AP kernels.");
    if (c.block_vec_size > 16) die("ERROR: block_vec_size > 16 is not supported by the GPU SpMMV kernels.");
    if (c.seg_method == "seg-metis") die("ERROR: seg-metis selected, but USE_METIS not defined in Makefile.");
    if (!is_ap(c.value_type) && c.ap_threshold_1 > 0.0) fprintf(stderr, "WARNING: First adaptive precision threshold entered, but not used.\n");
    if (c.value_type != "ap[dp_sp_hp]" && c.ap_threshold_2 > 0.0)
        fprintf(stderr, "WARNING: Second adaptive precision threshold entered, but three-way partitioning is not used.\n");
    if ((c.value_type == "ap[dp_sp]" || c.value_type == "ap[sp_hp]" || c.value_type == "ap[dp_hp]") && c.ap_threshold_1 == 0.0)
        fprintf(stderr, "WARNING: Two-way adaptive precision used, but the first threshold is not entered.\n");
    if (c.value_type == "ap[dp_sp_hp]") {
        if (c.ap_threshold_1 == 0.0) fprintf(stderr, "WARNING: Three-way adaptive precision used, but the first threshold is not entered.\n");
        if (c.ap_threshold_2 == 0.0) fprintf(stderr, "WARNING: Three-way adaptive precision used, but the second threshold is not entered.\n");
        if (c.ap_threshold_1 <= c.ap_threshold_2) die("ERROR: Three-way adaptive precision used, but the second threshold is larger than the first.");
    }
    if (c.dropout && c.dropout_threshold == 0.0) fprintf(stderr, "WARNING: Dropout selected, but dropout_threshold is 0.\n");
    if (c.kernel_format != "crs" && c.kernel_format != "csr" && c.kernel_format != "scs") die("ERROR: kernel format not recognized.");
    if (c.kernel_format != "scs") { c.chunk_size = 1; c.sigma = 1; }  // CRS is the C = 1, sigma = 1 instance
    if (c.n_gpus == 1) {
        printf("Single process: forcing comm_halos = 0.\n");
        c.comm_halos = 0;
    } else {
        if (c.comm_halos != 1) die("ERROR: -gpus N > 1 always exchanges the halo (comm_halos = 1): the kernels read remote x elements.");
        if (c.equilibrate && is_ap(c.value_type)) die("ERROR: -equilibrate with adaptive precision is single-GPU only.");
    }
}

struct Coo {
    ST n_rows = 0, n_cols = 0;
    std::vector<int> I, J;
    std::vector<double> V;
};

// read_mtx (utilities.hpp:2148-2309 + mmio.h:138-263): real/integer/pattern, general/symmetric.  Only the text is parsed here;
// the symmetric expansion ((i,j) immediately followed by (j,i)) and the stable sort by row run on the device
// (uspmv_coo_from_entries).
struct MtxFile {
    Coo entries;  // file order, 0-based
    bool symmetric = false;
};
MtxFile read_mtx(const std::string &path) {
    std::ifstream f(path);
    if (!f) die("Unable to open file");
    std::string line;
    std::getline(f, line);
    std::istringstream b(line);
    std::string banner, obj, fmt, field, symm;
    b >> banner >> obj >> fmt >> field >> symm;
    auto lower = [](std::string s) { std::transform(s.begin(), s.end(), s.begin(), ::tolower); return s; };
    obj = lower(obj); fmt = lower(fmt); field = lower(field); symm = lower(symm);
    if (banner != "%%MatrixMarket") die("mm_read_unsymetric: Could not process Matrix Market banner");
    if (obj != "matrix" || fmt != "coordinate") die("The matrix market file provided is not supported.\n Reason :\n * matrix has to be sparse");
    if (field != "real" && field != "integer" && field != "pattern") die("The matrix market file provided is not supported.\n Reason :\n * matrix has to be real or pattern");
    if (symm != "general" && symm != "symmetric") die("The matrix market file provided is not supported.\n Reason :\n * matrix has to be either general or symmetric");
    do { if (!std::getline(f, line)) die("read_unsymmetric_sparse(): could not parse matrix size."); } while (line.empty() || line[0] == '%');
    long M, N, nz;
    { std::istringstream s(line); if (!(s >> M >> N >> nz)) die("read_unsymmetric_sparse(): could not parse matrix size."); }
    if (M != N) die("Matrix not square. Currently only square matrices are supported");
    MtxFile m;
    m.symmetric = symm == "symmetric";
    m.entries.n_rows = M; m.entries.n_cols = N;
    m.entries.I.reserve(nz); m.entries.J.reserve(nz); m.entries.V.reserve(nz);
    const bool pattern = field == "pattern";
    for (long k = 0; k < nz; ++k) {
        long i, j;
        double v = 0.01;  // pattern matrices: mmio.h:195-203
        if (!(f >> i >> j)) die("Error in file reading");
        if (!pattern && !(f >> v)) die("Error in file reading");
        m.entries.I.push_back((int)i - 1); m.entries.J.push_back((int)j - 1); m.entries.V.push_back(v);
    }
    return m;
}

double now() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);  // timing.c:3-8
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

size_t vt_bytes(int vt) { return vt == USPMV_F64 ? 8 : vt == USPMV_F32 ? 4 : 2; }

// double -> fp16 bits (round to nearest even), for uploading x in half precision without _Float16 support
unsigned short f2h(double d) {
#if defined(__FLT16_MAX__)
    _Float16 h = (_Float16)d;
    unsigned short b;
    memcpy(&b, &h, 2);
    return b;
#else
    float f = (float)d;
    unsigned int x;
    memcpy(&x, &f, 4);
    unsigned int sign = (x >> 16) & 0x8000u;
    int e = (int)((x >> 23) & 0xff) - 127 + 15;
    unsigned int m = x & 0x7fffffu;
    if (e <= 0) return (unsigned short)sign;
    if (e >= 31) return (unsigned short)(sign | 0x7c00u);
    unsigned int h = sign | (e << 10) | (m >> 13);
    if ((m & 0x1000u) && ((m & 0x2fffu) != 0)) ++h;
    return (unsigned short)h;
#endif
}
double h2d(unsigned short b) {
#if defined(__FLT16_MAX__)
    _Float16 h;
    memcpy(&h, &b, 2);
    return (double)h;
#else
    int s = (b >> 15) & 1, e = (b >> 10) & 31, m = b & 1023;
    double v = e == 0 ? std::ldexp((double)m, -24) : e == 31 ? (m ? NAN : INFINITY) : std::ldexp((double)(m + 1024), e - 25);
    return s ? -v : v;
#endif
}

void upload(uspmv_ctx *ctx, void *dst, const std::vector<double> &src, int vt) {
    const size_t n = src.size();
    if (vt == USPMV_F64) ck(uspmv_memcpy_h2d(ctx, dst, src.data(), n * 8, nullptr));
    else if (vt == USPMV_F32) { std::vector<float> t(src.begin(), src.end()); ck(uspmv_memcpy_h2d(ctx, dst, t.data(), n * 4, nullptr)); }
    else { std::vector<unsigned short> t(n); for (size_t i = 0; i < n; ++i) t[i] = f2h(src[i]); ck(uspmv_memcpy_h2d(ctx, dst, t.data(), n * 2, nullptr)); }
}
std::vector<double> download(uspmv_ctx *ctx, const void *src, size_t n, int vt) {
    std::vector<double> out(n);
    if (vt == USPMV_F64) ck(uspmv_memcpy_d2h(ctx, out.data(), src, n * 8, nullptr));
    else if (vt == USPMV_F32) { std::vector<float> t(n); ck(uspmv_memcpy_d2h(ctx, t.data(), src, n * 4, nullptr)); std::copy(t.begin(), t.end(), out.begin()); }
    else { std::vector<unsigned short> t(n); ck(uspmv_memcpy_d2h(ctx, t.data(), src, n * 2, nullptr)); for (size_t i = 0; i < n; ++i) out[i] = h2d(t[i]); }
    return out;
}


// =====================================================================================================================
// Multi-GPU harness: one forked rank per GPU (stands in for the reference's MPI build, main.cpp:1035-1560).
// =====================================================================================================================
constexpr int MAXP = 16;
struct RankInfo {
    long n_local, n_halo, x_bytes, vec_length, nnz, n_elements, n_chunks, n_pad, part_nnz[3];
    double runtime;
    int recv_cumsum[MAXP + 1], need_ptr[MAXP + 1];
    unsigned char ipc[64];
};
struct Shared {
    pthread_barrier_t bar;
    RankInfo info[MAXP];
    long cap;        // ints per rank in the need-list area
    long y_len;      // doubles in the result area
};
inline int *need_area(Shared *sh, int q) { return reinterpret_cast<int *>(reinterpret_cast<unsigned char *>(sh) + 65536) + (size_t)q * sh->cap; }
inline double *y_area(Shared *sh, int P) {  // behind the need lists, 8-byte aligned
    const size_t off = (65536 + (size_t)P * sh->cap * sizeof(int) + 7) / 8 * 8;
    return reinterpret_cast<double *>(reinterpret_cast<unsigned char *>(sh) + off);
}

// global problem size without touching CUDA (the parent must not create a context before fork)
long peek_rows(const Config &cfg) {
    if (cfg.matrix_file_name.rfind("gen:", 0) == 0) {
        char kind[32];
        long n = 0;
        if (sscanf(cfg.matrix_file_name.c_str(), "gen:%31[^:]:%ld", kind, &n) != 2 || n < 1) die("ERROR: generator syntax is gen:laplace7:<n> or gen:stencil27:<n>");
        return n * n * n;
    }
    std::ifstream f(cfg.matrix_file_name);
    if (!f) die("Unable to open file");
    std::string line;
    std::getline(f, line);
    do { if (!std::getline(f, line)) die("read_unsymmetric_sparse(): could not parse matrix size."); } while (line.empty() || line[0] == '%');
    long M = 0, N = 0, nz = 0;
    std::istringstream s(line);
    if (!(s >> M >> N >> nz)) die("read_unsymmetric_sparse(): could not parse matrix size.");
    return M;
}

int rank_main(Config cfg, const int rank, const int P, Shared *sh) {
    auto barrier = [&] { pthread_barrier_wait(&sh->bar); };
    uspmv_ctx *ctx = nullptr;
    int ndev = 0;
    ck(uspmv_device_count(&ndev));
    // cudaSetDevice(my_rank % ndev) like the reference (main.cpp:1838-1842): more ranks than GPUs share devices (time-sliced; the
    // arenas are still exchanged through CUDA IPC, so the whole push / wait / acknowledge protocol runs with a real peer)
    ck(uspmv_ctx_create(rank % ndev, &ctx));
    const bool ap = is_ap(cfg.value_type);
    const int ap_mode = cfg.value_type == "ap[dp_sp]" ? USPMV_AP_DP_SP : cfg.value_type == "ap[dp_hp]" ? USPMV_AP_DP_HP
                        : cfg.value_type == "ap[sp_hp]" ? USPMV_AP_SP_HP : USPMV_AP_DP_SP_HP;
    const int vt = cfg.value_type == "sp" ? USPMV_F32 : cfg.value_type == "hp" ? USPMV_F16 : USPMV_F64;
    const int seg = cfg.seg_method == "seg-nnz" ? USPMV_SEG_NNZ : USPMV_SEG_ROWS;

    // ---- global matrix -> work_sharing_arr -> this rank's slab (seg_work_sharing_arr, seg_mtx_struct, localize_row_idx) ------------
    std::vector<int> wsa(P + 1, 0);
    uspmv_coo *coo = nullptr;  // rows [wsa[rank], wsa[rank+1]), local row ids, GLOBAL columns
    Coo host;                  // the whole matrix on the host (files only; rank 0 validates against it)
    bool have_host = false;
    long n_glob = 0;
    const bool gen = cfg.matrix_file_name.rfind("gen:", 0) == 0;
    int pts = 0;
    long gn = 0;
    if (gen) {
        char kind[32];
        sscanf(cfg.matrix_file_name.c_str(), "gen:%31[^:]:%ld", kind, &gn);
        pts = !strcmp(kind, "laplace7") ? 7 : !strcmp(kind, "stencil27") ? 27 : 0;
        if (!pts) die("ERROR: unknown generator");
        n_glob = gn * gn * gn;
        cfg.matrix_min = -1.0; cfg.matrix_max = pts - 1.0; cfg.matrix_mean = 0.0;
    }
    if (gen && seg == USPMV_SEG_ROWS && !cfg.equilibrate) {
        // seg-rows needs no look at the matrix (mpi_funcs.hpp:446-465): floor(n/P) rows per rank, the remainder on the last one —
        // every rank generates only its own rows, so matrices that do not fit one GPU (512^3 27-point) can be run
        if (n_glob > 2147483647L) die("ERROR: more than 2^31-1 rows (IT = int, like the reference)");
        for (int k = 0; k < P; ++k) wsa[k] = (int)(k * (n_glob / P));
        wsa[P] = (int)n_glob;
        ck(uspmv_coo_stencil(ctx, pts, gn, gn, gn, wsa[rank], wsa[rank + 1], &coo));
    } else {
        uspmv_coo *full = nullptr;
        if (gen) ck(uspmv_coo_stencil(ctx, pts, gn, gn, gn, 0, n_glob, &full));
        else {
            MtxFile file = read_mtx(cfg.matrix_file_name);
            if (file.entries.V.empty()) die("ERROR: empty matrix");
            if (cfg.dropout) {
                Coo k; k.n_rows = file.entries.n_rows; k.n_cols = file.entries.n_cols;
                for (size_t i = 0; i < file.entries.V.size(); ++i)
                    if (std::fabs(file.entries.V[i]) >= cfg.dropout_threshold) {
                        k.I.push_back(file.entries.I[i]); k.J.push_back(file.entries.J[i]); k.V.push_back(file.entries.V[i]);
                    }
                file.entries = k;
                if (file.entries.V.empty()) die("ERROR: dropout removed every element");
            }
            ck(uspmv_coo_from_entries(ctx, file.entries.n_rows, file.entries.n_cols, (long)file.entries.V.size(), file.entries.I.data(),
                                      file.entries.J.data(), file.entries.V.data(), file.symmetric ? 1 : 0, &full));
        }
        if (cfg.equilibrate) ck(uspmv_coo_equilibrate(full, nullptr, nullptr));
        long d3[3];
        ck(uspmv_coo_dims(full, d3));
        n_glob = d3[0];
        host.n_rows = d3[0]; host.n_cols = d3[1];
        host.I.resize(d3[2]); host.J.resize(d3[2]); host.V.resize(d3[2]);
        ck(uspmv_coo_export(full, host.I.data(), host.J.data(), host.V.data()));
        have_host = true;
        if (!gen) {
            cfg.matrix_min = *std::min_element(host.V.begin(), host.V.end());
            cfg.matrix_max = *std::max_element(host.V.begin(), host.V.end());
            cfg.matrix_mean = std::accumulate(host.V.begin(), host.V.end(), 0.0) / host.V.size();
        }
        ck(uspmv_seg_work_sharing_arr(seg, d3[0], d3[2], host.I.data(), P, wsa.data()));
        // seg_mtx_struct + localize_row_idx (mpi_funcs.hpp:636-674,862-877): this rank's slab, local rows, global columns
        ck(uspmv_coo_seg_mtx(full, wsa.data(), rank, P, &coo, nullptr));
        uspmv_coo_destroy(full);
        if (rank != 0 || !(cfg.mode == 's' && cfg.validate_result)) { Coo().I.swap(host.I); Coo().J.swap(host.J); Coo().V.swap(host.V); have_host = false; }
    }
    if (rank == 0) {
        printf("work_sharing_arr (%s):", cfg.seg_method.c_str());
        for (int k = 0; k <= P; ++k) printf(" %d", wsa[k]);
        printf("\n");
    }
    long cd[3];
    ck(uspmv_coo_dims(coo, cd));
    const long n_local = cd[0], nnz_local = cd[2];
    if (n_local < 1) die("ERROR: a rank received no rows; use fewer GPUs for this matrix");

    // ---- local format + halo discovery (convert_to_scs -> collect_local_needed_heri -> permute_scs_cols; main.cpp:1271-1308) ----------
    uspmv_scs *scs = nullptr, *part[3] = {nullptr, nullptr, nullptr};
    uspmv_halo *plan = nullptr;
    long dims[8];
    long part_nnz[3] = {0, 0, 0};
    const int first = ap && ap_mode == USPMV_AP_SP_HP ? 1 : 0;
    double t0 = now();
    if (!ap) {
        ck(uspmv_scs_build(ctx, coo, cfg.chunk_size, cfg.sigma, vt, nullptr, &scs));
        ck(uspmv_halo_plan_create(scs, wsa.data(), rank, P, &plan));
        ck(uspmv_scs_permute_cols(scs, nullptr));
        ck(uspmv_scs_dims(scs, dims));
    } else {
        // AP under MPI is refused by the reference (utilities.hpp:1445-1451); here: per-rank partition_precisions, one halo numbering
        // over the parts, x in the original local row order
        uspmv_coo *pc[3] = {nullptr, nullptr, nullptr};
        ck(uspmv_partition_precisions(ctx, coo, ap_mode, cfg.ap_threshold_1, cfg.ap_threshold_2, nullptr, nullptr, &pc[0], &pc[1], &pc[2]));
        const int vts[3] = {USPMV_F64, USPMV_F32, USPMV_F16};
        ck(uspmv_scs_build(ctx, pc[first], cfg.chunk_size, cfg.sigma, vts[first], nullptr, &part[first]));
        ck(uspmv_scs_dims(part[first], dims));
        std::vector<int> perm(dims[2]);
        ck(uspmv_scs_export(part[first], nullptr, nullptr, nullptr, nullptr, perm.data(), nullptr));
        std::vector<uspmv_scs *> used;
        for (int p = 0; p < 3; ++p) {
            if (!pc[p]) continue;
            long d3[3];
            ck(uspmv_coo_dims(pc[p], d3));
            part_nnz[p] = d3[2];
            if (p != first) ck(uspmv_scs_build(ctx, pc[p], cfg.chunk_size, cfg.sigma, vts[p], perm.data(), &part[p]));
            uspmv_coo_destroy(pc[p]);
            used.push_back(part[p]);
        }
        ck(uspmv_halo_plan_create_multi(used.data(), (int)used.size(), wsa.data(), rank, P, 0, &plan));
    }
    uspmv_coo_destroy(coo);
    const double t_convert = now() - t0;
    const long n_pad = dims[4], n_chunks = dims[5];
    long n_elements = dims[6];
    if (ap) { n_elements = 0; for (int p = 0; p < 3; ++p) if (part[p]) { long d[8]; ck(uspmv_scs_dims(part[p], d)); n_elements += d[6]; } }

    // ---- comm schedule (collect_comm_idxs / organize_cumsums, mpi_funcs.hpp:117-232): the need lists are transposed through the
    //      shared segment — send list to peer q = what q needs from this rank ----------------------------------------------------------
    RankInfo &me = sh->info[rank];
    long n_halo = 0;
    ck(uspmv_halo_plan_counts(plan, me.recv_cumsum, &n_halo));
    if (n_halo > sh->cap) die("ERROR: halo larger than the exchange area");
    {
        std::vector<int> flat(std::max<long>(n_halo, 1));
        ck(uspmv_halo_plan_need(plan, flat.data(), me.need_ptr));
        std::copy(flat.begin(), flat.begin() + n_halo, need_area(sh, rank));
    }
    me.n_local = n_local; me.n_halo = n_halo; me.nnz = nnz_local; me.n_elements = n_elements; me.n_chunks = n_chunks; me.n_pad = n_pad;
    for (int p = 0; p < 3; ++p) me.part_nnz[p] = part_nnz[p];
    barrier();
    std::vector<int> send_flat, send_ptr(P + 1, 0);
    for (int q = 0; q < P; ++q) {
        const int *nq = need_area(sh, q);
        const int a = sh->info[q].need_ptr[rank], b = sh->info[q].need_ptr[rank + 1];
        send_flat.insert(send_flat.end(), nq + a, nq + b);
        send_ptr[q + 1] = (int)send_flat.size();
    }
    if (send_flat.empty()) send_flat.push_back(0);
    ck(uspmv_halo_plan_set_send(plan, send_flat.data(), send_ptr.data()));
    if (cfg.print_comm_vol)
        printf("rank %d: receives %ld halo elements, sends %d\n", rank, n_halo, send_ptr[P]);
    long n_int = 0, n_bnd = 0;
    if (!ap) ck(uspmv_scs_split_chunks(scs, &n_int, &n_bnd));

    // ---- vectors + NVLink arena (main.cpp:1405-1431; init/finalize_halo_exchange, classes_structs.hpp:857-995) ---------------------------
    const int bvs = cfg.block_vec_size;
    const int layout = cfg.block_vec_layout == "rowwise" ? USPMV_ROWWISE : USPMV_COLWISE;
    const int xvt = ap ? (ap_mode == USPMV_AP_SP_HP ? USPMV_F32 : USPMV_F64) : vt;
    const long vec_length = n_local + std::max(n_pad - n_local, n_halo);  // n_local + max(scs_padding, halo_count)
    const bool two_buf = cfg.mode == 's' && !ap && bvs == 1 && cfg.n_repetitions > 1;
    uspmv_p2p *p2p = nullptr;
    void *xb[2] = {nullptr, nullptr};
    ck(uspmv_p2p_create_ex(plan, xvt, vec_length, bvs, layout, two_buf ? 2 : 1, &p2p, me.ipc, xb));
    me.vec_length = vec_length;
    me.x_bytes = ((long)vec_length * bvs * (long)vt_bytes(xvt) + 255) / 256 * 256;
    barrier();
    {
        std::vector<unsigned char> handles((size_t)P * 64);
        std::vector<long> pxb(P), pbase(P), pld(P);
        for (int q = 0; q < P; ++q) {
            memcpy(handles.data() + (size_t)q * 64, sh->info[q].ipc, 64);
            pxb[q] = sh->info[q].x_bytes;
            pbase[q] = sh->info[q].n_local + sh->info[q].recv_cumsum[rank];
            pld[q] = sh->info[q].vec_length;
        }
        ck(uspmv_p2p_connect_ex(p2p, handles.data(), pxb.data(), pbase.data(), pld.data()));
    }
    barrier();
    long nnz_total = 0, nel_total = 0, npad_total = 0, pn[3] = {0, 0, 0};
    for (int q = 0; q < P; ++q) {
        nnz_total += sh->info[q].nnz; nel_total += sh->info[q].n_elements; npad_total += sh->info[q].n_pad;
        for (int p = 0; p < 3; ++p) pn[p] += sh->info[q].part_nnz[p];
    }
    const double beta = nel_total ? (double)nnz_total / (double)nel_total : 0.0;
    printf("rank %d: rows [%d, %d), %ld nnz, %ld chunks (%ld interior / %ld boundary), %ld elements, halo %ld; conversion %.3f s\n", rank, wsa[rank],
           wsa[rank + 1], nnz_local, n_chunks, n_int, n_bnd, n_elements, n_halo, t_convert);

    std::vector<double> x_user(vec_length * bvs, 0.0);
    if (cfg.random_init_x == '1') {
        // every rank draws the same default-seeded sequence over ITS vector, like random_init under MPI (utilities.hpp:880-981;
        // bit-identical to the reference's x: tests/test_boundary_cpu.py)
        ck(uspmv_random_init_host(cfg.matrix_min, cfg.matrix_max, vec_length * bvs, USPMV_F64, x_user.data(), n_local, vec_length, bvs, layout));
    } else {
        for (long i = 0; i < vec_length * bvs; ++i) {
            const long r = layout == USPMV_ROWWISE ? i / bvs : i % vec_length;
            x_user[i] = r < n_local ? (cfg.random_init_x == 'm' ? cfg.matrix_mean : 5.0) : 0.0;
        }
    }
    std::vector<int> old_to_new(n_local), new_to_old(n_pad);
    ck(uspmv_scs_export(ap ? part[first] : scs, nullptr, nullptr, nullptr, nullptr, old_to_new.data(), new_to_old.data()));
    std::vector<double> x_perm(vec_length * bvs, 0.0);
    for (long v = 0; v < bvs; ++v)
        for (long i = 0; i < n_local; ++i) {
            const long dst = ap ? i : old_to_new[i];
            if (layout == USPMV_ROWWISE) x_perm[dst * bvs + v] = x_user[i * bvs + v];
            else x_perm[dst + v * vec_length] = x_user[i + v * vec_length];
        }
    upload(ctx, xb[0], x_perm, xvt);
    void *y_d = nullptr, *comm = nullptr;
    ck(uspmv_malloc(ctx, vec_length * bvs * vt_bytes(xvt), &y_d));
    ck(uspmv_memset(ctx, y_d, 0, vec_length * bvs * vt_bytes(xvt), nullptr));
    ck(uspmv_stream_create(ctx, &comm));
    ck(uspmv_ctx_sync(ctx));
    barrier();

    auto execute = [&]() {  // begin/finish halo exchange + kernel (main.cpp:464-468), one call
        if (ap) ck(uspmv_p2p_ap_spmv(p2p, ap_mode, part[0], part[1], part[2], y_d, nullptr, comm));
        else if (bvs > 1) ck(uspmv_p2p_spmmv(p2p, scs, 0, y_d, nullptr, comm));
        else ck(uspmv_p2p_spmv(p2p, scs, y_d, nullptr, comm));
    };
    auto max_runtime = [&](double mine) {
        me.runtime = mine;
        barrier();
        double m = 0.0;
        for (int q = 0; q < P; ++q) m = std::max(m, sh->info[q].runtime);
        barrier();
        return m;
    };
    int rc = 0;
    if (cfg.mode == 'b') {
        const int WARM_UP_REPS = 100;
        double tw = now();
        for (int k = 0; k < WARM_UP_REPS; ++k) execute();
        ck(uspmv_ctx_sync(ctx));
        const double t_warm = max_runtime(now() - tw);
        if (rank == 0) std::cout << "warm up time: " << t_warm << std::endl;
        long n_iter = 2;
        double runtime = 0.0;
        do {
            ck(uspmv_ctx_sync(ctx));
            barrier();  // ba_synch: MPI_Barrier before the timed loop (main.cpp:440-447)
            const double tb = now();
            for (long k = 0; k < n_iter; ++k) execute();
            ck(uspmv_ctx_sync(ctx));
            runtime = max_runtime(now() - tb);  // the slowest rank decides, so every rank takes the same branch
            n_iter *= 2;
        } while (runtime < cfg.bench_time);
        n_iter /= 2;
        int err = 0;
        long ep = 0;
        ck(uspmv_p2p_status(p2p, &err, &ep));
        if (err) { fprintf(stderr, "rank %d: halo exchange timed out (a peer never signalled)\n", rank); rc = 3; }
        if (rank == 0) {
            const double t_kernel = runtime / n_iter;
            const double gflops = (double)nnz_total * 2.0 * bvs / t_kernel / 1e9;
            printf("Total Gflops: %.6f   time per SpM(M)V: %.3f us   revisions: %ld   ranks: %d\n", gflops, t_kernel * 1e6, n_iter, P);
            std::fstream out(cfg.output_filename_bench, std::fstream::in | std::fstream::out | std::fstream::app);
            const int width = 32;
            out << cfg.matrix_file_name << " with " << P << " MPI processes, and " << (npad_total + 255) / 256 << " block(s), and " << 256
                << " thread(s) per block" << std::endl;
            out << "kernel: " << cfg.kernel_format << ", block_vec_size: " << bvs;
            if (cfg.kernel_format == "scs") out << ", C: " << cfg.chunk_size << " sigma: " << cfg.sigma << std::fixed << std::setprecision(8) << ", beta: " << beta;
            out << ", block_vec_layout: " << cfg.block_vec_layout;
            const double pct[3] = {nnz_total ? 100.0 * pn[0] / nnz_total : 0, nnz_total ? 100.0 * pn[1] / nnz_total : 0, nnz_total ? 100.0 * pn[2] / nnz_total : 0};
            out << std::fixed << std::setprecision(2);
            if (cfg.value_type == "ap[dp_sp]") out << ", data_type: ap[dp_sp], threshold: " << cfg.ap_threshold_1 << ", % dp elems: " << pct[0] << ", % sp elems: " << pct[1];
            else if (cfg.value_type == "ap[dp_hp]") out << ", data_type: ap[dp_hp], threshold: " << cfg.ap_threshold_1 << ", % dp elems: " << pct[0] << ", % hp elems: " << pct[2];
            else if (cfg.value_type == "ap[sp_hp]") out << ", data_type: ap[sp_hp], threshold: " << cfg.ap_threshold_1 << ", % sp elems: " << pct[1] << ", % hp elems: " << pct[2];
            else if (cfg.value_type == "ap[dp_sp_hp]") out << ", data_type: ap[dp_sp_hp], threshold 1: " << cfg.ap_threshold_1 << ", threshold 2: " << cfg.ap_threshold_2 << ", % dp elems: " << pct[0] << ", % sp elems: " << pct[1] << ", % hp elems: " << pct[2];
            else out << ", data_type: " << (cfg.value_type == "dp" ? "double" : cfg.value_type == "sp" ? "float" : "half");
            out << ", revisions: " << n_iter << ", seg_method: " << cfg.seg_method << std::endl << std::endl;
            out << std::left << std::setw(width) << "Total Gflops:" << std::left << std::setw(width) << "Total Walltime:" << std::endl;
            out << std::left << std::setw(width) << "-------------" << std::left << std::setw(width) << "-------------" << std::endl;
            out << std::left << std::setprecision(16) << std::left << std::setw(width) << gflops << std::left << std::setw(width) << runtime << std::endl;
            out << std::endl << std::endl;
        }
    } else {
        // ---- solve mode: rev x { exchange ; SpMV ; swap } (main.cpp:528-631), then every rank un-permutes its rows into the
        //      shared result vector and rank 0 validates against the host COO product ----------------------------------------
        if (cfg.n_repetitions > 1 && (ap || bvs > 1)) die("ERROR: multi-GPU solve mode with -rev > 1 supports plain SpMV only");
        const void *result = y_d;
        if (two_buf) {
            int b = 0;
            for (unsigned long it = 0; it < cfg.n_repetitions; ++it, b ^= 1) ck(uspmv_p2p_spmv_buf(p2p, scs, b, b ^ 1, nullptr, nullptr, comm));
            result = xb[b];  // the last y, in permuted row order, rows < n_local
        } else execute();
        ck(uspmv_ctx_sync(ctx));
        int err = 0;
        long ep = 0;
        ck(uspmv_p2p_status(p2p, &err, &ep));
        if (err) { fprintf(stderr, "rank %d: halo exchange timed out\n", rank); rc = 3; }
        std::vector<double> y_perm = download(ctx, result, (two_buf ? n_local : vec_length * bvs), xvt);
        double *yg = y_area(sh, P);  // global y, vector v at yg[v * n_glob + row]
        for (long v = 0; v < bvs; ++v)
            for (long i = 0; i < n_local; ++i)
                // (an empty row that tied with the padding rows of the last sigma-window may sit at a position >= n_local, which the
                //  two-buffer loop does not store: its y is 0 by definition)
                yg[v * n_glob + wsa[rank] + i] = two_buf ? (old_to_new[i] < n_local ? y_perm[old_to_new[i]] : 0.0)
                                                 : layout == USPMV_ROWWISE ? y_perm[(long)old_to_new[i] * bvs + v] : y_perm[old_to_new[i] + v * vec_length];
        // the x every rank used (identical local sequences), for the validator
        double *xg = yg + (size_t)bvs * n_glob;
        for (long v = 0; v < bvs; ++v)
            for (long i = 0; i < n_local; ++i)
                xg[v * n_glob + wsa[rank] + i] = layout == USPMV_ROWWISE ? x_user[i * bvs + v] : x_user[i + v * vec_length];
        barrier();
        if (rank == 0) {
            printf("solve mode: %lu revision(s) on %d GPUs; y[0..3] =", cfg.n_repetitions, P);
            for (long i = 0; i < std::min<long>(4, n_glob); ++i) printf(" %.10g", yg[i]);
            printf("\n");
            if (cfg.validate_result && have_host) {
                const bool sp_x = ap && ap_mode == USPMV_AP_SP_HP;
                std::vector<double> ref((size_t)bvs * n_glob), got(yg, yg + (size_t)bvs * n_glob);
                for (long v = 0; v < bvs; ++v) {
                    std::vector<double> xa(xg + v * n_glob, xg + (v + 1) * n_glob), ya(n_glob);
                    if (sp_x) for (double &t : xa) t = (double)(float)t;
                    for (unsigned long it = 0; it < cfg.n_repetitions; ++it) {
                        std::fill(ya.begin(), ya.end(), 0.0);
                        for (size_t k = 0; k < host.V.size(); ++k) ya[host.I[k]] += host.V[k] * xa[host.J[k]];
                        if (it + 1 < cfg.n_repetitions) std::swap(xa, ya);
                    }
                    std::copy(ya.begin(), ya.end(), ref.begin() + v * n_glob);
                }
                ResultReport rr;
                rr.matrix_file_name = cfg.matrix_file_name; rr.kernel_format = cfg.kernel_format; rr.value_type = cfg.value_type;
                rr.block_vec_layout = cfg.block_vec_layout; rr.seg_method = cfg.seg_method; rr.chunk_size = cfg.chunk_size; rr.sigma = cfg.sigma;
                rr.n_blocks = (npad_total + 255) / 256; rr.ranks = P; rr.verbose = cfg.verbose; rr.revisions = cfg.n_repetitions; rr.beta = beta;
                const double max_rel = write_result_to_file(rr, ref, got, n_glob);  // spmv_mkl_compare_<type>.txt, write_results.hpp:160-440
                const char *verdict = (max_rel > 1e-2 || std::isnan(max_rel)) ? "ERROR" : max_rel > 1e-4 ? "WARNING" : "OK";
                printf("validation vs host COO product: max relative difference %.3e -> %s\n", max_rel, verdict);
                if ((max_rel > 1e-2 || std::isnan(max_rel)) && cfg.value_type == "dp") rc = 2;
            }
        }
    }
    fflush(stdout);
    // collective teardown of the IPC arenas: every rank's device is idle (sync) before anybody unmaps a neighbour's arena, and
    // every mapping is closed before anybody frees
    if (uspmv_p2p_sync(p2p)) { fprintf(stderr, "rank %d: %s\n", rank, uspmv_last_error()); rc = 3; }
    barrier();
    uspmv_p2p_disconnect(p2p);
    barrier();
    uspmv_stream_destroy(ctx, comm);
    uspmv_free(ctx, y_d);
    uspmv_p2p_destroy(p2p);
    uspmv_halo_destroy(plan);
    if (scs) uspmv_scs_destroy(scs);
    for (auto *p : part) if (p) uspmv_scs_destroy(p);
    uspmv_ctx_destroy(ctx);
    return rc;
}

int run_multi(const Config &cfg) {
    const int P = cfg.n_gpus;
    const long n_glob = peek_rows(cfg);
    if (n_glob < P) die("ERROR: fewer rows than GPUs");
    const int bvs = cfg.block_vec_size;
    // shared segment: control block | need lists (P x n_glob ints, sparse) | y and x for the validator (2 x bvs x n_glob doubles)
    const size_t bytes = 65536 + (size_t)P * n_glob * sizeof(int) + 2 * (size_t)bvs * n_glob * sizeof(double) + 4096;
    void *mem = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    if (mem == MAP_FAILED) die("ERROR: mmap of the rank exchange segment failed");
    static_assert(sizeof(Shared) <= 65536, "control block");
    Shared *sh = new (mem) Shared();
    sh->cap = n_glob;
    sh->y_len = (long)bvs * n_glob;
    pthread_barrierattr_t at;
    pthread_barrierattr_init(&at);
    pthread_barrierattr_setpshared(&at, PTHREAD_PROCESS_SHARED);
    pthread_barrier_init(&sh->bar, &at, P);
    fflush(stdout);
    fflush(stderr);
    std::vector<pid_t> kids(P, 0);
    for (int r = 0; r < P; ++r) {
        pid_t pid = fork();
        if (pid < 0) die("ERROR: fork failed");
        if (pid == 0) {
            int rc = rank_main(cfg, r, P, sh);
            fflush(stdout);
            _exit(rc);
        }
        kids[r] = pid;
    }
    // a rank that dies would leave the others at a barrier forever: the first failure takes the rest down (exact PIDs)
    int worst = 0, left = P;
    while (left > 0) {
        int st = 0;
        pid_t pid = wait(&st);
        if (pid < 0) break;
        --left;
        const int rc = WIFEXITED(st) ? WEXITSTATUS(st) : 128 + (WIFSIGNALED(st) ? WTERMSIG(st) : 0);
        for (pid_t &k : kids) if (k == pid) k = 0;
        if (rc != 0) {
            worst = std::max(worst, rc);
            if (rc != 2)  // 2 = validation verdict ERROR on rank 0, the others finish normally
                for (pid_t k : kids) if (k > 0) kill(k, SIGKILL);
        }
    }
    munmap(mem, bytes);
    return worst;
}

}  // namespace

int main(int argc, char **argv) {
    Config cfg;
    parse_cli(argc, argv, cfg);
    if (cfg.n_gpus > 1) return run_multi(cfg);
    uspmv_ctx *ctx = nullptr;
    ck(uspmv_ctx_create(0, &ctx));

    // ---- matrix ------------------------------------------------------------------------------------------------------
    uspmv_coo *coo = nullptr;
    Coo host;  // kept only when a file was read (needed for validation / x statistics)
    bool have_host = false;
    if (cfg.matrix_file_name.rfind("gen:", 0) == 0) {
        char kind[32];
        long n = 0;
        if (sscanf(cfg.matrix_file_name.c_str(), "gen:%31[^:]:%ld", kind, &n) != 2 || n < 1) die("ERROR: generator syntax is gen:laplace7:<n> or gen:stencil27:<n>");
        int pts = !strcmp(kind, "laplace7") ? 7 : !strcmp(kind, "stencil27") ? 27 : 0;
        if (!pts) die("ERROR: unknown generator");
        ck(uspmv_coo_stencil(ctx, pts, n, n, n, 0, n * n * n, &coo));
        cfg.matrix_min = -1.0; cfg.matrix_max = pts - 1.0; cfg.matrix_mean = 0.0;
    } else {
        MtxFile file = read_mtx(cfg.matrix_file_name);
        if (file.entries.V.empty()) die("ERROR: empty matrix");
        if (cfg.dropout) {  // -dropout: remove elements below the threshold (utilities.hpp, -dt); (i,j) and (j,i) share the value
            Coo k; k.n_rows = file.entries.n_rows; k.n_cols = file.entries.n_cols;
            for (size_t i = 0; i < file.entries.V.size(); ++i)
                if (std::fabs(file.entries.V[i]) >= cfg.dropout_threshold) {
                    k.I.push_back(file.entries.I[i]); k.J.push_back(file.entries.J[i]); k.V.push_back(file.entries.V[i]);
                }
            file.entries = k;
            if (file.entries.V.empty()) die("ERROR: dropout removed every element");
        }
        ck(uspmv_coo_from_entries(ctx, file.entries.n_rows, file.entries.n_cols, (long)file.entries.V.size(), file.entries.I.data(),
                                  file.entries.J.data(), file.entries.V.data(), file.symmetric ? 1 : 0, &coo));
        long d3[3];
        ck(uspmv_coo_dims(coo, d3));
        host.n_rows = d3[0]; host.n_cols = d3[1];
        host.I.resize(d3[2]); host.J.resize(d3[2]); host.V.resize(d3[2]);
        ck(uspmv_coo_export(coo, host.I.data(), host.J.data(), host.V.data()));
        have_host = true;
        cfg.matrix_min = *std::min_element(host.V.begin(), host.V.end());
        cfg.matrix_max = *std::max_element(host.V.begin(), host.V.end());
        cfg.matrix_mean = std::accumulate(host.V.begin(), host.V.end(), 0.0) / host.V.size();
    }
    // -equilibrate (main.cpp:1118-1153): scale rows, then columns; AP also hands the maxima to partition_precisions
    std::vector<double> rowmax, colmax;
    if (cfg.equilibrate) {
        long d3[3];
        ck(uspmv_coo_dims(coo, d3));
        rowmax.resize(d3[0]); colmax.resize(d3[1]);
        ck(uspmv_coo_equilibrate(coo, rowmax.data(), colmax.data()));
        if (have_host) ck(uspmv_coo_export(coo, nullptr, nullptr, host.V.data()));  // validation uses the scaled matrix (main.cpp:1753)
    }
    long cd[3];
    ck(uspmv_coo_dims(coo, cd));
    const long n_rows = cd[0], nnz = cd[2];

    // ---- format conversion (convert_to_scs, partition_precisions, permute_scs_cols; main.cpp:1128-1221,1308) ---------
    const bool ap = is_ap(cfg.value_type);
    const int ap_mode = cfg.value_type == "ap[dp_sp]" ? USPMV_AP_DP_SP : cfg.value_type == "ap[dp_hp]" ? USPMV_AP_DP_HP
                        : cfg.value_type == "ap[sp_hp]" ? USPMV_AP_SP_HP : USPMV_AP_DP_SP_HP;
    const int vt = cfg.value_type == "sp" ? USPMV_F32 : cfg.value_type == "hp" ? USPMV_F16 : USPMV_F64;
    uspmv_scs *scs = nullptr, *part[3] = {nullptr, nullptr, nullptr};
    long dims[8];
    double t0 = now();
    long part_nnz[3] = {0, 0, 0};
    if (!ap) {
        ck(uspmv_scs_build(ctx, coo, cfg.chunk_size, cfg.sigma, vt, nullptr, &scs));
        ck(uspmv_scs_permute_cols(scs, nullptr));
        ck(uspmv_scs_dims(scs, dims));
    } else {
        uspmv_coo *pc[3] = {nullptr, nullptr, nullptr};
        ck(uspmv_partition_precisions(ctx, coo, ap_mode, cfg.ap_threshold_1, cfg.ap_threshold_2, cfg.equilibrate ? rowmax.data() : nullptr,
                                      cfg.equilibrate ? colmax.data() : nullptr, &pc[0], &pc[1], &pc[2]));
        const int first = ap_mode == USPMV_AP_SP_HP ? 1 : 0;
        const int vts[3] = {USPMV_F64, USPMV_F32, USPMV_F16};
        ck(uspmv_scs_build(ctx, pc[first], cfg.chunk_size, cfg.sigma, vts[first], nullptr, &part[first]));
        ck(uspmv_scs_dims(part[first], dims));
        std::vector<int> perm(dims[2]);
        ck(uspmv_scs_export(part[first], nullptr, nullptr, nullptr, nullptr, perm.data(), nullptr));
        for (int p = 0; p < 3; ++p) {
            if (!pc[p]) continue;
            long d3[3];
            ck(uspmv_coo_dims(pc[p], d3));
            part_nnz[p] = d3[2];
            if (p != first) ck(uspmv_scs_build(ctx, pc[p], cfg.chunk_size, cfg.sigma, vts[p], perm.data(), &part[p]));  // fixed_permutation, main.cpp:1175-1219
            uspmv_coo_destroy(pc[p]);
        }
        // the AP structs keep ORIGINAL column numbering (they are never column-permuted in the reference either,
        // main.cpp:1308-1332); x is therefore used un-permuted and only y comes out in permuted row order.
    }
    const double t_convert = now() - t0;
    const long n_pad = dims[4], n_chunks = dims[5];
    long n_elements = dims[6];
    if (ap) { n_elements = 0; for (int p = 0; p < 3; ++p) if (part[p]) { long d[8]; ck(uspmv_scs_dims(part[p], d)); n_elements += d[6]; } }
    const double beta = n_elements ? (double)nnz / (double)n_elements : 0.0;
    printf("matrix: %ld rows, %ld nnz; C=%ld sigma=%ld: %ld chunks, %ld elements, beta=%.8f; conversion on device took %.3f s\n", n_rows, nnz,
           cfg.chunk_size, cfg.sigma, n_chunks, n_elements, beta, t_convert);

    // ---- vectors (main.cpp:1405-1431; utilities.hpp:880-981) -------------------------------------------------------------
    const int bvs = cfg.block_vec_size;
    const int layout = cfg.block_vec_layout == "rowwise" ? USPMV_ROWWISE : USPMV_COLWISE;
    const long vec_length = std::max(n_pad, n_rows);  // n_local + per_vector_padding (no halo in a single process)
    const int xvt = ap ? (ap_mode == USPMV_AP_SP_HP ? USPMV_F32 : USPMV_F64) : vt;
    std::vector<double> x_user(vec_length * bvs, 0.0);
    if (cfg.random_init_x == '1') {
        // random_init + the padding rule (utilities.hpp:880-981), bit-identical to the reference's x (tests/test_boundary_cpu.py)
        ck(uspmv_random_init_host(cfg.matrix_min, cfg.matrix_max, vec_length * bvs, USPMV_F64, x_user.data(), n_rows, vec_length, bvs, layout));
    } else {
        for (long i = 0; i < vec_length * bvs; ++i) {
            const long r = layout == USPMV_ROWWISE ? i / bvs : i % vec_length;
            x_user[i] = r < n_rows ? (cfg.random_init_x == 'm' ? cfg.matrix_mean : 5.0) : 0.0;  // DefaultValues x = 5.0; padding zeroed
        }
    }
    std::vector<int> old_to_new(n_rows), new_to_old(n_pad);
    ck(uspmv_scs_export(ap ? part[ap_mode == USPMV_AP_SP_HP ? 1 : 0] : scs, nullptr, nullptr, nullptr, nullptr, old_to_new.data(), new_to_old.data()));
    // x_perm[i] = x[new_to_old[i]] (main.cpp:86-102); AP: x stays in user order (see above)
    std::vector<double> x_perm(vec_length * bvs, 0.0);
    for (long v = 0; v < bvs; ++v)
        for (long i = 0; i < n_rows; ++i) {
            const long dst = ap ? i : old_to_new[i];
            if (layout == USPMV_ROWWISE) x_perm[dst * bvs + v] = x_user[i * bvs + v];
            else x_perm[dst + v * vec_length] = x_user[i + v * vec_length];
        }
    void *x_d = nullptr, *y_d = nullptr;
    ck(uspmv_malloc(ctx, vec_length * bvs * vt_bytes(xvt), &x_d));
    ck(uspmv_malloc(ctx, vec_length * bvs * vt_bytes(xvt), &y_d));
    ck(uspmv_memset(ctx, y_d, 0, vec_length * bvs * vt_bytes(xvt), nullptr));
    upload(ctx, x_d, x_perm, xvt);

    auto execute = [&]() {  // SpmvKernel::execute (classes_structs.hpp:997-1127)
        if (ap) ck(uspmv_ap_spmv(ap_mode, part[0], part[1], part[2], x_d, y_d, nullptr));
        else if (bvs > 1) ck(uspmv_spmmv(scs, x_d, y_d, bvs, vec_length, layout, nullptr));
        else ck(uspmv_spmv(scs, x_d, y_d, nullptr));
    };

    if (cfg.mode == 'b') {
        const int WARM_UP_REPS = 100;  // main.cpp:22
        ck(uspmv_ctx_sync(ctx));
        double tw = now();
        for (int k = 0; k < WARM_UP_REPS; ++k) execute();
        ck(uspmv_ctx_sync(ctx));
        std::cout << "warm up time: " << now() - tw << std::endl;
        long n_iter = 2;
        double runtime = 0.0;
        do {
            ck(uspmv_ctx_sync(ctx));
            const double tb = now();
            for (long k = 0; k < n_iter; ++k) execute();
            ck(uspmv_ctx_sync(ctx));
            runtime = now() - tb;
            n_iter *= 2;
        } while (runtime < cfg.bench_time);
        n_iter /= 2;
        const double t_kernel = runtime / n_iter;
        const double gflops = (double)nnz * 2.0 * bvs / t_kernel / 1e9;  // "only count useful flops", main.cpp:521-526
        const size_t vs = vt_bytes(vt);
        const double bytes = ap ? 0.0 : (double)n_elements * (vs + 4) + 8.0 * n_chunks + (double)bvs * vs * (n_rows + n_pad);
        printf("Total Gflops: %.6f   time per SpM(M)V: %.3f us   revisions: %ld", gflops, t_kernel * 1e6, n_iter);
        if (!ap) printf("   model bandwidth: %.1f GB/s", bytes / t_kernel / 1e9);
        printf("\n");
        // ---- spmv_bench.txt, same layout as write_bench_to_file (write_results.hpp:43-157, nvcc branch) ----
        std::fstream out(cfg.output_filename_bench, std::fstream::in | std::fstream::out | std::fstream::app);
        const int width = 32;
        out << cfg.matrix_file_name << " with " << (n_pad + 255) / 256 << " block(s), and " << 256 << " thread(s) per block" << std::endl;
        out << "kernel: " << cfg.kernel_format << ", block_vec_size: " << bvs;
        if (cfg.kernel_format == "scs") out << ", C: " << cfg.chunk_size << " sigma: " << cfg.sigma << std::fixed << std::setprecision(8) << ", beta: " << beta;
        out << ", block_vec_layout: " << cfg.block_vec_layout;
        const double pct[3] = {nnz ? 100.0 * part_nnz[0] / nnz : 0, nnz ? 100.0 * part_nnz[1] / nnz : 0, nnz ? 100.0 * part_nnz[2] / nnz : 0};
        out << std::fixed << std::setprecision(2);
        if (cfg.value_type == "ap[dp_sp]") out << ", data_type: ap[dp_sp], threshold: " << cfg.ap_threshold_1 << ", % dp elems: " << pct[0] << ", % sp elems: " << pct[1];
        else if (cfg.value_type == "ap[dp_hp]") out << ", data_type: ap[dp_hp], threshold: " << cfg.ap_threshold_1 << ", % dp elems: " << pct[0] << ", % hp elems: " << pct[2];
        else if (cfg.value_type == "ap[sp_hp]") out << ", data_type: ap[sp_hp], threshold: " << cfg.ap_threshold_1 << ", % sp elems: " << pct[1] << ", % hp elems: " << pct[2];
        else if (cfg.value_type == "ap[dp_sp_hp]") out << ", data_type: ap[dp_sp_hp], threshold 1: " << cfg.ap_threshold_1 << ", threshold 2: " << cfg.ap_threshold_2 << ", % dp elems: " << pct[0] << ", % sp elems: " << pct[1] << ", % hp elems: " << pct[2];
        else out << ", data_type: " << (cfg.value_type == "dp" ? "double" : cfg.value_type == "sp" ? "float" : "half");
        out << ", revisions: " << n_iter << std::endl << std::endl;
        out << std::left << std::setw(width) << "Total Gflops:" << std::left << std::setw(width) << "Total Walltime:" << std::endl;
        out << std::left << std::setw(width) << "-------------" << std::left << std::setw(width) << "-------------" << std::endl;
        out << std::left << std::setprecision(16) << std::left << std::setw(width) << gflops << std::left << std::setw(width) << runtime << std::endl;
        out << std::endl << std::endl;
    } else {
        // ---- solve mode: rev x { SpMV; swap } then un-permute and validate -------------------------------------------
        if (ap && cfg.n_repetitions > 1) die("ERROR: solve mode with adaptive precision supports -rev 1 only (x stays in user order)");
        for (unsigned long it = 0; it < cfg.n_repetitions; ++it) {
            execute();
            if (it + 1 < cfg.n_repetitions) std::swap(x_d, y_d);  // swap_local_vectors, classes_structs.hpp:1130-1165
        }
        ck(uspmv_ctx_sync(ctx));
        std::vector<double> y_perm = download(ctx, y_d, vec_length * bvs, xvt);
        std::vector<double> y(n_rows * bvs);  // sorted_y[i] = y_perm[old_to_new[i]] (utilities.hpp:3856-3868)
        for (long v = 0; v < bvs; ++v)
            for (long i = 0; i < n_rows; ++i)
                y[i + v * n_rows] = layout == USPMV_ROWWISE ? y_perm[(long)old_to_new[i] * bvs + v] : y_perm[old_to_new[i] + v * vec_length];
        printf("solve mode: %lu revision(s); y[0..3] =", cfg.n_repetitions);
        for (long i = 0; i < std::min<long>(4, n_rows); ++i) printf(" %.10g", y[i]);
        printf("\n");
        if (cfg.validate_result && have_host) {
            // host-side COO reference in double (stands in for the reference's MKL validator, write_results.hpp:442-556)
            std::vector<double> ref((size_t)bvs * n_rows);
            for (long v = 0; v < bvs; ++v) {
                std::vector<double> xa(n_rows), ya(n_rows);
                for (long i = 0; i < n_rows; ++i) xa[i] = layout == USPMV_ROWWISE ? x_user[i * bvs + v] : x_user[i + v * vec_length];
                for (unsigned long it = 0; it < cfg.n_repetitions; ++it) {
                    std::fill(ya.begin(), ya.end(), 0.0);
                    for (size_t k = 0; k < host.V.size(); ++k) ya[host.I[k]] += host.V[k] * xa[host.J[k]];
                    if (it + 1 < cfg.n_repetitions) std::swap(xa, ya);
                }
                std::copy(ya.begin(), ya.end(), ref.begin() + v * n_rows);
            }
            ResultReport rr;
            rr.matrix_file_name = cfg.matrix_file_name; rr.kernel_format = cfg.kernel_format; rr.value_type = cfg.value_type;
            rr.block_vec_layout = cfg.block_vec_layout; rr.seg_method = cfg.seg_method; rr.chunk_size = cfg.chunk_size; rr.sigma = cfg.sigma;
            rr.n_blocks = (n_pad + 255) / 256; rr.ranks = 1; rr.verbose = cfg.verbose; rr.revisions = cfg.n_repetitions; rr.beta = beta;
            // same file and table as write_result_to_file (write_results.hpp:160-440); thresholds :378-383,422-428
            const double max_rel = write_result_to_file(rr, ref, y, n_rows);
            const char *verdict = (max_rel > 1e-2 || std::isnan(max_rel)) ? "ERROR" : max_rel > 1e-4 ? "WARNING" : "OK";
            printf("validation vs host COO product: max relative difference %.3e -> %s\n", max_rel, verdict);
            if ((max_rel > 1e-2 || std::isnan(max_rel)) && cfg.value_type == "dp") return 2;
        }
    }
    uspmv_free(ctx, x_d);
    uspmv_free(ctx, y_d);
    if (scs) uspmv_scs_destroy(scs);
    for (auto *p : part) if (p) uspmv_scs_destroy(p);
    uspmv_coo_destroy(coo);
    uspmv_ctx_destroy(ctx);
    return 0;
}
