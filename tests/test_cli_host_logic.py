"""CPU: the `uspmv` harness clone's host logic that runs before any CUDA call — the reference's argument checks
(utilities.hpp:1371-1545) — and its refusal to run without a GPU (no CPU fallback)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "ultimate-spmv_b200", "bin", "uspmv")


def run(args, cwd):
    if not os.path.exists(BIN):  # built by __graft_entry__.build(); cheap to make when only the host binaries are missing
        subprocess.run(["make", "-C", os.path.join(ROOT, "ultimate-spmv_b200")], check=True, capture_output=True, timeout=1800)
    return subprocess.run([BIN] + args, cwd=cwd, capture_output=True, text=True, timeout=60)


@pytest.fixture(scope="module")
def mtx(tmp_path_factory):
    p = tmp_path_factory.mktemp("cli") / "m.mtx"
    p.write_text("%%MatrixMarket matrix coordinate real general\n3 3 4\n1 1 2.0\n2 2 3.0\n3 1 -1.0\n3 3 4.0\n")
    return str(p)


@pytest.mark.parametrize("args,msg", [
    (["scs", "-c", "0"], "chunk size must be >= 1"),
    (["scs", "-s", "0"], "sigma must be >= 1"),
    (["scs", "-block_vec_size", "2", "-ap[dp_sp]"], "SpMMV is not yet implemented for AP kernels"),
    (["scs", "-block_vec_layout", "rowwise"], "Row-wise block vector layout selected, but block vector width is 1"),
    (["ell"], "kernel format not recognized"),
    (["scs", "-ap[dp_sp_hp]", "-apt1", "1", "-apt2", "2"], "second threshold is larger than the first"),
    (["scs", "-seg_metis"], "USE_METIS not defined"),
    (["scs", "-mode", "x"], "Only bench (b) and solve (s) modes are supported"),
    (["scs", "-gpus", "0"], "-gpus must be in [1,16]"),
    (["scs", "-gpus", "2", "-comm_halos", "0"], "always exchanges the halo"),
    (["scs", "-bogus"], "unknown argument"),
])
def test_rejections_before_cuda(mtx, tmp_path, args, msg):
    r = run([mtx] + args, str(tmp_path))
    assert r.returncode != 0 and msg in r.stderr, (args, r.stderr[-400:])


def test_usage_lists_the_reference_flags(tmp_path):
    r = run([], str(tmp_path))
    assert r.returncode != 0
    for flag in ("-block_vec_size", "-c ", "-s ", "-rev", "-rand_x", "-seg_metis / seg_nnz / seg_rows", "-validate", "-mode", "-bench_time",
                 "-equilibrate", "-ap_threshold_1", "-ap_threshold_2", "-dropout_threshold", "-gpus"):
        assert flag in r.stderr, flag


def test_no_cpu_fallback(mtx, tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = run([mtx, "scs", "-c", "32", "-s", "1", "-mode", "s"], str(tmp_path))
    assert r.returncode != 0 and "no CPU fallback" in r.stderr, r.stderr[-400:]
    r = run([mtx, "scs", "-c", "32", "-s", "1", "-mode", "s", "-gpus", "2"], str(tmp_path))
    assert r.returncode != 0 and "no CPU fallback" in r.stderr, r.stderr[-400:]


def test_solve_mode_report_file_layout(tmp_path):
    """host/result_report.hpp writes spmv_mkl_compare_<type>.txt with the reference's layout (write_results.hpp:160-440): run
    description, table header, one summary row (verbose 0) or one row per element (verbose 1), ERROR / WARNING marks."""
    import re
    src = tmp_path / "t.cpp"
    src.write_text('''
#include "result_report.hpp"
#include <cstdio>
int main() {
    ResultReport c;
    c.matrix_file_name = "m.mtx"; c.kernel_format = "scs"; c.value_type = "dp"; c.block_vec_layout = "colwise"; c.seg_method = "seg-nnz";
    c.chunk_size = 32; c.sigma = 512; c.n_blocks = 8; c.revisions = 2; c.beta = 0.94599197;
    std::vector<double> ref = {1.0, 2.0, 0.0, 4.0, 1.0, 2.0, 3.0, 4.0}, got = {1.0, 2.0, 0.0, 4.0, 1.0, 2.0 * (1 + 1e-3), 3.0, 4.0};
    double m0 = write_result_to_file(c, ref, got, 4);
    c.verbose = 1; c.ranks = 2;
    double m1 = write_result_to_file(c, ref, got, 4);
    c.value_type = "ap[dp_sp]"; c.verbose = 0; got[7] = 5.0;
    double m2 = write_result_to_file(c, ref, got, 4);
    std::printf("%.6e %.6e %.6e\\n", m0, m1, m2);
    return 0;
}
''')
    exe = tmp_path / "t"
    inc = os.path.join(ROOT, "ultimate-spmv_b200", "host")
    subprocess.run(["g++", "-std=c++17", "-O1", "-I", inc, "-o", str(exe), str(src)], check=True, capture_output=True, timeout=120)
    r = subprocess.run([str(exe)], cwd=str(tmp_path), capture_output=True, text=True, timeout=30)
    assert r.returncode == 0, r.stderr
    m0, m1, m2 = (float(v) for v in r.stdout.split())
    assert abs(m0 - 1e-3) < 1e-9 and m0 == m1 and abs(m2 - 0.25) < 1e-12
    dp = (tmp_path / "spmv_mkl_compare_dp.txt").read_text().splitlines()
    assert dp[0] == "m.mtx with 8 block(s), and 256 thread(s) per block"
    assert dp[1] == "kernel: scs, C: 32, sigma: 512, beta: 0.94599197, block_vec_layout: colwise, data_type: dp, revisions: 2"
    assert dp[2].startswith("mkl rel. elem:") and "MAX rel. diff(%):" in dp[2] and "||mkl - uspmv||/||mkl||_2" in dp[2]
    assert dp[3].startswith("-------------")
    assert re.match(r"2\.0+e\+00\s+2\.002\d*e\+00\s+1\.0\d*e-01\s", dp[4]) and dp[4].rstrip().endswith("WARNING")
    # second block: two ranks, verbose: one row per element with vector / row index
    i = dp.index("m.mtx with 2 MPI processes, and 8 block(s), and 256 thread(s) per block")
    assert dp[i + 1].endswith("revisions: 2, seg_method: seg-nnz, MPI_mode: bulkvec")
    assert dp[i + 2].startswith("vec idx:") and "uspmv results:" in dp[i + 2]
    rows = dp[i + 4:i + 12]
    assert len(rows) == 8 and rows[5].split()[:2] == ["1", "1"] and rows[5].rstrip().endswith("WARNING") and not rows[0].rstrip().endswith("WARNING")
    ap = (tmp_path / "spmv_mkl_compare_ap.txt").read_text()
    assert "data_type: ap[dp_sp]" in ap and ap.rstrip().endswith("ERROR")


def test_bench_report_block_is_scrapable_like_the_reference():
    """tests/golden/spmv_bench_sample.txt is a block this harness appended to spmv_bench.txt on two B200s (`-gpus 2`).  The reference's
    scripts/scrape_perf.py finds the line holding "Total Gflops:" and reads the first token two lines below it; its nvcc / MPI builds
    start a block with "<matrix> with N MPI processes, and B block(s), and T thread(s) per block" (write_results.hpp:66-75)."""
    import re
    lines = open(os.path.join(ROOT, "tests", "golden", "spmv_bench_sample.txt")).read().splitlines(keepends=True)
    vals = [float(lines[k + 2].split()[0]) for k, ln in enumerate(lines) if "Total Gflops:" in ln]
    assert vals == [1683.3778070149201085]
    assert re.match(r"\S+ with 2 MPI processes, and \d+ block\(s\), and 256 thread\(s\) per block$", lines[0].rstrip("\n"))
    assert re.match(r"kernel: scs, block_vec_size: 1, C: 32 sigma: 1, beta: 0\.\d{8}, block_vec_layout: colwise, data_type: double, revisions: \d+", lines[1])
