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
