"""CPU checks of the drop-in boundary: the shared library loads without a GPU and exports every symbol that
include/uspmv_b200.h declares; compute entry points fail loudly (no CPU fallback) when no device is present."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "uspmv_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(uspmv_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_boundary():
    syms = declared_symbols()
    for must in ("uspmv_scs_build", "uspmv_scs_export", "uspmv_scs_permute_cols", "uspmv_spmv", "uspmv_spmmv", "uspmv_ap_spmv",
                 "uspmv_partition_precisions", "uspmv_seg_work_sharing_arr", "uspmv_halo_plan_create", "uspmv_halo_pack",
                 "uspmv_scs_gpu", "uspmv_csr_gpu", "uspmv_apply_permutation", "uspmv_spmv_host"):
        assert must in syms
    assert len(syms) >= 40


def test_library_exports_every_declared_symbol(pkg):
    lib = pkg.capi.lib
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, f"declared in include/uspmv_b200.h but not exported: {missing}"


def test_every_header_entry_cites_the_reference():
    src = open(HEADER).read()
    assert src.count(".hpp:") + src.count(".cpp:") + src.count(".mk:") >= 30, "entry points must cite the reference file:line they replace"


def test_no_cpu_fallback(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = ctypes.c_void_p()
    rc = pkg.capi.lib.uspmv_ctx_create(0, ctypes.byref(h))
    assert rc != 0
    msg = pkg.capi.lib.uspmv_last_error().decode()
    assert "no CPU fallback" in msg or "CUDA" in msg
    with pytest.raises(pkg.capi.UspmvError):
        pkg.engine.Context(0)


def test_product_does_not_touch_the_oracle():
    """Nothing under ultimate-spmv_b200/ or include/ may import, link or load oracle/."""
    bad = []
    for base in ("ultimate-spmv_b200", "include"):
        for dp, _, files in os.walk(os.path.join(ROOT, base)):
            if os.sep + "build" in dp or os.sep + "lib" in dp or "__pycache__" in dp:
                continue
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp", "Makefile")):
                    txt = open(os.path.join(dp, f), errors="replace").read()
                    if re.search(r"liboracle|oracle\.bindings|from oracle|import oracle|oracle/_ref|uspmv_ref_", txt):
                        bad.append(os.path.join(dp, f))
    assert not bad, bad


def test_baseline_drivers_load_without_a_gpu_and_stay_out_of_the_product():
    """The two GPU baselines (the reference's own CUDA kernels recompiled for sm_100, and cuSPARSE the way the reference's
    USE_CUSPARSE mode calls it) are measurement infrastructure under oracle/_ref: they must load here (symbols only, no CUDA call) so that
    bench.py finds them on the GPU box, and the product library must not depend on them or on cuSPARSE."""
    import subprocess
    from oracle import bindings
    if not bindings.cusparse_available():
        pytest.skip("oracle/_ref/libuspmv_cusparse.so not built")
    lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libuspmv_cusparse.so"))
    assert hasattr(lib, "cusp_spmv") and hasattr(lib, "cusp_last_error")
    if bindings.ref_gpu_available():
        ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libuspmv_ref_gpu.so"))
    needed = subprocess.run(["ldd", os.path.join(ROOT, "ultimate-spmv_b200", "lib", "libuspmv_b200.so")], capture_output=True, text=True).stdout
    assert "cusparse" not in needed and "uspmv_ref" not in needed and "oracle" not in needed, needed
