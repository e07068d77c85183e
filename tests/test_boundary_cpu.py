"""CPU: pieces of the drop-in boundary that are host logic — the x initialisation (random_init + padding rule, utilities.hpp:880-981)
against fixtures produced by the unmodified reference and against the reference itself, and the compile-time fact that the harness
adapter's launchers are assignable to the reference's OnePrecFuncPtr / MultiPrecFuncPtr typedefs (classes_structs.hpp:283-333)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, ROOT

NPT = {"dp": np.float64, "sp": np.float32}
CODE = {"dp": 0, "sp": 1}


def _product_x(capi, vt, lo, hi, n_rows, n_pad, bvs, layout):
    out = np.zeros(n_pad * bvs, NPT[vt])
    capi.call("uspmv_random_init_host", float(lo), float(hi), n_pad * bvs, CODE[vt], out.ctypes.data_as(C.c_void_p), int(n_rows), int(n_pad),
              int(bvs), 1 if layout == "row" else 0)
    return out


@pytest.mark.parametrize("layout", ["col", "row"])
@pytest.mark.parametrize("vt", ["dp", "sp"])
def test_x_init_equals_reference_fixture(pkg, layout, vt):
    z = np.load(os.path.join(GOLDEN, "ref_xinit.npz"))
    for k, (lo, hi, n_rows, n_pad, bvs) in enumerate(z["cases"]):
        want = z[f"{layout}|{vt}|{k}"]
        got = _product_x(pkg.capi, vt, lo, hi, int(n_rows), int(n_pad), int(bvs), layout)
        assert np.array_equal(got.view(np.uint8), want.view(np.uint8)), (layout, vt, k)
        # the padding rule really zeroes something, and the random part is inside [min, max)
        nz = want[want != 0]
        assert (want == 0).sum() >= (int(n_pad) - int(n_rows)) and nz.min() >= min(lo, 0) and nz.max() < hi


def test_x_init_equals_reference_live(pkg, refs):
    for layout, ref in (("col", refs.col), ("row", refs.row)):
        for vt in ("dp", "sp"):
            for (lo, hi, n_rows, n_pad, bvs) in ((-7.0, 3.0, 513, 544, 1), (1e-3, 2.0, 4097, 4128, 5)):
                want = ref.random_x(vt, lo, hi, n_pad * bvs, n_rows, n_pad, bvs)
                got = _product_x(pkg.capi, vt, lo, hi, n_rows, n_pad, bvs, layout)
                assert np.array_equal(got.view(np.uint8), want.view(np.uint8)), (layout, vt, n_rows, bvs)


def test_x_init_hp_and_no_padding_rule(pkg):
    out = np.zeros(100, np.float16)
    pkg.capi.call("uspmv_random_init_host", -1.0, 1.0, 100, 2, out.ctypes.data_as(C.c_void_p), -1, 0, 1, 0)
    dp = _product_x(pkg.capi, "dp", -1.0, 1.0, 100, 100, 1, "col")
    assert np.array_equal(out, dp.astype(np.float16))  # one rounding double -> fp16, like static_cast<_Float16>(double)


@pytest.mark.parametrize("layout", ["col", "row"])
def test_adapter_compiles_against_the_reference_typedefs(layout):
    """oracle/_ref/adapter_check_* is built by oracle/Makefile from oracle/adapter_check.cpp, which assigns every launcher of
    include/uspmv_harness_adapter.hpp to SpmvKernel<VT,IT>::OnePrecFuncPtr / MultiPrecFuncPtr of the reference's own
    classes_structs.hpp; that it exists means the assignment compiled.  (It is RUN on the GPU by tests/test_gpu_boundary.py.)"""
    exe = os.path.join(ROOT, "oracle", "_ref", f"adapter_check_{layout}")
    if not os.path.exists(exe):
        if not os.path.isdir("/root/reference/code"):
            pytest.skip("oracle/_ref not built and /root/reference absent")
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "all"], check=True)
    r = subprocess.run([exe, "--compile-only"], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and "assignable to OnePrecFuncPtr / MultiPrecFuncPtr" in r.stdout and f"{layout}wise" in r.stdout, r.stdout + r.stderr
