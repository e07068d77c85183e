import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


@pytest.fixture(scope="session")
def pkg():
    return importlib.import_module("ultimate-spmv_b200")


@pytest.fixture(scope="session")
def mats(pkg):
    return pkg.matrices


@pytest.fixture(scope="session")
def orc():
    from oracle.bindings import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def refs():
    """The compiled, unmodified reference (oracle/_ref).  Built in the dev container; prebuilt on the GPU box."""
    from oracle import bindings
    if not bindings.ref_available():
        if os.path.isdir("/root/reference/code"):
            bindings.build("ref")
        else:
            pytest.skip("oracle/_ref not built and /root/reference absent")
    from types import SimpleNamespace
    return SimpleNamespace(col=bindings.Ref("col"), row=bindings.Ref("row"), iface=bindings.RefIface())


@pytest.fixture(scope="session")
def eng(pkg):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return pkg.engine


def load_matrix(name):
    """COO of one of the reference's small matrices, from the committed fixture (tests/golden/matrices.npz)."""
    z = np.load(os.path.join(GOLDEN, "matrices.npz"))
    n = int(z[f"{name}__n"])
    return n, n, z[f"{name}__I"], z[f"{name}__J"], z[f"{name}__V"]


MATRIX_NAMES = ["FDM-2d-16", "bcsstk13", "impcol_e", "matrix1", "matrix1int", "matrix1ones", "matrix_band_klein", "myBigMat", "myMat",
                "mySymmMat"]
