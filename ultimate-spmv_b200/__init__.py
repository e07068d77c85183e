"""uspmv-b200: B200-native (sm_100a) SELL-C-sigma SpMV/SpMMV engine, drop-in for the compute path of
RRZE-HPC/Ultimate-SpMV.

Layout
    csrc/        hand-written CUDA + the C ABI (include/uspmv_b200.h) -> lib/libuspmv_b200.so
    host/        C++ host side: the `uspmv` harness CLI clone on top of include/uspmv_interface.hpp
    capi.py      ctypes binding of the C ABI
    engine.py    Python mirror of the reference's interface.hpp names (tests / bench plumbing)
    matrices.py  Matrix Market reader + synthetic generators (host logic)
    dist.py      one-process-per-GPU row partitioning + halo exchange plumbing (torch.distributed)
    validate.py  checked steps: y against the stencil formula / the COO triplets with x = f(global row)

The directory name contains a hyphen, so import it with
    importlib.import_module("ultimate-spmv_b200")
(tests/conftest.py and bench.py do exactly that).
"""
from . import matrices  # noqa: F401  (pure host logic, importable without the CUDA library)


def __getattr__(name):
    # capi/engine need the built shared library; import lazily so host-only logic stays importable,
    # but fail loudly (ImportError from capi) the moment a compute entry point is requested.
    if name in ("capi", "engine", "dist", "validate"):
        import importlib
        return importlib.import_module(f"{__name__}.{name}")
    raise AttributeError(name)
