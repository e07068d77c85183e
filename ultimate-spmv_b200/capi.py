"""ctypes binding of the C ABI in include/uspmv_b200.h (ultimate-spmv_b200/lib/libuspmv_b200.so).

This is the only way Python reaches the engine: there is no CPU fallback and no PyTorch re-implementation
of any kernel.  If the shared library has not been built (``make -C ultimate-spmv_b200``) importing this
module raises, loudly.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("USPMV_B200_LIB") or os.path.join(HERE, "lib", "libuspmv_b200.so")  # the override is for A/B builds of the same ABI

F64, F32, F16 = 0, 1, 2
COLWISE, ROWWISE = 0, 1
AP_DP_SP, AP_DP_HP, AP_SP_HP, AP_DP_SP_HP = 0, 1, 2, 3
SEG_ROWS, SEG_NNZ = 0, 1


class UspmvError(RuntimeError):
    pass


if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build the CUDA extension first (python -c 'import __graft_entry__ as g; g.build()' "
        "or make -C ultimate-spmv_b200). There is no CPU fallback.")

lib = C.CDLL(LIB_PATH)
lib.uspmv_last_error.restype = C.c_char_p
lib.uspmv_kernel_launches.restype = C.c_long

vp = C.c_void_p
_sigs = {
    "uspmv_set_option": [C.c_char_p, C.c_long],
    "uspmv_ctx_set_option": [vp, C.c_char_p, C.c_long],
    "uspmv_ctx_get_option": [vp, C.c_char_p, C.POINTER(C.c_long)],
    "uspmv_device_count": [C.POINTER(C.c_int)],
    "uspmv_ctx_create": [C.c_int, C.POINTER(vp)],
    "uspmv_ctx_sync": [vp],
    "uspmv_malloc": [vp, C.c_size_t, C.POINTER(vp)],
    "uspmv_free": [vp, vp],
    "uspmv_memcpy_h2d": [vp, vp, vp, C.c_size_t, vp],
    "uspmv_memcpy_d2h": [vp, vp, vp, C.c_size_t, vp],
    "uspmv_memset": [vp, vp, C.c_int, C.c_size_t, vp],
    "uspmv_host_alloc": [C.c_size_t, C.POINTER(vp)],
    "uspmv_host_free": [vp],
    "uspmv_coo_from_host": [vp, C.c_long, C.c_long, C.c_long, vp, vp, vp, C.c_int, C.POINTER(vp)],
    "uspmv_coo_from_device": [vp, C.c_long, C.c_long, C.c_long, vp, vp, vp, C.c_int, C.POINTER(vp)],
    "uspmv_coo_stencil": [vp, C.c_int, C.c_long, C.c_long, C.c_long, C.c_long, C.c_long, C.POINTER(vp)],
    "uspmv_coo_powerlaw": [vp, C.c_long, C.c_long, C.c_long, C.c_double, C.c_double, C.c_int, C.c_ulong, C.POINTER(vp)],
    "uspmv_coo_device_arrays": [vp] + [C.POINTER(vp)] * 3,
    "uspmv_coo_from_entries": [vp, C.c_long, C.c_long, C.c_long, vp, vp, vp, C.c_int, C.POINTER(vp)],
    "uspmv_coo_equilibrate": [vp, vp, vp],
    "uspmv_coo_dims": [vp, C.POINTER(C.c_long)],
    "uspmv_coo_export": [vp, vp, vp, vp],
    "uspmv_scs_build": [vp, vp, C.c_long, C.c_long, C.c_int, vp, C.POINTER(vp)],
    "uspmv_scs_from_arrays": [vp, C.c_int, C.c_long, C.c_long, C.c_long, C.c_long, C.c_long, vp, vp, vp, vp, vp, C.c_int, C.POINTER(vp)],
    "uspmv_apply_strided_permutation": [vp, vp, vp, vp, C.c_long, C.c_long, C.c_int, vp],
    "uspmv_generate_inv_perm": [vp, vp, vp, C.c_long, C.c_long, vp],
    "uspmv_random_init_host": [C.c_double, C.c_double, C.c_long, C.c_int, vp, C.c_long, C.c_long, C.c_int, C.c_int],
    "uspmv_pointer_is_device": [vp, C.POINTER(C.c_int)],
    "uspmv_block_spmv_gpu": [vp, C.c_int, C.c_long, C.c_long, vp, vp, vp, vp, vp, vp, C.c_int, C.c_long, C.c_int, vp],
    "uspmv_scs_ap_gpu": [vp, C.c_int, C.c_long, C.c_long, C.POINTER(vp), vp, vp, vp],
    "uspmv_coo_seg_mtx": [vp, vp, C.c_int, C.c_int, C.POINTER(vp), C.POINTER(C.c_long)],
    "uspmv_scs_dims": [vp, C.POINTER(C.c_long)],
    "uspmv_scs_export": [vp, vp, vp, vp, vp, vp, vp],
    "uspmv_scs_permute_cols": [vp, vp],
    "uspmv_scs_device_arrays": [vp] + [C.POINTER(vp)] * 6,
    "uspmv_apply_permutation": [vp, vp, vp, vp, C.c_long, C.c_int, vp],
    "uspmv_apply_permutation_block": [vp, vp, vp, vp, C.c_long, C.c_int, C.c_int, C.c_long, C.c_int, vp],
    "uspmv_scs_gpu": [vp, C.c_int, C.c_long, C.c_long, vp, vp, vp, vp, vp, vp, vp],
    "uspmv_csr_gpu": [vp, C.c_int, C.c_long, vp, vp, vp, vp, vp, vp],
    "uspmv_spmv": [vp, vp, vp, vp],
    "uspmv_spmv_unpermuted": [vp, vp, vp, vp],
    "uspmv_spmmv": [vp, vp, vp, C.c_int, C.c_long, C.c_int, vp],
    "uspmv_spmv_host": [vp, vp, C.c_long, vp, C.c_long],
    "uspmv_spmv_host_submit": [vp, vp, C.c_long, vp, C.c_long, C.c_int],
    "uspmv_spmv_host_wait": [vp, C.c_int],
    "uspmv_partition_precisions": [vp, vp, C.c_int, C.c_double, C.c_double, vp, vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)],
    "uspmv_ap_spmv": [C.c_int, vp, vp, vp, vp, vp, vp],
    "uspmv_seg_work_sharing_arr": [C.c_int, C.c_long, C.c_long, vp, C.c_int, vp],
    "uspmv_halo_plan_create": [vp, vp, C.c_int, C.c_int, C.POINTER(vp)],
    "uspmv_halo_plan_create_multi": [vp, C.c_int, vp, C.c_int, C.c_int, C.c_int, C.POINTER(vp)],
    "uspmv_p2p_ap_spmv": [vp, C.c_int, vp, vp, vp, vp, vp, vp],
    "uspmv_halo_plan_counts": [vp, vp, C.POINTER(C.c_long)],
    "uspmv_halo_plan_need": [vp, vp, vp],
    "uspmv_halo_plan_set_send": [vp, vp, vp],
    "uspmv_scs_split_chunks": [vp, C.POINTER(C.c_long), C.POINTER(C.c_long)],
    "uspmv_spmv_part": [vp, C.c_int, vp, vp, vp],
    "uspmv_p2p_create": [vp, C.c_int, C.c_long, C.POINTER(vp), vp, C.POINTER(vp)],
    "uspmv_p2p_connect": [vp, vp, vp, vp],
    "uspmv_p2p_spmv": [vp, vp, vp, vp, vp],
    "uspmv_p2p_set_overlap": [vp, C.c_int],
    "uspmv_p2p_create_ex": [vp, C.c_int, C.c_long, C.c_int, C.c_int, C.c_int, C.POINTER(vp), vp, C.POINTER(vp)],
    "uspmv_p2p_connect_ex": [vp, vp, vp, vp, vp],
    "uspmv_p2p_spmv_buf": [vp, vp, C.c_int, C.c_int, vp, vp, vp],
    "uspmv_p2p_spmmv": [vp, vp, C.c_int, vp, vp, vp],
    "uspmv_spmmv_part": [vp, C.c_int, vp, vp, C.c_int, C.c_long, C.c_int, vp],
    "uspmv_spmmv_part_supported": [vp, C.c_int],
    "uspmv_p2p_status": [vp, C.POINTER(C.c_int), C.POINTER(C.c_long)],
    "uspmv_p2p_exchange": [vp, C.c_int, vp, vp],
    "uspmv_p2p_sync": [vp],
    "uspmv_p2p_disconnect": [vp],
    "uspmv_banded_build": [vp, vp, C.c_long, C.c_long, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, C.POINTER(vp)],
    "uspmv_banded_dims": [vp, vp],
    "uspmv_banded_perm": [vp, vp],
    "uspmv_banded_spmv": [vp, vp, vp, vp],
    "uspmv_p2p_spmv_host_submit": [vp, vp, vp, vp, C.c_int],
    "uspmv_p2p_spmv_host_wait": [vp, C.c_int],
    "uspmv_halo_pack": [vp, vp, vp, C.c_int, C.c_int, C.c_long, C.c_int, vp],
}
for _name, _args in _sigs.items():
    _fn = getattr(lib, _name, None)
    if _fn is not None:  # symbol presence is asserted by tests/test_capi_symbols.py against the header
        _fn.argtypes = _args
        _fn.restype = C.c_int
for _name in ("uspmv_ctx_destroy", "uspmv_coo_destroy", "uspmv_scs_destroy", "uspmv_halo_destroy", "uspmv_p2p_destroy", "uspmv_banded_destroy"):
    _fn = getattr(lib, _name, None)
    if _fn is not None:
        _fn.argtypes = [vp]
        _fn.restype = None


def check(rc: int) -> None:
    if rc != 0:
        raise UspmvError(lib.uspmv_last_error().decode(errors="replace"))


def call(name: str, *args) -> None:
    check(getattr(lib, name)(*args))


def set_option(name: str, value: int) -> None:
    call("uspmv_set_option", name.encode(), int(value))


def kernel_launches() -> int:
    return int(lib.uspmv_kernel_launches())
