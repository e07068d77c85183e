"""CPU: properties of the compiled sm_100a kernels that the performance design depends on, read from the built library with
cuobjdump (no GPU needed): the library holds sm_100a code only, the streamed kernels use the bulk-copy engine + mbarriers
(B200_PROFILING.md: UBLKCP / SYNCS), and the kernels whose occupancy plan is 32 warps per SM stay within 64 registers with no spills."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "ultimate-spmv_b200", "lib", "libuspmv_b200.so")
CUOBJDUMP = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"

pytestmark = pytest.mark.skipif(not os.path.exists(CUOBJDUMP), reason="cuobjdump not available")


def _res_usage():
    out = subprocess.run([CUOBJDUMP, "-res-usage", LIB], capture_output=True, text=True, timeout=300).stdout
    usage = {}
    for m in re.finditer(r"Function (\S+):\s*\n\s*REG:(\d+) STACK:(\d+)", out):
        usage[m.group(1)] = (int(m.group(2)), int(m.group(3)))
    return out, usage


def test_library_is_sm_100a_only():
    out = subprocess.run([CUOBJDUMP, "-lelf", LIB], capture_output=True, text=True, timeout=120).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_streamed_kernels_register_budget():
    _, usage = _res_usage()
    assert len(usage) > 300

    def of(pattern):
        hits = {k: v for k, v in usage.items() if re.search(pattern, k)}
        assert hits, pattern
        return hits
    # SELL-32 SpMV, the dominant kernel: (8 slots, depth 2, 16 warps) instances, fused and plain: <= 64 registers, nothing spilled
    for name, (reg, stack) in of(r"k_scs32_streamI[df6].*Li8ELi2ELi16E").items():
        # the single-GPU instances spill nothing; the fused halo-exchange instances (push / flag code on top) may keep <= 32 bytes of stack
        assert reg <= 64 and stack <= (32 if "Lb0ELb1E" in name or "Lb1ELb1E" in name else 0), (name, reg, stack)
    # wide chunks (C = 64 / 128): 64-register budget; at most a few bytes of stack in the fp64 H = 4 un-permuted instance
    for name, (reg, stack) in of(r"k_scsw_stream").items():
        assert reg <= 64 and stack <= 16, (name, reg, stack)
    # narrow chunks (C = 16): the fp64 instance that is dispatched has 12 warps per CTA and must not spill (its pending metadata loads)
    for name, (reg, stack) in of(r"k_scsn_streamId.*Li2ELi2ELi12E").items():
        assert reg <= 85 and stack == 0, (name, reg, stack)
    # streamed CRS
    for name, (reg, stack) in of(r"k_csr_stream").items():
        assert reg <= 64 and stack == 0, (name, reg, stack)


def test_streamed_kernels_use_the_bulk_copy_engine():
    sass = subprocess.run([CUOBJDUMP, "-sass", "-fun", "k_scs32_stream", LIB], capture_output=True, text=True, timeout=600).stdout
    if "UBLKCP" not in sass:  # -fun wants the mangled name on some versions: fall back to the whole library
        sass = subprocess.run([CUOBJDUMP, "-sass", LIB], capture_output=True, text=True, timeout=900).stdout
    assert "UBLKCP.S.G" in sass, "cp.async.bulk global -> shared (TMA bulk copy) missing"
    assert "SYNCS.ARRIVE.TRANS64" in sass and "SYNCS.PHASECHK.TRANS64.TRYWAIT" in sass, "mbarrier expect_tx / try_wait missing"


def _gather_destinations(mangled_pattern):
    """Destination registers of the 128-bit read-only gathers (LDG.E.128.CONSTANT) of one kernel instance, in SASS order."""
    _, usage = _res_usage()
    names = [k for k in usage if re.search(mangled_pattern, k)]
    assert len(names) == 1, (mangled_pattern, names)
    sass = subprocess.run([CUOBJDUMP, "-sass", "-fun", names[0], LIB], capture_output=True, text=True, timeout=300).stdout
    return re.findall(r"LDG\.E\.128\.CONSTANT\s+(R\d+)", sass)


@pytest.mark.parametrize("what,pattern", [
    ("sp block_vec_size 8, 8 warps, MINB 4", r"k_scs32_stream_mmvIf.*Li8ELi2ELi8ELi8ELb1ELb1ELb0ELi4E"),
    ("dp block_vec_size 8, 8 warps, MINB 3", r"k_scs32_stream_mmvId.*Li8ELi2ELi8ELi8ELb1ELb1ELb0ELi3E"),
    ("dp block_vec_size 4, 8 warps, MINB 3", r"k_scs32_stream_mmvId.*Li8ELi2ELi8ELi4ELb1ELb1ELb0ELi3E"),
])
def test_spmmv_kernels_keep_their_gathers_in_flight(what, pattern):
    """Round 2 (profiles/r02q_mmv_fused_instance.md): without a register budget in __launch_bounds__ ptxas re-serialises the gathers a
    piece issues together (destinations R16 R16 R16 ... or R20 R16 R20 R16 ...: one or two in flight per lane) and the kernel loses
    3-25 %.  The shipped row-major SpMMV instances state a budget; here: some run of 8 consecutive gathers lands in >= 6 DIFFERENT
    registers, i.e. they really are in flight together."""
    dst = _gather_destinations(pattern)
    assert len(dst) >= 16, (what, len(dst))
    best = max(len(set(dst[i:i + 8])) for i in range(len(dst) - 7))
    assert best >= 6, (what, best, dst[:40])
