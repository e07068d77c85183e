"""CPU: the reference arm of bench.py (`--impl reference`) prints ONE JSON line with the keys of the measurement contract, also when
launched the way torchrun launches it (OMP_NUM_THREADS=1 in the environment, RANK / WORLD_SIZE set); ranks > 0 print nothing."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra, *args):
    env = dict(os.environ, **env_extra)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "laplace7_48", "--steps", "3",
                           "--warmup", "1", *args], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)


@pytest.fixture(scope="module")
def ref_built():
    from oracle import bindings
    if not bindings.ref_available():
        pytest.skip("oracle/_ref not built (needs /root/reference)")


def test_reference_arm_json_line(ref_built):
    r = _run({"OMP_NUM_THREADS": "1", "TORCHELASTIC_RUN_ID": "x", "RANK": "0", "WORLD_SIZE": "2"}, "--gpus", "2")
    assert r.returncode == 0, r.stderr[-800:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
              "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "GFLOP/s" and d["n_gpus"] == 2 and d["dtype"] == "f64" and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["value"] == d["value"] == d["e2e"]["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    # torchrun exports OMP_NUM_THREADS=1: the arm must still use every host core it may run on
    assert d["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))
    assert d["value"] > 0 and "workload" in d["config"]


def test_reference_arm_other_ranks_are_silent(ref_built):
    r = _run({"RANK": "1", "WORLD_SIZE": "2"}, "--gpus", "2")
    assert r.returncode == 0 and r.stdout.strip() == ""
