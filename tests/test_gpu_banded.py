"""GPU: column-banded execution plan (uspmv_banded_*, opt-in) against the un-banded kernels and the COO sum.
The plan was measured in round 2 (BASELINE config 4 at sigma = 512: 15.0 ms against 10.2 ms for the fused kernel, bench.py
other_configs) and is therefore not a default, but it is a shipped entry point and stays covered (21 cases, 11 s on a B200:
profiles/r02M_pytest_banded.log)."""
import numpy as np
import pytest

pytestmark = [pytest.mark.gpu]


def _matrix(n, seed):
    rng = np.random.default_rng(seed)
    cnt = rng.integers(0, 12, n)
    cnt[rng.choice(n, 6, replace=False)] = rng.integers(300, 900, 6)  # a few long rows: the segment path of the streamed kernels
    I = np.repeat(np.arange(n), cnt).astype(np.int32)
    J = rng.integers(0, n, len(I)).astype(np.int32)
    V = np.sign(rng.standard_normal(len(I))) * 10.0 ** rng.uniform(-3, 1, len(I))
    return I, J, V


@pytest.mark.parametrize("C,sigma", [(32, 128), (16, 16), (1, 1)])
@pytest.mark.parametrize("vt", ["dp", "sp"])
@pytest.mark.parametrize("n_bands", [1, 3, 8])
def test_banded_spmv_matches_coo_sum(eng, vt, C, sigma, n_bands):
    import torch as t
    n = 20000
    I, J, V = _matrix(n, seed=C + n_bands)
    mtx = eng.MtxData.from_host(n, n, I, J, V)
    plan = eng.BandedPlan(mtx, C, sigma, vt, n_bands=n_bands)
    assert plan.n_bands == n_bands and plan.nnz == len(I)
    npt = {"dp": np.float64, "sp": np.float32}[vt]
    x = np.random.default_rng(1).uniform(-1, 1, n).astype(npt)
    y = t.zeros(plan.n_rows_padded, dtype=t.float64 if vt == "dp" else t.float32, device="cuda")
    plan.spmv(t.from_numpy(x).cuda(), y)
    t.cuda.synchronize()
    got = y.cpu().numpy().astype(np.float64)[plan.old_to_new]
    Vs = V.astype(npt).astype(np.float64)
    ref = np.zeros(n)
    np.add.at(ref, I, Vs * x.astype(np.float64)[J])
    scale = np.zeros(n)
    np.add.at(scale, I, np.abs(Vs * x.astype(np.float64)[J]))
    tol = 1e-12 if vt == "dp" else 1e-5
    assert np.all(np.abs(got - ref) <= tol * np.maximum(scale, 1e-300))


@pytest.mark.parametrize("mode", ["ap[dp_sp_hp]", "ap[dp_sp]", "ap[sp_hp]"])
def test_banded_ap_matches_unbanded(eng, mode):
    import torch as t
    n = 20000
    I, J, V = _matrix(n, seed=7)
    mtx = eng.MtxData.from_host(n, n, I, J, V)
    one = eng.BandedPlan(mtx, 32, 128, ap=mode, t1=0.5, t2=0.01, n_bands=1)     # = the un-banded fused AP kernel
    many = eng.BandedPlan(mtx, 32, 128, ap=mode, t1=0.5, t2=0.01, n_bands=5)
    assert np.array_equal(one.old_to_new, many.old_to_new)
    dt = t.float32 if mode == "ap[sp_hp]" else t.float64
    x = t.from_numpy(np.random.default_rng(2).uniform(-1, 1, n)).to(dt).cuda()
    y1 = t.zeros(one.n_rows_padded, dtype=dt, device="cuda")
    y5 = t.zeros(many.n_rows_padded, dtype=dt, device="cuda")
    one.spmv(x, y1)
    many.spmv(x, y5)
    t.cuda.synchronize()
    xs = x.double().cpu().numpy()
    scale = np.zeros(n)
    np.add.at(scale, I, np.abs(V * xs[J]))
    sp = np.zeros(one.n_rows_padded)
    sp[one.old_to_new] = scale
    tol = 1e-5 if mode == "ap[sp_hp]" else 1e-12
    assert np.all(np.abs((y1 - y5).double().cpu().numpy()) <= tol * np.maximum(sp, 1e-300))
