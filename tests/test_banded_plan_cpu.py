"""CPU: the column-banded execution plan (uspmv_banded_*) restated with the oracle and with the compiled,
unmodified reference: per column band a SELL-C-sigma structure built with the reference's own fixed_permutation mechanism on the
sigma-sorted row order of the whole matrix; the band products, added in band order, give the product of the whole matrix."""
import numpy as np
import pytest


def _matrix(n, seed):
    rng = np.random.default_rng(seed)
    cnt = rng.integers(0, 10, n)
    cnt[rng.choice(n, 4, replace=False)] = rng.integers(100, 300, 4)
    I = np.repeat(np.arange(n), cnt).astype(np.int32)
    J = rng.integers(0, n, len(I)).astype(np.int32)
    V = rng.uniform(-1, 1, len(I))
    return I, J, V


@pytest.mark.parametrize("C,sigma,K", [(32, 64, 4), (8, 16, 3), (32, 1, 8)])
def test_banded_structures_with_oracle_and_reference(orc, refs, C, sigma, K):
    n = 1024  # a multiple of every C used: fixed_permutation never maps a real row onto a padding position (SURVEY.md section 8a' item 6)
    I, J, V = _matrix(n, seed=C + K)
    full = orc.convert_to_scs(n, n, I, J, V, C, sigma)
    perm = full.old_to_new.copy()
    x = np.random.default_rng(1).uniform(-1, 1, n)
    y_full = orc.spmv_scs(full, x)  # columns not permuted: x in the original numbering, y in permuted row order
    w = (n + K - 1) // K
    y_sum = np.zeros_like(y_full)
    slots = 0
    for b in range(K):
        sel = (J >= b * w) & (J < (b + 1) * w)
        s_o = orc.convert_to_scs(n, n, I[sel], J[sel], V[sel], C, sigma, fixed_perm=perm)
        s_r = refs.col.convert_to_scs(n, n, I[sel], J[sel], V[sel], C, sigma, fixed_perm=perm)
        for k in ("chunk_ptrs", "chunk_lengths", "col_idxs"):
            assert np.array_equal(getattr(s_o, k), getattr(s_r, k)), (b, k)
        assert np.array_equal(s_o.values.view(np.uint8), s_r.values.view(np.uint8))
        assert np.all(s_o.col_idxs[s_o.values != 0] // w == b)  # only this band's slice of x is touched (padding slots carry column 0)
        y_sum += refs.col.spmv_scs(s_r, x)
        slots += s_o.n_elements
    scale = np.zeros(n)
    np.add.at(scale, I, np.abs(V * x[J]))
    sp = np.zeros(len(y_full))
    sp[perm] = scale
    assert np.all(np.abs(y_sum - y_full) <= 1e-12 * np.maximum(sp, 1e-300))
    assert slots >= full.n_elements  # the price of the plan: more padding
