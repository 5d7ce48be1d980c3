"""Degenerate inputs through the C-ABI path, both engines, against the oracle: single rows / columns / components,
all-zero and all-one data, rows and columns without a single observed entry, everything unobserved but one entry.
A row without observed entries has a zero multiplicative step, which the reference's L1 renormalisation turns into
0/0 = NaN (_solver.py:57): the drop-in reproduces the NaN pattern, it does not paper over it."""
import numpy as np
import pytest

import nbmf_oracle as orc
from conftest import rel_err
from nbmf_mm_b200 import NBMF, nbmf_mm_update_beta_dir

pytestmark = pytest.mark.gpu


def _factors(m, n, k, seed):
    rng = np.random.default_rng(seed)
    W = rng.uniform(0.1, 0.9, (k, m)); W /= W.sum(axis=0, keepdims=True)
    return W, rng.uniform(0.05, 0.95, (k, n))


CASES = {
    "one_row": lambda rng: ((rng.random((1, 40)) < 0.3).astype(float), None),
    "one_col": lambda rng: ((rng.random((40, 1)) < 0.3).astype(float), None),
    "all_zero": lambda rng: (np.zeros((37, 45)), (rng.random((37, 45)) < 0.8).astype(float)),
    "all_one": lambda rng: (np.ones((37, 45)), None),
    "dead_row_and_col": lambda rng: _dead(rng),
    "one_observed_entry": lambda rng: ((rng.random((20, 30)) < 0.5).astype(float), _single(20, 30)),
}


def _dead(rng):
    X = (rng.random((50, 60)) < 0.3).astype(float)
    mask = (rng.random((50, 60)) < 0.9).astype(float)
    mask[7, :] = 0.0
    mask[:, 11] = 0.0
    return X, mask


def _single(m, n):
    mk = np.zeros((m, n)); mk[3, 4] = 1.0
    return mk


@pytest.mark.parametrize("name", sorted(CASES))
@pytest.mark.parametrize("k", [1, 5])
@pytest.mark.parametrize("engine,dtype,tol", [("simt", "float64", 1e-9), ("simt", "float32", 5e-5), ("tensor", "float32", 5e-5)])
def test_one_step_on_degenerate_inputs(name, k, engine, dtype, tol):
    X, mask = CASES[name](np.random.default_rng(5))
    W, H = _factors(X.shape[0], X.shape[1], k, seed=k)
    Wo, Ho = orc.mm_step(X, W, H, mask, 1.2, 1.3)
    with np.errstate(all="ignore"):
        Wo, Ho = orc.mm_step(X, W, H, mask, 1.2, 1.3)
        Ws, Hs = orc.mm_step(X, W, H, mask, 1.2, 1.3, mask_semantics="strict")
    W1, H1 = nbmf_mm_update_beta_dir(X, W, H, mask, 1.2, 1.3, dtype=dtype, engine=engine)
    W2, H2 = nbmf_mm_update_beta_dir(X, W, H, mask, 1.2, 1.3, dtype=dtype, engine=engine, mask_semantics="strict")
    for got, want in ((W1, Wo), (H1, Ho), (W2, Ws), (H2, Hs)):
        assert np.array_equal(np.isnan(got), np.isnan(want))          # same NaN pattern as the reference arithmetic
        ok = ~np.isnan(want)
        assert np.all(np.isfinite(got[ok])) and rel_err(got[ok], want[ok]) < tol
    assert np.isnan(Wo).any() == (name in ("dead_row_and_col", "one_observed_entry"))


@pytest.mark.parametrize("projection", ["normalize", "duchi"])
def test_fit_with_an_unobserved_column_stays_finite(projection):
    X, mask = _dead(np.random.default_rng(2))
    mask[7, :] = 1.0                                               # keep the dead column, revive the dead row
    for dtype in ("float64", "float32"):
        est = NBMF(n_components=4, max_iter=25, tol=0.0, random_state=0, dtype=dtype, projection_method=projection).fit(X, mask=mask)
        assert np.all(np.isfinite(est.W_)) and np.all(np.isfinite(est.components_)) and np.all(np.isfinite(est.loss_curve_))
        assert np.max(np.abs(est.W_.sum(axis=1) - 1.0)) < 1e-6
    _, _, losses, _ = orc.fit(X, 4, max_iter=25, tol=0.0, random_state=0, mask=mask, projection=projection)
    assert abs(est.loss_curve_[-1] - losses[-1]) < 1e-4 * abs(losses[-1])
