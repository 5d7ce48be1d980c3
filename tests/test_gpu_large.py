"""Parity where the oracle cannot follow (BASELINE.json configs 3 and 4 are too big for NumPy to
finish in seconds): size-independent properties plus oracle checks on sub-problems.

 * the W half-step is row-local (_solver.py:53-57), so a row block of a big problem must equal
   the oracle run on that block alone;
 * the H half-step partials are sums over rows, so H' of the full problem must not depend on
   how rows are split across CTAs / shards (checked against fp64 oracle on a thin slice);
 * objective monotone, simplex exact, loss equals an independent evaluation of the factors."""
import numpy as np
import pytest

import nbmf_oracle as orc
from conftest import rel_err
from nbmf_mm_b200 import NBMF, BitMatrix
from nbmf_mm_b200.device import synth_bits_device
from nbmf_mm_b200.solver import PreparedData, make_problem

pytestmark = pytest.mark.gpu


def big_problem(m, n, k, seed=0, obs=0.9, dtype="float32"):
    hstar = (np.random.default_rng(seed).random((k, n)) * 0.2).astype(np.float32)
    P, M = synth_bits_device(seed, 0, m, n, hstar, obs, "cuda")
    data = PreparedData(m, n, "bits", P, M, None, float(M.count()))
    return data


@pytest.mark.parametrize("dtype,tol", [("float64", 1e-9), ("float32", 1e-4)])
def test_config3_shape_step_vs_oracle_blocks(dtype, tol):
    """Internal problem of config 3 (dir-beta on 20000x5000 -> 5000 x 20000), K=20, 90% observed."""
    m, n, k = 5000, 20000, 20
    data = big_problem(m, n, k, seed=3, dtype=dtype)
    rng = np.random.RandomState(0)
    W0 = rng.uniform(0.1, 0.9, (m, k)); H0 = rng.uniform(0.1, 0.9, (k, n))
    prob = make_problem(data, k, dtype=dtype, alpha=1.2, beta=1.2, eps=1e-8, mask_semantics="reference",
                        projection="normalize", max_iter_cap=4, device=None)
    with prob:
        prob.set_factors(W0, H0, normalize_w=True)
        Wn, _ = prob.get_factors()
        prob.h_half_step()
        _, H1 = prob.get_factors()
        prob.w_half_step()
        W1, _ = prob.get_factors()
        loss = prob.objective()
    # H' on a thin column slice: needs all rows but only 64 columns of V
    cols = slice(9984, 10048)
    Y = data.P.to_dense()[:, cols]; Mk = data.M.to_dense()[:, cols]
    Ho = orc.h_half_step(Y, Wn.T, H0[:, cols], Mk, 1.2, 1.2)
    assert rel_err(H1[:, cols], Ho) < tol
    # W' on a row block: row-local, needs all columns of 96 rows
    rows = slice(2400, 2496)
    Yr = data.P.to_dense()[rows]; Mr = data.M.to_dense()[rows]
    Wo = orc.w_half_step(Yr, Wn[rows].T, H1, Mr)
    assert rel_err(W1[rows], Wo.T) < tol
    assert np.max(np.abs(W1.sum(axis=1) - 1.0)) < (1e-12 if dtype == "float64" else 1e-6)
    # the fused loss equals an independent blocked evaluation of (W1, H1)
    ll, nobs = 0.0, 0.0
    for r0 in range(0, m, 1000):
        Yb = data.P.rows(r0, r0 + 1000).to_dense(); Mb = data.M.rows(r0, r0 + 1000).to_dense()
        th = W1[r0:r0 + 1000] @ H1
        pos = Yb * Mb
        ll += np.sum(pos * np.log(th + 1e-8) + (1 - pos) * np.log(1 - th + 1e-8)); nobs += Mb.sum()
    want = -(ll + 0.2 * np.sum(np.log(H1 + 1e-8)) + 0.2 * np.sum(np.log(1 - H1 + 1e-8))) / nobs
    assert abs(loss - want) < (1e-10 if dtype == "float64" else 2e-6) * abs(want)


def test_fit_properties_at_scale_fp32():
    """2e8 entries, K=32 (config 4's kernel variant), masked: monotone, simplex, deterministic."""
    m, n, k = 20000, 10000, 32
    data = big_problem(m, n, k, seed=4)
    rng = np.random.RandomState(1)
    W0 = rng.uniform(0.1, 0.9, (m, k)); H0 = rng.uniform(0.1, 0.9, (k, n))
    runs = []
    for _ in range(2):
        prob = make_problem(data, k, dtype="float32", alpha=1.2, beta=1.2, eps=1e-8, mask_semantics="reference",
                            projection="normalize", max_iter_cap=12, device=None)
        with prob:
            prob.set_factors(W0, H0, normalize_w=True)
            losses, n_iter, conv = prob.fit(12, 0.0)
            W, H = prob.get_factors()
        runs.append((losses, W, H))
    losses, W, H = runs[0]
    assert len(losses) == 12 and np.all(np.diff(losses) < 0)
    assert np.max(np.abs(W.sum(axis=1) - 1.0)) < 1e-6 and H.min() > 0 and H.max() <= 1
    assert np.array_equal(losses, runs[1][0]) and np.array_equal(W, runs[1][1]) and np.array_equal(H, runs[1][2])


def test_estimator_accepts_device_resident_bitmatrix():
    m, n, k = 3000, 2500, 12
    data = big_problem(m, n, k, seed=5)
    est = NBMF(n_components=k, max_iter=15, tol=0.0, random_state=0, dtype="float32").fit(data.P, mask=data.M)
    host = NBMF(n_components=k, max_iter=15, tol=0.0, random_state=0, dtype="float32").fit(
        data.P.to_dense(), mask=data.M.to_dense())
    assert np.array_equal(est.W_, host.W_) and np.array_equal(est.components_, host.components_)
    t = NBMF(n_components=k, max_iter=5, tol=0.0, random_state=0, dtype="float32", orientation="dir-beta").fit(
        data.P, mask=data.M)
    t2 = NBMF(n_components=k, max_iter=5, tol=0.0, random_state=0, dtype="float32", orientation="dir-beta").fit(
        data.P.to_dense(), mask=data.M.to_dense())
    assert np.array_equal(t.W_, t2.W_)
