"""Behavioural contract of the drop-in estimator on the GPU: the properties the reference's own
suite checks through ``NBMF``, restated for this repo -- shapes, simplex
and box constraints, strict monotonicity in FP64, bit-exact reproducibility, input kinds."""
import numpy as np
import pytest

from nbmf_mm_b200 import NBMF, NBMFMM, BitMatrix
from nbmf_mm_b200._utils import generate_synthetic_binary_data

pytestmark = pytest.mark.gpu


def toy(m=60, n=80, p=0.25, seed=0):
    return (np.random.default_rng(seed).random((m, n)) < p).astype(float)


@pytest.mark.parametrize("orientation,k", [("beta-dir", 6), ("dir-beta", 8), ("Aspect Bernoulli", 3), ("bICA", 5)])
def test_shapes_constraints_and_canonical_orientation(orientation, k):
    X = toy()
    est = NBMF(n_components=k, orientation=orientation, max_iter=150, tol=1e-6, random_state=0).fit(X)
    assert est.orientation in ("beta-dir", "dir-beta")
    assert est.W_.shape == (60, k) and est.components_.shape == (k, 80)
    assert np.isfinite(est.objective_history_[-1]) and len(est.objective_history_) == est.n_iter_
    if est.orientation == "beta-dir":
        assert np.allclose(est.W_.sum(axis=1), 1.0, atol=1e-10)
        cont = est.components_
    else:
        assert np.allclose(est.components_.sum(axis=0), 1.0, atol=1e-10)
        cont = est.W_
    assert np.all(cont > 0) and np.all(cont < 1) and len(np.unique(cont)) > 100
    assert np.all(est.W_ >= 0)


@pytest.mark.parametrize("orientation", ["beta-dir", "dir-beta"])
@pytest.mark.parametrize("masked", [False, True])
def test_strictly_monotone_map_objective_fp64(orientation, masked):
    rng = np.random.default_rng(42)
    X = (rng.random((50, 70)) < 0.3).astype(float)
    mask = (rng.random(X.shape) < 0.8) if masked else None                  # bool mask, as the reference accepts
    est = NBMF(n_components=8, orientation=orientation, alpha=1.1, beta=1.3, max_iter=100, tol=1e-8,
               random_state=1).fit(X, mask=mask)
    h = np.asarray(est.loss_curve_)
    assert np.all(h[1:] <= h[:-1] + 1e-12)
    assert h[0] - h[-1] > 1e-3


def test_prior_pulls_H_in_the_right_direction():
    X = toy(100, 50, 0.3, seed=42)
    lo = NBMF(n_components=10, alpha=0.5, beta=2.0, max_iter=100, random_state=42).fit(X).components_.mean()
    hi = NBMF(n_components=10, alpha=2.0, beta=0.5, max_iter=100, random_state=42).fit(X).components_.mean()
    assert lo < hi


def test_reproducible_bit_for_bit_and_seed_sensitivity():
    X, _, _ = generate_synthetic_binary_data(50, 30, 5, random_state=42)
    for dtype in ("float64", "float32"):
        a = NBMFMM(n_components=5, random_state=42, max_iter=50, dtype=dtype).fit(X)
        b = NBMFMM(n_components=5, random_state=42, max_iter=50, dtype=dtype).fit(X)
        assert np.array_equal(a.components_, b.components_) and np.array_equal(a.W_, b.W_)
        assert a.loss_curve_ == b.loss_curve_
    c = NBMFMM(n_components=5, random_state=43, max_iter=50).fit(X)
    assert not np.allclose(a.components_, c.components_)


def test_tolerance_controls_iterations():
    X, _, _ = generate_synthetic_binary_data(50, 30, 5, random_state=42)
    fast = NBMFMM(n_components=5, tol=0.1, max_iter=1000, random_state=42).fit(X)
    slow = NBMFMM(n_components=5, tol=1e-8, max_iter=1000, random_state=42).fit(X)
    assert fast.n_iter_ < 50 < slow.n_iter_


def test_custom_init_unnormalised_W_and_binary_H():
    X, W0, H0 = generate_synthetic_binary_data(50, 30, 5, random_state=42)     # H0 is {0,1}: exercises the clip
    est = NBMFMM(n_components=5, init="custom", W_init=W0, H_init=H0, max_iter=10).fit(X)
    assert est.n_iter_ <= 10 and np.all(np.isfinite(est.W_)) and np.all(np.isfinite(est.components_))
    assert np.allclose(est.W_.sum(axis=1), 1.0, atol=1e-10)
    rng = np.random.default_rng(5)
    Wd = rng.dirichlet(np.ones(4), size=20)
    Hc = np.clip(rng.random((4, 25)), 0.05, 0.95)
    Xs = (rng.random((20, 25)) < 0.4).astype(float)
    est = NBMF(n_components=4, W_init=Wd, H_init=Hc, max_iter=40, tol=0.0).fit(Xs)
    h = np.asarray(est.loss_curve_)
    assert np.all(h[1:] <= h[:-1] + 1e-12) and np.allclose(est.W_.sum(axis=1), 1.0, atol=1e-10)


def test_input_kinds_sparse_bool_probabilities_bitmatrix():
    sp = pytest.importorskip("scipy.sparse")
    X = toy()
    mask = (np.random.default_rng(1).random(X.shape) < 0.8).astype(float)
    kw = dict(n_components=4, max_iter=40, tol=1e-6, random_state=0)
    dense = NBMF(**kw).fit(X, mask=mask)
    sparse = NBMF(**kw).fit(sp.csr_matrix(X), mask=sp.csr_matrix(mask))
    boolm = NBMF(**kw).fit(X, mask=mask.astype(bool))
    packed = NBMF(**kw).fit(BitMatrix.from_dense(X), mask=BitMatrix.from_dense(mask))
    for other in (sparse, boolm, packed):
        assert np.array_equal(other.W_, dense.W_) and np.array_equal(other.components_, dense.components_)
    Xp = np.random.default_rng(2).random((40, 30))                             # probabilities are accepted
    est = NBMF(n_components=5, max_iter=30, random_state=0).fit(Xp)
    assert est.W_.shape == (40, 5) and np.isfinite(est.loss_)
    assert np.isfinite(NBMF(**kw).fit(X, mask=mask * 0.5).loss_)              # weighted masks: tests/test_gpu_weighted_mask.py


def test_fit_transform_and_reconstruction_quality():
    X, _, _ = generate_synthetic_binary_data(100, 50, 5, sparsity=0.3, random_state=42)
    est = NBMFMM(n_components=5, max_iter=200, random_state=42)
    W = est.fit_transform(X)
    assert W.shape == (100, 5) and np.array_equal(W, est.W_)
    Xhat = est.inverse_transform(est.W_)
    assert Xhat.shape == X.shape and np.all((Xhat >= 0) & (Xhat <= 1))
    assert np.mean(np.abs(X - (Xhat > 0.5))) < 0.4


def test_symmetry_between_orientations():
    """dir-beta on X == beta-dir on X^T with the same seed (same internal problem and RNG stream)."""
    X = toy(25, 30, 0.3, seed=7)
    a = NBMF(n_components=5, orientation="dir-beta", alpha=1.3, beta=1.7, max_iter=60, random_state=3).fit(X)
    b = NBMF(n_components=5, orientation="beta-dir", alpha=1.3, beta=1.7, max_iter=60, random_state=3).fit(X.T)
    assert np.allclose(a.W_ @ a.components_, (b.W_ @ b.components_).T, atol=1e-12)
    assert np.allclose(a.loss_curve_, b.loss_curve_, rtol=1e-12)


def test_n_init_keeps_the_best_restart():
    X = toy(40, 35, 0.3, seed=9)
    kw = dict(n_components=4, max_iter=60, tol=0.0, engine="simt")      # one engine on both sides: equal bit for bit
    singles = [NBMF(random_state=10 + r, **kw).fit(X) for r in range(3)]
    multi = NBMF(random_state=10, n_init=3, **kw).fit(X)
    finals = [s.loss_ for s in singles]
    assert multi.loss_ == min(finals) and multi.best_init_ == int(np.argmin(finals))
    assert np.array_equal(multi.W_, singles[multi.best_init_].W_)
    one = NBMF(random_state=10, n_init=1, **kw).fit(X)
    assert np.array_equal(one.W_, singles[0].W_)                               # r = 0 reproduces the reference


def test_duchi_is_near_identical_to_normalize():
    """README.md:27-30 of the reference: the two projections give near-identical fits."""
    X = toy(60, 80, 0.25, seed=3)
    mask = (np.random.default_rng(4).random(X.shape) < 0.9).astype(float)
    a = NBMF(n_components=6, max_iter=100, tol=0.0, random_state=0).fit(X, mask=mask)
    b = NBMF(n_components=6, max_iter=100, tol=0.0, random_state=0, projection_method="duchi").fit(X, mask=mask)
    assert abs(a.loss_ - b.loss_) / a.loss_ < 1e-3
    assert np.allclose(b.W_.sum(axis=1), 1.0, atol=1e-10) and np.all(b.W_ >= 0)
    with pytest.raises(ValueError, match="projection_method"):
        NBMF(projection_method="softmax").fit(X)
