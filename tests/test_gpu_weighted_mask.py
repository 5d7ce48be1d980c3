"""Weighted (non-0/1) observation masks.  The reference multiplies by the mask VALUES: ``Y * mask`` in the H half-step and
the loss, ``(1 - Y).T * mask.T`` in the W half-step, ``count_nonzero(mask)`` as the loss normaliser
(``_solver.py:30-32,43,153-155``).  The oracle restates exactly that (``oracle/nbmf_oracle.py:_weights``, pinned bit for bit
by the golden vectors for 0/1 masks; the same expressions serve any mask value).  The CUDA path carries such a mask in
the dense layout: V * mask, the mask values, and the bit plane of (mask != 0)."""
import numpy as np
import pytest
import scipy.sparse as sp

import nbmf_oracle as orc
from conftest import rel_err
from nbmf_mm_b200 import NBMF, BitMatrix, nbmf_mm_solver, nbmf_mm_update_beta_dir

pytestmark = pytest.mark.gpu


def _problem(m, n, k, seed, binary=True):
    rng = np.random.default_rng(seed)
    Y = (rng.random((m, n)) < 0.3).astype(np.float64) if binary else rng.random((m, n))
    mask = rng.choice([0.0, 0.25, 0.5, 1.0, 1.0, 2.0], size=(m, n))        # weights, some zero, some above one
    W = rng.uniform(0.1, 0.9, (k, m)); W /= W.sum(axis=0, keepdims=True)
    H = rng.uniform(0.05, 0.95, (k, n))
    return Y, mask, W, H


@pytest.mark.parametrize("m,n,k,binary", [(70, 90, 5, True), (257, 130, 12, True), (64, 300, 33, False), (300, 64, 7, False)])
@pytest.mark.parametrize("dtype,tol", [("float64", 1e-9), ("float32", 5e-5)])
def test_one_step_with_a_weighted_mask(m, n, k, binary, dtype, tol):
    Y, mask, W, H = _problem(m, n, k, seed=m + k, binary=binary)
    Wo, Ho = orc.mm_step(Y, W, H, mask, 1.2, 1.3)
    W1, H1 = nbmf_mm_update_beta_dir(Y, W, H, mask, 1.2, 1.3, dtype=dtype)
    assert rel_err(H1, Ho) < tol and rel_err(W1, Wo) < tol
    Wb, Hb = orc.mm_step(Y, W, H, (mask != 0).astype(float), 1.2, 1.3)    # the weights do matter
    assert rel_err(Ho, Hb) > 1e-3 and rel_err(Wo, Wb) > 1e-3


@pytest.mark.parametrize("orientation", ["beta-dir", "dir-beta"])
def test_trajectory_with_a_weighted_mask(orientation):
    Y, mask, _, _ = _problem(120, 80, 6, seed=11)
    W, H, losses, _, n_iter = nbmf_mm_solver(Y, 6, max_iter=60, tol=1e-7, mask=mask, random_state=3, orientation=orientation)
    Wo, Ho, lo, no = orc.fit(Y, 6, max_iter=60, tol=1e-7, mask=mask, random_state=3, orientation=orientation)
    assert n_iter == no
    assert np.max(np.abs(np.asarray(losses) - np.asarray(lo)) / np.abs(lo)) < 1e-9
    assert rel_err(W, Wo) < 1e-8 and rel_err(H, Ho) < 1e-8


def test_estimator_inputs_and_limits():
    Y, mask, _, _ = _problem(90, 70, 4, seed=5)
    kw = dict(n_components=4, max_iter=25, tol=0.0, random_state=1)
    dense = NBMF(**kw).fit(Y, mask=mask)
    sparse = NBMF(**kw).fit(sp.csr_matrix(Y), mask=sp.csr_matrix(mask))
    mixed = NBMF(**kw).fit(sp.csr_matrix(Y), mask=mask)
    for other in (sparse, mixed):
        assert np.array_equal(other.W_, dense.W_) and np.array_equal(other.components_, dense.components_)
    ref = orc.fit(Y, 4, max_iter=25, tol=0.0, mask=mask, random_state=1)
    assert abs(dense.loss_curve_[-1] - ref[2][-1]) < 1e-9 * abs(ref[2][-1])
    with pytest.raises(ValueError, match="weighted"):
        NBMF(**kw).fit(BitMatrix.from_dense(Y), mask=mask)                    # bit-packed X cannot carry mask values
    with pytest.raises(ValueError, match="reference"):
        NBMF(mask_semantics="strict", **kw).fit(Y, mask=mask)
