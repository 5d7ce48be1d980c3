"""One-step parity of the CUDA path (through the C-ABI) against the golden vectors of the
reference's ``nbmf_mm_update_beta_dir`` (_solver.py:5-59) and against the oracle.

Bars (BASELINE.json north_star): FP64 one-step W/H within 1e-9 relative; FP32 checked at 2e-5
(one step of fp32 rounding; the 1e-4 bar is on the final NLL, see test_gpu_trajectory.py)."""
import numpy as np
import pytest

import nbmf_oracle as orc
from conftest import rel_err
from nbmf_mm_b200 import nbmf_mm_update_beta_dir
from nbmf_mm_b200.solver import make_problem, prepare_data

pytestmark = pytest.mark.gpu

CASES = ["bin_nomask", "bin_mask", "bin_mask_k32", "prob_nomask", "prob_mask", "bin_alpha_lt1", "bin_k1", "bin_wide",
         "bin_mask_k100"]


@pytest.mark.parametrize("name", CASES)
def test_fp64_one_step_matches_reference(golden_onestep, name):
    c = golden_onestep[name]
    W1, H1 = nbmf_mm_update_beta_dir(c["Y"], c["W"], c["H"], c.get("mask"), float(c["alpha"]), float(c["beta"]))
    assert W1.shape == c["W1"].shape and H1.shape == c["H1"].shape
    assert rel_err(H1, c["H1"]) < 1e-9, name
    assert rel_err(W1, c["W1"]) < 1e-9, name
    # element-wise as well (entries are O(1e-3..1); floor keeps clipped-to-eps entries meaningful)
    assert np.max(np.abs(H1 - c["H1"]) / np.maximum(np.abs(c["H1"]), 1e-6)) < 1e-9
    assert np.max(np.abs(W1 - c["W1"]) / np.maximum(np.abs(c["W1"]), 1e-6)) < 1e-9
    assert np.allclose(W1.sum(axis=0), 1.0, atol=1e-12)


@pytest.mark.parametrize("name", CASES)
def test_fp32_one_step_close(golden_onestep, name):
    c = golden_onestep[name]
    W1, H1 = nbmf_mm_update_beta_dir(c["Y"], c["W"], c["H"], c.get("mask"), float(c["alpha"]), float(c["beta"]),
                                     dtype="float32")
    assert rel_err(H1, c["H1"]) < 2e-5, name
    assert rel_err(W1, c["W1"]) < 2e-5, name
    assert np.allclose(W1.sum(axis=0), 1.0, atol=1e-6)           # simplex bar of north_star


@pytest.mark.parametrize("dtype,tol", [("float64", 1e-11), ("float32", 2e-6)])
@pytest.mark.parametrize("name", CASES)
def test_objective_matches_reference_loss(golden_onestep, name, dtype, tol):
    """The fused NLL reduction (_solver.py:148-162) on the reference's own W1, H1."""
    c = golden_onestep[name]
    mask = c.get("mask")
    data = prepare_data(c["Y"], mask, transpose=False, dtype=dtype, device=None)
    prob = make_problem(data, c["W"].shape[0], dtype=dtype, alpha=float(c["alpha"]), beta=float(c["beta"]), eps=1e-8,
                        mask_semantics="reference", projection="normalize", max_iter_cap=1, device=None)
    with prob:
        prob.set_factors(np.ascontiguousarray(c["W1"].T), c["H1"], normalize_w=False)
        loss = prob.objective()
    assert abs(loss - float(c["loss1"])) <= tol * abs(float(c["loss1"])), (name, loss, float(c["loss1"]))


@pytest.mark.parametrize("k", [1, 2, 3, 4, 5, 7, 8, 9, 12, 13, 16, 17, 20, 21, 24, 25, 31, 32, 33, 40, 48, 49, 64, 65, 80, 96, 97, 128])
def test_every_k_variant_against_oracle(k):
    """All padded-K kernel variants (4..128) in both dtypes, masked, ragged shapes."""
    rng = np.random.default_rng(k)
    m, n = 150 + 3 * k, 1100 + k                  # n crosses the 1024-column pitch, m is not a tile multiple
    Y = (rng.random((m, n)) < 0.2).astype(np.float64)
    mask = (rng.random((m, n)) < 0.9).astype(np.float64)
    W = rng.uniform(0.1, 0.9, (k, m)); W /= W.sum(axis=0, keepdims=True)
    H = rng.uniform(0.05, 0.95, (k, n))
    Wo, Ho = orc.mm_step(Y, W, H, mask, 1.2, 1.3)
    W1, H1 = nbmf_mm_update_beta_dir(Y, W, H, mask, 1.2, 1.3)
    assert rel_err(H1, Ho) < 1e-9 and rel_err(W1, Wo) < 1e-9
    W1, H1 = nbmf_mm_update_beta_dir(Y, W, H, mask, 1.2, 1.3, dtype="float32")
    assert rel_err(H1, Ho) < 5e-5 and rel_err(W1, Wo) < 5e-5


@pytest.mark.parametrize("dtype,tol", [("float64", 1e-9), ("float32", 5e-5)])
def test_strict_mask_semantics_against_oracle(dtype, tol):
    rng = np.random.default_rng(11)
    for prob_x in (False, True):
        Y = rng.random((70, 90)) if prob_x else (rng.random((70, 90)) < 0.3).astype(np.float64)
        mask = (rng.random((70, 90)) < 0.7).astype(np.float64)
        W = rng.uniform(0.1, 0.9, (6, 70)); W /= W.sum(axis=0, keepdims=True)
        H = rng.uniform(0.05, 0.95, (6, 90))
        Wo, Ho = orc.mm_step(Y, W, H, mask, 1.4, 1.1, mask_semantics="strict")
        W1, H1 = nbmf_mm_update_beta_dir(Y, W, H, mask, 1.4, 1.1, mask_semantics="strict", dtype=dtype)
        assert rel_err(H1, Ho) < tol and rel_err(W1, Wo) < tol
        # and it must differ from the reference quirk, otherwise the flag is dead
        _, Hq = orc.mm_step(Y, W, H, mask, 1.4, 1.1)
        assert rel_err(Hq, Ho) > 1e-3


@pytest.mark.parametrize("dtype,tol", [("float64", 1e-9), ("float32", 5e-5)])
def test_duchi_projection_against_oracle(dtype, tol):
    rng = np.random.default_rng(12)
    Y = (rng.random((80, 120)) < 0.25).astype(np.float64)
    for mask in (None, (rng.random((80, 120)) < 0.8).astype(np.float64)):
        W = rng.uniform(0.1, 0.9, (9, 80)); W /= W.sum(axis=0, keepdims=True)
        H = rng.uniform(0.05, 0.95, (9, 120))
        Wo, Ho = orc.mm_step(Y, W, H, mask, 1.2, 1.2, projection="duchi")
        W1, H1 = nbmf_mm_update_beta_dir(Y, W, H, mask, 1.2, 1.2, projection_method="duchi", dtype=dtype)
        assert rel_err(H1, Ho) < tol and rel_err(W1, Wo) < tol
        assert np.all(W1 >= 0) and np.allclose(W1.sum(axis=0), 1.0, atol=1e-6)


def test_empty_and_degenerate_inputs():
    """All-zero, all-one and fully unobserved rows: same values (incl. NaN pattern) as the reference formulas."""
    rng = np.random.default_rng(13)
    Y = (rng.random((40, 50)) < 0.3).astype(np.float64)
    Y[3, :] = 0.0
    Y[5, :] = 1.0
    Y[:, 7] = 0.0
    mask = (rng.random((40, 50)) < 0.8).astype(np.float64)
    W = rng.uniform(0.1, 0.9, (4, 40)); W /= W.sum(axis=0, keepdims=True)
    H = rng.uniform(0.05, 0.95, (4, 50))
    Wo, Ho = orc.mm_step(Y, W, H, mask, 1.2, 1.2)
    W1, H1 = nbmf_mm_update_beta_dir(Y, W, H, mask, 1.2, 1.2)
    assert rel_err(H1, Ho) < 1e-9 and rel_err(W1, Wo) < 1e-9
    mask[9, :] = 0.0                               # a fully unobserved row: G = 0 -> 0/0 in the reference
    with np.errstate(invalid="ignore", divide="ignore"):
        Wo, Ho = orc.mm_step(Y, W, H, mask, 1.2, 1.2)
    W1, H1 = nbmf_mm_update_beta_dir(Y, W, H, mask, 1.2, 1.2)
    assert np.array_equal(np.isnan(W1), np.isnan(Wo)) and np.isnan(W1[:, 9]).all()
    ok = ~np.isnan(Wo)
    assert rel_err(W1[ok], Wo[ok]) < 1e-9 and rel_err(H1, Ho) < 1e-9


@pytest.mark.parametrize("dtype,tol", [("float64", 1e-9), ("float32", 5e-5)])
def test_more_than_64_components(dtype, tol):
    """64 < K <= 128 (the K axis split over 8 lanes; 64-column H-pass tiles and a single-staged W pass in fp64; four
    components per lane in the W epilogue): Duchi projection, strict mask semantics, probabilistic X and a short
    trajectory against the oracle; beyond 128 the C-ABI says so."""
    rng = np.random.default_rng(21)
    m, n, k = 210, 1300, 100
    Y = (rng.random((m, n)) < 0.2).astype(np.float64)
    mask = (rng.random((m, n)) < 0.85).astype(np.float64)
    W = rng.uniform(0.1, 0.9, (k, m)); W /= W.sum(axis=0, keepdims=True)
    H = rng.uniform(0.05, 0.95, (k, n))
    Wo, Ho = orc.mm_step(Y, W, H, mask, 1.2, 1.2, projection="duchi")
    W1, H1 = nbmf_mm_update_beta_dir(Y, W, H, mask, 1.2, 1.2, projection_method="duchi", dtype=dtype)
    assert rel_err(H1, Ho) < tol and rel_err(W1, Wo) < tol
    assert np.all(W1 >= 0) and np.allclose(W1.sum(axis=0), 1.0, atol=1e-6)
    Wo, Ho = orc.mm_step(Y, W, H, mask, 1.4, 1.1, mask_semantics="strict")
    W1, H1 = nbmf_mm_update_beta_dir(Y, W, H, mask, 1.4, 1.1, mask_semantics="strict", dtype=dtype)
    assert rel_err(H1, Ho) < tol and rel_err(W1, Wo) < tol
    Yp = rng.random((90, 140))                                   # probabilistic X: the dense layout, K = 72 and 128
    for kk in (72, 128):
        Wp = rng.uniform(0.1, 0.9, (kk, 90)); Wp /= Wp.sum(axis=0, keepdims=True)
        Hp = rng.uniform(0.05, 0.95, (kk, 140))
        Wo, Ho = orc.mm_step(Yp, Wp, Hp, mask[:90, :140], 1.2, 1.3)
        W1, H1 = nbmf_mm_update_beta_dir(Yp, Wp, Hp, mask[:90, :140], 1.2, 1.3, dtype=dtype)
        assert rel_err(H1, Ho) < tol and rel_err(W1, Wo) < tol
    from nbmf_mm_b200 import nbmf_mm_solver
    for orientation in ("beta-dir", "dir-beta"):
        want = orc.fit(Y, 90, max_iter=12, tol=0.0, mask=mask, random_state=3, orientation=orientation)
        got = nbmf_mm_solver(Y, 90, max_iter=12, tol=0.0, mask=mask, random_state=3, orientation=orientation, dtype=dtype)
        assert got[4] == want[3] == 12
        assert rel_err(got[2], want[2]) < (1e-9 if dtype == "float64" else 1e-4)
        if dtype == "float64":
            assert rel_err(got[0], want[0]) < 1e-8 and rel_err(got[1], want[1]) < 1e-8
    with pytest.raises(RuntimeError, match="1..128"):
        nbmf_mm_update_beta_dir(Y, np.full((129, m), 1 / 129), rng.uniform(0.1, 0.9, (129, n)), mask, 1.2, 1.2, dtype=dtype)
