"""The oracle (oracle/nbmf_oracle.py) against the committed golden vectors, which were produced
by the REAL reference in the authoring container (oracle/make_golden.py).  This is what pins
the oracle: every later GPU parity test trusts it."""
import numpy as np
import pytest

import nbmf_oracle as orc
from conftest import cfg1_matrix


def test_one_step_cases_bit_exact(golden_onestep):
    assert len(golden_onestep) >= 8
    for name, c in golden_onestep.items():
        mask = c.get("mask")
        W1, H1 = orc.mm_step(c["Y"], c["W"], c["H"], mask, float(c["alpha"]), float(c["beta"]))
        assert np.array_equal(H1, c["H1"]), name
        assert np.array_equal(W1, c["W1"]), name
        loss = orc.map_objective(c["Y"], W1, H1, mask, float(c["alpha"]), float(c["beta"]))
        assert loss == float(c["loss1"]), name


def test_cfg1_trajectory_bit_exact(golden_traj):
    g = golden_traj["cfg1"]
    W, H, losses, n_iter = orc.fit(cfg1_matrix(), 6, max_iter=2000, tol=1e-5, alpha=1.2, beta=1.2, random_state=0)
    assert n_iter == int(g["n_iter"]) == 286
    assert np.array_equal(np.asarray(losses), g["losses"])
    assert np.array_equal(W, g["W"]) and np.array_equal(H, g["H"])
    # regression anchors recorded in SURVEY.md section 8c
    assert abs(losses[0] - 0.583919790545) < 1e-11 and abs(losses[-1] - 0.537433210001) < 1e-11


@pytest.mark.parametrize("name,prefix", [("animals", None), ("paleo", 40), ("lastfm", 25)])
def test_cfg2_datasets(golden_traj, datasets, name, prefix):
    g = golden_traj[f"cfg2_{name}"]
    max_iter = 500 if prefix is None else prefix
    W, H, losses, n_iter = orc.fit(datasets[name], 10, max_iter=max_iter, tol=1e-5, random_state=0)
    assert np.array_equal(np.asarray(losses), g["losses"][:len(losses)])
    if prefix is None:
        assert n_iter == int(g["n_iter"])
        assert np.array_equal(W, g["W"]) and np.array_equal(H, g["H"])


def test_masked_and_dir_beta(golden_traj, datasets):
    g = golden_traj["cfg2_animals_train"]
    W, H, losses, n_iter = orc.fit(datasets["animals"], 10, max_iter=500, tol=1e-5, random_state=0,
                                   mask=datasets["animals_train_mask"])
    assert n_iter == int(g["n_iter"]) and np.array_equal(np.asarray(losses), g["losses"])
    g = golden_traj["cfg3s"]
    W, H, losses, n_iter = orc.fit(g["X"], 7, max_iter=150, tol=1e-7, random_state=0, mask=g["mask"],
                                   orientation="dir-beta")
    assert np.array_equal(np.asarray(losses), g["losses"])
    assert np.array_equal(W, g["W"]) and np.array_equal(H, g["H"])
    assert np.allclose(H.sum(axis=0), 1.0, atol=1e-10)          # dir-beta: columns of H on the simplex


def test_probabilistic_X(golden_traj):
    g = golden_traj["prob"]
    W, H, losses, n_iter = orc.fit(g["X"], 5, max_iter=120, tol=1e-9, alpha=1.3, beta=1.7, random_state=3)
    assert np.array_equal(np.asarray(losses), g["losses"])


def test_transform_and_score(golden_transform, datasets):
    g = golden_transform
    X = datasets["animals"]
    for tag, mk in (("nomask", None), ("mask", g["mask"])):
        Wt = orc.transform(X, g["components"], mask=mk, W0=g[f"{tag}/W0"])
        assert np.array_equal(Wt, g[f"{tag}/Wt"])
        W_nomask = orc.transform(X, g["components"], None, g[f"{tag}/W0"])
        s = orc.mean_loglik(X, orc.inverse_transform(W_nomask, g["components"]), mk)
        assert s == float(g[f"{tag}/score"])


def test_duchi_projection_properties():
    """Unpinned by the reference: checked against the defining properties of the projection."""
    rng = np.random.default_rng(0)
    U = rng.random((7, 40)) * 3 - 0.5
    Pj = orc.project_columns_duchi(U)
    assert np.all(Pj >= 0) and np.allclose(Pj.sum(axis=0), 1.0, atol=1e-12)
    on = rng.dirichlet(np.ones(7), size=40).T                   # already on the simplex -> fixed point
    assert np.allclose(orc.project_columns_duchi(on), on, atol=1e-12)
    # optimality: no other simplex point is closer (spot check against random simplex points)
    for _ in range(20):
        other = rng.dirichlet(np.ones(7), size=40).T
        assert np.all(((Pj - U) ** 2).sum(0) <= ((other - U) ** 2).sum(0) + 1e-12)


def test_duchi_and_strict_fits_behave():
    rng = np.random.default_rng(2)
    X = (rng.random((40, 60)) < 0.3).astype(float)
    mask = (rng.random(X.shape) < 0.85).astype(float)
    base = orc.fit(X, 5, max_iter=60, tol=0, random_state=0, mask=mask)
    duchi = orc.fit(X, 5, max_iter=60, tol=0, random_state=0, mask=mask, projection="duchi")
    strict = orc.fit(X, 5, max_iter=60, tol=0, random_state=0, mask=mask, mask_semantics="strict")
    assert np.allclose(duchi[0].sum(axis=1), 1.0, atol=1e-10)
    assert abs(duchi[2][-1] - base[2][-1]) / base[2][-1] < 1e-3   # README: "near-identical"
    ls = np.asarray(strict[2])
    assert np.all(np.diff(ls) <= 1e-12)                           # strict semantics is a true MM: monotone


def test_restart_schedule():
    rng = np.random.default_rng(4)
    X = (rng.random((30, 25)) < 0.3).astype(float)
    best = orc.fit_restarts(X, 4, n_init=3, random_state=5, max_iter=40, tol=0)
    finals = [orc.fit(X, 4, random_state=5 + r, max_iter=40, tol=0)[2][-1] for r in range(3)]
    assert best[2][-1] == min(finals)


def test_weighted_mask_train_splits_and_k40(golden_traj, datasets):
    """Round-2 golden cases from the real reference: weighted (non-0/1) masks in both orientations (the reference
    multiplies by the mask values, _solver.py:30-32), the seeded train masks on lastfm / paleo (a prefix of the 500
    iterations), K = 40."""
    g = golden_traj["wmask"]
    W, H, losses, n_iter = orc.fit(g["X"], 6, max_iter=80, tol=1e-8, alpha=1.2, beta=1.4, random_state=2, mask=g["mask"])
    assert n_iter == int(g["n_iter"]) and np.array_equal(np.asarray(losses), g["losses"]) and np.array_equal(W, g["W"])
    gd = golden_traj["wmask_dirbeta"]
    W, H, losses, n_iter = orc.fit(g["X"], 6, max_iter=60, tol=1e-8, alpha=1.2, beta=1.4, random_state=2, mask=g["mask"],
                                   orientation="dir-beta")
    assert np.array_equal(np.asarray(losses), gd["losses"]) and np.array_equal(H, gd["H"])
    unbits = lambda a, n: np.unpackbits(a, axis=1, bitorder="little")[:, :n].astype(np.float64)
    for name, prefix in (("lastfm", 12), ("paleo", 20)):
        gt = golden_traj[f"cfg2_{name}_train"]
        mask = unbits(gt["mask_bits"], datasets[name].shape[1])
        _, _, losses, _ = orc.fit(datasets[name], 10, max_iter=prefix, tol=1e-5, random_state=0, mask=mask)
        assert np.array_equal(np.asarray(losses), gt["losses"][:prefix])
    gk = golden_traj["k40"]
    _, _, losses, _ = orc.fit(unbits(gk["X_bits"], 600), 40, max_iter=6, tol=0.0, random_state=1, mask=unbits(gk["mask_bits"], 600))
    assert np.array_equal(np.asarray(losses), gk["losses"][:6])


def test_more_than_64_components_against_the_reference(golden_onestep, golden_traj):
    """Golden cases from the real reference for the 64 < K <= 128 kernels: one step at K = 100, 30 iterations at K = 70."""
    c = golden_onestep["bin_mask_k100"]
    Wo, Ho = orc.mm_step(c["Y"], c["W"], c["H"], c["mask"], float(c["alpha"]), float(c["beta"]))
    assert c["W"].shape[0] == 100 and np.array_equal(Wo, c["W1"]) and np.array_equal(Ho, c["H1"])
    unbits = lambda a, n: np.unpackbits(a, axis=1, bitorder="little")[:, :n].astype(np.float64)
    g = golden_traj["k70"]
    W, H, losses, n_iter = orc.fit(unbits(g["X_bits"], 100), 70, max_iter=30, tol=0.0, alpha=1.1, beta=1.3, random_state=4,
                                   mask=unbits(g["mask_bits"], 100))
    assert n_iter == int(g["n_iter"]) == 30 and np.array_equal(np.asarray(losses), g["losses"])
    assert np.array_equal(W, g["W"]) and np.array_equal(H, g["H"])
