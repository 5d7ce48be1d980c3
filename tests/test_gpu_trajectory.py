"""Objective-trajectory parity of the device-resident fit loop against the reference's own
trajectories (golden vectors from the real nbmf_mm_solver, _solver.py:61-216).

FP64: every recorded loss within 1e-9 relative, identical n_iter (so the device-side stop rule
fires at the same iteration), final factors within 1e-7.  FP32: final NLL within 1e-4 relative
after the same iteration count, simplex to 1e-6, objective monotone (slack 5e-7 relative:
fp32 evaluation noise of a 1e4..1e6-term sum)."""
import numpy as np
import pytest

from conftest import cfg1_matrix, rel_err
from nbmf_mm_b200 import NBMF, nbmf_mm_solver

pytestmark = pytest.mark.gpu


def _bits(a, n):
    return np.unpackbits(a, axis=1, bitorder="little")[:, :n].astype(np.float64)


def _cases(datasets, golden_traj):
    g = golden_traj
    return {
        "cfg2_lastfm_train": (datasets["lastfm"], _bits(g["cfg2_lastfm_train"]["mask_bits"], datasets["lastfm"].shape[1]),
                              dict(n_components=10, max_iter=500, tol=1e-5, random_state=0)),
        "cfg2_paleo_train": (datasets["paleo"], _bits(g["cfg2_paleo_train"]["mask_bits"], datasets["paleo"].shape[1]),
                             dict(n_components=10, max_iter=500, tol=1e-5, random_state=0)),
        "wmask": (g["wmask"]["X"], g["wmask"]["mask"],
                  dict(n_components=6, max_iter=80, tol=1e-8, alpha=1.2, beta=1.4, random_state=2)),
        "wmask_dirbeta": (g["wmask"]["X"], g["wmask"]["mask"],
                          dict(n_components=6, orientation="dir-beta", max_iter=60, tol=1e-8, alpha=1.2, beta=1.4, random_state=2)),
        "k40": (_bits(g["k40"]["X_bits"], 600), _bits(g["k40"]["mask_bits"], 600),
                dict(n_components=40, max_iter=40, tol=0.0, random_state=1)),
        "k70": (_bits(g["k70"]["X_bits"], 100), _bits(g["k70"]["mask_bits"], 100),
                dict(n_components=70, max_iter=30, tol=0.0, alpha=1.1, beta=1.3, random_state=4)),
        "cfg1": (cfg1_matrix(), None, dict(n_components=6, alpha=1.2, beta=1.2, random_state=0)),
        "cfg2_animals": (datasets["animals"], None, dict(n_components=10, max_iter=500, tol=1e-5, random_state=0)),
        "cfg2_lastfm": (datasets["lastfm"], None, dict(n_components=10, max_iter=500, tol=1e-5, random_state=0)),
        "cfg2_paleo": (datasets["paleo"], None, dict(n_components=10, max_iter=500, tol=1e-5, random_state=0)),
        "cfg2_animals_train": (datasets["animals"], datasets["animals_train_mask"],
                               dict(n_components=10, max_iter=500, tol=1e-5, random_state=0)),
        "cfg3s": (golden_traj["cfg3s"]["X"], golden_traj["cfg3s"]["mask"],
                  dict(n_components=7, orientation="dir-beta", max_iter=150, tol=1e-7, alpha=1.2, beta=1.2, random_state=0)),
        "prob": (golden_traj["prob"]["X"], None,
                 dict(n_components=5, max_iter=120, tol=1e-9, alpha=1.3, beta=1.7, random_state=3)),
    }


NAMES = ["cfg1", "cfg2_animals", "cfg2_lastfm", "cfg2_paleo", "cfg2_animals_train", "cfg3s", "prob",
         "cfg2_lastfm_train", "cfg2_paleo_train", "wmask", "wmask_dirbeta", "k40", "k70"]


@pytest.mark.parametrize("name", NAMES)
def test_fp64_trajectory(datasets, golden_traj, name):
    X, mask, kw = _cases(datasets, golden_traj)[name]
    g = golden_traj[name]
    est = NBMF(**kw).fit(X, mask=mask)
    assert est.n_iter_ == int(g["n_iter"]) == len(est.loss_curve_)
    ours, ref = np.asarray(est.loss_curve_), g["losses"]
    assert np.max(np.abs(ours - ref) / np.abs(ref)) < 1e-9
    assert rel_err(est.W_, g["W"]) < 1e-7 and rel_err(est.components_, g["H"]) < 1e-7
    assert est.W_.shape == g["W"].shape and est.components_.shape == g["H"].shape
    assert isinstance(est.reconstruction_err_, float) and est.loss_ == est.loss_curve_[-1]
    assert est.objective_history_ is est.loss_curve_


@pytest.mark.parametrize("name", NAMES)
def test_fp32_final_nll_simplex_monotone(datasets, golden_traj, name):
    X, mask, kw = _cases(datasets, golden_traj)[name]
    g = golden_traj[name]
    kw = dict(kw, max_iter=int(g["n_iter"]), tol=0.0, dtype="float32")       # same iteration count
    est = NBMF(**kw).fit(X, mask=mask)
    assert est.n_iter_ == int(g["n_iter"])
    ours, ref = np.asarray(est.loss_curve_), g["losses"]
    assert abs(ours[-1] - ref[-1]) / abs(ref[-1]) < 1e-4
    assert np.max(np.abs(ours - ref) / np.abs(ref)) < 1e-4                    # the whole curve, in fact
    simplex = est.components_.sum(axis=0) if est.orientation == "dir-beta" else est.W_.sum(axis=1)
    assert np.max(np.abs(simplex - 1.0)) < 1e-6
    assert np.all(np.diff(ours) <= 5e-7 * np.abs(ours[:-1]))


def test_stop_rule_semantics():
    """n_iter, the kept factors and the history length follow _solver.py:169-175,215."""
    X = cfg1_matrix()
    W, H, losses, t, n_iter = nbmf_mm_solver(X, 6, max_iter=2000, tol=1e-3, random_state=0)
    assert t == 0.0 and isinstance(n_iter, int) and len(losses) == n_iter and isinstance(losses[0], np.float64)
    rel = [abs(a - b) / abs(a) for a, b in zip(losses, losses[1:])]
    assert rel[-1] < 1e-3 and all(r >= 1e-3 for r in rel[:-1])
    # max_iter reached: exactly max_iter losses, no early stop
    _, _, losses2, _, n2 = nbmf_mm_solver(X, 6, max_iter=7, tol=0.0, random_state=0)
    assert n2 == 7 and len(losses2) == 7 and np.allclose(losses2, losses[:7], rtol=1e-12)
    # one iteration only
    _, _, losses3, _, n3 = nbmf_mm_solver(X, 6, max_iter=1, tol=1e-5, random_state=0)
    assert n3 == 1 and abs(losses3[0] - losses[0]) < 1e-12
    with pytest.raises(UnboundLocalError):
        nbmf_mm_solver(X, 6, max_iter=0)


def test_global_rng_side_effect_and_init_stream():
    """random_state reseeds the GLOBAL legacy RNG and draws W (m x k) then H (k x n) with the
    internal (post-orientation) m, n -- _solver.py:102-103,126-129."""
    X = (np.random.default_rng(3).random((25, 30)) < 0.3).astype(float)
    nbmf_mm_solver(X, 5, max_iter=2, random_state=9, orientation="dir-beta")
    after = np.random.uniform()
    rs = np.random.RandomState(9)
    rs.uniform(0.1, 0.9, (30, 5)); rs.uniform(0.1, 0.9, (5, 25))
    assert after == rs.uniform()
    # explicit inits equal to that stream reproduce the seeded run
    rs = np.random.RandomState(9)
    W0 = rs.uniform(0.1, 0.9, (30, 5)); H0 = rs.uniform(0.1, 0.9, (5, 25))
    a = nbmf_mm_solver(X, 5, max_iter=5, random_state=9, orientation="dir-beta")
    b = nbmf_mm_solver(X, 5, max_iter=5, orientation="dir-beta", W_init=H0.T, H_init=W0.T)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_verbose_lines(capsys):
    X = cfg1_matrix()
    nbmf_mm_solver(X, 6, max_iter=25, tol=0.0, random_state=0, verbose=1)
    out = capsys.readouterr().out.strip().splitlines()
    assert out[0].startswith("Iter    0: Loss = 0.583920") and len(out) == 3
