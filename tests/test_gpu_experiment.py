"""The reference's experiment driver as batched multi-fits with held-out evaluation on the device (SURVEY.md section 8
f1 + f2; ``examples/reproduce_magron2022.py:40-47,74-152,242-329``).  The checker is the oracle's restatement of the
driver's ``compute_perplexity`` applied to the factors the call returns, and the corresponding solver calls."""
import numpy as np
import pytest

from nbmf_mm_b200 import experiment, nbmf_mm_multifit, nbmf_mm_solver
from nbmf_mm_b200.datasets import make_split
from nbmf_oracle import heldout_perplexity

pytestmark = pytest.mark.gpu


def _data(m=226, n=285, seed=0):
    rng = np.random.default_rng(seed)
    X = (rng.random((m, n)) < 0.08).astype(np.float64)
    train, val, test = make_split(X.shape, seed=seed + 1)
    return X, train, val, test


@pytest.mark.parametrize("orientation,dtype,tol", [("beta-dir", "float64", 1e-11), ("dir-beta", "float64", 1e-11),
                                                     ("beta-dir", "float32", 1e-5)])
def test_held_out_perplexities_of_a_multifit_equal_the_driver_formula(orientation, dtype, tol):
    """Batched group (same K: an alpha / beta grid) and single jobs (other K) in one call; every job's sixth element is
    compute_perplexity(Y, W @ H, mask) of the factors that job returns, for every evaluation mask."""
    X, train, val, test = _data()
    jobs = [dict(n_components=6, alpha=a, beta=b, random_state=3) for a, b in [(0.5, 1.0), (1.2, 1.2), (2.0, 3.0), (1.0, 1.0)]]
    jobs += [dict(n_components=9, alpha=1.5, beta=1.5, random_state=4), dict(n_components=3, random_state=5)]
    stats = {}
    got = nbmf_mm_multifit(X, jobs, mask=train, orientation=orientation, max_iter=50, tol=1e-6, dtype=dtype, stats=stats,
                           eval_masks={"train": train, "val": val, "test": test})
    assert stats["batched"] == 4 and len(got) == len(jobs)
    for out in got:
        assert len(out) == 6
        W, H, ho = out[0], out[1], out[5]
        assert set(ho) == {"train", "val", "test"}
        for name, mk in (("train", train), ("val", val), ("test", test)):
            want = heldout_perplexity(X, W @ H, mk)
            assert abs(ho[name]["perplexity"] - want) <= tol * want, (name, ho[name]["perplexity"], want)
            assert ho[name]["n_entries"] == np.count_nonzero(mk)
            assert abs(ho[name]["nll"] - np.log(want)) <= tol * max(1.0, abs(np.log(want)))
    # without eval_masks the return value is the solver's five-tuple, and the fits are the same fits
    plain = nbmf_mm_multifit(X, jobs, mask=train, orientation=orientation, max_iter=50, tol=1e-6, dtype=dtype)
    for a, b in zip(plain, got):
        assert len(a) == 5 and a[4] == b[4] and np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_grid_search_is_the_loop_of_figure_1():
    """Every record of the grid is what the driver's loop body computes for that point: the solver call of
    train_nbmf_mm (beta-dir, random_state 12345) followed by compute_perplexity on the train and validation masks; the
    best record is the first arg-min of the validation perplexity; records are in the driver's order."""
    X, train, val, _ = _data(120, 150, seed=7)
    alphas, betas = (0.5, 1.5, 3.0), (1.0, 2.0)
    out = experiment.grid_search(X, train, val, n_components=4, alphas=alphas, betas=betas, max_iter=80, tol=1e-5,
                                 engine="simt")
    recs = out["records"]
    assert [(r["alpha"], r["beta"]) for r in recs] == [(a, b) for a in alphas for b in betas]
    for r in recs:
        W, H, losses, _, n_iter = nbmf_mm_solver(X, 4, max_iter=80, tol=1e-5, alpha=r["alpha"], beta=r["beta"], mask=train,
                                                 random_state=12345, orientation="beta-dir", engine="simt")
        assert r["n_iter"] == n_iter and r["final_loss"] == losses[-1] and r["k"] == 4
        for name, mk in (("train", train), ("val", val)):
            want = heldout_perplexity(X, W @ H, mk)
            assert abs(r[f"{name}_perplexity"] - want) <= 1e-11 * want
        assert "test_perplexity" not in r
    vals = [r["val_perplexity"] for r in recs]
    assert out["best"] is recs[int(np.argmin(vals))]


def test_components_sweep_and_final_fit_on_the_paper_data(datasets):
    """Figures 3 and 2 on the animals data with its stored split: one record per K with the three perplexities, and the
    single long fit whose record carries the factors."""
    Y = datasets["animals"]
    train, val, test = (datasets[f"animals_{k}"] for k in ("train_mask", "val_mask", "test_mask"))
    recs = experiment.components_sweep(Y, train, val, test, k_values=(2, 4, 8), alpha=2.0, beta=2.0, max_iter=60, tol=0.0)
    assert [r["k"] for r in recs] == [2, 4, 8]
    for r in recs:
        W, H, _, _, n_iter = nbmf_mm_solver(Y, r["k"], max_iter=60, tol=0.0, alpha=2.0, beta=2.0, mask=train, random_state=12345)
        assert r["n_iter"] == n_iter
        for name, mk in (("train", train), ("val", val), ("test", test)):
            want = heldout_perplexity(Y, W @ H, mk)
            assert abs(r[f"{name}_perplexity"] - want) <= 1e-8 * want      # auto engine: fits equal to rounding
    one = experiment.fit_and_test(Y, train, test, n_components=4, alpha=2.0, beta=2.0, max_iter=100)
    assert one["W"].shape == (Y.shape[0], 4) and one["H"].shape == (4, Y.shape[1])
    want = heldout_perplexity(Y, one["W"] @ one["H"], test)
    assert abs(one["test_perplexity"] - want) <= 1e-11 * want and "val_perplexity" not in one


def test_eval_mask_errors():
    X, train, val, _ = _data(40, 64, seed=2)
    with pytest.raises(ValueError, match="None"):
        nbmf_mm_multifit(X, [dict(n_components=3, random_state=0)], mask=train, max_iter=3, eval_masks={"val": None})
    with pytest.raises(ValueError):
        nbmf_mm_multifit(X, [dict(n_components=3, random_state=0)], mask=train, max_iter=3, eval_masks={"val": val[:, :10]})
