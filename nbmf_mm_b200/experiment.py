"""The reference's experiment driver as batched multi-fits.

``examples/reproduce_magron2022.py`` runs three loops of ``NBMF(...).fit(Y, mask=train_mask)`` calls, each followed by a
dense ``W @ H`` and ``compute_perplexity`` on the train / validation / test masks (``:40-47``):

* figure 1 (``:74-152``): a 6 x 6 grid of (alpha, beta) at a fixed K per data set, validation perplexity of every point,
  arg-min;
* figure 2 (``:154-240``): one fit with the chosen hyper-parameters and ``max_iter=1000``, test perplexity;
* figure 3 (``:242-329``): a sweep over K at the chosen (alpha, beta), train / validation / test perplexity.

Here every loop is ONE ``nbmf_mm_multifit`` call: the data planes and the planes of every evaluation mask are packed and
uploaded once, the points of a grid advance together (one launch per kernel for the whole grid: the Beta prior lives in
each fit's device-side state), and every fit is scored on its held-out masks on the device, from the factors where they
are.  Each record is what the reference's loop computes for that point -- the fit is the solver call of
``train_nbmf_mm`` (``:49-72``: beta-dir, ``random_state=12345``; bit for bit on the same engine), the perplexities agree
with ``compute_perplexity`` to rounding (tests/test_gpu_experiment.py).  Plotting, pickling and the comparison against the stored results of the paper's
authors (``outputs/magron2022``) are not part of the path and stay with the caller."""
from __future__ import annotations

import time

import numpy as np

from .multifit import nbmf_mm_multifit

# the driver's tables (examples/reproduce_magron2022.py:86-94,167-171,254-265)
ALPHA_VALUES = (0.5, 1.0, 1.5, 2.0, 2.5, 3.0)
BETA_VALUES = (0.5, 1.0, 1.5, 2.0, 2.5, 3.0)
VALIDATION_K = {"animals": 4, "lastfm": 8, "paleo": 4}
BEST_PARAMS = {"animals": (2.0, 2.0), "lastfm": (1.0, 1.0), "paleo": (2.0, 2.0)}
K_RANGE = (2, 4, 8, 16)
SEED = 12345


def _masks(train_mask, val_mask, test_mask):
    em = {"train": train_mask}
    if val_mask is not None:
        em["val"] = val_mask
    if test_mask is not None:
        em["test"] = test_mask
    return em


def _record(job, res, seconds):
    rec = {key: job[key] for key in ("alpha", "beta") if key in job}
    rec["k"] = int(job["n_components"])
    for name, ho in res[5].items():
        rec[f"{name}_perplexity"] = ho["perplexity"]
    rec["n_iter"] = int(res[4])
    rec["final_loss"] = float(res[2][-1]) if len(res[2]) else float("nan")
    rec["time"] = seconds
    return rec


def evaluate_jobs(Y, train_mask, jobs, *, val_mask=None, test_mask=None, max_iter=500, tol=1e-5, random_state=SEED,
                  orientation="beta-dir", dtype="float64", device=None, engine="auto", return_factors=False, **multifit_kw):
    """Fit every job of ``jobs`` (dicts with ``n_components``, ``alpha``, ``beta``; ``random_state`` defaults to the
    driver's 12345) on the entries of ``train_mask`` and score it on the train / validation / test masks.  Returns one
    record per job: ``alpha, beta, k, train_perplexity[, val_perplexity][, test_perplexity], n_iter, final_loss, time``
    (``time`` = wall-clock of the whole call divided by the number of jobs: the fits run together)."""
    jobs = [dict(j) for j in jobs]
    for j in jobs:
        j.setdefault("random_state", random_state)
    t0 = time.perf_counter()
    res = nbmf_mm_multifit(Y, jobs, mask=train_mask, orientation=orientation, max_iter=max_iter, tol=tol, dtype=dtype,
                           device=device, engine=engine, eval_masks=_masks(train_mask, val_mask, test_mask), **multifit_kw)
    per = (time.perf_counter() - t0) / max(1, len(jobs))
    recs = [_record(j, r, per) for j, r in zip(jobs, res)]
    if return_factors:
        for rec, r in zip(recs, res):
            rec["W"], rec["H"] = r[0], r[1]
    return recs


def grid_search(Y, train_mask, val_mask, *, n_components, alphas=ALPHA_VALUES, betas=BETA_VALUES, test_mask=None, **kw):
    """Figure 1: validation perplexity over the (alpha, beta) grid at a fixed K.  Returns ``{"records": [...], "best":
    record}`` with the records in the driver's order (alpha outer, beta inner) and ``best`` the first record with the
    lowest validation perplexity (``idxmin``, ``:146-147``)."""
    jobs = [dict(n_components=int(n_components), alpha=float(a), beta=float(b)) for a in alphas for b in betas]
    recs = evaluate_jobs(Y, train_mask, jobs, val_mask=val_mask, test_mask=test_mask, **kw)
    vals = np.array([r["val_perplexity"] for r in recs], dtype=np.float64)
    best = recs[int(np.nanargmin(vals))] if np.isfinite(vals).any() else None
    return {"records": recs, "best": best}


def components_sweep(Y, train_mask, val_mask, test_mask, *, k_values=K_RANGE, alpha, beta, **kw):
    """Figure 3: train / validation / test perplexity for every K of ``k_values`` at fixed (alpha, beta)."""
    jobs = [dict(n_components=int(k), alpha=float(alpha), beta=float(beta)) for k in k_values]
    return evaluate_jobs(Y, train_mask, jobs, val_mask=val_mask, test_mask=test_mask, **kw)


def fit_and_test(Y, train_mask, test_mask, *, n_components, alpha, beta, max_iter=1000, **kw):
    """Figure 2: one fit with the chosen hyper-parameters (``max_iter=1000``, ``:180-186``) and its test perplexity; the
    record also carries the factors."""
    return evaluate_jobs(Y, train_mask, [dict(n_components=int(n_components), alpha=float(alpha), beta=float(beta))],
                         test_mask=test_mask, max_iter=max_iter, return_factors=True, **kw)[0]


def run_magron2022(data_dir="data", datasets=("animals", "lastfm", "paleo"), split_dir=None, figures=(1, 2, 3), **kw):
    """The three loops of the reference's driver over the paper's data sets (``data/<name>.rda`` + splits, read by
    ``nbmf_mm_b200.datasets``).  Returns ``{"figure1": {name: grid_search result}, "figure2": {name: record},
    "figure3": {name: [records]}}`` for the requested figures."""
    from .datasets import load_dataset_and_splits
    out = {f"figure{f}": {} for f in figures}
    for name in datasets:
        Y, train, val, test = load_dataset_and_splits(name, data_dir=data_dir, split_dir=split_dir)
        a, b = BEST_PARAMS.get(name, (1.0, 1.0))
        k = VALIDATION_K.get(name, 4)
        if 1 in figures:
            out["figure1"][name] = grid_search(Y, train, val, n_components=k, **kw)
        if 2 in figures:
            out["figure2"][name] = fit_and_test(Y, train, test, n_components=k, alpha=a, beta=b, **kw)
        if 3 in figures:
            out["figure3"][name] = components_sweep(Y, train, val, test, alpha=a, beta=b, **kw)
    return out
