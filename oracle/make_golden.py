#!/usr/bin/env python
"""Generate ``tests/golden/*.npz`` from the REAL reference (authoring container only).

Run:  ``python oracle/make_golden.py``  (needs ``/root/reference``; read-only use).

For every case it (1) runs the unmodified reference (``/root/reference/src``),
(2) runs the oracle restatement (``oracle/nbmf_oracle.py``) on the same inputs
and asserts agreement, (3) stores inputs + reference outputs as small fixtures.
The fixtures travel to the GPU box; ``/root/reference`` does not.
"""
from __future__ import annotations

import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
REF = Path("/root/reference")
sys.dont_write_bytecode = True
sys.path.insert(0, str(REF / "src"))
sys.path.insert(0, str(ROOT / "oracle"))

import nbmf_oracle as orc                       # noqa: E402
import importlib.util as _ilu                   # noqa: E402
_spec = _ilu.spec_from_file_location("nbmf_datasets", ROOT / "nbmf_mm_b200" / "datasets.py")   # the product's .rda reader,
_ds = _ilu.module_from_spec(_spec)              # loaded by path: ROOT must not enter sys.path here (its nbmf_mm/ shim
_spec.loader.exec_module(_ds)                   # would shadow the real reference package imported below)
read_rda_matrix = _ds.read_rda_matrix
from nbmf_mm import NBMF                        # noqa: E402  (the reference)
from nbmf_mm._solver import nbmf_mm_solver, nbmf_mm_update_beta_dir   # noqa: E402

OUT = ROOT / "tests" / "golden"
OUT.mkdir(parents=True, exist_ok=True)


def pack(a):
    a = np.asarray(a)
    return np.packbits(a.astype(bool), axis=1, bitorder="little"), np.array(a.shape)


def check(name, a, b, exact=True):
    a, b = np.asarray(a), np.asarray(b)
    d = float(np.max(np.abs(a - b))) if a.size else 0.0
    if exact:
        assert d == 0.0, f"{name}: oracle != reference (max abs diff {d})"
    else:
        assert d <= 1e-14 * max(1.0, float(np.max(np.abs(b)))), f"{name}: diff {d}"
    return d


# --------------------------------------------------------------------------- datasets
def datasets():
    out = {}
    for name in ("animals", "lastfm", "paleo"):
        tag, M = read_rda_matrix(REF / "data" / f"{name}.rda")
        assert tag == name and set(np.unique(M)) <= {0.0, 1.0}
        out[f"{name}_bits"], out[f"{name}_shape"] = pack(M)
    sp = np.load(REF / "data" / "magron2022" / "animals_split.npz")
    for key in ("train_mask", "val_mask", "test_mask"):
        out[f"animals_{key}_bits"], _ = pack(sp[key])
    np.savez_compressed(OUT / "datasets.npz", **out)
    return {n: np.unpackbits(out[f"{n}_bits"], axis=1, bitorder="little")[:, :out[f"{n}_shape"][1]].astype(np.float64)
            for n in ("animals", "lastfm", "paleo")}, sp["train_mask"].astype(np.float64)


# --------------------------------------------------------------------------- one-step cases
def onestep_cases():
    cases = {}
    specs = [
        # name, m, n, k, density (None -> probabilistic X), mask_frac, alpha, beta
        ("bin_nomask", 37, 53, 6, 0.25, None, 1.2, 1.2),
        ("bin_mask", 64, 45, 8, 0.3, 0.8, 1.1, 1.3),
        ("bin_mask_k32", 96, 130, 32, 0.1, 0.9, 1.2, 1.2),
        ("prob_nomask", 29, 41, 5, None, None, 1.5, 1.2),
        ("prob_mask", 40, 33, 4, None, 0.7, 2.0, 0.5),
        ("bin_alpha_lt1", 50, 30, 5, 0.4, None, 0.5, 2.0),
        ("bin_k1", 17, 19, 1, 0.5, 0.9, 1.2, 1.2),
        ("bin_wide", 5, 300, 3, 0.2, 0.85, 1.2, 1.2),
        ("bin_mask_k100", 48, 130, 100, 0.15, 0.9, 1.2, 1.3),      # 64 < K <= 128: the 8-lane K split of the CUDA-core engine
    ]
    for idx, (name, m, n, k, dens, mfrac, a, b) in enumerate(specs):
        rng = np.random.default_rng(1000 + idx)
        Y = (rng.random((m, n)) < dens).astype(np.float64) if dens is not None else rng.random((m, n))
        mask = (rng.random((m, n)) < mfrac).astype(np.float64) if mfrac is not None else None
        W = rng.uniform(0.1, 0.9, (k, m)); W = W / W.sum(axis=0, keepdims=True)
        H = rng.uniform(0.05, 0.95, (k, n))
        Wr, Hr = nbmf_mm_update_beta_dir(Y, W, H, mask, a, b)
        Wo, Ho = orc.mm_step(Y, W, H, mask, a, b)
        check(name + ".H", Ho, Hr); check(name + ".W", Wo, Wr)
        # the reference computes the loss inline in the solver; take it from a 1-iteration solve
        _, _, losses, _, _ = nbmf_mm_solver(Y, k, max_iter=1, alpha=a, beta=b, W_init=W.T, H_init=H, mask=mask)
        lo = orc.map_objective(Y, Wo, Ho, mask, a, b)
        check(name + ".loss", lo, losses[0])
        d = dict(Y=Y, W=W, H=H, W1=Wr, H1=Hr, loss1=np.float64(losses[0]), alpha=a, beta=b)
        if mask is not None:
            d["mask"] = mask
        for key, val in d.items():
            cases[f"{name}/{key}"] = val
    np.savez_compressed(OUT / "onestep.npz", **cases)


# --------------------------------------------------------------------------- trajectories
def trajectories(data, animals_train):
    out = {}

    def run(name, X, k, **kw):
        mask = kw.pop("mask", None)
        orientation = kw.get("orientation", "beta-dir")
        est = NBMF(n_components=k, **kw).fit(X, mask=mask)
        solver_kw = {p: kw[p] for p in ("max_iter", "tol", "alpha", "beta", "random_state") if p in kw}
        solver_kw.setdefault("max_iter", 2000); solver_kw.setdefault("tol", 1e-5)
        Wo, Ho, lo, nit = orc.fit(X, k, mask=mask, orientation=orientation, **solver_kw)
        assert nit == est.n_iter_, (name, nit, est.n_iter_)
        check(name + ".losses", lo, est.loss_curve_)
        check(name + ".W", Wo, est.W_); check(name + ".H", Ho, est.components_)
        out[f"{name}/losses"] = np.asarray(est.loss_curve_)
        out[f"{name}/W"] = est.W_; out[f"{name}/H"] = est.components_
        out[f"{name}/n_iter"] = np.int64(est.n_iter_)
        print(f"{name}: n_iter={est.n_iter_} loss0={est.loss_curve_[0]:.12f} final={est.loss_curve_[-1]:.12f}")

    # config 1 (README quick start)
    X1 = (np.random.default_rng(0).random((100, 500)) < 0.25).astype(float)
    run("cfg1", X1, 6, orientation="beta-dir", alpha=1.2, beta=1.2, random_state=0)
    # config 2 (paper datasets), unmasked and (animals) with the shipped train mask
    for name in ("animals", "lastfm", "paleo"):
        run(f"cfg2_{name}", data[name], 10, orientation="beta-dir", max_iter=500, tol=1e-5, random_state=0)
    run("cfg2_animals_train", data["animals"], 10, orientation="beta-dir", max_iter=500, tol=1e-5,
        random_state=0, mask=animals_train)
    # dir-beta + mask, small (shape of config 3, scaled down)
    rng = np.random.default_rng(3)
    X3 = (rng.random((120, 70)) < 0.15).astype(float)
    M3 = (rng.random((120, 70)) < 0.9).astype(float)
    out["cfg3s/X"], out["cfg3s/mask"] = X3, M3
    run("cfg3s", X3, 7, orientation="dir-beta", max_iter=150, tol=1e-7, alpha=1.2, beta=1.2, random_state=0, mask=M3)
    # config 2 with train masks for the two data sets the reference ships no split for (seeded 70 / 15 / 15 partition of
    # the product's datasets.make_split; the masks are stored bit-packed so that the tests use exactly these)
    for name in ("lastfm", "paleo"):
        tr, _, _ = _ds.make_split(data[name].shape, seed=12345)
        out[f"cfg2_{name}_train/mask_bits"] = np.packbits(tr.astype(bool), axis=1, bitorder="little")
        run(f"cfg2_{name}_train", data[name], 10, orientation="beta-dir", max_iter=500, tol=1e-5, random_state=0, mask=tr)
    # weighted (non-0/1) mask: the reference multiplies by the mask values (_solver.py:30-32)
    rng = np.random.default_rng(8)
    Xw = (rng.random((90, 110)) < 0.25).astype(float)
    Mw = rng.choice([0.0, 0.25, 0.5, 1.0, 1.0], size=Xw.shape)
    out["wmask/X"], out["wmask/mask"] = Xw, Mw
    run("wmask", Xw, 6, orientation="beta-dir", max_iter=80, tol=1e-8, alpha=1.2, beta=1.4, random_state=2, mask=Mw)
    run("wmask_dirbeta", Xw, 6, orientation="dir-beta", max_iter=60, tol=1e-8, alpha=1.2, beta=1.4, random_state=2, mask=Mw)
    # K = 40: the 32 < K <= 64 kernels in float32 mode
    rng = np.random.default_rng(9)
    Xk = (rng.random((700, 600)) < 0.1).astype(float)
    Mk = (rng.random(Xk.shape) < 0.9).astype(float)
    out["k40/X_bits"] = np.packbits(Xk.astype(bool), axis=1, bitorder="little")
    out["k40/mask_bits"] = np.packbits(Mk.astype(bool), axis=1, bitorder="little")
    run("k40", Xk, 40, orientation="beta-dir", max_iter=40, tol=0.0, random_state=1, mask=Mk)
    # K = 70: the 64 < K <= 128 kernels (K padded to 96), both dtypes on the CUDA-core engine
    rng = np.random.default_rng(10)
    X70 = (rng.random((120, 100)) < 0.12).astype(float)
    M70 = (rng.random(X70.shape) < 0.9).astype(float)
    out["k70/X_bits"] = np.packbits(X70.astype(bool), axis=1, bitorder="little")
    out["k70/mask_bits"] = np.packbits(M70.astype(bool), axis=1, bitorder="little")
    run("k70", X70, 70, orientation="beta-dir", max_iter=30, tol=0.0, alpha=1.1, beta=1.3, random_state=4, mask=M70)
    # probabilistic X through the estimator
    Xp = np.random.default_rng(5).random((45, 60))
    out["prob/X"] = Xp
    run("prob", Xp, 5, orientation="beta-dir", max_iter=120, tol=1e-9, alpha=1.3, beta=1.7, random_state=3)
    np.savez_compressed(OUT / "trajectories.npz", **out)


# --------------------------------------------------------------------------- transform / score
def transform_cases(data):
    out = {}
    X = data["animals"]
    est = NBMF(n_components=4, max_iter=80, tol=0.0, random_state=1).fit(X)
    rng = np.random.default_rng(7)
    mask = (rng.random(X.shape) < 0.8).astype(float)
    for tag, mk in (("nomask", None), ("mask", mask)):
        np.random.seed(99)
        Wt_ref = est.transform(X, mask=mk)
        W0 = np.random.RandomState(99).uniform(0.1, 0.9, (X.shape[0], 4))
        Wt_orc = orc.transform(X, est.components_, mask=mk, W0=W0)
        check(f"transform.{tag}", Wt_orc, Wt_ref)
        out[f"{tag}/W0"], out[f"{tag}/Wt"] = W0, Wt_ref
        np.random.seed(99)
        s_ref = est.score(X, mask=mk)                 # transform is called WITHOUT the mask (_base.py:235)
        s_orc = orc.mean_loglik(X, orc.inverse_transform(orc.transform(X, est.components_, None, W0), est.components_), mk)
        check(f"score.{tag}", s_orc, s_ref)
        out[f"{tag}/score"] = np.float64(s_ref)
    out["components"], out["mask"] = est.components_, mask
    np.savez_compressed(OUT / "transform.npz", **out)


if __name__ == "__main__":
    os.environ.setdefault("OMP_NUM_THREADS", "8")
    data, train = datasets()
    onestep_cases()
    trajectories(data, train)
    transform_cases(data)
    print("golden fixtures written to", OUT)
    for p in sorted(OUT.glob("*.npz")):
        print(f"  {p.name}: {p.stat().st_size/1024:.1f} KiB")
