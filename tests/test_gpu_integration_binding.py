"""INTEGRATION.md section 2, executed: the ctypes binding a maintainer of the reference would add
(``tools/reference_binding_b200.py``, shown verbatim in INTEGRATION.md) replaces the iteration loop of the solver
(``_solver.py:143-175``) while seeding, orientation handling and normalisation stay in "reference" Python -- here the
oracle's restatement of those lines.  The golden reference trajectories must come back."""
import importlib.util
import os
from pathlib import Path

import numpy as np
import pytest

from conftest import cfg1_matrix

ROOT = Path(__file__).resolve().parents[1]
pytestmark = pytest.mark.gpu


def _binding():
    os.environ.setdefault("NBMF_B200_LIB", str(ROOT / "nbmf_mm_b200" / "libnbmf_b200.so"))
    spec = importlib.util.spec_from_file_location("reference_binding_b200", ROOT / "tools" / "reference_binding_b200.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_integration_md_shows_the_tested_file():
    text = (ROOT / "INTEGRATION.md").read_text()
    code = (ROOT / "tools" / "reference_binding_b200.py").read_text()
    body = code[code.index("# src/nbmf_mm/_b200.py"):]
    assert body in text


@pytest.mark.parametrize("case", ["cfg1", "cfg2_animals_train", "cfg3s"])
def test_reference_loop_routed_through_the_c_abi(case, datasets, golden_traj):
    b = _binding()
    g = golden_traj[case]
    if case == "cfg1":
        X, mask, kw = cfg1_matrix(), None, dict(k=6, alpha=1.2, beta=1.2, seed=0, orientation="beta-dir", max_iter=2000, tol=1e-5)
    elif case == "cfg2_animals_train":
        X, mask, kw = datasets["animals"], datasets["animals_train_mask"], dict(k=10, alpha=1.2, beta=1.2, seed=0, orientation="beta-dir", max_iter=500, tol=1e-5)
    else:
        X, mask, kw = g["X"], g["mask"], dict(k=7, alpha=1.2, beta=1.2, seed=0, orientation="dir-beta", max_iter=150, tol=1e-7)
    max_iter, tol = kw["max_iter"], kw["tol"]                 # as oracle/make_golden.py ran the reference
    # the lines of nbmf_mm_solver around the loop, as the reference has them (_solver.py:102-136)
    np.random.seed(kw["seed"])
    Y = np.asarray(X, dtype=np.float64)
    if kw["orientation"] == "dir-beta":                       # :113-123
        Y = np.ascontiguousarray(Y.T)
        mask = None if mask is None else np.ascontiguousarray(np.asarray(mask).T)
    m, n = Y.shape
    W_init = np.random.uniform(0.1, 0.9, (m, kw["k"]))        # :126-129
    H_init = np.random.uniform(0.1, 0.9, (kw["k"], n))
    W = W_init.T
    W = W / W.sum(axis=0, keepdims=True)                      # :136
    Wf, Hf, losses, n_iter = b.fit_loop_b200(Y, mask, np.ascontiguousarray(W), H_init, kw["alpha"], kw["beta"], max_iter, tol)
    ref = np.asarray(g["losses"])
    assert n_iter == int(g["n_iter"]) == len(losses)
    assert np.max(np.abs(np.asarray(losses) - ref) / np.abs(ref)) < 1e-9
    assert np.max(np.abs(Wf.sum(axis=0) - 1.0)) < 1e-12
