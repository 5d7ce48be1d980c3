"""The tcgen05/TMEM engine (tc_passes.cuh; float32, bit-packed V, K <= 64, TF32 + bf16 split precision) against the
oracle, the golden reference trajectories and the SIMT engine.  Same FP32-mode bars as the SIMT
kernels: one step <= 5e-5 relative, final NLL <= 1e-4 relative after the same iteration count,
simplex <= 1e-6, monotone objective."""
import numpy as np
import pytest

import nbmf_oracle as orc
from conftest import cfg1_matrix, rel_err
from nbmf_mm_b200 import NBMF, nbmf_mm_update_beta_dir
from nbmf_mm_b200.solver import make_problem, prepare_data

pytestmark = pytest.mark.gpu


def problem(m, n, k, seed, masked=True, density=0.2):
    rng = np.random.default_rng(seed)
    Y = (rng.random((m, n)) < density).astype(np.float64)
    mask = (rng.random((m, n)) < 0.9).astype(np.float64) if masked else None
    W = rng.uniform(0.1, 0.9, (k, m)); W /= W.sum(axis=0, keepdims=True)
    H = rng.uniform(0.05, 0.95, (k, n))
    return Y, mask, W, H


@pytest.mark.parametrize("m,n,k,masked", [
    (128, 128, 32, True),          # exactly one tile each way
    (130, 70, 32, False),          # ragged rows and columns, single partial blocks
    (517, 1300, 32, True),         # several row tiles, columns across the 1024 pitch
    (300, 2111, 20, True),         # K padded to 32
    (1000, 333, 6, False),
    (77, 4100, 1, True),
    (2500, 190, 25, True),
    (300, 700, 8, True),           # K <= 16 instantiation (skips the padded half of the K extent)
    (1111, 300, 16, False),
    (517, 1300, 33, True),         # K <= 64 instantiation (three pipelines, one accumulator set)
    (2200, 400, 48, True),
    (700, 2111, 64, False),
    (130, 70, 64, True),
])
def test_one_step_against_oracle(m, n, k, masked):
    Y, mask, W, H = problem(m, n, k, seed=m + n + k, masked=masked)
    Wo, Ho = orc.mm_step(Y, W, H, mask, 1.2, 1.3)
    W1, H1 = nbmf_mm_update_beta_dir(Y, W, H, mask, 1.2, 1.3, dtype="float32", engine="tensor")
    assert rel_err(H1, Ho) < 5e-5 and rel_err(W1, Wo) < 5e-5
    assert np.max(np.abs(W1.sum(axis=0) - 1.0)) < 1e-6
    Ws, Hs = nbmf_mm_update_beta_dir(Y, W, H, mask, 1.2, 1.3, dtype="float32", engine="simt")
    assert rel_err(H1, Hs) < 2e-5 and rel_err(W1, Ws) < 2e-5


@pytest.mark.parametrize("m,n,k", [(517, 1300, 32), (130, 70, 10), (900, 600, 17), (640, 900, 40)])
def test_strict_mask_semantics_on_the_tensor_engine(m, n, k):
    """README / paper mask semantics (unobserved entries contribute nothing to the H step and the loss): the strict
    variant of the tensor H pass against the oracle, one step, the fused objective and a short monotone fit."""
    Y, mask, W, H = problem(m, n, k, seed=3 * m + n)
    Wo, Ho = orc.mm_step(Y, W, H, mask, 1.3, 1.1, mask_semantics="strict")
    W1, H1 = nbmf_mm_update_beta_dir(Y, W, H, mask, 1.3, 1.1, mask_semantics="strict", dtype="float32", engine="tensor")
    assert rel_err(H1, Ho) < 5e-5 and rel_err(W1, Wo) < 5e-5
    Wr, Hr = orc.mm_step(Y, W, H, mask, 1.3, 1.1)                    # the reference quirk gives a different H
    assert rel_err(Hr, Ho) > 1e-3
    want = orc.map_objective(Y, W, H, mask, 1.3, 1.1, mask_semantics="strict")
    data = prepare_data(Y, mask, transpose=False, dtype="float32", device=None)
    with make_problem(data, k, dtype="float32", alpha=1.3, beta=1.1, eps=1e-8, mask_semantics="strict",
                      projection="normalize", max_iter_cap=30, device=None, engine="tensor") as prob:
        assert prob.engine == "tensor"
        prob.set_factors(np.ascontiguousarray(W.T), H, normalize_w=False)
        got = prob.objective()
        assert abs(got - want) < 5e-6 * abs(want)
        losses, n_iter, _ = prob.fit(30, 0.0)
    assert n_iter == 30 and np.all(np.diff(losses) <= 2e-6 * np.abs(losses[:-1]))   # strict semantics is a true MM


@pytest.mark.parametrize("m,n,k", [(517, 1300, 32), (130, 70, 10), (700, 333, 57)])
def test_fused_objective(m, n, k):
    Y, mask, W, H = problem(m, n, k, seed=5)
    want = orc.map_objective(Y, W, H, mask, 1.2, 1.3)
    data = prepare_data(Y, mask, transpose=False, dtype="float32", device=None)
    with make_problem(data, k, dtype="float32", alpha=1.2, beta=1.3, eps=1e-8, mask_semantics="reference",
                      projection="normalize", max_iter_cap=1, device=None, engine="tensor") as prob:
        assert prob.engine == "tensor"
        prob.set_factors(np.ascontiguousarray(W.T), H, normalize_w=False)
        got = prob.objective()
    assert abs(got - want) < 5e-6 * abs(want)


def test_golden_trajectories_fp32(datasets, golden_traj):
    cases = {
        "cfg1": (cfg1_matrix(), None, dict(n_components=6, alpha=1.2, beta=1.2, random_state=0)),
        "cfg2_lastfm": (datasets["lastfm"], None, dict(n_components=10, random_state=0)),
        "cfg2_animals_train": (datasets["animals"], datasets["animals_train_mask"], dict(n_components=10, random_state=0)),
        "cfg3s": (golden_traj["cfg3s"]["X"], golden_traj["cfg3s"]["mask"],
                  dict(n_components=7, orientation="dir-beta", alpha=1.2, beta=1.2, random_state=0)),
    }
    for name, (X, mask, kw) in cases.items():
        g = golden_traj[name]
        est = NBMF(max_iter=int(g["n_iter"]), tol=0.0, dtype="float32", engine="tensor", **kw).fit(X, mask=mask)
        assert est.transfer_stats_["engine"] == "tensor"
        ours, ref = np.asarray(est.loss_curve_), g["losses"]
        assert abs(ours[-1] - ref[-1]) / abs(ref[-1]) < 1e-4, name
        assert np.max(np.abs(ours - ref) / np.abs(ref)) < 1e-4, name
        simplex = est.components_.sum(axis=0) if est.orientation == "dir-beta" else est.W_.sum(axis=1)
        assert np.max(np.abs(simplex - 1.0)) < 1e-6
        assert np.all(np.diff(ours) <= 2e-6 * np.abs(ours[:-1])), name


def test_stop_rule_and_determinism_on_tensor_engine():
    X = cfg1_matrix()
    a = NBMF(n_components=6, random_state=0, dtype="float32", engine="tensor").fit(X)
    b = NBMF(n_components=6, random_state=0, dtype="float32", engine="tensor").fit(X)
    assert a.n_iter_ == b.n_iter_ and np.array_equal(a.W_, b.W_) and np.array_equal(a.components_, b.components_)
    assert abs(a.n_iter_ - 286) <= 3 and len(a.loss_curve_) == a.n_iter_      # reference stops at 286 in fp64


def test_auto_engine_selection_and_ineligible_requests():
    Y, mask, W, H = problem(1800, 1500, 8, seed=1)
    data = prepare_data(Y, mask, transpose=False, dtype="float32", device=None)
    kw = dict(alpha=1.2, beta=1.2, eps=1e-8, projection="normalize", max_iter_cap=1, device=None)
    with make_problem(data, 8, dtype="float32", mask_semantics="reference", **kw) as p:
        assert p.engine == "tensor" and not p.fit_is_fused
    with make_problem(data, 8, dtype="float32", mask_semantics="strict", **kw) as p:
        assert p.engine == "tensor"                                  # strict H pass variant reads the mask plane too
    with make_problem(data, 40, dtype="float32", mask_semantics="reference", **kw) as p:
        assert p.engine == "tensor"                                  # K <= 64: the 3-pipeline instantiation
    Y, mask = Y[:600, :700].copy(), mask[:600, :700].copy()
    data = prepare_data(Y, mask, transpose=False, dtype="float32", device=None)
    with make_problem(data, 8, dtype="float32", mask_semantics="reference", **kw) as p:
        assert p.engine == "fused" and p.fit_is_fused                # a single small fit: the persistent small-fit kernel
    with make_problem(data, 8, dtype="float32", mask_semantics="reference", engine="tensor", **kw) as p:
        assert p.engine == "tensor" and not p.fit_is_fused           # ... unless asked (batches of fits: multifit.py)
    with make_problem(data, 40, dtype="float32", mask_semantics="reference", **kw) as p:
        assert p.engine == "tensor"                                  # K > 32: not covered by the small-fit kernel
    d64 = prepare_data(Y, mask, transpose=False, dtype="float64", device=None)
    with make_problem(d64, 8, dtype="float64", mask_semantics="reference", **kw) as p:
        assert p.engine == "fused"
    with make_problem(d64, 8, dtype="float64", mask_semantics="reference", engine="simt", **kw) as p:
        assert p.engine == "simt" and not p.fit_is_fused
    with make_problem(d64, 40, dtype="float64", mask_semantics="reference", **kw) as p:
        assert p.engine == "simt"                                    # fp64, K > 32: the pass kernels
    with pytest.raises(RuntimeError, match="tensor engine"):
        make_problem(d64, 8, dtype="float64", mask_semantics="reference", engine="tensor", **kw)
    with pytest.raises(RuntimeError, match="fused engine"):
        make_problem(d64, 40, dtype="float64", mask_semantics="reference", engine="fused", **kw)
    small = prepare_data(Y[:100, :100], None, transpose=False, dtype="float32", device=None)
    with make_problem(small, 8, dtype="float32", mask_semantics="reference", **kw) as p:
        assert p.engine == "fused"
    with make_problem(small, 8, dtype="float32", mask_semantics="reference", engine="simt", **kw) as p:
        assert p.engine == "simt"
