"""The persistent small-fit kernel (``csrc/fused_small.cu``): whole MM iterations of a small fit in one cooperative launch.

It must (1) actually be the path small fits take, (2) agree with the regular pass kernels (``engine="simt"``) to
rounding -- same ``n_iter``, same stop decisions -- in every mode it covers (both dtypes, the mask quirk and strict
semantics, both projections, K = 1 ... 32, ragged shapes), (3) agree with the oracle, and (4) give the same bits however the
iterations are cut into launches and however many CTAs a fit gets (batched restarts)."""
import sys
from pathlib import Path

import numpy as np
import pytest

sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "oracle"))

from nbmf_mm_b200 import nbmf_mm_multifit, nbmf_mm_solver                 # noqa: E402
from nbmf_mm_b200.solver import make_problem, prepare_data                  # noqa: E402

pytestmark = pytest.mark.gpu


def _problem(data, k, dtype, projection="normalize", sem="reference", cap=2000, engine="fused"):
    return make_problem(data, k, dtype=dtype, alpha=1.2, beta=1.2, eps=1e-8, mask_semantics=sem, projection=projection,
                        max_iter_cap=cap, device=None, engine=engine)


def _data(m, n, seed, density=0.2, obs=0.85):
    rng = np.random.default_rng(seed)
    X = (rng.random((m, n)) < density).astype(np.float64)
    mask = (rng.random((m, n)) < obs).astype(np.float64)
    return X, mask


CASES = [
    # m, n, k, dtype, masked, orientation, projection, mask_semantics
    (100, 500, 6, "float64", False, "beta-dir", "normalize", "reference"),      # config 1
    (253, 902, 10, "float64", True, "beta-dir", "normalize", "reference"),      # paleo shape, train mask
    (1226, 285, 10, "float64", True, "beta-dir", "normalize", "strict"),
    (226, 285, 32, "float64", True, "beta-dir", "normalize", "reference"),      # K = 32: the widest instantiation
    (50, 85, 1, "float64", True, "beta-dir", "normalize", "reference"),         # K = 1
    (37, 33, 5, "float64", False, "dir-beta", "normalize", "reference"),        # one word + one column, transposed
    (120, 70, 7, "float64", True, "dir-beta", "duchi", "reference"),            # Duchi with per-row counts
    (300, 260, 17, "float32", True, "beta-dir", "normalize", "reference"),
    (100, 500, 6, "float32", False, "beta-dir", "duchi", "strict"),
    (1226, 100, 12, "float32", True, "beta-dir", "normalize", "strict"),        # n < 128: not tensor-eligible
]


@pytest.mark.parametrize("m,n,k,dtype,masked,orientation,projection,sem", CASES)
def test_fused_fit_equals_the_pass_kernels(monkeypatch, m, n, k, dtype, masked, orientation, projection, sem):
    X, mask = _data(m, n, seed=m + n + k)
    kw = dict(max_iter=60, tol=1e-5, mask=mask if masked else None, random_state=3, dtype=dtype, orientation=orientation,
              projection_method=projection, mask_semantics=sem)
    data = prepare_data(X, kw["mask"], transpose=orientation == "dir-beta", dtype=dtype, device=None)
    for engine in ("auto", "fused"):                               # every case here is small enough for the auto rule
        prob = _problem(data, k, dtype, projection, sem, engine=engine)
        assert prob.fit_is_fused and prob.engine == "fused"
        prob.close()
    Wf, Hf, lf, _, nf = nbmf_mm_solver(X, k, engine="auto", **kw)
    prob = _problem(data, k, dtype, projection, sem, engine="simt")
    assert not prob.fit_is_fused and prob.engine == "simt"
    prob.close()
    monkeypatch.setenv("NBMF_NO_FUSED", "1")                       # the auto rule without the small-fit kernel
    prob = _problem(data, k, dtype, projection, sem, engine="auto")
    assert not prob.fit_is_fused
    prob.close()
    monkeypatch.delenv("NBMF_NO_FUSED")
    Wr, Hr, lr, _, nr = nbmf_mm_solver(X, k, engine="simt", **kw)
    assert nf == nr
    tol = 1e-11 if dtype == "float64" else 2e-5
    assert np.max(np.abs(np.asarray(lf) - np.asarray(lr)) / np.abs(lr)) < tol
    assert np.max(np.abs(Wf - Wr)) < 50 * tol and np.max(np.abs(Hf - Hr)) < 50 * tol


def test_fused_fit_against_the_oracle():
    import nbmf_oracle as orc
    X, mask = _data(180, 333, seed=9, density=0.1)
    W, H, losses, _, n_iter = nbmf_mm_solver(X, 9, max_iter=150, tol=1e-6, mask=mask, random_state=1, alpha=1.3, beta=1.1)
    Wo, Ho, lo, no = orc.fit(X, 9, max_iter=150, tol=1e-6, mask=mask, random_state=1, alpha=1.3, beta=1.1)
    assert n_iter == no
    assert np.max(np.abs(np.asarray(losses) - np.asarray(lo)) / np.abs(lo)) < 1e-11
    assert np.max(np.abs(W - Wo)) < 1e-9 and np.max(np.abs(H - Ho)) < 1e-9


@pytest.mark.parametrize("dtype", ["float64", "float32"])
def test_result_does_not_depend_on_how_the_iterations_are_cut_into_launches(dtype):
    X, mask = _data(200, 300, seed=4)
    data = prepare_data(X, mask, transpose=False, dtype=dtype, device=None)
    rng = np.random.RandomState(0)
    W0, H0 = rng.uniform(0.1, 0.9, (200, 8)), rng.uniform(0.1, 0.9, (8, 300))
    outs = []
    for chunks in ([1000], [1] * 40, [3, 7, 1, 29], [39, 1], [40]):
        prob = _problem(data, 8, dtype, cap=40)
        assert prob.fit_is_fused
        prob.set_factors(W0, H0, normalize_w=True)
        prob.fit_begin(40, 0.0)
        for c in chunks:
            prob.fit_enqueue(c)
        done, n_hist = prob.fit_poll(wait=True)
        assert done and n_hist == 40
        hist, conv = prob.fit_history(n_hist)
        W, H = prob.get_factors_f64()
        outs.append((W.copy(), H.copy(), hist.copy()))
        prob.close()
    for W, H, hist in outs[1:]:
        assert np.array_equal(W, outs[0][0]) and np.array_equal(H, outs[0][1]) and np.array_equal(hist, outs[0][2])


@pytest.mark.parametrize("dtype,n_jobs", [("float64", 7), ("float32", 7), ("float64", 200)])
def test_batched_restarts_in_the_fused_kernel_equal_single_fits(dtype, n_jobs):
    """A batch gives each fit a few CTAs (200 fits: more than one co-resident grid, run in slices); a single fit gets the
    whole GPU.  The summation orders depend on the shape only, so the bits agree."""
    X, mask = _data(150, 100, seed=6)
    jobs = [dict(n_components=5, random_state=r, alpha=1.0 + 0.01 * (r % 5)) for r in range(n_jobs)]
    stats = {}
    got = nbmf_mm_multifit(X, jobs, mask=mask, max_iter=30, tol=1e-4, dtype=dtype, stats=stats, engine="fused")
    assert stats["batched"] == n_jobs and stats["engine"] == "fused"
    for j, out in list(zip(jobs, got))[:: max(1, n_jobs // 7)]:
        W, H, losses, _, n_iter = nbmf_mm_solver(X, mask=mask, max_iter=30, tol=1e-4, dtype=dtype, **j)   # auto: fused
        assert n_iter == out[4] and np.array_equal(losses, out[2])
        assert np.array_equal(W, out[0]) and np.array_equal(H, out[1])
