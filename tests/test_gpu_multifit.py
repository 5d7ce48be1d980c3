"""Many independent fits of one data set (restarts, alpha/beta grids, K sweeps): one
upload, concurrent device-resident loops on separate CUDA streams.  Every job must equal the corresponding
``nbmf_mm_solver`` call bit for bit when both run on the same engine (``engine="auto"`` gives a single small fit to the
persistent small-fit kernel and a batch to the launch-per-kernel engines: equal to rounding only)."""
import numpy as np
import pytest

from nbmf_mm_b200 import NBMF, nbmf_mm_multifit, nbmf_mm_solver

pytestmark = pytest.mark.gpu


def _data(m=226, n=285, seed=0):
    rng = np.random.default_rng(seed)
    X = (rng.random((m, n)) < 0.05).astype(np.float64)          # lastfm-like sparsity
    mask = (rng.random((m, n)) < 0.85).astype(np.float64)
    return X, mask


@pytest.mark.parametrize("orientation,dtype", [("beta-dir", "float64"), ("dir-beta", "float64"), ("beta-dir", "float32")])
def test_grid_of_jobs_equals_sequential_solver_calls(orientation, dtype):
    X, mask = _data()
    jobs = [dict(n_components=k, alpha=a, beta=b, random_state=s)
            for k, a, b, s in [(6, 1.2, 1.2, 0), (10, 1.0, 1.4, 1), (10, 2.0, 1.0, 2), (16, 1.2, 1.2, 3), (33, 1.1, 1.3, 4),
                               (6, 1.2, 1.2, 0), (8, 1.5, 1.5, 7), (12, 1.2, 1.0, 8), (4, 1.0, 1.0, 9)]]
    stats = {}
    got = nbmf_mm_multifit(X, jobs, mask=mask, orientation=orientation, max_iter=40, tol=1e-4, dtype=dtype, n_streams=4,
                           stats=stats, engine="simt")
    assert stats["n_streams"] == 4 and len(got) == len(jobs)
    for j, out in zip(jobs, got):
        W, H, losses, _, n_iter = nbmf_mm_solver(X, mask=mask, orientation=orientation, max_iter=40, tol=1e-4, dtype=dtype,
                                                 engine="simt", **j)
        assert n_iter == out[4] and np.array_equal(losses, out[2])
        assert np.array_equal(W, out[0]) and np.array_equal(H, out[1])
    assert np.array_equal(got[0][0], got[5][0])                  # identical jobs give identical results


def test_n_init_uses_one_upload_and_keeps_the_best_restart():
    X, mask = _data(seed=3)
    kw = dict(n_components=7, max_iter=30, tol=0.0, random_state=11, engine="simt")
    est = NBMF(n_init=6, **kw).fit(X, mask=mask)
    singles = [NBMF(n_components=7, max_iter=30, tol=0.0, random_state=11 + r, engine="simt").fit(X, mask=mask) for r in range(6)]
    best = int(np.argmin([s.loss_curve_[-1] for s in singles]))
    assert est.best_init_ == best
    assert np.array_equal(est.W_, singles[best].W_) and np.array_equal(est.components_, singles[best].components_)
    assert est.transfer_stats_["n_streams"] >= 1
    # restart 0 of an n_init run is the n_init=1 run with the same random_state
    assert np.array_equal(singles[0].W_, NBMF(**kw).fit(X, mask=mask).W_)


def test_graph_replay_on_the_tensor_engine_equals_per_kernel_launches():
    """Mid-size float32 problems take the tcgen05 engine AND (on the fits' own streams) the CUDA-graph replay of four
    iterations per launch; the sequential solver calls run on the default stream, kernel by kernel.  Same bits,
    same n_iter -- including fits that the device-side stop rule ends in the middle of a replayed graph."""
    X, mask = _data(700, 640, seed=5)
    jobs = [dict(n_components=k, random_state=s, tol=t, max_iter=mi)
            for k, s, t, mi in [(8, 0, 0.0, 23), (8, 1, 3e-4, 60), (16, 2, 1e-3, 60), (32, 3, 0.0, 9), (5, 4, 5e-4, 41)]]
    stats = {}
    got = nbmf_mm_multifit(X, jobs, mask=mask, dtype="float32", n_streams=3, stats=stats, engine="tensor")
    assert stats["engine"] == "tensor"
    stopped_early = 0
    for j, out in zip(jobs, got):
        W, H, losses, _, n_iter = nbmf_mm_solver(X, mask=mask, dtype="float32", engine="tensor", **j)
        assert n_iter == out[4] and np.array_equal(losses, out[2])
        assert np.array_equal(W, out[0]) and np.array_equal(H, out[1])
        stopped_early += n_iter < j["max_iter"]
    assert stopped_early >= 1


@pytest.mark.parametrize("engine", ["simt", "fused"])
@pytest.mark.parametrize("orientation,dtype,projection,tol", [("beta-dir", "float32", "normalize", 0.0),
                                                              ("dir-beta", "float64", "duchi", 2e-4),
                                                              ("beta-dir", "float64", "normalize", 7.5e-4)])
def test_batched_restarts_equal_sequential_solver_calls(orientation, dtype, projection, tol, engine):
    """Restarts (same K, alpha, beta, max_iter, tol) advance together: one launch per kernel for the whole group
    (nbmf_batch_bind), every fit with its own device-side loss history and stop rule.  Bit-identical to the loop of
    solver calls, including restarts that stop at different iterations; odd jobs take the per-stream path."""
    X, mask = _data(seed=9)
    jobs = [dict(n_components=7, alpha=1.3, beta=1.1, random_state=r) for r in range(7)]
    jobs.insert(3, dict(n_components=5, random_state=99))        # a job of another group (singleton)
    stats = {}
    got = nbmf_mm_multifit(X, jobs, mask=mask, orientation=orientation, max_iter=45, tol=tol, dtype=dtype,
                           projection_method=projection, stats=stats, engine=engine)
    assert stats["batched"] == 7
    n_iters = set()
    for j, out in zip(jobs, got):
        W, H, losses, _, n_iter = nbmf_mm_solver(X, mask=mask, orientation=orientation, max_iter=45, tol=tol, dtype=dtype,
                                                 projection_method=projection, engine=engine, **j)
        assert n_iter == out[4] and np.array_equal(losses, out[2]), (j, n_iter, out[4])
        assert np.array_equal(W, out[0]) and np.array_equal(H, out[1]), (j, float(np.abs(W - out[0]).max()))
        n_iters.add(n_iter)
    if tol == 7.5e-4 and engine == "simt":                         # (CPU oracle: 12, 45, 11, 10, 10, 10, 11 iterations)
        assert len(n_iters) > 1, n_iters                           # the restarts did stop at different iterations
    plain = nbmf_mm_multifit(X, jobs, mask=mask, orientation=orientation, max_iter=45, tol=tol, dtype=dtype,
                             projection_method=projection, batch=False, engine=engine)
    for a, b in zip(got, plain):
        assert a[4] == b[4] and np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


@pytest.mark.parametrize("engine", ["simt", "fused"])
@pytest.mark.parametrize("dtype", ["float64", "float32"])
def test_batched_alpha_beta_grid_equals_sequential_solver_calls(dtype, engine):
    """The 36-point alpha / beta grid of the reference's experiment driver (examples/reproduce_magron2022.py:87-117): the
    Beta prior lives in each fit's device-side state, so the whole grid (same K, max_iter, tol) advances with one launch
    per kernel.  Bit-identical to the loop of solver calls."""
    X, mask = _data(seed=3)
    grid = [1.0, 1.2, 1.4, 1.6, 1.8, 2.0]
    jobs = [dict(n_components=8, alpha=a, beta=b, random_state=12345) for a in grid for b in grid]
    stats = {}
    got = nbmf_mm_multifit(X, jobs, mask=mask, max_iter=30, tol=1e-5, dtype=dtype, stats=stats, engine=engine)
    assert stats["batched"] == 36
    finals = set()
    for j, out in list(zip(jobs, got))[::5]:
        W, H, losses, _, n_iter = nbmf_mm_solver(X, mask=mask, max_iter=30, tol=1e-5, dtype=dtype, engine=engine, **j)
        assert n_iter == out[4] and np.array_equal(losses, out[2]), j
        assert np.array_equal(W, out[0]) and np.array_equal(H, out[1]), j
        finals.add(float(losses[-1]))
    assert len(finals) > 3                                        # the priors did differ


@pytest.mark.parametrize("dtype,tol_f", [("float64", 1e-12), ("float32", 2e-5)])
def test_batch_planned_launches_agree_with_solver_calls_to_rounding(dtype, tol_f):
    """``batch_plan="batch"``: the launches of a group are planned for the whole group (fewer, larger splits per fit),
    which changes the summation order of the split partials -- same results to rounding, same iteration counts."""
    rng = np.random.default_rng(0)
    X = (rng.random((1226, 285)) < 0.0435).astype(np.float64)            # config-5 shape
    jobs = [dict(n_components=16, random_state=r) for r in range(12)]
    stats = {}
    got = nbmf_mm_multifit(X, jobs, max_iter=40, tol=0.0, dtype=dtype, stats=stats, batch_plan="batch", engine="simt")
    assert stats["batched"] == 12
    for j, out in list(zip(jobs, got))[::4]:
        W, H, losses, _, n_iter = nbmf_mm_solver(X, max_iter=40, tol=0.0, dtype=dtype, engine="simt", **j)
        assert n_iter == out[4]
        assert np.max(np.abs(np.asarray(losses) - np.asarray(out[2])) / np.abs(losses)) < tol_f
        assert np.max(np.abs(W - out[0])) < 50 * tol_f and np.max(np.abs(H - out[1])) < 50 * tol_f
