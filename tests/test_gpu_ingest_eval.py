"""Callers and data formats either side of the hot path: CSR ingestion without a dense
M x N copy, inverse_transform on the device, and held-out evaluation of the fitted factors.  Bit planes are
compared bit for bit; floating-point results in fp64 parity mode against plain NumPy formulas."""
import numpy as np
import pytest

from nbmf_mm_b200 import NBMF, BitMatrix
from nbmf_mm_b200.device import pack_csr_device

pytestmark = pytest.mark.gpu
sp = pytest.importorskip("scipy.sparse")


def _xy(m, n, seed, density=0.15):
    rng = np.random.default_rng(seed)
    X = (rng.random((m, n)) < density).astype(np.float64)
    mask = (rng.random((m, n)) < 0.8).astype(np.float64)
    return X, mask


@pytest.mark.parametrize("m,n", [(1, 1), (37, 1025), (300, 70), (64, 4096)])
def test_csr_packing_is_bit_exact(m, n):
    X, _ = _xy(m, n, seed=m + n)
    A = sp.csr_matrix(X)
    A.data[::7] = 0.0                                           # explicit zeros must stay zero bits
    want = BitMatrix.from_dense(A.toarray())
    got, flags, h2d = pack_csr_device(A, None)
    assert flags == 0 and (h2d < X.nbytes or m * n < 64)
    assert np.array_equal(got.words.cpu().numpy().view(np.uint32), want.words)
    empty, flags, _ = pack_csr_device(sp.csr_matrix((m, n)), None)
    assert flags == 0 and not empty.words.cpu().numpy().any()


def test_csr_value_checks_match_the_reference_errors():
    X, mask = _xy(40, 50, seed=1)
    bad = sp.csr_matrix(X * 2.0)
    with pytest.raises(ValueError, match="X must be binary"):
        NBMF(n_components=3, max_iter=2).fit(bad)
    weighted = sp.csr_matrix(mask * 0.5)                          # a weighted mask is accepted (dense layout), as the reference does
    est = NBMF(n_components=3, max_iter=2).fit(sp.csr_matrix(X), mask=weighted)
    assert np.isfinite(est.loss_)


@pytest.mark.parametrize("orientation", ["beta-dir", "dir-beta"])
def test_sparse_fit_equals_dense_fit(orientation):
    X, mask = _xy(90, 140, seed=3)
    kw = dict(n_components=5, max_iter=30, tol=0.0, random_state=0, orientation=orientation)
    dense = NBMF(**kw).fit(X, mask=mask)
    sparse = NBMF(**kw).fit(sp.csr_matrix(X), mask=sp.csr_matrix(mask))
    assert np.array_equal(dense.W_, sparse.W_) and np.array_equal(dense.components_, sparse.components_)
    assert np.array_equal(dense.loss_curve_, sparse.loss_curve_)
    assert sparse.transfer_stats_["h2d_bytes"] < dense.transfer_stats_["h2d_bytes"] + 8 * (X.size + mask.size)
    mixed = NBMF(**kw).fit(sp.csr_matrix(X), mask=mask)          # sparse X, dense mask
    assert np.array_equal(dense.W_, mixed.W_)
    # probabilistic sparse X (values strictly inside (0,1)) takes the dense layout, like dense input does
    Xp = X * 0.7
    a = NBMF(**kw).fit(Xp)
    b = NBMF(**kw).fit(sp.csr_matrix(Xp))
    assert np.array_equal(a.W_, b.W_) and np.array_equal(a.loss_curve_, b.loss_curve_)


def test_inverse_transform_on_device():
    X, _ = _xy(130, 257, seed=4)
    est = NBMF(n_components=7, max_iter=20, tol=0.0, random_state=1).fit(X)
    rng = np.random.default_rng(0)
    W = rng.uniform(0, 0.4, (33, 7))
    want = np.clip(W @ est.components_, 0.0, 1.0)               # _base.py:208
    got = est.inverse_transform(W)
    assert got.shape == want.shape and np.max(np.abs(got - want)) < 1e-14
    assert np.max(np.abs(est.inverse_transform(est.W_) - np.clip(est.W_ @ est.components_, 0, 1))) < 1e-14
    with pytest.raises(ValueError):
        est.inverse_transform(W[:, :5])


@pytest.mark.parametrize("dtype,tol", [("float64", 1e-12), ("float32", 2e-6)])
def test_held_out_evaluation_matches_the_experiment_driver_formula(dtype, tol):
    X, train = _xy(120, 200, seed=5)
    rng = np.random.default_rng(9)
    val = ((1 - train) * (rng.random(X.shape) < 0.5)).astype(np.float64)
    est = NBMF(n_components=6, max_iter=60, tol=0.0, random_state=2, dtype=dtype).fit(X, mask=train)
    Yhat = est.W_ @ est.components_

    def perplexity(mask):                                       # examples/reproduce_magron2022.py:40-47
        ll = X * np.log(Yhat + 1e-8) + (1 - X) * np.log(1 - Yhat + 1e-8)
        return np.exp(-np.sum(mask * ll) / np.count_nonzero(mask))

    for mk in (val, train, None):
        out = est.evaluate(X, mask=mk)
        want = perplexity(np.ones_like(X) if mk is None else mk)
        assert abs(out["perplexity"] - want) < tol * want
        assert out["n_entries"] == (X.size if mk is None else np.count_nonzero(mk))
    out = est.evaluate(sp.csr_matrix(X), mask=sp.csr_matrix(val))   # sparse in, same number
    assert abs(out["perplexity"] - perplexity(val)) < tol * perplexity(val)


@pytest.mark.parametrize("orientation", ["beta-dir", "dir-beta"])
def test_fp16_layout_of_probabilistic_v(orientation):
    """dense_storage='float16' stores V*mask as fp16 (half the bytes per pass), arithmetic stays float32: values
    that are exact in fp16 give bit-identical factors, arbitrary values differ by the storage rounding only."""
    rng = np.random.default_rng(11)
    X = rng.integers(0, 65, (150, 260)) / 64.0                  # exact in fp16
    mask = (rng.random(X.shape) < 0.85).astype(np.float64)
    kw = dict(n_components=9, max_iter=25, tol=0.0, random_state=4, orientation=orientation, dtype="float32")
    a = NBMF(**kw).fit(X, mask=mask)
    b = NBMF(dense_storage="float16", **kw).fit(X, mask=mask)
    assert np.array_equal(a.W_, b.W_) and np.array_equal(a.components_, b.components_)
    assert np.array_equal(a.loss_curve_, b.loss_curve_)
    Xr = rng.random(X.shape)
    c = NBMF(**kw).fit(Xr, mask=mask)
    d = NBMF(dense_storage="float16", **kw).fit(Xr, mask=mask)
    assert abs(c.loss_curve_[-1] - d.loss_curve_[-1]) < 1e-3 * abs(c.loss_curve_[-1])
    assert np.max(np.abs(c.components_ - d.components_)) < 5e-3
    with pytest.raises(ValueError, match="float32"):
        NBMF(n_components=3, max_iter=2, dtype="float64", dense_storage="float16").fit(Xr)


@pytest.mark.parametrize("dtype", ["float32", "float64"])
def test_solver_tail_on_device_equals_the_oracle_rule(dtype):
    """fp64 export + final simplex clean-up (_solver.py:192-213) run on the device; same result as the oracle's
    restatement of the rule (pinned to the reference by the golden trajectories) applied to the plain export."""
    from nbmf_oracle import final_simplex_cleanup
    from nbmf_mm_b200.solver import make_problem, prepare_data
    X, mask = _xy(257, 130, seed=8)
    data = prepare_data(X, mask, transpose=False, dtype=dtype, device=None)
    rs = np.random.RandomState(1)
    W0, H0 = rs.uniform(0.1, 0.9, (257, 5)), rs.uniform(0.1, 0.9, (5, 130))
    with make_problem(data, 5, dtype=dtype, alpha=1.2, beta=1.2, eps=1e-8, mask_semantics="reference",
                      projection="normalize", max_iter_cap=10, device=None) as prob:
        prob.set_factors(W0, H0, normalize_w=True)
        prob.fit(10, 0.0)
        W, H = prob.get_factors()
        dev = prob.simplex_deviation()
        assert abs(dev - np.max(np.abs(W.sum(axis=1) - 1.0))) < 1e-15
        want_W, want_H = final_simplex_cleanup(W.copy(), H.copy(), "beta-dir")
        got_W, got_H = prob.get_factors_f64(normalize_w=bool(dev > 1e-9))
    assert (dev > 1e-9) == (dtype == "float32")                  # fp64 keeps the simplex to ~1e-16: no renormalisation
    assert np.max(np.abs(got_W - want_W)) < 1e-15 and np.array_equal(got_H, want_H)
    assert np.max(np.abs(got_W.sum(axis=1) - 1.0)) < (1e-15 if dtype == "float32" else 1e-12)


@pytest.mark.parametrize("m,n,k,masked,engine", [(1000, 1300, 8, True, "tensor"), (1000, 1300, 8, False, "tensor"),
                                                  (777, 600, 5, True, "simt"), (130, 70, 3, True, "auto"),
                                                  (2100, 520, 32, True, "tensor")])
def test_streamed_host_planes_equal_device_planes(m, n, k, masked, engine):
    """Host bit planes go up in row chunks while the device prepares the chunks that have landed
    (nbmf_ingest_bits_rows); the fit must be bit-identical to the one on planes that were on the device before."""
    import torch
    from nbmf_mm_b200 import nbmf_mm_solver
    from nbmf_mm_b200.device import DeviceProblem
    X, mask = _xy(m, n, seed=m + k)
    P, M = BitMatrix.from_dense(X), (BitMatrix.from_dense(mask) if masked else None)   # P is NOT pre-masked
    kw = dict(max_iter=6, tol=0.0, random_state=3, dtype="float32", engine=engine)
    st = {}
    W1, H1, l1, _, n1 = nbmf_mm_solver(P, k, mask=M, stats=st, **kw)
    assert st["streamed"] and st["h2d_bytes"] >= P.words.nbytes * (2 if masked else 1)
    W2, H2, l2, _, n2 = nbmf_mm_solver(P.to_device("cuda"), k, mask=None if M is None else M.to_device("cuda"), stats=st, **kw)
    assert not st["streamed"]
    assert n1 == n2 and np.array_equal(W1, W2) and np.array_equal(H1, H2) and l1 == l2
    # chunk boundaries: many small chunks, planes and count checked directly
    with DeviceProblem(m, n, k, dtype="float32", has_mask=masked, engine=engine) as prob:
        prob.stream_bits_from_host(P, M, n_chunks=5)
        cnt = prob.finish_bits()
        assert cnt == (np.count_nonzero(mask) if masked else m * n)
        want = P.words & M.words if masked else P.words
        assert np.array_equal(prob._keep[0].cpu().numpy().view(np.uint32), want)
    with pytest.raises(ValueError, match="shapes"):
        with DeviceProblem(m, n, k, dtype="float32", has_mask=True, engine=engine) as prob:
            prob.stream_bits_from_host(P, BitMatrix.from_dense(mask[:-1]))


@pytest.mark.parametrize("xdt,mdt", [(np.float64, np.float64), (np.float32, np.bool_), (np.uint8, np.uint8),
                                     (np.int64, np.float32), (np.bool_, None)])
def test_dense_front_end_on_the_device_is_bit_exact(xdt, mdt):
    """Dense host X / mask are uploaded in row chunks as they are and checked + packed on the device
    (nbmf_pack_bits_checked): same planes as the host packing, same flags as the NumPy tests it replaces."""
    from nbmf_mm_b200.device import pack_host_dense_checked
    X, mask = _xy(301, 1100, seed=5)
    Xa = X.astype(xdt)
    Ma = None if mdt is None else mask.astype(mdt)
    want_p = BitMatrix.from_dense((X != 0) & ((mask != 0) if Ma is not None else True))
    # pageable chunks (small inputs) and the pinned, threaded staging pipeline large inputs take (forced here)
    for kw in (dict(), dict(pinned_from=0, n_threads=3)):
        P, M, flags, h2d = pack_host_dense_checked(Xa, Ma, None, chunk_bytes=64 * 1100 * 8, **kw)   # 5 chunks, the last ragged
        assert flags == 0 and h2d == Xa.nbytes + (0 if Ma is None else Ma.nbytes)
        assert np.array_equal(P.words.cpu().numpy().view(np.uint32), want_p.words)
        if Ma is not None:
            assert np.array_equal(M.words.cpu().numpy().view(np.uint32), BitMatrix.from_dense(mask).words)
        else:
            assert M is None
    assert pack_host_dense_checked(np.asfortranarray(X), mask, None, chunk_bytes=64 * 1100 * 8, pinned_from=0)[0] is not None
    Pf = pack_host_dense_checked(np.asfortranarray(X), mask, None, chunk_bytes=64 * 1100 * 8, pinned_from=0)[0]
    assert np.array_equal(Pf.words.cpu().numpy().view(np.uint32), BitMatrix.from_dense((X != 0) & (mask != 0)).words)   # any strides
    # flags: probabilistic values, out-of-range values, NaN, weighted mask
    Xp = X.copy(); Xp[7, 3] = 0.25
    assert pack_host_dense_checked(Xp, mask, None)[2] == 1
    Xo = X.copy(); Xo[300, 1099] = 1.5
    assert pack_host_dense_checked(Xo, None, None)[2] == 3
    Xn = X.copy(); Xn[0, 0] = np.nan
    assert pack_host_dense_checked(Xn, None, None)[2] & 2
    assert pack_host_dense_checked(X, mask * 0.5, None)[2] == 4


def test_dense_front_end_errors_and_large_x_range_check():
    from nbmf_mm_b200 import nbmf_mm_solver
    X, mask = _xy(60, 90, seed=2)
    with pytest.raises(ValueError, match="weighted"):
        nbmf_mm_solver(BitMatrix.from_dense(X), 3, max_iter=2, mask=mask * 0.5)     # bit-packed X cannot carry mask values
    with pytest.raises(ValueError, match="mask has shape"):
        nbmf_mm_solver(X, 3, max_iter=2, mask=mask[:-1])
    with pytest.raises(ValueError, match="X must be binary"):
        nbmf_mm_solver(X * 2.0, 3, max_iter=2, check_range=True)
    # a large X takes the device-side range check of the estimator (no NumPy passes over X)
    big = np.zeros((2100, 2048)); big[5, 7] = 1.0; big[2099, 2047] = 1.0
    NBMF(n_components=2, max_iter=2, dtype="float32").fit(big)
    big[1000, 1000] = -0.5
    with pytest.raises(ValueError, match="X must be binary"):
        NBMF(n_components=2, max_iter=2, dtype="float32").fit(big)
    with pytest.raises(ValueError, match="X must be binary"):            # reference order: X before orientation
        NBMF(n_components=2, max_iter=2, orientation="nope").fit(big)
    # NaN / inf in a large X: found by the device pass too (no host finiteness pass), reported with sklearn's message,
    # which the reference's check_array raises before anything else (_base.py:83)
    big[1000, 1000] = np.nan
    with pytest.raises(ValueError, match="contains NaN"):
        NBMF(n_components=2, max_iter=2, dtype="float32").fit(big)
    big[1000, 1000] = np.inf
    with pytest.raises(ValueError, match="infinity"):
        NBMF(n_components=2, max_iter=2, n_init=2).fit(big)
