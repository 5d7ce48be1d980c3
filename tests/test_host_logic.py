"""Host-side logic of the drop-in estimator that needs no GPU: validation order and error
messages, orientation aliases, bit packing, shard planning, the simplex clean-up tail."""
import numpy as np
import pytest

from nbmf_mm_b200 import NBMF, NBMFMM, BitMatrix
from nbmf_mm_b200.bits import words_per_row
from nbmf_mm_b200.solver import _row_shard
from nbmf_oracle import final_simplex_cleanup       # the rule itself lives in the oracle; the product applies it on the device
from nbmf_mm_b200._utils import check_is_fitted, generate_synthetic_binary_data


def test_constructor_matches_reference_signature():
    est = NBMFMM()
    p = est.get_params()
    # reference defaults, _base.py:63-66
    assert (p["n_components"], p["alpha"], p["beta"], p["max_iter"], p["tol"]) == (10, 1.2, 1.2, 2000, 1e-5)
    assert p["orientation"] == "beta-dir" and p["W_init"] is None and p["init"] is None and p["verbose"] == 0
    # README-only kwargs the reference lacks
    assert p["projection_method"] == "normalize" and p["n_init"] == 1
    assert NBMF is NBMFMM
    from sklearn.base import clone
    assert clone(NBMF(n_components=3, n_init=4)).get_params()["n_init"] == 4


def test_non_binary_X_rejected_before_any_device_work():
    with pytest.raises(ValueError, match="must be binary"):
        NBMFMM(n_components=3).fit(np.random.default_rng(0).standard_normal((20, 10)))
    with pytest.raises(ValueError):
        NBMFMM(n_components=3).fit(np.array([[0.0, np.nan], [1.0, 0.0]]))
    with pytest.raises(ValueError):
        NBMFMM(n_components=3).fit(np.zeros(5))


def test_orientation_aliases_exact_keys():
    est = NBMF()
    table = {"beta-dir": "beta-dir", "dir-beta": "dir-beta", "Beta-Dir": "beta-dir", "Dir-Beta": "dir-beta",
             "Dir Beta": "dir-beta", "binary ICA": "beta-dir", "Binary ICA": "beta-dir", "bICA": "beta-dir",
             "Aspect Bernoulli": "dir-beta"}
    for alias, canon in table.items():
        assert est._normalize_orientation(alias) == canon
    for bad in ("Dir-Dir", "BETA-DIR", "aspect bernoulli", ""):
        with pytest.raises(ValueError, match="Unknown orientation"):
            est._normalize_orientation(bad)
    X = (np.random.default_rng(0).random((8, 6)) < 0.4).astype(float)
    with pytest.raises(ValueError, match="Unknown orientation"):
        NBMF(n_components=2, orientation="Dir-Dir").fit(X)


def test_unfitted_estimator_raises_like_reference():
    with pytest.raises(ValueError, match="not fitted yet"):
        NBMF().transform(np.zeros((3, 3)))
    with pytest.raises(ValueError, match="not fitted yet"):
        NBMF().inverse_transform(np.zeros((3, 10)))
    with pytest.raises(ValueError, match="not fitted yet"):
        check_is_fitted(NBMF(), "components_")


def test_bitmatrix_roundtrip_and_layout():
    rng = np.random.default_rng(1)
    for m, n in [(1, 1), (3, 31), (5, 32), (7, 33), (9, 1024), (4, 1025), (50, 85)]:
        A = rng.random((m, n)) < 0.4
        B = BitMatrix.from_dense(A)
        assert B.words.shape == (m, words_per_row(n)) and B.words.dtype == np.uint32
        assert np.array_equal(B.to_dense(bool), A)
        assert B.count() == int(A.sum())
        # bit j of word j // 32, little-endian inside the word; padding is zero
        i, j = m - 1, n - 1
        assert ((int(B.words[i, j // 32]) >> (j % 32)) & 1) == int(A[i, j])
        assert B.count() == int(np.unpackbits(B.words.view(np.uint8)).sum())
        T = B.transpose()
        assert T.shape == (n, m) and np.array_equal(T.to_dense(bool), A.T)
    with pytest.raises(ValueError):
        BitMatrix(np.zeros((3, 5), dtype=np.uint32), (3, 40))


def test_row_shards_cover_everything_once():
    for m in (1, 31, 32, 33, 1000, 1226, 10**6):
        for world in (1, 2, 3, 4, 8):
            spans = [_row_shard(m, r, world) for r in range(world)]
            assert spans[0][0] == 0 and max(s[1] for s in spans) == m
            covered = sum(max(0, b - a) for a, b in spans)
            assert covered == m
            for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
                assert a1 == b0 or b1 == b0        # contiguous; trailing ranks may be empty
            assert all(a % 32 == 0 for a, b in spans if b > a)


def test_final_simplex_cleanup_rules():
    W = np.array([[0.2, 0.8], [0.5, 0.5000001], [0.0, 0.0]])
    H = np.ones((2, 3))
    W2, _ = final_simplex_cleanup(W.copy(), H.copy(), "beta-dir")
    assert np.allclose(W2[:2].sum(axis=1), 1.0, atol=1e-15) and np.array_equal(W2[2], [0.0, 0.0])
    ok = np.array([[0.25, 0.75], [0.5, 0.5 + 5e-10]])
    W3, _ = final_simplex_cleanup(ok.copy(), H.copy(), "beta-dir")
    assert np.array_equal(W3, ok)                  # deviation below 1e-9: untouched
    Hd = np.array([[0.3, 0.6], [0.3, 0.6]])
    _, H4 = final_simplex_cleanup(W.copy(), Hd.copy(), "dir-beta")
    assert np.allclose(H4.sum(axis=0), 1.0)


def test_synthetic_generator_contract():
    X, W, H = generate_synthetic_binary_data(40, 25, 4, sparsity=0.3, random_state=7)
    assert X.shape == (40, 25) and W.shape == (40, 4) and H.shape == (4, 25)
    assert set(np.unique(X)) <= {0.0, 1.0} and set(np.unique(H)) <= {0.0, 1.0}
    assert W.min() >= 0.1 and W.max() <= 0.9
    X2, _, _ = generate_synthetic_binary_data(40, 25, 4, sparsity=0.3, random_state=7)
    assert np.array_equal(X, X2)


def test_import_shim_exposes_reference_names():
    import nbmf_mm
    from nbmf_mm import NBMF as A, NBMFMM as B, nbmf_mm_solver  # noqa: F401
    from nbmf_mm._utils import generate_synthetic_binary_data as g  # noqa: F401
    assert A is B and set(nbmf_mm.__all__) == {"NBMFMM", "NBMF", "nbmf_mm_solver"}


def test_dense_inputs_are_uploaded_in_a_dtype_the_packing_kernels_read():
    """Host side of the device front end (nbmf_pack_bits_checked): f32 / f64 / u8 go up as they are, bool is a u8 view
    (no copy), everything else becomes f64."""
    from nbmf_mm_b200.device import _packable
    a = np.arange(6).reshape(2, 3)
    for dt in (np.float32, np.float64, np.uint8):
        x = a.astype(dt)
        assert _packable(x) is x
    b = a.astype(bool)
    v = _packable(b)
    assert v.dtype == np.uint8 and np.shares_memory(v, b)
    for dt in (np.int64, np.int32, np.float16):
        assert _packable(a.astype(dt)).dtype == np.float64
    assert _packable([[0, 1], [1, 0]]).dtype == np.float64


def test_package_import_asks_for_more_hardware_queues_without_overriding_the_user():
    """nbmf_mm_multifit runs one fit per stream; the package only sets a DEFAULT for CUDA_DEVICE_MAX_CONNECTIONS."""
    import os, subprocess, sys as _sys
    code = "import os, nbmf_mm_b200; print(os.environ['CUDA_DEVICE_MAX_CONNECTIONS'])"
    root = str(__import__("pathlib").Path(__file__).resolve().parents[1])
    env = {k: v for k, v in os.environ.items() if k != "CUDA_DEVICE_MAX_CONNECTIONS"}
    env["PYTHONPATH"] = root
    assert subprocess.run([_sys.executable, "-c", code], env=env, capture_output=True, text=True).stdout.strip() == "32"
    env["CUDA_DEVICE_MAX_CONNECTIONS"] = "4"
    assert subprocess.run([_sys.executable, "-c", code], env=env, capture_output=True, text=True).stdout.strip() == "4"


def test_solver_rejects_bad_shapes_before_touching_the_device():
    from nbmf_mm_b200 import nbmf_mm_solver
    with pytest.raises(ValueError, match="2-D"):
        nbmf_mm_solver(np.zeros(5), 2)
    with pytest.raises(ValueError, match="Unknown orientation"):
        nbmf_mm_solver(np.zeros((4, 4)), 2, orientation="Dir Beta")     # the solver takes canonical names only
    with pytest.raises(UnboundLocalError):
        nbmf_mm_solver(np.zeros((4, 4)), 2, max_iter=0)
    with pytest.raises(ValueError, match="shard"):
        nbmf_mm_solver(np.zeros((4, 4)), 2, shard=(0, 8))


def test_engine_is_resolved_from_the_global_problem_size():
    """Row shards must agree on the engine (the tensor engine pads K to its own tile: ranks that disagreed would
    all-reduce differently shaped buffers).  m = 2000 on 4 ranks gives shards of 512, 512, 512 and 464 rows: decided per
    shard, three ranks would take the tensor engine and one the SIMT engine."""
    from nbmf_mm_b200.solver import resolve_engine
    kw = dict(dtype="float32", vkind="bits", k=20, eps=1e-8, n=4096)
    assert [_row_shard(2000, r, 4) for r in range(4)] == [(0, 512), (512, 1024), (1024, 1536), (1536, 2000)]
    assert resolve_engine("auto", m_total=2000, **kw) == "tensor"
    assert resolve_engine("auto", m_total=400, **kw) == "simt"
    assert resolve_engine("auto", m_total=2000, dtype="float32", vkind="bits", k=20, eps=1e-8, n=100) == "simt"
    assert resolve_engine("simt", m_total=2000, **kw) == "simt"
    assert resolve_engine("auto", m_total=2000, dtype="float64", vkind="bits", k=20, eps=1e-8, n=4096) == "simt"
    assert resolve_engine("auto", m_total=2000, dtype="float32", vkind="dense", k=20, eps=1e-8, n=4096) == "simt"
    assert resolve_engine("auto", m_total=2000, dtype="float32", vkind="bits", k=65, eps=1e-8, n=4096) == "simt"


def test_dataset_reader_and_splits(datasets):
    """The data side of the experiment driver (examples/reproduce_magron2022.py:25-38): the stdlib .rda reader (checked
    against the committed bit-packed fixtures where the reference tree is present) and the seeded split."""
    from pathlib import Path
    from nbmf_mm_b200.datasets import load_dataset_and_splits, make_split
    tr, va, te = make_split((50, 85), seed=12345)
    assert np.array_equal(tr + va + te, np.ones((50, 85))) and abs(tr.mean() - 0.70) < 0.03 and abs(va.mean() - 0.15) < 0.03
    tr2, _, _ = make_split((50, 85), seed=12345)
    assert np.array_equal(tr, tr2)
    ref = Path("/root/reference/data")
    if ref.is_dir():
        Y, trm, vam, tem = load_dataset_and_splits("animals", ref)
        assert np.array_equal(Y, datasets["animals"]) and np.array_equal(trm, datasets["animals_train_mask"])
        assert np.array_equal(trm + vam + tem, np.ones_like(Y))
        Yl, a, b, c = load_dataset_and_splits("lastfm", ref)
        assert np.array_equal(Yl, datasets["lastfm"]) and np.array_equal(a + b + c, np.ones_like(Yl))


def test_concurrent_init_draws_equal_the_sequential_loop():
    """nbmf_mm_multifit draws the inits of seeded jobs concurrently from private generators; the arrays and the state the
    global NumPy stream is left in must be those of the reference's loop of solver calls (_solver.py:102-103,122-129)."""
    from nbmf_mm_b200.multifit import _draw_inits, draw_all_inits
    rng = np.random.default_rng(0)
    m, n = 37, 29
    for transpose in (False, True):
        for trial in range(6):
            jobs = []
            for r in range(7):
                k = int(rng.integers(2, 6))
                j = dict(n_components=k, random_state=int(rng.integers(0, 1000)))
                em, en = (n, m) if transpose else (m, n)                  # given inits are in the caller's orientation
                if rng.integers(0, 4) == 1:
                    j["W_init"], j["H_init"] = rng.random((em, k)), rng.random((k, en))
                jobs.append(j)
            if trial == 4:
                jobs[-1]["W_init"], jobs[-1]["H_init"] = rng.random((n if transpose else m, jobs[-1]["n_components"])), \
                    rng.random((jobs[-1]["n_components"], m if transpose else n))  # the last job draws nothing
            if trial == 5:
                jobs[3]["random_state"] = None                            # an unseeded job: the sequential loop is kept
            np.random.seed(123)
            seq = [_draw_inits(j.get("random_state"), m, n, int(j["n_components"]), j.get("W_init"), j.get("H_init"), transpose)
                   for j in jobs]
            after_seq = np.random.uniform()
            np.random.seed(123)
            par = draw_all_inits(jobs, m, n, transpose, n_threads=3)
            after_par = np.random.uniform()
            assert after_seq == after_par
            for (Ws, Hs), (Wp, Hp) in zip(seq, par):
                assert np.array_equal(Ws, Wp) and np.array_equal(Hs, Hp)


def test_experiment_driver_builds_the_loops_of_the_reference_script(monkeypatch):
    """Host side of nbmf_mm_b200.experiment (examples/reproduce_magron2022.py:74-152,242-329): job order (alpha outer,
    beta inner), the driver's seed, record fields, first arg-min of the validation perplexity -- with the device call
    replaced by a stub that returns a perplexity computed from the job."""
    from nbmf_mm_b200 import experiment
    seen = {}

    def stub(Y, jobs, *, mask=None, eval_masks=None, **kw):
        seen["jobs"], seen["kw"], seen["masks"] = jobs, kw, eval_masks
        out = []
        for j in jobs:
            val = abs(j["alpha"] - 1.5) + abs(j["beta"] - 2.0) + 1.0
            ho = {name: {"nll": np.log(val), "perplexity": val, "n_entries": 3} for name in eval_masks}
            out.append((np.zeros((4, j["n_components"])), np.zeros((j["n_components"], 5)), [np.float64(0.7), np.float64(0.6)],
                        0.0, 2, ho))
        return out
    monkeypatch.setattr(experiment, "nbmf_mm_multifit", stub)
    Y = np.zeros((4, 5))
    tr, va, te = np.ones((4, 5)), np.ones((4, 5)), np.ones((4, 5))
    out = experiment.grid_search(Y, tr, va, n_components=3, max_iter=7)
    assert [(j["alpha"], j["beta"]) for j in seen["jobs"]] == [(a, b) for a in experiment.ALPHA_VALUES for b in experiment.BETA_VALUES]
    assert all(j["random_state"] == 12345 and j["n_components"] == 3 for j in seen["jobs"])
    assert set(seen["masks"]) == {"train", "val"} and seen["kw"]["max_iter"] == 7 and seen["kw"]["orientation"] == "beta-dir"
    assert len(out["records"]) == 36 and (out["best"]["alpha"], out["best"]["beta"]) == (1.5, 2.0)
    assert set(out["best"]) == {"alpha", "beta", "k", "train_perplexity", "val_perplexity", "n_iter", "final_loss", "time"}
    assert out["best"]["n_iter"] == 2 and out["best"]["final_loss"] == 0.6
    recs = experiment.components_sweep(Y, tr, va, te, alpha=2.0, beta=2.0)
    assert [r["k"] for r in recs] == list(experiment.K_RANGE) and all("test_perplexity" in r for r in recs)
    one = experiment.fit_and_test(Y, tr, te, n_components=4, alpha=1.0, beta=1.0)
    assert seen["kw"]["max_iter"] == 1000 and one["W"].shape == (4, 4) and "val_perplexity" not in one
