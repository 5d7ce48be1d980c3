"""transform / inverse_transform / score / perplexity on the GPU against the golden vectors of the
reference estimator (_base.py:162-265), including its RNG and mask quirks."""
import numpy as np
import pytest

from conftest import rel_err
from nbmf_mm_b200 import NBMF

pytestmark = pytest.mark.gpu


def fitted(datasets, golden_transform, dtype="float64"):
    est = NBMF(n_components=4, max_iter=80, tol=0.0, random_state=1, dtype=dtype).fit(datasets["animals"])
    assert rel_err(est.components_, golden_transform["components"]) < (1e-9 if dtype == "float64" else 1e-3)
    return est


@pytest.mark.parametrize("tag", ["nomask", "mask"])
def test_transform_matches_reference(datasets, golden_transform, tag):
    est = fitted(datasets, golden_transform)
    est.components_ = golden_transform["components"]               # identical H: isolates transform
    mk = None if tag == "nomask" else golden_transform["mask"]
    np.random.seed(99)                                             # transform draws from the GLOBAL rng (_base.py:175)
    Wt = est.transform(datasets["animals"], mask=mk)
    assert Wt.shape == golden_transform[f"{tag}/Wt"].shape
    assert rel_err(Wt, golden_transform[f"{tag}/Wt"]) < 1e-9
    assert np.allclose(Wt.sum(axis=1), 1.0, atol=1e-12) and Wt.min() >= 0 and Wt.max() <= 1
    np.random.seed(99)
    assert np.array_equal(est.transform(datasets["animals"], mask=mk), Wt)
    assert not np.array_equal(est.transform(datasets["animals"], mask=mk), Wt)   # unseeded second call differs


@pytest.mark.parametrize("tag", ["nomask", "mask"])
def test_score_and_perplexity_match_reference(datasets, golden_transform, tag):
    est = fitted(datasets, golden_transform)
    est.components_ = golden_transform["components"]
    mk = None if tag == "nomask" else golden_transform["mask"]
    np.random.seed(99)
    s = est.score(datasets["animals"], mask=mk)
    assert isinstance(s, float) and abs(s - float(golden_transform[f"{tag}/score"])) < 1e-9 * abs(s)
    np.random.seed(99)
    p = est.perplexity(datasets["animals"], mask=mk)
    assert p >= 1.0 and abs(p - np.exp(-s)) < 1e-12 * p


def test_fp32_transform_close(datasets, golden_transform):
    est = fitted(datasets, golden_transform, dtype="float32")
    est.components_ = golden_transform["components"]
    np.random.seed(99)
    Wt = est.transform(datasets["animals"])
    assert rel_err(Wt, golden_transform["nomask/Wt"]) < 1e-4
    assert np.allclose(Wt.sum(axis=1), 1.0, atol=1e-6)


def test_transform_new_rows_and_feature_mismatch(datasets, golden_transform):
    est = fitted(datasets, golden_transform)
    Xnew = (np.random.default_rng(3).random((20, 85)) < 0.3).astype(float)
    W = est.transform(Xnew)
    assert W.shape == (20, 4) and np.all(W >= 0) and np.all(W <= 1)
    with pytest.raises(ValueError, match="features"):
        est.transform(np.zeros((5, 84)))
    # dir-beta models also use the beta-dir W step in transform, exactly like the reference
    est2 = NBMF(n_components=4, orientation="dir-beta", max_iter=30, random_state=0).fit(datasets["animals"])
    W2 = est2.transform(datasets["animals"])
    assert W2.shape == (50, 4) and np.allclose(W2.sum(axis=1), 1.0, atol=1e-12)
