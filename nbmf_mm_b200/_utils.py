"""Host helpers kept for drop-in compatibility with ``nbmf_mm._utils`` (reference
``src/nbmf_mm/_utils.py``): the fitted check and the synthetic generator the reference's
tests import."""
from __future__ import annotations

import numpy as np


def check_is_fitted(estimator, attributes):
    """Raise ``ValueError`` unless every name in ``attributes`` is set (``_utils.py:3-9``)."""
    names = [attributes] if isinstance(attributes, str) else list(attributes)
    missing = [a for a in names if not hasattr(estimator, a)]
    if missing:
        raise ValueError(f"This {type(estimator).__name__} instance is not fitted yet.")


def generate_synthetic_binary_data(n_samples=100, n_features=50, n_components=5, sparsity=0.3, random_state=None):
    """Binary data from a logistic-link low-rank model (``_utils.py:11-47``).

    Draw order (kept so seeded data equals the reference's): W ~ U(0.1, 0.9) (samples x k),
    binary H (k x features) with P(1) = ``sparsity``, then X ~ Bernoulli(sigmoid(W H)).
    Returns ``(X, W_true, H_true)``."""
    rng = np.random.RandomState(random_state)
    W_true = rng.uniform(0.1, 0.9, size=(n_samples, n_components))
    H_true = (rng.random((n_components, n_features)) < sparsity).astype(float)
    logits = W_true @ H_true
    X = (rng.random((n_samples, n_features)) < 1.0 / (1.0 + np.exp(-logits))).astype(float)
    return X, W_true, H_true
