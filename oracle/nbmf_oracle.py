"""CPU oracle for the NBMF-MM hot path (TEST INFRASTRUCTURE -- not the product).

This module is a NumPy fp64 restatement of the reference algorithm
(siddC/nbmf_mm, ``src/nbmf_mm/_solver.py`` and the W-step/NLL parts of
``src/nbmf_mm/_base.py``).  It exists only so that ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl
reference`` legs have something to check the CUDA path against and to time on
the host cores.  Nothing under ``nbmf_mm_b200/`` imports it.

Pinning: ``oracle/make_golden.py`` imports the real reference from
``/root/reference/src`` in the authoring container, runs it and this module on
identical inputs, asserts agreement (bit-exact where the op order is kept) and
writes ``tests/golden/*.npz``.  ``tests/test_oracle_golden.py`` re-checks this
module against those committed vectors on every run (no reference needed).

Unpinned parts (the reference has no code for them):
``projection="duchi"``, ``mask_semantics="strict"`` and the ``n_init`` restart
schedule.  Their spec is README prose + Duchi et al. 2008; parity for those is
"unpinned" and they are only checked against this restatement.

Internal notation follows the reference: Y is m x n, W is k x m (columns on the
simplex), H is k x n in (0, 1).
"""
from __future__ import annotations

import numpy as np

EPS_DEFAULT = 1e-8


# ---------------------------------------------------------------------------
# weights derived from the observation mask
# ---------------------------------------------------------------------------
def _weights(Y, mask, mask_semantics):
    """Return (pos, neg_for_H, posT, negT_for_W).

    reference (``_solver.py:21-32,43``): the H half-step and the loss use
    ``1 - Y*mask`` as the negative weight (unobserved entries count as observed
    zeros -- the reference quirk), the W half-step uses ``(1-Y)*mask``.
    strict: both use ``(1-Y)*mask`` (README/paper semantics; unpinned).
    """
    if mask is None:
        pos = Y
        negH = 1 - pos
        posT = Y.T
        negT = (1 - Y).T
        return pos, negH, posT, negT
    if hasattr(mask, "toarray"):
        mask = mask.toarray()
    pos = Y * mask
    posT = Y.T * mask.T
    negT = (1 - Y).T * mask.T
    if mask_semantics == "reference":
        negH = 1 - pos
    elif mask_semantics == "strict":
        negH = (1 - Y) * mask
    else:
        raise ValueError(f"unknown mask_semantics {mask_semantics!r}")
    return pos, negH, posT, negT


# ---------------------------------------------------------------------------
# simplex projections for the W half-step
# ---------------------------------------------------------------------------
def project_columns_duchi(U):
    """Euclidean projection of every column of U (k x m) onto the simplex.

    Duchi et al. 2008 sort/threshold: u sorted descending, rho = max{j :
    u_j - (sum_{i<=j} u_i - 1)/j > 0}, theta = (sum_{i<=rho} u_i - 1)/rho,
    w = max(v - theta, 0).  (README.md:27-30,177-182 of the reference name
    the method; there is no reference code -- parity unpinned.)
    """
    k, m = U.shape
    S = -np.sort(-U, axis=0)
    css = np.cumsum(S, axis=0) - 1.0
    idx = np.arange(1, k + 1, dtype=U.dtype)[:, None]
    cond = S - css / idx > 0
    rho = k - 1 - np.argmax(cond[::-1, :], axis=0)          # last True per column
    theta = css[rho, np.arange(m)] / (rho + 1.0)
    return np.maximum(U - theta[None, :], 0.0)


# ---------------------------------------------------------------------------
# one MM iteration (H half-step, then W half-step with the NEW H)
# ---------------------------------------------------------------------------
def h_half_step(Y, W, H, mask, alpha, beta, eps=EPS_DEFAULT, mask_semantics="reference", weights=None):
    """H half-step, ``_solver.py:34-47``.  Returns H_new (k x n).  ``weights`` = the tuple ``_weights`` returns,
    built once per iteration by ``mm_step`` exactly as the reference builds its masked copies once
    (``_solver.py:21-32``)."""
    pos, negH, _, _ = _weights(Y, mask, mask_semantics) if weights is None else weights
    prior_a = np.ones_like(H) * (alpha - 1)
    prior_b = np.ones_like(H) * (beta - 1)
    theta = W.T @ H                                             # :39
    num = H * (W @ (pos / (theta + eps))) + prior_a             # :42
    den = (1 - H) * (W @ (negH / (1 - theta + eps))) + prior_b  # :43
    H_new = num / (num + den + eps)                             # :46
    return np.clip(H_new, eps, 1 - eps)                         # :47


def w_half_step(Y, W, H_new, mask, eps=EPS_DEFAULT, projection="normalize", weights=None):
    """W half-step, ``_solver.py:50-57`` (same formulas as ``_base.py:180-193``).

    Always properly masked.  ``projection="normalize"``: multiplicative step,
    ``/n`` (full column count) then L1 renormalisation of every column.
    ``projection="duchi"`` (unpinned): multiplicative step divided by the
    per-row observed count, then Euclidean projection onto the simplex.
    """
    n = Y.shape[1]
    _, _, posT, negT = _weights(Y, mask, "reference") if weights is None else weights
    thetaT = H_new.T @ W                                        # :50  (n x m)
    G = H_new @ (posT / (thetaT + eps)) + (1 - H_new) @ (negT / (1 - thetaT + eps))  # :53
    step = W * G
    if projection == "normalize":
        step = step / n                                         # :54
        return step / step.sum(axis=0, keepdims=True)           # :57
    if projection == "duchi":
        if mask is None:
            counts = np.full(Y.shape[0], float(n))
        else:
            mk = mask.toarray() if hasattr(mask, "toarray") else np.asarray(mask)
            counts = (mk != 0).sum(axis=1).astype(np.float64)
        return project_columns_duchi(step / counts[None, :])
    raise ValueError(f"unknown projection {projection!r}")


def mm_step(Y, W, H, mask, alpha, beta, eps=EPS_DEFAULT,
            mask_semantics="reference", projection="normalize"):
    """One full MM iteration == ``nbmf_mm_update_beta_dir`` (``_solver.py:5-59``)."""
    weights = _weights(Y, mask, mask_semantics)          # once per iteration, as _solver.py:21-32
    H_new = h_half_step(Y, W, H, mask, alpha, beta, eps, mask_semantics, weights)
    W_new = w_half_step(Y, W, H_new, mask, eps, projection, weights)
    return W_new, H_new


def map_objective(Y, W, H, mask, alpha, beta, eps=EPS_DEFAULT, mask_semantics="reference"):
    """Per-observed-entry MAP objective, ``_solver.py:148-162``."""
    theta = W.T @ H
    if mask is None:
        ll = Y * np.log(theta + eps) + (1 - Y) * np.log(1 - theta + eps)
        n_obs = Y.size
    else:
        mk = mask.toarray() if hasattr(mask, "toarray") else mask
        pos = Y * mk
        if mask_semantics == "reference":
            ll = pos * np.log(theta + eps) + (1 - pos) * np.log(1 - theta + eps)
        else:
            ll = pos * np.log(theta + eps) + ((1 - Y) * mk) * np.log(1 - theta + eps)
        n_obs = np.count_nonzero(mk)
    prior_a = (alpha - 1) * np.sum(np.log(H + eps))
    prior_b = (beta - 1) * np.sum(np.log(1 - H + eps))
    return -(np.sum(ll) + prior_a + prior_b) / n_obs


# ---------------------------------------------------------------------------
# the fit loop
# ---------------------------------------------------------------------------
def draw_init(rng, m, n, k):
    """Init stream of ``_solver.py:126-129``: W (m x k) first, then H (k x n),
    both U(0.1, 0.9), with the INTERNAL (post-orientation) m, n."""
    W0 = rng.uniform(0.1, 0.9, (m, k))
    H0 = rng.uniform(0.1, 0.9, (k, n))
    return W0, H0


def fit(Y, n_components, max_iter=500, tol=1e-5, alpha=1.2, beta=1.2,
        W_init=None, H_init=None, mask=None, random_state=None,
        orientation="beta-dir", eps=EPS_DEFAULT,
        mask_semantics="reference", projection="normalize", return_trace=False):
    """Restatement of ``nbmf_mm_solver`` (``_solver.py:61-216``).

    Returns ``(W (m x k), H (k x n), losses, n_iter)`` in EXTERNAL orientation.
    The RNG stream equals the reference's ``np.random.seed(rs); uniform; uniform``
    because ``RandomState(rs)`` is the same legacy MT19937 stream.
    """
    rng = np.random.RandomState(random_state) if random_state is not None else np.random
    if mask is not None and hasattr(mask, "toarray"):
        mask = mask.toarray()
    Y = np.asarray(Y, dtype=np.float64)
    m, n = Y.shape
    k = n_components
    if orientation == "dir-beta":                    # :113-123
        Y = Y.T
        m, n = n, m
        if mask is not None:
            mask = mask.T
        if W_init is not None and H_init is not None:
            W_init, H_init = H_init.T, W_init.T
    elif orientation != "beta-dir":
        raise ValueError(f"orientation must be canonical, got {orientation!r}")
    if W_init is None:
        W_init = rng.uniform(0.1, 0.9, (m, k))
    if H_init is None:
        H_init = rng.uniform(0.1, 0.9, (k, n))
    W = np.asarray(W_init).T                          # :132
    H = np.asarray(H_init)
    W = W / W.sum(axis=0, keepdims=True)              # :136

    losses = []
    prev = np.inf
    trace = []
    it = -1
    for it in range(max_iter):                        # :143
        W, H = mm_step(Y, W, H, mask, alpha, beta, eps, mask_semantics, projection)
        loss = map_objective(Y, W, H, mask, alpha, beta, eps, mask_semantics)
        losses.append(loss)
        if return_trace:
            trace.append((W.copy(), H.copy()))
        if it > 0:                                    # :169-174
            if abs(prev - loss) / abs(prev) < tol:
                break
        prev = loss
    if it < 0:
        raise UnboundLocalError("max_iter=0: the reference leaves `iteration` unbound (_solver.py:215)")

    W_out, H_out = W.T, H                             # :178-184
    if orientation == "dir-beta":
        W_out, H_out = H_out.T, W_out.T
    W_out, H_out = final_simplex_cleanup(np.array(W_out), np.array(H_out), orientation)
    if return_trace:
        return W_out, H_out, losses, it + 1, trace
    return W_out, H_out, losses, it + 1


def final_simplex_cleanup(W_out, H_out, orientation):
    """``_solver.py:192-213``: renormalise the simplex factor only if the worst
    deviation exceeds 1e-9, skipping vectors whose sum is <= 1e-12."""
    tiny, dev_tol = 1e-12, 1e-9
    if orientation == "beta-dir":
        if W_out.size:
            rs = W_out.sum(axis=1, keepdims=True)
            dev = np.max(np.abs(rs - 1.0))
            if np.isfinite(dev) and dev > dev_tol:
                ok = (rs > tiny).ravel()
                if ok.any():
                    W_out[ok, :] = W_out[ok, :] / rs[ok]
    else:
        if H_out.size:
            cs = H_out.sum(axis=0, keepdims=True)
            dev = np.max(np.abs(cs - 1.0))
            if np.isfinite(dev) and dev > dev_tol:
                ok = (cs > tiny).ravel()
                if ok.any():
                    H_out[:, ok] = H_out[:, ok] / cs[:, ok]
    return W_out, H_out


# ---------------------------------------------------------------------------
# transform / score (``_base.py:162-265``)
# ---------------------------------------------------------------------------
def transform(X, components, mask=None, W0=None, n_steps=50, rng=None):
    """Fixed-H W-solver of ``NBMFMM.transform`` (``_base.py:162-199``).

    ``W0`` (m x k) replaces the reference's draw from the unseeded global RNG
    (``_base.py:175``) so that the CUDA path can be compared on identical inits.
    """
    X = np.asarray(X, dtype=np.float64)
    m, n = X.shape
    k = components.shape[0]
    if W0 is None:
        W0 = (rng or np.random).uniform(0.1, 0.9, (m, k))
    Wt = np.asarray(W0, dtype=np.float64).T
    for _ in range(n_steps):                          # :178-193
        thetaT = components.T @ Wt
        if mask is None:
            posT, negT = X.T, (1 - X).T
        else:
            posT, negT = X.T * mask.T, (1 - X).T * mask.T
        Wt = Wt * (components @ (posT / (thetaT + 1e-8))
                   + (1 - components) @ (negT / (1 - thetaT + 1e-8)))
        Wt = Wt / n
        Wt = Wt / Wt.sum(axis=0, keepdims=True)
    W = np.clip(Wt.T, 1e-8, 1.0)                      # :196
    return W / W.sum(axis=1, keepdims=True)           # :198


def inverse_transform(W, components):
    """``_base.py:201-210``."""
    return np.clip(W @ components, 0.0, 1.0)


def mean_loglik(X, X_recon, mask=None, eps=1e-8):
    """The NLL part of ``NBMFMM.score`` (``_base.py:238-247``), quirk-masked."""
    if mask is None:
        ll = X * np.log(X_recon + eps) + (1 - X) * np.log(1 - X_recon + eps)
        n_obs = X.size
    else:
        pos = X * mask
        ll = pos * np.log(X_recon + eps) + (1 - pos) * np.log(1 - X_recon + eps)
        n_obs = np.count_nonzero(mask)
    return np.sum(ll) / n_obs


def heldout_perplexity(Y, Y_hat, mask=None, eps=1e-8):
    """Properly-masked perplexity of ``examples/reproduce_magron2022.py:40-47``."""
    if mask is None:
        mask = np.ones_like(Y)
    ll = Y * np.log(Y_hat + eps) + (1 - Y) * np.log(1 - Y_hat + eps)
    return np.exp(-np.sum(mask * ll) / np.count_nonzero(mask))


# ---------------------------------------------------------------------------
# n_init restarts (unpinned: README.md:133,144 "keep the best NLL")
# ---------------------------------------------------------------------------
def fit_restarts(Y, n_components, n_init, random_state, **kw):
    """Restart r uses seed ``random_state + r``; r = 0 reproduces the
    reference.  Best = lowest final loss; ties keep the earliest restart."""
    best = None
    for r in range(n_init):
        seed = None if random_state is None else random_state + r
        out = fit(Y, n_components, random_state=seed, **kw)
        if best is None or out[2][-1] < best[2][-1]:
            best = out
    return best
