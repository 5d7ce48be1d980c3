"""scikit-learn style estimator, drop-in for ``nbmf_mm.NBMFMM`` / ``nbmf_mm.NBMF``
(reference ``src/nbmf_mm/_base.py``), backed by the B200 kernels.

Same constructor arguments, attributes, error messages and RNG side effects as the
reference, plus the keyword arguments its README advertises but its code lacks
(``projection_method``, ``n_init``; README.md:124-144) and the device knobs
(``dtype``, ``mask_semantics``, ``device``, ``distributed``).
"""
from __future__ import annotations

import numpy as np
from sklearn.base import BaseEstimator, TransformerMixin
from sklearn.utils import check_array

from ._utils import check_is_fitted
from .bits import BitMatrix
from .device import reconstruct_device
from .multifit import nbmf_mm_multifit
from .solver import make_problem, nbmf_mm_solver, prepare_data

# exact-key alias table of the reference (_base.py:127-137); keys are NOT case-folded
_ORIENTATION_ALIASES = {
    "beta-dir": "beta-dir",
    "dir-beta": "dir-beta",
    "Beta-Dir": "beta-dir",
    "Dir-Beta": "dir-beta",
    "Dir Beta": "dir-beta",
    "binary ICA": "beta-dir",
    "Binary ICA": "beta-dir",
    "bICA": "beta-dir",
    "Aspect Bernoulli": "dir-beta",
}


def partition_restarts(n_init, rank, world):
    """Restarts of this rank: r = rank, rank + world, ... (independent fits, no collective)."""
    return list(range(int(rank), int(n_init), int(world)))


def reduce_best_restart(local_best, group=None):
    """``local_best`` = (final_loss, restart_index, payload) of this rank's best restart, or None if it ran
    none.  Every rank gets (payload, restart_index) of the globally best restart: lowest final loss, ties to
    the lowest restart index (= what a sequential loop over the restarts keeps).  Uses only object
    collectives of torch.distributed, so it runs on any backend (gloo in the CPU tests)."""
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    scores = [None] * world
    dist.all_gather_object(scores, None if local_best is None else (float(local_best[0]), int(local_best[1])), group=group)
    ranked = sorted((s[0], s[1], r) for r, s in enumerate(scores) if s is not None)
    if not ranked:
        raise ValueError("no rank ran a restart")
    _, best_idx, owner = ranked[0]
    box = [local_best[2] if rank == owner else None]
    dist.broadcast_object_list(box, src=dist.get_global_rank(group, owner) if group is not None else owner, group=group)
    return box[0], best_idx


class NBMFMM(BaseEstimator, TransformerMixin):
    """Non-negative Binary Matrix Factorization via Majorization-Minimization on B200.

    Parameters (reference ``_base.py:63-66`` first, extensions after ``orientation``)
    ----------
    n_components, alpha, beta, max_iter, tol, W_init, H_init, init, random_state, verbose,
    orientation : as in the reference (``init`` is accepted and ignored there too).
    projection_method : {"normalize", "duchi"}, default "normalize"
        Simplex step of the W update: multiplicative + L1 renormalisation (the reference's
        only behaviour) or Euclidean projection (Duchi et al. 2008; README-only, unpinned).
    n_init : int, default 1
        Random restarts; restart r uses ``random_state + r`` and the lowest final loss wins.
    dtype : {"float64", "float32"}, default "float64"
        Arithmetic type of the device path.  float64 reproduces the reference to ~1e-12;
        float32 is the throughput mode (packed FFMA2).
    mask_semantics : {"reference", "strict"}, default "reference"
        "reference" reproduces the reference's H-step/loss treatment of unobserved entries as
        observed zeros (``_solver.py:43,153-154``); "strict" is the README/paper behaviour.
    device : torch device or None.
    distributed : {False, True, "rows", "restarts"}
        True / "rows": one problem row-sharded over the ranks of torch.distributed (one rank per GPU), H
        partials summed with an NCCL allreduce per iteration.  "restarts": the ``n_init`` restarts are dealt
        round-robin to the ranks (restart r on rank r % world, whole problem on that rank's GPU, no
        data-path collective); the final losses are compared across ranks and the winner's factors are
        broadcast, so every rank ends with the same fitted estimator.
    dense_storage : {None, "float16"}: device layout of probabilistic X (values strictly inside (0,1)); None = the
        compute dtype, "float16" halves the bytes each pass reads (float32 arithmetic only).
    engine : {"auto", "simt", "tensor", "fused"}: CUDA-core (packed FFMA2) kernels, the tcgen05/TMEM split-precision (TF32 +
        bf16) kernels (float32, binary X, K <= 64), or the persistent small-fit kernel (binary X, K <= 32: whole iterations in
        one launch).  "auto": a single small fit -> fused; else tensor when eligible and m >= 512, n >= 128; else simt.
        Restarts (``n_init`` > 1) advance as one batch on tensor / simt; pin an engine to get bit-identical single fits.
    """

    def __init__(self, n_components=10, alpha=1.2, beta=1.2, max_iter=2000, tol=1e-5,
                 W_init=None, H_init=None, init=None, random_state=None, verbose=0,
                 orientation="beta-dir", projection_method="normalize", n_init=1,
                 dtype="float64", mask_semantics="reference", device=None, distributed=False, engine="auto",
                 dense_storage=None):
        self.n_components = n_components
        self.alpha = alpha
        self.beta = beta
        self.max_iter = max_iter
        self.tol = tol
        self.W_init = W_init
        self.H_init = H_init
        self.init = init
        self.random_state = random_state
        self.verbose = verbose
        self.orientation = orientation
        self.projection_method = projection_method
        self.n_init = n_init
        self.dtype = dtype
        self.mask_semantics = mask_semantics
        self.device = device
        self.distributed = distributed
        self.engine = engine
        self.dense_storage = dense_storage

    # ------------------------------------------------------------------ helpers
    def _normalize_orientation(self, orientation):
        if orientation in _ORIENTATION_ALIASES:
            return _ORIENTATION_ALIASES[orientation]
        raise ValueError(f"Unknown orientation: {orientation}. "
                         f"Must be one of {list(_ORIENTATION_ALIASES.keys())}")

    @staticmethod
    def _validate_X(X):
        if isinstance(X, BitMatrix):
            return X
        # _base.py:83; CSR stays CSR (packed on the device).  Large dense X: sklearn's finiteness test is a host pass of
        # 70 ms per 10^8 entries; NaN and inf fail the range test of the device pass that packs X instead (fit() then
        # re-runs the host test to raise sklearn's message, as the reference would have)
        big = isinstance(X, np.ndarray) and X.size >= (1 << 22)
        return check_array(X, accept_sparse="csr", dtype=np.float64, ensure_all_finite=not big)

    # ------------------------------------------------------------------ fit
    def fit(self, X, y=None, mask=None):
        """Fit the model to binary (or [0,1]-valued) data ``X`` (``_base.py:80-122``)."""
        X = self._validate_X(X)
        # _base.py:90-91 "X must be binary".  Small X: tested here, before any device work.  Large X (three NumPy passes
        # cost 0.3 s per 10^8 entries): tested on the device in the pass that packs X (check_range below); the
        # reference raises this error before the orientation error, so that order is kept
        host_check = isinstance(X, np.ndarray) and X.size < (1 << 22)
        if host_check and not np.all((X >= 0) & (X <= 1)):
            raise ValueError("X must be binary")
        try:
            orientation = self._normalize_orientation(self.orientation)
        except ValueError:
            if isinstance(X, np.ndarray) and not host_check and not np.all((X >= 0) & (X <= 1)):
                raise ValueError("X must be binary") from None
            raise
        self.orientation = orientation                                 # reference mutates it too (_base.py:95)
        if self.projection_method not in ("normalize", "duchi"):
            raise ValueError(f"projection_method must be 'normalize' or 'duchi', got {self.projection_method!r}")
        n_init = int(self.n_init)
        if n_init < 1:
            raise ValueError("n_init must be >= 1")

        if self.distributed not in (False, True, "rows", "restarts"):
            raise ValueError(f"distributed must be False, True, 'rows' or 'restarts', got {self.distributed!r}")
        rank, world = 0, 1
        by_restart = self.distributed == "restarts"
        if by_restart:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                rank, world = dist.get_rank(), dist.get_world_size()
        row_sharded = self.distributed in (True, "rows")

        try:
            return self._fit_checked(X, mask, orientation, n_init, by_restart, row_sharded, rank, world)
        except ValueError as e:
            if str(e) == "X must be binary" and isinstance(X, np.ndarray) and not host_check:
                check_array(X, dtype=np.float64)                       # NaN / inf: sklearn's error comes first (_base.py:83)
            raise

    def _fit_checked(self, X, mask, orientation, n_init, by_restart, row_sharded, rank, world):
        best = None
        mine = partition_restarts(n_init, rank, world) if by_restart else list(range(n_init))
        if n_init > 1 and not row_sharded:
            # restarts share ONE upload of X / mask and run concurrently on separate streams (multifit.py)
            stats = {}
            jobs = [dict(n_components=self.n_components, alpha=self.alpha, beta=self.beta, W_init=self.W_init,
                         H_init=self.H_init, random_state=None if self.random_state is None else self.random_state + r)
                    for r in mine]
            outs = nbmf_mm_multifit(X, jobs, mask=mask, orientation=orientation, max_iter=self.max_iter, tol=self.tol,
                                    projection_method=self.projection_method, mask_semantics=self.mask_semantics,
                                    dtype=self.dtype, device=self.device, engine=self.engine,
                                    dense_storage=self.dense_storage, stats=stats, check_range=True,
                                    verbose=self.verbose)
            for r, out in zip(mine, outs):
                if best is None or out[2][-1] < best[0][2][-1]:
                    best = (out, stats, r)
            mine = []
        for r in mine:
            seed = self.random_state if (self.random_state is None or n_init == 1) else self.random_state + r
            stats = {}
            out = nbmf_mm_solver(
                Y=X, n_components=self.n_components, max_iter=self.max_iter, tol=self.tol,
                alpha=self.alpha, beta=self.beta, W_init=self.W_init, H_init=self.H_init, mask=mask,
                random_state=seed, verbose=self.verbose, orientation=orientation,
                projection_method=self.projection_method, mask_semantics=self.mask_semantics,
                dtype=self.dtype, device=self.device, distributed=row_sharded, stats=stats,
                engine=self.engine, dense_storage=self.dense_storage, check_range=True)
            if best is None or out[2][-1] < best[0][2][-1]:
                best = (out, stats, r)
        if by_restart and world > 1:
            local = None if best is None else (best[0][2][-1], best[2], (best[0], best[1]))
            (out, stats), best_r = reduce_best_restart(local)
            best = (out, stats, best_r)
        (W, H, losses, _, n_iter), stats, best_r = best

        self.W_ = W
        self.components_ = H
        self.loss_curve_ = losses
        self.objective_history_ = losses
        self.loss_ = losses[-1] if losses else np.inf
        self.n_iter_ = n_iter
        self.reconstruction_err_ = float(losses[-1]) if losses else np.inf
        self.best_init_ = best_r
        self.transfer_stats_ = stats
        return self

    def fit_transform(self, X, y=None):
        self.fit(X)
        return self.W_

    # ------------------------------------------------------------------ transform & co
    def _fixed_h_problem(self, X, mask, n_obs=None):
        data = prepare_data(X, mask, transpose=False, dtype=self.dtype, device=self.device)
        if data.n != self.components_.shape[1]:
            raise ValueError(f"X has {data.n} features, the model was fitted with {self.components_.shape[1]}")
        prob = make_problem(data, self.n_components, dtype=self.dtype, alpha=1.0, beta=1.0, eps=1e-8,
                            mask_semantics="reference", projection="normalize", max_iter_cap=1,
                            device=self.device, n_obs=n_obs, engine=self.engine)
        return data, prob

    def _transform_device(self, X, mask):
        """50 fixed-H W half-steps from W ~ U(0.1, 0.9) drawn from the global NumPy RNG, then clip
        and row-normalise -- ``_base.py:162-199``.  Always the beta-dir W step, whatever the
        orientation, exactly like the reference."""
        data, prob = self._fixed_h_problem(X, mask)
        try:
            W0 = np.random.uniform(0.1, 0.9, (data.m, self.n_components))      # _base.py:175
            prob.set_factors(W0, self.components_, normalize_w=False)
            prob.transform(50)
            W, _ = prob.get_factors()
        finally:
            prob.close()
        return W

    def transform(self, X, mask=None):
        check_is_fitted(self, ["components_"])
        X = self._validate_X(X)
        return self._transform_device(X, mask)

    def inverse_transform(self, W):
        """``clip(W @ components_, 0, 1)`` (``_base.py:201-210``); dense M x N by contract."""
        check_is_fitted(self, ["components_"])
        W = check_array(W, dtype=np.float64)
        if W.shape[1] != self.components_.shape[0]:
            raise ValueError(f"W has {W.shape[1]} components, the model has {self.components_.shape[0]}")
        return reconstruct_device(W, self.components_, self.dtype, self.device)

    def score(self, X, mask=None):
        """Average log-likelihood per observed entry (``_base.py:212-247``).

        As in the reference, ``transform`` is called WITHOUT the mask (``_base.py:235``) and the
        log-likelihood treats unobserved entries as observed zeros while dividing by the number of
        observed ones.  Theta = W.components_ is evaluated on the device and never materialised
        (the reference's clip to [0,1] is a no-op for simplex W and H in (0,1))."""
        check_is_fitted(self, ["components_"])
        X = self._validate_X(X)
        W = self._transform_device(X, None)
        data, prob = self._fixed_h_problem(X, mask)
        try:
            prob.set_factors(W, self.components_, normalize_w=False)
            loss = prob.objective()
        finally:
            prob.close()
        return float(-loss)

    def perplexity(self, X, mask=None):
        return float(np.exp(-self.score(X, mask)))

    def evaluate(self, X, mask=None, W=None):
        """Held-out evaluation of the FITTED factors on the entries selected by ``mask`` (validation / test
        split), entirely on the device: properly masked mean negative log-likelihood per selected entry and its
        exponential, i.e. ``compute_perplexity(Y, W_ @ components_, mask)`` of the reference's experiment
        driver (``examples/reproduce_magron2022.py:40-47``) without materialising Theta.  ``W`` defaults to
        ``W_`` (``X`` must then have the training rows); pass ``transform(X_new, mask=observed)`` for new rows.
        Returns ``{"nll": float, "perplexity": float, "n_entries": int}``."""
        check_is_fitted(self, ["components_"])
        X = self._validate_X(X)
        W = self.W_ if W is None else np.asarray(W, dtype=np.float64)
        data = prepare_data(X, mask, transpose=False, dtype=self.dtype, device=self.device)
        if (data.m, data.n) != (W.shape[0], self.components_.shape[1]):
            raise ValueError(f"X is {data.m} x {data.n}, the factors are {W.shape[0]} x {self.components_.shape[1]}")
        prob = make_problem(data, self.n_components, dtype=self.dtype, alpha=1.0, beta=1.0, eps=1e-8,
                            mask_semantics="strict", projection="normalize", max_iter_cap=1, device=self.device,
                            engine="simt")
        try:
            prob.set_factors(W, self.components_, normalize_w=False)
            nll = float(prob.objective())
        finally:
            prob.close()
        return {"nll": nll, "perplexity": float(np.exp(nll)), "n_entries": int(data.n_obs)}


NBMF = NBMFMM
