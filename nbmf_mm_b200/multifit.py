"""Many independent fits of ONE data set on one GPU: restarts (``n_init``), alpha/beta grids, K sweeps.

The reference runs such sweeps as a Python loop of full solver calls (``examples/reproduce_magron2022.py:87-117,
252-285``; README.md:133,144 for ``n_init``).  Here the data planes are validated, packed and uploaded ONCE and the
fits run concurrently, each on its own CUDA stream with its own device-resident loop (loss, stop rule and ``n_iter``
live on the device, so a fit never needs the host between its first and last kernel): small problems, whose kernels
occupy a few of the 148 SMs, overlap instead of queueing behind each other.  Every job returns exactly what
``nbmf_mm_solver`` returns for the same arguments (same RNG stream per job, bit-identical factors)."""
from __future__ import annotations

import threading
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from .device import require_cuda
from .solver import _CANON, make_problem, prepare_data, resolve_engine


def _draw_inits(random_state, m, n, k, W_init, H_init, transpose):
    """RNG side effects and draw order of ``_solver.py:102-103,122-129`` (internal orientation m, n)."""
    if random_state is not None:
        np.random.seed(random_state)
    if transpose and W_init is not None and H_init is not None:
        W_init, H_init = np.asarray(H_init).T, np.asarray(W_init).T
    if W_init is None:
        W_init = np.random.uniform(0.1, 0.9, (m, k))
    if H_init is None:
        H_init = np.random.uniform(0.1, 0.9, (k, n))
    return np.asarray(W_init, dtype=np.float64), np.asarray(H_init, dtype=np.float64)


def draw_all_inits(jobs, m, n, transpose, n_threads=None):
    """Inits of every job, as the loop of solver calls would draw them (``_draw_inits`` per job, in order), including the
    state it leaves the global NumPy stream in.  When every job is seeded (the usual case: restart r uses random_state +
    r) the jobs are independent streams and are drawn concurrently from private generators -- NumPy releases the GIL, and
    64 x (1226 x 32 + 32 x 285) doubles cost 30 ms on one thread, as much as the fits themselves on the tensor engine.
    Unseeded jobs continue the global stream and keep the sequential loop."""
    need = [i for i, j in enumerate(jobs) if j.get("W_init") is None or j.get("H_init") is None]
    if len(need) < 2 or any(j.get("random_state") is None for j in jobs):
        return [_draw_inits(j.get("random_state"), m, n, int(j["n_components"]), j.get("W_init"), j.get("H_init"), transpose)
                for j in jobs]

    def draw(i):
        j = jobs[i]
        rs = np.random.RandomState(j["random_state"])
        k = int(j["n_components"])
        W_i, H_i = j.get("W_init"), j.get("H_init")
        if W_i is None:
            W_i = rs.uniform(0.1, 0.9, (m, k))
        if H_i is None:
            H_i = rs.uniform(0.1, 0.9, (k, n))
        return np.asarray(W_i, dtype=np.float64), np.asarray(H_i, dtype=np.float64), rs

    import os
    out = [None] * len(jobs)
    last_rs = None
    workers = n_threads or max(1, (os.cpu_count() or 2) // 2)
    with ThreadPoolExecutor(max_workers=min(len(need), workers)) as pool:
        for i, (W_i, H_i, rs) in zip(need, pool.map(draw, need)):
            out[i] = (W_i, H_i)
            if i == len(jobs) - 1:
                last_rs = rs
    for i, j in enumerate(jobs):
        if out[i] is None:                                   # both inits given: no draw (the swap of _solver.py:122-125)
            W_i, H_i = j["W_init"], j["H_init"]
            if transpose:
                W_i, H_i = np.asarray(H_i).T, np.asarray(W_i).T
            out[i] = (np.asarray(W_i, dtype=np.float64), np.asarray(H_i, dtype=np.float64))
    # the global stream ends where the LAST job leaves it: seeded, then its draws (if it drew)
    if last_rs is not None:
        np.random.set_state(last_rs.get_state())
    else:
        np.random.seed(jobs[-1]["random_state"])
    return out


_BATCH_STREAMS = {}


def _batch_stream(dev):
    import torch
    key = (dev.type, dev.index if dev.index is not None else torch.cuda.current_device())
    if key not in _BATCH_STREAMS:
        _BATCH_STREAMS[key] = torch.cuda.Stream(device=dev)
    return _BATCH_STREAMS[key]


_WORKER_STREAMS = {}


def _worker_streams(dev, n):
    """n streams for the worker threads of one call, the same ones every call."""
    import torch
    key = (dev.type, dev.index if dev.index is not None else torch.cuda.current_device())
    pool = _WORKER_STREAMS.setdefault(key, [])
    while len(pool) < n:
        pool.append(torch.cuda.Stream(device=dev))
    return list(pool[:n])


class HeldOut:
    """Held-out evaluation of fitted factors without leaving the device: the properly masked
    mean negative log-likelihood per selected entry and its exponential under each of ``eval_masks`` (name -> 0/1 mask
    of Y's shape) -- ``compute_perplexity(Y, W @ H, mask)`` of the reference's experiment driver
    (``examples/reproduce_magron2022.py:40-47``, eps = 1e-8) -- through the fused objective kernel with strict mask
    semantics and a flat prior; Theta is never materialised.  The planes (Y & mask, mask) of every evaluation mask are
    packed and uploaded once per call, in the fits' internal orientation, so that the factors go from the fit's context
    to the evaluation context as device tensors."""

    def __init__(self, Y, eval_masks, *, transpose, dtype, device):
        self.dtype, self.device = dtype, device
        self.sets = {}
        for name, em in dict(eval_masks).items():
            if em is None:
                raise ValueError(f"eval_masks[{name!r}] is None: an evaluation mask selects the held-out entries")
            d = prepare_data(Y, em, transpose=transpose, dtype=dtype, device=device)
            if d.vkind != "bits":
                raise ValueError("held-out evaluation needs binary Y and 0/1 evaluation masks")
            self.sets[name] = d

    def contexts(self, k):
        """One evaluation context per mask for factors with ``k`` components (on the current stream); the caller closes
        them (``close``)."""
        return {name: make_problem(d, k, dtype=self.dtype, alpha=1.0, beta=1.0, eps=1e-8, mask_semantics="strict",
                                   projection="normalize", max_iter_cap=1, device=self.device, engine="simt")
                for name, d in self.sets.items()}

    def score(self, ctxs, Wd, Hd):
        """``{name: {"nll", "perplexity", "n_entries"}}`` of the internal-orientation factors ``Wd`` (m x k), ``Hd``
        (k x n) -- device tensors."""
        out = {}
        for name, prob in ctxs.items():
            prob.set_factors(Wd, Hd, normalize_w=False)
            nll = float(prob.objective())
            out[name] = {"nll": nll, "perplexity": float(np.exp(nll)), "n_entries": int(self.sets[name].n_obs)}
        return out

    @staticmethod
    def close(ctxs):
        for prob in ctxs.values():
            prob.close()


def nbmf_mm_multifit(Y, jobs, *, mask=None, orientation="beta-dir", max_iter=500, tol=1e-5, eps=1e-8,
                     projection_method="normalize", mask_semantics="reference", dtype="float64", device=None,
                     engine="auto", dense_storage=None, n_streams=None, stats=None, check_range=False, batch=True,
                     batch_plan="fit", verbose=0, eval_masks=None):
    """Fit ``len(jobs)`` models to the same ``Y`` / ``mask``.

    ``jobs``: sequence of dicts with ``n_components`` and optionally ``alpha``, ``beta`` (default 1.2),
    ``random_state``, ``W_init``, ``H_init``, ``max_iter``, ``tol`` (defaults: the keyword arguments).  Returns a
    list of ``(W, H, losses, 0.0, n_iter)`` in job order, each identical to
    ``nbmf_mm_solver(Y, mask=mask, orientation=orientation, **job)``.  ``batch``: jobs with the same K, max_iter and
    tol (restarts, alpha / beta grids) advance together on small problems, one launch per kernel for the whole group
    (``nbmf_batch_bind``).  ``batch_plan``: "fit" (default) keeps the launch plan of a single fit, so every result is
    bit-identical to the solver call whatever the batch size (and the number of GPUs the restarts are dealt to);
    "batch" plans the launches of a group for the whole group -- the batch fills the SMs, so every fit is cut into
    fewer, larger row / column splits (config 5: 64 restarts 25 % faster, profiles/r02_batch_plan_cfg5.txt); the split
    partials are then summed in a different order, so results agree with the solver call to rounding (fp64 ~1e-15, fp32
    ~1e-7) instead of bit for bit.
    ``n_streams``: concurrent fits for the remaining jobs
    (default: one per hardware queue, 8..32, for problems up to 2^24 entries, else 1: a large fit fills the GPU on its own).
    ``eval_masks``: ``{name: mask}`` of held-out entry sets (validation / test splits; ``mask`` itself may be listed to
    get the training perplexity).  Every result then carries a sixth element ``{name: {"nll", "perplexity",
    "n_entries"}}``: the properly masked perplexity of the job's returned factors (``HeldOut``), computed on the device
    right after the fit -- the train -> validate loop of ``examples/reproduce_magron2022.py:87-117`` without a host
    round trip of the factors or an M x N reconstruction."""
    import torch
    if orientation not in _CANON:
        raise ValueError(f"Unknown orientation: {orientation}. Must be one of {list(_CANON)}")
    if batch_plan not in ("batch", "fit"):
        raise ValueError(f"batch_plan must be 'batch' or 'fit', got {batch_plan!r}")
    jobs = [dict(j) for j in jobs]
    if not jobs:
        return []
    dev = require_cuda(device)
    transpose = orientation == "dir-beta"
    data = prepare_data(Y, mask, transpose=transpose, dtype=dtype, device=device, dense_storage=dense_storage,
                        check_range=check_range)
    m, n = data.m, data.n
    held = HeldOut(Y, eval_masks, transpose=transpose, dtype=dtype, device=device) if eval_masks else None
    if n_streams is None:                                    # one stream per hardware queue, see __init__.py
        import os
        try:
            queues = int(os.environ.get("CUDA_DEVICE_MAX_CONNECTIONS", "8"))
        except ValueError:
            queues = 8
        n_streams = max(8, min(queues, 32)) if m * n <= (1 << 24) else 1
    n_streams = max(1, min(int(n_streams), len(jobs)))
    for j in jobs:
        if int(j.get("max_iter", max_iter)) < 1:
            raise UnboundLocalError("max_iter must be >= 1")
    inits = draw_all_inits(jobs, m, n, transpose)
    prepared = []
    for j, (W0, H0) in zip(jobs, inits):
        prepared.append((int(j["n_components"]), float(j.get("alpha", 1.2)), float(j.get("beta", 1.2)),
                         int(j.get("max_iter", max_iter)), float(j.get("tol", tol)), W0, H0))

    main_stream = torch.cuda.current_stream(dev)
    ready = torch.cuda.Event()
    ready.record(main_stream)                                # the data planes were produced on this stream
    local = threading.local()
    free_streams = _worker_streams(dev, n_streams)           # cached per device: see _batch_stream

    def run(idx):
        k, alpha, beta, mi, tl, W0, H0 = prepared[idx]
        if not hasattr(local, "stream"):
            local.stream = free_streams.pop()                # one per worker thread of this call (list.pop is atomic)
        with torch.cuda.stream(local.stream):
            local.stream.wait_event(ready)
            prob = make_problem(data, k, dtype=dtype, alpha=alpha, beta=beta, eps=eps, mask_semantics=mask_semantics,
                                projection=projection_method, max_iter_cap=mi, device=device, engine=engine)
            try:
                prob.set_factors(W0, H0, normalize_w=True)
                losses_arr, n_iter, converged = prob.fit(mi, tl)
                dv = prob.simplex_deviation()                # solver tail, _solver.py:192-213
                renorm = bool(np.isfinite(dv) and dv > 1e-9)
                ho = None
                if held is None:
                    W, H = prob.get_factors_f64(normalize_w=renorm)
                else:                                        # the returned factors, scored where they are
                    Wd = torch.empty((m, k), dtype=torch.float64, device=dev)
                    Hd = torch.empty((k, n), dtype=torch.float64, device=dev)
                    prob.get_factors_f64_device(Wd, Hd, normalize_w=renorm)
                    ctxs = held.contexts(k)
                    try:
                        ho = held.score(ctxs, Wd, Hd)
                    finally:
                        held.close(ctxs)
                    W, H = Wd.cpu().numpy(), Hd.cpu().numpy()
                eng = prob.engine
            finally:
                prob.close()
        if transpose:                                        # _solver.py:182-184
            W, H = np.ascontiguousarray(H.T), np.ascontiguousarray(W.T)
        return W, H, [np.float64(v) for v in losses_arr], 0.0, n_iter, converged, eng, ho

    def run_batch(idxs):
        """Fits with the same K and hyper-parameters advance TOGETHER: contexts with workspaces at a uniform stride in one
        allocation, one launch per kernel for the whole group (``nbmf_batch_bind``; both engines).  Returns the results in
        the order of ``idxs``."""
        k, _, _, mi, tl = prepared[idxs[0]][:5]
        B = len(idxs)
        # engine="auto" gives a SINGLE small fit to the persistent small-fit kernel (latency: the GPU is otherwise idle); a
        # batch fills the GPU, and there the tensor engine wins by 3-5x wherever it is eligible (one launch per kernel for
        # all fits).  Results then differ from a solver call with engine="auto" by fp32 rounding; pass an engine for bits.
        beng = resolve_engine(engine, dtype=dtype, vkind=data.vkind, k=k, eps=eps, m_total=m, n=n)
        big = {}

        def slice_of(b):
            def provide(nbytes):
                if "t" not in big:
                    big["S"] = (nbytes + 255) // 256 * 256
                    big["t"] = torch.empty(big["S"] * B, dtype=torch.uint8, device=dev)
                return big["t"][b * big["S"]: b * big["S"] + nbytes]
            return provide

        # not the default stream (the loop is replayed from a graph), and the SAME stream on every call: torch's caching
        # allocator keeps one pool per stream, so a fresh stream per call finds none of the blocks earlier calls released
        # and goes back to cudaMalloc for half a gigabyte of workspaces (sporadic 100 ms stalls in any phase)
        stream = _batch_stream(dev)
        probs, out = [], []
        tdt = getattr(torch, np.dtype(dtype).name)
        import os, time
        timing = os.environ.get("NBMF_MULTIFIT_TIMING")
        marks = [("start", time.perf_counter())]

        def mark(name):
            if timing:
                torch.cuda.synchronize(dev)
                marks.append((name, time.perf_counter()))
        with torch.cuda.stream(stream):
            stream.wait_event(ready)
            try:
                # the inits of the whole batch cross PCIe in two copies, the results in two (per-fit copies and their
                # synchronisations cost more than the fits themselves once a batch of small fits runs on the tensor engine)
                # (pinned staging in the compute dtype, filled by a few threads: stacking 64 fp64 inits on one thread and
                # copying them from pageable memory cost 15 ms at K = 64, their device time is under 2 ms)
                W0p = torch.empty((B, m, k), dtype=tdt, pin_memory=True)
                H0p = torch.empty((B, k, n), dtype=tdt, pin_memory=True)
                W0v, H0v = W0p.numpy(), H0p.numpy()

                def stage(b):
                    np.copyto(W0v[b], prepared[idxs[b]][5], casting="same_kind")
                    np.copyto(H0v[b], prepared[idxs[b]][6], casting="same_kind")
                with ThreadPoolExecutor(max_workers=min(B, 8)) as pool:
                    list(pool.map(stage, range(B)))
                W0s, H0s = W0p.to(dev, non_blocking=True), H0p.to(dev, non_blocking=True)
                mark("inits uploaded")
                for b, idx in enumerate(idxs):
                    prob = make_problem(data, k, dtype=dtype, alpha=prepared[idx][1], beta=prepared[idx][2], eps=eps,
                                        mask_semantics=mask_semantics, projection=projection_method, max_iter_cap=mi,
                                        device=device, engine=beng, workspace=slice_of(b),
                                        batch_hint=B if batch_plan == "batch" else 0)
                    probs.append(prob)
                    prob.set_factors(W0s[b], H0s[b], normalize_w=True)
                    prob.fit_begin(mi, tl)
                mark("contexts created")
                leader = probs[0]
                leader.batch_bind(B, big["S"])
                chunk, n_iters = 16, [0] * B
                for _ in range(mi + 8):                      # every pass enqueues >= 1 iteration until the tail is in
                    leader.fit_enqueue(chunk)
                    done, n_iters = leader.batch_poll()
                    if done:
                        break
                    chunk = min(chunk * 2, 256)
                else:
                    raise RuntimeError("batched fit loop did not terminate")
                mark("loop")
                hist, conv, devs = leader.batch_tail(max(n_iters))   # one synchronisation for all fits
                leader.batch_bind(1, 0)
                Wd = torch.empty((B, m, k), dtype=torch.float64, device=dev)
                Hd = torch.empty((B, k, n), dtype=torch.float64, device=dev)
                meta = []
                for b, (prob, n_iter) in enumerate(zip(probs, n_iters)):
                    dv = devs[b]                             # solver tail, _solver.py:192-213
                    prob.get_factors_f64_device(Wd[b], Hd[b], normalize_w=bool(np.isfinite(dv) and dv > 1e-9))
                    meta.append(([np.float64(v) for v in hist[b, :n_iter]], n_iter, conv[b], prob.engine))
                mark("tails")
                held_out = [None] * B
                if held is not None:                         # one evaluation context per mask for the whole group
                    ctxs = held.contexts(k)
                    try:
                        held_out = [held.score(ctxs, Wd[b], Hd[b]) for b in range(B)]
                    finally:
                        held.close(ctxs)
                    mark("held-out")
                if transpose:                                # _solver.py:182-184, on the device
                    Wd, Hd = Hd.transpose(1, 2).contiguous(), Wd.transpose(1, 2).contiguous()
                # one DMA each into pinned staging, then a threaded copy into the arrays the caller keeps.  (A plain .cpu()
                # of 50 MB at K = 64 costs 19 ms: pageable D2H + page faults on one thread.  Handing out views of the pinned
                # blocks instead ties their lifetime to the caller's, and every call then page-locks fresh memory: 10-100 ms,
                # highly variable.  The staging blocks go back to torch's pinned cache when this function returns.)
                Wp = torch.empty(Wd.shape, dtype=torch.float64, pin_memory=True)
                Hp = torch.empty(Hd.shape, dtype=torch.float64, pin_memory=True)
                Wp.copy_(Wd, non_blocking=True)
                Hp.copy_(Hd, non_blocking=True)
                Wh, Hh = np.empty(tuple(Wd.shape)), np.empty(tuple(Hd.shape))
                stream.synchronize()
                Wpn, Hpn = Wp.numpy(), Hp.numpy()

                def unstage(b):
                    np.copyto(Wh[b], Wpn[b])
                    np.copyto(Hh[b], Hpn[b])
                with ThreadPoolExecutor(max_workers=min(B, 8)) as pool:
                    list(pool.map(unstage, range(B)))
                mark("results downloaded")
                for b, (losses, n_iter, converged, eng) in enumerate(meta):
                    out.append((Wh[b], Hh[b], losses, 0.0, n_iter, converged, eng, held_out[b]))
            finally:
                for prob in probs:
                    prob.close()
        if timing:
            marks.append(("closed", time.perf_counter()))
            print("[multifit batch of %d, K=%d] " % (B, k) + ", ".join(f"{n} {1e3 * (t - marks[i][1]):.1f} ms" for i, (n, t) in enumerate(marks[1:])),
                  flush=True)
        return out

    # groups of fits that can advance together (same K, max_iter, tol: restarts AND alpha / beta grids -- the Beta prior
    # lives in each fit's device-side state) on small problems
    results = [None] * len(jobs)
    batched = 0
    if batch and m * n <= (1 << 24) and data.vkind == "bits":
        groups = {}
        for i, pj in enumerate(prepared):
            groups.setdefault((pj[0], pj[3], pj[4]), []).append(i)
        for idxs in groups.values():
            if len(idxs) < 2:
                continue
            for c0 in range(0, len(idxs), 256):               # bounded batches: workspace memory, gridDim.z
                part = idxs[c0:c0 + 256]
                if len(part) < 2:
                    continue
                got = run_batch(part)
                if got is None:
                    break
                for i, r in zip(part, got):
                    results[i] = r
                batched += len(part)
    rest = [i for i, r in enumerate(results) if r is None]
    if n_streams == 1 or len(rest) <= 1:
        for i in rest:
            results[i] = run(i)
    else:
        with ThreadPoolExecutor(max_workers=min(n_streams, len(rest))) as pool:
            for i, r in zip(rest, pool.map(run, rest)):
                results[i] = r
    if verbose > 0:                                          # the lines of _solver.py:165-166,172-173, per job
        for i, r in enumerate(results):
            for it in range(0, r[4], 10):
                print(f"[fit {i}] Iter {it:4d}: Loss = {r[2][it]:.6f}")
            if r[5]:
                print(f"[fit {i}] Converged at iteration {r[4] - 1}")
    if stats is not None:
        stats.update(h2d_bytes=data.h2d_bytes, n_streams=n_streams, engine=results[0][6], batched=batched,
                     converged=[r[5] for r in results])
    if held is not None:
        return [r[:5] + (r[7],) for r in results]
    return [r[:5] for r in results]
