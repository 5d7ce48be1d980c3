"""Host-side mirror of the reference solver interface, running on the B200 kernels.

``nbmf_mm_solver`` and ``nbmf_mm_update_beta_dir`` keep the reference's names, argument
meaning, return tuple and side effects (``src/nbmf_mm/_solver.py:5-13,61-75``); everything
between input validation and the returned NumPy arrays happens on the device through the
C-ABI.  There is no CPU fallback.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import contextlib
import threading

import numpy as np

from .bits import BitMatrix
from .device import (PIN_THRESHOLD, DeviceProblem, pack_bits_device, pack_csr_device, pack_dense_device,
                     pack_host_dense_checked, pinned_factor_buffers, require_cuda)

_CANON = ("beta-dir", "dir-beta")


@dataclass
class PreparedData:
    """Device-resident data planes of one problem in INTERNAL orientation."""
    m: int
    n: int
    vkind: str                 # "bits" | "dense"
    P: Optional[BitMatrix]     # V & mask             (bits)
    M: Optional[BitMatrix]     # mask or None
    Vm: object                 # torch tensor V*mask  (dense)
    n_obs: float               # Y.size or count_nonzero(mask), _solver.py:151,155
    h2d_bytes: int = 0
    Wm: object = None          # torch tensor: values of a weighted (non-0/1) mask, dense layout (vkind "dense" only)
    pending: object = None     # deferred device work (P &= M, orientation, n_obs): see finish()

    def finish(self):
        """Run the device-side part of the preparation that was deferred so that the host could do other
        work (drawing the inits) while the asynchronous H2D copies of the bit planes were in flight."""
        if self.pending is not None:
            todo, self.pending = self.pending, None
            todo(self)
        return self


def _densify(a):
    return a.toarray() if hasattr(a, "toarray") else a


_WEIGHTED_BITS = ("a weighted (non-0/1) mask needs dense X / mask arrays: the reference multiplies by the mask values "
                  "(_solver.py:30-32), which the bit-packed inputs cannot carry")


def _as_bool_mask(mask, shape):
    mask = np.asarray(_densify(mask))
    if mask.shape != tuple(shape):
        raise ValueError(f"mask has shape {mask.shape}, expected {tuple(shape)}")
    if mask.dtype != np.bool_:
        if not np.all((mask == 0) | (mask == 1)):
            raise ValueError(_WEIGHTED_BITS)
        mask = mask != 0
    return mask


def prepare_data(Y, mask, *, transpose, dtype, device, defer=False, dense_storage=None, check_range=False) -> PreparedData:
    """Validate, orient, pack and upload V and the observation mask.

    Binary V goes to two 1-bit planes (packed on the host, so the H2D copy is 32-64x smaller
    than the reference's fp64 arrays); V with values strictly inside (0,1) goes to the dense
    ``V*mask`` layout in the compute dtype plus a mask bit plane.  ``transpose`` implements
    dir-beta == beta-dir on V^T (``_solver.py:113-123``).
    """
    import torch
    dev = require_cuda(device)
    extra_h2d = 0
    if hasattr(Y, "tocsr"):
        # sparse X: indptr / indices / data go to the device and are packed there; nothing M x N is ever built
        # (the reference densifies: _base.py:83-87, _solver.py:106-107)
        Pd, flags, extra_h2d = pack_csr_device(Y, dev)
        if flags & 2:
            raise ValueError("X must be binary")                      # _base.py:90-91
        if flags & 1:
            Y = Y.toarray()                                           # values strictly inside (0,1): dense layout
            extra_h2d = 0
        elif (mask is not None and not isinstance(mask, BitMatrix) and not hasattr(mask, "tocsr")
              and np.asarray(mask).dtype != np.bool_ and not np.all((np.asarray(mask) == 0) | (np.asarray(mask) == 1))):
            Y, extra_h2d = Y.toarray(), 0                             # dense weighted mask: the dense layout carries it
        else:
            if mask is not None and hasattr(mask, "tocsr"):
                if mask.shape != Y.shape:
                    raise ValueError(f"mask has shape {mask.shape}, expected {tuple(Y.shape)}")
                mflags = 1 if (mask.data.size and not np.all((mask.data == 0) | (mask.data == 1))) else 0
                if mflags & 1:                                        # weighted sparse mask: the dense layout carries it
                    Y, mask, extra_h2d = Y.toarray(), mask.toarray(), 0
                else:
                    mask, mflags, mh = pack_csr_device(mask, dev)
                    extra_h2d += mh
            if hasattr(Y, "tocsr"):
                Y = Pd
    if isinstance(Y, BitMatrix):
        P = Y
        M = mask
        if M is not None and not isinstance(M, BitMatrix):
            M = BitMatrix.from_dense(_as_bool_mask(M, P.shape))
        h2d = extra_h2d
        if not P.is_device:
            h2d += P.words.nbytes if isinstance(P.words, np.ndarray) else P.words.numel() * 4
            P = P.to_device(dev)
        if M is not None and not M.is_device:
            h2d += M.words.nbytes if isinstance(M.words, np.ndarray) else M.words.numel() * 4
            M = M.to_device(dev)
        m, n = (P.shape[1], P.shape[0]) if transpose else P.shape

        def device_part(d):                                   # everything that waits for the copies
            Pd, Md = d.P, d.M
            if Md is not None:
                Pd = Pd & Md
            if transpose:
                Pd = Pd.transpose()
                Md = Md.transpose() if Md is not None else None
            d.P, d.M = Pd, Md
            d.n_obs = float(Md.count()) if Md is not None else float(d.m) * float(d.n)

        data = PreparedData(m, n, "bits", P, M, None, float("nan"), h2d, pending=device_part)
        return data if defer else data.finish()

    Y = np.asarray(_densify(Y))
    if Y.ndim != 2:
        raise ValueError("Y must be 2-D")
    if mask is not None:
        mask = np.asarray(_densify(mask))
        if mask.shape != Y.shape:
            raise ValueError(f"mask has shape {mask.shape}, expected {tuple(Y.shape)}")
    # dense host arrays: the value checks and the packing run on the device, on row chunks of the arrays as they are
    # (external orientation; dir-beta transposes the packed planes)
    P, M, flags, h2d = pack_host_dense_checked(Y, mask, dev)
    weighted = bool(flags & 4)                                        # mask values other than 0 / 1: dense layout below
    if check_range and (flags & 2):
        raise ValueError("X must be binary")                          # _base.py:90-91
    if not (flags & 1) and not weighted:
        if transpose:
            P = P.transpose()
            M = None if M is None else M.transpose()
        m, n = P.shape
        n_obs = float(m) * float(n) if M is None else float(M.count())
        return PreparedData(m, n, "bits", P, M, None, n_obs, h2d)
    del P, M
    Y = np.asarray(Y, dtype=np.float64)
    mk = None if mask is None else (mask != 0)
    if transpose:
        Y = Y.T
        mk = None if mk is None else mk.T
    m, n = Y.shape
    n_obs = float(Y.size) if mk is None else float(np.count_nonzero(mk))
    # probabilistic V (or a weighted mask): dense V*mask in the compute dtype + mask bits
    Yd = torch.from_numpy(np.ascontiguousarray(Y)).to(dev)
    h2d = Y.nbytes
    md = None
    M = None
    if mk is not None:
        md = torch.from_numpy(np.ascontiguousarray(mk).view(np.uint8)).to(dev)
        h2d += mk.nbytes
        M, _ = pack_bits_device(md, None)
    if weighted:
        # the reference multiplies by the mask VALUES (_solver.py:30-32): V * mask for the H half-step and the loss,
        # (1 - V) * mask = mask - V * mask for the W half-step; the bit plane (mask != 0) only counts the observed entries
        if dense_storage == "float16":
            raise ValueError("a weighted mask needs the dense layout in the compute dtype (dense_storage=None)")
        wv = np.asarray(mask, dtype=np.float64)
        wv = np.ascontiguousarray(wv.T if transpose else wv)
        wd = torch.from_numpy(wv).to(dev)
        h2d += wv.nbytes
        Vm = pack_dense_device(Yd, wd, dtype)
        Wm = pack_dense_device(wd, None, dtype)
        return PreparedData(m, n, "dense", None, M, Vm, n_obs, h2d, Wm=Wm)
    if dense_storage not in (None, "float16", "float32", "float64"):
        raise ValueError(f"dense_storage must be None, 'float16', 'float32' or 'float64', got {dense_storage!r}")
    if dense_storage == "float16":                             # fp16 layout: half the HBM bytes per pass
        if np.dtype(dtype) != np.float32:
            raise ValueError("dense_storage='float16' needs dtype='float32'")
        return PreparedData(m, n, "dense16", None, M, pack_dense_device(Yd, md, np.float16), n_obs, h2d)
    Vm = pack_dense_device(Yd, md, dtype)
    return PreparedData(m, n, "dense", None, M, Vm, n_obs, h2d)


def make_problem(data: PreparedData, k, *, dtype, alpha, beta, eps, mask_semantics, projection, max_iter_cap,
                 device, n_obs=None, engine="auto", workspace=None, batch_hint=0) -> DeviceProblem:
    prob = DeviceProblem(data.m, data.n, k, dtype=dtype, vkind=data.vkind, has_mask=data.M is not None,
                         alpha=alpha, beta=beta, eps=eps, n_obs=data.n_obs if n_obs is None else n_obs,
                         mask_semantics=mask_semantics, projection=projection, max_iter_cap=max_iter_cap,
                         device=device, engine=engine, workspace=workspace, batch_hint=batch_hint)
    if data.vkind == "bits":
        prob.set_bits(data.P, data.M)
    else:
        if data.Wm is not None and mask_semantics == "strict":
            raise ValueError("weighted masks are defined for mask_semantics='reference' only")
        prob.set_dense(data.Vm, data.M, data.Wm)
    return prob


TENSOR_MAX_K = 64            # largest n_components the tcgen05 engine covers (capi.cu make_plan)

_FIT_STREAMS = threading.local()


def _fit_stream(dev):
    """The side stream single-GPU fits run on when the caller is on the default stream: one per device and calling
    thread (a stream that is being captured must not receive another thread's launches), reused from call to call
    (torch's caching allocator pools blocks per stream)."""
    import torch
    pool = _FIT_STREAMS.__dict__.setdefault("streams", {})
    key = dev.index if dev.index is not None else torch.cuda.current_device()
    if key not in pool:
        pool[key] = torch.cuda.Stream(device=dev)
    return pool[key]


def resolve_engine(engine, *, dtype, vkind, k, eps, m_total, n):
    """The engine every rank of a row-sharded fit must use, decided from GLOBAL quantities (``m_total``, not the
    local shard): ranks that disagreed would all-reduce differently shaped [C | D] buffers (the tensor engine pads K to
    its own tile).  Also the engine of a BATCH of fits (multifit.py).  ``make_plan`` in capi.cu applies the same rule to a
    single fit on one GPU, after first giving small problems to the persistent small-fit kernel ("fused")."""
    if engine != "auto":
        return engine
    eligible = np.dtype(dtype) == np.float32 and vkind == "bits" and k <= TENSOR_MAX_K and eps >= 1e-9
    return "tensor" if eligible and m_total >= 512 and n >= 128 else "simt"


def draw_shard_inits(seed, m, n, k, r0, r1, *, h_part=(0, 1), set_global_state=True):
    """Rows [r0, r1) of ``W_init`` and part ``h_part = (i, parts)`` of the flattened ``H_init`` exactly as
    ``np.random.seed(seed); uniform(0.1, 0.9, (m, k)); uniform(0.1, 0.9, (k, n))`` would produce them
    (``_solver.py:102-103,126-129``), without drawing what lies before them: ``nbmf_mt19937_uniform`` jumps ahead in the
    MT19937 stream.  Returns ``(W_rows, H_flat_part, (c0, c1))`` with ``H_init.ravel()[c0:c1] == H_flat_part``.
    ``set_global_state``: leave NumPy's global stream where the reference's two draws leave it."""
    import ctypes as C
    from . import _lib
    lib = _lib.load()
    seed = int(seed) & 0xFFFFFFFF

    def draw(skip, count, state=None):
        out = np.empty(int(count), dtype=np.float64)
        _lib.check(lib.nbmf_mt19937_uniform(seed, int(skip), int(count), 0.1, 0.9, out.ctypes.data,
                                            None if state is None else state.ctypes.data), "nbmf_mt19937_uniform")
        return out

    W_rows = draw(r0 * k, (r1 - r0) * k).reshape(r1 - r0, k)
    i, parts = h_part
    per = (k * n + parts - 1) // parts
    c0, c1 = min(k * n, i * per), min(k * n, (i + 1) * per)
    H_part = draw(m * k + c0, c1 - c0)
    if set_global_state:
        st = np.zeros(625, dtype=np.uint32)
        draw(m * k + k * n, 0, st)
        np.random.set_state(("MT19937", st[:624], int(st[624])))
    return W_rows, H_part, (c0, c1)


def _gather_h_parts(H_part, k, n, world, dtype, dev):
    """Every rank drew 1/world of the flattened H_init: all-gather the parts on the devices (fp64 over NVLink) and
    convert to the compute dtype there.  Returns the (k x n) device tensor ``set_factors`` takes."""
    import torch
    import torch.distributed as dist
    per = (k * n + world - 1) // world
    mine = torch.zeros(per, dtype=torch.float64, device=dev)
    mine[: H_part.size].copy_(torch.from_numpy(H_part), non_blocking=False)
    full = torch.empty(per * world, dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(full, mine)
    return full[: k * n].view(k, n).to(getattr(torch, np.dtype(dtype).name))


def _rank_block(A, r0, r1, transpose):
    """Rows [r0, r1) of the INTERNAL orientation of a host input (dense, scipy sparse or BitMatrix)."""
    if isinstance(A, BitMatrix):
        return A.rows(r0, r1)
    if hasattr(A, "tocsr"):
        A = A.tocsr()
        return A[:, r0:r1] if transpose else A[r0:r1]
    A = np.asarray(A)
    return A[:, r0:r1] if transpose else A[r0:r1]


def _row_shard(m, rank, world):
    """Contiguous row block of this rank (boundaries on multiples of 32 rows)."""
    per = (m + world - 1) // world
    per = (per + 31) // 32 * 32
    r0 = min(m, rank * per)
    return r0, min(m, r0 + per)


def nbmf_mm_solver(Y, n_components, max_iter=500, tol=1e-5, alpha=1.2, beta=1.2, W_init=None, H_init=None,
                   mask=None, random_state=None, verbose=0, orientation="beta-dir", eps=1e-8, *,
                   projection_method="normalize", mask_semantics="reference", dtype="float64", device=None,
                   distributed=False, shard=None, stats=None, engine="auto", dense_storage=None, check_range=False):
    """NBMF-MM solver, drop-in for ``nbmf_mm._solver.nbmf_mm_solver`` (``_solver.py:61-216``).

    Returns ``(W (m x k), H (k x n), losses, 0.0, n_iter)`` exactly as the reference does
    (``time_elapsed`` is hard-coded 0.0 there, ``_solver.py:216``).  Keyword-only extras:
    ``projection_method`` ("normalize" | "duchi"), ``mask_semantics`` ("reference" quirk |
    "strict"), ``dtype`` ("float64" parity mode | "float32" throughput mode), ``device``,
    ``distributed`` (row-shard over the default ``torch.distributed`` group, one rank per GPU:
    every rank passes the FULL ``Y`` and gets the full result), ``shard=(row0, m_total)`` (with
    ``distributed``: ``Y``/``mask`` are already this rank's row block of an ``m_total``-row
    problem in internal orientation; the returned W is the local block, H is global),
    ``engine`` ("auto" | "simt" | "tensor" | "fused": packed-FFMA2 CUDA-core kernels, the tcgen05/TMEM
    split-precision kernels (float32, binary V, K <= 64) or the persistent small-fit kernel (binary V, K <= 32, one GPU:
    whole iterations inside one launch; what "auto" gives a small problem), ``dense_storage`` ("float16":
    probabilistic V is stored as fp16 on the device, float32 arithmetic; default = the compute dtype),
    ``check_range`` (raise the estimator's ``ValueError("X must be binary")``, ``_base.py:90-91``, when a dense ``Y``
    holds a value outside [0, 1]: the test runs on the device, in the pass that packs ``Y``).
    """
    if orientation not in _CANON:
        raise ValueError(f"Unknown orientation: {orientation}. Must be one of {list(_CANON)}")
    if max_iter < 1:
        # the reference leaves `iteration` unbound for max_iter=0 (_solver.py:215)
        raise UnboundLocalError("max_iter must be >= 1")
    if random_state is not None:
        np.random.seed(random_state)                      # global legacy stream, as _solver.py:102-103
    transpose = orientation == "dir-beta"
    k = int(n_components)
    rank, world = 0, 1
    if distributed:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            rank, world = dist.get_rank(), dist.get_world_size()
    shape = tuple(getattr(Y, "shape", None) or np.shape(Y))
    if len(shape) != 2:
        raise ValueError("Y must be 2-D")
    m, n = (shape[1], shape[0]) if transpose else shape          # internal orientation (_solver.py:113-123)
    if shard is not None:
        if transpose or not distributed:
            raise ValueError("shard=(row0, m_total) needs distributed=True and the internal (beta-dir) orientation")
        shard_row0, m = int(shard[0]), int(shard[1])
        r0, r1 = shard_row0, shard_row0 + shape[0]
    elif world > 1:
        r0, r1 = _row_shard(m, rank, world)
        empty = [r for r in range(world) if _row_shard(m, r, world)[1] <= _row_shard(m, r, world)[0]]
        if empty:                                             # the same table on every rank: all of them raise
            raise ValueError(f"ranks {empty} of {world} would have no rows (m={m}, shards start on multiples of 32 rows); "
                             "use fewer ranks")
    else:
        r0, r1 = 0, m
    # distributed=True: every rank was handed the FULL matrix; only this rank's row block (internal orientation) is
    # uploaded and packed.  (Bit-packed input with dir-beta cannot be sliced by columns on the host: that one case still
    # packs the whole plane and slices after the device-side transpose.)
    local_input = shard is not None
    if world > 1 and shard is None and not (transpose and isinstance(Y, BitMatrix)):
        Y, mask = _rank_block(Y, r0, r1, transpose), (None if mask is None else _rank_block(mask, r0, r1, transpose))
        shape = tuple(Y.shape)
        local_input = True
    # large problems: the inits go up from, and the results come back into, pinned host memory (plain DMA instead of
    # staged pageable copies and page faults of a fresh array).  Page-lock BEFORE the big copies start: cudaHostAlloc
    # stalls DMA submission while it runs (measured: 0.4 s lost when it overlapped the bit-plane upload).
    pinned = None
    full_w = world > 1 and shard is None                   # every rank returns the whole W (all-gathered on the devices)
    if (r1 - r0) * k + k * n >= PIN_THRESHOLD:
        require_cuda(device)
        pinned = pinned_factor_buffers(r1 - r0, k, n, dtype, result_rows=m if full_w else None)
    # Host bit planes in the internal orientation are STREAMED: row chunks go up on a copy stream and each chunk is
    # prepared on the compute stream as it lands (P &= M, mask count, re-tiling for the tensor engine) while the next
    # ones are still crossing PCIe.  Other inputs: asynchronous H2D copies now, device-side preparation in
    # data.finish().  Either way the inits are drawn on the host while the copies are in flight.
    streamed = (isinstance(Y, BitMatrix) and not Y.is_device and not transpose
                and (mask is None or (isinstance(mask, BitMatrix) and not mask.is_device))
                and (world == 1 or local_input))
    data = prob = None
    if streamed:
        if world > 1:
            engine = resolve_engine(engine, dtype=dtype, vkind="bits", k=k, eps=eps, m_total=m, n=n)
        prob = DeviceProblem(shape[0], n, k, dtype=dtype, vkind="bits", has_mask=mask is not None, alpha=alpha, beta=beta,
                             eps=eps, n_obs=None, mask_semantics=mask_semantics, projection=projection_method,
                             max_iter_cap=max_iter, device=device, engine=engine)
        try:
            data_h2d = prob.stream_bits_from_host(Y, mask)
        except BaseException:
            prob.close()
            raise
    else:
        data = prepare_data(Y, mask, transpose=transpose, dtype=dtype, device=device, defer=True, dense_storage=dense_storage,
                            check_range=check_range)
        assert (data.m, data.n) == ((r1 - r0) if local_input else m, n)
    try:
        if transpose and W_init is not None and H_init is not None:      # _solver.py:122-123
            W_init, H_init = np.asarray(H_init).T, np.asarray(W_init).T
        rng, shard_seed = np.random, None
        if world > 1 and W_init is None and H_init is None:
            # Row shards: a rank needs rows [r0, r1) of W_init and its share of H_init.  Drawing the whole (m x k) W_init
            # on every rank to reach them costs 0.25 s at 10^6 x 32 and does not shrink with the number of GPUs, so the
            # reference's stream is entered by MT19937 jump-ahead (draw_shard_inits: same numbers, same global-stream
            # position afterwards).  Without a seed every rank (its own process, its own global stream) would start from
            # a different "global" H and stop at different iterations: rank 0 draws the seed and broadcasts it.
            shard_seed = random_state
            if shard_seed is None:
                import torch.distributed as dist
                box = [int(np.random.randint(0, 2**31 - 1)) if rank == 0 else None]
                dist.broadcast_object_list(box, src=0)
                shard_seed = box[0]
        elif world > 1 and random_state is None and (W_init is None or H_init is None):
            import torch.distributed as dist
            box = [int(np.random.randint(0, 2**31 - 1)) if rank == 0 else None]
            dist.broadcast_object_list(box, src=0)
            rng = np.random.RandomState(box[0])
        H_part = None
        if shard_seed is not None:
            W_local, H_part, _ = draw_shard_inits(shard_seed, m, n, k, r0, r1, h_part=(rank, world),
                                                  set_global_state=random_state is not None)
            n_w_init, n_h_init = W_local.size, H_part.size
        else:
            if W_init is None:
                W_init = rng.uniform(0.1, 0.9, (m, k))        # W first, then H: _solver.py:126-129
            if H_init is None:
                H_init = rng.uniform(0.1, 0.9, (k, n))
            W_init = np.asarray(W_init, dtype=np.float64)
            H_init = np.asarray(H_init, dtype=np.float64)
            if tuple(W_init.shape) != (m, k) or tuple(H_init.shape) != (k, n):
                raise ValueError(f"W_init / H_init have shapes {W_init.shape} / {H_init.shape}, expected {(m, k)} / {(k, n)}")
            W_local = W_init[r0:r1]
            n_w_init, n_h_init = W_local.size, H_init.size
        W_up, H_up, result_buffers = W_local, H_init, None
        if pinned is not None:
            (W_up, H_up), result_buffers = pinned
            W_up.numpy()[...] = W_local                        # fp64 -> compute dtype on the host, copies still in flight
            if H_part is None:
                H_up.numpy()[...] = H_init
        if H_part is not None:
            H_up = _gather_h_parts(H_part, k, n, world, dtype, require_cuda(device))
    except BaseException:
        if prob is not None:                                   # streamed upload in flight: release the context
            prob.close()
        raise

    def all_ranks_sum(x):
        import torch
        import torch.distributed as dist
        t = torch.tensor([x], dtype=torch.float64, device=require_cuda(device))
        dist.all_reduce(t)
        return float(t.item())

    fit_ctx = contextlib.ExitStack()
    if not streamed:
        data.finish()
        data_h2d = data.h2d_bytes
        n_obs_global = data.n_obs
        if local_input:
            if world > 1:
                n_obs_global = all_ranks_sum(data.n_obs)
        elif world > 1:
            data = PreparedData(r1 - r0, n, data.vkind,
                                None if data.P is None else data.P.rows(r0, r1),
                                None if data.M is None else data.M.rows(r0, r1),
                                None if data.Vm is None else data.Vm[r0:r1], n_obs_global, data.h2d_bytes,
                                Wm=None if data.Wm is None else data.Wm[r0:r1])
        if world > 1:
            engine = resolve_engine(engine, dtype=dtype, vkind=data.vkind, k=k, eps=eps, m_total=m, n=n)
        # Mid-size single-GPU fits (up to 2^24 entries) are bound by launch overhead (55-85 us per iteration); the C
        # side replays four iterations per CUDA-graph launch, but the legacy default stream cannot be captured.  Run the
        # fit on a cached side stream then (measured: 15-20 us per iteration less).
        if world == 1 and data.m * data.n <= (1 << 24):
            import torch
            dev_ = require_cuda(device)
            cur = torch.cuda.current_stream(dev_)
            if cur == torch.cuda.default_stream(dev_):
                side = _fit_stream(dev_)
                side.wait_stream(cur)                      # the data planes were produced on the caller's stream
                fit_ctx.enter_context(torch.cuda.stream(side))
                fit_ctx.callback(cur.wait_stream, side)    # ... and the caller's stream continues after the fit
        try:
            prob = make_problem(data, k, dtype=dtype, alpha=alpha, beta=beta, eps=eps, mask_semantics=mask_semantics,
                                projection=projection_method, max_iter_cap=max_iter, device=device, n_obs=n_obs_global,
                                engine=engine)
        except BaseException:
            fit_ctx.close()
            raise
    try:
        if world > 1:
            prob.init_comm()
        prob.set_factors(W_up, H_up, normalize_w=True)
        if streamed:                                       # count_nonzero(mask) of this shard, _solver.py:151,155
            n_obs_local = prob.finish_bits()
            prob.set_n_obs(all_ranks_sum(n_obs_local) if world > 1 else n_obs_local)
        if prob.release_planes() and data is not None and data.vkind == "bits" and not isinstance(Y, BitMatrix):
            data.P = data.M = None                         # tensor engine: the row-major planes packed from X are not read again
        losses_arr, n_iter, converged = prob.fit(max_iter, tol)
        # tail of the reference solver (_solver.py:192-213) on the device: the simplex factor is the internal W in
        # both orientations; it is renormalised (fp64) only when its worst deviation exceeds 1e-9 -- the worst over
        # ALL rows, so row shards agree on the decision first
        dev = prob.simplex_deviation()
        if world > 1:
            import torch
            import torch.distributed as dist
            t = torch.tensor([dev if np.isfinite(dev) else np.inf], dtype=torch.float64, device=require_cuda(device))
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dev = float(t.item())
        norm = bool(np.isfinite(dev) and dev > 1e-9)
        if full_w:
            # every rank returns the whole W: the fp64 row blocks are all-gathered on the devices (equal-sized,
            # zero-padded blocks of _row_shard's stride) and cross PCIe once, into pinned memory
            W, H = prob.get_factors_f64(normalize_w=norm, out=result_buffers,
                                        gather=(m, _row_shard(m, 0, world)[1], world))
            n_w_out = (r1 - r0) * k
        else:
            W, H = prob.get_factors_f64(normalize_w=norm, out=result_buffers)
            n_w_out = W.size
    finally:
        prob.close()
        fit_ctx.close()

    losses = [np.float64(v) for v in losses_arr]
    if verbose > 0:                                        # same lines as _solver.py:165-166,172-173
        for i in range(0, n_iter, 10):
            print(f"Iter {i:4d}: Loss = {losses[i]:.6f}")
        if converged:
            print(f"Converged at iteration {n_iter - 1}")
    if stats is not None:
        isz = np.dtype(dtype).itemsize                     # factors cross PCIe in the compute dtype
        stats.update(h2d_bytes=data_h2d + isz * n_w_init + (8 if H_part is not None else isz) * n_h_init, converged=converged, streamed=streamed,
                     d2h_bytes=8 * (W.size + H.size) + losses_arr.nbytes, w_rows_local=n_w_out // max(k, 1), world=world,
                     engine=prob.engine)

    W_final, H_final = W, H                                # (m x k), (k x n) internal
    if transpose:                                          # _solver.py:182-184
        W_final, H_final = np.ascontiguousarray(H_final.T), np.ascontiguousarray(W_final.T)
    return W_final, H_final, losses, 0.0, n_iter


def nbmf_mm_update_beta_dir(Y, W, H, mask, alpha, beta, eps=1e-8, *, projection_method="normalize",
                            mask_semantics="reference", dtype="float64", device=None, engine="auto"):
    """One MM iteration, drop-in for ``nbmf_mm._solver.nbmf_mm_update_beta_dir``
    (``_solver.py:5-59``): ``W`` is ``k x m`` with columns on the simplex, ``H`` is ``k x n``;
    returns ``(W_new (k x m), H_new (k x n))``.  H is updated first and the W step uses the new H."""
    W = np.asarray(W, dtype=np.float64)
    H = np.asarray(H, dtype=np.float64)
    data = prepare_data(Y, mask, transpose=False, dtype=dtype, device=device)
    k = W.shape[0]
    prob = make_problem(data, k, dtype=dtype, alpha=alpha, beta=beta, eps=eps, mask_semantics=mask_semantics,
                        projection=projection_method, max_iter_cap=1, device=device, engine=engine)
    try:
        prob.set_factors(np.ascontiguousarray(W.T), H, normalize_w=False)
        prob.h_half_step()
        prob.w_half_step()
        W_new, H_new = prob.get_factors()
    finally:
        prob.close()
    return np.ascontiguousarray(W_new.T), H_new
