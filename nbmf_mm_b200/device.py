"""Device-side plumbing between the Python host and the C-ABI (``include/nbmf_b200.h``).

PyTorch is used for what it is good at here -- device memory, streams and
``torch.distributed`` -- while all arithmetic of the hot path runs in the hand-written
sm_100a kernels of ``libnbmf_b200.so``.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .bits import BitMatrix, words_per_row


def _torch():
    import torch
    return torch


def require_cuda(device=None):
    torch = _torch()
    if not torch.cuda.is_available():
        raise RuntimeError("nbmf_mm_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)


def _ptr(t):
    return C.c_void_p(0 if t is None else t.data_ptr())


def _stream(dev):
    return C.c_void_p(_torch().cuda.current_stream(dev).cuda_stream)


_TORCH_DT = {"float32": "float32", "float64": "float64"}


def dtype_code(dtype) -> int:
    name = np.dtype(dtype).name
    if name == "float32":
        return _lib.NBMF_F32
    if name == "float64":
        return _lib.NBMF_F64
    raise ValueError(f"dtype must be float32 or float64, got {dtype!r}")


def _elem_code(t) -> int:
    torch = _torch()
    if t.dtype == torch.float32:
        return _lib.NBMF_F32
    if t.dtype == torch.float64:
        return _lib.NBMF_F64
    if t.dtype in (torch.uint8, torch.bool):
        return _lib.NBMF_U8
    raise ValueError(f"unsupported element type {t.dtype}")


# ----------------------------------------------------------------------------- data layer
def pack_bits_device(X_dev, mask_dev=None, want_mask_plane=True):
    """Dense device matrix (+ optional dense mask) -> (P = X!=0 & mask, M = mask) bit planes."""
    torch = _torch()
    lib = _lib.load()
    m, n = X_dev.shape
    wpr = words_per_row(n)
    dev = X_dev.device
    X_dev = X_dev.contiguous()
    P = torch.empty((m, wpr), dtype=torch.int32, device=dev)
    M = None
    if mask_dev is not None:
        mask_dev = mask_dev.contiguous()
        if want_mask_plane:
            M = torch.empty((m, wpr), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.nbmf_pack_bits(_ptr(X_dev), _elem_code(X_dev), n, _ptr(mask_dev),
                                      _elem_code(mask_dev) if mask_dev is not None else 0, n,
                                      m, n, _ptr(P), _ptr(M), _stream(dev)), "nbmf_pack_bits")
    return BitMatrix(P, (m, n)), (BitMatrix(M, (m, n)) if M is not None else None)


def _packable(a):
    """Host array in a dtype the packing kernels read (f32, f64, u8); bool is viewed as u8, the rest becomes f64."""
    a = np.asarray(a)
    if a.dtype == np.bool_:
        return a.view(np.uint8)
    if a.dtype in (np.float32, np.float64, np.uint8):
        return a
    return a.astype(np.float64)


def pack_host_dense_checked(X, mask, device, chunk_bytes=16 << 20, n_threads=None, pinned_from=64 << 20):
    """Dense HOST X (m x n) [+ dense host mask] -> device bit planes (P = X != 0 & mask, M = mask != 0) and the value
    flags of ``nbmf_pack_bits_checked``, without a NumPy pass over the data: row chunks are uploaded as they are
    (fp64 / fp32 / u8) and checked + packed on the device.  At 20 000 x 5 000 the host front end it replaces (range
    test, binary test, mask test, ``np.packbits``) costs 3.4 s; this costs the PCIe time of the raw arrays.

    Large inputs go through page-locked staging: a few threads copy chunk i + 1 into a pinned buffer (NumPy copies release
    the GIL) while chunk i crosses PCIe by DMA and is packed; a plain ``.to(device)`` of pageable memory is a synchronous
    copy through the driver's own staging at a third of the rate (1.6 GB of fp64 X + mask: 0.24 s -> see DESIGN 6b)."""
    torch = _torch()
    lib = _lib.load()
    dev = require_cuda(device)
    X = _packable(X)
    mask = None if mask is None else _packable(mask)
    m, n = X.shape
    wpr = words_per_row(n)
    P = torch.empty((m, wpr), dtype=torch.int32, device=dev)
    M = None if mask is None else torch.empty((m, wpr), dtype=torch.int32, device=dev)
    flags = torch.zeros(1, dtype=torch.int32, device=dev)
    h2d = X.nbytes + (0 if mask is None else mask.nbytes)

    def pack(Xc, Mc, r0, r1):
        _lib.check(lib.nbmf_pack_bits_checked(_ptr(Xc), _elem_code(Xc), n, _ptr(Mc), _elem_code(Mc) if Mc is not None else 0,
                                              n, r1 - r0, n, _ptr(P[r0:r1]), _ptr(None if M is None else M[r0:r1]),
                                              _ptr(flags), _stream(dev)), "nbmf_pack_bits_checked")

    rows = max(1, int(chunk_bytes) // max(1, n * max(X.dtype.itemsize, 1 if mask is None else mask.dtype.itemsize)))
    with torch.cuda.device(dev):
        if h2d < pinned_from or m <= rows:                     # small: not worth page-locking anything
            for r0 in range(0, m, rows):
                r1 = min(m, r0 + rows)
                Xc = torch.from_numpy(np.ascontiguousarray(X[r0:r1])).to(dev)
                Mc = None if mask is None else torch.from_numpy(np.ascontiguousarray(mask[r0:r1])).to(dev)
                pack(Xc, Mc, r0, r1)
            return BitMatrix(P, (m, n)), (None if M is None else BitMatrix(M, (m, n))), int(flags.item()), h2d
        from concurrent.futures import ThreadPoolExecutor
        import os
        nt = n_threads or max(1, min(8, (os.cpu_count() or 2) // 2))
        srcs = [X] + ([] if mask is None else [mask])
        tdt = [getattr(torch, a.dtype.name) for a in srcs]
        stage = [[torch.empty((rows, n), dtype=t, pin_memory=True) for t in tdt] for _ in range(2)]
        stage_np = [[t.numpy() for t in pair] for pair in stage]
        devbuf = [[torch.empty((rows, n), dtype=t, device=dev) for t in tdt] for _ in range(2)]
        done = [None, None]                                     # DMA out of staging pair b has finished

        def fill(b, r0, r1):
            step = -(-(r1 - r0) // nt)
            jobs = []
            for si, a in enumerate(srcs):
                for q0 in range(r0, r1, step):
                    q1 = min(r1, q0 + step)
                    jobs.append(pool.submit(np.copyto, stage_np[b][si][q0 - r0:q1 - r0], a[q0:q1]))
            for j in jobs:
                j.result()

        with ThreadPoolExecutor(max_workers=nt) as pool:
            for i, r0 in enumerate(range(0, m, rows)):
                r1, b = min(m, r0 + rows), i % 2
                if done[b] is not None:
                    done[b].synchronize()
                fill(b, r0, r1)
                for si in range(len(srcs)):
                    devbuf[b][si][: r1 - r0].copy_(stage[b][si][: r1 - r0], non_blocking=True)
                done[b] = torch.cuda.Event()
                done[b].record()
                pack(devbuf[b][0][: r1 - r0], devbuf[b][1][: r1 - r0] if mask is not None else None, r0, r1)
    return BitMatrix(P, (m, n)), (None if M is None else BitMatrix(M, (m, n))), int(flags.item()), h2d


def pack_dense_device(X_dev, mask_dev, dtype):
    """Dense device matrix (+ mask) -> V*mask in ``dtype`` with the padded leading dimension."""
    torch = _torch()
    lib = _lib.load()
    m, n = X_dev.shape
    ldv = words_per_row(n) * 32
    dev = X_dev.device
    X_dev = X_dev.contiguous()
    if mask_dev is not None:
        mask_dev = mask_dev.contiguous()
    half = np.dtype(dtype) == np.float16                      # fp16 storage layout (NBMF_V_DENSE_F16)
    Vm = torch.empty((m, ldv), dtype=getattr(torch, np.dtype(dtype).name), device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.nbmf_pack_dense(_ptr(X_dev), _elem_code(X_dev), n, _ptr(mask_dev),
                                       _elem_code(mask_dev) if mask_dev is not None else 0, n,
                                       m, n, _lib.NBMF_F16 if half else dtype_code(dtype), _ptr(Vm), _stream(dev)),
                   "nbmf_pack_dense")
    return Vm


def pack_csr_device(A, device):
    """scipy.sparse matrix -> (BitMatrix on ``device``, flags) without a dense host or device copy: only
    indptr / indices / data cross PCIe.  flags bit 0: a stored value is not 0/1, bit 1: outside [0, 1]."""
    torch = _torch()
    lib = _lib.load()
    dev = require_cuda(device)
    A = A.tocsr()
    m, n = A.shape
    indptr = torch.from_numpy(np.ascontiguousarray(A.indptr, dtype=np.int64)).to(dev)
    indices = torch.from_numpy(np.ascontiguousarray(A.indices, dtype=np.int32)).to(dev)
    data = None
    if A.data.dtype != np.bool_:
        dt = np.float32 if A.data.dtype == np.float32 else np.float64
        data = torch.from_numpy(np.ascontiguousarray(A.data, dtype=dt)).to(dev)
    P = torch.empty((m, words_per_row(n)), dtype=torch.int32, device=dev)
    flags_dev = torch.zeros(1, dtype=torch.int32, device=dev)
    flags = C.c_int32(0)
    with torch.cuda.device(dev):
        _lib.check(lib.nbmf_pack_csr(_ptr(indptr), _ptr(indices), _ptr(data), (0 if data.dtype == torch.float32 else 1) if data is not None else 0,
                                     m, n, _ptr(P), _ptr(flags_dev), C.byref(flags), _stream(dev)), "nbmf_pack_csr")
    h2d = indptr.numel() * 8 + indices.numel() * 4 + (0 if data is None else data.numel() * data.element_size())
    return BitMatrix(P, (m, n)), int(flags.value), h2d


def reconstruct_device(W, H, dtype, device):
    """clip(W @ H, 0, 1) on the device (``_base.py:201-210``); returns a host float64 array."""
    torch = _torch()
    lib = _lib.load()
    dev = require_cuda(device)
    tdt = getattr(torch, np.dtype(dtype).name)
    Wd = torch.from_numpy(np.ascontiguousarray(W, dtype=np.dtype(dtype))).to(dev)
    Hd = torch.from_numpy(np.ascontiguousarray(H, dtype=np.dtype(dtype))).to(dev)
    m, k = Wd.shape
    n = Hd.shape[1]
    out = torch.empty((m, n), dtype=tdt, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.nbmf_reconstruct(dtype_code(dtype), _ptr(Wd), _ptr(Hd), m, n, k, _ptr(out), _stream(dev)),
                   "nbmf_reconstruct")
    return out.cpu().numpy().astype(np.float64)


def transpose_device(B: BitMatrix) -> BitMatrix:
    torch = _torch()
    lib = _lib.load()
    m, n = B.shape
    dev = B.words.device
    out = torch.empty((n, words_per_row(m)), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.nbmf_transpose_bits(_ptr(B.words), m, n, _ptr(out), _stream(dev)), "nbmf_transpose_bits")
    return BitMatrix(out, (n, m))


def popcount_device(B: BitMatrix) -> int:
    torch = _torch()
    lib = _lib.load()
    dev = B.words.device
    scratch = torch.zeros(1, dtype=torch.int64, device=dev)
    out = C.c_uint64(0)
    with torch.cuda.device(dev):
        _lib.check(lib.nbmf_popcount_bits(_ptr(B.words), B.shape[0], B.shape[1], _ptr(scratch), C.byref(out),
                                          _stream(dev)), "nbmf_popcount_bits")
    return int(out.value)


def synth_bits_device(seed, row0, m, n, hstar, obs_frac, device, with_mask=True):
    """Counter-based generator of config 4: returns (P, M) planes on ``device``."""
    torch = _torch()
    lib = _lib.load()
    dev = require_cuda(device)
    hs = torch.as_tensor(np.ascontiguousarray(hstar, dtype=np.float32), device=dev)
    wpr = words_per_row(n)
    P = torch.empty((m, wpr), dtype=torch.int32, device=dev)
    M = torch.empty((m, wpr), dtype=torch.int32, device=dev) if with_mask else None
    with torch.cuda.device(dev):
        _lib.check(lib.nbmf_synth_bits(int(seed), int(row0), int(m), int(n), _ptr(hs), int(hs.shape[0]),
                                       float(obs_frac), _ptr(P), _ptr(M), _stream(dev)), "nbmf_synth_bits")
    return BitMatrix(P, (m, n)), (BitMatrix(M, (m, n)) if with_mask else None)


# ----------------------------------------------------------------------------- fit context
class DeviceProblem:
    """One NBMF-MM problem (or one row shard of it) resident on one GPU.

    Thin owner of an ``nbmf_ctx``: the workspace and the data planes are torch tensors,
    every method is one C-ABI call.  Internal orientation: V is ``m x n``, W ``m x k``,
    H ``k x n`` (``_solver.py:113-136``).
    """

    def __init__(self, m, n, k, *, dtype="float64", vkind="bits", has_mask=False, alpha=1.2, beta=1.2,
                 eps=1e-8, n_obs=None, mask_semantics="reference", projection="normalize",
                 max_iter_cap=2000, device=None, engine="auto", workspace=None, batch_hint=0):
        torch = _torch()
        self.lib = _lib.load()
        self.dev = require_cuda(device)
        self.m, self.n, self.k = int(m), int(n), int(k)
        self.np_dtype = np.dtype(dtype)
        self.t_dtype = getattr(torch, self.np_dtype.name)
        if mask_semantics not in ("reference", "strict"):
            raise ValueError(f"mask_semantics must be 'reference' or 'strict', got {mask_semantics!r}")
        if projection not in ("normalize", "duchi"):
            raise ValueError(f"projection_method must be 'normalize' or 'duchi', got {projection!r}")
        cfg = _lib.NbmfConfig()
        cfg.m, cfg.n, cfg.k = self.m, self.n, self.k
        cfg.dtype = dtype_code(dtype)
        if vkind not in ("bits", "dense", "dense16"):
            raise ValueError(f"vkind must be 'bits', 'dense' or 'dense16', got {vkind!r}")
        cfg.vkind = {"bits": _lib.NBMF_V_BITS, "dense": _lib.NBMF_V_DENSE, "dense16": _lib.NBMF_V_DENSE_F16}[vkind]
        cfg.mask_semantics = _lib.NBMF_MASK_STRICT if mask_semantics == "strict" else _lib.NBMF_MASK_REFERENCE
        cfg.projection = _lib.NBMF_PROJ_DUCHI if projection == "duchi" else _lib.NBMF_PROJ_NORMALIZE
        cfg.has_mask = 1 if has_mask else 0
        cfg.alpha, cfg.beta, cfg.eps = float(alpha), float(beta), float(eps)
        cfg.n_obs = float(self.m * self.n if n_obs is None else n_obs)
        cfg.max_iter_cap = int(max_iter_cap)
        if engine not in _lib.ENGINES:
            raise ValueError(f"engine must be one of {sorted(_lib.ENGINES)}, got {engine!r}")
        cfg.engine = _lib.ENGINES[engine]
        cfg.batch_hint = int(batch_hint)
        self.cfg = cfg
        nbytes = self.lib.nbmf_workspace_bytes(C.byref(cfg))
        if nbytes < 0:
            _lib.check(-1, "nbmf_workspace_bytes")
        # workspace: one allocation per problem, or (batched small fits) a slice handed out by `workspace(nbytes)`
        self.workspace = (torch.empty(int(nbytes), dtype=torch.uint8, device=self.dev) if workspace is None
                          else workspace(int(nbytes)))
        self._ctx = C.c_void_p(0)
        with torch.cuda.device(self.dev):
            _lib.check(self.lib.nbmf_create(C.byref(cfg), _ptr(self.workspace), nbytes, _stream(self.dev),
                                            C.byref(self._ctx)), "nbmf_create")
        self._keep = []          # tensors the context borrows
        self.world = 1
        self.engine = {_lib.NBMF_ENGINE_TENSOR: "tensor", _lib.NBMF_ENGINE_FUSED: "fused"}.get(self.lib.nbmf_engine(self._ctx), "simt")

    @property
    def fit_is_fused(self) -> bool:
        """The fit loop of this problem runs inside the persistent small-fit kernel (``nbmf_fit_is_fused``)."""
        return bool(self.lib.nbmf_fit_is_fused(self._ctx))

    # -- lifetime
    def close(self):
        if self._ctx:
            self.lib.nbmf_destroy(self._ctx)         # synchronises the context's stream
            self._ctx = C.c_void_p(0)
            self._keep = []
            self.workspace = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _call(self, name, *args):
        with _torch().cuda.device(self.dev):
            _lib.check(getattr(self.lib, name)(self._ctx, *args), name)

    # -- data
    def set_bits(self, P: BitMatrix, M: BitMatrix | None = None):
        if P.shape != (self.m, self.n):
            raise ValueError(f"P has shape {P.shape}, expected {(self.m, self.n)}")
        P = P if P.is_device else P.to_device(self.dev)
        if M is not None:
            M = M if M.is_device else M.to_device(self.dev)
        self._keep = [P.words, None if M is None else M.words]
        self._call("nbmf_set_data_bits", _ptr(P.words), _ptr(None if M is None else M.words))

    def stream_bits_from_host(self, P: BitMatrix, M: BitMatrix | None = None, n_chunks=8):
        """Upload HOST bit planes chunk by chunk on a copy stream; the context's stream processes each chunk as it
        lands (P &= M, mask count, re-tiling for the tensor engine: ``nbmf_ingest_bits_rows``) while the next ones are
        still crossing PCIe.  Everything is enqueued at once; call ``finish_bits()`` for the mask count."""
        torch = _torch()
        if P.shape != (self.m, self.n) or (M is not None and M.shape != P.shape):
            raise ValueError(f"planes have shapes {P.shape} / {None if M is None else M.shape}, expected {(self.m, self.n)}")
        if P.is_device or (M is not None and M.is_device):
            raise ValueError("stream_bits_from_host takes host planes; use set_bits for device planes")
        host = lambda w: torch.from_numpy(np.ascontiguousarray(w).view(np.int32)) if isinstance(w, np.ndarray) else w
        Ph, Mh = host(P.words), (None if M is None else host(M.words))
        wpr = Ph.shape[1]
        with torch.cuda.device(self.dev):
            compute = torch.cuda.current_stream(self.dev)
            Pd = torch.empty((self.m, wpr), dtype=torch.int32, device=self.dev)
            Md = None if Mh is None else torch.empty((self.m, wpr), dtype=torch.int32, device=self.dev)
            self._keep = [Pd, Md]
            self._call("nbmf_ingest_bits_begin", _ptr(Pd), _ptr(Md))
            copy = torch.cuda.Stream(self.dev)
            copy.wait_stream(compute)
            step = max(128, (-(-self.m // max(1, int(n_chunks))) + 127) // 128 * 128)
            for r0 in range(0, self.m, step):
                r1 = min(self.m, r0 + step)
                with torch.cuda.stream(copy):
                    Pd[r0:r1].copy_(Ph[r0:r1], non_blocking=True)
                    if Md is not None:
                        Md[r0:r1].copy_(Mh[r0:r1], non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(copy)
                compute.wait_event(ev)
                self._call("nbmf_ingest_bits_rows", r0, r1)
            Pd.record_stream(copy)
            if Md is not None:
                Md.record_stream(copy)
        return Ph.numel() * 4 + (0 if Mh is None else Mh.numel() * 4)

    def finish_bits(self) -> float:
        """Wait for the streamed ingestion; returns count_nonzero(mask) (m * n without a mask)."""
        cnt = C.c_double(0.0)
        self._call("nbmf_ingest_bits_end", C.byref(cnt))
        return float(cnt.value)

    def release_planes(self) -> bool:
        """Drop this object's references to the caller's bit planes when the context no longer reads them (tensor engine:
        it works on re-tiled copies).  Returns True if they were dropped; the memory goes back to torch's pool as soon as
        the caller holds no other reference."""
        if self._ctx and not self.lib.nbmf_planes_in_use(self._ctx):
            torch = _torch()
            for t in self._keep:                        # frees are stream-ordered after the re-tiling that reads them
                if t is not None:
                    t.record_stream(torch.cuda.current_stream(self.dev))
            self._keep = []
            return True
        return False

    def set_n_obs(self, n_obs):
        self._call("nbmf_set_n_obs", float(n_obs))

    def set_dense(self, Vm, M: BitMatrix | None = None, Wm=None):
        """Dense layout: ``Vm`` = V * mask, ``M`` = bit plane of (mask != 0), ``Wm`` = the VALUES of a weighted mask in the
        same dense layout (``None`` for a 0/1 mask)."""
        if M is not None:
            M = M if M.is_device else M.to_device(self.dev)
        self._keep = [Vm, None if M is None else M.words, Wm]
        self._call("nbmf_set_data_dense", _ptr(Vm), _ptr(None if M is None else M.words))
        if Wm is not None:
            self._call("nbmf_set_mask_weights", _ptr(Wm))

    # -- factors
    def _to_dev(self, A, shape):
        torch = _torch()
        if A is None:
            return None
        if not isinstance(A, torch.Tensor):
            A = torch.from_numpy(np.ascontiguousarray(A, dtype=self.np_dtype))
        A = A.to(device=self.dev, dtype=self.t_dtype, non_blocking=(not A.is_cuda) and A.is_pinned()).contiguous()
        if tuple(A.shape) != shape:
            raise ValueError(f"factor has shape {tuple(A.shape)}, expected {shape}")
        return A

    def set_factors(self, W=None, H=None, normalize_w=True):
        Wd = self._to_dev(W, (self.m, self.k))
        Hd = self._to_dev(H, (self.k, self.n))
        self._call("nbmf_set_factors", _ptr(Wd), _ptr(Hd), 1 if normalize_w else 0)

    def get_factors_device(self):
        torch = _torch()
        W = torch.empty((self.m, self.k), dtype=self.t_dtype, device=self.dev)
        H = torch.empty((self.k, self.n), dtype=self.t_dtype, device=self.dev)
        self._call("nbmf_get_factors", _ptr(W), _ptr(H))
        return W, H

    def get_factors(self):
        W, H = self.get_factors_device()
        return W.cpu().numpy().astype(np.float64), H.cpu().numpy().astype(np.float64)

    def simplex_deviation(self) -> float:
        """max_i |sum_k W[i,k] - 1| in fp64 (NaN if a row sum is not finite): the test of ``_solver.py:195-199``."""
        dev = C.c_double(0.0)
        self._call("nbmf_simplex_deviation", C.byref(dev))
        return float(dev.value)

    def get_factors_f64_device(self, W_out, H_out, normalize_w=False):
        """fp64 export (+ the optional row renormalisation of the solver tail) into caller-provided DEVICE tensors: no
        copy to the host, no synchronisation (batched fits collect all results and cross PCIe once)."""
        self._call("nbmf_get_factors_f64", _ptr(W_out), _ptr(H_out), 1 if normalize_w else 0)

    def get_factors_f64(self, normalize_w=False, out=None, gather=None):
        """Host fp64 copies of W (m x k) and H (k x n); conversion to fp64 and the optional row renormalisation
        of the solver tail (``_solver.py:200-204``) run on the device.  ``out`` = (W, H) pinned host fp64 tensors
        (``pinned_result_buffers``) makes the D2H copy a plain DMA and the returned arrays views of them: at
        10^6 x 32 the pageable path costs 112 ms, most of it page faults of the fresh array, the DMA 5 ms."""
        torch = _torch()
        W = torch.empty((self.m, self.k), dtype=torch.float64, device=self.dev)
        H = torch.empty((self.k, self.n), dtype=torch.float64, device=self.dev)
        self._call("nbmf_get_factors_f64", _ptr(W), _ptr(H), 1 if normalize_w else 0)
        if gather is not None:
            # row shards: all-gather the W blocks on the devices.  ``gather = (m_total, rows_per_rank, world)``: every
            # rank contributes rows_per_rank rows (its block, zero padded), rank r's block starts at r * rows_per_rank
            import torch.distributed as dist
            m_total, per, world = gather
            mine = W
            if self.m != per:
                mine = torch.zeros((per, self.k), dtype=torch.float64, device=self.dev)
                mine[: self.m] = W
            full = torch.empty((per * world, self.k), dtype=torch.float64, device=self.dev)
            dist.all_gather_into_tensor(full, mine)
            W = full[:m_total]
        if out is None:
            return W.cpu().numpy(), H.cpu().numpy()
        Wh, Hh = out
        with torch.cuda.device(self.dev):
            Wh.copy_(W, non_blocking=True)
            Hh.copy_(H, non_blocking=True)
            torch.cuda.current_stream(self.dev).synchronize()
        return Wh.numpy(), Hh.numpy()

    # -- steps
    def h_half_step(self):
        self._call("nbmf_h_half_step")

    def w_half_step(self):
        self._call("nbmf_w_half_step")

    def objective(self) -> float:
        out = C.c_double(0.0)
        self._call("nbmf_objective", C.byref(out))
        return float(out.value)

    def loglik_partials(self):
        """(row splits x column blocks) array of the per-CTA log-likelihood sums of the most recent H pass."""
        info = self.plan_info()
        n = info["h_col_blocks"] * info["h_row_splits"]
        out = np.zeros(n, dtype=np.float64)
        cb, rs = C.c_int32(0), C.c_int32(0)
        self._call("nbmf_loglik_partials", out.ctypes.data_as(C.POINTER(C.c_double)), n, C.byref(cb), C.byref(rs))
        return out.reshape(rs.value, cb.value)

    # -- loop
    def fit(self, max_iter, tol):
        """Run the device-resident loop; returns (losses ndarray, n_iter, converged)."""
        hist = np.zeros(int(max_iter) + 2, dtype=np.float64)
        n_iter = C.c_int32(0)
        conv = C.c_int32(0)
        self._call("nbmf_fit", int(max_iter), float(tol), hist.ctypes.data_as(C.POINTER(C.c_double)),
                   C.byref(n_iter), C.byref(conv))
        return hist[: n_iter.value].copy(), int(n_iter.value), bool(conv.value)

    def fit_begin(self, max_iter, tol):
        self._call("nbmf_fit_begin", int(max_iter), float(tol))

    def fit_enqueue(self, n_iters):
        self._call("nbmf_fit_enqueue", int(n_iters))

    def fit_poll(self, wait=False):
        done, n_iter = C.c_int32(0), C.c_int32(0)
        self._call("nbmf_fit_poll", 1 if wait else 0, C.byref(done), C.byref(n_iter))
        return int(done.value), int(n_iter.value)

    def fit_history(self, count):
        hist = np.zeros(max(int(count), 1), dtype=np.float64)
        conv = C.c_int32(0)
        self._call("nbmf_fit_history", hist.ctypes.data_as(C.POINTER(C.c_double)), int(count), C.byref(conv))
        return hist[: int(count)].copy(), bool(conv.value)

    def batch_bind(self, n, stride_bytes):
        """Make this context the leader of ``n`` contexts of identical configuration whose workspaces lie
        ``stride_bytes`` apart (``nbmf_batch_bind``): its ``fit_enqueue`` then advances all of them per launch."""
        self._call("nbmf_batch_bind", int(n), int(stride_bytes))
        self._batch_n = int(n)

    def batch_poll(self):
        """(every fit of the batch has stopped, [losses recorded by fit i]); synchronises the stream."""
        n = getattr(self, "_batch_n", 1)
        done = C.c_int32(0)
        iters = (C.c_int32 * n)()
        self._call("nbmf_batch_poll", C.byref(done), iters)
        return bool(done.value), [int(v) for v in iters]

    def batch_tail(self, hist_len):
        """(loss histories [n, hist_len], converged [n], simplex deviations [n]) of the batch this context leads, with one
        synchronisation (``nbmf_batch_tail``)."""
        n = getattr(self, "_batch_n", 1)
        hist = np.zeros((n, max(int(hist_len), 1)), dtype=np.float64)
        conv = (C.c_int32 * n)()
        dev = np.zeros(n, dtype=np.float64)
        self._call("nbmf_batch_tail", hist.ctypes.data_as(C.POINTER(C.c_double)), hist.shape[1], conv,
                   dev.ctypes.data_as(C.POINTER(C.c_double)))
        return hist, [bool(v) for v in conv], dev

    def transform(self, n_steps=50):
        self._call("nbmf_transform", int(n_steps))

    # -- measurement
    def profile(self, enable=True):
        self._call("nbmf_profile_enable", 1 if enable else 0)

    def profile_read(self):
        """(H-pass total ms, launches, W-pass total ms, launches) since profile(True)."""
        hm, wm, hc, wc = C.c_double(0), C.c_double(0), C.c_int32(0), C.c_int32(0)
        self._call("nbmf_profile_read", C.byref(hm), C.byref(hc), C.byref(wm), C.byref(wc))
        return hm.value, hc.value, wm.value, wc.value

    def plan_info(self):
        a, b, c_, d = C.c_int32(0), C.c_int32(0), C.c_int32(0), C.c_int32(0)
        self._call("nbmf_plan_info", C.byref(a), C.byref(b), C.byref(c_), C.byref(d))
        return dict(h_col_blocks=a.value, h_row_splits=b.value, w_row_blocks=c_.value, w_col_splits=d.value)

    # -- multi-GPU
    def init_comm(self, group=None):
        """Join the row-shard communicator of ``group``: rank 0 creates an NCCL unique id, torch.distributed
        (any backend) carries it to the other ranks, every rank calls ncclCommInitRank.  The communicator is
        created once per (group, device) and re-attached to later problems (ncclCommInitRank costs ~1 s)."""
        import torch.distributed as dist
        world = dist.get_world_size(group)
        rank = dist.get_rank(group)
        if world == 1:
            return
        key = (id(group) if group is not None else None, world, rank, self.dev.index)
        comm = _COMM_CACHE.get(key)
        if comm is None:
            _lib.nccl_library_hint()
            buf = (C.c_ubyte * 128)()
            if rank == 0:
                _lib.check(self.lib.nbmf_comm_unique_id(buf), "nbmf_comm_unique_id")
            payload = [bytes(buf)]
            dist.broadcast_object_list(payload, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
            raw = (C.c_ubyte * 128).from_buffer_copy(payload[0])
            handle = C.c_void_p()
            _lib.check(self.lib.nbmf_comm_create(raw, rank, world, C.byref(handle)), "nbmf_comm_create")
            comm = _COMM_CACHE[key] = handle
        # every rank must reduce the same [C | D] layout: same engine, same padded K (fail loudly instead of hanging)
        torch = _torch()
        mine = [1 if self.engine == "tensor" else 0, int(self.k), self.np_dtype.itemsize, int(self.n)]
        t = torch.tensor(mine + [-v for v in mine], dtype=torch.int64, device=self.dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)         # max(x) and -min(x) in one small collective
        t = t.tolist()
        if any(t[i] != -t[i + 4] for i in range(4)):
            raise RuntimeError(f"row shards disagree on (tensor engine, k, itemsize, n): max {t[:4]}, min {[-v for v in t[4:]]}; "
                               "pass an explicit engine= or let nbmf_mm_solver decide it from the global problem size")
        self._call("nbmf_comm_attach", comm, rank, world)
        self.world = world


_COMM_CACHE = {}


PIN_THRESHOLD = 1 << 22        # factor elements from which pinned host staging pays (torch caches pinned blocks)


def pinned_factor_buffers(m, k, n, dtype, result_rows=None):
    """Pinned host tensors for a large problem's factors: (W0, H0) in the compute dtype for the upload of the
    inits and (W, H) in fp64 for the results.  Page-locking ~0.4 GB costs ~0.15 s the first time; the solver calls
    this while the H2D copies of the bit planes are in flight, when the host has nothing else to do."""
    torch = _torch()
    tdt = getattr(torch, np.dtype(dtype).name)
    mk = lambda shape, dt: torch.empty(shape, dtype=dt, pin_memory=True)
    return (mk((m, k), tdt), mk((k, n), tdt)), (mk((m if result_rows is None else result_rows, k), torch.float64),
                                                mk((k, n), torch.float64))


def destroy_cached_comms():
    """Destroy the cached NCCL communicators (call before torch.distributed.destroy_process_group)."""
    lib = _lib.load()
    for comm in _COMM_CACHE.values():
        lib.nbmf_comm_destroy(comm)
    _COMM_CACHE.clear()
