"""The binding a maintainer of siddC/nbmf_mm would add to route the loop of ``nbmf_mm_solver`` through libnbmf_b200.so
(INTEGRATION.md section 2 shows this file verbatim; ``tests/test_gpu_integration_binding.py`` executes it).  It uses
nothing from ``nbmf_mm_b200``: ctypes against ``include/nbmf_b200.h`` and torch for device memory only."""
# src/nbmf_mm/_b200.py  (new file in the reference)
import ctypes as C, os, numpy as np, torch

lib = C.CDLL(os.environ.get("NBMF_B200_LIB", "libnbmf_b200.so"))

class Cfg(C.Structure):                       # struct nbmf_config, include/nbmf_b200.h
    _fields_ = [("m", C.c_int64), ("n", C.c_int64), ("k", C.c_int32), ("dtype", C.c_int32),
                ("vkind", C.c_int32), ("mask_semantics", C.c_int32), ("projection", C.c_int32),
                ("has_mask", C.c_int32), ("alpha", C.c_double), ("beta", C.c_double), ("eps", C.c_double),
                ("n_obs", C.c_double), ("max_iter_cap", C.c_int32), ("engine", C.c_int32)]

lib.nbmf_workspace_bytes.restype = C.c_int64
lib.nbmf_last_error.restype = C.c_char_p

def _ok(rc):
    if rc: raise RuntimeError(lib.nbmf_last_error().decode())

def fit_loop_b200(Y, mask, W, H, alpha, beta, max_iter, tol, eps=1e-8):
    """Y (m x n) binary, W (k x m) column-normalised, H (k x n): the state nbmf_mm_solver holds at
    _solver.py:139.  Returns (W, H, losses, n_iter) as the loop at _solver.py:143-175 would."""
    k, m = W.shape; n = H.shape[1]
    dev, st = torch.device("cuda"), C.c_void_p(torch.cuda.current_stream().cuda_stream)
    wpr = lib.nbmf_words_per_row(C.c_int64(n))
    Yd = torch.from_numpy(Y).to(dev)
    Md = None if mask is None else torch.from_numpy(np.ascontiguousarray(mask, dtype=np.float64)).to(dev)
    P = torch.empty((m, wpr), dtype=torch.int32, device=dev)
    M = None if mask is None else torch.empty((m, wpr), dtype=torch.int32, device=dev)
    vp = lambda t: C.c_void_p(0 if t is None else t.data_ptr())
    _ok(lib.nbmf_pack_bits(vp(Yd), 1, C.c_int64(n), vp(Md), 1, C.c_int64(n), C.c_int64(m), C.c_int64(n),
                           vp(P), vp(M), st))                        # replaces _solver.py:21-32
    cfg = Cfg(m, n, k, 1, 0, 0, 0, int(mask is not None), alpha, beta, eps,
              float(Y.size if mask is None else np.count_nonzero(mask)), max_iter, 0)
    ws = torch.empty(lib.nbmf_workspace_bytes(C.byref(cfg)), dtype=torch.uint8, device=dev)
    ctx = C.c_void_p()
    _ok(lib.nbmf_create(C.byref(cfg), vp(ws), C.c_int64(ws.numel()), st, C.byref(ctx)))
    try:
        _ok(lib.nbmf_set_data_bits(ctx, vp(P), vp(M)))
        Wd = torch.from_numpy(np.ascontiguousarray(W.T)).to(dev); Hd = torch.from_numpy(H).to(dev)
        _ok(lib.nbmf_set_factors(ctx, vp(Wd), vp(Hd), 0))
        hist = (C.c_double * (max_iter + 2))(); n_iter = C.c_int32(); conv = C.c_int32()
        _ok(lib.nbmf_fit(ctx, max_iter, C.c_double(tol), hist, C.byref(n_iter), C.byref(conv)))
        _ok(lib.nbmf_get_factors(ctx, vp(Wd), vp(Hd)))
        torch.cuda.synchronize()
    finally:
        lib.nbmf_destroy(ctx)
    return Wd.cpu().numpy().T, Hd.cpu().numpy(), list(hist[: n_iter.value]), n_iter.value
