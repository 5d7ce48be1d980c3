#!/usr/bin/env python
"""How should the fp64 factors of config 4 (W: 10^6 x 32) come back to the host?  (development tool; GPU box)"""
import time
import numpy as np
import torch

m, k = 1_000_000, 32
Wd = torch.rand((m, k), dtype=torch.float64, device="cuda")
torch.cuda.synchronize()

def t(f, name, reps=3):
    for r in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        out = f()
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print(f"{name} rep {r}: {dt * 1e3:.1f} ms", flush=True)
        del out

t(lambda: Wd.cpu().numpy(), "pageable .cpu().numpy()")
def pinned_view():
    h = torch.empty((m, k), dtype=torch.float64, pin_memory=True)
    h.copy_(Wd, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    return h.numpy()
t(pinned_view, "fresh pinned tensor, numpy view")
def f32_then_convert():
    h = torch.empty((m, k), dtype=torch.float32, pin_memory=True)
    h.copy_(Wd.float(), non_blocking=True)
    torch.cuda.current_stream().synchronize()
    return h.numpy().astype(np.float64)
t(f32_then_convert, "fp32 pinned + host astype")
def registered():
    out = np.empty((m, k), dtype=np.float64)
    th = torch.from_numpy(out)
    rc = torch.cuda.cudart().cudaHostRegister(th.data_ptr(), th.numel() * 8, 0)
    th.copy_(Wd, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    torch.cuda.cudart().cudaHostUnregister(th.data_ptr())
    return out
t(registered, "np.empty + cudaHostRegister")
def chunked(nchunk=8):
    out = np.empty((m, k), dtype=np.float64)
    step = (m + nchunk - 1) // nchunk
    bufs = [torch.empty((step, k), dtype=torch.float64, pin_memory=True) for _ in range(2)]
    evs = [torch.cuda.Event() for _ in range(2)]
    pend = []
    for i in range(nchunk):
        r0, r1 = i * step, min(m, (i + 1) * step)
        b = bufs[i % 2]
        if len(pend) >= 2:
            j, q0, q1 = pend.pop(0)
            evs[j].synchronize(); out[q0:q1] = bufs[j][: q1 - q0].numpy()
        b[: r1 - r0].copy_(Wd[r0:r1], non_blocking=True); evs[i % 2].record()
        pend.append((i % 2, r0, r1))
    for j, q0, q1 in pend:
        evs[j].synchronize(); out[q0:q1] = bufs[j][: q1 - q0].numpy()
    return out
t(chunked, "chunked pinned double buffer -> np.empty")
