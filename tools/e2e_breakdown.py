#!/usr/bin/env python
"""Where does the end-to-end time of nbmf_mm_solver go at config 4?  (development tool; GPU box)"""
import sys, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from nbmf_mm_b200 import BitMatrix, nbmf_mm_solver
from nbmf_mm_b200.device import synth_bits_device
import nbmf_mm_b200.solver as S, nbmf_mm_b200.device as D

m, n, k, steps = (int(x) for x in (sys.argv[1:5] or (1000000, 100000, 32, 10)))
hstar = (np.random.default_rng(4).random((k, n)) * 0.2).astype(np.float32)
P, M = synth_bits_device(4, 0, m, n, hstar, 0.9, "cuda")
Ph = torch.empty(P.words.shape, dtype=torch.int32, pin_memory=True).copy_(P.words)
Mh = torch.empty(M.words.shape, dtype=torch.int32, pin_memory=True).copy_(M.words)
del P, M
torch.cuda.empty_cache()
marks = []
def wrap(mod, name):
    f = getattr(mod, name)
    def g(*a, **kw):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        r = f(*a, **kw)
        torch.cuda.synchronize(); marks.append((name, time.perf_counter() - t0))
        return r
    setattr(mod, name, g)
for cls, names in ((D.DeviceProblem, ["__init__", "set_bits", "set_factors", "fit", "simplex_deviation", "get_factors_f64", "close"]),):
    for nm in names: wrap(cls, nm)
wrap(S.PreparedData, "finish")
for rep in range(2):
    marks.clear()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out = nbmf_mm_solver(BitMatrix(Ph, (m, n)), k, max_iter=steps, tol=0.0, alpha=1.2, beta=1.2, mask=BitMatrix(Mh, (m, n)),
                         random_state=0, dtype="float32")
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"rep {rep}: total {dt:.3f} s; " + ", ".join(f"{a} {b:.3f}" for a, b in marks) + f"; unaccounted {dt - sum(b for _, b in marks):.3f}")
