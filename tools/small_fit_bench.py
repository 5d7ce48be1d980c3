#!/usr/bin/env python
"""Small fits (config 5 shape, 1226 x 285, fp32): SIMT vs tensor engine, single fit and a batch of 64 restarts."""
import os, sys, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from nbmf_mm_b200 import nbmf_mm_multifit, nbmf_mm_solver
X = (np.random.default_rng(0).random((1226, 285)) < 0.0435).astype(np.float64)
for k in (6, 16, 32, 64):
    line = f"K={k:2d}:"
    for engine in ("simt", "tensor"):
        ref = None
        for rep in range(3):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            out = nbmf_mm_solver(X, k, max_iter=200, tol=0.0, random_state=0, dtype="float32", engine=engine)
            torch.cuda.synchronize(); dt = time.perf_counter() - t0
            if rep: ref = dt if ref is None else min(ref, dt)
        jobs = [dict(n_components=k, random_state=r) for r in range(64)]
        best = None
        for rep in range(3):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            res = nbmf_mm_multifit(X, jobs, max_iter=200, tol=0.0, dtype="float32", engine=engine)
            torch.cuda.synchronize(); dt = time.perf_counter() - t0
            if rep: best = dt if best is None else min(best, dt)
        same = np.array_equal(res[0][0], out[0]) and np.array_equal(res[0][2], out[2])
        line += f"  {engine}: single fit {ref * 1e3:.1f} ms, 64 restarts {best * 1e3:.1f} ms (best loss {min(r[2][-1] for r in res):.9f}; restart 0 == solver call: {same})"
    print(line, flush=True)
