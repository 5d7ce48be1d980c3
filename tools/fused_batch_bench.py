#!/usr/bin/env python
"""Batches of restarts that are not tensor-eligible (fp64, or too few rows / columns): fused small-fit kernel (a few CTAs per
fit) vs the regular batched launches (NBMF_NO_FUSED=1)."""
import os, sys, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from nbmf_mm_b200 import nbmf_mm_multifit
for (m, n, dtype) in [(1226, 285, "float64"), (253, 902, "float64"), (100, 500, "float32"), (400, 285, "float32")]:
    X = (np.random.default_rng(0).random((m, n)) < 0.0435).astype(np.float64)
    for k in (6, 16, 32):
        for B in (8, 64):
            jobs = [dict(n_components=k, random_state=r) for r in range(B)]
            line = f"{m}x{n} {dtype} K={k:2d} batch of {B:2d}:"
            for fused in (True, False):
                best = None
                for rep in range(3):
                    torch.cuda.synchronize(); t0 = time.perf_counter()
                    res = nbmf_mm_multifit(X, jobs, max_iter=200, tol=0.0, dtype=dtype, engine="fused" if fused else "simt")
                    torch.cuda.synchronize(); dt = time.perf_counter() - t0
                    if rep: best = dt if best is None else min(best, dt)
                line += f"  {'fused' if fused else 'regular'} {best * 1e3:7.1f} ms"
            print(line, flush=True)
