#!/usr/bin/env python
"""Microseconds per MM iteration of a single fit: fused small-fit kernel vs the regular launch-per-kernel path (SIMT, and the
tensor engine where eligible), over problem sizes -- sets NBMF_FUSED_MAX_WORK / the auto rule.  Fixed costs are removed
by differencing a 260- and a 60-iteration fit."""
import os, sys, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from nbmf_mm_b200 import nbmf_mm_solver

def per_iter(X, k, dtype, engine):
    ts = {}
    for iters in (60, 260):
        best = None
        for rep in range(3):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            nbmf_mm_solver(X, k, max_iter=iters, tol=0.0, random_state=0, dtype=dtype, engine=engine)
            torch.cuda.synchronize(); dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        ts[iters] = best
    return (ts[260] - ts[60]) / 200 * 1e6

shapes = [(50, 85), (100, 500), (253, 902), (1226, 285), (600, 600), (1000, 1000), (1500, 1500), (2000, 2000), (3000, 3000), (4000, 1000)]
for dtype in ("float64", "float32"):
    for k in (6, 16, 32):
        for (m, n) in shapes:
            X = (np.random.default_rng(0).random((m, n)) < 0.1).astype(np.float64)
            f = per_iter(X, k, dtype, "fused")
            r = per_iter(X, k, dtype, "simt")
            line = f"{dtype} K={k:2d} {m}x{n} ({m*n/1e6:.2f}M entries): fused {f:7.1f} us/it, regular simt {r:7.1f} us/it"
            if dtype == "float32" and m >= 512 and n >= 128:
                t = per_iter(X, k, dtype, "tensor")
                line += f", tensor {t:7.1f} us/it"
            print(line, flush=True)
