#!/usr/bin/env python
"""cfg5-style workload: 64 restarts on lastfm-shaped data, K sweep; sequential solver calls vs nbmf_mm_multifit."""
import sys, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from nbmf_mm_b200 import nbmf_mm_multifit, nbmf_mm_solver
rng = np.random.default_rng(0)
X = (rng.random((1226, 285)) < 0.0435).astype(np.float64)
n_init, iters = 64, 200
for k in (6, 16, 32, 64):
    for dtype in ("float32",):
        nbmf_mm_solver(X, k, max_iter=5, tol=0.0, random_state=0, dtype=dtype)         # warm-up
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for r in range(8):
            nbmf_mm_solver(X, k, max_iter=iters, tol=0.0, random_state=r, dtype=dtype)
        torch.cuda.synchronize(); t_seq = (time.perf_counter() - t0) / 8 * n_init
        res = {}
        for ns in (1, 4, 8, 16):
            jobs = [dict(n_components=k, random_state=r) for r in range(n_init)]
            torch.cuda.synchronize(); t0 = time.perf_counter()
            nbmf_mm_multifit(X, jobs, max_iter=iters, tol=0.0, dtype=dtype, n_streams=ns)
            torch.cuda.synchronize(); res[ns] = time.perf_counter() - t0
        upd = 1226 * 285 * iters * n_init
        print(f"K={k:2d} {dtype}: 64 sequential solver calls {t_seq:.2f} s ({upd / t_seq:.2e} upd/s) | multifit " +
              ", ".join(f"{ns} streams {t:.2f} s ({upd / t:.2e})" for ns, t in res.items()))
