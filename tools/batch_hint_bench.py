#!/usr/bin/env python
"""Config-5 batch (64 restarts, 1226 x 285, 200 iterations, fp32) with and without batch-aware launch plans
(NBMF_BATCH_HINT: the row / column splits of a fit are chosen for the whole batch, not for one fit)."""
import os, sys, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from nbmf_mm_b200 import nbmf_mm_multifit
X = (np.random.default_rng(0).random((1226, 285)) < 0.0435).astype(np.float64)
for k in (6, 16, 32, 64):
    jobs = [dict(n_components=k, random_state=r) for r in range(64)]
    out = {}
    for hint in (1, 8, 64):
        os.environ["NBMF_BATCH_HINT"] = str(hint)
        best = None
        for rep in range(3):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            res = nbmf_mm_multifit(X, jobs, max_iter=200, tol=0.0, dtype="float32")
            torch.cuda.synchronize(); dt = time.perf_counter() - t0
            if rep: best = dt if best is None else min(best, dt)
        out[hint] = (best, min(r[2][-1] for r in res))
    print(f"K={k:2d}: " + " | ".join(f"hint {h}: {t * 1e3:.1f} ms (best loss {l:.9f})" for h, (t, l) in out.items()), flush=True)
