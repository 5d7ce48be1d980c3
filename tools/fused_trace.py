#!/usr/bin/env python
"""Per-phase times of the fused small-fit kernel (NBMF_FUSED_TRACE hook in capi.cu) on the shapes of configs 1 and 2."""
import os, sys, time
from pathlib import Path
os.environ["NBMF_FUSED_TRACE"] = "1"
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from nbmf_mm_b200 import nbmf_mm_solver
for (m, n, k, dtype) in [(50, 85, 10, "float64"), (100, 500, 6, "float64"), (100, 500, 6, "float32"), (253, 902, 10, "float64"),
                         (1226, 285, 10, "float64"), (1226, 285, 32, "float64"), (2000, 2000, 16, "float32")]:
    X = (np.random.default_rng(0).random((m, n)) < 0.1).astype(np.float64)
    best = None
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        out = nbmf_mm_solver(X, k, max_iter=300, tol=0.0, random_state=0, dtype=dtype, engine="fused")
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    print(f"{m}x{n} K={k} {dtype}: {best * 1e3:.1f} ms for 300 iterations ({best / 300 * 1e6:.1f} us per iteration incl. fixed costs)", flush=True)
