#!/usr/bin/env python
"""One batch of 64 restarts on the config-5 shape (for ncu launch lists): python tools/small_fit_one.py K engine iters"""
import sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from nbmf_mm_b200 import nbmf_mm_multifit
k, engine, iters = int(sys.argv[1]), sys.argv[2], int(sys.argv[3])
X = (np.random.default_rng(0).random((1226, 285)) < 0.0435).astype(np.float64)
jobs = [dict(n_components=k, random_state=r) for r in range(64)]
res = nbmf_mm_multifit(X, jobs, max_iter=iters, tol=0.0, dtype="float32", engine=engine)
print(min(r[2][-1] for r in res))
