#!/usr/bin/env python
"""One small fit through the fused kernel (for ncu): python tools/fused_one.py M N K DTYPE ITERS"""
import sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from nbmf_mm_b200 import nbmf_mm_solver
m, n, k, dtype, iters = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4], int(sys.argv[5])
X = (np.random.default_rng(0).random((m, n)) < 0.1).astype(np.float64)
out = nbmf_mm_solver(X, k, max_iter=iters, tol=0.0, random_state=0, dtype=dtype, engine="fused")
print(out[2][-1], out[4])
