import sys, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from nbmf_mm_b200 import nbmf_mm_solver
k = int(sys.argv[1]) if len(sys.argv) > 1 else 32
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
rng = np.random.default_rng(0)
X = (rng.random((1226, 285)) < 0.0435).astype(np.float64)
nbmf_mm_solver(X, k, max_iter=3, tol=0.0, random_state=0, dtype="float32")
torch.cuda.synchronize(); t0 = time.perf_counter()
nbmf_mm_solver(X, k, max_iter=iters, tol=0.0, random_state=1, dtype="float32")
torch.cuda.synchronize(); print(f"K={k}: {iters} iterations in {(time.perf_counter()-t0)*1e3:.2f} ms")
