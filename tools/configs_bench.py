#!/usr/bin/env python
"""BASELINE.json configs[0], [1], [2], [4] end to end through the public API (dense fp64 inputs, as a user of the
reference would pass them), next to the CPU oracle port of the reference on the same box (bounded samples where
the full CPU run would take minutes).  configs[3] is bench.py.  One line per case; development / documentation tool.
The oracle is test infrastructure: it is only the timed CPU baseline here, as in bench.py's cpu_baseline leg."""
import sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
for p in (ROOT, ROOT / "oracle"):
    sys.path.insert(0, str(p))
import torch
import nbmf_oracle as orc
from nbmf_mm_b200 import NBMF, nbmf_mm_multifit


def timed(f, reps=2):
    best, out = None, None
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        out = f()
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return best, out


def cpu_time(f):
    t0 = time.perf_counter(); out = f(); return time.perf_counter() - t0, out


def line(name, gpu_s, gpu_iters, cpu_s, cpu_iters, entries, extra=""):
    g, c = entries * gpu_iters / gpu_s, entries * cpu_iters / cpu_s
    print(f"{name}: ours {gpu_s * 1e3:9.1f} ms for {gpu_iters} iterations ({g:.2e} updates/s) | CPU oracle {cpu_s:8.2f} s for "
          f"{cpu_iters} iterations ({c:.2e} updates/s) | ratio {g / c:8.1f}x {extra}", flush=True)


# configs[0]: quick start
X = (np.random.default_rng(0).random((100, 500)) < 0.25).astype(float)
for dtype in ("float64", "float32"):
    gs, est = timed(lambda: NBMF(n_components=6, orientation="beta-dir", alpha=1.2, beta=1.2, random_state=0, dtype=dtype).fit(X))
    cs, ref = cpu_time(lambda: orc.fit(X, 6, max_iter=2000, tol=1e-5, random_state=0))
    line(f"cfg1 quick start 100x500 K=6 {dtype}", gs, est.n_iter_, cs, ref[3], X.size,
         f"n_iter {est.n_iter_} vs {ref[3]}, final loss {est.loss_curve_[-1]:.9f} vs {ref[2][-1]:.9f}")

# configs[1]: paper datasets
z = np.load(ROOT / "tests" / "golden" / "datasets.npz")
for name in ("animals", "paleo", "lastfm"):
    n = int(z[f"{name}_shape"][1])
    D = np.unpackbits(z[f"{name}_bits"], axis=1, bitorder="little")[:, :n].astype(np.float64)
    gs, est = timed(lambda: NBMF(n_components=10, max_iter=500, tol=1e-5, random_state=0, dtype="float64").fit(D))
    cs, ref = cpu_time(lambda: orc.fit(D, 10, max_iter=500, tol=1e-5, random_state=0))
    line(f"cfg2 {name} {D.shape[0]}x{D.shape[1]} K=10 float64", gs, est.n_iter_, cs, ref[3], D.size,
         f"n_iter {est.n_iter_} vs {ref[3]}, final loss {est.loss_curve_[-1]:.9f} vs {ref[2][-1]:.9f}")

# configs[2]: masked completion, dir-beta, duchi
rng = np.random.default_rng(0)
Ws = rng.dirichlet(np.ones(20), size=20000); Hs = rng.random((20, 5000)) * 0.2
V = (rng.random((20000, 5000)) < Ws @ Hs).astype(np.float64)
mask = (rng.random((20000, 5000)) < 0.9).astype(np.float64)
gs, est = timed(lambda: NBMF(n_components=20, orientation="dir-beta", projection_method="duchi", max_iter=100, tol=0.0,
                             random_state=0, dtype="float32").fit(V, mask=mask))
cs, ref = cpu_time(lambda: orc.fit(V, 20, max_iter=2, tol=0.0, mask=mask, random_state=0, orientation="dir-beta", projection="duchi"))
line("cfg3 20000x5000 K=20 dir-beta duchi 90% mask float32", gs, 100, cs, 2, V.size, f"final loss {est.loss_curve_[-1]:.6f}")

# configs[4]: 64 restarts, K sweep on lastfm-shaped data
L = (np.random.default_rng(0).random((1226, 285)) < 0.0435).astype(np.float64)
for k in (6, 16, 32, 64):
    jobs = [dict(n_components=k, random_state=r) for r in range(64)]
    gs, _ = timed(lambda: nbmf_mm_multifit(L, jobs, max_iter=200, tol=0.0, dtype="float32"))
    cs, _ = cpu_time(lambda: orc.fit(L, k, max_iter=200, tol=0.0, random_state=0))
    line(f"cfg5 64 restarts 1226x285 K={k} float32", gs, 200 * 64, cs * 64, 200 * 64, L.size, "(CPU: one restart timed, x 64)")
