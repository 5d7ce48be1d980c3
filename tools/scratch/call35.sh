python - <<'PY'
import time, numpy as np, torch, sys, cProfile, pstats
sys.path.insert(0, '.')
from nbmf_mm_b200 import NBMF
rng = np.random.default_rng(3)
V = (rng.random((20000, 5000)) < 0.15).astype(np.float64); mask = (rng.random((20000, 5000)) < 0.9).astype(np.float64)
def run():
    return NBMF(n_components=20, orientation="dir-beta", projection_method="duchi", max_iter=100, tol=0.0, random_state=0, dtype="float32").fit(V, mask=mask)
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter(); est = run(); torch.cuda.synchronize(); print("fit %.1f ms" % ((time.perf_counter() - t0) * 1e3), est.transfer_stats_)
pr = cProfile.Profile(); pr.enable(); run(); torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(35)
PY
