timeout 900 python -m pytest tests/test_gpu_ingest_eval.py tests/test_gpu_estimator.py -q -m gpu > gpurun_out/r2_tests_c34.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2_tests_c34.log
timeout 900 python bench.py --configs > gpurun_out/r2_configs_c34.log 2>&1; echo "rc=$?"; grep "cfg3\|cfg2 lastfm\|cfg1" gpurun_out/r2_configs_c34.log | cut -c1-330
python - <<'PY'
import time, numpy as np, torch, sys
sys.path.insert(0, '.')
from nbmf_mm_b200.device import pack_host_dense_checked
rng = np.random.default_rng(0)
X = (rng.random((20000, 5000)) < 0.15).astype(np.float64); M = (rng.random((20000, 5000)) < 0.9).astype(np.float64)
for kw in (dict(pinned_from=1 << 60), dict(), dict(n_threads=4), dict(n_threads=8), dict(chunk_bytes=64 << 20), dict(chunk_bytes=16 << 20)):
    ts = []
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter(); pack_host_dense_checked(X, M, None, **kw); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    print(kw, ["%.1f ms" % (t * 1e3) for t in ts], flush=True)
PY
