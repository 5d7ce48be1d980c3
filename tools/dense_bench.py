#!/usr/bin/env python
"""Dense (probabilistic) V: the HBM-bound side of the metric (BASELINE.json: 'HBM GB/s for dense float V').
One JSON line per storage layout (fp32, fp16): updates/s and the achieved fraction of the measured HBM roofline
of the two pass kernels.  Synthetic V in (0,1), 90 % mask, small K so that the passes are bandwidth bound."""
import json, sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch
from nbmf_mm_b200.device import DeviceProblem, pack_bits_device, pack_dense_device

m, n, k, steps = (int(x) for x in (sys.argv[1:5] or (131072, 32768, 8, 5)))
dev = torch.device("cuda")
peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {}
hbm = peaks.get("hbm_gbs", 6650.0)
g = torch.Generator(device=dev); g.manual_seed(0)
X = torch.rand((m, n), generator=g, device=dev, dtype=torch.float32) * 0.5
mask = (torch.rand((m, n), generator=g, device=dev, dtype=torch.float32) < 0.9).to(torch.uint8)
M, _ = pack_bits_device(mask, None)
n_obs = float(M.count())
for storage, vkind in (("float32", "dense"), ("float16", "dense16")):
    Vm = pack_dense_device(X, mask, np.float16 if storage == "float16" else np.float32)
    prob = DeviceProblem(m, n, k, dtype="float32", vkind=vkind, has_mask=True, alpha=1.2, beta=1.2, eps=1e-8, n_obs=n_obs,
                         max_iter_cap=steps + 8, device=dev)
    prob.set_dense(Vm, M)
    W0 = torch.rand((m, k), generator=g, device=dev) * 0.8 + 0.1
    H0 = torch.rand((k, n), generator=g, device=dev) * 0.8 + 0.1
    prob.set_factors(W0, H0, normalize_w=True)
    prob.fit_begin(steps + 4, 0.0)
    prob.fit_enqueue(3)
    torch.cuda.synchronize()
    prob.profile(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); prob.fit_enqueue(steps); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    h_ms, h_cnt, w_ms, w_cnt = prob.profile_read()
    prob.profile(False)
    done, nh = prob.fit_poll(wait=True); hist, _ = prob.fit_history(nh)
    vb = 2 if storage == "float16" else 4
    h_bytes, w_bytes = m * n * vb, m * n * (vb + 0.125)          # H pass: V*mask; W pass: V*mask + mask bits
    line = {"metric": "observed-entry MM updates/s (M*N*iters/s)", "value": m * n * steps / (ms * 1e-3), "unit": "updates/s",
            "n_gpus": 1, "steps": steps, "ms_per_step": ms / steps, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"dense probabilistic V {m}x{n}, K={k}, 90% observed, V*mask stored as {storage}", "engine": prob.engine},
            "roofline": {"bound": "hbm", "kernel": "w_pass_kernel (dense V)", "unit": "GB/s", "peak": hbm,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6.65 TB/s",
                         "achieved": w_bytes / (w_ms / w_cnt * 1e-3) * 1e-9, "frac": w_bytes / (w_ms / w_cnt * 1e-3) * 1e-9 / hbm,
                         "avg_launch_ms": w_ms / w_cnt, "traffic": None,
                         "h_pass": {"achieved": h_bytes / (h_ms / h_cnt * 1e-3) * 1e-9, "frac": h_bytes / (h_ms / h_cnt * 1e-3) * 1e-9 / hbm,
                                    "avg_launch_ms": h_ms / h_cnt},
                         "algorithmic_bytes_per_entry": {"h_pass": vb, "w_pass": vb + 0.125}},
            "loss_first_last": [float(hist[0]), float(hist[-1])]}
    print(json.dumps(line), flush=True)
    prob.close(); del Vm
