#!/usr/bin/env python
"""Per-pass times of the tensor engine by K (development tool), with and without the per-column choice of the directly
accumulated plane (NBMF_TC_NOFLIP=1 turns it off: results must not change where no column flips).  Round 2 used this
script, with a library that also held the round-1 kernels, for the A/B runs quoted in DESIGN.md section 4.1."""
import os
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from nbmf_mm_b200.device import synth_bits_device
from nbmf_mm_b200.solver import PreparedData, make_problem

m, n = (int(x) for x in (sys.argv[1:3] or (65536, 32768)))
ks = [int(x) for x in (sys.argv[3].split(",") if len(sys.argv) > 3 else (32, 20, 12))]
iters = 4


def run(data, k, W0, H0, env):
    for key in ("NBMF_TC_V1", "NBMF_TC_NOFLIP"):
        os.environ.pop(key, None)
    os.environ.update(env)
    with make_problem(data, k, dtype="float32", alpha=1.2, beta=1.2, eps=1e-8, mask_semantics="reference",
                      projection="normalize", max_iter_cap=iters + 1, device=None, engine="tensor") as prob:
        prob.set_factors(W0, H0, normalize_w=True)
        l0 = prob.objective()
        prob.h_half_step(); _, H1 = prob.get_factors()
        prob.w_half_step(); W1, _ = prob.get_factors()
        prob.set_factors(W0, H0, normalize_w=True)
        prob.profile(True)
        losses, _, _ = prob.fit(iters, 0.0)
        hm, hc, wm, wc = prob.profile_read()
        prob.profile(False)
    return dict(l0=l0, H1=H1, W1=W1, losses=np.asarray(losses), h_ms=hm / max(hc, 1), w_ms=wm / max(wc, 1))


for k in ks:
    hstar = (np.random.default_rng(4).random((min(k, 32), n)) * 0.2).astype(np.float32)
    P, M = synth_bits_device(4, 0, m, n, hstar, 0.9, "cuda")
    data = PreparedData(m, n, "bits", P, M, None, float(M.count()))
    rs = np.random.RandomState(0)
    W0, H0 = rs.uniform(0.1, 0.9, (m, k)), rs.uniform(0.1, 0.9, (k, n))
    v2 = run(data, k, W0, H0, {"NBMF_TC_NOFLIP": "1"})
    v2f = run(data, k, W0, H0, {})
    line = f"K={k:2d} {m}x{n}: no-flip H {v2['h_ms']:.3f} ms W {v2['w_ms']:.3f} ms | default H {v2f['h_ms']:.3f} ms W {v2f['w_ms']:.3f} ms"
    line += f" | flip vs noflip max |dH| {np.max(np.abs(v2f['H1'] - v2['H1'])):.2e} losses {v2['losses'][-1]:.9f} {v2f['losses'][-1]:.9f}"
    print(line, flush=True)
    del P, M, data
