#!/usr/bin/env python
"""Accuracy of the fp32 engines at scale: one MM step and a short fit on a large synthetic problem,
SIMT-fp32 and tensor-fp32 against SIMT-fp64 (which is pinned to the reference at 1e-9)."""
import sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from nbmf_mm_b200.device import synth_bits_device
from nbmf_mm_b200.solver import PreparedData, make_problem

m, n, k = (int(x) for x in (sys.argv[1:4] or (200000, 20000, 32)))
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 6
hstar = (np.random.default_rng(4).random((min(k, 32), n)) * 0.2).astype(np.float32)
P, M = synth_bits_device(4, 0, m, n, hstar, 0.9, "cuda")
data = PreparedData(m, n, "bits", P, M, None, float(M.count()))
rs = np.random.RandomState(0)
W0 = rs.uniform(0.1, 0.9, (m, k)); H0 = rs.uniform(0.1, 0.9, (k, n))
out = {}
for name, dtype, engine in (("f64", "float64", "simt"), ("simt32", "float32", "simt"), ("tc32", "float32", "tensor")):
    with make_problem(data, k, dtype=dtype, alpha=1.2, beta=1.2, eps=1e-8, mask_semantics="reference",
                      projection="normalize", max_iter_cap=iters + 1, device=None, engine=engine) as prob:
        prob.set_factors(W0, H0, normalize_w=True)
        l0 = prob.objective()
        prob.h_half_step(); _, H1 = prob.get_factors()
        prob.w_half_step(); W1, _ = prob.get_factors()
        l1 = prob.objective()
        prob.set_factors(W0, H0, normalize_w=True)
        losses, nit, _ = prob.fit(iters, 0.0)
        Wf, Hf = prob.get_factors()
    out[name] = dict(l0=l0, l1=l1, H1=H1, W1=W1, losses=losses, Wf=Wf, Hf=Hf)
ref = out["f64"]
rel = lambda a, b: float(np.max(np.abs(a - b)) / np.max(np.abs(b)))
print(f"problem {m}x{n} K={k}, {iters} iterations; reference = SIMT fp64")
for name in ("simt32", "tc32"):
    o = out[name]
    print(f"{name:7s} loss(init) rel {abs(o['l0']-ref['l0'])/abs(ref['l0']):.2e} | one step: H' {rel(o['H1'], ref['H1']):.2e} "
          f"W' {rel(o['W1'], ref['W1']):.2e} loss {abs(o['l1']-ref['l1'])/abs(ref['l1']):.2e} | "
          f"fit: loss curve max rel {np.max(np.abs(o['losses']-ref['losses'])/np.abs(ref['losses'])):.2e} "
          f"final {abs(o['losses'][-1]-ref['losses'][-1])/abs(ref['losses'][-1]):.2e} "
          f"H {rel(o['Hf'], ref['Hf']):.2e} W {rel(o['Wf'], ref['Wf']):.2e} simplex {np.max(np.abs(o['Wf'].sum(1)-1)):.1e}")
print("losses f64   ", np.array2string(ref["losses"], precision=10))
print("losses tc32  ", np.array2string(out["tc32"]["losses"], precision=10))
