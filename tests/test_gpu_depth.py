"""Parity of the tcgen05 engine at the DEPTH of BASELINE.json config 4 (10^6 x 10^5, K=32): what the small oracle cases
cannot reach is the length of the accumulation chains one CTA runs through -- 10 417 row blocks (651 TMEM flush chains)
per H-pass CTA and 1 563 column blocks per W-pass CTA at config 4.  The fp64 SIMT engine is the checker here: it is
pinned to the reference at 1e-9 by the golden one-step vectors and trajectories (test_gpu_onestep / test_gpu_trajectory).

Also: the regime in which forming D = S - C (or C = S - D) from the unmasked sum cancels (theta on the ones tiny, so the
ones' ratios dominate S), on the tensor engine, against the oracle."""
import numpy as np
import pytest

import nbmf_oracle as orc
from conftest import rel_err
from nbmf_mm_b200 import nbmf_mm_update_beta_dir
from nbmf_mm_b200.device import synth_bits_device
from nbmf_mm_b200.solver import PreparedData, make_problem

pytestmark = pytest.mark.gpu


def _run(data, k, dtype, engine, W0, H0, iters):
    with make_problem(data, k, dtype=dtype, alpha=1.2, beta=1.2, eps=1e-8, mask_semantics="reference",
                      projection="normalize", max_iter_cap=iters + 1, device=None, engine=engine) as prob:
        assert prob.engine == engine
        plan = prob.plan_info()
        prob.set_factors(W0, H0, normalize_w=True)
        prob.h_half_step(); _, H1 = prob.get_factors()
        prob.w_half_step(); W1, _ = prob.get_factors()
        prob.set_factors(W0, H0, normalize_w=True)
        losses, _, _ = prob.fit(iters, 0.0)
    return H1, W1, np.asarray(losses), plan


@pytest.mark.parametrize("m,n,k,what", [
    (400_000, 18_944, 32, "H pass: 148 column blocks x 1 row split -> 12 500 row blocks per CTA (config 4: 10 417)"),
    (18_944, 100_000, 32, "W pass: 148 row blocks x 1 column split -> 1 563 column blocks per CTA (config 4: 1 563)"),
    (131_072, 18_944, 64, "K <= 64 kernels, H pass: 4 096 row blocks = 8 192 half-blocks per CTA over three pipelines"),
    (18_944, 50_048, 48, "K <= 64 kernels, W pass: 782 column blocks per CTA"),
    (200_000, 18_944, 12, "K <= 16 instantiation, H pass: 6 250 row blocks per CTA"),
])
def test_tensor_engine_at_config4_depth_vs_fp64(m, n, k, what):
    iters = 3
    hstar = (np.random.default_rng(4).random((min(k, 32), n)) * 0.2).astype(np.float32)
    P, M = synth_bits_device(4, 0, m, n, hstar, 0.9, "cuda")
    data = PreparedData(m, n, "bits", P, M, None, float(M.count()))
    rs = np.random.RandomState(0)
    W0, H0 = rs.uniform(0.1, 0.9, (m, k)), rs.uniform(0.1, 0.9, (k, n))
    Hr, Wr, Lr, _ = _run(data, k, "float64", "simt", W0, H0, iters)
    Ht, Wt, Lt, plan = _run(data, k, "float32", "tensor", W0, H0, iters)
    if m > n:
        assert plan["h_row_splits"] == 1, plan           # the whole row range is ONE chain of blocks per CTA
    else:
        assert plan["w_col_splits"] == 1, plan
    assert rel_err(Ht, Hr) < 1e-5, what
    assert rel_err(Wt, Wr) < 1e-5, what
    assert np.max(np.abs(Lt - Lr) / np.abs(Lr)) < 1e-6, what
    assert np.all(np.diff(Lt) < 0)


@pytest.mark.parametrize("h_lo,h_hi,density,alpha", [(0.5e-4, 1.5e-4, 0.01, 0.5), (1.5e-4, 4.5e-4, 0.05, 1.2),
                                                      (0.45, 0.9, 0.002, 1.2)])
def test_tensor_engine_where_the_unmasked_sum_cancels(h_lo, h_hi, density, alpha):
    """C >> D (Theta far below the density of ones: every one contributes 1/theta ~ 1e4, every zero ~1) and D >> C
    (Theta far above it).  Whichever of C, D the kernel derives by subtraction from the unmasked sum loses
    log2(big / small) bits there; the H' bar of the fp32 engines (5e-5 relative) must still hold."""
    m, n, k = 4096, 1536, 32
    rng = np.random.default_rng(7)
    Y = (rng.random((m, n)) < density).astype(np.float64)
    mask = (rng.random((m, n)) < 0.9).astype(np.float64)
    W = rng.uniform(0.1, 0.9, (k, m)); W /= W.sum(axis=0, keepdims=True)
    H = rng.uniform(h_lo, h_hi, (k, n))
    Wo, Ho = orc.mm_step(Y, W, H, mask, alpha, 1.2)
    pos = Y * mask
    C = W @ (pos / (W.T @ H + 1e-8)); D = W @ ((1 - pos) / (1 - W.T @ H + 1e-8))
    ratio = float(np.median(C / D))
    assert ratio > 30 or ratio < 1 / 30, ratio                      # the regime this test is about
    W1, H1 = nbmf_mm_update_beta_dir(Y, W, H, mask, alpha, 1.2, dtype="float32", engine="tensor")
    assert rel_err(H1, Ho) < 5e-5, (ratio, rel_err(H1, Ho))
    if alpha >= 1.0:                                                 # (alpha < 1: the numerator H C + (alpha - 1) itself cancels in
        big = Ho > 1e-5                                              #  fp32 wherever H C ~ 1 - alpha, in every fp32 engine)
        assert np.max(np.abs(H1 - Ho)[big] / Ho[big]) < 2e-4, ratio  # element-wise too: H' itself is small here
    assert rel_err(W1, Wo) < 5e-5
    Ws, Hs = nbmf_mm_update_beta_dir(Y, W, H, mask, alpha, 1.2, dtype="float32", engine="simt")
    assert rel_err(Hs, Ho) < 5e-5 and rel_err(Ws, Wo) < 5e-5
