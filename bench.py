#!/usr/bin/env python
"""Headline benchmark: observed-entry MM updates/s (M*N*iters/s) of the NBMF-MM fit loop.

  python bench.py --gpus N --steps K --warmup W            # our arm (B200 kernels)
  python bench.py --impl reference --steps K --warmup W    # CPU arm: the oracle port of the reference

A "step" is one MM iteration (H half-step + loss/stop rule + W half-step, reference
_solver.py:143-175) over the whole matrix.  Workload = BASELINE.json configs[3]: synthetic
bit-packed binary 1,000,000 x 100,000, K=32, 90 % observed, beta-dir, normalize, FP32,
row-sharded over the N GPUs (fixed total size: strong scaling) with one NCCL allreduce of the
K x N H-partials per iteration.  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
for p in (ROOT, ROOT / "oracle"):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))

METRIC = "observed-entry MM updates/s (M*N*iters/s)"
# tensor-pipe occupancy of the tensor engine relative to the algorithmic flop, in TF32-rate units (one unit = one M=128,
# N=32 MMA instruction = 16 clk; a bf16 MMA covers twice the K extent in the same time): every product chain is one TF32
# chain (hi.hi) plus one bf16 chain (hi.lo + lo.hi), i.e. 2 units per algorithmic unit in both passes.
H_EXEC, W_EXEC = 2.0, 2.0
UNIT = "updates/s"
SEED = 4


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows", type=int, default=1_000_000)
    ap.add_argument("--cols", type=int, default=100_000)
    ap.add_argument("--k", type=int, default=32)
    ap.add_argument("--obs", type=float, default=0.9)
    ap.add_argument("--dtype", default="float32", choices=["float32", "float64"])
    ap.add_argument("--engine", default="auto", choices=["auto", "simt", "tensor"])
    ap.add_argument("--configs", action="store_true", help="time BASELINE configs 1, 2, 3, 5 end to end instead (text lines)")
    ap.add_argument("--restarts-only", action="store_true", help="with --configs: only configs[4] (64 restarts, K sweep), as the multi-GPU mode runs it")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-warmup", type=int, default=1, help="untimed end-to-end calls before the timed one")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the post-run parity check against the oracle")
    ap.add_argument("--cpu-rows", type=int, default=0, help="rows of the CPU-baseline sample (0 = auto)")
    return ap.parse_args()


def workload_config(a, extra=None):
    cfg = {
        "workload": f"configs[3]: synthetic bit-packed binary {a.rows}x{a.cols}, K={a.k}, {int(a.obs * 100)}% observed, "
                    f"beta-dir, normalize, alpha=beta=1.2, row-sharded with NCCL allreduce of H partials",
        "rows": a.rows, "cols": a.cols, "k": a.k, "observed_fraction": a.obs,
        "orientation": "beta-dir", "projection_method": "normalize", "mask_semantics": "reference",
        "l2": "inputs (bit planes, 2 x rows x cols / 8 bytes per pass) are far larger than the 126 MB L2; no flush needed",
    }
    if extra:
        cfg.update(extra)
    return cfg


# --------------------------------------------------------------------------- clocks
class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons through NVML every 200 ms while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
        }
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# --------------------------------------------------------------------------- CPU arm
def cpu_port_run(a, steps, warmup, rows=None, cols=None):
    """Time the oracle port of the reference (NumPy + BLAS, all host threads) on a bounded sample of
    the workload: a `rows x cols` block with the same K, mask fraction and generative recipe."""
    import nbmf_oracle as orc
    # all the host cores this process may use, set explicitly: torchrun exports OMP_NUM_THREADS=1 to its workers, which
    # would time the reference's BLAS single-threaded at N > 1
    try:
        ncores = len(os.sched_getaffinity(0))
    except Exception:
        ncores = os.cpu_count() or 1
    limiter = None
    try:
        from threadpoolctl import threadpool_limits
        limiter = threadpool_limits(limits=ncores)
    except Exception:
        pass
    total = max(1, steps + warmup)
    if rows is None:
        rows = a.cpu_rows or (4096 if total <= 6 else 2048)
    cols = cols or 4096
    rows, cols = min(rows, a.rows), min(cols, a.cols)
    rng = np.random.default_rng(SEED)
    Wst = rng.dirichlet(np.ones(a.k), size=rows)
    Hst = rng.random((a.k, cols)) * 0.2
    Y = (rng.random((rows, cols)) < Wst @ Hst).astype(np.float64)
    mask = (rng.random((rows, cols)) < a.obs).astype(np.float64)
    rs = np.random.RandomState(0)
    W = rs.uniform(0.1, 0.9, (rows, a.k)).T
    W = W / W.sum(axis=0, keepdims=True)
    H = rs.uniform(0.1, 0.9, (a.k, cols))
    for _ in range(warmup):
        W, H = orc.mm_step(Y, W, H, mask, 1.2, 1.2)
        orc.map_objective(Y, W, H, mask, 1.2, 1.2)
    t0 = time.perf_counter()
    for _ in range(steps):
        W, H = orc.mm_step(Y, W, H, mask, 1.2, 1.2)            # one reference iteration ...
        orc.map_objective(Y, W, H, mask, 1.2, 1.2)             # ... including its per-iteration loss
    dt = time.perf_counter() - t0
    try:
        from threadpoolctl import threadpool_info
        blas_threads = max([d.get("num_threads", 1) for d in threadpool_info()] or [1])
    except Exception:
        blas_threads = os.cpu_count() or 1
    if limiter is not None:
        limiter.restore_original_limits()
    return {
        "value": rows * cols * steps / dt, "unit": UNIT, "cores": int(blas_threads), "kind": "port",
        "sample": f"{rows}x{cols} block of the workload (same K={a.k}, {int(a.obs * 100)}% mask), {steps} iterations, "
                  f"{dt / max(steps, 1):.2f} s/iter; oracle/nbmf_oracle.py (NumPy fp64 + BLAS, {blas_threads} BLAS threads "
                  f"set explicitly, {os.cpu_count()} host cpus; masked copies built once per iteration as _solver.py:21-32: "
                  f"within 3 % of the unmodified reference in the authoring container)",
    }, dt


def run_reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    base, dt = cpu_port_run(a, a.steps, a.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": a.gpus,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": dt / max(a.steps, 1) * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(a, {"note": "reference arm: CPU oracle port timed on a bounded sample, see cpu_baseline.sample"}),
        "cpu_baseline": base,
        "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------- parity of the timed configuration
def parity_check(prob, P, Mk, a, m_local, world, rank, dev, tol_factor=5e-5, tol_loss=1e-5):
    """Check the configuration that was just timed against the CPU oracle (fp64), at its own size and launch plan.

    From the factors (W_t, H_t) the timed iterations left on the device, one more H half-step and W half-step run on the
    device; the host then recomputes, in fp64 from the same W_t / H_t and the same bits (downloaded from the device
    planes), (1) H' for one 128-column block over ALL rows (`_solver.py:39-47`; row shards add their C / D partials),
    (2) the log-likelihood sum of that column block (`_solver.py:150-161`), compared with the per-CTA partials of the
    fused NLL, and (3) W' for one 128-row block over ALL columns (`_solver.py:50-57`).  The oracle is the checker."""
    import torch
    import torch.distributed as dist
    import nbmf_oracle as orc
    from nbmf_mm_b200 import _lib
    K, N, eps, alpha, beta = a.k, a.cols, 1e-8, 1.2, 1.2
    t_start = time.perf_counter()
    W_t, H_t = prob.get_factors_device()
    prob.h_half_step()
    ll_part = prob.loglik_partials()                       # (row splits, column blocks) of the H pass that just ran
    _, H_n = prob.get_factors_device()
    prob.w_half_step()
    W_n, _ = prob.get_factors_device()
    bw = 128
    if prob.engine != "tensor":
        import ctypes
        hc = ctypes.c_int32(0)
        _lib.check(_lib.load().nbmf_variant_info(0 if a.dtype == "float32" else 1, 0, K, ctypes.byref(hc), None, None))
        bw = int(hc.value)
    ncb = ll_part.shape[1]
    jb = min(ncb // 2, max(0, N // bw - 1))                # a column block fully inside n
    c0, c1 = jb * bw, min(N, jb * bw + bw)
    Hj = H_t[:, c0:c1].double().cpu().numpy()
    Pj = P.words[:, c0 // 32:(c1 + 31) // 32].contiguous().cpu().numpy().view(np.uint8)
    Wh = W_t.double().cpu().numpy()
    Cs, Ds, ll = np.zeros((K, c1 - c0)), np.zeros((K, c1 - c0)), 0.0
    for i0 in range(0, m_local, 1 << 16):
        i1 = min(m_local, i0 + (1 << 16))
        pos = np.unpackbits(Pj[i0:i1], axis=1, bitorder="little")[:, :c1 - c0].astype(np.float64)
        Wc = Wh[i0:i1]
        th = Wc @ Hj                                        # _solver.py:39
        Cs += Wc.T @ (pos / (th + eps))                     # :42
        Ds += Wc.T @ ((1 - pos) / (1 - th + eps))           # :43 (reference mask quirk: neg = 1 - Y*mask)
        ll += float(np.sum(pos * np.log(th + eps) + (1 - pos) * np.log(1 - th + eps)))   # :153-154
    ll_dev = float(ll_part[:, jb].sum())
    if world > 1:
        t = torch.from_numpy(np.concatenate([Cs.ravel(), Ds.ravel(), [ll, ll_dev]])).to(dev)
        dist.all_reduce(t)
        t = t.cpu().numpy()
        Cs, Ds = t[:Cs.size].reshape(Cs.shape), t[Cs.size:2 * Cs.size].reshape(Cs.shape)
        ll, ll_dev = float(t[-2]), float(t[-1])
    num = Hj * Cs + (alpha - 1)
    den = (1 - Hj) * Ds + (beta - 1)
    H_ref = np.clip(num / (num + den + eps), eps, 1 - eps)  # :46-47
    H_dev = H_n[:, c0:c1].double().cpu().numpy()
    h_rel = float(np.max(np.abs(H_dev - H_ref)) / np.max(np.abs(H_ref)))
    loss_rel = abs(ll_dev - ll) / abs(ll)
    # W' of one 128-row block of rank 0's shard over all columns, with the device's own H'
    w_rel, rows = None, None
    if rank == 0:
        i0 = (m_local // 2) // 128 * 128
        i1 = min(m_local, i0 + 128)
        Pi = P.rows(i0, i1).to_dense()
        Mi = Mk.rows(i0, i1).to_dense()
        W_ref = orc.w_half_step(Pi, Wh[i0:i1].T, H_n.double().cpu().numpy(), Mi, eps)     # (K x rows)
        W_dev = W_n[i0:i1].double().cpu().numpy().T
        w_rel = float(np.max(np.abs(W_dev - W_ref)) / np.max(np.abs(W_ref)))
        rows = [int(i0), int(i1)]
    ok = bool(h_rel < tol_factor and loss_rel < tol_loss and (w_rel is None or w_rel < tol_factor))
    return {"ok": ok, "h_rel": h_rel, "w_rel": w_rel, "loss_rel": float(loss_rel), "tol": {"h_rel": tol_factor, "w_rel": tol_factor, "loss_rel": tol_loss},
            "checker": "oracle/nbmf_oracle.py formulas in fp64 on the host, same W_t / H_t / bits as the device",
            "h_block": {"columns": [int(c0), int(c1)], "rows": "all (row shards all-reduce their fp64 C / D / LL partials)"},
            "w_block": {"rows_of_rank0": rows, "columns": "all"},
            "loss": "log-likelihood sum of the H column block vs the fused-NLL partials of the same launch",
            "seconds": time.perf_counter() - t_start}


# --------------------------------------------------------------------------- our arm
def run_ours(a):
    import torch
    import torch.distributed as dist
    from nbmf_mm_b200 import BitMatrix, _lib, nbmf_mm_solver
    from nbmf_mm_b200.device import DeviceProblem, synth_bits_device
    from nbmf_mm_b200.solver import _row_shard

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (ours) needs a GPU: there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    if world != a.gpus and rank == 0:
        print(f"# note: --gpus {a.gpus} but WORLD_SIZE={world}; using {world}", file=sys.stderr)
    lib = _lib.load()

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()

    def max_over_ranks(x):
        if world == 1:
            return float(x)
        t = torch.tensor([float(x)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    M_rows, N, K = a.rows, a.cols, a.k
    r0, r1 = _row_shard(M_rows, rank, world)
    m_local = r1 - r0
    hstar = (np.random.default_rng(SEED).random((K if K <= 32 else 32, N)) * 0.2).astype(np.float32)
    P, Mk = synth_bits_device(SEED, r0, m_local, N, hstar, a.obs, dev)
    n_obs_local = float(Mk.count())
    n_obs = n_obs_local
    if world > 1:
        t = torch.tensor([n_obs_local], dtype=torch.float64, device=dev)
        dist.all_reduce(t)
        n_obs = float(t.item())

    steps, warmup = a.steps, a.warmup
    prob = DeviceProblem(m_local, N, K, dtype=a.dtype, vkind="bits", has_mask=True, alpha=1.2, beta=1.2, eps=1e-8,
                         n_obs=n_obs, max_iter_cap=steps + warmup + 2, device=dev, engine=a.engine)
    prob.set_bits(P, Mk)
    if world > 1:
        prob.init_comm()
    tdt = torch.float32 if a.dtype == "float32" else torch.float64
    g = torch.Generator(device=dev); g.manual_seed(1234 + rank)
    W0 = torch.rand((m_local, K), generator=g, device=dev, dtype=tdt) * 0.8 + 0.1
    H0 = torch.from_numpy(np.random.RandomState(0).uniform(0.1, 0.9, (K, N))).to(dev, tdt)
    prob.set_factors(W0, H0, normalize_w=True)

    # ---- device-resident timing: W warm-up iterations, then exactly K timed iterations
    prob.fit_begin(steps + warmup + 1, 0.0)
    prob.fit_enqueue(warmup)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    prob.profile(True)
    lib.nbmf_launch_count(1)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    prob.fit_enqueue(steps)
    ev1.record()
    barrier()
    sampler.stop_flag = True
    launches = int(lib.nbmf_launch_count(0))
    ms_total = max_over_ranks(ev0.elapsed_time(ev1))
    h_ms, h_cnt, w_ms, w_cnt = prob.profile_read()
    prob.profile(False)
    done, n_hist = prob.fit_poll(wait=True)
    hist, _ = prob.fit_history(n_hist)
    sampler.join(timeout=2)
    value = M_rows * N * steps / (ms_total * 1e-3)
    plan = prob.plan_info()
    parity = None if a.no_parity else parity_check(prob, P, Mk, a, m_local, world, rank, dev)

    # ---- FMA-pipe peak (roofline denominator), measured on this GPU right after the timed region
    scratch = torch.zeros(16, dtype=torch.float32, device=dev)
    peak = C_double()
    _lib.check(lib.nbmf_fma_peak(0 if a.dtype == "float32" else 1, 4000, scratch.data_ptr(),
                                 torch.cuda.current_stream(dev).cuda_stream, peak.ref()), "nbmf_fma_peak")
    peak_tf = peak.value
    h_avg_ms = h_ms / max(h_cnt, 1)
    w_avg_ms = w_ms / max(w_cnt, 1)
    entries_local = float(m_local) * N
    h_flop = 6.0 * K * entries_local          # Theta dot 2K + C and D accumulations 4K flop per entry
    w_flop = 4.0 * K * entries_local          # Theta' dot 2K + one accumulation of H'(p-q) 2K
    h_ach = h_flop / (h_avg_ms * 1e-3) * 1e-12 if h_cnt else None
    w_ach = w_flop / (w_avg_ms * 1e-3) * 1e-12 if w_cnt else None
    nominal = 2 * 128 * 148 * (sampler.max_mhz or 1965) * 1e6 * 1e-12
    engine = prob.engine
    mp = measured_peaks()
    common = {
        "unit": "TFLOP/s", "traffic": None, "engine": engine,
        "algorithmic_flop_per_entry": {"h_pass": 6 * K, "w_pass": 4 * K, "iteration": 10 * K},
        "avg_launch_ms": h_avg_ms, "launches_timed": h_cnt, "share_of_step": h_ms / ms_total if ms_total else None,
        "fp32_simt_peak": {"value": peak_tf, "source": "measured in this run: nbmf_fma_peak (packed FFMA2 chains, 148x8 CTAs); "
                           "MEASURED_PEAKS.json has no FP32 entry", "nominal": nominal,
                           "h_pass_frac": (h_ach / peak_tf) if h_ach else None, "w_pass_frac": (w_ach / peak_tf) if w_ach else None,
                           "iteration_frac": (10.0 * K * entries_local * steps / (ms_total * 1e-3) * 1e-12 / peak_tf)},
        "hbm_context": {"algorithmic_bytes_per_entry_iteration": 0.375, "hbm_gbs_used": 0.375 * value / world * 1e-9,
                        "hbm_peak_gbs": mp.get("hbm_gbs")},
    }
    try:   # DRAM bytes per launch of the dominant kernel from the committed ncu capture of this exact shape
        tr = json.loads((ROOT / "profiles" / "dram_traffic.json").read_text()).get(f"{m_local}x{N}x{K}:{engine}")
    except Exception:
        tr = None
    if tr:
        common["traffic"] = tr["h_pass"]
        common["traffic_detail"] = dict(tr, unit="bytes per launch (regenerated per build by tools/ncu_regen.py)")
    if engine == "tensor":
        # kind::tf32 runs at half the bf16 rate; the kernels are timed inside a long step -> sustained figure
        bf16 = mp.get("bf16_tflops_sustained") or mp.get("bf16_tflops")
        tpeak = bf16 / 2 if bf16 else 1590.0 / 2
        src = ("MEASURED_PEAKS.json bf16_tflops_sustained / 2 (the file has no TF32 entry; kind::tf32 is half the bf16 rate)"
               if bf16 else "fallback 1.59 PFLOP/s bf16 / 2 (B200_PROFILING.md); MEASURED_PEAKS.json absent")
        roofline = dict(common, **{
            "bound": "tensor", "kernel": "h_pass_tc_kernel (H half-step + fused NLL, tcgen05 TF32 + bf16 split precision)",
            "achieved": h_ach, "peak": tpeak, "frac": (h_ach / tpeak) if h_ach else None, "peak_source": src,
            "executed_tensor_tflops": H_EXEC * h_ach if h_ach else None,
            "frac_executed": (H_EXEC * h_ach / tpeak) if h_ach else None,
            "note": "achieved = algorithmic flop (6K per entry) / launch time; the split-precision scheme occupies the tensor "
                    "pipe for 2x that in TF32-rate units: hi.hi as a TF32 MMA chain plus one bf16 MMA chain for the "
                    "two correction terms hi.lo + lo.hi (frac_executed), so frac cannot exceed 1/2",
            "w_pass": {"kernel": "w_pass_tc_kernel", "achieved": w_ach, "frac": (w_ach / tpeak) if w_ach else None,
                       "frac_executed": (W_EXEC * w_ach / tpeak) if w_ach else None, "avg_launch_ms": w_avg_ms,
                       "share_of_step": w_ms / ms_total if ms_total else None},
            "iteration_frac": (10.0 * K * entries_local * steps / (ms_total * 1e-3) * 1e-12 / tpeak),
        })
        # what actually bounds the two kernels: warp-instruction issue slots of the SIMT stage between the two MMAs
        # (ratio arithmetic, tf32 / bf16 splits, masking, TMEM traffic, barrier code).  Executed warp instructions per
        # entry from the committed ncu capture x entries / launch time, against 4 issue slots per SM and clock.
        clk = (sampler.summary().get("sm_mhz") or sampler.max_mhz or 1965) * 1e6
        if tr and tr.get("h_pass_warp_instructions") and tr.get("w_pass_warp_instructions"):
            wi = {"h_pass": tr["h_pass_warp_instructions"] / entries_local, "w_pass": tr["w_pass_warp_instructions"] / entries_local}
            roofline["issue_slots"] = {
                "warp_instructions_per_entry": wi,
                "source": f"smsp__inst_executed.sum of this shape in {tr.get('source')} (tools/ncu_regen.py, this build)",
                "peak": "4 warp instructions per clock and SM x 148 SMs at the median SM clock of the timed region",
                "h_pass_frac": (wi["h_pass"] * entries_local / (h_avg_ms * 1e-3) / (4 * 148 * clk)) if h_cnt else None,
                "w_pass_frac": (wi["w_pass"] * entries_local / (w_avg_ms * 1e-3) / (4 * 148 * clk)) if w_cnt else None,
            }
    else:
        roofline = dict(common, **{
            "bound": "fp32" if a.dtype == "float32" else "fp64", "kernel": "h_pass_kernel (H half-step + fused NLL)",
            "achieved": h_ach, "peak": peak_tf, "frac": (h_ach / peak_tf) if h_ach else None,
            "peak_source": common["fp32_simt_peak"]["source"],
            "nominal_peak": nominal, "frac_of_nominal": (h_ach / nominal) if h_ach else None,
            "w_pass": {"achieved": w_ach, "frac": (w_ach / peak_tf) if w_ach else None, "avg_launch_ms": w_avg_ms,
                       "share_of_step": w_ms / ms_total if ms_total else None},
            "iteration_frac": (10.0 * K * entries_local * steps / (ms_total * 1e-3) * 1e-12 / peak_tf),
        })

    # ---- end-to-end through the public API with HOST buffers (pinned), H2D/D2H inside the timed region
    e2e = None
    if not a.no_e2e:
        Ph = torch.empty(P.words.shape, dtype=torch.int32, pin_memory=True).copy_(P.words)
        Mh = torch.empty(Mk.words.shape, dtype=torch.int32, pin_memory=True).copy_(Mk.words)
        prob.close()
        del P, Mk, W0, H0
        torch.cuda.empty_cache()
        def e2e_call(st):
            return nbmf_mm_solver(BitMatrix(Ph, (m_local, N)), K, max_iter=steps, tol=0.0, alpha=1.2, beta=1.2,
                                  mask=BitMatrix(Mh, (m_local, N)), random_state=0, dtype=a.dtype, device=dev,
                                  distributed=True, shard=(r0, M_rows), stats=st, engine=a.engine)
        cold = None
        for i in range(max(0, a.e2e_warmup)):              # like the W warm-up steps of the device-timed figure: page-locked
            barrier()                                      # factor staging, allocator pools, the NCCL communicator and the
            t0 = time.perf_counter()                       # jump polynomials of the init stream exist afterwards.  The
            e2e_call({})                                   # first call is timed too and reported as `cold_seconds`.
            barrier()
            if i == 0:
                cold = max_over_ranks(time.perf_counter() - t0)
        stats = {}
        barrier()
        t0 = time.perf_counter()
        out = e2e_call(stats)
        barrier()
        dt = max_over_ranks(time.perf_counter() - t0)
        assert out[4] == steps
        e2e = {"value": M_rows * N * steps / dt, "unit": UNIT,
               "h2d_bytes_per_step": stats["h2d_bytes"] / steps, "d2h_bytes_per_step": stats["d2h_bytes"] / steps,
               "seconds": dt, "warmup_calls": max(0, a.e2e_warmup), "cold_seconds": cold,
               "cold_value": (M_rows * N * steps / cold) if cold else None, "api": "nbmf_mm_b200.nbmf_mm_solver(BitMatrix(pinned host), mask=BitMatrix(pinned host), "
                                    "max_iter=steps, tol=0, dtype=float32): H2D of both bit planes and the inits, "
                                    "the fit loop, D2H of W, H and the loss history",
               "final_loss": float(out[2][-1])}
    else:
        prob.close()

    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu:
        cpu, _ = cpu_port_run(a, 2, 1, rows=a.cpu_rows or 4096)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": ms_total / steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32" if a.dtype == "float32" else "f64", "data": "synthetic",
            "config": workload_config(a, {"rows_per_gpu": m_local, "launch_plan": plan, "n_obs": n_obs,
                                          "loss_first_last": [float(hist[0]), float(hist[-1])] if len(hist) else None}),
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches,
            "clocks": sampler.summary(), "parity_check": parity,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        from nbmf_mm_b200.device import destroy_cached_comms
        destroy_cached_comms()
        dist.destroy_process_group()
    if parity is not None and not parity["ok"]:
        raise SystemExit(f"parity_check failed: {parity}")



# --------------------------------------------------------------------------- the other BASELINE configs
def run_configs():
    """`python bench.py --configs`: BASELINE.json configs[0], [1], [2], [4] end to end through the public API (dense fp64
    host inputs, as a user of the reference passes them) next to the CPU oracle port on the same box (the cpu_baseline
    leg for those configs: bounded samples where the full CPU run would take minutes).  One text line per case; the
    JSON contract line of configs[3] is the default mode."""
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
    l_rel = max(abs(est.loss_curve_[i] - ref[2][i]) / abs(ref[2][i]) for i in range(2))
    line("cfg3 20000x5000 K=20 dir-beta duchi 90% mask float32", gs, 100, cs, 2, V.size,
         f"final loss {est.loss_curve_[-1]:.6f}; losses of iterations 0, 1: {est.loss_curve_[0]:.9f} {est.loss_curve_[1]:.9f} vs oracle "
         f"{ref[2][0]:.9f} {ref[2][1]:.9f} (max rel {l_rel:.1e}, bar 1e-4)")
    assert l_rel < 1e-4, "cfg3: the first two losses disagree with the oracle"

    # configs[4]: 64 restarts, K sweep on lastfm-shaped data
    L = (np.random.default_rng(0).random((1226, 285)) < 0.0435).astype(np.float64)
    # (the GPU legs first, back to back: after a minute of CPU-only work the GPU has dropped to its idle clocks and the
    # first calls, tens of milliseconds long, are over before it has ramped up again)
    gpu_s = {}
    t_warm = time.perf_counter()
    while time.perf_counter() - t_warm < 1.5:                  # untimed: clocks back up after the CPU leg of configs[2]
        nbmf_mm_multifit(L, [dict(n_components=32, random_state=r) for r in range(64)], max_iter=200, tol=0.0, dtype="float32")
    for k in (6, 16, 32, 64):
        jobs = [dict(n_components=k, random_state=r) for r in range(64)]
        gpu_s[k], res = timed(lambda: nbmf_mm_multifit(L, jobs, max_iter=200, tol=0.0, dtype="float32"), reps=4)
        gpu_s[k] = (gpu_s[k], min(r[2][-1] for r in res))
    for k in (6, 16, 32, 64):
        cs, ref = cpu_time(lambda: orc.fit(L, k, max_iter=200, tol=0.0, random_state=0))
        line(f"cfg5 64 restarts 1226x285 K={k} float32", gpu_s[k][0], 200 * 64, cs * 64, 200 * 64, L.size,
             f"(CPU: one restart timed, x 64; best final loss {gpu_s[k][1]:.9f}, restart 0 on the CPU {ref[2][-1]:.9f})")


def run_config5_restarts(a):
    """`python bench.py --configs --gpus N` (under torchrun for N > 1): BASELINE.json configs[4] -- n_init = 64 restarts on
    lastfm-shaped data, K sweep, FP32 -- with the restarts partitioned across the N GPUs (`distributed="restarts"`: restart r
    on rank r % N, no data-path collective, the best restart's factors broadcast at the end).  One JSON line per K on rank
    0: 64 restarts x 200 iterations, wall time of the whole `NBMF(n_init=64).fit` call, max over ranks."""
    import torch
    import torch.distributed as dist
    from nbmf_mm_b200 import NBMF
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    L = (np.random.default_rng(0).random((1226, 285)) < 0.0435).astype(np.float64)
    n_init, iters = 64, 200
    for k in (6, 8, 12, 16, 24, 32, 48, 64):
        best = None
        for rep in range(3):                                   # first repetition warms up (module load, pinned pools, graphs)
            est = NBMF(n_components=k, n_init=n_init, random_state=0, max_iter=iters, tol=0.0, dtype="float32",
                       distributed="restarts" if world > 1 else False, device=dev)
            torch.cuda.synchronize(dev)
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            est.fit(L)
            torch.cuda.synchronize(dev)
            dt = time.perf_counter() - t0
            if world > 1:
                t = torch.tensor([dt], dtype=torch.float64, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dt = float(t.item())
            if rep > 0:
                best = dt if best is None else min(best, dt)
        if rank == 0:
            print(json.dumps({"config": "configs[4]: 64 restarts on 1226x285 (4.35% ones), 200 iterations each, float32",
                              "k": k, "n_gpus": world, "seconds": best, "unit": UNIT,
                              "value": L.size * iters * n_init / best, "best_restart": int(est.best_init_),
                              "final_loss": float(est.loss_curve_[-1]), "engine": est.transfer_stats_.get("engine"),
                              "restarts_per_gpu": -(-n_init // world)}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


class C_double:
    def __init__(self):
        import ctypes
        self._c = ctypes.c_double(0.0)
        self._ctypes = ctypes

    def ref(self):
        return self._ctypes.byref(self._c)

    @property
    def value(self):
        return float(self._c.value)


def measured_peaks():
    try:
        return json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
    except Exception:
        return {}


if __name__ == "__main__":
    args = parse()
    if args.configs and (args.gpus > 1 or int(os.environ.get("WORLD_SIZE", "1")) > 1 or args.restarts_only):
        run_config5_restarts(args)
    elif args.configs:
        run_configs()
    elif args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)
