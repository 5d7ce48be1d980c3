#!/usr/bin/env python
"""Per-phase wall time of the row-sharded end-to-end path (development tool; run under torchrun on a GPU box):
    python -m torch.distributed.run --nproc-per-node N tools/e2e_breakdown_dist.py [rows cols k steps]
Every phase is bracketed by a device synchronisation, so the phases do not overlap here as they do in the real call
(the total of the un-instrumented call is printed next to the sum)."""
import os, sys, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
import torch.distributed as dist
import nbmf_mm_b200.solver as S, nbmf_mm_b200.device as D
from nbmf_mm_b200 import BitMatrix, nbmf_mm_solver
from nbmf_mm_b200.device import synth_bits_device

m, n, k, steps = (int(x) for x in (sys.argv[1:5] or (1000000, 100000, 32, 20)))
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
r0, r1 = S._row_shard(m, rank, world)
hstar = (np.random.default_rng(4).random((k, n)) * 0.2).astype(np.float32)
P, M = synth_bits_device(4, r0, r1 - r0, n, hstar, 0.9, dev)
Ph = torch.empty(P.words.shape, dtype=torch.int32, pin_memory=True).copy_(P.words)
Mh = torch.empty(M.words.shape, dtype=torch.int32, pin_memory=True).copy_(M.words)
del P, M
torch.cuda.empty_cache()
marks, on = [], [False]

def wrap(owner, name):
    f = getattr(owner, name)
    def g(*a, **kw):
        if not on[0]:
            return f(*a, **kw)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        r = f(*a, **kw)
        torch.cuda.synchronize(); marks.append((name, time.perf_counter() - t0))
        return r
    setattr(owner, name, g)

for nm in ("__init__", "stream_bits_from_host", "init_comm", "set_factors", "finish_bits", "fit", "simplex_deviation",
           "get_factors_f64", "release_planes", "close"):
    wrap(D.DeviceProblem, nm)
for nm in ("draw_shard_inits", "_gather_h_parts", "pinned_factor_buffers"):
    wrap(S, nm)

def call():
    return nbmf_mm_solver(BitMatrix(Ph, (r1 - r0, n)), k, max_iter=steps, tol=0.0, alpha=1.2, beta=1.2, mask=BitMatrix(Mh, (r1 - r0, n)),
                          random_state=0, dtype="float32", device=dev, distributed=True, shard=(r0, m))

def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()

call()                                            # first call: pools, communicator, jump polynomials
for instrumented in (False, True, False):
    on[0] = instrumented
    marks.clear()
    barrier(); t0 = time.perf_counter()
    call()
    barrier(); dt = time.perf_counter() - t0
    if rank == 0:
        if instrumented:
            print(f"instrumented: total {dt:.3f} s; " + ", ".join(f"{a} {b * 1e3:.1f} ms" for a, b in marks) +
                  f"; unaccounted {(dt - sum(b for _, b in marks)) * 1e3:.1f} ms", flush=True)
        else:
            print(f"plain call: {dt:.3f} s for {steps} iterations at N={world} ({m * n * steps / dt:.3e} updates/s)", flush=True)
if world > 1:
    D.destroy_cached_comms()
    dist.destroy_process_group()
