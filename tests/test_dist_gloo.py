"""Host-side logic of the multi-GPU layer on CPU: world_size-2 `gloo` processes.

The CUDA kernels cannot run here, so the per-shard arithmetic is done by the oracle (test
infrastructure); what is under test is the decomposition the multi-GPU path relies on
(SURVEY.md section 8e): contiguous row shards from `_row_shard`, W half-step row-local, H
half-step = allreduce(sum) of the per-shard K x N partials [C | D | LL], identical epilogue
on every rank; plus the 128-byte unique-id hand-off that `DeviceProblem.init_comm` does."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import nbmf_oracle as orc
from nbmf_mm_b200.solver import _row_shard


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _partials(Y, W, H, mask, eps=1e-8):
    """Per-shard partial numerators/denominators and log-likelihood of the H half-step
    (_solver.py:39-43,150-154), reference mask semantics."""
    pos = Y * mask
    theta = W.T @ H
    C = W @ (pos / (theta + eps))
    D = W @ ((1 - pos) / (1 - theta + eps))
    LL = np.sum(pos * np.log(theta + eps) + (1 - pos) * np.log(1 - theta + eps))
    return C, D, LL


def _worker(rank, world, port, m, n, k, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(0)                         # every rank builds the same global problem
        Y = (rng.random((m, n)) < 0.3).astype(np.float64)
        mask = (rng.random((m, n)) < 0.85).astype(np.float64)
        W = rng.uniform(0.1, 0.9, (k, m)); W /= W.sum(axis=0, keepdims=True)
        H = rng.uniform(0.05, 0.95, (k, n))
        r0, r1 = _row_shard(m, rank, world)
        # 1. the id hand-off used for ncclCommInitRank
        payload = [bytes(range(128)) if rank == 0 else None]
        dist.broadcast_object_list(payload, src=0)
        assert payload[0] == bytes(range(128))
        # 2. H half-step: local partials -> allreduce -> identical epilogue everywhere
        C, D, LL = _partials(Y[r0:r1], W[:, r0:r1], H, mask[r0:r1])
        buf = torch.from_numpy(np.concatenate([C.ravel(), D.ravel(), [LL]]))
        dist.all_reduce(buf)
        C = buf[: k * n].numpy().reshape(k, n); D = buf[k * n: 2 * k * n].numpy().reshape(k, n); LL = float(buf[-1])
        num = H * C + 0.2
        den = (1 - H) * D + 0.3
        H1 = np.clip(num / (num + den + 1e-8), 1e-8, 1 - 1e-8)
        # 3. W half-step is row-local: no communication
        W1_local = orc.w_half_step(Y[r0:r1], W[:, r0:r1], H1, mask[r0:r1])
        gathered = [None] * world
        dist.all_gather_object(gathered, (r0, W1_local))
        if rank == 0:
            W1 = np.zeros((k, m))
            for g0, blk in gathered:
                W1[:, g0:g0 + blk.shape[1]] = blk
            np.savez(out, H1=H1, W1=W1, LL=LL)
        n_obs = torch.tensor([float(mask[r0:r1].sum())], dtype=torch.float64)
        dist.all_reduce(n_obs)
        assert n_obs.item() == mask.sum()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("m,n,k", [(97, 60, 5), (64, 40, 3)])
def test_row_sharded_step_equals_unsharded(tmp_path, m, n, k):
    out = str(tmp_path / "res.npz")
    mp.spawn(_worker, args=(2, _free_port(), m, n, k, out), nprocs=2, join=True)
    res = np.load(out)
    rng = np.random.default_rng(0)
    Y = (rng.random((m, n)) < 0.3).astype(np.float64)
    mask = (rng.random((m, n)) < 0.85).astype(np.float64)
    W = rng.uniform(0.1, 0.9, (k, m)); W /= W.sum(axis=0, keepdims=True)
    H = rng.uniform(0.05, 0.95, (k, n))
    W1, H1 = orc.mm_step(Y, W, H, mask, 1.2, 1.3)
    assert np.max(np.abs(res["H1"] - H1)) < 1e-13             # differs only by the summation order of the allreduce
    assert np.max(np.abs(res["W1"] - W1)) < 1e-13
    _, _, LL = _partials(Y, W, H, mask)
    assert abs(res["LL"] - LL) < 1e-10 * abs(LL)


def _restart_worker(rank, world, port, n_init, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from nbmf_mm_b200.estimator import partition_restarts, reduce_best_restart
        losses = np.random.default_rng(7).permutation(n_init).astype(np.float64) // 2     # ties on purpose
        mine = partition_restarts(n_init, rank, world)
        best = None
        for r in mine:                                          # stand-in for the per-restart fit
            if best is None or losses[r] < best[0]:
                best = (losses[r], r, {"restart": r, "W": np.full((2, 2), float(r))})
        payload, idx = reduce_best_restart(best)
        assert payload["restart"] == idx and np.all(payload["W"] == idx)
        if rank == 0:
            np.savez(out, idx=idx, mine=np.asarray(mine))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_init", [1, 5, 8])
def test_restart_partition_keeps_what_a_sequential_loop_keeps(tmp_path, n_init):
    """n_init restarts dealt round-robin to the ranks, no data-path collective; the winner is the restart a
    sequential loop with `<` keeps (lowest loss, first index among ties) and every rank receives it."""
    from nbmf_mm_b200.estimator import partition_restarts
    assert sorted(partition_restarts(n_init, 0, 2) + partition_restarts(n_init, 1, 2)) == list(range(n_init))
    out = str(tmp_path / "best.npz")
    mp.spawn(_restart_worker, args=(2, _free_port(), n_init, out), nprocs=2, join=True)
    losses = np.random.default_rng(7).permutation(n_init).astype(np.float64) // 2
    want = int(np.argmin(losses))                                # first index among ties
    assert int(np.load(out)["idx"]) == want


def _init_worker(rank, world, port, m, n, k, seed, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from nbmf_mm_b200.solver import _gather_h_parts, draw_shard_inits
        r0, r1 = _row_shard(m, rank, world)
        W_local, H_part, _ = draw_shard_inits(seed, m, n, k, r0, r1, h_part=(rank, world), set_global_state=True)
        H = _gather_h_parts(H_part, k, n, world, "float64", torch.device("cpu")).numpy()
        after = np.random.uniform(0.1, 0.9, 4)                  # the global stream sits where the reference leaves it
        np.savez(f"{out}.{rank}.npz", W=W_local, H=H, r0=r0, after=after)
    finally:
        dist.destroy_process_group()


def test_sharded_inits_equal_the_reference_stream(tmp_path):
    """Row shards enter the reference's init stream (seed; W_init m x k; H_init k x n: _solver.py:102-103,126-129) by
    MT19937 jump-ahead: each rank draws only its W rows and its 1/world of H_init, the H parts are all-gathered."""
    m, n, k, seed = 5000, 333, 9, 5
    out = str(tmp_path / "init")
    mp.spawn(_init_worker, args=(2, _free_port(), m, n, k, seed, out), nprocs=2, join=True)
    rs = np.random.RandomState(seed)
    W = rs.uniform(0.1, 0.9, (m, k)); H = rs.uniform(0.1, 0.9, (k, n)); after = rs.uniform(0.1, 0.9, 4)
    for rank in range(2):
        z = np.load(f"{out}.{rank}.npz")
        r0 = int(z["r0"])
        assert np.array_equal(z["W"], W[r0:r0 + z["W"].shape[0]]) and np.array_equal(z["H"], H)
        assert np.array_equal(z["after"], after)
