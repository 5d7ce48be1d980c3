"""Multi-GPU path (needs >= 2 GPUs; skipped otherwise): rows sharded over ranks, one ncclAllReduce
of the K x N H-partials per iteration.  1-GPU vs 2-GPU results may differ
only by summation order; run-to-run at fixed world size must be bitwise identical."""
import os
import socket

import numpy as np
import pytest

pytestmark = [pytest.mark.gpu, pytest.mark.multigpu]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _problem():
    rng = np.random.default_rng(21)
    X = (rng.random((700, 1500)) < 0.2).astype(np.float64)
    mask = (rng.random(X.shape) < 0.9).astype(np.float64)
    return X, mask


def _worker(rank, world, port, dtype, out):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from nbmf_mm_b200 import BitMatrix, nbmf_mm_solver
        from nbmf_mm_b200.solver import _row_shard
        X, mask = _problem()
        runs = []
        for _ in range(2):
            W, H, losses, _, n_iter = nbmf_mm_solver(X, 12, max_iter=60, tol=1e-6, mask=mask, random_state=5,
                                                     dtype=dtype, distributed=True)
            runs.append((W, H, np.asarray(losses), n_iter))
        assert np.array_equal(runs[0][0], runs[1][0]) and np.array_equal(runs[0][2], runs[1][2])   # deterministic
        # shard-local inputs: each rank only ever sees its own rows
        r0, r1 = _row_shard(X.shape[0], rank, world)
        Ws, Hs, ls, _, ns = nbmf_mm_solver(BitMatrix.from_dense(X[r0:r1]), 12, max_iter=60, tol=1e-6,
                                           mask=BitMatrix.from_dense(mask[r0:r1]), random_state=5, dtype=dtype,
                                           distributed=True, shard=(r0, X.shape[0]))
        assert np.array_equal(Ws, runs[0][0][r0:r1]) and np.array_equal(Hs, runs[0][1]) and ns == runs[0][3]
        if rank == 0:
            np.savez(out, W=runs[0][0], H=runs[0][1], losses=runs[0][2], n_iter=runs[0][3])
    finally:
        from nbmf_mm_b200.device import destroy_cached_comms
        destroy_cached_comms()
        dist.destroy_process_group()


@pytest.mark.parametrize("dtype,tol", [("float64", 1e-11), ("float32", 2e-5)])
def test_two_gpus_match_one(tmp_path, dtype, tol):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from nbmf_mm_b200 import nbmf_mm_solver
    out = str(tmp_path / "res.npz")
    mp.spawn(_worker, args=(2, _free_port(), dtype, out), nprocs=2, join=True)
    res = np.load(out)
    X, mask = _problem()
    W, H, losses, _, n_iter = nbmf_mm_solver(X, 12, max_iter=60, tol=1e-6, mask=mask, random_state=5, dtype=dtype)
    assert int(res["n_iter"]) == n_iter
    assert np.max(np.abs(res["losses"] - np.asarray(losses)) / np.abs(losses)) < tol
    assert np.max(np.abs(res["W"] - W)) < tol * 10 and np.max(np.abs(res["H"] - H)) < tol * 10


def _restart_worker(rank, world, port, out):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from nbmf_mm_b200 import NBMF
        X, mask = _problem()
        est = NBMF(n_components=6, max_iter=40, tol=0.0, random_state=3, n_init=5, dtype="float64",
                   distributed="restarts").fit(X, mask=mask)
        if rank == 1:                                            # the non-owner ranks hold the winner too
            np.savez(out, W=est.W_, H=est.components_, losses=np.asarray(est.loss_curve_), best=est.best_init_)
    finally:
        dist.destroy_process_group()


def test_restarts_partitioned_over_two_gpus_match_sequential(tmp_path):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from nbmf_mm_b200 import NBMF
    out = str(tmp_path / "restarts.npz")
    mp.spawn(_restart_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = np.load(out)
    X, mask = _problem()
    seq = NBMF(n_components=6, max_iter=40, tol=0.0, random_state=3, n_init=5, dtype="float64").fit(X, mask=mask)
    assert int(got["best"]) == seq.best_init_
    assert np.array_equal(got["W"], seq.W_) and np.array_equal(got["H"], seq.components_)
    assert np.array_equal(got["losses"], np.asarray(seq.loss_curve_))


def _unseeded_worker(rank, world, port, out):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from nbmf_mm_b200 import nbmf_mm_solver
        X, mask = _problem()
        np.random.seed(100 + rank)                               # every rank's own global stream is in a different state
        W, H, losses, _, n_iter = nbmf_mm_solver(X, 9, max_iter=50, tol=1e-5, mask=mask, random_state=None,
                                                 dtype="float32", distributed=True)
        # dir-beta with dense inputs: each rank uploads only its block of COLUMNS of X (internal rows)
        Wd, Hd, ld, _, nd = nbmf_mm_solver(X, 7, max_iter=30, tol=0.0, mask=mask, random_state=4, dtype="float64",
                                           distributed=True, orientation="dir-beta")
        np.savez(f"{out}.{rank}.npz", W=W, H=H, losses=np.asarray(losses), n_iter=n_iter, Wd=Wd, Hd=Hd, ld=np.asarray(ld))
    finally:
        from nbmf_mm_b200.device import destroy_cached_comms
        destroy_cached_comms()
        dist.destroy_process_group()


def test_unseeded_row_shards_agree_and_dir_beta_slices_columns(tmp_path):
    """random_state=None: every rank is its own process with its own global NumPy stream; rank 0's seed is broadcast so
    that all ranks start from the same H (they would otherwise diverge, and ranks that stop at different iterations
    enqueue different numbers of collectives).  And dir-beta on dense inputs: a rank's row block of the internal problem
    is a block of columns of X."""
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from nbmf_mm_b200 import nbmf_mm_solver
    out = str(tmp_path / "unseeded")
    mp.spawn(_unseeded_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    a, b = np.load(f"{out}.0.npz"), np.load(f"{out}.1.npz")
    assert int(a["n_iter"]) == int(b["n_iter"]) and np.array_equal(a["losses"], b["losses"])
    assert np.array_equal(a["H"], b["H"]) and np.array_equal(a["W"], b["W"])
    assert np.all(np.diff(a["losses"]) <= 2e-6 * np.abs(a["losses"][:-1]))
    X, mask = _problem()
    W1, H1, l1, _, _ = nbmf_mm_solver(X, 7, max_iter=30, tol=0.0, mask=mask, random_state=4, dtype="float64", orientation="dir-beta")
    assert np.array_equal(a["Wd"], b["Wd"]) and np.max(np.abs(a["Wd"] - W1)) < 1e-10 and np.max(np.abs(a["Hd"] - H1)) < 1e-10
    assert np.max(np.abs(a["ld"] - np.asarray(l1)) / np.abs(l1)) < 1e-11
