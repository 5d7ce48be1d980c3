"""Device data layer: bit packing, packed transpose, counts, dense V*mask, synthetic generator."""
import numpy as np
import pytest

from nbmf_mm_b200 import BitMatrix
from nbmf_mm_b200.bits import words_per_row
from nbmf_mm_b200.device import pack_bits_device, pack_dense_device, synth_bits_device

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape", [(1, 1), (5, 31), (33, 32), (40, 1023), (17, 1025), (300, 2100)])
@pytest.mark.parametrize("dt", ["float64", "float32", "uint8"])
def test_pack_bits_matches_host_packing(shape, dt):
    import torch
    rng = np.random.default_rng(shape[0] * 7 + shape[1])
    X = (rng.random(shape) < 0.3)
    mask = rng.random(shape) < 0.8
    Xd = torch.from_numpy(X.astype(dt)).cuda()
    Md = torch.from_numpy(mask.astype(np.uint8)).cuda()
    P, M = pack_bits_device(Xd, Md)
    assert np.array_equal(P.words.cpu().numpy().view(np.uint32), BitMatrix.from_dense(X & mask).words)
    assert np.array_equal(M.words.cpu().numpy().view(np.uint32), BitMatrix.from_dense(mask).words)
    P2, M2 = pack_bits_device(Xd, None)
    assert M2 is None and np.array_equal(P2.to_dense(bool), X)
    assert P.count() == int((X & mask).sum()) and M.count() == int(mask.sum())


@pytest.mark.parametrize("shape", [(1, 1), (31, 33), (64, 64), (100, 1500), (1500, 70), (1030, 2050)])
def test_packed_transpose(shape):
    rng = np.random.default_rng(shape[0] + shape[1])
    A = rng.random(shape) < 0.4
    B = BitMatrix.from_dense(A).to_device("cuda")
    T = B.transpose()
    assert T.shape == (shape[1], shape[0]) and T.words.shape == (shape[1], words_per_row(shape[0]))
    assert np.array_equal(T.to_dense(bool), A.T)
    assert np.array_equal(T.transpose().words.cpu().numpy(), B.words.cpu().numpy())   # involution, padding stays zero


def test_pack_dense_applies_mask_and_pads():
    import torch
    rng = np.random.default_rng(5)
    X = rng.random((37, 1100))
    mask = rng.random((37, 1100)) < 0.7
    for dt in ("float64", "float32"):
        Vm = pack_dense_device(torch.from_numpy(X).cuda(), torch.from_numpy(mask.astype(np.uint8)).cuda(), dt)
        out = Vm.cpu().numpy()
        assert out.shape == (37, 2048) and out.dtype == np.dtype(dt)
        assert np.array_equal(out[:, :1100], (X * mask).astype(dt)) and not out[:, 1100:].any()


def test_synthetic_generator_is_counter_based():
    """Any row block can be regenerated independently; statistics follow the generative recipe."""
    n, k = 3000, 8
    hstar = np.random.default_rng(0).random((k, n)).astype(np.float32) * 0.3
    P, M = synth_bits_device(7, 0, 512, n, hstar, 0.9, "cuda")
    Pb, Mb = synth_bits_device(7, 128, 64, n, hstar, 0.9, "cuda")
    assert np.array_equal(P.to_dense(bool)[128:192], Pb.to_dense(bool))
    assert np.array_equal(M.to_dense(bool)[128:192], Mb.to_dense(bool))
    P2, _ = synth_bits_device(8, 0, 512, n, hstar, 0.9, "cuda")
    assert not np.array_equal(P.to_dense(bool), P2.to_dense(bool))
    obs, pos = M.to_dense(bool), P.to_dense(bool)
    assert abs(obs.mean() - 0.9) < 0.01 and not (pos & ~obs).any()
    # column means of V follow mean_k H*[k, j] (Dirichlet(1) rows have mean 1/k per component)
    want = hstar.mean(axis=0)
    got = pos.sum(axis=0) / np.maximum(obs.sum(axis=0), 1)
    assert abs(got.mean() - want.mean()) < 0.01 and np.corrcoef(got, want)[0, 1] > 0.5
    Pn, Mn = synth_bits_device(7, 0, 16, n, hstar, 1.0, "cuda", with_mask=False)
    assert Mn is None
