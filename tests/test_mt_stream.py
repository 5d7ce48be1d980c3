"""``nbmf_mt19937_uniform`` (host C++ in libnbmf_b200.so, no GPU): a window of the reference's init stream --
``np.random.seed(s); uniform(0.1, 0.9, (m, k)); uniform(0.1, 0.9, (k, n))``, ``_solver.py:102-103,126-129`` -- reached
by MT19937 polynomial jump-ahead instead of by drawing everything before it.  Must be bit-identical to NumPy."""
import numpy as np
import pytest

from nbmf_mm_b200 import _lib
from nbmf_mm_b200.solver import draw_shard_inits


def _draw(seed, skip, count, lo=0.1, hi=0.9):
    lib = _lib.load()
    out = np.empty(count)
    st = np.zeros(625, dtype=np.uint32)
    assert lib.nbmf_mt19937_uniform(seed, skip, count, lo, hi, out.ctypes.data, st.ctypes.data) == 0
    return out, st


@pytest.mark.parametrize("seed", [0, 1, 12345, 2**32 - 1])
def test_windows_of_the_stream_equal_numpy(seed):
    ref = np.random.RandomState(seed).uniform(0.1, 0.9, 200_000)
    for skip, count in ((0, 700), (5, 100), (311, 2), (312, 5), (2047, 10), (2048, 10), (5000, 1000), (123_457, 2000), (199_000, 1000)):
        got, _ = _draw(seed, skip, count)
        assert np.array_equal(got, ref[skip:skip + count]), (seed, skip, count)


def test_other_ranges_and_the_state_afterwards():
    rs = np.random.RandomState(3)
    a = rs.uniform(-2.0, 5.0, 70_001)
    nxt = rs.uniform(0.1, 0.9, 7)
    got, st = _draw(3, 60_000, 10_001, -2.0, 5.0)
    assert np.array_equal(got, a[60_000:])
    saved = np.random.get_state()
    try:
        np.random.set_state(("MT19937", st[:624], int(st[624])))
        assert np.array_equal(np.random.uniform(0.1, 0.9, 7), nxt)       # the global stream continues where NumPy's would
    finally:
        np.random.set_state(saved)


def test_shard_inits_equal_the_reference_draw_order():
    m, n, k, seed = 4000, 300, 7, 11
    rs = np.random.RandomState(seed)
    W = rs.uniform(0.1, 0.9, (m, k))
    H = rs.uniform(0.1, 0.9, (k, n))
    after = rs.uniform(0.1, 0.9, 3)
    saved = np.random.get_state()
    try:
        for r0, r1 in ((0, 1024), (1024, 2048), (3072, 4000)):
            Wl, Hc, (c0, c1) = draw_shard_inits(seed, m, n, k, r0, r1, h_part=(1, 3), set_global_state=True)
            assert np.array_equal(Wl, W[r0:r1])
            assert np.array_equal(Hc, H.ravel()[c0:c1]) and (c0, c1) == ((k * n + 2) // 3, 2 * ((k * n + 2) // 3))
            assert np.array_equal(np.random.uniform(0.1, 0.9, 3), after)
    finally:
        np.random.set_state(saved)
