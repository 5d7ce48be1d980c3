"""The C-ABI library loads on a CPU-only box and exports exactly what include/nbmf_b200.h
declares; the ctypes prototype table covers every declared symbol.  No compute calls here."""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
HEADER = ROOT / "include" / "nbmf_b200.h"


def declared_symbols():
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    return sorted(set(re.findall(r"\b(nbmf_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_boundary():
    names = declared_symbols()
    for must in ("nbmf_create", "nbmf_fit", "nbmf_h_half_step", "nbmf_w_half_step", "nbmf_objective",
                 "nbmf_transform", "nbmf_pack_bits", "nbmf_transpose_bits", "nbmf_comm_init"):
        assert must in names


def test_library_exports_every_declared_symbol():
    from nbmf_mm_b200 import _lib
    lib = _lib.load()
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} is declared in the header but not exported"


def test_ctypes_table_matches_header():
    from nbmf_mm_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared_symbols()


def test_pure_host_entry_points():
    from nbmf_mm_b200 import _lib
    lib = _lib.load()
    assert lib.nbmf_version() >= 100
    assert lib.nbmf_words_per_row(1) == 32 and lib.nbmf_words_per_row(1024) == 32
    assert lib.nbmf_words_per_row(1025) == 64 and lib.nbmf_words_per_row(100000) == 3136
    assert lib.nbmf_padded_cols(500) == 1024
    h, w, kp = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32()
    for k, want in [(1, 4), (4, 4), (6, 8), (10, 12), (16, 16), (20, 20), (32, 32), (33, 48), (64, 64), (65, 96), (97, 128), (128, 128)]:
        assert lib.nbmf_variant_info(0, 0, k, ctypes.byref(h), ctypes.byref(w), ctypes.byref(kp)) == 0
        assert kp.value == want and 1024 % h.value == 0 and w.value % 32 == 0
    assert lib.nbmf_variant_info(1, 0, 128, ctypes.byref(h), ctypes.byref(w), ctypes.byref(kp)) == 0 and h.value == 64    # fp64: 64-column tiles
    assert lib.nbmf_variant_info(0, 0, 129, None, None, None) != 0
    assert b"variant" in lib.nbmf_last_error()


def test_workspace_planning_is_deterministic_and_sane():
    from nbmf_mm_b200 import _lib
    lib = _lib.load()
    cfg = _lib.NbmfConfig()
    cfg.m, cfg.n, cfg.k, cfg.dtype, cfg.vkind = 1_000_000, 100_000, 32, 0, 0
    cfg.has_mask, cfg.alpha, cfg.beta, cfg.eps, cfg.n_obs, cfg.max_iter_cap = 1, 1.2, 1.2, 1e-8, 9e10, 50
    cfg.engine = _lib.NBMF_ENGINE_SIMT
    a = lib.nbmf_workspace_bytes(ctypes.byref(cfg))
    b = lib.nbmf_workspace_bytes(ctypes.byref(cfg))
    assert a == b and 128e6 < a < 8e9          # W alone is 128 MB; partial buffers stay bounded
    cfg.engine = _lib.NBMF_ENGINE_TENSOR       # + re-tiled bit planes (12.5 GB + 25 GB) and the split/swizzled factor blocks
    t = lib.nbmf_workspace_bytes(ctypes.byref(cfg))
    assert a + 37.5e9 < t < a + 39e9
    cfg.dtype = 1                              # the tensor engine is float32 only
    assert lib.nbmf_workspace_bytes(ctypes.byref(cfg)) < 0 and b"tensor engine" in lib.nbmf_last_error()
    cfg.dtype, cfg.engine = 0, _lib.NBMF_ENGINE_AUTO
    cfg.k = 0
    assert lib.nbmf_workspace_bytes(ctypes.byref(cfg)) < 0


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import numpy as np
    from nbmf_mm_b200 import NBMF
    X = (np.random.default_rng(0).random((10, 12)) < 0.3).astype(float)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        NBMF(n_components=3, max_iter=3).fit(X)


def test_product_never_imports_the_oracle():
    for path in (ROOT / "nbmf_mm_b200").rglob("*.py"):
        text = path.read_text()
        assert "nbmf_oracle" not in text and "import oracle" not in text and "from oracle" not in text, path
