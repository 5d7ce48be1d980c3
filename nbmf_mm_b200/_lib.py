"""ctypes binding of ``libnbmf_b200.so`` (C-ABI declared in ``include/nbmf_b200.h``).

There is no CPU fallback: if the shared library is missing, or a compute entry point is
called without a CUDA device, this module raises.  Build with
``python -c "import __graft_entry__ as g; g.build()"`` or ``make -C nbmf_mm_b200/csrc``.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "libnbmf_b200.so"

NBMF_F32, NBMF_F64, NBMF_U8, NBMF_F16 = 0, 1, 2, 3
NBMF_V_BITS, NBMF_V_DENSE, NBMF_V_DENSE_F16 = 0, 1, 2
NBMF_MASK_REFERENCE, NBMF_MASK_STRICT = 0, 1
NBMF_PROJ_NORMALIZE, NBMF_PROJ_DUCHI = 0, 1
NBMF_ENGINE_AUTO, NBMF_ENGINE_SIMT, NBMF_ENGINE_TENSOR, NBMF_ENGINE_FUSED = 0, 1, 2, 3
ENGINES = {"auto": NBMF_ENGINE_AUTO, "simt": NBMF_ENGINE_SIMT, "tensor": NBMF_ENGINE_TENSOR, "fused": NBMF_ENGINE_FUSED}


class NbmfConfig(C.Structure):
    _fields_ = [
        ("m", C.c_int64), ("n", C.c_int64), ("k", C.c_int32), ("dtype", C.c_int32),
        ("vkind", C.c_int32), ("mask_semantics", C.c_int32), ("projection", C.c_int32),
        ("has_mask", C.c_int32), ("alpha", C.c_double), ("beta", C.c_double), ("eps", C.c_double),
        ("n_obs", C.c_double), ("max_iter_cap", C.c_int32), ("engine", C.c_int32),
        ("batch_hint", C.c_int32),
    ]


# name -> (restype, argtypes); every symbol include/nbmf_b200.h declares is listed here and
# tests/test_capi_symbols.py checks the two stay in sync.
_P, _I64, _I32, _INT, _DBL = C.c_void_p, C.c_int64, C.c_int32, C.c_int, C.c_double
SIGNATURES = {
    "nbmf_version": (_INT, []),
    "nbmf_last_error": (C.c_char_p, []),
    "nbmf_words_per_row": (_I64, [_I64]),
    "nbmf_padded_cols": (_I64, [_I64]),
    "nbmf_pack_bits": (_INT, [_P, _INT, _I64, _P, _INT, _I64, _I64, _I64, _P, _P, _P]),
    "nbmf_pack_bits_checked": (_INT, [_P, _INT, _I64, _P, _INT, _I64, _I64, _I64, _P, _P, _P, _P]),
    "nbmf_pack_dense": (_INT, [_P, _INT, _I64, _P, _INT, _I64, _I64, _I64, _INT, _P, _P]),
    "nbmf_transpose_bits": (_INT, [_P, _I64, _I64, _P, _P]),
    "nbmf_pack_csr": (_INT, [_P, _P, _P, _INT, _I64, _I64, _P, _P, C.POINTER(_I32), _P]),
    "nbmf_reconstruct": (_INT, [_INT, _P, _P, _I64, _I64, _I32, _P, _P]),
    "nbmf_popcount_bits": (_INT, [_P, _I64, _I64, _P, C.POINTER(C.c_uint64), _P]),
    "nbmf_synth_bits": (_INT, [C.c_uint64, _I64, _I64, _I64, _P, _I32, C.c_float, _P, _P, _P]),
    "nbmf_workspace_bytes": (_I64, [C.POINTER(NbmfConfig)]),
    "nbmf_create": (_INT, [C.POINTER(NbmfConfig), _P, _I64, _P, C.POINTER(_P)]),
    "nbmf_destroy": (_INT, [_P]),
    "nbmf_set_data_bits": (_INT, [_P, _P, _P]),
    "nbmf_planes_in_use": (_INT, [_P]),
    "nbmf_ingest_bits_begin": (_INT, [_P, _P, _P]),
    "nbmf_ingest_bits_rows": (_INT, [_P, _I64, _I64]),
    "nbmf_ingest_bits_end": (_INT, [_P, C.POINTER(_DBL)]),
    "nbmf_set_n_obs": (_INT, [_P, _DBL]),
    "nbmf_set_data_dense": (_INT, [_P, _P, _P]),
    "nbmf_set_mask_weights": (_INT, [_P, _P]),
    "nbmf_set_factors": (_INT, [_P, _P, _P, _INT]),
    "nbmf_get_factors": (_INT, [_P, _P, _P]),
    "nbmf_simplex_deviation": (_INT, [_P, C.POINTER(_DBL)]),
    "nbmf_get_factors_f64": (_INT, [_P, _P, _P, _INT]),
    "nbmf_h_half_step": (_INT, [_P]),
    "nbmf_w_half_step": (_INT, [_P]),
    "nbmf_objective": (_INT, [_P, C.POINTER(_DBL)]),
    "nbmf_loglik_partials": (_INT, [_P, C.POINTER(_DBL), _I64, C.POINTER(_I32), C.POINTER(_I32)]),
    "nbmf_fit": (_INT, [_P, _I32, _DBL, C.POINTER(_DBL), C.POINTER(_I32), C.POINTER(_I32)]),
    "nbmf_fit_begin": (_INT, [_P, _I32, _DBL]),
    "nbmf_batch_bind": (_INT, [_P, _I32, _I64]),
    "nbmf_batch_poll": (_INT, [_P, C.POINTER(_I32), C.POINTER(_I32)]),
    "nbmf_batch_tail": (_INT, [_P, C.POINTER(_DBL), _I32, C.POINTER(_I32), C.POINTER(_DBL)]),
    "nbmf_fit_enqueue": (_INT, [_P, _I32]),
    "nbmf_fit_poll": (_INT, [_P, _INT, C.POINTER(_I32), C.POINTER(_I32)]),
    "nbmf_fit_history": (_INT, [_P, C.POINTER(_DBL), _I32, C.POINTER(_I32)]),
    "nbmf_transform": (_INT, [_P, _I32]),
    "nbmf_comm_unique_id": (_INT, [_P]),
    "nbmf_comm_init": (_INT, [_P, _P, _I32, _I32]),
    "nbmf_comm_world": (_INT, [_P]),
    "nbmf_comm_create": (_INT, [_P, _I32, _I32, C.POINTER(_P)]),
    "nbmf_comm_attach": (_INT, [_P, _P, _I32, _I32]),
    "nbmf_comm_destroy": (_INT, [_P]),
    "nbmf_engine": (_INT, [_P]),
    "nbmf_fit_is_fused": (_INT, [_P]),
    "nbmf_mt19937_uniform": (_INT, [C.c_uint32, C.c_uint64, C.c_uint64, _DBL, _DBL, _P, _P]),
    "nbmf_fma_peak": (_INT, [_INT, _I32, _P, _P, C.POINTER(_DBL)]),
    "nbmf_profile_enable": (_INT, [_P, _INT]),
    "nbmf_profile_read": (_INT, [_P, C.POINTER(_DBL), C.POINTER(_I32), C.POINTER(_DBL), C.POINTER(_I32)]),
    "nbmf_plan_info": (_INT, [_P, C.POINTER(_I32), C.POINTER(_I32), C.POINTER(_I32), C.POINTER(_I32)]),
    "nbmf_launch_count": (_I64, [_INT]),
    "nbmf_variant_info": (_INT, [_INT, _INT, _INT, C.POINTER(_I32), C.POINTER(_I32), C.POINTER(_I32)]),
}

_lib = None


class NbmfError(RuntimeError):
    """An entry point of libnbmf_b200.so returned a non-zero status."""


def load():
    """Load the shared library (once) and attach the prototypes.  Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    path = Path(os.environ.get("NBMF_B200_LIB", LIB_PATH))
    if not path.exists():
        raise ImportError(
            f"{path} not found: the CUDA extension is not built and nbmf_mm_b200 has no CPU fallback. "
            "Run `make -C nbmf_mm_b200/csrc` (needs nvcc, cross-compiles sm_100a without a GPU).")
    lib = C.CDLL(str(path), mode=C.RTLD_GLOBAL)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        msg = load().nbmf_last_error().decode("utf-8", "replace")
        raise NbmfError(f"{what or 'libnbmf_b200'} failed ({rc}): {msg}")


def nccl_library_hint():
    """Point the dlopen in capi.cu at the NCCL that ships with torch (same one torch.distributed uses)."""
    if "NBMF_NCCL_LIB" in os.environ:
        return os.environ["NBMF_NCCL_LIB"]
    try:
        import nvidia.nccl  # type: ignore
        for base in list(getattr(nvidia.nccl, "__path__", [])):
            cand = Path(base) / "lib" / "libnccl.so.2"
            if cand.exists():
                os.environ["NBMF_NCCL_LIB"] = str(cand)
                return str(cand)
    except Exception:
        pass
    return None
