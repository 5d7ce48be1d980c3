"""nbmf_mm_b200 -- B200-native NBMF-MM (mean-parameterised Bernoulli NMF, Magron & Fevotte 2022).

Drop-in for the fit path of ``siddC/nbmf_mm``: same public names (``NBMFMM``, ``NBMF``,
``nbmf_mm_solver``), hand-written sm_100a CUDA kernels behind a C-ABI, no CPU fallback.
"""
import os as _os

# Many small fits run concurrently on their own streams (n_init, grid sweeps: nbmf_mm_multifit).  The driver maps
# streams onto CUDA_DEVICE_MAX_CONNECTIONS hardware queues (default 8); beyond that, independent streams serialise
# behind each other.  Measured on config 5 (64 restarts, 1226 x 285): 128 -> 76 ms at K=6, 263 -> 193 ms at K=32 with
# 32 queues.  Only a default: an explicit setting wins, and it has no effect once the CUDA context exists.
_os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

from .bits import BitMatrix  # noqa: E402
from .estimator import NBMF, NBMFMM  # noqa: E402
from .multifit import nbmf_mm_multifit  # noqa: E402
from .solver import nbmf_mm_solver, nbmf_mm_update_beta_dir  # noqa: E402
from . import datasets, experiment  # noqa: E402,F401

__version__ = "0.1.0"
__all__ = ["NBMFMM", "NBMF", "nbmf_mm_solver", "nbmf_mm_update_beta_dir", "nbmf_mm_multifit", "BitMatrix"]
