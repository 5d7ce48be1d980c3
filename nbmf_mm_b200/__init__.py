"""nbmf_mm_b200 -- B200-native NBMF-MM (mean-parameterised Bernoulli NMF, Magron & Fevotte 2022).

Drop-in for the fit path of ``siddC/nbmf_mm``: same public names (``NBMFMM``, ``NBMF``,
``nbmf_mm_solver``), hand-written sm_100a CUDA kernels behind a C-ABI, no CPU fallback.
"""
from .bits import BitMatrix
from .estimator import NBMF, NBMFMM
from .multifit import nbmf_mm_multifit
from .solver import nbmf_mm_solver, nbmf_mm_update_beta_dir

__version__ = "0.1.0"
__all__ = ["NBMFMM", "NBMF", "nbmf_mm_solver", "nbmf_mm_update_beta_dir", "nbmf_mm_multifit", "BitMatrix"]
