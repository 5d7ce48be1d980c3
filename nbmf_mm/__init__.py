"""Import shim: ``from nbmf_mm import NBMF`` resolves to the B200 implementation, so code
written against siddC/nbmf_mm runs unchanged (public names of the reference's
``src/nbmf_mm/__init__.py:10-18``)."""
from nbmf_mm_b200 import NBMF, NBMFMM, nbmf_mm_solver, __version__  # noqa: F401

__all__ = ["NBMFMM", "NBMF", "nbmf_mm_solver"]
