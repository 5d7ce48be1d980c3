"""Shim for ``nbmf_mm._solver``."""
from nbmf_mm_b200.solver import nbmf_mm_solver, nbmf_mm_update_beta_dir  # noqa: F401
