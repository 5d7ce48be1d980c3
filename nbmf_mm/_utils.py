"""Shim for ``nbmf_mm._utils`` (imported by the reference's tests)."""
from nbmf_mm_b200._utils import check_is_fitted, generate_synthetic_binary_data  # noqa: F401
