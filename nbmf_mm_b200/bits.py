"""Bit-packed binary matrices: the device-resident data layer's host-visible type.

Binary V and the observation mask are stored at 1 bit per entry: ``uint32`` words,
bit ``j % 32`` of word ``j // 32`` of row ``i``; rows are padded to
``words_per_row(n)`` words (a multiple of 32 words = 1024 columns) with zero bits.
This replaces the dense fp64 ``Y``, ``Y*mask``, ``Y.T*mask.T`` ... arrays the reference
rebuilds every iteration (``_solver.py:21-32``).
"""
from __future__ import annotations

import numpy as np


def words_per_row(n: int) -> int:
    return (int(n) + 1023) // 1024 * 32


class BitMatrix:
    """An ``m x n`` binary matrix, bit-packed.  ``words`` is either a NumPy ``uint32`` array
    (host) or a ``torch.int32`` CUDA tensor (device) of shape ``(m, words_per_row(n))``."""

    def __init__(self, words, shape):
        m, n = int(shape[0]), int(shape[1])
        if tuple(words.shape) != (m, words_per_row(n)):
            raise ValueError(f"words has shape {tuple(words.shape)}, expected {(m, words_per_row(n))}")
        self.words = words
        self.shape = (m, n)

    # ------------------------------------------------------------------ construction
    @classmethod
    def from_dense(cls, A) -> "BitMatrix":
        """Pack a dense host array (non-zero -> 1)."""
        A = np.asarray(A)
        if A.ndim != 2:
            raise ValueError("expected a 2-D array")
        m, n = A.shape
        wpr = words_per_row(n)
        by = np.packbits(A != 0, axis=1, bitorder="little")
        out = np.zeros((m, wpr * 4), dtype=np.uint8)
        out[:, : by.shape[1]] = by
        return cls(out.view("<u4"), (m, n))

    @property
    def is_device(self) -> bool:
        return bool(getattr(self.words, "is_cuda", False))

    def _host_words(self) -> np.ndarray:
        w = self.words
        if not isinstance(w, np.ndarray):          # torch tensor (CPU, possibly pinned, or CUDA)
            w = w.cpu().numpy()
        return np.ascontiguousarray(w).view(np.uint32)

    def to_dense(self, dtype=np.float64) -> np.ndarray:
        by = self._host_words().view(np.uint8)
        return np.unpackbits(by, axis=1, bitorder="little")[:, : self.shape[1]].astype(dtype)

    def to_device(self, device):
        """Copy the words to ``device`` (async from pinned host memory)."""
        import torch
        if self.is_device:
            return BitMatrix(self.words.to(device), self.shape)
        w = self.words
        t = torch.from_numpy(np.ascontiguousarray(w).view(np.int32)) if isinstance(w, np.ndarray) else w
        return BitMatrix(t.to(device, non_blocking=t.is_pinned()), self.shape)

    def count(self) -> int:
        """Number of set bits (``np.count_nonzero`` of the dense matrix)."""
        if self.is_device:
            from .device import popcount_device
            return popcount_device(self)
        return int(np.unpackbits(self._host_words().view(np.uint8)).sum())

    def transpose(self) -> "BitMatrix":
        if self.is_device:
            from .device import transpose_device
            return transpose_device(self)
        return BitMatrix.from_dense(self.to_dense(np.uint8).T)

    def __and__(self, other: "BitMatrix") -> "BitMatrix":
        if self.shape != other.shape:
            raise ValueError("shape mismatch")
        return BitMatrix(self.words & other.words, self.shape)

    def rows(self, r0: int, r1: int) -> "BitMatrix":
        return BitMatrix(self.words[r0:r1], (r1 - r0, self.shape[1]))
