"""Minimal stdlib reader for the R ``.rda`` files the reference ships in ``data/``.

TEST INFRASTRUCTURE (used by ``oracle/make_golden.py`` only).  The reference
loads these with ``pyreadr`` (``examples/reproduce_magron2022.py:28-29``), which
is not installed here.  The files are bz2-compressed ``RDX2`` XDR
serialisations of one named numeric matrix each (column-major, with ``dim`` and
``dimnames`` attributes).  Only the node types those three files use are parsed.
"""
from __future__ import annotations

import bz2
import gzip
import struct

import numpy as np

_NIL, _SYM, _LIST, _CHAR, _LGL, _INT, _REAL, _STR, _VEC = 254, 1, 2, 9, 10, 13, 14, 16, 19
_REF, _NAMESPACE, _GLOBALENV, _EMPTYENV, _BASEENV = 255, 249, 253, 242, 241


class _Reader:
    def __init__(self, buf: bytes):
        self.b = buf
        self.o = 0
        self.syms = []

    def i32(self):
        v = struct.unpack_from(">i", self.b, self.o)[0]
        self.o += 4
        return v

    def take(self, n):
        v = self.b[self.o:self.o + n]
        self.o += n
        return v

    def item(self):
        flags = self.i32()
        ty = flags & 0xFF
        has_attr = bool(flags & 0x200)
        has_tag = bool(flags & 0x400)
        if ty in (_NIL, _GLOBALENV, _EMPTYENV, _BASEENV):
            return None
        if ty == _REF:
            idx = flags >> 8
            if idx == 0:
                idx = self.i32()
            return self.syms[idx - 1]
        if ty == _SYM:
            name = self.item()
            self.syms.append(name)
            return name
        if ty == _LIST:
            out = []
            while True:
                attr = self.item() if has_attr else None
                tag = self.item() if has_tag else None
                car = self.item()
                out.append((tag, car, attr))
                flags = self.i32()
                ty = flags & 0xFF
                has_attr = bool(flags & 0x200)
                has_tag = bool(flags & 0x400)
                if ty != _LIST:
                    if ty != _NIL:
                        raise ValueError(f"unexpected pairlist tail type {ty}")
                    return out
        if ty == _CHAR:
            n = self.i32()
            return None if n == -1 else self.take(n).decode("utf-8", "replace")
        if ty in (_INT, _LGL):
            n = self.i32()
            v = np.frombuffer(self.take(4 * n), dtype=">i4").astype(np.int64)
        elif ty == _REAL:
            n = self.i32()
            v = np.frombuffer(self.take(8 * n), dtype=">f8").astype(np.float64)
        elif ty == _STR:
            n = self.i32()
            v = [self.item() for _ in range(n)]
        elif ty == _VEC:
            n = self.i32()
            v = [self.item() for _ in range(n)]
        else:
            raise ValueError(f"unsupported SEXP type {ty} at offset {self.o}")
        attrs = {}
        if has_attr:
            for tag, car, _ in self.item() or []:
                attrs[tag] = car
        return {"value": v, "attrs": attrs}


def read_rda_matrix(path):
    """Return ``(name, ndarray[float64] (rows x cols))`` of the single matrix in ``path``."""
    raw = open(path, "rb").read()
    if raw[:3] == b"BZh":
        raw = bz2.decompress(raw)
    elif raw[:2] == b"\x1f\x8b":
        raw = gzip.decompress(raw)
    if raw[:5] != b"RDX2\n" or raw[5:7] != b"X\n":
        raise ValueError("not an RDX2 XDR file")
    r = _Reader(raw)
    r.o = 7
    r.i32(); r.i32(); r.i32()                      # format version, writer, min reader
    top = r.item()
    tag, obj, _ = top[0]
    dims = obj["attrs"]["dim"]["value"]
    mat = np.asarray(obj["value"], dtype=np.float64).reshape(int(dims[1]), int(dims[0])).T
    return tag, np.ascontiguousarray(mat)
