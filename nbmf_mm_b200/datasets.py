"""Data side of the reference's experiment driver: the paper's data sets and their splits.

``read_rda_matrix``: minimal stdlib reader for the R ``.rda`` files the reference ships in ``data/`` (``animals.rda``,
``lastfm.rda``, ``paleo.rda``).  The reference loads them with ``pyreadr`` (``examples/reproduce_magron2022.py:28-29``);
the files are bz2-compressed ``RDX2`` XDR serialisations of one named numeric matrix each (column-major, with ``dim`` and
``dimnames`` attributes), and only the node types those three files use are parsed.  ``load_dataset_and_splits`` mirrors
the driver's function of that name (``:25-38``): the matrix plus the train / validation / test masks of
``data/magron2022/<name>_split.npz``; the reference ships that file for ``animals`` only, for the other two data sets
``make_split`` draws a seeded 70 / 15 / 15 partition of the entries (the driver's seed 12345; the partition itself is
this package's definition -- the reference has no generator to follow)."""
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


def make_split(shape, seed=12345, fractions=(0.70, 0.15, 0.15)):
    """Disjoint train / validation / test masks (float 0/1) covering every entry of a ``shape`` matrix, drawn from
    ``np.random.RandomState(seed)``.  Unpinned: the reference only ships a stored split for ``animals``."""
    rs = np.random.RandomState(seed)
    u = rs.uniform(size=shape)
    a, b = fractions[0], fractions[0] + fractions[1]
    return (u < a).astype(np.float64), ((u >= a) & (u < b)).astype(np.float64), (u >= b).astype(np.float64)


def load_dataset_and_splits(dataset_name, data_dir="data", split_dir=None, seed=12345):
    """``(Y, train_mask, val_mask, test_mask)`` as ``examples/reproduce_magron2022.py:25-38`` returns them: the stored
    split when ``<split_dir>/<name>_split.npz`` exists, else ``make_split(Y.shape, seed)``."""
    from pathlib import Path
    data_dir = Path(data_dir)
    split_dir = data_dir / "magron2022" if split_dir is None else Path(split_dir)
    Y = read_rda_matrix(str(data_dir / f"{dataset_name}.rda"))
    if isinstance(Y, tuple):
        Y = Y[1]
    Y = np.asarray(Y, dtype=np.float64)
    f = split_dir / f"{dataset_name}_split.npz"
    if f.is_file():
        with np.load(f) as z:
            return Y, z["train_mask"].astype(np.float64), z["val_mask"].astype(np.float64), z["test_mask"].astype(np.float64)
    return (Y,) + make_split(Y.shape, seed)
