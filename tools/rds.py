"""Minimal reader for R `.rds` files (gzip + XDR serialisation, format version 2 or 3).

Test/fixture tooling only: it decodes the stored `training_VEM` outputs under
`/root/reference/Simulations/Training/*.rds` so that golden vectors can be committed as plain
JSON under `tests/golden/` (R itself is not available in the build image).

Only the SEXP types that occur in those fixtures are handled: NILSXP, SYMSXP, LISTSXP (attribute
pairlists), CHARSXP, LGLSXP, INTSXP, REALSXP, STRSXP, VECSXP, ALTREP compact sequences, REFSXP.
R objects map to Python as: atomic vectors -> numpy arrays (with `.attrs` dropped unless the
vector carries `names`/`dim`), generic vectors (lists) -> `RList` (a list with `.names`/`.attrs`).
"""
from __future__ import annotations

import gzip
import struct

import numpy as np


class RList(list):
    """An R generic vector; `names` is a list of str (or None), `attrs` the other attributes."""

    names = None
    attrs = None

    def __getitem__(self, key):
        if isinstance(key, str):
            return list.__getitem__(self, self.names.index(key))
        return list.__getitem__(self, key)

    def keys(self):
        return list(self.names or [])

    def as_dict(self):
        return {n: v for n, v in zip(self.names, self)}


class RVector(np.ndarray):
    """numpy array carrying R attributes (names, dim, class...) in `.attrs`."""

    attrs = None

    def __new__(cls, arr, attrs=None):
        obj = np.asarray(arr).view(cls)
        obj.attrs = attrs or {}
        return obj

    def __array_finalize__(self, obj):
        self.attrs = getattr(obj, "attrs", None) or {}


class _Reader:
    def __init__(self, buf: bytes):
        self.b = buf
        self.o = 0
        self.refs = []

    def i32(self) -> int:
        (v,) = struct.unpack_from(">i", self.b, self.o)
        self.o += 4
        return v

    def raw(self, n: int) -> bytes:
        v = self.b[self.o:self.o + n]
        self.o += n
        return v

    def item(self):
        flags = self.i32()
        ty = flags & 0xFF
        has_attr = bool(flags & (1 << 9))
        has_tag = bool(flags & (1 << 10))
        if ty == 254:  # NILVALUE_SXP
            return None
        if ty == 255:  # REFSXP
            return self.refs[(flags >> 8) - 1]
        if ty == 1:  # SYMSXP
            name = self.item()
            self.refs.append(name)
            return name
        if ty == 2:  # LISTSXP: pairlist -> list of (tag, value)
            out = []
            while True:
                if has_attr:
                    self.item()
                tag = self.item() if has_tag else None
                out.append((tag, self.item()))
                flags = self.i32()
                ty = flags & 0xFF
                if ty == 254:
                    return out
                if ty != 2:
                    raise ValueError(f"unexpected pairlist cdr type {ty}")
                has_attr = bool(flags & (1 << 9))
                has_tag = bool(flags & (1 << 10))
        if ty == 9:  # CHARSXP
            n = self.i32()
            return None if n == -1 else self.raw(n).decode("utf-8", "replace")
        if ty == 238:  # ALTREP
            info = self.item()
            state = self.item()
            self.item()  # attributes
            cls = info[0][1] if isinstance(info, list) else None
            if cls == "compact_intseq":
                n, start, step = (int(x) for x in np.asarray(state)[:3])
                return RVector(np.arange(start, start + n * step, step, dtype=np.int32))
            if cls == "compact_realseq":
                n, start, step = (float(x) for x in np.asarray(state)[:3])
                return RVector(start + step * np.arange(int(n)))
            raise ValueError(f"unsupported ALTREP class {cls}")
        if ty in (10, 13):
            n = self.i32()
            v = np.frombuffer(self.raw(4 * n), dtype=">i4").astype(np.int32)
            val = RVector(v.astype(bool) if ty == 10 else v)
        elif ty == 14:
            n = self.i32()
            val = RVector(np.frombuffer(self.raw(8 * n), dtype=">f8").astype(np.float64))
        elif ty == 16:
            n = self.i32()
            val = RList(self.item() for _ in range(n))
        elif ty == 19:
            n = self.i32()
            val = RList(self.item() for _ in range(n))
        else:
            raise ValueError(f"unsupported SEXP type {ty} at offset {self.o}")
        if has_attr:
            attrs = {k: v for k, v in self.item()}
            names = attrs.pop("names", None)
            if isinstance(val, RList):
                val.names = list(names) if names is not None else None
                val.attrs = attrs
            else:
                if names is not None:
                    attrs["names"] = list(names)
                val.attrs = attrs
        return val


def read_rds(path: str):
    with gzip.open(path, "rb") as f:
        buf = f.read()
    if buf[:2] != b"X\n":
        raise ValueError("not an XDR-serialised R object")
    r = _Reader(buf)
    r.o = 2
    version = r.i32()
    r.i32()
    r.i32()
    if version == 3:
        n = r.i32()
        r.raw(n)
    return r.item()
