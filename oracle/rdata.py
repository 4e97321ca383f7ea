"""Minimal reader for R's RDX3 (XDR) serialisation, enough for the reference's
`.rda` fixtures (R-pkg/data/SVCdata.rda, R-pkg/data/house.rda).

TEST INFRASTRUCTURE ONLY (oracle/): used by tests/golden/make_golden.py to turn
the reference's own data sets into committed fixtures.  Never imported by the
product package.
"""
from __future__ import annotations

import gzip
import lzma
import struct

import numpy as np


class _Reader:
    def __init__(self, buf: bytes):
        self.b = buf
        self.i = 0
        self.refs = []

    def int(self) -> int:
        (v,) = struct.unpack_from(">i", self.b, self.i)
        self.i += 4
        return v

    def ints(self, n):
        a = np.frombuffer(self.b, dtype=">i4", count=n, offset=self.i).astype(np.int64)
        self.i += 4 * n
        return a

    def reals(self, n):
        a = np.frombuffer(self.b, dtype=">f8", count=n, offset=self.i).astype(np.float64)
        self.i += 8 * n
        return a

    def bytes_(self, n):
        v = self.b[self.i:self.i + n]
        self.i += n
        return v

    def item(self):
        flags = self.int()
        typ = flags & 0xFF
        has_attr = bool(flags & 0x200)
        has_tag = bool(flags & 0x400)
        if typ == 254:  # NILVALUE
            return None
        if typ == 255:  # REFSXP
            return self.refs[(flags >> 8) - 1]
        if typ == 1:  # SYMSXP
            name = self.item()
            self.refs.append(name)
            return name
        if typ == 2:  # LISTSXP (pairlist)
            out = []
            while True:
                attr = self.item() if has_attr else None
                tag = self.item() if has_tag else None
                car = self.item()
                out.append((tag, car))
                flags = self.int()
                typ = flags & 0xFF
                has_attr = bool(flags & 0x200)
                has_tag = bool(flags & 0x400)
                if typ == 254:
                    break
                if typ != 2:
                    raise ValueError("unexpected cdr type %d" % typ)
            return out
        if typ == 9:  # CHARSXP
            n = self.int()
            if n == -1:
                return None
            return self.bytes_(n).decode("utf-8", "replace")
        if typ in (10, 13):  # LGLSXP / INTSXP
            n = self.int()
            v = {"type": "int", "data": self.ints(n)}
        elif typ == 14:  # REALSXP
            n = self.int()
            v = {"type": "real", "data": self.reals(n)}
        elif typ == 16:  # STRSXP
            n = self.int()
            v = {"type": "str", "data": [self.item() for _ in range(n)]}
        elif typ == 19:  # VECSXP
            n = self.int()
            v = {"type": "list", "data": [self.item() for _ in range(n)]}
        else:
            raise ValueError("unsupported SEXP type %d at %d" % (typ, self.i))
        if has_attr:
            v["attr"] = dict(self.item())
        return v


def read_rda(path: str) -> dict:
    """Return {object name: parsed object} of an .rda file (gzip or xz)."""
    raw = open(path, "rb").read()
    if raw[:2] == b"\x1f\x8b":
        raw = gzip.decompress(raw)
    elif raw[:6] == b"\xfd7zXZ\x00":
        raw = lzma.decompress(raw)
    if raw[:5] != b"RDX3\n" or raw[5:7] != b"X\n":
        raise ValueError("not an RDX3/XDR file")
    r = _Reader(raw)
    r.i = 7
    r.int(), r.int(), r.int()
    n = r.int()
    r.bytes_(n)  # native encoding
    top = r.item()
    return {tag: car for tag, car in top}


def as_python(obj):
    """Convert a parsed SEXP into numpy arrays / dicts (names + dim honoured)."""
    if obj is None or isinstance(obj, str):
        return obj
    attr = obj.get("attr", {})
    t = obj["type"]
    if t in ("real", "int"):
        a = obj["data"]
        if "dim" in attr:
            a = a.reshape(tuple(int(x) for x in attr["dim"]["data"]), order="F")
        if "levels" in attr:
            return {"codes": a, "levels": attr["levels"]["data"]}
        return a
    if t == "str":
        return list(obj["data"])
    if t == "list":
        vals = [as_python(v) for v in obj["data"]]
        if "names" in attr:
            return dict(zip(attr["names"]["data"], vals))
        return vals
    raise ValueError(t)
