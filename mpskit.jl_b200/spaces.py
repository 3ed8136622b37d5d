"""Host-side graded spaces (charges are integers, 2q for half-integer U(1) irreps; ascending order).
Mirrors the TensorKit space operations used by src/states/finitemps.jl:183-192."""
from __future__ import annotations
import ctypes as C
import numpy as np
from . import _lib


class Space:
    __slots__ = ("charges", "dims", "modulus")

    def __init__(self, charges, dims, modulus=0):
        items = sorted((int(c), int(n)) for c, n in zip(charges, dims) if int(n) > 0)
        self.charges = tuple(c for c, _ in items)
        self.dims = tuple(n for _, n in items)
        self.modulus = int(modulus)
        if len(set(self.charges)) != len(self.charges):
            raise ValueError("duplicate sector")

    @staticmethod
    def from_dict(d, modulus=0):
        return Space(list(d.keys()), list(d.values()), modulus)

    @staticmethod
    def trivial(n):
        return Space((0,), (n,))

    def as_dict(self):
        return dict(zip(self.charges, self.dims))

    @property
    def dim(self):
        return int(sum(self.dims))

    @property
    def nsec(self):
        return len(self.charges)

    def offsets(self):
        return np.concatenate([[0], np.cumsum(self.dims)]).astype(np.int64)

    def index_charges(self):
        return np.repeat(np.asarray(self.charges, dtype=np.int64), self.dims)

    def norm(self, c):
        return int(c) % self.modulus if self.modulus else int(c)

    def __eq__(self, o):
        return isinstance(o, Space) and (self.charges, self.dims, self.modulus) == (o.charges, o.dims, o.modulus)

    def __hash__(self):
        return hash((self.charges, self.dims, self.modulus))

    def __repr__(self):
        return f"Space({self.as_dict()})"

    def c_struct(self):
        """(c_space, keepalive) for the ABI"""
        q = (C.c_int32 * self.nsec)(*self.charges)
        n = (C.c_int64 * self.nsec)(*self.dims)
        return _lib.c_space(self.nsec, q, n, self.modulus), (q, n)


def fuse(a: Space, b: Space) -> Space:
    out = {}
    for ca, na in zip(a.charges, a.dims):
        for cb, nb in zip(b.charges, b.dims):
            c = a.norm(ca + cb)
            out[c] = out.get(c, 0) + na * nb
    return Space.from_dict(out, a.modulus)


def infimum(a: Space, b: Space) -> Space:
    db = b.as_dict()
    return Space.from_dict({c: min(n, db[c]) for c, n in a.as_dict().items() if c in db}, a.modulus)


def dual(a: Space) -> Space:
    return Space.from_dict({a.norm(-c): n for c, n in a.as_dict().items()}, a.modulus)
