"""Abelian (U(1) / Z_N / trivial) graded vector spaces.

Charges are plain integers (2q for half-integer U(1) irreps, as SURVEY.md 8b proposes for the C
ABI).  Sector order is ascending charge; TensorKit's own U1Irrep iteration order (|q| ascending,
+ before -) only matters inside the Julia adapter and is isolated there.
Restates the TensorKit operations the path uses through src/states/finitemps.jl:183-192
(``fuse``, ``infimum``, ``dual``, ``oneunit``).
"""
from __future__ import annotations
from dataclasses import dataclass
import numpy as np


@dataclass(frozen=True)
class Space:
    charges: tuple
    dims: tuple
    modulus: int = 0  # 0: U(1) (integers); N>0: Z_N

    def __post_init__(self):
        ch = tuple(int(c) for c in self.charges)
        dm = tuple(int(d) for d in self.dims)
        order = sorted(range(len(ch)), key=lambda i: ch[i])
        object.__setattr__(self, "charges", tuple(ch[i] for i in order))
        object.__setattr__(self, "dims", tuple(dm[i] for i in order))
        assert len(set(self.charges)) == len(self.charges), "duplicate sector"
        assert all(d > 0 for d in self.dims), "empty sectors must be dropped"

    @staticmethod
    def from_dict(d, modulus=0):
        items = [(c, n) for c, n in d.items() if n > 0]
        return Space(tuple(c for c, _ in items), tuple(n for _, n in items), modulus)

    @staticmethod
    def trivial(n):
        return Space((0,), (n,))

    @property
    def dim(self):
        return int(sum(self.dims))

    @property
    def nsec(self):
        return len(self.charges)

    def as_dict(self):
        return dict(zip(self.charges, self.dims))

    def offsets(self):
        return np.concatenate([[0], np.cumsum(self.dims)]).astype(np.int64)

    def index_charges(self):
        """charge of every basis index, in storage order."""
        return np.repeat(np.asarray(self.charges, dtype=np.int64), self.dims)

    def sector(self, charge):
        """(index, offset, dim) of the sector with that charge or None."""
        try:
            i = self.charges.index(int(charge))
        except ValueError:
            return None
        return i, int(self.offsets()[i]), self.dims[i]

    def norm(self, c):
        return int(c) % self.modulus if self.modulus else int(c)


def _norm(c, modulus):
    return int(c) % modulus if modulus else int(c)


def fuse(a: Space, b: Space) -> Space:
    out = {}
    for ca, na in zip(a.charges, a.dims):
        for cb, nb in zip(b.charges, b.dims):
            c = _norm(ca + cb, a.modulus)
            out[c] = out.get(c, 0) + na * nb
    return Space.from_dict(out, a.modulus)


def infimum(a: Space, b: Space) -> Space:
    da, db = a.as_dict(), b.as_dict()
    return Space.from_dict({c: min(n, db[c]) for c, n in da.items() if c in db}, a.modulus)


def dual(a: Space) -> Space:
    return Space.from_dict({_norm(-c, a.modulus): n for c, n in a.as_dict().items()}, a.modulus)
