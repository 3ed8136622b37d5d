"""Block-sparse abelian tensors of the generic three-leg shape T[l, m, r].

``l`` and ``r`` are graded bond spaces, ``m`` is a flat list of labels each carrying a definite
charge; charge conservation is  q_r = q_l + mch[m].  Every tensor on the hot path is of this
shape (SURVEY.md section 8 notation):

  MPS tensor  A[l p; r]            m = physical basis state p,         mch = s_p
  two-site    x[l p1; r p2]        m = p1*d + p2,                      mch = s_p1 + s_p2
  bond        C[l; r]              m = single label,                   mch = 0
  left env    GL[l_bra w; l_ket]   m = MPO level i (width-1 levels),   mch = -c_i
  right env   GR[r_ket w; r_bra]   m = MPO level k,                    mch = +c_k

(the env rules follow from transfer.jl:105-110 applied to the FinEnv start vectors,
FinEnv.jl:41-70; MPO levels wider than one are flattened into width-1 levels).

One block per (m, l-sector) whose r-sector exists: a dense (n_l, n_r) array.  This mirrors the
TensorKit data handed across the C ABI (per-sector contiguous matrices).
"""
from __future__ import annotations
import numpy as np
from .spaces import Space


class BlockTensor:
    def __init__(self, Vl: Space, mch, Vr: Space, dtype=np.complex128, blocks=None):
        self.Vl, self.Vr = Vl, Vr
        self.mch = tuple(int(c) for c in mch)
        self.dtype = np.dtype(dtype)
        self.blocks = {}
        for m, jl, jr in self.block_keys(Vl, self.mch, Vr):
            self.blocks[(m, jl)] = np.zeros((Vl.dims[jl], Vr.dims[jr]), dtype=self.dtype)
        if blocks is not None:
            for k, v in blocks.items():
                assert k in self.blocks and self.blocks[k].shape == v.shape, k
                self.blocks[k] = np.array(v, dtype=self.dtype)

    @staticmethod
    def block_keys(Vl, mch, Vr):
        """bit-exact bookkeeping: ordered list of (m, l-sector index, r-sector index)."""
        out = []
        rpos = {c: i for i, c in enumerate(Vr.charges)}
        for m, cm in enumerate(mch):
            for jl, ql in enumerate(Vl.charges):
                jr = rpos.get(Vl.norm(ql + cm))
                if jr is not None:
                    out.append((m, jl, jr))
        return out

    def keys3(self):
        return self.block_keys(self.Vl, self.mch, self.Vr)

    @property
    def nm(self):
        return len(self.mch)

    def numel(self):
        return sum(b.size for b in self.blocks.values())

    def copy(self):
        return BlockTensor(self.Vl, self.mch, self.Vr, self.dtype,
                           {k: v.copy() for k, v in self.blocks.items()})

    def to_dense(self):
        out = np.zeros((self.Vl.dim, self.nm, self.Vr.dim), dtype=self.dtype)
        lo, ro = self.Vl.offsets(), self.Vr.offsets()
        for m, jl, jr in self.keys3():
            out[lo[jl]:lo[jl + 1], m, ro[jr]:ro[jr + 1]] = self.blocks[(m, jl)]
        return out

    @staticmethod
    def from_dense(arr, Vl, mch, Vr, dtype=None, check=True):
        dtype = arr.dtype if dtype is None else dtype
        t = BlockTensor(Vl, mch, Vr, dtype)
        lo, ro = Vl.offsets(), Vr.offsets()
        for m, jl, jr in t.keys3():
            t.blocks[(m, jl)] = np.array(arr[lo[jl]:lo[jl + 1], m, ro[jr]:ro[jr + 1]], dtype=dtype)
        if check:
            resid = np.linalg.norm(arr - t.to_dense())
            assert resid <= 1e-12 * max(1.0, np.linalg.norm(arr)), f"dense array is not symmetric: {resid}"
        return t

    @staticmethod
    def random(rng, Vl, mch, Vr, dtype=np.complex128):
        t = BlockTensor(Vl, mch, Vr, dtype)
        for k, b in t.blocks.items():
            if t.dtype.kind == "c":
                t.blocks[k] = rng.standard_normal(b.shape) + 1j * rng.standard_normal(b.shape)
            else:
                t.blocks[k] = rng.standard_normal(b.shape)
        return t

    def norm(self):
        return float(np.sqrt(sum(np.vdot(b, b).real for b in self.blocks.values())))


def mask3(Vl: Space, mch, Vr: Space):
    """boolean (Dl, nm, Dr) mask of the symmetry-allowed entries."""
    ql = Vl.index_charges()[:, None, None]
    qm = np.asarray(mch, dtype=np.int64)[None, :, None]
    qr = Vr.index_charges()[None, None, :]
    s = ql + qm
    if Vl.modulus:
        s = s % Vl.modulus
    return s == qr
