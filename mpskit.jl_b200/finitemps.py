"""Device-resident FiniteMPS with the gauge bookkeeping of src/states/finitemps.jl and
src/states/orthoview.jl (``AL/AR/AC/CR`` views, setters that invalidate dependent gauges).
Sites and bonds are 0-based: ``CR[i]`` is the bond matrix to the right of site i (``CR[-1]`` left edge)."""
from __future__ import annotations
import numpy as np
from . import _lib
from ._lib import default_context
from .spaces import Space, fuse, infimum, dual
from .tensor import DeviceTensor
from ._hostlinalg import qrpos, lqpos
import ctypes as C


def virtual_spaces(pspace: Space, L, maxV, left=None, right=None):
    """finitemps.jl:183-192 (full-rank-limited bond spaces)."""
    left = left or Space((0,), (1,), pspace.modulus)
    right = right or Space((0,), (1,), pspace.modulus)
    maxVs = list(maxV) if isinstance(maxV, (list, tuple)) else [maxV] * (L - 1)
    V = [None] * (L + 1)
    V[0] = left
    for k in range(1, L):
        V[k] = infimum(fuse(V[k - 1], pspace), maxVs[k - 1])
    V[L] = right
    for k in range(L - 1, 0, -1):
        V[k] = infimum(V[k], fuse(V[k + 1], dual(pspace)))
    return V


def _mask(Vl, pq, Vr):
    s = Vl.index_charges()[:, None, None] + np.asarray(pq)[None, :, None]
    if Vl.modulus:
        s = s % Vl.modulus
    return s == Vr.index_charges()[None, None, :]


class _View:
    def __init__(self, get, set_=None):
        self._get, self._set = get, set_

    def __getitem__(self, i):
        return self._get(i)

    def __setitem__(self, i, v):
        self._set(i, v)


class FiniteMPS:
    def __init__(self, As, Vs, pspace: Space, ctx=None, normalize=True, dtype=np.complex128):
        """As: host tensors A[l,p,r] (dense charged arrays); canonicalised on the host like
        finitemps.jl:143-170 (QRpos chain) followed by the LQpos chain that the lazy AR/CR getters of
        orthoview.jl:49-60 would run on first use; the gauged tensors are then uploaded once."""
        self.ctx = ctx or default_context()
        self.L = len(As)
        self.pspace = pspace
        self.pq = tuple(int(c) for c in pspace.index_charges())
        self.dtype = np.dtype(dtype)
        V = list(Vs)
        As = [np.array(a, dtype=self.dtype) for a in As]
        mod = pspace.modulus
        R = None
        for i in range(self.L):
            Dl, d, Dr = As[i].shape
            rowq = (V[i].index_charges()[:, None] + np.asarray(self.pq)[None, :]).reshape(-1)
            if mod:
                rowq = rowq % mod
            Q, R, newV = qrpos(As[i].reshape(Dl * d, Dr), rowq, V[i + 1].index_charges(), mod)
            if normalize:
                R = R / np.linalg.norm(R)
            if i < self.L - 1:
                V[i + 1] = newV
                As[i] = Q.reshape(Dl, d, newV.dim)
                As[i + 1] = np.einsum('ab,bpc->apc', R, As[i + 1])
            else:
                As[i] = Q.reshape(Dl, d, newV.dim)
                V[i + 1] = newV
        ALh = As
        Ch = [None] * (self.L + 1)
        Ch[self.L] = R
        ARh = [None] * self.L
        ACh = [None] * self.L
        for i in range(self.L - 1, -1, -1):
            ACh[i] = np.einsum('apb,bc->apc', ALh[i], Ch[i + 1])
            Dl, d, Dr = ACh[i].shape
            colq = (V[i + 1].index_charges()[None, :] - np.asarray(self.pq)[:, None]).reshape(-1)
            if mod:
                colq = colq % mod
            Lm, Qm, newV = lqpos(ACh[i].reshape(Dl, d * Dr), V[i].index_charges(), colq, mod)
            assert newV == V[i], "left-canonical state changed its bond space under LQ"
            Ch[i] = Lm
            ARh[i] = Qm.reshape(newV.dim, d, Dr)
        self.V = V
        up = lambda a, Vl, Vr: DeviceTensor.from_dense(self.ctx, a, Vl, self.pq, Vr, self.dtype)
        self.ALs = [up(ALh[i], V[i], V[i + 1]) for i in range(self.L)]
        self.ARs = [up(ARh[i], V[i], V[i + 1]) for i in range(self.L)]
        self.ACs = [None] * self.L
        self.CLs = [DeviceTensor.from_dense(self.ctx, Ch[i][:, None, :], V[i], (0,), V[i], self.dtype) for i in range(self.L + 1)]
        self._views()

    def _views(self):
        self.AL = _View(self._get_AL)
        self.AR = _View(self._get_AR)
        self.AC = _View(self._get_AC, self._set_AC)
        self.CR = _View(lambda i: self._get_CL(i + 1))

    @staticmethod
    def random(rng, pspace, L, maxV, ctx=None, dtype=np.complex128, left=None, right=None):
        """FiniteMPS(rand, eltype, Pspaces, maxVspaces) (finitemps.jl:172-198) with a seeded generator."""
        V = virtual_spaces(pspace, L, maxV, left, right)
        pq = pspace.index_charges()
        As = []
        for i in range(L):
            m = _mask(V[i], pq, V[i + 1])
            a = rng.standard_normal(m.shape)
            if np.dtype(dtype).kind == "c":
                a = a + 1j * rng.standard_normal(m.shape)
            As.append(a * m)
        return FiniteMPS(As, V, pspace, ctx, dtype=dtype)

    def copy(self):
        """Base.copy(psi): shallow copy of the tensor lists (finitemps.jl:210)."""
        o = object.__new__(FiniteMPS)
        o.__dict__.update(self.__dict__)
        o.ALs, o.ARs, o.ACs, o.CLs, o.V = list(self.ALs), list(self.ARs), list(self.ACs), list(self.CLs), list(self.V)
        o._views()
        return o

    def __len__(self):
        return self.L

    # ---- orthoview.jl getters ------------------------------------------------------------------
    def _get_CL(self, i):
        if self.CLs[i] is None:
            raise NotImplementedError("lazy QR/LQ re-gauging on the device (orthoview.jl:49-60) is not available yet; "
                                      "two-site sweeps never need it")
        return self.CLs[i]

    def _get_AL(self, i):
        if self.ALs[i] is None:
            raise NotImplementedError("AL[%d] is stale and device QRpos is not available yet" % i)
        return self.ALs[i]

    def _get_AR(self, i):
        if self.ARs[i] is None:
            raise NotImplementedError("AR[%d] is stale and device LQpos is not available yet" % i)
        return self.ARs[i]

    def _wrap(self, h):
        return DeviceTensor.wrap(self.ctx, h, self.pspace.modulus)

    def _get_AC(self, i):
        """orthoview.jl:95-106"""
        if self.ACs[i] is None and self.ARs[i] is not None:
            h = C.c_void_p()
            self.ctx.check(self.ctx.lib.mpsk_mul_bond_left(self.ctx.h, self._get_CL(i).h, self.ARs[i].h, C.byref(h)))
            self.ACs[i] = self._wrap(h)
        elif self.ACs[i] is None and self.ALs[i] is not None:
            h = C.c_void_p()
            self.ctx.check(self.ctx.lib.mpsk_mul_bond_right(self.ctx.h, self.ALs[i].h, self._get_CL(i + 1).h, C.byref(h)))
            self.ACs[i] = self._wrap(h)
        return self.ACs[i]

    # ---- orthoview.jl:108-143 setters ------------------------------------------------------------
    def _invalidate(self, i):
        if self.ACs[i] is None:
            if i < self.L - 1:
                self._get_AR(i + 1)
            if i > 0:
                self._get_AL(i - 1)
        self.ACs = [None] * self.L
        self.CLs = [None] * (self.L + 1)
        for k in range(i, self.L):
            self.ALs[k] = None
        for k in range(0, i + 1):
            self.ARs[k] = None

    def _set_AC(self, i, v):
        if isinstance(v, tuple):
            a, b = v
            if getattr(a, "_is_bond", False):      # (c, ar)
                self._invalidate(i)
                self.V[i] = a.Vr
                self.CLs[i] = a
                self.ARs[i] = b
            elif getattr(b, "_is_bond", False):    # (al, c)
                self._invalidate(i)
                self.V[i + 1] = b.Vl
                self.CLs[i + 1] = b
                self.ALs[i] = a
            else:
                raise ValueError("invalid value types")
        else:
            self._invalidate(i)
            self.ACs[i] = v

    # ---- read-outs ------------------------------------------------------------------------------
    def entanglement_spectrum(self, i):
        """toolbox.jl:65-69: singular values of CR[i] per sector (host read-out of a small bond matrix)."""
        Cm = self.CR[i]
        blocks = Cm.to_blocks()
        return {Cm.Vl.charges[jl]: np.linalg.svd(b, compute_uv=False) for (m, jl), b in blocks.items()}
