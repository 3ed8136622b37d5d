"""Device-resident FiniteMPS with the gauge bookkeeping of src/states/finitemps.jl and
src/states/orthoview.jl (``AL/AR/AC/CR`` views, setters that invalidate dependent gauges).
Sites and bonds are 0-based: ``CR[i]`` is the bond matrix to the right of site i (``CR[-1]`` left edge)."""
from __future__ import annotations
import numpy as np
from . import _lib
from ._lib import default_context
from .spaces import Space, fuse, infimum, dual
from .tensor import DeviceTensor
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


def qrpos(ctx, A, modulus=0):
    """Q, R = leftorth(A; alg=QRpos())  (device; orthoview.jl:58, finitemps.jl:150)"""
    q, r = C.c_void_p(), C.c_void_p()
    ctx.check(ctx.lib.mpsk_qrpos(ctx.h, A.h, C.byref(q), C.byref(r)))
    Q, R = DeviceTensor.wrap(ctx, q, modulus), DeviceTensor.wrap(ctx, r, modulus)
    R._is_bond = True
    return Q, R


def lqpos(ctx, A, modulus=0):
    """L, Q = rightorth(A; alg=LQpos())  (device; orthoview.jl:52-56)"""
    l, q = C.c_void_p(), C.c_void_p()
    ctx.check(ctx.lib.mpsk_lqpos(ctx.h, A.h, C.byref(l), C.byref(q)))
    Lm, Q = DeviceTensor.wrap(ctx, l, modulus), DeviceTensor.wrap(ctx, q, modulus)
    Lm._is_bond = True
    return Lm, Q


def decompose_localmps(T, pspace: Space, left=None, right=None, tol=1e-12):
    """decompose_localmps (src/utility/utility.jl:57-68) for abelian charges: T[left, p_1 .. p_L, right] -> site arrays
    A_k[l, p, r] and their bond spaces; one SVD per charge of the cut bond (the rows (l, p) are grouped by q_l + s_p), the
    singular values stay on the left factor (A = U S, B = V) as in the reference."""
    T = np.asarray(T)
    L = T.ndim - 2
    pq = np.asarray(pspace.index_charges())
    d, mod = len(pq), pspace.modulus
    Vl = left if left is not None else Space((0,), (T.shape[0],), mod)
    As, Vs = [], [Vl]
    cur, curV = T, Vl
    scale = np.abs(T).max() or 1.0
    for _ in range(L - 1):
        Dl = cur.shape[0]
        M = cur.reshape(Dl * d, -1)
        rowq = (curV.index_charges()[:, None] + pq[None, :]).reshape(-1)
        if mod:
            rowq = rowq % mod
        cols, rows_out, charges, dims = [], [], [], []
        for q in sorted(set(int(x) for x in rowq)):
            rows = np.nonzero(rowq == q)[0]
            U, sv, Vh = np.linalg.svd(M[rows], full_matrices=False)
            keep = int(np.sum(sv > tol * scale))
            if keep == 0:
                continue
            blk = np.zeros((M.shape[0], keep), dtype=M.dtype)
            blk[rows] = U[:, :keep] * sv[None, :keep]
            cols.append(blk); rows_out.append(Vh[:keep])
            charges.append(q); dims.append(keep)
        newV = Space(charges, dims, mod)
        As.append(np.concatenate(cols, axis=1).reshape(Dl, d, newV.dim))
        cur = np.concatenate(rows_out, axis=0).reshape((newV.dim,) + cur.shape[2:])
        curV = newV
        Vs.append(newV)
    if right is None:
        rowq = (curV.index_charges()[:, None] + pq[None, :]).reshape(-1)
        nz = np.abs(cur.reshape(len(rowq), -1)).max(axis=1) > tol * scale
        tot = sorted(set(int(x) % mod if mod else int(x) for x in rowq[nz])) or [0]
        if len(tot) != 1 or cur.shape[-1] != 1:
            raise ValueError("decompose_localmps: pass the right edge space of a state without a definite total charge")
        right = Space((tot[0],), (1,), mod)
    As.append(cur.reshape(cur.shape[0], d, cur.shape[-1]))
    Vs.append(right)
    return As, Vs


class FiniteMPS:
    def __init__(self, As, Vs, pspace: Space, ctx=None, normalize=True, dtype=np.complex128):
        """As: site tensors A[l,p,r] -- host arrays (dense charged) or DeviceTensors.  Like finitemps.jl:143-170 the
        constructor left-canonicalises with a QRpos chain (on the device) and keeps the last C; right-canonical
        tensors are produced lazily by the AR / CR getters (orthoview.jl:49-60)."""
        self.ctx = ctx or default_context()
        self.L = len(As)
        self.pspace = pspace
        self.pq = tuple(int(c) for c in pspace.index_charges())
        self.dtype = np.dtype(dtype)
        mod = pspace.modulus
        V = list(Vs)
        cur = [a if isinstance(a, DeviceTensor) else DeviceTensor.from_dense(self.ctx, np.asarray(a), V[i], self.pq, V[i + 1], self.dtype)
               for i, a in enumerate(As)]
        R = None
        for i in range(self.L):
            Q, R = qrpos(self.ctx, cur[i], mod)
            if normalize:
                R.scale(1.0 / R.norm())
            cur[i] = Q
            V[i + 1] = Q.Vr
            if i < self.L - 1:
                h = C.c_void_p()
                self.ctx.check(self.ctx.lib.mpsk_mul_bond_left(self.ctx.h, R.h, cur[i + 1].h, C.byref(h)))
                cur[i + 1] = DeviceTensor.wrap(self.ctx, h, mod)
        self.V = V
        self.ALs = cur
        self.ARs = [None] * self.L
        self.ACs = [None] * self.L
        self.CLs = [None] * (self.L + 1)
        self.CLs[self.L] = R
        self._views()

    def _views(self):
        self.AL = _View(self._get_AL)
        self.AR = _View(self._get_AR)
        self.AC = _View(self._get_AC, self._set_AC)
        self.CR = _View(lambda i: self._get_CL(i + 1), lambda i, v: self._set_CR(i, v))

    @staticmethod
    def random(rng, pspace, L, maxV, ctx=None, dtype=np.complex128, left=None, right=None, on_device=False):
        """FiniteMPS(rand, eltype, Pspaces, maxVspaces) (finitemps.jl:172-198).  `rng`: numpy Generator (host-generated
        normal entries, shared with the oracle in the parity tests) or, with on_device=True, an integer seed for
        device-generated uniform entries (large synthetic states)."""
        ctx = ctx or default_context()
        V = virtual_spaces(pspace, L, maxV, left, right)
        pq = pspace.index_charges()
        As = []
        for i in range(L):
            if on_device:
                As.append(DeviceTensor.random(ctx, V[i], [int(c) for c in pq], V[i + 1], int(rng) + i, dtype))
                continue
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

    # ---- the state as a vector (finitemps.jl:372-469) -----------------------------------------------------------------
    def _site_tensors(self):
        """(AC[0], AR[1], ..., AR[L-1]): site tensors whose product is the state including its norm"""
        return [self.AC[0]] + [self.AR[i] for i in range(1, self.L)]

    def norm(self):
        """norm(psi) = norm(psi.AC[1]) (finitemps.jl:467)"""
        return float(self.AC[0].norm())

    def dot(self, other):
        """dot(self, other) = <self|other> (finitemps.jl:459-464), through the rectangular overlap environments of
        `approximate` (FinEnv with a unit operator and `other` above) and the projection ac_proj (derivatives.jl:216-221)"""
        from .environments import environments
        from .approximate import ac_proj
        from .operators import MPOHamiltonian, SparseMPOSlice
        if len(self) != len(other):
            raise ValueError("MPS with different length")
        unit = MPOHamiltonian([SparseMPOSlice([[1.0]], (0,), (0,), self.pq, self.pspace.modulus)])
        envs = environments(self, unit, other)
        y = ac_proj(0, self, envs)
        val = complex(self.AC[0].dot(y))
        y.free()
        return val

    def __mul__(self, a):
        """psi * a (finitemps.jl:372-373,445-455): the factor sits on the centre tensors"""
        T = [t.copy() for t in self._site_tensors()]
        T[0].scale(a)
        return FiniteMPS(T, list(self.V), self.pspace, self.ctx, normalize=False, dtype=self.dtype)

    __rmul__ = __mul__

    def normalize(self):
        """normalize(psi) (finitemps.jl:468-469)"""
        return self * (1.0 / self.norm())

    def __add__(self, other):
        """psi1 + psi2 (finitemps.jl:375-441): direct sum of the bond spaces on the device (mpsk_tensor_cat) -- first tensor
        [A1 A2], inner tensors block diagonal, last tensor [A1; A2] -- a right-to-left LQ sweep that cuts the summed bonds of
        the right half and the constructor's QR chain for the left half (the leftorth! / rightorth! steps of the reference)."""
        from .changebonds import cat
        if len(self) != len(other):
            raise ValueError(f"Cannot add states of length {len(self)} and {len(other)}")
        L, ctx, mod = self.L, self.ctx, self.pspace.modulus
        assert L > 1, "not implemented for length < 2"
        Ta, Tb = self._site_tensors(), other._site_tensors()
        out = []
        for i in range(L):
            if i == 0:
                out.append(cat(Ta[i], Tb[i], 1))
            elif i == L - 1:
                out.append(cat(Ta[i], Tb[i], 0))
            else:
                za = DeviceTensor.zeros(ctx, Ta[i].Vl, Ta[i].mch, Tb[i].Vr, self.dtype)
                zb = DeviceTensor.zeros(ctx, Tb[i].Vl, Tb[i].mch, Ta[i].Vr, self.dtype)
                top, bot = cat(Ta[i], za, 1), cat(zb, Tb[i], 1)
                out.append(cat(top, bot, 0))
                for t in (za, zb, top, bot):
                    t.free()
        for i in range(L - 1, 0, -1):
            Lm, Q = lqpos(ctx, out[i], mod)
            out[i] = Q
            h = C.c_void_p()
            ctx.check(ctx.lib.mpsk_mul_bond_right(ctx.h, out[i - 1].h, Lm.h, C.byref(h)))
            out[i - 1] = DeviceTensor.wrap(ctx, h, mod)
        V = [out[0].Vl] + [t.Vr for t in out]
        return FiniteMPS(out, V, self.pspace, ctx, normalize=False, dtype=self.dtype)

    def to_dense(self):
        """convert(TensorMap, psi) (finitemps.jl:351-362): the state as one dense tensor [left, p_1 .. p_L, right] on the host
        (exponential in L: small chains, tests, exact diagonalisation cross-checks)"""
        T = None
        for t in self._site_tensors():
            a = t.to_dense()
            T = a if T is None else np.tensordot(T, a, axes=([T.ndim - 1], [0]))
        return T

    @staticmethod
    def from_dense(T, pspace, ctx=None, left=None, right=None, tol=1e-12, dtype=np.complex128):
        """FiniteMPS(decompose_localmps(T)) (utility.jl:57-68, finitemps.jl:143-170): successive charged SVDs of the dense
        tensor T[left, p_1 .. p_L, right] on the host (state construction), singular values below tol * s_max dropped, then the
        device constructor.  left / right: the edge bond spaces (default: one-dimensional, charge 0 on the left and the total
        charge found in T on the right)."""
        As, Vs = decompose_localmps(T, pspace, left, right, tol)
        return FiniteMPS(As, Vs, pspace, ctx, normalize=False, dtype=dtype)

    def __neg__(self):
        return self * (-1.0)

    def __sub__(self, other):                      # finitemps.jl:443
        return self + (other * (-1.0))

    # ---- orthoview.jl getters ------------------------------------------------------------------
    def _get_CL(self, i):
        """CR[i-1] (orthoview.jl:49-60): LQpos of AC[i] when the left neighbour is left-canonical, else QRpos of AC[i-1]."""
        if self.CLs[i] is None:
            mod = self.pspace.modulus
            if i == 0 or self.ALs[i - 1] is not None:
                Lm, Q = lqpos(self.ctx, self._get_AC(i), mod)
                self.CLs[i], self.ARs[i] = Lm, Q
            else:
                Q, R = qrpos(self.ctx, self._get_AC(i - 1), mod)
                self.ALs[i - 1], self.CLs[i] = Q, R
        return self.CLs[i]

    def _get_AL(self, i):
        """orthoview.jl:6-9"""
        if self.ALs[i] is None:
            self._get_CL(i + 1)
        return self.ALs[i]

    def _get_AR(self, i):
        """orthoview.jl:28-32"""
        if self.ARs[i] is None:
            self._get_CL(i)
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
            def isbond(t):
                return getattr(t, "_is_bond", False) or (len(t.mch) == 1 and t.mch[0] == 0 and len(self.pq) != 1)
            if isbond(a):      # (c, ar)
                self._invalidate(i)
                self.V[i] = a.Vr
                self.CLs[i] = a
                self.ARs[i] = b
            elif isbond(b):    # (al, c)
                self._invalidate(i)
                self.V[i + 1] = b.Vl
                self.CLs[i + 1] = b
                self.ALs[i] = a
            else:
                raise ValueError("invalid value types")
        else:
            self._invalidate(i)
            self.ACs[i] = v

    def _set_CR(self, i, vec):
        """orthoview.jl:62-77 (CRView setindex!)"""
        if self.CLs[i + 1] is None:
            self._get_CL(i + 1)
        if i >= 0 and self.ALs[i] is None:
            self._get_AL(i)
        if i + 1 < self.L and self.ARs[i + 1] is None:
            self._get_AR(i + 1)
        self.CLs = [None] * (self.L + 1)
        self.ACs = [None] * self.L
        for k in range(i + 1, self.L):
            self.ALs[k] = None
        for k in range(0, i + 1):
            self.ARs[k] = None
        vec._is_bond = True
        self.CLs[i + 1] = vec

    # ---- read-outs ------------------------------------------------------------------------------
    def entanglement_spectrum(self, i):
        """toolbox.jl:65-69: singular values of CR[i] per sector, computed on the device (the bond matrix is factorised by
        the library's per-sector SVD, only the spectrum comes back to the host)."""
        Cm = self.CR[i]
        ctx = self.ctx
        pc = (C.c_int32 * 1)(0)
        U, S, Vh = C.c_void_p(), C.c_void_p(), C.c_void_p()
        eps = C.c_double()
        ctx.check(ctx.lib.mpsk_tsvd(ctx.h, Cm.h, 1, pc, _lib.c_trunc(_lib.NOTRUNC | 0x100, 0.0), C.byref(U), C.byref(S), C.byref(Vh), C.byref(eps)))
        for h in (U, S, Vh):
            ctx.lib.mpsk_tensor_free(ctx.h, h)
        from .dmrg import tsvd_spectrum
        return {q: vals for q, (vals, kept) in tsvd_spectrum(ctx).items()}


# ---- space accessors (src/states/abstractmps.jl:141-161; 0-based sites) ---------------------------------------------------
def left_virtualspace(psi, i):
    """left_virtualspace(psi, i): the bond space to the left of site i (periodic for an InfiniteMPS)"""
    return psi.V[i % psi.N] if hasattr(psi, "N") and not hasattr(psi, "L") else psi.V[i]


def right_virtualspace(psi, i):
    """right_virtualspace(psi, i): the bond space to the right of site i"""
    return psi.V[(i + 1) % psi.N] if hasattr(psi, "N") and not hasattr(psi, "L") else psi.V[i + 1]


def physicalspace(psi, i=0):
    """physicalspace(psi, i): one physical space for all sites in this package"""
    return psi.pspace


def max_Ds(psi):
    """max_Ds(psi::FiniteMPS) (finitemps.jl:297-303): min of the cumulative products of the physical dimensions from the two
    edge spaces -- the largest dimension each of the L + 1 bonds could carry"""
    L, d = len(psi), psi.pspace.dim
    dl, dr = psi.V[0].dim, psi.V[L].dim
    return [min(dl * d ** k, dr * d ** (L - k)) for k in range(L + 1)]
