"""changebonds for finite MPS.  `SvdCut` (changebonds/svdcut.jl:14-24): truncated SVD of every bond matrix.
`changebonds(psi, H, OptimalExpand(trscheme))`: mirror of src/algorithms/changebonds/optimalexpand.jl:80-110.
Every tensor operation runs on the device: two-site matvec, leftnull / rightnull (full Householder QR per coupled sector),
the projections onto the null spaces, the truncated SVD of the projected two-site tensor, direct sums along the bonds, QRpos."""
from __future__ import annotations
import ctypes as C
from dataclasses import dataclass, field
import numpy as np
from .tensor import DeviceTensor
from .derivatives import ddAC2
from .environments import environments
from .dmrg import contract_ac2, truncdim, notrunc, tsvd, _mul_right
from .finitemps import qrpos
from .operators import SparseMPOSlice


def _call(ctx, fn, *args, modulus=0):
    h = C.c_void_p()
    ctx.check(fn(ctx.h, *args, C.byref(h)))
    return DeviceTensor.wrap(ctx, h, modulus)


def leftnull(A):
    """leftnull(A): N[l p; k'] with N' N = 1 and N' A = 0 (optimalexpand.jl:95)"""
    return _call(A.ctx, A.ctx.lib.mpsk_leftnull, A.h, modulus=A.Vl.modulus)


def rightnull(A):
    """rightnull(_transpose_tail(A)): N[k' p; r] with N N' = 1 and A N' = 0 (optimalexpand.jl:96)"""
    return _call(A.ctx, A.ctx.lib.mpsk_rightnull, A.h, modulus=A.Vl.modulus)


def contract_adjoint_left(N, x2):
    """adjoint(N) * x2: B[k' p2; r] = sum_{l p1} conj(N[l p1; k']) x2[l p1; r p2]"""
    return _call(N.ctx, N.ctx.lib.mpsk_contract_adjoint_left, N.h, x2.h, modulus=N.Vl.modulus)


def cat(A, B, axis):
    """catcodomain (axis 0: left bond) / catdomain (axis 1: right bond)"""
    return _call(A.ctx, A.ctx.lib.mpsk_tensor_cat, A.h, B.h, int(axis), modulus=A.Vl.modulus)


@dataclass
class OptimalExpand:
    """optimalexpand.jl:12-14"""
    trscheme: tuple = field(default_factory=lambda: truncdim(1))


_unit = {}


def _right_overlap(psi, B, N):
    """sum_{p, r} B[k' p r] conj(N[k'' p r]) = B * adjoint(N): the MPO-free right transfer of the identity (transfer.jl:38-45)"""
    ctx = psi.ctx
    key = (psi.pq, psi.pspace.modulus)
    if key not in _unit:
        _unit[key] = SparseMPOSlice([[1.0]], (0,), (0,), psi.pq, psi.pspace.modulus)
    V = B.Vr
    eye = np.zeros((V.dim, 1, V.dim), dtype=psi.dtype)
    eye[:, 0, :] = np.eye(V.dim)
    one = DeviceTensor.from_dense(ctx, eye, V, (0,), V, psi.dtype)
    h = C.c_void_p()
    ctx.check(ctx.lib.mpsk_transfer_right(ctx.h, _unit[key].device(ctx), B.h, N.h, one.h, C.byref(h)))
    one.free()
    return DeviceTensor.wrap(ctx, h, psi.pspace.modulus)


def changebonds_inplace(psi, H, alg: OptimalExpand, envs=None, log=None):
    """changebonds!(psi::FiniteMPS, H, alg::OptimalExpand, envs) (optimalexpand.jl:84-110): the state is unchanged, every bond
    gains the dominant directions of the two-site update orthogonal to the current tensors (zeros on site i, optimal
    vectors on site i + 1)."""
    envs = envs if envs is not None else environments(psi, H)
    ctx = psi.ctx
    L = len(psi)
    for i in range(L - 1):
        ac = psi.AC[i]; ar = psi.AR[i + 1]
        NL = leftnull(ac)
        NR = rightnull(ar)
        if NL.Vr.dim == 0 or NR.Vl.dim == 0:          # the bond is already full rank on one side: nothing to add
            continue
        ac2 = contract_ac2(psi, ac, ar)
        h = ddAC2(i, psi, H, envs, ac2)
        AC2 = h(ac2); h.free(); ac2.free()
        B = contract_adjoint_left(NL, AC2)            # adjoint(NL) * AC2
        inter = _right_overlap(psi, B, NR)            # ... * adjoint(NR): bond tensor [k'; k'']

        class _P:                                      # tsvd of a bond tensor: one trivial label on the left
            pass
        P = _P(); P.ctx, P.pq, P.pspace = ctx, (0,), psi.pspace
        U, S, V, eps = tsvd(P, inter, alg.trscheme, normalize=False)
        if log is not None:
            from .dmrg import tsvd_spectrum
            log.append(dict(pos=i, added=V.Vl.as_dict(), spectrum={q: s for q, (s, k) in tsvd_spectrum(ctx).items()}))
        Vb = V                                          # [n; k''] as a bond tensor (single label of charge 0)
        h2 = C.c_void_p()
        ctx.check(ctx.lib.mpsk_mul_bond_left(ctx.h, Vb.h, NR.h, C.byref(h2)))
        ar_re = DeviceTensor.wrap(ctx, h2, psi.pspace.modulus)                    # V * NR: [n p; r]
        ar_le = DeviceTensor.zeros(ctx, ac.Vl, ac.mch, ar_re.Vl, psi.dtype)     # zeros: [l p; n]
        nal, nc = qrpos(ctx, cat(ac, ar_le, 1), psi.pspace.modulus)              # leftorth!(catdomain(AC, 0); QRpos)
        nar = cat(ar, ar_re, 0)                                                   # catcodomain(AR, V * NR)
        psi.AC[i] = (nal, nc)
        psi.AC[i + 1] = (nc, nar)
        for t in (NL, NR, AC2, B, inter, U, S, ar_le):
            t.free()
        envs.drop_stale(psi)
    return psi, envs


def changebonds(psi, H=None, alg=None, envs=None, log=None):
    """changebonds(psi, H, alg::OptimalExpand, envs) = changebonds!(copy(psi), ...) (optimalexpand.jl:80-82);
    changebonds(psi, alg::SvdCut) = changebonds!(copy(psi), alg) (svdcut.jl:13)"""
    if isinstance(H, SvdCut) and alg is None:
        H, alg = None, H
    if isinstance(alg, SvdCut):
        return svdcut_inplace(psi.copy(), alg, log)
    return changebonds_inplace(psi.copy(), H, alg, envs, log)


@dataclass
class SvdCut:
    """svdcut.jl:9-11"""
    trscheme: tuple = field(default_factory=notrunc)


class _BondLeg:
    """what `tsvd` needs to know to factorise a bond tensor: one trivial label on the left"""
    def __init__(self, psi):
        self.ctx, self.pq, self.pspace = psi.ctx, (0,), psi.pspace


def svdcut_inplace(psi, alg: SvdCut, log=None):
    """changebonds!(psi::AbstractFiniteMPS, alg::SvdCut) (svdcut.jl:14-24): for i = L-1 ... 1: U, S, V = tsvd(CR[i]; trunc);
    AC[i] = (AL[i] * U, S); AC[i+1] = (S, V * AR[i+1]); normalize!(psi).  The SVD of the bond matrix, the two bond
    multiplications and the final scaling all run on the device."""
    ctx = psi.ctx
    P = _BondLeg(psi)
    for i in range(len(psi) - 2, -1, -1):
        Cm = psi.CR[i]
        AL, AR = psi.AL[i], psi.AR[i + 1]
        U, S, V, eps = tsvd(P, Cm, alg.trscheme, normalize=False)
        if log is not None:
            from .dmrg import tsvd_spectrum
            log.append(dict(pos=i, kept=S.Vl.as_dict(), eps_trunc=eps, spectrum={q: s for q, (s, k) in tsvd_spectrum(ctx).items()}))
        al = _mul_right(psi, AL, U)                                           # AL[i] * U
        h = C.c_void_p()
        ctx.check(ctx.lib.mpsk_mul_bond_left(ctx.h, V.h, AR.h, C.byref(h)))     # V * AR[i+1]
        ar = DeviceTensor.wrap(ctx, h, psi.pspace.modulus)
        psi.AC[i] = (al, S)
        psi.AC[i + 1] = (S, ar)
        U.free(); V.free()
    c = psi.CR[0]
    c.scale(1.0 / c.norm())                                                   # normalize!(psi)
    psi.ACs = [None] * psi.L
    return psi
