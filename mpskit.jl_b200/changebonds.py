"""changebonds for finite and infinite MPS.  `SvdCut` (changebonds/svdcut.jl:14-24): truncated SVD of every bond matrix.
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
from .spaces import infimum
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
    rand = isinstance(alg, RandExpand)                # changebonds!(psi, alg::RandExpand) (randexpand.jl:41-63): random AC2
    if not rand:
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
        if rand:
            AC2 = DeviceTensor.random(ctx, ac2.Vl, ac2.mch, ac2.Vr, alg.seed + i, psi.dtype); ac2.free()
        else:
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
        if envs is not None:
            envs.drop_stale(psi)
    if rand:
        k = max(L - 2, 0)
        c = psi.CR[k].copy()                          # normalize!(psi) (randexpand.jl:62); psi is a shallow copy: do not scale
        c.scale(1.0 / c.norm())                       # a bond tensor it may share with the input state
        c._is_bond = True
        psi.CLs[k + 1] = c
        psi.ACs = [None] * psi.L
        return psi
    return psi, envs


def changebonds(psi, H=None, alg=None, envs=None, log=None):
    """changebonds(psi, H, alg::OptimalExpand, envs) = changebonds!(copy(psi), ...) (optimalexpand.jl:80-82);
    changebonds(psi, alg::SvdCut) = changebonds!(copy(psi), alg) (svdcut.jl:13)"""
    if alg is None:                                   # changebonds(psi, alg)
        H, alg = None, H
    from .operators import DenseMPO
    if isinstance(psi, DenseMPO) and isinstance(alg, SvdCut):      # svdcut.jl:26-28
        return svdcut_densempo(psi, alg, log=log)
    from .algorithms import UnionAlg
    if isinstance(alg, UnionAlg):                     # unionalg.jl:13-23
        if H is None:
            return changebonds(changebonds(psi, alg.alg1, log=log), alg.alg2, log=log)
        psi, envs = changebonds(psi, H, alg.alg1, envs, log)
        return changebonds(psi, H, alg.alg2, envs, log)
    if hasattr(psi, "N") and not hasattr(psi, "L"):   # InfiniteMPS
        return changebonds_infinite(psi, H, alg, envs, log)
    if isinstance(alg, SvdCut):
        out = svdcut_inplace(psi.copy(), alg, log)
        return (out, envs) if H is not None else out   # svdcut.jl:48-50
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


# ------------------------------------------------------------------------------------------ InfiniteMPS
@dataclass
class RandExpand:
    """randexpand.jl:11-13"""
    trscheme: tuple = field(default_factory=lambda: truncdim(1))
    seed: int = 0


@dataclass
class VUMPSSvdCut:
    """vumpssvd.jl:12-17"""
    tol_gauge: float = 1e-14
    tol: float = 1e-12
    tol_eigenval: float = 1e-12
    trscheme: tuple = field(default_factory=notrunc)


def _bond_leg(st):
    class _P:
        pass
    P = _P(); P.ctx, P.pq, P.pspace = st.ctx, (0,), st.pspace
    return P


def _mul_left(Cm, A):
    return _call(A.ctx, A.ctx.lib.mpsk_mul_bond_left, Cm.h, A.h, modulus=A.Vl.modulus)


def _mul_right_t(A, Cm):
    return _call(A.ctx, A.ctx.lib.mpsk_mul_bond_right, A.h, Cm.h, modulus=A.Vl.modulus)


def _expansion_vectors(st, i, AC2, trscheme, log):
    """VL = leftnull(AL[i]); VR = rightnull(AR[i+1]); U, _, V = tsvd(VL' * AC2 * VR'; trunc) -> (VL * U, V * VR)
    (optimalexpand.jl:26-33, randexpand.jl:23-30), all on the device"""
    N = st.N
    VL = leftnull(st.AL[i % N])
    VR = rightnull(st.AR[(i + 1) % N])
    B = contract_adjoint_left(VL, AC2)
    inter = _right_overlap(st, B, VR)
    U, S, V, eps = tsvd(_bond_leg(st), inter, trscheme, normalize=False)
    if log is not None:
        from .dmrg import tsvd_spectrum
        log.append(dict(pos=i, added=V.Vl.as_dict(), spectrum={q: s for q, (s, k) in tsvd_spectrum(st.ctx).items()}))
    alp = _mul_right_t(VL, U)
    arp = _mul_left(V, VR)
    for t in (VL, VR, B, inter, U, S, V):
        t.free()
    return alp, arp


def _expand(st, ALp, ARp):
    """_expand!(psi, AL', AR') (changebonds.jl:13-38): AL -> [AL AL'; 0 0], AR -> [AR 0; AR' 0], CR -> [C 0; 0 0], AC = AL * CR,
    normalize!(psi); direct sums along the bonds on the device (mpsk_tensor_cat).  ALp[i]: [l p; n_i] on the bond right of
    site i; ARp[i]: [n_{i-1} p; r] on the bond left of site i."""
    from .infinite import InfiniteMPS, mul_right
    ctx, N, dt = st.ctx, st.N, st.dtype
    AL, AR, CR, AC = [None] * N, [None] * N, [None] * N, [None] * N
    for i in range(N):
        nl = ARp[i].Vl                    # n_{i-1}: space added on the bond left of site i  (== _lastspace(AL'[i-1])')
        nr = ALp[i].Vr                    # n_i
        al = cat(st.AL[i], ALp[i], 1)
        lz = DeviceTensor.zeros(ctx, nl, al.mch, al.Vr, dt)
        AL[i] = cat(al, lz, 0)
        ar = cat(st.AR[i], ARp[i], 0)
        rz = DeviceTensor.zeros(ctx, ar.Vl, ar.mch, nr, dt)
        AR[i] = cat(ar, rz, 1)
        c = st.CR[i]
        cz = DeviceTensor.zeros(ctx, c.Vl, c.mch, nr, dt)
        c1 = cat(c, cz, 1)
        cz2 = DeviceTensor.zeros(ctx, nr, c.mch, c1.Vr, dt)
        CR[i] = cat(c1, cz2, 0)
        CR[i]._is_bond = True
        for t in (al, lz, ar, rz, cz, c1, cz2):
            t.free()
    for i in range(N):                    # normalize!(psi) (infinitemps.jl:252-256): CR and AC separately
        CR[i].scale(1.0 / CR[i].norm())
        AC[i] = mul_right(AL[i], CR[i])
        AC[i].scale(1.0 / AC[i].norm())
    V = [AL[i].Vl for i in range(N)]
    return InfiniteMPS(AL, AR, CR, AC, V, st.pspace, ctx)


def _inf_ddAC2(loc, st, H, envs, ac2):
    """ddAC2(loc, psi::InfiniteMPS, H, envs) (derivatives.jl:55-58)"""
    from .derivatives import MPO_ddAC2
    return MPO_ddAC2(st.ctx, H[loc], H[loc + 1], envs.leftenv(loc, st), envs.rightenv(loc + 1, st), ac2)


def changebonds_infinite(st, H=None, alg=None, envs=None, log=None):
    """changebonds(psi::InfiniteMPS, [H,] alg[, envs]) for OptimalExpand (optimalexpand.jl:16-38), RandExpand
    (randexpand.jl:15-35), SvdCut (svdcut.jl:36-46) and VUMPSSvdCut (vumpssvd.jl:19-86).  Returns (state, envs) when a
    Hamiltonian is given, else the state."""
    from . import infinite as di
    if alg is None:
        H, alg = None, H
    N = st.N
    from .operators import DenseMPO
    if isinstance(alg, OptimalExpand) and isinstance(H, DenseMPO):
        # changebonds(psi::InfiniteMPS, H::DenseMPO, alg::OptimalExpand, envs) (optimalexpand.jl:40-44): a one-row multiline
        from .statmech import environments as _statmech_environments
        envs = envs if envs is not None else _statmech_environments(st, H)
        rows, envs = changebonds_multiline([st], [[H[c] for c in range(N)]], alg, envs, log)
        return rows[0], envs
    if isinstance(alg, (OptimalExpand, RandExpand)):
        if isinstance(alg, OptimalExpand):
            envs = envs if envs is not None else di.environments(st, H)
        ALp, ARp = [None] * N, [None] * N
        for i in range(N):
            ac2 = contract_ac2(st, st.AC[i], st.AR[(i + 1) % N])
            if isinstance(alg, OptimalExpand):
                h = _inf_ddAC2(i, st, H, envs, ac2)
                AC2 = h(ac2); h.free(); ac2.free()
            else:                                                # randomize!(AC2)
                AC2 = DeviceTensor.random(st.ctx, ac2.Vl, ac2.mch, ac2.Vr, alg.seed + i, st.dtype); ac2.free()
            ALp[i], ARp[(i + 1) % N] = _expansion_vectors(st, i, AC2, alg.trscheme, log)
            AC2.free()
        new = _expand(st, ALp, ARp)
        return (new, envs) if H is not None else new
    if isinstance(alg, SvdCut):
        copied = list(st.AL)
        ncr = st.CR[0]
        P = _bond_leg(st)
        for i in range(N):
            U, S, V, eps = tsvd(P, st.CR[i], alg.trscheme, normalize=False)
            if log is not None:
                log.append(dict(pos=i, kept=S.Vl.as_dict(), eps_trunc=eps))
            ncr = S
            copied[i] = _mul_right_t(copied[i], U)
            copied[(i + 1) % N] = _mul_left(U.adjoint(), copied[(i + 1) % N])
            V.free()
        Vs = [copied[i].Vl for i in range(N)]
        new = di.InfiniteMPS.from_AL(copied, ncr, Vs, st.pspace, st.ctx)
        for i in range(N):                                       # normalize!
            new.CR[i].scale(1.0 / new.CR[i].norm()); new.AC[i].scale(1.0 / new.AC[i].norm())
        return (new, envs) if H is not None else new
    if isinstance(alg, VUMPSSvdCut):
        return _vumps_svdcut(st, H, alg, envs, log)
    raise TypeError(f"changebonds: unsupported algorithm {alg!r}")


def changebonds_multiline(rows, H=None, alg=None, envs=None, log=None):
    """changebonds(psi::MPSMultiline, [H::MPOMultiline,] alg[, envs]).  rows: list of InfiniteMPS (one per row, equal bond
    spaces); H[r][c]: MPO slices of row r, column c; envs: statmech.MultilineEnv (or PerMPOInfEnv for a single row).
    OptimalExpand (optimalexpand.jl:46-66): the two-site tensor of row i - 1 through ddAC2(i - 1, j) lives in the bond spaces
    of row i, whose null spaces (leftnull(AL[i, j]), rightnull(AR[i, j + 1])) select the expansion of row i; RandExpand
    (randexpand.jl:36-38) and SvdCut (svdcut.jl:32-34) are row-wise maps.  The rows keep equal bond spaces when the truncation
    scheme keeps the same sector dimensions in every row (truncdim on a trivial symmetry always does).
    Returns (rows, envs) when an operator is given, else the rows."""
    if alg is None:
        H, alg = None, H
    R, N = len(rows), rows[0].N
    if isinstance(alg, OptimalExpand):
        from .derivatives import MPO_ddAC2
        from .statmech import MultilineEnv, PerMPOInfEnv
        envs = envs if envs is not None else MultilineEnv(rows, H)
        if isinstance(envs, PerMPOInfEnv):                                  # one row: lw[c] / rw[c]
            envs.leftenv(0, rows[0])                                        # recalculates when the state changed
            lw, rw = [envs.lw], [envs.rw]
        else:
            if envs.dependency is not rows:
                envs.recalculate(rows)
            lw, rw = envs.lw, envs.rw
        out = []
        for i in range(R):
            r = (i - 1) % R
            top = rows[r]
            ALp, ARp = [None] * N, [None] * N
            for j in range(N):
                ac2 = contract_ac2(top, top.AC[j], top.AR[(j + 1) % N])
                h = MPO_ddAC2(top.ctx, H[r][j], H[r][(j + 1) % N], lw[r][j], rw[r][(j + 1) % N], ac2)
                AC2 = h(ac2); h.free(); ac2.free()
                ALp[j], ARp[(j + 1) % N] = _expansion_vectors(rows[i], j, AC2, alg.trscheme, log)
                AC2.free()
            out.append(_expand(rows[i], ALp, ARp))
        return out, envs
    if isinstance(alg, (RandExpand, SvdCut)):
        out = [changebonds_infinite(st, alg, log=log) for st in rows]
        return (out, envs) if H is not None else out
    raise TypeError(f"changebonds on MPSMultiline: unsupported algorithm {alg!r}")


def _isometry(ctx, Vrow, Vcol, dtype):
    """isometry(Vcol, Vrow)' as a bond tensor [Vrow; Vcol]: the identity on the leading min(n_row, n_col) indices of every
    common sector (vumpssvd.jl:72)"""
    M = np.zeros((Vrow.dim, 1, Vcol.dim), dtype=dtype)
    ro, co = Vrow.offsets(), Vcol.offsets()
    for jl, q in enumerate(Vrow.charges):
        if q in Vcol.charges:
            jr = Vcol.charges.index(q)
            k = min(Vrow.dims[jl], Vcol.dims[jr])
            M[ro[jl]:ro[jl] + k, 0, co[jr]:co[jr] + k] = np.eye(k)
    t = DeviceTensor.from_dense(ctx, M, Vrow, (0,), Vcol, dtype)
    t._is_bond = True
    return t


def _vumps_svdcut_n(st, H, alg, envs, log):
    """changebonds_n (vumpssvd.jl:49-86)"""
    from . import infinite as di
    from .fixedpoint import fixedpoint, Arnoldi
    from .finitemps import qrpos as _qrpos
    ctx, mod = st.ctx, st.pspace.modulus
    eig = Arnoldi(tol=alg.tol_eigenval, krylovdim=30, maxiter=100, hermitian=True)
    for loc in range(st.N):
        N = st.N
        if envs is None or envs.dependency is not st:      # leftenv(envs, pos, state) recalculates for a new state
            envs = di.environments(st, H)
        ac2 = contract_ac2(st, st.AC[loc], st.AR[(loc + 1) % N])
        h2 = _inf_ddAC2(loc, st, H, envs, ac2)
        _, nac2 = fixedpoint(h2, ac2, "SR", eig); h2.free(); ac2.free()
        hc = di._inf_ddC((loc + 1) % N, st, H, envs)
        _, nc2 = fixedpoint(hc, st.CR[(loc + 1) % N], "SR", eig); hc.free()
        AL1, S, Vt, eps = tsvd(st, nac2, alg.trscheme, normalize=False)
        if log is not None:
            log.append(dict(pos=loc, kept=S.Vl.as_dict(), eps_trunc=eps))
        AC = _mul_left(S, Vt)                               # S * V: [n p; r]
        QAC, r1 = _qrpos(ctx, AC, mod)
        QC, r2 = _qrpos(ctx, nc2, mod)
        if QAC.Vr == QC.Vr:
            AL2 = _mul_right_t(QAC, QC.adjoint())
        else:                                               # AL2 = QAC * dom_map' * QC'
            E = _isometry(ctx, QAC.Vr, QC.Vr, st.dtype)
            t = _mul_right_t(QAC, E)
            AL2 = _mul_right_t(t, QC.adjoint())
            E.free(); t.free()
        copied = list(st.AL)
        copied[loc] = AL1
        copied[(loc + 1) % N] = AL2
        Vs = [copied[i].Vl for i in range(N)]
        for t in (nac2, nc2, S, Vt, AC, QAC, r1, QC, r2):
            t.free()
        st = di.InfiniteMPS.from_tensors(copied, Vs, st.pspace, ctx, tol=alg.tol_gauge)
    return st, di.environments(st, H)


def _truncspace_cut(st, target):
    """changebonds(psi, SvdCut(trscheme = truncspace(target))) for InfiniteMPS (svdcut.jl:36-46 with TensorKit's truncspace):
    per sector the leading min(n, target_n) singular vectors; the restriction is a rectangular identity applied on the device"""
    from . import infinite as di
    N = st.N
    copied = list(st.AL)
    ncr = st.CR[0]
    P = _bond_leg(st)
    for i in range(N):
        U, S, V, eps = tsvd(P, st.CR[i], notrunc(), normalize=False)
        keep = infimum(S.Vl, target)
        E = _isometry(st.ctx, S.Vl, keep, st.dtype)        # [n; n']
        Uc = _mul_right_t(U, E)
        t = _mul_right_t(S, E)
        ncr = _mul_left(E.adjoint(), t)
        ncr._is_bond = True
        copied[i] = _mul_right_t(copied[i], Uc)
        copied[(i + 1) % N] = _mul_left(Uc.adjoint(), copied[(i + 1) % N])
        for x in (U, S, V, E, t):
            x.free()
    Vs = [copied[i].Vl for i in range(N)]
    new = di.InfiniteMPS.from_AL(copied, ncr, Vs, st.pspace, st.ctx)
    for i in range(N):
        new.CR[i].scale(1.0 / new.CR[i].norm()); new.AC[i].scale(1.0 / new.AC[i].norm())
    return new


def _vumps_svdcut(st, H, alg, envs, log):
    """changebonds(state, H, alg::VUMPSSvdCut, envs) (vumpssvd.jl:81-86); unit cells of one site are doubled, expanded and
    collapsed again (changebonds_1, vumpssvd.jl:19-47)"""
    from . import infinite as di
    if st.N > 1:
        return _vumps_svdcut_n(st, H, alg, envs, log)
    from .operators import MPOHamiltonian
    two = di.InfiniteMPS.from_tensors([st.AL[0], st.AL[0]], [st.V[0], st.V[0]], st.pspace, st.ctx)
    H2 = MPOHamiltonian([H[0], H[0]])
    two, _ = _vumps_svdcut_n(two, H2, alg, None, log)
    D1, D2 = two.AL[0].Vl, two.AL[1].Vl
    if D1 != D2:
        two = _truncspace_cut(two, infimum(D1, D2))
    one = di.InfiniteMPS.from_AL([two.AL[0]], two.CR[0], [two.AL[0].Vl], st.pspace, st.ctx, alg.tol_gauge)
    return one, di.environments(one, H)


def dot_infinite(a, b, krylovdim=30, tol=1e-12):
    """dot(psi1::InfiniteMPS, psi2::InfiniteMPS): dominant eigenvalue of the mixed transfer matrix of the AL tensors, by the
    device Arnoldi on the unit-cell transfer operator"""
    from .toolbox import transfer_spectrum
    return transfer_spectrum(b, below=a, num_vals=1, tol=tol, krylovdim=krylovdim)[0]


def _union(a, b):
    from .algorithms import UnionAlg
    return UnionAlg(a, b)


for _cls in (SvdCut, OptimalExpand, RandExpand, VUMPSSvdCut):          # alg1 & alg2 (unionalg.jl:11)
    _cls.__and__ = _union


# ------------------------------------------------------------------------------------------ DenseMPO products (state construction)
def _slice_dense(sl):
    """O[w, p_out, p_in, v] of a SparseMPOSlice (scalar entries are multiples of the identity, sparseslice.jl:40-55)"""
    out = np.zeros((sl.odim, sl.d, sl.d, sl.odim), dtype=np.complex128)
    for (i, j) in sl.keys():
        e = sl.Os[i][j]
        out[i, :, :, j] = e * np.eye(sl.d) if np.isscalar(e) else np.asarray(e)
    return out


def _fused(qa, qw, modulus):
    from .spaces import Space
    q = (np.asarray(qa)[:, None] + np.asarray(qw)[None, :]).reshape(-1)
    if modulus:
        q = q % modulus
    perm = np.argsort(q, kind="stable")
    cs, ns = np.unique(q, return_counts=True)
    return Space([int(c) for c in cs], [int(n) for n in ns], modulus), perm


def mpo_times_infinite(W, st):
    """Base.:*(mpo::DenseMPO, st::InfiniteMPS) (src/operators/densempo.jl:31-45): A'[(a w), p, (b v)] = sum_q AL[a, q, b]
    O[w, p, q, v], the fusers sorting the pair index into the sectors of charge q_a + c_w, then the InfiniteMPS constructor (device
    uniform gauge).  State construction, not the hot path: the product tensors (chi^2 D^2 d numbers) are formed once on the host
    from the downloaded AL and uploaded; the constructor's makefullrank! (abstractmps.jl:94-108) is not mirrored -- the fused bond
    sectors must be fillable by the neighbouring tensors (always the case without charges)."""
    from .infinite import InfiniteMPS
    N = st.N
    if len(W) % N and N % len(W):
        raise ValueError("dimension mismatch")                      # densempo.jl:32
    tensors, Vf = _product_tensors(W, [a.to_dense() for a in st.AL], st.V, st.pspace.modulus)
    A = [DeviceTensor.from_dense(st.ctx, tensors[i], Vf[i], list(st.pq), Vf[(i + 1) % N], st.dtype) for i in range(N)]
    return InfiniteMPS.from_tensors(A, Vf, st.pspace, st.ctx)


def _product_tensors(W, AL, V, modulus=0):
    """host part of mpo_times_infinite: dense (D chi, d, D chi) product tensors in the sector order of the fused bond spaces"""
    N = len(AL)
    fus = [_fused(V[i].index_charges(), W[i].domcharges, modulus) for i in range(N)]
    out = []
    for i in range(N):
        t = np.einsum('aqb,wpqv->awpbv', AL[i], _slice_dense(W[i]))
        Dl, chi, d, Dr, chir = t.shape
        t = t.reshape(Dl * chi, d, Dr * chir)[fus[i][1]][:, :, fus[(i + 1) % N][1]]
        out.append(np.ascontiguousarray(t))
    return out, [f[0] for f in fus]


def densempo_as_state(W, ctx=None):
    """convert(InfiniteMPS, mpo::DenseMPO) (densempo.jl:18-22): the two physical legs fused into one of charge p_out - p_in;
    returns the state and the permutation that sorted the fused physical index"""
    from .infinite import InfiniteMPS
    from .spaces import Space
    from ._lib import default_context
    ctx = ctx or default_context()
    N = len(W)
    pch = np.asarray(W[0].pcharges)
    d, mod = len(pch), W[0].modulus
    pq = (pch[:, None] - pch[None, :]).reshape(-1)
    if mod:
        pq = pq % mod
    order = np.argsort(pq, kind="stable")
    cs, ns = np.unique(pq, return_counts=True)
    pspace = Space([int(c) for c in cs], [int(n) for n in ns], mod)
    A, V = [], []
    for i in range(N):
        cl, cr = np.asarray(W[i].domcharges), np.asarray(W[i].imcharges)
        pl, pr = np.argsort(cl, kind="stable"), np.argsort(cr, kind="stable")
        t = _slice_dense(W[i]).reshape(len(cl), d * d, len(cr))[:, order, :][pl][:, :, pr]
        ul, nl = np.unique(cl, return_counts=True)
        ur, nr = np.unique(cr, return_counts=True)
        Vl, Vr = Space([int(c) for c in ul], [int(n) for n in nl], mod), Space([int(c) for c in ur], [int(n) for n in nr], mod)
        A.append(DeviceTensor.from_dense(ctx, np.ascontiguousarray(t), Vl, [int(c) for c in pq[order]], Vr, np.complex128))
        V.append(Vl)
    return InfiniteMPS.from_tensors(A, V, pspace, ctx), order


def svdcut_densempo(W, alg: SvdCut, ctx=None, log=None):
    """changebonds(mpo::DenseMPO, alg::SvdCut) (svdcut.jl:26-28): through the InfiniteMPS form, the device SvdCut, and back
    (densempo.jl:24-28: the AL tensors with the physical leg split again; every bond index becomes a level of width 1)."""
    from .operators import DenseMPO, SparseMPOSlice as _Slice
    st, order = densempo_as_state(W, ctx)
    cut = changebonds_infinite(st, alg, log=log)
    d = W[0].d
    inv = np.argsort(order)
    slices = []
    for i in range(cut.N):
        a = cut.AL[i].to_dense()[:, inv, :]
        O = a.reshape(a.shape[0], d, d, a.shape[2])
        Os = [[(O[w, :, :, v] if np.abs(O[w, :, :, v]).max() > 0 else 0) for v in range(O.shape[3])] for w in range(O.shape[0])]
        if O.shape[0] != O.shape[3]:
            raise ValueError("SvdCut of a DenseMPO left different bond dimensions in the unit cell: not a square slice")
        slices.append(_Slice(Os, tuple(cut.V[i].index_charges()), tuple(cut.V[(i + 1) % cut.N].index_charges()), W[0].pcharges, W[0].modulus))
    return DenseMPO(slices)
