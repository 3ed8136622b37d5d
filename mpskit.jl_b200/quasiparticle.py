"""Finite quasiparticle ansatz on the device: excitations(H, QuasiparticleAnsatz(), psi::FiniteMPS, envs; num)
(src/algorithms/excitation/quasiparticleexcitation.jl:130-170,234-362) with LeftGaugedQP states
(src/states/quasiparticle_state.jl:30-42,68,73-76: B_i = VL_i X_i, VL_i = leftnull(AL_i)) and the finite QP environments
(src/environments/qpenv.jl:145-166).  Trivial excitation sector, left_gs == right_gs.

One application of the effective excitation Hamiltonian is built from library calls only: the B-environments lBs / rBs are
chains of sparse-slice transfers with different ket and bra tensors (mpsk_transfer_left / _right: the excitation-leg transfers
of transfer.jl:48-87,113-126 reduce to these for a one-dimensional excitation leg), the three local terms are single-site
effective-Hamiltonian applies (mpsk_op_apply) with (GL, GR), (lB, GR) and (GL, rB), the projection X' = VL' B' is an MPO-free
transfer.  The Krylov vector is the tuple of X tensors (RecursiveVec); the Lanczos loop is host-driven over the device
VectorInterface like KrylovKit drives the reference."""
from __future__ import annotations
from dataclasses import dataclass
import numpy as np
from .tensor import DeviceTensor, block_table
from .derivatives import MPO_ddAC
from .environments import environments
from .krylov import RecursiveVec, eigsolve_lanczos
from .changebonds import leftnull
from .infinite import transfer_left, transfer_right, mul_right
from .algorithms import _overlap


@dataclass
class QuasiparticleAnsatz:
    """quasiparticleexcitation.jl:20-30: the eigensolver parameters (Defaults.eigsolver)"""
    tol: float = 1e-10
    krylovdim: int = 30
    maxiter: int = 100


class FiniteQP:
    """LeftGaugedQP(rand, psi) for a FiniteMPS (quasiparticle_state.jl:30-42): VLs[i] = leftnull(AL[i]) on the device; sites
    whose AL is square (no null space) carry no X tensor"""

    def __init__(self, gs, seed=0):
        self.gs = gs
        L = len(gs)
        self.VLs, self.active = [], []
        for i in range(L):
            N = leftnull(gs.AL[i])
            if N.Vr.dim == 0 or not block_table(N.Vr, (0,), gs.AR[i].Vr):    # no null space / no X[k'; r] block of trivial charge
                N.free(); N = None
            self.VLs.append(N)
            if N is not None:
                self.active.append(i)
        self.seed = seed

    def random(self, seed=None):
        seed = self.seed if seed is None else seed
        gs = self.gs
        return RecursiveVec([DeviceTensor.random(gs.ctx, self.VLs[i].Vr, (0,), gs.AR[i].Vr, seed + 17 * i, gs.dtype)
                             for i in self.active])

    def B(self, X, k):
        """exci[i] = VLs[i] * Xs[i] (quasiparticle_state.jl:68); k indexes the active sites"""
        return mul_right(self.VLs[self.active[k]], X.vecs[k])


def renormalization_energy(H, gs, envs):
    """effective_excitation_renormalization_energy (quasiparticleexcitation.jl:336-349): <AC| dAC |AC> per site"""
    out = []
    for loc in range(len(gs)):
        ac = gs.AC[loc]
        h = MPO_ddAC(gs.ctx, H[loc], envs.leftenv(loc, gs), envs.rightenv(loc, gs), ac)
        y = h(ac); h.free()
        out.append(complex(ac.dot(y))); y.free()
    return out


class _HEff:
    """phi -> effective_excitation_hamiltonian(H, phi, environments(phi, H, lenvs, renvs), E)"""

    def __init__(self, H, qp: FiniteQP, envs, E):
        self.H, self.qp, self.envs, self.E = H, qp, envs, E
        gs = qp.gs
        self.L = len(gs)
        self.GL = [envs.leftenv(i, gs) for i in range(self.L)]
        self.GR = [envs.rightenv(i, gs) for i in range(self.L)]
        self.AL = [gs.AL[i] for i in range(self.L)]
        self.AR = [gs.AR[i] for i in range(self.L)]
        self.center = {}                                   # persistent (GL, GR) operators of the active sites
        for i in qp.active:
            self.center[i] = MPO_ddAC(gs.ctx, H[i], self.GL[i], self.GR[i], self.AL[i])

    def __call__(self, X):
        H, qp, L = self.H, self.qp, self.L
        gs = qp.gs
        ctx = gs.ctx
        Bs = {i: qp.B(X, k) for k, i in enumerate(qp.active)}
        # ---- environments(exci::FiniteQP, H, lenvs, renvs) (qpenv.jl:154-164); None stands for a zero environment
        lB = [None] * L
        for pos in range(L - 1):
            acc = None
            if lB[pos] is not None:
                acc = transfer_left(lB[pos], H[pos], self.AR[pos], self.AL[pos])
            if pos in Bs:
                t = transfer_left(self.GL[pos], H[pos], Bs[pos], self.AL[pos])
                if acc is None:
                    acc = t
                else:
                    acc.axpy(1.0, t); t.free()
            lB[pos + 1] = acc
        rB = [None] * L
        for pos in range(L - 1, 0, -1):
            acc = None
            if rB[pos] is not None:
                acc = transfer_right(rB[pos], H[pos], self.AL[pos], self.AR[pos])
            if pos in Bs:
                t = transfer_right(self.GR[pos], H[pos], Bs[pos], self.AR[pos])
                if acc is None:
                    acc = t
                else:
                    acc.axpy(1.0, t); t.free()
            rB[pos - 1] = acc
        # ---- _effective_excitation_local_apply (quasiparticleexcitation.jl:295-334)
        out = []
        for k, loc in enumerate(qp.active):
            B = Bs[loc]
            Bp = self.center[loc](B)
            Bp.axpy(-self.E[loc], B)
            if loc > 0 and lB[loc] is not None:
                h = MPO_ddAC(ctx, H[loc], lB[loc], self.GR[loc], self.AR[loc])
                t = h(self.AR[loc]); h.free()
                Bp.axpy(1.0, t); t.free()
            if loc < L - 1 and rB[loc] is not None:
                h = MPO_ddAC(ctx, H[loc], self.GL[loc], rB[loc], self.AL[loc])
                t = h(self.AL[loc]); h.free()
                Bp.axpy(1.0, t); t.free()
            out.append(_overlap(gs, Bp, qp.VLs[loc]))                       # setindex!: X = VL' * B (quasiparticle_state.jl:73-76)
            Bp.free()
        for t in list(Bs.values()) + [e for e in lB + rB if e is not None]:
            t.free()
        return RecursiveVec(out)

    def free(self):
        for o in self.center.values():
            o.free()
        self.center = {}


def excitations_qp(H, alg: QuasiparticleAnsatz, gs, envs=None, num=1, seed=0, weight=100.0, log=None):
    """excitations(H, alg::QuasiparticleAnsatz, psi::FiniteMPS, envs; num) (quasiparticleexcitation.jl:130-170): energies of the
    `num` lowest quasiparticle states relative to the ground state and their X tuples.  Several states are found one after the
    other with the converged ones shifted up by `weight` (the effective Hamiltonian is Hermitian in the X parametrisation)."""
    alg = alg or QuasiparticleAnsatz()
    envs = envs if envs is not None else environments(gs, H)
    qp = FiniteQP(gs, seed)
    E = renormalization_energy(H, gs, envs)
    heff = _HEff(H, qp, envs, E)
    found, energies = [], []

    def f(v):
        y = heff(v)
        for u in found:
            y.axpy(weight * complex(u.dot(v)), u)
        return y
    for n in range(num):
        x0 = qp.random(seed + 1000 * n)
        val, vec, info = eigsolve_lanczos(f, x0, "SR", alg.tol, alg.krylovdim, alg.maxiter)
        if info["converged"] == 0:
            import warnings
            warnings.warn(f"excitation failed to converge: normres = {info['normres']}")     # quasiparticleexcitation.jl:139-140
        if log is not None:
            log.append(dict(state=n, energy=float(np.real(val)), **info))
        energies.append(float(np.real(val))); found.append(vec)
    heff.free()
    return energies, found, qp


# ------------------------------------------------------------------------------------------------------------------------
# Infinite quasiparticle ansatz (trivial excitation sector, left_gs == right_gs)
class InfiniteQP:
    """LeftGaugedQP(rand, psi::InfiniteMPS; momentum) (quasiparticle_state.jl:30-42): VLs[i] = leftnull(AL[i]) on the device;
    the momentum is per site, as in the reference (qpenv.jl:68-71)"""

    def __init__(self, gs, momentum=0.0, seed=0):
        self.gs, self.momentum, self.seed = gs, float(momentum), seed
        self.VLs = [leftnull(gs.AL[i]) for i in range(gs.N)]
        self.active = list(range(gs.N))

    def random(self, seed=None):
        seed = self.seed if seed is None else seed
        gs = self.gs
        return RecursiveVec([DeviceTensor.random(gs.ctx, self.VLs[i].Vr, (0,), gs.AR[i].Vr, seed + 17 * i, gs.dtype)
                             for i in range(gs.N)])

    def B(self, X, k):
        return mul_right(self.VLs[k], X.vecs[k])


def _replace_level(t, i, new):
    """stacked level tensor with level i replaced"""
    nlev = len(t.mch)
    parts = [new if j == i else t.select([j]) for j in range(nlev)]
    out = DeviceTensor.stack(parts)
    for j, p in enumerate(parts):
        if j != i:
            p.free()
    return out


def _remove_component(level, Cm):
    """level -= <C, level> C : the regularisation of the identity levels (qpenv.jl:73-82,90-99; exci_transfer_system.jl:15-21)"""
    level.axpy(-complex(Cm.dot(level)), Cm)
    return level


class _InfHEff:
    """phi -> effective_excitation_hamiltonian(H, phi, environments(phi, H, lenvs, renvs), E) for an InfiniteQP
    (quasiparticleexcitation.jl:234-256,295-334; qpenv.jl:56-143; exci_transfer_system.jl:1-85)"""

    def __init__(self, H, qp: InfiniteQP, envs, E, solver_tol=1e-12, solver_maxiter=100):
        from .infinite import TransferMatrix, regularize, _entry_slice, _isid
        self.H, self.qp, self.envs, self.E = H, qp, envs, E
        self.tol, self.maxiter = solver_tol, solver_maxiter
        gs = qp.gs
        N, od = gs.N, H.odim
        self.N, self.od = N, od
        self.GL = [envs.leftenv(i, gs) for i in range(N)]
        self.GR = [envs.rightenv(i, gs) for i in range(N)]
        self.AL, self.AR, self.CR = list(gs.AL), list(gs.AR), list(gs.CR)
        self.isid = [_isid(H, i) for i in range(od)]
        self.ids = [i for i in range(1, od - 1) if self.isid[i]]
        self.center = [MPO_ddAC(gs.ctx, H[i], self.GL[i], self.GR[i], self.AL[i]) for i in range(N)]
        Cn = self.CR[N - 1]
        Cadj = Cn.adjoint()
        # diagonal transfer operators of the two linear systems (persistent device operators)
        self.tl, self.tr = {}, {}
        for i in range(od):
            if not all(sl.contains(i, i) for sl in H.slices):
                continue
            mids = None if self.isid[i] else [_entry_slice(H[s], i, i) for s in range(N)]
            tl = TransferMatrix(self.AR, mids, self.AL).flip()
            tr = TransferMatrix(self.AL, mids, self.AR)
            if self.isid[i]:
                tl = regularize(tl, Cadj, Cn)
                tr = regularize(tr, Cadj, Cn)
            self.tl[i], self.tr[i] = tl, tr

    # ---- exci_transfer_system.jl
    def _left_system(self, lB0):
        from .fixedpoint import linsolve
        H, N, od, p = self.H, self.N, self.od, self.qp.momentum
        Cn = self.CR[N - 1]
        found = lB0.zerovector()
        for i in range(od):
            t = found.copy()
            for s in range(N):
                nt = transfer_left(t, H[s], self.AR[s], self.AL[s]); t.free(); t = nt
            start = t.select([i]); t.free()
            start.scale(np.exp(-1j * p * N))
            if self.isid[i]:
                _remove_component(start, Cn)
            b0 = lB0.select([i])
            start.axpy(1.0, b0); b0.free()
            if i in self.tl:
                sol = linsolve(self.tl[i], start, start, 1.0, -np.exp(-1j * p * N), self.tol, 30, self.maxiter)
                start.free(); start = sol
            nf = _replace_level(found, i, start)
            found.free(); start.free(); found = nf
        return found

    def _right_system(self, rB0):
        from .fixedpoint import linsolve
        H, N, od, p = self.H, self.N, self.od, self.qp.momentum
        Cn = self.CR[N - 1]
        found = rB0.zerovector()
        for i in range(od - 1, -1, -1):
            t = found.copy()
            for s in range(N - 1, -1, -1):
                nt = transfer_right(t, H[s], self.AL[s], self.AR[s]); t.free(); t = nt
            start = t.select([i]); t.free()
            start.scale(np.exp(1j * p * N))
            if self.isid[i]:
                _remove_component(start, Cn)
            b0 = rB0.select([i])
            start.axpy(1.0, b0); b0.free()
            if i in self.tr:
                sol = linsolve(self.tr[i], start, start, 1.0, -np.exp(1j * p * N), self.tol, 30, self.maxiter)
                start.free(); start = sol
            nf = _replace_level(found, i, start)
            found.free(); start.free(); found = nf
        return found

    def _regularise_ids(self, t, Cm):
        for k in self.ids:
            lev = _remove_component(t.select([k]), Cm)
            nt = _replace_level(t, k, lev)
            t.free(); lev.free(); t = nt
        return t

    def __call__(self, X):
        H, qp, N = self.H, self.qp, self.N
        gs = qp.gs
        ctx = gs.ctx
        ph = np.exp(1j * qp.momentum)
        Bs = [qp.B(X, k) for k in range(N)]
        # ---- environments(exci::InfiniteQP, H, lenvs, renvs) (qpenv.jl:68-137)
        lB = [self.GL[i].zerovector() for i in range(N)]
        for pos in range(N):
            new = transfer_left(lB[pos], H[pos], self.AR[pos], self.AL[pos])
            t = transfer_left(self.GL[pos], H[pos], Bs[pos], self.AL[pos])
            new.axpy(1.0, t); t.free()
            new.scale(1.0 / ph)
            new = self._regularise_ids(new, self.CR[pos])
            lB[(pos + 1) % N].free()
            lB[(pos + 1) % N] = new
        rB = [self.GR[i].zerovector() for i in range(N)]
        for pos in range(N - 1, -1, -1):
            new = transfer_right(rB[pos], H[pos], self.AL[pos], self.AR[pos])
            t = transfer_right(self.GR[pos], H[pos], Bs[pos], self.AR[pos])
            new.axpy(1.0, t); t.free()
            new.scale(ph)
            new = self._regularise_ids(new, self.CR[(pos - 1) % N])
            rB[(pos - 1) % N].free()
            rB[(pos - 1) % N] = new
        s0 = self._left_system(lB[0]); lB[0].free(); lB[0] = s0
        s1 = self._right_system(rB[N - 1]); rB[N - 1].free(); rB[N - 1] = s1
        cur = lB[0].copy()
        for i in range(N - 1):
            nt = transfer_left(cur, H[i], self.AR[i], self.AL[i]); cur.free(); cur = nt
            cur.scale(1.0 / ph)
            cur = self._regularise_ids(cur, self.CR[i])
            lB[i + 1].axpy(1.0, cur)
        cur.free()
        cur = rB[N - 1].copy()
        for i in range(N - 1, 0, -1):
            nt = transfer_right(cur, H[i], self.AL[i], self.AR[i]); cur.free(); cur = nt
            cur.scale(ph)
            cur = self._regularise_ids(cur, self.CR[i - 1])
            rB[i - 1].axpy(1.0, cur)
        cur.free()
        # ---- _effective_excitation_local_apply (quasiparticleexcitation.jl:295-334): all three terms on every site
        out = []
        for loc in range(N):
            B = Bs[loc]
            Bp = self.center[loc](B)
            Bp.axpy(-self.E[loc], B)
            h = MPO_ddAC(ctx, H[loc], lB[loc], self.GR[loc], self.AR[loc])
            t = h(self.AR[loc]); h.free()
            Bp.axpy(1.0, t); t.free()
            h = MPO_ddAC(ctx, H[loc], self.GL[loc], rB[loc], self.AL[loc])
            t = h(self.AL[loc]); h.free()
            Bp.axpy(1.0, t); t.free()
            out.append(_overlap(gs, Bp, qp.VLs[loc]))
            Bp.free()
        for t in Bs + lB + rB:
            t.free()
        return RecursiveVec(out)

    def free(self):
        for o in self.center:
            o.free()
        self.center = []


def excitations_qp_infinite(H, alg: QuasiparticleAnsatz, momentum, gs, envs=None, num=1, seed=0, weight=100.0, log=None):
    """excitations(H, alg::QuasiparticleAnsatz, momentum::Number, psi::InfiniteMPS, envs; num)
    (quasiparticleexcitation.jl:39-92): energies of the `num` lowest quasiparticle states of (per-site) momentum `momentum`."""
    from . import infinite as di
    alg = alg or QuasiparticleAnsatz()
    envs = envs if envs is not None else di.environments(gs, H)
    qp = InfiniteQP(gs, momentum, seed)
    E = []
    for loc in range(gs.N):                              # effective_excitation_renormalization_energy (…:336-349)
        ac = gs.AC[loc]
        h = MPO_ddAC(gs.ctx, H[loc], envs.leftenv(loc, gs), envs.rightenv(loc, gs), ac)
        y = h(ac); h.free()
        E.append(complex(ac.dot(y))); y.free()
    heff = _InfHEff(H, qp, envs, E)
    found, energies = [], []

    def f(v):
        y = heff(v)
        for u in found:
            y.axpy(weight * complex(u.dot(v)), u)
        return y
    for n in range(num):
        x0 = qp.random(seed + 1000 * n)
        val, vec, info = eigsolve_lanczos(f, x0, "SR", alg.tol, alg.krylovdim, alg.maxiter)
        if info["converged"] == 0:
            import warnings
            warnings.warn(f"excitation failed to converge: normres = {info['normres']}")
        if log is not None:
            log.append(dict(state=n, energy=float(np.real(val)), **info))
        energies.append(float(np.real(val))); found.append(vec)
    heff.free()
    return energies, found, qp


# ------------------------------------------------------------------------------------------------------------------------
def fidelity_susceptibility(gs, H0, Vs, henvs=None, tol=1e-12, maxiter=100):
    """fidelity_susceptibility(state, H0, Vs, henvs; maxiter, tol) (src/algorithms/fidelity_susceptibility.jl:4-29) for a
    FiniteMPS or an InfiniteMPS: for every perturbation V the tangent vector with B_i = dAC_V(i) * AC_i (projected on the left
    null spaces, LeftGaugedQP) is the right-hand side of  H_eff x = Tos  with the effective excitation Hamiltonian of H0 at zero
    momentum, solved by GMRES over the tuple of X tensors (x0 = Tos, as there); the result is the matrix dot(x_a, x_b)."""
    from .krylov import linsolve_gmres
    from . import infinite as di
    infinite = hasattr(gs, "N") and not hasattr(gs, "L")
    ctx = gs.ctx
    if infinite:
        henvs = henvs if henvs is not None else di.environments(gs, H0)
        qp = InfiniteQP(gs, 0.0)
        E = []
        for loc in range(gs.N):
            ac = gs.AC[loc]
            h = MPO_ddAC(ctx, H0[loc], henvs.leftenv(loc, gs), henvs.rightenv(loc, gs), ac)
            y = h(ac); h.free()
            E.append(complex(ac.dot(y))); y.free()
        heff = _InfHEff(H0, qp, henvs, E, solver_tol=min(tol, 1e-12))
    else:
        henvs = henvs if henvs is not None else environments(gs, H0)
        qp = FiniteQP(gs)
        heff = _HEff(H0, qp, henvs, renormalization_energy(H0, gs, henvs))
    vecs = []
    for V in Vs:
        venvs = di.environments(gs, V) if infinite else environments(gs, V)
        Xs = []
        for i in qp.active:
            ac = gs.AC[i]
            h = MPO_ddAC(ctx, V[i], venvs.leftenv(i, gs), venvs.rightenv(i, gs), ac)
            temp = h(ac); h.free()
            Xs.append(_overlap(gs, temp, qp.VLs[i]))               # Tos[i] = temp (x) utility leg -> X_i = VL_i' * temp
            temp.free()
        b = RecursiveVec(Xs)
        x, info = linsolve_gmres(heff, b, b, 0.0, 1.0, tol, 30, maxiter)
        if info["converged"] == 0:
            import warnings
            warnings.warn(f"failed to converge: normres = {info['normres']}")      # fidelity_susceptibility.jl:21
        vecs.append(x)
    heff.free()
    return np.array([[complex(a.dot(b)) for b in vecs] for a in vecs])
