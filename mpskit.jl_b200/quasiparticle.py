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
