"""TEST INFRASTRUCTURE (CPU oracle, never imported by the product path).

Finite quasiparticle ansatz of the reference restated on dense charged tensors:
  * LeftGaugedQP                         src/states/quasiparticle_state.jl:30-42,68,73-76   (B_i = VL_i X_i)
  * environments(exci::FiniteQP, H)      src/environments/qpenv.jl:145-166                  (lBs / rBs)
  * effective_excitation_hamiltonian     src/algorithms/excitation/quasiparticleexcitation.jl:234-334
  * excitations(H, QuasiparticleAnsatz(), psi::FiniteMPS, envs; num)                        :130-170
Trivial excitation sector, left_gs == right_gs.  The Krylov vector is the tuple of X tensors, flattened."""
from __future__ import annotations
import numpy as np
from .finite import FiniteMPS, FinEnv, _null_blocks
from .contract import dAC_sparse, transfer_left_sparse, transfer_right_sparse
from .krylov import eigsolve_lanczos


class FiniteQP:
    def __init__(self, gs: FiniteMPS):
        self.gs = gs
        L, pq = gs.L, gs.pq
        d = len(pq)
        self.VLs, self.masks = [], []
        for i in range(L):
            AL = gs.get_AL(i)
            Vl, Vr = gs.V[i], gs.V[i + 1]
            rowq = (Vl.index_charges()[:, None] + pq[None, :]).reshape(-1)
            N, nq = _null_blocks(AL.reshape(Vl.dim * d, Vr.dim), rowq, Vr.index_charges())       # leftnull(AL[i])
            self.VLs.append(N.reshape(Vl.dim, d, N.shape[1]))
            diff = nq[:, None] - Vr.index_charges()[None, :]
            self.masks.append((diff % gs.p.modulus == 0) if gs.p.modulus else (diff == 0))       # X[k'; r]: trivial sector
        self.sizes = [int(m.sum()) for m in self.masks]

    def unpack(self, vec):
        Xs, off = [], 0
        for m, n in zip(self.masks, self.sizes):
            X = np.zeros(m.shape, dtype=np.complex128)
            X[m] = vec[off:off + n]
            Xs.append(X); off += n
        return Xs

    def pack(self, Xs):
        return np.concatenate([X[m] for X, m in zip(Xs, self.masks)]) if Xs else np.zeros(0, dtype=np.complex128)

    def B(self, Xs, i):
        return np.einsum('apk,kr->apr', self.VLs[i], Xs[i])


def renormalization_energy(H, gs, envs):
    """effective_excitation_renormalization_energy (quasiparticleexcitation.jl:336-362): <AC| dAC |AC> per site"""
    out = []
    for loc in range(gs.L):
        AC = gs.get_AC(loc)
        out.append(np.vdot(AC, dAC_sparse(AC, H[loc], envs.leftenv(loc, gs), envs.rightenv(loc, gs))))
    return out


def effective_excitation_hamiltonian(H, qp: FiniteQP, Xs, envs, E):
    gs = qp.gs
    L = gs.L
    Bs = [qp.B(Xs, i) for i in range(L)]
    GL = [envs.leftenv(i, gs) for i in range(L)]
    GR = [envs.rightenv(i, gs) for i in range(L)]
    lB = [None] * L
    rB = [None] * L
    lB[0] = np.zeros_like(GL[0])
    for pos in range(L - 1):                                        # qpenv.jl:154-158
        lB[pos + 1] = transfer_left_sparse(lB[pos], H[pos], gs.get_AR(pos), gs.get_AL(pos)) + \
            transfer_left_sparse(GL[pos], H[pos], Bs[pos], gs.get_AL(pos))
    rB[L - 1] = np.zeros_like(GR[L - 1])
    for pos in range(L - 1, 0, -1):                                 # qpenv.jl:160-164
        rB[pos - 1] = transfer_right_sparse(rB[pos], H[pos], gs.get_AL(pos), gs.get_AR(pos)) + \
            transfer_right_sparse(GR[pos], H[pos], Bs[pos], gs.get_AR(pos))
    out = []
    for loc in range(L):                                            # quasiparticleexcitation.jl:295-334
        Bp = -E[loc] * Bs[loc] + dAC_sparse(Bs[loc], H[loc], GL[loc], GR[loc])
        if loc > 0:
            Bp = Bp + dAC_sparse(gs.get_AR(loc), H[loc], lB[loc], GR[loc])
        if loc < L - 1:
            Bp = Bp + dAC_sparse(gs.get_AL(loc), H[loc], GL[loc], rB[loc])
        out.append(np.einsum('apk,apr->kr', np.conj(qp.VLs[loc]), Bp))          # setindex!: X = VL' * B
    return out


def excitations(H, gs: FiniteMPS, envs=None, num=1, tol=1e-10, krylovdim=30, maxiter=100, seed=0, weight=100.0):
    """excitations(H, QuasiparticleAnsatz(), psi::FiniteMPS, envs; num): the `num` lowest eigenvalues of the effective
    excitation Hamiltonian (energies relative to the ground state) and the X tuples; several states by deflation"""
    envs = envs or FinEnv(gs, H)
    qp = FiniteQP(gs)
    E = renormalization_energy(H, gs, envs)
    rng = np.random.default_rng(seed)
    n = sum(qp.sizes)
    found, energies = [], []

    def f(v):
        y = qp.pack(effective_excitation_hamiltonian(H, qp, qp.unpack(v), envs, E))
        for u in found:
            y = y + weight * np.vdot(u, v) * u
        return y
    for _ in range(num):
        x0 = rng.standard_normal(n) + 1j * rng.standard_normal(n)
        val, vec, info = eigsolve_lanczos(f, x0, "SR", tol, krylovdim, maxiter)
        energies.append(float(np.real(val))); found.append(vec)
    return energies, [qp.unpack(v) for v in found], qp


# ------------------------------------------------------------------------------------------------------------------------
# Infinite quasiparticle ansatz (trivial excitation sector, left_gs == right_gs):
#   environments(exci::InfiniteQP, H::MPOHamiltonian, lenvs, renvs)   src/environments/qpenv.jl:56-143
#   left / right_excitation_transfer_system                            src/algorithms/excitation/exci_transfer_system.jl:1-85
#   excitations(H, QuasiparticleAnsatz(), momentum, psi, envs)          quasiparticleexcitation.jl:39-92
class InfiniteQP:
    def __init__(self, gs, momentum=0.0):
        self.gs, self.momentum = gs, float(momentum)
        pq, mod = gs.pq, gs.modulus
        d = len(pq)
        self.VLs, self.masks = [], []
        for i in range(gs.N):
            AL = gs.AL[i]
            Vl, Vr = gs.V[i], gs.V[(i + 1) % gs.N]
            rowq = (Vl.index_charges()[:, None] + pq[None, :]).reshape(-1)
            N, nq = _null_blocks(AL.reshape(Vl.dim * d, Vr.dim), rowq, Vr.index_charges())
            self.VLs.append(N.reshape(Vl.dim, d, N.shape[1]))
            diff = nq[:, None] - Vr.index_charges()[None, :]
            self.masks.append((diff % mod == 0) if mod else (diff == 0))
        self.sizes = [int(m.sum()) for m in self.masks]

    unpack = FiniteQP.unpack
    pack = FiniteQP.pack
    B = FiniteQP.B


def _isid(H, i):
    return all(np.isscalar(s.Os[i][i]) and abs(s.Os[i][i] - 1) < 1e-14 for s in H.slices)


def _left_transfer_system(lB0, H, gs, p, tol=1e-12):
    """left_excitation_transfer_system (exci_transfer_system.jl:1-42)"""
    from .krylov import linsolve_gmres
    N, od = gs.N, H.odim
    Cn = gs.CR[N - 1]
    found = np.zeros_like(lB0)
    for i in range(od):
        t = found
        for s in range(N):
            t = transfer_left_sparse(t, H[s], gs.AR[s], gs.AL[s])
        start = t[:, i, :] * np.exp(-1j * p * N)
        idl = _isid(H, i)
        if idl:
            start = start - np.vdot(Cn, start) * Cn
        cur = start + lB0[:, i, :]
        if all(sl.contains(i, i) for sl in H.slices):
            def tm(v):
                for s in range(N):
                    e = H[s].Os[i][i]
                    if np.isscalar(e):
                        v = e * np.einsum('ab,bpc,apd->dc', v, gs.AR[s], np.conj(gs.AL[s]))
                    else:
                        v = np.einsum('ab,bqc,pq,apd->dc', v, gs.AR[s], np.asarray(e), np.conj(gs.AL[s]))
                if idl:
                    v = v - np.vdot(Cn, v) * Cn
                return v
            cur, info = linsolve_gmres(tm, cur, cur, 1.0, -np.exp(-1j * p * N), tol, 30, 100)
        found = found.copy()
        found[:, i, :] = cur
    return found


def _right_transfer_system(rB0, H, gs, p, tol=1e-12):
    """right_excitation_transfer_system (exci_transfer_system.jl:44-85)"""
    from .krylov import linsolve_gmres
    N, od = gs.N, H.odim
    Cn = gs.CR[N - 1]
    found = np.zeros_like(rB0)
    for i in range(od - 1, -1, -1):
        t = found
        for s in range(N - 1, -1, -1):
            t = transfer_right_sparse(t, H[s], gs.AL[s], gs.AR[s])
        start = t[:, i, :] * np.exp(1j * p * N)
        idl = _isid(H, i)
        if idl:
            start = start - np.vdot(Cn, start) * Cn
        cur = start + rB0[:, i, :]
        if all(sl.contains(i, i) for sl in H.slices):
            def tm(v):
                for s in range(N - 1, -1, -1):
                    e = H[s].Os[i][i]
                    if np.isscalar(e):
                        v = e * np.einsum('apb,bc,dpc->ad', gs.AL[s], v, np.conj(gs.AR[s]))
                    else:
                        v = np.einsum('aqb,pq,bc,dpc->ad', gs.AL[s], np.asarray(e), v, np.conj(gs.AR[s]))
                if idl:
                    v = v - np.vdot(Cn, v) * Cn
                return v
            cur, info = linsolve_gmres(tm, cur, cur, 1.0, -np.exp(1j * p * N), tol, 30, 100)
        found = found.copy()
        found[:, i, :] = cur
    return found


def qp_environments_infinite(H, qp: InfiniteQP, Bs, envs, tol=1e-12):
    """environments(exci::InfiniteQP, H, lenvs, renvs) (qpenv.jl:56-143): lB[pos] / rB[pos] as stacked level tensors"""
    gs, p = qp.gs, qp.momentum
    N, od = gs.N, H.odim
    ids = [i for i in range(1, od - 1) if _isid(H, i)]
    GL = [envs.leftenv(i, gs) for i in range(N)]
    GR = [envs.rightenv(i, gs) for i in range(N)]
    ph = np.exp(1j * p)
    lB = [np.zeros_like(GL[i]) for i in range(N)]
    for pos in range(N):
        new = (transfer_left_sparse(lB[pos], H[pos], gs.AR[pos], gs.AL[pos]) +
               transfer_left_sparse(GL[pos], H[pos], Bs[pos], gs.AL[pos])) / ph
        for i in ids:
            new[:, i, :] -= np.vdot(gs.CR[pos], new[:, i, :]) * gs.CR[pos]
        lB[(pos + 1) % N] = new
    rB = [np.zeros_like(GR[i]) for i in range(N)]
    for pos in range(N - 1, -1, -1):
        new = (transfer_right_sparse(rB[pos], H[pos], gs.AL[pos], gs.AR[pos]) +
               transfer_right_sparse(GR[pos], H[pos], Bs[pos], gs.AR[pos])) * ph
        for i in ids:
            new[:, i, :] -= np.vdot(gs.CR[(pos - 1) % N], new[:, i, :]) * gs.CR[(pos - 1) % N]
        rB[(pos - 1) % N] = new
    lB[0] = _left_transfer_system(lB[0], H, gs, p, tol)
    rB[N - 1] = _right_transfer_system(rB[N - 1], H, gs, p, tol)
    cur = lB[0]
    for i in range(N - 1):
        cur = transfer_left_sparse(cur, H[i], gs.AR[i], gs.AL[i]) / ph
        for k in ids:
            cur[:, k, :] -= np.vdot(gs.CR[i], cur[:, k, :]) * gs.CR[i]
        lB[i + 1] = lB[i + 1] + cur
    cur = rB[N - 1]
    for i in range(N - 1, 0, -1):
        cur = transfer_right_sparse(cur, H[i], gs.AL[i], gs.AR[i]) * ph
        for k in ids:
            cur[:, k, :] -= np.vdot(gs.CR[i - 1], cur[:, k, :]) * gs.CR[i - 1]
        rB[i - 1] = rB[i - 1] + cur
    return lB, rB, GL, GR


def effective_excitation_hamiltonian_infinite(H, qp: InfiniteQP, Xs, envs, E, tol=1e-12):
    gs = qp.gs
    N = gs.N
    Bs = [qp.B(Xs, i) for i in range(N)]
    lB, rB, GL, GR = qp_environments_infinite(H, qp, Bs, envs, tol)
    out = []
    for loc in range(N):
        Bp = -E[loc] * Bs[loc] + dAC_sparse(Bs[loc], H[loc], GL[loc], GR[loc])
        Bp = Bp + dAC_sparse(gs.AR[loc], H[loc], lB[loc], GR[loc])
        Bp = Bp + dAC_sparse(gs.AL[loc], H[loc], GL[loc], rB[loc])
        out.append(np.einsum('apk,apr->kr', np.conj(qp.VLs[loc]), Bp))
    return out


def excitations_infinite(H, momentum, gs, envs, num=1, tol=1e-10, krylovdim=30, maxiter=100, seed=0, weight=100.0, envtol=1e-12):
    """excitations(H, QuasiparticleAnsatz(), momentum, psi::InfiniteMPS, envs; num) (quasiparticleexcitation.jl:39-92)"""
    qp = InfiniteQP(gs, momentum)
    E = [np.vdot(gs.AC[loc], dAC_sparse(gs.AC[loc], H[loc], envs.leftenv(loc, gs), envs.rightenv(loc, gs))) for loc in range(gs.N)]
    rng = np.random.default_rng(seed)
    n = sum(qp.sizes)
    found, energies = [], []

    def f(v):
        y = qp.pack(effective_excitation_hamiltonian_infinite(H, qp, qp.unpack(v), envs, E, envtol))
        for u in found:
            y = y + weight * np.vdot(u, v) * u
        return y
    for _ in range(num):
        x0 = rng.standard_normal(n) + 1j * rng.standard_normal(n)
        val, vec, info = eigsolve_lanczos(f, x0, "SR", tol, krylovdim, maxiter)
        energies.append(float(np.real(val))); found.append(vec)
    return energies, [qp.unpack(v) for v in found], qp


# ------------------------------------------------------------------------------------------------------------------------
def fidelity_susceptibility(gs, H0, Vs, henvs=None, tol=1e-12, maxiter=100):
    """fidelity_susceptibility(state, H0, Vs, henvs; maxiter, tol) (src/algorithms/fidelity_susceptibility.jl:4-29) for a
    FiniteMPS or InfiniteMPS: for every perturbation V the tangent vector Tos with B_i = dAC_V(i) * AC_i (projected on the
    left null spaces) is fed to the linear system  H_eff x = Tos  (GMRES, x0 = Tos) with the effective excitation Hamiltonian
    of H0 at zero momentum; the susceptibility matrix is dot(x_a, x_b)."""
    from .krylov import linsolve_gmres
    from .infinite import MPOHamInfEnv
    infinite = hasattr(gs, "N")
    if infinite:
        henvs = henvs or MPOHamInfEnv(gs, H0)
        qp = InfiniteQP(gs, 0.0)
        L = gs.N
        AC = lambda i: gs.AC[i]
        E = [np.vdot(AC(i), dAC_sparse(AC(i), H0[i], henvs.leftenv(i, gs), henvs.rightenv(i, gs))) for i in range(L)]
        heff = lambda v: qp.pack(effective_excitation_hamiltonian_infinite(H0, qp, qp.unpack(v), henvs, E, 1e-13))
    else:
        henvs = henvs or FinEnv(gs, H0)
        qp = FiniteQP(gs)
        L = gs.L
        AC = lambda i: gs.get_AC(i)
        E = renormalization_energy(H0, gs, henvs)
        heff = lambda v: qp.pack(effective_excitation_hamiltonian(H0, qp, qp.unpack(v), henvs, E))
    vecs = []
    for V in Vs:
        venvs = MPOHamInfEnv(gs, V) if infinite else FinEnv(gs, V)
        Xs = []
        for i in range(L):
            temp = dAC_sparse(AC(i), V[i], venvs.leftenv(i, gs), venvs.rightenv(i, gs))
            Xs.append(np.einsum('apk,apr->kr', np.conj(qp.VLs[i]), temp))
        b = qp.pack(Xs)
        x, info = linsolve_gmres(heff, b, b, 0.0, 1.0, tol, 30, maxiter)
        vecs.append(x)
    return np.array([[np.vdot(a, b) for b in vecs] for a in vecs])
