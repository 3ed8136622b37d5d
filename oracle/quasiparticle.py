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
