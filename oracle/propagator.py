"""TEST INFRASTRUCTURE (CPU oracle, never imported by the product path).

Dynamical DMRG of the reference restated on dense charged tensors (src/algorithms/propagator/corvector.jl:44-158):
propagator(A, z, H, DynamicalDMRG(flavour)) -> (<A| (z - H)^-1 |A>, the solution MPS), for finite MPS."""
from __future__ import annotations
import copy
import numpy as np
from .finite import FiniteMPS, FinEnv, unit_mpo
from .contract import dAC_sparse
from .krylov import linsolve_gmres
from .mpo import mpo_product


def _copy(psi: FiniteMPS):
    out = copy.copy(psi)
    out.AL, out.AR, out.AC, out.CL, out.V = list(psi.AL), list(psi.AR), list(psi.AC), list(psi.CL), list(psi.V)
    return out


def propagator(A: FiniteMPS, z, H, flavour="NaiveInvert", tol=1e-11, maxiter=100, solver_tol=1e-12, log=None):
    """corvector.jl:57-98 (NaiveInvert): sweeps of (H_AC - z) AC' = -ac_proj(i) ; corvector.jl:109-158 (Jeckelmann):
    ((H_AC - w)^2 + eta^2) AC' = -eta ac_proj(i), value = b / eta - w / eta a + i a"""
    init = _copy(A)
    L = A.L
    envs1 = FinEnv(init, H)
    mixed = FinEnv(init, unit_mpo(init), A)
    w, eta = np.real(z), np.imag(z)
    if flavour == "Jeckelmann":
        H2 = mpo_product(H, H)                     # conj(H) * H for a Hermitian MPO with real entries
        envs2 = FinEnv(init, H2)
    eps = 2 * tol
    for it in range(maxiter):
        eps = 0.0
        for i in list(range(0, L - 1)) + list(range(L - 1, 0, -1)):
            tos = dAC_sparse(A.get_AC(i), mixed.opp[i], mixed.leftenv(i, init), mixed.rightenv(i, init))
            AC = init.get_AC(i)
            GL, GR = envs1.leftenv(i, init), envs1.rightenv(i, init)
            if flavour == "NaiveInvert":
                new, info = linsolve_gmres(lambda x: dAC_sparse(x, H[i], GL, GR), -tos, AC, -z, 1.0, solver_tol, 30, 100)
            else:
                GL2, GR2 = envs2.leftenv(i, init), envs2.rightenv(i, init)
                f = lambda x: -2 * w * dAC_sparse(x, H[i], GL, GR) + dAC_sparse(x, H2[i], GL2, GR2)
                new, info = linsolve_gmres(f, -eta * tos, AC, abs(z) ** 2, 1.0, solver_tol, 30, 100)
            eps = max(eps, float(np.linalg.norm(new - AC)))
            init.set_AC(i, new)
        if log is not None:
            log.append(dict(iter=it, eps=eps))
        if eps <= tol:
            break
    tos = dAC_sparse(A.get_AC(0), mixed.opp[0], mixed.leftenv(0, init), mixed.rightenv(0, init))
    a = np.vdot(tos, init.get_AC(0))               # dot(ac_proj(1), init.AC[1]) = <A|init>
    if flavour == "NaiveInvert":
        return a, init
    eA = FinEnv(A, H, init)                         # <A| H |init>
    y = dAC_sparse(init.get_AC(0), H[0], eA.leftenv(0, A), eA.rightenv(0, A))
    b = np.vdot(A.get_AC(0), y)
    return b / eta - w / eta * a + 1j * a, init
