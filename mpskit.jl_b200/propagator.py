"""Dynamical DMRG on the device: propagator(A, z, H, DynamicalDMRG(flavour)) (src/algorithms/propagator/corvector.jl:44-158)
for finite MPS.  Every local step is a linear solve with the single-site effective Hamiltonian: the NaiveInvert flavour runs
the library's GMRES on the effective-Hamiltonian handle itself (mpsk_linsolve with a0 = -z, a1 = 1); the Jeckelmann flavour
solves with the LinearCombination (-2w H_AC + (H^2)_AC) through the host-driven GMRES over the device VectorInterface.  The
right-hand sides are the overlap projections ac_proj (derivatives.jl:210-215) of the source state."""
from __future__ import annotations
from dataclasses import dataclass
from .derivatives import ddAC, LinearCombinationOp
from .environments import environments
from .fixedpoint import linsolve
from .approximate import ac_proj
from .excitations import unit_mpo
from .operators import mpo_product


@dataclass
class DynamicalDMRG:
    """corvector.jl:21-27; flavour: "NaiveInvert" (corvector.jl:44-52) or "Jeckelmann" (corvector.jl:100-107);
    solver_*: the GMRES parameters (Defaults.linearsolver)"""
    flavour: str = "NaiveInvert"
    tol: float = 1e-11
    maxiter: int = 100
    verbosity: int = 0
    solver_tol: float = 1e-12
    solver_krylovdim: int = 30
    solver_maxiter: int = 100


def propagator(A, z, H, alg: DynamicalDMRG = None, init=None, log=None):
    """propagator(psi0, z, H, alg::DynamicalDMRG; init = copy(psi0)) -> (<psi0| (z - H)^-1 |psi0>, solution MPS)"""
    alg = alg or DynamicalDMRG()
    if alg.flavour not in ("NaiveInvert", "Jeckelmann"):
        raise ValueError(f"unknown dynamical DMRG flavour {alg.flavour!r}")
    z = complex(z)
    w, eta = z.real, z.imag
    init = init if init is not None else A.copy()
    L = len(A)
    envs1 = environments(init, H)                                  # environments for h
    mixed = environments(init, unit_mpo(init), A)                  # environments for <init | A>
    jeck = alg.flavour == "Jeckelmann"
    if jeck:
        H2 = mpo_product(H, H)                                     # conj(H) * H (corvector.jl:161) for a Hermitian MPO
        envs2 = environments(init, H2)                             # squaredenvs: trivial boundaries for a finite MPS
    eps = 2 * alg.tol
    for it in range(alg.maxiter):
        eps = 0.0
        for i in list(range(0, L - 1)) + list(range(L - 1, 0, -1)):
            tos = ac_proj(i, init, mixed)
            AC = init.AC[i]
            h1 = ddAC(i, init, H, envs1)
            if jeck:
                h2 = ddAC(i, init, H2, envs2)
                hop = LinearCombinationOp([h1, h2], (-2 * w, 1.0))
                tos.scale(-eta)
                new = linsolve(hop, tos, AC, abs(z) ** 2, 1.0, alg.solver_tol, alg.solver_krylovdim, alg.solver_maxiter)
                h2.free()
            else:
                tos.scale(-1.0)
                new = linsolve(h1, tos, AC, -z, 1.0, alg.solver_tol, alg.solver_krylovdim, alg.solver_maxiter)
            h1.free(); tos.free()
            d = new.copy(); d.axpy(-1.0, AC)
            eps = max(eps, d.norm())
            d.free()
            init.AC[i] = new
            for e in (envs1, mixed) + ((envs2,) if jeck else ()):
                e.drop_stale(init)
        if log is not None:
            log.append(dict(iter=it, eps=eps))
        if eps <= alg.tol:
            break
    tos = ac_proj(0, init, mixed)
    a = complex(tos.dot(init.AC[0]))                               # dot(ac_proj(1, init, mixedenvs), init.AC[1]) = <A | init>
    tos.free()
    if not jeck:
        return a, init
    eA = environments(A, H, init)                                  # <A| H |init> (the cb / b contraction of corvector.jl:146-152)
    y = ac_proj(0, A, eA)
    b = complex(A.AC[0].dot(y))
    y.free()
    return b / eta - w / eta * a + 1j * a, init
