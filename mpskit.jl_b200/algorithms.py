"""Single-site DMRG (src/algorithms/groundstate/dmrg.jl:14-55), calc_galerkin (src/algorithms/toolbox.jl:17-22) and
finite TDVP / TDVP2 (src/algorithms/timestep/tdvp.jl:61-151) on device handles."""
from __future__ import annotations
import ctypes as C
from dataclasses import dataclass, field
import numpy as np
from . import _lib
from .tensor import DeviceTensor
from .operators import SparseMPOSlice
from .derivatives import ddAC, ddC, ddAC2
from .fixedpoint import fixedpoint, integrate, Arnoldi, Lanczos
from .environments import environments
from .dmrg import tsvd, contract_ac2, truncerr, DMRG2, find_groundstate_inplace as _dmrg2_inplace


@dataclass
class DMRG:
    """dmrg.jl:14-20"""
    tol: float = 1e-12
    maxiter: int = 100
    eigalg: object = field(default_factory=Arnoldi)
    verbosity: int = 0
    finalize: object = None


@dataclass
class TDVP:
    """tdvp.jl:14-19"""
    integrator: object = field(default_factory=lambda: Lanczos(tol=1e-12))
    tolgauge: float = 1e-14
    gaugemaxiter: int = 100
    finalize: object = None


@dataclass
class TDVP2:
    """tdvp.jl:107-113"""
    integrator: object = field(default_factory=lambda: Lanczos(tol=1e-12))
    trscheme: tuple = field(default_factory=lambda: truncerr(1e-3))
    finalize: object = None


_unit_slices = {}


def _overlap(psi, A, Abar):
    """(Abar^dagger A)[a'; b'] = sum_{a,s} conj(Abar[a s; a']) A[a s; b'] -- the MPO-free transfer of the identity
    (transfer.jl:18-25) expressed through the sparse-slice transfer with a one-level unit MPO."""
    ctx = psi.ctx
    key = (psi.pq, psi.pspace.modulus)
    if key not in _unit_slices:
        _unit_slices[key] = SparseMPOSlice([[1.0]], (0,), (0,), psi.pq)
    sl = _unit_slices[key]
    V = A.Vl
    eye = np.zeros((V.dim, 1, V.dim), dtype=psi.dtype)
    eye[:, 0, :] = np.eye(V.dim)
    one = DeviceTensor.from_dense(ctx, eye, V, (0,), V, psi.dtype)
    h = C.c_void_p()
    ctx.check(ctx.lib.mpsk_transfer_left(ctx.h, sl.device(ctx), A.h, Abar.h, one.h, C.byref(h)))
    out = DeviceTensor.wrap(ctx, h, psi.pspace.modulus)
    one.free()
    return out


def calc_galerkin(psi, loc, envs):
    """toolbox.jl:17-22: || AC' - AL AL^dagger AC' || with AC' = normalize(ddAC * AC)"""
    h = ddAC(loc, psi, envs.opp, envs)
    ac = h(psi.AC[loc])
    h.free()
    ac.scale(1.0 / ac.norm())
    AL = psi.AL[loc]
    ov = _overlap(psi, ac, AL)                       # AL^dagger AC'   [k; r]
    hh = C.c_void_p()
    psi.ctx.check(psi.ctx.lib.mpsk_mul_bond_right(psi.ctx.h, AL.h, ov.h, C.byref(hh)))
    proj = DeviceTensor.wrap(psi.ctx, hh, psi.pspace.modulus)
    ac.axpy(-1.0, proj)
    out = ac.norm()
    for t in (ac, ov, proj):
        t.free()
    return out


def find_groundstate_dmrg_inplace(psi, H, alg: DMRG, envs=None, log=None):
    """find_groundstate!(psi, H, alg::DMRG, envs) (dmrg.jl:22-55)"""
    envs = envs if envs is not None else environments(psi, H)
    L = len(psi)
    eps = max(calc_galerkin(psi, pos, envs) for pos in range(L))
    from .iterlog import IterLog
    from .excitations import expectation_value as _total_energy
    ilog = IterLog("DMRG", alg.verbosity)                        # dmrg.jl:25-28
    obj = lambda: np.sum(_total_energy(psi, H, envs))
    ilog.init(eps, obj)
    for it in range(alg.maxiter):
        eps_s = np.zeros(L)
        for pos in list(range(0, L - 1)) + list(range(L - 1, 0, -1)):
            h = ddAC(pos, psi, H, envs)
            val, vec, info = fixedpoint(h, psi.AC[pos], "SR", alg.eigalg, return_info=True)
            h.free()
            eps_s[pos] = max(eps_s[pos], calc_galerkin(psi, pos, envs))
            psi.AC[pos] = vec
            if log is not None:
                log.append(dict(sweep=it, pos=pos, energy=val.real, numops=info["numops"]))
        eps = eps_s.max()
        if alg.finalize is not None:
            psi, envs = alg.finalize(it, psi, H, envs)
        if eps <= alg.tol:
            ilog.finish(it + 1, eps, obj)                        # dmrg.jl:42-49
            break
        if it + 1 == alg.maxiter:
            ilog.cancel(it + 1, eps, obj)
        else:
            ilog.iteration(it + 1, eps, obj)
    return psi, envs, eps


@dataclass
class UnionAlg:
    """unionalg.jl:4-9: alg1 & alg2 -- run alg1, then alg2 from its result"""
    alg1: object
    alg2: object

    def __and__(self, other):
        return UnionAlg(self, other)


def _is_infinite(psi):
    return hasattr(psi, "N") and not hasattr(psi, "L")


def default_groundstate_algorithm(psi, tol=1e-12, maxiter=100, verbosity=0, trscheme=None):
    """the algorithm find_groundstate(psi, H; tol, maxiter, verbosity, trscheme) picks (find_groundstate.jl:19-41).  The
    reference finishes InfiniteMPS runs below 1e-4 with GradientGrassmann, which is not built here: VUMPS runs to `tol`."""
    from .infinite import VUMPS, IDMRG2
    if _is_infinite(psi):
        alg = VUMPS(tol=tol, maxiter=maxiter, verbosity=verbosity)
        if trscheme is not None:
            alg = UnionAlg(IDMRG2(tol=min(1e-2, 100 * tol), verbosity=verbosity, trscheme=trscheme), alg)
        return alg
    alg = DMRG(tol=tol, maxiter=maxiter, verbosity=verbosity)
    if trscheme is not None:
        alg = UnionAlg(DMRG2(tol=min(1e-2, 100 * tol), verbosity=verbosity, trscheme=trscheme), alg)
    return alg


def find_groundstate(psi, H, alg=None, envs=None, log=None, *, tol=1e-12, maxiter=100, verbosity=0, trscheme=None):
    """find_groundstate(psi, H, alg, envs) for FiniteMPS (DMRG, DMRG2: find_groundstate!(copy(psi), ...), dmrg.jl:139-141) and
    InfiniteMPS (VUMPS vumps.jl:29-92, IDMRG1 / IDMRG2 idmrg.jl:21-204), alg1 & alg2 (unionalg.jl:22-25), and the keyword form
    find_groundstate(psi, H; tol, maxiter, verbosity, trscheme) of find_groundstate.jl:19-41 (alg = None)."""
    if alg is None:
        alg = default_groundstate_algorithm(psi, tol, maxiter, verbosity, trscheme)
    if isinstance(alg, UnionAlg):
        psi, envs, _ = find_groundstate(psi, H, alg.alg1, envs, log)
        return find_groundstate(psi, H, alg.alg2, envs, log)
    if _is_infinite(psi):
        from . import infinite as di
        if isinstance(alg, di.VUMPS):
            return di.find_groundstate_vumps(psi, H, alg, envs, log)
        if isinstance(alg, di.IDMRG1):
            return di.find_groundstate_idmrg1(psi, H, alg, envs, log)
        if isinstance(alg, di.IDMRG2):
            return di.find_groundstate_idmrg2(psi, H, alg, envs, log)
        raise TypeError("find_groundstate on InfiniteMPS: alg must be VUMPS, IDMRG1, IDMRG2 or a union of them")
    if isinstance(alg, DMRG2):
        return _dmrg2_inplace(psi.copy(), H, alg, envs, log)
    if isinstance(alg, DMRG):
        return find_groundstate_dmrg_inplace(psi.copy(), H, alg, envs, log)
    raise TypeError("find_groundstate on FiniteMPS: alg must be DMRG, DMRG2 or a union of them")


def timestep_inplace(psi, H, t, dt, alg, envs=None):
    """timestep!(psi::FiniteMPS, H, t, dt, alg::TDVP / TDVP2, envs) (tdvp.jl:61-91, 115-151)"""
    envs = envs if envs is not None else environments(psi, H)
    L = len(psi)
    integ = alg.integrator
    if isinstance(alg, TDVP):
        for i in range(0, L - 1):                                        # tdvp.jl:65-73
            h = ddAC(i, psi, H, envs)
            psi.AC[i] = integrate(h, psi.AC[i], t, dt / 2, integ); h.free()
            h = ddC(i, psi, H, envs)
            psi.CR[i] = integrate(h, psi.CR[i], t, -dt / 2, integ); h.free()
        h = ddAC(L - 1, psi, H, envs)                                    # tdvp.jl:75-77
        psi.AC[L - 1] = integrate(h, psi.AC[L - 1], t, dt / 2, integ); h.free()
        for i in range(L - 1, 0, -1):                                    # tdvp.jl:79-87
            h = ddAC(i, psi, H, envs)
            psi.AC[i] = integrate(h, psi.AC[i], t + dt / 2, dt / 2, integ); h.free()
            h = ddC(i - 1, psi, H, envs)
            psi.CR[i - 1] = integrate(h, psi.CR[i - 1], t + dt / 2, -dt / 2, integ); h.free()
        h = ddAC(0, psi, H, envs)
        psi.AC[0] = integrate(h, psi.AC[0], t + dt / 2, dt / 2, integ); h.free()
        return psi, envs
    # TDVP2 (tdvp.jl:115-151)
    for i in range(0, L - 1):
        ac2 = contract_ac2(psi, psi.AC[i], psi.AR[i + 1])
        h = ddAC2(i, psi, H, envs, ac2)
        nac2 = integrate(h, ac2, t, dt / 2, integ); h.free()
        al, c, ar, _ = tsvd(psi, nac2, alg.trscheme, normalize=False)
        psi.AC[i] = (al, c)
        psi.AC[i + 1] = (c, ar)
        if i != L - 2:
            h = ddAC(i + 1, psi, H, envs)
            psi.AC[i + 1] = integrate(h, psi.AC[i + 1], t, -dt / 2, integ); h.free()
    for i in range(L - 1, 0, -1):
        ac2 = contract_ac2(psi, psi.AL[i - 1], psi.AC[i])
        h = ddAC2(i - 1, psi, H, envs, ac2)
        nac2 = integrate(h, ac2, t + dt / 2, dt / 2, integ); h.free()
        al, c, ar, _ = tsvd(psi, nac2, alg.trscheme, normalize=False)
        psi.AC[i - 1] = (al, c)
        psi.AC[i] = (c, ar)
        if i != 1:
            h = ddAC(i - 1, psi, H, envs)
            psi.AC[i - 1] = integrate(h, psi.AC[i - 1], t + dt / 2, -dt / 2, integ); h.free()
    return psi, envs


def timestep(psi, H, t, dt, alg, envs=None):
    """timestep(psi, H, t, dt, alg, envs) = timestep!(copy(psi), ...) (tdvp.jl:93-96, 153-157)"""
    return timestep_inplace(psi.copy(), H, t, dt, alg, envs)


def time_evolve(psi, H, t_span, alg=None, envs=None):
    """time_evolve(psi, H, t_span, alg, envs) (src/algorithms/timestep/time_evolve.jl:17-40): step through the time points
    of t_span with timestep (finite or infinite state), calling alg.finalize(t, psi, H, envs) after every step."""
    from .infinite import InfiniteMPS, timestep_infinite
    alg = alg or TDVP()
    infinite = isinstance(psi, InfiniteMPS)
    if not infinite:
        psi = psi.copy()
        envs = envs if envs is not None else environments(psi, H)
    for k in range(len(t_span) - 1):
        t, dt = t_span[k], t_span[k + 1] - t_span[k]
        if infinite:
            psi, envs = timestep_infinite(psi, H, t, dt, alg, envs)      # applies alg.finalize itself
        else:
            psi, envs = timestep_inplace(psi, H, t, dt, alg, envs)
            if alg.finalize is not None:
                psi, envs = alg.finalize(t, psi, H, envs)
    return psi, envs


def _enable_union_operator():
    """Base.:&(alg1::Algorithm, alg2::Algorithm) = UnionAlg(alg1, alg2) (unionalg.jl:11) for the algorithm dataclasses"""
    from .infinite import VUMPS, IDMRG1, IDMRG2
    for cls in (DMRG, DMRG2, TDVP, TDVP2, VUMPS, IDMRG1, IDMRG2):
        cls.__and__ = lambda a, b: UnionAlg(a, b)


_enable_union_operator()


def exact_diagonalization(H, pspace, L=None, sector=0, num=1, which="SR", alg=None, ctx=None, seed=0, weight=1000.0):
    """exact_diagonalization(opp::MPOHamiltonian; sector, len, num, which, alg) (src/algorithms/ED.jl:4-53): the largest possible
    FiniteMPS of that length -- bond spaces limited only by the Hilbert spaces to their left and right, total charge `sector`
    on the right edge -- whose middle site then parametrises the whole (symmetry sector of the) Hilbert space, so that the
    single-site effective Hamiltonian IS the Hamiltonian; its lowest `num` eigenpairs come from the device Lanczos (further
    ones by shifting the converged ones up by `weight`, host-driven over the device VectorInterface).
    Returns (values, states, infos)."""
    from .finitemps import FiniteMPS
    from .spaces import Space, fuse, infimum, dual
    from .krylov import eigsolve_lanczos
    L = len(H) if L is None else int(L)
    alg = alg or Lanczos(tol=1e-12, krylovdim=30, maxiter=100)
    mod = pspace.modulus
    Vl = [Space((0,), (1,), mod)]
    for _ in range(L):
        Vl.append(fuse(Vl[-1], pspace))
    Vr = [Space((int(sector),), (1,), mod)]
    for _ in range(L):
        Vr.append(fuse(Vr[-1], dual(pspace)))
    Vr = Vr[::-1]
    V = [infimum(Vl[k], Vr[k]) for k in range(L + 1)]
    if any(v.dim == 0 for v in V):
        raise ValueError("invalid sector")                                    # ED.jl:36
    psi = FiniteMPS.random(np.random.default_rng(seed), pspace, L, V[1:L], ctx, left=V[0], right=V[L])
    mid = max(int(round(L / 2)) - 1, 0)                                       # ED.jl:11 (1-based there)
    envs = environments(psi, H)
    h = ddAC(mid, psi, H, envs)
    vals, states, infos, found = [], [], [], []
    for n in range(num):
        if n == 0:
            val, vec, info = fixedpoint(h, psi.AC[mid], which, alg, return_info=True)
        else:
            def f(v):
                y = h(v)
                for u in found:
                    y.axpy(weight * complex(u.dot(v)), u)
                return y
            x0 = DeviceTensor.random(psi.ctx, psi.AC[mid].Vl, psi.AC[mid].mch, psi.AC[mid].Vr, seed + 7 * n, psi.dtype)
            val, vec, info = eigsolve_lanczos(f, x0, which, alg.tol, alg.krylovdim, alg.maxiter)
        found.append(vec)
        st = psi.copy()
        st.AC[mid] = vec
        vals.append(complex(val)); states.append(st); infos.append(info)
    h.free()
    return vals, states, infos
