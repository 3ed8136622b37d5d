"""DMRG2 / find_groundstate: mirror of src/algorithms/groundstate/dmrg.jl:57-141 (two-site sweep), with
every tensor operation executed by libmpskit_b200 on the device."""
from __future__ import annotations
import ctypes as C
from dataclasses import dataclass, field
import numpy as np
from . import _lib
from .tensor import DeviceTensor
from .derivatives import ddAC2, ddAC
from .fixedpoint import fixedpoint, Arnoldi, Lanczos
from .environments import environments


# TensorKit truncation schemes used at dmrg.jl:75,96
def notrunc():
    return (_lib.NOTRUNC, 0.0)


def truncdim(D):
    return (_lib.TRUNCDIM, float(D))


def truncerr(eps):
    return (_lib.TRUNCERR, float(eps))


def truncbelow(eps):
    return (_lib.TRUNCBELOW, float(eps))


@dataclass
class DMRG2:
    """dmrg.jl:71-78"""
    tol: float = 1e-12
    maxiter: int = 100
    eigalg: object = field(default_factory=Arnoldi)
    trscheme: tuple = field(default_factory=lambda: truncerr(1e-6))
    verbosity: int = 0
    finalize: object = None


def tsvd(psi, x, trscheme, normalize=True):
    """al, c, ar, eps = tsvd!(x; trunc, alg=SVD()) followed by normalize!(c) (dmrg.jl:96-97)."""
    ctx = psi.ctx
    d1 = len(psi.pq)
    pc = (C.c_int32 * d1)(*psi.pq)
    U, S, Vh = C.c_void_p(), C.c_void_p(), C.c_void_p()
    eps = C.c_double()
    ctx.check(ctx.lib.mpsk_tsvd(ctx.h, x.h, d1, pc, _lib.c_trunc(trscheme[0] | (0 if normalize else 0x100), trscheme[1]), C.byref(U), C.byref(S), C.byref(Vh), C.byref(eps)))
    mod = psi.pspace.modulus
    al, c, ar = DeviceTensor.wrap(ctx, U, mod), DeviceTensor.wrap(ctx, S, mod), DeviceTensor.wrap(ctx, Vh, mod)
    c._is_bond = True
    return al, c, ar, eps.value


def tsvd_spectrum(ctx):
    """per coupled sector: all singular values of the last tsvd and how many were kept."""
    n = C.c_int32()
    ctx.lib.mpsk_tsvd_spectrum(ctx.h, C.byref(n), None, None, None, None)
    q = (C.c_int32 * max(n.value, 1))(); cnt = (C.c_int32 * max(n.value, 1))(); kept = (C.c_int32 * max(n.value, 1))()
    ctx.lib.mpsk_tsvd_spectrum(ctx.h, C.byref(n), q, cnt, kept, None)
    tot = sum(cnt[i] for i in range(n.value))
    vals = (C.c_double * max(tot, 1))()
    ctx.lib.mpsk_tsvd_spectrum(ctx.h, C.byref(n), q, cnt, kept, vals)
    out, off = {}, 0
    for i in range(n.value):
        out[q[i]] = (np.array(vals[off:off + cnt[i]]), kept[i])
        off += cnt[i]
    return out


def contract_ac2(psi, a1, a2):
    h = C.c_void_p()
    psi.ctx.check(psi.ctx.lib.mpsk_contract_ac2(psi.ctx.h, a1.h, a2.h, C.byref(h)))
    return DeviceTensor.wrap(psi.ctx, h, psi.pspace.modulus)


def _mul_right(psi, A, Cm):
    h = C.c_void_p()
    psi.ctx.check(psi.ctx.lib.mpsk_mul_bond_right(psi.ctx.h, A.h, Cm.h, C.byref(h)))
    return DeviceTensor.wrap(psi.ctx, h, psi.pspace.modulus)


def find_groundstate_inplace(psi, H, alg: DMRG2, envs=None, log=None):
    """find_groundstate!(psi, H, alg::DMRG2, envs) (dmrg.jl:80-137)."""
    envs = envs if envs is not None else environments(psi, H)
    L = len(psi)
    eps = 0.0
    from .iterlog import IterLog
    ilog = IterLog("DMRG2", alg.verbosity)                       # dmrg.jl:84-87
    obj = lambda: np.sum(expectation_value(psi, H, envs)) if not hasattr(H, "_environments") else None
    ilog.init(2 * alg.tol, obj)
    for it in range(alg.maxiter):
        eps_s = np.zeros(L)
        for direction, rng_ in (("lr", range(0, L - 1)), ("rl", range(L - 3, -1, -1))):
            for pos in rng_:
                if direction == "lr":
                    ac2 = contract_ac2(psi, psi.AC[pos], psi.AR[pos + 1])        # dmrg.jl:92
                else:
                    ac2 = contract_ac2(psi, psi.AL[pos], psi.AC[pos + 1])        # dmrg.jl:108
                h = ddAC2(pos, psi, H, envs, ac2)
                val, new, info = fixedpoint(h, ac2, "SR", alg.eigalg, return_info=True)   # dmrg.jl:94,110
                h.free()
                al, c, ar, eps_tr = tsvd(psi, new, alg.trscheme)                     # dmrg.jl:96-97
                # v = <ac2 | al c ar>  (dmrg.jl:98-100)
                alc = _mul_right(psi, al, c)
                rec = contract_ac2(psi, alc, ar)
                v = rec.dot(ac2)
                eps_s[pos] = max(eps_s[pos], abs(1 - abs(v)))
                if log is not None:
                    spec = tsvd_spectrum(psi.ctx)
                    log.append(dict(sweep=it, dir=direction, pos=pos, energy=val.real, eps_trunc=eps_tr,
                                    kept={q: k for q, (s, k) in spec.items()}, numops=info["numops"],
                                    spectrum={q: s for q, (s, k) in spec.items()}))
                if direction == "lr":
                    psi.AC[pos] = (al, c)                                            # dmrg.jl:102-103
                    psi.AC[pos + 1] = (c, ar)
                else:
                    psi.AC[pos + 1] = (c, ar)                                        # dmrg.jl:118-119
                    psi.AC[pos] = (al, c)
                for t in (ac2, new, alc, rec):
                    t.free()
                envs.drop_stale(psi)
        eps = eps_s.max()
        if alg.finalize is not None:
            psi, envs = alg.finalize(it, psi, H, envs)
        if eps <= alg.tol:
            ilog.finish(it + 1, eps, obj)                        # dmrg.jl:124-132
            break
        if it + 1 == alg.maxiter:
            ilog.cancel(it + 1, eps, obj)
        else:
            ilog.iteration(it + 1, eps, obj)
    return psi, envs, eps


def find_groundstate(psi, H, alg, envs=None, log=None):
    """find_groundstate(psi, H, alg) = find_groundstate!(copy(psi), ...) (dmrg.jl:139-141)."""
    return find_groundstate_inplace(psi.copy(), H, alg, envs, log)


def expectation_value(psi, H, envs):
    """sum(expectation_value(psi, H, envs)) (expval.jl:92-109): total energy <psi|H|psi>/<psi|psi>, evaluated as
    <AC| ddAC |AC> / <AC|AC> at the current centre (the per-site split of the reference sums to this number)."""
    pos = next((i for i in range(len(psi)) if psi.ACs[i] is not None or (psi.ARs[i] is not None and psi.CLs[i] is not None)
                or (psi.ALs[i] is not None and psi.CLs[i + 1] is not None)), 0)
    ac = psi.AC[pos]
    h = ddAC(pos, psi, H, envs)
    y = h(ac)
    e = ac.dot(y) / ac.dot(ac)
    h.free(); y.free()
    return e.real if isinstance(e, complex) else e


def variance(psi, H, envs=None):
    """variance(state::FiniteMPS, H, envs) (src/algorithms/toolbox.jl:141-145; InfiniteMPS: 136-139): Re(<H * H> - <H>^2) with H * H the product MPO
    (sparsempo.jl:232-264) and its own finite environments; everything but the host-side operator product runs on the device."""
    from .operators import mpo_product
    if hasattr(H, "ops") and hasattr(H, "sum"):
        # variance(psi, H::LazySum, envs = environments(psi, sum(H))) (toolbox.jl:176-181)
        if envs is not None and hasattr(envs, "envs"):
            raise ValueError("The environment cannot be Lazy i.e. use environments of sum(H)")
        H = H.sum()
    if hasattr(psi, "N") and not hasattr(psi, "L"):
        # variance(state::InfiniteMPS, H, envs) (toolbox.jl:136-139): rescaled_H = H - <H> per site,
        # Re(sum(<rescaled_H * rescaled_H>)) -- the extensive parts cancel, the variance density remains
        from . import infinite as di
        envs = envs if envs is not None else di.environments(psi, H)
        e = di.expectation_value(psi, H, envs)
        Hr = H - [complex(x) for x in e]
        H2 = Hr * Hr
        return float(np.sum(di.expectation_value(psi, H2, di.environments(psi, H2))).real)
    envs = envs if envs is not None else environments(psi, H)
    H2 = mpo_product(H, H)
    e2 = expectation_value(psi, H2, environments(psi, H2))
    e1 = expectation_value(psi, H, envs)
    return float(e2 - e1 * e1)
