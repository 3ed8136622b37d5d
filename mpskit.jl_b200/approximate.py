"""approximate(psi, (O, above), alg) for finite MPS: mirror of src/algorithms/approximate/fvomps.jl:1-114 with the projections
ac_proj / ac2_proj of src/algorithms/derivatives.jl:209-232 evaluated by RECTANGULAR effective-Hamiltonian operators on the
device (ket legs in the spaces of `above`, bra legs in the spaces of the approximating state)."""
from __future__ import annotations
import numpy as np
from .derivatives import MPO_ddAC, MPO_ddAC2
from .environments import environments
from .dmrg import DMRG2, tsvd, contract_ac2, _mul_right
from .algorithms import DMRG


def ac_proj(pos, below, envs):
    """derivatives.jl:216-221: dAC(above.AC[pos], opp[pos], leftenv(envs, pos, below), rightenv(envs, pos, below))"""
    le, re = envs.leftenv(pos, below), envs.rightenv(pos, below)
    x = envs.above.AC[pos]
    op = MPO_ddAC(below.ctx, envs.opp[pos], le, re, x)
    y = op(x); op.free()
    return y


def ac2_proj(pos, below, envs):
    """derivatives.jl:226-232: dAC2(above.AC[pos] * above.AR[pos + 1], opp[pos], opp[pos + 1], leftenv, rightenv)"""
    le, re = envs.leftenv(pos, below), envs.rightenv(pos + 1, below)
    x = contract_ac2(envs.above, envs.above.AC[pos], envs.above.AR[pos + 1])
    op = MPO_ddAC2(below.ctx, envs.opp[pos], envs.opp[pos + 1], le, re, x)
    y = op(x); op.free(); x.free()
    return y


def approximate_inplace(psi, toapprox, alg, envs=None, log=None):
    """approximate!(psi::FiniteMPS, (O, above), alg::DMRG2 / DMRG, envs) (fvomps.jl:11-114)"""
    O, above = toapprox
    envs = envs if envs is not None else environments(psi, O, above)
    L = len(psi)
    eps = 2 * alg.tol
    for it in range(alg.maxiter):
        eps = 0.0
        if isinstance(alg, DMRG2):
            for pos in list(range(0, L - 1)) + list(range(L - 3, -1, -1)):
                new = ac2_proj(pos, psi, envs)
                al, c, ar, _ = tsvd(psi, new, alg.trscheme, normalize=False)            # fvomps.jl:26 (no normalize!)
                old = contract_ac2(psi, psi.AC[pos], psi.AR[pos + 1])
                alc = _mul_right(psi, al, c)
                rec = contract_ac2(psi, alc, ar)
                nold = old.norm()
                if rec.Vr == old.Vr and rec.Vl == old.Vl:
                    rec.axpy(-1.0, old)
                    eps = max(eps, rec.norm() / nold)                                   # fvomps.jl:29
                else:
                    eps = max(eps, 1.0)
                psi.AC[pos] = (al, c)
                psi.AC[pos + 1] = (c, ar)
                for t in (new, old, alc, rec):
                    t.free()
                envs.drop_stale(psi)
        elif isinstance(alg, DMRG):
            for pos in list(range(0, L - 1)) + list(range(L - 1, 0, -1)):
                new = ac_proj(pos, psi, envs)
                d = new.copy(); d.axpy(-1.0, psi.AC[pos])
                eps = max(eps, d.norm() / new.norm())                                   # fvomps.jl:70
                d.free()
                psi.AC[pos] = new
                envs.drop_stale(psi)
        else:
            raise TypeError("approximate: alg must be DMRG or DMRG2")
        if log is not None:
            log.append(dict(iter=it, eps=eps))
        if alg.finalize is not None:
            psi, envs = alg.finalize(it, psi, toapprox, envs)
        if eps < alg.tol:
            break
    return psi, envs, eps


def approximate(psi, toapprox, alg, envs=None, log=None):
    """approximate(psi, toapprox, alg) = approximate!(copy(psi), ...) (fvomps.jl:2-4); InfiniteMPS with VOMPS:
    src/algorithms/approximate/vomps.jl:1-11 (statmech.approximate_vomps)"""
    if hasattr(psi, "N") and not hasattr(psi, "L"):
        from .statmech import VOMPS, approximate_vomps, approximate_idmrg1, approximate_idmrg2
        from .infinite import IDMRG1, IDMRG2
        if isinstance(alg, VOMPS):
            return approximate_vomps(psi, toapprox, alg, envs, log)
        if isinstance(alg, IDMRG1):                    # approximate/idmrg.jl:1-58
            return approximate_idmrg1(psi, toapprox, alg, envs, log)
        if isinstance(alg, IDMRG2):                    # approximate/idmrg.jl:60-145
            return approximate_idmrg2(psi, toapprox, alg, envs, log)
        raise TypeError("approximate on InfiniteMPS: alg must be VOMPS, IDMRG1 or IDMRG2")
    return approximate_inplace(psi.copy(), toapprox, alg, envs, log)
