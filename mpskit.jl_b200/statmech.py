"""leading_boundary(psi::InfiniteMPS, O::DenseMPO, VUMPS()) on the device: src/algorithms/statmech/vumps.jl:15-84,
PerMPOInfEnv + mixed_fixpoints (src/environments/permpoinfenv.jl:23-63,138-189), expectation_value
(src/algorithms/expval.jl:165-173).  Single-row unit cells (an InfiniteMPS with any number of columns); the
non-Hermitian eigenproblems are solved by the library's device-resident Krylov-Schur Arnoldi (mpsk_eigsolve, hermitian = 0)."""
from __future__ import annotations
import ctypes as C
import numpy as np
from .tensor import DeviceTensor
from .derivatives import MPO_ddAC, MPO_ddC
from .fixedpoint import fixedpoint, Arnoldi
from .infinite import InfiniteMPS, TransferMatrix, regauge, dyn_tol, TOLGAUGE, mul_right, VUMPS
from .algorithms import _overlap


def _tl(psi, sl, A, v):
    ctx = psi.ctx
    h = C.c_void_p()
    ctx.check(ctx.lib.mpsk_transfer_left(ctx.h, sl.device(ctx), A.h, A.h, v.h, C.byref(h)))
    return DeviceTensor.wrap(ctx, h, psi.pspace.modulus)


def _tr(psi, sl, A, v):
    ctx = psi.ctx
    h = C.c_void_p()
    ctx.check(ctx.lib.mpsk_transfer_right(ctx.h, sl.device(ctx), A.h, A.h, v.h, C.byref(h)))
    return DeviceTensor.wrap(ctx, h, psi.pspace.modulus)


class PerMPOInfEnv:
    def __init__(self, psi: InfiniteMPS, opp, tol=1e-12):
        self.opp, self.tol = opp, tol
        self.lw = self.rw = None
        self.dependency = None
        self.recalculate(psi)

    def recalculate(self, psi, tol=None):
        """mixed_fixpoints (permpoinfenv.jl:138-189) for one row"""
        tol = self.tol if tol is None else tol
        N, ctx = psi.N, psi.ctx
        lc = [-c for c in self.opp[0].domcharges]
        rc = [c for c in self.opp[N - 1].imcharges]
        V0 = psi.V[0]
        same = self.lw is not None and self.lw[0].Vl == V0
        L0 = self.lw[0] if same else DeviceTensor.random(ctx, V0, lc, V0, 17, psi.dtype)
        R0 = self.rw[N - 1] if same else DeviceTensor.random(ctx, V0, rc, V0, 18, psi.dtype)

        # TransferMatrix(AL, O, AL) over the unit cell (permpoinfenv.jl:150-170) as single device operators: the Arnoldi
        # (Krylov-Schur) solves run inside the library
        mids = [self.opp[i] for i in range(N)]
        TL = TransferMatrix(list(psi.AL), mids, list(psi.AL)).flip()
        TR = TransferMatrix(list(psi.AR), mids, list(psi.AR))
        alg = Arnoldi(tol=tol, krylovdim=30, maxiter=100, hermitian=False)
        GL, GR = [None] * N, [None] * N
        _, GL[0] = fixedpoint(TL, L0, "LM", alg)
        for c in range(1, N):
            GL[c] = _tl(psi, self.opp[c - 1], psi.AL[c - 1], GL[c - 1])
        _, GR[N - 1] = fixedpoint(TR, R0, "LM", alg)
        for c in range(N - 2, -1, -1):
            GR[c] = _tr(psi, self.opp[c + 1], psi.AR[c + 1], GR[c + 1])
        for c in range(N):                                  # permpoinfenv.jl:180-185 (+ removal of the arbitrary Arnoldi phase)
            h = MPO_ddC(ctx, GL[(c + 1) % N], GR[c], psi.CR[c])
            lam = complex(psi.CR[c].dot(h(psi.CR[c]))); h.free()
            GL[(c + 1) % N].scale(np.conj(lam) / abs(lam) / np.sqrt(abs(lam)))
            GR[c].scale(1.0 / np.sqrt(abs(lam)))
        self.lw, self.rw, self.dependency = GL, GR, psi
        return self

    def leftenv(self, pos, psi):
        if self.dependency is not psi:
            self.recalculate(psi)
        return self.lw[pos % psi.N]

    def rightenv(self, pos, psi):
        if self.dependency is not psi:
            self.recalculate(psi)
        return self.rw[pos % psi.N]


def environments(psi: InfiniteMPS, opp, tol=1e-12):
    """environments(state::InfiniteMPS, opp::DenseMPO) (permpoinfenv.jl:19-27)"""
    return PerMPOInfEnv(psi, opp, tol)


def expectation_value(psi, envs):
    """expval.jl:165-173: <AC| O |AC> with the environments, per column"""
    out = []
    for i in range(psi.N):
        h = MPO_ddAC(psi.ctx, envs.opp[i], envs.leftenv(i, psi), envs.rightenv(i, psi), psi.AC[i])
        y = h(psi.AC[i]); h.free()
        out.append(complex(psi.AC[i].dot(y)))
        y.free()
    return np.array(out)


def calc_galerkin(psi, loc, envs):
    """toolbox.jl:26-38 for one row"""
    h = MPO_ddAC(psi.ctx, envs.opp[loc], envs.leftenv(loc, psi), envs.rightenv(loc, psi), psi.AC[loc])
    ac = h(psi.AC[loc]); h.free()
    ac.scale(1.0 / ac.norm())
    AL = psi.AL[loc]
    ov = _overlap(psi, ac, AL)
    proj = mul_right(AL, ov)
    ac.axpy(-1.0, proj)
    return ac.norm()


def leading_boundary(psi: InfiniteMPS, opp, alg: VUMPS = None, envs=None, log=None):
    """leading_boundary(psi, opp, alg::VUMPS, envs) (statmech/vumps.jl:15-84)"""
    alg = alg or VUMPS(tol=1e-10)
    envs = envs or environments(psi, opp)
    N = psi.N
    eps = max(calc_galerkin(psi, i, envs) for i in range(N))
    for it in range(1, alg.maxiter + 1):
        etol = dyn_tol(eps, it, 1e-12, 1e-5, 1e-5)
        eig = Arnoldi(tol=etol, krylovdim=30, maxiter=100, hermitian=False)
        newAL = []
        for col in range(N):
            hac = MPO_ddAC(psi.ctx, opp[col], envs.leftenv(col, psi), envs.rightenv(col, psi), psi.AC[col])
            _, ac = fixedpoint(hac, psi.AC[col], "LM", eig); hac.free()
            hc = MPO_ddC(psi.ctx, envs.leftenv(col + 1, psi), envs.rightenv(col, psi), psi.CR[col])
            _, c = fixedpoint(hc, psi.CR[col], "LM", eig); hc.free()
            newAL.append(regauge(ac, c))
        gtol = dyn_tol(eps, it, TOLGAUGE, 1e-5, 1e-8)
        psi = InfiniteMPS.from_AL(newAL, psi.CR[N - 1], psi.V, psi.pspace, psi.ctx, gtol, 100)
        envs.recalculate(psi, dyn_tol(eps, it, 1e-12, 1e-5, 1e-5))
        if alg.finalize is not None:
            psi, envs = alg.finalize(it, psi, opp, envs)
        eps = max(calc_galerkin(psi, i, envs) for i in range(N))
        if log is not None:
            log.append(dict(iter=it, eps=eps, value=expectation_value(psi, envs)))
        if eps <= alg.tol:
            break
    return psi, envs, eps
