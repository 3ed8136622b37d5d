"""Infinite-MPS side of the path on device handles: uniform gauge (src/states/ortho.jl:127-257,
src/states/infinitemps.jl:139-204), TransferMatrix (src/transfermatrix/transfermatrix.jl:4-90), MPOHamInfEnv
(src/environments/mpohaminfenv.jl:4-215), VUMPS (src/algorithms/groundstate/vumps.jl:18-92), the infinite
expectation value (src/algorithms/expval.jl:111-124) and calc_galerkin (src/algorithms/toolbox.jl:17-25).
Every tensor operation is a libmpskit_b200 call; Krylov solves on non-Hermitian maps and linear systems are
host-driven over the VectorInterface (krylov.py) exactly like KrylovKit drives the reference."""
from __future__ import annotations
import ctypes as C
import warnings
from dataclasses import dataclass, field
import numpy as np
from ._lib import default_context
from .spaces import Space
from .tensor import DeviceTensor
from .operators import SparseMPOSlice, MPOHamiltonian
from .finitemps import qrpos, lqpos, _mask
from .derivatives import MPO_ddAC, MPO_ddC
from .fixedpoint import fixedpoint, linsolve, Arnoldi, Lanczos
from . import krylov

TOLGAUGE = 1e-14


def dyn_tol(eps, it, tol_min, tol_max, factor):
    """updatetol(::DynamicTol, iter, eps) (dynamictols.jl:50-53)"""
    return float(np.clip(eps * factor / np.sqrt(it), tol_min, tol_max))


# ------------------------------------------------------------------------------------------ small helpers
def _call2(ctx, fn, a, b, mod):
    h = C.c_void_p()
    ctx.check(fn(ctx.h, a.h, b.h, C.byref(h)))
    return DeviceTensor.wrap(ctx, h, mod)


def mul_right(A, Cm):
    return _call2(A.ctx, A.ctx.lib.mpsk_mul_bond_right, A, Cm, A.Vl.modulus)


def mul_left(Cm, A):
    return _call2(A.ctx, A.ctx.lib.mpsk_mul_bond_left, Cm, A, A.Vl.modulus)


def identity_bond(ctx, V: Space, dtype):
    eye = np.zeros((V.dim, 1, V.dim), dtype=dtype)
    eye[:, 0, :] = np.eye(V.dim)
    t = DeviceTensor.from_dense(ctx, eye, V, (0,), V, dtype)
    t._is_bond = True
    return t


_slices = {}


def _entry_slice(sl: SparseMPOSlice, j, k):
    """1 x 1 slice holding entry (j, k) (None: the unit MPO, i.e. an MPO-free transfer, transfer.jl:18-45)"""
    key = (id(sl), j, k)
    if key not in _slices:
        if j is None:
            _slices[key] = SparseMPOSlice([[1.0]], (0,), (0,), sl.pcharges)
        else:
            _slices[key] = SparseMPOSlice([[sl.Os[j][k]]], (sl.domcharges[j],), (sl.imcharges[k],), sl.pcharges)
        _slices[key]._parent = sl          # keep the parent alive so that id() stays unique
    return _slices[key]


def transfer_left(v, sl1, A, Ab):
    """v * TransferMatrix(A, O, Ab) for a single level tensor v and a 1x1 slice (transfer.jl:18-25,105-107)"""
    ctx = A.ctx
    h = C.c_void_p()
    ctx.check(ctx.lib.mpsk_transfer_left(ctx.h, sl1.device(ctx), A.h, Ab.h, v.h, C.byref(h)))
    return DeviceTensor.wrap(ctx, h, A.Vl.modulus)


def transfer_right(v, sl1, A, Ab):
    ctx = A.ctx
    h = C.c_void_p()
    ctx.check(ctx.lib.mpsk_transfer_right(ctx.h, sl1.device(ctx), A.h, Ab.h, v.h, C.byref(h)))
    return DeviceTensor.wrap(ctx, h, A.Vl.modulus)


class _DeviceOp:
    """an mpsk_op handle created by mpsk_tmop_create (persistent plans; freed with the wrapper)"""

    def __init__(self, ctx, h):
        self.ctx, self.h = ctx, h

    def __call__(self, v):
        y = v.zerovector()
        self.ctx.check(self.ctx.lib.mpsk_op_apply(self.ctx.h, self.h, v.h, y.h))
        return y

    def free(self):
        if self.h is not None and self.ctx.h:
            self.ctx.lib.mpsk_op_free(self.ctx.h, self.h)
        self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def _tmop(ctx, left, middle, above, below, v_like, ladj=None, rvec=None):
    n = len(above)
    sl = (C.c_void_p * n)(*[(m_.device(ctx) if m_ is not None else None) for m_ in middle])
    A = (C.c_void_p * n)(*[a.h for a in above])
    Ab = (C.c_void_p * n)(*[a.h for a in below])
    h = C.c_void_p()
    ctx.check(ctx.lib.mpsk_tmop_create(ctx.h, 1 if left else 0, n, sl, A, Ab, v_like.h,
                                       ladj.h if ladj is not None else None, rvec.h if rvec is not None else None, C.byref(h)))
    return _DeviceOp(ctx, h)


class TransferMatrix:
    """TransferMatrix(above, middle, below) over a unit cell; `tm * v` / `v * tm` (transfermatrix.jl:43-66).
    above / below: lists of site tensors; middle: list of MPO slices (or None: MPO-free transfer).  The product over the
    unit cell is ONE device operator (mpsk_tmop_create) whose plans are built once per vector structure."""

    def __init__(self, above, middle, below, isflipped=False):
        self.above = list(above) if isinstance(above, (list, tuple)) else [above]
        self.below = list(below) if isinstance(below, (list, tuple)) else [below]
        mid = middle if isinstance(middle, (list, tuple)) else [middle] * len(self.above)
        self.middle = list(mid)
        self.isflipped = isflipped
        self.ctx = self.above[0].ctx
        self._ops = {}

    def flip(self):
        return TransferMatrix(self.above, self.middle, self.below, not self.isflipped)

    def device_op(self, v, ladj=None, rvec=None):
        key = (v.Vl, v.mch, v.Vr, id(ladj), id(rvec))
        if key not in self._ops:
            self._ops[key] = _tmop(self.ctx, self.isflipped, self.middle, self.above, self.below, v, ladj, rvec)
        return self._ops[key]

    def __call__(self, v):
        return self.device_op(v)(v)

    def __mul__(self, v):
        return self(v)

    def __rmul__(self, v):
        return self.flip()(v)


class RegTransferMatrix:
    """RegTransferMatrix (transfermatrix.jl:28-33,55,68-76): v -> tm(v) - <lvec, tm(v)> rvec with <l, v> = sum l[1;2] v[2;1];
    the projection runs inside the device operator (the coefficient never leaves the GPU)."""

    def __init__(self, tm, lvec, rvec):
        self.tm, self.ladj, self.rvec, self.ctx = tm, lvec.adjoint(), rvec, tm.ctx

    def device_op(self, v):
        return self.tm.device_op(v, self.ladj, self.rvec)

    def __call__(self, v):
        return self.device_op(v)(v)


def regularize(tm, lvec, rvec):
    return RegTransferMatrix(tm, lvec, rvec)


# ------------------------------------------------------------------------------------------ InfiniteMPS
class InfiniteMPS:
    def __init__(self, AL, AR, CR, AC, V, pspace, ctx):
        self.AL, self.AR, self.CR, self.AC = AL, AR, CR, AC      # lists over the unit cell; CR[i] right of site i
        self.V, self.pspace, self.ctx = V, pspace, ctx
        self.pq = tuple(int(c) for c in pspace.index_charges())
        self.N = len(AL)
        self.dtype = AL[0].dtype

    def __len__(self):
        return self.N

    @staticmethod
    def from_tensors(A, V, pspace, ctx=None, C0=None, tol=TOLGAUGE, maxiter=100):
        """InfiniteMPS(A::Vector) (infinitemps.jl:139-171): gaugefix! order :LR, AC = AL * CR"""
        ctx = ctx or default_context()
        N = len(A)
        C0 = C0 if C0 is not None else identity_bond(ctx, V[0], A[0].dtype)
        AL, CR = uniform_leftorth(A, C0, tol, maxiter)
        AR, CR = uniform_rightorth(AL, CR[N - 1], tol, maxiter)
        AC = [mul_right(AL[i], CR[i]) for i in range(N)]
        return InfiniteMPS(AL, AR, CR, AC, V, pspace, ctx)

    @staticmethod
    def from_AL(AL, C0, V, pspace, ctx, tol=TOLGAUGE, maxiter=100):
        """InfiniteMPS(AL, C0) (infinitemps.jl:173-187): gaugefix! order :R"""
        N = len(AL)
        AR, CR = uniform_rightorth(AL, C0, tol, maxiter)
        AC = [mul_right(AL[i], CR[i]) for i in range(N)]
        return InfiniteMPS(list(AL), AR, CR, AC, V, pspace, ctx)

    @staticmethod
    def random(rng, pspace: Space, Vs, ctx=None, dtype=np.complex128):
        ctx = ctx or default_context()
        N = len(Vs)
        pq = pspace.index_charges()
        A = []
        for i in range(N):
            Vl, Vr = Vs[i], Vs[(i + 1) % N]
            msk = _mask(Vl, pq, Vr)
            a = rng.standard_normal(msk.shape) + 1j * rng.standard_normal(msk.shape)
            A.append(DeviceTensor.from_dense(ctx, (a * msk).astype(dtype), Vl, [int(c) for c in pq], Vr, dtype))
        return InfiniteMPS.from_tensors(A, list(Vs), pspace, ctx)


def _normalize(t):
    t.scale(1.0 / t.norm())
    return t


def uniform_leftorth(A, C0, tol=TOLGAUGE, maxiter=100, eig_miniter=10):
    """ortho.jl:143-201 (LeftCanonical): QRpos sweeps, Arnoldi fixed point of flip(TransferMatrix(A, AL)) after 10 sweeps"""
    ctx = A[0].ctx
    mod = A[0].Vl.modulus
    N = len(A)
    CR = [None] * N
    CR[N - 1] = _normalize(C0.copy())
    AL = [None] * N
    eps = np.inf
    for it in range(maxiter + 2):
        if it >= eig_miniter:
            etol = dyn_tol(eps ** 2, 1, TOLGAUGE / 10, np.inf, 1)
            tm = TransferMatrix(list(A), None, list(AL)).flip()
            _, vec = fixedpoint(tm, CR[N - 1], "LM", Arnoldi(tol=etol, krylovdim=30, hermitian=False))
            _, CR[N - 1] = qrpos(ctx, vec, mod)
        Cprev = CR[N - 1]
        for i in range(N):
            CA = mul_left(CR[(i - 1) % N], A[i])
            AL[i], CR[i] = qrpos(ctx, CA, mod)
        _normalize(CR[N - 1])
        d = Cprev.copy(); d.axpy(-1.0, CR[N - 1]); eps = d.norm()
        if eps < tol or it + 1 > maxiter:
            break
    return AL, CR


def uniform_rightorth(A, C0, tol=TOLGAUGE, maxiter=100, eig_miniter=10):
    """ortho.jl:203-257 (RightCanonical)"""
    ctx = A[0].ctx
    mod = A[0].Vl.modulus
    N = len(A)
    CR = [None] * N
    CR[N - 1] = _normalize(C0.copy())
    AR = [None] * N
    eps = np.inf
    for it in range(maxiter + 2):
        if it >= eig_miniter:
            etol = dyn_tol(eps ** 2, 1, TOLGAUGE / 10, np.inf, 1)
            tm = TransferMatrix(list(A), None, list(AR))
            _, vec = fixedpoint(tm, CR[N - 1], "LM", Arnoldi(tol=etol, krylovdim=30, hermitian=False))
            CR[N - 1], _ = lqpos(ctx, vec, mod)
        Cprev = CR[N - 1]
        for i in range(N - 1, -1, -1):
            AC = mul_right(A[i], CR[i])
            CR[(i - 1) % N], AR[i] = lqpos(ctx, AC, mod)
        _normalize(CR[N - 1])
        d = Cprev.copy(); d.axpy(-1.0, CR[N - 1]); eps = d.norm()
        if eps < tol or it + 1 > maxiter:
            break
    return AR, CR


def regauge(AC, Cm):
    """regauge!(AC, CR) (ortho.jl:127-131): AL = Q_AC * Q_C'"""
    ctx, mod = AC.ctx, AC.Vl.modulus
    Qac, _ = qrpos(ctx, AC, mod)
    Qc, _ = qrpos(ctx, Cm, mod)
    return mul_right(Qac, Qc.adjoint())


# ------------------------------------------------------------------------------------------ MPOHamInfEnv
def _isid(H: MPOHamiltonian, i):
    return all(s.isscal(i, i) and abs(s.Os[i][i] - 1) < 1e-14 for s in H.slices)


class MPOHamInfEnv:
    """mpohaminfenv.jl:4-74; lw[j][i] / rw[j][i]: level-j environments at site i (single-label tensors)"""

    def __init__(self, psi: InfiniteMPS, H: MPOHamiltonian, tol=1e-12, maxiter=100):
        self.opp, self.tol, self.maxiter = H, tol, maxiter
        N, od = psi.N, H.odim
        ctx = psi.ctx
        self.lw = [[DeviceTensor.zeros(ctx, psi.V[i], (-H[i].domcharges[j],), psi.V[i], psi.dtype) for i in range(N)] for j in range(od)]
        self.rw = [[DeviceTensor.zeros(ctx, psi.V[(i + 1) % N], (H[i].imcharges[j],), psi.V[(i + 1) % N], psi.dtype) for i in range(N)] for j in range(od)]
        self.dependency = None
        self._stacked = {}
        self.recalculate(psi)

    def recalculate(self, psi, tol=None):
        tol = self.tol if tol is None else tol
        calclw(self.lw, psi, self.opp, tol, self.maxiter)
        calcrw(self.rw, psi, self.opp, tol, self.maxiter)
        self.dependency = psi
        self._stacked = {}
        return self

    def leftenv(self, pos, psi):
        """vector of levels as ONE stacked environment tensor (what the matvec consumes)"""
        if self.dependency is not psi:
            self.recalculate(psi)
        key = ("l", pos % psi.N)
        if key not in self._stacked:
            self._stacked[key] = DeviceTensor.stack([self.lw[j][pos % psi.N] for j in range(self.opp.odim)])
        return self._stacked[key]

    def rightenv(self, pos, psi):
        if self.dependency is not psi:
            self.recalculate(psi)
        key = ("r", pos % psi.N)
        if key not in self._stacked:
            self._stacked[key] = DeviceTensor.stack([self.rw[j][pos % psi.N] for j in range(self.opp.odim)])
        return self._stacked[key]


def environments(psi: InfiniteMPS, H: MPOHamiltonian, tol=1e-12, maxiter=100):
    """environments(psi::InfiniteMPS, H::MPOHamiltonian; solver) (mpohaminfenv.jl:40-45)"""
    return MPOHamInfEnv(psi, H, tol, maxiter)


def _left_cyclethrough(index, fp, H, st):
    """mpohaminfenv.jl:177-195"""
    N = st.N
    for i in range(N):
        nxt = (i + 1) % N
        acc = None
        for j in range(index, -1, -1):
            if not H[i].contains(j, index):
                continue
            t = transfer_left(fp[j][i], _entry_slice(H[i], j, index), st.AL[i], st.AL[i])
            if acc is None:
                acc = t
            else:
                acc.axpy(1.0, t); t.free()
        fp[index][nxt] = acc if acc is not None else fp[index][nxt].zerovector()


def _right_cyclethrough(index, fp, H, st):
    """mpohaminfenv.jl:197-215"""
    N = st.N
    for i in range(N - 1, -1, -1):
        prv = (i - 1) % N
        acc = None
        for j in range(index, len(fp)):
            if not H[i].contains(index, j):
                continue
            t = transfer_right(fp[j][i], _entry_slice(H[i], index, j), st.AR[i], st.AR[i])
            if acc is None:
                acc = t
            else:
                acc.axpy(1.0, t); t.free()
        fp[index][prv] = acc if acc is not None else fp[index][prv].zerovector()


def _gmres(f, b, x0, tol, maxiter):
    """linsolve(flip(tm), b, x0, solver, 1, -1) (mpohaminfenv.jl:95): device GMRES on the transfer-matrix operator"""
    return linsolve(f, b, x0, 1.0, -1.0, tol, 30, maxiter)


def calclw(fp, st: InfiniteMPS, H, tol, maxiter):
    """mpohaminfenv.jl:76-125"""
    N, od = st.N, H.odim
    ctx = st.ctx
    ident = [identity_bond(ctx, st.V[i], st.dtype) for i in range(N)]                   # l_LL(st, i)
    rho = [mul_right_bond(st.CR[i], st.CR[i].adjoint()) for i in range(N)]              # r_LL(st, i) = CR CR'
    fp[0][0] = ident[0].copy()
    if N > 1:
        _left_cyclethrough(0, fp, H, st)
    for i in range(1, od):
        prev = fp[i][0]
        fp[i][0] = prev.zerovector()
        _left_cyclethrough(i, fp, H, st)
        if _isid(H, i):
            tm = regularize(TransferMatrix(st.AL, None, st.AL).flip(), rho[N - 1], ident[0])
            fp[i][0] = _gmres(tm, fp[i][0], prev, tol, maxiter)
            if N > 1:
                _left_cyclethrough(i, fp, H, st)
            for s in range(N):                                                          # mpohaminfenv.jl:102-107
                sc = rho[(s - 1) % N].adjoint().dot(fp[i][s])
                fp[i][s].axpy(-sc, ident[s])
        else:
            if all(sl.contains(i, i) for sl in H.slices):
                tm = TransferMatrix(st.AL, [_entry_slice(H[s], i, i) for s in range(N)], st.AL).flip()
                fp[i][0] = _gmres(tm, fp[i][0], prev, tol, maxiter)
            if N > 1:
                _left_cyclethrough(i, fp, H, st)
    return fp


def calcrw(fp, st: InfiniteMPS, H, tol, maxiter):
    """mpohaminfenv.jl:127-175"""
    N, od = st.N, H.odim
    ctx = st.ctx
    ident = [identity_bond(ctx, st.V[(i + 1) % N], st.dtype) for i in range(N)]         # r_RR(st, i)
    lrr = [mul_right_bond(st.CR[(i - 1) % N].adjoint(), st.CR[(i - 1) % N]) for i in range(N)]   # l_RR(st, i) = CR[i-1]' CR[i-1]
    fp[od - 1][N - 1] = ident[N - 1].copy()
    if N > 1:
        _right_cyclethrough(od - 1, fp, H, st)
    for i in range(od - 2, -1, -1):
        prev = fp[i][N - 1]
        fp[i][N - 1] = prev.zerovector()
        _right_cyclethrough(i, fp, H, st)
        if _isid(H, i):
            tm = regularize(TransferMatrix(st.AR, None, st.AR), lrr[0], ident[N - 1])
            fp[i][N - 1] = _gmres(tm, fp[i][N - 1], prev, tol, maxiter)
            if N > 1:
                _right_cyclethrough(i, fp, H, st)
            for s in range(N):                                                          # mpohaminfenv.jl:153-159
                sc = lrr[(s + 1) % N].adjoint().dot(fp[i][s])
                fp[i][s].axpy(-sc, ident[s])
        else:
            if all(sl.contains(i, i) for sl in H.slices):
                tm = TransferMatrix(st.AR, [_entry_slice(H[s], i, i) for s in range(N)], st.AR)
                fp[i][N - 1] = _gmres(tm, fp[i][N - 1], prev, tol, maxiter)
            if N > 1:
                _right_cyclethrough(i, fp, H, st)
    return fp


def mul_right_bond(A, B):
    """bond * bond"""
    return mul_right(A, B)


# ------------------------------------------------------------------------------------------ read-outs / VUMPS
def expectation_value(st: InfiniteMPS, H, envs: MPOHamInfEnv):
    """expval.jl:111-124: energy density per site"""
    N, od = st.N, H.odim
    ens = np.zeros(N, dtype=np.complex128)
    for i in range(N):
        if envs.dependency is not st:
            envs.recalculate(st)
        rho_adj = mul_right_bond(st.CR[i], st.CR[i].adjoint()).adjoint()
        for j in range(od - 1, -1, -1):
            if not H[i].contains(j, od - 1):
                continue
            apl = transfer_left(envs.lw[j][i], _entry_slice(H[i], j, od - 1), st.AL[i], st.AL[i])
            ens[i] += rho_adj.dot(apl)
    return ens


def calc_galerkin(st: InfiniteMPS, loc, envs):
    """toolbox.jl:17-22 (envs.opp may be a lazy sum: the derivative constructor dispatches, derivatives.jl:298-323)"""
    from .algorithms import _overlap

    class _P:       # what _overlap needs
        pass
    P = _P(); P.ctx, P.pq, P.pspace, P.dtype = st.ctx, st.pq, st.pspace, st.dtype
    h = _inf_ddAC(loc, st, envs.opp, envs)
    ac = h(st.AC[loc]); h.free()
    ac.scale(1.0 / ac.norm())
    AL = st.AL[loc]
    ov = _overlap(P, ac, AL)
    proj = mul_right(AL, ov)
    ac.axpy(-1.0, proj)
    return ac.norm()


@dataclass
class VUMPS:
    """vumps.jl:18-27 with the default dynamic tolerances of defaults.jl:38-57"""
    tol: float = 1e-12
    maxiter: int = 100
    verbosity: int = 0
    finalize: object = None


def find_groundstate_vumps(st: InfiniteMPS, H, alg: VUMPS = None, envs=None, log=None):
    """find_groundstate(psi::InfiniteMPS, H, alg::VUMPS, envs) (vumps.jl:29-92)"""
    alg = alg or VUMPS()
    if envs is None:
        from .lazyops import _environments as _lazy_environments
        envs = _lazy_environments(st, H)                             # MPOHamInfEnv, or MultipleEnvironments for a LazySum
    N = st.N
    eps = max(calc_galerkin(st, i, envs) for i in range(N))
    from .iterlog import IterLog
    ilog = IterLog("VUMPS", alg.verbosity)                       # vumps.jl:33-35
    obj = lambda: np.sum(expectation_value_lazy(st, H, envs))
    ilog.init(eps, obj)
    for it in range(1, alg.maxiter + 1):
        etol = dyn_tol(eps, it, 1e-12, 1e-5, 1e-5)
        eig = Lanczos(tol=etol, krylovdim=30, maxiter=100)
        newAL = []
        for loc in range(N):                                       # _vumps_localupdate (vumps.jl:77-92)
            hac = _inf_ddAC(loc, st, H, envs)
            _, ac = fixedpoint(hac, st.AC[loc], "SR", eig); hac.free()
            hc = _inf_ddC(loc, st, H, envs)
            _, c = fixedpoint(hc, st.CR[loc], "SR", eig); hc.free()
            newAL.append(regauge(ac, c))
        gtol = dyn_tol(eps, it, TOLGAUGE, 1e-5, 1e-8)
        st = InfiniteMPS.from_AL(newAL, st.CR[N - 1], st.V, st.pspace, st.ctx, gtol, 100)     # vumps.jl:52
        envs.recalculate(st, dyn_tol(eps, it, 1e-12, 1e-5, 1e-5))                               # vumps.jl:54-55
        if alg.finalize is not None:
            st, envs = alg.finalize(it, st, H, envs)
        eps = max(calc_galerkin(st, i, envs) for i in range(N))
        if log is not None:
            log.append(dict(iter=it, eps=eps, energy=float(np.sum(expectation_value(st, H, envs)).real)))
        if eps <= alg.tol:
            ilog.finish(it, eps, obj)                            # vumps.jl:63-66
            break
        if it == alg.maxiter:
            ilog.cancel(it, eps, obj)
        else:
            ilog.iteration(it, eps, obj)
    return st, envs, eps


def _inf_ddAC(loc, st, H, envs):
    """ddAC(loc, psi::InfiniteMPS, H, envs) (derivatives.jl:44-46) with the lazy-operator dispatch of derivatives.jl:298-323"""
    from .lazyops import LazySum, MultipliedOperator
    from .derivatives import LazySumOp, MultipliedOp
    if isinstance(H, LazySum):
        return LazySumOp([_inf_ddAC(loc, st, o, e) for o, e in zip(H.ops, envs.envs)])
    if isinstance(H, MultipliedOperator):
        return MultipliedOp(_inf_ddAC(loc, st, H.op, envs), H.f)
    return MPO_ddAC(st.ctx, H[loc], envs.leftenv(loc, st), envs.rightenv(loc, st), st.AC[loc])


def _inf_ddC(loc, st, H, envs):
    """ddC(loc, psi::InfiniteMPS, H, envs) (derivatives.jl:34-36, 298-323)"""
    from .lazyops import LazySum, MultipliedOperator
    from .derivatives import LazySumOp, MultipliedOp
    if isinstance(H, LazySum):
        return LazySumOp([_inf_ddC(loc, st, o, e) for o, e in zip(H.ops, envs.envs)])
    if isinstance(H, MultipliedOperator):
        return MultipliedOp(_inf_ddC(loc, st, H.op, envs), H.f)
    return MPO_ddC(st.ctx, envs.leftenv(loc + 1, st), envs.rightenv(loc, st), st.CR[loc])


def expectation_value_lazy(st: InfiniteMPS, H, envs):
    """expval.jl:182-192 for InfiniteMPS: LazySum -> sum of the parts, UntimedOperator -> f * <op>"""
    from .lazyops import LazySum, MultipliedOperator
    if isinstance(H, LazySum):
        return sum(expectation_value_lazy(st, o, e) for o, e in zip(H.ops, envs.envs))
    if isinstance(H, MultipliedOperator):
        if H.istimed:
            raise ValueError("expectation value of a time-dependent operator: evaluate it at a time first, H(t)")
        return H.f * expectation_value_lazy(st, H.op, envs)
    return expectation_value(st, H, envs)


def timestep_infinite(st: InfiniteMPS, H, t, dt, alg=None, envs=None):
    """timestep(psi::InfiniteMPS, H, t, dt, alg::TDVP, envs; leftorthflag = true) (src/algorithms/timestep/tdvp.jl:21-59):
    AC and CR of every site of the unit cell are evolved on the device (mpsk_expv through `integrate`), regauged
    (ortho.jl:127-131), the uniform gauge is restored and the environments are recalculated."""
    from .fixedpoint import integrate
    from .algorithms import TDVP
    alg = alg or TDVP()
    if envs is None:
        from .lazyops import _environments as _lazy_environments
        envs = _lazy_environments(st, H)
    N = st.N
    newAL = []
    for loc in range(N):
        hac = _inf_ddAC(loc, st, H, envs)
        ac = integrate(hac, st.AC[loc], t, dt, alg.integrator); hac.free()
        hc = _inf_ddC(loc, st, H, envs)
        c = integrate(hc, st.CR[loc], t, dt, alg.integrator); hc.free()
        newAL.append(regauge(ac, c))
    st = InfiniteMPS.from_AL(newAL, st.CR[N - 1], st.V, st.pspace, st.ctx, alg.tolgauge, alg.gaugemaxiter)
    envs.recalculate(st)
    if alg.finalize is not None:
        st, envs = alg.finalize(t, st, H, envs)
    return st, envs


# ------------------------------------------------------------------------------------------ IDMRG1
@dataclass
class IDMRG1:
    """idmrg.jl:13-19"""
    tol: float = 1e-12
    tol_gauge: float = TOLGAUGE
    eigalg: object = field(default_factory=lambda: Lanczos(tol=1e-12))
    maxiter: int = 100
    verbosity: int = 0


class IDMRGEnv:
    """idmrgenv.jl:6-15: manually updated copies of the infinite environments (stacked level vectors)"""

    def __init__(self, psi: InfiniteMPS, env: MPOHamInfEnv):
        if env.dependency is not psi:
            env.recalculate(psi)
        self.opp = env.opp
        self.N = psi.N
        self.lw = [env.leftenv(i, psi).copy() for i in range(psi.N)]
        self.rw = [env.rightenv(i, psi).copy() for i in range(psi.N)]

    def leftenv(self, pos, psi=None):
        return self.lw[pos % self.N]

    def rightenv(self, pos, psi=None):
        return self.rw[pos % self.N]


def update_leftenv(envs: IDMRGEnv, AL_prev, H, pos):
    """update_leftenv!(envs, st, H, pos) (idmrgenv.jl:68-72): lw[pos] = lw[pos-1] * TransferMatrix(AL[pos-1], H[pos-1], AL[pos-1])"""
    envs.lw[pos % envs.N] = transfer_left(envs.lw[(pos - 1) % envs.N], H[pos - 1], AL_prev, AL_prev)


def update_rightenv(envs: IDMRGEnv, AR_next, H, pos):
    """update_rightenv!(envs, st, H, pos) (idmrgenv.jl:63-66): rw[pos] = TransferMatrix(AR[pos+1], H[pos+1], AR[pos+1]) * rw[pos+1]"""
    envs.rw[pos % envs.N] = transfer_right(envs.rw[(pos + 1) % envs.N], H[pos + 1], AR_next, AR_next)


def find_groundstate_idmrg1(ost: InfiniteMPS, H, alg: IDMRG1 = None, oenvs=None, log=None):
    """find_groundstate(ost::InfiniteMPS, H, alg::IDMRG1, oenvs) (idmrg.jl:21-75)"""
    alg = alg or IDMRG1()
    oenvs = oenvs or environments(ost, H)
    ctx, mod, N = ost.ctx, ost.pspace.modulus, ost.N
    AL, AR, AC, CR = list(ost.AL), list(ost.AR), list(ost.AC), list(ost.CR)
    envs = IDMRGEnv(ost, oenvs)
    eps = np.inf
    for it in range(1, alg.maxiter + 1):
        Ccur = CR[N - 1]
        for pos in range(N):                                        # idmrg.jl:35-40
            h = MPO_ddAC(ctx, H[pos], envs.leftenv(pos), envs.rightenv(pos), AC[pos])
            _, AC[pos] = fixedpoint(h, AC[pos], "SR", alg.eigalg); h.free()
            AL[pos], CR[pos] = qrpos(ctx, AC[pos], mod)
            update_leftenv(envs, AL[pos], H, pos + 1)
        for pos in range(N - 1, -1, -1):                            # idmrg.jl:43-51
            h = MPO_ddAC(ctx, H[pos], envs.leftenv(pos), envs.rightenv(pos), AC[pos])
            _, AC[pos] = fixedpoint(h, AC[pos], "SR", alg.eigalg); h.free()
            CR[(pos - 1) % N], AR[pos] = lqpos(ctx, AC[pos], mod)
            update_rightenv(envs, AR[pos], H, pos - 1)
        d = Ccur.copy(); d.axpy(-1.0, CR[N - 1]); eps = d.norm()
        if log is not None:
            log.append(dict(iter=it, eps=eps))
        if eps < alg.tol:
            break
    nst = InfiniteMPS.from_tensors(AR, ost.V, ost.pspace, ctx, tol=alg.tol_gauge)     # idmrg.jl:72
    nenvs = environments(nst, H)
    return nst, nenvs, eps


# ------------------------------------------------------------------------------------------ IDMRG2
def bond_inverse(Cm):
    """inv(C) of a bond tensor on the device (idmrg.jl:130-134,172-173)"""
    h = C.c_void_p()
    Cm.ctx.check(Cm.ctx.lib.mpsk_bond_inverse(Cm.ctx.h, Cm.h, C.byref(h)))
    out = DeviceTensor.wrap(Cm.ctx, h, Cm.Vl.modulus)
    out._is_bond = True
    return out


@dataclass
class IDMRG2:
    """idmrg.jl:84-91"""
    tol: float = 1e-12
    tol_gauge: float = TOLGAUGE
    eigalg: object = field(default_factory=lambda: Lanczos(tol=1e-12))
    maxiter: int = 100
    verbosity: int = 0
    trscheme: tuple = None


def find_groundstate_idmrg2(ost: InfiniteMPS, H, alg: IDMRG2 = None, oenvs=None, log=None):
    """find_groundstate(ost::InfiniteMPS, H, alg::IDMRG2, oenvs) (idmrg.jl:92-204): two-site updates through the unit cell
    and across its edge, every tensor operation on the device (two-site matvec + Lanczos, truncated SVD, bond-matrix inverse,
    manual environment updates)."""
    from .dmrg import tsvd, contract_ac2, truncerr
    from .derivatives import MPO_ddAC2
    alg = alg or IDMRG2()
    trscheme = alg.trscheme if alg.trscheme is not None else truncerr(1e-6)
    N = ost.N
    if N < 2:
        raise ValueError("unit cell should be >= 2")
    oenvs = oenvs or environments(ost, H)
    ctx = ost.ctx
    AL, AR, AC, CR = list(ost.AL), list(ost.AR), list(ost.AC), list(ost.CR)
    envs = IDMRGEnv(ost, oenvs)
    lw, rw = envs.lw, envs.rw

    def update(pos, ac2):
        """eigensolve + tsvd of the two-site tensor on sites pos, pos + 1 (periodic)"""
        h = MPO_ddAC2(ctx, H[pos], H[pos + 1], lw[pos % N], rw[(pos + 1) % N], ac2)
        val, new = fixedpoint(h, ac2, "SR", alg.eigalg); h.free()
        al, c, ar, _ = tsvd(ost, new, trscheme)                     # normalize!(c) inside
        new.free()
        return al, c, ar, val

    def tl(i):      # update_leftenv!(envs, psi, H, i + 1)
        lw[(i + 1) % N] = transfer_left(lw[i % N], H[i], AL[i % N], AL[i % N])

    def tr(i):      # update_rightenv!(envs, psi, H, i - 1)
        rw[(i - 1) % N] = transfer_right(rw[i % N], H[i], AR[i % N], AR[i % N])
    eps, val = np.inf, 0.0
    for it in range(1, alg.maxiter + 1):
        for pos in range(N - 1):                                                    # idmrg.jl:111-126
            ac2 = contract_ac2(ost, AC[pos], AR[pos + 1])
            al, c, ar, val = update(pos, ac2); ac2.free()
            AL[pos], CR[pos], AR[pos + 1] = al, c, ar
            AC[pos + 1] = mul_left(c, ar)
            tl(pos); tr(pos + 1)
        # edge: sites N-1 and 0 (idmrg.jl:129-148)
        left = mul_right(AC[N - 1], bond_inverse(CR[N - 1]))
        ac2 = contract_ac2(ost, left, mul_right(AL[0], CR[0]))
        al, c, ar, val = update(N - 1, ac2); ac2.free()
        AC[N - 1] = mul_right(al, c); AL[N - 1] = al; CR[N - 1] = c; AR[0] = ar
        AC[0] = mul_left(c, ar)
        AL[0] = mul_right(AC[0], bond_inverse(CR[0]))
        Ccur = c
        tl(N - 1); tr(0)
        for pos in range(N - 2, -1, -1):                                            # idmrg.jl:151-167
            ac2 = contract_ac2(ost, AL[pos], AC[pos + 1])
            al, c, ar, val = update(pos, ac2); ac2.free()
            AL[pos], CR[pos], AR[pos + 1] = al, c, ar
            AC[pos] = mul_right(al, c)
            AC[pos + 1] = mul_left(c, ar)
            tl(pos); tr(pos + 1)
        # edge (idmrg.jl:170-186)
        left = mul_left(CR[N - 2], AR[N - 1])
        ac2 = contract_ac2(ost, mul_right(left, bond_inverse(CR[N - 1])), AC[0])
        al, c, ar, val = update(N - 1, ac2); ac2.free()
        AR[N - 1] = mul_left(bond_inverse(CR[N - 2]), mul_right(al, c))
        AL[N - 1] = al; CR[N - 1] = c; AR[0] = ar
        AC[0] = mul_left(c, ar)
        tl(N - 1); tr(0)
        # error on the smallest common space (idmrg.jl:189-192); the bond matrices are small: host read-out
        ba, bb = Ccur.to_blocks(), c.to_blocks()
        acc = 0.0
        for jl_a, q in enumerate(Ccur.Vl.charges):
            if q not in c.Vl.charges:
                continue
            x, y = ba[(0, jl_a)], bb[(0, c.Vl.charges.index(q))]
            k = min(x.shape[0], y.shape[0])
            acc += float(np.linalg.norm(y[:k, :k] - x[:k, :k]) ** 2)
        eps = float(np.sqrt(acc))
        if log is not None:
            log.append(dict(iter=it, eps=eps, energy=val.real, D=c.Vl.dim))
        if eps < alg.tol:
            break
    V = [AR[i].Vl for i in range(N)]
    nst = InfiniteMPS.from_tensors(AR, V, ost.pspace, ctx, tol=alg.tol_gauge)       # idmrg.jl:201
    nenvs = environments(nst, H)
    return nst, nenvs, eps
