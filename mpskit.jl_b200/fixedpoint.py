"""``fixedpoint`` and ``integrate``: mirrors of src/algorithms/fixedpoint.jl:9-30 and
src/algorithms/timestep/integrators.jl:20-25.  The Krylov loop runs inside the library on the device."""
from __future__ import annotations
import ctypes as C
import warnings
from dataclasses import dataclass
from . import _lib

_WHICH = {"SR": _lib.SR, "LM": _lib.LM, "LR": _lib.LR}


@dataclass
class Lanczos:
    """KrylovKit.Lanczos parameters (Hermitian maps)."""
    tol: float = 1e-12
    krylovdim: int = 30
    maxiter: int = 100


@dataclass
class Arnoldi:
    """KrylovKit.Arnoldi parameters (Defaults.eigsolver, defaults.jl:33). ``hermitian=True`` lets the library use
    the Lanczos recurrence for the Hermitian effective Hamiltonians of DMRG (same converged eigenpair)."""
    tol: float = 1e-12
    krylovdim: int = 30
    maxiter: int = 100
    eager: bool = True
    hermitian: bool = True


def _device_handle(A, x0):
    """mpsk_op handle of A acting on vectors like x0, or None when A is not expressible as one device operator"""
    if hasattr(A, "h") and A.h is not None:
        return A.h
    if hasattr(A, "device_op") and hasattr(x0, "h"):
        return A.device_op(x0).h
    return None


def fixedpoint(A, x0, which, alg, return_info=False):
    """val, vec = fixedpoint(A, x0, which, alg)   (fixedpoint.jl:9-30).  Single device operators (effective Hamiltonians,
    transfer matrices) run inside the library: Lanczos (thick restart) for Hermitian maps, Krylov-Schur Arnoldi otherwise."""
    herm = 1 if isinstance(alg, Lanczos) or getattr(alg, "hermitian", False) else 0
    handle = _device_handle(A, x0) if hasattr(x0, "h") else None
    if handle is None:
        # sums of operators / tuples of vectors (LazySum, RecursiveVec): host-driven Krylov over the VectorInterface
        # (krylov.py), the way KrylovKit drives the reference
        from . import krylov
        solver = krylov.eigsolve_lanczos if herm else krylov.eigsolve_arnoldi
        val, vec, info = solver(A, x0, which, alg.tol, alg.krylovdim, alg.maxiter)
        if info["converged"] == 0:
            warnings.warn(f"fixedpoint not converged after {info['numiter']} iterations: normres = {info['normres']}")
        return (complex(val), vec, info) if return_info else (complex(val), vec)
    ctx = A.ctx
    vec = x0.zerovector()
    re, im = C.c_double(), C.c_double()
    info = _lib.c_info()
    ctx.check(ctx.lib.mpsk_eigsolve(ctx.h, handle, x0.h, _WHICH[which], herm, alg.krylovdim, alg.maxiter, alg.tol,
                                    C.byref(re), C.byref(im), vec.h, C.byref(info)))
    if info.converged == 0:   # fixedpoint.jl:12-14
        warnings.warn(f"fixedpoint not converged after {info.numiter} iterations: normres = {info.normres}")
    val = complex(re.value, im.value)
    if return_info:
        return val, vec, dict(converged=info.converged, numiter=info.numiter, numops=info.numops, normres=info.normres)
    return val, vec


def linsolve(f, b, x0, a0=1.0, a1=-1.0, tol=1e-12, krylovdim=30, maxiter=100, return_info=False):
    """x = linsolve(f, b, x0, GMRES(tol, krylovdim, maxiter), a0, a1): (a0 + a1 f) x = b (mpohaminfenv.jl:95,113,146,164).
    Device transfer-matrix operators are solved by the library's GMRES (mpsk_linsolve); anything else by krylov.py."""
    handle = _device_handle(f, x0) if hasattr(x0, "h") else None
    if handle is None:
        from . import krylov
        x, info = krylov.linsolve_gmres(f, b, x0, a0, a1, tol, krylovdim, maxiter)
    else:
        ctx = x0.ctx
        x = x0.zerovector()
        ci = _lib.c_info()
        a0, a1 = complex(a0), complex(a1)
        ctx.check(ctx.lib.mpsk_linsolve(ctx.h, handle, b.h, x0.h, a0.real, a0.imag, a1.real, a1.imag, krylovdim, maxiter, tol,
                                        x.h, C.byref(ci)))
        info = dict(converged=ci.converged, numops=ci.numops, normres=ci.normres)
    if info["converged"] == 0:
        warnings.warn(f"linsolve failed to converge: normres = {info['normres']}")
    return (x, info) if return_info else x


def integrate(f, y0, t, dt, alg: Lanczos = None, return_info=False):
    """y = exponentiate(f(t + dt/2), -im*dt, y0, alg)   (integrators.jl:20-25); f is a DerivativeOperator."""
    alg = alg or Lanczos(tol=1e-12)
    ctx = f.ctx
    tau = -1j * dt
    if not hasattr(f, "h") or f.h is None:
        from .derivatives import MultipliedOp
        if isinstance(f, MultipliedOp) and getattr(f.op, "h", None) is not None:
            # f(t) * h with h a device operator: the coefficient at t + dt/2 scales the exponent (derivatives.jl:284-285)
            c = f.f(t + dt / 2) if callable(f.f) else f.f
            return integrate(f.op, y0, t, dt * c, alg, return_info)
        # LazySum of derivatives: host-driven exponentiate of x -> f(x, t + dt/2)
        from . import krylov
        y, info = krylov.exponentiate_lanczos(lambda x: f(x, t + dt / 2), tau, y0, alg.tol, alg.krylovdim, alg.maxiter)
        if info["converged"] == 0:
            warnings.warn("integrate not converged")
        return (y, info) if return_info else y
    y = y0.zerovector()
    info = _lib.c_info()
    ctx.check(ctx.lib.mpsk_expv(ctx.h, f.h, tau.real, tau.imag, y0.h, alg.krylovdim, alg.maxiter, alg.tol, y.h, C.byref(info)))
    if info.converged == 0:   # integrators.jl:22-23
        warnings.warn("integrate not converged")
    if return_info:
        return y, dict(converged=info.converged, numops=info.numops)
    return y
