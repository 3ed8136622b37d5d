"""Host-driven Krylov solvers over the VectorInterface of device tensors (inner, norm, add!!, scale!!), the way
KrylovKit drives MPSKit in the reference: the vectors and every matvec live on the device, only the Krylov scalars
and the small projected problem come to the host.  Used where the operator is not a single `mpsk_op`:
non-Hermitian maps (Arnoldi: src/algorithms/fixedpoint.jl:19-30, transfer-matrix fixed points ortho.jl:184,241,
permpoinfenv.jl:157,168), sums of operators (LazySum, derivatives.jl:283-323), linear systems
(GMRES: mpohaminfenv.jl:95,113,146,164).  The Hermitian single-operator case uses the library's device-resident
Lanczos (mpsk_eigsolve) instead."""
from __future__ import annotations
import numpy as np


class RecursiveVec:
    """KrylovKit.RecursiveVec: a tuple of vectors behaving as one vector (statmech/vumps.jl:33-36)."""

    def __init__(self, vecs):
        self.vecs = list(vecs)

    def copy(self):
        return RecursiveVec([v.copy() for v in self.vecs])

    def zerovector(self):
        return RecursiveVec([v.zerovector() for v in self.vecs])

    def dot(self, o):
        return sum(a.dot(b) for a, b in zip(self.vecs, o.vecs))

    def norm(self):
        return float(np.sqrt(sum(v.norm() ** 2 for v in self.vecs)))

    def axpy(self, a, x):
        for v, w in zip(self.vecs, x.vecs):
            v.axpy(a, w)
        return self

    def scale(self, a):
        for v in self.vecs:
            v.scale(a)
        return self

    def free(self):
        for v in self.vecs:
            v.free()


def _orth(w, V, passes=2):
    """classical Gram-Schmidt with one refinement pass (KrylovKit ModifiedGramSchmidtIR quality)"""
    h = np.zeros(len(V), dtype=np.complex128)
    for _ in range(passes):
        c = np.array([complex(v.dot(w)) for v in V])
        for ci, v in zip(c, V):
            w.axpy(-ci, v)
        h += c
    return w, h


def _combine(coef, V):
    x = V[0].copy().scale(coef[0])
    for c, v in zip(coef[1:], V[1:]):
        x.axpy(c, v)
    return x


def _select(ev, which):
    if which == "LM":
        return int(np.argmax(np.abs(ev)))
    if which == "SR":
        return int(np.argmin(ev.real))
    if which == "LR":
        return int(np.argmax(ev.real))
    raise ValueError(which)


def eigsolve_arnoldi(f, x0, which="LM", tol=1e-12, krylovdim=30, maxiter=100):
    """one extremal eigenpair of a general linear map; restarted Arnoldi with eager convergence test."""
    x = x0.copy()
    x.scale(1.0 / x.norm())
    numops = 0
    val, res = 0.0, np.inf
    for it in range(1, maxiter + 1):
        V = [x]
        Hm = np.zeros((krylovdim + 1, krylovdim), dtype=np.complex128)
        conv = False
        for k in range(krylovdim):
            w = f(V[k]); numops += 1
            w, h = _orth(w, V)
            Hm[:k + 1, k] = h
            beta = w.norm()
            Hm[k + 1, k] = beta
            ev, U = np.linalg.eig(Hm[:k + 1, :k + 1])
            idx = _select(ev, which)
            val = ev[idx]
            res = abs(beta * U[k, idx])
            if res < tol or beta < 1e-14:
                conv = True
                break
            V.append(w.scale(1.0 / beta))
        y = U[:, idx]
        x = _combine(y, V[:len(y)])
        x.scale(1.0 / x.norm())
        if conv:
            return val, x, dict(converged=1, numiter=it, numops=numops, normres=res)
    return val, x, dict(converged=0, numiter=maxiter, numops=numops, normres=res)


def eigsolve_lanczos(f, x0, which="SR", tol=1e-12, krylovdim=30, maxiter=100):
    """Hermitian maps that are not a single device operator (sums of operators)."""
    x = x0.copy()
    x.scale(1.0 / x.norm())
    numops = 0
    val, res = 0.0, np.inf
    for it in range(1, maxiter + 1):
        V = [x]
        T = np.zeros((krylovdim + 1, krylovdim + 1))
        conv = False
        for k in range(krylovdim):
            w = f(V[k]); numops += 1
            w, h = _orth(w, V)
            T[k, k] = h[k].real
            beta = w.norm()
            T[k + 1, k] = T[k, k + 1] = beta
            ev, U = np.linalg.eigh(T[:k + 1, :k + 1])
            idx = 0 if which == "SR" else (len(ev) - 1 if which == "LR" else int(np.argmax(np.abs(ev))))
            val = ev[idx]
            res = abs(beta * U[k, idx])
            if res < tol or beta < 1e-14:
                conv = True
                break
            V.append(w.scale(1.0 / beta))
        y = U[:, idx]
        x = _combine(y, V[:len(y)])
        x.scale(1.0 / x.norm())
        if conv:
            return val, x, dict(converged=1, numiter=it, numops=numops, normres=res)
    return val, x, dict(converged=0, numiter=maxiter, numops=numops, normres=res)


def linsolve_gmres(f, b, x0, a0=1.0, a1=-1.0, tol=1e-12, krylovdim=30, maxiter=100):
    """solve (a0 + a1 f) x = b with restarted GMRES (KrylovKit.linsolve(f, b, x0, GMRES, a0, a1))."""
    x = x0.copy()

    def op(v):
        y = f(v)
        y.scale(a1)
        y.axpy(a0, v)
        return y
    numops = 0
    beta = np.inf
    for it in range(maxiter):
        r = op(x); numops += 1
        r.scale(-1.0)
        r.axpy(1.0, b)
        beta = r.norm()
        if beta < tol:
            return x, dict(converged=1, numops=numops, normres=beta)
        V = [r.scale(1.0 / beta)]
        Hm = np.zeros((krylovdim + 1, krylovdim), dtype=np.complex128)
        yk = None
        for k in range(krylovdim):
            w = op(V[k]); numops += 1
            w, h = _orth(w, V)
            Hm[:k + 1, k] = h
            hn = w.norm()
            Hm[k + 1, k] = hn
            e1 = np.zeros(k + 2, dtype=np.complex128); e1[0] = beta
            yk, *_ = np.linalg.lstsq(Hm[:k + 2, :k + 1], e1, rcond=None)
            res = np.linalg.norm(Hm[:k + 2, :k + 1] @ yk - e1)
            if res < tol or hn < 1e-14:
                break
            V.append(w.scale(1.0 / hn))
        for c, v in zip(yk, V[:len(yk)]):
            x.axpy(c, v)
    r = op(x); r.scale(-1.0); r.axpy(1.0, b)
    return x, dict(converged=0, numops=numops, normres=r.norm())


def exponentiate_lanczos(f, tau, y0, tol=1e-12, krylovdim=30, maxiter=100):
    """y = exp(tau f) y0 for a Hermitian map f that is not a single device operator (KrylovKit.exponentiate with Lanczos,
    called by integrate, integrators.jl:20-25): Krylov projections of dimension <= krylovdim on the device vectors, the small
    exponential on the host; if the a-posteriori error estimate |beta_m [exp(dtau T)]_{m,1}| of a full-dimension projection
    exceeds the tolerance the step is halved and the remaining interval is continued from the intermediate vector."""
    y = y0.copy()
    done, numops = 0.0, 0
    step = 1.0
    converged = 1
    it = 0
    while done < 1.0 - 1e-15:
        it += 1
        if it > maxiter:
            converged = 0
            break
        beta0 = y.norm()
        if beta0 == 0.0:
            break
        V = [y.copy().scale(1.0 / beta0)]
        alphas, betas = [], []
        w = None
        happy = False
        for k in range(krylovdim):
            w = f(V[k]); numops += 1
            w, h = _orth(w, V)
            alphas.append(h[k].real)
            b = w.norm()
            betas.append(b)
            if b < 1e-14 * max(1.0, abs(alphas[-1])):
                happy = True
                break
            if k + 1 < krylovdim:
                V.append(w.scale(1.0 / b))
        m = len(alphas)
        T = np.diag(alphas) + np.diag(betas[:m - 1], 1) + np.diag(betas[:m - 1], -1)
        ev, U = np.linalg.eigh(T)
        while True:
            dt_ = min(step, 1.0 - done)
            c = U @ (np.exp(tau * dt_ * ev) * U[0, :].conj())
            err = 0.0 if happy else abs(betas[m - 1] * c[m - 1]) * beta0
            if err <= tol * dt_ or dt_ < 1e-6:
                if err > tol * dt_:
                    converged = 0
                break
            step *= 0.5
        ynew = _combine(c * beta0, V[:m])
        for v in V:
            v.free() if hasattr(v, "free") else None
        y.free() if hasattr(y, "free") else None
        y = ynew
        done += dt_
    return y, dict(converged=converged, numops=numops)
