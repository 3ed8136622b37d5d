"""Krylov solvers restating the KrylovKit calls on the path (not in /root/reference: KrylovKit
0.4-0.7 per Project.toml:24-41; call sites src/algorithms/fixedpoint.jl:9-30,
src/algorithms/timestep/integrators.jl:20-25, src/environments/mpohaminfenv.jl:95-164).

Published algorithms restated: Lanczos with full re-orthogonalisation and thick restart
(``eigsolve(..., Lanczos)``), Arnoldi with Krylov-Schur style restart (``schursolve``), Lanczos
``exponentiate`` with sub-stepping, restarted GMRES ``linsolve(f,b,x0,GMRES,a0,a1)``.
North-star parity is on converged quantities, so iteration counts are
reported but not matched.  Vectors are numpy arrays of any shape; inner product is vdot.
"""
from __future__ import annotations
import numpy as np
import scipy.linalg as sla


def _orth(w, V, passes=2):
    h = np.zeros(len(V), dtype=np.result_type(w, np.complex128) if np.iscomplexobj(w) else w.dtype)
    for _ in range(passes):
        for i, v in enumerate(V):
            c = np.vdot(v, w)
            w = w - c * v
            h[i] += c
    return w, h


def eigsolve_lanczos(f, x0, which="SR", tol=1e-12, krylovdim=30, maxiter=100):
    """extremal eigenpair of a Hermitian map. returns (val, vec, info)."""
    x = x0 / np.linalg.norm(x0)
    numops = 0
    val = None
    for it in range(1, maxiter + 1):
        V = [x]
        T = np.zeros((krylovdim + 1, krylovdim + 1))
        k = 0
        conv = False
        for k in range(krylovdim):
            w = f(V[k]); numops += 1
            w, h = _orth(w, V)
            T[k, k] = np.real(h[k])
            beta = np.linalg.norm(w)
            T[k + 1, k] = T[k, k + 1] = beta
            ev, U = np.linalg.eigh(T[:k + 1, :k + 1])
            idx = 0 if which == "SR" else (np.argmax(ev) if which == "LR" else np.argmax(np.abs(ev)))
            val = ev[idx]
            res = abs(beta * U[k, idx])
            if res < tol or beta < 1e-14:
                conv = True
                break
            V.append(w / beta)
        y = U[:, idx]
        x = sum(c * v for c, v in zip(y, V[:len(y)]))
        x = x / np.linalg.norm(x)
        if conv:
            return val, x, dict(converged=1, numiter=it, numops=numops, normres=res)
    return val, x, dict(converged=0, numiter=maxiter, numops=numops, normres=res)


def eigsolve_arnoldi(f, x0, which="LM", tol=1e-12, krylovdim=30, maxiter=100):
    """dominant eigenpair of a general map by restarted Arnoldi (restart with the Ritz vector)."""
    x = x0 / np.linalg.norm(x0)
    numops = 0
    for it in range(1, maxiter + 1):
        V = [x]
        Hm = np.zeros((krylovdim + 1, krylovdim), dtype=np.complex128)
        conv = False
        for k in range(krylovdim):
            w = f(V[k]); numops += 1
            w, h = _orth(w, V)
            Hm[:k + 1, k] = h
            beta = np.linalg.norm(w)
            Hm[k + 1, k] = beta
            ev, U = np.linalg.eig(Hm[:k + 1, :k + 1])
            if which == "LM":
                idx = np.argmax(np.abs(ev))
            elif which == "SR":
                idx = np.argmin(ev.real)
            else:
                idx = np.argmax(ev.real)
            res = abs(beta * U[k, idx])
            if res < tol or beta < 1e-14:
                conv = True
                break
            V.append(w / beta)
        y = U[:, idx]
        x = sum(c * v for c, v in zip(y, V[:len(y)]))
        x = x / np.linalg.norm(x)
        if conv:
            return ev[idx], x, dict(converged=1, numiter=it, numops=numops, normres=res)
    return ev[idx], x, dict(converged=0, numiter=maxiter, numops=numops, normres=res)


def exponentiate_lanczos(f, tau, x0, tol=1e-12, krylovdim=30, maxiter=100):
    """y = exp(tau * f) x0 for Hermitian f (integrators.jl:20-25 calls it with tau = -i dt)."""
    y = np.array(x0, dtype=np.complex128)
    t_done = 0.0
    numops = 0
    total = 1.0
    for it in range(maxiter):
        nrm = np.linalg.norm(y)
        if nrm == 0:
            return y, dict(converged=1, numops=numops)
        V = [y / nrm]
        T = np.zeros((krylovdim + 1, krylovdim + 1))
        m = 0
        for k in range(krylovdim):
            w = f(V[k]); numops += 1
            w, h = _orth(w, V)
            T[k, k] = np.real(h[k])
            beta = np.linalg.norm(w)
            m = k + 1
            if beta < 1e-14:
                break
            T[k + 1, k] = T[k, k + 1] = beta
            V.append(w / beta)
        dt = total - t_done
        while True:
            E = sla.expm(tau * dt * T[:m, :m])
            err = abs(T[m, m - 1] * E[m - 1, 0]) * abs(tau * dt) if m == krylovdim and len(V) > m else 0.0
            if err <= tol * max(dt / total, 1e-3) or dt < 1e-12:
                break
            dt *= 0.5
        y = nrm * sum(c * v for c, v in zip(E[:, 0], V[:m]))
        t_done += dt
        if t_done >= total - 1e-15:
            return y, dict(converged=1, numops=numops)
    return y, dict(converged=0, numops=numops)


def linsolve_gmres(f, b, x0, a0=1.0, a1=-1.0, tol=1e-12, krylovdim=30, maxiter=100):
    """solve (a0 + a1 f) x = b by restarted GMRES (mpohaminfenv.jl:95,113,146,164)."""
    x = np.array(x0, dtype=np.result_type(b, np.complex128))
    op = lambda v: a0 * v + a1 * f(v)
    bn = np.linalg.norm(b)
    numops = 0
    for it in range(maxiter):
        r = b - op(x); numops += 1
        beta = np.linalg.norm(r)
        if beta <= tol * max(bn, 1e-300) or beta < tol:
            return x, dict(converged=1, numops=numops, normres=beta)
        V = [r / beta]
        Hm = np.zeros((krylovdim + 1, krylovdim), dtype=np.complex128)
        for k in range(krylovdim):
            w = op(V[k]); numops += 1
            w, h = _orth(w, V)
            Hm[:k + 1, k] = h
            hn = np.linalg.norm(w)
            Hm[k + 1, k] = hn
            e1 = np.zeros(k + 2, dtype=np.complex128); e1[0] = beta
            yk, *_ = np.linalg.lstsq(Hm[:k + 2, :k + 1], e1, rcond=None)
            res = np.linalg.norm(Hm[:k + 2, :k + 1] @ yk - e1)
            if res < tol or hn < 1e-14:
                break
            V.append(w / hn)
        x = x + sum(c * v for c, v in zip(yk, V[:len(yk)]))
    r = b - op(x)
    return x, dict(converged=0, numops=numops, normres=np.linalg.norm(r))
