"""Truncated SVD and positive QR/LQ with TensorKit semantics (not in /root/reference: TensorKit
0.12 per Project.toml:24-41; call sites dmrg.jl:96,112, tdvp.jl:124,140, orthoview.jl:52-56,
ortho.jl:128-135).  Published behaviour restated: ``tsvd!`` does one
LAPACK gesvd per coupled sector and truncates globally over all sectors; ``leftorth!(QRpos)`` /
``rightorth!(LQpos)`` do one QR/LQ per coupled sector with a non-negative real diagonal.

Matrices are given densely with an integer charge per row and per column; the coupled sectors are
the common charges.  New bond sectors are in ascending charge order.
"""
from __future__ import annotations
import numpy as np
from .spaces import Space


def _sectors(rowq, colq):
    return sorted(set(rowq.tolist()) & set(colq.tolist()))


def tsvd(M, rowq, colq, trunc=("notrunc",), modulus=0):
    """returns U (rows x Dn), S (Dn,), Vh (Dn x cols), newspace, eps_trunc, info.

    trunc: ("notrunc",) | ("truncdim", D) | ("truncerr", eps) | ("truncbelow", eps).
    truncdim keeps the D largest singular values over all sectors; truncerr drops the smallest
    values while the discarded 2-norm stays <= eps (absolute); truncbelow drops s < eps.
    """
    secs = _sectors(rowq, colq)
    res = {}
    allS = []
    for q in secs:
        ri = np.nonzero(rowq == q)[0]
        ci = np.nonzero(colq == q)[0]
        u, s, vh = np.linalg.svd(M[np.ix_(ri, ci)], full_matrices=False)
        res[q] = (ri, ci, u, s, vh)
        allS += [(sv, q, i) for i, sv in enumerate(s)]
    keep = {q: len(res[q][3]) for q in secs}
    # sort ascending by value; ties: larger sector order dropped first (documented tie rule of this port)
    order = sorted(allS, key=lambda t: (t[0], -secs.index(t[1]), -t[2]))
    kind = trunc[0]
    dropped = []
    if kind == "truncdim":
        excess = len(allS) - int(trunc[1])
        dropped = order[:max(excess, 0)]
    elif kind == "truncerr":
        acc = 0.0
        for t in order:
            if np.sqrt(acc + t[0] ** 2) <= trunc[1]:
                acc += t[0] ** 2
                dropped.append(t)
            else:
                break
    elif kind == "truncbelow":
        dropped = [t for t in order if t[0] < trunc[1]]
    elif kind != "notrunc":
        raise ValueError(kind)
    for (_, q, _) in dropped:
        keep[q] -= 1
    eps = float(np.sqrt(sum(t[0] ** 2 for t in dropped)))
    newsp = Space.from_dict({q: keep[q] for q in secs}, modulus)
    Dn = newsp.dim
    U = np.zeros((M.shape[0], Dn), dtype=M.dtype)
    Vh = np.zeros((Dn, M.shape[1]), dtype=M.dtype)
    S = np.zeros(Dn)
    off = 0
    for q in newsp.charges:
        ri, ci, u, s, vh = res[q]
        k = keep[q]
        U[np.ix_(ri, np.arange(off, off + k))] = u[:, :k]
        Vh[np.ix_(np.arange(off, off + k), ci)] = vh[:k, :]
        S[off:off + k] = s[:k]
        off += k
    return U, S, Vh, newsp, eps, dict(kept={q: keep[q] for q in secs},
                                      spectrum={q: res[q][3] for q in secs})


def qrpos(M, rowq, colq, modulus=0):
    """leftorth!(alg=QRpos()): M = Q R per sector, diag(R) >= 0; new bond sector dim min(m, n)."""
    secs = _sectors(rowq, colq)
    parts = {}
    dims = {}
    for q in secs:
        ri = np.nonzero(rowq == q)[0]
        ci = np.nonzero(colq == q)[0]
        Q, R = np.linalg.qr(M[np.ix_(ri, ci)])
        ph = np.diagonal(R).copy()
        ph = np.where(np.abs(ph) > 0, ph / np.where(np.abs(ph) > 0, np.abs(ph), 1), 1.0)
        Q = Q * ph[None, :]
        R = np.conj(ph)[:, None] * R
        parts[q] = (ri, ci, Q, R)
        dims[q] = Q.shape[1]
    newsp = Space.from_dict(dims, modulus)
    Dn = newsp.dim
    Qf = np.zeros((M.shape[0], Dn), dtype=M.dtype)
    Rf = np.zeros((Dn, M.shape[1]), dtype=M.dtype)
    off = 0
    for q in newsp.charges:
        ri, ci, Q, R = parts[q]
        k = Q.shape[1]
        Qf[np.ix_(ri, np.arange(off, off + k))] = Q
        Rf[np.ix_(np.arange(off, off + k), ci)] = R
        off += k
    return Qf, Rf, newsp


def lqpos(M, rowq, colq, modulus=0):
    """rightorth!(alg=LQpos()): M = L Q with diag(L) >= 0."""
    Qh, Rh, newsp = qrpos(M.conj().T, colq, rowq, modulus)
    return Rh.conj().T, Qh.conj().T, newsp
