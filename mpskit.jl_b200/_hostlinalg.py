"""Host-side per-sector QRpos / LQpos used only by the FiniteMPS *constructor* (state construction is
host-side in the reference too: src/states/finitemps.jl:143-170 runs TensorKit leftorth on the CPU).
Not used inside sweeps."""
from __future__ import annotations
import numpy as np
from .spaces import Space


def qrpos(M, rowq, colq, modulus=0):
    secs = sorted(set(rowq.tolist()) & set(colq.tolist()))
    parts, dims = {}, {}
    for q in secs:
        ri = np.nonzero(rowq == q)[0]; ci = np.nonzero(colq == q)[0]
        Q, R = np.linalg.qr(M[np.ix_(ri, ci)])
        ph = np.diagonal(R).copy()
        ph = np.where(np.abs(ph) > 0, ph / np.where(np.abs(ph) > 0, np.abs(ph), 1), 1.0)
        parts[q] = (ri, ci, Q * ph[None, :], np.conj(ph)[:, None] * R)
        dims[q] = Q.shape[1]
    newsp = Space.from_dict(dims, modulus)
    Qf = np.zeros((M.shape[0], newsp.dim), dtype=M.dtype)
    Rf = np.zeros((newsp.dim, M.shape[1]), dtype=M.dtype)
    off = 0
    for q in newsp.charges:
        ri, ci, Q, R = parts[q]
        k = Q.shape[1]
        Qf[np.ix_(ri, np.arange(off, off + k))] = Q
        Rf[np.ix_(np.arange(off, off + k), ci)] = R
        off += k
    return Qf, Rf, newsp


def lqpos(M, rowq, colq, modulus=0):
    Qh, Rh, newsp = qrpos(M.conj().T, colq, rowq, modulus)
    return Rh.conj().T, Qh.conj().T, newsp
