"""Finite-MPS sweep algorithms of the reference restated on charged dense tensors.

State layout: ``A[l, p, r]`` dense arrays plus one integer charge per basis index of every bond
(``Space.index_charges``) and per physical state.  Symmetric tensors are exactly zero outside the
charge-conserving entries, so dense einsum reproduces what TensorKit computes block-wise.

Restates: FiniteMPS random construction + gauge (src/states/finitemps.jl:143-197,
src/states/orthoview.jl:49-143), FinEnv (src/environments/FinEnv.jl:41-70,114-145), DMRG and
DMRG2 sweeps (src/algorithms/groundstate/dmrg.jl:22-55,80-137), calc_galerkin
(src/algorithms/toolbox.jl:17-22), expectation_value (src/algorithms/expval.jl:92-109),
entanglement_spectrum (toolbox.jl:65-69), TDVP/TDVP2 (src/algorithms/timestep/tdvp.jl:61-151).
"""
from __future__ import annotations
import numpy as np
from .spaces import Space, fuse, infimum, dual
from .blocks import mask3
from .contract import (transfer_left_sparse, transfer_right_sparse, dAC_sparse, dAC2_sparse, dC_dense)
from .krylov import eigsolve_lanczos, exponentiate_lanczos
from .linalg import tsvd, qrpos, lqpos
from .mpo import MPOHamiltonian, SparseMPOSlice


def virtual_spaces(pspace: Space, L: int, maxV, left: Space = None, right: Space = None):
    """finitemps.jl:183-192: full-rank-limited bond spaces. maxV: Space or list of L-1 Spaces."""
    left = left or Space((0,), (1,), pspace.modulus)
    right = right or Space((0,), (1,), pspace.modulus)
    maxVs = list(maxV) if isinstance(maxV, (list, tuple)) else [maxV] * (L - 1)
    V = [None] * (L + 1)
    V[0] = left
    for k in range(1, L):
        V[k] = infimum(fuse(V[k - 1], pspace), maxVs[k - 1])
    V[L] = right
    for k in range(L - 1, 0, -1):
        V[k] = infimum(V[k], fuse(V[k + 1], dual(pspace)))
    return V


def gaussian_bond_space(D, parity, sigma=1.5, qmax=5, modulus=0):
    """SURVEY.md 8d config 3 recipe: n_q = round(D g(q)/sum g), g = exp(-q^2/2 sigma^2), |q| <= qmax,
    fix-up on the q ~ 0 sector so that sum n_q = D.  Charges are 2q; ``parity`` selects integer (0)
    or half-integer (1) q."""
    cs = [c for c in range(-2 * qmax, 2 * qmax + 1) if (c % 2) == parity]
    g = np.array([np.exp(-(c / 2) ** 2 / (2 * sigma ** 2)) for c in cs])
    n = np.rint(D * g / g.sum()).astype(int)
    i0 = int(np.argmin(np.abs(np.array(cs))))
    n[i0] += D - n.sum()
    return Space.from_dict({c: int(k) for c, k in zip(cs, n) if k > 0}, modulus)


class FiniteMPS:
    """gauge bookkeeping of finitemps.jl / orthoview.jl on dense charged tensors (0-based sites)."""

    def __init__(self, As, Vs, pspace: Space, normalize=True):
        # finitemps.jl:143-170: left-canonicalise with QRpos, keep last C
        self.L = len(As)
        self.p = pspace
        self.pq = pspace.index_charges()
        self.V = list(Vs)
        self.AL = [None] * self.L
        self.AR = [None] * self.L
        self.AC = [None] * self.L
        self.CL = [None] * (self.L + 1)   # CL[i] sits to the left of site i; CR[i] = CL[i+1]
        As = [np.array(a) for a in As]
        for i in range(self.L):
            Q, R, newV = self._leftorth(As[i], self.V[i], self.V[i + 1])
            As[i] = Q
            if normalize:
                R = R / np.linalg.norm(R)
            self.V[i + 1] = newV if i < self.L - 1 else self.V[i + 1]
            if i < self.L - 1:
                As[i + 1] = np.einsum('ab,bpc->apc', R, As[i + 1])
        self.AL = As
        self.CL[self.L] = R

    @staticmethod
    def random(rng, pspace, L, maxV, dtype=np.complex128, left=None, right=None):
        V = virtual_spaces(pspace, L, maxV, left, right)
        As = []
        for i in range(L):
            m = mask3(V[i], pspace.index_charges(), V[i + 1])
            a = rng.standard_normal(m.shape)
            if np.dtype(dtype).kind == "c":
                a = a + 1j * rng.standard_normal(m.shape)
            As.append((a * m).astype(dtype))
        return FiniteMPS(As, V, pspace)

    # --- per-sector factorizations on (l p; r) / (l; p r) groupings
    def _rowq(self, Vl):
        return (Vl.index_charges()[:, None] + self.pq[None, :]).reshape(-1)

    def _leftorth(self, A, Vl, Vr):
        Dl, d, Dr = A.shape
        Q, R, newV = qrpos(A.reshape(Dl * d, Dr), self._rowq(Vl), Vr.index_charges(), self.p.modulus)
        return Q.reshape(Dl, d, newV.dim), R, newV

    def _rightorth(self, A, Vl, Vr):
        Dl, d, Dr = A.shape
        colq = (Vr.index_charges()[None, :] - self.pq[:, None]).reshape(-1)  # (p, r) order
        Lm, Q, newV = lqpos(A.reshape(Dl, d * Dr), Vl.index_charges(), colq, self.p.modulus)
        return Lm, Q.reshape(newV.dim, d, Dr), newV

    # --- orthoview.jl getters (lazy gauge)
    def get_CR(self, i):
        """CR[i] = bond matrix right of site i (i = -1 .. L-1); orthoview.jl:49-60"""
        if self.CL[i + 1] is None:
            if i == -1 or self.AL[i] is not None:
                Lm, Q, _ = self._rightorth(self.get_AC(i + 1), self.V[i + 1], self.V[i + 2])
                self.CL[i + 1] = Lm
                self.AR[i + 1] = Q
            else:
                Q, R, _ = self._leftorth(self.get_AC(i), self.V[i], self.V[i + 1])
                self.AL[i] = Q
                self.CL[i + 1] = R
        return self.CL[i + 1]

    def get_AC(self, i):
        """orthoview.jl:95-106"""
        if self.AC[i] is None and self.AR[i] is not None:
            self.AC[i] = np.einsum('ab,bpc->apc', self.get_CR(i - 1), self.AR[i])
        elif self.AC[i] is None and self.AL[i] is not None:
            self.AC[i] = np.einsum('apb,bc->apc', self.AL[i], self.get_CR(i))
        return self.AC[i]

    def get_AL(self, i):
        if self.AL[i] is None:
            self.get_AL(i - 1) if i > 0 and self.AL[i - 1] is None else None
            self.get_CR(i)  # triggers leftorth of AC[i] when AL[i] is missing and AL[i-1] present
            if self.AL[i] is None:
                Q, R, _ = self._leftorth(self.get_AC(i), self.V[i], self.V[i + 1])
                self.AL[i], self.CL[i + 1] = Q, R
        return self.AL[i]

    def get_AR(self, i):
        if self.AR[i] is None:
            if i < self.L - 1 and self.AR[i + 1] is None:
                self.get_AR(i + 1)
            Lm, Q, _ = self._rightorth(self.get_AC(i), self.V[i], self.V[i + 1])
            self.CL[i], self.AR[i] = Lm, Q
        return self.AR[i]

    # --- orthoview.jl:108-143 setters
    def _invalidate(self, i):
        self.AC = [None] * self.L
        self.CL = [None] * (self.L + 1)
        for k in range(i, self.L):
            self.AL[k] = None
        for k in range(0, i + 1):
            self.AR[k] = None

    def set_AC(self, i, vec):
        if self.AC[i] is None:
            if i < self.L - 1:
                self.get_AR(i + 1)
            if i > 0:
                self.get_AL(i - 1)
        self._invalidate(i)
        self.AC[i] = vec

    def set_AC_alc(self, i, al, c, newV):
        """ψ.AC[i] = (al, c): orthoview.jl:121-143 (b is the bond tensor)"""
        if self.AC[i] is None:
            if i < self.L - 1:
                self.get_AR(i + 1)
            if i > 0:
                self.get_AL(i - 1)
        self._invalidate(i)
        self.V[i + 1] = newV
        self.CL[i + 1] = c
        self.AL[i] = al

    def set_AC_car(self, i, c, ar, newV):
        """ψ.AC[i] = (c, ar)"""
        if self.AC[i] is None:
            if i < self.L - 1:
                self.get_AR(i + 1)
            if i > 0:
                self.get_AL(i - 1)
        self._invalidate(i)
        self.V[i] = newV
        self.CL[i] = c
        self.AR[i] = ar

    def entanglement_spectrum(self, i):
        """toolbox.jl:65-69: per-sector singular values of CR[i]."""
        C = self.get_CR(i)
        V = self.V[i + 1]
        q = V.index_charges()
        return {c: np.linalg.svd(C[np.ix_(q == c, q == c)], compute_uv=False) for c in V.charges}


class FinEnv:
    """FinEnv.jl:9-145 with identity-based staleness (`===`)."""

    def __init__(self, psi: FiniteMPS, H: MPOHamiltonian, above: FiniteMPS = None):
        self.opp = H
        self.above = above           # FinEnv.jl:72-106: environments between `above` (ket) and psi (bra) -- approximate
        L = psi.L
        odim = H.odim
        # FinEnv.jl:41-70: l_LL (x) 1 on level 1, r_RR (x) 1 on level odim (MPOHamiltonian)
        D0, DL = psi.V[0].dim, psi.V[L].dim
        ls = np.zeros((D0, odim, D0), dtype=np.complex128); ls[:, 0, :] = np.eye(D0)
        rs = np.zeros((DL, odim, DL), dtype=np.complex128); rs[:, odim - 1, :] = np.eye(DL)
        self.leftenvs = [ls] + [None] * L
        self.rightenvs = [None] * L + [rs]
        self.ldep = [None] * L
        self.rdep = [None] * L

    def rightenv(self, ind, psi):
        """FinEnv.jl:114-129: env contractable with the tensor on site ind"""
        a = None
        for i in range(psi.L - 1, ind, -1):
            if psi.get_AR(i) is not self.rdep[i]:
                a = i
                break
        if a is not None:
            for j in range(a, ind, -1):
                AR = psi.get_AR(j)
                ket = AR if self.above is None else self.above.get_AR(j)
                self.rightenvs[j] = transfer_right_sparse(self.rightenvs[j + 1], self.opp[j], ket, AR)
                self.rdep[j] = AR
        return self.rightenvs[ind + 1]

    def leftenv(self, ind, psi):
        """FinEnv.jl:131-145"""
        a = None
        for i in range(0, ind):
            if psi.get_AL(i) is not self.ldep[i]:
                a = i
                break
        if a is not None:
            for j in range(a, ind):
                AL = psi.get_AL(j)
                ket = AL if self.above is None else self.above.get_AL(j)
                self.leftenvs[j + 1] = transfer_left_sparse(self.leftenvs[j], self.opp[j], ket, AL)
                self.ldep[j] = AL
        return self.leftenvs[ind]


def expectation_value(psi, H, envs):
    """expval.jl:92-109: per-site energies with the half-weight rule for boundary-row/column entries."""
    ens = np.zeros(psi.L, dtype=np.complex128)
    for i in range(psi.L):
        AC = psi.get_AC(i)
        GL, GR = envs.leftenv(i, psi), envs.rightenv(i, psi)
        for (j, k) in H[i].keys():
            if not ((j == 0 and k != 0) or (k == H.odim - 1 and j != H.odim - 1)):
                continue
            cur = np.einsum('ac,cpe,ef,aqf,qp->', GL[:, j, :], AC, GR[:, k, :], np.conj(AC), H[i].op(j, k),
                            optimize=True)
            if not (j == 0 and k == H.odim - 1):
                cur = cur / 2
            ens[i] += cur
    n = np.linalg.norm(psi.get_AC(psi.L - 1)) ** 2
    return ens / n


def calc_galerkin(psi, pos, envs):
    """toolbox.jl:17-22"""
    AC = psi.get_AC(pos)
    y = dAC_sparse(AC, envs.opp[pos], envs.leftenv(pos, psi), envs.rightenv(pos, psi))
    y = y / np.linalg.norm(y)
    AL = psi.get_AL(pos)
    proj = np.einsum('apb,cqb,cqd->apd', AL, np.conj(AL), y, optimize=True)
    return float(np.linalg.norm(y - proj))


def ac2_space_charges(psi, pos):
    Vl, Vr = psi.V[pos], psi.V[pos + 2]
    rowq = (Vl.index_charges()[:, None] + psi.pq[None, :]).reshape(-1)          # (l, p1)
    colq = (Vr.index_charges()[:, None] - psi.pq[None, :]).reshape(-1)          # (r, p2)
    return rowq, colq


def dmrg2_sweeps(psi: FiniteMPS, H, nsweeps=1, trunc=("truncerr", 1e-6), tol=1e-12, krylovdim=30,
                 maxiter=100, envs=None, log=None):
    """dmrg.jl:80-137 (fixed number of outer iterations unless converged to `tol`)."""
    envs = envs or FinEnv(psi, H)
    L = psi.L
    eps_all = 0.0
    for it in range(nsweeps):
        eps_s = np.zeros(L)
        for direction, rng_ in (("lr", range(0, L - 1)), ("rl", range(L - 3, -1, -1))):
            for pos in rng_:
                if direction == "lr":
                    ac2 = np.einsum('apk,kqr->aprq', psi.get_AC(pos), psi.get_AR(pos + 1))
                else:
                    ac2 = np.einsum('apk,kqr->aprq', psi.get_AL(pos), psi.get_AC(pos + 1))
                GL, GR = envs.leftenv(pos, psi), envs.rightenv(pos + 1, psi)
                f = lambda x: dAC2_sparse(x, H[pos], H[pos + 1], GL, GR)
                val, new, info = eigsolve_lanczos(f, ac2, "SR", tol, krylovdim, maxiter)
                Dl, d1, Dr, d2 = new.shape
                rowq, colq = ac2_space_charges(psi, pos)
                U, S, Vh, newV, eps_tr, sinfo = tsvd(new.reshape(Dl * d1, Dr * d2), rowq, colq, trunc,
                                                     psi.p.modulus)
                S = S / np.linalg.norm(S)
                al = U.reshape(Dl, d1, newV.dim)
                ar_t = Vh.reshape(newV.dim, Dr, d2)                         # [k; r p2]
                c = np.diag(S).astype(np.complex128)
                v = np.einsum('aprq,apk,k,krq->', ac2, np.conj(al), S, np.conj(ar_t), optimize=True)
                eps_s[pos] = max(eps_s[pos], abs(1 - abs(v)))
                ar = ar_t.transpose(0, 2, 1)                                # _transpose_front: [k p2; r]
                if direction == "lr":
                    psi.set_AC_alc(pos, al, c, newV)
                    psi.set_AC_car(pos + 1, c, ar, newV)
                else:
                    psi.set_AC_car(pos + 1, c, ar, newV)
                    psi.set_AC_alc(pos, al, c, newV)
                if log is not None:
                    log.append(dict(sweep=it, dir=direction, pos=pos, energy=float(np.real(val)),
                                    eps_trunc=eps_tr, kept=dict(sinfo["kept"]), numops=info["numops"],
                                    spectrum={q: s.copy() for q, s in sinfo["spectrum"].items()}))
        eps_all = eps_s.max()
        if eps_all <= tol:
            break
    return psi, envs, eps_all


def dmrg_sweeps(psi: FiniteMPS, H, nsweeps=1, tol=1e-12, krylovdim=30, maxiter=100, envs=None, log=None):
    """dmrg.jl:22-55 single-site DMRG."""
    envs = envs or FinEnv(psi, H)
    L = psi.L
    eps = 0.0
    for it in range(nsweeps):
        eps_s = np.zeros(L)
        for pos in list(range(0, L - 1)) + list(range(L - 1, 0, -1)):
            GL, GR = envs.leftenv(pos, psi), envs.rightenv(pos, psi)
            f = lambda x: dAC_sparse(x, H[pos], GL, GR)
            val, vec, info = eigsolve_lanczos(f, psi.get_AC(pos), "SR", tol, krylovdim, maxiter)
            eps_s[pos] = max(eps_s[pos], calc_galerkin(psi, pos, envs))
            psi.set_AC(pos, vec)
            if log is not None:
                log.append(dict(sweep=it, pos=pos, energy=float(np.real(val)), numops=info["numops"]))
        eps = eps_s.max()
        if eps <= tol:
            break
    return psi, envs, eps


def tdvp_step(psi: FiniteMPS, H, dt, envs=None, tol=1e-12, krylovdim=30):
    """tdvp.jl:61-91: one finite TDVP sweep (left->right then right->left, half steps)."""
    envs = envs or FinEnv(psi, H)
    L = psi.L
    for i in range(0, L - 1):
        GL, GR = envs.leftenv(i, psi), envs.rightenv(i, psi)
        f = lambda x: dAC_sparse(x, H[i], GL, GR)
        ac, _ = exponentiate_lanczos(f, -1j * dt / 2, psi.get_AC(i), tol, krylovdim)
        psi.set_AC(i, ac)
        GLn, GRn = envs.leftenv(i + 1, psi), envs.rightenv(i, psi)
        fc = lambda c: dC_dense(c, GLn, GRn)
        c, _ = exponentiate_lanczos(fc, +1j * dt / 2, psi.get_CR(i), tol, krylovdim)
        psi_set_CR(psi, i, c)
    GL, GR = envs.leftenv(L - 1, psi), envs.rightenv(L - 1, psi)
    f = lambda x: dAC_sparse(x, H[L - 1], GL, GR)
    ac, _ = exponentiate_lanczos(f, -1j * dt / 2, psi.get_AC(L - 1), tol, krylovdim)
    psi.set_AC(L - 1, ac)
    for i in range(L - 1, 0, -1):
        GL, GR = envs.leftenv(i, psi), envs.rightenv(i, psi)
        f = lambda x: dAC_sparse(x, H[i], GL, GR)
        ac, _ = exponentiate_lanczos(f, -1j * dt / 2, psi.get_AC(i), tol, krylovdim)
        psi.set_AC(i, ac)
        GLn, GRn = envs.leftenv(i, psi), envs.rightenv(i - 1, psi)
        fc = lambda c: dC_dense(c, GLn, GRn)
        c, _ = exponentiate_lanczos(fc, +1j * dt / 2, psi.get_CR(i - 1), tol, krylovdim)
        psi_set_CR(psi, i - 1, c)
    GL, GR = envs.leftenv(0, psi), envs.rightenv(0, psi)
    f = lambda x: dAC_sparse(x, H[0], GL, GR)
    ac, _ = exponentiate_lanczos(f, -1j * dt / 2, psi.get_AC(0), tol, krylovdim)
    psi.set_AC(0, ac)
    return psi, envs


def psi_set_CR(psi: FiniteMPS, i, vec):
    """orthoview.jl:62-77 CRView setindex!"""
    if psi.CL[i + 1] is None:
        psi.get_CR(i)
    if psi.AL[i] is None:
        psi.get_AL(i)
    if i + 1 < psi.L and psi.AR[i + 1] is None:
        psi.get_AR(i + 1)
    psi.CL = [None] * (psi.L + 1)
    psi.AC = [None] * psi.L
    for k in range(i + 1, psi.L):
        psi.AL[k] = None
    for k in range(0, i + 1):
        psi.AR[k] = None
    psi.CL[i + 1] = vec


def tdvp2_step(psi: FiniteMPS, H, dt, envs=None, trunc=("truncerr", 1e-3), tol=1e-12, krylovdim=30):
    """tdvp.jl:115-151: one two-site TDVP sweep (no normalisation of the singular values)."""
    envs = envs or FinEnv(psi, H)
    L = psi.L

    def two_site(i, ac2, sub_dt):
        GL, GR = envs.leftenv(i, psi), envs.rightenv(i + 1, psi)
        f = lambda x: dAC2_sparse(x, H[i], H[i + 1], GL, GR)
        new, _ = exponentiate_lanczos(f, -1j * sub_dt, ac2, tol, krylovdim)
        Dl, d1, Dr, d2 = new.shape
        rowq, colq = ac2_space_charges(psi, i)
        U, S, Vh, newV, eps, _ = tsvd(new.reshape(Dl * d1, Dr * d2), rowq, colq, trunc, psi.p.modulus)
        al = U.reshape(Dl, d1, newV.dim)
        ar = Vh.reshape(newV.dim, Dr, d2).transpose(0, 2, 1)
        c = np.diag(S).astype(np.complex128)
        psi.set_AC_alc(i, al, c, newV)
        psi.set_AC_car(i + 1, c, ar, newV)

    def one_site_back(i, sub_dt):
        GL, GR = envs.leftenv(i, psi), envs.rightenv(i, psi)
        f = lambda x: dAC_sparse(x, H[i], GL, GR)
        ac, _ = exponentiate_lanczos(f, +1j * sub_dt, psi.get_AC(i), tol, krylovdim)
        psi.set_AC(i, ac)
    for i in range(0, L - 1):
        ac2 = np.einsum('apk,kqr->aprq', psi.get_AC(i), psi.get_AR(i + 1))
        two_site(i, ac2, dt / 2)
        if i != L - 2:
            one_site_back(i + 1, dt / 2)
    for i in range(L - 1, 0, -1):
        ac2 = np.einsum('apk,kqr->aprq', psi.get_AL(i - 1), psi.get_AC(i))
        two_site(i - 1, ac2, dt / 2)
        if i != 1:
            one_site_back(i - 1, dt / 2)
    return psi, envs


def approximate(psi: FiniteMPS, O, above: FiniteMPS, nsweeps=1, trunc=("truncerr", 1e-6), two_site=True, tol=1e-12, log=None):
    """approximate!(psi, (O, above), DMRG2 / DMRG) (src/algorithms/approximate/fvomps.jl:11-114) with ac2_proj / ac_proj
    (src/algorithms/derivatives.jl:209-232): psi <- the best MPS approximation of O |above> in psi's manifold."""
    envs = FinEnv(psi, O, above)
    L = psi.L
    eps = 0.0
    for it in range(nsweeps):
        eps = 0.0
        if two_site:
            for pos in list(range(0, L - 1)) + list(range(L - 3, -1, -1)):
                x = np.einsum('apk,kqr->aprq', above.get_AC(pos), above.get_AR(pos + 1))
                new = dAC2_sparse(x, O[pos], O[pos + 1], envs.leftenv(pos, psi), envs.rightenv(pos + 1, psi))
                Dl, d1, Dr, d2 = new.shape
                rowq, colq = ac2_space_charges(psi, pos)
                U, S, Vh, newV, eps_tr, sinfo = tsvd(new.reshape(Dl * d1, Dr * d2), rowq, colq, trunc, psi.p.modulus)
                old = np.einsum('apk,kqr->aprq', psi.get_AC(pos), psi.get_AR(pos + 1))
                rec = (U * S[None, :]) @ Vh
                eps = max(eps, float(np.linalg.norm(rec.reshape(old.shape) - old) / np.linalg.norm(old)))
                al = U.reshape(Dl, d1, newV.dim)
                ar = Vh.reshape(newV.dim, Dr, d2).transpose(0, 2, 1)
                c = np.diag(S).astype(np.complex128)
                psi.set_AC_alc(pos, al, c, newV)
                psi.set_AC_car(pos + 1, c, ar, newV)
        else:
            for pos in list(range(0, L - 1)) + list(range(L - 1, 0, -1)):
                new = dAC_sparse(above.get_AC(pos), O[pos], envs.leftenv(pos, psi), envs.rightenv(pos, psi))
                eps = max(eps, float(np.linalg.norm(new - psi.get_AC(pos)) / np.linalg.norm(new)))
                psi.set_AC(pos, new)
        if log is not None:
            log.append(dict(iter=it, eps=eps))
        if eps < tol:
            break
    return psi, envs, eps


def _null_blocks(M, rowq, colq):
    """orthonormal basis N (rows x nnull, with a charge per column) of the orthogonal complement of range(M), sector by sector"""
    cols, charges = [], []
    for q in sorted(set(rowq.tolist())):
        ri = np.nonzero(rowq == q)[0]
        ci = np.nonzero(colq == q)[0]
        if len(ci) == 0:
            basis = np.eye(len(ri), dtype=np.complex128)
        else:
            Qf = np.linalg.qr(M[np.ix_(ri, ci)], mode="complete")[0]
            basis = Qf[:, min(len(ri), len(ci)):]
        for k in range(basis.shape[1]):
            v = np.zeros(M.shape[0], dtype=np.complex128); v[ri] = basis[:, k]
            cols.append(v); charges.append(q)
    N = np.stack(cols, axis=1) if cols else np.zeros((M.shape[0], 0), dtype=np.complex128)
    return N, np.asarray(charges, dtype=np.int64)


def optimal_expand(psi: FiniteMPS, H, trunc=("truncdim", 1), envs=None, log=None):
    """changebonds!(psi::FiniteMPS, H, OptimalExpand(trscheme)) (src/algorithms/changebonds/optimalexpand.jl:84-110)"""
    from .spaces import Space
    envs = envs or FinEnv(psi, H)
    L, pq, mod = psi.L, psi.pq, psi.p.modulus
    d = len(pq)
    for i in range(L - 1):
        AC, AR = psi.get_AC(i), psi.get_AR(i + 1)
        Vl, Vk, Vr = psi.V[i], psi.V[i + 1], psi.V[i + 2]
        Dl, Dk, Dr = Vl.dim, Vk.dim, Vr.dim
        rowq = (Vl.index_charges()[:, None] + pq[None, :]).reshape(-1)             # (l, p1)
        colq = (pq[:, None] * -1 + Vr.index_charges()[None, :]).reshape(-1)        # (p2, r): charge of the k index q_r - s_p2
        NL, nlq = _null_blocks(AC.reshape(Dl * d, Dk), rowq, Vk.index_charges())
        NRh, nrq = _null_blocks(AR.reshape(Dk, d * Dr).conj().T, colq, Vk.index_charges())     # columns of NR^H
        if NL.shape[1] == 0 or NRh.shape[1] == 0:
            continue
        ac2 = np.einsum('apk,kqr->aprq', AC, AR)
        AC2 = dAC2_sparse(ac2, H[i], H[i + 1], envs.leftenv(i, psi), envs.rightenv(i + 1, psi))
        Mat = AC2.transpose(0, 1, 3, 2).reshape(Dl * d, d * Dr)                    # (l p1) x (p2 r)
        inter = NL.conj().T @ Mat @ NRh                                           # adjoint(NL) * AC2 * adjoint(NR)
        U, S, Vh, newV, eps, info = tsvd(inter, nlq, nrq, trunc, mod)
        if log is not None:
            log.append(dict(pos=i, added=newV.as_dict(), spectrum={q: s.copy() for q, s in info["spectrum"].items()}))
        ar_re = (Vh @ NRh.conj().T).reshape(newV.dim, d, Dr)                       # V * NR: [n p; r]
        # direct sums, sector by sector (old indices first inside every sector)
        both = Space.from_dict({q: Vk.as_dict().get(q, 0) + newV.as_dict().get(q, 0) for q in set(Vk.charges) | set(newV.charges)}, mod)
        idx_old, idx_new = [], []
        off = 0
        for q, n in zip(both.charges, both.dims):
            no = Vk.as_dict().get(q, 0)
            idx_old += list(range(off, off + no)); idx_new += list(range(off + no, off + n)); off += n
        ACp = np.zeros((Dl, d, both.dim), dtype=np.complex128); ACp[:, :, idx_old] = AC
        ARp = np.zeros((both.dim, d, Dr), dtype=np.complex128); ARp[idx_old] = AR; ARp[idx_new] = ar_re
        Q, R, nV = qrpos(ACp.reshape(Dl * d, both.dim), rowq, both.index_charges(), mod)
        nal = Q.reshape(Dl, d, nV.dim)
        # new bond space of the QR can be smaller than `both` only if a sector is rank deficient; keep R rectangular then
        assert nV.dim == both.dim, "rank-deficient expansion"
        psi.set_AC_alc(i, nal, R, nV)
        psi.set_AC_car(i + 1, R, ARp, nV)
    return psi, envs


def svdcut(psi: FiniteMPS, trunc=("notrunc",), log=None):
    """changebonds!(psi::AbstractFiniteMPS, alg::SvdCut) (src/algorithms/changebonds/svdcut.jl:14-24): right-to-left sweep,
    truncated SVD of every bond matrix CR[i], absorbed into AL[i] / AR[i+1]; normalize!(psi) at the end."""
    mod = psi.p.modulus
    for i in range(psi.L - 2, -1, -1):
        AL = psi.get_AL(i)
        # CR[i] through the getter's own route (orthoview.jl:49-60): LQpos of AC[i+1]; after a truncation of the bond to the
        # right the L factor is rectangular (V[i+1] x V'), which the square-matrix getter of this port does not track
        Cm, AR, Vp = psi._rightorth(psi.get_AC(i + 1), psi.V[i + 1], psi.V[i + 2])
        U, S, Vh, newV, eps, info = tsvd(Cm, psi.V[i + 1].index_charges(), Vp.index_charges(), trunc, mod)
        if log is not None:
            log.append(dict(pos=i, kept=newV.as_dict(), eps_trunc=eps, spectrum={c: s.copy() for c, s in info["spectrum"].items()}))
        al = np.einsum('apk,kn->apn', AL, U)
        ar = np.einsum('nk,kpr->npr', Vh, AR)
        c = np.diag(S).astype(np.complex128)
        psi.set_AC_alc(i, al, c, newV)
        psi.set_AC_car(i + 1, c, ar, newV)
    c = psi.get_CR(0)
    psi.CL[1] = c / np.linalg.norm(c)
    psi.AC = [None] * psi.L
    return psi


# ---------------------------------------------------------------------------------------------------------------------
# ProjectionOperator / AC_EffProj / LinearCombination / FiniteExcited
def unit_mpo(psi: FiniteMPS):
    """one-level identity MPO: environments of a ProjectionOperator are plain overlap environments between the state (bra)
    and opp.ket (ket) -- FinEnv.jl:101-105 builds them with opp = fill(nothing) and the trivial start vectors l_LL (x) l_LL"""
    return MPOHamiltonian([SparseMPOSlice([[1.0]], (0,), (0,), psi.pq)])


def proj_vector(pos, psi, ket: FiniteMPS, penv: "FinEnv"):
    """|p> = le * ket.AC[pos] * re in the bond spaces of psi: the vector AC_EffProj projects on (derivatives.jl:249-265:
    y = conj(v) * (le a1 re), v = <x | le a1 re>, i.e. y = |p><p|x>)"""
    return dAC_sparse(ket.get_AC(pos), penv.opp[pos], penv.leftenv(pos, psi), penv.rightenv(pos, psi))


def finite_excited(H, states, init: FiniteMPS, weight=10.0, nsweeps=1, tol=1e-12, krylovdim=30, maxiter=100, log=None):
    """excitations(H, FiniteExcited(gsalg = DMRG(), weight), states; init) for ONE new state
    (src/algorithms/excitation/dmrgexcitation.jl:16-36): single-site DMRG (dmrg.jl:22-55) on
    LinearCombination((H, ProjectionOperator.(states)...), (1.0, weight, ...)), whose local operator is
    dAC(H) + weight * sum_k |p_k><p_k| (derivatives.jl:234-247, 249-281) with the environments of a LazyLincoCache
    (lazylincocache.jl:1-8).  Returns (energy of the new state under H, new state)."""
    psi = init
    envH = FinEnv(psi, H)
    unit = unit_mpo(psi)
    penvs = [FinEnv(psi, unit, ket) for ket in states]
    L = psi.L

    def local_op(pos):
        GL, GR = envH.leftenv(pos, psi), envH.rightenv(pos, psi)
        ps = [proj_vector(pos, psi, ket, pe) for ket, pe in zip(states, penvs)]
        return lambda x: dAC_sparse(x, H[pos], GL, GR) + weight * sum(p * np.vdot(p, x) for p in ps)

    def galerkin(pos):
        y = local_op(pos)(psi.get_AC(pos))
        y = y / np.linalg.norm(y)
        AL = psi.get_AL(pos)
        return float(np.linalg.norm(y - np.einsum('apb,cqb,cqd->apd', AL, np.conj(AL), y, optimize=True)))
    eps = 0.0
    for it in range(nsweeps):
        eps_s = np.zeros(L)
        for pos in list(range(0, L - 1)) + list(range(L - 1, 0, -1)):
            val, vec, info = eigsolve_lanczos(local_op(pos), psi.get_AC(pos), "SR", tol, krylovdim, maxiter)
            eps_s[pos] = max(eps_s[pos], galerkin(pos))
            psi.set_AC(pos, vec)
            if log is not None:
                log.append(dict(sweep=it, pos=pos, energy=float(np.real(val)), numops=info["numops"]))
        eps = eps_s.max()
        if eps <= tol:
            break
    return float(np.real(sum(expectation_value(psi, H, FinEnv(psi, H))))), psi


# ------------------------------------------------------------------------------------------ FiniteMPS as a vector
def _site_tensors(psi: FiniteMPS):
    """(AC[0], AR[1], ..., AR[L-1]): site tensors whose product is the state including its norm"""
    return [psi.get_AC(0)] + [psi.get_AR(i) for i in range(1, psi.L)]


def mps_dot(a: FiniteMPS, b: FiniteMPS):
    """dot(psi1, psi2) (src/states/finitemps.jl:459-464), antilinear in the first argument"""
    assert a.L == b.L, "MPS with different length"
    v = np.ones((1, 1), dtype=np.complex128)
    for x, y in zip(_site_tensors(a), _site_tensors(b)):
        v = np.einsum('ab,apc,bpd->cd', v, x.conj(), y)
    return complex(np.trace(v))


def mps_norm(psi: FiniteMPS):
    """norm(psi) = norm(psi.AC[1]) (finitemps.jl:467)"""
    return float(np.linalg.norm(psi.get_AC(0)))


def mps_scale(psi: FiniteMPS, c):
    """psi * a / a * psi (finitemps.jl:372-373, 445-455): every AC and C of the state carries the factor"""
    T = _site_tensors(psi)
    T[0] = T[0] * c
    return FiniteMPS(T, list(psi.V), psi.p, normalize=False)


def mps_add(a: FiniteMPS, b: FiniteMPS):
    """psi1 + psi2 (finitemps.jl:375-441): direct sum of the bond spaces -- first tensor [A1 A2], inner tensors block
    diagonal, last tensor [A1; A2] -- then the canonical form (whose QR chain also cuts the summed bond spaces back to full rank,
    as the leftorth! / rightorth! steps of the reference do)."""
    from .changebonds import _sum_space
    assert a.L == b.L and a.L > 1
    mod = a.p.modulus
    Ta, Tb = _site_tensors(a), _site_tensors(b)
    d = len(a.pq)
    sums = [None] + [_sum_space(a.V[i], b.V[i], mod) for i in range(1, a.L)] + [None]
    out, V = [], [a.V[0]]
    for i in range(a.L):
        lo = (np.arange(a.V[i].dim), np.arange(b.V[i].dim)) if sums[i] is None else sums[i][1:]
        ro = (np.arange(a.V[i + 1].dim), np.arange(b.V[i + 1].dim)) if sums[i + 1] is None else sums[i + 1][1:]
        Dl = a.V[i].dim if sums[i] is None else sums[i][0].dim
        Dr = a.V[i + 1].dim if sums[i + 1] is None else sums[i + 1][0].dim
        t = np.zeros((Dl, d, Dr), dtype=np.complex128)
        t[np.ix_(lo[0], range(d), ro[0])] += Ta[i]
        t[np.ix_(lo[1], range(d), ro[1])] += Tb[i]
        out.append(t)
        V.append(a.V[i + 1] if sums[i + 1] is None else sums[i + 1][0])
    for i in range(a.L - 1, 0, -1):               # right-to-left LQ sweep: cuts the summed bonds on the right half
        Lm, Q, newV = a._rightorth(out[i], V[i], V[i + 1])
        out[i], V[i] = Q, newV
        out[i - 1] = np.einsum('apb,bc->apc', out[i - 1], Lm)
    return FiniteMPS(out, V, a.p, normalize=False)
