"""Infinite-MPS algorithms of the reference restated on dense (charged) tensors: uniform gauge
(src/states/ortho.jl:127-257, src/states/infinitemps.jl:139-204), MPOHamInfEnv (src/environments/mpohaminfenv.jl:56-215),
VUMPS (src/algorithms/groundstate/vumps.jl:29-92), infinite expectation values (src/algorithms/expval.jl:111-124) and
calc_galerkin (src/algorithms/toolbox.jl:23-25).  Sites are 0-based and periodic; environments are stored per MPO level
as bond matrices lw[j][i][bra, ket] (width-1 levels), rw[j][i][ket, bra]."""
from __future__ import annotations
import numpy as np
from .spaces import Space
from .linalg import qrpos, lqpos
from .krylov import eigsolve_arnoldi, eigsolve_lanczos, linsolve_gmres
from .contract import transfer_left_bond, transfer_right_bond, dAC_sparse, dC_dense
from .mpo import MPOHamiltonian

TOLGAUGE = 1e-14


def dyn_tol(eps, it, tol_min, tol_max, factor):
    """updatetol (dynamictols.jl:50-53)"""
    return float(np.clip(eps * factor / np.sqrt(it), tol_min, tol_max))


class InfiniteMPS:
    def __init__(self, AL, AR, CR, AC, V, pq, modulus=0):
        self.AL, self.AR, self.CR, self.AC = AL, AR, CR, AC    # CR[i]: bond right of site i; CR[-1] == CR[N-1]
        self.V, self.pq, self.modulus = V, np.asarray(pq), modulus   # V[i]: bond space left of site i
        self.N = len(AL)

    # --- per-sector factorisations
    def _rowq(self, Vl):
        return (Vl.index_charges()[:, None] + self.pq[None, :]).reshape(-1)

    @staticmethod
    def leftorth(A, Vl, Vr, pq, mod=0):
        Dl, d, Dr = A.shape
        rowq = (Vl.index_charges()[:, None] + np.asarray(pq)[None, :]).reshape(-1)
        Q, R, nV = qrpos(A.reshape(Dl * d, Dr), rowq, Vr.index_charges(), mod)
        return Q.reshape(Dl, d, nV.dim), R

    @staticmethod
    def rightorth(A, Vl, Vr, pq, mod=0):
        Dl, d, Dr = A.shape
        colq = (Vr.index_charges()[None, :] - np.asarray(pq)[:, None]).reshape(-1)
        Lm, Q, nV = lqpos(A.reshape(Dl, d * Dr), Vl.index_charges(), colq, mod)
        return Lm, Q.reshape(nV.dim, d, Dr)

    @staticmethod
    def from_tensors(A, V, pq, C0=None, tol=TOLGAUGE, maxiter=100, mod=0):
        """InfiniteMPS(A::Vector) (infinitemps.jl:139-171): makefullrank!, gaugefix! order :LR, then AC = AL * CR"""
        N = len(A)
        A, V = makefullrank(A, V, pq, mod)
        if C0 is not None and C0.shape[0] != A[0].shape[0]:
            C0 = None
        D0 = A[0].shape[0]
        C0 = np.eye(D0, dtype=np.complex128) if C0 is None else C0
        AL, CR = uniform_leftorth(A, C0, V, pq, tol, maxiter, mod)
        AR, CR = uniform_rightorth(AL, CR[N - 1], V, pq, tol, maxiter, mod)
        AC = [np.einsum('apb,bc->apc', AL[i], CR[i]) for i in range(N)]
        return InfiniteMPS(AL, AR, CR, AC, V, pq, mod)

    @staticmethod
    def from_AL(AL, C0, V, pq, tol=TOLGAUGE, maxiter=100, mod=0):
        """InfiniteMPS(AL, C0) (infinitemps.jl:173-187): gaugefix! order :R"""
        N = len(AL)
        AL = [a.copy() for a in AL]
        AR, CR = uniform_rightorth(AL, C0, V, pq, tol, maxiter, mod)
        AC = [np.einsum('apb,bc->apc', AL[i], CR[i]) for i in range(N)]
        return InfiniteMPS(AL, AR, CR, AC, V, pq, mod)

    @staticmethod
    def random(rng, pspace: Space, Vs, dtype=np.complex128):
        from .blocks import mask3
        N = len(Vs)
        pq = pspace.index_charges()
        A = []
        for i in range(N):
            Vl, Vr = Vs[i], Vs[(i + 1) % N]
            msk = mask3(Vl, pq, Vr)
            a = rng.standard_normal(msk.shape) + 1j * rng.standard_normal(msk.shape)
            A.append((a * msk).astype(dtype))
        return InfiniteMPS.from_tensors(A, list(Vs), pq, mod=pspace.modulus)


def isfullrank(A, Vl, Vr, pq, mod=0, side="both"):
    """isfullrank(A; side) (src/states/abstractmps.jl:74-87): a statement about sector dimensions only --
    :right  Vl (x) P >= Vr  and  :left  Vl <= P (x) Vr  sector by sector."""
    pq = np.asarray(pq)
    rowq = (Vl.index_charges()[:, None] + pq[None, :]).reshape(-1)
    colq = (Vr.index_charges()[None, :] - pq[:, None]).reshape(-1)
    if mod:
        rowq, colq = rowq % mod, colq % mod
    right = all(n <= int(np.sum(rowq == c)) for c, n in zip(Vr.charges, Vr.dims))
    left = all(n <= int(np.sum(colq == c)) for c, n in zip(Vl.charges, Vl.dims))
    return {"both": right and left, "right": right, "left": left}[side]


def makefullrank(A, V, pq, mod=0):
    """makefullrank!(A) (src/states/abstractmps.jl:94-108, called by the InfiniteMPS constructor, infinitemps.jl:150): bond
    sectors larger than what the neighbouring tensor can fill are cut by QR / LQ sweeps.  V[i]: bond left of site i.  Returns
    the input objects unchanged when every tensor is full rank (the usual case)."""
    N = len(A)
    if all(isfullrank(A[i], V[i], V[(i + 1) % N], pq, mod) for i in range(N)):
        return A, V
    A, V = [np.array(a, dtype=np.complex128) for a in A], list(V)
    for _ in range(100 * N):
        bad = [i for i in range(N) if not isfullrank(A[i], V[i], V[(i + 1) % N], pq, mod)]
        if not bad:
            return A, V
        i = bad[0]
        Vl, Vr = V[i], V[(i + 1) % N]
        if not isfullrank(A[i], Vl, Vr, pq, mod, "left"):
            Dl, d, Dr = A[i].shape
            colq = (Vr.index_charges()[None, :] - np.asarray(pq)[:, None]).reshape(-1)
            Lm, Q, nV = lqpos(A[i].reshape(Dl, d * Dr), Vl.index_charges(), colq, mod)
            A[i] = Q.reshape(nV.dim, d, Dr)
            A[(i - 1) % N] = np.einsum('apb,bc->apc', A[(i - 1) % N], Lm)
            V[i] = nV
        else:
            Dl, d, Dr = A[i].shape
            rowq = (Vl.index_charges()[:, None] + np.asarray(pq)[None, :]).reshape(-1)
            Q, R, nV = qrpos(A[i].reshape(Dl * d, Dr), rowq, Vr.index_charges(), mod)
            A[i] = Q.reshape(Dl, d, nV.dim)
            A[(i + 1) % N] = np.einsum('ab,bpc->apc', R, A[(i + 1) % N])
            V[(i + 1) % N] = nV
    raise RuntimeError("makefullrank did not terminate")


def _transfer_left_cell(v, A, B):
    for a, b in zip(A, B):
        v = transfer_left_bond(v, a, b)
    return v


def _transfer_right_cell(v, A, B):
    for a, b in zip(reversed(A), reversed(B)):
        v = transfer_right_bond(v, a, b)
    return v


def uniform_leftorth(A, C0, V, pq, tol=TOLGAUGE, maxiter=100, mod=0, eig_miniter=10):
    """ortho.jl:143-201"""
    N = len(A)
    CR = [None] * N
    CR[N - 1] = C0 / np.linalg.norm(C0)
    AL = [None] * N
    eps = np.inf
    for it in range(maxiter + 2):
        # gauge_eigsolve_step!
        if it >= eig_miniter:
            etol = dyn_tol(eps ** 2, 1, TOLGAUGE / 10, np.inf, 1)
            _, vec, _ = eigsolve_arnoldi(lambda v: _transfer_left_cell(v, A, AL), CR[N - 1], "LM", etol, 30, 100)
            Vb = V[0]
            _, R, _ = qrpos(vec, Vb.index_charges(), Vb.index_charges(), mod)
            CR[N - 1] = R
        Cprev = CR[N - 1]
        # gauge_orth_step!
        for i in range(N):
            CA = np.einsum('ab,bpc->apc', CR[(i - 1) % N], A[i])
            AL[i], CR[i] = InfiniteMPS.leftorth(CA, V[i], V[(i + 1) % N], pq, mod)
        CR[N - 1] = CR[N - 1] / np.linalg.norm(CR[N - 1])
        eps = np.linalg.norm(Cprev - CR[N - 1])
        if eps < tol or it + 1 > maxiter:
            break
    return AL, CR


def uniform_rightorth(A, C0, V, pq, tol=TOLGAUGE, maxiter=100, mod=0, eig_miniter=10):
    """ortho.jl:203-257"""
    N = len(A)
    CR = [None] * N
    CR[N - 1] = C0 / np.linalg.norm(C0)
    AR = [None] * N
    eps = np.inf
    for it in range(maxiter + 2):
        if it >= eig_miniter:
            etol = dyn_tol(eps ** 2, 1, TOLGAUGE / 10, np.inf, 1)
            _, vec, _ = eigsolve_arnoldi(lambda v: _transfer_right_cell(v, A, AR), CR[N - 1], "LM", etol, 30, 100)
            Vb = V[0]
            Lm, _, _ = lqpos(vec, Vb.index_charges(), Vb.index_charges(), mod)
            CR[N - 1] = Lm
        Cprev = CR[N - 1]
        for i in range(N - 1, -1, -1):
            AC = np.einsum('apb,bc->apc', A[i], CR[i])
            CR[(i - 1) % N], AR[i] = InfiniteMPS.rightorth(AC, V[i], V[(i + 1) % N], pq, mod)
        CR[N - 1] = CR[N - 1] / np.linalg.norm(CR[N - 1])
        eps = np.linalg.norm(Cprev - CR[N - 1])
        if eps < tol or it + 1 > maxiter:
            break
    return AR, CR


# ------------------------------------------------------------------------------------------ MPOHamInfEnv
def _isid(H: MPOHamiltonian, i):
    """mpohamiltonian.jl:52-57"""
    return all(np.isscalar(s.Os[i][i]) and abs(s.Os[i][i] - 1) < 1e-14 for s in H.slices)


def _tl_entry(v, sl, j, k, A):
    """fp[j] * TransferMatrix(A, H[j,k], A)"""
    if sl.isscal(j, k):
        return sl.Os[j][k] * transfer_left_bond(v, A, A)
    return np.einsum('ad,dec,fe,afh->hc', v, A, sl.op(j, k), np.conj(A), optimize=True)


def _tr_entry(v, sl, j, k, A):
    """TransferMatrix(A, H[j,k], A) * fp[k]"""
    if sl.isscal(j, k):
        return sl.Os[j][k] * transfer_right_bond(v, A, A)
    return np.einsum('xba,db,zde,ae->xz', A, sl.op(j, k), np.conj(A), v, optimize=True)


class MPOHamInfEnv:
    def __init__(self, psi: InfiniteMPS, H: MPOHamiltonian, tol=1e-12, maxiter=100):
        self.opp, self.tol, self.maxiter = H, tol, maxiter
        N, od = psi.N, H.odim
        rng = np.random.default_rng(0)
        D = [psi.AL[i].shape[0] for i in range(N)]
        self.lw = [[np.zeros((D[i], D[i]), dtype=np.complex128) for i in range(N)] for _ in range(od)]
        self.rw = [[np.zeros((psi.AR[i].shape[2], psi.AR[i].shape[2]), dtype=np.complex128) for i in range(N)] for _ in range(od)]
        self.dependency = None
        self.recalculate(psi)

    def recalculate(self, psi, tol=None):
        tol = self.tol if tol is None else tol
        calclw(self.lw, psi, self.opp, tol, self.maxiter)
        calcrw(self.rw, psi, self.opp, tol, self.maxiter)
        self.dependency = psi
        return self

    def leftenv(self, pos, psi):
        if self.dependency is not psi:
            self.recalculate(psi)
        return np.stack([self.lw[j][pos % psi.N] for j in range(self.opp.odim)], axis=1)     # [bra, w, ket]

    def rightenv(self, pos, psi):
        if self.dependency is not psi:
            self.recalculate(psi)
        return np.stack([self.rw[j][pos % psi.N] for j in range(self.opp.odim)], axis=1)     # [ket, w, bra]


def _left_cyclethrough(index, fp, H, st):
    N = st.N
    for i in range(N):
        nxt = (i + 1) % N
        fp[index][nxt] = np.zeros_like(fp[index][nxt])
        for j in range(index, -1, -1):
            if not H[i].contains(j, index):
                continue
            fp[index][nxt] = fp[index][nxt] + _tl_entry(fp[j][i], H[i], j, index, st.AL[i])


def _right_cyclethrough(index, fp, H, st):
    N = st.N
    for i in range(N - 1, -1, -1):
        prv = (i - 1) % N
        fp[index][prv] = np.zeros_like(fp[index][prv])
        for j in range(index, len(fp)):
            if not H[i].contains(index, j):
                continue
            fp[index][prv] = fp[index][prv] + _tr_entry(fp[j][i], H[i], index, j, st.AR[i])


def calclw(fp, st: InfiniteMPS, H, tol, maxiter):
    """mpohaminfenv.jl:76-125.  fp[j][i]: left env of level j at site i ([bra; ket])"""
    N, od = st.N, H.odim
    D0 = st.AL[0].shape[0]
    fp[0][0] = np.eye(D0, dtype=np.complex128)                      # l_LL (x) 1
    if N > 1:
        _left_cyclethrough(0, fp, H, st)
    rho = [st.CR[i] @ st.CR[i].conj().T for i in range(N)]          # r_LL(st, i)
    for i in range(1, od):
        prev = fp[i][0].copy()
        fp[i][0] = np.zeros_like(fp[i][0])
        _left_cyclethrough(i, fp, H, st)
        if _isid(H, i):
            r_end = rho[N - 1]

            def tm(v):                                              # flip(regularize(TransferMatrix(AL, AL), l_LL, r_LL))
                w = _transfer_left_cell(v, st.AL, st.AL)
                return w - np.einsum('ab,ba->', r_end, w) * np.eye(D0)
            fp[i][0], info = linsolve_gmres(tm, fp[i][0], prev, 1.0, -1.0, tol, 30, maxiter)
            if N > 1:
                _left_cyclethrough(i, fp, H, st)
            for s in range(N):                                      # subtract the fixed-point component on every site
                fp[i][s] = fp[i][s] - np.einsum('ab,ba->', fp[i][s], rho[(s - 1) % N]) * np.eye(fp[i][s].shape[0])
        else:
            if all(sl.contains(i, i) for sl in H.slices):
                def tm(v):
                    for s in range(N):
                        v = _tl_entry(v, H[s], i, i, st.AL[s])
                    return v
                fp[i][0], info = linsolve_gmres(tm, fp[i][0], prev, 1.0, -1.0, tol, 30, maxiter)
            if N > 1:
                _left_cyclethrough(i, fp, H, st)
    return fp


def calcrw(fp, st: InfiniteMPS, H, tol, maxiter):
    """mpohaminfenv.jl:127-175.  fp[j][i]: right env of level j at site i ([ket; bra])"""
    N, od = st.N, H.odim
    De = st.AR[N - 1].shape[2]
    fp[od - 1][N - 1] = np.eye(De, dtype=np.complex128)
    if N > 1:
        _right_cyclethrough(od - 1, fp, H, st)
    lrr = [st.CR[(i - 1) % N].conj().T @ st.CR[(i - 1) % N] for i in range(N)]     # l_RR(st, i)
    for i in range(od - 2, -1, -1):
        prev = fp[i][N - 1].copy()
        fp[i][N - 1] = np.zeros_like(fp[i][N - 1])
        _right_cyclethrough(i, fp, H, st)
        if _isid(H, i):
            l0 = lrr[0]

            def tm(v):                                              # regularize(TransferMatrix(AR, AR), l_RR, r_RR)
                w = _transfer_right_cell(v, st.AR, st.AR)
                return w - np.einsum('ab,ba->', l0, w) * np.eye(De)
            fp[i][N - 1], info = linsolve_gmres(tm, fp[i][N - 1], prev, 1.0, -1.0, tol, 30, maxiter)
            if N > 1:
                _right_cyclethrough(i, fp, H, st)
            for s in range(N):
                fp[i][s] = fp[i][s] - np.einsum('ab,ba->', fp[i][s], lrr[(s + 1) % N]) * np.eye(fp[i][s].shape[0])
        else:
            if all(sl.contains(i, i) for sl in H.slices):
                def tm(v):
                    for s in range(N - 1, -1, -1):
                        v = _tr_entry(v, H[s], i, i, st.AR[s])
                    return v
                fp[i][N - 1], info = linsolve_gmres(tm, fp[i][N - 1], prev, 1.0, -1.0, tol, 30, maxiter)
            if N > 1:
                _right_cyclethrough(i, fp, H, st)
    return fp


def expectation_value(st: InfiniteMPS, H, envs: MPOHamInfEnv):
    """expval.jl:111-124: energy density per site"""
    N, od = st.N, H.odim
    ens = np.zeros(N, dtype=np.complex128)
    for i in range(N):
        GL = envs.leftenv(i, st)
        rho = st.CR[i] @ st.CR[i].conj().T
        for j in range(od - 1, -1, -1):
            if not H[i].contains(j, od - 1):
                continue
            apl = _tl_entry(GL[:, j, :], H[i], j, od - 1, st.AL[i])
            ens[i] += np.einsum('ab,ba->', apl, rho)
    return ens


def calc_galerkin(st: InfiniteMPS, loc, envs):
    """toolbox.jl:17-22"""
    y = dAC_sparse(st.AC[loc], envs.opp[loc], envs.leftenv(loc, st), envs.rightenv(loc, st))
    y = y / np.linalg.norm(y)
    AL = st.AL[loc]
    return float(np.linalg.norm(y - np.einsum('apb,cqb,cqd->apd', AL, np.conj(AL), y, optimize=True)))


def regauge(AC, C, Vl, Vr, pq, mod=0):
    """ortho.jl:127-131: AL = Q_AC * Q_C'"""
    Qac, _ = InfiniteMPS.leftorth(AC, Vl, Vr, pq, mod)
    Qc, _, _ = qrpos(C, Vr.index_charges(), Vr.index_charges(), mod)
    return np.einsum('apb,cb->apc', Qac, np.conj(Qc))


def vumps(st: InfiniteMPS, H, tol=1e-12, maxiter=100, envs=None, log=None, project=None):
    """find_groundstate(psi::InfiniteMPS, H, VUMPS()) (vumps.jl:29-92) with the default dynamic tolerances.
    project = (pac, pc): optional projectors pac(loc, x), pc(loc, x) onto a symmetric subspace that the effective Hamiltonians
    leave invariant; applied after every application, i.e. the eigenproblems are solved inside that subspace -- what the
    reference does when its tensors carry a larger (non-abelian) symmetry than the U(1) blocks of this restatement
    (tests/golden/make_su2_pins.py)."""
    envs = envs or MPOHamInfEnv(st, H)
    N = st.N
    eps = max(calc_galerkin(st, i, envs) for i in range(N))
    pac, pc = project if project is not None else ((lambda loc, x: x), (lambda loc, x: x))
    for it in range(1, maxiter + 1):
        etol = dyn_tol(eps, it, 1e-12, 1e-5, 1e-5)
        newAL = []
        for loc in range(N):
            GL, GR = envs.leftenv(loc, st), envs.rightenv(loc, st)
            _, ac, _ = eigsolve_lanczos(lambda x: pac(loc, dAC_sparse(x, H[loc], GL, GR)), pac(loc, st.AC[loc]), "SR", etol, 30, 100)
            GLc, GRc = envs.leftenv(loc + 1, st), envs.rightenv(loc, st)
            _, c, _ = eigsolve_lanczos(lambda x: pc(loc, dC_dense(x, GLc, GRc)), pc(loc, st.CR[loc]), "SR", etol, 30, 100)
            newAL.append(regauge(ac, c, st.V[loc], st.V[(loc + 1) % N], st.pq, st.modulus))
        gtol = dyn_tol(eps, it, TOLGAUGE, 1e-5, 1e-8)
        st = InfiniteMPS.from_AL(newAL, st.CR[N - 1], st.V, st.pq, gtol, 100, st.modulus)
        envtol = dyn_tol(eps, it, 1e-12, 1e-5, 1e-5)
        envs.recalculate(st, envtol)
        eps = max(calc_galerkin(st, i, envs) for i in range(N))
        if log is not None:
            log.append(dict(iter=it, eps=eps, energy=float(np.sum(expectation_value(st, H, envs)).real)))
        if eps <= tol:
            break
    return st, envs, eps


def tdvp_step(st: InfiniteMPS, H, dt, envs=None, tol=1e-12, krylovdim=30, tolgauge=TOLGAUGE, gaugemaxiter=100):
    """timestep(psi::InfiniteMPS, H, t, dt, TDVP(), envs) with leftorthflag = true (src/algorithms/timestep/tdvp.jl:21-59):
    every AC and CR of the unit cell is evolved with its own effective Hamiltonian (integrators.jl:20-25), the pairs are
    regauged to left-orthonormal tensors (ortho.jl:127-131) and the environments are recalculated."""
    from .krylov import exponentiate_lanczos
    envs = envs or MPOHamInfEnv(st, H)
    N = st.N
    newAL = []
    for loc in range(N):
        GL, GR = envs.leftenv(loc, st), envs.rightenv(loc, st)
        ac, _ = exponentiate_lanczos(lambda x: dAC_sparse(x, H[loc], GL, GR), -1j * dt, st.AC[loc], tol, krylovdim)
        GLc, GRc = envs.leftenv(loc + 1, st), envs.rightenv(loc, st)
        c, _ = exponentiate_lanczos(lambda x: dC_dense(x, GLc, GRc), -1j * dt, st.CR[loc], tol, krylovdim)
        newAL.append(regauge(ac, c, st.V[loc], st.V[(loc + 1) % N], st.pq, st.modulus))
    st = InfiniteMPS.from_AL(newAL, st.CR[N - 1], st.V, st.pq, tolgauge, gaugemaxiter, st.modulus)
    envs.recalculate(st)
    return st, envs


# ------------------------------------------------------------------------------------------ IDMRG1 (next-row preparation)
def idmrg1(st: InfiniteMPS, H, tol=1e-12, maxiter=100, envs=None, log=None):
    """find_groundstate(psi::InfiniteMPS, H, IDMRG1()) (src/algorithms/groundstate/idmrg.jl:21-75) with the manually
    updated IDMRGEnv (src/environments/idmrgenv.jl:6-72).  Oracle only in round 1."""
    from .contract import transfer_left_sparse, transfer_right_sparse
    envs0 = envs or MPOHamInfEnv(st, H)
    N, od = st.N, H.odim
    lw = [envs0.leftenv(i, st).copy() for i in range(N)]          # IDMRGEnv: deep copies, updated by hand
    rw = [envs0.rightenv(i, st).copy() for i in range(N)]
    AL, AR, AC, CR = list(st.AL), list(st.AR), list(st.AC), list(st.CR)
    eps = np.inf
    for it in range(1, maxiter + 1):
        etol = max(tol / 10, 1e-14)
        Ccur = CR[N - 1]
        for pos in range(N):                                       # left -> right
            GL, GR = lw[pos], rw[pos]
            _, AC[pos], _ = eigsolve_lanczos(lambda x: dAC_sparse(x, H[pos], GL, GR), AC[pos], "SR", etol, 30, 100)
            AL[pos], CR[pos] = InfiniteMPS.leftorth(AC[pos], st.V[pos], st.V[(pos + 1) % N], st.pq, st.modulus)
            lw[(pos + 1) % N] = transfer_left_sparse(lw[pos], H[pos], AL[pos], AL[pos])           # update_leftenv!
        for pos in range(N - 1, -1, -1):                           # right -> left
            GL, GR = lw[pos], rw[pos]
            _, AC[pos], _ = eigsolve_lanczos(lambda x: dAC_sparse(x, H[pos], GL, GR), AC[pos], "SR", etol, 30, 100)
            CR[(pos - 1) % N], AR[pos] = InfiniteMPS.rightorth(AC[pos], st.V[pos], st.V[(pos + 1) % N], st.pq, st.modulus)
            rw[(pos - 1) % N] = transfer_right_sparse(rw[pos], H[pos], AR[pos], AR[pos])          # update_rightenv!
        eps = float(np.linalg.norm(Ccur - CR[N - 1]))
        if log is not None:
            log.append(dict(iter=it, eps=eps))
        if eps < tol:
            break
    nst = InfiniteMPS.from_tensors(AR, st.V, st.pq, mod=st.modulus)
    nenvs = MPOHamInfEnv(nst, H)
    return nst, nenvs, eps


def idmrg2(st: InfiniteMPS, H, trunc=("truncerr", 1e-6), tol=1e-12, maxiter=100, envs=None, log=None):
    """find_groundstate(psi::InfiniteMPS, H, IDMRG2()) (src/algorithms/groundstate/idmrg.jl:92-204): two-site updates
    through the unit cell and across its edge (with inv(CR) re-gauging), truncated SVD after every update, manually
    updated IDMRGEnv.  Unit cell >= 2.  Bond spaces change with the truncation (tracked in V)."""
    from .contract import transfer_left_sparse, transfer_right_sparse, dAC2_sparse
    from .linalg import tsvd
    from .spaces import Space
    N = st.N
    assert N >= 2, "unit cell should be >= 2"
    envs0 = envs or MPOHamInfEnv(st, H)
    lw = [envs0.leftenv(i, st).copy() for i in range(N)]
    rw = [envs0.rightenv(i, st).copy() for i in range(N)]
    AL, AR, AC, CR, V = list(st.AL), list(st.AR), list(st.AC), list(st.CR), list(st.V)
    pq, mod = np.asarray(st.pq), st.modulus
    state = {"E": 0.0}

    def update(GL, GR, h1, h2, ac2, Vl, Vr, etol):
        val, new, _ = eigsolve_lanczos(lambda x: dAC2_sparse(x, h1, h2, GL, GR), ac2, "SR", etol, 30, 100)
        state["E"] = float(np.real(val))
        Dl, d1, Dr, d2 = new.shape
        rowq = (Vl.index_charges()[:, None] + pq[None, :]).reshape(-1)
        colq = (Vr.index_charges()[:, None] - pq[None, :]).reshape(-1)
        U, S, Vh, newV, _, _ = tsvd(new.reshape(Dl * d1, Dr * d2), rowq, colq, trunc, mod)
        S = S / np.linalg.norm(S)                                                 # normalize!(c)
        return U.reshape(Dl, d1, newV.dim), np.diag(S).astype(np.complex128), Vh.reshape(newV.dim, Dr, d2).transpose(0, 2, 1), newV

    def tl(i):      # update_leftenv!(envs, psi, H, i + 1) (idmrgenv.jl:68-72)
        lw[(i + 1) % N] = transfer_left_sparse(lw[i], H[i], AL[i], AL[i])

    def tr(i):      # update_rightenv!(envs, psi, H, i - 1) (idmrgenv.jl:63-66)
        rw[(i - 1) % N] = transfer_right_sparse(rw[i], H[i], AR[i], AR[i])
    eps = np.inf
    for it in range(1, maxiter + 1):
        etol = max(tol / 10, 1e-14)
        for pos in range(N - 1):                                                  # idmrg.jl:111-126
            ac2 = np.einsum('apk,kqr->aprq', AC[pos], AR[pos + 1])
            al, c, ar, nV = update(lw[pos], rw[pos + 1], H[pos], H[pos + 1], ac2, V[pos], V[(pos + 2) % N], etol)
            AL[pos], CR[pos], AR[pos + 1], V[pos + 1] = al, c, ar, nV
            AC[pos + 1] = np.einsum('ab,bpc->apc', c, ar)
            tl(pos); tr(pos + 1)
        # edge (idmrg.jl:129-148): sites N-1 and 0
        ac2 = np.einsum('apk,kl,lqm,mr->aprq', AC[N - 1], np.linalg.inv(CR[N - 1]), AL[0], CR[0])
        al, c, ar, nV = update(lw[N - 1], rw[0], H[N - 1], H[0], ac2, V[N - 1], V[1 % N], etol)
        AC[N - 1] = np.einsum('apb,bc->apc', al, c); AL[N - 1] = al; CR[N - 1] = c; AR[0] = ar; V[0] = nV
        AC[0] = np.einsum('ab,bpc->apc', c, ar)
        AL[0] = np.einsum('apb,bc->apc', AC[0], np.linalg.inv(CR[0]))
        Ccur, Vcur = c, nV
        tl(N - 1); tr(0)
        for pos in range(N - 2, -1, -1):                                          # idmrg.jl:151-167
            ac2 = np.einsum('apk,kqr->aprq', AL[pos], AC[pos + 1])
            al, c, ar, nV = update(lw[pos], rw[pos + 1], H[pos], H[pos + 1], ac2, V[pos], V[(pos + 2) % N], etol)
            AL[pos], CR[pos], AR[pos + 1], V[pos + 1] = al, c, ar, nV
            AC[pos] = np.einsum('apb,bc->apc', al, c)
            AC[pos + 1] = np.einsum('ab,bpc->apc', c, ar)
            tl(pos); tr(pos + 1)
        # edge (idmrg.jl:170-186)
        ac2 = np.einsum('ab,bpc,cd,dqe->apeq', CR[N - 2], AR[N - 1], np.linalg.inv(CR[N - 1]), AC[0])
        al, c, ar, nV = update(lw[N - 1], rw[0], H[N - 1], H[0], ac2, V[N - 1], V[1 % N], etol)
        AR[N - 1] = np.einsum('ab,bpc->apc', np.linalg.inv(CR[N - 2]), np.einsum('apb,bc->apc', al, c))
        AL[N - 1] = al; CR[N - 1] = c; AR[0] = ar; V[0] = nV
        AC[0] = np.einsum('ab,bpc->apc', c, ar)
        tl(N - 1); tr(0)
        # error (idmrg.jl:189-192): compare on the smallest common space
        qa, qb = Vcur.index_charges(), nV.index_charges()
        ia, ib = [], []
        for q in sorted(set(qa.tolist()) & set(qb.tolist())):
            xa, xb = np.nonzero(qa == q)[0], np.nonzero(qb == q)[0]
            k = min(len(xa), len(xb))
            ia += list(xa[:k]); ib += list(xb[:k])
        eps = float(np.linalg.norm(c[np.ix_(ib, ib)] - Ccur[np.ix_(ia, ia)]))
        if log is not None:
            log.append(dict(iter=it, eps=eps, energy=state["E"], D=nV.dim))
        if eps < tol:
            break
    nst = InfiniteMPS.from_tensors(AR, V, st.pq, mod=mod)                          # idmrg.jl:201
    nenvs = MPOHamInfEnv(nst, H)
    return nst, nenvs, eps
