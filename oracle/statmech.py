"""Statistical-mechanics VUMPS for a single-row unit cell (leading_boundary of an InfiniteMPS with a DenseMPO):
src/algorithms/statmech/vumps.jl:15-84, src/environments/permpoinfenv.jl:23-63,138-189 (mixed_fixpoints),
src/algorithms/expval.jl:165-173.  Dense restatement; the MPO is given as a list of flattened slices (oracle/mpo.py
dense_mpo_slice).  Multi-row unit cells (MPSMultiline: RecursiveVec + circshift, derivatives.jl:107-110,195-197;
mixed_fixpoints row by row, permpoinfenv.jl:138-189) are restated in the *_multiline functions at the end."""
from __future__ import annotations
import numpy as np
from .infinite import InfiniteMPS, regauge, dyn_tol, TOLGAUGE
from .krylov import eigsolve_arnoldi
from .contract import transfer_left_sparse, transfer_right_sparse, dAC_sparse, dC_dense


class PerMPOInfEnv:
    def __init__(self, psi: InfiniteMPS, slices, tol=1e-12):
        self.opp, self.tol = slices, tol
        self.lw = self.rw = None
        self.dependency = None
        self.recalculate(psi)

    def recalculate(self, psi, tol=None):
        tol = self.tol if tol is None else tol
        N = psi.N
        chi = self.opp[0].odim
        D0 = psi.AL[0].shape[0]
        rng = np.random.default_rng(0)
        L0 = self.lw[0] if self.lw is not None and self.lw[0].shape[0] == D0 else rng.standard_normal((D0, chi, D0)) + 0j
        R0 = self.rw[N - 1] if self.rw is not None and self.rw[N - 1].shape[0] == D0 else rng.standard_normal((D0, chi, D0)) + 0j

        def TL(v):
            for i in range(N):
                v = transfer_left_sparse(v, self.opp[i], psi.AL[i], psi.AL[i])
            return v

        def TR(v):
            for i in range(N - 1, -1, -1):
                v = transfer_right_sparse(v, self.opp[i], psi.AR[i], psi.AR[i])
            return v
        GL = [None] * N
        GR = [None] * N
        _, GL[0], _ = eigsolve_arnoldi(TL, L0, "LM", tol, 30, 100)
        for c in range(1, N):
            GL[c] = transfer_left_sparse(GL[c - 1], self.opp[c - 1], psi.AL[c - 1], psi.AL[c - 1])
        _, GR[N - 1], _ = eigsolve_arnoldi(TR, R0, "LM", tol, 30, 100)
        for c in range(N - 2, -1, -1):
            GR[c] = transfer_right_sparse(GR[c + 1], self.opp[c + 1], psi.AR[c + 1], psi.AR[c + 1])
        for c in range(N):                      # permpoinfenv.jl:180-185
            lam = np.vdot(psi.CR[c], dC_dense(psi.CR[c], GL[(c + 1) % N], GR[c]))
            # the reference divides both by sqrt(|lam|); the Arnoldi eigenvectors carry an arbitrary phase, which we
            # remove here as well (lam -> real positive) so that read-outs are gauge independent
            GL[(c + 1) % N] = GL[(c + 1) % N] * (np.conj(lam) / abs(lam)) / np.sqrt(abs(lam))
            GR[c] = GR[c] / np.sqrt(abs(lam))
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


def expectation_value(psi, envs):
    """expval.jl:165-173"""
    out = []
    for i in range(psi.N):
        y = dAC_sparse(psi.AC[i], envs.opp[i], envs.leftenv(i, psi), envs.rightenv(i, psi))
        out.append(np.vdot(psi.AC[i], y))
    return np.array(out)


def calc_galerkin(psi, loc, envs):
    y = dAC_sparse(psi.AC[loc], envs.opp[loc], envs.leftenv(loc, psi), envs.rightenv(loc, psi))
    y = y / np.linalg.norm(y)
    AL = psi.AL[loc]
    return float(np.linalg.norm(y - np.einsum('apb,cqb,cqd->apd', AL, np.conj(AL), y, optimize=True)))


def leading_boundary(psi: InfiniteMPS, slices, tol=1e-10, maxiter=100, log=None):
    """statmech/vumps.jl:20-84 for one row"""
    envs = PerMPOInfEnv(psi, slices)
    N = psi.N
    eps = max(calc_galerkin(psi, i, envs) for i in range(N))
    for it in range(1, maxiter + 1):
        etol = dyn_tol(eps, it, 1e-12, 1e-5, 1e-5)
        newAL = []
        for col in range(N):
            GL, GR = envs.leftenv(col, psi), envs.rightenv(col, psi)
            _, ac, _ = eigsolve_arnoldi(lambda x: dAC_sparse(x, slices[col], GL, GR), psi.AC[col], "LM", etol, 30, 100)
            GLc = envs.leftenv(col + 1, psi)
            _, c, _ = eigsolve_arnoldi(lambda x: dC_dense(x, GLc, GR), psi.CR[col], "LM", etol, 30, 100)
            newAL.append(regauge(ac, c, psi.V[col], psi.V[(col + 1) % N], psi.pq, psi.modulus))
        gtol = dyn_tol(eps, it, TOLGAUGE, 1e-5, 1e-8)
        psi = InfiniteMPS.from_AL(newAL, psi.CR[N - 1], psi.V, psi.pq, gtol, 100, psi.modulus)
        envs.recalculate(psi, dyn_tol(eps, it, 1e-12, 1e-5, 1e-5))
        eps = max(calc_galerkin(psi, i, envs) for i in range(N))
        if log is not None:
            log.append(dict(iter=it, eps=eps, value=expectation_value(psi, envs)))
        if eps <= tol:
            break
    return psi, envs, eps


# ------------------------------------------------------------------------------------------ multi-row unit cells
class MultilineEnv:
    """PerMPOInfEnv for an MPSMultiline (permpoinfenv.jl:4-63,138-189): lw[r][c] / rw[r][c] between row r (top, ket)
    and row r + 1 (bottom, bra); slices[r][c] is the MPO tensor of row r, column c."""

    def __init__(self, rows, slices, tol=1e-12):
        self.opp, self.tol = slices, tol
        self.lw = self.rw = None
        self.dependency = None
        self.recalculate(rows)

    def recalculate(self, rows, tol=None):
        tol = self.tol if tol is None else tol
        R, N = len(rows), rows[0].N
        GLs = [[None] * N for _ in range(R)]
        GRs = [[None] * N for _ in range(R)]
        rng = np.random.default_rng(0)
        for r in range(R):
            top, bot = rows[r], rows[(r + 1) % R]
            Os = self.opp[r]
            chi = Os[0].odim
            D0 = top.AL[0].shape[0]
            same = self.lw is not None and self.lw[r][0].shape[0] == D0
            L0 = self.lw[r][0] if same else rng.standard_normal((D0, chi, D0)) + 0j
            R0 = self.rw[r][N - 1] if same else rng.standard_normal((D0, chi, D0)) + 0j

            def TL(v):
                for i in range(N):
                    v = transfer_left_sparse(v, Os[i], top.AL[i], bot.AL[i])
                return v

            def TR(v):
                for i in range(N - 1, -1, -1):
                    v = transfer_right_sparse(v, Os[i], top.AR[i], bot.AR[i])
                return v
            _, GLs[r][0], _ = eigsolve_arnoldi(TL, L0, "LM", tol, 30, 100)
            for c in range(1, N):
                GLs[r][c] = transfer_left_sparse(GLs[r][c - 1], Os[c - 1], top.AL[c - 1], bot.AL[c - 1])
            _, GRs[r][N - 1], _ = eigsolve_arnoldi(TR, R0, "LM", tol, 30, 100)
            for c in range(N - 2, -1, -1):
                GRs[r][c] = transfer_right_sparse(GRs[r][c + 1], Os[c + 1], top.AR[c + 1], bot.AR[c + 1])
            for c in range(N):                      # permpoinfenv.jl:180-185 (+ removal of the arbitrary Arnoldi phase)
                lam = np.vdot(bot.CR[c], dC_dense(top.CR[c], GLs[r][(c + 1) % N], GRs[r][c]))
                GLs[r][(c + 1) % N] = GLs[r][(c + 1) % N] * (np.conj(lam) / abs(lam)) / np.sqrt(abs(lam))
                GRs[r][c] = GRs[r][c] / np.sqrt(abs(lam))
        self.lw, self.rw, self.dependency = GLs, GRs, rows
        return self


def expectation_value_multiline(rows, envs):
    """expval.jl:165-173: <AC[r+1, c]| O[r, c] |AC[r, c]> with the environments, per (row, column)"""
    R, N = len(rows), rows[0].N
    out = np.zeros((R, N), dtype=np.complex128)
    for r in range(R):
        for c in range(N):
            y = dAC_sparse(rows[r].AC[c], envs.opp[r][c], envs.lw[r][c], envs.rw[r][c])
            out[r, c] = np.vdot(rows[(r + 1) % R].AC[c], y)
    return out


def calc_galerkin_multiline(rows, envs):
    """toolbox.jl:26-38"""
    R, N = len(rows), rows[0].N
    eps = 0.0
    for r in range(R):
        for c in range(N):
            y = dAC_sparse(rows[r].AC[c], envs.opp[r][c], envs.lw[r][c], envs.rw[r][c])
            y = y / np.linalg.norm(y)
            AL = rows[(r + 1) % R].AL[c]
            eps = max(eps, float(np.linalg.norm(y - np.einsum('apb,cqb,cqd->apd', AL, np.conj(AL), y, optimize=True))))
    return eps


def leading_boundary_multiline(rows, slices, tol=1e-10, maxiter=100, log=None):
    """leading_boundary(psi::MPSMultiline, H::MPOMultiline, VUMPS()) (statmech/vumps.jl:20-84): every column update is ONE
    eigenproblem on the RecursiveVec of all rows, (dAC x)[r + 1] = dAC(x[r], O[r, col], GL[r, col], GR[r, col])."""
    envs = MultilineEnv(rows, slices)
    R, N = len(rows), rows[0].N
    eps = calc_galerkin_multiline(rows, envs)
    for it in range(1, maxiter + 1):
        etol = dyn_tol(eps, it, 1e-12, 1e-5, 1e-5)
        newAL = [[None] * N for _ in range(R)]
        for col in range(N):
            def HAC(x):
                return np.stack([dAC_sparse(x[(r - 1) % R], slices[(r - 1) % R][col], envs.lw[(r - 1) % R][col], envs.rw[(r - 1) % R][col]) for r in range(R)])

            def HC(x):
                return np.stack([dC_dense(x[(r - 1) % R], envs.lw[(r - 1) % R][(col + 1) % N], envs.rw[(r - 1) % R][col]) for r in range(R)])
            _, ac, _ = eigsolve_arnoldi(HAC, np.stack([rows[r].AC[col] for r in range(R)]), "LM", etol, 30, 100)
            _, c, _ = eigsolve_arnoldi(HC, np.stack([rows[r].CR[col] for r in range(R)]), "LM", etol, 30, 100)
            for r in range(R):
                newAL[r][col] = regauge(ac[r], c[r], rows[r].V[col], rows[r].V[(col + 1) % N], rows[r].pq, rows[r].modulus)
        gtol = dyn_tol(eps, it, TOLGAUGE, 1e-5, 1e-8)
        rows = [InfiniteMPS.from_AL(newAL[r], rows[r].CR[N - 1], rows[r].V, rows[r].pq, gtol, 100, rows[r].modulus) for r in range(R)]
        envs.recalculate(rows, dyn_tol(eps, it, 1e-12, 1e-5, 1e-5))
        eps = calc_galerkin_multiline(rows, envs)
        if log is not None:
            log.append(dict(iter=it, eps=eps, value=expectation_value_multiline(rows, envs)))
        if eps <= tol:
            break
    return rows, envs, eps


def leading_boundary_multiline_vomps(rows, slices, tol=1e-10, maxiter=100, log=None):
    """leading_boundary(psi::MPSMultiline, O::MPOMultiline, VOMPS()) (statmech/vomps.jl:27-87): the power method -- one
    application of the multi-row ddAC / ddC (row r feeds row r + 1, derivatives.jl:107-110) instead of the eigensolve of VUMPS."""
    envs = MultilineEnv(rows, slices)
    R, N = len(rows), rows[0].N
    eps = calc_galerkin_multiline(rows, envs)
    for it in range(1, maxiter + 1):
        newAL = [[None] * N for _ in range(R)]
        for col in range(N):
            for r in range(R):
                t = (r - 1) % R
                ac = dAC_sparse(rows[t].AC[col], slices[t][col], envs.lw[t][col], envs.rw[t][col])
                c = dC_dense(rows[t].CR[col], envs.lw[t][(col + 1) % N], envs.rw[t][col])
                newAL[r][col] = regauge(ac, c, rows[r].V[col], rows[r].V[(col + 1) % N], rows[r].pq, rows[r].modulus)
        gtol = dyn_tol(eps, it, TOLGAUGE, 1e-5, 1e-8)
        rows = [InfiniteMPS.from_AL(newAL[r], rows[r].CR[N - 1], rows[r].V, rows[r].pq, gtol, 100, rows[r].modulus) for r in range(R)]
        envs.recalculate(rows, dyn_tol(eps, it, 1e-12, 1e-5, 1e-5))
        eps = calc_galerkin_multiline(rows, envs)
        if log is not None:
            log.append(dict(iter=it, eps=eps, value=expectation_value_multiline(rows, envs)))
        if eps <= tol:
            break
    return rows, envs, eps


# ------------------------------------------------------------------------------------------ VOMPS (approximate / power method)
class MixedPerMPOInfEnv:
    """environments(below, (opp, above)) for one row (permpoinfenv.jl:29-63, mixed_fixpoints 138-189): fixed points of
    TransferMatrix(above.AL, O, below.AL) / TransferMatrix(above.AR, O, below.AR), normalised with
    lambda = <C_below| dC(GL, GR) |C_above>; above = None means above = the state itself (leading_boundary)."""

    def __init__(self, below: InfiniteMPS, slices, above: InfiniteMPS = None, tol=1e-12):
        self.opp, self.above, self.tol = slices, above, tol
        self.lw = self.rw = None
        self.dependency = None
        self.recalculate(below)

    def recalculate(self, below, tol=None):
        tol = self.tol if tol is None else tol
        above = below if self.above is None else self.above
        N = below.N
        chi = self.opp[0].odim
        Db, Da = below.AL[0].shape[0], above.AL[0].shape[0]
        rng = np.random.default_rng(0)
        ok = self.lw is not None and self.lw[0].shape == (Db, chi, Da)
        L0 = self.lw[0] if ok else rng.standard_normal((Db, chi, Da)) + 0j
        R0 = self.rw[N - 1] if ok else rng.standard_normal((Da, chi, Db)) + 0j

        def TL(v):
            for i in range(N):
                v = transfer_left_sparse(v, self.opp[i], above.AL[i], below.AL[i])
            return v

        def TR(v):
            for i in range(N - 1, -1, -1):
                v = transfer_right_sparse(v, self.opp[i], above.AR[i], below.AR[i])
            return v
        GL, GR = [None] * N, [None] * N
        _, GL[0], _ = eigsolve_arnoldi(TL, L0, "LM", tol, 30, 100)
        for c in range(1, N):
            GL[c] = transfer_left_sparse(GL[c - 1], self.opp[c - 1], above.AL[c - 1], below.AL[c - 1])
        _, GR[N - 1], _ = eigsolve_arnoldi(TR, R0, "LM", tol, 30, 100)
        for c in range(N - 2, -1, -1):
            GR[c] = transfer_right_sparse(GR[c + 1], self.opp[c + 1], above.AR[c + 1], below.AR[c + 1])
        for c in range(N):                      # permpoinfenv.jl:180-185
            lam = np.vdot(below.CR[c], dC_dense(above.CR[c], GL[(c + 1) % N], GR[c]))
            GL[(c + 1) % N] = GL[(c + 1) % N] * (np.conj(lam) / abs(lam)) / np.sqrt(abs(lam))
            GR[c] = GR[c] / np.sqrt(abs(lam))
        self.lw, self.rw, self.dependency = GL, GR, below
        return self

    def leftenv(self, pos, psi):
        if self.dependency is not psi:
            self.recalculate(psi)
        return self.lw[pos % psi.N]

    def rightenv(self, pos, psi):
        if self.dependency is not psi:
            self.recalculate(psi)
        return self.rw[pos % psi.N]


def calc_galerkin_mixed(psi, loc, envs):
    """toolbox.jl:26-38: AC' = normalize(dAC * above.AC); || AC' - AL AL' AC' ||"""
    above = psi if envs.above is None else envs.above
    y = dAC_sparse(above.AC[loc], envs.opp[loc], envs.leftenv(loc, psi), envs.rightenv(loc, psi))
    y = y / np.linalg.norm(y)
    AL = psi.AL[loc]
    return float(np.linalg.norm(y - np.einsum('apb,cqb,cqd->apd', AL, np.conj(AL), y, optimize=True)))


def vomps(psi: InfiniteMPS, slices, above: InfiniteMPS = None, tol=1e-12, maxiter=100, log=None):
    """approximate(psi, (O, above), VOMPS()) (src/algorithms/approximate/vomps.jl:19-80) for one row; above = None:
    leading_boundary(psi, O, VOMPS()) (src/algorithms/statmech/vomps.jl:27-87), the power method on the current state"""
    envs = MixedPerMPOInfEnv(psi, slices, above)
    N = psi.N
    eps = max(calc_galerkin_mixed(psi, i, envs) for i in range(N))
    for it in range(1, maxiter + 1):
        top = psi if above is None else above
        newAL = []
        for col in range(N):
            ac = dAC_sparse(top.AC[col], slices[col], envs.leftenv(col, psi), envs.rightenv(col, psi))        # ac_proj
            c = dC_dense(top.CR[col], envs.leftenv(col + 1, psi), envs.rightenv(col, psi))                    # c_proj
            newAL.append(regauge(ac, c, psi.V[col], psi.V[(col + 1) % N], psi.pq, psi.modulus))
        gtol = dyn_tol(eps, it, TOLGAUGE, 1e-5, 1e-8)
        psi = InfiniteMPS.from_AL(newAL, psi.CR[N - 1], psi.V, psi.pq, gtol, 100, psi.modulus)
        envs.recalculate(psi, dyn_tol(eps, it, 1e-12, 1e-5, 1e-5))
        eps = max(calc_galerkin_mixed(psi, i, envs) for i in range(N))
        if log is not None:
            log.append(dict(iter=it, eps=eps))
        if eps <= tol:
            break
    return psi, envs, eps


def approximate_idmrg1(ost: InfiniteMPS, slices, above: InfiniteMPS, tol=1e-12, tol_gauge=TOLGAUGE, maxiter=100, log=None):
    """approximate(ost, (mpo, above), IDMRG1()) (src/algorithms/approximate/idmrg.jl:1-58) for one row"""
    from .linalg import qrpos as _qr, lqpos as _lq
    envs0 = MixedPerMPOInfEnv(ost, slices, above)
    N, pq, mod = ost.N, ost.pq, ost.modulus
    lw = [envs0.leftenv(i, ost).copy() for i in range(N)]
    rw = [envs0.rightenv(i, ost).copy() for i in range(N)]
    AL, AR, AC, CR = list(ost.AL), list(ost.AR), list(ost.AC), list(ost.CR)
    eps = 2 * tol
    for it in range(1, maxiter + 1):
        Ccur = CR[N - 1]
        for col in range(N):
            ac = dAC_sparse(above.AC[col], slices[col], lw[col], rw[col])
            ac = ac / np.linalg.norm(ac)
            AC[col] = ac
            AL[col], CR[col] = InfiniteMPS.leftorth(ac, ost.V[col], ost.V[(col + 1) % N], pq, mod)
            nl = transfer_left_sparse(lw[col], slices[col], above.AL[col], AL[col])
            lw[(col + 1) % N] = nl / np.linalg.norm(nl)
        for col in range(N - 1, -1, -1):
            ac = dAC_sparse(above.AC[col], slices[col], lw[col], rw[col])
            ac = ac / np.linalg.norm(ac)
            AC[col] = ac
            CR[(col - 1) % N], AR[col] = InfiniteMPS.rightorth(ac, ost.V[col], ost.V[(col + 1) % N], pq, mod)
            nr = transfer_right_sparse(rw[col], slices[col], above.AR[col], AR[col])
            rw[(col - 1) % N] = nr / np.linalg.norm(nr)
        eps = float(np.linalg.norm(Ccur - CR[N - 1]))
        if log is not None:
            log.append(dict(iter=it, eps=eps))
        if eps < tol:
            break
    nst = InfiniteMPS.from_tensors(AR, ost.V, pq, tol=tol_gauge, mod=mod)
    return nst, MixedPerMPOInfEnv(nst, slices, above), eps


def approximate_idmrg2(ost: InfiniteMPS, slices, above: InfiniteMPS, trunc=("truncerr", 1e-6), tol=1e-12, tol_gauge=TOLGAUGE,
                       maxiter=100, log=None):
    """approximate(ost, (mpo, above), IDMRG2(trscheme)) (src/algorithms/approximate/idmrg.jl:60-145) for one row"""
    from .linalg import tsvd as _tsvd
    from .contract import dAC2_sparse
    if ost.N < 2:
        raise ValueError("unit cell should be >= 2")
    envs0 = MixedPerMPOInfEnv(ost, slices, above)
    N, pq, mod = ost.N, ost.pq, ost.modulus
    d = len(pq)
    lw = [envs0.leftenv(i, ost).copy() for i in range(N)]
    rw = [envs0.rightenv(i, ost).copy() for i in range(N)]
    AL, AR, CR, V = list(ost.AL), list(ost.AR), list(ost.CR), list(ost.V)
    eps = 2 * tol

    def step(col, x2):
        c1 = (col + 1) % N
        y = dAC2_sparse(x2, slices[col % N], slices[c1], lw[col % N], rw[c1])
        Dl, _, Dr, _ = y.shape
        Vl, Vr = V[col % N], V[(col + 2) % N]
        rowq = (Vl.index_charges()[:, None] + pq[None, :]).reshape(-1)
        colq = (-pq[:, None] + Vr.index_charges()[None, :]).reshape(-1)
        U, S, Vh, newV, _, _ = _tsvd(y.transpose(0, 1, 3, 2).reshape(Dl * d, d * Dr), rowq, colq, trunc, mod)
        S = S / np.linalg.norm(S)
        AL[col % N] = U.reshape(Dl, d, newV.dim)
        CR[col % N] = np.diag(S).astype(np.complex128)
        AR[c1] = Vh.reshape(newV.dim, d, Dr)
        V[c1] = newV
        nl = transfer_left_sparse(lw[col % N], slices[col % N], above.AL[col % N], AL[col % N])
        lw[c1] = nl / np.linalg.norm(nl)
        nr = transfer_right_sparse(rw[c1], slices[c1], above.AR[c1], AR[c1])
        rw[col % N] = nr / np.linalg.norm(nr)
    for it in range(1, maxiter + 1):
        Ccur, Vcur = CR[N - 1], V[0]
        for col in range(N):
            step(col, np.einsum('apk,kqr->aprq', above.AC[col], above.AR[(col + 1) % N]))
        for col in range(N - 2, -2, -1):
            step(col, np.einsum('apk,kqr->aprq', above.AL[col % N], above.AC[(col + 1) % N]))
        c2, V2 = CR[N - 1], V[0]
        acc = 0.0
        q1, q2 = Vcur.index_charges(), V2.index_charges()
        for q in set(Vcur.charges) & set(V2.charges):
            x, y = Ccur[np.ix_(q1 == q, q1 == q)], c2[np.ix_(q2 == q, q2 == q)]
            k = min(x.shape[0], y.shape[0])
            acc += float(np.linalg.norm(y[:k, :k] - x[:k, :k]) ** 2)
        eps = float(np.sqrt(acc))
        if log is not None:
            log.append(dict(iter=it, eps=eps, D=V[0].dim))
        if eps < tol:
            break
    nst = InfiniteMPS.from_tensors(AR, V, pq, tol=tol_gauge, mod=mod)
    return nst, MixedPerMPOInfEnv(nst, slices, above), eps


def vomps_multiline(rows, slices, above_rows, tol=1e-12, maxiter=100, log=None):
    """approximate(psi::MPSMultiline, (O::MPOMultiline, psi0::MPSMultiline), VOMPS()) (src/algorithms/approximate/vomps.jl:19-80):
    ac_proj / c_proj of row r (ket psi0[r], operator O[r], bra psi[r + 1]; derivatives.jl:204-219) are circshifted by one row,
    i.e. row r + 1 of psi approximates O[r] psi0[r]; the mixed environments of different rows do not couple, so the update is
    the row-wise map of the single-row one (the reference stops on the maximum of the Galerkin errors; so does this)."""
    R = len(rows)
    out, envs, eps = [None] * R, [None] * R, 0.0
    for r in range(R):
        out[(r + 1) % R], envs[r], e = vomps(rows[(r + 1) % R], slices[r], above_rows[r], tol, maxiter, log)
        eps = max(eps, e)
    return out, envs, eps
