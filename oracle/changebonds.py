"""TEST INFRASTRUCTURE (CPU oracle, never imported by the product path).

changebonds for InfiniteMPS restated on dense charged tensors:
  * _expand!                      src/algorithms/changebonds/changebonds.jl:13-38
  * OptimalExpand (infinite)      src/algorithms/changebonds/optimalexpand.jl:16-38
  * RandExpand (infinite)         src/algorithms/changebonds/randexpand.jl:15-35
  * SvdCut (infinite)             src/algorithms/changebonds/svdcut.jl:36-46
  * VUMPSSvdCut                   src/algorithms/changebonds/vumpssvd.jl:19-86
  * dot(psi1, psi2)               src/states/infinitemps.jl (leading eigenvalue of the mixed transfer matrix)
Sites are 0-based and periodic; V[i] is the bond space to the left of site i."""
from __future__ import annotations
import numpy as np
from .spaces import Space
from .linalg import tsvd, qrpos
from .finite import _null_blocks
from .contract import dAC2_sparse, dC_dense, transfer_left_bond
from .krylov import eigsolve_lanczos
from .infinite import InfiniteMPS, MPOHamInfEnv, TOLGAUGE


def _sum_space(Vold: Space, Vnew: Space, mod):
    """direct sum with the old indices first inside every sector; returns (space, positions of old, positions of new)"""
    do, dn = Vold.as_dict(), Vnew.as_dict()
    both = Space.from_dict({q: do.get(q, 0) + dn.get(q, 0) for q in set(do) | set(dn)}, mod)
    idx_old, idx_new, off = [], [], 0
    for q, n in zip(both.charges, both.dims):
        no = do.get(q, 0)
        idx_old += list(range(off, off + no)); idx_new += list(range(off + no, off + n)); off += n
    return both, np.array(idx_old, dtype=int), np.array(idx_new, dtype=int)


def expand(st: InfiniteMPS, ALp, ARp, newV):
    """_expand!(psi, AL', AR') (changebonds.jl:13-38): AL -> [AL AL'; 0 0], AR -> [AR 0; AR' 0], C -> [C 0; 0 0], AC = AL C;
    ALp[i][l, p, n_i] lives on the bond right of site i, ARp[i][n_{i-1}, p, r] on the bond left of site i, newV[i] = n_i"""
    N, mod = st.N, st.modulus
    d = len(st.pq)
    sums = [_sum_space(st.V[(i + 1) % N], newV[i], mod) for i in range(N)]        # bond right of site i
    AL, AR, CR, AC, V = [None] * N, [None] * N, [None] * N, [None] * N, [None] * N
    for i in range(N):
        (Sl, lo, ln), (Sr, ro, rn) = sums[(i - 1) % N], sums[i]
        a = np.zeros((Sl.dim, d, Sr.dim), dtype=np.complex128)
        a[np.ix_(lo, range(d), ro)] = st.AL[i]
        a[np.ix_(lo, range(d), rn)] = ALp[i]
        AL[i] = a
        b = np.zeros((Sl.dim, d, Sr.dim), dtype=np.complex128)
        b[np.ix_(lo, range(d), ro)] = st.AR[i]
        b[np.ix_(ln, range(d), ro)] = ARp[i]
        AR[i] = b
        c = np.zeros((Sr.dim, Sr.dim), dtype=np.complex128)
        c[np.ix_(ro, ro)] = st.CR[i]
        CR[i] = c
        V[i] = Sl
    nrm = [np.linalg.norm(c) for c in CR]
    CR = [c / n for c, n in zip(CR, nrm)]                                        # normalize!(psi): CR and AC separately
    for i in range(N):
        AC[i] = np.einsum('apb,bc->apc', AL[i], CR[i])
        AC[i] = AC[i] / np.linalg.norm(AC[i])
    return InfiniteMPS(AL, AR, CR, AC, V, st.pq, mod)


def _expansion_vectors(st, i, AC2mat, trunc, log):
    """VL = leftnull(AL[i]), VR = rightnull(AR[i+1]); U, _, V = tsvd(VL' AC2 VR'; trunc); returns VL U, V VR, the added space"""
    N, pq, mod = st.N, st.pq, st.modulus
    d = len(pq)
    Vl, Vk, Vr = st.V[i % N], st.V[(i + 1) % N], st.V[(i + 2) % N]
    Dl, Dk, Dr = Vl.dim, Vk.dim, Vr.dim
    rowq = (Vl.index_charges()[:, None] + pq[None, :]).reshape(-1)
    colq = (-pq[:, None] + Vr.index_charges()[None, :]).reshape(-1)
    NL, nlq = _null_blocks(st.AL[i % N].reshape(Dl * d, Dk), rowq, Vk.index_charges())
    NRh, nrq = _null_blocks(st.AR[(i + 1) % N].reshape(Dk, d * Dr).conj().T, colq, Vk.index_charges())
    inter = NL.conj().T @ AC2mat @ NRh
    U, S, Vh, newV, eps, info = tsvd(inter, nlq, nrq, trunc, mod)
    if log is not None:
        log.append(dict(pos=i, added=newV.as_dict(), spectrum={q: s.copy() for q, s in info["spectrum"].items()}))
    return (NL @ U).reshape(Dl, d, newV.dim), (Vh @ NRh.conj().T).reshape(newV.dim, d, Dr), newV


def optimal_expand(st: InfiniteMPS, H, trunc=("truncdim", 1), envs=None, log=None):
    """changebonds(psi::InfiniteMPS, H, alg::OptimalExpand, envs) (optimalexpand.jl:16-38)"""
    envs = envs or MPOHamInfEnv(st, H)
    N, d = st.N, len(st.pq)
    ALp, ARp, newV = [None] * N, [None] * N, [None] * N
    for i in range(N):
        ac2 = np.einsum('apk,kqr->aprq', st.AC[i], st.AR[(i + 1) % N])
        AC2 = dAC2_sparse(ac2, H[i], H[i + 1], envs.leftenv(i, st), envs.rightenv(i + 1, st))
        Mat = AC2.transpose(0, 1, 3, 2).reshape(ac2.shape[0] * d, d * ac2.shape[2])
        ALp[i], ARp[(i + 1) % N], newV[i] = _expansion_vectors(st, i, Mat, trunc, log)
    return expand(st, ALp, ARp, newV), envs


def rand_expand(st: InfiniteMPS, trunc=("truncdim", 1), rng=None, log=None):
    """changebonds(psi::InfiniteMPS, alg::RandExpand) (randexpand.jl:15-35): the same with a random two-site tensor"""
    rng = rng or np.random.default_rng(0)
    N, d, pq = st.N, len(st.pq), st.pq
    ALp, ARp, newV = [None] * N, [None] * N, [None] * N
    for i in range(N):
        Vl, Vr = st.V[i], st.V[(i + 2) % N]
        rowq = (Vl.index_charges()[:, None] + pq[None, :]).reshape(-1)
        colq = (-pq[:, None] + Vr.index_charges()[None, :]).reshape(-1)
        Mat = rng.standard_normal((len(rowq), len(colq))) + 1j * rng.standard_normal((len(rowq), len(colq)))
        diff = rowq[:, None] - colq[None, :]
        Mat = Mat * ((diff % st.modulus == 0) if st.modulus else (diff == 0))
        ALp[i], ARp[(i + 1) % N], newV[i] = _expansion_vectors(st, i, Mat, trunc, log)
    return expand(st, ALp, ARp, newV)


def svdcut(st: InfiniteMPS, trunc=("notrunc",), log=None, tol=TOLGAUGE):
    """changebonds(psi::InfiniteMPS, alg::SvdCut) (svdcut.jl:36-46)"""
    N, mod = st.N, st.modulus
    copied = [a.astype(np.complex128) for a in st.AL]
    V = list(st.V)
    ncr = st.CR[0]
    for i in range(N):
        Vb = st.V[(i + 1) % N]
        U, S, Vh, newV, eps, info = tsvd(st.CR[i], Vb.index_charges(), Vb.index_charges(), trunc, mod)
        if log is not None:
            log.append(dict(pos=i, kept=newV.as_dict(), eps_trunc=eps))
        ncr = np.diag(S).astype(np.complex128)
        copied[i] = np.einsum('apk,kn->apn', copied[i], U)
        copied[(i + 1) % N] = np.einsum('kn,kpr->npr', U.conj(), copied[(i + 1) % N])
        V[(i + 1) % N] = newV
    out = InfiniteMPS.from_AL(copied, ncr, V, st.pq, tol, 100, mod)
    out.CR = [c / np.linalg.norm(c) for c in out.CR]
    out.AC = [a / np.linalg.norm(a) for a in out.AC]
    return out


def dot(a: InfiniteMPS, b: InfiniteMPS):
    """dot(psi1, psi2) for InfiniteMPS: dominant eigenvalue of the mixed transfer matrix of the AL tensors (dense)"""
    Da, Db = a.AL[0].shape[0], b.AL[0].shape[0]
    n = Da * Db
    if n > 600:                                    # large mixed transfer matrices: implicitly restarted Arnoldi instead of dense
        import scipy.sparse.linalg as spl

        def mv(x):
            v = x.reshape(Da, Db)
            for s_ in range(a.N):
                v = transfer_left_bond(v, b.AL[s_], a.AL[s_])
            return v.reshape(-1)
        rng = np.random.default_rng(0)
        ev = spl.eigs(spl.LinearOperator((n, n), matvec=mv, dtype=np.complex128), k=1, which="LM",
                      v0=rng.standard_normal(n) + 1j * rng.standard_normal(n), tol=1e-13)[0]
        return ev[0]
    T = np.zeros((n, n), dtype=np.complex128)
    for k in range(n):
        v = np.zeros(n, dtype=np.complex128); v[k] = 1
        v = v.reshape(Da, Db)
        for s in range(a.N):
            v = transfer_left_bond(v, b.AL[s], a.AL[s])
        T[:, k] = v.reshape(-1)
    ev = np.linalg.eigvals(T)
    return ev[np.argmax(np.abs(ev))]


def vumps_svdcut(st: InfiniteMPS, H, trunc=("notrunc",), tol_gauge=TOLGAUGE, tol_eigenval=1e-12, envs=None, log=None):
    """changebonds(state::InfiniteMPS, H, alg::VUMPSSvdCut, envs) for unit cells of >= 2 sites (changebonds_n,
    vumpssvd.jl:49-86): per site a two-site eigensolve, a bond eigensolve, the truncated SVD of the two-site tensor and the
    VUMPS regauging of the right half; the state is re-gauged after every site and the environments follow the new state
    (leftenv(envs, pos, state) recalculates when the state object changed)."""
    envs = envs or MPOHamInfEnv(st, H)
    if st.N < 2:
        raise ValueError("unit cell of one site: double it first (changebonds_1, vumpssvd.jl:19-47)")
    pq, mod = st.pq, st.modulus
    d = len(pq)
    for loc in range(st.N):
        N = st.N
        if envs.dependency is not st:                 # leftenv(envs, pos, state) recalculates for a new state (abstractinfenv)
            envs = MPOHamInfEnv(st, H)
        A1, A2 = st.AC[loc], st.AR[(loc + 1) % N]
        ac2 = np.einsum('apk,kqr->aprq', A1, A2)
        GL, GR = envs.leftenv(loc, st), envs.rightenv(loc + 1, st)
        _, nac2, _ = eigsolve_lanczos(lambda x: dAC2_sparse(x, H[loc], H[loc + 1], GL, GR), ac2, "SR", tol_eigenval, 30, 100)
        GLc, GRc = envs.leftenv(loc + 2, st), envs.rightenv(loc + 1, st)
        _, nc2, _ = eigsolve_lanczos(lambda x: dC_dense(x, GLc, GRc), st.CR[(loc + 1) % N], "SR", tol_eigenval, 30, 100)
        Vl, Vr = st.V[loc], st.V[(loc + 2) % N]
        Dl, Dr = Vl.dim, Vr.dim
        rowq = (Vl.index_charges()[:, None] + pq[None, :]).reshape(-1)
        colq = (-pq[:, None] + Vr.index_charges()[None, :]).reshape(-1)
        Mat = nac2.transpose(0, 1, 3, 2).reshape(Dl * d, d * Dr)
        U, S, Vh, newV, eps, info = tsvd(Mat, rowq, colq, trunc, mod)
        if log is not None:
            log.append(dict(pos=loc, kept=newV.as_dict(), eps_trunc=eps))
        AL1 = U.reshape(Dl, d, newV.dim)
        AC = (np.diag(S) @ Vh).reshape(newV.dim, d, Dr)
        rq2 = (newV.index_charges()[:, None] + pq[None, :]).reshape(-1)
        QAC, _, Vq = qrpos(AC.reshape(newV.dim * d, Dr), rq2, Vr.index_charges(), mod)
        QC, _, Vc = qrpos(nc2, Vr.index_charges(), Vr.index_charges(), mod)
        assert Vq == Vc, "rectangular dom_map (vumpssvd.jl:72) not needed for the configurations tested"
        AL2 = (QAC @ QC.conj().T).reshape(newV.dim, d, Dr)
        copied = [a.copy() for a in st.AL]
        copied[loc] = AL1
        copied[(loc + 1) % N] = AL2
        V = list(st.V)
        V[(loc + 1) % N] = newV
        st = InfiniteMPS.from_tensors(copied, V, pq, tol=tol_gauge, mod=mod)
    return st, MPOHamInfEnv(st, H)


# ------------------------------------------------------------------------------------------ MPSMultiline
def optimal_expand_multiline(rows, slices, trunc=("truncdim", 1), envs=None, log=None):
    """changebonds(psi::MPSMultiline, H::MPOMultiline, alg::OptimalExpand, envs) (optimalexpand.jl:46-66): the two-site tensor of
    row i - 1, pushed through ddAC2(i - 1, j), lives in the bond spaces of row i; the null spaces of AL[i, j] and AR[i, j + 1]
    pick the expansion of row i.  rows: list of InfiniteMPS, slices[r][c]: MPO tensors, envs: statmech.MultilineEnv.  The
    single-row case is changebonds(psi::InfiniteMPS, H::DenseMPO, alg::OptimalExpand) (optimalexpand.jl:40-44)."""
    from .statmech import MultilineEnv
    envs = envs or MultilineEnv(rows, slices)
    R, N = len(rows), rows[0].N
    d = len(rows[0].pq)
    out = []
    for i in range(R):
        top, r = rows[(i - 1) % R], (i - 1) % R
        ALp, ARp, newV = [None] * N, [None] * N, [None] * N
        for j in range(N):
            ac2 = np.einsum('apk,kqr->aprq', top.AC[j], top.AR[(j + 1) % N])
            AC2 = dAC2_sparse(ac2, slices[r][j], slices[r][(j + 1) % N], envs.lw[r][j], envs.rw[r][(j + 1) % N])
            Mat = AC2.transpose(0, 1, 3, 2).reshape(ac2.shape[0] * d, d * ac2.shape[2])
            ALp[j], ARp[(j + 1) % N], newV[j] = _expansion_vectors(rows[i], j, Mat, trunc, log)
        out.append(expand(rows[i], ALp, ARp, newV))
    return out, envs


def rand_expand_multiline(rows, trunc=("truncdim", 1), rng=None, log=None):
    """changebonds(psi::MPSMultiline, alg::RandExpand) (randexpand.jl:36-38): row-wise"""
    rng = rng or np.random.default_rng(0)
    return [rand_expand(st, trunc, rng, log) for st in rows]


def svdcut_multiline(rows, trunc=("notrunc",), log=None):
    """changebonds(psi::MPSMultiline, alg::SvdCut) (svdcut.jl:32-34): row-wise"""
    return [svdcut(st, trunc, log) for st in rows]


# ------------------------------------------------------------------------------------------ DenseMPO products
def _fused(qa, qw, mod):
    """fuser of a bond index a (charges qa) with an MPO level w (charges qw): the pair index (a, w) sorted into the sectors of
    the fused space (charge qa + qw); returns (Space, permutation of the a-major pair index)"""
    q = (np.asarray(qa)[:, None] + np.asarray(qw)[None, :]).reshape(-1)
    if mod:
        q = q % mod
    perm = np.argsort(q, kind="stable")
    cs, ns = np.unique(q, return_counts=True)
    return Space(tuple(int(c) for c in cs), tuple(int(n) for n in ns), mod), perm


def mpo_times_state(slices, st: InfiniteMPS, tol=TOLGAUGE):
    """Base.:*(mpo::DenseMPO, st::InfiniteMPS) (src/operators/densempo.jl:31-45): A'[(a w), p, (b v)] = sum_q AL[a, q, b]
    O[w, p, q, v] with the fusers sorting (a, w) into the sectors of charge q_a + c_w, then the InfiniteMPS constructor"""
    N, mod = st.N, st.modulus
    fus = [_fused(st.V[i].index_charges(), slices[i].cl, mod) for i in range(N)]
    A = []
    for i in range(N):
        t = np.einsum('aqb,wpqv->awpbv', st.AL[i], slices[i].dense())
        Dl, chi, d, Dr, chir = t.shape
        t = t.reshape(Dl * chi, d, Dr * chir)
        A.append(t[fus[i][1]][:, :, fus[(i + 1) % N][1]])
    return InfiniteMPS.from_tensors(A, [f[0] for f in fus], st.pq, tol=tol, mod=mod)


def densempo_as_state(slices):
    """convert(InfiniteMPS, mpo::DenseMPO) (densempo.jl:19-25): the two physical legs fused into one of charge p_out - p_in"""
    N = len(slices)
    pch = np.asarray(slices[0].pch)
    d = len(pch)
    pq = (pch[:, None] - pch[None, :]).reshape(-1)
    order = np.argsort(pq, kind="stable")
    A, V = [], []
    for i in range(N):
        cl, cr = np.asarray(slices[i].cl), np.asarray(slices[i].cr)
        pl, pr = np.argsort(cl, kind="stable"), np.argsort(cr, kind="stable")
        t = slices[i].dense().reshape(len(cl), d * d, len(cr))[:, order, :]
        A.append(t[pl][:, :, pr])
        cs, ns = np.unique(cl, return_counts=True)
        V.append(Space(tuple(int(c) for c in cs), tuple(int(n) for n in ns)))
    return InfiniteMPS.from_tensors(A, V, pq[order], mod=0), order


def svdcut_densempo(slices, trunc=("notrunc",)):
    """changebonds(mpo::DenseMPO, alg::SvdCut) (svdcut.jl:26-28): through the InfiniteMPS form (densempo.jl:18-22), SvdCut, and
    back (densempo.jl:24-28: the AL tensors with the physical leg split again).  Returns the dense tensors O'[w, p_out, p_in, v]
    per site and the bond spaces (level charges) of the cut state."""
    st, order = densempo_as_state(slices)
    cut = svdcut(st, trunc)
    d = len(slices[0].pch)
    inv = np.argsort(order)
    tensors = [cut.AL[i][:, inv, :].reshape(cut.AL[i].shape[0], d, d, cut.AL[i].shape[2]) for i in range(cut.N)]
    return tensors, list(cut.V)
