"""TEST INFRASTRUCTURE (CPU oracle, never imported by the product path).

Read-outs of the reference restated on dense (charged) numpy tensors:
  * transfer_spectrum / marek_gap / correlation_length   src/algorithms/toolbox.jl:42-58, 88-133
  * entropy / entanglement_spectrum                      src/algorithms/toolbox.jl:1-5, 65-69
  * expectation_value of local operators / local MPOs    src/algorithms/expval.jl:22-65
  * correlator                                           src/algorithms/correlators.jl:10-46
States are oracle.infinite.InfiniteMPS or oracle.finite.FiniteMPS (0-based sites).  A local MPO tensor is a dense array
O[wl, p_out, p_in, wr] (SURVEY.md Appendix A index order); single-site operators are d x d matrices O[p_out, p_in]."""
from __future__ import annotations
from fractions import Fraction
import numpy as np
from .contract import transfer_left_bond, transfer_left_dense


# ------------------------------------------------------------------------------------------ state accessors
def _is_inf(st):
    return hasattr(st, "N")


def _AL(st, i):
    return st.AL[i % st.N] if _is_inf(st) else st.get_AL(i)


def _AR(st, i):
    return st.AR[i % st.N] if _is_inf(st) else st.get_AR(i)


def _AC(st, i):
    return st.AC[i % st.N] if _is_inf(st) else st.get_AC(i)


def _CR(st, i):
    return st.CR[i % st.N] if _is_inf(st) else st.get_CR(i)


def _len(st):
    return st.N if _is_inf(st) else st.L


# ------------------------------------------------------------------------------------------ transfer spectrum
def transfer_spectrum(above, below=None, num_vals=20, sector=0):
    """toolbox.jl:42-58: the `num_vals` largest-magnitude eigenvalues of v -> v * TransferMatrix(above.AL, below.AL) over
    the unit cell on vectors v[bra (below); ket (above)] of charge `sector` (q_ket = q_bra + sector), by dense
    diagonalisation of the transfer matrix restricted to that sector."""
    below = above if below is None else below
    qb = below.V[0].index_charges(); qa = above.V[0].index_charges()
    mod = above.modulus
    diff = qa[None, :] - qb[:, None] - int(sector)
    mask = (diff % mod == 0) if mod else (diff == 0)
    idx = np.argwhere(mask)
    n = len(idx)
    T = np.zeros((n, n), dtype=np.complex128)
    for k, (a, b) in enumerate(idx):
        v = np.zeros(mask.shape, dtype=np.complex128); v[a, b] = 1.0
        for s in range(above.N):
            v = transfer_left_bond(v, above.AL[s], below.AL[s])
        T[:, k] = v[mask]
    ev = np.linalg.eigvals(T)
    ev = ev[np.argsort(-np.abs(ev), kind="stable")]
    return ev[:min(num_vals, n)]


def rationalize(x, tol):
    """simplest fraction within tol of x (what Julia's rationalize(x; tol) returns for the angles used here):
    continued-fraction convergents until the error is within tol."""
    x = float(x)
    if abs(x - round(x)) <= tol:
        return Fraction(int(round(x)), 1)
    sign = -1 if x < 0 else 1
    y = abs(x)
    p0, q0, p1, q1 = 0, 1, 1, 0
    r = y
    for _ in range(64):
        a = int(np.floor(r))
        p0, q0, p1, q1 = p1, q1, a * p1 + p0, a * q1 + q0
        if abs(y - p1 / q1) <= tol:
            break
        f = r - a
        if f == 0:
            break
        r = 1.0 / f
    return Fraction(sign * p1, q1)


def approx_angles(spectrum, tol_angle=0.1):
    """toolbox.jl:74-83"""
    ang = np.angle(np.asarray(spectrum)) / np.pi
    out = [rationalize(a, tol_angle) for a in ang]
    out = [Fraction(1, 1) if a == -1 else a for a in out]
    return np.array([float(a) * np.pi for a in out])


def marek_gap(spectrum, tol_angle=0.1):
    """toolbox.jl:93-117: (epsilon, delta, theta) from a transfer spectrum"""
    spectrum = np.asarray(spectrum)
    spectrum = spectrum[np.abs(spectrum) < 1 - 1e-12]
    angles = approx_angles(spectrum, tol_angle)
    theta = angles[0]
    at = spectrum[angles == theta]
    lambdas = -np.log(np.abs(at))
    eps = lambdas[0]
    delta = np.inf
    if len(lambdas) > 2:
        delta = lambdas[1] - lambdas[0]
    return float(eps), float(delta), float(theta)


def correlation_length(st_or_spectrum, **kw):
    """toolbox.jl:122-133"""
    spec = transfer_spectrum(st_or_spectrum, **kw) if hasattr(st_or_spectrum, "AL") else st_or_spectrum
    return 1.0 / marek_gap(spec)[0]


# ------------------------------------------------------------------------------------------ entanglement
def entanglement_spectrum(st, site=0):
    """toolbox.jl:65-69: per-sector singular values of CR[site] (descending)"""
    Cm = _CR(st, site)
    V = st.V[(site + 1) % st.N] if _is_inf(st) else st.V[site + 1]
    q = V.index_charges()
    return {int(c): np.linalg.svd(Cm[np.ix_(q == c, q == c)], compute_uv=False) for c in V.charges}


def entropy(st, loc=None):
    """toolbox.jl:1-5: -tr(safe_xlogx(CR CR')) at one bond, or at every bond of the unit cell"""
    if loc is None:
        return [entropy(st, i) for i in range(_len(st))]
    s2 = np.concatenate([v ** 2 for v in entanglement_spectrum(st, loc).values()])
    s2 = s2[s2 > 0]
    return float(-np.sum(s2 * np.log(s2)))


# ------------------------------------------------------------------------------------------ local expectation values
def expectation_value_site(st, O):
    """expval.jl:22-36: <O_i> for a single-site operator (or one operator per site) on every site"""
    n = _len(st)
    ops = [O] * n if np.asarray(O).ndim == 2 else list(O)
    return np.array([np.einsum('asb,st,atb->', np.conj(_AC(st, i)), np.asarray(ops[i]), _AC(st, i)) for i in range(n)])


def expectation_value_localmpo(st, Os, at):
    """expval.jl:45-65: identity on the left bond of site `at`, transferred through AL[at : at + n] with the local MPO
    tensors O[wl, p_out, p_in, wr] (trivial outer legs) in the middle, closed with CR CR' of the last site"""
    n = len(Os)
    D = _AL(st, at).shape[0]
    v = np.zeros((D, Os[0].shape[0], D), dtype=np.complex128)
    v[:, 0, :] = np.eye(D)
    for k in range(n):
        A = _AL(st, at + k)
        v = transfer_left_dense(v, np.asarray(Os[k], dtype=np.complex128), A, A)
    Cm = _CR(st, at + n - 1)
    return complex(np.einsum('ac,cd,ad->', v[:, 0, :], Cm, np.conj(Cm)))


def decompose_localmpo(O, tol=1e-14):
    """decompose an n-site operator O[p1_out .. pn_out, p1_in .. pn_in] into n MPO tensors [wl, p_out, p_in, wr]
    (utility.jl decompose_localmpo after add_util_leg) by successive SVDs"""
    O = np.asarray(O, dtype=np.complex128)
    n = O.ndim // 2
    d = O.shape[0]
    perm = [k for i in range(n) for k in (i, n + i)]                    # (p1o p1i p2o p2i ...)
    rest = O.transpose(perm).reshape(1, *([d, d] * n))
    out = []
    for k in range(n - 1):
        w = rest.shape[0]
        M = rest.reshape(w * d * d, -1)
        U, s, Vh = np.linalg.svd(M, full_matrices=False)
        keep = max(1, int(np.sum(s > tol * s[0])))
        out.append(U[:, :keep].reshape(w, d, d, keep))
        rest = (s[:keep, None] * Vh[:keep]).reshape(keep, *([d, d] * (n - 1 - k)))
    out.append(rest.reshape(rest.shape[0], d, d, 1))
    return out


def expectation_value_nsite(st, O, at):
    """expval.jl:41-44: n-site operator at site `at`"""
    return expectation_value_localmpo(st, decompose_localmpo(O), at)


def correlator(st, O1, O2, i, js):
    """correlators.jl:14-39: G[k] = <O1[i] O2[js[k]]> with MPO tensors O1[1, p_out, p_in, w], O2[w, p_out, p_in, 1]"""
    js = list(js)
    assert js[0] > i
    O1 = np.asarray(O1, dtype=np.complex128); O2 = np.asarray(O2, dtype=np.complex128)
    AC = _AC(st, i)
    Vl = np.einsum('cdb,efdg,cfa->agb', AC, O1[0:1], np.conj(AC))[:, :, :]      # [bra', w, ket']
    ctr = i + 1
    G = []
    for j in js:
        while ctr < j:
            A = _AR(st, ctr)
            Vl = np.einsum('awb,bpc,apd->dwc', Vl, A, np.conj(A))
            ctr += 1
        A = _AR(st, j)
        G.append(complex(np.einsum('bce,efg,cdfa,bdg->', Vl, A, O2[:, :, :, 0:1], np.conj(A))))
        ctr = j
    return np.array(G)
