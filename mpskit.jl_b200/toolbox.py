"""Read-outs of src/algorithms/toolbox.jl, src/algorithms/expval.jl (local operators) and src/algorithms/correlators.jl on
device handles.  Everything that scales with the bond dimension runs in the library: the transfer spectrum is a
Krylov-Schur eigensolve of the transfer-matrix operator (mpsk_tmop_create + mpsk_eigsolve_values), entanglement spectra are
per-sector SVDs of the bond matrix (mpsk_tsvd), local expectation values and correlators are sparse-slice transfers
(mpsk_transfer_left) closed by an inner product (mpsk_dot); only scalars and spectra reach the host.

A local MPO tensor is a dense array O[wl, p_out, p_in, wr] (the reference's MPOTensor, docs/src/man/intro.md:77-83) with
optional integer charges of its virtual levels; single-site operators are d x d matrices O[p_out, p_in]."""
from __future__ import annotations
import ctypes as C
import warnings
from fractions import Fraction
import numpy as np
from . import _lib
from .tensor import DeviceTensor
from .operators import SparseMPOSlice


# ------------------------------------------------------------------------------------------ state accessors
def _is_inf(st):
    return hasattr(st, "N") and not hasattr(st, "L")


def _at(st, name, i):
    view = getattr(st, name)
    return view[i % st.N] if _is_inf(st) else view[i]


def _len(st):
    return st.N if _is_inf(st) else len(st)


def _identity(ctx, V, dtype, levels=1, level_charges=None):
    """[V; levels; V] with the identity on level 0 and zeros elsewhere (isomorphism (x) conj(ut), expval.jl:56-58)"""
    mch = [0] * levels if level_charges is None else [-int(c) for c in level_charges]
    eye = np.zeros((V.dim, levels, V.dim), dtype=dtype)
    eye[:, 0, :] = np.eye(V.dim)
    return DeviceTensor.from_dense(ctx, eye, V, mch, V, dtype)


def _transfer_left(v, sl, A, Ab):
    ctx = A.ctx
    h = C.c_void_p()
    ctx.check(ctx.lib.mpsk_transfer_left(ctx.h, sl.device(ctx), A.h, Ab.h, v.h, C.byref(h)))
    return DeviceTensor.wrap(ctx, h, A.Vl.modulus)


# ------------------------------------------------------------------------------------------ transfer spectrum
def transfer_spectrum(above, below=None, tol=1e-12, num_vals=20, sector=0, krylovdim=None, maxiter=100, seed=1234, return_info=False):
    """transfer_spectrum(above; below, tol, num_vals, sector) (toolbox.jl:42-58): the `num_vals` largest-magnitude
    eigenvalues of flip(TransferMatrix(above.AL, below.AL)) on vectors v[bra (below); ket (above)] of charge `sector`
    (q_ket = q_bra + sector), by the library's Krylov-Schur Arnoldi on the unit-cell transfer operator."""
    from .infinite import TransferMatrix
    below = above if below is None else below
    ctx = above.ctx
    init = DeviceTensor.random(ctx, below.V[0], (int(sector),), above.V[0], seed, above.dtype)
    nelem = sum(below.V[0].dims[jl] * above.V[0].dims[jr] for (_, jl, jr) in _block_table(init))
    if nelem == 0:
        raise _lib.SpaceMismatch(_lib.ESPACE, f"no transfer-matrix vectors of charge {sector}")
    num = min(int(num_vals), nelem)                   # toolbox.jl:49: we can ask at most this many values
    kd = int(krylovdim) if krylovdim is not None else max(30, 2 * num)
    kd = max(min(kd, max(nelem, 2)), (5 * num + 2) // 3 + 1)
    tm = TransferMatrix(list(above.AL), None, list(below.AL)).flip()
    op = tm.device_op(init)
    vals = (C.c_double * (2 * num))()
    nvals, nconv = C.c_int32(), C.c_int32()
    info = _lib.c_info()
    ctx.check(ctx.lib.mpsk_eigsolve_values(ctx.h, op.h, init.h, _lib.LM, num, kd, maxiter, tol, vals, C.byref(nvals), C.byref(nconv),
                                           C.byref(info)))
    init.free()
    if nconv.value < num:                              # toolbox.jl:54-55
        warnings.warn(f"correlation length failed to converge: normres = {info.normres}")
    out = np.array([complex(vals[2 * j], vals[2 * j + 1]) for j in range(nvals.value)])
    if return_info:
        return out, dict(converged=nconv.value, numiter=info.numiter, numops=info.numops, normres=info.normres)
    return out


def _block_table(t):
    from .tensor import block_table
    return block_table(t.Vl, t.mch, t.Vr)


def rationalize(x, tol):
    """simplest fraction within tol of x (Julia's rationalize(x; tol) for the angles of toolbox.jl:74-83)"""
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
    """toolbox.jl:74-83: the closest fractions of pi"""
    ang = np.angle(np.asarray(spectrum)) / np.pi
    out = [rationalize(a, tol_angle) for a in ang]
    out = [Fraction(1, 1) if a == -1 else a for a in out]
    return np.array([float(a) * np.pi for a in out])


def marek_gap(above_or_spectrum, tol_angle=0.1, **kw):
    """marek_gap (toolbox.jl:88-117): (epsilon, delta, theta) of a state's transfer spectrum (or of a given spectrum)"""
    spectrum = transfer_spectrum(above_or_spectrum, **kw) if hasattr(above_or_spectrum, "AL") else np.asarray(above_or_spectrum)
    inds = np.abs(spectrum) < 1 - 1e-12
    if len(spectrum) - int(inds.sum()) >= 2:
        warnings.warn("Non-injective mps?")
    spectrum = spectrum[inds]
    angles = approx_angles(spectrum, tol_angle)
    theta = angles[0]
    lambdas = -np.log(np.abs(spectrum[angles == theta]))
    eps = lambdas[0]
    delta = np.inf
    if len(lambdas) > 2:
        delta = lambdas[1] - lambdas[0]
    return float(eps), float(delta), float(theta)


def correlation_length(above_or_spectrum, **kw):
    """correlation_length (toolbox.jl:122-133) = 1 / epsilon"""
    return 1.0 / marek_gap(above_or_spectrum, **kw)[0]


# ------------------------------------------------------------------------------------------ entanglement
def entanglement_spectrum(st, site=0):
    """entanglement_spectrum(psi, site) (toolbox.jl:65-69): {charge: singular values of CR[site]}; the bond matrix is
    factorised per sector on the device, only the spectrum comes back"""
    Cm = _at(st, "CR", site)
    ctx = st.ctx
    pc = (C.c_int32 * 1)(0)
    U, S, Vh = C.c_void_p(), C.c_void_p(), C.c_void_p()
    eps = C.c_double()
    ctx.check(ctx.lib.mpsk_tsvd(ctx.h, Cm.h, 1, pc, _lib.c_trunc(_lib.NOTRUNC | 0x100, 0.0), C.byref(U), C.byref(S), C.byref(Vh), C.byref(eps)))
    for h in (U, S, Vh):
        ctx.lib.mpsk_tensor_free(ctx.h, h)
    from .dmrg import tsvd_spectrum
    return {q: vals for q, (vals, kept) in tsvd_spectrum(ctx).items()}


def entropy(st, loc=None):
    """entropy(state) / entropy(state, loc) (toolbox.jl:1-5): -tr(safe_xlogx(CR CR'))"""
    if loc is None:
        return [entropy(st, i) for i in range(_len(st))]
    s2 = np.concatenate([np.asarray(v) ** 2 for v in entanglement_spectrum(st, loc).values()])
    s2 = s2[s2 > 0]
    return float(-np.sum(s2 * np.log(s2)))


# ------------------------------------------------------------------------------------------ local operators
def local_mpo_slices(Os, pcharges, level_charges=None, modulus=0):
    """local MPO tensors O_k[wl, p_out, p_in, wr] -> square SparseMPOSlices of one common number of levels (missing levels
    are empty rows / columns; the sparse-slice transfer contracts them to zero like transfer.jl:182-184).
    level_charges[k]: charges of the virtual bond to the left of tensor k (len(Os) + 1 lists; default: all zero)."""
    Os = [np.asarray(O, dtype=np.complex128) for O in Os]
    n = len(Os)
    od = max(max(O.shape[0], O.shape[3]) for O in Os)
    if level_charges is None:
        level_charges = [[0] * (Os[0].shape[0])] + [[0] * O.shape[3] for O in Os]
    out = []
    for k, O in enumerate(Os):
        cl = list(level_charges[k]) + [0] * (od - len(level_charges[k]))
        cr = list(level_charges[k + 1]) + [0] * (od - len(level_charges[k + 1]))
        ent = [[0 for _ in range(od)] for _ in range(od)]
        for i in range(O.shape[0]):
            for j in range(O.shape[3]):
                if np.any(O[i, :, :, j] != 0):
                    ent[i][j] = O[i, :, :, j]
        out.append(SparseMPOSlice(ent, cl, cr, pcharges, modulus))
    return out, od, level_charges


def decompose_localmpo(O, pq=None, tol=1e-14, return_charges=False):
    """n-site operator O[p1_out .. pn_out, p1_in .. pn_in] -> n MPO tensors [wl, p_out, p_in, wr] (decompose_localmpo of
    add_util_leg(O), expval.jl:41-44) by successive SVDs, one per charge of the cut virtual bond (the rows (w, p_out, p_in) of
    each unfolding are grouped by c_w + s_out - s_in so that every new level carries a definite charge); host-side model
    building on d^n x d^n numbers"""
    O = np.asarray(O, dtype=np.complex128)
    n = O.ndim // 2
    d = O.shape[0]
    pq = [0] * d if pq is None else [int(c) for c in pq]
    perm = [k for i in range(n) for k in (i, n + i)]
    rest = O.transpose(perm).reshape(1, *([d, d] * n))
    out, charges = [], [[0]]
    for k in range(n - 1):
        w = rest.shape[0]
        M = rest.reshape(w * d * d, -1)
        rowq = np.array([charges[-1][i] + pq[a] - pq[b] for i in range(w) for a in range(d) for b in range(d)])
        scale = np.abs(M).max() or 1.0
        Ucols, Rrows, cr = [], [], []
        for q in sorted(set(rowq.tolist())):
            rows = np.nonzero(rowq == q)[0]
            U, s, Vh = np.linalg.svd(M[rows], full_matrices=False)
            keep = int(np.sum(s > tol * scale))
            for t in range(keep):
                col = np.zeros(M.shape[0], dtype=np.complex128); col[rows] = U[:, t]
                Ucols.append(col); Rrows.append(s[t] * Vh[t]); cr.append(q)
        if not Ucols:                                   # the zero operator
            Ucols, Rrows, cr = [np.zeros(M.shape[0], dtype=np.complex128)], [np.zeros(M.shape[1], dtype=np.complex128)], [0]
        out.append(np.stack(Ucols, axis=1).reshape(w, d, d, len(cr)))
        rest = np.stack(Rrows, axis=0).reshape(len(cr), *([d, d] * (n - 1 - k)))
        charges.append(cr)
    out.append(rest.reshape(rest.shape[0], d, d, 1))
    charges.append([0])
    return (out, charges) if return_charges else out


def _level_charges_from_ops(Os, pq):
    """charges of the virtual levels of a chain of local MPO tensors with a trivial first leg: level j to the right of
    tensor k carries c_left + s_out - s_in of any non-zero element feeding it (abelian charge conservation)"""
    pq = [int(c) for c in pq]
    charges = [[0] * Os[0].shape[0]]
    for O in Os:
        cl = charges[-1]
        cr = []
        for j in range(O.shape[3]):
            found = None
            for i in range(O.shape[0]):
                nz = np.argwhere(np.abs(O[i, :, :, j]) > 1e-13)
                for a, b in nz:
                    c = cl[i] + pq[a] - pq[b]
                    if found is None:
                        found = c
                    elif found != c:
                        raise _lib.SpaceMismatch(_lib.ESPACE, "local MPO level mixes charges: decompose per sector or pass level_charges")
            cr.append(0 if found is None else found)
        charges.append(cr)
    return charges


def expectation_value_local(st, O, at=None, level_charges=None):
    """Local expectation values (expval.jl:22-65).
      * O a d x d matrix, or a list of them (one per site), `at` None: <O_i> on every site from AC[i] (expval.jl:22-36);
      * O a list of local MPO tensors [wl, p_out, p_in, wr] and `at` a site: the identity on the bond left of `at` is
        transferred through AL[at : at + n] with the MPO tensors in the middle and closed with CR CR' (expval.jl:45-65);
      * O an n-site operator O[p_out ..., p_in ...] and `at` a site: decomposed into local MPO tensors (expval.jl:41-44)."""
    ctx = st.ctx
    pq = st.pq
    mod = st.pspace.modulus
    if at is None:
        n = _len(st)
        arr = np.asarray(O)
        ops = [arr] * n if arr.ndim == 2 else [np.asarray(o) for o in O]
        out = np.zeros(n, dtype=np.complex128)
        for i in range(n):
            sl = SparseMPOSlice([[np.asarray(ops[i], dtype=np.complex128)]], (0,), (0,), pq, mod)
            ac = _at(st, "AC", i)
            one_l = _identity(ctx, ac.Vl, st.dtype)
            res = _transfer_left(one_l, sl, ac, ac)
            one_r = _identity(ctx, ac.Vr, st.dtype)
            out[i] = one_r.dot(res)
            for t in (one_l, res, one_r):
                t.free()
        return out
    arr = O if isinstance(O, (list, tuple)) else np.asarray(O)
    if not isinstance(arr, (list, tuple)):
        arr, level_charges = decompose_localmpo(arr, pq, return_charges=True)
    Os = [np.asarray(o, dtype=np.complex128) for o in arr]
    if Os[0].shape[0] != 1 or Os[-1].shape[3] != 1:
        raise ValueError("localmpo should start and end in a trivial leg")        # expval.jl:49-51
    if level_charges is None:
        level_charges = _level_charges_from_ops(Os, pq)
    slices, od, level_charges = local_mpo_slices(Os, pq, level_charges, mod)
    A0 = _at(st, "AL", at)
    v = _identity(ctx, A0.Vl, st.dtype, od, list(level_charges[0]) + [0] * (od - len(level_charges[0])))
    for k, sl in enumerate(slices):
        A = _at(st, "AL", at + k)
        nv = _transfer_left(v, sl, A, A)
        v.free()
        v = nv
    Cm = _at(st, "CR", at + len(Os) - 1)
    from .infinite import mul_right
    rho_adj = mul_right(Cm, Cm.adjoint()).adjoint()
    top = v.select([0])
    val = rho_adj.dot(top)
    for t in (v, top, rho_adj):
        t.free()
    return complex(val)


def correlator(st, O1, O2, i, js=None, level_charges=None):
    """correlator(state, O1, O2, i, js) (correlators.jl:10-39): G[k] = <O1[i] O2[js[k]]> for MPO tensors O1[1, p_out, p_in, w],
    O2[w, p_out, p_in, 1]; correlator(state, O12, i, js) for a two-site operator O12[p1_out, p2_out, p1_in, p2_in]
    (correlators.jl:41-45).  `js` may be an int (returns a scalar) or an increasing range."""
    if js is None:                                    # correlator(state, O12, i, js): the arguments shift by one
        O12, i, js = np.asarray(O1), O2, i
        (O1, O2), lcs = decompose_localmpo(O12, st.pq, return_charges=True)
        level_charges = lcs[1]
    scalar = isinstance(js, (int, np.integer))
    js = [int(js)] if scalar else [int(j) for j in js]
    if not js[0] > i:
        raise ValueError(f"i should be smaller than j ({i}, {js[0]})")          # correlators.jl:16
    ctx, pq, mod = st.ctx, st.pq, st.pspace.modulus
    O1 = np.asarray(O1, dtype=np.complex128); O2 = np.asarray(O2, dtype=np.complex128)
    if O1.shape[0] != 1:
        raise ValueError("O1 should start with a trivial leg.")
    if O2.shape[3] != 1:
        raise ValueError("O2 should end with a trivial leg.")
    w = O1.shape[3]
    if level_charges is None:
        level_charges = _level_charges_from_ops([O1], pq)[1]
    lc = [int(c) for c in level_charges]
    (s1,), _, _ = local_mpo_slices([O1], pq, [[0], lc], mod)                   # w x w slice, row 0 filled
    ident = SparseMPOSlice([[1.0 if a == b else 0 for b in range(w)] for a in range(w)], lc, lc, pq, mod)
    (s2,), _, _ = local_mpo_slices([O2], pq, [lc, [0]], mod)                    # w x w slice, column 0 filled
    ac = _at(st, "AC", i)
    v0 = _identity(ctx, ac.Vl, st.dtype, w, [0] * w)
    Vl = _transfer_left(v0, s1, ac, ac)
    v0.free()
    ctr = i + 1
    G = np.zeros(len(js), dtype=np.complex128)
    for k, j in enumerate(js):
        while ctr < j:                                # Vl * TransferMatrix(AR[ctr : j - 1]) (correlators.jl:30-32)
            A = _at(st, "AR", ctr)
            nv = _transfer_left(Vl, ident, A, A)
            Vl.free(); Vl = nv
            ctr += 1
        A = _at(st, "AR", j)
        res = _transfer_left(Vl, s2, A, A)
        top = res.select([0])
        one = _identity(ctx, A.Vr, st.dtype)
        G[k] = one.dot(top)
        for t in (res, top, one):
            t.free()
        ctr = j
    Vl.free()
    return G[0] if scalar else G
