"""Sparse MPO slices with width-1 levels and the model Hamiltonians of the configs.

Mirrors ``SparseMPOSlice`` (src/operators/sparsempo/sparseslice.jl:13-106): ``Os[i, j]`` is
either the number zero (absent), a number (scalar entry c * identity, ``isscal``) or a d x d
operator ``O[p_out, p_in]``; ``keys``, ``opkeys``, ``scalkeys``, ``contains`` and ``isscal`` have
the reference semantics (sparseslice.jl:74-106).  Level charges obey
c_j = c_i + s_{p_out} - s_{p_in} for every non-zero operator element (the TensorMap
O[wl p_out; p_in wr] of docs/src/man/intro.md:77-83 restricted to width-1 levels).
``MPOHamiltonian`` structure (upper triangular, Os[1,1] = Os[end,end] = 1) follows
src/operators/mpohamiltonian.jl:19-31,63-74.
"""
from __future__ import annotations
import numpy as np


class SparseMPOSlice:
    def __init__(self, Os, level_charges_left, level_charges_right, phys_charges):
        self.odim = len(Os)
        self.Os = [[Os[i][j] for j in range(self.odim)] for i in range(self.odim)]
        self.cl = tuple(int(c) for c in level_charges_left)
        self.cr = tuple(int(c) for c in level_charges_right)
        self.pch = tuple(int(c) for c in phys_charges)
        self.d = len(self.pch)
        for (i, j) in self.keys():
            if not self.isscal(i, j):
                O = np.asarray(self.Os[i][j])
                assert O.shape == (self.d, self.d)
                for a in range(self.d):
                    for b in range(self.d):
                        if O[a, b] != 0:
                            assert self.cr[j] == self.cl[i] + self.pch[a] - self.pch[b], \
                                f"entry ({i},{j}) element ({a},{b}) violates charge conservation"
            else:
                assert self.cr[j] == self.cl[i]

    # sparseslice.jl:100-106
    def contains(self, i, j):
        e = self.Os[i][j]
        return not (np.isscalar(e) and e == 0)

    def isscal(self, i, j):
        return np.isscalar(self.Os[i][j]) and self.contains(i, j)

    # sparseslice.jl:74-98 ; product(1:odim,1:odim) iterates the first index fastest
    def keys(self):
        return [(i, j) for j in range(self.odim) for i in range(self.odim) if self.contains(i, j)]

    def keys_col(self, k):
        return [i for i in range(self.odim) if self.contains(i, k)]

    def keys_row(self, j):
        return [k for k in range(self.odim) if self.contains(j, k)]

    def opkeys(self):
        return [(i, j) for (i, j) in self.keys() if not self.isscal(i, j)]

    def scalkeys(self):
        return [(i, j) for (i, j) in self.keys() if self.isscal(i, j)]

    # sparseslice.jl:40-55: getindex materialises scalars as c * (1 (x) 1)
    def op(self, i, j):
        e = self.Os[i][j]
        if np.isscalar(e):
            return e * np.eye(self.d, dtype=np.complex128)
        return np.asarray(e, dtype=np.complex128)

    def dense(self):
        """O[wl, p_out, p_in, wr] (SURVEY.md Appendix A index order)."""
        out = np.zeros((self.odim, self.d, self.d, self.odim), dtype=np.complex128)
        for (i, j) in self.keys():
            out[i, :, :, j] = self.op(i, j)
        return out

    def kinds(self):
        """integer bookkeeping compared bit-exactly with the device side: 0 absent, 1 scalar, 2 operator."""
        return [[0 if not self.contains(i, j) else (1 if self.isscal(i, j) else 2)
                 for j in range(self.odim)] for i in range(self.odim)]


class MPOHamiltonian:
    """finite/periodic list of slices; ``H[i]`` is the slice at site i (0-based here)."""

    def __init__(self, slices):
        self.slices = list(slices)
        self.odim = self.slices[0].odim

    def __len__(self):
        return len(self.slices)

    def __getitem__(self, i):
        return self.slices[i % len(self.slices)]


# ----------------------------------------------------------------------------- local operators
def spin_ops(d=2):
    """S+, S-, Sz for spin (d-1)/2 in the basis m = +S ... -S (charges 2m descending -> we sort ascending)."""
    S = (d - 1) / 2
    ms = np.array([-S + k for k in range(d)])  # ascending m  => ascending charge 2m
    Sz = np.diag(ms)
    Sp = np.zeros((d, d))
    for k in range(d - 1):
        m = ms[k]
        Sp[k + 1, k] = np.sqrt(S * (S + 1) - m * (m + 1))
    Sm = Sp.T.copy()
    charges = tuple(int(round(2 * m)) for m in ms)
    return Sp, Sm, Sz, charges


def fsm_hamiltonian(L, odim, entries, level_charges, phys_charges):
    """build a uniform upper-triangular FSM MPO from {(i,j): scalar|matrix}."""
    def one_slice():
        Os = [[0 for _ in range(odim)] for _ in range(odim)]
        for (i, j), v in entries.items():
            Os[i][j] = v
        return SparseMPOSlice(Os, level_charges, level_charges, phys_charges)
    return MPOHamiltonian([one_slice() for _ in range(L)])


def transverse_field_ising(L, g=1.0):
    """test/setup.jl:38-44: H = -sum ZZ - g sum X (g/2 per bond side -> g per site in the bulk...).

    The reference builds it from the two-site term  -(Z(x)Z + g/2 (X(x)E + E(x)X)) through
    ``decompose_localmpo``; the resulting operator on an infinite chain is
    -sum_i Z_i Z_{i+1} - g sum_i X_i, which is what this FSM encodes (trivial symmetry, chi=3).
    """
    X = np.array([[0, 1], [1, 0]], dtype=np.complex128)
    Z = np.array([[1, 0], [0, -1]], dtype=np.complex128)
    ent = {(0, 0): 1.0, (2, 2): 1.0, (0, 1): -Z, (1, 2): Z, (0, 2): -g * X}
    return fsm_hamiltonian(L, 3, ent, (0, 0, 0), (0, 0))


def heisenberg_u1(L, d=2, J=1.0):
    """nearest-neighbour Heisenberg, U(1) FSM: levels (1, S+, S-, Sz, end), charges (0,+2,-2,0,0)
    (2 scalar + 6 operator entries)."""
    Sp, Sm, Sz, pch = spin_ops(d)
    ent = {(0, 0): 1.0, (4, 4): 1.0,
           (0, 1): Sp, (0, 2): Sm, (0, 3): Sz,
           (1, 4): 0.5 * J * Sm, (2, 4): 0.5 * J * Sp, (3, 4): J * Sz}
    # level after S+ (raises: s_out - s_in = +2) carries +2
    return fsm_hamiltonian(L, 5, ent, (0, 2, -2, 0, 0), pch)


def heisenberg_longrange_u1(L, r=6, d=2, power=2.0):
    """range-r Heisenberg J_k = 1/k^power as a U(1) FSM with chi = 2 + 3r:
    r=6 -> chi=20, 38 non-zeros (17 scalars, 21 operators)."""
    Sp, Sm, Sz, pch = spin_ops(d)
    odim = 2 + 3 * r
    end = odim - 1
    ent = {(0, 0): 1.0, (end, end): 1.0}
    ch = [0] * odim
    ops_first = [Sp, Sm, Sz]
    ops_last = [0.5 * Sm, 0.5 * Sp, Sz]
    qs = [2, -2, 0]
    for t in range(3):
        for k in range(r):
            lev = 1 + t * r + k
            ch[lev] = qs[t]
            if k == 0:
                ent[(0, lev)] = ops_first[t]
            else:
                ent[(lev - 1, lev)] = 1.0
            ent[(lev, end)] = (1.0 / (k + 1) ** power) * ops_last[t]
    return fsm_hamiltonian(L, odim, ent, ch, pch)


def cylinder_bonds(W, Lx, J1=1.0, J2=0.5):
    """{(i, j): J} with i < j for the J1-J2 Heisenberg model on a cylinder of circumference W (periodic) and length Lx
    (open), snake order n = x * W + y (BASELINE.json config 5: W = 6, Lx = 16, J2 = 0.5).  J1: nearest neighbours
    (around the ring and along the cylinder), J2: the two diagonals of every plaquette."""
    bonds = {}

    def add(a, b, J):
        i, j = (a, b) if a < b else (b, a)
        if i != j:
            bonds[(i, j)] = bonds.get((i, j), 0.0) + J
    for x in range(Lx):
        for y in range(W):
            n = x * W + y
            if W > 2 or y == 0:
                add(n, x * W + (y + 1) % W, J1)                       # around the ring
            if x + 1 < Lx:
                add(n, (x + 1) * W + y, J1)                           # along the cylinder
                for dy in ((1, -1) if W > 2 else (1,)):
                    add(n, (x + 1) * W + (y + dy) % W, J2)            # plaquette diagonals
    return bonds


def heisenberg_bonds_u1(L, bonds, d=2):
    """U(1) FSM MPO of sum_{(i,j)} J_ij S_i . S_j for an arbitrary bond list on a chain of L sites: three delay lines
    (S+, S-, Sz) of length R = max |i - j|; level (t, k) = "operator t was applied k + 1 sites ago".  Site-dependent
    slices, chi = 2 + 3 R (snake MPO of a width-6 cylinder: R = 11, chi = 35)."""
    Sp, Sm, Sz, pch = spin_ops(d)
    R = max(j - i for (i, j) in bonds)
    odim = 2 + 3 * R
    end = odim - 1
    first, last, qs = [Sp, Sm, Sz], [0.5 * Sm, 0.5 * Sp, Sz], [2, -2, 0]
    ch = [0] * odim
    for t in range(3):
        for k in range(R):
            ch[1 + t * R + k] = qs[t]
    # a delay-line level is only needed while some bond still waits for it: (t, k) at the left of site n exists iff a
    # bond (n - k - 1, j) with j >= n is present
    def pending(n, k):
        i = n - k - 1
        return i >= 0 and any((i, j) in bonds for j in range(n, min(L, i + R + 1)))
    slices = []
    for n in range(L):
        Os = [[0 for _ in range(odim)] for _ in range(odim)]
        Os[0][0] = 1.0; Os[end][end] = 1.0
        for t in range(3):
            if any((n, j) in bonds for j in range(n + 1, min(L, n + R + 1))):
                Os[0][1 + t * R] = first[t]
            for k in range(R):
                lev = 1 + t * R + k
                if not pending(n, k):
                    continue
                J = bonds.get((n - k - 1, n), 0.0)
                if J != 0.0:
                    Os[lev][end] = J * last[t]
                if k + 1 < R and pending(n + 1, k + 1):
                    Os[lev][lev + 1] = 1.0
        slices.append(SparseMPOSlice(Os, ch, ch, pch))
    return MPOHamiltonian(slices)


def j1j2_cylinder_u1(W=6, Lx=16, J1=1.0, J2=0.5):
    """BASELINE.json config 5: J1-J2 Heisenberg on a width-W cylinder as a snake MPO (U(1), chi = 2 + 3 (2 W - 1))."""
    return heisenberg_bonds_u1(W * Lx, cylinder_bonds(W, Lx, J1, J2))


def bonds_dense_hamiltonian(L, bonds, d=2):
    """sum_{(i,j)} J_ij S_i . S_j built directly from Kronecker products (independent of the MPO construction)"""
    Sp, Sm, Sz, _ = spin_ops(d)
    def site(op, i):
        out = np.eye(1)
        for k in range(L):
            out = np.kron(out, op if k == i else np.eye(d))
        return out
    H = np.zeros((d ** L, d ** L))
    for (i, j), J in bonds.items():
        H += J * (0.5 * (site(Sp, i) @ site(Sm, j) + site(Sm, i) @ site(Sp, j)) + site(Sz, i) @ site(Sz, j))
    return H


def mpo_product(B, A):
    """B * A for two sparse MPOs with width-1 levels (src/operators/sparsempo/sparsempo.jl:232-264): new level
    (i, k) -> i + k * odim_A (first index fastest, as LinearIndices), entry [(i,k),(j,l)] = A[i,j] then B[k,l]
    (operator product B[k,l] @ A[i,j]; two scalars multiply), level charges add.  Used by `variance` (H * H)."""
    out = []
    for sa, sb in zip(A.slices, B.slices):
        oa, ob = sa.odim, sb.odim
        n = oa * ob
        Os = [[0 for _ in range(n)] for _ in range(n)]
        for (i, j) in sa.keys():
            for (k, l) in sb.keys():
                if sa.isscal(i, j) and sb.isscal(k, l):
                    v = sa.Os[i][j] * sb.Os[k][l]
                else:
                    ea = sa.Os[i][j] * np.eye(sa.d) if sa.isscal(i, j) else np.asarray(sa.Os[i][j], dtype=np.complex128)
                    eb = sb.Os[k][l] * np.eye(sb.d) if sb.isscal(k, l) else np.asarray(sb.Os[k][l], dtype=np.complex128)
                    v = eb @ ea
                    if np.abs(v).max() == 0:
                        continue
                Os[i + k * oa][j + l * oa] = v
        cl = [sa.cl[i] + sb.cl[k] for k in range(ob) for i in range(oa)]
        cr = [sa.cr[i] + sb.cr[k] for k in range(ob) for i in range(oa)]
        out.append(SparseMPOSlice(Os, cl, cr, sa.pch))
    return MPOHamiltonian(out)


def heisenberg_dense(L, d=2, J=1.0):
    """same operator without symmetry (all charges zero): the dense/trivial path."""
    H = heisenberg_u1(L, d, J)
    sl = []
    for s in H.slices:
        sl.append(SparseMPOSlice(s.Os, (0,) * s.odim, (0,) * s.odim, (0,) * s.d))
    return MPOHamiltonian(sl)


def dense_hamiltonian(H: MPOHamiltonian):
    """full 2^L matrix of a finite MPOHamiltonian with the FinEnv boundary vectors
    (level 1 on the left, level odim on the right; FinEnv.jl:41-70). Exact-diagonalisation anchor."""
    L = len(H)
    cur = None
    for i in range(L):
        W = H[i].dense()  # wl, po, pi, wr
        if cur is None:
            cur = W[0]  # po, pi, wr
        else:
            cur = np.einsum('abw,wcdv->acbdv', cur, W)
            s = cur.shape
            cur = cur.reshape(s[0] * s[1], s[2] * s[3], s[4])
    return cur[:, :, -1]


# ----------------------------------------------------------------------------- dense MPOs (classic2d)
def dense_mpo_slice(O, level_charges=None, phys_charges=None):
    """flatten a DenseMPO tensor O[wl, p_out, p_in, wr] (src/operators/densempo.jl) into width-1 levels: a chi x chi
    SparseMPOSlice whose entries are the d x d operators O[i, :, :, j] (entries that vanish identically are absent)."""
    chi, d = O.shape[0], O.shape[1]
    Os = [[0 for _ in range(chi)] for _ in range(chi)]
    for i in range(chi):
        for j in range(chi):
            blk = np.asarray(O[i, :, :, j], dtype=np.complex128)
            if np.abs(blk).max() > 0:
                Os[i][j] = blk
    lc = level_charges if level_charges is not None else (0,) * chi
    return SparseMPOSlice(Os, lc, lc, phys_charges if phys_charges is not None else (0,) * d)


def classical_ising(beta=None):
    """test/setup.jl:78-94: the 2D classical Ising transfer MPO at beta_c = log(1 + sqrt 2) / 2, O[wl, p_out; p_in, wr]
    with the reference's leg order (all four legs of dimension 2)."""
    beta = np.log(1 + np.sqrt(2)) / 2 if beta is None else beta
    t = np.array([[np.exp(beta), np.exp(-beta)], [np.exp(-beta), np.exp(beta)]])
    w, v = np.linalg.eigh(t)
    nt = v @ np.diag(np.sqrt(w)) @ v.T          # test/setup.jl:78-83 uses r.vectors * sqrt(D) * r.vectors (symmetric square root)
    delta = np.zeros((2, 2, 2, 2))
    delta[0, 0, 0, 0] = delta[1, 1, 1, 1] = 1
    o = np.einsum('abcd,ia,jb,kc,ld->ijkl', delta, nt, nt, nt, nt)
    return o      # legs (1,2;3,4) of the TensorMap: here read as O[wl, p_out, p_in, wr]


def periodic_boundary_conditions(H: MPOHamiltonian, L=None):
    """src/algorithms/toolbox.jl:186-306 for width-1 levels: every fuser is the 1 x 1 identity and
    every braid trivial, so the construction is a relabelled finite-state machine over triples (a, b, c) = (progress in
    the upper copy, level lent across the boundary, progress in the lower copy), charge c_a - c_b + c_c."""
    L = len(H) if L is None else L
    period = len(H)
    chi = H.odim
    top = chi - 1
    indmap = {}
    for b in range(top, 0, -1):
        for c in range(b, chi):
            indmap[(0, b, c)] = len(indmap)
    for a in range(1, chi):
        for b in range(top, a - 1, -1):
            indmap[(a, b, top)] = len(indmap)
    chi2 = len(indmap)
    cl = H[0].cl
    charges = [0] * chi2
    for (a, b, c), idx in indmap.items():
        charges[idx] = cl[a] - cl[b] + cl[c]

    def empty():
        return [[0 for _ in range(chi2)] for _ in range(chi2)]

    def bulk(loc):
        Os = empty()
        sl = H[loc]
        for (j, k) in sl.keys():
            for i in range(1, chi):                      # apply (j, k) above
                if k <= i:
                    Os[indmap[(j, i, top)]][indmap[(k, i, top)]] = sl.Os[j][k]
            for l in range(1, top):                      # apply (j, k) below
                if l <= j:
                    Os[indmap[(0, l, j)]][indmap[(0, l, k)]] = sl.Os[j][k]
        return Os
    starter = empty()
    for (j, k) in H[0].keys():
        if j == 0:
            starter[0][indmap[(k, top, top)]] = H[0].Os[j][k]
        if 0 < j < top:
            starter[0][indmap[(0, j, k)]] = H[0].Os[j][k]
    starter[0][0] = 1.0
    starter[chi2 - 1][chi2 - 1] = 1.0
    ender = empty()
    last = H[period - 1]
    for (j, k) in last.keys():
        if k > 0:
            ender[indmap[(j, k, top)]][chi2 - 1] = last.Os[j][k]
    ender[0][0] = 1.0
    ender[chi2 - 1][chi2 - 1] = 1.0
    pch = H[0].pch
    slices = [SparseMPOSlice(starter, charges, charges, pch)]
    for i in range(1, L - 1):
        slices.append(SparseMPOSlice(bulk(i % period), charges, charges, pch))
    slices.append(SparseMPOSlice(ender, charges, charges, pch))
    return MPOHamiltonian(slices)
