"""Host mirror of the operator containers the hot path reads (construction itself is host-side model
building and out of scope): ``SparseMPOSlice`` (src/operators/sparsempo/sparseslice.jl:13-106) with
width-1 levels, ``MPOHamiltonian`` (src/operators/mpohamiltonian.jl:9-31), and the model Hamiltonians
of the configs in BASELINE.json.  Each slice owns the device-side ``mpsk_mposlice`` handle."""
from __future__ import annotations
import ctypes as C
import numpy as np
from . import _lib


class SparseMPOSlice:
    def __init__(self, Os, level_charges_left, level_charges_right, phys_charges, modulus=0):
        self.modulus = int(modulus)                                    # 0: U(1) / trivial charges, N: Z_N
        self.odim = len(Os)
        self.Os = [[Os[i][j] for j in range(self.odim)] for i in range(self.odim)]
        self.domcharges = tuple(int(c) for c in level_charges_left)    # domspaces (left virtual levels)
        self.imcharges = tuple(int(c) for c in level_charges_right)    # imspaces  (right virtual levels)
        self.pcharges = tuple(int(c) for c in phys_charges)            # pspace
        self.d = len(self.pcharges)
        self._dev = {}

    # sparseslice.jl:100-106
    def contains(self, i, j):
        e = self.Os[i][j]
        return not (np.isscalar(e) and e == 0)

    def isscal(self, i, j):
        return np.isscalar(self.Os[i][j]) and self.contains(i, j)

    # sparseslice.jl:74-98 (product iterates the first index fastest)
    def keys(self):
        return [(i, j) for j in range(self.odim) for i in range(self.odim) if self.contains(i, j)]

    def opkeys(self):
        return [k for k in self.keys() if not self.isscal(*k)]

    def scalkeys(self):
        return [k for k in self.keys() if self.isscal(*k)]

    def kinds(self):
        return [[0 if not self.contains(i, j) else (1 if self.isscal(i, j) else 2) for j in range(self.odim)]
                for i in range(self.odim)]

    def device(self, ctx):
        """mpsk_mposlice handle (created once per context)."""
        key = id(ctx)
        if key not in self._dev:
            od, d = self.odim, self.d
            kind = (C.c_int32 * (od * od))()
            scal = (C.c_double * (2 * od * od))()
            ops = []
            for i in range(od):
                for j in range(od):
                    k = self.kinds()[i][j]
                    kind[i * od + j] = k
                    if k == 1:
                        v = complex(self.Os[i][j])
                        scal[2 * (i * od + j)], scal[2 * (i * od + j) + 1] = v.real, v.imag
                    elif k == 2:
                        O = np.asarray(self.Os[i][j], dtype=np.complex128)
                        assert O.shape == (d, d)
                        ops.append(O.reshape(-1))
            flat = np.concatenate(ops).view(np.float64) if ops else np.zeros(2)
            flat = np.ascontiguousarray(flat)
            cl = (C.c_int32 * od)(*self.domcharges); cr = (C.c_int32 * od)(*self.imcharges)
            pc = (C.c_int32 * d)(*self.pcharges)
            h = C.c_void_p()
            ctx.check(ctx.lib.mpsk_mposlice_create_zn(ctx.h, od, cl, cr, d, pc, kind, scal,
                                                      flat.ctypes.data_as(C.POINTER(C.c_double)), self.modulus, C.byref(h)))
            self._dev[key] = h
        return self._dev[key]

    def device_keys(self, ctx):
        """keys(x) as computed by the library (bit-exact bookkeeping check)."""
        h = self.device(ctx)
        n = C.c_int32()
        ctx.lib.mpsk_mposlice_keys(h, C.byref(n), None)
        ij = (C.c_int32 * (2 * max(n.value, 1)))()
        ctx.lib.mpsk_mposlice_keys(h, C.byref(n), ij)
        return [(ij[2 * k], ij[2 * k + 1]) for k in range(n.value)]


class MPOHamiltonian:
    """``H[i]`` is the slice at site i (0-based, periodic like the reference's PeriodicArray)."""

    def __init__(self, slices):
        self.slices = list(slices)
        self.odim = self.slices[0].odim

    def __len__(self):
        return len(self.slices)

    def __getitem__(self, i):
        return self.slices[i % len(self.slices)]

    # ---- arithmetic (src/operators/mpohamiltonian.jl:76-157); host-side operator construction -------------------------
    def _map_slices(self, f):
        return MPOHamiltonian([f(c, sl) for c, sl in enumerate(self.slices)])

    def __neg__(self):                                      # mpohamiltonian.jl:77
        return self * (-1.0)

    def __mul__(self, b):
        """a * b::Number scales the entries that end in the last level (mpohamiltonian.jl:147-154); b * a::MPOHamiltonian is
        the operator product (mpohamiltonian.jl:156, sparsempo.jl:232-264)"""
        if isinstance(b, MPOHamiltonian):
            return mpo_product(self, b)
        if hasattr(b, "AL"):                                # H * mps::InfiniteMPS = convert(DenseMPO, H) * mps (:164)
            if hasattr(b, "L"):
                raise NotImplementedError("mpo * FiniteMPS (densempo.jl:46-71) is not built: "
                                          "approximate(psi, (H, above), alg) applies an operator to a FiniteMPS")
            return DenseMPO(self.slices) * b
        if not isinstance(b, (int, float, complex, np.number)):
            return NotImplemented                           # e.g. f(t) * H -> MultipliedOperator (lazyops.py)

        def scaled(c, sl):
            Os = [row[:] for row in sl.Os]
            for j in range(sl.odim - 1):
                if sl.contains(j, sl.odim - 1):
                    Os[j][sl.odim - 1] = Os[j][sl.odim - 1] * b
            return SparseMPOSlice(Os, sl.domcharges, sl.imcharges, sl.pcharges, sl.modulus)
        return self._map_slices(scaled)

    def __rmul__(self, b):                                  # mpohamiltonian.jl:146
        return self * b

    def __add__(self, other):
        """a + e::Vector adds e[c] * 1 per site to the entry (first, last) (mpohamiltonian.jl:78-90); a + b merges the two
        automata: shared first and last level, b's inner levels appended after a's (mpohamiltonian.jl:97-142)"""
        if not isinstance(other, MPOHamiltonian):
            if not isinstance(other, (list, tuple, np.ndarray, int, float, complex, np.number)):
                return NotImplemented                       # H + LazySum / MultipliedOperator: the lazy algebra (lazyops.py)
            e = [complex(x) for x in np.atleast_1d(other)]
            if len(e) != len(self):
                raise ValueError(f"periodicity should match {len(self)} != {len(e)}")

            def shifted(c, sl):
                Os = [row[:] for row in sl.Os]
                Os[0][sl.odim - 1] = _entry_add(Os[0][sl.odim - 1], e[c], sl.d)
                return SparseMPOSlice(Os, sl.domcharges, sl.imcharges, sl.pcharges, sl.modulus)
            return self._map_slices(shifted)
        if len(self) != len(other):
            raise ValueError(f"periodicity should match {len(self)} != {len(other)}")
        out = []
        for sa, sb in zip(self.slices, other.slices):
            oa, ob = sa.odim, sb.odim
            n = oa + ob - 2
            Os = [[0 for _ in range(n)] for _ in range(n)]
            for (i, j) in sa.keys():
                if i < oa - 1 and j < oa - 1:
                    Os[i][j] = sa.Os[i][j]
                elif i < oa - 1 and j == oa - 1:
                    Os[i][n - 1] = sa.Os[i][j]
            for (i, j) in sb.keys():
                if i == 0 and j > 0:
                    Os[0][oa + j - 2] = _entry_add(Os[0][oa + j - 2], sb.Os[i][j], sb.d)
                elif i > 0 and j > 0:
                    Os[oa + i - 2][oa + j - 2] = sb.Os[i][j]
            cl = list(sa.domcharges[:oa - 1]) + list(sb.domcharges[1:])
            cr = list(sa.imcharges[:oa - 1]) + list(sb.imcharges[1:])
            out.append(SparseMPOSlice(Os, cl, cr, sa.pcharges, sa.modulus))
        return MPOHamiltonian(out)

    def __radd__(self, e):                                  # mpohamiltonian.jl:93
        return self + e

    def __sub__(self, other):                               # mpohamiltonian.jl:94,143
        if isinstance(other, MPOHamiltonian):
            return self + (-other)
        if not isinstance(other, (list, tuple, np.ndarray, int, float, complex, np.number)):
            return NotImplemented
        return self + [-complex(x) for x in np.atleast_1d(other)]

    def __rsub__(self, e):                                  # mpohamiltonian.jl:92
        return self * (-1.0) + e

    def repeat(self, n):                                    # mpohamiltonian.jl:157
        return MPOHamiltonian(self.slices * int(n))

    @staticmethod
    def from_local(O, phys_charges=None, tol=1e-14):
        """MPOHamiltonian(t::TensorMap) = MPOHamiltonian(decompose_localmpo(add_util_leg(t))) (mpohamiltonian.jl:16,19-31): the
        n-site interaction O[p1_out .. pn_out, p1_in .. pn_in] on every n consecutive sites of a one-site unit cell; every index
        of the k-th virtual bond of the decomposition becomes one level (width 1, with its charge) of the automaton."""
        from .toolbox import decompose_localmpo
        O = np.asarray(O, dtype=np.complex128)
        d = O.shape[0]
        pq = [0] * d if phys_charges is None else [int(c) for c in phys_charges]
        Ts, charges = decompose_localmpo(O, pq, tol, return_charges=True)
        scale = np.abs(O).max() or 1.0
        widths = [T.shape[3] for T in Ts[:-1]]
        offs = [1 + int(sum(widths[:k])) for k in range(len(widths) + 1)]          # first level of bond k (bond 0: level 0)
        od = 2 + int(sum(widths))
        lc = [0] * od
        for k, w in enumerate(widths):
            for a in range(w):
                lc[offs[k] + a] = int(charges[k + 1][a])
        Os = [[0 for _ in range(od)] for _ in range(od)]
        Os[0][0] = 1.0
        Os[od - 1][od - 1] = 1.0
        for k, T in enumerate(Ts):
            for a in range(T.shape[0]):
                for b in range(T.shape[3]):
                    row = 0 if k == 0 else offs[k - 1] + a
                    col = od - 1 if k == len(Ts) - 1 else offs[k] + b
                    allowed = np.array([[lc[col] == lc[row] + pq[x] - pq[y] for y in range(d)] for x in range(d)])
                    blk = T[a, :, :, b]
                    if np.abs(blk * ~allowed).max() > 1e-10 * scale:
                        raise ValueError("from_local: the operator does not conserve the physical charges")
                    blk = blk * allowed                         # rounding noise of the SVDs outside the allowed elements
                    if np.abs(blk).max() == 0:
                        continue
                    Os[row][col] = _entry_add(Os[row][col], blk, d)
        return MPOHamiltonian([SparseMPOSlice(Os, lc, lc, pq)])


def _entry_add(a, b, d):
    """sum of two slice entries; scalars stand for multiples of the identity (sparseslice.jl:40-55)"""
    if np.isscalar(a) and np.isscalar(b):
        return a + b
    ea = a * np.eye(d) if np.isscalar(a) else np.asarray(a, dtype=np.complex128)
    eb = b * np.eye(d) if np.isscalar(b) else np.asarray(b, dtype=np.complex128)
    return ea + eb


# --------------------------------------------------------------------------------------- models
def spin_ops(d=2):
    S = (d - 1) / 2
    ms = np.array([-S + k for k in range(d)])
    Sz = np.diag(ms)
    Sp = np.zeros((d, d))
    for k in range(d - 1):
        Sp[k + 1, k] = np.sqrt(S * (S + 1) - ms[k] * (ms[k] + 1))
    return Sp, Sp.T.copy(), Sz, tuple(int(round(2 * m)) for m in ms)


def _fsm(L, odim, entries, level_charges, phys_charges):
    out = []
    for _ in range(L):
        Os = [[0 for _ in range(odim)] for _ in range(odim)]
        for (i, j), v in entries.items():
            Os[i][j] = v
        out.append(SparseMPOSlice(Os, level_charges, level_charges, phys_charges))
    return MPOHamiltonian(out)


def transverse_field_ising(L, g=1.0):
    """-sum ZZ - g sum X (test/setup.jl:38-44), trivial symmetry, odim 3."""
    X = np.array([[0, 1], [1, 0]], dtype=np.complex128)
    Z = np.array([[1, 0], [0, -1]], dtype=np.complex128)
    return _fsm(L, 3, {(0, 0): 1.0, (2, 2): 1.0, (0, 1): -Z, (1, 2): Z, (0, 2): -g * X}, (0, 0, 0), (0, 0))


def transverse_field_ising_z2(L, g=1.0):
    """the same operator in the X eigenbasis with its Z2 parity symmetry: -sum X X - g sum Z (|+>: charge 0, |->: charge 1);
    the coupling operator flips the parity, so its MPO level carries charge 1 (mod 2)."""
    X = np.array([[0, 1], [1, 0]], dtype=np.complex128)
    Z = np.array([[1, 0], [0, -1]], dtype=np.complex128)
    out = []
    for _ in range(L):
        Os = [[0] * 3 for _ in range(3)]
        for (i, j), v in {(0, 0): 1.0, (2, 2): 1.0, (0, 1): -X, (1, 2): X, (0, 2): -g * Z}.items():
            Os[i][j] = v
        out.append(SparseMPOSlice(Os, (0, 1, 0), (0, 1, 0), (0, 1), modulus=2))
    return MPOHamiltonian(out)


def heisenberg_u1(L, d=2, J=1.0):
    Sp, Sm, Sz, pch = spin_ops(d)
    ent = {(0, 0): 1.0, (4, 4): 1.0, (0, 1): Sp, (0, 2): Sm, (0, 3): Sz,
           (1, 4): 0.5 * J * Sm, (2, 4): 0.5 * J * Sp, (3, 4): J * Sz}
    return _fsm(L, 5, ent, (0, 2, -2, 0, 0), pch)


def heisenberg_longrange_u1(L, r=6, d=2, power=2.0):
    """range-r Heisenberg J_k = k^-power; odim = 2 + 3r (r=6: 38 non-zeros = 17 scalars + 21 operators)."""
    Sp, Sm, Sz, pch = spin_ops(d)
    odim = 2 + 3 * r
    end = odim - 1
    ent = {(0, 0): 1.0, (end, end): 1.0}
    ch = [0] * odim
    first, last, qs = [Sp, Sm, Sz], [0.5 * Sm, 0.5 * Sp, Sz], [2, -2, 0]
    for t in range(3):
        for k in range(r):
            lev = 1 + t * r + k
            ch[lev] = qs[t]
            if k == 0:
                ent[(0, lev)] = first[t]
            else:
                ent[(lev - 1, lev)] = 1.0
            ent[(lev, end)] = (1.0 / (k + 1) ** power) * last[t]
    return _fsm(L, odim, ent, ch, pch)


def cylinder_bonds(W, Lx, J1=1.0, J2=0.5):
    """bond list {(i, j): J}, i < j, of the J1-J2 Heisenberg model on a width-W cylinder (periodic around, open along),
    snake order n = x W + y (BASELINE.json config 5)."""
    bonds = {}
    for x in range(Lx):
        for y in range(W):
            n = x * W + y
            nbrs = []
            if W > 2 or y == 0:
                nbrs.append((x * W + (y + 1) % W, J1))
            if x + 1 < Lx:
                nbrs.append(((x + 1) * W + y, J1))
                nbrs += [((x + 1) * W + (y + dy) % W, J2) for dy in ((1, -1) if W > 2 else (1,))]
            for o, J in nbrs:
                if o != n:
                    key = (min(n, o), max(n, o))
                    bonds[key] = bonds.get(key, 0.0) + J
    return bonds


def heisenberg_bonds_u1(L, bonds, d=2):
    """sum_{(i,j)} J_ij S_i.S_j as a U(1) finite-state-machine MPO with three delay lines of length R = max |i - j|
    (level 1 + t R + k: operator t in {S+, S-, Sz} applied k + 1 sites ago); levels exist only while a bond waits for them.
    Host-side model construction (the reference builds such MPOs in MPSKitModels)."""
    Sp, Sm, Sz, pch = spin_ops(d)
    R = max(j - i for (i, j) in bonds)
    odim, first, last, qs = 2 + 3 * R, [Sp, Sm, Sz], [0.5 * Sm, 0.5 * Sp, Sz], [2, -2, 0]
    end = odim - 1
    ch = [0] + [qs[t] for t in range(3) for _ in range(R)] + [0]
    starts = {i for (i, j) in bonds}
    waiting = [set() for _ in range(L + 1)]            # waiting[n]: sites i < n whose operator is still needed at some j >= n
    for (i, j) in bonds:
        for n in range(i + 1, j + 1):
            waiting[n].add(i)
    out = []
    for n in range(L):
        Os = [[0] * odim for _ in range(odim)]
        Os[0][0] = 1.0; Os[end][end] = 1.0
        for t in range(3):
            if n in starts:
                Os[0][1 + t * R] = first[t]
            for i in waiting[n]:
                k = n - i - 1
                lev = 1 + t * R + k
                J = bonds.get((i, n), 0.0)
                if J != 0.0:
                    Os[lev][end] = J * last[t]
                if i in waiting[n + 1]:
                    Os[lev][lev + 1] = 1.0
        out.append(SparseMPOSlice(Os, ch, ch, pch))
    return MPOHamiltonian(out)


def j1j2_cylinder_u1(W=6, Lx=16, J1=1.0, J2=0.5):
    """config 5: J1-J2 Heisenberg on a width-W cylinder of length Lx, snake MPO with chi = 2 + 3 (2 W - 1) (35 for W = 6)."""
    return heisenberg_bonds_u1(W * Lx, cylinder_bonds(W, Lx, J1, J2))


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
        cl = [sa.domcharges[i] + sb.domcharges[k] for k in range(ob) for i in range(oa)]
        cr = [sa.imcharges[i] + sb.imcharges[k] for k in range(ob) for i in range(oa)]
        out.append(SparseMPOSlice(Os, cl, cr, sa.pcharges, sa.modulus))
    return MPOHamiltonian(out)


def drop_symmetry(H: MPOHamiltonian):
    """same operator with all charges set to zero (trivial-symmetry / dense path)."""
    return MPOHamiltonian([SparseMPOSlice(s.Os, (0,) * s.odim, (0,) * s.odim, (0,) * s.d) for s in H.slices])


# --------------------------------------------------------------------------------------- dense MPOs (classic2d)
def dense_mpo_slice(O, level_charges=None, phys_charges=None):
    """DenseMPO tensor O[wl, p_out, p_in, wr] flattened into width-1 levels (chi x chi slice of d x d operator entries)."""
    O = np.asarray(O)
    chi, d = O.shape[0], O.shape[1]
    Os = [[0 for _ in range(chi)] for _ in range(chi)]
    for i in range(chi):
        for j in range(chi):
            blk = np.asarray(O[i, :, :, j], dtype=np.complex128)
            if np.abs(blk).max() > 0:
                Os[i][j] = blk
    lc = tuple(level_charges) if level_charges is not None else (0,) * chi
    return SparseMPOSlice(Os, lc, lc, tuple(phys_charges) if phys_charges is not None else (0,) * d)


def classical_ising(beta=None):
    """test/setup.jl:78-94: 2D classical Ising transfer MPO at beta_c (default), O[wl, p_out, p_in, wr]."""
    beta = np.log(1 + np.sqrt(2)) / 2 if beta is None else beta
    t = np.array([[np.exp(beta), np.exp(-beta)], [np.exp(-beta), np.exp(beta)]])
    w, v = np.linalg.eigh(t)
    nt = v @ np.diag(np.sqrt(w)) @ v.T
    delta = np.zeros((2, 2, 2, 2))
    delta[0, 0, 0, 0] = delta[1, 1, 1, 1] = 1
    return np.einsum('abcd,ia,jb,kc,ld->ijkl', delta, nt, nt, nt, nt)


class DenseMPO:
    """src/operators/densempo.jl: periodic list of dense MPO tensors, stored as flattened slices."""

    def __init__(self, tensors, level_charges=None, phys_charges=None):
        self.slices = [t if isinstance(t, SparseMPOSlice) else dense_mpo_slice(t, level_charges, phys_charges) for t in tensors]
        self.odim = self.slices[0].odim

    def __len__(self):
        return len(self.slices)

    def __getitem__(self, i):
        return self.slices[i % len(self.slices)]

    def __mul__(self, other):
        """mpo * st::InfiniteMPS (densempo.jl:31-45); mpo1 * mpo2 (densempo.jl:73-87: mpo2 acts first)"""
        if isinstance(other, DenseMPO):
            if len(self) != len(other):
                raise ValueError("dimension mismatch")
            return DenseMPO(mpo_product(self, other).slices)
        if hasattr(other, "L"):
            raise NotImplementedError("mpo * FiniteMPS (densempo.jl:46-71) is not built: "
                                      "approximate(psi, (O, above), alg) applies an operator to a FiniteMPS")
        from .changebonds import mpo_times_infinite
        return mpo_times_infinite(self, other)


def periodic_boundary_conditions(H: MPOHamiltonian, L=None):
    """periodic_boundary_conditions(H, L) (src/algorithms/toolbox.jl:186-306) for width-1 levels: relabelled FSM over
    triples (a, b, c) with charge c_a - c_b + c_c (all fusers are 1 x 1 identities); host-side operator construction."""
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
    cl = H[0].domcharges
    charges = [0] * chi2
    for (a, b, c), idx in indmap.items():
        charges[idx] = cl[a] - cl[b] + cl[c]

    def empty():
        return [[0 for _ in range(chi2)] for _ in range(chi2)]

    def bulk(loc):
        Os = empty()
        sl = H[loc]
        for (j, k) in sl.keys():
            for i in range(1, chi):
                if k <= i:
                    Os[indmap[(j, i, top)]][indmap[(k, i, top)]] = sl.Os[j][k]
            for l in range(1, top):
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
    pch = H[0].pcharges
    slices = [SparseMPOSlice(starter, charges, charges, pch)]
    for i in range(1, L - 1):
        slices.append(SparseMPOSlice(bulk(i % period), charges, charges, pch))
    slices.append(SparseMPOSlice(ender, charges, charges, pch))
    return MPOHamiltonian(slices)


# --------------------------------------------------------------------------------------- time-evolution MPO (W^II)
def make_time_mpo(H: MPOHamiltonian, dt, alg="WII"):
    """make_time_mpo(H::MPOHamiltonian, dt, WII()) (src/algorithms/timestep/timeevmpo.jl:149-207; Zaletel et al.,
    arXiv:1407.1832): the second-order-accurate MPO approximation of exp(-i dt H) for an FSM Hamiltonian
    H = [[1, C, D], [0, A, B], [0, 0, 1]] with width-1 levels.  For every pair (j, k) of inner levels the reference exponentiates
    the block-triangular generator
        F = [[dD, 0, 0, 0], [sqrt(d) C_k, dD, 0, 0], [sqrt(d) B_j, 0, dD, 0], [A_jk, sqrt(d) B_j, sqrt(d) C_k, dD]],  d = -i dt,
    applied to (1, 0, 0, 0) (there with KrylovKit's exponentiate; here the 4d x 4d matrix exponential, the same numbers): the
    first block column of exp(F) gives W_D, W_C[k], W_B[j], W_A[j, k].  The result has odim - 1 levels -- the first and last
    level of H merge into level 0 -- and is used as a (periodic) list of slices, e.g. by approximate(psi, (W, psi), VOMPS()).
    Host-side model building on d x d numbers per entry; nothing here touches the device."""
    if alg == "WI":
        alg = ("TaylorCluster", 1)                                     # const WI = TaylorCluster{1} (timeevmpo.jl:10)
    if isinstance(alg, tuple) and alg[0] == "TaylorCluster":
        return _taylor_cluster(H, dt, int(alg[1]))
    if alg != "WII":
        raise ValueError("make_time_mpo: alg must be 'WII', 'WI' or ('TaylorCluster', N)")
    from scipy.linalg import expm
    delta = -1j * dt
    sq = np.sqrt(delta + 0j)
    out = []
    for sl in H.slices:
        od, d = sl.odim, sl.d
        eye = np.eye(d, dtype=np.complex128)

        def op(i, j):
            e = sl.Os[i][j]
            if np.isscalar(e):
                return complex(e) * eye
            return np.asarray(e, dtype=np.complex128)
        n = od - 1
        D = op(0, od - 1)
        W = [[0 for _ in range(n)] for _ in range(n)]
        W[0][0] = expm(delta * D)
        for j in range(1, od - 1):
            for k in range(1, od - 1):
                A, B, Cm = op(j, k), op(j, od - 1), op(0, k)
                F = np.zeros((4 * d, 4 * d), dtype=np.complex128)
                for b in range(4):
                    F[b * d:(b + 1) * d, b * d:(b + 1) * d] = delta * D
                F[d:2 * d, 0:d] = sq * Cm
                F[2 * d:3 * d, 0:d] = sq * B
                F[3 * d:4 * d, 0:d] = A
                F[3 * d:4 * d, d:2 * d] = sq * B
                F[3 * d:4 * d, 2 * d:3 * d] = sq * Cm
                y = expm(F)[:, 0:d]
                W[j][k] = y[3 * d:4 * d]
                W[j][0] = y[2 * d:3 * d]
                W[0][k] = y[d:2 * d]
        cl = (0,) + tuple(sl.domcharges[1:od - 1])
        cr = (0,) + tuple(sl.imcharges[1:od - 1])
        pch = sl.pcharges
        for j in range(n):
            for k in range(n):
                if np.isscalar(W[j][k]):
                    continue
                # the matrix exponential leaves rounding noise on elements that charge conservation forbids: they must vanish
                for a in range(d):
                    for b in range(d):
                        viol = cr[k] - cl[j] - pch[a] + pch[b]
                        if (viol % sl.modulus if sl.modulus else viol) != 0:
                            if abs(W[j][k][a, b]) > 1e-12:
                                raise _lib.SpaceMismatch(_lib.ESPACE, f"time MPO entry ({j},{k}) violates charge conservation")
                            W[j][k][a, b] = 0.0
                if np.abs(W[j][k]).max() == 0.0:
                    W[j][k] = 0
        out.append(SparseMPOSlice(W, cl, cr, sl.pcharges, sl.modulus))
    return MPOHamiltonian(out)


def _taylor_cluster(H: MPOHamiltonian, dt, N):
    """make_time_mpo(H, dt, TaylorCluster{N}()) (src/algorithms/timestep/timeevmpo.jl:12-108; Van Damme et al.,
    arXiv:2302.14181): the N-fold operator product MPO H * ... * H (levels = N-tuples of levels of H, first index fastest) with
    (i) the next order embedded into the entries that continue, (ii) the tuples that have finished (only first / last levels)
    looped back to level 1 with the Taylor weights tau^order (N - order)! / N!, (iii) permutation-equivalent rows / columns
    merged, (iv) the approximate compression of the partially finished tuples, (v) orphan levels removed.  exp(-i dt H) is
    reproduced to O(dt^(N+1)) per site.  Host-side model building on d x d matrices."""
    import itertools
    from math import factorial
    tau = -1j * dt
    out_slices = []
    for sl in H.slices:
        od, d = sl.odim, sl.d
        eye = np.eye(d, dtype=np.complex128)

        def hop(i, j):
            e = sl.Os[i][j]
            if np.isscalar(e):
                return None if e == 0 else complex(e) * eye
            return np.asarray(e, dtype=np.complex128)
        base = [[hop(i, j) for j in range(od)] for i in range(od)]

        def prod_elem(ta, tb):
            """calc_prod_elem: H[a_n, b_n] ... H[a_2, b_2] H[a_1, b_1] (the first tuple entry acts first); None if absent"""
            o = None
            for a, b in zip(ta, tb):
                e = base[a][b]
                if e is None:
                    return None
                o = e if o is None else e @ o
            return o
        tuples = list(itertools.product(range(od), repeat=N))
        tuples.sort(key=lambda t: sum(x * od ** k for k, x in enumerate(t)))      # LinearIndices: first index fastest
        index = {t: k for k, t in enumerate(tuples)}
        n = len(tuples)
        W = [[None] * n for _ in range(n)]
        for a in tuples:
            for b in tuples:
                e = prod_elem(a, b)
                if e is not None:
                    W[index[a]][index[b]] = e
        first, last = 0, od - 1

        def add(i, j, e):
            W[i][j] = e if W[i][j] is None else W[i][j] + e

        def interweave(x, t):
            """all tuples obtained by inserting the single element x somewhere into t (timeevmpo.jl:123-147 for length(a) == 1)"""
            return [t[:k] + (x,) + t[k:] for k in range(len(t) + 1)]
        # (i) embed the next order (timeevmpo.jl:26-45)
        for a in tuples:
            for b in tuples:
                if not all(x > first for x in b):
                    continue
                if all(x in (first, last) for x in a) and any(x == last for x in a):
                    continue
                n3 = sum(x == last for x in b) + 1
                n1 = sum(x == first for x in a) + 1
                for eb in interweave(last, b):
                    for ea in interweave(first, a):
                        e = prod_elem(ea, eb)
                        if e is not None:
                            add(index[a], index[b], e * tau * factorial(N) / (factorial(N + 1) * n1 * n3))

        def zero_rowcol(c):
            for k in range(n):
                W[c][k] = None
                W[k][c] = None
        # (ii) loop back the finished tuples (timeevmpo.jl:48-57)
        for a in itertools.product((first, last), repeat=N):
            a = a[::-1]                                                 # Iterators.product: first index fastest
            if all(x == first for x in a):
                continue
            order = sum(x == last for x in a)
            c = index[a]
            for r in range(c):
                if W[r][c] is not None:
                    add(r, 0, W[r][c] * tau ** order * factorial(N - order) / factorial(N))
            zero_rowcol(c)
        # (iii) merge permutation-equivalent columns / rows (timeevmpo.jl:60-91)
        for c in tuples:
            s_tc = tuple(sorted(c, key=lambda x: 2 if x == first else 1))
            n1 = sum(x == first for x in c); n3 = sum(x == last for x in c)
            if n1 >= n3 and c != s_tc:
                for k in range(n):
                    if W[index[c]][k] is not None:
                        add(index[s_tc], k, W[index[c]][k])
                zero_rowcol(index[c])
        for c in tuples:
            s_tc = tuple(sorted(c, key=lambda x: 2 if x == last else 1))
            n1 = sum(x == first for x in c); n3 = sum(x == last for x in c)
            if n3 > n1 and c != s_tc:
                for k in range(n):
                    if W[k][index[c]] is not None:
                        add(k, index[s_tc], W[k][index[c]])
                zero_rowcol(index[c])
        # (iv) approximate compression (timeevmpo.jl:93-105)
        for c in tuples:
            nl = sum(x == last for x in c)
            if not (all(x > first for x in c) and nl > 0):
                continue
            tr = tuple(first if x == last else x for x in c)
            for k in range(n):
                if W[k][index[c]] is not None:
                    add(k, index[tr], W[k][index[c]] * tau ** nl * factorial(N - nl) / factorial(N))
            zero_rowcol(index[c])
        cl = [sum(sl.domcharges[x] for x in t) for t in tuples]
        cr = [sum(sl.imcharges[x] for x in t) for t in tuples]
        out_slices.append((W, cl, cr, sl))
    # (v) remove_orphans (sparsempo.jl:310-350): levels that nothing enters or nothing leaves, repeated until stable
    n = len(out_slices[0][0])
    keep = list(range(n))
    L = len(out_slices)
    changed = True
    while changed:
        changed = False
        for loc in range(L):
            W = out_slices[loc][0]
            Wp = out_slices[(loc - 1) % L][0]
            Wn = out_slices[(loc + 1) % L][0]
            for i in list(keep):
                dead_start = all(W[i][k] is None for k in keep)         # nothing leaves level i at this site
                dead_end = all(W[k][i] is None for k in keep)           # nothing enters level i at this site
                if dead_start and any(Wp[k][i] is not None for k in keep):
                    for k in keep:
                        Wp[k][i] = None
                    changed = True
                if dead_end and any(Wn[i][k] is not None for k in keep):
                    for k in keep:
                        Wn[i][k] = None
                    changed = True
        alive = [i for i in keep if not all(all(out_slices[loc][0][i][k] is None for k in keep) for loc in range(L))
                 and not all(all(out_slices[loc][0][k][i] is None for k in keep) for loc in range(L))]
        if len(alive) != len(keep):
            keep = alive
            changed = True
    out = []
    for W, cl, cr, sl in out_slices:
        d = sl.d
        ent = [[0 for _ in keep] for _ in keep]
        ncl = [cl[i] for i in keep]; ncr = [cr[i] for i in keep]
        for a, i in enumerate(keep):
            for b, j in enumerate(keep):
                e = W[i][j]
                if e is None:
                    continue
                e = np.array(e)
                for p in range(d):
                    for q in range(d):
                        viol = ncr[b] - ncl[a] - sl.pcharges[p] + sl.pcharges[q]
                        if (viol % sl.modulus if sl.modulus else viol) != 0:
                            if abs(e[p, q]) > 1e-12:
                                raise _lib.SpaceMismatch(_lib.ESPACE, f"time MPO entry ({a},{b}) violates charge conservation")
                            e[p, q] = 0.0
                if np.abs(e).max() > 0:
                    ent[a][b] = e
        out.append(SparseMPOSlice(ent, ncl, ncr, sl.pcharges, sl.modulus))
    return MPOHamiltonian(out)
