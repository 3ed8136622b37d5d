"""Generates tests/golden/su2_pins_oracle.json: oracle (oracle/infinite.py VUMPS, oracle/toolbox.py entropy) runs that reproduce
numbers the REFERENCE ITSELF printed in its documentation for SU(2)-symmetric states:

* docs/src/examples/quantum1d/5.haldane-spt/index.md:385-404,1011-1019 -- spin-1 Heisenberg chain,
  V_minus = SU2Space(1/2 => 10, 3/2 => 5, 5/2 => 3):  E_minus = -1.4014839739630869, S_minus + log(2) = 1.5486227235451633;
  V_plus = SU2Space(0 => 10, 1 => 5, 2 => 3):  E_plus = -1.401419331339301, S_plus = 1.5450323530414438.
* docs/src/examples/quantum1d/4.xxz-heisenberg/index.md:579-671 -- spin-1/2 Heisenberg chain, two-site unit cell,
  V1 = Rep[SU2](1/2 => 10, 3/2 => 5, 5/2 => 2), V2 = Rep[SU2](0 => 15, 1 => 10, 2 => 5):
  "VUMPS conv 73: obj = -8.862878980874e-01" (sum over the unit cell).
(Not used: 2.haldane/index.md:209-300, Rep[SU2](1/2 => 16, 3/2 => 16, 5/2 => 8, 7/2 => 4) -- there VUMPS stops at err 1e-5 and a
conjugate-gradient refinement, which the oracle does not restate, reaches f = -1.401484038967; the projected oracle VUMPS from a
random symmetric start settles, after 22 iterations / 25 min at D = 176, near a different stationary point, -1.4014840254.)

The oracle has U(1) blocks only.  An SU(2)-symmetric MPS in the U(1) basis: every multiplet (j, alpha) contributes its 2j+1 members
m = -j..j to the U(1) sectors of charge 2m, and the tensor is  A[(j1 m1 a), ms, (j2 m2 b)] = R_{j1 j2}[a, b] <j1 m1; s ms | j2 m2>
with random reduced tensors R and Clebsch-Gordan coefficients in the Condon-Shortley convention (the convention of
oracle/mpo.py:spin_ops).  The effective Hamiltonians of a symmetric state commute with the symmetry, so VUMPS restricted to the
invariant subspace (oracle.infinite.vumps(project=...): every application of dAC / dC is followed by the orthogonal projector onto
invariant tensors) is exactly the iteration the reference runs on its reduced tensors, and must land on the reference's numbers.
Without the projector rounding noise in the non-invariant sectors is amplified by the Krylov solvers and the U(1) run leaves the
manifold within five iterations: it converges to a LOWER, symmetry-broken energy (-1.40148402648 for V_minus; recorded as
minus_generic_u1), as it may by the variational principle -- the U(1) manifold of the same sectors is larger.
The requested multiplet spaces first go through the reference's makefullrank! (fullrank_cells below): V_plus keeps 5 of its 10
singlets, the xxz cell 10 of 15.  V_plus is a non-injective state (the "wrong" SPT phase); the projected iteration handles it too.
    python tests/golden/make_su2_pins.py [minus plus xxz minus_generic_u1]     (45 s, 20 s, 3 min, 1 min on 8 cores)"""
import json
import os
import sys
import time
from fractions import Fraction
from math import factorial, sqrt
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from oracle.spaces import Space
from oracle import mpo as ompo, infinite as oi, toolbox as ot

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "su2_pins_oracle.json")


def clebsch_gordan(j1, m1, j2, m2, j, m):
    """<j1 m1; j2 m2 | j m>, Racah's formula, Condon-Shortley phases; arguments are Fractions (or ints)."""
    j1, m1, j2, m2, j, m = (Fraction(x) for x in (j1, m1, j2, m2, j, m))
    if m1 + m2 != m or not (abs(j1 - j2) <= j <= j1 + j2) or abs(m1) > j1 or abs(m2) > j2 or abs(m) > j:
        return 0.0
    if (j1 + j2 + j).denominator != 1:
        return 0.0
    f = lambda x: factorial(int(x))
    pre = Fraction((2 * j + 1) * f(j + j1 - j2) * f(j - j1 + j2) * f(j1 + j2 - j), f(j1 + j2 + j + 1))
    pre *= f(j + m) * f(j - m) * f(j1 - m1) * f(j1 + m1) * f(j2 - m2) * f(j2 + m2)
    s = Fraction(0)
    for k in range(0, int(j1 + j2 - j) + 1):
        args = (k, j1 + j2 - j - k, j1 - m1 - k, j2 + m2 - k, j - j2 + m1 + k, j - j1 - m2 + k)
        if any(a < 0 for a in args):
            continue
        den = 1
        for a in args:
            den *= f(a)
        s += Fraction((-1) ** k, den)
    return float(s) * sqrt(float(pre))


def u1_space(irreps):
    """{j: n} -> (Space with charges 2m, index map {(j, m, alpha): position inside the dense bond index})"""
    dims = {}
    for j, n in irreps.items():
        j = Fraction(j)
        for k in range(int(2 * j) + 1):
            q = int(2 * (-j + k))
            dims[q] = dims.get(q, 0) + n
    V = Space.from_dict(dims)
    qs = V.index_charges()
    start = {int(c): int(np.nonzero(qs == c)[0][0]) for c in V.charges}
    fill = {c: 0 for c in start}
    pos = {}
    for j in sorted(Fraction(x) for x in irreps):
        n = irreps[j] if j in irreps else irreps[float(j)]
        for k in range(int(2 * j) + 1):
            m = -j + k
            q = int(2 * m)
            for a in range(n):
                pos[(j, m, a)] = start[q] + fill[q]
                fill[q] += 1
    return V, pos


def symmetric_tensor(rng, irl, s, irr):
    """random SU(2)-invariant A: V(irl) x spin-s -> V(irr) in the U(1) basis (physical index m = -s..s ascending)."""
    irl = {Fraction(k): v for k, v in irl.items()}
    irr = {Fraction(k): v for k, v in irr.items()}
    s = Fraction(s)
    Vl, pl = u1_space(irl)
    Vr, pr = u1_space(irr)
    d = int(2 * s) + 1
    A = np.zeros((Vl.dim, d, Vr.dim), dtype=np.complex128)
    for j1, n1 in irl.items():
        for j2, n2 in irr.items():
            if not (abs(j1 - s) <= j2 <= j1 + s) or (j1 + s + j2).denominator != 1:
                continue
            R = rng.standard_normal((n1, n2)) + 1j * rng.standard_normal((n1, n2))
            for k1 in range(int(2 * j1) + 1):
                m1 = -j1 + k1
                for ks in range(d):
                    ms = -s + ks
                    m2 = m1 + ms
                    if abs(m2) > j2:
                        continue
                    cg = clebsch_gordan(j1, m1, s, ms, j2, m2)
                    rows = [pl[(j1, m1, a)] for a in range(n1)]
                    cols = [pr[(j2, m2, b)] for b in range(n2)]
                    A[np.ix_(rows, [ks], cols)] = (cg * R)[:, None, :]
    return A, Vl, Vr


def tensor_projector(irl, s, irr):
    """orthogonal projector of (Dl, d, Dr) tensors onto the SU(2)-invariant ones  sum_{j1 j2} R_{j1 j2} (x) CG"""
    irl = {Fraction(k): v for k, v in irl.items()}
    irr = {Fraction(k): v for k, v in irr.items()}
    s = Fraction(s)
    _, pl = u1_space(irl)
    _, pr = u1_space(irr)
    d = int(2 * s) + 1
    plan = []
    for j1, n1 in irl.items():
        for j2, n2 in irr.items():
            if not (abs(j1 - s) <= j2 <= j1 + s) or (j1 + s + j2).denominator != 1:
                continue
            terms = []
            for k1 in range(int(2 * j1) + 1):
                m1 = -j1 + k1
                for ks in range(d):
                    m2 = m1 - s + ks
                    if abs(m2) > j2:
                        continue
                    cg = clebsch_gordan(j1, m1, s, -s + ks, j2, m2)
                    if cg != 0.0:
                        terms.append((cg, np.array([pl[(j1, m1, a)] for a in range(n1)]), ks,
                                      np.array([pr[(j2, m2, b)] for b in range(n2)])))
            plan.append((float(2 * j2 + 1), terms))

    def proj(x):
        y = np.zeros_like(x)
        for nrm, terms in plan:
            R = sum(cg * x[np.ix_(r, [ks], c)][:, 0, :] for cg, r, ks, c in terms) / nrm
            for cg, r, ks, c in terms:
                y[np.ix_(r, [ks], c)] = (cg * R)[:, None, :]
        return y
    return proj


def matrix_projector(ir):
    """orthogonal projector of bond matrices onto  sum_j c_j (x) 1_{2j+1}"""
    ir = {Fraction(k): v for k, v in ir.items()}
    _, pos = u1_space(ir)
    plan = [[np.array([pos[(j, -j + k, a)] for a in range(n)]) for k in range(int(2 * j) + 1)] for j, n in ir.items()]

    def proj(x):
        y = np.zeros_like(x)
        for rows in plan:
            c = sum(x[np.ix_(r, r)] for r in rows) / len(rows)
            for r in rows:
                y[np.ix_(r, r)] = c
        return y
    return proj


def multiplet_defect(st, irreps_per_bond):
    """largest deviation from the (2j+1)-fold degeneracy the entanglement spectrum of a symmetric state must show: in the sector
    of charge 2m the singular values are those of all multiplets with j >= |m|, so sector m = j_max holds the j_max values,
    sector j_max - 1 those plus the (j_max - 1) values, ...; checked as: the spectrum of charge q equals that of charge -q."""
    worst = 0.0
    for site in range(st.N):
        spec = ot.entanglement_spectrum(st, site)
        for q, v in spec.items():
            if -q in spec:
                worst = max(worst, float(np.max(np.abs(np.sort(v) - np.sort(spec[-q])))))
    return worst


def fullrank_cells(cells, spin):
    """makefullrank! (src/states/abstractmps.jl:94-108, called by the InfiniteMPS constructor, infinitemps.jl:150) on the level of
    multiplet multiplicities: n_r[j2] <= sum_{j1 : j2 in j1 x s} n_l[j1] and n_l[j1] <= sum_{j2 in j1 x s} n_r[j2], to a fixed
    point.  E.g. Rep[SU2](0 => 15, 1 => 10, 2 => 5) next to Rep[SU2](1/2 => 10, 3/2 => 5, 5/2 => 2) with spin 1/2 keeps only 10
    of the 15 singlets (the printed dim(V2) = 70 of 4.xxz-heisenberg/index.md:596-597 is the space before this reduction)."""
    s = Fraction(spin)
    cells = [{Fraction(k): int(v) for k, v in c.items()} for c in cells]
    N = len(cells)
    couples = lambda j1, j2: abs(j1 - s) <= j2 <= j1 + s and (j1 + s + j2).denominator == 1
    changed = True
    while changed:
        changed = False
        for i in range(N):
            l, r = cells[i], cells[(i + 1) % N]
            for j2 in list(r):
                cap = sum(n for j1, n in l.items() if couples(j1, j2))
                if r[j2] > cap:
                    r[j2], changed = cap, True
            for j1 in list(l):
                cap = sum(n for j2, n in r.items() if couples(j1, j2))
                if l[j1] > cap:
                    l[j1], changed = cap, True
        cells = [{j: n for j, n in c.items() if n > 0} for c in cells]
    return cells


def run(name, spin, cells, tol, maxiter, seed=7, symmetric=True, save_state=None):
    requested = [{str(Fraction(k)): v for k, v in c.items()} for c in cells]
    cells = fullrank_cells(cells, spin)
    d = int(2 * spin) + 1
    N = len(cells)
    H = ompo.heisenberg_u1(N, d=d)
    p = Space(tuple(int(2 * (-spin + k)) for k in range(d)), (1,) * d)
    rng = np.random.default_rng(seed)
    t0 = time.time()
    if symmetric:
        A, Vs = [], []
        for i in range(N):
            a, Vl, _ = symmetric_tensor(rng, cells[i], spin, cells[(i + 1) % N])
            A.append(a)
            Vs.append(Vl)
        st = oi.InfiniteMPS.from_tensors(A, Vs, p.index_charges(), mod=0)
        pa = [tensor_projector(cells[i], spin, cells[(i + 1) % N]) for i in range(N)]
        pc = [matrix_projector(cells[(i + 1) % N]) for i in range(N)]
        project = (lambda loc, x: pa[loc](x), lambda loc, x: pc[loc](x))
    else:
        Vs = [u1_space(c)[0] for c in cells]
        st = oi.InfiniteMPS.random(rng, p, Vs)
        project = None
    class _Log(list):
        def append(self, rec):
            list.append(self, rec)
            if os.environ.get("MPSK_VERBOSE"):
                print(f"{name} {rec['iter']:4d}: obj = {rec['energy']:+.12e} err = {rec['eps']:.4e}", flush=True)
    log = _Log()
    st, envs, eps = oi.vumps(st, H, tol=tol, maxiter=maxiter, log=log, project=project)
    S = ot.entropy(st)
    if save_state:       # the converged left-gauged tensors, for the device read-out test (tests/unverified_gpu_cases.py)
        np.savez_compressed(save_state, **{f"AL{i}": st.AL[i] for i in range(N)}, **{f"C{i}": st.CR[i] for i in range(N)},
                            charges=np.array(Vs[0].charges), dims=np.array(Vs[0].dims),
                            **{f"charges{i}": np.array(Vs[i].charges) for i in range(N)},
                            **{f"dims{i}": np.array(Vs[i].dims) for i in range(N)})
    rec = dict(name=name, symmetric_start=symmetric, spin=float(spin), requested_irreps=requested,
               irreps=[{str(Fraction(k)): v for k, v in c.items()} for c in cells],
               bond_spaces=[{str(k): v for k, v in V.as_dict().items()} for V in Vs], D=[V.dim for V in Vs],
               vumps_iterations=len(log), galerkin=eps, energy_unit_cell=log[-1]["energy"], entropy=[float(x) for x in S],
               multiplet_defect=multiplet_defect(st, cells), seconds=time.time() - t0)
    print(rec, flush=True)
    return rec


H = Fraction(1, 2)
CASES = {
    "minus": lambda: run("haldane-spt V_minus", 1, [{H: 10, 3 * H: 5, 5 * H: 3}], 1e-10, 100,
                         save_state=os.path.join(os.path.dirname(OUT), "su2_haldane_spt_minus_state.npz")),
    "minus_generic_u1": lambda: run("haldane-spt V_minus sectors, generic U(1) start", 1, [{H: 10, 3 * H: 5, 5 * H: 3}], 1e-10, 100,
                                    seed=3, symmetric=False),
    "plus": lambda: run("haldane-spt V_plus", 1, [{0: 10, 1: 5, 2: 3}], 1e-10, 100),
    "xxz": lambda: run("xxz-heisenberg SU(2) two-site cell", H, [{0: 15, 1: 10, 2: 5}, {H: 10, 3 * H: 5, 5 * H: 2}], 1e-12, 400,
                       save_state=os.path.join(os.path.dirname(OUT), "su2_xxz_state.npz")),
}
REFERENCE = {
    "minus": dict(source="docs/src/examples/quantum1d/5.haldane-spt/index.md:397-404,1018", energy=-1.4014839739630869,
                  entropy=1.5486227235451633 - float(np.log(2.0))),
    "plus": dict(source="docs/src/examples/quantum1d/5.haldane-spt/index.md:385-393,1019", energy=-1.401419331339301,
                 entropy=1.5450323530414438),
    "xxz": dict(source="docs/src/examples/quantum1d/4.xxz-heisenberg/index.md:671", energy=-8.862878980874e-01),
}

if __name__ == "__main__":
    todo = sys.argv[1:] or ["minus", "minus_generic_u1", "plus", "xxz"]
    for k in todo:
        rec = CASES[k]()
        out = json.load(open(OUT)) if os.path.exists(OUT) else dict(reference=REFERENCE, runs={})     # re-read: runs may overlap
        out["reference"] = REFERENCE
        out["runs"][k] = rec
        json.dump(out, open(OUT, "w"), indent=1)
