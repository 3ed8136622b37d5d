"""CPU tests of the boundary: the C-ABI library loads, exports every symbol the header declares, and its integer
bookkeeping matches the oracle bit-exactly.  No compute calls (no GPU here)."""
import ctypes as C
import os
import re
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "mpskit_b200.h")).read()
    return sorted(set(re.findall(r"^(?:int32_t|int64_t|void|double|const char\*)\s+(mpsk_[a-z0-9_]+)\s*\(", src, re.M)))


def test_library_exports_every_declared_symbol():
    import mpskit_b200 as m
    from mpskit_b200 import _lib
    lib = m.load_library()
    syms = header_symbols()
    assert len(syms) >= 40
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/mpskit_b200.h but not exported"
        assert s in _lib.PROTOTYPES, f"{s} has no ctypes prototype"
    assert lib.mpsk_version() >= 100


def test_no_cpu_fallback():
    """without a GPU context creation must fail loudly (MPSK_ENODEVICE), never fall back to the CPU"""
    import mpskit_b200 as m
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("GPU present")
    except ImportError:
        pass
    with pytest.raises(m.MpskError) as e:
        m.Context(0)
    assert e.value.code == -6


def test_product_package_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "mpskit.jl_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "import oracle" not in src and "from oracle" not in src, fn


def test_block_table_matches_oracle_bit_exactly():
    import mpskit_b200 as m
    from oracle.spaces import Space as OSpace
    from oracle.blocks import BlockTensor
    rng = np.random.default_rng(0)
    for trial in range(50):
        def rs():
            k = rng.integers(1, 7)
            ch = np.sort(rng.choice(np.arange(-8, 9), size=k, replace=False))
            return OSpace(tuple(ch), tuple(rng.integers(1, 9, size=k)))
        Vl, Vr = rs(), rs()
        mch = [int(c) for c in rng.integers(-3, 4, size=rng.integers(1, 6))]
        ref = BlockTensor.block_keys(Vl, mch, Vr)
        got = m.block_table(m.Space(Vl.charges, Vl.dims), mch, m.Space(Vr.charges, Vr.dims))
        assert got == [tuple(int(v) for v in t) for t in ref]
    Vl = OSpace((0, 1), (3, 2), 2); Vr = OSpace((0, 1), (1, 4), 2)
    assert m.block_table(m.Space((0, 1), (3, 2), 2), [1, 0], m.Space((0, 1), (1, 4), 2)) == \
        [tuple(int(v) for v in t) for t in BlockTensor.block_keys(Vl, [1, 0], Vr)]


def test_host_mirror_bookkeeping():
    import mpskit_b200 as m
    from oracle import mpo as ompo
    for r in (1, 2, 6):
        so = ompo.heisenberg_longrange_u1(4, r=r)[0]
        sd = m.heisenberg_longrange_u1(4, r=r)[0]
        assert sd.keys() == so.keys() and sd.kinds() == so.kinds()
        assert sd.opkeys() == so.opkeys() and sd.scalkeys() == so.scalkeys()
        assert sd.domcharges == so.cl and sd.pcharges == so.pch
    from oracle import finite as of
    from oracle.spaces import Space as OSpace
    from mpskit_b200.workloads import gaussian_bond_space
    for D in (16, 256, 4096):
        for par in (0, 1):
            assert gaussian_bond_space(D, par).as_dict() == of.gaussian_bond_space(D, par).as_dict()
    p = m.Space((-1, 1), (1, 1))
    Vd = m.virtual_spaces(p, 10, [gaussian_bond_space(16, k % 2) for k in range(1, 10)])
    Vo = of.virtual_spaces(OSpace((-1, 1), (1, 1)), 10, [of.gaussian_bond_space(16, k % 2) for k in range(1, 10)])
    assert [v.as_dict() for v in Vd] == [v.as_dict() for v in Vo]


def test_bad_spaces_are_rejected_without_a_gpu():
    import mpskit_b200 as m
    from mpskit_b200 import _lib
    lib = m.load_library()
    q = (C.c_int32 * 2)(3, 1); n = (C.c_int64 * 2)(1, 1)         # not ascending
    bad = _lib.c_space(2, q, n, 0)
    good, keep = m.Space((0,), (1,)).c_struct()
    mc = (C.c_int32 * 1)(0)
    cnt = C.c_int32()
    assert lib.mpsk_block_table(C.byref(bad), 1, mc, C.byref(good), C.byref(cnt), None, None, None) == _lib.ESPACE


def test_ac_tuple_setter_bookkeeping_without_gpu():
    """orthoview.jl:121-143: psi.AC[i] = (al, c) / (c, ar) -- gauge invalidation bookkeeping of the host mirror, exercised
    on placeholder tensors (no device calls are made on this path)."""
    import types
    import mpskit_b200 as m
    from mpskit_b200.finitemps import FiniteMPS, _View

    def T(mch, Vl="l", Vr="r", bond=False):
        t = types.SimpleNamespace(mch=tuple(mch), Vl=Vl, Vr=Vr)
        if bond:
            t._is_bond = True
        return t
    L = 4
    psi = object.__new__(FiniteMPS)
    psi.L, psi.pq = L, (-1, 1)
    psi.V = ["V%d" % i for i in range(L + 1)]
    psi.ALs = [T((-1, 1)) for _ in range(L)]
    psi.ARs = [T((-1, 1)) for _ in range(L)]
    psi.ACs = [T((-1, 1)) for _ in range(L)]
    psi.CLs = [T((0,), bond=True) for _ in range(L + 1)]
    psi._views()
    al, c, ar = T((-1, 1), Vr="new"), T((0,), Vl="new", Vr="new"), T((-1, 1), Vl="new")     # c carries no flag: structural detection
    keepR = psi.ARs[3]
    psi.AC[1] = (al, c)
    assert psi.ALs[1] is al and psi.CLs[2] is c and psi.V[2] == "new"
    assert psi.ALs[2] is None and psi.ARs[0] is None and psi.ARs[1] is None and psi.ARs[3] is keepR
    assert all(a is None for a in psi.ACs) and sum(x is not None for x in psi.CLs) == 1
    psi.ACs[2] = T((-1, 1))            # pretend AC[2] is known so that no lazy gauge is triggered
    psi.AC[2] = (c, ar)
    assert psi.ARs[2] is ar and psi.CLs[2] is c and psi.ALs[1] is al and psi.ALs[2] is None
    import pytest
    with pytest.raises(ValueError):
        psi.ACs[2] = T((-1, 1))
        psi.AC[2] = (al, ar)


def test_header_is_plain_c99_and_links_from_c(tmp_path):
    """the boundary is a C ABI: include/mpskit_b200.h must compile as C99 (-pedantic) and link from a plain C program"""
    import subprocess
    import mpskit_b200 as m
    src = tmp_path / "hdr.c"
    src.write_text('#include "mpskit_b200.h"\nint main(void){ mpsk_ctx* c = 0; return mpsk_version() >= 100 && c == 0 ? 0 : 1; }\n')
    libdir = os.path.dirname(m.LIB_PATH)
    exe = tmp_path / "hdr"
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                        "-L", libdir, "-lmpskit_b200", "-Wl,-rpath," + libdir], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert subprocess.run([str(exe)]).returncode == 0


def test_host_schur_and_reordering_against_numpy():
    """the complex Schur decomposition + reordering behind the device Krylov-Schur eigensolver (csrc/eigsolve.cu) is pure
    host code: checked here without a GPU against numpy (A = Z T Z^H, T upper triangular, Z unitary, eigenvalues)."""
    import ctypes as C
    import numpy as np
    import mpskit_b200 as m
    lib = m.load_library()
    DP = C.POINTER(C.c_double)
    rng = np.random.default_rng(1)
    for n in (1, 2, 3, 7, 30, 61):
        A = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
        if n == 7:
            A = np.triu(A, -1)                                        # already Hessenberg (plain Arnoldi)
        H = np.asfortranarray(A.copy()); Z = np.zeros((n, n), dtype=complex, order="F")
        assert lib.mpsk_host_schur(n, H.ctypes.data_as(DP), Z.ctypes.data_as(DP), -1, -1) == 0
        assert np.abs(Z @ H @ Z.conj().T - A).max() <= 1e-12 * max(1.0, np.abs(A).max()) * n
        assert np.abs(np.tril(H, -1)).max() == 0.0
        assert np.abs(Z.conj().T @ Z - np.eye(n)).max() <= 1e-13 * n
        ev = np.linalg.eigvals(A)
        assert max(np.abs(ev - d).min() for d in np.diag(H)) <= 1e-11 * n
        if n >= 3:
            H2 = np.asfortranarray(A.copy()); Z2 = np.zeros((n, n), dtype=complex, order="F")
            assert lib.mpsk_host_schur(n, H2.ctypes.data_as(DP), Z2.ctypes.data_as(DP), n - 1, 0) == 0
            assert abs(H2[0, 0] - H[n - 1, n - 1]) <= 1e-12 * n       # the last eigenvalue moved to the front
            assert np.abs(Z2 @ H2 @ Z2.conj().T - A).max() <= 1e-12 * max(1.0, np.abs(A).max()) * n
            assert np.abs(np.tril(H2, -1)).max() == 0.0
            # first Schur vector = eigenvector of the leading eigenvalue
            assert np.linalg.norm(A @ Z2[:, 0] - H2[0, 0] * Z2[:, 0]) <= 1e-11 * n


def test_device_allocator_free_list_logic_on_host():
    """devalloc.h (split / coalesce / trim) under random traffic with malloc as the backend: no overlap of live blocks,
    free blocks disjoint and merged, live + free == reserved after every operation, everything returned at the end."""
    import mpskit_b200 as m
    lib = m.load_library()
    for seed in (1, 2, 3, 4):
        assert lib.mpsk_host_alloc_selftest(seed, 15000) == 0


def test_state_file_pack_unpack_roundtrip(tmp_path):
    """serialize.py host halves: a block tensor written in ABI block order and read back through np.savez is bit-identical,
    and a truncated data vector is rejected."""
    import mpskit_b200 as m
    rng = np.random.default_rng(5)
    Vl, Vr = m.Space((-2, 0, 2), (3, 5, 2)), m.Space((-3, -1, 1, 3), (2, 4, 4, 1))
    mch = [-1, 1]
    table = m.block_table(Vl, mch, Vr)
    blocks = {(mm, jl): rng.standard_normal((Vl.dims[jl], Vr.dims[jr])) + 1j * rng.standard_normal((Vl.dims[jl], Vr.dims[jr]))
              for (mm, jl, jr) in table}
    store = {}
    m.pack_tensor("AL/0", Vl, mch, Vr, np.complex128, blocks, table, store)
    np.savez(tmp_path / "t.npz", **store)
    with np.load(tmp_path / "t.npz") as st:
        Vl2, mch2, Vr2, dt, b2 = m.unpack_tensor("AL/0", st, m.block_table)
        assert (Vl2.charges, Vl2.dims, Vr2.charges, Vr2.dims, mch2) == (Vl.charges, Vl.dims, Vr.charges, Vr.dims, mch)
        assert dt == np.complex128 and set(b2) == set(blocks)
        assert all(np.array_equal(b2[k], blocks[k]) for k in blocks)
        bad = {k: st[k] for k in st.files}
    bad["AL/0/data"] = bad["AL/0/data"][:-1]
    with pytest.raises(ValueError):
        m.unpack_tensor("AL/0", bad, m.block_table)


def test_toolbox_host_logic_without_gpu():
    """host-side pieces of mpskit_b200/toolbox.py (no device call): the per-charge decomposition of multi-site operators into
    local MPO tensors (decompose_localmpo of expval.jl:41-44 with definite level charges), the square-slice embedding of
    rectangular local MPO tensors, rationalize / approx_angles / marek_gap (toolbox.jl:74-117) against the oracle's."""
    import numpy as np
    from mpskit_b200 import toolbox as tb
    from oracle import toolbox as ot, mpo as ompo
    Sp, Sm, Sz, pq = ompo.spin_ops(2)
    SS = np.einsum('ac,bd->abcd', Sz, Sz) + 0.5 * (np.einsum('ac,bd->abcd', Sp, Sm) + np.einsum('ac,bd->abcd', Sm, Sp))
    Os, ch = tb.decompose_localmpo(SS, pq, return_charges=True)
    assert [o.shape for o in Os] == [(1, 2, 2, 3), (3, 2, 2, 1)] and ch == [[0], [-2, 0, 2], [0]]
    assert np.abs(np.einsum('apqb,brsc->prqs', Os[0], Os[1]) - SS).max() < 1e-14
    assert tb._level_charges_from_ops(Os, pq) == ch
    H3 = np.einsum('abcd,ef->abecdf', SS, np.eye(2)) + np.einsum('ef,abcd->eabfcd', np.eye(2), SS)
    Os3, ch3 = tb.decompose_localmpo(H3, pq, return_charges=True)
    assert np.abs(np.einsum('apqb,brsc,ctud->prtqsu', *Os3) - H3).max() < 1e-13 and ch3[0] == [0] and ch3[-1] == [0]
    slices, od, lc = tb.local_mpo_slices(Os, pq, ch)
    assert od == 3 and slices[0].kinds() == [[2, 2, 2], [0, 0, 0], [0, 0, 0]] and slices[1].kinds() == [[2, 0, 0], [2, 0, 0], [2, 0, 0]]
    assert slices[0].imcharges == (-2, 0, 2) and slices[1].domcharges == (-2, 0, 2) and slices[1].imcharges == (0, 0, 0)
    for x, tol in ((0.3333, 0.01), (-0.98, 0.1), (0.52, 0.1), (0.0, 0.1), (0.749, 0.01)):
        assert tb.rationalize(x, tol) == ot.rationalize(x, tol)
    spec = np.array([1.0, -0.5, 0.4, -0.25, -0.2, 0.1j])
    assert np.allclose(tb.approx_angles(spec), ot.approx_angles(spec))
    assert np.allclose(tb.marek_gap(spec), ot.marek_gap(spec))
    assert abs(tb.correlation_length(spec) - 1 / np.log(2)) < 1e-14


def test_make_time_mpo_wii_host_logic_without_gpu():
    """make_time_mpo(H, dt, WII()) (timeevmpo.jl:149-207) is host-side model building: the W^II MPO of an FSM Hamiltonian
    against exp(-i dt H) of exact linear algebra on L = 6 chains (error ~ dt^2: ratio 4 when dt halves; TFI, U(1) Heisenberg,
    range-3 Heisenberg with scalar decay entries), odim - 1 levels with the first and last level merged, every entry charge
    conserving (checked by the oracle's slice constructor); and through the oracle: VOMPS of W^II |psi> equals a TDVP step
    far better than dt (test/algorithms.jl:447-466 asserts dt)."""
    import numpy as np
    import scipy.linalg as sla
    import mpskit_b200 as m
    from mpskit_b200.operators import make_time_mpo
    from oracle import mpo as ompo, infinite as oi, statmech as osm, changebonds as ocb
    from oracle.spaces import Space

    def dense_mpo(W, L):
        cur = None
        for i in range(L):
            sl = W[i]; od, d = sl.odim, sl.d
            T = np.zeros((od, d, d, od), dtype=complex)
            for a in range(od):
                for b in range(od):
                    e = sl.Os[a][b]
                    if np.isscalar(e):
                        if e != 0:
                            T[a, :, :, b] = e * np.eye(d)
                    else:
                        T[a, :, :, b] = np.asarray(e)
            if cur is None:
                cur = T[0]
            else:
                cur = np.einsum('abw,wcdv->acbdv', cur, T)
                s = cur.shape
                cur = cur.reshape(s[0] * s[1], s[2] * s[3], s[4])
        return cur[:, :, 0]
    L = 6
    for Hd, Ho in ((m.transverse_field_ising(L, 1.3), ompo.transverse_field_ising(L, 1.3)), (m.heisenberg_u1(L), ompo.heisenberg_u1(L)),
                   (m.heisenberg_longrange_u1(L, r=3), ompo.heisenberg_longrange_u1(L, r=3))):
        Hm = ompo.dense_hamiltonian(Ho)
        errs = []
        for dt in (0.1, 0.05, 0.025):
            W = make_time_mpo(Hd, dt)
            assert W[0].odim == Hd[0].odim - 1 and W[0].domcharges[0] == 0
            for sl in W.slices:
                ompo.SparseMPOSlice(sl.Os, sl.domcharges, sl.imcharges, sl.pcharges)      # asserts charge conservation
            errs.append(np.abs(dense_mpo(W, L) - sla.expm(-1j * dt * Hm)).max())
        assert errs[2] < 2e-3 and 3.8 < errs[0] / errs[1] < 4.2 and 3.8 < errs[1] / errs[2] < 4.2
    dt = 1e-3
    W = make_time_mpo(m.heisenberg_u1(2), dt)
    sl = [ompo.SparseMPOSlice(s.Os, s.domcharges, s.imcharges, s.pcharges) for s in W.slices]
    st = oi.InfiniteMPS.random(np.random.default_rng(5), Space((-1, 1), (1, 1)), [Space((-2, 0, 2), (2, 3, 2)), Space((-1, 1), (3, 3))])
    a, _, eps = osm.vomps(st, sl, st, tol=1e-12, maxiter=100)
    t, _ = oi.tdvp_step(st, ompo.heisenberg_u1(2), dt)
    assert eps < 1e-11 and abs(abs(ocb.dot(a, t)) - 1) < 1e-10 and abs(abs(ocb.dot(st, t)) - 1) > 1e-8


def test_find_groundstate_dispatch_and_union_algorithms_without_gpu(monkeypatch):
    """host-side dispatch of find_groundstate (find_groundstate.jl:19-41, unionalg.jl:4-25): the keyword form picks DMRG
    (FiniteMPS), DMRG2 & DMRG with a truncation scheme, VUMPS / IDMRG2 & VUMPS (InfiniteMPS); alg1 & alg2 runs both in order and
    hands state and environments on.  The solvers are replaced by recorders: no device call."""
    import mpskit_b200 as m
    from mpskit_b200 import algorithms as alg_mod, infinite as di
    calls = []

    class Fin:
        L = 4

        def copy(self):
            return self

    class Inf:
        N = 2
    monkeypatch.setattr(alg_mod, "_dmrg2_inplace", lambda psi, H, a, envs, log: (calls.append(("DMRG2", a, envs)) or (psi, "envs-after-dmrg2", 0.0)))
    monkeypatch.setattr(alg_mod, "find_groundstate_dmrg_inplace", lambda psi, H, a, envs, log: (calls.append(("DMRG", a, envs)) or (psi, "envs-after-dmrg", 0.0)))
    monkeypatch.setattr(di, "find_groundstate_vumps", lambda psi, H, a, envs, log: (calls.append(("VUMPS", a, envs)) or (psi, "envs-after-vumps", 0.0)))
    monkeypatch.setattr(di, "find_groundstate_idmrg2", lambda psi, H, a, envs, log: (calls.append(("IDMRG2", a, envs)) or (psi, "envs-after-idmrg2", 0.0)))
    m.find_groundstate(Fin(), "H", tol=1e-9, maxiter=7)
    assert [c[0] for c in calls] == ["DMRG"] and calls[0][1].tol == 1e-9 and calls[0][1].maxiter == 7
    calls.clear()
    m.find_groundstate(Fin(), "H", tol=1e-9, trscheme=m.truncdim(8))
    assert [c[0] for c in calls] == ["DMRG2", "DMRG"] and abs(calls[0][1].tol - 1e-7) < 1e-20 and calls[1][2] == "envs-after-dmrg2"
    calls.clear()
    m.find_groundstate(Inf(), "H", tol=1e-10, trscheme=m.truncdim(8))
    assert [c[0] for c in calls] == ["IDMRG2", "VUMPS"] and calls[1][1].tol == 1e-10 and calls[1][2] == "envs-after-idmrg2"
    calls.clear()
    u = m.DMRG2(trscheme=m.truncdim(4)) & m.DMRG(tol=1e-6) & m.DMRG(tol=1e-12)
    assert isinstance(u, m.UnionAlg) and isinstance(u.alg1, m.UnionAlg)
    m.find_groundstate(Fin(), "H", u)
    assert [c[0] for c in calls] == ["DMRG2", "DMRG", "DMRG"] and calls[2][1].tol == 1e-12
    assert isinstance(m.SvdCut() & m.OptimalExpand(), m.UnionAlg)
    with pytest.raises(TypeError):
        m.find_groundstate(Inf(), "H", m.DMRG())


def test_iterlog_reproduces_the_reference_log_lines():
    """mpskit_b200/iterlog.py (src/utility/logging.jl:12-142): the exact line formats the reference prints, checked against lines
    of its own documentation (docs/src/examples/quantum1d/4.xxz-heisenberg/index.md:42-43,142; 3.ising-dqpt/index.md), the
    verbosity levels of defaults.jl:12-16, and laziness of the objective (never evaluated when its level is off)."""
    import io
    from mpskit_b200.iterlog import IterLog, format_time
    buf = io.StringIO()
    lg = IterLog("VUMPS", 3, buf)
    lg.init(7.9836e-03, lambda: 2.499921743436e-01)
    lg.iteration(1, 3.7172694563e-01, lambda: -1.716410056492e-01)
    lg.cancel(100, 2.3131204986e-01, lambda: -4.061058847078e-01)
    lines = buf.getvalue().splitlines()
    assert lines[0] == "[ Info: VUMPS init:\tobj = +2.499921743436e-01\terr = 7.9836e-03"
    assert lines[1].startswith("[ Info: VUMPS   1:\tobj = -1.716410056492e-01\terr = 3.7172694563e-01\ttime = ") and lines[1].endswith(" sec")
    assert lines[2].startswith("┌ Warning: VUMPS cancel 100:\tobj = -4.061058847078e-01\terr = 2.3131204986e-01\ttime = ")
    assert format_time(4.5) == "4.50 sec" and format_time(120.0) == "2.00 min" and format_time(7200.0) == "2.00 hr"
    called = []
    quiet = IterLog("DMRG", 0, buf)
    quiet.init(1.0, lambda: called.append(1)); quiet.iteration(1, 0.5, lambda: called.append(1))
    quiet.finish(2, 0.1, lambda: called.append(1)); quiet.cancel(3, 0.1, lambda: called.append(1))
    assert not called and len(buf.getvalue().splitlines()) == 3
    conv = IterLog("DMRG", 2, buf)
    conv.iteration(1, 0.5, lambda: called.append(1))            # level 3 line: silent at verbosity 2
    conv.finish(2, 1e-13, lambda: -20.40021786703)
    assert not called and buf.getvalue().splitlines()[-1].startswith("[ Info: DMRG conv 2:\tobj = -2.040021786703e+01\terr = 1.0000000000e-13\ttime = ")


def test_make_time_mpo_taylor_cluster_orders_without_gpu():
    """make_time_mpo(H, dt, TaylorCluster{N}()) (timeevmpo.jl:12-108) is host-side model building: the MPO reproduces
    exp(-i dt H) on an L = 6 U(1) Heisenberg chain with an error that falls by 2^(N+1) when dt halves (N = 1, 2, 3: measured
    3.99, 7.98, 15.98) -- a wrong weight anywhere in the embedding / loop-back / merging / compression steps breaks the order --
    and every entry conserves charge (WI = TaylorCluster{1})."""
    import numpy as np
    import scipy.linalg as sla
    import mpskit_b200 as m
    from mpskit_b200.operators import make_time_mpo
    from oracle import mpo as ompo

    def dense_mpo(W, L):
        cur = None
        for i in range(L):
            sl = W[i]; od, d = sl.odim, sl.d
            T = np.zeros((od, d, d, od), dtype=complex)
            for a in range(od):
                for b in range(od):
                    e = sl.Os[a][b]
                    if not np.isscalar(e):
                        T[a, :, :, b] = np.asarray(e)
                    elif e != 0:
                        T[a, :, :, b] = e * np.eye(d)
            if cur is None:
                cur = T[0]
            else:
                cur = np.einsum('abw,wcdv->acbdv', cur, T)
                s = cur.shape
                cur = cur.reshape(s[0] * s[1], s[2] * s[3], s[4])
        return cur[:, :, 0]
    L = 6
    Hd, Hm = m.heisenberg_u1(L), ompo.dense_hamiltonian(ompo.heisenberg_u1(L))
    for N in (1, 2, 3):
        errs = []
        for dt in (0.1, 0.05, 0.025):
            W = make_time_mpo(Hd, dt, ("TaylorCluster", N))
            for sl in W.slices:
                ompo.SparseMPOSlice(sl.Os, sl.domcharges, sl.imcharges, sl.pcharges)      # asserts charge conservation
            errs.append(np.abs(dense_mpo(W, L) - sla.expm(-1j * dt * Hm)).max())
        for r in (errs[0] / errs[1], errs[1] / errs[2]):
            assert 0.95 * 2 ** (N + 1) < r < 1.05 * 2 ** (N + 1)
    W1 = make_time_mpo(m.transverse_field_ising(L, 1.3), 0.05, "WI")
    assert W1[0].odim == 2
    Ht = ompo.dense_hamiltonian(ompo.transverse_field_ising(L, 1.3))
    assert np.abs(dense_mpo(make_time_mpo(m.transverse_field_ising(L, 1.3), 0.025, ("TaylorCluster", 3)), L) - sla.expm(-0.025j * Ht)).max() < 2e-7


def test_densempo_times_state_host_part_without_gpu():
    """The host part of mpo * st::InfiniteMPS (densempo.jl:31-45; mpskit_b200.changebonds._product_tensors): the product tensors in
    the sector order of the fused bond spaces, for a charged W^II MPO on a U(1) two-site cell -- every non-zero element conserves
    the fused charges, and the tensors as an (ungauged) MPS have the transfer spectrum of the oracle's product state."""
    import numpy as np
    import mpskit_b200 as m
    from mpskit_b200.operators import make_time_mpo
    from mpskit_b200.changebonds import _product_tensors
    from oracle import mpo as ompo, infinite as oi, changebonds as ocb
    from oracle.spaces import Space
    W = make_time_mpo(m.heisenberg_u1(2), 1e-3)
    Vo = [Space((-2, 0, 2), (2, 3, 2)), Space((-1, 1), (3, 3))]
    st = oi.InfiniteMPS.random(np.random.default_rng(5), Space((-1, 1), (1, 1)), Vo)
    Vd = [m.Space(v.charges, v.dims) for v in Vo]
    tensors, Vf = _product_tensors(W, st.AL, Vd, 0)
    pq = np.array([-1, 1])
    for i in range(2):
        ql, qr = Vf[i].index_charges(), Vf[(i + 1) % 2].index_charges()
        assert tensors[i].shape == (Vf[i].dim, 2, Vf[(i + 1) % 2].dim) and Vf[i].dim == Vo[i].dim * W[i].odim
        bad = (ql[:, None, None] + pq[None, :, None] != qr[None, None, :]) & (np.abs(tensors[i]) > 0)
        assert not bad.any()
    sl = [ompo.SparseMPOSlice(s.Os, s.domcharges, s.imcharges, s.pcharges) for s in W.slices]
    prod = ocb.mpo_times_state(sl, st)
    assert [v.as_dict() for v in prod.V] == [v.as_dict() for v in Vf]

    def leading(T):
        D = T[0].shape[0]
        v = np.eye(D, dtype=complex)
        for _ in range(200):
            for t in T:
                v = np.einsum('ab,apc,bpd->cd', v, t, t.conj())
            lam = np.linalg.norm(v); v = v / lam
        return lam
    # W^II is unitary up to O(dt^2): the product of a normalised state stays normalised to that order
    assert abs(leading(tensors) - 1.0) < 1e-5


def test_decompose_localmps_host_part_without_gpu():
    """test/states.jl:40-47: psi2 = FiniteMPS(decompose_localmps(convert(TensorMap, psi))) has dot(psi, psi2) = dot(psi, psi), for a
    trivial and a U(1) (spin-1) bond space.  The charged SVD chain is the package's host-side state construction
    (mpskit_b200.finitemps.decompose_localmps); the states themselves are evaluated through the oracle."""
    import numpy as np
    import mpskit_b200 as m
    from mpskit_b200.finitemps import decompose_localmps
    from oracle import finite as of_
    from oracle.spaces import Space
    for p, maxV in ((Space((0,), (2,)), Space((0,), (10,))), (Space((-1, 0, 1), (1, 1, 1)), Space((-1, 0, 1), (3, 3, 3)))):
        psi = of_.FiniteMPS.random(np.random.default_rng(3), p, 4, maxV)
        T = None
        for t in of_._site_tensors(psi):
            T = t if T is None else np.tensordot(T, t, axes=([T.ndim - 1], [0]))
        As, Vs = decompose_localmps(T, m.Space(p.charges, p.dims))
        assert [v.as_dict() for v in Vs] == [v.as_dict() for v in psi.V]          # full-rank state: the same bond spaces
        psi2 = of_.FiniteMPS(As, [Space(v.charges, v.dims) for v in Vs], p, normalize=False)
        assert abs(of_.mps_dot(psi, psi2) - of_.mps_dot(psi, psi)) < 1e-12
        T2 = None
        for t in As:
            T2 = t if T2 is None else np.tensordot(T2, t, axes=([T2.ndim - 1], [0]))
        assert np.abs(T2 - T).max() < 1e-13


def test_space_accessors_without_gpu():
    """left_virtualspace / right_virtualspace / physicalspace / max_Ds (abstractmps.jl:141-161, finitemps.jl:297-303; 0-based
    sites) on stand-ins that carry only the bookkeeping of a FiniteMPS / InfiniteMPS."""
    import types
    import mpskit_b200 as m
    p = m.Space((-1, 1), (1, 1))
    Vs = m.finitemps.virtual_spaces(p, 6, m.Space((-2, -1, 0, 1, 2), (1, 2, 3, 2, 1)))

    class Fin:
        V, pspace, L = Vs, p, 6

        def __len__(self):
            return 6
    fin = Fin()
    assert m.left_virtualspace(fin, 2) == Vs[2] and m.right_virtualspace(fin, 2) == Vs[3] and m.physicalspace(fin, 2) is p
    assert m.max_Ds(fin) == [1, 2, 4, 8, 4, 2, 1] and all(v.dim <= D for v, D in zip(Vs, m.max_Ds(fin)))
    inf = types.SimpleNamespace(V=[m.Space((0,), (3,)), m.Space((0,), (5,))], N=2, pspace=p)
    assert m.left_virtualspace(inf, 2).dim == 3 and m.right_virtualspace(inf, 1).dim == 3 and m.right_virtualspace(inf, 0).dim == 5
