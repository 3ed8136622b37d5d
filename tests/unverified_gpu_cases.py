"""GPU cases written after the round's GPU budget was spent: device compositions of verified primitives that have never been
executed on a GPU.  This file is NOT collected by default (its name does not match test_*.py); tests/test_gpu_round2.py runs every
case in its own subprocess with a timeout (test_unverified_device_compositions), so that neither a failure nor a crash or a hang
of an unexecuted composition can take the verified suite down with it.  Run directly:
    python -m pytest tests/unverified_gpu_cases.py -m gpu -q"""
import numpy as np
import pytest
import mpskit_b200 as m
from oracle.spaces import Space as OSpace
from oracle import mpo as ompo, finite as of
from helpers import *

pytestmark = pytest.mark.gpu


def test_haldane_gap_device(ctx):
    """The reference's own quasiparticle assertion on the device (test/algorithms.jl:204-211: energies[1] ~ 0.41047925,
    atol 1e-4, D = 48): spin-1 Heisenberg chain with U(1) blocks, device VUMPS + device infinite quasiparticle ansatz at
    momentum pi, against the reference constant (1e-4) and the oracle's recorded run in the same bond space
    (tests/golden/haldane_gap_oracle.json, 1e-6).  Every primitive on this path is verified by the tests above (U(1) VUMPS,
    U(1) finite quasiparticles, the infinite ansatz on the TFI chain); this combination itself has not been executed on a GPU,
    hence the non-strict xfail marker."""
    import json
    import os
    rec = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "haldane_gap_oracle.json")))
    d48 = [r for r in rec["runs"] if r["D"] == 48][0]
    Ho = ompo.heisenberg_u1(1, d=3)
    Hd = dev_hamiltonian(Ho)
    V = m.Space.from_dict({int(k): v for k, v in d48["bond_space"].items()})
    st = m.InfiniteMPS.random(np.random.default_rng(3), m.Space((-2, 0, 2), (1, 1, 1)), [V], ctx)
    log = []
    st, envs, eps = m.find_groundstate_vumps(st, Hd, m.VUMPS(tol=1e-10, maxiter=400), log=log)
    en, _, _ = m.excitations_qp_infinite(Hd, m.QuasiparticleAnsatz(tol=1e-8, maxiter=20), np.pi, st, envs)     # <= 600 applications
    print("Haldane gap on the device", en[0], d48["gap"], log[-1]["energy"], d48["energy_per_site"], eps)
    assert abs(log[-1]["energy"] - d48["energy_per_site"]) <= 1e-9
    assert abs(en[0] - 0.41047925) <= 1e-4 and abs(en[0] - d48["gap"]) <= 1e-6


def test_time_mpo_wii_approximate_vs_tdvp(ctx):
    """test/algorithms.jl:447-466 with the reference's own time-evolution MPO: W = make_time_mpo(H, dt, WII())
    (timeevmpo.jl:149-207; host-side construction, validated against exp(-i dt H) in tests/test_abi_cpu.py) applied to a U(1)
    two-site InfiniteMPS and approximated with VOMPS and IDMRG2 on the device, against a device TDVP step (the reference asserts
    |dot| = 1 within dt; here 1e-9, the oracle gives 3e-14).  The device algorithms are verified with a W^I MPO by the tests
    above; this charged four-level W^II MPO has not been executed on a GPU, hence the non-strict xfail marker."""
    from mpskit_b200.operators import make_time_mpo
    dt = 1e-3
    Hd = m.heisenberg_u1(2)
    W = make_time_mpo(Hd, dt)
    Vs = [m.Space((-2, 0, 2), (2, 3, 2)), m.Space((-1, 1), (3, 3))]
    st = m.InfiniteMPS.random(np.random.default_rng(5), m.Space((-1, 1), (1, 1)), Vs, ctx)
    a, _, eps = m.approximate(st, (W, st), m.VOMPS(tol=1e-12, maxiter=100))
    b, _, eps2 = m.approximate(st, (W, st), m.IDMRG2(tol=1e-12, maxiter=200, trscheme=m.truncdim(7)))
    t, _ = m.timestep_infinite(st, Hd, 0.0, dt)
    ov_a, ov_b, ov_0 = abs(m.dot_infinite(a, t)), abs(m.dot_infinite(b, t)), abs(m.dot_infinite(st, t))
    print("W^II approximate vs TDVP", eps, eps2, ov_a, ov_b, ov_0)
    assert eps <= 1e-11 and abs(ov_a - 1.0) <= 1e-9 and abs(ov_b - 1.0) <= 1e-9 and abs(ov_0 - 1.0) > 1e-8


def test_fidelity_susceptibility_device(ctx):
    """fidelity_susceptibility(psi, H, [H_X], envs) (fidelity_susceptibility.jl:4-29; test/algorithms.jl:368-396 asserts the
    analytic value |1 / (16 l^2 (l^2 - 1))| of the transverse-field Ising chain within 1e-2): infinite D = 16 states at l = 2 and
    l = 4 against the analytic value (1e-8; the oracle gives 3e-13 / 2e-17) and a finite chain against the oracle (1e-8).  Built
    from the quasiparticle effective Hamiltonian verified above plus the host-driven GMRES; this combination has not been
    executed on a GPU, hence the non-strict xfail marker."""
    from oracle import quasiparticle as oq
    X = np.array([[0, 1], [1, 0]], dtype=complex)
    for lam in (2.0, 4.0):
        Hd = dev_hamiltonian(ompo.transverse_field_ising(1, lam))
        Vd = dev_hamiltonian(ompo.fsm_hamiltonian(1, 2, {(0, 0): 1.0, (1, 1): 1.0, (0, 1): X}, (0, 0), (0, 0)))
        st = m.InfiniteMPS.random(np.random.default_rng(3), m.Space((0,), (2,)), [m.Space((0,), (16,))], ctx)
        st, envs, _ = m.find_groundstate_vumps(st, Hd, m.VUMPS(tol=1e-11, maxiter=300))
        chi = m.fidelity_susceptibility(st, Hd, [Vd], envs)
        exact = abs(1 / (16 * lam ** 2 * (lam ** 2 - 1)))
        print("fidelity susceptibility", lam, chi[0, 0], exact)
        assert abs(chi[0, 0] - exact) <= 1e-8
    L, lam = 8, 2.0
    Ho = ompo.transverse_field_ising(L, lam)
    Vo = ompo.fsm_hamiltonian(L, 2, {(0, 0): 1.0, (1, 1): 1.0, (0, 1): X}, (0, 0), (0, 0))
    psi_o = of.FiniteMPS.random(np.random.default_rng(1), OSpace((0,), (2,)), L, OSpace((0,), (16,)))
    psi_o, envs_o, _ = of.dmrg2_sweeps(psi_o, Ho, nsweeps=8, trunc=("truncdim", 16), tol=1e-13)
    chi_o = oq.fidelity_susceptibility(psi_o, Ho, [Vo], envs_o)
    psi_d = m.FiniteMPS.random(np.random.default_rng(1), m.Space((0,), (2,)), L, m.Space((0,), (16,)), ctx)
    psi_d, envs_d, _ = m.find_groundstate(psi_d, dev_hamiltonian(Ho), m.DMRG2(tol=1e-13, maxiter=8, eigalg=m.Lanczos(tol=1e-13), trscheme=m.truncdim(16)))
    chi_d = m.fidelity_susceptibility(psi_d, dev_hamiltonian(Ho), [dev_hamiltonian(Vo)], envs_d)
    print("finite fidelity susceptibility", chi_d[0, 0], chi_o[0, 0])
    assert abs(chi_d[0, 0] - chi_o[0, 0]) <= 1e-8


def test_find_groundstate_keyword_form_and_union(ctx):
    """find_groundstate(psi, H; tol, trscheme) (find_groundstate.jl:19-41) and alg1 & alg2 (unionalg.jl): the README quickstart
    form on a finite chain (DMRG2 & DMRG) against exact diagonalisation, and on an InfiniteMPS (IDMRG2 & VUMPS) against the exact
    TFI energy density.  Pure dispatch over the solvers verified above (the dispatch itself is tested on the CPU)."""
    import scipy.integrate as si
    L = 8
    Ho = ompo.transverse_field_ising(L, 1.5)
    psi = m.FiniteMPS.random(np.random.default_rng(1), m.Space((0,), (2,)), L, m.Space((0,), (4,)), ctx)
    Hd = dev_hamiltonian(Ho)
    psi, envs, eps = m.find_groundstate(psi, Hd, tol=1e-10, trscheme=m.truncdim(16))
    E = float(np.sum(m.expectation_value(psi, Hd, envs)).real)
    assert abs(E - np.linalg.eigvalsh(ompo.dense_hamiltonian(Ho))[0]) <= 1e-9
    g = 0.5
    H2 = dev_hamiltonian(ompo.transverse_field_ising(2, g))
    st = m.InfiniteMPS.random(np.random.default_rng(3), m.Space((0,), (2,)), [m.Space((0,), (4,))] * 2, ctx)
    st, ienvs, eps = m.find_groundstate(st, H2, m.IDMRG2(tol=1e-4, maxiter=50, trscheme=m.truncdim(12)) & m.VUMPS(tol=1e-11, maxiter=100))
    from mpskit_b200 import infinite as di
    e = float(np.mean(di.expectation_value(st, H2, ienvs)).real)
    exact = -si.quad(lambda k: np.sqrt(1 + g * g + 2 * g * np.cos(k)), 0, np.pi)[0] / np.pi
    print("keyword form / union", E, e, exact, st.V[0].dim)
    assert st.V[0].dim == 12 and abs(e - exact) <= 1e-9


def test_exact_diagonalization_device(ctx):
    """exact_diagonalization(H; sector, num) (ED.jl:4-53): a full-bond-dimension FiniteMPS whose middle site spans the symmetry
    sector; the two lowest eigenvalues of the single-site effective Hamiltonian against dense diagonalisation (TFI chain, and
    the S_z = 0 and S_z = 1 sectors of the U(1) Heisenberg chain)."""
    L = 8
    Ho = ompo.transverse_field_ising(L, 1.5)
    vals, states, infos = m.exact_diagonalization(dev_hamiltonian(Ho), m.Space((0,), (2,)), num=2, ctx=ctx)
    ev = np.linalg.eigvalsh(ompo.dense_hamiltonian(Ho))
    assert abs(vals[0].real - ev[0]) <= 1e-10 and abs(vals[1].real - ev[1]) <= 1e-9
    Hh = ompo.heisenberg_u1(L)
    Hm = ompo.dense_hamiltonian(Hh)
    for sector, nup in ((0, L // 2), (2, L // 2 + 1)):
        keep = [k for k in range(2 ** L) if bin(k).count("1") == nup]
        evs = np.linalg.eigvalsh(Hm[np.ix_(keep, keep)])
        vals, states, infos = m.exact_diagonalization(dev_hamiltonian(Hh), m.Space((-1, 1), (1, 1)), sector=sector, num=1, ctx=ctx)
        print("ED sector", sector, vals[0], evs[0])
        # the sign convention of the sector label (S_z = +1 or -1) does not matter for the Heisenberg chain
        assert abs(vals[0].real - evs[0]) <= 1e-10


def test_reference_printed_su2_state_readouts_device(ctx):
    """docs/src/examples/quantum1d/5.haldane-spt/index.md:397-404,1018 prints E_minus = -1.4014839739630869 and
    S_minus + log(2) = 1.5486227235451633 for the SU(2)-symmetric VUMPS state in V_minus = SU2Space(1/2 => 10, 3/2 => 5, 5/2 => 3).
    The oracle reproduces both (tests/golden/make_su2_pins.py; 4e-16 / 2e-12) and recorded the converged left-gauged tensor in U(1)
    blocks (D = 58).  Here that tensor goes through the device's own path: upload, uniform gauge (device QR / Arnoldi), the
    MPOHamInfEnv GMRES solves, expectation_value, entropy, calc_galerkin -- against the numbers the reference printed.  (A device
    VUMPS iteration from this state would leave the SU(2) manifold, as the unprojected oracle does, hence read-outs only.)  The same
    for the two-site SU(2) cell of 4.xxz-heisenberg/index.md:579-671 (D = 65 / 52): "VUMPS conv 73: obj = -8.862878980874e-01"."""
    import os
    from mpskit_b200 import infinite as di
    from mpskit_b200.tensor import DeviceTensor
    rec = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "su2_haldane_spt_minus_state.npz"))
    V = m.Space(tuple(int(c) for c in rec["charges"]), tuple(int(d) for d in rec["dims"]))
    p = m.Space((-2, 0, 2), (1, 1, 1))
    A = DeviceTensor.from_dense(ctx, rec["AL0"], V, [-2, 0, 2], V, np.complex128)
    st = m.InfiniteMPS.from_tensors([A], [V], p, ctx)
    Hd = dev_hamiltonian(ompo.heisenberg_u1(1, d=3))
    envs = di.MPOHamInfEnv(st, Hd)
    e = float(di.expectation_value(st, Hd, envs)[0].real)
    S = m.entropy(st, 0)
    gal = di.calc_galerkin(st, 0, envs)
    print("device read-outs of the recorded SU(2) state", repr(e), repr(S + np.log(2.0)), gal)
    assert abs(e - (-1.4014839739630869)) <= 1e-11
    assert abs(S + np.log(2.0) - 1.5486227235451633) <= 1e-8
    assert gal <= 1e-8
    # 4.xxz-heisenberg/index.md:671: "VUMPS conv 73: obj = -8.862878980874e-01" (two-site SU(2) cell, sum over the cell)
    rec = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "su2_xxz_state.npz"))
    Vs = [m.Space(tuple(int(c) for c in rec[f"charges{i}"]), tuple(int(d) for d in rec[f"dims{i}"])) for i in range(2)]
    A = [DeviceTensor.from_dense(ctx, rec[f"AL{i}"], Vs[i], [-1, 1], Vs[(i + 1) % 2], np.complex128) for i in range(2)]
    st = m.InfiniteMPS.from_tensors(A, Vs, m.Space((-1, 1), (1, 1)), ctx)
    Hx = dev_hamiltonian(ompo.heisenberg_u1(2))
    envs = di.MPOHamInfEnv(st, Hx)
    ex = float(np.sum(di.expectation_value(st, Hx, envs)).real)
    gx = max(di.calc_galerkin(st, i, envs) for i in range(2))
    print("device read-outs of the recorded xxz SU(2) state", repr(ex), gx)
    assert abs(ex - (-8.862878980874e-01)) <= 1e-11 and gx <= 1e-8


def test_reference_printed_su2_numbers_device_vumps(ctx):
    """The device VUMPS against a number the REFERENCE printed: docs/src/examples/quantum1d/5.haldane-spt/index.md:397-404,1018,
    E_minus = -1.4014839739630869 and S_minus + log(2) = 1.5486227235451633 for V_minus = SU2Space(1/2 => 10, 3/2 => 5, 5/2 => 3).
    As in tests/golden/make_su2_pins.py (where the oracle reproduces both to 4e-16 / 2e-12): the SU(2)-invariant start tensor in
    U(1) blocks, and VUMPS restricted to the invariant subspace -- the loop of mpskit_b200.infinite.find_groundstate_vumps
    (vumps.jl:29-92) with every application of the device ddAC / ddC operators followed by the orthogonal projector (download,
    numpy projector, upload), so that the eigensolver is the host-driven Lanczos over device vectors that the lazy-sum tests verify.
    Gauge fixing (device QR / Arnoldi), the MPOHamInfEnv GMRES solves, regauge, expectation_value and entropy are the product
    path unchanged.  Energy 1e-11, entropy 1e-8."""
    import importlib.util
    import os
    from fractions import Fraction
    from mpskit_b200 import infinite as di
    from mpskit_b200.tensor import DeviceTensor
    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location("make_su2_pins", os.path.join(here, "golden", "make_su2_pins.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    h = Fraction(1, 2)
    cell = {h: 10, 3 * h: 5, 5 * h: 3}
    A0, Vo, _ = mod.symmetric_tensor(np.random.default_rng(7), cell, 1, cell)
    pa, pc = mod.tensor_projector(cell, 1, cell), mod.matrix_projector(cell)
    V = m.Space(tuple(Vo.charges), tuple(Vo.dims))
    p = m.Space((-2, 0, 2), (1, 1, 1))

    def project(x):
        dense = x.to_dense()
        dense = pa(dense) if len(x.mch) == 3 else pc(dense[:, 0, :])[:, None, :]
        return DeviceTensor.from_dense(ctx, dense, x.Vl, list(x.mch), x.Vr, x.dtype)

    class Projected:
        def __init__(self, op):
            self.op = op

        def __call__(self, x):
            y = self.op(x)
            z = project(y)
            y.free()
            return z

    st = m.InfiniteMPS.from_tensors([DeviceTensor.from_dense(ctx, A0, V, [-2, 0, 2], V, np.complex128)], [V], p, ctx)
    Hd = dev_hamiltonian(ompo.heisenberg_u1(1, d=3))
    envs = di.MPOHamInfEnv(st, Hd)
    eps = di.calc_galerkin(st, 0, envs)
    for it in range(1, 101):
        etol = di.dyn_tol(eps, it, 1e-12, 1e-5, 1e-5)
        eig = m.Lanczos(tol=etol, krylovdim=30, maxiter=100)
        hac = di._inf_ddAC(0, st, Hd, envs)
        _, ac = di.fixedpoint(Projected(hac), project(st.AC[0]), "SR", eig); hac.free()
        hc = di._inf_ddC(0, st, Hd, envs)
        _, c = di.fixedpoint(Projected(hc), project(st.CR[0]), "SR", eig); hc.free()
        newAL = [di.regauge(ac, c)]
        st = m.InfiniteMPS.from_AL(newAL, st.CR[0], st.V, st.pspace, st.ctx, di.dyn_tol(eps, it, di.TOLGAUGE, 1e-5, 1e-8), 100)
        envs.recalculate(st, di.dyn_tol(eps, it, 1e-12, 1e-5, 1e-5))
        eps = di.calc_galerkin(st, 0, envs)
        if eps <= 1e-9:
            break
    e = float(di.expectation_value(st, Hd, envs)[0].real)
    S = m.entropy(st, 0)
    print("device SU(2)-restricted VUMPS", it, eps, repr(e), repr(S + np.log(2.0)))
    assert eps <= 1e-9
    assert abs(e - (-1.4014839739630869)) <= 1e-11
    assert abs(S + np.log(2.0) - 1.5486227235451633) <= 1e-8


def test_changebonds_multiline_device(ctx):
    """changebonds on MPSMultiline on the device (test/algorithms.jl:318-338; optimalexpand.jl:40-66, randexpand.jl:36-38,
    svdcut.jl:32-34) against the oracle restatement (oracle/changebonds.py *_multiline): a random dense MPO on a one-row multiline
    through the InfiniteMPS / DenseMPO form, and two classical-Ising rows where row i is expanded with the two-site tensor of row
    i - 1: added dimensions bit-exact, singular values of the projected two-site tensors 1e-9, overlaps with the old rows 1 (the
    reference asserts 1e-8; here 1e-10), SvdCut back to the old dimension.  Composition of verified device primitives
    (MultilineEnv, MPO_ddAC2 on dense slices, the expansion step of the single-row changebonds test)."""
    import importlib
    from oracle import infinite as oi, statmech as osm, changebonds as ocb
    dcb = importlib.import_module(m.__name__ + ".changebonds")
    rng = np.random.default_rng(11)
    o = rng.standard_normal((2, 2, 2, 2)) + 1j * rng.standard_normal((2, 2, 2, 2))
    p_o, D_o = OSpace((0,), (2,)), OSpace((0,), (3,))
    st_o = oi.InfiniteMPS.random(np.random.default_rng(2), p_o, [D_o])
    st_d = m.InfiniteMPS.random(np.random.default_rng(2), m.Space((0,), (2,)), [m.Space((0,), (3,))], ctx)
    log_o, log_d = [], []
    oe_o, _ = ocb.optimal_expand_multiline([st_o], [[ompo.dense_mpo_slice(o)]], ("truncdim", 9), None, log_o)
    oe_d, _ = m.changebonds(st_d, m.DenseMPO([o]), m.OptimalExpand(trscheme=m.truncdim(9)), log=log_d)
    assert oe_d.V[0].as_dict() == oe_o[0].V[0].as_dict() == {0: 6}
    for q, s in log_o[0]["spectrum"].items():
        assert np.abs(np.asarray(log_d[0]["spectrum"][q]) - s).max() <= 1e-9 * max(1.0, np.abs(s).max())
    assert abs(abs(dcb.dot_infinite(st_d, oe_d)) - 1.0) <= 1e-8                 # test/algorithms.jl:333
    cut = m.changebonds(oe_d, m.SvdCut(trscheme=m.truncdim(3)))
    assert cut.V[0].dim == 3 < oe_d.V[0].dim                                     # test/algorithms.jl:337-338
    # two rows
    b1, b2 = 0.35, 0.30
    so = [[ompo.dense_mpo_slice(ompo.classical_ising(b1))], [ompo.dense_mpo_slice(ompo.classical_ising(b2))]]
    sd = [[m.dense_mpo_slice(m.classical_ising(b1))], [m.dense_mpo_slice(m.classical_ising(b2))]]
    rows_o = [oi.InfiniteMPS.random(np.random.default_rng(3 + r), p_o, [OSpace((0,), (4,))]) for r in range(2)]
    rows_d = [m.InfiniteMPS.random(np.random.default_rng(3 + r), m.Space((0,), (2,)), [m.Space((0,), (4,))], ctx) for r in range(2)]
    rows_o, envs_o, _ = osm.leading_boundary_multiline(rows_o, so, tol=1e-11, maxiter=80)
    rows_d, envs_d, _ = m.leading_boundary_multiline(rows_d, sd, m.VUMPS(tol=1e-11, maxiter=80))
    log_o, log_d = [], []
    new_o, _ = ocb.optimal_expand_multiline(rows_o, so, ("truncdim", 2), envs_o, log_o)
    new_d, _ = m.changebonds_multiline(rows_d, sd, m.OptimalExpand(trscheme=m.truncdim(2)), envs_d, log=log_d)
    for r in range(2):
        assert new_d[r].V[0].as_dict() == new_o[r].V[0].as_dict() == {0: 6}
        for q, s in log_o[r]["spectrum"].items():
            assert np.abs(np.asarray(log_d[r]["spectrum"][q]) - s).max() <= 1e-8 * np.abs(s).max()
        assert abs(abs(dcb.dot_infinite(rows_d[r], new_d[r])) - 1.0) <= 1e-10
    back = m.changebonds_multiline(new_d, m.SvdCut(trscheme=m.truncdim(4)))
    assert all(b.V[0].dim == 4 for b in back)
    assert all(abs(abs(dcb.dot_infinite(rows_d[r], back[r])) - 1.0) <= 1e-10 for r in range(2))


def test_densempo_times_state_and_svdcut_device(ctx):
    """test/algorithms.jl:447-471 on the device: st6 = changebonds(W1 * st, SvdCut(truncdim(10))) with the TaylorCluster{3} MPO of
    the TFI chain (g = 4, two-site cell, D = 10) against a device TDVP step (`abs(dot(st6, st5)) ~ 1 atol dt`; here 1e-3 dt) and
    the oracle's product state (overlap 1 to 1e-9); nW1 = changebonds(W1, SvdCut(truncerr(dt))) is a trivial MPO (level dimension
    1); and the charged W^II MPO of the U(1) Heisenberg chain times a U(1) state (cut back to its own bond spaces, overlap with
    TDVP 1 to 1e-7).  The product tensors are formed on the host (state construction, CPU-tested); gauge fixing, SvdCut, TDVP and
    dot are the verified device paths."""
    import importlib
    from mpskit_b200.operators import make_time_mpo
    dcb = importlib.import_module(m.__name__ + ".changebonds")
    dt = 1e-3
    Hd = m.transverse_field_ising(2, 4.0)
    W1 = m.DenseMPO(make_time_mpo(Hd, dt, ("TaylorCluster", 3)).slices)
    st = m.InfiniteMPS.random(np.random.default_rng(5), m.Space((0,), (2,)), [m.Space((0,), (10,))] * 2, ctx)
    prod = W1 * st
    assert prod.V[0].dim == 10 * W1[0].odim
    st6 = m.changebonds(prod, m.SvdCut(trscheme=m.truncdim(10)))
    st5, _ = m.timestep_infinite(st, Hd, 0.0, dt)
    d65, d05 = abs(abs(dcb.dot_infinite(st6, st5)) - 1.0), abs(abs(dcb.dot_infinite(st, st5)) - 1.0)
    print("W1 * st, SvdCut vs TDVP", d65, d05)
    assert st6.V[0].dim == 10 and d65 <= 1e-3 * dt and d65 <= 0.05 * d05
    nW1 = m.changebonds(W1, m.SvdCut(trscheme=m.truncerr(dt)))
    assert nW1[0].odim == 1                                                    # test/algorithms.jl:470-471
    Hh = m.heisenberg_u1(2)
    W2 = m.DenseMPO(make_time_mpo(Hh, dt).slices)
    Vs = [m.Space((-2, 0, 2), (2, 4, 2)), m.Space((-1, 1), (4, 4))]
    su = m.InfiniteMPS.random(np.random.default_rng(5), m.Space((-1, 1), (1, 1)), Vs, ctx)
    cut = m.changebonds(W2 * su, m.SvdCut(trscheme=m.truncdim(8)))
    tu, _ = m.timestep_infinite(su, Hh, 0.0, dt)
    du = abs(abs(dcb.dot_infinite(cut, tu)) - 1.0)
    print("U(1) W^II * st, SvdCut vs TDVP", du, [v.as_dict() for v in cut.V])
    assert [v.as_dict() for v in cut.V] == [v.as_dict() for v in Vs] and du <= 1e-7


def test_densempo_products_device(ctx):
    """test/operators.jl:159-171 ("DenseMPO") on the device: W = convert(DenseMPO, make_time_mpo(ham, 1im * 0.5, WII())) for the
    TFI and the spin-1 Heisenberg chain (trivial and U(1)), psi = InfiniteMPS([P], [oneunit (+) P]):
    abs(dot(W * (W * psi), (W * W) * psi)) ~ 1 atol 1e-10 with the device constructor (uniform gauge) and the device dot."""
    import importlib
    from mpskit_b200.operators import make_time_mpo
    dcb = importlib.import_module(m.__name__ + ".changebonds")
    cases = [(m.transverse_field_ising(1, 1.0), m.Space((0,), (2,)), m.Space((0,), (3,))),
             (m.drop_symmetry(m.heisenberg_u1(1, d=3)), m.Space((0,), (3,)), m.Space((0,), (4,))),
             (m.heisenberg_u1(1, d=3), m.Space((-2, 0, 2), (1, 1, 1)), m.Space((-2, 0, 2), (1, 2, 1)))]
    for Hd, p, V in cases:
        W = m.DenseMPO(make_time_mpo(Hd, 0.5j).slices)
        psi = m.InfiniteMPS.random(np.random.default_rng(2), p, [V], ctx)
        a = W * (W * psi)
        b = (W * W) * psi
        ov = abs(dcb.dot_infinite(a, b))
        print("DenseMPO products", a.V[0].as_dict(), b.V[0].as_dict(), ov)
        assert abs(ov - 1.0) <= 1e-10 and abs(abs(dcb.dot_infinite(a, psi)) - 1.0) > 1e-3


def test_finite_mps_vector_operations_device(ctx):
    """test/states.jl:15-38 ("FiniteMPS") on the device: dot(psi, psi) = norm(psi.AC[1])^2, norm(3 psi)^2 = 9 norm(psi)^2,
    norm(2 psi + psi - 3 psi) ~ 0 (sqrt(eps)), and dot / + / scalar * against the oracle's dense state vectors (1e-12); trivial
    and U(1) bond spaces.  dot runs through the rectangular overlap environments of `approximate`, + through mpsk_tensor_cat and
    the device QR / LQ (verified primitives; this composition has not been executed on a GPU)."""
    from oracle import finite as of_
    for p_o, maxV_o, L in ((OSpace((0,), (2,)), OSpace((0,), (10,)), 7),
                           (OSpace((-1, 1), (1, 1)), OSpace((-2, -1, 0, 1, 2), (2, 3, 4, 3, 2)), 6)):
        p_d, maxV_d = to_dev_space(p_o), to_dev_space(maxV_o)
        psi_o = of_.FiniteMPS.random(np.random.default_rng(12), p_o, L, maxV_o)
        phi_o = of_.FiniteMPS.random(np.random.default_rng(13), p_o, L, maxV_o)
        psi = m.FiniteMPS.random(np.random.default_rng(12), p_d, L, maxV_d, ctx)
        phi = m.FiniteMPS.random(np.random.default_rng(13), p_d, L, maxV_d, ctx)
        ovl = psi.dot(psi)
        assert abs(ovl - psi.norm() ** 2) <= 1e-12 and abs(ovl - 1.0) <= 1e-12
        assert abs(psi.dot(phi) - of_.mps_dot(psi_o, phi_o)) <= 1e-12
        p3 = 3 * (psi * 3)
        assert abs(p3.norm() ** 2 - 81 * ovl.real) <= 1e-10
        zero = 2 * psi + psi - 3 * psi
        assert zero.norm() <= 1e-7
        c = 0.5 - 0.25j
        s_d, s_o = psi + c * phi, of_.mps_add(psi_o, of_.mps_scale(phi_o, c))
        assert abs(s_d.norm() - of_.mps_norm(s_o)) <= 1e-12
        assert abs(s_d.dot(psi) - of_.mps_dot(s_o, psi_o)) <= 1e-12 and abs(s_d.dot(phi) - of_.mps_dot(s_o, phi_o)) <= 1e-12
        assert [v.dim for v in s_d.V] == [v.dim for v in s_o.V]
        assert abs(s_d.normalize().norm() - 1.0) <= 1e-13
        # test/states.jl:40-47: psi2 = FiniteMPS(decompose_localmps(convert(TensorMap, psi))) -> dot(psi, psi2) = dot(psi, psi)
        small = m.FiniteMPS.random(np.random.default_rng(3), p_d, 4, maxV_d, ctx)
        again = m.FiniteMPS.from_dense(small.to_dense(), p_d, ctx)
        assert abs(small.dot(again) - small.dot(small)) <= 1e-12


def test_multiline_power_method_device(ctx):
    """leading_boundary(psi::MPSMultiline, O::MPOMultiline, VOMPS()) (statmech/vomps.jl:27-87) on the device: the power method
    on two classical-Ising rows (one application of the verified multi-row operator per column instead of its eigensolve)
    reaches the value of the device multi-row VUMPS and of the oracle's power method (product of the row values, 1e-9)."""
    from oracle import infinite as oi, statmech as osm
    b1, b2 = 0.35, 0.30
    so = [[ompo.dense_mpo_slice(ompo.classical_ising(b1))], [ompo.dense_mpo_slice(ompo.classical_ising(b2))]]
    sd = [[m.dense_mpo_slice(m.classical_ising(b1))], [m.dense_mpo_slice(m.classical_ising(b2))]]
    rows_o = [oi.InfiniteMPS.random(np.random.default_rng(3 + r), OSpace((0,), (2,)), [OSpace((0,), (4,))]) for r in range(2)]
    mk = lambda: [m.InfiniteMPS.random(np.random.default_rng(3 + r), m.Space((0,), (2,)), [m.Space((0,), (4,))], ctx) for r in range(2)]
    lo, la, lb = [], [], []
    osm.leading_boundary_multiline_vomps(rows_o, so, tol=1e-10, maxiter=400, log=lo)
    m.leading_boundary_multiline(mk(), sd, m.VUMPS(tol=1e-11, maxiter=80), log=la)
    _, _, eps = m.leading_boundary_multiline(mk(), sd, m.VOMPS(tol=1e-10, maxiter=400), log=lb)
    va, vb, vo = np.prod(la[-1]["value"]), np.prod(lb[-1]["value"]), np.prod(lo[-1]["value"])
    print("multi-row power method", va, vb, vo, eps, len(la), len(lb), len(lo))
    assert eps <= 1e-10 and abs(va - vb) <= 1e-9 * abs(va) and abs(vo - vb) <= 1e-9 * abs(vo)


def test_multiline_approximate_device(ctx):
    """approximate(psi::MPSMultiline, (O::MPOMultiline, psi0::MPSMultiline), VOMPS()) (approximate/vomps.jl:19-80) on the device:
    two rows with W^I MPOs of dt and 2 dt on two different states; row r + 1 of the result equals a device TDVP step of psi0[r]
    with the matching time step (overlap 1 to 1e-8) and the oracle's multi-row result (1e-8).  Row-wise map of the verified
    single-row approximate_vomps."""
    import importlib
    from mpskit_b200.operators import make_time_mpo
    from oracle import infinite as oi, statmech as osm
    dcb = importlib.import_module(m.__name__ + ".changebonds")
    dt = 1e-3
    Hd = m.transverse_field_ising(1, 2.0)
    p, V = m.Space((0,), (2,)), m.Space((0,), (6,))
    psi0 = [m.InfiniteMPS.random(np.random.default_rng(20 + r), p, [V], ctx) for r in range(2)]
    start = [m.InfiniteMPS.random(np.random.default_rng(30 + r), p, [V], ctx) for r in range(2)]
    Ws = [m.DenseMPO(make_time_mpo(Hd, dt * (r + 1), "WI").slices) for r in range(2)]
    out, _, eps = m.approximate_vomps_multiline(start, ([[W[0]] for W in Ws], psi0), m.VOMPS(tol=1e-12, maxiter=200))
    assert eps <= 1e-11
    for r in range(2):
        ref, _ = m.timestep_infinite(psi0[r], Hd, 0.0, dt * (r + 1))
        ov = abs(dcb.dot_infinite(out[(r + 1) % 2], ref))
        print("multiline approximate", r, ov, eps)
        assert abs(ov - 1.0) <= 1e-8 and abs(abs(dcb.dot_infinite(out[r], ref)) - 1.0) > 1e-3


def test_infinite_variance_device(ctx):
    """variance(state::InfiniteMPS, H, envs) (toolbox.jl:136-139) on the device: rescaled_H = H - <H> (host-side operator
    arithmetic), the product automaton rescaled_H * rescaled_H with identity levels in its middle through the device MPOHamInfEnv
    (GMRES with the regularised transfer operators).  A sigma_z product state of H = -sum ZZ - g sum X has variance density g^2
    exactly; VUMPS ground states have a small non-negative one that shrinks with the bond dimension; both against the oracle."""
    from oracle import infinite as oi
    g = 0.7
    Hd, Ho = m.transverse_field_ising(1, g), ompo.transverse_field_ising(1, g)
    from mpskit_b200.tensor import DeviceTensor
    A = np.zeros((1, 2, 1), dtype=complex)
    A[0, 1, 0] = 1.0
    V1, p = m.Space((0,), (1,)), m.Space((0,), (2,))
    up = m.InfiniteMPS.from_tensors([DeviceTensor.from_dense(ctx, A, V1, [0, 0], V1, np.complex128)], [V1], p, ctx)
    v_up = m.variance(up, Hd)
    assert abs(v_up - g ** 2) <= 1e-10
    vs = []
    for D in (4, 12):
        st = m.InfiniteMPS.random(np.random.default_rng(3), p, [m.Space((0,), (D,))], ctx)
        st, envs, _ = m.find_groundstate_vumps(st, Hd, m.VUMPS(tol=1e-11, maxiter=200))
        vs.append(m.variance(st, Hd, envs))
    print("infinite variance", v_up, vs)
    assert -1e-10 < vs[1] < vs[0] < 1e-3
