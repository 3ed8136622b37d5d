"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.
Tolerances: matvec outputs <= 1e-12 relative to max|y|; energies 1e-10 relative; singular
values / truncation errors 1e-8; sector and block bookkeeping bit-exact."""
import numpy as np
import pytest
import mpskit_b200 as m
from oracle.spaces import Space as OSpace
from oracle.blocks import BlockTensor
from oracle import mpo as ompo, contract as oc, linalg as ol, finite as of, krylov as ok
from helpers import *

pytestmark = pytest.mark.gpu

VL = OSpace((-2, 0, 2), (3, 5, 2))
VR = OSpace((-2, 0, 2, 4), (2, 6, 3, 1))
VBIG_L = OSpace((-3, -1, 1, 3), (37, 150, 141, 29))
VBIG_R = OSpace((-4, -2, 0, 2, 4), (19, 133, 170, 90, 11))
VBIG_R2 = OSpace((-5, -3, -1, 1, 3), (17, 140, 163, 97, 13))   # same parity as VBIG_L: two-site tensors


def _ctxs(ctx, dbg_ctx):
    return [("dmma", ctx), ("debug", dbg_ctx)]


@pytest.mark.parametrize("dtype", [np.complex128, np.float64])
def test_roundtrip_and_vector_interface(ctx, dtype):
    rng = np.random.default_rng(3)
    a = BlockTensor.random(rng, VL, [1, -1, 1], VR, dtype)
    b = BlockTensor.random(rng, VL, [1, -1, 1], VR, dtype)
    da, db = dev_tensor(ctx, a), dev_tensor(ctx, b)
    assert m.block_table(da.Vl, da.mch, da.Vr) == a.keys3()          # bit-exact bookkeeping
    back = da.to_blocks()
    for k, v in a.blocks.items():
        assert np.array_equal(back[k], v)
    ref = sum(np.vdot(a.blocks[k], b.blocks[k]) for k in a.blocks)
    assert abs(da.dot(db) - ref) <= 1e-13 * abs(ref)
    assert abs(da.norm() - a.norm()) <= 1e-13 * a.norm()
    coef = 0.3 - 0.7j if np.dtype(dtype).kind == "c" else 0.3
    db.axpy(coef, da)
    db.scale(2.0)
    out = db.to_blocks()
    for k in a.blocks:
        assert relerr(out[k], 2.0 * (b.blocks[k] + coef * a.blocks[k])) < 1e-14
    # fortran-ordered / strided host blocks (TensorKit column-major sub-blocks)
    blocks_f = {k: np.asfortranarray(v) for k, v in a.blocks.items()}
    df = m.DeviceTensor.from_blocks(ctx, da.Vl, da.mch, da.Vr, blocks_f, dtype)
    assert np.array_equal(df.to_dense(), a.to_dense())


@pytest.mark.parametrize("which", ["dmma", "debug"])
@pytest.mark.parametrize("dtype", [np.complex128, np.float64])
@pytest.mark.parametrize("size", ["small", "big"])
def test_dAC2_matvec(ctx, dbg_ctx, which, dtype, size):
    c = ctx if which == "dmma" else dbg_ctx
    H = ompo.heisenberg_longrange_u1(4, r=2)
    h = H[0]
    Vl, Vr = (VL, VR) if size == "small" else (VBIG_L, VBIG_R2)
    x, GL, GR = random_problem(11, h, Vl, Vr, "ac2", dtype)
    d = h.d
    # small: the reference's entry loop on dense arrays; big: its block-sparse twin (same order, fast on the CPU)
    yref = x2_from_dense(oc.dAC2_sparse(x2_dense(x, d), h, h, GL.to_dense(), GR.to_dense()), d) if size == "small" \
        else oc.dAC2_blocks(x, h, h, GL, GR).to_dense()
    hs = dev_slice(h)
    dx, dGL, dGR = dev_tensor(c, x), dev_tensor(c, GL), dev_tensor(c, GR)
    op = m.MPO_ddAC2(c, hs, hs, dGL, dGR, dx)
    y = op(dx).to_dense()
    err = relerr(y, yref)
    print("dAC2", which, dtype.__name__, size, "relerr", err, "flops", op.flops)
    assert err <= 1e-12 and op.flops > 0 and np.abs(y).max() > 0
    # second application with a different vector through the same plan
    x2 = BlockTensor.random(np.random.default_rng(5), Vl, x.mch, Vr, dtype)
    y2 = op(dev_tensor(c, x2)).to_dense()
    y2ref = oc.dAC2_blocks(x2, h, h, GL, GR).to_dense()
    assert relerr(y2, y2ref) <= 1e-12


@pytest.mark.parametrize("dtype", [np.complex128, np.float64])
def test_dAC_and_dC(ctx, dtype):
    H = ompo.heisenberg_u1(4)
    h = H[0]
    x, GL, GR = random_problem(12, h, VBIG_L, VBIG_R, "ac", dtype)
    yref = oc.dAC_sparse(x.to_dense(), h, GL.to_dense(), GR.to_dense())
    hs = dev_slice(h)
    dx, dGL, dGR = dev_tensor(ctx, x), dev_tensor(ctx, GL), dev_tensor(ctx, GR)
    y = m.MPO_ddAC(ctx, hs, dGL, dGR, dx)(dx).to_dense()
    assert relerr(y, yref) <= 1e-12
    # dC: GL at the bond, GR at the bond (derivatives.jl:171-193)
    rng = np.random.default_rng(13)
    V = VBIG_R
    Cm = BlockTensor.random(rng, V, [0], V, dtype)
    GLc = BlockTensor.random(rng, V, [-c for c in h.cl], V, dtype)
    GRc = BlockTensor.random(rng, V, [c for c in h.cr], V, dtype)
    cref = oc.dC_dense(Cm.to_dense()[:, 0, :], GLc.to_dense(), GRc.to_dense())
    dC = dev_tensor(ctx, Cm)
    yc = m.MPO_ddC(ctx, dev_tensor(ctx, GLc), dev_tensor(ctx, GRc), dC)(dC).to_dense()[:, 0, :]
    assert relerr(yc, cref) <= 1e-12


def test_trivial_symmetry_dense_path(ctx):
    """no symmetry: one sector per bond -> plain dense contractions (config 1 / 2 style), TFI slice."""
    H = ompo.transverse_field_ising(4, g=0.7)
    h = H[0]
    Vl, Vr = OSpace((0,), (37,)), OSpace((0,), (45,))
    x, GL, GR = random_problem(14, h, Vl, Vr, "ac2")
    yref = oc.dAC2_dense(x2_dense(x, 2), h.dense(), h.dense(), GL.to_dense(), GR.to_dense())
    hs = dev_slice(h)
    dx = dev_tensor(ctx, x)
    y = m.MPO_ddAC2(ctx, hs, hs, dev_tensor(ctx, GL), dev_tensor(ctx, GR), dx)(dx).to_dense()
    assert relerr(y, x2_from_dense(yref, 2)) <= 1e-12


@pytest.mark.parametrize("dtype", [np.complex128, np.float64])
def test_transfers(ctx, dtype):
    """transfer_left / transfer_right over a SparseMPOSlice (transfer.jl:166-259) incl. an empty column."""
    import ctypes as C
    H = ompo.heisenberg_longrange_u1(4, r=2)
    h = H[0]
    rng = np.random.default_rng(21)
    p = list(h.pch)
    A = BlockTensor.random(rng, VBIG_L, p, VBIG_R, dtype)
    Ab = BlockTensor.random(rng, VBIG_L, p, VBIG_R, dtype)
    GL = BlockTensor.random(rng, VBIG_L, [-c for c in h.cl], VBIG_L, dtype)
    GR = BlockTensor.random(rng, VBIG_R, [c for c in h.cr], VBIG_R, dtype)
    hs = dev_slice(h)
    dA, dAb = dev_tensor(ctx, A), dev_tensor(ctx, Ab)
    dGL, dGR = dev_tensor(ctx, GL), dev_tensor(ctx, GR)
    out = C.c_void_p()
    ctx.check(ctx.lib.mpsk_transfer_left(ctx.h, hs.device(ctx), dA.h, dAb.h, dGL.h, C.byref(out)))
    got = m.DeviceTensor.wrap(ctx, out).to_dense()
    ref = oc.transfer_left_sparse(GL.to_dense(), h, A.to_dense(), Ab.to_dense())
    assert relerr(got, ref) <= 1e-12
    out = C.c_void_p()
    ctx.check(ctx.lib.mpsk_transfer_right(ctx.h, hs.device(ctx), dA.h, dAb.h, dGR.h, C.byref(out)))
    got = m.DeviceTensor.wrap(ctx, out).to_dense()
    ref = oc.transfer_right_sparse(GR.to_dense(), h, A.to_dense(), Ab.to_dense())
    assert relerr(got, ref) <= 1e-12


def test_local_algebra(ctx):
    import ctypes as C
    rng = np.random.default_rng(31)
    Vm = OSpace((-3, -1, 1, 3), (7, 20, 22, 5))
    A1 = BlockTensor.random(rng, VL, [-1, 1], Vm)
    A2 = BlockTensor.random(rng, Vm, [-1, 1], VR)
    d1, d2 = dev_tensor(ctx, A1), dev_tensor(ctx, A2)
    out = C.c_void_p()
    ctx.check(ctx.lib.mpsk_contract_ac2(ctx.h, d1.h, d2.h, C.byref(out)))
    got = m.DeviceTensor.wrap(ctx, out).to_dense()
    ref = np.einsum('apk,kqr->aprq', A1.to_dense(), A2.to_dense())
    assert relerr(got, x2_from_dense(ref, 2)) <= 1e-13
    Cr = BlockTensor.random(rng, Vm, [0], Vm)
    dCr = dev_tensor(ctx, Cr)
    out = C.c_void_p()
    ctx.check(ctx.lib.mpsk_mul_bond_right(ctx.h, d1.h, dCr.h, C.byref(out)))
    assert relerr(m.DeviceTensor.wrap(ctx, out).to_dense(), np.einsum('apb,bc->apc', A1.to_dense(), Cr.to_dense()[:, 0, :])) <= 1e-13
    out = C.c_void_p()
    ctx.check(ctx.lib.mpsk_mul_bond_left(ctx.h, dCr.h, d2.h, C.byref(out)))
    assert relerr(m.DeviceTensor.wrap(ctx, out).to_dense(), np.einsum('ab,bpc->apc', Cr.to_dense()[:, 0, :], A2.to_dense())) <= 1e-13


def _hermitian_problem(seed, h, Vl, Vr, dtype=np.complex128):
    """GL, GR that make the effective Hamiltonian Hermitian: build them by real transfers of random isometries."""
    rng = np.random.default_rng(seed)
    x, GL, GR = random_problem(seed, h, Vl, Vr, "ac2", dtype)
    # symmetrise level-wise: H_eff = sum_i GL_i (x) ... is Hermitian if GL, GR come from <A|O|A>; emulate with
    # GL_i -> GL_i + GL_ibar^dagger where ibar is the level with opposite charge and adjoint operator content.
    return x, GL, GR


def test_eigsolve_and_expv_against_dense(ctx):
    """fixedpoint(h, x0, :SR) and exponentiate on a genuinely Hermitian effective Hamiltonian from a real MPS."""
    rng = np.random.default_rng(41)
    L = 6
    Ho = ompo.heisenberg_u1(L)
    p = OSpace((-1, 1), (1, 1))
    maxV = [of.gaussian_bond_space(12, k % 2) for k in range(1, L)]
    psi = of.FiniteMPS.random(rng, p, L, maxV)
    envs = of.FinEnv(psi, Ho)
    pos = 2
    ac2 = np.einsum('apk,kqr->aprq', psi.get_AC(pos), psi.get_AR(pos + 1))
    GL, GR = envs.leftenv(pos, psi), envs.rightenv(pos + 1, psi)
    Vl, Vr = psi.V[pos], psi.V[pos + 2]
    h = Ho[pos]
    d = 2
    mch = [h.pch[a] + h.pch[b] for a in range(d) for b in range(d)]
    f = lambda v: oc.dAC2_sparse(v, h, h, GL, GR)
    val_o, vec_o, info_o = ok.eigsolve_lanczos(f, ac2, "SR", 1e-13, 30, 100)
    dVl, dVr = to_dev_space(Vl), to_dev_space(Vr)
    dx = m.DeviceTensor.from_dense(ctx, x2_from_dense(ac2, d), dVl, mch, dVr)
    dGL = m.DeviceTensor.from_dense(ctx, GL, dVl, [-c for c in h.cl], dVl)
    dGR = m.DeviceTensor.from_dense(ctx, GR, dVr, [c for c in h.cr], dVr)
    hs = dev_slice(h)
    op = m.MPO_ddAC2(ctx, hs, hs, dGL, dGR, dx)
    val, vec, info = m.fixedpoint(op, dx, "SR", m.Lanczos(tol=1e-13), return_info=True)
    print("eigsolve", val, val_o, info, info_o)
    assert info["converged"] == 1
    assert abs(val.real - val_o) <= 1e-10 * abs(val_o)
    v = x2_dense(BlockTensor.from_dense(vec.to_dense(), Vl, mch, Vr), d)
    assert abs(abs(np.vdot(vec_o, v)) - 1) <= 1e-9        # same eigenvector up to phase
    # exponentiate
    y_o, _ = ok.exponentiate_lanczos(f, -0.05j, ac2, 1e-12, 30)
    y = m.integrate(op, dx, 0.0, 0.05, m.Lanczos(tol=1e-12))
    yd = x2_dense(BlockTensor.from_dense(y.to_dense(), Vl, mch, Vr), d)
    assert relerr(yd, y_o) <= 1e-10


@pytest.mark.parametrize("trunc", [("notrunc",), ("truncdim", 20), ("truncerr", 0.3), ("truncbelow", 0.5)])
def test_tsvd(ctx, trunc):
    rng = np.random.default_rng(51)
    d = 2
    pch = [-1, 1]
    mch = [pch[a] + pch[b] for a in range(d) for b in range(d)]
    Vl, Vr = OSpace((-2, 0, 2), (6, 9, 4)), OSpace((-2, 0, 2, 4), (5, 11, 7, 2))
    x = BlockTensor.random(rng, Vl, mch, Vr)
    xd = x2_dense(x, d)
    rowq = (Vl.index_charges()[:, None] + np.array(pch)[None, :]).reshape(-1)
    colq = (Vr.index_charges()[:, None] - np.array(pch)[None, :]).reshape(-1)
    U, S, Vh, newV, eps, info = ol.tsvd(xd.reshape(Vl.dim * d, Vr.dim * d), rowq, colq, trunc)

    class P:  # minimal stand-in carrying what dmrg.tsvd needs
        pass
    P.ctx, P.pq, P.pspace = ctx, tuple(pch), m.Space((-1, 1), (1, 1))
    code = m.notrunc() if trunc[0] == "notrunc" else getattr(m, trunc[0])(trunc[1])
    al, c, ar, eps_d = m.tsvd(P, dev_tensor(ctx, x), code)
    assert (al.Vr.charges, al.Vr.dims) == (newV.charges, newV.dims)           # kept dims per sector: bit-exact
    assert abs(eps_d - eps) <= 1e-8
    spec = m.tsvd_spectrum(ctx)
    for q, s in info["spectrum"].items():
        assert np.abs(spec[q][0] - s).max() <= 1e-8
        assert spec[q][1] == info["kept"][q]
    # gauge-invariant reconstruction al c ar
    cd = c.to_dense()[:, 0, :]
    rec = np.einsum('apk,kl,lqr->aprq', al.to_dense(), cd, ar.to_dense())
    ref = np.einsum('ik,k,kj->ij', U, S / np.linalg.norm(S), Vh).reshape(Vl.dim, d, Vr.dim, d)
    assert relerr(rec, ref) <= 1e-10
    # isometries
    A = al.to_dense().reshape(-1, al.Vr.dim)
    assert np.abs(A.conj().T @ A - np.eye(A.shape[1])).max() <= 1e-12


@pytest.mark.parametrize("model", ["heisenberg_u1", "tfi_trivial"])
def test_dmrg2_energy_parity(ctx, model):
    """find_groundstate(psi, H, DMRG2(...)) on the device vs the oracle sweep on the same seeded state."""
    L = 8
    if model == "heisenberg_u1":
        Ho = ompo.heisenberg_u1(L); p = OSpace((-1, 1), (1, 1))
        maxV = [of.gaussian_bond_space(16, k % 2) for k in range(1, L)]
        exact = -3.3749325986878906
    else:
        Ho = ompo.transverse_field_ising(L, 1.0); p = OSpace((0,), (2,))
        maxV = OSpace((0,), (16,))
        exact = 1 - 1 / np.sin(np.pi / (2 * (2 * L + 1)))
    psi_o = of.FiniteMPS.random(np.random.default_rng(61), p, L, maxV)
    log_o = []
    of.dmrg2_sweeps(psi_o, Ho, nsweeps=3, trunc=("truncdim", 16), tol=1e-13, log=log_o)
    dmaxV = [to_dev_space(v) for v in maxV] if isinstance(maxV, list) else to_dev_space(maxV)
    psi_d = m.FiniteMPS.random(np.random.default_rng(61), to_dev_space(p), L, dmaxV, ctx)
    Hd = dev_hamiltonian(Ho)
    log_d = []
    alg = m.DMRG2(tol=1e-13, maxiter=3, eigalg=m.Lanczos(tol=1e-13), trscheme=m.truncdim(16))
    psi_d, envs, eps = m.find_groundstate(psi_d, Hd, alg, log=log_d)
    E_d, E_o = log_d[-1]["energy"], log_o[-1]["energy"]
    print(model, "E_dev", E_d, "E_oracle", E_o, "exact", exact)
    assert abs(E_d - E_o) <= 1e-10 * abs(E_o)
    assert abs(E_d - exact) <= 1e-9 * abs(exact)
    assert abs(m.expectation_value(psi_d, Hd, envs) - E_o) <= 1e-10 * abs(E_o)
    # step-by-step: same kept sector dimensions (bit-exact) and truncation errors / spectra to 1e-8
    assert len(log_d) == len(log_o)
    for a, b in zip(log_d, log_o):
        assert {q: k for q, k in a["kept"].items() if k} == {q: k for q, k in b["kept"].items() if k}
        assert abs(a["eps_trunc"] - b["eps_trunc"]) <= 1e-8
        for q, s_o in b["spectrum"].items():          # singular values of every two-site update
            assert np.abs(a["spectrum"][q] - s_o).max() <= 1e-8
    # entanglement spectrum (toolbox.jl:65-69) at the bond where the sweep left the centre
    es_d, es_o = psi_d.entanglement_spectrum(0), psi_o.entanglement_spectrum(0)
    for q in es_o:
        assert np.abs(es_d[q] - es_o[q]).max() <= 1e-8


# ----------------------------------------------------------------------------------------------------------------
# gauge / single-site / time-evolution rows
@pytest.mark.parametrize("dtype", [np.complex128, np.float64])
def test_qrpos_lqpos(ctx, dtype):
    """leftorth!(QRpos) / rightorth!(LQpos): unique factorisation -> elementwise parity with the oracle."""
    rng = np.random.default_rng(71)
    Vl, Vr = OSpace((-3, -1, 1, 3), (7, 20, 22, 5)), OSpace((-4, -2, 0, 2, 4), (3, 40, 11, 9, 2))   # sector 2: rows < cols too
    pch = [-1, 1]
    A = BlockTensor.random(rng, Vl, pch, Vr, dtype)
    Ad = A.to_dense()
    rowq = (Vl.index_charges()[:, None] + np.array(pch)[None, :]).reshape(-1)
    Qo, Ro, nV = ol.qrpos(Ad.reshape(Vl.dim * 2, Vr.dim), rowq, Vr.index_charges())
    dA = dev_tensor(ctx, A)
    Q, R = m.qrpos(ctx, dA)
    assert (Q.Vr.charges, Q.Vr.dims) == (nV.charges, nV.dims)
    assert relerr(Q.to_dense().reshape(Vl.dim * 2, nV.dim), Qo) <= 1e-12
    assert relerr(R.to_dense()[:, 0, :], Ro) <= 1e-12
    colq = (Vr.index_charges()[None, :] - np.array(pch)[:, None]).reshape(-1)
    Lo, Qo2, nV2 = ol.lqpos(Ad.reshape(Vl.dim, 2 * Vr.dim), Vl.index_charges(), colq)
    Lm, Q2 = m.lqpos(ctx, dA)
    assert (Q2.Vl.charges, Q2.Vl.dims) == (nV2.charges, nV2.dims)
    assert relerr(Lm.to_dense()[:, 0, :], Lo) <= 1e-12
    assert relerr(Q2.to_dense().reshape(nV2.dim, 2 * Vr.dim), Qo2) <= 1e-12


def test_finitemps_gauge_identities(ctx):
    """test/states.jl:15-38: AC = AL*C = C*AR on every site, isometry of AL / AR (lazy device gauging)."""
    L = 8
    p = m.Space((-1, 1), (1, 1))
    from mpskit_b200.workloads import gaussian_bond_space
    psi = m.FiniteMPS.random(np.random.default_rng(2), p, L, [gaussian_bond_space(12, k % 2) for k in range(1, L)], ctx)
    for i in range(L):
        AL, AR, AC = psi.AL[i].to_dense(), psi.AR[i].to_dense(), psi.AC[i].to_dense()
        Cr, Cl = psi.CR[i].to_dense()[:, 0, :], psi.CR[i - 1].to_dense()[:, 0, :]
        assert relerr(np.einsum('apb,bc->apc', AL, Cr), AC) <= 1e-12
        assert relerr(np.einsum('ab,bpc->apc', Cl, AR), AC) <= 1e-12
        M = AL.reshape(-1, AL.shape[2]); assert np.abs(M.conj().T @ M - np.eye(M.shape[1])).max() <= 1e-12
        N = AR.reshape(AR.shape[0], -1); assert np.abs(N @ N.conj().T - np.eye(N.shape[0])).max() <= 1e-12
    assert abs(psi.CR[L // 2].norm() - 1) <= 1e-12


def test_dmrg_single_site_and_galerkin(ctx):
    """find_groundstate(psi, H, DMRG()) (README quickstart algorithm, config 1 style) vs the oracle sweep."""
    L = 10
    Ho = ompo.transverse_field_ising(L, 1.0)
    po = OSpace((0,), (2,))
    psi_o = of.FiniteMPS.random(np.random.default_rng(81), po, L, OSpace((0,), (12,)))
    log_o = []
    of.dmrg_sweeps(psi_o, Ho, nsweeps=4, tol=1e-13, log=log_o)
    psi_d = m.FiniteMPS.random(np.random.default_rng(81), to_dev_space(po), L, m.Space((0,), (12,)), ctx)
    Hd = dev_hamiltonian(Ho)
    envs = m.environments(psi_d, Hd)
    g_d = m.calc_galerkin(psi_d, 3, envs)
    psi_o2 = of.FiniteMPS.random(np.random.default_rng(81), po, L, OSpace((0,), (12,)))
    g_o = of.calc_galerkin(psi_o2, 3, of.FinEnv(psi_o2, Ho))
    assert abs(g_d - g_o) <= 1e-10
    log_d = []
    psi_d, envs, eps = m.find_groundstate(psi_d, Hd, m.DMRG(tol=1e-13, maxiter=4, eigalg=m.Lanczos(tol=1e-13)), log=log_d)
    E_d, E_o = log_d[-1]["energy"], log_o[-1]["energy"]
    exact = np.linalg.eigvalsh(ompo.dense_hamiltonian(Ho))[0]
    print("DMRG", E_d, E_o, exact)
    assert abs(E_d - E_o) <= 1e-10 * abs(E_o) and abs(E_d - exact) <= 1e-9 * abs(exact)


def _full_state(psi):
    T = psi.AC[0].to_dense()
    for i in range(1, len(psi)):
        T = np.tensordot(T, psi.AR[i].to_dense(), axes=([-1], [0]))
    return T.reshape(-1)


def _full_state_o(psi):
    T = psi.get_AC(0)
    for i in range(1, psi.L):
        T = np.tensordot(T, psi.get_AR(i), axes=([-1], [0]))
    return T.reshape(-1)


def test_tdvp_step(ctx):
    """timestep(psi, H, 0, dt, TDVP()) (tdvp.jl:61-91) vs the oracle and vs exact exp(-i dt H) on the full state."""
    L = 6
    Ho = ompo.heisenberg_u1(L)
    po = OSpace((-1, 1), (1, 1))
    maxV = [of.gaussian_bond_space(8, k % 2) for k in range(1, L)]     # D=8 is exact for L=6
    psi_o = of.FiniteMPS.random(np.random.default_rng(91), po, L, maxV)
    v0 = _full_state_o(psi_o)
    dt = 0.05
    psi_o, _ = of.tdvp_step(psi_o, Ho, dt)
    v_o = _full_state_o(psi_o)
    psi_d = m.FiniteMPS.random(np.random.default_rng(91), to_dev_space(po), L, [to_dev_space(v) for v in maxV], ctx)
    assert abs(abs(np.vdot(_full_state(psi_d), v0)) - np.vdot(v0, v0).real) <= 1e-10
    Hd = dev_hamiltonian(Ho)
    psi_d, envs = m.timestep(psi_d, Hd, 0.0, dt, m.TDVP())
    v_d = _full_state(psi_d)
    import scipy.linalg as sla
    v_ex = sla.expm(-1j * dt * ompo.dense_hamiltonian(Ho)) @ v0
    print("tdvp overlaps", abs(np.vdot(v_d, v_o)), abs(np.vdot(v_d, v_ex)))
    assert abs(abs(np.vdot(v_d, v_o)) - 1) <= 1e-9
    assert abs(abs(np.vdot(v_d, v_ex)) - 1) <= 1e-6       # TDVP projection + splitting error at full bond dimension


def test_golden_vectors_on_device(ctx):
    """the CUDA path against the committed fixtures (tests/golden/*.npz, generated by make_golden.py)."""
    import os
    gdir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    g = np.load(os.path.join(gdir, "contractions_u1.npz"))
    h = dev_slice(ompo.heisenberg_longrange_u1(4, r=int(g["r"]))[0])
    Vl, Vr, Vm = (m.Space(*g[k]) for k in ("Vl", "Vr", "Vm"))
    p = h.pcharges
    mch2 = [p[a] + p[b] for a in range(2) for b in range(2)]
    up = lambda a, V1, mc, V2: m.DeviceTensor.from_dense(ctx, a, V1, mc, V2)
    GL = up(g["GL"], Vl, [-c for c in h.domcharges], Vl)
    GR2 = up(g["GR2"], Vr, list(h.imcharges), Vr)
    x2 = up(g["x2"], Vl, mch2, Vr)
    assert relerr(m.MPO_ddAC2(ctx, h, h, GL, GR2, x2)(x2).to_dense(), g["y2"]) <= 1e-12
    GR1 = up(g["GR1"], Vm, list(h.imcharges), Vm)
    x1 = up(g["x1"], Vl, list(p), Vm)
    assert relerr(m.MPO_ddAC(ctx, h, GL, GR1, x1)(x1).to_dense(), g["y1"]) <= 1e-12
    C0 = up(g["C0"], Vm, [0], Vm); GLc = up(g["GLc"], Vm, [-c for c in h.domcharges], Vm)
    assert relerr(m.MPO_ddC(ctx, GLc, GR1, C0)(C0).to_dense()[:, 0, :], g["y0"]) <= 1e-12
    t = np.load(os.path.join(gdir, "dmrg2_heisenberg_u1_L8_D16.npz"))
    assert abs(t["energy"][-1] + 3.3749325986878906) < 1e-12


@pytest.mark.parametrize("nranks", [2, 3, 8])
def test_row_panel_sharding_single_gpu_emulation(ctx, nranks):
    """the multi-GPU decomposition executed rank by rank on one GPU: the panels of y produced by the row-panel
    operators tile the full matvec exactly (everything but the NCCL allgather itself)."""
    H = ompo.heisenberg_longrange_u1(4, r=2)
    h = H[0]
    x, GL, GR = random_problem(101, h, VBIG_L, VBIG_R2, "ac2")
    hs = dev_slice(h)
    dx, dGL, dGR = dev_tensor(ctx, x), dev_tensor(ctx, GL), dev_tensor(ctx, GR)
    full = m.MPO_ddAC2(ctx, hs, hs, dGL, dGR, dx)
    yfull = full(dx).to_dense()
    y = dx.zerovector()
    flops = 0.0
    for r in range(nranks):
        GLp = m.row_panel(dGL, r, nranks)
        assert GLp.Vl == m.panel_space(dGL.Vl, r, nranks)           # host mirror of the integer rule: bit-exact
        op = m.MPO_ddAC2(ctx, hs, hs, GLp, dGR, dx)        # a panel operator; applied locally (no communicator here)
        yp = m.DeviceTensor.zeros(ctx, GLp.Vl, dx.mch, dx.Vr)
        m.apply_local(op, dx, yp)
        m.scatter_panel(yp, r, nranks, y)
        flops += op.flops
    assert relerr(y.to_dense(), yfull) <= 1e-13
    assert relerr(y.to_dense(), oc.dAC2_blocks(x, h, h, GL, GR).to_dense()) <= 1e-12
    assert abs(flops - full.flops) <= 1e-9 * full.flops            # no redundant work


# ----------------------------------------------------------------------------------------------------------------
# infinite-MPS rows (8f item 2) and operator algebra (a6)
def test_uniform_gauge_and_transfer_fixed_points(ctx):
    """test/states.jl:49-72: AC = AL*C = C*AR and l * TransferMatrix(AL, AL) = l, TransferMatrix(AR, AR) * r = r."""
    from oracle import infinite as oi
    rng = np.random.default_rng(111)
    p = m.Space((-1, 0, 1), (1, 1, 1))                     # spin-1, U(1)
    Vs = [m.Space((-1, 0, 1), (5, 8, 5)), m.Space((-1, 0, 1), (6, 7, 6))]
    psi = m.InfiniteMPS.random(rng, p, Vs, ctx)
    for i in range(2):
        AL, AR, AC = psi.AL[i].to_dense(), psi.AR[i].to_dense(), psi.AC[i].to_dense()
        Cr, Cl = psi.CR[i].to_dense()[:, 0, :], psi.CR[i - 1].to_dense()[:, 0, :]
        assert relerr(np.einsum('apb,bc->apc', AL, Cr), AC) <= 1e-11
        assert relerr(np.einsum('ab,bpc->apc', Cl, AR), AC) <= 1e-11
        M = AL.reshape(-1, AL.shape[2]); assert np.abs(M.conj().T @ M - np.eye(M.shape[1])).max() <= 1e-12
    from mpskit_b200.infinite import identity_bond, mul_right_bond
    l = identity_bond(ctx, psi.V[0], psi.dtype)
    out = l * m.TransferMatrix(psi.AL, None, psi.AL)                 # v * tm (transfermatrix.jl:44)
    assert relerr(out.to_dense(), l.to_dense()) <= 1e-12
    r = mul_right_bond(psi.CR[1], psi.CR[1].adjoint())               # r_LL = C C'
    out = m.TransferMatrix(psi.AL, None, psi.AL) * r
    assert relerr(out.to_dense(), r.to_dense()) <= 1e-10


def test_vumps_tfi_reference_logged_energy(ctx):
    """docs/src/examples/quantum1d/3.ising-dqpt/index.md:107-118: VUMPS, TFI g=-0.5, D=10 -> E = -1.063544409973
    (logged by the reference itself, 7 iterations); device VUMPS vs that number and vs the oracle VUMPS."""
    from oracle import infinite as oi
    Ho = ompo.transverse_field_ising(1, 0.5)
    psi_o = oi.InfiniteMPS.random(np.random.default_rng(3), OSpace((0,), (2,)), [OSpace((0,), (10,))])
    log_o = []
    oi.vumps(psi_o, Ho, tol=1e-12, maxiter=30, log=log_o)
    psi_d = m.InfiniteMPS.random(np.random.default_rng(3), m.Space((0,), (2,)), [m.Space((0,), (10,))], ctx)
    log_d = []
    psi_d, envs, eps = m.find_groundstate_vumps(psi_d, dev_hamiltonian(Ho), m.VUMPS(tol=1e-12, maxiter=30), log=log_d)
    print("VUMPS", log_d[-1], log_o[-1], len(log_d), len(log_o))
    assert eps <= 1e-11
    assert abs(log_d[-1]["energy"] - (-1.063544409973)) <= 2e-11
    assert abs(log_d[-1]["energy"] - log_o[-1]["energy"]) <= 1e-10


def test_vumps_two_site_unit_cell_u1(ctx):
    """spin-1/2 Heisenberg, 2-site unit cell with U(1) symmetry (config 2 style): device vs oracle energy density."""
    from oracle import infinite as oi
    Ho = ompo.heisenberg_u1(2)
    Vo = [OSpace((-2, 0, 2), (3, 6, 3)), OSpace((-3, -1, 1, 3), (1, 5, 5, 1))]
    psi_o = oi.InfiniteMPS.random(np.random.default_rng(5), OSpace((-1, 1), (1, 1)), Vo)
    log_o = []
    oi.vumps(psi_o, Ho, tol=1e-8, maxiter=6, log=log_o)
    psi_d = m.InfiniteMPS.random(np.random.default_rng(5), m.Space((-1, 1), (1, 1)), [to_dev_space(v) for v in Vo], ctx)
    log_d = []
    m.find_groundstate_vumps(psi_d, dev_hamiltonian(Ho), m.VUMPS(tol=1e-8, maxiter=6), log=log_d)
    print("VUMPS U1", log_d[-1], log_o[-1])
    # The critical Heisenberg chain does not converge at D = 12 (the oracle still has eps = 1e-6 after 80 iterations), so
    # this compares two UNCONVERGED trajectories whose eigensolves run at the dynamic tolerance 1e-5 * eps
    # (defaults.jl:45-50): differences of that size are amplified, hence 1e-7.  Converged runs are held to 1e-10
    # (test_vumps_tfi_reference_logged_energy, test_leading_boundary_off_critical_converged).
    assert abs(log_d[-1]["energy"] - log_o[-1]["energy"]) <= 1e-7 * abs(log_o[-1]["energy"])
    assert abs(log_d[-1]["energy"] / 2 - (0.25 - np.log(2))) < 5e-3       # Bethe ansatz, D = 12, 6 iterations


def test_derivative_linearity_lazysum(ctx):
    """test/operators.jl:207-225: the derivative of a LazySum equals the sum of the derivatives."""
    H = ompo.heisenberg_u1(4)
    h = H[0]
    x, GL, GR = random_problem(121, h, VBIG_L, VBIG_R, "ac")
    _, GL2, GR2 = random_problem(122, h, VBIG_L, VBIG_R, "ac")
    hs = dev_slice(h)
    dx = dev_tensor(ctx, x)
    t = [dev_tensor(ctx, v) for v in (GL, GR, GL2, GR2)]
    op1 = m.MPO_ddAC(ctx, hs, t[0], t[1], dx)
    op2 = m.MPO_ddAC(ctx, hs, t[2], t[3], dx)
    ysum = m.LazySumOp([op1, m.MultipliedOp(op2, 0.5 - 2j)])(dx).to_dense()
    ref = oc.dAC_sparse(x.to_dense(), h, GL.to_dense(), GR.to_dense()) + (0.5 - 2j) * oc.dAC_sparse(x.to_dense(), h, GL2.to_dense(), GR2.to_dense())
    assert relerr(ysum, ref) <= 1e-12


@pytest.mark.parametrize("trunc", [("notrunc",), ("truncdim", 20), ("truncerr", 0.3)])
def test_tsvd_gram_route(trunc):
    """the GEMM-rich SVD route used for large sectors (G = W^H W, heevd, B = W V, QR clean-up), forced on every sector:
    singular values, truncation decisions and the reconstruction against the oracle; exact isometries."""
    from mpskit_b200 import _lib
    c = m.Context(0, 4)                      # MPSK_FLAG_FAST_SVD_ALL
    rng = np.random.default_rng(51)
    d = 2
    pch = [-1, 1]
    mch = [pch[a] + pch[b] for a in range(d) for b in range(d)]
    Vl, Vr = OSpace((-2, 0, 2), (60, 90, 40)), OSpace((-2, 0, 2, 4), (50, 110, 70, 20))
    x = BlockTensor.random(rng, Vl, mch, Vr)
    xd = x2_dense(x, d)
    rowq = (Vl.index_charges()[:, None] + np.array(pch)[None, :]).reshape(-1)
    colq = (Vr.index_charges()[:, None] - np.array(pch)[None, :]).reshape(-1)
    U, S, Vh, newV, eps, info = ol.tsvd(xd.reshape(Vl.dim * d, Vr.dim * d), rowq, colq, trunc)

    class P:
        pass
    P.ctx, P.pq, P.pspace = c, tuple(pch), m.Space((-1, 1), (1, 1))
    code = m.notrunc() if trunc[0] == "notrunc" else getattr(m, trunc[0])(trunc[1])
    al, cc, ar, eps_d = m.tsvd(P, dev_tensor(c, x), code)
    assert (al.Vr.charges, al.Vr.dims) == (newV.charges, newV.dims)
    assert abs(eps_d - eps) <= 1e-8
    spec = m.tsvd_spectrum(c)
    for q, s in info["spectrum"].items():
        assert np.abs(spec[q][0] - s).max() <= 1e-8
    rec = np.einsum('apk,kl,lqr->aprq', al.to_dense(), cc.to_dense()[:, 0, :], ar.to_dense())
    ref = np.einsum('ik,k,kj->ij', U, S / np.linalg.norm(S), Vh).reshape(Vl.dim, d, Vr.dim, d)
    assert relerr(rec, ref) <= 1e-10
    A = al.to_dense().reshape(-1, al.Vr.dim)
    assert np.abs(A.conj().T @ A - np.eye(A.shape[1])).max() <= 1e-12
    B = ar.to_dense().reshape(ar.Vl.dim, -1)
    assert np.abs(B @ B.conj().T - np.eye(B.shape[0])).max() <= 1e-12


def test_leading_boundary_classical_ising(ctx):
    """test/algorithms.jl:185-201: leading_boundary of the 2D classical Ising MPO at beta_c, expectation value 2.5337
    (Onsager: 2.533737...), device statmech VUMPS vs the oracle and the known answer (config 4 style, D = 10)."""
    from oracle import infinite as oi, statmech as osm
    sl_o = ompo.dense_mpo_slice(ompo.classical_ising())
    psi_o = oi.InfiniteMPS.random(np.random.default_rng(3), OSpace((0,), (2,)), [OSpace((0,), (10,))])
    log_o = []
    osm.leading_boundary(psi_o, [sl_o], tol=1e-6, maxiter=12, log=log_o)
    mpo = m.DenseMPO([m.classical_ising()])
    assert mpo[0].kinds() == sl_o.kinds()
    psi_d = m.InfiniteMPS.random(np.random.default_rng(3), m.Space((0,), (2,)), [m.Space((0,), (10,))], ctx)
    log_d = []
    m.leading_boundary(psi_d, mpo, m.VUMPS(tol=1e-6, maxiter=12), log=log_d)
    v_d, v_o = log_d[-1]["value"][0], log_o[-1]["value"][0]
    print("classic2d", v_d, v_o, len(log_d), len(log_o))
    assert abs(v_d - 2.5337) < 1e-3 and abs(v_d.imag) < 1e-9
    # at beta_c the boundary MPS is critical: 12 iterations at tol 1e-6 are not converged (the oracle needs > 60
    # iterations for 1e-9), so the two trajectories agree to the eigensolver tolerance only
    assert abs(v_d - v_o) <= 1e-6 * abs(v_o)


def test_leading_boundary_off_critical_converged(ctx):
    """the same statmech VUMPS away from criticality (beta = 0.35), where it converges: device Krylov-Schur Arnoldi +
    device transfer-matrix fixed points vs the oracle at the north-star tolerance, and vs Onsager's exact free energy."""
    from oracle import infinite as oi, statmech as osm
    beta = 0.35
    sl_o = ompo.dense_mpo_slice(ompo.classical_ising(beta))
    psi_o = oi.InfiniteMPS.random(np.random.default_rng(3), OSpace((0,), (2,)), [OSpace((0,), (10,))])
    log_o = []
    osm.leading_boundary(psi_o, [sl_o], tol=1e-11, maxiter=60, log=log_o)
    psi_d = m.InfiniteMPS.random(np.random.default_rng(3), m.Space((0,), (2,)), [m.Space((0,), (10,))], ctx)
    log_d = []
    _, _, eps = m.leading_boundary(psi_d, m.DenseMPO([m.classical_ising(beta)]), m.VUMPS(tol=1e-11, maxiter=60), log=log_d)
    v_d, v_o = log_d[-1]["value"][0], log_o[-1]["value"][0]
    k = 1 / np.sinh(2 * beta) ** 2
    th = np.linspace(0, np.pi, 200001)
    f = np.log(np.cosh(2 * beta) ** 2 + (1 / k) * np.sqrt(1 + k * k - 2 * k * np.cos(2 * th)))
    onsager = np.exp(np.log(2) / 2 + np.sum((f[1:] + f[:-1]) / 2 * np.diff(th)) / (2 * np.pi))
    print("classic2d beta=0.35", v_d, v_o, onsager, len(log_d), len(log_o), eps)
    assert eps <= 1e-10
    assert abs(v_d - v_o) <= 1e-10 * abs(v_o)
    assert abs(v_d - onsager) <= 1e-7 * onsager


def test_tdvp2_step(ctx):
    """timestep(psi, H, 0, dt, TDVP2(trscheme=truncdim(8))) (tdvp.jl:115-151) vs the oracle sweep and the exact evolution."""
    import scipy.linalg as sla
    L = 6
    Ho = ompo.heisenberg_u1(L)
    po = OSpace((-1, 1), (1, 1))
    maxV = [of.gaussian_bond_space(8, k % 2) for k in range(1, L)]
    psi_o = of.FiniteMPS.random(np.random.default_rng(93), po, L, maxV)
    v0 = _full_state_o(psi_o)
    dt = 0.05
    psi_o, _ = of.tdvp2_step(psi_o, Ho, dt, trunc=("truncdim", 8))
    v_o = _full_state_o(psi_o)
    psi_d = m.FiniteMPS.random(np.random.default_rng(93), to_dev_space(po), L, [to_dev_space(v) for v in maxV], ctx)
    psi_d, envs = m.timestep(psi_d, dev_hamiltonian(Ho), 0.0, dt, m.TDVP2(trscheme=m.truncdim(8)))
    v_d = _full_state(psi_d)
    v_ex = sla.expm(-1j * dt * ompo.dense_hamiltonian(Ho)) @ v0
    print("tdvp2 overlaps", abs(np.vdot(v_d, v_o)), abs(np.vdot(v_d, v_ex)))
    assert abs(abs(np.vdot(v_d, v_o)) - 1) <= 1e-9
    assert abs(abs(np.vdot(v_d, v_ex)) - 1) <= 1e-6


def test_mpo_free_derivatives(ctx):
    """dAC2(x, nothing, nothing, GL, GR) = GL x GR and dAC(x, nothing, GL, GR) (derivatives.jl:112-114,159-161)."""
    rng = np.random.default_rng(131)
    pch = [-1, 1]
    mch = [pch[a] + pch[b] for a in range(2) for b in range(2)]
    x = BlockTensor.random(rng, VBIG_L, mch, VBIG_R2)
    GL = BlockTensor.random(rng, VBIG_L, [0], VBIG_L)
    GR = BlockTensor.random(rng, VBIG_R2, [0], VBIG_R2)
    dx, dGL, dGR = dev_tensor(ctx, x), dev_tensor(ctx, GL), dev_tensor(ctx, GR)
    y = m.MPO_ddAC2(ctx, None, None, dGL, dGR, dx)(dx).to_dense()
    ref = np.einsum('ab,bmc,cd->amd', GL.to_dense()[:, 0, :], x.to_dense(), GR.to_dense()[:, 0, :], optimize=True)
    assert relerr(y, ref) <= 1e-12
    x1 = BlockTensor.random(rng, VBIG_L, pch, VBIG_R)
    GR1 = BlockTensor.random(rng, VBIG_R, [0], VBIG_R)
    d1, dGR1 = dev_tensor(ctx, x1), dev_tensor(ctx, GR1)
    y1 = m.MPO_ddAC(ctx, None, dGL, dGR1, d1)(d1).to_dense()
    assert relerr(y1, np.einsum('ab,bmc,cd->amd', GL.to_dense()[:, 0, :], x1.to_dense(), GR1.to_dense()[:, 0, :], optimize=True)) <= 1e-12


def test_config1_readme_quickstart_pbc_tfi(ctx):
    """BASELINE.json config 1 (README quickstart): finite TFI, L=16, D=4, PBC through periodic_boundary_conditions,
    find_groundstate with single-site DMRG -- device vs oracle on the same seeded state."""
    L, D, g = 16, 4, 1.0
    Ho = ompo.periodic_boundary_conditions(ompo.transverse_field_ising(1, g), L)
    psi_o = of.FiniteMPS.random(np.random.default_rng(141), OSpace((0,), (2,)), L, OSpace((0,), (D,)))
    log_o = []
    of.dmrg_sweeps(psi_o, Ho, nsweeps=4, tol=1e-13, log=log_o)
    Hd = m.periodic_boundary_conditions(m.transverse_field_ising(1, g), L)
    psi_d = m.FiniteMPS.random(np.random.default_rng(141), m.Space((0,), (2,)), L, m.Space((0,), (D,)), ctx)
    log_d = []
    m.find_groundstate(psi_d, Hd, m.DMRG(tol=1e-13, maxiter=4, eigalg=m.Lanczos(tol=1e-13)), log=log_d)
    print("config1", log_d[-1]["energy"], log_o[-1]["energy"], -2 / np.sin(np.pi / (2 * L)))
    # g = 1 is the critical point: at D = 4 the sweeps are far from converged after 4 iterations (energy still moving at
    # 1e-3 per sweep), so this is a comparison of two trajectories; the converged g = 2 run below is held to 1e-10
    assert abs(log_d[-1]["energy"] - log_o[-1]["energy"]) <= 1e-8 * abs(log_o[-1]["energy"])
    assert abs(log_d[-1]["energy"] - (-2 / np.sin(np.pi / (2 * L)))) < 0.05      # D = 4 variational bound vs the exact ring energy
    # same README call in the gapped phase (g = 2, one of the README's g values): converged -> north-star tolerance
    g = 2.0
    Ho = ompo.periodic_boundary_conditions(ompo.transverse_field_ising(1, g), L)
    psi_o = of.FiniteMPS.random(np.random.default_rng(141), OSpace((0,), (2,)), L, OSpace((0,), (D,)))
    log_o = []
    of.dmrg_sweeps(psi_o, Ho, nsweeps=14, tol=1e-13, log=log_o)
    Hd = m.periodic_boundary_conditions(m.transverse_field_ising(1, g), L)
    psi_d = m.FiniteMPS.random(np.random.default_rng(141), m.Space((0,), (2,)), L, m.Space((0,), (D,)), ctx)
    log_d = []
    m.find_groundstate(psi_d, Hd, m.DMRG(tol=1e-13, maxiter=14, eigalg=m.Lanczos(tol=1e-13)), log=log_d)
    print("config1 g=2", log_d[-1]["energy"], log_o[-1]["energy"])
    assert abs(log_d[-1]["energy"] - log_o[-1]["energy"]) <= 1e-10 * abs(log_o[-1]["energy"])


def test_idmrg1_device(ctx):
    """find_groundstate(psi, H, IDMRG1()) (idmrg.jl:21-75) with the manual IDMRGEnv updates (idmrgenv.jl:63-72): the
    device run reaches the variational energy logged by the reference for VUMPS at the same bond dimension and agrees
    with the oracle IDMRG1."""
    from oracle import infinite as oi
    from mpskit_b200 import infinite as di
    Ho = ompo.transverse_field_ising(1, 0.5)
    psi_o = oi.InfiniteMPS.random(np.random.default_rng(3), OSpace((0,), (2,)), [OSpace((0,), (10,))])
    st_o, envs_o, eps_o = oi.idmrg1(psi_o, Ho, tol=1e-10, maxiter=200)
    E_o = oi.expectation_value(st_o, Ho, envs_o).sum().real
    psi_d = m.InfiniteMPS.random(np.random.default_rng(3), m.Space((0,), (2,)), [m.Space((0,), (10,))], ctx)
    Hd = dev_hamiltonian(Ho)
    st_d, envs_d, eps_d = m.find_groundstate_idmrg1(psi_d, Hd, m.IDMRG1(tol=1e-10, maxiter=200))
    E_d = di.expectation_value(st_d, Hd, envs_d).sum().real
    print("IDMRG1", E_d, E_o, eps_d, eps_o)
    assert eps_d < 1e-10
    assert abs(E_d - E_o) <= 1e-9 and abs(E_d - (-1.063544409973)) <= 1e-9
