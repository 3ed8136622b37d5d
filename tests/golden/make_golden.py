"""Generates tests/golden/*.npz: seeded inputs and oracle outputs for every row of the hot path.
Run from the repo root:  python tests/golden/make_golden.py
The Julia reference cannot be executed in this image, so these vectors come from the numpy oracle; they freeze the
oracle's behaviour (CPU regression) and give the GPU tests fixed vectors that do not depend on oracle code at run time."""
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.spaces import Space  # noqa: E402
from oracle.blocks import BlockTensor  # noqa: E402
from oracle import mpo as ompo, contract as oc, linalg as ol, finite as of  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
SEED = 20240213


def sp(s):
    return np.array([s.charges, s.dims], dtype=np.int64)


def main():
    # ---- contractions on a U(1) range-2 Heisenberg slice (odim 8: scalar + operator entries)
    h = ompo.heisenberg_longrange_u1(4, r=2)[0]
    Vl, Vr = Space((-3, -1, 1, 3), (5, 9, 8, 3)), Space((-5, -3, -1, 1, 3), (2, 7, 10, 6, 3))
    rng = np.random.default_rng(SEED + 1)
    d = h.d
    mch2 = [h.pch[a] + h.pch[b] for a in range(d) for b in range(d)]
    x2 = BlockTensor.random(rng, Vl, mch2, Vr)
    GL = BlockTensor.random(rng, Vl, [-c for c in h.cl], Vl)
    GR2 = BlockTensor.random(rng, Vr, [c for c in h.cr], Vr)
    x2d = x2.to_dense().reshape(Vl.dim, d, d, Vr.dim).transpose(0, 1, 3, 2)
    y2 = oc.dAC2_sparse(x2d, h, h, GL.to_dense(), GR2.to_dense())
    Vm = Space((-4, -2, 0, 2, 4), (3, 8, 11, 7, 2))
    x1 = BlockTensor.random(rng, Vl, list(h.pch), Vm)
    GR1 = BlockTensor.random(rng, Vm, [c for c in h.cr], Vm)
    y1 = oc.dAC_sparse(x1.to_dense(), h, GL.to_dense(), GR1.to_dense())
    C0 = BlockTensor.random(rng, Vm, [0], Vm)
    GLc = BlockTensor.random(rng, Vm, [-c for c in h.cl], Vm)
    y0 = oc.dC_dense(C0.to_dense()[:, 0, :], GLc.to_dense(), GR1.to_dense())
    A = BlockTensor.random(rng, Vl, list(h.pch), Vm)
    Ab = BlockTensor.random(rng, Vl, list(h.pch), Vm)
    tl = oc.transfer_left_sparse(GL.to_dense(), h, A.to_dense(), Ab.to_dense())
    tr = oc.transfer_right_sparse(GR1.to_dense(), h, A.to_dense(), Ab.to_dense())
    np.savez_compressed(os.path.join(OUT, "contractions_u1.npz"), Vl=sp(Vl), Vr=sp(Vr), Vm=sp(Vm), r=2,
                        x2=x2.to_dense(), GL=GL.to_dense(), GR2=GR2.to_dense(), y2=y2.transpose(0, 1, 3, 2).reshape(Vl.dim, d * d, Vr.dim),
                        x1=x1.to_dense(), GR1=GR1.to_dense(), y1=y1, C0=C0.to_dense(), GLc=GLc.to_dense(), y0=y0,
                        A=A.to_dense(), Ab=Ab.to_dense(), tl=tl, tr=tr)
    # ---- truncated SVD
    rng = np.random.default_rng(SEED + 2)
    Wl, Wr = Space((-2, 0, 2), (6, 9, 4)), Space((-2, 0, 2, 4), (5, 11, 7, 2))
    pch = [-1, 1]
    m2 = [pch[a] + pch[b] for a in range(2) for b in range(2)]
    xs = BlockTensor.random(rng, Wl, m2, Wr)
    xd = xs.to_dense().reshape(Wl.dim, 2, 2, Wr.dim).transpose(0, 1, 3, 2)
    rowq = (Wl.index_charges()[:, None] + np.array(pch)[None, :]).reshape(-1)
    colq = (Wr.index_charges()[:, None] - np.array(pch)[None, :]).reshape(-1)
    out = dict(Wl=sp(Wl), Wr=sp(Wr), x=xs.to_dense())
    for name, tr_ in (("notrunc", ("notrunc",)), ("truncdim", ("truncdim", 20)), ("truncerr", ("truncerr", 0.3)), ("truncbelow", ("truncbelow", 0.5))):
        U, S, Vh, newV, eps, info = ol.tsvd(xd.reshape(Wl.dim * 2, Wr.dim * 2), rowq, colq, tr_)
        out[name + "_newV"] = sp(newV); out[name + "_eps"] = eps; out[name + "_S"] = S
        out[name + "_rec"] = np.einsum('ik,k,kj->ij', U, S / np.linalg.norm(S), Vh).reshape(Wl.dim, 2, Wr.dim, 2)
    np.savez_compressed(os.path.join(OUT, "tsvd_u1.npz"), **out)
    # ---- DMRG2 trajectories (energies, truncation errors, kept dims per step)
    for name, Ho, p, maxV in (
            ("heisenberg_u1_L8_D16", ompo.heisenberg_u1(8), Space((-1, 1), (1, 1)), [of.gaussian_bond_space(16, k % 2) for k in range(1, 8)]),
            ("tfi_L8_D16", ompo.transverse_field_ising(8, 1.0), Space((0,), (2,)), Space((0,), (16,)))):
        psi = of.FiniteMPS.random(np.random.default_rng(61), p, 8, maxV)
        log = []
        of.dmrg2_sweeps(psi, Ho, nsweeps=3, trunc=("truncdim", 16), tol=1e-13, log=log)
        np.savez_compressed(os.path.join(OUT, f"dmrg2_{name}.npz"), energy=np.array([s["energy"] for s in log]),
                            eps_trunc=np.array([s["eps_trunc"] for s in log]), pos=np.array([s["pos"] for s in log]),
                            kept=np.array([sum(s["kept"].values()) for s in log]))
    # ---- infinite-MPS / statmech / time-evolution trajectories (frozen oracle outputs)
    from oracle import infinite as oi, statmech as osm
    psi = oi.InfiniteMPS.random(np.random.default_rng(3), Space((0,), (2,)), [Space((0,), (10,))])
    log = []
    oi.vumps(psi, ompo.transverse_field_ising(1, 0.5), tol=1e-12, maxiter=30, log=log)
    np.savez_compressed(os.path.join(OUT, "vumps_tfi_g0.5_D10.npz"), energy=np.array([s["energy"] for s in log]),
                        eps=np.array([s["eps"] for s in log]))
    psi = oi.InfiniteMPS.random(np.random.default_rng(3), Space((0,), (2,)), [Space((0,), (10,))])
    log = []
    osm.leading_boundary(psi, [ompo.dense_mpo_slice(ompo.classical_ising())], tol=1e-6, maxiter=12, log=log)
    np.savez_compressed(os.path.join(OUT, "statmech_ising_D10.npz"), value=np.array([s["value"][0] for s in log]),
                        eps=np.array([s["eps"] for s in log]))
    L = 6
    Ho = ompo.heisenberg_u1(L)
    st = of.FiniteMPS.random(np.random.default_rng(91), Space((-1, 1), (1, 1)), L, [of.gaussian_bond_space(8, k % 2) for k in range(1, L)])
    st, _ = of.tdvp_step(st, Ho, 0.05)
    T = st.get_AC(0)
    for i in range(1, L):
        T = np.tensordot(T, st.get_AR(i), axes=([-1], [0]))
    np.savez_compressed(os.path.join(OUT, "tdvp_heisenberg_L6.npz"), state=T.reshape(-1))
    print("golden vectors written to", OUT)


if __name__ == "__main__":
    main()
