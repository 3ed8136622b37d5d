"""2+ GPU check (torchrun): DMRG2 / DMRG / TDVP with the matvec, the Krylov basis, the environment updates, the stored
environments and the SVD sectors distributed over the ranks give (i) bit-identical numbers on every rank, (ii) the oracle's
energies, (iii) the same energies as a context that keeps the local steps replicated (MPSK_FLAG_REPLICATE_LOCAL).
torchrun --nproc-per-node 2 tools/dist_dmrg2_check.py"""
import os, sys
sys.path.insert(0, ".")
import numpy as np
import torch, torch.distributed as dist
import mpskit_b200 as m
from mpskit_b200 import _lib
from oracle.spaces import Space as OSpace
from oracle import mpo as ompo, finite as of
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))


def bcast(uid):
    box = [uid]; dist.broadcast_object_list(box, src=0); return box[0]


def same_on_all_ranks(vals, what):
    t = torch.tensor(vals, dtype=torch.float64, device="cuda")
    lst = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(lst, t)
    for o in lst[1:]:
        assert torch.equal(o, lst[0]), f"ranks diverged in {what}"


ctx = m.Context(local)
m.init_communicator(ctx, rank, world, bcast)
ctx_rep = m.Context(local, _lib.FLAG_REPLICATE_LOCAL)
m.init_communicator(ctx_rep, rank, world, bcast)

# ---- (a) tiny bond dimension (some panels are empty): DMRG2 vs oracle
L = 8
Ho = ompo.heisenberg_u1(L)
maxV = [of.gaussian_bond_space(16, k % 2) for k in range(1, L)]
psi_o = of.FiniteMPS.random(np.random.default_rng(61), OSpace((-1, 1), (1, 1)), L, maxV)
log_o = []
of.dmrg2_sweeps(psi_o, Ho, nsweeps=2, trunc=("truncdim", 16), tol=1e-13, log=log_o)
H = m.MPOHamiltonian([m.SparseMPOSlice(s.Os, s.cl, s.cr, s.pch) for s in Ho.slices])
alg = m.DMRG2(tol=1e-13, maxiter=2, eigalg=m.Lanczos(tol=1e-13), trscheme=m.truncdim(16))
psi = m.FiniteMPS.random(np.random.default_rng(61), m.Space((-1, 1), (1, 1)), L, [m.Space(v.charges, v.dims) for v in maxV], ctx)
log = []
psi, envs, eps = m.find_groundstate(psi, H, alg, log=log)
E = [e["energy"] for e in log]
same_on_all_ranks(E, "DMRG2 energies (D=16)")
assert abs(E[-1] - log_o[-1]["energy"]) <= 1e-10 * abs(E[-1]), (E[-1], log_o[-1]["energy"])
for a, b in zip(log, log_o):
    assert {q: k for q, k in a["kept"].items() if k} == {q: k for q, k in b["kept"].items() if k}
    assert abs(a["eps_trunc"] - b["eps_trunc"]) <= 1e-8

# ---- (b) larger bond dimension, long-range MPO (chi = 8): distributed vs replicated local steps vs oracle
L, D = 10, 48
Ho = ompo.heisenberg_longrange_u1(L, r=2)
maxV = [of.gaussian_bond_space(D, k % 2) for k in range(1, L)]
psi_o = of.FiniteMPS.random(np.random.default_rng(62), OSpace((-1, 1), (1, 1)), L, maxV)
log_o = []
of.dmrg2_sweeps(psi_o, Ho, nsweeps=2, trunc=("truncdim", D), tol=1e-13, log=log_o)
H = m.MPOHamiltonian([m.SparseMPOSlice(s.Os, s.cl, s.cr, s.pch) for s in Ho.slices])
alg = m.DMRG2(tol=1e-13, maxiter=2, eigalg=m.Lanczos(tol=1e-13), trscheme=m.truncdim(D))
res = []
for c in (ctx, ctx_rep):
    psi = m.FiniteMPS.random(np.random.default_rng(62), m.Space((-1, 1), (1, 1)), L, [m.Space(v.charges, v.dims) for v in maxV], c)
    log = []
    psi, envs, eps = m.find_groundstate(psi, H, alg, log=log)
    res.append([e["energy"] for e in log])
    same_on_all_ranks(res[-1], "DMRG2 energies (D=48)")
    assert abs(res[-1][-1] - log_o[-1]["energy"]) <= 1e-10 * abs(res[-1][-1]), (res[-1][-1], log_o[-1]["energy"])
assert np.abs(np.array(res[0]) - np.array(res[1])).max() <= 1e-10 * abs(res[0][-1])
# entanglement spectrum and energy read-outs on the sharded context
es = psi.entanglement_spectrum(L // 2)
E_dev = m.expectation_value(psi, H, envs)
same_on_all_ranks([E_dev], "expectation value")

# ---- (c) single-site DMRG (sharded dAC, QR) and a TDVP step (sharded exponentiate on dAC and dC)
psi = m.FiniteMPS.random(np.random.default_rng(63), m.Space((-1, 1), (1, 1)), L, [m.Space(v.charges, v.dims) for v in maxV], ctx)
log1 = []
psi, envs, eps = m.find_groundstate(psi, H, m.DMRG(tol=1e-12, maxiter=2, eigalg=m.Lanczos(tol=1e-13)), log=log1)
same_on_all_ranks([e["energy"] for e in log1], "DMRG energies")
psi_o = of.FiniteMPS.random(np.random.default_rng(63), OSpace((-1, 1), (1, 1)), L, maxV)
log_o1 = []
of.dmrg_sweeps(psi_o, Ho, nsweeps=2, tol=1e-13, log=log_o1)
assert abs(log1[-1]["energy"] - log_o1[-1]["energy"]) <= 1e-9 * abs(log1[-1]["energy"]), (log1[-1]["energy"], log_o1[-1]["energy"])
E0 = m.expectation_value(psi, H, envs)
psi2, envs2 = m.timestep(psi, H, 0.0, 0.05, m.TDVP())
E1 = m.expectation_value(psi2, H, envs2)
same_on_all_ranks([E0, E1], "TDVP energies")
assert abs(E1 - E0) <= 1e-8 * abs(E0), (E0, E1)             # TDVP conserves the energy (test/algorithms.jl:96-122)
if rank == 0:
    print("energies", res[0][-1], res[1][-1], log_o[-1]["energy"], "| DMRG", log1[-1]["energy"], log_o1[-1]["energy"], "| TDVP", E0, E1)
    print("DIST DMRG2 OK")
dist.destroy_process_group()
