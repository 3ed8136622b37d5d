"""2+ GPU check (torchrun): a DMRG2 sweep with the matvec sharded over ranks gives the same energies on every rank
and the same as the oracle.  torchrun --nproc-per-node 2 tools/dist_dmrg2_check.py"""
import os, sys
sys.path.insert(0, ".")
import numpy as np
import torch, torch.distributed as dist
import mpskit_b200 as m
from oracle.spaces import Space as OSpace
from oracle import mpo as ompo, finite as of
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = m.Context(local)
def bcast(uid):
    box = [uid]; dist.broadcast_object_list(box, src=0); return box[0]
m.init_communicator(ctx, rank, world, bcast)
L = 8
Ho = ompo.heisenberg_u1(L)
maxV = [of.gaussian_bond_space(16, k % 2) for k in range(1, L)]
psi_o = of.FiniteMPS.random(np.random.default_rng(61), OSpace((-1, 1), (1, 1)), L, maxV)
log_o = []
of.dmrg2_sweeps(psi_o, Ho, nsweeps=2, trunc=("truncdim", 16), tol=1e-13, log=log_o)
psi = m.FiniteMPS.random(np.random.default_rng(61), m.Space((-1, 1), (1, 1)), L, [m.Space(v.charges, v.dims) for v in maxV], ctx)
H = m.MPOHamiltonian([m.SparseMPOSlice(s.Os, s.cl, s.cr, s.pch) for s in Ho.slices])
log = []
psi, envs, eps = m.find_groundstate(psi, H, m.DMRG2(tol=1e-13, maxiter=2, eigalg=m.Lanczos(tol=1e-13), trscheme=m.truncdim(16)), log=log)
E = torch.tensor([log[-1]["energy"]], dtype=torch.float64, device="cuda")
lst = [torch.zeros_like(E) for _ in range(world)]
dist.all_gather(lst, E)
if rank == 0:
    vals = [float(t) for t in lst]
    print("energies per rank", vals, "oracle", log_o[-1]["energy"])
    assert max(vals) - min(vals) == 0.0, "ranks diverged"
    assert abs(vals[0] - log_o[-1]["energy"]) <= 1e-10 * abs(vals[0])
    print("DIST DMRG2 OK")
dist.destroy_process_group()
