"""One full DMRG2 sweep (left->right then right->left) of config 3(ii), timed per phase, on 1 GPU or under torchrun on N GPUs
(matvec / Krylov basis / environment updates / stored environments / SVD sectors distributed).
python tools/full_sweep.py D L krylovdim [maxiter]  ->  gpurun_out/full_sweep_D{D}_n{N}.json"""
import os, sys, time, json
sys.path.insert(0, ".")
import numpy as np
import mpskit_b200 as m
from mpskit_b200.workloads import gaussian_bond_space
from mpskit_b200 import dmrg as dm
D = int(sys.argv[1]); L = int(sys.argv[2]); kd = int(sys.argv[3]); mi = int(sys.argv[4]) if len(sys.argv) > 4 else 1
rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
if world > 1:
    import torch, torch.distributed as dist
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = m.Context(local)
if world > 1:
    def bcast(uid):
        box = [uid]; dist.broadcast_object_list(box, src=0); return box[0]
    m.init_communicator(ctx, rank, world, bcast)
p = m.Space((-1, 1), (1, 1))
H = m.heisenberg_longrange_u1(L, r=6)
t0 = time.time()
psi = m.FiniteMPS.random(20240213 + 3000, p, L, [gaussian_bond_space(D, k % 2) for k in range(1, L)], ctx, on_device=True)
ctx.synchronize(); prep = time.time() - t0
if rank == 0:
    print("state prep", prep, "s; mem", ctx.bytes_in_use() / 1e9, "GB"); sys.stdout.flush()
envs = m.environments(psi, H)
alg = m.Lanczos(tol=1e-10, krylovdim=kd, maxiter=mi)
tot = dict(ac2=0.0, env_op=0.0, eig=0.0, tsvd=0.0, fidelity=0.0)
numops = 0; flops = 0.0; steps = 0; peak = 0
energies = []
import warnings; warnings.simplefilter("ignore")
tsweep = time.time()
for direction, rng_ in (("lr", range(0, L - 1)), ("rl", range(L - 3, -1, -1))):
    for pos in rng_:
        t0 = time.time()
        ac2 = dm.contract_ac2(psi, psi.AC[pos], psi.AR[pos + 1]) if direction == "lr" else dm.contract_ac2(psi, psi.AL[pos], psi.AC[pos + 1])
        ctx.synchronize(); t1 = time.time()
        h = m.ddAC2(pos, psi, H, envs, ac2)
        ctx.synchronize(); t2 = time.time()
        val, new, info = m.fixedpoint(h, ac2, "SR", alg, return_info=True)
        ctx.synchronize(); t3 = time.time()
        flops += h.flops * info["numops"] * world; numops += info["numops"]; h.free()
        al, c, ar, eps = dm.tsvd(psi, new, m.truncdim(D))
        ctx.synchronize(); t4 = time.time()
        alc = dm._mul_right(psi, al, c); rec = dm.contract_ac2(psi, alc, ar); v = rec.dot(ac2)
        if direction == "lr":
            psi.AC[pos] = (al, c); psi.AC[pos + 1] = (c, ar)
        else:
            psi.AC[pos + 1] = (c, ar); psi.AC[pos] = (al, c)
        for t in (ac2, new, alc, rec):
            t.free()
        envs.drop_stale(psi)
        ctx.synchronize(); t5 = time.time()
        tot["ac2"] += t1 - t0; tot["env_op"] += t2 - t1; tot["eig"] += t3 - t2; tot["tsvd"] += t4 - t3; tot["fidelity"] += t5 - t4
        steps += 1; energies.append(val.real); peak = max(peak, ctx.bytes_in_use())
        if rank == 0 and (steps % 10 == 0 or pos in (0, L - 2)):
            print(f"{direction} pos {pos}: E={val.real:.8f} D'={al.Vr.dim} eps={eps:.1e} numops={info['numops']} step={t5-t0:.2f}s "
                  f"elapsed={time.time()-tsweep:.0f}s mem={ctx.bytes_in_use()/1e9:.0f}GB"); sys.stdout.flush()
sweep = time.time() - tsweep
e = np.array(energies)
out = dict(D=D, L=L, chi=20, n_gpus=world, krylovdim=kd, maxiter=mi, steps=steps, sweep_seconds=sweep, matvecs=numops, matvec_flops=flops,
           eig_tflops=flops / tot["eig"] / 1e12, phases=tot, state_prep_seconds=prep, energy=float(e[-1]),
           peak_hbm_bytes_per_rank=int(peak),
           energy_increases_above_1e_8=int(np.sum(np.diff(e[L - 1:]) > 1e-8 * abs(e[-1]))),
           note="one DMRG2 iteration (l->r + r->l) from a random U(1) state, truncdim(D), Lanczos tol 1e-10; energy_increases counts "
                "updates of the second half-sweep whose Lanczos eigenvalue went up")
if rank == 0:
    print(json.dumps(out))
    open(f"gpurun_out/full_sweep_D{D}_n{world}.json", "w").write(json.dumps(out))
if world > 1:
    dist.destroy_process_group()
