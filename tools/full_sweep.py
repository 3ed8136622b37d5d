"""One full DMRG2 sweep (left->right then right->left) of config 3(ii) on one B200, timed per phase.
python tools/full_sweep.py D L krylovdim  ->  gpurun_out/full_sweep_D{D}.json"""
import sys, time, json
sys.path.insert(0, ".")
import numpy as np
import mpskit_b200 as m
from mpskit_b200.workloads import gaussian_bond_space
from mpskit_b200 import dmrg as dm
D = int(sys.argv[1]); L = int(sys.argv[2]); kd = int(sys.argv[3])
ctx = m.default_context()
p = m.Space((-1, 1), (1, 1))
H = m.heisenberg_longrange_u1(L, r=6)
t0 = time.time()
psi = m.FiniteMPS.random(20240213 + 3000, p, L, [gaussian_bond_space(D, k % 2) for k in range(1, L)], ctx, on_device=True)
ctx.synchronize(); prep = time.time() - t0
print("state prep", prep, "s; mem", ctx.bytes_in_use() / 1e9, "GB"); sys.stdout.flush()
envs = m.environments(psi, H)
alg = m.Lanczos(tol=1e-10, krylovdim=kd, maxiter=1)
tot = dict(ac2=0.0, env_op=0.0, eig=0.0, tsvd=0.0, fidelity=0.0)
numops = 0; flops = 0.0; steps = 0
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
        flops += h.flops * info["numops"]; numops += info["numops"]; h.free()
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
        steps += 1
        if steps % 10 == 0 or pos in (0, L - 2):
            print(f"{direction} pos {pos}: E={val.real:.8f} D'={al.Vr.dim} eps={eps:.1e} numops={info['numops']} step={t5-t0:.2f}s "
                  f"elapsed={time.time()-tsweep:.0f}s mem={ctx.bytes_in_use()/1e9:.0f}GB"); sys.stdout.flush()
sweep = time.time() - tsweep
out = dict(D=D, L=L, chi=20, krylovdim=kd, steps=steps, sweep_seconds=sweep, matvecs=numops, matvec_flops=flops,
           eig_tflops=flops / tot["eig"] / 1e12, phases=tot, state_prep_seconds=prep, energy=val.real,
           note="one DMRG2 iteration (l->r + r->l) from a random U(1) state, truncdim(D), Lanczos tol 1e-10, maxiter 1")
print(json.dumps(out))
open(f"gpurun_out/full_sweep_D{D}.json", "w").write(json.dumps(out))
