"""DMRG2 partial-sweep timing (development aid): python tools/sweep_perf.py D L nsteps"""
import sys, time
sys.path.insert(0, ".")
import numpy as np
import mpskit_b200 as m
from mpskit_b200.workloads import gaussian_bond_space
from mpskit_b200 import dmrg as dm
D = int(sys.argv[1]); L = int(sys.argv[2]); nsteps = int(sys.argv[3])
ctx = m.default_context()
p = m.Space((-1, 1), (1, 1))
H = m.heisenberg_longrange_u1(L, r=6)
t0 = time.time()
psi = m.FiniteMPS.random(1234, p, L, [gaussian_bond_space(D, k % 2) for k in range(1, L)], ctx, on_device=True)
ctx.synchronize()
print("state prep s", time.time() - t0, "centre bond", psi.V[L // 2]); sys.stdout.flush()
envs = m.environments(psi, H)
alg = m.Lanczos(tol=1e-10, krylovdim=30, maxiter=1)
pos0 = L // 2 - nsteps // 2
t0 = time.time(); envs.leftenv(pos0, psi); envs.rightenv(pos0 + 1, psi); ctx.synchronize()
print("env build s", time.time() - t0); sys.stdout.flush()
# emulate the l->r sweep body on a window: need AC[pos0]
import ctypes as C
for step in range(nsteps):
    pos = pos0 + step
    tt = {}
    t0 = time.time(); ac2 = dm.contract_ac2(psi, psi.AC[pos], psi.AR[pos + 1]); ctx.synchronize(); tt["ac2"] = time.time() - t0
    t0 = time.time(); h = m.ddAC2(pos, psi, H, envs, ac2); ctx.synchronize(); tt["envs+op"] = time.time() - t0
    t0 = time.time(); val, new, info = m.fixedpoint(h, ac2, "SR", alg, return_info=True); ctx.synchronize(); tt["eig"] = time.time() - t0
    flops = h.flops; h.free()
    t0 = time.time(); al, c, ar, eps = dm.tsvd(psi, new, m.truncdim(D)); ctx.synchronize(); tt["tsvd"] = time.time() - t0
    t0 = time.time(); psi.AC[pos] = (al, c); psi.AC[pos + 1] = (c, ar); ctx.synchronize(); tt["set"] = time.time() - t0
    print(f"pos {pos}: E={val.real:.10f} numops={info['numops']} eps={eps:.2e} D'={al.Vr.dim} "
          + " ".join(f"{k}={v:.3f}s" for k, v in tt.items()) + f" eig TF/s={flops*info['numops']/tt['eig']/1e12:.1f} mem={ctx.bytes_in_use()/1e9:.1f}GB")
    sys.stdout.flush()
