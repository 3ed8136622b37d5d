"""One full DMRG2 sweep (l->r, r->l) of the config-3(ii) chain at bond dimension D with the default (two-level Gram) SVD
route and with MPSK_FLAG_EXACT_SVD (cuSOLVER gesvd everywhere): energies of every update must agree to 1e-10 relative and
the Lanczos eigenvalues must decrease along the sweep.   python tools/sweep_svd_check.py [D] [L] [krylovdim]"""
import sys, time, json, warnings
sys.path.insert(0, ".")
import numpy as np
import mpskit_b200 as m
from mpskit_b200 import _lib
from mpskit_b200.workloads import gaussian_bond_space
D = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
L = int(sys.argv[2]) if len(sys.argv) > 2 else 32
kd = int(sys.argv[3]) if len(sys.argv) > 3 else 10
warnings.simplefilter("ignore")
res = {}
for name, flags in (("gram2", 0), ("gesvd", _lib.FLAG_EXACT_SVD)):
    ctx = m.Context(0, flags)
    p = m.Space((-1, 1), (1, 1))
    H = m.heisenberg_longrange_u1(L, r=6)
    psi = m.FiniteMPS.random(20240213 + 3000, p, L, [gaussian_bond_space(D, k % 2) for k in range(1, L)], ctx, on_device=True)
    log = []
    t0 = time.time()
    psi, envs, eps = m.find_groundstate(psi, H, m.DMRG2(tol=1e-12, maxiter=1, eigalg=m.Lanczos(tol=1e-10, krylovdim=kd, maxiter=1),
                                                        trscheme=m.truncdim(D)), log=log)
    ctx.synchronize()
    res[name] = dict(seconds=time.time() - t0, energies=[e["energy"] for e in log], eps=[e["eps_trunc"] for e in log],
                     dims=[sum(e["kept"].values()) for e in log])
    ctx.close()
a, b = np.array(res["gram2"]["energies"]), np.array(res["gesvd"]["energies"])
out = dict(D=D, L=L, krylovdim=kd, updates=len(a), seconds_gram2=res["gram2"]["seconds"], seconds_gesvd=res["gesvd"]["seconds"],
           max_rel_energy_diff=float(np.abs(a - b).max() / np.abs(a).max()), final_energy=float(a[-1]),
           max_eps_trunc_diff=float(np.abs(np.array(res["gram2"]["eps"]) - np.array(res["gesvd"]["eps"])).max()),
           kept_dims_identical=res["gram2"]["dims"] == res["gesvd"]["dims"],
           energy_increases_above_1e_9=int(np.sum(np.diff(a) > 1e-9 * np.abs(a).max())), max_bond_dim=int(max(res["gram2"]["dims"])))
print(json.dumps(out))
open("gpurun_out/sweep_svd_check.json", "w").write(json.dumps(out, indent=1))
