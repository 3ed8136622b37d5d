"""BASELINE.json configs 2 and 4 at their named sizes on one B200 (timing + known-answer parity):
  config 2: infinite TFI (g = 0.5) VUMPS at D = 256 to convergence -> exact energy density (free fermions; the reference
            logs the same number, docs/src/examples/quantum1d/3.ising-dqpt/index.md:112-118), and spin-1 Heisenberg, 2-site unit
            cell, D = 256 -> Haldane-chain energy -1.401484038971 per site (docs/.../2.haldane/index.md:207-225 logs -1.401484038553)
  config 4: 2D classical Ising at beta_c, leading_boundary VUMPS at D = 1024 -> Onsager 2.5337372...
python tools/config24.py [D2] [D4] [iters4]  ->  one JSON line (also written to gpurun_out/config24.json)"""
import sys, time, json, warnings
sys.path.insert(0, ".")
import numpy as np
import mpskit_b200 as m
from mpskit_b200 import infinite as di, statmech as ds

D2 = int(sys.argv[1]) if len(sys.argv) > 1 else 256
D4 = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
it4 = int(sys.argv[3]) if len(sys.argv) > 3 else 6
ctx = m.default_context()
warnings.simplefilter("ignore")
out = {}

# ---- config 2a: TFI, unit cell 1
g = 0.5
k = np.linspace(0, np.pi, 400001)
f = np.sqrt(1 + g * g + 2 * g * np.cos(k))
e_exact = -np.sum((f[1:] + f[:-1]) / 2 * np.diff(k)) / np.pi
H = m.transverse_field_ising(1, g)
psi = m.InfiniteMPS.random(np.random.default_rng(3), m.Space((0,), (2,)), [m.Space((0,), (D2,))], ctx)
log = []
ctx.synchronize(); t0 = time.perf_counter(); l0 = ctx.launch_count()
psi, envs, eps = m.find_groundstate_vumps(psi, H, m.VUMPS(tol=1e-10, maxiter=40), log=log)
ctx.synchronize(); dt = time.perf_counter() - t0
out["config2_tfi"] = dict(D=D2, iterations=len(log), seconds=dt, seconds_per_iteration=dt / len(log), galerkin=eps, energy=log[-1]["energy"],
                          exact=e_exact, reference_logged=-1.063544409973, abs_err_vs_exact=abs(log[-1]["energy"] - e_exact),
                          kernel_launches=ctx.launch_count() - l0)

# ---- config 2b: spin-1 Heisenberg, 2-site unit cell, trivial symmetry (dense), fixed number of iterations
H1 = m.drop_symmetry(m.heisenberg_u1(2, d=3))
psi = m.InfiniteMPS.random(np.random.default_rng(4), m.Space((0,), (3,)), [m.Space((0,), (D2,)), m.Space((0,), (D2,))], ctx)
log = []
ctx.synchronize(); t0 = time.perf_counter()
psi, envs, eps = m.find_groundstate_vumps(psi, H1, m.VUMPS(tol=1e-9, maxiter=25), log=log)
ctx.synchronize(); dt = time.perf_counter() - t0
out["config2_spin1"] = dict(D=D2, unit_cell=2, iterations=len(log), seconds=dt, seconds_per_iteration=dt / len(log), galerkin=eps,
                            energy_per_site=log[-1]["energy"] / 2, haldane_chain=-1.401484038971,
                            abs_err=abs(log[-1]["energy"] / 2 + 1.401484038971))

# ---- config 4: classical Ising at beta_c, D = 1024
mpo = m.DenseMPO([m.classical_ising()])
psi = m.InfiniteMPS.random(np.random.default_rng(5), m.Space((0,), (2,)), [m.Space((0,), (D4,))], ctx)
envs4 = ds.environments(psi, mpo)
# matvec roofline of the dAC operator at this size
h = m.MPO_ddAC(ctx, mpo[0], envs4.leftenv(0, psi), envs4.rightenv(0, psi), psi.AC[0])
y = psi.AC[0].zerovector()
for _ in range(3):
    h.apply_into(psi.AC[0], y)
ctx.synchronize(); t0 = time.perf_counter()
for _ in range(20):
    h.apply_into(psi.AC[0], y)
ctx.synchronize(); mv = (time.perf_counter() - t0) / 20
fl = h.flops; h.free()
log = []
ctx.synchronize(); t0 = time.perf_counter()
psi, envs4, eps = m.leading_boundary(psi, mpo, m.VUMPS(tol=1e-8, maxiter=it4), envs=envs4, log=log)
ctx.synchronize(); dt = time.perf_counter() - t0
onsager = np.exp(2 * 0.915965594177219 / np.pi + np.log(2) / 2)
out["config4_ising"] = dict(D=D4, iterations=len(log), seconds=dt, seconds_per_iteration=dt / len(log), galerkin=eps,
                            value=log[-1]["value"][0].real, onsager=onsager, abs_err=abs(log[-1]["value"][0].real - onsager),
                            dAC_matvec_ms=mv * 1e3, dAC_flops=fl, dAC_tflops=fl / mv / 1e12,
                            history=[float(l["value"][0].real) for l in log])
print(json.dumps(out))
open("gpurun_out/config24.json", "w").write(json.dumps(out, indent=1))
