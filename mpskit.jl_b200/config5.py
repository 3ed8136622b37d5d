"""BASELINE.json config 5: J1-J2 Heisenberg (J2 = 0.5) on a width-6 cylinder, snake MPO with chi = 35 (U(1)), D = 8192 with
the sigma = 1.8 sector recipe, sharded over the GPUs of one box.  `bench.py --config 5` runs
  (1) the centre-bond ddAC2 matvec (strong scaling over the ranks, panel Krylov vectors),
  (2) a DMRG2 half-sweep window: consecutive two-site centre updates (dmrg.jl:80-104) on a 6 x 6 section of the cylinder
      (36 sites reach the full bond dimension at the centre; a 16-column half-sweep pays this per-update cost ~70 times),
  (3) a TDVP window: single-site integrations AC(dt/2) / C(-dt/2) of timestep! (tdvp.jl:61-73) on the same state, dt = 0.05.
Environments are stored as row panels (1/N per rank), the Krylov basis, the environment updates and the SVD sectors are
distributed; nothing of size chi D^2 is replicated except the active GR of the matvec."""
from __future__ import annotations
import json
import time
import warnings
import numpy as np


def run_config5(args, ctx, m, torch, dist, rank, world, local, stream, ClockSampler):
    from .workloads import gaussian_bond_space
    from . import dmrg as dm
    D, W = args.D, 6
    sigma = 1.8

    def allmax(v):
        if world == 1:
            return v
        t = torch.tensor([v], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(v):
        if world == 1:
            return v
        t = torch.tensor([v], device="cuda", dtype=torch.float64)
        dist.all_reduce(t)
        return float(t.item())

    # ---- (1) matvec at the centre of the 96-site snake MPO
    Hfull = m.j1j2_cylinder_u1(W, 16)
    h1, h2 = Hfull[47], Hfull[48]
    V = gaussian_bond_space(D, 0, sigma)
    p = h1.pcharges
    mch = [p[a] + p[b] for a in range(2) for b in range(2)]
    seed = 20240213 + 5000
    x = m.DeviceTensor.random(ctx, V, mch, V, seed)
    GL = m.DeviceTensor.random(ctx, V, [-c for c in h1.domcharges], V, seed + 1)
    GR = m.DeviceTensor.random(ctx, V, [c for c in h2.imcharges], V, seed + 2)
    op = m.MPO_ddAC2(ctx, h1, h2, GL, GR, x)
    GL.free()
    flops = allsum(op.flops) if world > 1 else op.flops
    xp = m.row_panel(x, rank, world) if world > 1 else x
    yp = xp.zerovector()
    xbytes = sum(b.nbytes for b in x.to_blocks().values()) if rank == 0 else 0

    def step():
        ctx.check(ctx.lib.mpsk_op_apply_panel(ctx.h, op.h, xp.h, yp.h))
    for _ in range(max(args.warmup, 3)):
        step()
    ctx.synchronize()
    sampler = ClockSampler(local) if rank == 0 else None
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    l0 = ctx.launch_count()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    torch.cuda.synchronize()
    ms = allmax(e0.elapsed_time(e1) / args.steps)
    launches = ctx.launch_count() - l0
    clocks = sampler.stop() if sampler else None
    mem_matvec = ctx.bytes_in_use()
    for t in (op, GR, xp, yp, x):
        try:
            t.free()
        except Exception:
            pass

    # ---- (2) DMRG2 half-sweep window on a 6 x 6 section
    Lx = 6
    L = W * Lx
    H = m.j1j2_cylinder_u1(W, Lx)
    pspace = m.Space((-1, 1), (1, 1))
    t0 = time.perf_counter()
    psi = m.FiniteMPS.random(seed, pspace, L, [gaussian_bond_space(D, k % 2, sigma) for k in range(1, L)], ctx, on_device=True)
    envs = m.environments(psi, H)
    nsteps = max(args.sweep_steps, 1)
    pos0 = L // 2 - nsteps // 2
    envs.leftenv(pos0 - 1, psi); envs.rightenv(pos0, psi)
    ctx.synchronize()
    prep = time.perf_counter() - t0
    alg = m.Lanczos(tol=1e-10, krylovdim=args.sweep_krylovdim, maxiter=args.sweep_maxiter)
    names = ("ac2", "env_op", "eigsolve", "tsvd", "update")
    phases = {k: 0.0 for k in names}
    times, matvecs, energies, fl_tot = [], 0, [], 0.0
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for stepi in range(nsteps + 1):                  # first update untimed (allocator / cuSOLVER warm-up)
            pos = pos0 - 1 + stepi
            ctx.synchronize(); t0 = time.perf_counter()
            ac2 = dm.contract_ac2(psi, psi.AC[pos], psi.AR[pos + 1])
            ctx.synchronize(); t1 = time.perf_counter()
            h = m.ddAC2(pos, psi, H, envs, ac2)
            ctx.synchronize(); t2 = time.perf_counter()
            val, new, info = m.fixedpoint(h, ac2, "SR", alg, return_info=True)
            ctx.synchronize(); t3 = time.perf_counter()
            fl = h.flops; h.free()
            al, c, ar, eps = dm.tsvd(psi, new, m.truncdim(D))
            ctx.synchronize(); t4 = time.perf_counter()
            psi.AC[pos] = (al, c); psi.AC[pos + 1] = (c, ar)
            ac2.free(); new.free(); envs.drop_stale(psi)
            ctx.synchronize(); t5 = time.perf_counter()
            if stepi > 0:
                times.append(t5 - t0); matvecs += info["numops"]; energies.append(val.real); fl_tot += fl * info["numops"]
                for k, dt_ in zip(names, (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4)):
                    phases[k] += dt_
    n = len(times)
    per = allmax(float(np.mean(times)))
    dmrg = {"seconds_per_centre_update": per, "timed_updates": n, "matvecs_per_update": matvecs / n,
            "phase_seconds_per_update": {k: v / n for k, v in phases.items()}, "energies": energies,
            "eigsolve_tflops": allsum(fl_tot) / max(allmax(phases["eigsolve"]), 1e-30) / 1e12,
            "state_and_environment_setup_seconds": prep, "krylovdim": args.sweep_krylovdim, "maxiter": args.sweep_maxiter,
            "estimated_half_sweep_seconds_L96": per * 70 + 0.06 * per * 25,
            "hbm_bytes_in_use_per_rank": ctx.bytes_in_use(), "section": f"{W} x {Lx} cylinder section, updates at sites {pos0}..{pos0 + n - 1}"}

    # ---- (3) TDVP window (tdvp.jl:65-73): AC forward dt/2, C backward dt/2, site after site
    dt = 0.05
    integ = m.Lanczos(tol=1e-10, krylovdim=30, maxiter=100)
    pos = pos0 + n
    times, mv = [], 0
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for stepi in range(nsteps + 1):
            i = pos + stepi
            ctx.synchronize(); t0 = time.perf_counter()
            hh = m.ddAC(i, psi, H, envs)
            newAC, inf1 = m.integrate(hh, psi.AC[i], 0.0, dt / 2, integ, return_info=True); hh.free()
            psi.AC[i] = newAC
            hh = m.ddC(i, psi, H, envs)
            newC, inf2 = m.integrate(hh, psi.CR[i], 0.0, -dt / 2, integ, return_info=True); hh.free()
            psi.CR[i] = newC
            envs.drop_stale(psi)
            ctx.synchronize(); t1 = time.perf_counter()
            if stepi > 0:
                times.append(t1 - t0); mv += inf1["numops"] + inf2["numops"]
    per_t = allmax(float(np.mean(times)))
    tdvp = {"seconds_per_site": per_t, "timed_sites": len(times), "matvecs_per_site": mv / len(times), "dt": dt,
            "estimated_timestep_seconds_L96": per_t * 2 * (70 + 0.06 * 25),
            "note": "one site = exponentiate(dAC, -i dt/2) + QR + exponentiate(dC, +i dt/2); a full timestep! visits every site twice"}

    if rank == 0:
        line = {"metric": "ddAC2_matvec_fp64_tflops", "value": flops / (ms * 1e-3) / 1e12, "unit": "TFLOP/s", "n_gpus": world,
                "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "c128", "data": "synthetic",
                "config": {"workload": f"config 5: J1-J2 (J2 = 0.5) Heisenberg on a width-6 cylinder, snake MPO chi = {h1.odim}, U(1), D = {D}, "
                                       f"sigma = {sigma} sector recipe, centre-bond ddAC2 matvec + DMRG2 half-sweep window + TDVP window",
                           "D": D, "chi": h1.odim, "flops_per_matvec": flops, "x_bytes": xbytes,
                           "parallelism": "single GPU" if world == 1 else f"{world} GPUs: row panels (GL, Krylov basis), environment "
                                                                          "updates over the ket index, SVD sectors over ranks, stored environments sharded"},
                "clocks": clocks, "gpu_launches": launches, "hbm_bytes_in_use_matvec_per_rank": mem_matvec,
                "dmrg2_half_sweep_window": dmrg, "tdvp_window": tdvp}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
