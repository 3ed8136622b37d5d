#!/usr/bin/env python
"""Benchmark of the north-star hot path: the two-site effective-Hamiltonian (ddAC2) matvec of config 3(ii)
(finite spin-1/2 Heisenberg, U(1), D=4096, range-6 MPO with chi=20), complex128.

One step = one matvec y = H_eff x through the C ABI of libmpskit_b200 (3 kernel launches + 2 staging copies).
`value`: inputs resident in HBM.  `e2e`: host x -> upload -> matvec -> download y inside the timed region.
`--impl reference`: the CPU oracle port (numpy/OpenBLAS block-sparse restatement in the reference's own
contraction order) on the box's host cores; Julia is not available so the reference itself cannot run.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "ddAC2_matvec_fp64_tflops"
UNIT = "TFLOP/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--D", type=int, default=4096)
    ap.add_argument("--range", type=int, default=6, dest="r")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--sweep-steps", type=int, default=3, help="DMRG2 centre steps timed for the sweep estimate (0: skip)")
    return ap.parse_args()


def workload_name(D, r):
    return (f"config 3(ii): finite spin-1/2 Heisenberg L=100 U(1) DMRG2, centre-bond ddAC2 matvec, D={D}, "
            f"range-{r} MPO chi={2 + 3 * r}, sigma=1.5 sector recipe")


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                       "-i", str(index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush(); self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f.read().strip().splitlines():
            c = [t.strip() for t in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if sm:
            busy = [s for s in sm if s > 0.5 * max(sm)] or sm
            out = {"sm_mhz": statistics.median(busy), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        return out


def oracle_problem(w, d=2):
    """device inputs -> oracle BlockTensors (same numbers) + oracle slice"""
    from oracle.spaces import Space as OSpace
    from oracle.blocks import BlockTensor
    from oracle import mpo as ompo

    def conv(t):
        Vl = OSpace(t.Vl.charges, t.Vl.dims, t.Vl.modulus); Vr = OSpace(t.Vr.charges, t.Vr.dims, t.Vr.modulus)
        return BlockTensor(Vl, t.mch, Vr, t.dtype, t.to_blocks())
    h = w["h"]
    oh = ompo.SparseMPOSlice(h.Os, h.domcharges, h.imcharges, h.pcharges)
    return conv(w["x"]), conv(w["GL"]), conv(w["GR"]), oh


def cpu_matvec_time(x, GL, GR, oh, reps=1):
    from oracle.contract import dAC2_blocks
    best = 1e30
    for _ in range(reps):
        t0 = time.perf_counter()
        y = dAC2_blocks(x, oh, oh, GL, GR)
        best = min(best, time.perf_counter() - t0)
    return best, y


def run_reference(args):
    """CPU arm: the oracle port of the reference algorithm, all host threads (OpenBLAS), bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np
    from oracle.spaces import Space as OSpace
    from oracle.blocks import BlockTensor
    from oracle import mpo as ompo
    from oracle.contract import dAC2_blocks
    sys.path.insert(0, os.path.join(ROOT, "mpskit.jl_b200"))
    # bounded sample: largest D <= args.D of the same recipe whose (warmup + steps) matvecs fit in ~200 s of CPU time
    # (measured: 6.6 s per D=4096 matvec on 16 cores, ~cubic in D at these sizes)
    t4096 = 14.0 * 8.0 / max(os.cpu_count() or 8, 1)
    Ds = args.D
    while Ds > 512 and (args.steps + args.warmup) * t4096 * (Ds / 4096.0) ** 3 > 200.0:
        Ds //= 2
    cs = [c for c in range(-10, 11) if c % 2 == 0]
    g = np.array([np.exp(-(c / 2) ** 2 / (2 * 1.5 ** 2)) for c in cs])
    n = np.rint(Ds * g / g.sum()).astype(int)
    n[int(np.argmin(np.abs(np.array(cs))))] += Ds - n.sum()
    V = OSpace.from_dict({c: int(k) for c, k in zip(cs, n) if k > 0})
    H = ompo.heisenberg_longrange_u1(2, r=args.r)
    h = H[0]
    rng = np.random.default_rng(20240213 + 3000)
    mch = [h.pch[a] + h.pch[b] for a in range(2) for b in range(2)]
    x = BlockTensor.random(rng, V, mch, V)
    GL = BlockTensor.random(rng, V, [-c for c in h.cl], V)
    GR = BlockTensor.random(rng, V, [c for c in h.cr], V)
    flops = algorithmic_flops(V, h, mch)
    for _ in range(args.warmup):
        dAC2_blocks(x, h, h, GL, GR)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        dAC2_blocks(x, h, h, GL, GR)
    dt = (time.perf_counter() - t0) / max(args.steps, 1)
    val = flops / dt / 1e12
    sample = f"same workload recipe at D={Ds} (algorithmic flops {flops:.3e} per matvec), numpy/OpenBLAS block-sparse port in reference contraction order"
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "c128", "data": "synthetic", "config": {"workload": workload_name(args.D, args.r), "sample_D": Ds},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))


def algorithmic_flops(V, h, mch):
    """SURVEY.md 8d F_AC2 for square centre-bond spaces: optimal-order GEMM flops (complex: kappa = 8)."""
    q = dict(zip(V.charges, V.dims))
    cl = h.cl if hasattr(h, "cl") else h.domcharges
    cr = h.cr if hasattr(h, "cr") else h.imcharges
    F = 0.0
    for qb, nb in q.items():
        MA = sum(q.get(qb + c, 0) for c in cl)
        NA = sum(q.get(qb + s, 0) for s in mch)
        F += 8.0 * MA * NA * nb
    for qa, na in q.items():
        for s in mch:
            nz = q.get(qa + s, 0)
            if not nz:
                continue
            K = sum(q.get(qa + s - c, 0) for c in cr)
            F += 8.0 * nz * na * K
    return F


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)
    import numpy as np
    import torch
    import torch.distributed as dist
    import mpskit_b200 as m
    from mpskit_b200.workloads import cfg3_matvec

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: there is no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = m.Context(local)
    stream = torch.cuda.Stream()          # a real (non-default) stream: torch events and the library share it
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)

    # N > 1: strong scaling of ONE matvec: rank r owns the r-th row panel of the bra index (mpskit_b200/sharding.py);
    # GL is sharded, GR and the Krylov vector x are replicated, one NCCL allgather per matvec assembles y everywhere.
    w = cfg3_matvec(ctx, args.D, r=args.r, build_op=(world == 1))
    x = w["x"]
    if world == 1:
        op = w["op"]
        flops = op.flops
    else:
        def bcast(uid):
            box = [uid]
            dist.broadcast_object_list(box, src=0)
            return box[0]
        m.init_communicator(ctx, rank, world, bcast)
        op = m.MPO_ddAC2(ctx, w["h"], w["h"], w["GL"], w["GR"], x)    # row-panels GL for this rank (derivatives.py)
        w["GL"].free()
        t = torch.tensor([op.flops], device="cuda", dtype=torch.float64)
        dist.all_reduce(t)
        flops = float(t.item())            # sum of the panel flops == flops of the full operator (no redundant work)
    y = x.zerovector()
    for _ in range(max(args.warmup, 3)):
        op.apply_into(x, y)
    ctx.synchronize()
    sampler = ClockSampler(local) if rank == 0 else None
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    l0 = ctx.launch_count()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        op.apply_into(x, y)
    e1.record(stream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = e0.elapsed_time(e1) / args.steps
    launches = ctx.launch_count() - l0
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    clocks = sampler.stop() if sampler else None
    value = flops / (ms * 1e-3) / 1e12             # whole-job: one full matvec per step, whatever the number of ranks

    # per-kernel timing of the dominant kernel (grouped DMMA GEMM): time stage A + stage C launches alone
    e2e = None
    if not args.no_e2e:
        # every rank holds the (replicated) Krylov vector in pinned host memory: upload, (sharded) matvec, download
        xb = {k: torch.from_numpy(v).pin_memory().numpy() for k, v in x.to_blocks().items()}   # pinned host blocks
        yb = {k: torch.from_numpy(np.empty_like(v)).pin_memory().numpy() for k, v in xb.items()}
        bytes_x = sum(b.nbytes for b in xb.values())
        for _ in range(2):
            xd = m.DeviceTensor.from_blocks(ctx, x.Vl, x.mch, x.Vr, xb); op.apply_into(xd, y); y.to_blocks(yb); xd.free()
        nrep = max(3, min(args.steps, 10))
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(nrep):
            xd = m.DeviceTensor.from_blocks(ctx, x.Vl, x.mch, x.Vr, xb)
            op.apply_into(xd, y)
            y.to_blocks(yb)
            xd.free()
        dt = (time.perf_counter() - t0) / nrep
        if world > 1:
            t = torch.tensor([dt], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e = {"value": flops / dt / 1e12, "unit": UNIT, "h2d_bytes_per_step": bytes_x * world, "d2h_bytes_per_step": bytes_x * world,
               "ms_per_step": dt * 1e3}

    # roofline denominator: FP64 tensor (DMMA) peak is not in MEASURED_PEAKS.json -> measure cuBLAS zgemm live
    roofline = None
    if rank == 0 and world == 1:
        n = 4096
        a = torch.randn(n, n, dtype=torch.complex128, device="cuda"); b = torch.randn(n, n, dtype=torch.complex128, device="cuda")
        (a @ b); torch.cuda.synchronize()
        best = 1e30
        for _ in range(3):
            s0 = torch.cuda.Event(enable_timing=True); s1 = torch.cuda.Event(enable_timing=True)
            s0.record(); (a @ b); s1.record(); torch.cuda.synchronize()
            best = min(best, s0.elapsed_time(s1))
        peak = 8.0 * n ** 3 / (best * 1e-3) / 1e12
        del a, b
        # the two GEMM launches carry all the algorithmic flops; time them in isolation over the same inputs
        gm = gemm_only_ms(ctx, op, x, y, stream, torch, max(3, min(args.steps, 10))) if world == 1 else float("nan")
        roofline = {"bound": "tensor", "kernel": "gemm_dmma_kernel<complex> (stage A + stage C of one matvec)",
                    "achieved": flops / (gm * 1e-3) / 1e12, "peak": peak, "unit": "TFLOP/s",
                    "frac": flops / (gm * 1e-3) / 1e12 / peak,
                    "traffic": 55.85e9 if (args.D == 4096 and args.r == 6) else None,
                    "traffic_source": "ncu dram__bytes_read+write per matvec (stage A 42.89+3.56 GB, stage C 9.21+0.19 GB), profiles/r01_SUMMARY.md; algorithmic operand bytes 10.6 GB",
                    "peak_source": "live torch.matmul complex128 4096^3 (cuBLAS zgemm) on this GPU; DMMA microbenchmark "
                                   "profiles/r01_peak_fp64_microbench.json gives 36.99; MEASURED_PEAKS.json has no FP64 entry",
                    "gemm_ms_per_matvec": gm, "share_of_step": gm / ms}

    cpu = None
    if rank == 0 and not args.no_cpu_baseline and world == 1:
        try:
            ox, oGL, oGR, oh = oracle_problem(w)
            t, yo = cpu_matvec_time(ox, oGL, oGR, oh)
            yd = y.to_blocks()
            err = max(float(np.abs(yd[k] - v).max()) for k, v in yo.blocks.items()) / max(float(np.abs(v).max()) for v in yo.blocks.values())
            cpu = {"value": flops / t / 1e12, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                   "sample": f"one full-size matvec of the same inputs ({t:.1f} s), numpy/OpenBLAS block-sparse port in the reference's "
                             f"contraction order; max rel. deviation GPU vs CPU {err:.2e}"}
        except Exception as exc:       # the baseline is a report, never a reason to lose the GPU number
            cpu = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {exc!r}"}

    sweep = None
    if rank == 0 and world == 1 and args.sweep_steps > 0:
        for t in (w.get("op"), w.get("GL"), w.get("GR")):
            if t is not None:
                t.free()
        try:
            sweep = sweep_window(ctx, m, args.D, args.r, args.sweep_steps)
        except Exception as exc:       # secondary measurement: report the failure, keep the headline line
            sweep = {"error": repr(exc)}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "c128",
                "data": "synthetic",
                "config": {"workload": workload_name(args.D, args.r), "D": args.D, "chi": 2 + 3 * args.r, "flops_per_matvec": flops,
                           "l2": "inputs larger than L2 (x 192 MB, environments 2 x 0.97 GB, intermediates 7.6 GB)",
                           "parallelism": "single GPU" if world == 1 else f"{world} GPUs: row panels of the bra index, GL sharded, 1 NCCL allgather of y per matvec"},
                "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu, "dmrg2_sweep": sweep}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def sweep_window(ctx, m, D, r, nsteps):
    """DMRG2 sweep-time estimate: `nsteps` consecutive two-site updates at the centre of an L=36 chain of the same
    recipe (full bond dimension D there), each = ac2 build + environment update + Lanczos eigsolve (krylovdim 15, one
    cycle) + truncated SVD + state update; extrapolated to the 197 updates of a full L=100 sweep by the measured
    full-sweep profile (profiles/r01_full_sweep_D4096*.json: 150 of 197 updates run at full D)."""
    import warnings
    import numpy as np
    from mpskit_b200.workloads import gaussian_bond_space
    from mpskit_b200 import dmrg as dm
    L = 36
    p = m.Space((-1, 1), (1, 1))
    H = m.heisenberg_longrange_u1(L, r=r)
    psi = m.FiniteMPS.random(20240213 + 3000, p, L, [gaussian_bond_space(D, k % 2) for k in range(1, L)], ctx, on_device=True)
    envs = m.environments(psi, H)
    alg = m.Lanczos(tol=1e-10, krylovdim=15, maxiter=1)
    pos0 = L // 2 - nsteps // 2
    envs.leftenv(pos0, psi); envs.rightenv(pos0 + 1, psi)
    # one untimed step absorbs cuSOLVER / allocator warm-up
    times, matvecs = [], 0
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for step in range(nsteps + 1):
            pos = pos0 - 1 + step
            ctx.synchronize(); t0 = time.perf_counter()
            ac2 = dm.contract_ac2(psi, psi.AC[pos], psi.AR[pos + 1])
            h = m.ddAC2(pos, psi, H, envs, ac2)
            val, new, info = m.fixedpoint(h, ac2, "SR", alg, return_info=True)
            h.free()
            al, c, ar, eps = dm.tsvd(psi, new, m.truncdim(D))
            psi.AC[pos] = (al, c); psi.AC[pos + 1] = (c, ar)
            ac2.free(); new.free(); envs.drop_stale(psi)
            ctx.synchronize()
            if step > 0:
                times.append(time.perf_counter() - t0); matvecs += info["numops"]
    per = float(np.mean(times))
    return {"seconds_per_centre_update": per, "matvecs_per_update": matvecs / len(times), "timed_updates": len(times),
            "estimated_sweep_seconds_L100": per * 150 + 0.06 * per * 47,
            "note": "estimate = 150 full-D updates + 47 edge updates (~6 % of a centre update each, from the measured full sweep); "
                    "measured full sweeps are in profiles/"}


def gemm_only_ms(ctx, op, x, y, stream, torch, reps):
    """average duration of the GEMM launches of one matvec (CUDA events on the launching stream)."""
    import ctypes as C
    lib = ctx.lib
    if not hasattr(lib, "mpsk_op_time_gemms"):
        return float("nan")
    lib.mpsk_op_time_gemms.restype = C.c_int32
    lib.mpsk_op_time_gemms.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.POINTER(C.c_double)]
    out = C.c_double()
    ctx.check(lib.mpsk_op_time_gemms(ctx.h, op.h, reps, C.byref(out)))
    return out.value


if __name__ == "__main__":
    main()
