#!/usr/bin/env python
"""Benchmark of the north-star hot path: the two-site effective-Hamiltonian (ddAC2) matvec of config 3(ii)
(finite spin-1/2 Heisenberg, U(1), D=4096, range-6 MPO with chi=20), complex128, and the DMRG2 centre
update built around it.

One step = one matvec y = H_eff x through the C ABI of libmpskit_b200.
  * N = 1: 3 kernel launches (grouped GEMM, mix, grouped GEMM) + 2 staging copies.
  * N > 1 (strong scaling, one process per GPU): the Krylov vector lives as row panels (rank r owns rows r/N of every
    sector of the left bond); one step = NCCL allgather of the x panels into the operator input + the rank-local 1/N of
    the three launches, y stays a panel (exactly the step the sharded eigensolver iterates).
`value`: inputs resident in HBM.  `e2e`: host buffers -> upload -> matvec -> download inside the timed region (N > 1:
every rank moves only its own panel).  `dmrg2_sweep`: timed DMRG2 centre updates (ac2 build, environment update, Lanczos
eigensolve with krylovdim 30 / tol 1e-10, truncated SVD, state update) on all ranks.
`--impl reference`: the CPU oracle port (numpy/OpenBLAS block-sparse restatement in the reference's own contraction order)
on the box's host cores with all threads; Julia is not available so the reference itself cannot run.
`--config 5`: the J1-J2 width-6 cylinder workload (D=8192, chi=32) -- see run_config5.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

if "--impl=reference" in sys.argv or ("--impl" in sys.argv and "reference" in sys.argv):
    # torchrun exports OMP_NUM_THREADS=1; the CPU arm must use every host core (OpenBLAS reads these at import time)
    for _k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[_k] = str(os.cpu_count() or 1)

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "ddAC2_matvec_fp64_tflops"
UNIT = "TFLOP/s"
HBM_PEAK_FALLBACK = 6553.9


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=3, choices=[3, 5])
    ap.add_argument("--D", type=int, default=None)
    ap.add_argument("--range", type=int, default=6, dest="r")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--sweep-steps", type=int, default=3, help="DMRG2 centre updates timed for the sweep figure (0: skip)")
    ap.add_argument("--sweep-krylovdim", type=int, default=30)
    ap.add_argument("--sweep-maxiter", type=int, default=3)
    a = ap.parse_args()
    if a.D is None:
        a.D = 4096 if a.config == 3 else 8192
    return a


def workload_name(D, r):
    return (f"config 3(ii): finite spin-1/2 Heisenberg L=100 U(1) DMRG2, centre-bond ddAC2 matvec, D={D}, "
            f"range-{r} MPO chi={2 + 3 * r}, sigma=1.5 sector recipe")


def bond_space_dims(D, parity=0, sigma=1.5, qmax=5):
    """SURVEY.md 8d config 3 recipe (pure python: both arms build the identical config dict from it)"""
    import math
    cs = [c for c in range(-2 * qmax, 2 * qmax + 1) if (c % 2) == parity]
    g = [math.exp(-(c / 2) ** 2 / (2 * sigma ** 2)) for c in cs]
    tot = sum(g)
    n = [int(round(D * x / tot)) for x in g]
    i0 = min(range(len(cs)), key=lambda i: abs(cs[i]))
    n[i0] += D - sum(n)
    return {c: k for c, k in zip(cs, n) if k > 0}


def algorithmic_flops_cfg3(D, r):
    """SURVEY.md 8d F_AC2 for the square centre-bond spaces of config 3(ii): optimal-order GEMM flops (complex: kappa = 8)."""
    q = bond_space_dims(D)
    qs = [2, -2, 0]
    cl = [0] + [qs[t] for t in range(3) for _ in range(r)] + [0]
    mch = [a + b for a in (-1, 1) for b in (-1, 1)]
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
            K = sum(q.get(qa + s - c, 0) for c in cl)
            F += 8.0 * nz * na * K
    return F


def config_dict(args):
    """identical in both arms (the driver compares them)"""
    n = args.gpus
    return {"workload": workload_name(args.D, args.r), "D": args.D, "chi": 2 + 3 * args.r,
            "flops_per_matvec": algorithmic_flops_cfg3(args.D, args.r),
            "l2": "inputs larger than L2 (x 192 MB, environments 2 x 0.97 GB, intermediates 7.6 GB at D=4096)",
            "parallelism": "single GPU" if n == 1 else f"{n} GPUs: row panels of the bra index (GL and the Krylov basis sharded), "
                                                        "1 NCCL allgather of x per matvec"}


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


# ------------------------------------------------------------------------------------------------ CPU arm
def oracle_problem(w):
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


def cpu_problem(D, r, seed=20240213 + 3000):
    import numpy as np
    from oracle.spaces import Space as OSpace
    from oracle.blocks import BlockTensor
    from oracle import mpo as ompo
    V = OSpace.from_dict(bond_space_dims(D))
    h = ompo.heisenberg_longrange_u1(2, r=r)[0]
    rng = np.random.default_rng(seed)
    mch = [h.pch[a] + h.pch[b] for a in range(2) for b in range(2)]
    x = BlockTensor.random(rng, V, mch, V)
    GL = BlockTensor.random(rng, V, [-c for c in h.cl], V)
    GR = BlockTensor.random(rng, V, [c for c in h.cr], V)
    return x, GL, GR, h


def run_reference(args):
    """CPU arm: the oracle port of the reference algorithm, all host threads (OpenBLAS), bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle.contract import dAC2_blocks
    ncores = os.cpu_count() or 1
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(ncores)
    except Exception:
        pass
    # bounded sample: calibrate on a small instance of the same recipe, then take the largest D <= args.D whose
    # (warmup + steps) matvecs fit in ~190 s of CPU time (cost grows ~ D^3)
    prob = cpu_problem(512, args.r)
    dAC2_blocks(prob[0], prob[3], prob[3], prob[1], prob[2])
    t0 = time.perf_counter()
    dAC2_blocks(prob[0], prob[3], prob[3], prob[1], prob[2])
    t512 = time.perf_counter() - t0
    nrun = args.steps + args.warmup
    Ds = args.D
    while Ds > 512 and nrun * t512 * (Ds / 512.0) ** 3 > 190.0:
        Ds //= 2
    x, GL, GR, h = cpu_problem(Ds, args.r) if Ds != 512 else prob
    flops = algorithmic_flops_cfg3(Ds, args.r)
    for _ in range(args.warmup):
        dAC2_blocks(x, h, h, GL, GR)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        dAC2_blocks(x, h, h, GL, GR)
    dt = (time.perf_counter() - t0) / max(args.steps, 1)
    val = flops / dt / 1e12
    sample = (f"same workload recipe at D={Ds} (algorithmic flops {flops:.3e} per matvec, {dt:.3f} s each), numpy/OpenBLAS "
              f"block-sparse port in reference contraction order, {ncores} threads")
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "c128", "data": "synthetic", "config": config_dict(args),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": ncores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))


def cpu_matvec_time(x, GL, GR, oh, reps=1):
    from oracle.contract import dAC2_blocks
    best = 1e30
    for _ in range(reps):
        t0 = time.perf_counter()
        y = dAC2_blocks(x, oh, oh, GL, GR)
        best = min(best, time.perf_counter() - t0)
    return best, y


# ------------------------------------------------------------------------------------------------ GPU arm
def hbm_peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    except Exception:
        return HBM_PEAK_FALLBACK, "fallback 6553.9 GB/s (MEASURED_PEAKS.json absent)"


def ncu_traffic(D, r):
    """per-matvec DRAM bytes of the GEMM launches from the committed ncu capture (profiles/r02_traffic.json)"""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
        return t.get(f"D{D}_r{r}"), t.get("source")
    except Exception:
        return None, None


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
    if world > 1:
        def bcast(uid):
            box = [uid]
            dist.broadcast_object_list(box, src=0)
            return box[0]
        m.init_communicator(ctx, rank, world, bcast)
    if args.config == 5:
        from mpskit_b200.config5 import run_config5
        return run_config5(args, ctx, m, torch, dist, rank, world, local, stream, ClockSampler)

    def allmax(v):
        if world == 1:
            return v
        t = torch.tensor([v], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # N > 1: strong scaling of ONE matvec: rank r owns the r-th row panel of the bra index (mpskit_b200/sharding.py);
    # GL and the Krylov vectors are sharded, GR is replicated, one NCCL allgather of the x panels per matvec.
    w = cfg3_matvec(ctx, args.D, r=args.r, build_op=False)
    x = w["x"]
    op = m.MPO_ddAC2(ctx, w["h"], w["h"], w["GL"], w["GR"], x)        # N > 1: row-panels GL for this rank (derivatives.py)
    cfg = config_dict(args)
    if world == 1:
        flops = op.flops
    else:
        w["GL"].free()
        t = torch.tensor([op.flops], device="cuda", dtype=torch.float64)
        dist.all_reduce(t)
        flops = float(t.item())            # sum of the panel flops == flops of the full operator (no redundant work)
    flops_check = "library flop count == SURVEY.md 8d formula" if abs(flops - cfg["flops_per_matvec"]) <= 1e-9 * flops else \
        f"MISMATCH: library {flops:.6e} vs SURVEY formula {cfg['flops_per_matvec']:.6e}"
    xp = m.row_panel(x, rank, world) if world > 1 else x             # this rank's Krylov vector
    yp = xp.zerovector()

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
    if world > 1:
        dist.barrier()
    ms = allmax(e0.elapsed_time(e1) / args.steps)
    launches = ctx.launch_count() - l0
    clocks = sampler.stop() if sampler else None
    value = flops / (ms * 1e-3) / 1e12             # whole-job: one full matvec per step, whatever the number of ranks

    e2e = None
    if not args.no_e2e:
        # every rank keeps ITS panel of the Krylov vector in pinned host memory: upload it, matvec (allgather inside), download
        # its panel of y.  Bytes over PCIe per step = |x| in and |y| out summed over the ranks.
        xb = {k: torch.from_numpy(v).pin_memory().numpy() for k, v in xp.to_blocks().items()}   # pinned host blocks
        yb = {k: torch.from_numpy(np.empty_like(v)).pin_memory().numpy() for k, v in xb.items()}
        bytes_x = sum(b.nbytes for b in xb.values())

        def e2e_step():
            xd = m.DeviceTensor.from_blocks(ctx, xp.Vl, xp.mch, xp.Vr, xb)
            ctx.check(ctx.lib.mpsk_op_apply_panel(ctx.h, op.h, xd.h, yp.h))
            yp.to_blocks(yb)
            xd.free()
        for _ in range(2):
            e2e_step()
        nrep = max(3, min(args.steps, 10))
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(nrep):
            e2e_step()
        torch.cuda.synchronize()
        dt = allmax((time.perf_counter() - t0) / nrep)
        tb = torch.tensor([float(bytes_x)], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tb)
        # the same through the streamed entry (mpsk_matvec_stream_*): independent matvecs, H2D of step i+1 / matvec of step i /
        # D2H of step i-1 overlap on the copy engines; every step still moves its x panel in and its y panel out
        yb2 = {k: torch.from_numpy(np.empty_like(v)).pin_memory().numpy() for k, v in xb.items()}
        ms_ = m.MatvecStream(op, xp, yp, depth=2)
        for i in range(3):
            ms_.submit(xb, yb2 if i % 2 else yb)
        ms_.wait()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for i in range(nrep):
            ms_.submit(xb, yb2 if i % 2 else yb)
        ms_.wait()
        torch.cuda.synchronize()
        dts = allmax((time.perf_counter() - t0) / nrep)
        ms_.free()
        e2e = {"value": flops / dts / 1e12, "unit": UNIT, "h2d_bytes_per_step": int(tb.item()), "d2h_bytes_per_step": int(tb.item()),
               "ms_per_step": dts * 1e3, "mode": "streamed: double-buffered async copies overlap the matvec of the neighbouring steps",
               "blocking_value": flops / dt / 1e12, "blocking_ms_per_step": dt * 1e3,
               "blocking_mode": "one blocking upload -> matvec -> download per step (no overlap)"}

    # roofline of the dominant kernel (grouped DMMA GEMM): FP64 tensor peak is not in MEASURED_PEAKS.json -> cuBLAS zgemm live
    roofline = None; roofline_mix = None
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
        peak3m = cublas_zgemm3m_tflops(torch, a, b)
        del a, b
        reps = max(3, min(args.steps, 10))
        gm = gemm_only_ms(ctx, op, reps)
        traffic, tsrc = ncu_traffic(args.D, args.r)
        ach = flops / (gm * 1e-3) / 1e12
        lib_peak = max(peak, peak3m or 0.0)
        roofline = {"bound": "tensor", "kernel": "gemm_dmma_kernel<complex, 3M> (stage A + stage C of one matvec)",
                    "achieved": ach, "peak": lib_peak, "unit": "TFLOP/s", "frac": ach / lib_peak, "traffic": traffic, "traffic_source": tsrc,
                    "peak_source": "builder-measured, live on this GPU in this run (MEASURED_PEAKS.json has no FP64 entry): the faster of "
                                   "cuBLAS zgemm (torch.matmul complex128 4096^3, conventional complex product) and cublasZgemm3m (the "
                                   "library's 3-multiplication complex GEMM), both in algorithmic flops 8 n^3; the DMMA issue "
                                   "microbenchmark profiles/r01_peak_fp64_microbench.json gives 36.99 executed TFLOP/s",
                    "note": "achieved counts ALGORITHMIC flops (8 per complex multiply-add, SURVEY.md 8d: independent of any 3M trick); "
                            "the kernel uses the 3-multiplication complex product and executes 0.75 x that many DMMA flops",
                    "cublas_zgemm_tflops": peak, "frac_of_cublas_zgemm": ach / peak,
                    "cublas_zgemm3m_tflops": peak3m, "frac_of_cublas_zgemm3m": (ach / peak3m) if peak3m else None,
                    "executed_tflops": 0.75 * ach, "dmma_pipe_frac": 0.75 * ach / 36.99,
                    "gemm_ms_per_matvec": gm, "share_of_step": gm / ms}
        mm, mb = mix_only_ms(ctx, op, reps)
        hp, hsrc = hbm_peak()
        roofline_mix = {"bound": "hbm", "kernel": "mix_kernel<double2>", "achieved": mb / (mm * 1e-3) / 1e9, "peak": hp, "unit": "GB/s",
                        "frac": mb / (mm * 1e-3) / 1e9 / hp, "bytes_per_launch": mb, "ms": mm, "peak_source": hsrc}

    cpu = None
    if rank == 0 and not args.no_cpu_baseline and world == 1:
        try:
            y = x.zerovector()
            op.apply_into(x, y)
            ox, oGL, oGR, oh = oracle_problem(w)
            t, yo = cpu_matvec_time(ox, oGL, oGR, oh)
            yd = y.to_blocks()
            err = max(float(np.abs(yd[k] - v).max()) for k, v in yo.blocks.items()) / max(float(np.abs(v).max()) for v in yo.blocks.values())
            y.free()
            cpu = {"value": flops / t / 1e12, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                   "sample": f"one full-size matvec of the same inputs ({t:.1f} s), numpy/OpenBLAS block-sparse port in the reference's "
                             f"contraction order; max rel. deviation GPU vs CPU {err:.2e}"}
        except Exception as exc:       # the baseline is a report, never a reason to lose the GPU number
            cpu = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {exc!r}"}

    sweep = None
    if args.sweep_steps > 0:
        for t in (op, w.get("GL"), w.get("GR"), xp, yp, x):
            try:
                if t is not None:
                    t.free()
            except Exception:
                pass
        try:
            sweep = sweep_window(ctx, m, args, world, allmax)
        except Exception as exc:       # secondary measurement: report the failure, keep the headline line
            if world > 1:
                raise                  # a rank that drops out of the collectives would hang the others
            sweep = {"error": repr(exc)}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "c128",
                "data": "synthetic", "config": cfg, "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline,
                "roofline_mix": roofline_mix, "cpu_baseline": cpu, "dmrg2_sweep": sweep, "flops_check": flops_check}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def sweep_window(ctx, m, args, world, allmax):
    """DMRG2 sweep figure: `sweep_steps` consecutive two-site updates at the centre of an L=36 chain of the same recipe
    (full bond dimension D there), each = ac2 build + environment update + operator plans + Lanczos eigensolve
    (krylovdim 30, tol 1e-10, thick restarts) + truncated SVD + state update; all ranks take part (sharded
    matvec / Krylov basis / environment updates / SVD sectors).  The L=100 figure extrapolates with the measured profile of
    a full sweep (profiles/r01_full_sweep_D4096*.json: 150 of 197 updates run at full D, edge updates ~6 % of a centre one)."""
    import warnings
    import numpy as np
    from mpskit_b200.workloads import gaussian_bond_space
    from mpskit_b200 import dmrg as dm
    D, r, nsteps = args.D, args.r, args.sweep_steps
    L = 36
    p = m.Space((-1, 1), (1, 1))
    H = m.heisenberg_longrange_u1(L, r=r)
    psi = m.FiniteMPS.random(20240213 + 3000, p, L, [gaussian_bond_space(D, k % 2) for k in range(1, L)], ctx, on_device=True)
    envs = m.environments(psi, H)
    alg = m.Lanczos(tol=1e-10, krylovdim=args.sweep_krylovdim, maxiter=args.sweep_maxiter)
    pos0 = L // 2 - nsteps // 2
    envs.leftenv(pos0, psi); envs.rightenv(pos0 + 1, psi)
    names = ("ac2", "env_op", "eigsolve", "tsvd", "update")
    phases = {k: 0.0 for k in names}
    times, matvecs, conv, energies, flops_mv = [], 0, 0, [], 0.0
    sub = {}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for step in range(nsteps + 1):      # one untimed update absorbs cuSOLVER / allocator warm-up
            pos = pos0 - 1 + step
            ctx.synchronize(); t0 = time.perf_counter()
            ac2 = dm.contract_ac2(psi, psi.AC[pos], psi.AR[pos + 1])
            ctx.synchronize(); t1 = time.perf_counter()
            gl = envs.leftenv(pos, psi)                      # environment update (sharded transfer + allgather)
            ctx.synchronize(); t1a = time.perf_counter()
            gr = envs.rightenv(pos + 1, psi)                 # stored right environment: panels -> full (allgather)
            ctx.synchronize(); t1b = time.perf_counter()
            h = m.MPO_ddAC2(ctx, envs.opp[pos], envs.opp[pos + 1], gl, gr, ac2)      # = ddAC2(pos, psi, H, envs) (derivatives.jl:55-58)
            ctx.synchronize(); t2 = time.perf_counter()
            val, new, info = m.fixedpoint(h, ac2, "SR", alg, return_info=True)
            ctx.synchronize(); t3 = time.perf_counter()
            fl = h.flops
            h.free()
            al, c, ar, eps = dm.tsvd(psi, new, m.truncdim(D))
            ctx.synchronize(); t4 = time.perf_counter()
            psi.AC[pos] = (al, c); psi.AC[pos + 1] = (c, ar)
            ac2.free(); new.free(); envs.drop_stale(psi)
            ctx.synchronize(); t5 = time.perf_counter()
            if step > 0:
                times.append(t5 - t0); matvecs += info["numops"]; conv += int(info["converged"]); energies.append(val.real)
                flops_mv += fl * info["numops"]
                for k, dt in zip(names, (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4)):
                    phases[k] += dt
                for k, dt in zip(("leftenv_update", "rightenv_activate", "operator_create"), (t1a - t1, t1b - t1a, t2 - t1b)):
                    sub[k] = sub.get(k, 0.0) + dt
    per = allmax(float(np.mean(times)))
    n = len(times)
    eig = allmax(phases["eigsolve"] / n)
    return {"seconds_per_centre_update": per, "matvecs_per_update": matvecs / n, "converged_updates": conv, "timed_updates": n,
            "krylovdim": args.sweep_krylovdim, "tol": 1e-10, "maxiter": args.sweep_maxiter,
            "phase_seconds_per_update": {k: v / n for k, v in phases.items()},
            "env_op_breakdown": {k: v / n for k, v in sub.items()}, "normres_last": info["normres"],
            "non_eigsolve_share": 1.0 - eig / per, "energies": energies,
            "eigsolve_tflops": flops_mv * world / max(phases["eigsolve"], 1e-30) / 1e12 if world > 1 else flops_mv / max(phases["eigsolve"], 1e-30) / 1e12,
            "estimated_sweep_seconds_L100": per * 150 + 0.06 * per * 47,
            "note": "estimate = 150 full-D updates + 47 edge updates (~6 % of a centre update each, from the measured full sweep); "
                    "per-update time is the max over ranks of the mean over the timed updates"}


def cublas_zgemm3m_tflops(torch, a, b):
    """live cublasZgemm3m (the library's own 3-multiplication complex GEMM) on the same 4096^3 problem, in algorithmic
    TFLOP/s (8 n^3): the like-for-like library number for the 3M kernel.  None if the entry point cannot be bound."""
    import ctypes as C
    try:
        lib = None
        for name in ("libcublas.so.12", "libcublas.so"):
            try:
                lib = C.CDLL(name)
                break
            except OSError:
                continue
        if lib is None:
            return None
        fn = lib.cublasZgemm3m
        fn.restype = C.c_int
        fn.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                       C.c_void_p, C.c_void_p, C.c_int]
        n = a.shape[0]
        c = torch.empty_like(a)
        handle = torch.cuda.current_blas_handle()
        alpha = (C.c_double * 2)(1.0, 0.0); beta = (C.c_double * 2)(0.0, 0.0)

        def run():
            st = fn(handle, 0, 0, n, n, n, alpha, a.data_ptr(), n, b.data_ptr(), n, beta, c.data_ptr(), n)
            if st != 0:
                raise RuntimeError(f"cublasZgemm3m status {st}")
        run(); torch.cuda.synchronize()
        best = 1e30
        for _ in range(3):
            s0 = torch.cuda.Event(enable_timing=True); s1 = torch.cuda.Event(enable_timing=True)
            s0.record(); run(); s1.record(); torch.cuda.synchronize()
            best = min(best, s0.elapsed_time(s1))
        ref = a.T @ b.T          # column-major C = A B on row-major storage: C^T = B^T A^T -> c (as row-major) = (a^T b^T)^T
        if not torch.allclose(c.T, ref, rtol=1e-9, atol=1e-9 * float(ref.abs().max())):
            return None
        return 8.0 * n ** 3 / (best * 1e-3) / 1e12
    except Exception:
        return None


def gemm_only_ms(ctx, op, reps):
    """average duration of the GEMM launches of one matvec (CUDA events on the launching stream, inside the library)."""
    import ctypes as C
    out = C.c_double()
    ctx.check(ctx.lib.mpsk_op_time_gemms(ctx.h, op.h, reps, C.byref(out)))
    return out.value


def mix_only_ms(ctx, op, reps):
    import ctypes as C
    ms, by = C.c_double(), C.c_double()
    ctx.check(ctx.lib.mpsk_op_time_mix(ctx.h, op.h, reps, C.byref(ms), C.byref(by)))
    return ms.value, by.value


if __name__ == "__main__":
    main()
