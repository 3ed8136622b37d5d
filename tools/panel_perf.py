"""per-rank matvec time of the row-panel decomposition, emulated rank by rank on ONE GPU (everything but the NCCL
allgather): python tools/panel_perf.py [D] [N]  -> time of the slowest emulated rank = the multi-GPU compute time"""
import sys, time, json
sys.path.insert(0, ".")
import ctypes as C
import numpy as np
import mpskit_b200 as m
from mpskit_b200.workloads import cfg3_matvec
D = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
N = int(sys.argv[2]) if len(sys.argv) > 2 else 8
ctx = m.default_context()
w = cfg3_matvec(ctx, D, build_op=False)
x, GL, GR, h = w["x"], w["GL"], w["GR"], w["h"]
out = {}
tot_flops = 0.0
for r in sorted(set([0, N // 2, N - 1])):
    GLp = m.row_panel(GL, r, N)
    op = m.MPO_ddAC2(ctx, h, h, GLp, GR, x)
    yp = m.DeviceTensor.zeros(ctx, GLp.Vl, x.mch, x.Vr)
    for _ in range(3):
        m.apply_local(op, x, yp)
    ctx.synchronize()
    n = 10
    t0 = time.perf_counter()
    for _ in range(n):
        m.apply_local(op, x, yp)
    ctx.synchronize()
    dt = (time.perf_counter() - t0) / n
    gm = C.c_double(); ctx.check(ctx.lib.mpsk_op_time_gemms(ctx.h, op.h, 5, C.byref(gm)))
    out[r] = dict(ms=dt * 1e3, gemm_ms=gm.value, flops=op.flops, tflops_rank=op.flops / dt / 1e12)
    op.free(); GLp.free(); yp.free()
full_flops = 3084892580736.0 if D == 4096 else None
worst = max(v["ms"] for v in out.values())
print(json.dumps(dict(D=D, N=N, per_rank=out, slowest_ms=worst,
                      implied_tflops=(full_flops / (worst * 1e-3) / 1e12 if full_flops else None))))
