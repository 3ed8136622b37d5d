"""3M vs 4M complex product of the grouped DMMA GEMM on the config 3(ii) matvec (one GPU):
python tools/gemm3m_perf.py [D] [Npanels ...]  -> per mode: matvec ms, GEMM-only ms, TFLOP/s (algorithmic flops), the deviation
between the two modes, and the same for emulated row panels of an N-GPU run (narrow tile configurations)."""
import sys, time, json
sys.path.insert(0, ".")
import ctypes as C
import numpy as np
import mpskit_b200 as m
from mpskit_b200 import _lib
from mpskit_b200.workloads import cfg3_matvec
D = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
panels = [int(a) for a in sys.argv[2:]]
res = {"D": D}
ys = {}
for name, flags in (("3m", 0), ("4m", _lib.FLAG_GEMM_4M)):
    ctx = m.Context(0, flags)
    w = cfg3_matvec(ctx, D, build_op=False)
    x, GL, GR, h = w["x"], w["GL"], w["GR"], w["h"]
    op = m.MPO_ddAC2(ctx, h, h, GL, GR, x)
    y = x.zerovector()
    for _ in range(3):
        op.apply_into(x, y)
    ctx.synchronize()
    n = 10
    t0 = time.perf_counter()
    for _ in range(n):
        op.apply_into(x, y)
    ctx.synchronize()
    dt = (time.perf_counter() - t0) / n
    gm = C.c_double(); ctx.check(ctx.lib.mpsk_op_time_gemms(ctx.h, op.h, 5, C.byref(gm)))
    ys[name] = y.to_blocks()
    r = dict(ms=dt * 1e3, gemm_ms=gm.value, tflops=op.flops / dt / 1e12, gemm_tflops=op.flops / (gm.value * 1e-3) / 1e12)
    op.free(); y.free()
    for N in panels:
        worst = 0.0; wg = 0.0
        for rk in sorted(set([0, N // 2, N - 1])):
            GLp = m.row_panel(GL, rk, N)
            opp = m.MPO_ddAC2(ctx, h, h, GLp, GR, x)
            yp = m.DeviceTensor.zeros(ctx, GLp.Vl, x.mch, x.Vr)
            for _ in range(3):
                m.apply_local(opp, x, yp)
            ctx.synchronize()
            t0 = time.perf_counter()
            for _ in range(n):
                m.apply_local(opp, x, yp)
            ctx.synchronize()
            dtp = (time.perf_counter() - t0) / n
            g2 = C.c_double(); ctx.check(ctx.lib.mpsk_op_time_gemms(ctx.h, opp.h, 5, C.byref(g2)))
            worst = max(worst, dtp * 1e3); wg = max(wg, g2.value)
            opp.free(); GLp.free(); yp.free()
        r[f"panel_N{N}_slowest_ms"] = worst; r[f"panel_N{N}_gemm_ms"] = wg
    res[name] = r
    for t in (x, GL, GR):
        t.free()

num = max(float(np.abs(ys["3m"][k] - ys["4m"][k]).max()) for k in ys["3m"])
den = max(float(np.abs(v).max()) for v in ys["4m"].values())
res["max_rel_deviation_3m_vs_4m"] = num / den
res["speedup_gemm"] = res["4m"]["gemm_ms"] / res["3m"]["gemm_ms"]
print(json.dumps(res))
