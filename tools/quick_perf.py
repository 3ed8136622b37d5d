"""quick matvec timing (development aid): python tools/quick_perf.py D [real]"""
import sys, time
sys.path.insert(0, ".")
import numpy as np
import mpskit_b200 as m
from mpskit_b200.workloads import cfg3_matvec
D = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
dtype = np.float64 if len(sys.argv) > 2 and sys.argv[2] == "real" else np.complex128
ctx = m.default_context()
t0 = time.time()
w = cfg3_matvec(ctx, D, dtype=dtype)
ctx.synchronize()
print("setup s", time.time() - t0, "V", w["V"], "flops", w["op"].flops, "bytes in use GB", ctx.bytes_in_use() / 1e9)
op, x = w["op"], w["x"]
y = x.zerovector()
for _ in range(3):
    op.apply_into(x, y)
ctx.synchronize()
n = 10
t0 = time.time()
for _ in range(n):
    op.apply_into(x, y)
ctx.synchronize()
dt = (time.time() - t0) / n
print(f"D={D} {dtype.__name__}: {dt*1e3:.3f} ms/matvec  {op.flops/dt/1e12:.2f} TFLOP/s  ({op.flops/dt/36.99e12*100:.1f}% of 36.99 TF DMMA peak)")
