"""Measure cuBLAS FP64 GEMM peaks (roofline denominators for the contraction kernels)."""
import json, time, torch
def bench(dtype, n, reps=10):
    a = torch.randn(n, n, device="cuda", dtype=torch.float64).to(dtype)
    b = torch.randn(n, n, device="cuda", dtype=torch.float64).to(dtype)
    c = a @ b
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); c = a @ b; e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    # sustained: 3 s back to back
    t0 = time.time(); k = 0
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    while time.time() - t0 < 3.0:
        for _ in range(5):
            c = a @ b
        k += 5
        torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    sus = e0.elapsed_time(e1) / k
    fl = (8.0 if dtype.is_complex else 2.0) * n ** 3
    return fl / best / 1e9, fl / sus / 1e9
out = {"gpu": torch.cuda.get_device_name(0)}
for n in (4096, 8192):
    b, s = bench(torch.float64, n); out[f"dgemm_{n}_tflops_burst"] = round(b, 2); out[f"dgemm_{n}_tflops_sustained"] = round(s, 2)
    b, s = bench(torch.complex128, n, reps=5); out[f"zgemm_{n}_tflops_burst"] = round(b, 2); out[f"zgemm_{n}_tflops_sustained"] = round(s, 2)
print(json.dumps(out))
