import torch, time
torch.manual_seed(0)
for n in (600, 1320, 1964):
    a = torch.randn(n, n, dtype=torch.complex128, device="cuda")
    for drv in ("gesvd", "gesvdj", "gesvda"):
        try:
            torch.linalg.svd(a, full_matrices=False, driver=drv); torch.cuda.synchronize()
            t0 = time.time(); u, s, vh = torch.linalg.svd(a, full_matrices=False, driver=drv); torch.cuda.synchronize()
            dt = time.time() - t0
            err = (u @ torch.diag(s).to(a.dtype) @ vh - a).abs().max().item()
            orth = (u.conj().T @ u - torch.eye(n, device="cuda", dtype=a.dtype)).abs().max().item()
            print(f"n={n} {drv}: {dt*1e3:.1f} ms  recon {err:.1e} orth {orth:.1e}")
        except Exception as e:
            print(n, drv, "failed", str(e)[:80])
    g = a.conj().T @ a
    torch.linalg.eigh(g); torch.cuda.synchronize()
    t0 = time.time(); w, v = torch.linalg.eigh(g); torch.cuda.synchronize()
    print(f"n={n} eigh(gram): {(time.time()-t0)*1e3:.1f} ms")
    t0 = time.time(); q, r = torch.linalg.qr(a); torch.cuda.synchronize()
    print(f"n={n} qr: {(time.time()-t0)*1e3:.1f} ms")
