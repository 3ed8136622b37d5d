"""short Lanczos run on the config-3(ii) operator (for ncu captures of the Krylov BLAS-1 kernels)
python tools/krylov_profile.py [D] [krylovdim]"""
import sys
sys.path.insert(0, ".")
import mpskit_b200 as m
from mpskit_b200.workloads import cfg3_matvec
D = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
kd = int(sys.argv[2]) if len(sys.argv) > 2 else 12
ctx = m.default_context()
w = cfg3_matvec(ctx, D)
import warnings; warnings.simplefilter("ignore")
val, vec, info = m.fixedpoint(w["op"], w["x"], "SR", m.Lanczos(tol=1e-10, krylovdim=kd, maxiter=1), return_info=True)
print("krylov profile", val, info, "nelem bytes", sum(b.nbytes for b in w["x"].to_blocks().values()))
