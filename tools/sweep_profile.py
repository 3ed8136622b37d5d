"""DMRG2 centre-update window of config 3(ii) with the library's phase profiler on (MPSK_PROFILE=1): prints the tsvd
level breakdown and the Krylov phase timers.   python tools/sweep_profile.py [D] [nsteps] [krylovdim] [maxiter]"""
import os, sys, time, json
os.environ["MPSK_PROFILE"] = "1"
sys.path.insert(0, ".")
import bench
import mpskit_b200 as m

class A:
    pass
a = A()
a.D = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
a.sweep_steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
a.sweep_krylovdim = int(sys.argv[3]) if len(sys.argv) > 3 else 30
a.sweep_maxiter = int(sys.argv[4]) if len(sys.argv) > 4 else 1
a.r = 6
ctx = m.default_context()
out = bench.sweep_window(ctx, m, a, 1, lambda v: v)
print(json.dumps(out))
