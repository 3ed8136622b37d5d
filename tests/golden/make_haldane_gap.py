"""Generates tests/golden/haldane_gap_oracle.json: the oracle (oracle/infinite.py VUMPS + oracle/quasiparticle.py infinite
quasiparticle ansatz) on the spin-1 Heisenberg chain at momentum pi -- the computation the reference's own test asserts,
test/algorithms.jl:204-211:  energies[1] ~ 0.41047925 atol = 1e-4  (there: two-site unit cell, D = 48 without symmetry; here:
U(1) symmetry, one-site unit cell, the S_z = 0 member of the triplet).  Run time on 8 CPU cores: ~40 s (D = 28), ~4.5 min (D = 48).
    python tests/golden/make_haldane_gap.py"""
import json
import os
import sys
import time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from oracle.spaces import Space
from oracle import mpo as ompo, infinite as oi, quasiparticle as oq

REFERENCE_TEST_VALUE = 0.41047925          # test/algorithms.jl:209


def run(dims, seed=3, vumps_tol=1e-10, qp_tol=1e-8):
    H = ompo.heisenberg_u1(1, d=3)
    p = Space((-2, 0, 2), (1, 1, 1))
    V = Space.from_dict(dims)
    t0 = time.time()
    st = oi.InfiniteMPS.random(np.random.default_rng(seed), p, [V])
    log = []
    st, envs, eps = oi.vumps(st, H, tol=vumps_tol, maxiter=400, log=log)
    en, _, _ = oq.excitations_infinite(H, np.pi, st, envs, num=1, tol=qp_tol)
    return dict(bond_space={str(k): v for k, v in dims.items()}, D=V.dim, vumps_iterations=len(log), galerkin=eps,
                energy_per_site=log[-1]["energy"], gap=en[0], abs_dev_from_reference_test=abs(en[0] - REFERENCE_TEST_VALUE),
                seconds=time.time() - t0)


if __name__ == "__main__":
    out = dict(reference_test="test/algorithms.jl:204-211", reference_value=REFERENCE_TEST_VALUE, atol=1e-4, runs=[])
    for dims in ({0: 8, 2: 6, -2: 6, 4: 3, -4: 3, 6: 1, -6: 1}, {0: 12, 2: 10, -2: 10, 4: 6, -4: 6, 6: 2, -6: 2}):
        out["runs"].append(run(dims))
        print(out["runs"][-1], flush=True)
    json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "haldane_gap_oracle.json"), "w"), indent=1)
