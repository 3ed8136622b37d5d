"""Synthetic workloads of BASELINE.json's configs: bond-space recipes and random-init operators."""
from __future__ import annotations
import numpy as np
from .spaces import Space
from .tensor import DeviceTensor
from .operators import heisenberg_longrange_u1, heisenberg_u1
from .derivatives import MPO_ddAC2


def gaussian_bond_space(D, parity, sigma=1.5, qmax=5, modulus=0):
    """n_q = round(D g(q)/sum g), g = exp(-q^2 / 2 sigma^2), |q| <= qmax, fix-up on the q ~ 0 sector so that
    sum n_q = D.  Charges are 2q; parity 0: integer q, 1: half-integer q."""
    cs = [c for c in range(-2 * qmax, 2 * qmax + 1) if (c % 2) == parity]
    g = np.array([np.exp(-(c / 2) ** 2 / (2 * sigma ** 2)) for c in cs])
    n = np.rint(D * g / g.sum()).astype(int)
    n[int(np.argmin(np.abs(np.array(cs))))] += D - n.sum()
    return Space.from_dict({c: int(k) for c, k in zip(cs, n) if k > 0}, modulus)


def cfg3_matvec(ctx, D=4096, r=6, dtype=np.complex128, seed=20240213 + 3000, sigma=1.5, build_op=True):
    """config 3(ii): spin-1/2 U(1) chain, range-r Heisenberg FSM MPO (chi = 2 + 3r), centre-bond two-site
    effective Hamiltonian with random GL / GR / x of the right block structure."""
    H = heisenberg_longrange_u1(2, r=r)
    h = H[0]
    V = gaussian_bond_space(D, 0, sigma)
    p = h.pcharges
    mch = [p[a] + p[b] for a in range(len(p)) for b in range(len(p))]
    x = DeviceTensor.random(ctx, V, mch, V, seed, dtype)
    GL = DeviceTensor.random(ctx, V, [-c for c in h.domcharges], V, seed + 1, dtype)
    GR = DeviceTensor.random(ctx, V, [c for c in h.imcharges], V, seed + 2, dtype)
    op = MPO_ddAC2(ctx, h, h, GL, GR, x) if build_op else None
    return dict(op=op, x=x, GL=GL, GR=GR, h=h, V=V, mch=mch)
