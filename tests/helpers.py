"""shared builders for the parity tests: seeded inputs in oracle and device representations."""
import numpy as np
from oracle.spaces import Space as OSpace
from oracle.blocks import BlockTensor
from oracle import mpo as ompo
import mpskit_b200 as m


def to_dev_space(s: OSpace):
    return m.Space(s.charges, s.dims, s.modulus)


def dev_tensor(ctx, bt: BlockTensor):
    return m.DeviceTensor.from_blocks(ctx, to_dev_space(bt.Vl), bt.mch, to_dev_space(bt.Vr), bt.blocks, bt.dtype)


def dev_slice(os_: ompo.SparseMPOSlice):
    return m.SparseMPOSlice(os_.Os, os_.cl, os_.cr, os_.pch)


def dev_hamiltonian(H: ompo.MPOHamiltonian):
    return m.MPOHamiltonian([dev_slice(s) for s in H.slices])


def x2_dense(bt: BlockTensor, d):
    """[l,(p1 p2),r] block tensor -> dense x2[l,p1,r,p2]"""
    return bt.to_dense().reshape(bt.Vl.dim, d, d, bt.Vr.dim).transpose(0, 1, 3, 2)


def x2_from_dense(y, d):
    """dense x2[l,p1,r,p2] -> [l,(p1 p2),r]"""
    Dl, _, Dr, _ = y.shape
    return y.transpose(0, 1, 3, 2).reshape(Dl, d * d, Dr)


def random_problem(seed, h: ompo.SparseMPOSlice, Vl: OSpace, Vr: OSpace, kind="ac2", dtype=np.complex128):
    """random x, GL, GR with the right block structure for slice h (same slice on both sites)."""
    rng = np.random.default_rng(seed)
    d = h.d
    if kind == "ac2":
        mch = [h.pch[a] + h.pch[b] for a in range(d) for b in range(d)]
    elif kind == "ac":
        mch = list(h.pch)
    else:
        mch = [0]
    x = BlockTensor.random(rng, Vl, mch, Vr, dtype)
    GL = BlockTensor.random(rng, Vl, [-c for c in h.cl], Vl, dtype)
    GR = BlockTensor.random(rng, Vr, [c for c in h.cr], Vr, dtype)
    return x, GL, GR


def relerr(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))
