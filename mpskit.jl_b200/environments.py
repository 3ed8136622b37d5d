"""FinEnv: mirror of src/environments/FinEnv.jl:9-145 with device-resident environments.
Staleness is decided by object identity of ``AL[j]`` / ``AR[j]`` exactly like the reference's ``===``."""
from __future__ import annotations
import ctypes as C
import numpy as np
from .tensor import DeviceTensor


class _StoredPanel:
    """an environment that is not the active GL / GR of the sweep, kept as this rank's row panel (1/N of its memory per
    rank; multi-GPU only).  `full()` re-assembles it on every rank with one NCCL allgather."""

    def __init__(self, full: DeviceTensor):
        ctx = full.ctx
        self.ctx, self.Vl, self.mch, self.Vr, self.dtype = ctx, full.Vl, full.mch, full.Vr, full.dtype
        h = C.c_void_p()
        ctx.check(ctx.lib.mpsk_tensor_row_panel(ctx.h, full.h, ctx.rank, ctx.nranks, C.byref(h)))
        self.panel = DeviceTensor.wrap(ctx, h, full.Vl.modulus)

    def full(self) -> DeviceTensor:
        out = DeviceTensor.zeros(self.ctx, self.Vl, self.mch, self.Vr, self.dtype)
        self.ctx.check(self.ctx.lib.mpsk_tensor_allgather(self.ctx.h, self.panel.h, out.h))
        return out

    def free(self):
        self.panel.free()


class FinEnv:
    """FinEnv.jl:9-37.  Multi-GPU: the L+1 stored environments per side are kept SHARDED (row panels, `_StoredPanel`); only
    the environments next to the current centre are held in full on every rank (what the matvec and the next update read)."""

    def __init__(self, below, opp, leftstart, rightstart, above=None):
        self.above = above
        self.opp = opp
        L = len(below)
        self.ldependencies = [None] * L
        self.rdependencies = [None] * L
        self.leftenvs = [leftstart] + [None] * L
        self.rightenvs = [None] * L + [rightstart]
        self.shard_store = getattr(below.ctx, "nranks", 1) > 1

    def poison(self, ind):
        """FinEnv.jl:108-111"""
        self.ldependencies[ind] = None
        self.rdependencies[ind] = None

    def drop_stale(self, state):
        """free environments whose dependency is gone (the reference keeps them allocated until they are overwritten,
        FinEnv.jl:28-37; at D=4096 that is ~190 GB, so stale ones are released eagerly -- semantics are unchanged because a
        stale entry is always recomputed before use)."""
        L = len(state)
        for j in range(L):
            if self.rightenvs[j] is not None and self.rdependencies[j] is not state.ARs[j]:
                self.rightenvs[j].free(); self.rightenvs[j] = None; self.rdependencies[j] = None
            if self.leftenvs[j + 1] is not None and self.ldependencies[j] is not state.ALs[j]:
                self.leftenvs[j + 1].free(); self.leftenvs[j + 1] = None; self.ldependencies[j] = None

    # ---- sharded storage (multi-GPU) -------------------------------------------------------------
    @staticmethod
    def _materialise(store, idx):
        e = store[idx]
        if isinstance(e, _StoredPanel):
            full = e.full()
            e.free()
            store[idx] = full
        return store[idx]

    def _compress(self, store, keep):
        """turn every full environment of `store` except the indices in `keep` (and the trivial boundary vectors) into
        its row panel"""
        if not self.shard_store:
            return
        for i in range(1, len(store) - 1):
            e = store[i]
            if i in keep or e is None or isinstance(e, _StoredPanel):
                continue
            store[i] = _StoredPanel(e)
            e.free()

    def rightenv(self, ind, state):
        """FinEnv.jl:114-129: environment contractable with the tensor on site ind."""
        a = None
        for i in range(len(state) - 1, ind, -1):
            if state.AR[i] is not self.rdependencies[i]:
                a = i
                break
        if a is not None:
            ctx = state.ctx
            for j in range(a, ind, -1):
                above = state.AR[j] if self.above is None else self.above.AR[j]
                h = C.c_void_p()
                src = self._materialise(self.rightenvs, j + 1)
                ctx.check(ctx.lib.mpsk_transfer_right(ctx.h, self.opp[j].device(ctx), above.h, state.AR[j].h,
                                                      src.h, C.byref(h)))
                if isinstance(self.rightenvs[j], (DeviceTensor, _StoredPanel)):
                    self.rightenvs[j].free()
                self.rightenvs[j] = DeviceTensor.wrap(ctx, h, state.pspace.modulus)
                self.rdependencies[j] = state.AR[j]
                self._compress(self.rightenvs, (j, ind + 1))
        out = self._materialise(self.rightenvs, ind + 1)
        self._compress(self.rightenvs, (ind + 1, ind + 2))
        return out

    def leftenv(self, ind, state):
        """FinEnv.jl:131-145"""
        a = None
        for i in range(0, ind):
            if state.AL[i] is not self.ldependencies[i]:
                a = i
                break
        if a is not None:
            ctx = state.ctx
            for j in range(a, ind):
                above = state.AL[j] if self.above is None else self.above.AL[j]
                h = C.c_void_p()
                src = self._materialise(self.leftenvs, j)
                ctx.check(ctx.lib.mpsk_transfer_left(ctx.h, self.opp[j].device(ctx), above.h, state.AL[j].h,
                                                     src.h, C.byref(h)))
                if isinstance(self.leftenvs[j + 1], (DeviceTensor, _StoredPanel)):
                    self.leftenvs[j + 1].free()
                self.leftenvs[j + 1] = DeviceTensor.wrap(ctx, h, state.pspace.modulus)
                self.ldependencies[j] = state.AL[j]
                self._compress(self.leftenvs, (j + 1, ind))
        out = self._materialise(self.leftenvs, ind)
        self._compress(self.leftenvs, (ind, ind - 1))
        return out


def environments(below, O, above=None):
    """FinEnv.jl:41-70: start vectors l_LL (x) 1 on level 1 and r_RR (x) 1 on level odim (MPOHamiltonian).
    LinearCombination -> LazyLincoCache (lazylincocache.jl:6-8), ProjectionOperator -> overlap environments (FinEnv.jl:101-105)."""
    if hasattr(O, "_environments"):
        return O._environments(below)
    ctx = below.ctx
    L = len(below)
    odim = O.odim
    V0, VL = below.V[0], below.V[L]
    cl = O[0].domcharges
    cr = O[L - 1].imcharges
    dt = below.dtype
    ls = np.zeros((V0.dim, odim, V0.dim), dtype=dt); ls[:, 0, :] = np.eye(V0.dim)
    rs = np.zeros((VL.dim, odim, VL.dim), dtype=dt); rs[:, odim - 1, :] = np.eye(VL.dim)
    leftstart = DeviceTensor.from_dense(ctx, ls, V0, [-c for c in cl], V0, dt)
    rightstart = DeviceTensor.from_dense(ctx, rs, VL, [c for c in cr], VL, dt)
    return FinEnv(below, O, leftstart, rightstart, above)
