"""Multi-GPU sharding of the matvec by row panels of the left (bra) index: one process per GPU, NCCL over NVLink.

Rank r of N owns one contiguous row range per sector of the left bond (`panel_table`: an even split of every sector).  An operator built from the row
panel of GL computes exactly that panel of y = h(x): stage-A M rows, mix rows and stage-C N columns all split on the same
index, so there are no redundant flops and no partial sums; the panels are assembled with ONE allgather per matvec.
The Krylov vectors stay replicated (their BLAS-1 is ~1 % of a matvec), so every rank takes identical decisions."""
from __future__ import annotations
import ctypes as C
import numpy as np
from .spaces import Space
from .tensor import DeviceTensor


def panel_table(V: Space, nranks):
    """{(rank, sector index): (lo, hi)} -- bit-exact mirror of csrc/core.cu:panel_table: rank r owns rows
    [floor(n r/N), floor(n (r+1)/N)) of every sector, i.e. exactly 1/N of the flops of any operator."""
    out = {}
    for r in range(nranks):
        for j, n in enumerate(V.dims):
            lo, hi = (int(n) * r) // nranks, (int(n) * (r + 1)) // nranks
            if hi > lo:
                out[(r, j)] = (lo, hi)
    return out


def panel_bounds(V: Space, j, rank, nranks):
    return panel_table(V, nranks).get((rank, j), (0, 0))


def panel_space(V: Space, rank, nranks) -> Space:
    tab = panel_table(V, nranks)
    out = {}
    for j, c in enumerate(V.charges):
        lo, hi = tab.get((rank, j), (0, 0))
        if hi > lo:
            out[c] = hi - lo
    return Space.from_dict(out, V.modulus)


def row_panel(t: DeviceTensor, rank, nranks) -> DeviceTensor:
    h = C.c_void_p()
    t.ctx.check(t.ctx.lib.mpsk_tensor_row_panel(t.ctx.h, t.h, rank, nranks, C.byref(h)))
    return DeviceTensor.wrap(t.ctx, h, t.Vl.modulus)


def scatter_panel(panel: DeviceTensor, rank, nranks, full: DeviceTensor):
    full.ctx.check(full.ctx.lib.mpsk_tensor_scatter_panel(full.ctx.h, panel.h, rank, nranks, full.h))
    return full


def apply_local(op, x: DeviceTensor, y_panel: DeviceTensor):
    op.ctx.check(op.ctx.lib.mpsk_op_apply_local(op.ctx.h, op.h, x.h, y_panel.h))
    return y_panel


def init_communicator(ctx, rank, nranks, broadcast):
    """create the library's NCCL communicator; `broadcast(bytes_or_None) -> bytes` ships rank 0's unique id
    (e.g. through torch.distributed.broadcast_object_list)."""
    uid = None
    if rank == 0:
        buf = (C.c_uint8 * 128)()
        st = ctx.lib.mpsk_comm_unique_id(buf)
        if st != 0:
            raise RuntimeError("mpsk_comm_unique_id failed (libnccl.so.2 not loadable?)")
        uid = bytes(buf)
    uid = broadcast(uid)
    buf = (C.c_uint8 * 128)(*uid)
    ctx.check(ctx.lib.mpsk_comm_init(ctx.h, buf, rank, nranks))
    ctx.rank, ctx.nranks = rank, nranks      # derivative operators built on this context now shard their matvec
