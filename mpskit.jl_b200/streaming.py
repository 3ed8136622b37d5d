"""Streamed matvecs on host buffers (`mpsk_matvec_stream_*`): `y_i = h(x_i)` for independent vectors whose blocks live in
pinned host memory, with the H2D copy of the next vector, the matvec of the current one and the D2H copy of the previous
result overlapping.  Host-buffer form of `(h::MPO_∂∂AC2)(x)` (derivatives.jl:26-31) — what `bench.py` times as `e2e`."""
from __future__ import annotations
import ctypes as C
from . import _lib
from .tensor import block_table


class MatvecStream:
    def __init__(self, op, x_like, y_like=None, depth=2, order="C"):
        """op: an effective-Hamiltonian operator handle holder (`.h`), x_like / y_like: DeviceTensors giving the block
        structures of the inputs and outputs (row panels when the context has a communicator); order: "C" for row-major host
        blocks (numpy default), "F" for column-major ones (TensorKit's layout)."""
        if order not in ("C", "F"):
            raise ValueError("order must be 'C' or 'F'")
        self.order = order
        y_like = x_like if y_like is None else y_like
        self.ctx = x_like.ctx
        self._tx = block_table(x_like.Vl, x_like.mch, x_like.Vr)
        self._ty = block_table(y_like.Vl, y_like.mch, y_like.Vr)
        self._shape_x = [(x_like.Vl.dims[jl], x_like.Vr.dims[jr]) for (_, jl, jr) in self._tx]
        self._shape_y = [(y_like.Vl.dims[jl], y_like.Vr.dims[jr]) for (_, jl, jr) in self._ty]
        self.dtype = x_like.dtype
        self.h = C.c_void_p()
        self.ctx.check(self.ctx.lib.mpsk_matvec_stream_create(self.ctx.h, op.h, x_like.h, y_like.h, int(depth), 0 if order == "C" else 1, C.byref(self.h)))
        self._keep = []

    def _refs(self, blocks, table, shapes):
        refs = (_lib.c_blockref * max(len(table), 1))()
        for i, (m, jl, jr) in enumerate(table):
            b = blocks[(m, jl)]
            ok = b.flags.c_contiguous if self.order == "C" else b.flags.f_contiguous
            if b.dtype != self.dtype or b.shape != shapes[i] or not ok:
                raise _lib.SpaceMismatch(_lib.ESPACE, f"block {(m, jl)}: need a {self.order}-contiguous {self.dtype} array of shape {shapes[i]}")
            refs[i] = _lib.c_blockref(b.ctypes.data, shapes[i][1], 1) if self.order == "C" else _lib.c_blockref(b.ctypes.data, 1, shapes[i][0])
        return refs

    def submit(self, xblocks, yblocks):
        """enqueue y = h(x); xblocks / yblocks: {(m, l-sector): ndarray} (pinned for real overlap).  Returns immediately."""
        rx = self._refs(xblocks, self._tx, self._shape_x)
        ry = self._refs(yblocks, self._ty, self._shape_y)
        self._keep = (self._keep + [(rx, ry, xblocks, yblocks)])[-16:]
        self.ctx.check(self.ctx.lib.mpsk_matvec_stream_submit(self.ctx.h, self.h, rx, len(self._tx), ry, len(self._ty)))

    def wait(self):
        self.ctx.check(self.ctx.lib.mpsk_matvec_stream_wait(self.ctx.h, self.h))
        self._keep = []

    def free(self):
        if self.h:
            self.ctx.lib.mpsk_matvec_stream_free(self.ctx.h, self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass
