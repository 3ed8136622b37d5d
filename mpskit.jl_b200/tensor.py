"""Device-resident block tensors T[l, m, r] (see include/mpskit_b200.h for the data model).

``DeviceTensor`` owns an ``mpsk_tensor*``.  Host data crosses the ABI as per-block pointers + strides:
numpy blocks here, TensorKit column-major sub-blocks in the Julia shim."""
from __future__ import annotations
import ctypes as C
import numpy as np
from . import _lib
from .spaces import Space


def block_table(Vl: Space, mch, Vr: Space):
    """[(m, l-sector, r-sector)] in ABI order (integer bookkeeping done by the library)."""
    lib = _lib.load()
    sl, kl = Vl.c_struct()
    sr, kr = Vr.c_struct()
    mc = (C.c_int32 * len(mch))(*[int(c) for c in mch])
    n = C.c_int32(0)
    st = lib.mpsk_block_table(C.byref(sl), len(mch), mc, C.byref(sr), C.byref(n), None, None, None)
    if st != _lib.OK:
        raise _lib.SpaceMismatch(st, "invalid spaces")
    m = (C.c_int32 * n.value)(); a = (C.c_int32 * n.value)(); b = (C.c_int32 * n.value)()
    lib.mpsk_block_table(C.byref(sl), len(mch), mc, C.byref(sr), C.byref(n), m, a, b)
    return [(m[i], a[i], b[i]) for i in range(n.value)]


class DeviceTensor:
    def __init__(self, ctx, handle, Vl, mch, Vr, dtype):
        self.ctx, self.h = ctx, handle
        self.Vl, self.Vr = Vl, Vr
        self.mch = tuple(int(c) for c in mch)
        self.dtype = np.dtype(dtype)

    # ---- construction -------------------------------------------------------------------------
    @staticmethod
    def _codes(dtype):
        dtype = np.dtype(dtype)
        if dtype == np.float64:
            return _lib.F64
        if dtype == np.complex128:
            return _lib.C128
        raise TypeError("only float64 / complex128 (Defaults.eltype, defaults.jl:18)")

    @staticmethod
    def from_blocks(ctx, Vl, mch, Vr, blocks, dtype=np.complex128):
        """blocks: {(m, l-sector index): ndarray (n_l, n_r)}; any strides."""
        dtype = np.dtype(dtype)
        table = block_table(Vl, mch, Vr)
        refs = (_lib.c_blockref * max(len(table), 1))()
        keep = []
        for i, (m, jl, jr) in enumerate(table):
            b = np.asarray(blocks[(m, jl)], dtype=dtype)
            if b.shape != (Vl.dims[jl], Vr.dims[jr]):
                raise _lib.SpaceMismatch(_lib.ESPACE, f"block {(m, jl)} has shape {b.shape}")
            keep.append(b)
            refs[i] = _lib.c_blockref(b.ctypes.data, b.strides[0] // b.itemsize, b.strides[1] // b.itemsize)
        sl, kl = Vl.c_struct(); sr, kr = Vr.c_struct()
        mc = (C.c_int32 * len(mch))(*[int(c) for c in mch])
        h = C.c_void_p()
        ctx.check(ctx.lib.mpsk_tensor_upload(ctx.h, C.byref(sl), len(mch), mc, C.byref(sr), DeviceTensor._codes(dtype),
                                             refs, len(table), C.byref(h)))
        return DeviceTensor(ctx, h, Vl, mch, Vr, dtype)

    @staticmethod
    def from_dense(ctx, arr, Vl, mch, Vr, dtype=None):
        """arr: (Dl, nm, Dr) dense embedding; only the symmetry-allowed blocks are read."""
        dtype = np.dtype(arr.dtype if dtype is None else dtype)
        lo, ro = Vl.offsets(), Vr.offsets()
        blocks = {(m, jl): arr[lo[jl]:lo[jl + 1], m, ro[jr]:ro[jr + 1]] for (m, jl, jr) in block_table(Vl, mch, Vr)}
        return DeviceTensor.from_blocks(ctx, Vl, mch, Vr, blocks, dtype)

    @staticmethod
    def zeros(ctx, Vl, mch, Vr, dtype=np.complex128):
        sl, kl = Vl.c_struct(); sr, kr = Vr.c_struct()
        mc = (C.c_int32 * len(mch))(*[int(c) for c in mch])
        h = C.c_void_p()
        ctx.check(ctx.lib.mpsk_tensor_zeros(ctx.h, C.byref(sl), len(mch), mc, C.byref(sr), DeviceTensor._codes(dtype), C.byref(h)))
        return DeviceTensor(ctx, h, Vl, mch, Vr, dtype)

    @staticmethod
    def random(ctx, Vl, mch, Vr, seed, dtype=np.complex128):
        """device-generated synthetic tensor (uniform(-1,1) in every allowed block), reproducible from seed"""
        t = DeviceTensor.zeros(ctx, Vl, mch, Vr, dtype)
        ctx.check(ctx.lib.mpsk_tensor_fill_random(ctx.h, t.h, int(seed)))
        return t

    @staticmethod
    def wrap(ctx, handle, modulus=0):
        """wrap a handle returned by the library, reading its structure back (two-phase query)."""
        nl, nr, nm, dt = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
        ctx.lib.mpsk_tensor_structure(handle, C.byref(nl), C.byref(nr), C.byref(nm), C.byref(dt))
        ql = (C.c_int32 * max(nl.value, 1))(); dl = (C.c_int64 * max(nl.value, 1))()
        qr = (C.c_int32 * max(nr.value, 1))(); dr = (C.c_int64 * max(nr.value, 1))()
        mc = (C.c_int32 * max(nm.value, 1))()
        ctx.lib.mpsk_tensor_spaces(handle, ql, dl, qr, dr, mc)
        Vl = Space(list(ql)[:nl.value], list(dl)[:nl.value], modulus)
        Vr = Space(list(qr)[:nr.value], list(dr)[:nr.value], modulus)
        return DeviceTensor(ctx, handle, Vl, list(mc)[:nm.value], Vr, np.complex128 if dt.value == _lib.C128 else np.float64)

    # ---- host transfer ------------------------------------------------------------------------
    def to_blocks(self, out=None):
        """download into {(m, l-sector): ndarray}; `out` may hold preallocated (e.g. pinned) blocks of any strides"""
        table = block_table(self.Vl, self.mch, self.Vr)
        out = {} if out is None else out
        refs = (_lib.c_blockref * max(len(table), 1))()
        for i, (m, jl, jr) in enumerate(table):
            b = out.get((m, jl))
            if b is None:
                b = np.empty((self.Vl.dims[jl], self.Vr.dims[jr]), dtype=self.dtype)
            out[(m, jl)] = b
            refs[i] = _lib.c_blockref(b.ctypes.data, b.strides[0] // b.itemsize, b.strides[1] // b.itemsize)
        self.ctx.check(self.ctx.lib.mpsk_tensor_download(self.ctx.h, self.h, refs, len(table)))
        return out

    def to_dense(self):
        out = np.zeros((self.Vl.dim, len(self.mch), self.Vr.dim), dtype=self.dtype)
        lo, ro = self.Vl.offsets(), self.Vr.offsets()
        blocks = self.to_blocks()
        for (m, jl, jr) in block_table(self.Vl, self.mch, self.Vr):
            out[lo[jl]:lo[jl + 1], m, ro[jr]:ro[jr + 1]] = blocks[(m, jl)]
        return out

    # ---- label algebra ----------------------------------------------------------------------------
    @staticmethod
    def stack(ts):
        ctx = ts[0].ctx
        arr = (C.c_void_p * len(ts))(*[t.h for t in ts])
        h = C.c_void_p()
        ctx.check(ctx.lib.mpsk_tensor_stack_labels(ctx.h, arr, len(ts), C.byref(h)))
        return DeviceTensor.wrap(ctx, h, ts[0].Vl.modulus)

    def select(self, labels):
        lab = (C.c_int32 * len(labels))(*[int(x) for x in labels])
        h = C.c_void_p()
        self.ctx.check(self.ctx.lib.mpsk_tensor_select_labels(self.ctx.h, self.h, lab, len(labels), C.byref(h)))
        return DeviceTensor.wrap(self.ctx, h, self.Vl.modulus)

    def adjoint(self):
        h = C.c_void_p()
        self.ctx.check(self.ctx.lib.mpsk_tensor_adjoint(self.ctx.h, self.h, C.byref(h)))
        return DeviceTensor.wrap(self.ctx, h, self.Vl.modulus)

    # ---- VectorInterface ----------------------------------------------------------------------
    def copy(self):
        h = C.c_void_p()
        self.ctx.check(self.ctx.lib.mpsk_tensor_copy(self.ctx.h, self.h, C.byref(h)))
        return DeviceTensor(self.ctx, h, self.Vl, self.mch, self.Vr, self.dtype)

    def zerovector(self):
        return DeviceTensor.zeros(self.ctx, self.Vl, self.mch, self.Vr, self.dtype)

    def dot(self, other):
        re, im = C.c_double(), C.c_double()
        self.ctx.check(self.ctx.lib.mpsk_dot(self.ctx.h, self.h, other.h, C.byref(re), C.byref(im)))
        return complex(re.value, im.value) if self.dtype.kind == "c" else re.value

    def norm(self):
        out = C.c_double()
        self.ctx.check(self.ctx.lib.mpsk_norm(self.ctx.h, self.h, C.byref(out)))
        return out.value

    def axpy(self, a, x):
        """self += a * x"""
        a = complex(a)
        self.ctx.check(self.ctx.lib.mpsk_axpy(self.ctx.h, a.real, a.imag, x.h, self.h))
        return self

    def scale(self, a):
        a = complex(a)
        self.ctx.check(self.ctx.lib.mpsk_scale(self.ctx.h, a.real, a.imag, self.h))
        return self

    def free(self):
        if self.h is not None and self.ctx.h:
            self.ctx.lib.mpsk_tensor_free(self.ctx.h, self.h)
        self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass
