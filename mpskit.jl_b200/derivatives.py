"""Effective-Hamiltonian operators, mirror of src/algorithms/derivatives.jl:6-71.

``MPO_ddC / MPO_ddAC / MPO_ddAC2`` are callables ``h(x)`` (derivatives.jl:26-31) backed by an ``mpsk_op``;
``ddC / ddAC / ddAC2(pos, psi, H, envs)`` are the constructors (derivatives.jl:34-66), which -- as in the
reference -- trigger the lazy environment update through ``leftenv`` / ``rightenv``."""
from __future__ import annotations
import ctypes as C
from . import _lib
from .tensor import DeviceTensor


class DerivativeOperator:
    kind = None

    def __init__(self, ctx, h1, h2, leftenv, rightenv, x_like):
        self.ctx = ctx
        self.o1, self.o2 = h1, h2
        self._bra_l, self._bra_r = leftenv.Vl, rightenv.Vr          # structure of the output: y[bra_l, m, bra_r]
        kind = self.kind
        if getattr(ctx, "nranks", 1) > 1 and leftenv.Vl == x_like.Vl:
            kind |= 0x100                      # MPSK_OP_SHARDED
            # multi-GPU: this rank applies the row panel of the operator that belongs to its rows of the bra index;
            # apply / eigsolve / expv allgather the panels of y inside the library (sharding.py)
            h = C.c_void_p()
            ctx.check(ctx.lib.mpsk_tensor_row_panel(ctx.h, leftenv.h, ctx.rank, ctx.nranks, C.byref(h)))
            leftenv = DeviceTensor.wrap(ctx, h, leftenv.Vl.modulus)
        self.leftenv, self.rightenv = leftenv, rightenv
        op = C.c_void_p()
        ctx.check(ctx.lib.mpsk_op_create(ctx.h, kind,
                                         h1.device(ctx) if h1 is not None else None,
                                         h2.device(ctx) if h2 is not None else None,
                                         leftenv.h, rightenv.h, x_like.h, C.byref(op)))
        self.h = op

    def __call__(self, x, *_t):
        """y = h(x); the optional second argument is the (ignored) time of derivatives.jl:31.  Rectangular operators
        (bra spaces of the environments != bond spaces of x: ac_proj / ac2_proj of `approximate`) return a vector in the
        bra spaces."""
        if self._bra_l == x.Vl and self._bra_r == x.Vr:
            y = x.zerovector()
        else:
            y = DeviceTensor.zeros(self.ctx, self._bra_l, x.mch, self._bra_r, x.dtype)
        self.ctx.check(self.ctx.lib.mpsk_op_apply(self.ctx.h, self.h, x.h, y.h))
        return y

    def apply_into(self, x, y):
        self.ctx.check(self.ctx.lib.mpsk_op_apply(self.ctx.h, self.h, x.h, y.h))
        return y

    __mul__ = __call__

    @property
    def flops(self):
        """algorithmic real flops of one application."""
        return float(self.ctx.lib.mpsk_op_flops(self.h))

    def free(self):
        if self.h is not None and self.ctx.h:
            self.ctx.lib.mpsk_op_free(self.ctx.h, self.h)
        self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class MPO_ddC(DerivativeOperator):
    kind = _lib.OP_DC

    def __init__(self, ctx, leftenv, rightenv, x_like):
        super().__init__(ctx, None, None, leftenv, rightenv, x_like)


class MPO_ddAC(DerivativeOperator):
    kind = _lib.OP_DAC

    def __init__(self, ctx, o, leftenv, rightenv, x_like):
        super().__init__(ctx, o, None, leftenv, rightenv, x_like)


class MPO_ddAC2(DerivativeOperator):
    kind = _lib.OP_DAC2

    def __init__(self, ctx, o1, o2, leftenv, rightenv, x_like):
        super().__init__(ctx, o1, o2, leftenv, rightenv, x_like)


def ddC(pos, psi, H, envs):
    """derivatives.jl:34-36: MPO_ddC(leftenv(envs, pos+1), rightenv(envs, pos)); LinearCombination / ProjectionOperator
    dispatch (derivatives.jl:234-247, 275-281) lives with those operator types (excitations.py)"""
    if hasattr(H, "_ddC"):
        return H._ddC(pos, psi, envs)
    return MPO_ddC(psi.ctx, envs.leftenv(pos + 1, psi), envs.rightenv(pos, psi), psi.CR[pos])


def ddAC(pos, psi, H, envs):
    """derivatives.jl:44-46"""
    if hasattr(H, "_ddAC"):
        return H._ddAC(pos, psi, envs)
    return MPO_ddAC(psi.ctx, envs.opp[pos], envs.leftenv(pos, psi), envs.rightenv(pos, psi), psi.AC[pos])


def ddAC2(pos, psi, H, envs, x_like):
    """derivatives.jl:55-58: (opp[pos], opp[pos+1], leftenv(pos), rightenv(pos+1))"""
    if hasattr(H, "_ddAC2"):
        return H._ddAC2(pos, psi, envs, x_like)
    return MPO_ddAC2(psi.ctx, envs.opp[pos], envs.opp[pos + 1], envs.leftenv(pos, psi), envs.rightenv(pos + 1, psi), x_like)


# ---- operator algebra on derivatives (derivatives.jl:234-247, 283-323) ---------------------------------------------
class MultipliedOp:
    """MultipliedOperator of a derivative: f * h, with f a number or a function of time (derivatives.jl:311-323)."""

    def __init__(self, op, f):
        self.op, self.f, self.ctx = op, f, op.ctx

    def __call__(self, x, t=None):
        c = self.f(t) if callable(self.f) else self.f
        return self.op(x).scale(c)

    def free(self):
        if hasattr(self.op, "free"):
            self.op.free()


class LazySumOp:
    """LazySum of derivative operators: (sum_n h_n)(x) = sum_n h_n(x) (derivatives.jl:283-309)."""

    def __init__(self, ops):
        self.ops, self.ctx = list(ops), ops[0].ctx

    def __call__(self, x, t=None):
        y = self.ops[0](x, t) if t is not None else self.ops[0](x)
        for o in self.ops[1:]:
            z = o(x, t) if t is not None else o(x)
            y.axpy(1.0, z)
            z.free()
        return y

    __mul__ = __call__

    def free(self):
        for o in self.ops:
            if hasattr(o, "free"):
                o.free()


class LinearCombinationOp(LazySumOp):
    """LinearCombination: sum_n c_n h_n (derivatives.jl:234-247)."""

    def __init__(self, ops, coeffs):
        super().__init__([MultipliedOp(o, c) for o, c in zip(ops, coeffs)])
