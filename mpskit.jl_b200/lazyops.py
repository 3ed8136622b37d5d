"""Operator-level lazy algebra: MultipliedOperator (TimedOperator / UntimedOperator), LazySum and MultipleEnvironments --
mirrors of src/operators/multipliedoperator.jl:6-52, src/operators/lazysum.jl:16-57, src/environments/multipleenv.jl:1-15 and
the derivative constructors of src/algorithms/derivatives.jl:283-323.  The operators stay lazy all the way down: the local
operator of a sum is a LazySumOp of device effective Hamiltonians (one `mpsk_op` per term, each with its own environments),
evaluated at `t + dt/2` by `integrate` (integrators.jl:20-25) and iterated by the host-driven Krylov solvers (krylov.py)."""
from __future__ import annotations
import numpy as np
from .derivatives import MultipliedOp, LazySumOp, ddAC as _ddAC, ddAC2 as _ddAC2, ddC as _ddC
from .environments import environments as _finite_environments


def _environments(state, op):
    """environments(st, op): FinEnv for a FiniteMPS, MPOHamInfEnv for an InfiniteMPS (multipleenv.jl:13-26)"""
    if hasattr(op, "_environments"):
        return op._environments(state)
    from .infinite import InfiniteMPS, environments as _infinite_environments
    if isinstance(state, InfiniteMPS):
        return _infinite_environments(state, op)
    return _finite_environments(state, op)


class MultipliedOperator:
    """multipliedoperator.jl:6-9: f * op with f a number (UntimedOperator) or a function of time (TimedOperator)"""

    def __init__(self, op, f=1.0):
        self.op, self.f = op, f

    @property
    def istimed(self):
        return callable(self.f)

    def __call__(self, t=None):
        """multipliedoperator.jl:31-37: evaluate the coefficient (at time t for a TimedOperator)"""
        if self.istimed:
            if t is None:
                raise ValueError("attempting to evaluate a time-dependent operator without specifying a time")
            return MultipliedOperator(self.op, self.f(t))
        return self

    def __mul__(self, b):           # multipliedoperator.jl:40-46
        f = self.f
        if callable(b) and callable(f):
            return MultipliedOperator(self.op, lambda t: b(t) * f(t))
        if callable(b):
            return MultipliedOperator(self.op, lambda t: b(t) * f)
        if callable(f):
            return MultipliedOperator(self.op, lambda t: b * f(t))
        return MultipliedOperator(self.op, b * f)

    __rmul__ = __mul__

    def __add__(self, other):       # lazysum.jl:54-57
        return LazySum([self]) + other

    def _environments(self, state):                # multipliedoperator.jl:50-52
        return _environments(state, self.op)

    def _ddC(self, pos, state, envs):              # derivatives.jl:298-300
        return MultipliedOp(_ddC(pos, state, self.op, envs), self.f)

    def _ddAC(self, pos, state, envs):
        return MultipliedOp(_ddAC(pos, state, self.op, envs), self.f)

    def _ddAC2(self, pos, state, envs, x_like=None):
        return MultipliedOp(_ddAC2(pos, state, self.op, envs, x_like), self.f)


def TimedOperator(op, f=None):
    """multipliedoperator.jl:16-17"""
    return MultipliedOperator(op, f if f is not None else (lambda t: 1.0))


def UntimedOperator(op, f=1.0):
    """multipliedoperator.jl:24-25"""
    return MultipliedOperator(op, f)


class LazySum:
    """lazysum.jl:16-18: a sum of operators that is only carried out where it is needed"""

    def __init__(self, ops, fs=None):
        ops = list(ops) if isinstance(ops, (list, tuple)) else [ops]
        if fs is not None:                          # lazysum.jl:33
            ops = [MultipliedOperator(o, f) for o, f in zip(ops, fs)]
        self.ops = ops

    def __len__(self):
        return len(self.ops)

    def __getitem__(self, i):
        return self.ops[i]

    def __iter__(self):
        return iter(self.ops)

    @property
    def istimed(self):                              # lazysum.jl:28-29
        return any(getattr(o, "istimed", False) for o in self.ops)

    def __call__(self, t=None):
        """lazysum.jl:36-48: `H(t)` evaluates the time-dependent coefficients and returns a time-independent LazySum"""
        if self.istimed and t is None:
            raise ValueError("attempting to evaluate time-dependent LazySum without specifying a time")
        if not self.istimed and t is not None:
            raise ValueError("attempting to evaluate time-independent LazySum at time")
        return LazySum([o(t) if isinstance(o, MultipliedOperator) else o for o in self.ops])

    def sum(self):
        """Base.sum(x::LazySum) = safe_eval(x) (lazysum.jl:48): the sum carried out with the operators' own arithmetic
        (MPOHamiltonian: mpohamiltonian.jl:97-154); time-dependent terms must be evaluated first, H(t).sum()"""
        tot = None
        for o in self.ops:
            if isinstance(o, MultipliedOperator):
                if o.istimed:
                    raise ValueError("attempting to evaluate a time-dependent LazySum without specifying a time")
                o = o.f * o.op
            tot = o if tot is None else tot + o
        return tot

    def repeat(self, n):
        """Base.repeat(x::LazySum, n) = LazySum(repeat.(x, n)) (lazysum.jl:59)"""
        return LazySum([MultipliedOperator(o.op.repeat(n), o.f) if isinstance(o, MultipliedOperator) else o.repeat(n) for o in self.ops])

    def __add__(self, other):                       # lazysum.jl:50-56
        more = other.ops if isinstance(other, LazySum) else [other]
        return LazySum(self.ops + list(more))

    __radd__ = lambda self, other: LazySum([other]) + self

    def _environments(self, state):                 # multipleenv.jl:13-15
        return MultipleEnvironments(self, [_environments(state, o) for o in self.ops])

    # derivatives.jl:311-323
    def _ddC(self, pos, state, cache):
        return LazySumOp([_ddC(pos, state, o, e) for o, e in zip(self.ops, cache.envs)])

    def _ddAC(self, pos, state, cache):
        return LazySumOp([_ddAC(pos, state, o, e) for o, e in zip(self.ops, cache.envs)])

    def _ddAC2(self, pos, state, cache, x_like=None):
        return LazySumOp([_ddAC2(pos, state, o, e, x_like) for o, e in zip(self.ops, cache.envs)])


class MultipleEnvironments:
    """multipleenv.jl:1-11"""

    def __init__(self, opp, envs):
        self.opp, self.envs = opp, list(envs)

    def __len__(self):
        return len(self.envs)

    def __getitem__(self, i):
        return self.envs[i]

    def __iter__(self):
        return iter(self.envs)

    def drop_stale(self, state):
        for e in self.envs:
            e.drop_stale(state)

    def recalculate(self, state, *args):            # multipleenv.jl:52-57
        for e in self.envs:
            e.recalculate(state, *args)
        return self
