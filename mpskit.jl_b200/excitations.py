"""ProjectionOperator, AC_EffProj / AC2_EffProj, operator-level LinearCombination with its LazyLincoCache and the
FiniteExcited algorithm: mirrors of src/operators/projection.jl:8-10, src/algorithms/derivatives.jl:234-281,
src/environments/lazylincocache.jl:1-8, src/environments/FinEnv.jl:101-105 and
src/algorithms/excitation/dmrgexcitation.jl:1-40.

The effective projector of a ProjectionOperator is y = |p><p|x> with |p> = le * ket.AC[pos] * re
(derivatives.jl:258-263: v = <x|le a1 re>, y = conj(v) le a1 re); le / re are the overlap environments between the state
(bra) and the projected ket, i.e. a FinEnv over a one-level identity MPO with `above = ket`, and |p> is one application of
the rectangular MPO-free derivative on the device.  Sums of such operators with effective Hamiltonians are iterated by the
host-driven Krylov solvers over the device VectorInterface (krylov.py), the way KrylovKit drives the reference."""
from __future__ import annotations
from dataclasses import dataclass, field
import numpy as np
from .operators import SparseMPOSlice, MPOHamiltonian
from .derivatives import MPO_ddAC, MPO_ddAC2, LinearCombinationOp, ddAC as _ddAC, ddAC2 as _ddAC2, ddC as _ddC
from .environments import environments as _environments
from .dmrg import contract_ac2
from .algorithms import DMRG


def unit_mpo(psi):
    """the `fill(nothing, length(state))` operator of FinEnv.jl:104 as a one-level identity MPO"""
    return MPOHamiltonian([SparseMPOSlice([[1.0]], (0,), (0,), psi.pq, psi.pspace.modulus)])


class ProjectionOperator:
    """projection.jl:8-10: |ket><ket|"""

    def __init__(self, ket):
        self.ket = ket

    # FinEnv.jl:101-105
    def _environments(self, state):
        envs = _environments(state, unit_mpo(state), self.ket)
        envs.opp_proj = self
        return envs

    # derivatives.jl:275-281
    def _ddAC(self, pos, state, envs):
        return AC_EffProj(self.ket.AC[pos], envs.leftenv(pos, state), envs.rightenv(pos, state), envs.opp[pos])

    def _ddAC2(self, pos, state, envs, x_like=None):
        return AC2_EffProj(self.ket, self.ket.AC[pos], self.ket.AR[pos + 1], envs.leftenv(pos, state),
                           envs.rightenv(pos + 1, state), envs.opp[pos], envs.opp[pos + 1])


class _EffProj:
    def __call__(self, x, *_t):
        y = self.p.copy()
        y.scale(self.p.dot(x))            # |p><p|x>
        return y

    __mul__ = __call__

    def free(self):
        if self.p is not None:
            self.p.free()
        self.p = None


class AC_EffProj(_EffProj):
    """derivatives.jl:249-253, 258-263"""

    def __init__(self, a1, le, re, unit_slice):
        self.ctx = a1.ctx
        op = MPO_ddAC(self.ctx, unit_slice, le, re, a1)
        self.p = op(a1)
        op.free()


class AC2_EffProj(_EffProj):
    """derivatives.jl:254-257, 264-269"""

    def __init__(self, ket, a1, a2, le, re, unit1, unit2):
        self.ctx = a1.ctx
        x = contract_ac2(ket, a1, a2)
        op = MPO_ddAC2(self.ctx, unit1, unit2, le, re, x)
        self.p = op(x)
        op.free(); x.free()


class LinearCombination:
    """LinearCombination of operators on the chain, sum_n c_n O_n (src/operators/lazysum ... used by dmrgexcitation.jl:21)"""

    def __init__(self, opps, coeffs):
        self.opps, self.coeffs = tuple(opps), tuple(coeffs)

    def _environments(self, state):          # lazylincocache.jl:6-8
        return LazyLincoCache(self, tuple(_environments(state, o) for o in self.opps))

    # derivatives.jl:234-247
    def _ddC(self, pos, state, cache):
        return LinearCombinationOp([_ddC(pos, state, h, e) for h, e in zip(self.opps, cache.envs)], self.coeffs)

    def _ddAC(self, pos, state, cache):
        return LinearCombinationOp([_ddAC(pos, state, h, e) for h, e in zip(self.opps, cache.envs)], self.coeffs)

    def _ddAC2(self, pos, state, cache, x_like=None):
        return LinearCombinationOp([_ddAC2(pos, state, h, e, x_like) for h, e in zip(self.opps, cache.envs)], self.coeffs)


class LazyLincoCache:
    """lazylincocache.jl:1-4"""

    def __init__(self, opp, envs):
        self.opp, self.envs = opp, envs

    def drop_stale(self, state):
        for e in self.envs:
            e.drop_stale(state)


def expectation_value(psi, O, envs=None):
    """expval.jl:195-212: LinearCombination -> sum_n c_n <O_n>; ProjectionOperator -> per-site <AC|dAC|AC> / |AC_end|^2
    (every entry equals |<ket|psi>|^2 / <psi|psi>); MPOHamiltonian -> total energy (dmrg.expectation_value)."""
    from .dmrg import expectation_value as _ev
    envs = envs if envs is not None else _environments(psi, O)
    from .lazyops import LazySum, MultipliedOperator
    if isinstance(O, LazySum):                      # expval.jl:186-192
        return sum(np.sum(expectation_value(psi, o, e)) for o, e in zip(O.ops, envs.envs))
    if isinstance(O, MultipliedOperator):           # expval.jl:182-184 (UntimedOperator)
        if O.istimed:
            raise ValueError("expectation value of a time-dependent operator: evaluate it at a time first, H(t)")
        return O.f * np.sum(expectation_value(psi, O.op, envs))
    if isinstance(O, LinearCombination):
        return sum(c * np.sum(expectation_value(psi, o, e)) for c, o, e in zip(O.coeffs, O.opps, envs.envs))
    if isinstance(O, ProjectionOperator):
        L = len(psi)
        ens = np.zeros(L, dtype=np.complex128)
        for i in range(L):
            op = O._ddAC(i, psi, envs)
            ac = psi.AC[i]
            y = op(ac)
            ens[i] = ac.dot(y)
            y.free(); op.free()
        last = psi.AC[L - 1]
        return ens / last.norm() ** 2
    return _ev(psi, O, envs)


@dataclass
class FiniteExcited:
    """dmrgexcitation.jl:10-13"""
    gsalg: object = field(default_factory=DMRG)
    weight: float = 10.0


def excitations(H, alg: FiniteExcited, states, init=None, num=1, log=None):
    """excitations(H, alg::FiniteExcited, states; init, num) (dmrgexcitation.jl:16-40): the next `num` states above the given
    ones, each the ground state of H + weight * sum_k |psi_k><psi_k| found by alg.gsalg.  Returns (energies, states)."""
    from .algorithms import find_groundstate
    states = tuple(states) if isinstance(states, (tuple, list)) else (states,)
    if num == 0:
        return [], []
    init = init if init is not None else states[0].copy()
    super_op = LinearCombination((H,) + tuple(ProjectionOperator(s) for s in states), (1.0,) + tuple(alg.weight for _ in states))
    envs = _environments(init, super_op)
    ne, _, _ = find_groundstate(init, super_op, alg.gsalg, envs, log)
    ens, excis = excitations(H, alg, states + (ne,), init=init, num=num - 1, log=log)
    e = expectation_value(ne, H)
    ens.append(float(np.real(e)))
    excis.append(ne)
    return ens, excis
