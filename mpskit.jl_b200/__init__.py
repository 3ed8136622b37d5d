"""mpskit_b200: B200-native (sm_100a) hot path of MPSKit.jl behind the reference's own operator API.

The numerical work lives in ``libmpskit_b200.so`` (csrc/, C ABI in include/mpskit_b200.h); this package is the
host-side mirror of the reference interface for that path (same names, argument meaning and error behaviour).
"""
from ._lib import (Context, default_context, MpskError, SpaceMismatch, load as load_library, LIB_PATH)  # noqa: F401
from .spaces import Space, fuse, infimum, dual  # noqa: F401
from .tensor import DeviceTensor, block_table  # noqa: F401
from .operators import (SparseMPOSlice, MPOHamiltonian, transverse_field_ising, heisenberg_u1,  # noqa: F401
                        heisenberg_longrange_u1, drop_symmetry, DenseMPO, dense_mpo_slice, classical_ising,
                        periodic_boundary_conditions, transverse_field_ising_z2, cylinder_bonds, heisenberg_bonds_u1,
                        j1j2_cylinder_u1, mpo_product)
from . import operators  # noqa: F401
from .derivatives import (MPO_ddC, MPO_ddAC, MPO_ddAC2, ddC, ddAC, ddAC2, LazySumOp, MultipliedOp,  # noqa: F401
                          LinearCombinationOp)
from .krylov import RecursiveVec, eigsolve_arnoldi, linsolve_gmres  # noqa: F401
from .fixedpoint import fixedpoint, integrate, linsolve, Lanczos, Arnoldi  # noqa: F401
from .finitemps import FiniteMPS, virtual_spaces, qrpos, lqpos  # noqa: F401
from .environments import FinEnv, environments  # noqa: F401
from .dmrg import (DMRG2, find_groundstate_inplace, expectation_value, tsvd, contract_ac2,  # noqa: F401
                   truncdim, truncerr, truncbelow, notrunc, tsvd_spectrum, variance)
from .algorithms import (DMRG, TDVP, TDVP2, find_groundstate, find_groundstate_dmrg_inplace, calc_galerkin,  # noqa: F401
                         timestep, timestep_inplace)
from .sharding import panel_bounds, panel_table, panel_space, row_panel, scatter_panel, apply_local, init_communicator  # noqa: F401
from .infinite import (InfiniteMPS, TransferMatrix, regularize, MPOHamInfEnv, VUMPS, find_groundstate_vumps,  # noqa: F401
                       uniform_leftorth, uniform_rightorth, regauge, IDMRG1, IDMRGEnv, find_groundstate_idmrg1, IDMRG2,
                       find_groundstate_idmrg2, bond_inverse)
from . import infinite  # noqa: F401
from .infinite import timestep_infinite, expectation_value_lazy  # noqa: F401
from .algorithms import time_evolve  # noqa: F401
from .statmech import (PerMPOInfEnv, leading_boundary, MultilineOp, MultilineEnv, leading_boundary_multiline,  # noqa: F401
                       expectation_value_multiline, fixedpoint_multiline)
from . import statmech  # noqa: F401
from .approximate import approximate, approximate_inplace, ac_proj, ac2_proj  # noqa: F401
from .streaming import MatvecStream  # noqa: F401
from .serialize import save_mps, load_mps, pack_tensor, unpack_tensor  # noqa: F401
from .changebonds import (SvdCut, svdcut_inplace, OptimalExpand, RandExpand, VUMPSSvdCut, changebonds, changebonds_inplace,  # noqa: F401
                          changebonds_infinite, changebonds_multiline, mpo_times_infinite, densempo_as_state, svdcut_densempo, dot_infinite, leftnull, rightnull, contract_adjoint_left, cat)
from .excitations import (ProjectionOperator, AC_EffProj, AC2_EffProj, LinearCombination, LazyLincoCache, FiniteExcited,  # noqa: F401
                          excitations, unit_mpo)
from .excitations import expectation_value  # noqa: F401,F811  (general form: MPOHamiltonian / LinearCombination / ProjectionOperator)
from .lazyops import MultipliedOperator, TimedOperator, UntimedOperator, LazySum, MultipleEnvironments  # noqa: F401
from . import toolbox  # noqa: F401
from .toolbox import (transfer_spectrum, marek_gap, correlation_length, entropy, entanglement_spectrum,  # noqa: F401
                      expectation_value_local, correlator, decompose_localmpo)
from .propagator import DynamicalDMRG, propagator  # noqa: F401
from .statmech import VOMPS, MixedPerMPOInfEnv, approximate_vomps, approximate_vomps_multiline  # noqa: F401
from .statmech import approximate_idmrg1, approximate_idmrg2  # noqa: F401
from .quasiparticle import QuasiparticleAnsatz, FiniteQP, excitations_qp  # noqa: F401
from .quasiparticle import InfiniteQP, excitations_qp_infinite  # noqa: F401
from .operators import make_time_mpo  # noqa: F401
from .quasiparticle import fidelity_susceptibility  # noqa: F401
from .algorithms import UnionAlg, default_groundstate_algorithm  # noqa: F401
from .iterlog import IterLog, VERBOSE_NONE, VERBOSE_WARN, VERBOSE_CONV, VERBOSE_ITER, VERBOSE_ALL  # noqa: F401
from .algorithms import exact_diagonalization  # noqa: F401
from .finitemps import left_virtualspace, right_virtualspace, physicalspace, max_Ds  # noqa: F401,E402
