"""CPU oracle for the MPSKit.jl hot path -- TEST INFRASTRUCTURE ONLY.

This package is a numpy (float64/complex128, OpenBLAS) restatement of the reference algorithm
for the path named in BASELINE.json (effective-Hamiltonian matvecs, Krylov eigsolve /
exponentiate, transfer-matrix / environment updates, truncated SVD, DMRG/DMRG2/TDVP sweep
bodies) and of the callers built on it (VUMPS, IDMRG, statmech boundaries, changebonds, approximate,
dynamical DMRG, quasiparticle excitations, read-outs).  Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl
reference`` legs of ``bench.py`` may import it; the product package ``mpskit_b200`` never does.

Parity pinning.  The reference is pure Julia and no Julia toolchain exists in this image, so the
reference itself cannot be executed here and it ships no golden vectors for this path (its own
tests pin the path only to ~1e-5 through physics identities).  At the
north-star tolerance (1e-10) the oracle is therefore pinned by
  * the reference's known answers that do exist (test/algorithms.jl:185-201 classical Ising
    2.5337; test/algorithms.jl:204-211 Haldane gap 0.41047925 from the quasiparticle ansatz (oracle: 0.41047684 at
    D = 48, tests/golden/haldane_gap_oracle.json); test/algorithms.jl:512-540 DMRG == exact diagonalisation; the logged
    energies of docs/src/examples/quantum1d/1.ising-cft/index.md:330-362 and 3.ising-dqpt/index.md:42-48; the energies
    and entropies the reference printed for SU(2)-symmetric VUMPS states in 5.haldane-spt/index.md:385-404,1011-1019
    (E_plus, S_plus, E_minus, S_minus) and 4.xxz-heisenberg/index.md:671, reproduced to 7e-15 / 2e-11 / 4e-16 / 2e-12 / 13 digits by
    this oracle's VUMPS restricted to SU(2)-invariant tensors written in U(1) blocks, tests/golden/make_su2_pins.py),
  * analytic known answers (free-fermion transverse-field Ising spectra, exact diagonalisation),
  * two independent implementations inside the oracle cross-checking each other (dense einsum
    restatements vs. the block-sparse reference-order implementation).
"parity unpinned by reference-executed outputs": no output of the Julia code itself is available.

All file:line citations are relative to /root/reference/.
"""
from .spaces import Space, fuse, infimum, dual  # noqa: F401
