"""world_size-2 gloo test (CPU) of the host-side multi-GPU logic: the row-panel partition rule, the panel operators
(oracle restatement of what each rank computes) and the allgather assembly reproduce the full matvec."""
import os
import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle.spaces import Space
    from oracle.blocks import BlockTensor
    from oracle import mpo as ompo, contract as oc
    import mpskit_b200 as m
    h = ompo.heisenberg_longrange_u1(4, r=2)[0]
    Vl, Vr = Space((-3, -1, 1, 3), (5, 9, 8, 3)), Space((-5, -3, -1, 1, 3), (2, 7, 10, 6, 3))
    rng = np.random.default_rng(7)                                  # same seed on every rank: replicated inputs
    mch = [h.pch[a] + h.pch[b] for a in range(2) for b in range(2)]
    x = BlockTensor.random(rng, Vl, mch, Vr)
    GL = BlockTensor.random(rng, Vl, [-c for c in h.cl], Vl)
    GR = BlockTensor.random(rng, Vr, [c for c in h.cr], Vr)
    # this rank's row panel of GL (host mirror of the library's integer rule)
    dVl = m.Space(Vl.charges, Vl.dims)
    Vp = m.panel_space(dVl, rank, world)
    Vpo = Space(Vp.charges, Vp.dims)
    blocks = {}
    for (lev, jl), b in GL.blocks.items():
        q = Vl.charges[jl]
        if q not in Vp.charges:
            continue
        lo, hi = m.panel_bounds(dVl, jl, rank, world)
        blocks[(lev, Vp.charges.index(q))] = b[lo:hi]
    GLp = BlockTensor(Vpo, GL.mch, Vl, GL.dtype, blocks)
    yp = oc.dAC2_blocks(x, h, h, GLp, GR)
    gathered = [None] * world
    dist.all_gather_object(gathered, {k: v for k, v in yp.blocks.items()})       # the allgather of the panels
    y = BlockTensor(Vl, mch, Vr)
    for r, blk in enumerate(gathered):
        Vr_ = m.panel_space(dVl, r, world)
        for (mm, jp), v in blk.items():
            jl = Vl.charges.index(Vr_.charges[jp])
            lo, hi = m.panel_bounds(dVl, jl, r, world)
            y.blocks[(mm, jl)][lo:hi] = v
    ref = oc.dAC2_blocks(x, h, h, GL, GR)
    err = max(np.abs(y.blocks[k] - ref.blocks[k]).max() for k in ref.blocks)
    # replicated Krylov scalars: every rank must see the same numbers
    t = torch.tensor([float(sum(np.vdot(b, b).real for b in y.blocks.values()))], dtype=torch.float64)
    lst = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(lst, t)
    same = all(float(a) == float(lst[0]) for a in lst)
    if rank == 0:
        out.put((err, same))
    dist.destroy_process_group()


def _worker_updates(rank, world, port, out):
    """round 2 exchange steps: (i) the environment update sharded over the KET index of the new environment (every rank
    contracts with its panel of the ket tensor's free bond, allgather assembles the full environment, no partial sums);
    (ii) Krylov inner products on row-panel vectors = local partial sums + allreduce, bit-identical on every rank."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle.spaces import Space
    from oracle.blocks import BlockTensor
    from oracle import mpo as ompo, contract as oc
    import mpskit_b200 as m
    h = ompo.heisenberg_longrange_u1(4, r=2)[0]
    Vl, Vr = Space((-3, -1, 1, 3), (5, 9, 8, 3)), Space((-4, -2, 0, 2, 4), (2, 7, 10, 6, 3))
    rng = np.random.default_rng(11)
    A = BlockTensor.random(rng, Vl, h.pch, Vr).to_dense()
    GL = BlockTensor.random(rng, Vl, [-c for c in h.cl], Vl).to_dense()
    full = oc.transfer_left_sparse(GL, h, A, A)                       # GL'[a' k; b']
    # this rank's panel of the ket index b' (library rule, mirrored by mpskit_b200.panel_bounds on the right bond of A)
    dVr = m.Space(Vr.charges, Vr.dims)
    off = Vr.offsets()
    cols = np.concatenate([np.arange(off[j] + m.panel_bounds(dVr, j, rank, world)[0], off[j] + m.panel_bounds(dVr, j, rank, world)[1])
                           for j in range(len(Vr.dims))]).astype(int)
    part = oc.transfer_left_sparse(GL, h, A[:, :, cols], A)           # contracts only with the panel of A's free bond
    gathered = [None] * world
    dist.all_gather_object(gathered, (cols, part))
    asm = np.zeros_like(full)
    for c, p_ in gathered:
        asm[:, :, c] = p_
    err = float(np.abs(asm - full).max())
    # Krylov dots on row panels
    x = BlockTensor.random(rng, Vl, h.pch, Vr); y = BlockTensor.random(rng, Vl, h.pch, Vr)
    dVl = m.Space(Vl.charges, Vl.dims)
    loc = 0.0 + 0.0j
    for (mm, jl), b in x.blocks.items():
        lo, hi = m.panel_bounds(dVl, jl, rank, world)
        loc += np.vdot(b[lo:hi], y.blocks[(mm, jl)][lo:hi])
    t = torch.tensor([loc.real, loc.imag], dtype=torch.float64)
    dist.all_reduce(t)
    ref = sum(np.vdot(b, y.blocks[k]) for k, b in x.blocks.items())
    lst = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(lst, t)
    same = all(torch.equal(a, lst[0]) for a in lst)
    if rank == 0:
        out.put((err, abs(complex(t[0], t[1]) - ref), same))
    dist.destroy_process_group()


def test_sharded_environment_update_and_panel_dots_world2_gloo():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    world = 2
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker_updates, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    err, derr, same = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert err == 0.0 or err < 1e-12
    assert derr < 1e-11 and same


def test_row_panel_matvec_world2_gloo():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    world = 2
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    err, same = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert err < 1e-12 and same


def test_panel_rule_tiles_every_space():
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import mpskit_b200 as m
    from mpskit_b200.workloads import gaussian_bond_space
    spaces = [gaussian_bond_space(D, par) for D in (16, 256, 4096, 8192) for par in (0, 1)] + [m.Space((0,), (n,)) for n in (1, 5, 129, 1024)]
    for V in spaces:
        for N in (1, 2, 3, 4, 8):
            tab = m.panel_table(V, N)
            for j, n in enumerate(V.dims):
                rng = sorted(tab[(r, j)] for r in range(N) if (r, j) in tab)
                assert rng[0][0] == 0 and rng[-1][1] == n and all(rng[i][1] == rng[i + 1][0] for i in range(len(rng) - 1))
            assert sum(m.panel_space(V, r, N).dim for r in range(N)) == V.dim
    # every rank gets the same share (+-1 row) of every sector: exactly 1/N of the flops of any operator
    V = gaussian_bond_space(4096, 0)
    tab = m.panel_table(V, 8)
    for j, n in enumerate(V.dims):
        w = [tab[(r, j)][1] - tab[(r, j)][0] for r in range(8) if (r, j) in tab]
        assert max(w) - min(w) <= 1 or n < 8
