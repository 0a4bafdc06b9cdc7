"""
Multi-rank host logic on CPU (gloo, world_size 2 and 3): slab ownership,
neighbour ring, and the halo exchange plan, checked bit-exactly against
np.pad(mode='wrap') — the semantics of the reference's DMDA globalToLocal
(KSFD/ksfdsym.py:703-705, 919-920).  The data-path exchange on GPUs
(ncclSend/Recv in libksfd_b200.so) posts exactly the messages of
parallel.exchange_plan.
"""
import os
import socket

import numpy as np
import pytest


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _run_ranks(target, world, extra, wait=150):
    """spawn `world` ranks of `target(rank, world, port, *extra, q)`; a rendezvous that does not
    come up (port taken in between, slow start) is retried once on another port; ranks that
    are still alive at the end are terminated, never left behind"""
    import queue
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    for attempt in range(2):
        q = ctx.Queue()
        port = _free_port()
        procs = [ctx.Process(target=target, args=(r, world, port) + tuple(extra) + (q,))
                 for r in range(world)]
        for p in procs:
            p.start()
        res = []
        try:
            for _ in range(world):
                res.append(q.get(timeout=wait))
        except queue.Empty:
            res = None
        finally:
            for p in procs:
                p.join(timeout=30 if res is not None else 1)
                if p.is_alive():
                    p.terminate()
                    p.join(timeout=10)
        if res is not None:
            return res
    raise AssertionError('ranks did not report (twice)')


def _worker(rank, world, port, n, dof, q):
    import torch.distributed as dist
    import datetime
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    os.environ.setdefault('GLOO_SOCKET_IFNAME', 'lo')       # the container hostname may not resolve
    dist.init_process_group('gloo', rank=rank, world_size=world,
                            timeout=datetime.timedelta(seconds=60))
    try:
        from ksfd_b200.core import dmda_ownership
        from ksfd_b200.grid import Comm, Grid
        from ksfd_b200 import parallel
        glob = np.arange(dof * int(np.prod(n)), dtype=float).reshape((dof,) + tuple(n), order='F')
        start, count = dmda_ownership(n[-1], world)[rank]
        local = glob[..., start:start + count]
        lo, hi = parallel.host_halo_exchange(local, n[-1])
        pad = np.pad(glob, [(0, 0)] * (glob.ndim - 1) + [(2, 2)], mode='wrap')
        ok = (np.array_equal(lo, pad[..., start:start + 2]) and
              np.array_equal(hi, pad[..., start + count + 2:start + count + 4]))
        # the index description agrees with what was received
        slo, shi = parallel.ghost_sources(n[-1], world, rank)
        ok = ok and np.array_equal(lo, glob[..., slo]) and np.array_equal(hi, glob[..., shi])
        # Grid + DMDA.globalToLocal over ranks == wrap pad of the global array
        kw = dict(dim=len(n), dof=dof, nx=n[0])
        if len(n) > 1:
            kw['ny'] = n[1]
        if len(n) > 2:
            kw['nz'] = n[2]
        g = Grid(comm=Comm(rank, world), **kw)
        ok = ok and g.ranges[-1] == (start, start + count)
        v = g.Vdmda.createGlobalVec()
        v.array = local.reshape(-1, order='F')
        l = g.Vdmda.createLocalVec()
        g.Vdmda.globalToLocal(v, l)
        full = np.pad(glob, [(0, 0)] + [(2, 2)] * len(n), mode='wrap')
        want = full[..., start:start + count + 4]
        ok = ok and np.array_equal(l.array.reshape(g.Vashape, order='F'), want)
        c = Comm(rank, world)
        ok = ok and c.allreduce(float(rank + 1), 'sum') == world * (world + 1) / 2
        ok = ok and c.allreduce(float(rank), 'max') == world - 1
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world,n,dof', [(2, (6, 10), 3), (3, (5, 4, 11), 2), (2, (9,), 3)])
def test_slab_halo_exchange_gloo(world, n, dof):
    res = _run_ranks(_worker, world, (n, dof))
    assert sorted(res) == [(r, True) for r in range(world)]


def test_exchange_plan_and_ring():
    from ksfd_b200 import parallel
    assert parallel.neighbours(0, 4) == (3, 1) and parallel.neighbours(3, 4) == (2, 0)
    plan = parallel.exchange_plan(1024, 8, 0)
    assert plan == [('send', 1, (126, 128)), ('recv', 7, 'lo'),
                    ('send', 7, (0, 2)), ('recv', 1, 'hi')]
    lo, hi = parallel.ghost_sources(1024, 8, 0)
    assert lo == [1022, 1023] and hi == [128, 129]
    lo, hi = parallel.ghost_sources(10, 3, 2)       # 4,3,3 split
    assert lo == [5, 6] and hi == [0, 1]


def _ic_worker(rank, world, port, tmp, q):
    """start_values / resume_values on `world` ranks against the single-rank run
    (the advertised torchrun command could not start a fresh run)."""
    import torch.distributed as dist
    import datetime
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    os.environ.setdefault('GLOO_SOCKET_IFNAME', 'lo')       # the container hostname may not resolve
    dist.init_process_group('gloo', rank=rank, world_size=world,
                            timeout=datetime.timedelta(seconds=60))
    try:
        from ksfd_b200 import solver
        from ksfd_b200.grid import Comm, Grid
        from ksfd_b200.params import SolutionParameters, parse_commandline
        from ksfd_b200.random import Generator
        from ksfd_b200.timeseries import TimeSeries
        here = os.path.dirname(os.path.abspath(__file__))
        argv = ['@' + os.path.join(here, 'options', 'options84.args'), '--seed', '793817931',
                'nwidth=32', 'nheight=24']
        cl = parse_commandline(argv)
        ps = SolutionParameters(cl)

        def grid_for(comm):
            return Grid(dim=ps.dim, dof=ps.nligands + 1, width=ps.width, height=ps.height,
                        depth=ps.depth, nx=ps.nwidth, ny=ps.nheight, nz=ps.ndepth, comm=comm)

        # the single-rank field, computed on every rank
        Generator(seed=cl.seed, comm=Comm(0, 1))
        g1 = grid_for(Comm(0, 1))
        v1, _ = solver.start_values(cl, g1, ps)
        full = np.asarray(v1.array).reshape(g1.Vlshape, order='F')
        # the distributed field
        comm = Comm(rank, world)
        Generator(seed=cl.seed, comm=comm)
        g = grid_for(comm)
        v, t0 = solver.start_values(cl, g, ps)
        lo, hi = g.ranges[-1]
        mine = np.asarray(v.array).reshape(g.Vlshape, order='F')
        ok = np.array_equal(mine, full[..., lo:hi]) and mine.shape[-1] == hi - lo
        # resume from a sequential (s1r0) file: every rank keeps its slab
        prefix = os.path.join(tmp, 'seq')
        if rank == 0:
            ts = TimeSeries(prefix, grid=g1, mode='w', comm=Comm(0, 1))
            ts.store(v1, 0.5)
            ts.store(v1, 1.5)
            ts.close()
        comm.Barrier()
        cl2 = parse_commandline(argv + ['--resume', prefix])
        ps2 = SolutionParameters(cl2)
        rv, t = solver.resume_values(cl2, g, ps2)
        ok = ok and t == 1.5 and np.array_equal(
            np.asarray(rv.array).reshape(g.Vlshape, order='F'), full[..., lo:hi])
        # vals given on a distributed coarse grid are gathered
        from ksfd_b200.random import random_function
        rg = Grid(dim=2, nx=8, ny=8, dof=1, comm=comm)
        rg1 = rg.serial()
        gv = rg1.Sdmda.createGlobalVec()
        gv.array = np.arange(64, dtype=float)
        lv = rg.Sdmda.createGlobalVec()
        l0, l1 = rg.ranges[-1]
        lv.array = np.asarray(gv.array).reshape(rg1.Slshape, order='F')[..., l0:l1].reshape(
            -1, order='F')
        fine = Grid(dim=2, nx=32, ny=16, dof=1, comm=comm)
        a = random_function(fine, randgrid=rg, vals=lv)
        b = random_function(fine.serial(), randgrid=rg1, vals=gv)
        f0, f1 = fine.ranges[-1]
        ok = ok and np.array_equal(
            np.asarray(a.array).reshape(fine.Slshape, order='F'),
            np.asarray(b.array).reshape(fine.serial().Slshape, order='F')[..., f0:f1])
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_start_and_resume_values_on_two_ranks(tmp_path):
    world = 2
    res = _run_ranks(_ic_worker, world, (str(tmp_path),))
    assert sorted(res) == [(r, True) for r in range(world)]
