"""
Command-line driver with the reference's `ksfdsolver2.py` behaviour
(ksfdsolver2.py:642-774): parse option file -> parameters -> grid -> sources
-> initial (or resumed) values -> Derivatives -> implicitTS -> monitors ->
solve.  Usage, unchanged:

    python ksfdsolver2.py @options93nx128dt1
    torchrun --nproc-per-node N ksfdsolver2.py @options84     (one rank per GPU)
"""
import sys

import numpy as np

from .derivs import Derivatives, SpatialExpression
from .grid import Grid, comm_world
from .params import (KSFDException, SolutionParameters, find_duplicates,
                     parse_commandline, petsc_init)
from .random import Generator, random_function
from .timeseries import TimeSeries, dillnp
from .ts import make_implicitTS


def decode_sources(sargs, ps, grid):
    """--source=name=expr list -> one SpatialExpression per dof
    (reference ksfdsolver2.py:473-498)."""
    names = ['rho'] + [l.name() for l in ps.groups.ligands()]
    keys = [a.split('=', 1)[0] for a in sargs]
    dups = find_duplicates(keys)
    if dups:
        raise KSFDException('duplicated sources: ' + ', '.join(dups))
    sources = [SpatialExpression(ps, grid, '0.0') for _ in names]
    for a in sargs:
        k, val = a.split('=', 1)
        if k not in names:
            raise KSFDException('unknown function: ' + k)
        sources[names.index(k)] = SpatialExpression(ps, grid, val)
    return sources


def start_values(clargs, grid, ps):
    """rho = rho0(x) + random field, U = U0(x) or rho*s/gamma
    (reference ksfdsolver2.py:580-639)."""
    p0, v0 = ps.params0, ps.values0
    rn = [p0['randgridnw'] or ps.nwidth // 4, p0['randgridnh'] or ps.nheight // 4,
          p0['randgridnd'] or ps.ndepth // 4]
    # the coarse random grid is always built whole (size-1 comm): on several ranks
    # every rank draws the same global sample from the single-rank stream and
    # random_function keeps the rank's slab (the coarse grid may have fewer planes
    # than ranks * stencil_width, and the field must not depend on the GPU count)
    from .grid import Comm
    rgrid = Grid(dim=ps.dim, width=ps.width, height=ps.height, depth=ps.depth,
                 nx=max(rn[0], 1), ny=max(rn[1], 1), nz=max(rn[2], 1), dof=1,
                 comm=Comm(0, 1))
    murho0 = v0['Nworms'] / (ps.width ** ps.dim)
    sigma = v0['srho0']
    rvals = rgrid.Sdmda.createGlobalVec()
    if sigma == 0.0:
        rvals.array[:] = murho0
    else:
        sig = SpatialExpression(ps, rgrid, sigma)()
        sample = Generator.global_rng().normal(size=rgrid.Slshape)
        rvals.array = (sig * sample + murho0).reshape(-1, order='F')
    rra = random_function(grid, randgrid=rgrid, vals=rvals).array.reshape(
        grid.Slshape, order='F')
    vec = grid.Vdmda.createGlobalVec()
    va = vec.array.reshape(grid.Vlshape, order='F')
    rho0 = v0['rho0']
    va[0] = SpatialExpression(ps, grid, rho0)() if rho0 else 0.0
    va[0] += rra
    for dof, lig in enumerate(ps.groups.ligands()):
        name = 'U0' + lig.name()[1:]
        val = v0.get(name)
        if val is not None and val is not False and val != '':
            va[dof + 1] = SpatialExpression(ps, grid, val)()
        else:
            va[dof + 1] = va[0] * float(lig.s / lig.gamma)
    vec.array = va.reshape(-1, order='F')
    return vec, ps.t0


def resume_values(clargs, grid, ps):
    """Last time point of a saved series (reference ksfdsolver2.py:525-578)."""
    name = clargs.resume or clargs.restart
    cpf = TimeSeries(name, grid=grid, mode='r',
                     retries=getattr(clargs, 'series_retries', 0) or 0,
                     retry_interval=getattr(clargs, 'series_retry_interval', 60) or 60)
    times = cpf.sorted_times()
    tlast = times[-1]
    given = {p.split('=', 1)[0] for p in clargs.params}
    if clargs.resume:
        t = tlast
        if 'dt' not in given:
            if 'dt' in cpf.info:
                ps.params0['dt'] = float(cpf.info['dt'])
            elif len(times) >= 2:
                ps.params0['dt'] = float(tlast - times[-2])
        if 'lastvart' not in given:
            ps.params0['lastvart'] = (float(cpf.info['lastvart'])
                                      if 'lastvart' in cpf.info else t)
    else:
        t = ps.t0
        if 'lastvart' not in given:
            ps.params0['lastvart'] = ps.t0
    values = np.asarray(cpf.retrieve_by_time(tlast))
    cpf.close()
    vec = grid.Vdmda.createGlobalVec()
    vec.array = np.asfortranarray(values).reshape(-1, order='F')
    return vec, t


def initial_values(clargs, grid, ps):
    if clargs.resume or clargs.restart:
        return resume_values(clargs, grid, ps)
    return start_values(clargs, grid, ps)


def init_distributed():
    """One process per GPU when launched under torchrun."""
    import os
    import torch
    if int(os.environ.get('WORLD_SIZE', '1')) > 1:
        import torch.distributed as dist
        local = int(os.environ.get('LOCAL_RANK', '0'))
        torch.cuda.set_device(local)
        if not dist.is_initialized():
            dist.init_process_group('nccl', device_id=torch.device('cuda', local))


def main(*args):
    argv = list(args) if args else sys.argv
    clargs = parse_commandline(argv[1:])
    petsc_init(clargs.petsc)
    if clargs.noperiodic:
        raise KSFDException('--periodic=false not implemented')
    init_distributed()
    comm = comm_world()
    ps = SolutionParameters(clargs)
    Generator(seed=clargs.seed, comm=comm)
    if clargs.showparams:
        for k, v in ps.params0.items():
            print('%s=%s' % (k, v))
        return 0
    grid = Grid(dim=ps.dim, dof=ps.nligands + 1, width=ps.width, height=ps.height,
                depth=ps.depth, nx=ps.nwidth, ny=ps.nheight, nz=ps.ndepth, comm=comm)
    sources = decode_sources(clargs.source, ps, grid)
    vec0, t = initial_values(clargs, grid, ps)
    tseries = None
    if clargs.save:
        tseries = TimeSeries(clargs.save, grid=grid, mode='w', comm=comm)
        tseries.info['commandlineArguments'] = dillnp(clargs)
        tseries.info['SolutionParameters'] = dillnp(ps, recurse=True)
        try:
            tseries.info['sources'] = dillnp(sources)
        except Exception:
            pass
        tseries.info['dt'] = float(ps.params0['dt'])
        if 'lastvart' in ps.params0:
            tseries.info['lastvart'] = float(ps.params0['lastvart'])
        tseries.flush()
    derivs = Derivatives(ps, grid, sources=sources, u0=vec0)
    maxsteps = 1 if clargs.onestep else ps.params0['maxsteps']
    ts = make_implicitTS(derivs, t0=t, restart=not bool(clargs.resume or clargs.restart),
                         rtol=ps.params0['rtol'], atol=ps.params0['atol'],
                         dt=ps.params0['dt'], tmax=ps.params0['tmax'],
                         maxsteps=maxsteps)
    ts.setMonitor(ts.printMonitor)
    if clargs.save:
        save, close_save = ts.makeSaveMonitor(timeseries=tseries)
        ts.setMonitor(save)
    if clargs.check:
        ts.setMonitor(ts.checkpointMonitor, (),
                      {'prefix': clargs.check, 'mpiok': clargs.mpiok})
    failed = None
    try:
        ts.solve()
    except KeyboardInterrupt as e:
        print('KeyboardInterrupt:', str(e))
    except Exception as e:          # noqa: BLE001 — as the reference (ksfdsolver2.py:747-756): report,
        print('Exception:', str(e))  # then close the series so that what was saved stays readable
        failed = e
    if clargs.save:
        close_save()
        tseries.close()
    if failed is not None:
        # (the reference goes on and returns 0; a job script is better served by the error)
        raise failed
    ts.cleanup()
    if comm.rank == 0:
        print('SNES failures = ', ts.getSNESFailures())
        print('KSP iterations = %d, steps = %d, rejections = %d' % (
            ts.getKSPIterations(), ts.getStepNumber(), ts.getStepRejections()))
    return 0
