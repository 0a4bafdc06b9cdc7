"""
Drop-in boundary on the GPU: the reference-named Python classes (Grid,
Derivatives, implicitTS, ksfdsolver2 main) driving the CUDA library, checked
against the oracle, the reference goldens and the manufactured exact solution
of options93nx128dt1 (reference options93nx128dt1:22,39-47).
"""
import json
import os

import numpy as np
import pytest

from helpers import check_field, cond_scale, load_golden, oracle_physics, phys84, relerr

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
OPT93_FILE = os.path.join(HERE, 'options', 'options93nx128dt1.args')
OPT93 = '@' + OPT93_FILE


def opt93_with(tmp_path, **over):
    """copy of the option file with some name=value lines replaced (the
    reference rejects duplicated parameters, so they cannot be appended)"""
    lines = []
    for line in open(OPT93_FILE):
        key = line.split('=', 1)[0].strip()
        if key in over:
            line = '%s=%s\n' % (key, over.pop(key))
        lines.append(line)
    f = tmp_path / 'opts93'
    f.write_text(''.join(lines))
    return '@' + str(f)


def build(args):
    from ksfd_b200 import SolutionParameters, parse_commandline, petsc_init
    from ksfd_b200.derivs import Derivatives
    from ksfd_b200.grid import Comm, Grid
    from ksfd_b200.solver import decode_sources, start_values
    cl = parse_commandline(args)
    petsc_init(cl.petsc)
    ps = SolutionParameters(cl)
    grid = Grid(dim=ps.dim, dof=ps.nligands + 1, width=ps.width, height=ps.height,
                depth=ps.depth, nx=ps.nwidth, ny=ps.nheight, nz=ps.ndepth,
                comm=Comm(0, 1))
    sources = decode_sources(cl.source, ps, grid)
    u0, t0 = start_values(cl, grid, ps)
    derivs = Derivatives(ps, grid, sources=sources, u0=u0)
    return cl, ps, grid, sources, u0, derivs


def exact93(x, t):
    lam = 0.003974930217658144
    s = np.exp(lam * t) * np.sin(2 * np.pi * (0.25 + 4.0 * x))
    return np.stack([9000 + s, 9000 + 0.6846227279629311 * s,
                     9000 + 0.088562372925828 * s])


def test_options93_initial_condition_and_operator_vs_reference_golden():
    cl, ps, grid, sources, u0, derivs = build([OPT93])
    x = grid.coordsNoGhosts[0]
    assert np.allclose(u0.array_r.reshape(grid.Vlshape, order='F'), exact93(x, 0.0),
                       rtol=0, atol=1e-9)
    g, physs, nrec = load_golden('opt93_1d128')
    for r in range(nrec):
        t = physs[r]['t']
        u = grid.Vdmda.createGlobalVec()
        u.array = g['u_%d' % r]
        f = derivs.dfdt(u, t=t)
        cond = cond_scale(oracle_physics(physs[r]), g['u_%d' % r])
        assert check_field(f.array_r, g['f_%d' % r], grid.dof, 1e-11, cond) < 1.0, r
        vel = derivs.velocity(u, t=t)
        assert vel.shape == (1,) + grid.Slshape
        assert np.allclose(vel.reshape(-1, order='F'), g['vel_%d' % r], rtol=1e-9, atol=1e-12)
        kJ = derivs.Jacobian(u, t=t)            # shift 0: mult gives -J v
        v = grid.Vdmda.createGlobalVec()
        v.array = g['v_%d' % r][0]
        y = kJ.mult(v, grid.Vdmda.createGlobalVec())
        assert relerr(-y.array_r, g['Jv_%d' % r][0], grid.dof) < 1e-11


def test_options93_trajectory_vs_oracle_and_exact_solution():
    from ksfd_b200.ts import make_implicitTS
    from oracle import ksfd_oracle as O
    nsteps = 30
    cl, ps, grid, sources, u0, derivs = build([OPT93])
    ts = make_implicitTS(derivs, t0=0.0, dt=ps.params0['dt'], tmax=ps.params0['tmax'],
                         maxsteps=nsteps, rtol=ps.params0['rtol'], atol=ps.params0['atol'])
    ts.setMonitor(ts.historyMonitor)
    ts.solve()
    assert ts.getStepNumber() == nsteps and not ts.diverged
    assert ts.getSNESFailures() == 0
    assert abs(ts.getTime() - nsteps * 1.0) < 1e-12          # -ts_adapt_type none, dt=1
    # oracle: same loop, direct LU
    v = ps.values0
    ph = O.Physics(1, [128], [1.0 / 128],
                   [(v['alpha_1'], v['beta_1'], [(1.0, v['s_1_1'], v['gamma_1_1'], v['D_1_1'])]),
                    (v['alpha_2'], v['beta_2'], [(1.0, v['s_2_1'], v['gamma_2_1'], v['D_2_1'])])],
                   v['s2'], v['rhomax'], v['cushion'], v['maxscale'], 'tophat',
                   v['rhomin'], v['Umin'])
    srcfn = lambda t: [np.asarray(s(t)) for s in sources]
    traj = O.integrate(u0.array_r.copy(), 0.0, 1.0, nsteps, ph, sources_fn=srcfn)
    x = grid.coordsNoGhosts[0]
    for k in (1, 10, nsteps):
        mine = ts.history[k]['u'].reshape(grid.Vlshape, order='F')
        ref = traj[k - 1][1]
        # deviation from the mean state is what evolves (amplitude ~1): compare it
        assert np.abs(mine - ref).max() < 1e-8, (k, np.abs(mine - ref).max())
        ex = exact93(x, float(k))
        assert np.abs(mine - ex).max() < 2e-6, (k, np.abs(mine - ex).max())
    # growth rate of the perturbation == lamda of the manufactured solution
    a0 = np.ptp(ts.history[0]['u'].reshape(grid.Vlshape, order='F')[0])
    aN = np.ptp(ts.history[nsteps]['u'].reshape(grid.Vlshape, order='F')[0])
    assert abs(np.log(aN / a0) / nsteps - 0.003974930217658144) < 1e-7


def test_operator_callbacks_match_reference_plugin_api():
    """implicitIF / implicitIJ signatures of the reference (ksfdts.py:563,598)."""
    from helpers import phys84, random_state
    from ksfd_b200.ts import make_implicitTS
    from oracle import ksfd_oracle as O
    args = ['dim=2', 'nwidth=40', 'nheight=24', 'width=%r' % (40 / 384), 'height=%r' % (24 / 384),
            'sigma=0.02357', 's2=sigma**2/2', 'ngroups=2', 'nligands_1=1', 'alpha_1=1500',
            'beta_1=5.56e-4', 's_1_1=0.01', 'gamma_1_1=0.01', 'D_1_1=1e-6', 'nligands_2=1',
            'alpha_2=1500', 'beta_2=-5.56e-4', 's_2_1=0.001', 'gamma_2_1=0.001',
            'D_2_1=1e-5', 'srho0=0']
    cl, ps, grid, sources, u0, derivs = build(args)
    p = phys84(2, (40, 24))
    ph = oracle_physics(p)
    ts = make_implicitTS(derivs)
    rng = np.random.default_rng(0)
    u, udot, f = (grid.Vdmda.createGlobalVec() for _ in range(3))
    u.array = random_state(p, 3)
    udot.array = rng.standard_normal(u.size)
    ts.implicitIF(ts, 0.0, u, udot, f)
    F_ref = O.ifunction(u.array_r, udot.array_r, ph).reshape(-1, order='F')
    assert check_field(f.array_r, F_ref, grid.dof, 1e-11, cond_scale(ph, u.array_r)) < 1.0
    shift = 123.0
    assert ts.implicitIJ(ts, 0.0, u, udot, shift, ts.kJ, ts.kJ) is True
    y = ts.kJ.mult(udot, grid.Vdmda.createGlobalVec())
    Jv_ref = O.jvp(u.array_r, udot.array_r, shift, ph).reshape(-1, order='F')
    assert relerr(y.array_r, Jv_ref, grid.dof) < 1e-11
    # host edits through .array are seen by the device (reference mutates u.array)
    a = u.array.reshape(grid.Vlshape, order='F')
    a[0] *= 1.01
    f2 = derivs.dfdt(u, t=0.0)
    ref2 = O.dfdt(u.array_r.copy(), ph).reshape(-1, order='F')
    assert check_field(f2.array_r, ref2, grid.dof, 1e-11, cond_scale(ph, u.array_r)) < 1.0
    # CFL number as the reference computes it
    assert abs(ts.CFL_step(u) - O.cfl_maxh(u.array_r, ph)) < 1e-9 * O.cfl_maxh(u.array_r, ph)
    # worm count / conservation
    n0 = ts.count_worms(u)
    assert abs(n0 - u.array_r[0::grid.dof].sum()) < 1e-6
    u.array.reshape(grid.Vlshape, order='F')[0] *= 1.5
    ts.conserve_worms(u, n0)
    assert abs(ts.count_worms(u) - n0) < 1e-6


def test_solver_main_save_resume_roundtrip(tmp_path, capsys):
    """ksfdsolver2 command line: --save, --check, then --resume continues from
    the saved state and time."""
    from ksfd_b200.grid import Comm, Grid
    from ksfd_b200.solver import main
    from ksfd_b200.timeseries import TimeSeries
    save = str(tmp_path / 'solutions' / 'run93')
    chk = str(tmp_path / 'checks' / 'run93')
    rc = main('ksfdsolver2.py', opt93_with(tmp_path, maxsteps=6), '--save=' + save,
              '--check=' + chk)
    assert rc == 0
    out = capsys.readouterr().out
    assert 'SNES failures =  0' in out and out.count('clock:') == 7
    g = Grid(dim=1, nx=128, dof=3, comm=Comm(0, 1))
    ser = TimeSeries(save, grid=g, mode='r')
    assert list(ser.sorted_times()) == [float(k) for k in range(7)]
    last6 = ser.retrieve_by_time(6.0)
    assert os.path.exists(TimeSeries(chk + '_6_', grid=g, mode='r').filename)
    save2 = str(tmp_path / 'solutions' / 'run93b')
    rc = main('ksfdsolver2.py', opt93_with(tmp_path, maxsteps=4), '--resume=' + save,
              '--save=' + save2)
    assert rc == 0
    ser2 = TimeSeries(save2, grid=g, mode='r')
    assert list(ser2.sorted_times()) == [6.0, 7.0, 8.0, 9.0, 10.0]
    assert np.array_equal(ser2.retrieve_by_time(6.0), last6)
    x = g.coordsNoGhosts[0]
    assert np.abs(ser2.retrieve_by_time(10.0) - exact93(x, 10.0)).max() < 2e-6


OPT84_FILE = os.path.join(HERE, 'options', 'options84.args')


def opt_with(path, tmp_path, name, **over):
    lines = []
    for line in open(path):
        key = line.split('=', 1)[0].strip()
        if key.startswith('--save') or key.startswith('--check'):
            continue
        if key in over:
            line = '%s=%s\n' % (key, over.pop(key))
        lines.append(line)
    f = tmp_path / name
    f.write_text(''.join(lines))
    return '@' + str(f)


def test_options84_reduced_adaptive_run_vs_oracle(tmp_path, capsys):
    """BASELINE configs[1]: the options84 option file (2-D, two ligand groups,
    TSAdapt basic from dt = 1e-8, '-pc_type lu' mapped to the iterative solve)
    run through the ksfdsolver2 entry at the SAME grid spacing h = 1/384 on a
    48x48 tile, 8 accepted steps; the saved trajectory (times and states) must
    match the oracle's adaptive ROSW with direct solves started from the saved
    initial state."""
    from ksfd_b200.grid import Comm, Grid
    from ksfd_b200.solver import main
    from ksfd_b200.timeseries import TimeSeries
    from oracle import ksfd_oracle as O
    save = str(tmp_path / 'solutions' / 'run84')
    n = 48
    args = opt_with(OPT84_FILE, tmp_path, 'opts84', nelements=n, width=n / 384.0,
                    height=n / 384.0, maxsteps=8)
    rc = main('ksfdsolver2.py', args, '--save=' + save)
    assert rc == 0
    out = capsys.readouterr().out
    assert 'SNES failures =  0' in out
    g = Grid(dim=2, nx=n, ny=n, width=n / 384.0, height=n / 384.0, dof=3, comm=Comm(0, 1))
    ser = TimeSeries(save, grid=g, mode='r')
    times = list(ser.sorted_times())
    assert len(times) == 9 and times[0] == 0.0
    u0 = np.asarray(ser.retrieve_by_time(0.0)).reshape(-1, order='F')
    p = phys84(2, (n, n))
    ph = oracle_physics(p)
    traj = O.integrate(u0, 0.0, 1e-8, 8, ph,
                       adapt=dict(atol=0.01, rtol=1e-6, clip=(0.1, 5.0), dt_min=1e-20,
                                  dt_max=1e4))
    for k, (t, u) in enumerate(traj):
        assert abs(times[k + 1] - t) <= 1e-9 * t, (k, times[k + 1], t)
        got = np.asarray(ser.retrieve_by_time(times[k + 1])).reshape(-1, order='F')
        ref = u.reshape(-1, order='F')
        assert relerr(got, ref, 3) < 1e-8, (k, relerr(got, ref, 3))


@pytest.mark.parametrize('name,over', [
    ('options80', dict(nelements=96, maxsteps=5)),                  # 1-D adaptive
    ('options81', dict(nelements=32, maxsteps=4)),                  # 2-D adaptive
    ('options92', dict(nelements=64, maxsteps=4)),                  # 1-D, expression ICs
    ('options113a', dict(nelements=96, randgridnw=96, randgridnh=96,
                         maxsteps=5)),     # noise injection, worm conservation, CFL limiter
])
def test_shipped_option_files_run_through_the_solver_entry(name, over, tmp_path, capsys):
    """every option file shipped with the reference (argument lines unchanged
    except for the grid size / step count) runs through the ksfdsolver2 entry on
    the device path: parser, parameters, initial condition, adaptive ROSW,
    monitors, time series."""
    from ksfd_b200.solver import main
    path = os.path.join(HERE, 'options', name + '.args')
    save = str(tmp_path / 'solutions' / name)
    rc = main('ksfdsolver2.py', opt_with(path, tmp_path, name, **over), '--save=' + save)
    assert rc == 0
    out = capsys.readouterr().out
    assert 'SNES failures =  0' in out, out[-2000:]
    assert out.count('clock:') >= over['maxsteps']


def test_noise_injection_on_device_uses_the_reference_stream():
    """KSFDTS.add_variance (reference ksfdts.py:268-284): rho *= exp(sd * N(0,1)) with the
    sample drawn from the rank's numpy stream as the reference does.  The field stays on
    the device (only the sample is uploaded); the result equals the host formula to
    rounding, the ligand fields are untouched, and conserve_worms restores the total."""
    from ksfd_b200.random import Generator
    from ksfd_b200.ts import make_implicitTS
    args = ['@' + os.path.join(HERE, 'options', 'options113a.args')]
    cl, ps, grid, sources, u0, derivs = build(args)
    ts = make_implicitTS(derivs, t0=0.0, dt=1e-3, tmax=1.0, maxsteps=1)
    before = np.array(u0.array_r).reshape(grid.Vlshape, order='F').copy()
    N0 = ts.count_worms(u0)
    assert abs(N0 - before[0].sum()) <= 1e-9 * abs(N0)
    Generator(seed=1234, comm=grid.comm)
    z = np.random.default_rng(np.random.SeedSequence(1234).spawn(1)[0]).normal(size=grid.Slshape)
    vrate = ps.values(0.0)['variance_rate']
    dt = 7.0
    ts.getTime = lambda: 0.0
    u = ts.add_variance(u0, dt)
    after = np.array(u.array_r).reshape(grid.Vlshape, order='F')
    want = before[0] * np.exp(np.sqrt(vrate * dt) * z)
    assert np.abs(after[0] - want).max() <= 4e-16 * np.abs(want).max()
    assert np.array_equal(after[1:], before[1:])
    u = ts.conserve_worms(u, N0)
    assert abs(ts.count_worms(u) - N0) <= 1e-12 * N0


def test_in_step_profile_counts_stencil_launches():
    """option 'profile': CUDA events around the stencil launches (what bench.py reports as
    the in-step kernel time)."""
    import torch
    from helpers import product_physics
    from ksfd_b200 import core
    p = phys84(2, (64, 48))
    ctx = core.Context(2, p['n'], 3)
    ctx.set_physics(product_physics(p))
    u = ctx.upload(np.full(64 * 48 * 3, 9000.0))
    v = torch.randn(ctx.npts * 3, device='cuda', dtype=torch.float64)
    ctx.jvp_setup(u, 100.0)
    ctx.set_option('profile', 1)
    for _ in range(5):
        ctx.jvp(v, precond=True)
    for _ in range(3):
        ctx.residual(u, v)
    d = ctx.profile_fetch()
    assert d['jvp_launches_all'] == 5 and d['residual_launches_all'] == 3
    assert d['jvp_ms_all'] > 0.0 and d['residual_ms_all'] > 0.0
    assert ctx.profile_fetch()['jvp_launches_all'] == 0          # fetch resets
    ctx.set_option('profile', 0)
    ctx.jvp(v)
    assert ctx.profile_fetch()['jvp_launches_all'] == 0
    ctx.close()
