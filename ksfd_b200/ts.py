"""
Time steppers with the reference interface (KSFD/ksfdts.py, ksfdtsmaker.py).

`KSFDTS` owns the Python step loop exactly as the reference does
(ksfdts.py:170-229): monitor, clamp, TS.step, noise injection, worm
conservation, CFL check, monitor.  `TS.step` — PETSc's ROSW/BEuler + SNES
ksponly + KSP + PC in the reference — is ONE call into the CUDA library
(`ksfd_ts_step`): device-resident ROSW ra34pw2 stages, fused residual kernel,
matrix-free J.v, GMRES with point-block Jacobi, TSAdapt basic.

PETSc options understood (from the '--petsc' block of the option files):
  -ts_type rosw|beuler            -ts_adapt_type none|basic
  -ts_adapt_clip lo,hi            -ts_adapt_dt_min/-ts_adapt_dt_max
  -ts_adapt_safety / -ts_adapt_reject_safety / -ts_adapt_scale_solve_failed
  -ksp_rtol / -ksp_atol / -ksp_divtol / -ksp_max_it / -ksp_gmres_restart
  -ksp_gmres_cgs_refinement_type refine_never|refine_always
  -ksp_type gmres|richardson (default: the library chooses — stationary block-Jacobi
   sweeps fused into the stencil kernel while they contract fast, GMRES otherwise)
  -pc_type none|pbjacobi|fft (lu, as shipped in the option files, is mapped to the
   iterative solve with a tight default tolerance and the automatic choice between
   block Jacobi and the spectral preconditioner: there is no LU on the device)
Anything else in the block is accepted and ignored, as PETSc would do for
options it has no consumer for (-options_left would list them).
"""
import gc
import math
import os
from datetime import datetime

import numpy as np

from . import core
from .params import petsc_options


class KSFDTS:
    """Base class for KSFD timesteppers (reference KSFD/ksfdts.py:53-497)."""

    default_rollback_factor = 0.25
    default_hmin = 1e-20

    def __init__(self, derivs, t0=0.0, dt=0.001, tmax=20, maxsteps=100, rtol=1e-5,
                 atol=1e-5, restart=True, tstype='rosw', finaltime='stepover',
                 rollback_factor=None, hmin=None, comm=None):
        self.comm = comm if comm is not None else derivs.grid.comm
        self.mpi_comm = self.comm
        self.derivs = derivs
        self.t0 = float(t0)
        self.tmax = float(tmax)
        self.maxsteps = maxsteps
        self.rtol = float(rtol)
        self.atol = float(atol)
        self.restart = restart
        self.tstype = tstype
        self.finaltime = finaltime
        popts = petsc_options()
        if rollback_factor is None:
            self.rollback_factor = popts.getReal('ts_adapt_scale_solve_failed',
                                                 self.default_rollback_factor)
        else:
            self.rollback_factor = rollback_factor
        self.hmin = float(hmin if hmin else self.default_hmin)
        self.history = []
        self.diverged = False
        self._monitors = []
        self._snes_failures = 0
        self._ksp_its = 0
        self._rejections = 0
        self._k = 0
        self._h = float(dt)
        self._t = self.t0
        self.u = self.derivs.u0.duplicate()
        self.f = self.u.duplicate()
        self.derivs.u0.copy(self.u)
        self.CFL_maxh = self.CFL_step(self.derivs.u0)
        self.kJ = self.derivs.Jacobian(self.derivs.u0)
        self._opts = None
        self.setFromOptions()

    # -- PETSc TS style accessors -------------------------------------------
    def setSolution(self, u):
        self.u = u

    def getSolution(self):
        return self.u

    def setTime(self, t):
        self._t = float(t)

    def getTime(self):
        return self._t

    def setTimeStep(self, h):
        self._h = float(h)

    def getTimeStep(self):
        return self._h

    def getStepNumber(self):
        return self._k

    def setMaxSteps(self, max_steps):
        self.maxsteps = max_steps

    def getMaxSteps(self):
        return self.maxsteps

    def setMaxTime(self, max_time):
        self.tmax = float(max_time)

    def getMaxTime(self):
        return self.tmax

    def getSNESFailures(self):
        return self._snes_failures

    def getKSPIterations(self):
        return self._ksp_its

    def getStepRejections(self):
        return self._rejections

    def setTolerances(self, rtol=None, atol=None):
        if rtol is not None:
            self.rtol = float(rtol)
        if atol is not None:
            self.atol = float(atol)
        self._opts = None

    def setMonitor(self, monitor, args=None, kargs=None):
        self._monitors.append((monitor, tuple(args or ()), dict(kargs or {})))

    def monitor(self, k, t, u):
        for fn, a, kw in self._monitors:
            fn(self, k, t, u, *a, **kw)

    def setFromOptions(self):
        """Read the '--petsc' options (see module docstring)."""
        po = petsc_options()
        tstype = po.get('ts_type', self.tstype if isinstance(self.tstype, str) else 'rosw')
        if tstype not in ('rosw', 'beuler'):
            raise ValueError('-ts_type %s is not supported on the device '
                             '(rosw, beuler)' % tstype)
        adapt = po.get('ts_adapt_type', 'basic')
        if adapt not in ('none', 'basic'):
            raise ValueError('-ts_adapt_type %s is not supported (none, basic)' % adapt)
        clip = po.getRealArray('ts_adapt_clip', [0.1, 10.0])
        pc = po.get('pc_type', 'pbjacobi')
        direct = pc in ('lu', 'cholesky')
        refine = po.get('ksp_gmres_cgs_refinement_type', 'refine_never')
        # -ksp_type gmres | richardson; anything else (preonly with a direct solver in
        # the shipped option files, or nothing) leaves the choice to the library
        ksp_type = po.get('ksp_type', 'auto')
        if ksp_type not in ('gmres', 'richardson'):
            ksp_type = 'auto'
        self._ts_kind = tstype
        self._opts = core.ts_options(
            ts_type=tstype, adapt=adapt, atol=self.atol, rtol=self.rtol,
            clip=(clip[0], clip[1]),
            dt_min=po.getReal('ts_adapt_dt_min', 1e-20),
            dt_max=po.getReal('ts_adapt_dt_max', 1e50),
            safety=po.getReal('ts_adapt_safety', 0.9),
            reject_safety=po.getReal('ts_adapt_reject_safety', 0.5),
            max_reject=po.getInt('ts_max_reject', 10),
            # a direct solver in the option file asks for an "exact" solve
            ksp_rtol=po.getReal('ksp_rtol', 1e-10 if direct else 1e-5),
            ksp_atol=po.getReal('ksp_atol', 1e-50),
            ksp_dtol=po.getReal('ksp_divtol', 1e5),
            ksp_max_it=max(po.getInt('ksp_max_it', 10000), 1),
            restart=po.getInt('ksp_gmres_restart', 30),
            reorth=0 if refine == 'refine_never' else 1,
            ksp_type=ksp_type,
            # lu/cholesky (the shipped option files): automatic choice between the
            # fused block-Jacobi and the spectral preconditioner; 'fft' forces the latter
            precond={'none': 0, 'pbjacobi': 1, 'bjacobi': 1, 'jacobi': 1, 'fft': 2}.get(pc, 3))
        return self._opts

    # -- the step -------------------------------------------------------------
    def step(self, groom=False, velocity_max=False):
        """One accepted step of the device integrator (PETSc TS.step()).
        groom / velocity_max: the clamp before and the CFL maxima after the step (the
        neighbours of TS.step() in the reference's loop, ksfdts.py:205-227) inside the same
        C call; the maxima are kept for the next CFL_check of the unchanged state."""
        d = self.derivs
        ctx = d.ctx
        if self._opts is None:
            self.setFromOptions()
        self._opts.flags = (1 if groom else 0) | (2 if velocity_max else 0)
        self._vmax_fresh = None
        u = self.u.device(ctx)
        td = d._phys_td or d._src_td
        src = d.source_device(self._t)
        if td:
            # stage-time refresh of physics / sources; the source buffer is
            # updated IN PLACE so the pointer handed to the library stays valid
            res = ctx.ts_step(u, self._t, self._h, self._opts, src=src,
                              time_cb=self._stage_time_cb)
        else:
            res = ctx.ts_step(u, self._t, self._h, self._opts, src=src)
        self.u.mark_device_written()
        self._ksp_its += res.ksp_its
        self._rejections += res.rejections
        self.last_result = res
        if res.ksp_fail or not res.accepted:
            self._snes_failures += 1
            self.diverged = True
            return res
        self._k += 1
        self._t = res.t_new
        self._h = res.h_next
        if res.have_vmax:
            self._vmax_fresh = np.array([res.vmax[i] for i in range(d.grid.dim)])
        return res

    def _stage_time_cb(self, tt):
        d = self.derivs
        old = d._src_dev
        if d._phys_td:
            d.ctx.set_physics(d.ps.physics(d.grid.spacing, tt))
            d._phys_t = float(tt)
        if d._have_src and d._src_td:
            src = np.stack([np.asarray(s(tt), dtype=float) for s in d.sources])
            new = d.ctx.upload(src.reshape(-1, order='F'))
            old.copy_(new)              # in place: the library holds this pointer
            d._src_t = float(tt)

    def solve(self, u=None):
        """Run the timestepper (reference KSFDTS.solve, ksfdts.py:170-229)."""
        if u:
            u.copy(self.u)
        else:
            u = self.u
        self.setSolution(u)
        self.setTime(self.t0)
        self.setFromOptions()
        tmax, kmax = self.getMaxTime(), self.getMaxSteps()
        k, h = self.getStepNumber(), self.getTimeStep()
        self.CFL_check()
        t = self.getTime()
        ps = self.derivs.ps
        Nworms = self.count_worms(u)
        self.lastvart = ps.params0['lastvart'] if 'lastvart' in ps.params0 else t
        cw = ps.params0['conserve_worms']
        conserve = False if cw == 'False' else bool(cw)
        self.monitor(k, t, u)
        while (not self.diverged) and k < kmax and t <= tmax and h >= self.hmin:
            # groom(u), TS.step() and the velocity maxima of CFL_check in ONE library call
            # (time-dependent parameters: CFL_check evaluates them at the new time itself)
            self.step(groom=True,
                      velocity_max=not (self.derivs._phys_td or self.derivs._src_td))
            if k % 20 == 0:
                gc.collect()
            k, h, t = self.getStepNumber(), self.getTimeStep(), self.getTime()
            u = self.getSolution()
            if self.diverged:
                break
            dt = t - self.lastvart
            if self.is_noise_time(t, self.lastvart):
                u = self.add_variance(u, dt)
                if conserve:
                    u = self.conserve_worms(u, Nworms)
                self.lastvart = t
                self._vmax_fresh = None         # the state changed after the step
            self.CFL_check()
            self.monitor(k, t, u)

    def groom(self, u):
        """Clamp rho >= rhomin, U >= Umin, NaN -> min, in place on the device."""
        return self.derivs.groom_vec(u)

    def count_worms(self, u):
        ctx = self.derivs.ctx
        return ctx.sum_dof0(u.device(ctx))          # all-reduced in the library

    def conserve_worms(self, u, Nworms):
        ctx = self.derivs.ctx
        n = ctx.sum_dof0(u.device(ctx))
        ctx.scale_dof0(u.device(ctx), Nworms / n)
        u.mark_device_written()
        return u

    def is_noise_time(self, t, lastvart):
        ps = self.derivs.ps
        vrate = ps.values(t)['variance_rate']
        if not vrate or vrate <= 0.0:
            return False
        flast = ps.values(lastvart)['variance_timing_function']
        fnow = ps.values(t)['variance_timing_function']
        return fnow - flast >= 1.0

    def add_variance(self, u, dt):
        """Lognormal multiplicative noise on rho (reference ksfdts.py:268-284).
        The standard-normal sample comes from the rank's numpy stream exactly as in
        the reference (same seed -> same sample); only the sample (8 bytes per point)
        goes to the device, where rho *= exp(sd*z) is applied in place — the field
        itself is not copied to the host and back."""
        from .random import Generator
        vrate = self.derivs.ps.values(self.getTime())['variance_rate']
        if not vrate or vrate <= 0.0:
            return u
        sd = np.sqrt(vrate * dt)
        z = Generator.get_rng().normal(size=self.derivs.grid.Slshape)
        ctx = self.derivs.ctx
        ctx.mul_exp_dof0(u.device(ctx), z, sd)
        u.mark_device_written()
        return u

    def CFL_check(self):
        h, t, u = self.getTimeStep(), self.getTime(), self.getSolution()
        self.CFL_maxh = self.CFL_step(u, t)
        safety = self.derivs.ps.values(t)['CFL_safety_factor']
        if safety > 0.0:
            maxh = safety * self.CFL_maxh
            if h > maxh:
                self.setTimeStep(maxh)

    def CFL_step(self, u, t=None):
        """min_d spacing_d*sw/max|v_d| (reference ksfdts.py:302-319); the max
        is reduced on the device and across ranks."""
        vmax = getattr(self, '_vmax_fresh', None)
        self._vmax_fresh = None
        if vmax is None:
            vmax = self.derivs.velocity_max(u, t)
        sw = self.derivs.grid.stencil_width
        hm = [float('inf') if v == 0.0 else s * sw / v
              for v, s in zip(vmax, self.derivs.grid.spacing)]
        return float(np.min(hm))

    def cleanup(self):
        self.kJ = None

    # -- monitors ------------------------------------------------------------
    def printMonitor(self, ts, k, t, u):
        if self.comm.rank == 0:
            h = ts.getTimeStep()
            clock = datetime.now().strftime('%H:%M:%S')
            if hasattr(self, 'lastt'):
                out = 'clock: %s, step %3d t=%8.3g dt=%8.3g h=%8.3g' % (
                    clock, k, t, t - self.lastt, h)
            else:
                out = 'clock: %s, step %3d t=%8.3g h=%8.3g' % (clock, k, t, h)
            if hasattr(self, 'CFL_maxh'):
                out += ' CFL=%8.3g' % self.CFL_maxh
            print(out, flush=True)
            self.lastt = t

    def historyMonitor(self, ts, k, t, u):
        self.history.append(dict(step=k, h=ts.getTimeStep(), t=t,
                                 u=np.array(u.array_r)))

    def checkpointMonitor(self, ts, k, t, u, prefix, mpiok=False):
        from .timeseries import TimeSeries, dillnp
        import zipfile
        zipit = prefix.endswith('.zip')
        real = prefix[:-4] if zipit else prefix
        cpf = TimeSeries(real + '_' + str(k) + '_', grid=self.derivs.grid, mode='w',
                         comm=self.comm)
        cpf.set_info('commandlineArguments', dillnp(self.derivs.ps.clargs))
        cpf.set_info('SolutionParameters', dillnp(self.derivs.ps, recurse=True))
        cpf.set_info('dt', float(ts.getTimeStep()))
        cpf.set_info('lastvart', float(getattr(self, 'lastvart', t)))
        try:
            cpf.set_info('sources', dillnp(self.derivs.sources))
        except Exception:
            pass
        cpf.store(u, t, k=k)
        name = cpf.filename
        cpf.close()
        if zipit:
            zname = real + 's%dr%d.zip' % (self.comm.size, self.comm.rank)
            with zipfile.ZipFile(zname, mode='w' if k == 0 else 'a',
                                 compression=zipfile.ZIP_DEFLATED) as zf:
                zf.write(name, arcname=os.path.basename(name))
            os.remove(name)

    def makeSaveMonitor(self, timeseries):
        self.timeseries = timeseries

        def closeSaveMonitor():
            pass

        def saveMonitor(ts, k, t, u):
            if not self.timeseries.is_open():
                self.timeseries.reopen()
            self.timeseries.store(u, t, k=k)
            self.timeseries.set_info('dt', float(ts.getTimeStep()))
            self.timeseries.temp_close()

        return saveMonitor, closeSaveMonitor


class implicitTS(KSFDTS):
    """Fully implicit timestepper (reference KSFD/ksfdts.py:500-640).  The
    operator callbacks of the reference's plugin API are kept for callers that
    drive the operator themselves."""

    def implicitIF(self, ts, t, u, udot, f):
        """f <- udot - b(u, t)   (one fused kernel)"""
        self.derivs.ifunction(u, udot, t=t, out=f)

    def implicitIJ(self, ts, t, u, udot, shift, J, B):
        """J (and B) <- shift*I - db/du at u, as a matrix-free operator."""
        self.derivs.Jacobian(u, t=t, out=J if J is not None else self.kJ, shift=shift)
        if B is not None and B is not J:
            B.shift, B._u = J.shift, J._u
        return True


def ksfdTS(derivs, **kw):
    return KSFDTS(derivs, **kw)


def make_implicitTS(derivs, t0=0.0, dt=0.001, tmax=20, maxsteps=100, rtol=1e-5,
                    atol=1e-5, restart=True, tstype=None, finaltime=None,
                    rollback_factor=None, hmin=None, comm=None):
    """Factory with the signature of KSFD.implicitTS
    (reference KSFD/ksfdtsmaker.py:101-168)."""
    return implicitTS(derivs, t0=t0, dt=dt, tmax=tmax, maxsteps=maxsteps, rtol=rtol,
                      atol=atol, restart=restart, tstype=tstype or 'rosw',
                      finaltime=finaltime or 'stepover',
                      rollback_factor=rollback_factor, hmin=hmin, comm=comm)
