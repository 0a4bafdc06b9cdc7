"""BASELINE configs[0]: the reference's options93nx128dt1 case (1-D, nx = 128, manufactured solution with a
time-dependent source, dt = 1, 2080 steps) run to the END through the host mirror of the reference's
time stepper on one GPU: wall time and the error against the exact solution at t = 2080."""
import os, sys, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, 'tests'))
import numpy as np
from test_gpu_dropin import OPT93, build, exact93
from ksfd_b200.ts import make_implicitTS

cl, ps, grid, sources, u0, derivs = build([OPT93])
ts = make_implicitTS(derivs, t0=0.0, dt=ps.params0['dt'], tmax=ps.params0['tmax'],
                     maxsteps=int(ps.params0['maxsteps']), rtol=ps.params0['rtol'], atol=ps.params0['atol'])
t0 = time.perf_counter()
ts.solve()
wall = time.perf_counter() - t0
k, t = ts.getStepNumber(), ts.getTime()
u = np.asarray(ts.getSolution().array_r).reshape(grid.Vlshape, order='F')
ex = exact93(grid.coordsNoGhosts[0], float(t))
amp = np.exp(0.003974930217658144 * t)
print('options93nx128dt1: %d steps to t = %g in %.2f s wall (%.2f ms/step), SNES failures %d, KSP iterations %d'
      % (k, t, wall, 1e3 * wall / max(k, 1), ts.getSNESFailures(), ts.getKSPIterations()))
print('max |u - exact| at t = %g: %.3e  (amplitude of the manufactured perturbation exp(lamda t) = %.1f: relative %.2e)'
      % (t, np.abs(u - ex).max(), amp, np.abs(u - ex).max() / amp))
