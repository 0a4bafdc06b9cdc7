"""BASELINE configs[0] on the CPU: options93nx128dt1 run to the END (2081 steps, dt = 1) by the numpy
oracle — ROSW ra34pw2 with DIRECT sparse LU solves (SuperLU), the solver class the reference's option
file asks for (-pc_type lu, MUMPS) — from the same option file, initial values and time-dependent
source as the GPU run of scripts/run_options93.py.  Prints the error against the manufactured exact
solution, to be read next to the GPU run's (profiles/r02_options93_full_run.txt).  No GPU needed:
parsing, grid, sources and initial values come from the host mirror (ksfd_b200.params / grid / solver),
the arithmetic from oracle/ksfd_oracle.py."""
import os, sys, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, 'tests'))
import numpy as np
from ksfd_b200 import SolutionParameters, parse_commandline, petsc_init
from ksfd_b200.grid import Comm, Grid
from ksfd_b200.solver import decode_sources, start_values
from oracle import ksfd_oracle as O

OPT93 = '@' + os.path.join(R, 'tests', 'options', 'options93nx128dt1.args')
LAM = 0.003974930217658144


def exact93(x, t):
    s = np.exp(LAM * t) * np.sin(2 * np.pi * (0.25 + 4.0 * x))
    return np.stack([9000 + s, 9000 + 0.6846227279629311 * s, 9000 + 0.088562372925828 * s])


cl = parse_commandline([OPT93])
petsc_init(cl.petsc)
ps = SolutionParameters(cl)
grid = Grid(dim=ps.dim, dof=ps.nligands + 1, width=ps.width, height=ps.height, depth=ps.depth,
            nx=ps.nwidth, ny=ps.nheight, nz=ps.ndepth, comm=Comm(0, 1))
sources = decode_sources(cl.source, ps, grid)
u0, t0 = start_values(cl, grid, ps)
v = ps.values0
ph = O.Physics(1, [128], [1.0 / 128],
               [(v['alpha_1'], v['beta_1'], [(1.0, v['s_1_1'], v['gamma_1_1'], v['D_1_1'])]),
                (v['alpha_2'], v['beta_2'], [(1.0, v['s_2_1'], v['gamma_2_1'], v['D_2_1'])])],
               v['s2'], v['rhomax'], v['cushion'], v['maxscale'], 'tophat', v['rhomin'], v['Umin'])
srcfn = lambda t: [np.asarray(s(t)) for s in sources]
dt, tmax = float(ps.params0['dt']), float(ps.params0['tmax'])
x = grid.coordsNoGhosts[0]
u = np.array(u0.array_r, dtype=float).reshape(ph.Vshape, order='F')
t, k = 0.0, 0
w0 = time.perf_counter()
marks = {}
while t <= tmax:                # the reference's loop condition (KSFD/ksfdts.py:198-202): 2081 steps
    u = O.groom(u, ph)
    u, _, _ = O.rosw_step(u, t, dt, ph, srcfn)
    t += dt
    k += 1
    if k in (30, 500, 1000, 2000):
        marks[k] = np.abs(u - exact93(x, t)).max()
wall = time.perf_counter() - w0
err = np.abs(u - exact93(x, t)).max()
amp = np.exp(LAM * t)
print('options93nx128dt1, numpy oracle + SuperLU on one CPU core: %d steps to t = %g in %.2f s wall (%.2f ms/step)'
      % (k, t, wall, 1e3 * wall / k))
print('max |u - exact| at t = %g: %.3e  (amplitude exp(lamda t) = %.1f: relative %.2e)' % (t, err, amp, err / amp))
print('on the way: ' + ', '.join('step %d: %.3e' % (kk, e) for kk, e in sorted(marks.items())))
