"""CPU figure beside a GPU step: the oracle's C + OpenMP restatement (oracle/ksfd_oracle_c.c) on the
benchmark problem (options84 physics, h = 1/384, dt = 1e-3, rtol 1e-8) at any grid, e.g.
    python scripts/cpu_oracle_step.py 256 256 256      # BASELINE configs[3], one process, all threads
Prints seconds per step (clamp + ROSW step + CFL maxima), sweeps per step, residual and J.v rates."""
import os, sys, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, 'tests'))
import numpy as np
from helpers import oracle_physics, phys84, random_state
from oracle import ksfd_oracle_c as OC

n = tuple(int(a) for a in sys.argv[1:4]) or (1024, 1024)
nsteps = int(os.environ.get('STEPS', '3'))
p = phys84(len(n), n)
c = OC.COracle(oracle_physics(p))
u = np.ascontiguousarray(random_state(p, 100, rel=0.0))
c.ts_step(u, 1e-3, rtol=1e-8)                   # warm-up
t0 = time.perf_counter()
its = 0
for _ in range(nsteps):
    its += c.ts_step(u, 1e-3, rtol=1e-8)[1]
dt = (time.perf_counter() - t0) / nsteps
npts = int(np.prod(n))
t1 = time.perf_counter(); c.dfdt(u); tr = time.perf_counter() - t1
c.jvp_setup(u, 1.0 / (0.435866521508459 * 1e-3))
v = np.random.default_rng(1).standard_normal(u.size)
t1 = time.perf_counter(); c.jvp(v); tj = time.perf_counter() - t1
print('C oracle, %d threads, grid %s: %.3f s per step (%.2f Mpts*steps/s), %.1f sweeps per step; '
      'residual %.1f Mpts/s, J.v %.1f Mpts/s'
      % (OC.threads(), 'x'.join(map(str, n)), dt, npts / dt / 1e6, its / nsteps, npts / tr / 1e6, npts / tj / 1e6))
