"""Driver for an ncu capture of the Krylov / spectral-preconditioner kernels:
one block-Jacobi solve (dt 1e-3) and one spectral solve (dt 1) on an N^2 grid.
usage: profile_aux.py N"""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, 'tests'))
import numpy as np, torch
from helpers import phys84, product_physics
from ksfd_b200 import core
N = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
n = (N, N)
ctx = core.Context(2, n, 3); ctx.set_physics(product_physics(phys84(2, n)))
rng = np.random.default_rng(1)
rho = 9000.0 + 90.0 * rng.standard_normal(ctx.npts)
u = ctx.upload(np.repeat(rho, 3))
F = ctx.residual(u)
for dt, pc in ((1e-3, 1), (1.0, 2)):
    ctx.jvp_setup(u, 1.0 / (0.435866521508459 * dt))
    for rep in range(2):            # second solve of each kind is the warm one
        x, r = ctx.gmres(F, rtol=1e-8, max_it=200, precond=pc)
    torch.cuda.synchronize()
    print('N %d dt %g pc %d: its %d reason %d' % (N, dt, pc, r.its, r.reason), flush=True)
ctx.close()
