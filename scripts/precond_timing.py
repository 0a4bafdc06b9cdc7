"""Arnoldi steps and wall time of one linear solve with point-block Jacobi (pc 1) and the
spectral preconditioner (pc 2) over dt = 1e-3 .. 100, 2-D 1024^2 and 3-D 128^3 (run on a GPU box;
output kept in profiles/r01_spectral_pc_timing.txt)."""
import os, sys, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, 'tests'))
import numpy as np, torch
from helpers import phys84, product_physics
from ksfd_b200 import core
for dim, n in ((2, (1024, 1024)), (3, (128, 128, 128))):
    ctx = core.Context(dim, n, 3); ctx.set_physics(product_physics(phys84(dim, n)))
    rng = np.random.default_rng(1)
    rho = 9000.0 + 90.0 * rng.standard_normal(ctx.npts)
    u = ctx.upload(np.repeat(rho, 3))
    F = ctx.residual(u)
    for dt in (1e-3, 1e-1, 1.0, 100.0):
        ctx.jvp_setup(u, 1.0 / (0.435866521508459 * dt))
        for pc in (1, 2):
            ctx.gmres(F, rtol=1e-8, max_it=3000, precond=pc)
            torch.cuda.synchronize(); t0 = time.perf_counter()
            x, r = ctx.gmres(F, rtol=1e-8, max_it=3000, precond=pc)
            torch.cuda.synchronize(); ms = (time.perf_counter() - t0) * 1e3
            true = ctx.norm2(F - ctx.jvp(x)) / r.rnorm0
            print('dim %d dt %-6g pc %d: its %4d reason %2d true %.2e  %.2f ms' % (dim, dt, pc, r.its, r.reason, true, ms), flush=True)
    ctx.close()
