"""GMRES experiment on the benchmark problem: iterations / recurrence vs true
residual for several cycle-closing factors (run on a GPU box)."""
import os
import sys
import time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
sys.path.insert(0, os.path.join(R, 'tests'))
import numpy as np
import torch
from helpers import phys84, product_physics
from ksfd_b200 import core

for dim, n, dts in ((2, (1024, 1024), (1e-3, 1e-1, 10.0)), (3, (96, 96, 96), (1e-3, 1.0))):
    ctx = core.Context(dim, n, 3)
    ctx.set_physics(product_physics(phys84(dim, n)))
    rng = np.random.default_rng(np.random.SeedSequence(793817931).spawn(1)[0])
    rho = 9000.0 + 90.0 * rng.standard_normal(ctx.npts)
    u = ctx.upload(np.repeat(rho, 3))
    F = ctx.residual(u)
    for dt in dts:
        shift = 1.0 / (0.435866521508459 * dt)
        ctx.jvp_setup(u, shift)
        for rtol in (1e-8, 1e-12):
            for cexp in (5, 7, 9, 0):
                ctx.set_option('gmres_cycle_exp', cexp)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                x, r = ctx.gmres(F, rtol=rtol, max_it=500)
                torch.cuda.synchronize()
                ms = (time.perf_counter() - t0) * 1e3
                res = F - ctx.jvp(x)
                print('dim %d dt %-6g rtol %-6g cycle 1e-%d: its %3d reason %2d rec %.2e true %.2e  %.2f ms'
                      % (dim, dt, rtol, cexp, r.its, r.reason, r.rnorm / r.rnorm0,
                         ctx.norm2(res) / r.rnorm0, ms))
    ctx.close()
