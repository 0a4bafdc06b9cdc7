import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, 'tests'))
import numpy as np, torch
from helpers import phys84, product_physics, random_state
from ksfd_b200 import core
p = phys84(2, (96, 64))
ctx = core.Context(2, (96, 64), 3); ctx.set_physics(product_physics(p))
u = ctx.upload(random_state(p, 5))
F = ctx.residual(u)
h = 1e-4
shift = 1.0 / (0.435866521508459 * h)
ctx.jvp_setup(u, shift)
for pipe in (0, 1, 1):
    ctx.set_option('gmres_pipeline', pipe)
    x, r = ctx.gmres(F, rtol=1e-8, max_it=500)
    res = F - ctx.jvp(x)
    print('h %g pipe %d: its %d reason %d rec %.2e true %.2e' % (h, pipe, r.its, r.reason, r.rnorm / r.rnorm0, ctx.norm2(res) / r.rnorm0), flush=True)
