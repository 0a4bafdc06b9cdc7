import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, 'tests'))
import numpy as np, torch
from helpers import phys84, product_physics
from ksfd_b200 import core
n = (1024, 1024)
ctx = core.Context(2, n, 3); ctx.set_physics(product_physics(phys84(2, n)))
rng = np.random.default_rng(np.random.SeedSequence(793817931).spawn(1)[0])
rho = 9000.0 + 90.0 * rng.standard_normal(ctx.npts)
u = ctx.upload(np.repeat(rho, 3))
if len(sys.argv) > 1:
    ctx.set_option('gmres_cycle_exp', int(sys.argv[1]))
opts = core.ts_options(ts_type='rosw', adapt='none', atol=0.01, rtol=1e-6, ksp_rtol=1e-8, ksp_max_it=2000, restart=30)
t = 0.0
for k in range(3):
    ctx.groom(u)
    r = ctx.ts_step(u, t, 1e-3, opts)
    t = r.t_new
    print('step', k, 'its', r.ksp_its, 'acc', r.accepted, 'enorm %.3e' % r.enorm, flush=True)
