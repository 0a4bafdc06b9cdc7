"""Wall time per ROSW step of the bench problem (2-D 1024^2, dt 1e-3), CUDA events,
no profiler.  usage: step_time.py [n] [nsteps]"""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, 'tests'))
import numpy as np, torch
from helpers import phys84, product_physics
from ksfd_b200 import core
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
NS = int(sys.argv[2]) if len(sys.argv) > 2 else 100
n = (N, N)
ctx = core.Context(2, n, 3); ctx.set_physics(product_physics(phys84(2, n)))
rng = np.random.default_rng(np.random.SeedSequence(793817931).spawn(1)[0])
rho = 9000.0 + 90.0 * rng.standard_normal(ctx.npts)
u = ctx.upload(np.repeat(rho, 3))
opts = core.ts_options(ts_type='rosw', adapt='none', atol=0.01, rtol=1e-6, ksp_rtol=1e-8, ksp_max_it=2000, restart=30)
t = 0.0; its = 0
def step():
    global t, its
    ctx.groom(u)
    r = ctx.ts_step(u, t, 1e-3, opts)
    t = r.t_new; its += r.ksp_its
    return ctx.velocity_max(u)
for _ in range(10):
    step()
torch.cuda.synchronize()
its = 0
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(NS):
    step()
e1.record(); torch.cuda.synchronize()
print('%dx%d: %.4f ms/step, %.1f its/step, spec=%s, checksum %.12e' % (N, N, e0.elapsed_time(e1) / NS, its / NS,
      os.environ.get('KSFD_GM_SPECULATE', 'default'), float(ctx.download(u).sum())))
