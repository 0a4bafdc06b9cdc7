"""Launch sequence of ONE warm time step at 1024^2 with per-launch GPU times and the idle gap in
front of every launch (torch.profiler / CUPTI).  With KSFD_SEQ_PLAIN=1: no profiler, just 5 warm +
3 steps (the target of an ncu launch-list pass with DRAM byte counters, --cache-control none)."""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, 'tests'))
import numpy as np, torch
from helpers import phys84, product_physics
from ksfd_b200 import core
N = int(os.environ.get('KSFD_SEQ_N', '1024'))
n = (N, N)
ctx = core.Context(2, n, 3); ctx.set_physics(product_physics(phys84(2, n)))
rng = np.random.default_rng(np.random.SeedSequence(793817931).spawn(1)[0])
rho = 9000.0 + 90.0 * rng.standard_normal(ctx.npts)
u = ctx.upload(np.repeat(rho, 3))
opts = core.ts_options(ts_type='rosw', adapt='none', atol=0.01, rtol=1e-6, ksp_rtol=1e-8, ksp_max_it=2000, restart=30)
t = 0.0
def step():
    global t
    ctx.groom(u)
    r = ctx.ts_step(u, t, 1e-3, opts)
    t = r.t_new
    return ctx.velocity_max(u)
for _ in range(5):
    step()
torch.cuda.synchronize()
if os.environ.get('KSFD_SEQ_PLAIN'):
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    sys.exit(0)
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        step()
    torch.cuda.synchronize()
ev = sorted([e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA],
            key=lambda e: e.time_range.start)
per = len(ev) // 3
ev = ev[per:2 * per]            # the middle step
prev_end = None
tot = gap = 0.0
for i, e in enumerate(ev):
    s, d = e.time_range.start, e.device_time
    g = 0.0 if prev_end is None else s - prev_end
    prev_end = s + d
    tot += d; gap += max(g, 0.0)
    print('%4d gap %6.2f dur %7.2f  %s' % (i, g, d, e.name[:110]))
print('launches %d  kernel time %.1f us  gaps %.1f us' % (len(ev), tot, gap))
