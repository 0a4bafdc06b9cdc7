"""Warm per-kernel GPU times of the time step (torch.profiler / CUPTI), to compare
summed kernel time with the wall time of a step."""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, 'tests'))
import numpy as np, torch
from torch.profiler import profile, ProfilerActivity
from helpers import phys84, product_physics
from ksfd_b200 import core
n = (1024, 1024)
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
NS = 10
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(NS):
    step()
e1.record(); torch.cuda.synchronize()
print('wall ms/step (no profiler): %.3f' % (e0.elapsed_time(e1) / NS))
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(NS):
        step()
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
tot = sum(e.device_time for e in ev)
print('kernels per step %.1f, summed GPU kernel time %.3f ms/step' % (len(ev) / NS, tot / NS / 1e3))
agg = {}
for e in ev:
    k = e.name.split('(')[0][:70]
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += e.device_time
for k, (c, tt) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:28]:
    print('%6.1f/step %9.1f us/step  avg %7.2f us  %s' % (c / NS, tt / NS, tt / c, k))
