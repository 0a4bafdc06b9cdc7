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
world = int(os.environ.get('WORLD_SIZE', '1')); rank = int(os.environ.get('RANK', '0'))
n = (N, N * world)          # several ranks (torchrun): one N^2 tile per GPU, rank 0's timeline
if world > 1:
    import torch.distributed as dist
    torch.cuda.set_device(int(os.environ['LOCAL_RANK']))
    dist.init_process_group('nccl', device_id=torch.device('cuda', int(os.environ['LOCAL_RANK'])))
ctx = core.Context(2, n, 3, device=int(os.environ.get('LOCAL_RANK', '0')), rank=rank, nranks=world)
ctx.set_physics(product_physics(phys84(2, n)))
if world > 1:
    from ksfd_b200 import parallel
    parallel.init_comm(ctx)
for k in ('sweep_test_lead', 'sweep_fuse_push', 'fuse_push_mask'):
    if os.environ.get('KSFD_OPT_' + k.upper()):
        ctx.set_option(k, int(os.environ['KSFD_OPT_' + k.upper()]))
rng = np.random.default_rng(np.random.SeedSequence(793817931).spawn(world)[rank])
rho = 9000.0 + 90.0 * rng.standard_normal(ctx.npts)
u = ctx.upload(np.repeat(rho, 3))
opts = core.ts_options(ts_type='rosw', adapt='none', atol=0.01, rtol=1e-6, ksp_rtol=1e-8, ksp_max_it=2000, restart=30,
                       groom=True, velocity_max=True)
t = 0.0
def step():
    global t
    r = ctx.ts_step(u, t, 1e-3, opts)       # clamp + step + CFL maxima in one call
    t = r.t_new
    return r.vmax[0]
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
if world > 1:
    dist.barrier()
prev_end = None
tot = gap = 0.0
for i, e in enumerate(ev):
    s, d = e.time_range.start, e.device_time
    g = 0.0 if prev_end is None else s - prev_end
    prev_end = s + d
    tot += d; gap += max(g, 0.0)
    if rank == 0:
        print('%4d gap %6.2f dur %7.2f  %s' % (i, g, d, e.name[:110]))
if world > 1:
    # start / duration of every stencil-sweep launch on every rank (CUPTI timestamps share the
    # host's time base): skew between the ranks
    rows = [(e.time_range.start, e.device_time) for e in ev if 'SweepOp' in e.name]
    t = torch.tensor(rows, dtype=torch.float64, device='cuda')
    allt = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(allt, t)
    if rank == 0:
        t0 = float(allt[0][0, 0])
        for i in range(t.shape[0]):
            print('sweep %2d ' % i + '  '.join('r%d start %8.1f dur %5.1f' % (r, float(a[i, 0]) - t0, float(a[i, 1]))
                                                for r, a in enumerate(allt)))
if rank == 0:
    sw = sorted(e.device_time for e in ev if 'SweepOp' in e.name and e.device_time > 5)
    print('sweeps %d median %.1f us; residual %s' % (len(sw), sw[len(sw) // 2], ['%.1f' % e.device_time for e in ev if 'ResidualOp' in e.name]))
    print('launches %d  kernel time %.1f us  gaps %.1f us' % (len(ev), tot, gap))
