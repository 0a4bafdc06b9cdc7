"""Kernel timing sweep on one GPU (development tool; not the bench)."""
import sys, os, json, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np, torch
from helpers import phys84, product_physics
from ksfd_b200 import core

def timeit(fn, nrot, reps=20, warm=3):
    for i in range(warm): fn(i % nrot)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps+1)]
    ev[0].record()
    for i in range(reps):
        fn(i % nrot); ev[i+1].record()
    torch.cuda.synchronize()
    ts = [ev[i].elapsed_time(ev[i+1]) for i in range(reps)]
    return float(np.median(ts))*1e3, float(np.min(ts))*1e3   # us

def run(dim, n, sweeps):
    p = phys84(dim, n)
    ctx = core.Context(dim, n, 3); ctx.set_physics(product_physics(p))
    npts = ctx.npts; N = npts*3
    nrot = max(2, int(400e6 // (N*8*3)) + 1)   # rotate buffers > 2x L2
    gen = torch.Generator(device='cuda').manual_seed(1)
    us = [(9000+90*torch.randn(npts, generator=gen, device='cuda', dtype=torch.float64)).repeat_interleave(3).contiguous() for _ in range(nrot)]
    uds = [torch.randn(N, generator=gen, device='cuda', dtype=torch.float64) for _ in range(nrot)]
    outs = [torch.empty(N, device='cuda', dtype=torch.float64) for _ in range(nrot)]
    ctx.jvp_setup(us[0], 1.0/(0.435866521508459*1e-3))
    out = []
    for opt in sweeps:
        ctx.set_option('variant', opt.get('variant', 0)); ctx.set_option('tile', opt.get('tile', -1)); ctx.set_option('rz', opt.get('rz', 0))
        try:
            r = timeit(lambda i: ctx.residual(us[i], uds[i], None, outs[i]), nrot)
            j = timeit(lambda i: ctx.jvp(uds[i], outs[i]), nrot)
            jp = timeit(lambda i: ctx.jvp(uds[i], outs[i], precond=True), nrot)
            v = timeit(lambda i: ctx.velocity_max(us[i]), nrot, reps=5)
        except Exception as e:
            print('ERR', opt, e); continue
        gp = npts/1e3
        rec = dict(n=list(n), opt=opt, residual_us=r, jvp_us=j, jvp_pc_us=jp, velmax_us=v,
                   residual_gpts=npts/r[0]/1e3, jvp_gpts=npts/j[0]/1e3,
                   residual_frac=npts*72/r[0]/1e3/6544.7, jvp_frac=npts*72/j[0]/1e3/6544.7)
        print(json.dumps(rec), flush=True); out.append(rec)
    ctx.close()
    return out

if __name__ == '__main__':
    which = sys.argv[1] if len(sys.argv) > 1 else 'all'
    if which in ('all','2d'):
        sw = [dict(variant=0), dict(variant=1)]
        for tile, rz in itertools.product((0, 1), (8, 12, 16, 24, 32, 64)):
            sw.append(dict(variant=2, tile=tile, rz=rz))
        run(2, (1024,1024), sw)
        run(2, (4096,4096), [dict(variant=0)] + [dict(variant=2, tile=t, rz=rz) for t in (0,1) for rz in (32, 64, 128)])
    if which in ('all','3d'):
        sw = [dict(variant=0), dict(variant=1)]
        for tile, rz in itertools.product((0, 1), (16, 32, 64, 128, 256)):
            sw.append(dict(variant=2, tile=tile, rz=rz))
        run(3, (256,256,256), sw)
