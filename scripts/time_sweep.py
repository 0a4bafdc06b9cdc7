"""Kernel-only timing of one Richardson sweep and the fused A*M^-1 v kernel (CUDA events, buffers
rotated over > 2.5x L2) with the tile candidate forced (`tile` option: 0 / 1 = first / second
candidate of the launcher, -1 = the planner's choice): A/B of the planner's cost model."""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, 'tests'))
import torch
import bench
from helpers import product_physics
from ksfd_b200 import core

for dim, n, reps in ((2, (1024, 1024), 20), (3, (256, 256, 256), 10)):
    for tile in (-1, 0, 1):
        ctx = core.Context(dim, n, 3)
        ctx.set_physics(product_physics(bench.phys_dict(dim, n)))
        ctx.set_option('tile', tile)
        npts, N = ctx.npts, ctx.npts * 3
        nrot = max(2, int(2.5 * 126e6 // (N * 8 * 3)) + 1)
        gen = torch.Generator(device='cuda').manual_seed(1)
        us = [(9000 + 90 * torch.randn(npts, generator=gen, device='cuda', dtype=torch.float64)).repeat_interleave(3).contiguous()
              for _ in range(nrot)]
        vs = [torch.randn(N, generator=gen, device='cuda', dtype=torch.float64) for _ in range(nrot)]
        outs = [torch.empty(N, device='cuda', dtype=torch.float64) for _ in range(nrot)]
        ctx.jvp_setup(us[0], 1.0 / (bench.ROSW_GAMMA * bench.DT))
        t_sw = bench.time_kernel(lambda i: ctx.sweep(vs[i], us[i], outs[i], norms=False), nrot, reps)
        t_jp = bench.time_kernel(lambda i: ctx.jvp(vs[i], outs[i], precond=True), nrot, reps)
        print('dim %d %s tile %2d: sweep %.1f us, A*M^-1 v %.1f us' % (dim, n, tile, t_sw, t_jp), flush=True)
        ctx.close()
        del us, vs, outs
        torch.cuda.empty_cache()
