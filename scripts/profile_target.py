"""Short program for `ncu --set full`: a few launches of the hot kernels."""
import sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, 'tests'))
import torch
from helpers import phys84, product_physics
from ksfd_b200 import core

def main():
    which = sys.argv[1] if len(sys.argv) > 1 else '2d1024'
    dim, n = {'2d1024': (2, (1024, 1024)), '2d4096': (2, (4096, 4096)),
              '3d256': (3, (256, 256, 256))}[which]
    ctx = core.Context(dim, n, 3)
    ctx.set_physics(product_physics(phys84(dim, n)))
    gen = torch.Generator(device='cuda').manual_seed(1)
    N = ctx.npts * 3
    u = (9000 + 90 * torch.randn(ctx.npts, generator=gen, device='cuda', dtype=torch.float64)).repeat_interleave(3).contiguous()
    v = torch.randn(N, generator=gen, device='cuda', dtype=torch.float64)
    out = torch.empty_like(v)
    ctx.jvp_setup(u, 1.0 / (0.435866521508459 * 1e-3))
    x = torch.zeros_like(v)
    for _ in range(3):
        ctx.residual(u, v, None, out)
        ctx.jvp(v, out, precond=True)
        ctx.jvp(v, out)
        ctx.sweep(v, x, out, norms=False)
    torch.cuda.synchronize()
    print('done', ctx.norm2(out))
    ctx.close()

if __name__ == '__main__':
    main()
