"""
Multi-GPU parity check (run under torchrun, one rank per GPU):
the slab-decomposed operator and time step over N ranks must reproduce the
single-GPU result of the same global problem.

    torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 scripts/multi_gpu_check.py
"""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, 'tests'))
import numpy as np, torch, torch.distributed as dist
from helpers import phys84, product_physics, random_state
from ksfd_b200 import core, parallel

def main():
    rank = int(os.environ['RANK']); world = int(os.environ['WORLD_SIZE']); local = int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    ok_all = True
    # >= 4 planes per rank at 8 ranks: same kernel family as one rank; the last case forces chunks of 3
    # planes so that the two-pass boundary CTAs of the sweep kernel (MarchArgs::rb) also run in 3-D
    for dim, n, rz in ((2, (96, 64), 0), (3, (20, 24, 32), 0), (1, (64,), 0), (2, (256, 1024), 0),
                       (3, (32, 32, 48), 3)):
        p = phys84(dim, n)
        dof = 3
        u_g = random_state(p, 5)
        rng = np.random.default_rng(6)
        ud_g = rng.standard_normal(u_g.size); v_g = rng.standard_normal(u_g.size)
        ctx = core.Context(dim, n, dof, device=local, rank=rank, nranks=world)
        ctx.set_physics(product_physics(p))
        parallel.init_comm(ctx)
        if rz:
            ctx.set_option('rz', rz)
        plane = dof * int(np.prod(n[:-1]))
        sl = slice(ctx.last_start * plane, (ctx.last_start + ctx.last_count) * plane)
        u, ud, v = ctx.upload(u_g[sl]), ctx.upload(ud_g[sl]), ctx.upload(v_g[sl])
        shift = 1.0 / (0.435866521508459 * 1e-3)
        F = ctx.download(ctx.residual(u, ud))
        ctx.jvp_setup(u, shift)
        Jv = ctx.download(ctx.jvp(v)); Jvp = ctx.download(ctx.jvp(v, precond=True))
        vm = ctx.velocity_max(u)
        nrm = ctx.norm2(v); wsum = ctx.sum_dof0(u)
        x, res = ctx.gmres(ud, rtol=1e-10, max_it=500)
        xs = ctx.download(x)
        # stationary sweeps fused into the stencil kernel (peer-memory exchange only)
        sweeps = dim >= 2 and os.environ.get('KSFD_HALO_P2P', '1') != '0'
        if sweeps:
            xw = ctx.zeros()
            rw, nw = ctx.sweep(v, xw, first=True)
            rws, xws = ctx.download(rw), ctx.download(xw)
            xr, rr = ctx.ksp_solve(ud, ksp_type='richardson', rtol=1e-10, max_it=100)
            xrs = ctx.download(xr)
        opts = core.ts_options(adapt='basic', atol=0.01, rtol=1e-6, clip=(0.1, 5.0), ksp_rtol=1e-12, ksp_max_it=500)
        uu = u.clone(); t, h = 0.0, 1e-6
        for k in range(3):
            ctx.groom(uu); r = ctx.ts_step(uu, t, h, opts); t, h = r.t_new, r.h_next
            assert r.accepted == 1 and r.ksp_fail == 0, ('multi-rank step failed', dim, n, k, r.ksp_its)
        us = ctx.download(uu)
        ctx.close()
        if rank == 0:
            c1 = core.Context(dim, n, dof, device=local)
            c1.set_physics(product_physics(p))
            if rz:
                c1.set_option('rz', rz)
            U, UD, V = c1.upload(u_g), c1.upload(ud_g), c1.upload(v_g)
            F1 = c1.download(c1.residual(U, UD)); c1.jvp_setup(U, shift)
            Jv1 = c1.download(c1.jvp(V)); Jvp1 = c1.download(c1.jvp(V, precond=True))
            vm1 = c1.velocity_max(U); nrm1 = c1.norm2(V); wsum1 = c1.sum_dof0(U)
            x1, res1 = c1.gmres(UD, rtol=1e-10, max_it=500); xs1 = c1.download(x1)
            if sweeps:
                xw1 = c1.zeros()
                rw1, nw1 = c1.sweep(V, xw1, first=True)
                rws1, xws1 = c1.download(rw1), c1.download(xw1)
                xr1, rr1 = c1.ksp_solve(UD, ksp_type='richardson', rtol=1e-10, max_it=100)
                xrs1 = c1.download(xr1)
            uu1 = U.clone(); t1, h1 = 0.0, 1e-6
            for k in range(3):
                c1.groom(uu1); r1 = c1.ts_step(uu1, t1, h1, opts); t1, h1 = r1.t_new, r1.h_next
                assert r1.accepted == 1 and r1.ksp_fail == 0, ('single-rank step failed', dim, n, k)
            us1 = c1.download(uu1)
            c1.close()
            def rel(a, b): return float(np.abs(a - b).max() / np.abs(b).max())
            errs = dict(F=rel(F, F1[sl]), Jv=rel(Jv, Jv1[sl]), Jvp=rel(Jvp, Jvp1[sl]),
                        vmax=rel(vm, vm1), norm=abs(nrm - nrm1) / nrm1, sum=abs(wsum - wsum1) / wsum1,
                        gmres=rel(xs, xs1[sl]), ts=rel(us, us1[sl]), t=abs(t - t1), h=abs(h - h1) / h1)
            if sweeps:
                errs.update(sweep_r=rel(rws, rws1[sl]), sweep_x=rel(xws, xws1[sl]),
                            sweep_norms=max(abs(nw[0] - nw1[0]) / nw1[0], abs(nw[1] - nw1[1]) / nw1[1]),
                            richardson=rel(xrs, xrs1[sl]), richardson_its=float(abs(rr.its - rr1.its)))
            ok = all(e < 1e-9 for e in errs.values()) and errs['F'] == 0.0 and errs['Jv'] == 0.0
            if sweeps:
                ok = ok and errs['sweep_r'] == 0.0 and errs['sweep_x'] == 0.0
            ok_all = ok_all and ok
            print('dim', dim, n, 'ranks', world, 'OK' if ok else 'FAIL', errs, 'its', res.its, res1.its, flush=True)
        dist.barrier()
    if rank == 0:
        print('MULTI_GPU_CHECK', 'PASS' if ok_all else 'FAIL', flush=True)
    dist.destroy_process_group()

if __name__ == '__main__':
    main()
