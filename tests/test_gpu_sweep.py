"""
The stationary (Richardson) block-Jacobi sweeps fused into the marching stencil kernel
(csrc/sweep_op.cuh; -ksp_type richardson / the automatic choice) against the direct solve of
the oracle's assembled operator, on grids that exercise every tiling case: exact tiles,
clamped (overlapping) last tiles of the TMA-fed marcher, odd extents (register-prefetch
marcher), 3-D.  Tolerance: the solves are made to rtol 1e-12 and must agree with SuperLU
to 1e-8 relative (the conditioning of the stage matrix), the TRUE residual recomputed
through ksfd_jvp must meet the tolerance the solver reported.
"""
import numpy as np
import pytest

from helpers import oracle_physics, phys84, random_state, relerr
from test_gpu_parity import make_ctx

pytestmark = pytest.mark.gpu

GRIDS = [('2d_exact', phys84(2, (96, 64))),
         ('2d_clamped_tile', phys84(2, (302, 40))),
         ('2d_odd', phys84(2, (51, 36))),
         ('3d_exact', phys84(3, (16, 16, 8))),
         ('3d_clamped_tile', phys84(3, (22, 18, 6))),
         ('3d_odd', phys84(3, (15, 12, 6)))]


@pytest.mark.parametrize('label,p', GRIDS)
def test_one_sweep_equals_its_parts(label, p):
    """The fused kernel against the separate kernels it fuses: r_out must equal
    r - (A M^-1 r) with A M^-1 r from ksfd_jvp_precond BIT FOR BIT (same arithmetic, one
    more subtraction), x must equal x + M^-1 r from ksfd_pc_apply, and the two norms the
    device epilogue reports must be those of r and r_out."""
    import torch
    ctx = make_ctx(p)
    u = ctx.upload(random_state(p, 7))
    ctx.jvp_setup(u, 1.0 / (0.435866521508459 * 1e-3))
    gen = torch.Generator(device='cuda').manual_seed(3)
    r = torch.randn(u.numel(), generator=gen, device='cuda', dtype=torch.float64)
    x0 = torch.randn(u.numel(), generator=gen, device='cuda', dtype=torch.float64)
    t = ctx.jvp(r, precond=True)
    z = ctx.pc_apply(r)
    for first in (True, False):
        x = x0.clone()
        rout, (n0, n1) = ctx.sweep(r, x, first=first)
        assert torch.equal(rout, r - t), (label, first, float((rout - (r - t)).abs().max()))
        xe = z if first else x0 + z
        assert float((x - xe).abs().max()) <= 4e-16 * float(xe.abs().max()), (label, first)
        assert abs(n0 - float(r.norm())) <= 1e-13 * n0, (label, n0)
        assert abs(n1 - float(rout.norm())) <= 1e-13 * n1, (label, n1)
    ctx.close()


@pytest.mark.parametrize('label,p', GRIDS)
def test_sweeps_against_direct_solve(label, p):
    import scipy.sparse.linalg as spla
    from oracle import ksfd_oracle as O
    ph = oracle_physics(p)
    u = random_state(p, 21)
    rng = np.random.default_rng(22)
    b = rng.standard_normal(u.size)
    shift = 1.0 / (O.ROSW_GAMMA * 1e-3)
    x_ref = spla.splu(O.ijacobian(u, shift, ph).tocsc()).solve(b)
    ctx = make_ctx(p)
    ctx.jvp_setup(ctx.upload(u), shift)
    bd = ctx.upload(b)
    its = []
    for rep in range(3):            # launch-ahead: the later solves use the predicted length
        x, res = ctx.ksp_solve(bd, ksp_type='richardson', rtol=1e-12, max_it=200)
        assert res.reason > 0, (label, res.reason, res.its)
        its.append(res.its)
        assert relerr(ctx.download(x), x_ref) < 1e-8, (label, rep, res.its)
        true = ctx.norm2(bd - ctx.jvp(x)) / res.rnorm0
        assert true <= 1.5e-12, (label, rep, true)
        assert abs(res.rnorm / res.rnorm0 - true) <= 0.05 * true + 1e-15, (label, res.rnorm, true)
    assert len(set(its)) == 1 and its[0] <= 20, (label, its)
    # the solution does not depend on the solver: GMRES to the same tolerance
    xg, rg = ctx.gmres(bd, rtol=1e-12, max_it=2000)
    assert relerr(ctx.download(xg), ctx.download(x)) < 1e-9, label
    print('%s: %d sweeps, GMRES %d steps' % (label, its[0], rg.its))
    ctx.close()


@pytest.mark.parametrize('label,p', [('2d', phys84(2, (96, 64))), ('3d', phys84(3, (16, 16, 8)))])
def test_automatic_choice_falls_back_to_gmres(label, p):
    """Large time steps: the stationary iteration contracts slowly (or diverges); the
    automatic choice must notice on the device, hand over to GMRES from the iterate
    reached, and still return the direct solve's answer; the next solves start with
    GMRES (back-off), and small steps return to the sweeps."""
    import scipy.sparse.linalg as spla
    from oracle import ksfd_oracle as O
    ph = oracle_physics(p)
    u = random_state(p, 21)
    rng = np.random.default_rng(23)
    b = rng.standard_normal(u.size)
    ctx = make_ctx(p)
    bd = ctx.upload(b)
    for dt in (0.2, 0.02, 1e-3, 1e-5):
        shift = 1.0 / (O.ROSW_GAMMA * dt)
        x_ref = spla.splu(O.ijacobian(u, shift, ph).tocsc()).solve(b)
        ctx.jvp_setup(ctx.upload(u), shift)
        for rep in range(2):
            x, res = ctx.ksp_solve(bd, ksp_type='auto', rtol=1e-12, max_it=2000)
            assert res.reason > 0, (label, dt, res.reason, res.its)
            assert relerr(ctx.download(x), x_ref) < 1e-8, (label, dt, rep, res.its)
            true = ctx.norm2(bd - ctx.jvp(x)) / res.rnorm0
            assert true <= 1.5e-12, (label, dt, rep, true)
    # pure richardson on a step it cannot handle reports a failure instead of hanging
    ctx.jvp_setup(ctx.upload(u), 1.0 / (O.ROSW_GAMMA * 0.2))
    x, res = ctx.ksp_solve(bd, ksp_type='richardson', rtol=1e-12, max_it=40)
    assert res.reason < 0 or res.its <= 40
    ctx.close()


def test_sweep_sign_and_zero_rhs():
    """ksfd_ts_step solves A y = -F (sign applied inside the first sweep); b = 0 gives
    x = 0 and a converged reason (sweeps the prediction from earlier solves asks for are
    not tested, so the count may exceed one)."""
    from ksfd_b200 import core
    p = phys84(2, (64, 48))
    ctx = make_ctx(p)
    u = ctx.upload(random_state(p, 5))
    opts_s = core.ts_options(adapt='none', ksp_rtol=1e-12, ksp_type='richardson')
    opts_g = core.ts_options(adapt='none', ksp_rtol=1e-12, ksp_type='gmres')
    us, ug = u.clone(), u.clone()
    rs = ctx.ts_step(us, 0.0, 1e-3, opts_s)
    rg = ctx.ts_step(ug, 0.0, 1e-3, opts_g)
    assert rs.accepted and rg.accepted and not rs.ksp_fail
    d = (us - ug).abs().max().item() / ug.abs().max().item()
    assert d < 1e-11, d
    ctx.jvp_setup(u, 1000.0)
    x, res = ctx.ksp_solve(ctx.zeros(), ksp_type='richardson', rtol=1e-8)
    assert res.reason > 0 and res.its >= 1 and float(x.abs().max()) == 0.0
    ctx.close()


def test_step_flags_do_the_loop_body_in_one_call():
    """KSFD_TS_GROOM | KSFD_TS_VELOCITY_MAX: the clamp before and the CFL maxima after the
    step inside ksfd_ts_step (the body of the reference's loop, ksfdts.py:205-227) give exactly
    what the three separate calls give."""
    import torch
    from ksfd_b200 import core
    for p in (phys84(2, (64, 48)), phys84(3, (16, 16, 8))):
        # two contexts: the solver's launch-ahead prediction is per context, so both see the
        # same history and must produce the same bits
        ca, cb = make_ctx(p), make_ctx(p)
        u0 = random_state(p, 9)
        ua, ub = ca.upload(u0), cb.upload(u0)
        # (1) a tiny step from a state the clamp changes (negative and NaN entries)
        bad = torch.arange(0, ua.numel(), 97, device=ua.device)
        ua[bad] = -1.0
        ub[bad] = -1.0
        ua[5] = float('nan')
        ub[5] = float('nan')
        oa = core.ts_options(adapt='none', ksp_rtol=1e-10)
        ob = core.ts_options(adapt='none', ksp_rtol=1e-10, groom=True, velocity_max=True)
        ca.groom(ua)
        ra = ca.ts_step(ua, 0.0, 1e-9, oa)
        rb = cb.ts_step(ub, 0.0, 1e-9, ob)
        assert ra.accepted == rb.accepted == 1 and rb.have_vmax
        assert torch.equal(ua, ub), float((ua - ub).abs().max())
        # (2) ordinary steps, with and without step-size control
        ua, ub = ca.upload(u0), cb.upload(u0)
        for adapt in ('none', 'basic'):
            oa = core.ts_options(adapt=adapt, atol=0.01, rtol=1e-6, ksp_rtol=1e-10)
            ob = core.ts_options(adapt=adapt, atol=0.01, rtol=1e-6, ksp_rtol=1e-10, groom=True,
                                 velocity_max=True)
            ca.groom(ua)
            ra = ca.ts_step(ua, 0.0, 1e-4, oa)
            va = ca.velocity_max(ua)
            rb = cb.ts_step(ub, 0.0, 1e-4, ob)
            assert ra.accepted and rb.accepted and rb.have_vmax and not ra.have_vmax
            assert torch.equal(ua, ub), float((ua - ub).abs().max())
            assert rb.h_next == ra.h_next and rb.enorm == ra.enorm
            assert np.array_equal(np.array([rb.vmax[i] for i in range(p['dim'])]), va)
        ca.close()
        cb.close()
