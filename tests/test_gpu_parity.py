"""
Parity of the CUDA path (through the C ABI) against
  (1) golden vectors from the reference's own code (tests/golden/*.npz),
  (2) the numpy oracle on seeded inputs at sizes it finishes in seconds,
  (3) size-independent properties at the BASELINE sizes (1024^2, 256^3).

Tolerances (fp64, max-norm relative to the largest entry of each dof):
  residual / velocity   1e-11   (CUDA libm log/tanh 1-2 ulp + FMA contraction,
                                 amplified by the stencil's cancellation)
  J.v vs assembled J@v  1e-11
  block-Jacobi blocks   1e-11
"""
import os

import numpy as np
import pytest

from helpers import (GOLD, check_field, cond_scale, golden_names, load_golden,
                     oracle_physics, phys84, product_physics, random_state,
                     relerr)

pytestmark = pytest.mark.gpu

TOL_F = 1e-11
TOL_J = 1e-11


def _torch():
    import torch
    return torch


def make_ctx(p, variant=0):
    from ksfd_b200 import core
    dof = 1 + sum(len(g[2]) for g in p['groups'])
    ctx = core.Context(p['dim'], p['n'], dof)
    ctx.set_physics(product_physics(p))
    ctx.set_option('variant', variant)
    return ctx


def dev(a):
    """plain H2D copy (layout-agnostic data: BLAS-1 tests)"""
    torch = _torch()
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).cuda()


def variants_for(p):
    dof = 1 + sum(len(g[2]) for g in p['groups'])
    if p['dim'] == 1 or dof - 1 > 4:
        return [1]
    # 1 direct kernels, 2 marching (J.v through the TMA-fed marcher where the grid is
    # eligible), 3 marching with the register-prefetch kernels only
    return [1, 2, 3]


def per_dof_err(a, b, dof):
    a = np.asarray(a).ravel()
    b = np.asarray(b).ravel()
    scale = max(np.abs(b).max(), 1e-300)
    worst = 0.0
    for c in range(dof):
        ref = max(np.abs(b[c::dof]).max(), 1e-9 * scale)
        worst = max(worst, np.abs(a[c::dof] - b[c::dof]).max() / ref)
    return worst


# ---------------------------------------------------------------------------
# (1) golden vectors from the reference
# ---------------------------------------------------------------------------
@pytest.mark.parametrize('name', golden_names())
def test_golden_residual_velocity_jvp(name):
    g, physs, nrec = load_golden(name)
    for r in range(nrec):
        p = physs[r]
        for variant in variants_for(p):
            ctx = make_ctx(p, variant)
            dof = ctx.dof
            u = ctx.upload(g['u_%d' % r])
            src = ctx.upload(g['src_%d' % r])
            f = ctx.download(ctx.residual(u, None, src))
            cond = cond_scale(oracle_physics(p), g['u_%d' % r])
            assert check_field(f, g['f_%d' % r], dof, TOL_F, cond) < 1.0, \
                (name, r, variant)
            # the field vector itself is not modified (clamp is on the fly)
            assert np.array_equal(ctx.download(u), g['u_%d' % r], equal_nan=True)
            vel = ctx.download(ctx.velocity(u), ctx.dim)
            vr = g['vel_%d' % r]
            assert (relerr(vel, vr) < TOL_F or np.abs(vel - vr).max() < 1e-12), \
                (name, r, variant)
            vm = ctx.velocity_max(u)
            vref = np.abs(vr.reshape(-1, p['dim'])).max(axis=0)
            assert np.allclose(vm, vref, rtol=1e-10, atol=1e-13)
            ctx.jvp_setup(u, 0.0)
            for m in range(3):
                v = ctx.upload(g['v_%d' % r][m])
                jv = -ctx.download(ctx.jvp(v))
                assert per_dof_err(jv, g['Jv_%d' % r][m], dof) < TOL_J, \
                    (name, r, variant, m)
            ctx.close()


# ---------------------------------------------------------------------------
# (2) oracle on seeded inputs
# ---------------------------------------------------------------------------
CASES = [
    ('1d', phys84(1, (200,))),
    ('2d', phys84(2, (96, 80))),
    ('2d_thin', phys84(2, (8, 40))),
    ('2d_odd', phys84(2, (131, 67))),
    ('3d', phys84(3, (24, 20, 28))),
    ('3d_odd', phys84(3, (37, 9, 11))),
]


@pytest.mark.parametrize('label,p', CASES)
def test_oracle_ifunction_jvp_pc(label, p):
    from oracle import ksfd_oracle as O
    ph = oracle_physics(p)
    u = random_state(p, 11)
    rng = np.random.default_rng(12)
    udot = rng.standard_normal(u.size)
    v = rng.standard_normal(u.size)
    shift = 1.0 / (O.ROSW_GAMMA * 1e-3)
    F_ref = O.ifunction(u, udot, ph).reshape(-1, order='F')
    Jv_ref = O.jvp(u, v, shift, ph).reshape(-1, order='F')
    blocks_ref = O.block_diagonal(u, shift, ph)
    for variant in variants_for(p):
        ctx = make_ctx(p, variant)
        F = ctx.download(ctx.residual(ctx.upload(u), ctx.upload(udot)))
        assert check_field(F, F_ref, ph.dof, TOL_F, cond_scale(ph, u)) < 1.0, \
            (label, variant)
        ctx.jvp_setup(ctx.upload(u), shift)
        Jv = ctx.download(ctx.jvp(ctx.upload(v)))
        assert per_dof_err(Jv, Jv_ref, ph.dof) < TOL_J, (label, variant)
        blocks = ctx.block_diagonal().cpu().numpy()
        assert relerr(blocks, blocks_ref) < TOL_J
        # M^{-1} really inverts the diagonal blocks
        z = ctx.download(ctx.pc_apply(ctx.upload(v))).reshape(-1, ph.dof)
        back = np.einsum('prc,pc->pr', blocks_ref, z).reshape(-1)
        assert relerr(back, v) < 1e-10
        # fused A*M^{-1} == A(M^{-1} v)
        fused = ctx.download(ctx.jvp(ctx.upload(v), precond=True))
        two = ctx.download(ctx.jvp(ctx.pc_apply(ctx.upload(v))))
        assert per_dof_err(fused, two, ph.dof) < 1e-12
        ctx.close()


def test_clamp_and_nan_handling():
    from oracle import ksfd_oracle as O
    p = phys84(2, (40, 36))
    ph = oracle_physics(p)
    u = random_state(p, 5)
    rng = np.random.default_rng(6)
    idx = rng.integers(0, u.size, 60)
    u[idx[:20]] = -3.0
    u[idx[20:40]] = np.nan
    u[idx[40:]] = 0.0
    f_ref = O.dfdt(u.copy(), ph).reshape(-1, order='F')
    for variant in (1, 2):
        ctx = make_ctx(p, variant)
        f = ctx.download(ctx.residual(ctx.upload(u)))
        assert check_field(f, f_ref, ph.dof, TOL_F, cond_scale(ph, u)) < 1.0
        ug = ctx.download(ctx.groom(ctx.upload(u)))
        ref = O.groom(u.copy().reshape(ph.Vshape, order='F'), ph).reshape(-1, order='F')
        assert np.array_equal(ug, ref)                       # bit exact
        assert np.array_equal(ctx.download(ctx.groom(ctx.upload(ug))), ug)   # idempotent
        ctx.close()


def test_blas1_against_numpy():
    p = phys84(2, (64, 48))
    ctx = make_ctx(p)
    rng = np.random.default_rng(3)
    n = ctx.npts * ctx.dof
    vs = [rng.standard_normal(n) for _ in range(11)]
    w = rng.standard_normal(n)
    d = ctx.mdot([dev(v) for v in vs], dev(w)).cpu().numpy()
    assert np.allclose(d, [v @ w for v in vs], rtol=1e-12, atol=1e-9)
    y = dev(w)
    coefs = rng.standard_normal(11)
    ctx.maxpy(y, coefs, [dev(v) for v in vs])
    assert np.allclose(y.cpu().numpy(), w + sum(c * v for c, v in zip(coefs, vs)),
                       rtol=1e-13, atol=1e-12)
    assert abs(ctx.norm2(dev(w)) - np.linalg.norm(w)) < 1e-10
    assert abs(ctx.sum_dof0(ctx.upload(w)) - w[0::ctx.dof].sum()) < 1e-9
    # layout converters are exact inverse permutations with the documented map
    wi = ctx.upload(w)
    assert np.array_equal(ctx.download(wi), w)
    nx, ny = ctx.local_shape
    ref = w.reshape(ny, nx, ctx.dof).transpose(0, 2, 1).reshape(-1)   # (k, c, x)
    assert np.array_equal(wi.cpu().numpy(), ref)
    ctx.close()


def _box_errors(ph_box, u_g, v_g, ud_box, F_box, Jv_box, shift, cond):
    """worst err/allowed of the residual and relative J.v error on one box"""
    from oracle import ksfd_oracle as O
    F_ref = ud_box - O.dfdt_ghosted(u_g, ph_box)
    Jv_ref = O.jvp_ghosted(u_g, v_g, shift, ph_box)
    dof = ph_box.dof
    wf = check_field(F_box.reshape(-1, order='F'), F_ref.reshape(-1, order='F'), dof, TOL_F,
                     cond)
    ej = per_dof_err(Jv_box.reshape(-1, order='F'), Jv_ref.reshape(-1, order='F'), dof)
    ef = per_dof_err(F_box.reshape(-1, order='F'), F_ref.reshape(-1, order='F'), dof)
    return wf, ef, ej


@pytest.mark.parametrize('label,p,boxes', [
    # the whole 1024^2 grid in one piece
    ('1024^2', phys84(2, (1024, 1024)), [((0, 0), (1024, 1024))]),
    # 256^3: boxes that straddle tile seams (multiples of 8/16/32 in x and y), every
    # chunk seam of the marching axis (full z extent) and all three periodic
    # boundaries (start near the upper end of an axis, wrap to its beginning)
    ('256^3', phys84(3, (256, 256, 256)), [((0, 0, 0), (40, 24, 256)),
                                           ((243, 250, 0), (26, 20, 256)),
                                           ((120, 56, 200), (48, 80, 112)),
                                           ((0, 100, 250), (256, 12, 12))]),
])
def test_full_size_vs_oracle(label, p, boxes):
    """BASELINE sizes against the numpy oracle itself (VERDICT r1 "What's weak" 1):
    the residual F = udot - f(u) and (shift*I - J) v of the CUDA path, compared on
    the full 1024^2 grid and on ghosted sub-boxes of the 256^3 grid — where tile
    pitch, planes per CTA, the periodic wrap of every box load and 32-bit indexing
    differ from the 8^2 .. 96^2 golden grids.  Prints the achieved errors."""
    from oracle import ksfd_oracle as O
    torch = _torch()
    ctx = make_ctx(p, 2)
    dof, n = ctx.dof, tuple(p['n'])
    rng = np.random.default_rng(793817931)
    ua = np.empty((dof,) + n)
    ua[0] = 9000.0 + 90.0 * rng.standard_normal(n)
    for l in range(1, dof):
        ua[l] = ua[0] * (1.0 + 0.01 * rng.standard_normal(n))
    uda = rng.standard_normal((dof,) + n)
    va = rng.standard_normal((dof,) + n)
    shift = 1.0 / (O.ROSW_GAMMA * 1e-3)
    u, ud, v = (ctx.upload(a.reshape(-1, order='F')) for a in (ua, uda, va))
    F = ctx.download(ctx.residual(u, ud)).reshape((dof,) + n, order='F')
    ctx.jvp_setup(u, shift)
    Jv = ctx.download(ctx.jvp(v)).reshape((dof,) + n, order='F')
    # A*M^-1 v (the fused kernel of the Krylov loop) == A applied to M^-1 v
    AMv = ctx.download(ctx.jvp(v, precond=True))
    AMv2 = ctx.download(ctx.jvp(ctx.pc_apply(v)))
    assert per_dof_err(AMv, AMv2, dof) < 1e-12, label
    cond = cond_scale(oracle_physics(phys84(p['dim'], (16,) * p['dim'])),
                      ua[(slice(None),) + (slice(0, 16),) * p['dim']].reshape(-1, order='F'))
    worst = (0.0, 0.0, 0.0)
    for lo, size in boxes:
        pb = phys84(p['dim'], size)
        phb = oracle_physics(pb)
        sl = (slice(None),) + tuple(slice(None) for _ in size)
        cut = lambda a: O.cut_box(a, lo, size, sw=0)
        e = _box_errors(phb, O.cut_box(ua, lo, size), O.cut_box(va, lo, size), cut(uda),
                        cut(F), cut(Jv), shift, cond)
        print('%s box lo=%s size=%s: residual err/allowed %.3f (rel %.2e), J.v rel %.2e'
              % (label, lo, size, e[0], e[1], e[2]))
        worst = tuple(max(a, b) for a, b in zip(worst, e))
    print('%s worst: residual err/allowed %.3f, residual rel %.2e, J.v rel %.2e' % (
        (label,) + worst))
    assert worst[0] < 1.0 and worst[2] < TOL_J, (label, worst)
    ctx.close()


@pytest.mark.parametrize('label,p', [('1d', phys84(1, (128,), h=1.0 / 128)),
                                     ('2d', phys84(2, (40, 32))),
                                     ('3d', phys84(3, (12, 10, 14)))])
def test_gmres_against_direct_solve(label, p):
    import scipy.sparse.linalg as spla
    from oracle import ksfd_oracle as O
    ph = oracle_physics(p)
    u = random_state(p, 21)
    rng = np.random.default_rng(22)
    b = rng.standard_normal(u.size)
    for shift in (1.0 / (O.ROSW_GAMMA * 1e-3), 1.0 / (O.ROSW_GAMMA * 1.0)):
        A = O.ijacobian(u, shift, ph).tocsc()
        x_ref = spla.splu(A).solve(b)
        ctx = make_ctx(p)
        ctx.jvp_setup(ctx.upload(u), shift)
        for reorth in (0, 1):
            x, res = ctx.gmres(ctx.upload(b), rtol=1e-12, max_it=2000, restart=30,
                               reorth=reorth)
            assert res.reason > 0, (label, shift, res.reason, res.its)
            assert relerr(ctx.download(x), x_ref) < 1e-8, (label, shift, res.its)
        ctx.close()


@pytest.mark.parametrize('label,p', [('2d', phys84(2, (96, 64))), ('3d', phys84(3, (20, 24, 16))),
                                     ('1d', phys84(1, (200,)))])
def test_gmres_pipeline_matches_sync(label, p):
    """The pipelined (device-decided, launch-ahead) GMRES and the host-driven
    one are the same algorithm: same iteration counts, same solution, and the
    TRUE residual meets the tolerance, over repeated solves and a range of
    shifts (time steps 1e-6 .. 1e-1)."""
    ctx = make_ctx(p)
    u = ctx.upload(random_state(p, 5))
    F = ctx.residual(u)
    for h in (1e-6, 1e-4, 1e-3, 1e-1):
        ctx.jvp_setup(u, 1.0 / (0.435866521508459 * h))
        for rtol in (1e-8, 1e-12):
            ctx.set_option('gmres_pipeline', 0)
            x0, r0 = ctx.gmres(F, rtol=rtol, max_it=500)
            x0 = x0.clone()
            ctx.set_option('gmres_pipeline', 1)
            for rep in range(4):
                x1, r1 = ctx.gmres(F, rtol=rtol, max_it=500)
                assert r1.reason > 0 and r1.its == r0.its, (label, h, rtol, rep, r0.its, r1.its)
                true = ctx.norm2(F - ctx.jvp(x1)) / r1.rnorm0
                assert true <= 1.5 * rtol, (label, h, rtol, rep, true)
                d = (x1 - x0).abs().max().item() / x0.abs().max().item()
                # both meet the tolerance; they differ in how a cancelled
                # Gram-Schmidt norm is handled, so agree to the tolerance only
                assert d < 10 * rtol, (label, h, rtol, rep, d)
    ctx.close()


@pytest.mark.parametrize('label,p,h', [('1d', phys84(1, (64,), h=1.0 / 64), 0.5),
                                       ('2d', phys84(2, (32, 24)), 1e-3),
                                       ('2d_big_dt', phys84(2, (32, 24)), 0.25),
                                       ('3d', phys84(3, (10, 12, 8)), 1e-3)])
def test_rosw_step_and_trajectory(label, p, h):
    """N-step ROSW trajectory vs the oracle (splu linear solves): rel 1e-8."""
    from ksfd_b200 import core
    from oracle import ksfd_oracle as O
    ph = oracle_physics(p)
    u0 = random_state(p, 31)
    nsteps = 5
    traj = O.integrate(u0, 0.0, h, nsteps, ph)
    ctx = make_ctx(p)
    opts = core.ts_options(adapt='none', ksp_rtol=1e-13, ksp_max_it=2000)
    u = ctx.upload(u0)
    t = 0.0
    for k in range(nsteps):
        ctx.groom(u)
        res = ctx.ts_step(u, t, h, opts)
        assert res.accepted == 1 and res.ksp_fail == 0
        t = res.t_new
        ref = traj[k][1].reshape(-1, order='F')
        assert per_dof_err(ctx.download(u), ref, ph.dof) < 1e-8, (label, k)
    assert abs(t - nsteps * h) < 1e-12
    ctx.close()


def test_rosw_adaptive_matches_oracle():
    from ksfd_b200 import core
    from oracle import ksfd_oracle as O
    p = phys84(2, (24, 20))
    ph = oracle_physics(p)
    u0 = random_state(p, 41)
    adapt = dict(atol=0.01, rtol=1e-6, clip=(0.1, 5.0), dt_min=1e-20, dt_max=1e4)
    # oracle loop
    u = u0.reshape(ph.Vshape, order='F').copy()
    t, h = 0.0, 1e-8
    ref = []
    for k in range(12):
        u = O.groom(u, ph)
        while True:
            un, ue, _ = O.rosw_step(u, t, h, ph)
            en = O.wnorm2(un, ue, adapt['atol'], adapt['rtol'])
            ok, hn = O.adapt_basic(h, en, clip=adapt['clip'], dt_min=adapt['dt_min'],
                                   dt_max=adapt['dt_max'])
            if ok:
                u, t = un, t + h
                h = hn
                break
            h = hn
        ref.append((t, h, u.copy()))
    ctx = make_ctx(p)
    opts = core.ts_options(adapt='basic', atol=0.01, rtol=1e-6, clip=(0.1, 5.0),
                           dt_max=1e4, ksp_rtol=1e-13, ksp_max_it=2000)
    ud = ctx.upload(u0)
    t, h = 0.0, 1e-8
    for k in range(12):
        ctx.groom(ud)
        res = ctx.ts_step(ud, t, h, opts)
        assert res.accepted == 1
        t, h = res.t_new, res.h_next
        assert abs(t - ref[k][0]) <= 1e-6 * abs(ref[k][0]), (k, t, ref[k][0])
        assert abs(h - ref[k][1]) <= 1e-5 * abs(ref[k][1]), (k, h, ref[k][1])
        assert per_dof_err(ctx.download(ud), ref[k][2].reshape(-1, order='F'),
                           ph.dof) < 1e-8
    ctx.close()


# ---------------------------------------------------------------------------
# (2b) BASELINE-size ROSW steps against the oracle's C restatement
# ---------------------------------------------------------------------------
def _c_oracle(ph):
    """the oracle in C + OpenMP (oracle/ksfd_oracle_c.c): what makes a full-size step a
    matter of seconds on the CPU; skipped if the box cannot build / load it"""
    try:
        from oracle import ksfd_oracle_c as OC
        OC.lib()
    except Exception as e:      # noqa: BLE001
        pytest.skip('oracle C library unavailable: %s' % e)
    return OC.COracle(ph)


@pytest.mark.parametrize('label,p,nsteps', [('1024^2', phys84(2, (1024, 1024)), 3),
                                            ('3d_144x80x112', phys84(3, (144, 80, 112)), 2)])
def test_full_size_rosw_steps_vs_c_oracle(label, p, nsteps):
    """The benchmark's own step (options84 physics, dt = 1e-3, clamp + ROSW ra34pw2 with the
    default solver choice, i.e. the fused Richardson sweeps) at the BASELINE 2-D size and on a
    3-D grid with several tiles, chunks and a clamped last tile, against the C oracle
    (pinned to the numpy oracle / SuperLU steps in tests/test_oracle_c.py).  Measured per dof
    relative to the largest INCREMENT of the step (the state itself is 9000 + noise: errors
    relative to it would hide a wrong step).  Tolerance 1e-8; both sides solve to rtol 1e-13."""
    from ksfd_b200 import core
    ph = oracle_physics(p)
    c = _c_oracle(ph)
    u0 = random_state(p, 31)
    ctx = make_ctx(p)
    opts = core.ts_options(adapt='none', ksp_rtol=1e-13, ksp_max_it=2000)
    u = ctx.upload(u0)
    uc = np.ascontiguousarray(u0.copy())
    t = 0.0
    for k in range(nsteps):
        before = uc.copy()
        ctx.groom(u)
        res = ctx.ts_step(u, t, 1e-3, opts)
        assert res.accepted == 1 and res.ksp_fail == 0
        t = res.t_new
        c.ts_step(uc, 1e-3, rtol=1e-13)
        got = ctx.download(u)
        worst = 0.0
        for d in range(ph.dof):
            inc = np.abs(uc[d::ph.dof] - before[d::ph.dof]).max()
            worst = max(worst, np.abs(got[d::ph.dof] - uc[d::ph.dof]).max() / inc)
        print('%s step %d: max error / max increment = %.2e (sweeps %d)' % (label, k, worst, res.ksp_its))
        assert worst < 1e-8, (label, k, worst)
    ctx.close()
    c.close()


# (3) BASELINE sizes: size-independent properties
# ---------------------------------------------------------------------------
@pytest.mark.parametrize('label,p', [('1024^2', phys84(2, (1024, 1024))),
                                     ('256^3', phys84(3, (256, 256, 256)))])
def test_full_size_properties(label, p):
    torch = _torch()
    ctx = make_ctx(p, 2)
    n = ctx.npts * ctx.dof
    gen = torch.Generator(device='cuda').manual_seed(793817931)
    rho = 9000.0 + 90.0 * torch.randn(ctx.npts, generator=gen, device='cuda',
                                      dtype=torch.float64)
    u = rho.repeat_interleave(ctx.dof).contiguous()
    udot = torch.randn(n, generator=gen, device='cuda', dtype=torch.float64)
    v = torch.randn(n, generator=gen, device='cuda', dtype=torch.float64)
    w = torch.randn(n, generator=gen, device='cuda', dtype=torch.float64)
    # marching kernels agree with the independent direct kernels
    F2 = ctx.residual(u, udot)
    ctx.set_option('variant', 1)
    F1 = ctx.residual(u, udot)
    ctx.set_option('variant', 2)
    F1r, F2r = ctx.from_internal(F1), ctx.from_internal(F2)
    for c in range(ctx.dof):
        d = (F2r[c::ctx.dof] - F1r[c::ctx.dof]).abs().max().item()
        assert d / F1r[c::ctx.dof].abs().max().item() < 1e-11, (label, c)
    # F(u, udot) - F(u, 0) == udot  and  dfdt == -F(u, 0) bit exactly
    f = ctx.residual(u)
    F0 = ctx.residual(u, torch.zeros_like(u))
    assert torch.equal(f, -F0)
    # uniform state with s = gamma is an equilibrium
    ueq = torch.full((n,), 9000.0, device='cuda', dtype=torch.float64)
    assert ctx.residual(ueq).abs().max().item() < 1e-8
    # translation equivariance under the periodic wrap (exact index mapping):
    # shifting the field by whole planes of the last axis shifts the residual
    plane = ctx.dof * int(np.prod(ctx.local_shape[:-1]))
    ur = torch.roll(u, 3 * plane)
    assert torch.equal(ctx.residual(ur), torch.roll(f, 3 * plane))
    nx = ctx.local_shape[0]
    ur = torch.roll(u.view(-1, nx), 5, dims=1).reshape(-1)      # every x-line
    fr = torch.roll(f.view(-1, nx), 5, dims=1).reshape(-1)
    assert torch.equal(ctx.residual(ur), fr)
    # J.v: linearity and agreement with the direct kernel and a finite difference
    shift = 1.0 / (0.435866521508459 * 1e-3)
    ctx.jvp_setup(u, shift)
    a, b = 0.37, -1.9
    lhs = ctx.jvp(a * v + b * w)
    rhs = a * ctx.jvp(v) + b * ctx.jvp(w)
    assert (lhs - rhs).abs().max().item() / rhs.abs().max().item() < 1e-13
    Jv2 = ctx.jvp(v)
    ctx.set_option('variant', 1)
    Jv1 = ctx.jvp(v)
    ctx.set_option('variant', 3)
    Jv3 = ctx.jvp(v)
    Jp3 = ctx.jvp(v, precond=True)
    ctx.set_option('variant', 2)
    assert (Jv2 - Jv1).abs().max().item() / Jv1.abs().max().item() < 1e-12
    # the TMA-fed and the register-prefetch marcher share the stage / emit arithmetic
    assert torch.equal(Jv2, Jv3) and torch.equal(ctx.jvp(v, precond=True), Jp3)
    eps = 1e-4
    fd = (ctx.residual(u + eps * v) - ctx.residual(u - eps * v)) / (2 * eps)
    Jv = ctx.from_internal(shift * v - ctx.jvp(v))
    fd = ctx.from_internal(fd)
    for c in range(ctx.dof):
        d = (Jv[c::ctx.dof] - fd[c::ctx.dof]).abs().max().item()
        assert d / fd[c::ctx.dof].abs().max().item() < 1e-6, (label, c)
    # GMRES solves the full-size system to the requested tolerance
    x, res = ctx.gmres(udot, rtol=1e-8, max_it=200)
    assert res.reason > 0
    r = udot - ctx.jvp(x)
    assert ctx.norm2(r) <= 1.5e-8 * ctx.norm2(udot)
    ctx.close()


@pytest.mark.parametrize('label,p', [('1d', phys84(1, (128,), h=1.0 / 128)),
                                     ('2d', phys84(2, (40, 32))),
                                     ('3d', phys84(3, (12, 10, 14)))])
def test_gmres_spectral_preconditioner(label, p):
    """precond=2: FFT inverse of the frozen-coefficient operator.  Same solution
    as the direct solve for small and large time steps, and far fewer Arnoldi
    steps than point-block Jacobi once the step is large (the regime options84
    reaches after its start-up phase)."""
    import scipy.sparse.linalg as spla
    from oracle import ksfd_oracle as O
    ph = oracle_physics(p)
    u = random_state(p, 21)
    rng = np.random.default_rng(22)
    b = rng.standard_normal(u.size)
    ctx = make_ctx(p)
    for dt in (1e-3, 1.0, 100.0):
        shift = 1.0 / (O.ROSW_GAMMA * dt)
        x_ref = spla.splu(O.ijacobian(u, shift, ph).tocsc()).solve(b)
        ctx.jvp_setup(ctx.upload(u), shift)
        x2, r2 = ctx.gmres(ctx.upload(b), rtol=1e-12, max_it=2000, precond=2)
        assert r2.reason > 0, (label, dt, r2.reason, r2.its)
        assert relerr(ctx.download(x2), x_ref) < 1e-8, (label, dt, r2.its)
        x1, r1 = ctx.gmres(ctx.upload(b), rtol=1e-12, max_it=2000, precond=1)
        assert r2.its <= r1.its + 2, (label, dt, r1.its, r2.its)
        if dt >= 1.0:
            assert r2.its <= max(30, r1.its // 4), (label, dt, r1.its, r2.its)
    ctx.close()


def test_gmres_spectral_preconditioner_on_patterned_state():
    """The pattern phase of options84 (fixture from oracle/make_pattern_state.py:
    rho between 5e2 and 2.6e4).  The scaled spectral preconditioner converges to
    the direct solution in a few dozen Arnoldi steps (numpy restatement:
    23 / 43, tests/test_spectral_pc_cpu.py)."""
    import scipy.sparse.linalg as spla
    from oracle import ksfd_oracle as O
    g = np.load(os.path.join(GOLD, 'host_pattern96.npz'))
    u, n = g['u'], tuple(int(x) for x in g['n'])
    p = phys84(2, n)
    ph = oracle_physics(p)
    b = O.dfdt(u, ph).reshape(-1, order='F')
    ctx = make_ctx(p)
    for dt, most in ((5.0, 40), (20.0, 70)):
        shift = 1.0 / (O.ROSW_GAMMA * dt)
        x_ref = spla.splu(O.ijacobian(u, shift, ph).tocsc()).solve(b)
        ctx.jvp_setup(ctx.upload(u), shift)
        x2, r2 = ctx.gmres(ctx.upload(b), rtol=1e-11, max_it=2000, precond=2)
        assert r2.reason > 0 and r2.its <= most * 11 // 8, (dt, r2.reason, r2.its)
        assert relerr(ctx.download(x2), x_ref) < 1e-7, (dt, r2.its)
    ctx.close()
