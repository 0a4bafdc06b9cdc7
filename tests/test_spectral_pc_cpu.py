"""
The spectral preconditioner's algorithm (ksfd_b200/csrc/fftpc.cuh) restated in
numpy and pinned against the oracle's assembled Jacobian (the reference's
implicitIJ, KSFD/ksfdts.py:598-640, as restated in oracle/ksfd_oracle.py):

* on a uniform state the frozen-coefficient operator A0 IS the Jacobian, so
  A0^-1 A = I to rounding — this pins the symbol of the 4th-order stencil, the
  sign conventions of the arrow block and the Schur step;
* on a random state right-preconditioned GMRES with A0^-1 converges in a handful
  of steps for dt from 1e-3 to 100 (hardware independent: the CUDA path must
  show the same counts, tests/test_gpu_parity.py).
CPU only; no product code is exercised here.
"""
import numpy as np
import pytest
import scipy.sparse.linalg as spla

from helpers import oracle_physics, phys84, random_state


def spectral_inverse(u, shift, ph):
    """numpy restatement of k_fft_means_* + k_fft_prescale + k_fft_symbol_solve +
    k_fft_postscale (+ the FFTs):  M^-1 = S_R A0^-1 S_L."""
    from oracle import ksfd_oracle as O
    ua = np.asarray(u, dtype=float).reshape(ph.Vshape, order='F').copy()
    O.groom(ua, ph)
    g_rho, g_U = O.dG_of(ua, ph)
    rho = ua[0]
    kbar, cbar, gl = (rho * g_rho).mean(), (1.0 / g_rho).mean(), [g.mean() for g in g_U]
    n = ph.Vshape[1:]
    axes = tuple(range(1, ph.dim + 1))
    lam = 0.0
    for ax in range(ph.dim):
        c = np.cos(2 * np.pi * np.arange(n[ax]) / n[ax])
        sh = [1] * ph.dim
        sh[ax] = n[ax]
        lam = lam + ((32 * c - 2 * (2 * c * c - 1) - 30) / (12 * ph.h[ax] ** 2)).reshape(sh)
    ligs = ph.ligands()

    def apply(r):
        ra = np.asarray(r, dtype=float).reshape(ph.Vshape, order='F').copy()
        ra[0] /= rho                                    # S_L: rho row divided by rho
        R = np.fft.fftn(ra, axes=axes)
        schur = shift / kbar - lam
        t = R[0].copy()
        invd = []
        for l, lig in enumerate(ligs):
            invd.append(1.0 / (shift + lig['gamma'] - lig['D'] * lam))
            bd = -gl[l] * lam * invd[l]
            schur = schur + bd * lig['s'] * cbar
            t = t - bd * R[1 + l]
        Z = np.empty_like(R)
        Z[0] = t / schur
        for l, lig in enumerate(ligs):
            Z[1 + l] = (R[1 + l] + lig['s'] * cbar * Z[0]) * invd[l]
        z = np.fft.ifftn(Z, axes=axes).real
        z[0] /= g_rho                                   # S_R: v_rho = y_0 / g_rho
        return z.reshape(-1, order='F')
    return apply


def patterned_state():
    """options84 at t = 2.7e3 on a 96^2 tile (oracle/make_pattern_state.py): rho
    between 5e2 and the density cap 2.6e4 — the pattern phase."""
    import os
    from helpers import GOLD
    g = np.load(os.path.join(GOLD, 'host_pattern96.npz'))
    return g['u'], tuple(int(x) for x in g['n'])


CASES = [('1d', phys84(1, (128,), h=1.0 / 128)), ('2d', phys84(2, (40, 32))),
         ('3d', phys84(3, (12, 10, 14)))]


@pytest.mark.parametrize('label,p', CASES)
def test_exact_on_uniform_state(label, p):
    from oracle import ksfd_oracle as O
    ph = oracle_physics(p)
    u = np.full(ph.Vshape, 9000.0).reshape(-1, order='F')
    v = np.random.default_rng(3).standard_normal(u.size)
    for dt in (1e-3, 1.0, 100.0):
        shift = 1.0 / (O.ROSW_GAMMA * dt)
        z = spectral_inverse(u, shift, ph)(O.ijacobian(u, shift, ph) @ v)
        assert np.linalg.norm(z - v) <= 1e-12 * np.linalg.norm(v), (label, dt)


@pytest.mark.parametrize('label,p', CASES)
def test_gmres_step_counts_on_random_state(label, p):
    from oracle import ksfd_oracle as O
    ph = oracle_physics(p)
    u = random_state(p, 21)
    b = np.random.default_rng(22).standard_normal(u.size)
    for dt, most in ((1e-3, 6), (1.0, 15), (100.0, 25)):
        shift = 1.0 / (O.ROSW_GAMMA * dt)
        A = O.ijacobian(u, shift, ph).tocsr()
        M = spectral_inverse(u, shift, ph)
        count = [0]

        def cb(_):
            count[0] += 1
        op = spla.LinearOperator(A.shape, matvec=lambda x: A @ M(x))
        y, info = spla.gmres(op, b, rtol=1e-10, restart=30, maxiter=10, callback=cb,
                             callback_type='pr_norm')
        assert info == 0 and count[0] <= most, (label, dt, count[0])
        x = M(y)
        assert np.linalg.norm(b - A @ x) <= 1e-9 * np.linalg.norm(b)


def test_scalings_keep_step_counts_low_on_a_patterned_state():
    """rho from 5e2 to 2.6e4: the diagonal scalings (row / rho, unknown
    g_rho*v_rho) halve the step count of the unscaled frozen-coefficient inverse
    (51 / 60 steps at dt = 5 / 20) and stay far below point-block Jacobi (which
    does not converge in 600 steps here)."""
    from oracle import ksfd_oracle as O
    u, n = patterned_state()
    p = phys84(2, n)
    ph = oracle_physics(p)
    b = O.dfdt(u, ph).reshape(-1, order='F')
    for dt, most in ((5.0, 30), (20.0, 50)):
        shift = 1.0 / (O.ROSW_GAMMA * dt)
        A = O.ijacobian(u, shift, ph).tocsr()
        M = spectral_inverse(u, shift, ph)
        count = [0]

        def cb(_):
            count[0] += 1
        op = spla.LinearOperator(A.shape, matvec=lambda x: A @ M(x))
        y, info = spla.gmres(op, b, rtol=1e-8, restart=30, maxiter=4, callback=cb,
                             callback_type='pr_norm')
        assert info == 0 and count[0] <= most, (dt, count[0])
