"""
The time integrator lives in PETSc, which is not in the reference tree.  Pin
the restated ROSW 'ra34pw2' tableau by its defining properties (Rang &
Angermann 2005): order-3 Rosenbrock conditions, stiff accuracy, gamma a root of
6g^3 - 18g^2 + 9g - 1, embedded order 2 — and the oracle stepper by convergence
order and by the manufactured exact solution of options93nx128dt1.  CPU only.
"""
import numpy as np

from oracle import ksfd_oracle as O


def test_order_conditions():
    A, G, b, be = O.RA34PW2_A, O.RA34PW2_GAMMA, O.RA34PW2_B, O.RA34PW2_BEMBED
    g = O.ROSW_GAMMA
    assert abs(6 * g ** 3 - 18 * g ** 2 + 9 * g - 1) < 1e-14
    B = A + G                       # beta_ij = alpha_ij + gamma_ij
    e = np.ones(4)
    alpha = A @ e
    beta = B @ e
    assert abs(b.sum() - 1) < 1e-14                         # order 1
    assert abs(b @ beta - 0.5) < 1e-14                      # order 2
    assert abs(b @ alpha ** 2 - 1 / 3) < 1e-14              # order 3a
    assert abs(b @ (B @ beta) - 1 / 6) < 1e-14              # order 3b
    assert abs(be.sum() - 1) < 1e-14                        # embedded order 1
    assert abs(be @ beta - 0.5) < 1e-14                     # embedded order 2
    assert np.allclose(B[-1], b, atol=1e-15)                # stiffly accurate


def test_transformed_tableau_consistency():
    T = O.rosw_transformed()
    Gi = T['GammaInv']
    assert np.allclose(Gi @ O.RA34PW2_GAMMA, np.eye(4), atol=1e-14)
    assert np.allclose(T['At'] @ O.RA34PW2_GAMMA, O.RA34PW2_A, atol=1e-14)
    assert np.allclose(T['bt'] @ O.RA34PW2_GAMMA, O.RA34PW2_B, atol=1e-14)
    assert abs(Gi[0, 0] - 1 / O.ROSW_GAMMA) < 1e-14


def _lin_problem():
    from helpers import phys84, oracle_physics
    p = phys84(1, (16,), h=1.0 / 16)
    return p, oracle_physics(p)


def test_rosw_third_order_convergence():
    """Global error of the oracle ROSW stepper shrinks like h^3 on the
    (nonlinear) Keller-Segel ODE system of a small 1-D grid."""
    from helpers import random_state
    p, ph = _lin_problem()
    u0 = random_state(p, 1).reshape(ph.Vshape, order='F')
    T = 40.0

    def run(nsteps):
        return O.integrate(u0, 0.0, T / nsteps, nsteps, ph,
                           groom_each_step=False)[-1][1]
    ref = run(256)
    e1 = np.abs(run(8) - ref).max()
    e2 = np.abs(run(16) - ref).max()
    e3 = np.abs(run(32) - ref).max()
    assert 6.0 < e1 / e2 < 11.0, (e1, e2, e3)
    assert 6.0 < e2 / e3 < 11.0, (e1, e2, e3)


def test_adapt_basic_rules():
    ok, hn = O.adapt_basic(1.0, 0.5)
    assert ok and abs(hn - 0.9 * 0.5 ** (-1 / 3)) < 1e-15
    ok, hn = O.adapt_basic(1.0, 8.0)
    assert (not ok) and abs(hn - 0.45 * 8.0 ** (-1 / 3)) < 1e-15
    ok, hn = O.adapt_basic(1.0, 1e-12, clip=(0.1, 5.0))
    assert ok and hn == 5.0
    ok, hn = O.adapt_basic(1.0, 1e9, clip=(0.1, 5.0))
    assert (not ok) and abs(hn - 0.1) < 1e-15
