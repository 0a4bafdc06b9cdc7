"""
Pin the oracle: oracle/ksfd_oracle.py (numpy restatement) against golden
vectors produced by the reference's own code (oracle/make_golden.py ran the
unmodified KSFD.Derivatives.dfdt / Jacobian / velocity in the build
container).  CPU only.
"""
import numpy as np
import pytest
import scipy.sparse as sp

from helpers import (check_field, cond_scale, golden_names, load_golden,
                     oracle_physics, relerr)
from oracle import ksfd_oracle as O

NAMES = golden_names()

# fp64 tolerances (max-norm relative to the largest entry of that dof).  The
# reference evaluates sympy-folded expressions in a different association
# order, so agreement is to rounding, not bitwise.
TOL_F = 5e-12
TOL_J = 1e-11


def test_fixtures_present():
    assert len(NAMES) >= 8, NAMES


@pytest.mark.parametrize('name', NAMES)
def test_dfdt_velocity(name):
    g, physs, nrec = load_golden(name)
    for r in range(nrec):
        ph = oracle_physics(physs[r])
        src = g['src_%d' % r].reshape((ph.dof,) + ph.n, order='F')
        f = O.dfdt(g['u_%d' % r], ph, sources=list(src)).reshape(-1, order='F')
        fr = g['f_%d' % r]
        # relative to max|f| per dof, plus the cancellation floor of the stencil
        # (smooth states: |f| << sum|w_k a_k|, see helpers.cond_scale)
        bad = check_field(f, fr, ph.dof, TOL_F, cond_scale(ph, g['u_%d' % r]))
        assert bad < 1.0, (name, r, bad)
        v = O.velocity(g['u_%d' % r], ph).reshape(-1, order='F')
        vr = g['vel_%d' % r]
        assert relerr(v, vr) < TOL_F or np.abs(v - vr).max() < 1e-13, (name, r)


@pytest.mark.parametrize('name', NAMES)
def test_jacobian(name):
    g, physs, nrec = load_golden(name)
    for r in range(nrec):
        ph = oracle_physics(physs[r])
        J = O.jacobian(g['u_%d' % r], ph)
        if 'Jnnz_%d' % r in g:
            # sparsity: 47 nnz/point in 2-D, 19 (5+... ) in 1-D etc.
            assert J.nnz == int(g['Jnnz_%d' % r][0])
        if 'Jval_%d' % r in g:
            Jr = sp.coo_matrix((g['Jval_%d' % r],
                                (g['Jrow_%d' % r], g['Jcol_%d' % r])),
                               shape=J.shape).tocsr()
            for c in range(ph.dof):
                rows = np.arange(c, J.shape[0], ph.dof)
                d = abs(J[rows] - Jr[rows]).max()
                assert d / abs(Jr[rows]).max() < TOL_J, (name, r, c)
        for m in range(3):
            v = g['v_%d' % r][m]
            ref = g['Jv_%d' % r][m]
            assert relerr(J @ v, ref, ph.dof) < TOL_J, (name, r, m)
            mf = -O.jvp(g['u_%d' % r], v, 0.0, ph).reshape(-1, order='F')
            assert relerr(mf, ref, ph.dof) < TOL_J, (name, r, m, 'matrix-free')


def test_uniform_equilibrium():
    """rho = U = const with s = gamma is an equilibrium: f == 0
    (SURVEY 8c known answer ii)."""
    from helpers import phys84
    ph = oracle_physics(phys84(2, (16, 12)))
    u = np.full(ph.dof * ph.npts, 9000.0)
    f = O.dfdt(u, ph)
    assert np.abs(f).max() < 1e-8      # ~ rho*|G|*sum|w2| * eps (weights do not cancel bitwise)


def test_jacobian_is_derivative_3d():
    """Central difference of dfdt vs assembled Jacobian (3-D, where the
    reference's own assembled matrix is defective)."""
    from helpers import phys84, random_state
    p = phys84(3, (6, 7, 8))
    ph = oracle_physics(p)
    u = random_state(p, 3)
    J = O.jacobian(u, ph)
    rng = np.random.default_rng(0)
    v = rng.standard_normal(u.size)
    eps = 1e-3
    fd = (O.dfdt(u + eps * v, ph) - O.dfdt(u - eps * v, ph)).reshape(-1, order='F') / (2 * eps)
    assert relerr(J @ v, fd, ph.dof) < 1e-6


def test_ghosted_subbox_equals_global():
    """dfdt_ghosted / jvp_ghosted on a wrapped sub-box == the same region of the global
    evaluation, bit for bit (the full-size GPU tests rely on it)."""
    from helpers import oracle_physics, phys84, random_state
    from oracle import ksfd_oracle as O
    for n, lo, size in (((20, 16), (14, 0), (12, 16)), ((10, 12, 9), (7, 10, 0), (6, 5, 9))):
        p = phys84(len(n), n)
        ph = oracle_physics(p)
        u = random_state(p, 3).reshape(ph.Vshape, order='F')
        v = np.random.default_rng(4).standard_normal(ph.Vshape)
        f = O.dfdt(u, ph)
        jv = O.jvp(u, v, 123.0, ph)
        phb = oracle_physics(phys84(len(n), size))
        fb = O.dfdt_ghosted(O.cut_box(u, lo, size), phb)
        jb = O.jvp_ghosted(O.cut_box(u, lo, size), O.cut_box(v, lo, size), 123.0, phb)
        assert np.array_equal(fb, O.cut_box(f, lo, size, sw=0))
        assert np.array_equal(jb, O.cut_box(jv, lo, size, sw=0))
