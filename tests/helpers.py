"""Shared helpers for the tests: golden fixtures, physics builders, inputs."""
import glob
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, 'golden')


def golden_names():
    """Operator fixtures (oracle/make_golden.py); host_*.npz hold host-mirror
    fixtures (oracle/make_golden_host.py) and are read by test_host_mirror."""
    return sorted(n for n in (os.path.splitext(os.path.basename(p))[0]
                              for p in glob.glob(os.path.join(GOLD, '*.npz')))
                  if not n.startswith('host_'))


def load_golden(name):
    g = np.load(os.path.join(GOLD, name + '.npz'))
    physs = json.loads(str(g['phys']))
    return g, physs, int(g['nrec'][0])


def oracle_physics(p):
    from oracle import ksfd_oracle as O
    return O.Physics(p['dim'], p['n'], p['h'], p['groups'], p['s2'], p['rhomax'],
                     p['cushion'], p['maxscale'], p['cap'], p['rhomin'], p['Umin'])


def product_physics(p):
    from ksfd_b200 import core
    groups = [(a, b, [tuple(l) for l in ligs]) for a, b, ligs in p['groups']]
    return core.make_physics(p['dim'], p['h'], groups, p['s2'], p['rhomax'],
                             p['cushion'], p['maxscale'], p['cap'], p['rhomin'],
                             p['Umin'])


OPT84 = dict(s2=0.02357 ** 2 / 2, rhomax=28000.0, cushion=2000.0, maxscale=2.0,
             cap='tophat', rhomin=1e-7, Umin=1e-7,
             groups=[[1500.0, 5.56e-4, [[1.0, 0.01, 0.01, 1e-6]]],
                     [1500.0, -5.56e-4, [[1.0, 0.001, 0.001, 1e-5]]]])


def phys84(dim, n, h=1.0 / 384):
    """options84 physics (reference options84:20-46) on an n grid, spacing h."""
    p = dict(OPT84)
    p.update(dim=dim, n=list(n), h=[h] * dim)
    return p


def random_state(p, seed=0, rel=0.01):
    """rho = 9000 + 90 N(0,1); U_l = rho*(1 + rel*N(0,1)); flat F-order."""
    rng = np.random.default_rng(seed)
    n = tuple(p['n'])
    dof = 1 + sum(len(g[2]) for g in p['groups'])
    a = np.empty((dof,) + n)
    a[0] = 9000.0 + 90.0 * rng.standard_normal(n)
    for l in range(1, dof):
        a[l] = a[0] * (1.0 + rel * rng.standard_normal(n))
    return a.reshape(-1, order='F')


def relerr(a, b, dof=None):
    """max-norm relative error, per dof if dof given."""
    a = np.asarray(a).ravel()
    b = np.asarray(b).ravel()
    if dof is None:
        return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))
    return max(float(np.abs(a[c::dof] - b[c::dof]).max()
                     / max(np.abs(b[c::dof]).max(), 1e-300)) for c in range(dof))


EPS = 2.220446049250313e-16


def cond_scale(ph, u):
    """
    Per-dof conditioning scale of the discrete f(u): the stencil sums cancel
    (f_rho = grad(rho).grad(G) + rho*lap(G) with G nearly constant), so the
    attainable absolute accuracy of ANY fp64 evaluation order is
    ~eps * sum|w_k a_k|, not eps*|f|.  Returns, per dof, sum over the stencil
    of |weight|*|value| bounds:  rho*sum|w2|*|G| + (sum|w1| rho)(sum|w1| |G|)
    for the rho row, gamma*U + s*rho + D*sum|w2|*U for the ligand rows.
    """
    from oracle import ksfd_oracle as O
    ua = np.array(u, dtype=float).reshape(ph.Vshape, order='F')
    farr = O.groom(O.ghost_fill(ua, ph.dim), ph)
    G = np.abs(O.G_of(farr, ph)).max()
    rho = np.abs(farr[0]).max()
    w1 = sum(np.abs(w).sum() for w in ph.w1)
    w2 = sum(np.abs(w).sum() for w in ph.w2)
    out = [rho * w2 * G + sum(np.abs(a).sum() * rho * np.abs(a).sum() * G
                              for a in ph.w1)]
    for l, lig in enumerate(ph.ligands()):
        U = np.abs(farr[l + 1]).max()
        out.append(lig['gamma'] * U + lig['s'] * rho + lig['D'] * w2 * U)
    return np.array(out)


def check_field(a, b, dof, rtol, cond=None, ncond=64.0):
    """
    assert-able error measure: per dof, max|a-b| must be below
    rtol*max|b_dof| + ncond*eps*cond_dof.  Returns the worst ratio
    err/allowed (pass iff < 1).
    """
    a = np.asarray(a).ravel()
    b = np.asarray(b).ravel()
    worst = 0.0
    for c in range(dof):
        err = np.abs(a[c::dof] - b[c::dof]).max()
        allowed = rtol * np.abs(b[c::dof]).max()
        if cond is not None:
            allowed += ncond * EPS * cond[c]
        worst = max(worst, err / max(allowed, 1e-300))
    return worst
