#!/usr/bin/env python3
"""
Generate golden vectors from the reference's OWN code.

Runs the unmodified /root/reference KSFD.Derivatives (sympy -> C99 -> numpy
ufunc path, compiled here with gcc) through oracle/refharness and stores small
input/output fixtures under tests/golden/*.npz.  Run ONCE in the build
container (needs /root/reference + gcc; takes ~10 min because the reference
re-generates and compiles its ufuncs for every distinct problem):

    python oracle/make_golden.py [case ...]

The fixtures (not this script) travel to the GPU box.  Every fixture stores
  phys   JSON description of the problem at the evaluation time(s)
  u      input field vectors, flat F-order (dof fastest)
  f      reference Derivatives.dfdt(u, t)            (ksfdsym.py:902-940)
  vel    reference Derivatives.velocity(u, t)        (ksfdsym.py:1188-1209)
  J_*    reference Derivatives.Jacobian(u, t) as COO (ksfdsym.py:814-886) for
         1-D/2-D; for 3-D the per-stencil values of rhoJacobian_arrays
         applied with correct x-fastest row indexing (the reference's own 3-D
         row order is defective: SURVEY.md 8a) as J@v products.
"""
import contextlib
import io
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle.refharness import harness  # noqa: E402

GOLD = os.path.join(ROOT, 'tests', 'golden')

OPT84_PHYS = [
    'sigma=0.02357', 's2=sigma**2/2', 'rhomin=1e-7', 'rhomax=28000',
    'cushion=2000', 'Nworms=0', 'murho=9000.0', 'rho0=murho', 'srho0=90',
    'ngroups=2',
    'nligands_1=1', 'alpha_1=1500', 'beta_1=5.56e-4', 's_1_1=0.01',
    'gamma_1_1=0.01', 'D_1_1=1e-6',
    'nligands_2=1', 'alpha_2=1500', 'beta_2=-5.56e-4', 's_2_1=0.001',
    'gamma_2_1=0.001', 'D_2_1=1e-5',
]


def opt93_args():
    """The shipped options93nx128dt1 file, read as the reference reads it."""
    return ['@' + os.path.join(harness.REFERENCE_ROOT, 'options93nx128dt1')]


CASES = {
    # name: (args, times, kind)
    'opt93_1d128': (opt93_args, [0.0, 5.0, 100.0], 'opt93'),
    'phys84_1d40': (lambda: ['dim=1', 'nelements=40', 'width=0.104166666666666667']
                    + OPT84_PHYS, [0.0], 'random'),
    'phys84_2d12': (lambda: ['dim=2', 'nelements=12', 'width=0.03125',
                             'height=0.03125'] + OPT84_PHYS, [0.0], 'random'),
    'phys84_2d48x40': (lambda: ['dim=2', 'nwidth=48', 'nheight=40',
                                'width=0.125', 'height=0.10416666666666667']
                       + OPT84_PHYS, [0.0], 'random'),
    'phys84_3d8x6x10': (lambda: ['dim=3', 'nwidth=8', 'nheight=6', 'ndepth=10',
                                 'width=0.020833333333333332', 'height=0.015625',
                                 'depth=0.026041666666666668'] + OPT84_PHYS,
                        [0.0], 'random'),
    'witch_2d10x14': (lambda: ['--cappotential=witch', 'dim=2', 'nwidth=10',
                               'nheight=14', 'width=0.5', 'height=0.7',
                               'rhomax=9500', 'cushion=300']
                      + [a for a in OPT84_PHYS
                         if not a.startswith(('rhomax', 'cushion'))],
                      [0.0], 'random'),
    'onelig_2d16': (lambda: ['dim=2', 'nelements=16', 'width=1.0', 'height=1.0',
                             's2=5.56e-4', 'ngroups=1', 'nligands_1=1',
                             'alpha_1=1500', 'beta_1=5.56e-4', 's_1_1=0.01',
                             'gamma_1_1=0.02', 'D_1_1=1e-4', 'Nworms=0',
                             'rho0=9000.0'], [0.0], 'random'),
    'threelig_2d12x9': (lambda: ['dim=2', 'nwidth=12', 'nheight=9', 'width=0.3',
                                 'height=0.225', 's2=2.5e-4', 'ngroups=2',
                                 'nligands_1=2', 'alpha_1=1200', 'beta_1=4e-4',
                                 'weight_1_1=0.7', 's_1_1=0.01',
                                 'gamma_1_1=0.01', 'D_1_1=2e-6',
                                 'weight_1_2=1.3', 's_1_2=0.02',
                                 'gamma_1_2=0.04', 'D_1_2=3e-6',
                                 'nligands_2=1', 'alpha_2=1800',
                                 'beta_2=-3e-4', 's_2_1=0.001',
                                 'gamma_2_1=0.002', 'D_2_1=1e-5', 'Nworms=0',
                                 'rho0=9000.0'], [0.0], 'random'),
    'tdparam_2d9x11': (lambda: ['dim=2', 'nwidth=9', 'nheight=11', 'width=0.09',
                                'height=0.11', 's2=2.7e-4*(1+0.01*t)',
                                'ngroups=1', 'nligands_1=1', 'alpha_1=1500',
                                'beta_1=5.56e-4', 's_1_1=0.01+0.001*t',
                                'gamma_1_1=0.01', 'D_1_1=1e-6', 'Nworms=0',
                                'rho0=9000.0'], [0.0, 3.0], 'random'),
}


def phys_json(ps, grid, t):
    """Plain-number problem description at time t, from the reference ps."""
    v = ps.values(t)
    groups = []
    for g in ps.Vgroups.groups:
        gn = g.groupnum
        ligs = []
        for lig in g.ligands:
            ln = lig.ligandnum
            ligs.append([float(v['weight_%d_%d' % (gn, ln)]),
                         float(v['s_%d_%d' % (gn, ln)]),
                         float(v['gamma_%d_%d' % (gn, ln)]),
                         float(v['D_%d_%d' % (gn, ln)])])
        groups.append([float(v['alpha_%d' % gn]), float(v['beta_%d' % gn]),
                       ligs])
    return dict(dim=int(grid.dim), n=[int(x) for x in grid.nps],
                h=[float(x) for x in grid.spacing], groups=groups,
                s2=float(v['s2']), rhomax=float(v['rhomax']),
                cushion=float(v['cushion']), maxscale=float(v['maxscale']),
                cap=ps.clargs.cappotential, rhomin=float(v['rhomin']),
                Umin=float(v['Umin']), t=float(t))


def make_inputs(kind, ps, grid, rng):
    """A few field vectors per case: plain random, one needing the clamp."""
    shape = grid.Vlshape
    us = []
    if kind == 'opt93':
        x = grid.coordsNoGhosts[0]
        v0 = ps.values0
        for amp in (1.0, 37.0):
            a = np.empty(shape)
            sn = np.sin(2 * np.pi * (0.25 + 4.0 * x))
            a[0] = 9000.0 + amp * sn
            a[1] = 9000.0 + amp * 0.6846227279629311 * sn
            a[2] = 9000.0 + amp * 0.088562372925828 * sn
            us.append(a)
    a = np.empty(shape)
    a[0] = 9000.0 + 90.0 * rng.standard_normal(shape[1:])
    for l in range(1, shape[0]):
        a[l] = a[0] * (1.0 + 0.02 * rng.standard_normal(shape[1:]))
    us.append(a)
    b = a.copy() * (1.0 + 0.3 * rng.standard_normal(shape))
    idx = rng.integers(0, b.size, size=max(3, b.size // 50))
    bf = b.reshape(-1)
    bf[idx[0::3]] = -5.0            # below the clamp
    bf[idx[1::3]] = np.nan          # NaN -> min
    bf[idx[2::3]] = 0.0
    us.append(b)
    return us


def run_case(name):
    argsf, times, kind = CASES[name]
    t0 = time.time()
    clargs, ps, grid, derivs = harness.make_problem(argsf())
    rng = np.random.default_rng(abs(hash(name)) % (2 ** 31) if False else
                                sum(map(ord, name)))
    us = make_inputs(kind, ps, grid, rng)
    out = {}
    physs = []
    dof = grid.dof
    n = tuple(int(x) for x in grid.nps)
    npts = int(np.prod(n))
    rec = 0
    for t in times:
        for a in us:
            u = grid.Vdmda.createGlobalVec()
            u.array = a
            uin = u.array.copy()
            f = derivs.dfdt(u, t=t)
            vel = derivs.velocity(u, t=t)
            out['u_%d' % rec] = uin
            out['f_%d' % rec] = f.array.copy()
            out['vel_%d' % rec] = np.array(vel).reshape(-1, order='F')
            # sources alone, for the product's source-term check
            src = np.stack([np.array(derivs.sources[c](t), dtype=float)
                            * np.ones(n) for c in range(dof)])
            out['src_%d' % rec] = src.reshape(-1, order='F')
            vs = rng.standard_normal((3, dof * npts))
            out['v_%d' % rec] = vs
            if grid.dim < 3:
                kJ = derivs.Jacobian(u, t=t)
                J = kJ.tocsr()
                J.sum_duplicates()
                coo = J.tocoo()
                if coo.nnz * 16 < 400000:
                    out['Jrow_%d' % rec] = coo.row.astype(np.int32)
                    out['Jcol_%d' % rec] = coo.col.astype(np.int32)
                    out['Jval_%d' % rec] = coo.data
                out['Jnnz_%d' % rec] = np.array([J.nnz])
                out['Jv_%d' % rec] = np.stack([J @ v for v in vs])
            else:
                # correct-indexing application of the reference's own
                # per-stencil Jacobian values (rho rows) + constant U rows
                rows, cos, values = derivs.rhoJacobian_arrays(u, t=t)
                vals = values.reshape(n + (len(cos),), order='F')
                Jv = np.zeros((3, dof * npts))
                pid = np.arange(npts).reshape(n, order='F')
                for k, (di, dj, dk, c) in enumerate(cos):
                    cp = np.roll(pid, (-di, -dj, -dk), axis=(0, 1, 2))
                    for m in range(3):
                        Jv[m][(0 + dof * pid).ravel(order='F')] += (
                            vals[..., k] * vs[m][c + dof * cp]
                        ).ravel(order='F')
                for lm1, (urows, ucos, uvals) in enumerate(
                        derivs.UJacobian_arrays(t)):
                    row_dof = lm1 + 1
                    for k, (di, dj, dk, c) in enumerate(ucos):
                        cp = np.roll(pid, (-di, -dj, -dk), axis=(0, 1, 2))
                        for m in range(3):
                            Jv[m][(row_dof + dof * pid).ravel(order='F')] += (
                                uvals[0, k] * vs[m][c + dof * cp]
                            ).ravel(order='F')
                out['Jv_%d' % rec] = Jv
            physs.append(phys_json(ps, grid, t))
            rec += 1
    out['phys'] = np.array(json.dumps(physs))
    out['nrec'] = np.array([rec])
    os.makedirs(GOLD, exist_ok=True)
    np.savez_compressed(os.path.join(GOLD, name + '.npz'), **out)
    return rec, time.time() - t0


def main():
    names = sys.argv[1:] or list(CASES)
    for name in names:
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            rec, dt = run_case(name)
        print('%-20s %d records  %.0f s' % (name, rec, dt), flush=True)


if __name__ == '__main__':
    main()
