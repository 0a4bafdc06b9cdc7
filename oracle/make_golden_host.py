#!/usr/bin/env python3
"""
Golden data for the HOST-side mirror (parameters, grid bookkeeping, random
initial condition), produced by the reference's own classes through
oracle/refharness.  Run once in the build container:

    python oracle/make_golden_host.py

Writes tests/golden/host_params.json and tests/golden/host_random.npz.
"""
import contextlib
import io
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle.refharness import harness  # noqa: E402

GOLD = os.path.join(ROOT, 'tests', 'golden')
OPTION_FILES = ['options80', 'options81', 'options84', 'options92',
                'options93nx128dt1', 'options113a']


def numeric(d):
    out = {}
    for k, v in d.items():
        try:
            out[k] = float(v)
        except (TypeError, ValueError):
            out[k] = str(v)
    return out


def main():
    C = harness.load_reference()
    KSFD = C['KSFD']
    solver = harness.load_solver_module()
    params = {}
    for name in OPTION_FILES:
        path = os.path.join(harness.REFERENCE_ROOT, name)
        clargs = solver.parse_commandline(['@' + path])
        ps = KSFD.SolutionParameters(clargs)
        rec = dict(petsc=list(clargs.petsc), save=clargs.save, check=clargs.check,
                   source=list(clargs.source), seed=clargs.seed,
                   cappotential=clargs.cappotential,
                   dim=int(ps.dim), nligands=int(ps.nligands),
                   nwidth=int(ps.nwidth), nheight=int(ps.nheight), ndepth=int(ps.ndepth),
                   ligand_names=[l.name() for l in ps.groups.ligands()],
                   tdnames=sorted(ps.tdfuncs.keys()),
                   values={str(t): numeric(ps.values(t)) for t in (0.0, 7.5)})
        params[name] = rec
    # grid bookkeeping for a few shapes
    grids = {}
    for key, kw in {'g1': dict(dim=1, width=2.0, nx=10, dof=3),
                    'g2': dict(dim=2, width=1.5, height=0.5, nx=6, ny=9, dof=2),
                    'g3': dict(dim=3, width=1.0, height=2.0, depth=3.0, nx=5, ny=6,
                               nz=7, dof=4)}.items():
        g = C['Grid'](**kw)
        grids[key] = dict(kw=kw, spacing=[float(x) for x in g.spacing],
                          Slshape=list(g.Slshape), Vlshape=list(g.Vlshape),
                          Sashape=list(g.Sashape), Vashape=list(g.Vashape),
                          stencil_width=int(g.stencil_width),
                          ranges=[list(r) for r in g.ranges],
                          coordsNoGhosts=np.asarray(g.coordsNoGhosts).ravel(order='F').tolist())
    json.dump(dict(params=params, grids=grids),
              open(os.path.join(GOLD, 'host_params.json'), 'w'), indent=1, sort_keys=True,
              default=lambda o: o.item() if hasattr(o, 'item') else str(o))
    # random_function with supplied coarse values (non-square, to pin the
    # reference's C-order/F-order index scramble)
    out = {}
    rng = np.random.default_rng(5)
    for key, (fine, coarse) in {'r1': (dict(dim=1, width=1.0, nx=32, dof=1),
                                       dict(dim=1, width=1.0, nx=8, dof=1)),
                                'r2': (dict(dim=2, width=1.0, height=1.0, nx=24, ny=24, dof=1),
                                       dict(dim=2, width=1.0, height=1.0, nx=6, ny=6, dof=1)),
                                'r2n': (dict(dim=2, width=2.0, height=1.0, nx=24, ny=16, dof=1),
                                        dict(dim=2, width=2.0, height=1.0, nx=6, ny=4, dof=1)),
                                }.items():
        g = C['Grid'](**fine)
        rg = C['Grid'](**coarse)
        vals = rg.Sdmda.createGlobalVec()
        vals.array = rng.standard_normal(vals.array.shape)
        with contextlib.redirect_stdout(io.StringIO()):
            f = KSFD.random_function(g, randgrid=rg, vals=vals)
        out[key + '_vals'] = vals.array.copy()
        out[key + '_field'] = f.array.copy()
        out[key + '_cfg'] = np.array(json.dumps(dict(fine=fine, coarse=coarse)))
    np.savez_compressed(os.path.join(GOLD, 'host_random.npz'), **out)
    print('host goldens written')


if __name__ == '__main__':
    main()
