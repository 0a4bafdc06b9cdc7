"""
Host-side mirror (option files, parameters, grid bookkeeping, random IC,
time-series container) against data produced by the reference's own classes
(oracle/make_golden_host.py -> tests/golden/host_*).  CPU only.
"""
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, 'golden')
OPTS = os.path.join(HERE, 'options')

G = json.load(open(os.path.join(GOLD, 'host_params.json')))


@pytest.mark.parametrize('name', sorted(G['params']))
def test_option_file_parses_like_reference(name):
    from ksfd_b200 import SolutionParameters, parse_commandline
    ref = G['params'][name]
    cl = parse_commandline(['@' + os.path.join(OPTS, name + '.args')])
    assert list(cl.petsc) == ref['petsc']
    assert cl.save == ref['save'] and cl.check == ref['check']
    assert list(cl.source) == ref['source'] and cl.seed == ref['seed']
    assert cl.cappotential == ref['cappotential']
    ps = SolutionParameters(cl)
    assert ps.dim == ref['dim'] and ps.nligands == ref['nligands']
    assert (ps.nwidth, ps.nheight, ps.ndepth) == (ref['nwidth'], ref['nheight'], ref['ndepth'])
    assert [l.name() for l in ps.groups.ligands()] == ref['ligand_names']
    assert sorted(ps.tdfuncs) == ref['tdnames']
    for tkey, vals in ref['values'].items():
        mine = ps.values(float(tkey))
        for k, v in vals.items():
            assert k in mine, k
            if isinstance(v, float):
                if mine[k] in (None, False, ''):     # '' defaults (U0_g_l)
                    continue
                assert abs(float(mine[k]) - v) <= 1e-14 * max(1.0, abs(v)), (name, k)


def test_petsc_option_database():
    from ksfd_b200.params import PetscOptions
    po = PetscOptions(['-ts_type', 'rosw', '-ts_adapt_clip', '0.1,5', '-info',
                       '-ts_adapt_dt_min', '1e-20', '-snes_view', 'ascii',
                       '-ksp_max_it', '2000', '-x', '-3'])
    assert po.get('ts_type') == 'rosw'
    assert po.getRealArray('ts_adapt_clip') == [0.1, 5.0]
    assert po.get('info') == '' and po.getReal('ts_adapt_dt_min') == 1e-20
    assert po.getInt('ksp_max_it') == 2000 and po.getReal('x') == -3.0
    with pytest.raises(KeyError):
        po.getReal('missing')


def test_parser_comments_quotes_and_petsc_block(tmp_path):
    from ksfd_b200 import parse_commandline
    f = tmp_path / 'opts'
    f.write_text("# comment\ndim=2 # trailing\n'rho0=murho + 1'\n--petsc\n-ts_type beuler\n--\n"
                 "--seed=7\nmurho=3.0\n")
    cl = parse_commandline(['@' + str(f), 'extra=1.5'])
    assert cl.params == ['dim=2', 'rho0=murho + 1', 'murho=3.0', 'extra=1.5']
    assert cl.petsc == ['-ts_type', 'beuler'] and cl.seed == 7


def test_duplicate_parameters_rejected():
    from ksfd_b200 import KSFDException, SolutionParameters, parse_commandline
    with pytest.raises(KSFDException):
        SolutionParameters(parse_commandline(['dim=1', 'dim=2']))


def test_time_dependent_parameters_and_physics_block():
    from ksfd_b200 import SolutionParameters, parse_commandline
    cl = parse_commandline(['dim=2', 'nelements=16', 's2=2.7e-4*(1+0.01*t)', 'ngroups=1',
                            'nligands_1=1', 'alpha_1=1500', 'beta_1=5.56e-4',
                            's_1_1=0.01+0.001*t', 'gamma_1_1=0.01', 'D_1_1=1e-6'])
    ps = SolutionParameters(cl)
    assert ps.physics_is_time_dependent()
    p0, p3 = ps.physics((0.1, 0.1), 0.0), ps.physics((0.1, 0.1), 3.0)
    assert abs(p0.s2 - 2.7e-4) < 1e-18 and abs(p3.s2 - 2.7e-4 * 1.03) < 1e-18
    assert abs(p3.s[0] - 0.013) < 1e-17 and p3.nlig == 1 and p3.ngroups == 1
    # weights are the reference's (same sympy path): 1/(12 h) * (1,-8,0,8,-1)
    assert abs(p0.w1[0][1] + 8 / (12 * 0.1)) < 1e-12


@pytest.mark.parametrize('key', sorted(G['grids']))
def test_grid_bookkeeping_matches_dmda(key):
    from ksfd_b200.grid import Comm, Grid
    ref = G['grids'][key]
    g = Grid(comm=Comm(0, 1), **ref['kw'])
    assert [float(x) for x in g.spacing] == ref['spacing']
    assert list(g.Slshape) == ref['Slshape'] and list(g.Vlshape) == ref['Vlshape']
    assert list(g.Sashape) == ref['Sashape'] and list(g.Vashape) == ref['Vashape']
    assert g.stencil_width == ref['stencil_width']
    assert [list(r) for r in g.ranges] == ref['ranges']
    assert np.allclose(np.asarray(g.coordsNoGhosts).ravel(order='F'),
                       ref['coordsNoGhosts'], rtol=0, atol=1e-15)


def test_ghost_fill_is_periodic_wrap_bit_exact():
    """DMDA globalToLocal == np.pad(mode='wrap'): integer index compare."""
    from ksfd_b200.grid import Comm, Grid
    g = Grid(dim=2, nx=7, ny=5, dof=3, comm=Comm(0, 1))
    v = g.Vdmda.createGlobalVec()
    v.array = np.arange(v.size, dtype=float)
    l = g.Vdmda.createLocalVec()
    g.Vdmda.globalToLocal(v, l)
    a = np.arange(v.size, dtype=float).reshape(g.Vlshape, order='F')
    want = np.pad(a, [(0, 0), (2, 2), (2, 2)], mode='wrap')
    assert np.array_equal(l.array.reshape(g.Vashape, order='F'), want)
    sl = g.stencil_slice([-2, 1, 0, 1], l.array.reshape(g.Vashape, order='F'))
    assert np.array_equal(sl, np.roll(np.roll(a[1], 2, axis=0), -1, axis=1))


def test_random_function_matches_reference():
    from ksfd_b200.grid import Comm, Grid
    from ksfd_b200.random import random_function
    z = np.load(os.path.join(GOLD, 'host_random.npz'))
    for key in ('r1', 'r2', 'r2n'):
        cfg = json.loads(str(z[key + '_cfg']))
        g = Grid(comm=Comm(0, 1), **cfg['fine'])
        rg = Grid(comm=Comm(0, 1), **cfg['coarse'])
        vals = rg.Sdmda.createGlobalVec()
        vals.array = z[key + '_vals']
        f = random_function(g, randgrid=rg, vals=vals)
        assert np.allclose(f.array_r, z[key + '_field'], rtol=1e-13, atol=1e-14), key


def test_generator_streams_are_the_reference_spawn():
    from ksfd_b200.grid import Comm
    from ksfd_b200.random import Generator
    ss = np.random.SeedSequence(793817931).spawn(4)
    for r in range(4):
        Generator._rng = None
        Generator(seed=793817931, comm=Comm(r, 4))
        a = Generator.get_rng().normal(size=5)
        assert np.array_equal(a, np.random.default_rng(ss[r]).normal(size=5))
    Generator._rng = None


def test_dmda_ownership_ranges():
    from ksfd_b200.core import dmda_ownership
    assert dmda_ownership(10, 3) == [(0, 4), (4, 3), (7, 3)]
    assert dmda_ownership(1024, 8)[3] == (384, 128)
    assert sum(c for _, c in dmda_ownership(257, 8)) == 257


def test_timeseries_roundtrip(tmp_path):
    from ksfd_b200.grid import Comm, Grid
    from ksfd_b200.timeseries import TimeSeries, dillnp, dillunp
    g = Grid(dim=2, nx=6, ny=5, dof=3, comm=Comm(0, 1))
    ts = TimeSeries(str(tmp_path / 'sol' / 'run'), grid=g, mode='w')
    assert os.path.basename(ts.filename).startswith('runs1r0.')
    ts.info['dt'] = 0.25
    ts.info['blob'] = dillnp({'a': 1})
    rng = np.random.default_rng(0)
    snaps = []
    for k, t in enumerate((0.0, 0.5, 1.5)):
        u = g.Vdmda.createGlobalVec()
        u.array = rng.standard_normal(u.size)
        snaps.append(u.array_r.copy())
        ts.store(u, t, k=k)
        ts.temp_close()
        ts.reopen()
    ts.close()
    rd = TimeSeries(str(tmp_path / 'sol' / 'run'), grid=g, mode='r')
    assert list(rd.sorted_times()) == [0.0, 0.5, 1.5]
    assert float(rd.info['dt']) == 0.25 and dillunp(np.asarray(rd.info['blob'])) == {'a': 1}
    got = rd.retrieve_by_time(1.5)
    assert got.shape == g.Vlshape        # (dof, nx, ny), C order on disk
    assert np.array_equal(got, snaps[2].reshape(g.Vlshape, order='F'))
    mid = rd.retrieve_by_time(1.0)
    want = 0.5 * (snaps[1] + snaps[2]).reshape(g.Vlshape, order='F')
    assert np.allclose(mid, want)


def test_ksfd_import_name_resolves_to_this_package():
    """existing user scripts do `from KSFD import ...` (reference ksfdsolver2.py:354-360)"""
    import KSFD
    from KSFD import (KSFDException, LigandGroups, ParameterList, Parser,  # noqa: F401
                      SolutionParameters, default_parameters)
    from KSFD.ksfddebug import log
    import ksfd_b200
    assert KSFD.Parser is ksfd_b200.Parser
    assert KSFD.implicitTS is ksfd_b200.implicitTS and KSFD.Grid is ksfd_b200.Grid
    assert KSFD.TimeSeries is ksfd_b200.TimeSeries and KSFD.dillnp is ksfd_b200.dillnp
    log('quiet unless KSFDDEBUG names the system', system='TEST')
    pl = ParameterList([('a', 1, 'first'), ('b', 2.5)])
    pl['c'] = 7
    pl.update({'a': 3})
    assert dict(pl.items()) == {'a': 3, 'b': 2.5, 'c': 7} and pl.get('zz', 9) == 9
    pl.decode(['b=4.5'])
    assert pl['b'] == 4.5 and pl.defaults['b'] == 2.5 and pl.helps['a'] == 'first'
    with pytest.raises(KSFDException):
        pl.decode(['a=1', 'a=2'])


def test_ksfd_alias_covers_the_reference_all_or_says_why():
    """Every name of the reference's KSFD.__all__ either resolves through the alias package or
    raises an AttributeError that names what replaces it (code generation, assembled matrix)."""
    import KSFD
    ref_all = ['getMat', 'Parser', 'KSFDException', 'Generator', 'random_function', 'TimeSeries',
               'dillnp', 'dillunp', 'remap_from_files', 'makeKSFDSolver', 'Parameter',
               'ParameterList', 'Ligand', 'LigandGroup', 'LigandGroups', 'find_duplicates',
               'SolutionParameters', 'Solution', 'default_parameters', 'UFUNC_MAXARGS',
               'UfuncifyCodeWrapperMultiple', 'ufuncify', 'Grid', 'safe_sympify',
               'cartesian_product', 'spatial_expression', 'SpatialExpression', 'StencilUfunc',
               'Derivatives', 'ksfdTS', 'implicitTS']       # /root/reference/KSFD/__init__.py
    replaced = []
    for name in ref_all:
        try:
            getattr(KSFD, name)
        except AttributeError as e:
            assert 'not provided by ksfd_b200' in str(e), (name, str(e))
            replaced.append(name)
    assert sorted(replaced) == sorted(KSFD._REPLACED), replaced
    for name in ('Grid', 'Derivatives', 'implicitTS', 'ksfdTS', 'TimeSeries', 'SpatialExpression',
                 'Generator', 'random_function'):
        assert name not in replaced
