"""
Gatherer / tsmerge (reference KSFD/ksfdtimeseries.py:674-828, tsmerge.py:40-111): the per-rank
files a multi-rank run writes are gathered into the global series, interrupted-and-resumed runs
are merged, and the result resumes on any number of ranks.  CPU only; runs on whichever backend
the box has (h5py or the .npz stand-in).
"""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _grids(size, dim=2, n=(6, 11), dof=3):
    from ksfd_b200.grid import Comm, Grid
    kw = dict(dim=dim, dof=dof, width=1.0, height=2.0, depth=1.5, nx=n[0],
              ny=n[1] if dim > 1 else 8, nz=n[2] if dim > 2 else 8)
    return [Grid(comm=Comm(r, size), **kw) for r in range(size)], Grid(comm=Comm(0, 1), **kw)


def _write(prefix, size, times, k0=0, dim=2, n=(6, 11), seed=0):
    """every 'rank' writes its slab of a known global field per time point"""
    from ksfd_b200.timeseries import TimeSeries, dillnp
    grids, gg = _grids(size, dim, n)
    rng = np.random.default_rng(seed)
    fields = [rng.standard_normal(gg.globalVshape) for _ in times]
    for r, g in enumerate(grids):
        ts = TimeSeries(prefix, grid=g, comm=g.comm, mode='w')
        ts.info['dt'] = 0.5
        ts.info['commandlineArguments'] = dillnp(['a=1', 'b=2'])
        lo, hi = g.ranges[-1]
        for i, (t, f) in enumerate(zip(times, fields)):
            ts.store(np.asfortranarray(f[..., lo:hi]).reshape(-1, order='F'), t, k=k0 + i)
        ts.close()
    return fields, gg


@pytest.mark.parametrize('dim,n', [(1, (17,)), (2, (6, 11)), (3, (4, 5, 9))])
def test_gatherer_reassembles_the_global_field(tmp_path, dim, n):
    from ksfd_b200.timeseries import Gatherer
    prefix = str(tmp_path / 'run')
    times = [0.0, 0.25, 1.5]
    fields, gg = _write(prefix, 3, times, dim=dim, n=n)
    g = Gatherer(prefix + 's3@')                    # the reference's special name
    assert g.size == 3 and tuple(g.grid.globalVshape) == tuple(gg.globalVshape)
    assert tuple(g.grid.Vlshape) == tuple(gg.globalVshape)         # one rank owns everything
    assert np.array_equal(g.sorted_times(), times)
    for j, k in enumerate(g.sorted_steps()):
        glob = np.full(gg.globalVshape, np.nan)
        ranks = []
        for s in g:
            assert s.shape == s.retrieve_by_number(k).shape
            glob[s.slice] = s.retrieve_by_number(k)
            ranks.append(s.rank)
        assert ranks == [0, 1, 2]
        assert np.array_equal(glob, fields[j])      # bit-exact, every point written once
    # a second pass over the same gatherer restarts at rank 0
    assert [s.rank for s in g] == [0, 1, 2]
    g.close()
    assert Gatherer(prefix, size=3).size == 3
    with pytest.raises(ValueError):
        Gatherer(prefix, size=0)
    with pytest.raises(FileNotFoundError):
        Gatherer(prefix + 's4@')


def test_tsmerge_gathers_merges_and_selects(tmp_path):
    from ksfd_b200.timeseries import Gatherer, TimeSeries, dillunp, tsmerge
    a = str(tmp_path / 'first')
    b = str(tmp_path / 'second')
    fa, gg = _write(a, 4, [0.0, 1.0, 2.0], seed=1, n=(6, 13))
    fb, _ = _write(b, 1, [3.0, 4.0], k0=7, seed=2, n=(6, 13))     # resumed sequentially
    out = str(tmp_path / 'sub' / 'merged')
    name = tsmerge(out, [a + 's4@', b])
    assert os.path.isfile(name) and os.path.basename(name).startswith('mergeds1r0')
    m = TimeSeries(out, mode='r')                   # grid comes from the file
    assert tuple(m.grid.globalVshape) == tuple(gg.globalVshape)
    assert np.array_equal(m.sorted_times(), [0.0, 1.0, 2.0, 3.0, 4.0])
    for t, f in zip([0.0, 1.0, 2.0, 3.0, 4.0], fa + fb):
        assert np.array_equal(m.retrieve_by_time(t), f)
    assert np.allclose(m.retrieve_by_time(2.5), 0.5 * (fa[2] + fb[0]))
    assert float(np.asarray(m.info['dt'])) == 0.5
    assert dillunp(np.asarray(m.info['commandlineArguments'])) == ['a=1', 'b=2']
    # the merged file is a sequential series: two ranks resume from it, each with its slab
    grids, _ = _grids(2, n=(6, 13))
    for g in grids:
        part = TimeSeries(out, grid=g, comm=g.comm, mode='r')
        lo, hi = g.ranges[-1]
        assert np.array_equal(part.retrieve_by_number(part.sorted_steps()[-1]), fb[1][..., lo:hi])
    # time window through the command line entry
    sel = str(tmp_path / 'sel')
    rc = subprocess.run([sys.executable, os.path.join(ROOT, 'tsmerge.py'), '-o', sel, '-s', '1',
                         '-e', '3', a + 's4@', b], capture_output=True, text=True, timeout=300)
    assert rc.returncode == 0, rc.stderr[-400:]
    assert np.array_equal(Gatherer(sel).sorted_times(), [1.0, 2.0, 3.0])


def test_store_slice_completes_a_time_point(tmp_path):
    from ksfd_b200.timeseries import TimeSeries
    _, gg = _grids(1)
    ts = TimeSeries(str(tmp_path / 'x'), grid=gg, comm=gg.comm, mode='w')
    f = np.random.default_rng(0).standard_normal(gg.globalVshape)
    ts.store_slice(((0, 6), (0, 4)), f[:, :, 0:4], 2.0)
    ts.store_slice(((0, 6), (4, 11)), f[:, :, 4:11], 2.0 * (1 + 1e-9))    # same point (tol 1e-7)
    ts.store_slice(((0, 6), (0, 11)), 2 * f, 2.1)                         # a new one
    assert np.array_equal(ts.times(), [2.0, 2.1])
    assert np.array_equal(ts.retrieve_by_number(0), f)
    assert np.array_equal(ts.retrieve_by_number(1), 2 * f)
    ts.close()


def test_timeseries_reference_surface(tmp_path):
    """the small members of the reference's KSFDTimeSeries the tools use (ksfdtimeseries.py:139-243)"""
    from ksfd_b200.timeseries import TimeSeries
    assert TimeSeries.parse_filename('bases2r1.h5') == ('base', 2, 1, False)
    assert TimeSeries.parse_filename('tests/test1MPI.h5') == ('tests/test1', 1, 0, True)
    with pytest.raises(ValueError):
        TimeSeries.parse_filename('nonsense.txt')
    prefix = str(tmp_path / 'seq')
    fields, gg = _write(prefix, 1, [0.0, 1.0])
    grids, _ = _grids(2)
    for g in grids:                                  # two ranks read the sequential file
        ts = TimeSeries(prefix, grid=g, comm=g.comm, mode='r')
        assert not ts.rank_owns_file and ts.tsf is ts.tsFile and ts.dim == 2 and ts.dof == 3
        assert ts.ranges == ((0, 6), (0, 11))        # the FILE holds the global ranges
        whole = np.array(ts.tsf['data1'])
        assert np.array_equal(whole[ts.myslice], ts.retrieve_by_number(1))
        assert np.array_equal(whole[ts.myslice], fields[1][..., g.ranges[-1][0]:g.ranges[-1][1]])
    own = TimeSeries(prefix, grid=gg, comm=gg.comm, mode='r')
    assert own.rank_owns_file and own.myslice == (slice(0, None),) * 3
    with pytest.raises(OSError):                     # nothing to read: an error, after the retries
        TimeSeries(str(tmp_path / 'missing'), grid=gg, comm=gg.comm, mode='r', retries=1,
                   retry_interval=0.01)
