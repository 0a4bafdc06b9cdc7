"""
HDF5 on-disk layout of TimeSeries (reference KSFD/ksfdtimeseries.py:188-262,
484-509).  Needs h5py, which neither the build container nor the GPU image has: the test
skips there (the .npz stand-in is covered by test_host_mirror / test_gpu_dropin) and runs
wherever the reference's own stack is installed.  It pins, with raw h5py calls, exactly what
the reference's readers (`TimeSeries(mode='r')`, `Gatherer`, tsmerge.py) look up:
file name <prefix>s<size>r<rank>.h5, /data<k> of shape Vlshape (C order, float64) with
attributes k and t, /times /ks /order /lastk, /size /rank /ranges, /grid/<attr>, /info/<name>.
"""
import os

import numpy as np
import pytest

h5py = pytest.importorskip('h5py')


def test_h5_layout_is_the_reference_layout(tmp_path):
    from ksfd_b200 import timeseries
    from ksfd_b200.grid import Comm, Grid
    assert timeseries.HAVE_H5
    grid = Grid(dim=2, nx=8, ny=6, dof=3, width=2.0, height=1.5, comm=Comm(0, 1))
    prefix = str(tmp_path / 'run' / 'series')
    ts = timeseries.TimeSeries(prefix, grid=grid, mode='w', comm=Comm(0, 1))
    ts.info['dt'] = 0.25
    ts.info['blob'] = timeseries.dillnp({'a': 1})
    rng = np.random.default_rng(0)
    frames = [rng.standard_normal(grid.Vlshape) for _ in range(3)]
    for k, f in enumerate(frames):
        ts.store(f.reshape(-1, order='F'), 0.5 * k)
    ts.close()
    fname = prefix + 's1r0.h5'
    assert os.path.isfile(fname)
    with h5py.File(fname, 'r') as f:
        for k, fr in enumerate(frames):
            d = f['data%d' % k]
            assert d.shape == grid.Vlshape and d.dtype == np.float64
            assert np.array_equal(d[()], fr)
            assert d.attrs['k'] == k and d.attrs['t'] == 0.5 * k
        assert np.array_equal(f['times'][()], [0.0, 0.5, 1.0])
        assert np.array_equal(f['ks'][()], [0, 1, 2]) and int(f['lastk'][()]) == 2
        assert np.array_equal(f['order'][()], [0, 1, 2])
        assert int(f['size'][()]) == 1 and int(f['rank'][()]) == 0
        assert np.array_equal(f['ranges'][()], np.array(grid.ranges))
        for a in timeseries.GRID_ATTRS:
            assert 'grid/' + a in f, a
        assert np.array_equal(f['grid/nps'][()], [8, 6])
        assert np.allclose(f['grid/coordsNoGhosts'][()], grid.coordsNoGhosts)
        assert float(f['info/dt'][()]) == 0.25
        assert timeseries.dillunp(f['info/blob'][()]) == {'a': 1}
    back = timeseries.TimeSeries(prefix, grid=grid, mode='r', comm=Comm(0, 1))
    assert np.array_equal(back.sorted_times(), [0.0, 0.5, 1.0])
    assert np.array_equal(back.retrieve_by_time(0.75), 0.5 * (frames[1] + frames[2]))
    back.close()
