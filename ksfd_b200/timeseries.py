"""
Time series / checkpoint files with the reference's layout
(KSFD/ksfdtimeseries.py; SURVEY 5 "Checkpoint / resume").

  file name   <prefix>s<size>r<rank>.h5
  /data<k>    dataset, shape Vlshape = (dof, nx_loc[, ny_loc[, nz_loc]]),
              C order, float64, attrs k, t
  /times /ks /order /lastk          bookkeeping (rewritten on flush/close)
  /size /rank /ranges               who wrote the file
  /grid/<attr>                      dim, dof, nps, bounds, spacing, order,
                                    stencil_width, stencil_type, boundary_type,
                                    global*shape, Slshape, Vlshape, ranges,
                                    Clshape, Cashape, coordsNoGhosts,
                                    coordsWithGhosts
  /info/<name>                      dill pickles as uint8 arrays
                                    (commandlineArguments, SolutionParameters,
                                    sources) and scalars dt, lastvart

Backend: h5py when importable.  Where h5py is absent (the build container) the
same keys are kept in a numpy .npz container with the suffix .npz — enough for
--save / --check / --resume round trips, not readable by the reference tools.
"""
import os

import numpy as np

try:
    import h5py
    HAVE_H5 = True
except ImportError:          # pragma: no cover - depends on the box
    h5py = None
    HAVE_H5 = False


def dillnp(*args, **kwargs):
    """pickle an object (dill) into a uint8 array (reference ksfdtsmaker.py:10-21)"""
    import dill
    return np.frombuffer(dill.dumps(*args, **kwargs), dtype=np.uint8).copy()


def dillunp(arr):
    import dill
    assert isinstance(arr, np.ndarray) and arr.dtype == np.uint8
    return dill.loads(arr.tobytes())


class _NpzStore(dict):
    """dict persisted as .npz; stands in for an h5py.File"""

    def __init__(self, filename, mode):
        super().__init__()
        self.filename = filename
        self.attrs = {}
        if mode in ('r', 'r+', 'a') and os.path.isfile(filename):
            with np.load(filename, allow_pickle=False) as z:
                for k in z.files:
                    self[k] = z[k]

    def flush(self):
        d = os.path.dirname(os.path.abspath(self.filename))
        os.makedirs(d, exist_ok=True)
        tmp = self.filename + '.tmp.npz'
        np.savez(tmp, **{k: np.asarray(v) for k, v in self.items()})
        os.replace(tmp, self.filename)

    def close(self):
        self.flush()


GRID_ATTRS = ['dim', 'dof', 'nps', 'bounds', 'spacing', 'order', 'stencil_width',
              'stencil_type', 'boundary_type', 'globalSshape', 'globalVshape',
              'globalCshape', 'Slshape', 'Vlshape', 'ranges', 'Clshape', 'Cashape',
              'coordsNoGhosts', 'coordsWithGhosts']


class TimeSeries:
    def __init__(self, basename, grid=None, comm=None, mpiok=False, mode='r+',
                 retries=0, retry_interval=60):
        self.grid = grid
        self.comm = comm if comm is not None else (grid.comm if grid is not None else None)
        self.size = self.comm.size if self.comm is not None else 1
        self.rank = self.comm.rank if self.comm is not None else 0
        self.mode = mode
        self.basename = basename
        suffix = '.h5' if HAVE_H5 else '.npz'
        name = '%ss%dr%d%s' % (basename, self.size, self.rank, suffix)
        seq = '%ss1r0%s' % (basename, suffix)
        if mode in ('r', 'a') and not os.path.isfile(name) and os.path.isfile(seq):
            name = seq
        self.filename = name
        self.creating = mode[0] in 'wx' or (mode != 'r' and not os.path.isfile(name))
        self._f = None
        self.ts = np.array([], dtype=float)
        self.ks = np.array([], dtype=int)
        self.lastk = -1
        self._open(mode)
        self.file_size = self.size
        if not self.creating:
            self._read_index()
            if self._has('size'):
                self.file_size = int(np.asarray(self._get('size')).reshape(-1)[0])
        else:
            self._set('size', self.size)
            self._set('rank', self.rank)
            if grid is not None:
                self._set('ranges', np.array(grid.ranges))
                self.grid_save()
        self.info = _InfoView(self)

    # -- backend ------------------------------------------------------------
    def _open(self, mode):
        if mode[0] in 'wxa' or mode == 'r+':
            d = os.path.dirname(os.path.abspath(self.filename))
            os.makedirs(d, exist_ok=True)
        if HAVE_H5:
            self._f = h5py.File(self.filename, mode)
        else:
            self._f = _NpzStore(self.filename, mode)

    def is_open(self):
        return self._f is not None

    tsFile = property(lambda s: s._f)

    def _set(self, key, val):
        if self.mode == 'r':
            return
        key = key.lstrip('/')
        if HAVE_H5:
            if key in self._f:
                del self._f[key]
            try:
                self._f[key] = val
            except (ValueError, TypeError):
                self._f[key] = str(val)
        else:
            self._f[key] = np.asarray(val) if not isinstance(val, str) else np.array(val)

    def _get(self, key):
        v = self._f[key.lstrip('/')]
        return v[()] if HAVE_H5 else (v[()] if v.shape == () else v)

    def _has(self, key):
        return key.lstrip('/') in self._f

    def set_info(self, name, val):
        self._set('info/' + name, val)

    def get_info(self, name):
        return self._get('info/' + name)

    def has_info(self, name):
        return self._has('info/' + name)

    def grid_save(self):
        for a in GRID_ATTRS:
            v = getattr(self.grid, a)
            self._set('grid/' + a, np.array(v) if not isinstance(v, str) else v)

    def _read_index(self):
        if self._has('times'):
            self.ts = np.array(self._get('times'), dtype=float).reshape(-1)
            self.ks = (np.array(self._get('ks'), dtype=int).reshape(-1)
                       if self._has('ks') else np.arange(len(self.ts)))
            self.lastk = int(self._get('lastk')) if self._has('lastk') else len(self.ts) - 1

    def _sort(self):
        self.order = np.argsort(self.ts, kind='stable')
        self.sts = np.sort(self.ts)
        self._set('times', self.ts)
        self._set('order', self.order)
        self._set('ks', self.ks)
        self._set('lastk', self.lastk)

    def flush(self):
        self._sort()
        self._f.flush()

    def temp_close(self):
        self._sort()
        self._f.close()
        self._f = None

    def reopen(self):
        self._open('r' if self.mode == 'r' else 'r+')

    def close(self):
        if self._f is None:
            self.reopen()
        self._sort()
        self._f.close()
        self._f = None

    # -- data ---------------------------------------------------------------
    def store(self, data, t, k=None):
        arr = data.array_r if hasattr(data, 'array_r') else np.asarray(data)
        vals = np.ascontiguousarray(arr.reshape(self.grid.Vlshape, order='F'))
        if k is None:
            k = self.lastk + 1
        self.lastk = k
        self.ks = np.append(self.ks, k)
        self.ts = np.append(self.ts, t)
        key = 'data' + str(k)
        if HAVE_H5:
            if key in self._f:
                del self._f[key]
            ds = self._f.create_dataset(key, data=vals)
            ds.attrs['k'] = k
            ds.attrs['t'] = t
        else:
            self._f[key] = vals
            self._f[key + '.k'] = np.array(k)
            self._f[key + '.t'] = np.array(t)
        self.flush()

    def times(self):
        return self.ts

    def steps(self):
        return self.ks

    def sorted_times(self):
        return np.sort(self.ts)

    def retrieve_by_number(self, k):
        """Values of step k on THIS rank's part of the grid.  A file written by a
        different number of ranks (the sequential <prefix>s1r0 file of a merged or
        single-rank run) holds the global array: the rank keeps its own slab, as the
        reference does with `myslice` (KSFD/ksfdtimeseries.py:560-600)."""
        a = np.array(self._get('data' + str(int(k))))
        g = self.grid
        if (g is not None and self.file_size != self.size
                and tuple(a.shape[1:]) == tuple(g.globalSshape)
                and tuple(g.Slshape) != tuple(g.globalSshape)):
            lo, hi = g.ranges[-1]
            a = np.ascontiguousarray(a[..., lo:hi])
        return a

    def retrieve_by_time(self, t):
        order = np.argsort(self.ts, kind='stable')
        sts = self.ts[order]
        if t <= sts[0]:
            return self.retrieve_by_number(self.ks[order[0]])
        if t >= sts[-1]:
            return self.retrieve_by_number(self.ks[order[-1]])
        b = int(np.searchsorted(sts, t))
        if sts[b] == t:
            return self.retrieve_by_number(self.ks[order[b]])
        a = b - 1
        A = self.retrieve_by_number(self.ks[order[a]])
        B = self.retrieve_by_number(self.ks[order[b]])
        return ((t - sts[a]) * B + (sts[b] - t) * A) / (sts[b] - sts[a])


class _InfoView:
    """`ts.info['name']` / `in` / item assignment, like the h5py group."""

    def __init__(self, ts):
        self._ts = ts

    def __contains__(self, name):
        return self._ts.has_info(name)

    def __getitem__(self, name):
        return self._ts.get_info(name)

    def __setitem__(self, name, val):
        self._ts.set_info(name, val)
