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

`Gatherer` reads the per-rank files of a multi-rank run one after the other and
`tsmerge` (the root `tsmerge.py` entry) gathers / merges them into one sequential
<prefix>s1r0 file, as the reference's Gatherer and tsmerge.py do
(KSFD/ksfdtimeseries.py:674-828, tsmerge.py:40-111).

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
        mpi = '%sMPI%s' % (basename, suffix)        # one file for all ranks (parallel HDF5 builds)
        self.rank_owns_file = True
        if mode in ('r', 'a') and not os.path.isfile(name):
            if os.path.isfile(seq):
                name = seq
                self.rank_owns_file = self.size == 1
            elif os.path.isfile(mpi):               # written elsewhere with parallel HDF5: readable
                name = mpi
                self.rank_owns_file = self.size == 1
        self.filename = name
        self.retries, self.retry_interval = int(retries), retry_interval
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
        if self.grid is None and not self.creating and self._has('grid/dim'):
            self.grid_load()            # as the reference: the grid comes from the file
        self.info = _InfoView(self)

    # -- backend ------------------------------------------------------------
    def _open(self, mode):
        if mode[0] in 'wxa' or mode == 'r+':
            d = os.path.dirname(os.path.abspath(self.filename))
            os.makedirs(d, exist_ok=True)
        # a file another process is still writing may fail to open: try again `retries`
        # times, `retry_interval` seconds apart (reference open_with_retry)
        import time
        left = getattr(self, 'retries', 0)
        while True:
            try:
                if HAVE_H5:
                    self._f = h5py.File(self.filename, mode)
                else:
                    if mode == 'r' and not os.path.isfile(self.filename):
                        raise OSError('no such series file: ' + self.filename)
                    self._f = _NpzStore(self.filename, mode)
                return
            except (OSError, ValueError, EOFError):
                if left <= 0:
                    raise
                left -= 1
                time.sleep(self.retry_interval)

    def is_open(self):
        return self._f is not None

    tsFile = tsf = property(lambda s: s._f)
    dim = property(lambda s: s.grid.dim)
    dof = property(lambda s: s.grid.dof)

    @property
    def ranges(self):
        """index ranges of the data in this file (the global ones for a shared / sequential file)"""
        if self.rank_owns_file or self.grid is None:
            return tuple(self.grid.ranges) if self.grid is not None else None
        return tuple((0, int(m)) for m in self.grid.nps)

    @property
    def myslice(self):
        """slice of the file's arrays that belongs to this rank (reference set_grid)"""
        if self.rank_owns_file:
            return (slice(0, None),) * (self.grid.dim + 1)
        return (slice(0, None),) + tuple(slice(*r) for r in self.grid.ranges)

    def set_grid(self, grid):
        self.grid = grid
        if self.mode != 'r':
            self._set('ranges', np.array(self.ranges))

    @staticmethod
    def parse_filename(filename):
        """'bases2r1.h5' -> ('base', 2, 1, False); 'baseMPI.h5' -> ('base', 1, 0, True)"""
        import re
        res = re.fullmatch(r'(.*)MPI\.(?:h5|npz)', filename)
        if res:
            return (res[1], 1, 0, True)
        res = re.fullmatch(r'(.*)s(\d+)r(\d+)\.(?:h5|npz)', filename)
        if res:
            return (res[1], int(res[2]), int(res[3]), False)
        raise ValueError("Couldn't parse filename %s" % filename)

    def get_filename(self, *a, **k):
        return self.filename

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

    def grid_read(self):
        """grid parameters of the open file as a dict (reference ksfdtimeseries.py:264-291)"""
        gd = {}
        for a in GRID_ATTRS:
            if not self._has('grid/' + a):
                gd[a] = None
                continue
            val = np.asarray(self._get('grid/' + a))
            if val.dtype.kind in 'SUO':
                v = val.item() if val.shape == () else val
                gd[a] = v.decode() if isinstance(v, bytes) else str(v)
            elif a.endswith('shape'):
                gd[a] = tuple(int(x) for x in val.reshape(-1))
            elif val.shape == ():
                gd[a] = val.item()
            else:
                gd[a] = val
        dim = gd['dim']
        gd['width'] = float(gd['bounds'][0])
        gd['height'] = float(gd['bounds'][1]) if dim > 1 else 1.0
        gd['depth'] = float(gd['bounds'][2]) if dim > 2 else 1.0
        gd['nx'] = int(gd['nps'][0])
        gd['ny'] = int(gd['nps'][1]) if dim > 1 else 8
        gd['nz'] = int(gd['nps'][2]) if dim > 2 else 8
        return gd

    def grid_load(self, gd=None):
        """a sequential Grid (one rank owns the whole domain) from the file's grid parameters
        (reference ksfdtimeseries.py:293-310)"""
        from .grid import Comm, Grid
        if gd is None:
            gd = self.grid_read()
        self.grid = Grid(dim=gd['dim'], width=gd['width'], height=gd['height'], depth=gd['depth'],
                         nx=gd['nx'], ny=gd['ny'], nz=gd['nz'], dof=gd['dof'], order=gd['order'],
                         stencil_width=gd['stencil_width'], stencil_type=gd['stencil_type'],
                         boundary_type=gd['boundary_type'], comm=Comm(0, 1))
        return self.grid

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

    def find_time(self, t):
        """(na, nb, ta, tb): step numbers and times of the stored points that bracket t
        (reference ksfdtimeseries.py:574-604)"""
        order = np.argsort(self.ts, kind='stable')
        sts = self.ts[order]
        if sts.size == 0:
            return 0, 0, t - 1.0, t - 1.0
        if t <= sts[0]:
            a = int(self.ks[order[0]])
            return a, a, float(sts[0]), float(sts[0])
        if t >= sts[-1]:
            b = int(self.ks[order[-1]])
            return b, b, float(sts[-1]), float(sts[-1])
        b = int(np.searchsorted(sts, t))
        if sts[b] == t:
            kb = int(self.ks[order[b]])
            return kb, kb, float(sts[b]), float(sts[b])
        return int(self.ks[order[b - 1]]), int(self.ks[order[b]]), float(sts[b - 1]), float(sts[b])

    def store_slice(self, ranges, data, t, tol=1e-7):
        """Store the part `ranges` (one (start, end) per axis) of the time point t: a point
        within `tol` (relative) of a stored time is completed, else a new point is appended
        (reference ksfdtimeseries.py:511-549; what tsmerge writes the gathered slabs with)."""
        shape = (self.grid.dof,) + tuple(int(r[1]) - int(r[0]) for r in ranges)
        slc = (slice(0, None),) + tuple(slice(int(r[0]), int(r[1])) for r in ranges)
        vals = np.asarray(data).reshape(shape, order='F')
        na, nb, ta, tb = self.find_time(t)
        n, tn = (na, ta) if abs(t - ta) <= abs(tb - t) else (nb, tb)
        new = self.ts.size == 0 or ((not (t == 0.0 and tn == 0.0)) and
                                    abs(t - tn) / max(abs(t), abs(tn)) > tol)
        if new:
            k = self.lastk + 1
            self.lastk = k
            self.ks = np.append(self.ks, k)
            self.ts = np.append(self.ts, t)
            full = np.zeros(self.grid.Vlshape, dtype=vals.dtype)
        else:
            k = n
            full = np.array(self._get('data' + str(k)))
        full[slc] = vals
        key = 'data' + str(k)
        if HAVE_H5:
            if key in self._f:
                del self._f[key]
            ds = self._f.create_dataset(key, data=full)
            ds.attrs['k'] = k
            ds.attrs['t'] = t if new else tn
        else:
            self._f[key] = full
            self._f[key + '.k'] = np.array(k)
            self._f[key + '.t'] = np.array(t if new else tn)

    def times(self):
        return self.ts

    def steps(self):
        return self.ks

    def sorted_times(self):
        return np.sort(self.ts)

    def sorted_steps(self):
        return self.ks[np.argsort(self.ts, kind='stable')]

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


class _Self:
    """rank/size of a sequential reader (MPI.COMM_SELF in the reference)"""

    def __init__(self, rank=0, size=1):
        self.rank, self.size = rank, size


class Gatherer:
    """
    Sequential reader of the per-rank files `<base>s<size>r<rank>` a multi-rank run wrote
    (reference KSFD/ksfdtimeseries.py:674-828, same use):

        gather = Gatherer('bases4@')            # or Gatherer('base', size=4)
        grid = gather.grid                      # the GLOBAL grid, owned by one rank
        glob = np.empty(grid.globalVshape)
        for series in gather:                   # one file after the other
            glob[series.slice] = series.retrieve_by_number(k)

    `basename` may have the reference's special form '<base>s<n>@...'; a plain prefix with
    size=None (or 1) reads the sequential file `<base>s1r0`.  Read-only.
    """

    def __init__(self, basename, size=None, retries=0, retry_interval=60):
        import re
        m = re.fullmatch(r'(.+)s(\d+)@.*', basename)
        if m:
            basename, size = m.group(1), int(m.group(2))
        if size is None:
            size = 1
        if not isinstance(size, int) or size <= 0:
            raise ValueError('size %r is not a positive int' % (size,))
        self.basename, self.size = basename, size
        self.retries, self.retry_interval = retries, retry_interval
        self._ts = None
        self._open_rank(0)
        self.grid = self._ts.grid_load()        # global: what the merged series is defined on
        self.iter_started = self.iter_stopped = False

    def _drop(self):
        """leave the open file without writing anything back (read-only)"""
        if self._ts is not None and self._ts._f is not None:
            if HAVE_H5:
                self._ts._f.close()
            self._ts._f = None

    def _open_rank(self, rank):
        self._drop()
        self.rank = rank
        suffix = '.h5' if HAVE_H5 else '.npz'
        name = '%ss%dr%d%s' % (self.basename, self.size, rank, suffix)
        if not os.path.isfile(name):
            raise FileNotFoundError(name)
        ts = self._ts = TimeSeries(self.basename, grid=None, comm=_Self(rank, self.size), mode='r')
        rg = np.asarray(ts._get('ranges')).reshape(-1, 2)
        self.ranges = tuple((int(a), int(b)) for a, b in rg)
        self.dof = int(np.asarray(ts._get('grid/dof')).reshape(-1)[0])
        self.shape = (self.dof,) + tuple(b - a for a, b in self.ranges)
        self.slice = (slice(0, None),) + tuple(slice(a, b) for a, b in self.ranges)

    # the open file answers like a TimeSeries
    tsf = tsFile = property(lambda s: s._ts._f)
    info = property(lambda s: s._ts.info)
    filename = property(lambda s: s._ts.filename)

    def times(self):
        return self._ts.times()

    def steps(self):
        return self._ts.steps()

    def sorted_times(self):
        return self._ts.sorted_times()

    def sorted_steps(self):
        return self._ts.sorted_steps()

    def retrieve_by_number(self, k):
        """the part of step k this file holds, shape `self.shape`"""
        return np.array(self._ts._get('data' + str(int(k))))

    def retrieve_by_time(self, t):
        return self._ts.retrieve_by_time(t)

    def close(self):
        self._drop()

    def __iter__(self):
        return self

    def __next__(self):
        if self.iter_stopped:                   # exhausted before: start again
            self._open_rank(0)
        elif self.iter_started:
            if self.rank + 1 >= self.size:
                self.iter_stopped = True
                self.iter_started = False
                raise StopIteration
            self._open_rank(self.rank + 1)
        self.iter_started, self.iter_stopped = True, False
        return self


def tsmerge(outfile, infiles, start=0.0, end=None, verbose=0):
    """
    Gather and/or merge time series into ONE sequential series `<outfile>s1r0`
    (reference tsmerge.py:40-111): every input is a prefix (a series in one file) or
    '<prefix>s<n>@' (n per-rank files to gather); time points outside [start, end] are left
    out; `/info` is taken from the first input.  Returns the output file name.
    """
    grid, plan = None, []
    for name in infiles:
        g = Gatherer(name)
        if grid is None:
            grid = g.grid
        order = np.argsort(g.times(), kind='stable')
        plan.append((name, g.steps()[order], g.times()[order]))
        g.close()
    out = TimeSeries(outfile, grid=grid, comm=_Self(0, 1), mode='w')
    have_info = False
    for name, fsteps, ftimes in plan:
        g = Gatherer(name)
        if not have_info:
            if HAVE_H5:
                if 'info' in g.tsf:
                    for key in g.tsf['info'].keys():
                        out._set('info/' + key, g.tsf['info'][key][()])
            else:
                for key in [k for k in g.tsf.keys() if k.startswith('info/')]:
                    out._f[key] = g.tsf[key]
            have_info = True
        for s in g:
            if verbose:
                print(s.filename, flush=True)
            for k, t in zip(fsteps, ftimes):
                if t < start or (end is not None and t > end):
                    continue
                out.store_slice(s.ranges, s.retrieve_by_number(k), t)
        g.close()
    out.close()
    return out.filename
