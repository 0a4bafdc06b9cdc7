"""
Stub modules that let the UNMODIFIED reference package at /root/reference be
imported in the build container, where PETSc, MPI, h5py and dogpile.cache are
absent.

TEST INFRASTRUCTURE ONLY.  This is used by oracle/make_golden.py (run once in
the build container, outputs committed under tests/golden/) and by nothing in
the product.  It reads /root/reference, which does not exist on the GPU box,
so nothing under tests/ -m gpu, bench.py or __graft_entry__ imports it.

What is faked and why (reference file:line that needs it):
  mpi4py.MPI            ksfdsym.py:16-18, ksfdufunc.py:253-259, ksfdts.py:244,311
  petsc4py(.PETSc)      ksfdgrid.py:136-139, ksfdsym.py:783,809,845-848
  dogpile.cache         ksfdsym.py:34-50, ksfdufunc.py:30-46,267
  h5py                  ksfdtimeseries.py:53 (import only)
  ksfdMat               ksfdmat.py:10 (the Cython shim; see harness.FakeKsfdMat)
"""
import sys
import types


class _Comm:
    rank = 0
    size = 1

    def bcast(self, x, root=0):
        return x

    def allreduce(self, x, op=None):
        return x

    def Barrier(self):
        pass

    def tompi4py(self):
        return self

    def Get_rank(self):
        return 0

    def Get_size(self):
        return 1


class _Chain:
    """Object whose arbitrary attribute chains resolve (enum stand-ins)."""

    def __init__(self, name):
        self._name = name

    def __getattr__(self, item):
        if item.startswith('__'):
            raise AttributeError(item)
        return _Chain(self._name + '.' + item)

    def __call__(self, *a, **k):
        return _Chain(self._name + '()')

    def __repr__(self):
        return self._name

    def __eq__(self, other):
        return isinstance(other, _Chain) and other._name == self._name

    def __hash__(self):
        return hash(self._name)

    def __reduce__(self):
        return (_Chain, (self._name,))


class _NullRegion:
    def configure(self, *a, **k):
        return self

    def get(self, key=None):
        return None

    def set(self, key=None, value=None):
        pass

    def cache_on_arguments(self, *a, **k):
        def deco(fn):
            return fn
        return deco


def install():
    """Insert the stub modules into sys.modules (idempotent)."""
    if 'petsc4py' in sys.modules and getattr(
            sys.modules['petsc4py'], '_ksfd_stub', False):
        return
    # ---- mpi4py
    mpi4py = types.ModuleType('mpi4py')
    MPI = types.ModuleType('mpi4py.MPI')
    MPI.COMM_WORLD = _Comm()
    MPI.COMM_SELF = _Comm()
    MPI.INT64_T = 'INT64_T'
    MPI.SUM = 'SUM'
    MPI.MAX = 'MAX'
    MPI.Comm = _Comm
    mpi4py.MPI = MPI
    sys.modules['mpi4py'] = mpi4py
    sys.modules['mpi4py.MPI'] = MPI

    # ---- petsc4py
    petsc4py = types.ModuleType('petsc4py')
    petsc4py._ksfd_stub = True

    class Vec:          # real class: isinstance() at ksfdsym.py:783,809
        pass

    class Comm:
        pass

    class TS:           # so that ksfdts.py (class KSFDTS(PETSc.TS)) imports
        Type = _Chain('TS.Type')
        ExactFinalTime = _Chain('TS.ExactFinalTime')
        ProblemType = _Chain('TS.ProblemType')
        EquationType = _Chain('TS.EquationType')

    class _PETSc(types.ModuleType):
        def __getattr__(self, item):
            if item.startswith('__'):
                raise AttributeError(item)
            return _Chain('PETSc.' + item)

    PETSc = _PETSc('petsc4py.PETSc')
    PETSc.Vec = Vec
    PETSc.Comm = Comm
    PETSc.TS = TS
    petsc4py.PETSc = PETSc
    petsc4py.init = lambda *a, **k: None
    sys.modules['petsc4py'] = petsc4py
    sys.modules['petsc4py.PETSc'] = PETSc

    # ---- dogpile.cache
    dogpile = types.ModuleType('dogpile')
    cache = types.ModuleType('dogpile.cache')
    cache.make_region = lambda *a, **k: _NullRegion()
    backends = types.ModuleType('dogpile.cache.backends')
    null = types.ModuleType('dogpile.cache.backends.null')
    null.NullBackend = object
    backends.null = null
    cache.backends = backends
    dogpile.cache = cache
    sys.modules['dogpile'] = dogpile
    sys.modules['dogpile.cache'] = cache
    sys.modules['dogpile.cache.backends'] = backends
    sys.modules['dogpile.cache.backends.null'] = null

    # ---- h5py (import only)
    if 'h5py' not in sys.modules:
        try:
            import h5py  # noqa: F401
        except ImportError:
            sys.modules['h5py'] = types.ModuleType('h5py')

    # ---- ksfdMat: top-level module, class supplied by harness
    from . import harness
    km = types.ModuleType('ksfdMat')
    km.ksfdMat = harness.FakeKsfdMat
    sys.modules['ksfdMat'] = km
