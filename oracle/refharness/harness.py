"""
Run the UNMODIFIED reference `KSFD.Derivatives` (and friends) from
/root/reference over numpy stand-ins for PETSc Vec / DMDA / Mat, one process,
periodic grid.  Used ONLY to generate golden vectors (oracle/make_golden.py)
and to cross-check the oracle restatement in the build container.

TEST INFRASTRUCTURE ONLY — never imported by the product, bench.py's own arm,
or any -m gpu test (it needs /root/reference, absent on the GPU box).

Semantics faked, with the reference call sites that rely on them:
  DMDA.globalToLocal   periodic ghost fill, width sw, == np.pad(mode='wrap')
                       (ksfdsym.py:703-705, 919-920, 1203)
  DMDA.getRanges       single rank owns everything (ksfdgrid.py:169)
  MatSetValuesStencil  (i+di, j+dj, k+dk, c) wraps periodically; ADD mode;
                       c<0 skips (ksfdMat.pyx:280-325)
"""
import os
import sys

import numpy as np

REFERENCE_ROOT = os.environ.get('KSFD_REFERENCE_ROOT', '/root/reference')


class FakeKsfdMat:
    """Records COO triplets with MatSetValuesStencil wrap semantics."""

    def __init__(self, mat):
        self.dmda = mat                     # FakeDMDA.getMatrix() returns the dmda
        self.rows = []
        self.cols = []
        self.vals = []

    def setOption(self, *a, **k):
        pass

    def setUp(self):
        pass

    def assemble(self):
        pass

    def zeroEntries(self):
        self.rows, self.cols, self.vals = [], [], []

    def _flat(self, ijkc):
        n = self.dmda.n3
        dof = self.dmda.dof
        i = np.mod(ijkc[..., 0], n[0])
        j = np.mod(ijkc[..., 1], n[1])
        k = np.mod(ijkc[..., 2], n[2])
        return ijkc[..., 3] + dof * (i + n[0] * (j + n[1] * k))

    def setValuesJacobian(self, rows, col_offsets, values, insert_mode=None):
        rows = np.asarray(rows)
        col_offsets = np.asarray(col_offsets)
        values = np.asarray(values, dtype=float)
        nr, nc = values.shape
        assert rows.shape == (nr, 4)
        assert col_offsets.shape == (nc, 4), 'only the (nc,4) form is used'
        rflat = self._flat(rows)
        for c in range(nc):
            if col_offsets[c, 3] < 0:
                continue
            cc = rows.copy()
            cc[:, :3] += col_offsets[c, :3]
            cc[:, 3] = col_offsets[c, 3]
            ok = rows[:, 3] >= 0
            self.rows.append(rflat[ok])
            self.cols.append(self._flat(cc)[ok])
            self.vals.append(values[ok, c])
        return dict(nrows=nr, ncols=nc, row_type=rows.dtype)

    def tocsr(self):
        import scipy.sparse as sp
        N = self.dmda.dof * int(np.prod(self.dmda.n3))
        r = np.concatenate(self.rows)
        c = np.concatenate(self.cols)
        v = np.concatenate(self.vals)
        return sp.coo_matrix((v, (r, c)), shape=(N, N)).tocsr()


_classes = {}


def load_reference():
    """Install stubs, import the reference package, build fake classes."""
    if _classes:
        return _classes
    from . import stubs
    stubs.install()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    # numpy 2 removed np.product (used at KSFD/ksfdrandom.py:197)
    if not hasattr(np, 'product'):
        np.product = np.prod
    import sympy as sy
    # sympy >= 1.9 raises when '' reaches Basic.subs (ksfdsoln.py:321); older
    # sympy silently skipped such pairs.  Restore the old behaviour.
    if not getattr(sy.Basic.subs, '_ksfd_patched', False):
        _orig_subs = sy.Basic.subs

        def _subs(self, *args, **kwargs):
            if len(args) == 1 and hasattr(args[0], 'items'):
                d = {k: v for k, v in args[0].items()
                     if not (v is None or (isinstance(v, str) and v == ''))}
                return _orig_subs(self, d, **kwargs)
            return _orig_subs(self, *args, **kwargs)
        _subs._ksfd_patched = True
        sy.Basic.subs = _subs
    import petsc4py
    import KSFD
    import KSFD.ksfdsym as _ksym
    # old sympy sympified arbitrary objects through str(); ksfdsolver2.py:497
    # relies on that to re-wrap a SpatialExpression.  Restore it.
    if not getattr(_ksym.safe_sympify, '_ksfd_patched', False):
        _orig_ss = _ksym.safe_sympify

        def _ss(exp):
            if isinstance(exp, _ksym.SpatialExpression):
                return exp.expression
            return _orig_ss(exp)
        _ss._ksfd_patched = True
        _ksym.safe_sympify = _ss

    class FakeVec(petsc4py.PETSc.Vec):
        def __init__(self, n):
            self._a = np.zeros(n, dtype=float)

        @property
        def array(self):
            return self._a

        @array.setter
        def array(self, v):
            self._a[:] = np.asarray(v, dtype=float).reshape(-1, order='F')

        def assemble(self):
            pass

        def setUp(self):
            pass

        def destroy(self):
            pass

        def zeroEntries(self):
            self._a[:] = 0.0

        def duplicate(self):
            return FakeVec(self._a.size)

        def copy(self, dst=None):
            if dst is None:
                dst = self.duplicate()
            dst._a[:] = self._a
            return dst

    class FakeDMDA:
        def __init__(self, grid, dof):
            self.grid = grid
            self.dof = dof
            self.dim = grid.dim
            self.n = tuple(int(x) for x in grid.globalSshape)
            self.n3 = self.n + (1,) * (3 - self.dim)
            self.sw = grid.stencil_width
            self._free = []

        def setUniformCoordinates(self, **k):
            pass

        def setFromOptions(self):
            pass

        def setUp(self):
            pass

        def destroy(self):
            pass

        def getRanges(self):
            return tuple((0, m) for m in self.n)

        def createGlobalVec(self):
            return FakeVec(self.dof * int(np.prod(self.n)))

        def createLocalVec(self):
            na = tuple(m + 2 * self.sw for m in self.n)
            return FakeVec(self.dof * int(np.prod(na)))

        def getLocalVec(self):
            return self._free.pop() if self._free else self.createLocalVec()

        def restoreLocalVec(self, v):
            self._free.append(v)

        def globalToLocal(self, g, l):
            a = g.array.reshape((self.dof,) + self.n, order='F')
            p = np.pad(a, [(0, 0)] + [(self.sw, self.sw)] * self.dim,
                       mode='wrap')
            l.array[:] = p.reshape(-1, order='F')

        def getCoordinates(self):
            g = self.grid
            axes = [np.arange(m) * (g.bounds[d] / m)
                    for d, m in enumerate(self.n)]
            mesh = np.meshgrid(*axes, indexing='ij')
            c = np.stack(mesh, axis=0)           # (dim,)+n
            v = FakeVec(c.size)
            v.array[:] = c.reshape(-1, order='F')
            return v

        def getMatrix(self):
            return self

    class HarnessGrid(KSFD.Grid):
        def make_dmda(self, dof=1):
            return FakeDMDA(self, dof)

    _classes.update(KSFD=KSFD, FakeVec=FakeVec, FakeDMDA=FakeDMDA,
                    Grid=HarnessGrid, sy=sy)
    return _classes


def load_solver_module():
    """Import /root/reference/ksfdsolver2.py (for parse_commandline etc.)."""
    import importlib.util
    load_reference()
    spec = importlib.util.spec_from_file_location(
        'ksfdsolver2_ref', os.path.join(REFERENCE_ROOT, 'ksfdsolver2.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def make_problem(args, with_sources=True):
    """
    Build (clargs, ps, grid, derivs) from a ksfdsolver2-style argument list,
    following ksfdsolver2.main (ksfdsolver2.py:647-705) without PETSc.
    """
    C = load_reference()
    KSFD = C['KSFD']
    solver = load_solver_module()
    clargs = solver.parse_commandline(list(args))
    ps = KSFD.SolutionParameters(clargs)
    grid = C['Grid'](dim=ps.dim, dof=ps.nligands + 1,
                     width=ps.width, height=ps.height, depth=ps.depth,
                     nx=ps.nwidth, ny=ps.nheight, nz=ps.ndepth)
    sources = None
    if with_sources:
        sources = solver.decode_sources(clargs.source, ps, grid)
    u0 = grid.Vdmda.createGlobalVec()
    derivs = KSFD.Derivatives(ps, grid, sources=sources, u0=u0)
    return clargs, ps, grid, derivs
