"""
`Derivatives` and `SpatialExpression` with the reference interface
(KSFD/ksfdsym.py:145-1209, 1515-1697), evaluated by the CUDA library.

The reference builds sympy stencil expressions, turns them into compiled numpy
ufuncs and assembles a PETSc matrix.  Here the stencil structure is fixed in
hand-written sm_100a kernels; what remains variable — every physical
parameter (possibly time dependent), the finite-difference weights, the number
and grouping of ligands, the cap potential — is passed as a plain-number block
(`ksfd_physics`) refreshed whenever t changes a parameter.

    derivs.dfdt(fvec, t, out)      -> ksfd_residual(udot = NULL)
    derivs.Jacobian(fvec, t, out)  -> ksfd_jvp_setup; returns a matrix-free
                                      operator handle (JacobianOperator)
    derivs.velocity(fvec, t, out)  -> ksfd_velocity
    derivs.groom(farr)             -> numpy clamp for host arrays (same rule as
                                      the device clamp, KSFD/ksfdsym.py:888-900)
"""
import numpy as np
import sympy as sy

from . import core
from .params import KSFDException, safe_sympify
from .vec import Vec


class SpatialExpression:
    """A function of space and time given as a sympy expression in x, y, z, t
    and command-line parameters (reference KSFD/ksfdsym.py:1515-1697)."""

    def __init__(self, ps, grid, expression='0.0'):
        self._ps = ps
        self._grid = grid
        self.expression = expression

    ps = property(lambda s: s._ps)
    grid = property(lambda s: s._grid)

    @property
    def expression(self):
        return self._expression

    @expression.setter
    def expression(self, exp):
        if isinstance(exp, SpatialExpression):
            exp = exp.expression
        exp = safe_sympify(exp)
        if exp is None:
            exp = sy.Float(0.0)
        exp = sy.sympify(exp)
        tds = self._ps.time_dependent_symbols()
        sub = {sy.Symbol(k): v for k, v in tds.items()
               if not (v is None or v == '' or isinstance(v, bool))}
        self._expression = exp.subs(sub)
        self._fn = None
        dim = self._grid.dim
        coords = tuple(sy.symbols('x y z')[:dim]) + (sy.Symbol('t'),)
        td = self._expression.free_symbols.difference(coords)
        unknown = td.difference(sy.symbols(list(self._ps.tdfuncs.keys())))
        if unknown:
            raise ValueError('unknown symbol(s) %s' % sorted(map(str, unknown)))
        self._inputs = list(coords) + sorted(td, key=str)

    inputs = property(lambda s: s._inputs)

    @property
    def time_dependent(self):
        fs = self._expression.free_symbols
        return bool(fs.difference(sy.symbols('x y z')[:self._grid.dim]))

    @property
    def is_zero(self):
        return self._expression == 0

    def __call__(self, t=None, out=None, **kw):
        return self.call(t=t, out=out)

    def call(self, t=None, out=None, **kw):
        """Evaluate on the grid's owned points; returns shape grid.Slshape."""
        t = self._ps.t0 if t is None else t
        if self._fn is None:
            self._fn = sy.lambdify(self._inputs, self._expression, 'numpy')
        c = self._grid.coordsNoGhosts
        vals = self._ps.values(t)
        args = [c[i] for i in range(self._grid.dim)] + [t]
        args += [vals[str(s)] for s in self._inputs[self._grid.dim + 1:]]
        r = np.broadcast_to(np.asarray(self._fn(*args), dtype=float),
                            self._grid.Slshape)
        if out is not None:
            tgt = out[0] if isinstance(out, tuple) else out
            tgt[...] = r
            return tgt
        return np.array(r)

    def __str__(self):
        return str(self._expression)

    __repr__ = __str__

    def __getstate__(self):
        return dict(ps=self._ps, grid=self._grid, expression=str(self._expression))

    def __setstate__(self, st):
        self._ps, self._grid = st['ps'], st['grid']
        self.expression = st['expression']


class JacobianOperator:
    """Matrix-free stand-in for the reference's assembled ksfdMat: the device
    holds the linearisation point's coefficient field and the block-Jacobi
    preconditioner of  A = shift*I - df/du  (reference implicitIJ,
    KSFD/ksfdts.py:598-640)."""

    def __init__(self, derivs):
        self.derivs = derivs
        self.shift = 0.0
        self._u = None

    def setup(self, u_dev, shift):
        self._u = u_dev
        self.shift = float(shift)
        self.derivs.ctx.jvp_setup(u_dev, self.shift)

    def mult(self, x, y):
        """y = A x  (Vecs)"""
        ctx = self.derivs.ctx
        ctx.jvp(x.device(ctx), y.device(ctx))
        y.mark_device_written()
        return y

    def assemble(self):
        pass

    def zeroEntries(self):
        pass

    def setOption(self, *a, **k):
        pass

    def setUp(self):
        pass

    def destroy(self):
        pass


class Derivatives:
    def __init__(self, ps, grid, sources=None, u0=None, device=None):
        self.ps = ps
        self.grid = grid
        if sources is None:
            self.sources = [SpatialExpression(ps, grid, '0.0')
                            for _ in range(ps.nligands + 1)]
        else:
            self.sources = sources
        self.own_u0 = u0 is None
        self.u0 = grid.Vdmda.createGlobalVec() if u0 is None else u0
        self.dim = grid.dim
        self.sw = grid.stencil_width
        self.n_stencil_points = 1 + 2 * self.sw * self.dim
        if grid.dof != ps.nligands + 1:
            raise KSFDException('grid.dof must be nligands+1')
        import torch
        if device is None:
            device = torch.cuda.current_device() if torch.cuda.is_available() else 0
        self.ctx = core.Context(grid.dim, grid.globalSshape, grid.dof, device=device,
                                rank=grid.comm.rank, nranks=grid.comm.size)
        if grid.comm.size > 1:
            from . import parallel
            parallel.init_comm(self.ctx)
        self._phys_t = None
        self._phys_td = ps.physics_is_time_dependent()
        self._src_dev = None
        self._src_t = None
        self._have_src = any(not s.is_zero for s in self.sources)
        self._src_td = any(s.time_dependent for s in self.sources)
        self.set_time(ps.t0)

    # -- time-dependent data -----------------------------------------------
    def set_time(self, t):
        """Refresh the kernel parameter block and the source field for time t
        (no-ops when nothing depends on t)."""
        t = self.ps.t0 if t is None else t
        if self._phys_t is None or (self._phys_td and self._phys_t != t):
            self.ctx.set_physics(self.ps.physics(self.grid.spacing, t))
            self._phys_t = float(t)
        if self._have_src and (self._src_dev is None or
                               (self._src_td and self._src_t != t)):
            src = np.stack([np.asarray(s(t), dtype=float) for s in self.sources])
            self._src_dev = self.ctx.upload(src.reshape(-1, order='F'))
            self._src_t = float(t)
        return self._src_dev

    def source_device(self, t):
        return self.set_time(t)

    # -- reference API -----------------------------------------------------
    def groom(self, farr):
        """Get rid of negatives and NaNs in a HOST array (dof first)."""
        v = self.ps.values0
        rhomin, Umin = v['rhomin'], v['Umin']
        farr[0] = np.fmax(farr[0], rhomin)
        farr[1:] = np.fmax(farr[1:], Umin)
        return farr

    def groom_vec(self, u):
        ctx = self.ctx
        ctx.groom(u.device(ctx))
        u.mark_device_written()
        return u

    def dfdt(self, fvec, t=None, out=None):
        ctx = self.ctx
        src = self.set_time(t)
        if out is None:
            out = self.grid.Vdmda.createGlobalVec()
        ctx.residual(fvec.device(ctx), None, src, out.device(ctx))
        out.mark_device_written()
        return out

    def ifunction(self, u, udot, t=None, out=None):
        """F = udot - f(u, t), fused (reference implicitIF)."""
        ctx = self.ctx
        src = self.set_time(t)
        if out is None:
            out = self.grid.Vdmda.createGlobalVec()
        ctx.residual(u.device(ctx), udot.device(ctx), src, out.device(ctx))
        out.mark_device_written()
        return out

    def Jacobian(self, fvec, t=None, out=None, cache=True, shift=0.0):
        """Linearise at fvec.  Returns a JacobianOperator for
        A = shift*I - df/du (shift = 0 gives -J, i.e. out.mult(v) = -J v)."""
        self.set_time(t)
        kJ = out if isinstance(out, JacobianOperator) else JacobianOperator(self)
        kJ.setup(fvec.device(self.ctx), shift)
        return kJ

    def velocity(self, fvec, t=None, out=None):
        """ndarray (dim,)+Slshape of grad G (reference :1188-1209)."""
        ctx = self.ctx
        self.set_time(t)
        v = ctx.download(ctx.velocity(fvec.device(ctx)), ctx.dim)
        v = v.reshape((self.dim,) + self.grid.Slshape, order='F')
        if isinstance(out, np.ndarray):
            out[...] = v
            return out
        return np.ascontiguousarray(v)

    def velocity_max(self, fvec, t=None):
        self.set_time(t)
        return self.ctx.velocity_max(fvec.device(self.ctx))
