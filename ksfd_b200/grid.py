"""
Distributed periodic uniform Cartesian grid with the reference `KSFD.Grid`
interface (KSFD/ksfdgrid.py:60-483), without PETSc.

Differences in mechanism, not in meaning:
  * decomposition is 1-D slabs along the LAST axis (ranges equal PETSc DMDA's
    for `-da_processors_<other axes> 1`: lx[i] = M/P + (M%P > i));
  * `Sdmda` / `Vdmda` / `Cdmda` are small host objects that create `Vec`s and
    answer the DMDA queries the rest of the code uses (getRanges,
    createGlobalVec, createLocalVec, globalToLocal, getCoordinates); the
    actual ghost exchange on the hot path happens inside the CUDA library.

Layout facts kept verbatim: Fortran order, dof fastest, then x, y, z;
`Vashape = (dof,) + (n_local + 2*stencil_width)`; STAR stencil,
`stencil_width = 1 + order//2`.
"""
import numpy as np

from .core import dmda_ownership
from .vec import Vec


class Comm:
    """rank/size view of the process group (torch.distributed when
    initialised, else a single rank)."""

    def __init__(self, rank=None, size=None):
        if rank is None:
            try:
                import torch.distributed as dist
                if dist.is_available() and dist.is_initialized():
                    rank, size = dist.get_rank(), dist.get_world_size()
            except Exception:
                pass
        self.rank = int(rank or 0)
        self.size = int(size or 1)

    def Get_rank(self):
        return self.rank

    def Get_size(self):
        return self.size

    def tompi4py(self):
        return self

    def allreduce(self, x, op='sum'):
        if self.size == 1:
            return x
        import torch
        import torch.distributed as dist
        t = torch.tensor([float(x)], dtype=torch.float64)
        if dist.get_backend() == 'nccl':
            t = t.cuda()
        dist.all_reduce(t, op=dist.ReduceOp.MAX if op in ('max', 'MAX') else dist.ReduceOp.SUM)
        return float(t.item())

    def bcast(self, x, root=0):
        if self.size == 1:
            return x
        import torch.distributed as dist
        box = [x]
        dist.broadcast_object_list(box, src=root)
        return box[0]

    def allgather(self, x):
        if self.size == 1:
            return [x]
        import torch.distributed as dist
        out = [None] * self.size
        dist.all_gather_object(out, x)
        return out

    def Barrier(self):
        if self.size > 1:
            import torch.distributed as dist
            dist.barrier()


COMM_WORLD = None


def comm_world():
    global COMM_WORLD
    if COMM_WORLD is None or COMM_WORLD.size == 1:
        COMM_WORLD = Comm()
    return COMM_WORLD


class DMDA:
    """The slice of PETSc's DMDA interface the package uses."""

    def __init__(self, grid, dof):
        self.grid = grid
        self.dof = dof
        self._pool = []

    def getRanges(self):
        return self.grid.ranges

    def getSizes(self):
        return self.grid.globalSshape

    def setUniformCoordinates(self, **kw):
        pass

    def setFromOptions(self):
        pass

    def setUp(self):
        pass

    def destroy(self):
        pass

    def createGlobalVec(self):
        return Vec(self.grid, self.dof)

    def createLocalVec(self):
        return Vec(self.grid, self.dof, ghosted=True)

    def getLocalVec(self):
        return self._pool.pop() if self._pool else self.createLocalVec()

    def restoreLocalVec(self, v):
        self._pool.append(v)

    def globalToLocal(self, g, l):
        """Host ghost fill (periodic; DMDA STAR semantics leave the corner
        ghosts unreferenced — they are filled by the wrap as well)."""
        grid = self.grid
        sw = grid.stencil_width
        a = np.asarray(g.array).reshape((self.dof,) + grid.Slshape, order='F')
        pads = [(0, 0)] + [(sw, sw)] * grid.dim
        if grid.comm.size == 1:
            p = np.pad(a, pads, mode='wrap')
        else:
            from . import parallel
            lo, hi = parallel.host_halo_exchange(a, grid.nps[-1], sw)
            a2 = np.concatenate([lo, a, hi], axis=-1)
            pads[-1] = (0, 0)
            p = np.pad(a2, pads, mode='wrap')
        l.array[:] = p.reshape(-1, order='F')

    def getCoordinates(self):
        grid = self.grid
        v = Vec(grid, grid.dim)
        v.array[:] = grid.coordsNoGhosts.reshape(-1, order='F')
        return v


class Grid:
    def __init__(self, dim=1, width=1.0, height=1.0, depth=1.0, nx=8, ny=8, nz=8,
                 dof=2, order=3, stencil_width=None, stencil_type=None,
                 boundary_type=None, comm=None):
        if dim not in (1, 2, 3):
            raise ValueError('KSFD.Grid dimension must be 1, 2, or 3')
        self._dim = int(dim)
        self._width, self._height, self._depth = width, height, depth
        self._bounds = np.array([width, height, depth][:dim], dtype=float)
        self._nx, self._ny, self._nz = int(nx), int(ny), int(nz)
        self._nps = np.array([nx, ny, nz][:dim], dtype=int)
        self._spacing = self._bounds / self._nps
        self._dof = int(dof)
        self._order = order
        self._stencil_width = stencil_width if stencil_width else 1 + order // 2
        self._stencil_type = stencil_type or 'STAR'
        self._boundary_type = boundary_type or 'PERIODIC'
        self._comm = comm if comm is not None else comm_world()
        self._globalSshape = tuple(int(x) for x in self._nps)
        self._globalVshape = (self._dof,) + self._globalSshape
        self._globalCshape = (self._dim,) + self._globalSshape
        start, count = dmda_ownership(self._globalSshape[-1], self._comm.size)[self._comm.rank]
        if count < self._stencil_width:
            raise ValueError('each rank must own at least stencil_width planes')
        self._ranges = tuple((0, m) for m in self._globalSshape[:-1]) + ((start, start + count),)
        self._Slshape = tuple(r[1] - r[0] for r in self._ranges)
        self._Vlshape = (self._dof,) + self._Slshape
        self._Clshape = (self._dim,) + self._Slshape
        self._Sashape = tuple(int(s) + 2 * self._stencil_width for s in self._Slshape)
        self._Vashape = (self._dof,) + self._Sashape
        self._Cashape = (self._dim,) + self._Sashape
        self._Sdmda = DMDA(self, 1)
        self._Vdmda = DMDA(self, self._dof)
        self._Cdmda = DMDA(self, self._dim)

    # plain attribute access for everything the reference exposes
    dim = property(lambda s: s._dim)
    width = property(lambda s: s._width)
    height = property(lambda s: s._height)
    depth = property(lambda s: s._depth)
    bounds = property(lambda s: s._bounds)
    nx = property(lambda s: s._nx)
    ny = property(lambda s: s._ny)
    nz = property(lambda s: s._nz)
    nps = property(lambda s: s._nps)
    spacing = property(lambda s: s._spacing)
    dof = property(lambda s: s._dof)
    order = property(lambda s: s._order)
    stencil_width = property(lambda s: s._stencil_width)
    stencil_type = property(lambda s: s._stencil_type)
    boundary_type = property(lambda s: s._boundary_type)
    comm = property(lambda s: s._comm)
    Sdmda = property(lambda s: s._Sdmda)
    Vdmda = property(lambda s: s._Vdmda)
    Cdmda = property(lambda s: s._Cdmda)
    globalSshape = property(lambda s: s._globalSshape)
    globalVshape = property(lambda s: s._globalVshape)
    globalCshape = property(lambda s: s._globalCshape)
    ranges = property(lambda s: s._ranges)
    Slshape = property(lambda s: s._Slshape)
    Vlshape = property(lambda s: s._Vlshape)
    Clshape = property(lambda s: s._Clshape)
    Sashape = property(lambda s: s._Sashape)
    Vashape = property(lambda s: s._Vashape)
    Cashape = property(lambda s: s._Cashape)

    def _axis_coords(self, d, ghosts):
        lo, hi = self._ranges[d]
        sw = self._stencil_width if ghosts else 0
        idx = np.arange(lo - sw, hi + sw)
        if ghosts:
            idx = np.mod(idx, self._globalSshape[d])      # DMDA ghost coords wrap
        return idx * self._spacing[d]

    @property
    def coordsNoGhosts(self):
        """(dim,)+Slshape, F-contiguous, read-only: x_i = i*width/nx."""
        if not hasattr(self, '_cng'):
            mesh = np.meshgrid(*[self._axis_coords(d, False) for d in range(self._dim)],
                               indexing='ij')
            c = np.asfortranarray(np.stack(mesh, axis=0))
            c.flags['WRITEABLE'] = False
            self._cng = c
        return self._cng

    @property
    def coordsWithGhosts(self):
        if not hasattr(self, '_cwg'):
            mesh = np.meshgrid(*[self._axis_coords(d, True) for d in range(self._dim)],
                               indexing='ij')
            c = np.asfortranarray(np.stack(mesh, axis=0))
            c.flags['WRITEABLE'] = False
            self._cwg = c
        return self._cwg

    def make_dmda(self, dof=1):
        return DMDA(self, dof)

    def serial(self):
        """The same grid owned by a single rank (global shapes)."""
        if self._comm.size == 1:
            return self
        st = self.__getstate__()
        st['comm'] = Comm(0, 1)
        return Grid(**st)

    def stencil_slice(self, stencil, array, G=None, requireF=True):
        """Shifted interior view of a ghosted array (reference
        KSFD/ksfdgrid.py:413-434)."""
        if not isinstance(array, np.ndarray):
            array = array.array
        assert (not requireF) or array.flags['F_CONTIGUOUS']
        assert array.shape in (self.Sashape, self.Vashape)
        sw = self._stencil_width
        sl = [slice(stencil[i] + sw, stencil[i] + sw + self._Slshape[i])
              for i in range(self._dim)]
        out = array if stencil[-1] != -1 else G
        if out.ndim > self._dim:
            sl.insert(0, stencil[-1])
        return out[tuple(sl)]

    def cleanup(self):
        pass

    def __getstate__(self):
        return dict(dim=self.dim, width=self.width, height=self.height,
                    depth=self.depth, nx=self.nx, ny=self.ny, nz=self.nz,
                    dof=self.dof, order=self.order,
                    stencil_width=self.stencil_width,
                    stencil_type=self.stencil_type,
                    boundary_type=self.boundary_type)

    def __setstate__(self, state):
        self.__init__(**state)
