"""
Random initial conditions and per-rank RNG streams (reference
KSFD/ksfdrandom.py).  Host-side, runs once per run.

Generator: numpy Generator seeded with SeedSequence(seed).spawn(size)[rank]
           (reference :13-60).
random_function: random values on a coarse periodic grid blended onto the
           fine grid with the cubic hat f(x) = 2x^3 - 3x^2 + 1 per axis
           (reference :108-220).  The reference does this with a KD-tree
           query per coarse vertex; the result is a separable (tensor
           product) interpolation, computed here with one small matrix per
           axis.  The reference flattens both grids in C order while its Vecs
           are in Fortran order (:180-183,:214-217); that index scramble is
           reproduced literally so that equal seeds give equal fields.
"""
import numpy as np
from numpy.random import SeedSequence, default_rng


class Generator:
    _rng = None
    _seeds = None
    _seed = None
    _global = None

    def __init__(self, seed=None, comm=None):
        if seed is None and type(self)._rng is not None:
            return
        rank = comm.rank if comm is not None else 0
        size = comm.size if comm is not None else 1
        seeds = SeedSequence(seed).spawn(size)
        type(self)._seeds = seeds
        type(self)._seed = seed
        type(self)._global = None
        type(self)._rng = default_rng(seeds[rank])

    @classmethod
    def global_rng(cls):
        """The stream a single-rank run of the same seed draws from, identical on
        every rank: used for fields every rank generates in full and then slices
        (initial condition), so that a run does not depend on the number of GPUs.
        With one rank this IS the rank stream."""
        if cls._rng is None:
            cls()
        if cls._seeds is not None and len(cls._seeds) == 1:
            return cls._rng
        if cls._global is None:
            cls._global = default_rng(SeedSequence(cls._seed).spawn(1)[0])
        return cls._global

    def __call__(self):
        return self.get_rng()

    @classmethod
    def get_rng(cls):
        if cls._rng is None:
            cls()
        return cls._rng


def _hat(x):
    return 2 * x ** 3 - 3 * x ** 2 + 1


def random_function(grid, randgrid=None, vals=None, mu=0.0, sigma=0.01, tol=1e-10,
                    seed=None):
    """Scalar random field on `grid`; returns a Vec (dof 1).

    Several ranks: every rank evaluates the GLOBAL field from the single-rank
    stream (Generator.global_rng) and keeps its own slab, so the field equals the
    one a single-rank run of the same seed produces.  (The reference draws
    per-rank streams on a block-partitioned coarse grid, KSFD/ksfdrandom.py:
    150-220, so its field depends on the rank count; there is nothing to match.)
    `vals` given on a distributed randgrid are gathered first."""
    if randgrid is None:
        randgrid = grid
    if grid.dim != randgrid.dim:
        raise ValueError('randgrid and grid must have the same dimension')
    if grid.comm.size != 1:
        g1, r1 = grid.serial(), randgrid.serial()
        v1 = None
        if vals is not None:
            if randgrid.comm.size == 1:
                v1 = vals
            else:
                parts = grid.comm.allgather(np.asarray(vals.array).reshape(
                    randgrid.Slshape, order='F'))
                v1 = r1.Sdmda.createGlobalVec()
                v1.array = np.concatenate(parts, axis=-1).reshape(-1, order='F')
        elif seed is not None:
            Generator(seed=seed, comm=grid.comm)
        if v1 is None:
            v1 = r1.Sdmda.createGlobalVec()
            v1.array = Generator.global_rng().normal(loc=mu, scale=sigma,
                                                     size=v1.array.shape)
        full = random_function(g1, randgrid=r1, vals=v1, tol=tol)
        lo, hi = grid.ranges[-1]
        out = grid.Sdmda.createGlobalVec()
        out.array = np.asarray(full.array).reshape(g1.Slshape, order='F')[
            ..., lo:hi].reshape(-1, order='F')
        return out
    if vals is None:
        vals = randgrid.Sdmda.createGlobalVec()
        vals.array = Generator(seed=seed, comm=grid.comm)().normal(
            loc=mu, scale=sigma, size=vals.array.shape)
    out = grid.Sdmda.createGlobalVec()
    if np.all(randgrid.nps == grid.nps) and np.all(randgrid.spacing == grid.spacing):
        out.array = vals.array
        return out
    dim = grid.dim
    sw = randgrid.stencil_width
    lvals = randgrid.Sdmda.createLocalVec()
    randgrid.Sdmda.globalToLocal(vals, lvals)
    # the reference indexes the (Fortran-ordered) local array with C-order
    # point numbers
    L = np.asarray(lvals.array).reshape(randgrid.Sashape, order='C')
    W = []
    for d in range(dim):
        hs = randgrid.spacing[d]
        # extended (non-wrapped) coarse coordinates, incl. ghost vertices
        # NOTE: same C-order scramble applies to the coordinates: vertex number
        # v = C-order index, coordinates taken from the C-order flattening too,
        # so per-axis coordinate tables are consistent with L's axes
        X = (np.arange(-sw, randgrid.nps[d] + sw)) * hs
        x = np.arange(grid.nps[d]) * grid.spacing[d]
        r = np.abs(x[:, None] - X[None, :]) / hs
        W.append(np.where(r < 1 - tol, _hat(r), 0.0))
    R = L
    for d in range(dim):
        R = np.moveaxis(np.tensordot(W[d], R, axes=([1], [d])), 0, d)
    out.array = R.reshape(-1, order='C')
    return out
