"""
`Vec`: the slice of PETSc's Vec interface the reference code uses, backed by a
device tensor in the library's internal layout plus a lazily synchronised host
mirror.

  vec.array        host-visible flat fp64 numpy array in the REFERENCE layout
                   (Fortran order, dof fastest; KSFD/ksfdtimeseries.py:485-488).
                   Reading it downloads (and converts) if the device copy is
                   newer; because the caller may modify it in place (the
                   reference does: ksfdts.py:231-284), the host copy is then
                   treated as authoritative and re-uploaded on next device use.
  vec.device(ctx)  torch CUDA tensor, internal plane-SoA layout.
"""
import numpy as np


class Vec:
    def __init__(self, grid, dof, ghosted=False):
        self.grid = grid
        self.dof = int(dof)
        self.ghosted = bool(ghosted)
        shape = grid.Sashape if ghosted else grid.Slshape
        self.size = self.dof * int(np.prod(shape))
        self._host = None
        self._dev = None
        self._ctx = None
        self._host_new = False      # host copy holds changes the device lacks
        self._dev_new = False       # device copy holds changes the host lacks

    # -- host side ---------------------------------------------------------
    def _ensure_host(self):
        if self._host is None:
            self._host = np.zeros(self.size, dtype=np.float64)
        if self._dev_new:
            self._host[:] = self._ctx.download(self._dev, self.dof)
            self._dev_new = False
        return self._host

    @property
    def array(self):
        h = self._ensure_host()
        self._host_new = True       # caller may write through the view
        return h

    @array.setter
    def array(self, values):
        h = self._ensure_host()
        h[:] = np.asarray(values, dtype=np.float64).reshape(-1, order='F')
        self._host_new = True
        self._dev_new = False

    @property
    def array_r(self):
        """read-only snapshot access (does not mark the host copy dirty)"""
        h = self._ensure_host()
        v = h.view()
        v.flags['WRITEABLE'] = False
        return v

    # -- device side -------------------------------------------------------
    def device(self, ctx):
        """device tensor (internal layout), uploading pending host changes."""
        if self.ghosted:
            raise ValueError('local (ghosted) Vecs live on the host only')
        if self._dev is None or self._ctx is not ctx:
            self._ctx = ctx
            if self._host is None:
                self._dev = ctx.zeros() if self.dof == ctx.dof else \
                    ctx.upload(np.zeros(self.size), self.dof)
                self._host_new = False
            else:
                if self._dev_new:
                    self._ensure_host()
                self._dev = ctx.upload(self._host, self.dof)
                self._host_new = False
        elif self._host_new:
            import torch
            ref = torch.from_numpy(self._host).to(ctx.tdev)
            ctx.to_internal(ref, self.dof, out=self._dev)
            self._host_new = False
        return self._dev

    def mark_device_written(self):
        self._dev_new = True
        self._host_new = False

    # -- PETSc-style API ---------------------------------------------------
    def assemble(self):
        pass

    assemblyBegin = assemblyEnd = setUp = assemble

    def destroy(self):
        self._dev = None
        self._host = None

    def getSize(self):
        return self.size

    def duplicate(self):
        return Vec(self.grid, self.dof, self.ghosted)

    def zeroEntries(self):
        if self._dev is not None and not self._host_new:
            self._dev.zero_()
            self.mark_device_written()
        else:
            self._ensure_host()[:] = 0.0
            self._host_new = True

    def copy(self, dst=None):
        if dst is None:
            dst = self.duplicate()
        if self._dev is not None and not self._host_new and self._ctx is not None:
            d = dst.device(self._ctx)
            d.copy_(self._dev)
            dst.mark_device_written()
        else:
            dst.array = self._ensure_host()
        return dst

    def _binary_dev(self, other):
        ctx = self._ctx or other._ctx
        if ctx is None:
            return None
        return self.device(ctx), other.device(ctx)

    def aypx(self, alpha, x):
        """self = alpha*self + x"""
        pair = self._binary_dev(x)
        if pair is None:
            h = self.array
            h *= alpha
            h += x._ensure_host()
        else:
            a, b = pair
            a.mul_(alpha).add_(b)
            self.mark_device_written()

    def axpy(self, alpha, x):
        """self += alpha*x"""
        pair = self._binary_dev(x)
        if pair is None:
            self.array[:] += alpha * x._ensure_host()
        else:
            a, b = pair
            a.add_(b, alpha=alpha)
            self.mark_device_written()

    def scale(self, alpha):
        if self._dev is not None and not self._host_new:
            self._dev.mul_(alpha)
            self.mark_device_written()
        else:
            self.array[:] *= alpha

    def norm(self):
        return float(np.linalg.norm(self.array_r))

    def max(self):
        a = self.array_r
        i = int(np.argmax(a))
        return i, float(a[i])

    def min(self):
        a = self.array_r
        i = int(np.argmin(a))
        return i, float(a[i])
