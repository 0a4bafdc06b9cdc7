"""
Thin object layer over the C ABI: `Context` owns a `ksfd_ctx` and passes torch
CUDA tensors (device memory + streams only) to it.  All arithmetic happens in
libksfd_b200.so.
"""
import ctypes as C
import functools

import numpy as np

from . import _lib
from ._lib import KSFDError, KspOpts, KspResult, Physics, TsOpts, TsResult

SW = 2      # stencil width for order 3 (reference KSFD/ksfdgrid.py:152-155)


@functools.lru_cache(maxsize=64)
def fd_weights(h, deriv):
    """
    Finite-difference weights on offsets (-2,-1,0,+1,+2)*h produced exactly
    the way the reference produces them (sympy finite_diff_weights on float
    points ordered (-2h,-h,h,2h,0); KSFD/ksfdsym.py:391-436), so that the
    last-bit asymmetries of the reference coefficients are kept.
    """
    import sympy as sy
    pts = [sy.Float(float(j) * float(h)) for j in (-2, -1, 1, 2)] + [sy.Float(0.0)]
    w = [float(x) for x in sy.finite_diff_weights(deriv, pts, sy.S(0))[deriv][-1]]
    return (w[0], w[1], w[4], w[2], w[3])


def make_physics(dim, h, groups, s2, rhomax, cushion, maxscale, cap='tophat',
                 rhomin=1e-7, Umin=1e-7):
    """
    Fill a `ksfd_physics` struct.
    groups: [(alpha, beta, [(weight, s, gamma, D), ...]), ...]; groups without
    ligands contribute nothing to V (KSFD/ksfdligand.py:542-543) and are dropped.
    """
    p = Physics()
    groups = [g for g in groups if len(g[2]) > 0]
    nlig = sum(len(g[2]) for g in groups)
    if nlig < 1 or nlig > _lib.MAX_LIGANDS:
        raise KSFDError('number of ligands must be in [1, %d]' % _lib.MAX_LIGANDS)
    if len(groups) > _lib.MAX_GROUPS:
        raise KSFDError('too many ligand groups')
    p.ngroups = len(groups)
    p.nlig = nlig
    if cap not in ('tophat', 'witch'):
        raise KSFDError('cappotential must be tophat or witch')
    p.cap_type = 1 if cap == 'witch' else 0
    p.s2, p.rhomax, p.cushion, p.maxscale = (float(s2), float(rhomax),
                                             float(cushion), float(maxscale))
    p.rhomin, p.Umin = float(rhomin), float(Umin)
    l = 0
    for gi, (alpha, beta, ligs) in enumerate(groups):
        p.alpha[gi] = float(alpha)
        p.beta[gi] = float(beta)
        for (w, s, gam, D) in ligs:
            p.lig_group[l] = gi
            p.weight[l], p.s[l], p.gamma[l], p.D[l] = (float(w), float(s),
                                                       float(gam), float(D))
            l += 1
    for ax in range(dim):
        w1 = fd_weights(float(h[ax]), 1)
        w2 = fd_weights(float(h[ax]), 2)
        for k in range(5):
            p.w1[ax][k] = w1[k]
            p.w2[ax][k] = w2[k]
    return p


def dmda_ownership(M, P):
    """PETSc DMDA default split of M points over P ranks along one axis:
    lx[i] = M/P + (M%P > i).  Returns list of (start, count)."""
    out, s = [], 0
    for i in range(P):
        c = M // P + (1 if (M % P) > i else 0)
        out.append((s, c))
        s += c
    return out


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _stream():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class Context:
    """One rank's grid slab + operator state on one GPU."""

    def __init__(self, dim, n_global, dof, device=0, rank=0, nranks=1):
        import torch
        if not torch.cuda.is_available():
            raise KSFDError('ksfd_b200 needs a CUDA device (no CPU fallback)')
        self.lib = _lib.load()
        self.dim = int(dim)
        self.n = tuple(int(x) for x in n_global)[:self.dim]
        self.dof = int(dof)
        self.device = int(device)
        self.rank, self.nranks = int(rank), int(nranks)
        start, count = dmda_ownership(self.n[-1], self.nranks)[self.rank]
        self.last_start, self.last_count = start, count
        self.local_shape = self.n[:-1] + (count,)
        self.npts = int(np.prod(self.local_shape))
        n3 = (C.c_int64 * 3)(*(list(self.n) + [1] * (3 - self.dim)))
        h = C.c_void_p()
        torch.cuda.set_device(self.device)
        _lib.check(self.lib.ksfd_ctx_create(C.byref(h), self.dim, n3, start,
                                            count, self.dof, self.device))
        self.h = h
        self.tdev = torch.device('cuda', self.device)
        self._cb_keep = None

    # -- lifetime ----------------------------------------------------------
    def close(self):
        if getattr(self, 'h', None):
            self.lib.ksfd_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- helpers -----------------------------------------------------------
    def zeros(self):
        import torch
        return torch.zeros(self.npts * self.dof, dtype=torch.float64,
                           device=self.tdev)

    def empty(self):
        import torch
        return torch.empty(self.npts * self.dof, dtype=torch.float64,
                           device=self.tdev)

    def _chk(self, t, n=None):
        import torch
        n = self.npts * self.dof if n is None else n
        if (not isinstance(t, torch.Tensor) or t.dtype != torch.float64
                or not t.is_cuda or not t.is_contiguous() or t.numel() != n):
            raise KSFDError('expected a contiguous float64 CUDA tensor of %d '
                            'elements' % n)
        return t

    # -- layout boundary ---------------------------------------------------
    def to_internal(self, ref, nfields=None, out=None):
        """device tensor in the reference layout (dof fastest) -> internal."""
        import torch
        nf = self.dof if nfields is None else int(nfields)
        self._chk(ref, self.npts * nf)
        out = torch.empty_like(ref) if out is None else self._chk(out, self.npts * nf)
        _lib.check(self.lib.ksfd_to_internal(self.h, _ptr(ref), _ptr(out), nf,
                                             _stream()))
        return out

    def from_internal(self, t, nfields=None, out=None):
        """internal layout -> device tensor in the reference layout."""
        import torch
        nf = self.dof if nfields is None else int(nfields)
        self._chk(t, self.npts * nf)
        out = torch.empty_like(t) if out is None else self._chk(out, self.npts * nf)
        _lib.check(self.lib.ksfd_from_internal(self.h, _ptr(t), _ptr(out), nf,
                                               _stream()))
        return out

    def upload(self, host_array, nfields=None):
        """numpy array in the reference layout (flat, dof fastest) -> device
        vector in the internal layout."""
        import torch
        a = np.ascontiguousarray(np.asarray(host_array, dtype=np.float64).ravel())
        return self.to_internal(torch.from_numpy(a).to(self.tdev), nfields)

    def download(self, t, nfields=None):
        """device vector (internal layout) -> numpy array, reference layout."""
        return self.from_internal(t, nfields).cpu().numpy()

    def set_physics(self, phys):
        self.phys = phys
        _lib.check(self.lib.ksfd_set_physics(self.h, C.byref(phys)))

    def set_option(self, key, value):
        _lib.check(self.lib.ksfd_set_option(self.h, key.encode(), int(value)))

    def profile_fetch(self):
        """in-situ stencil kernel timing (set_option('profile', 1)): dict of launch
        counts and summed device milliseconds since the last fetch"""
        out = (C.c_double * 32)()
        _lib.check(self.lib.ksfd_profile_fetch(self.h, out, _stream()))
        names = ['jvp', 'residual', 'mdot', 'orth', 'first_vector', 'cycle_begin', 'sweep']
        d = {}
        for k, nm in enumerate(names):
            d[nm + '_launches'] = int(out[4 * k])
            d[nm + '_ms'] = out[4 * k + 1]
            d[nm + '_launches_all'] = int(out[4 * k + 2])
            d[nm + '_ms_all'] = out[4 * k + 3]
        return d

    def comm_init(self, nccl_path, unique_id):
        _lib.check(self.lib.ksfd_comm_init(self.h, (nccl_path or '').encode(),
                                           self.nranks, self.rank, unique_id))

    # -- operator ----------------------------------------------------------
    def groom(self, u):
        _lib.check(self.lib.ksfd_groom(self.h, _ptr(self._chk(u)), _stream()))
        return u

    def residual(self, u, udot=None, src=None, out=None):
        """out = udot - (f(u)+src)  or  f(u)+src when udot is None."""
        out = self.empty() if out is None else self._chk(out)
        _lib.check(self.lib.ksfd_residual(
            self.h, _ptr(self._chk(u)),
            _ptr(self._chk(udot) if udot is not None else None),
            _ptr(self._chk(src) if src is not None else None),
            _ptr(out), _stream()))
        return out

    def velocity_max(self, u):
        import torch
        # persistent device result + pinned host copy: no fill kernel, no pageable staging
        # (the kernel zeroes the result itself)
        if getattr(self, '_vm', None) is None:
            self._vm = torch.empty(3, dtype=torch.float64, device=self.tdev)
            self._vm_host = torch.empty(3, dtype=torch.float64).pin_memory()
        _lib.check(self.lib.ksfd_velocity_max(self.h, _ptr(self._chk(u)),
                                              _ptr(self._vm), _stream()))
        self._vm_host.copy_(self._vm, non_blocking=True)
        torch.cuda.current_stream(self.tdev).synchronize()
        v = self._vm_host.numpy()[:self.dim].copy()
        if self.nranks > 1:
            arr = (C.c_double * self.dim)(*v)
            _lib.check(self.lib.ksfd_allreduce_max(self.h, arr, self.dim, _stream()))
            v = np.array(list(arr))
        return v

    def velocity(self, u):
        import torch
        out = torch.empty(self.npts * self.dim, dtype=torch.float64,
                          device=self.tdev)
        _lib.check(self.lib.ksfd_velocity(self.h, _ptr(self._chk(u)), _ptr(out),
                                          _stream()))
        return out

    def jvp_setup(self, u_lin, shift):
        _lib.check(self.lib.ksfd_jvp_setup(self.h, _ptr(self._chk(u_lin)),
                                           float(shift), _stream()))

    def jvp(self, v, out=None, precond=False):
        out = self.empty() if out is None else self._chk(out)
        fn = self.lib.ksfd_jvp_precond if precond else self.lib.ksfd_jvp
        _lib.check(fn(self.h, _ptr(self._chk(v)), _ptr(out), _stream()))
        return out

    def pc_apply(self, r, out=None):
        out = self.empty() if out is None else self._chk(out)
        _lib.check(self.lib.ksfd_pc_apply(self.h, _ptr(self._chk(r)), _ptr(out),
                                          _stream()))
        return out

    def block_diagonal(self):
        import torch
        out = torch.empty(self.npts * self.dof * self.dof, dtype=torch.float64,
                          device=self.tdev)
        _lib.check(self.lib.ksfd_block_diagonal(self.h, _ptr(out), _stream()))
        return out.view(self.npts, self.dof, self.dof)

    # -- BLAS-1 ------------------------------------------------------------
    def mdot(self, vs, w):
        import torch
        out = torch.empty(len(vs), dtype=torch.float64, device=self.tdev)
        arr = (C.c_void_p * len(vs))(*[self._chk(v).data_ptr() for v in vs])
        _lib.check(self.lib.ksfd_mdot(self.h, len(vs), arr, _ptr(self._chk(w)),
                                      _ptr(out), _stream()))
        return out

    def maxpy(self, y, coefs, vs):
        arr = (C.c_void_p * len(vs))(*[self._chk(v).data_ptr() for v in vs])
        cf = (C.c_double * len(vs))(*[float(c) for c in coefs])
        _lib.check(self.lib.ksfd_maxpy(self.h, len(vs), cf, arr,
                                       _ptr(self._chk(y)), _stream()))
        return y

    def norm2(self, x):
        out = C.c_double()
        _lib.check(self.lib.ksfd_norm2(self.h, _ptr(self._chk(x)), C.byref(out), _stream()))
        return out.value

    def sum_dof0(self, u):
        out = C.c_double()
        _lib.check(self.lib.ksfd_sum_dof0(self.h, _ptr(self._chk(u)),
                                          C.byref(out), _stream()))
        return out.value

    def scale_dof0(self, u, f):
        _lib.check(self.lib.ksfd_scale_dof0(self.h, _ptr(self._chk(u)), float(f),
                                            _stream()))

    def mul_exp_dof0(self, u, z_host, sd):
        """u[dof 0] *= exp(sd*z) on the device; z_host: numpy sample, one value per owned
        point in Fortran (x fastest) order"""
        import torch
        z = torch.from_numpy(np.ascontiguousarray(
            np.asarray(z_host, dtype=np.float64).reshape(-1, order='F'))).to(self.tdev)
        self._chk(z, self.npts)
        _lib.check(self.lib.ksfd_mul_exp_dof0(self.h, _ptr(self._chk(u)), _ptr(z), float(sd),
                                              _stream()))

    # -- solvers -----------------------------------------------------------
    def gmres(self, rhs, x=None, rtol=1e-5, atol=1e-50, dtol=1e5, max_it=10000,
              restart=30, reorth=0, precond=1):
        x = self.empty() if x is None else self._chk(x)
        o = KspOpts(rtol, atol, dtol, max_it, restart, reorth, precond, 0, 0)
        r = KspResult()
        _lib.check(self.lib.ksfd_gmres(self.h, _ptr(self._chk(rhs)), _ptr(x),
                                       C.byref(o), C.byref(r), _stream()))
        return x, r

    def ksp_solve(self, rhs, x=None, ksp_type='auto', rtol=1e-5, atol=1e-50, dtol=1e5,
                  max_it=10000, restart=30, reorth=0, precond=1):
        """The linear solve of a stage system with the solver -ksp_type names: 'gmres',
        'richardson' (stationary sweeps fused into the stencil kernel) or 'auto' (sweeps
        while they contract fast enough, else GMRES from the iterate reached)."""
        x = self.empty() if x is None else self._chk(x)
        o = KspOpts(rtol, atol, dtol, max_it, restart, reorth, precond, KSP_TYPES[ksp_type], 0)
        r = KspResult()
        _lib.check(self.lib.ksfd_ksp_solve(self.h, _ptr(self._chk(rhs)), _ptr(x),
                                           C.byref(o), C.byref(r), _stream()))
        return x, r

    def sweep(self, r_in, x, r_out=None, first=False, norms=True):
        """One stationary sweep (csrc/sweep_op.cuh): r_out = r_in - A M^-1 r_in,
        x = (0 if first else x) + M^-1 r_in.  Returns (r_out, (||r_in||, ||r_out||))."""
        r_out = self.empty() if r_out is None else self._chk(r_out)
        nr = (C.c_double * 2)()
        _lib.check(self.lib.ksfd_sweep(self.h, _ptr(self._chk(r_in)), _ptr(self._chk(x)),
                                       _ptr(r_out), 1 if first else 0,
                                       nr if norms else None, _stream()))
        return r_out, (nr[0], nr[1])

    def ts_step(self, u, t, h, opts, src=None, time_cb=None):
        """One accepted step (or a failure report); u is advanced in place."""
        res = TsResult()
        raised = []
        if time_cb is not None:
            def guarded(tt, user):
                # ctypes swallows exceptions raised inside a callback: keep the first
                # one and re-raise it once the C call has returned
                if raised:
                    return
                try:
                    time_cb(tt)
                except BaseException as e:          # noqa: BLE001
                    raised.append(e)
            cb = _lib.TIME_CB(guarded)
        else:
            cb = C.cast(None, _lib.TIME_CB)
        self._cb_keep = cb
        rc = self.lib.ksfd_ts_step(
            self.h, _ptr(self._chk(u)), float(t), float(h), C.byref(opts),
            _ptr(self._chk(src) if src is not None else None), cb, None,
            C.byref(res), _stream())
        if raised:
            raise raised[0]
        _lib.check(rc)
        return res


KSP_TYPES = {'gmres': 0, 'richardson': 1, 'auto': 2}


def ts_options(ts_type='rosw', adapt='none', atol=1e-5, rtol=1e-5,
               clip=(0.1, 10.0), dt_min=1e-20, dt_max=1e50, safety=0.9,
               reject_safety=0.5, max_reject=10, ksp_rtol=1e-5, ksp_atol=1e-50,
               ksp_dtol=1e5, ksp_max_it=10000, restart=30, reorth=0, precond=1,
               ksp_type='auto', groom=False, velocity_max=False):
    """groom / velocity_max: do the step loop's clamp before and CFL maxima after the step
    inside the one C call (result.vmax), saving two host round trips per step."""
    o = TsOpts()
    o.ts_type = {'rosw': 0, 'beuler': 1}[ts_type]
    o.adapt = {'none': 0, 'basic': 1}[adapt]
    o.atol, o.rtol = float(atol), float(rtol)
    o.clip_lo, o.clip_hi = float(clip[0]), float(clip[1])
    o.dt_min, o.dt_max = float(dt_min), float(dt_max)
    o.safety, o.reject_safety = float(safety), float(reject_safety)
    o.max_reject = int(max_reject)
    o.flags = (1 if groom else 0) | (2 if velocity_max else 0)
    o.ksp = KspOpts(float(ksp_rtol), float(ksp_atol), float(ksp_dtol),
                    int(ksp_max_it), int(restart), int(reorth), int(precond),
                    KSP_TYPES[ksp_type], 0)
    return o
