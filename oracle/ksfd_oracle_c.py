"""
ctypes front end of the oracle's C restatement (oracle/ksfd_oracle_c.c).

TEST INFRASTRUCTURE ONLY — same rule as oracle/ksfd_oracle.py: imported by tests/ and by
bench.py's cpu_baseline / --impl reference legs, never by the product.  Arrays are flat fp64 in
the reference layout (dof fastest, then x, y, z), exactly what the numpy oracle takes.
"""
import ctypes as C
import os

import numpy as np

from . import build_c
from . import ksfd_oracle as O

MAXLIG = 8


class _Phys(C.Structure):
    _fields_ = [('dim', C.c_int), ('n', C.c_int * 3), ('nlig', C.c_int), ('ngroups', C.c_int),
                ('witch', C.c_int), ('pad_', C.c_int),
                ('s2', C.c_double), ('rhomax', C.c_double), ('cushion', C.c_double),
                ('maxscale', C.c_double), ('rhomin', C.c_double), ('Umin', C.c_double),
                ('alpha', C.c_double * MAXLIG), ('beta', C.c_double * MAXLIG),
                ('lig_group', C.c_int * MAXLIG),
                ('lig_w', C.c_double * MAXLIG), ('lig_s', C.c_double * MAXLIG),
                ('lig_gamma', C.c_double * MAXLIG), ('lig_D', C.c_double * MAXLIG),
                ('w1', (C.c_double * 5) * 3), ('w2', (C.c_double * 5) * 3)]


class _Tableau(C.Structure):
    _fields_ = [('At', (C.c_double * 4) * 4), ('Gi', (C.c_double * 4) * 4),
                ('bt', C.c_double * 4), ('bet', C.c_double * 4), ('asum', C.c_double * 4),
                ('gamma', C.c_double)]


_lib = None
_dp = C.POINTER(C.c_double)


def lib():
    global _lib
    if _lib is None:
        path = build_c.build()
        L = C.CDLL(path)
        L.oc_create.restype = C.c_void_p
        L.oc_create.argtypes = [C.POINTER(_Phys)]
        L.oc_destroy.argtypes = [C.c_void_p]
        for name in ('oc_abi', 'oc_phys_size', 'oc_tableau_size', 'oc_threads'):
            getattr(L, name).restype = C.c_int
        if L.oc_phys_size() != C.sizeof(_Phys) or L.oc_tableau_size() != C.sizeof(_Tableau):
            raise RuntimeError('oracle C library: struct layout mismatch (rebuild oracle/_build)')
        L.oc_groom.argtypes = [C.c_void_p, _dp]
        L.oc_dfdt.argtypes = [C.c_void_p, _dp, _dp, _dp]
        L.oc_ifunction.argtypes = [C.c_void_p, _dp, _dp, _dp, _dp]
        L.oc_velocity_max.argtypes = [C.c_void_p, _dp, _dp]
        L.oc_velocity.argtypes = [C.c_void_p, _dp, _dp]
        L.oc_jvp_setup.argtypes = [C.c_void_p, _dp, C.c_double]
        L.oc_jvp_setup.restype = C.c_int
        L.oc_get_minv.argtypes = [C.c_void_p, _dp]
        L.oc_pc_apply.argtypes = [C.c_void_p, _dp, _dp]
        L.oc_jvp.argtypes = [C.c_void_p, _dp, _dp]
        L.oc_solve.argtypes = [C.c_void_p, _dp, _dp, C.c_double, C.c_double, C.c_int, C.c_int,
                               C.c_int, C.POINTER(C.c_int), _dp]
        L.oc_solve.restype = C.c_int
        L.oc_rosw_step.argtypes = [C.c_void_p, _dp, C.c_double, C.POINTER(_Tableau), C.c_double,
                                   C.c_double, C.c_int, C.c_int, C.c_int, _dp, _dp,
                                   C.POINTER(C.c_int)]
        L.oc_rosw_step.restype = C.c_int
        L.oc_beuler_step.argtypes = [C.c_void_p, _dp, C.c_double, C.c_double, C.c_double, C.c_int,
                                     C.c_int, C.c_int, _dp, C.POINTER(C.c_int)]
        L.oc_beuler_step.restype = C.c_int
        L.oc_get_stage.argtypes = [C.c_void_p, C.c_int, _dp]
        L.oc_ts_step.argtypes = [C.c_void_p, _dp, C.c_double, C.POINTER(_Tableau), C.c_double,
                                 C.c_int, _dp, C.POINTER(C.c_int)]
        L.oc_ts_step.restype = C.c_int
        _lib = L
    return _lib


def threads():
    return int(lib().oc_threads())


def _p(a):
    return a.ctypes.data_as(_dp)


def _vec(a, n):
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float64).reshape(-1, order='F'))
    if a.size != n:
        raise ValueError('expected %d values, got %d' % (n, a.size))
    return a


def tableau(tab=None):
    tab = tab or O.rosw_transformed()
    T = _Tableau()
    for i in range(4):
        for j in range(4):
            T.At[i][j] = float(tab['At'][i, j])
            T.Gi[i][j] = float(tab['GammaInv'][i, j])
        T.bt[i] = float(tab['bt'][i])
        T.bet[i] = float(tab['bembedt'][i])
        T.asum[i] = float(tab['ASum'][i])
    T.gamma = O.ROSW_GAMMA
    return T


class COracle:
    """One periodic grid + physics (an oracle `Physics`) evaluated by the C restatement."""

    KSP = {'auto': 0, 'richardson': 1, 'gmres': 2}

    def __init__(self, ph):
        self.ph = ph
        P = _Phys()
        P.dim = ph.dim
        for a in range(3):
            P.n[a] = ph.n[a] if a < ph.dim else 1
        groups = [g for g in ph.groups]
        if len(groups) > MAXLIG or ph.nlig > MAXLIG:
            raise ValueError('at most %d ligands / groups' % MAXLIG)
        P.nlig, P.ngroups = ph.nlig, len(groups)
        P.witch = 1 if ph.cap == 'witch' else 0
        P.s2, P.rhomax, P.cushion, P.maxscale = ph.s2, ph.rhomax, ph.cushion, ph.maxscale
        P.rhomin, P.Umin = ph.rhomin, ph.Umin
        l = 0
        for gi, (alpha, beta, ligs) in enumerate(groups):
            P.alpha[gi], P.beta[gi] = alpha, beta
            for (w, s, gam, D) in ligs:
                P.lig_group[l] = gi
                P.lig_w[l], P.lig_s[l], P.lig_gamma[l], P.lig_D[l] = w, s, gam, D
                l += 1
        for a in range(ph.dim):
            for k in range(5):
                P.w1[a][k] = float(ph.w1[a][k])
                P.w2[a][k] = float(ph.w2[a][k])
        self._P = P
        self.nv = ph.dof * ph.npts
        self.h = C.c_void_p(lib().oc_create(C.byref(P)))
        if not self.h:
            raise MemoryError('oc_create failed')
        self._tab = tableau()

    def close(self):
        if getattr(self, 'h', None):
            lib().oc_destroy(self.h)
            self.h = None

    __del__ = close

    # -- operators (flat in, flat out) --------------------------------------
    def groom(self, u):
        u = _vec(u, self.nv).copy()
        lib().oc_groom(self.h, _p(u))
        return u

    def dfdt(self, u, sources=None):
        u = _vec(u, self.nv)
        out = np.empty(self.nv)
        src = None if sources is None else _vec(sources, self.nv)
        lib().oc_dfdt(self.h, _p(u), None if src is None else _p(src), _p(out))
        return out

    def ifunction(self, u, udot, sources=None):
        u, ud = _vec(u, self.nv), _vec(udot, self.nv)
        out = np.empty(self.nv)
        src = None if sources is None else _vec(sources, self.nv)
        lib().oc_ifunction(self.h, _p(u), _p(ud), None if src is None else _p(src), _p(out))
        return out

    def velocity(self, u):
        u = _vec(u, self.nv)
        out = np.empty(self.ph.dim * self.ph.npts)
        lib().oc_velocity(self.h, _p(u), _p(out))
        return out

    def velocity_max(self, u):
        u = _vec(u, self.nv)
        vm = np.zeros(3)
        lib().oc_velocity_max(self.h, _p(u), _p(vm))
        return vm[:self.ph.dim]

    def jvp_setup(self, u_lin, shift):
        u = _vec(u_lin, self.nv)
        if lib().oc_jvp_setup(self.h, _p(u), float(shift)):
            raise RuntimeError('singular diagonal block')

    def jvp(self, v):
        v = _vec(v, self.nv)
        out = np.empty(self.nv)
        lib().oc_jvp(self.h, _p(v), _p(out))
        return out

    def pc_apply(self, r):
        r = _vec(r, self.nv)
        out = np.empty(self.nv)
        lib().oc_pc_apply(self.h, _p(r), _p(out))
        return out

    def minv(self):
        out = np.empty(self.ph.npts * self.ph.dof ** 2)
        lib().oc_get_minv(self.h, _p(out))
        return out.reshape(self.ph.npts, self.ph.dof, self.ph.dof)

    def solve(self, b, rtol=1e-8, atol=0.0, max_it=10000, restart=30, ksp_type='auto'):
        b = _vec(b, self.nv)
        x = np.empty(self.nv)
        info = (C.c_int * 2)()
        norms = np.zeros(2)
        rc = lib().oc_solve(self.h, _p(b), _p(x), rtol, atol, max_it, restart,
                            self.KSP[ksp_type], info, _p(norms))
        return x, dict(converged=rc == 0, its=int(info[0]), gmres=bool(info[1]),
                       bnorm=float(norms[0]), rnorm=float(norms[1]))

    def rosw_step(self, u, h, rtol=1e-8, atol=0.0, max_it=10000, restart=30, ksp_type='auto'):
        """-> (u_new, u_embedded, info); raises if a stage solve does not converge."""
        u = _vec(u, self.nv)
        un, ue = np.empty(self.nv), np.empty(self.nv)
        info = (C.c_int * 2)()
        rc = lib().oc_rosw_step(self.h, _p(u), float(h), C.byref(self._tab), rtol, atol, max_it,
                                restart, self.KSP[ksp_type], _p(un), _p(ue), info)
        if rc:
            raise RuntimeError('oc_rosw_step: stage %d did not converge' % (rc - 1))
        return un, ue, dict(its=int(info[0]), gmres_solves=int(info[1]))

    def beuler_step(self, u, h, rtol=1e-8, atol=0.0, max_it=10000, restart=30, ksp_type='auto'):
        """backward Euler, one Newton step (-snes_type ksponly) -> u_new"""
        u = _vec(u, self.nv)
        un = np.empty(self.nv)
        info = (C.c_int * 2)()
        if lib().oc_beuler_step(self.h, _p(u), float(h), rtol, atol, max_it, restart,
                                self.KSP[ksp_type], _p(un), info):
            raise RuntimeError('oc_beuler_step: the solve did not converge')
        return un

    def stage(self, j):
        out = np.empty(self.nv)
        lib().oc_get_stage(self.h, j, _p(out))
        return out

    def ts_step(self, u, h, rtol=1e-8, ksp_type='auto'):
        """clamp, ROSW step, CFL maxima — in place on the contiguous flat array u."""
        assert u.flags.c_contiguous and u.dtype == np.float64 and u.size == self.nv
        vm = np.zeros(3)
        info = (C.c_int * 2)()
        rc = lib().oc_ts_step(self.h, _p(u), float(h), C.byref(self._tab), rtol,
                              self.KSP[ksp_type], _p(vm), info)
        if rc:
            raise RuntimeError('oc_ts_step failed (%d)' % rc)
        return vm[:self.ph.dim], int(info[0])


def integrate(c, u0, t0, h, nsteps, groom_each_step=True, adapt=None, rtol=1e-12, ksp_type='auto'):
    """The reference's step loop (KSFDTS.solve, KSFD/ksfdts.py:202-228) without noise and monitors on
    the C restatement: clamp, ROSW step; adapt=None -> fixed step (-ts_adapt_type none),
    adapt=dict(atol, rtol[, clip, dt_min, dt_max]) -> TSAdapt basic exactly as
    ksfd_oracle.integrate (same `wnorm2` / `adapt_basic`).  `rtol` is the tolerance of the
    iterative stage solves.  Returns [(t, u)] after each accepted step and the number of
    rejected attempts."""
    u = np.ascontiguousarray(np.asarray(u0, dtype=np.float64).reshape(-1, order='F')).copy()
    t, k, rejected, out = float(t0), 0, 0, []
    while k < nsteps:
        if groom_each_step:
            u = c.groom(u)
        unew, uemb, _ = c.rosw_step(u, h, rtol=rtol, ksp_type=ksp_type)
        if adapt is None:
            u, t, k = unew, t + h, k + 1
            out.append((t, u.copy()))
            continue
        en = O.wnorm2(unew, uemb, adapt['atol'], adapt['rtol'])
        kw = {a: adapt[a] for a in ('clip', 'dt_min', 'dt_max') if a in adapt}
        ok, hn = O.adapt_basic(h, en, **kw)
        if ok:
            u, t, k = unew, t + h, k + 1
            out.append((t, u.copy()))
        else:
            rejected += 1
        h = hn
    return out, rejected
