"""
CPU oracle: a plain numpy/scipy restatement of the reference's algorithm for
the implicit time-stepping hot path of leonavery/KSFD.

TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module, and only as the
checker / CPU baseline.  The product (ksfd_b200/) never imports it.

Pinning: this restatement is checked against golden vectors produced by the
reference's OWN code (KSFD.Derivatives.dfdt / Jacobian / velocity run
unmodified through oracle/refharness, see oracle/make_golden.py and
tests/test_oracle_vs_golden.py).  The time integrator (PETSc TSROSW
'ra34pw2', TSAdapt basic) is NOT in the reference tree: it is restated from
the published scheme (Rang & Angermann 2005) and pinned only by (a) the
Rosenbrock order conditions (tests/test_rosw_tableau.py) and (b) the
manufactured exact solution of options93nx128dt1 -> "parity unpinned" at the
PETSc boundary.

Layout everywhere: flat fp64, Fortran order, dof fastest then x, y, z
(reference KSFD/ksfdgrid.py:10-28).  `Physics.n` is the global point count per
axis; one process owns the whole periodic grid.
"""
import numpy as np

SW = 2          # stencil width for order 3: 1 + order//2 (ksfdgrid.py:152-155)


# --------------------------------------------------------------------------
# Finite-difference weights (reference KSFD/ksfdsym.py:391-436)
# --------------------------------------------------------------------------
def fd_weights(h, deriv):
    """
    Weights of the `deriv`-th derivative on points (-2h,-h,0,h,2h), produced
    the way the reference does: sympy finite_diff_weights on FLOAT points in
    the order (-2h,-h,h,2h,0) (ksfdsym.py:423-432), so the last-bit
    asymmetries of the reference coefficients are reproduced.
    Returns a float array ordered by offset -2,-1,0,+1,+2.
    """
    import sympy as sy
    # mimic np.zeros + xcoords assignment: xcoords = j*spacing (ksfdsym.py:367)
    pts = [sy.Float(float(j) * float(h)) for j in (-2, -1, 1, 2)] + [sy.Float(0.0)]
    w = sy.finite_diff_weights(deriv, pts, sy.S(0))[deriv][-1]
    w = [float(x) for x in w]
    # reorder (-2,-1,+1,+2,0) -> (-2,-1,0,+1,+2)
    return np.array([w[0], w[1], w[4], w[2], w[3]], dtype=float)


class Physics:
    """
    Plain-number description of one Keller-Segel problem at one time t.

    groups: list of (alpha, beta, [ (weight, s, gamma, D), ... ])
    cap: 'tophat' | 'witch'  (reference KSFD/ksfdsoln.py:150-158)
    """

    def __init__(self, dim, n, h, groups, s2, rhomax, cushion, maxscale,
                 cap='tophat', rhomin=1e-7, Umin=1e-7):
        self.dim = int(dim)
        self.n = tuple(int(x) for x in n)[:self.dim]
        self.h = tuple(float(x) for x in h)[:self.dim]
        self.groups = [(float(a), float(b), [tuple(float(x) for x in l)
                                              for l in ligs])
                       for a, b, ligs in groups]
        self.s2 = float(s2)
        self.rhomax = float(rhomax)
        self.cushion = float(cushion)
        self.maxscale = float(maxscale)
        self.cap = cap
        self.rhomin = float(rhomin)
        self.Umin = float(Umin)
        self.nlig = sum(len(g[2]) for g in self.groups)
        self.dof = self.nlig + 1
        self.w1 = [fd_weights(hh, 1) for hh in self.h]
        self.w2 = [fd_weights(hh, 2) for hh in self.h]

    # flat per-ligand views
    def ligands(self):
        out = []
        for gi, (a, b, ligs) in enumerate(self.groups):
            for (w, s, gam, D) in ligs:
                out.append(dict(group=gi, weight=w, s=s, gamma=gam, D=D))
        return out

    @property
    def npts(self):
        return int(np.prod(self.n))

    @property
    def Vshape(self):
        return (self.dof,) + self.n


# --------------------------------------------------------------------------
# ghost fill and clamp
# --------------------------------------------------------------------------
def ghost_fill(arr, dim, sw=SW):
    """DMDA globalToLocal on one periodic rank == np.pad(mode='wrap')
    (reference call sites ksfdsym.py:703-705, 919-920, 1203)."""
    return np.pad(arr, [(0, 0)] * (arr.ndim - dim) + [(sw, sw)] * dim,
                  mode='wrap')


def groom(farr, ph):
    """Clamp rho >= rhomin, U >= Umin, NaN -> min; in place
    (reference ksfdsym.py:888-900)."""
    farr[0] = np.maximum(farr[0], ph.rhomin)
    farr[0][np.isnan(farr[0])] = ph.rhomin
    farr[1:] = np.maximum(farr[1:], ph.Umin)
    farr[1:][np.isnan(farr[1:])] = ph.Umin
    return farr


# --------------------------------------------------------------------------
# pointwise free energy G and its partial derivatives
# --------------------------------------------------------------------------
def G_of(farr, ph):
    """
    G = V(U, rho) + s2*log(rho)  (reference ksfdsym.py:983-990),
    V = sum_g -beta_g*log(alpha_g + sum_l w_gl U_gl) + Vcap(rho)
    (ksfdligand.py:527-547, 720-746; ksfdsoln.py:147-161).
    farr: (dof,)+shape, already clamped.
    """
    rho = farr[0]
    G = ph.s2 * np.log(rho)
    l = 1
    for (alpha, beta, ligs) in ph.groups:
        if not ligs:
            continue
        sU = 0.0
        for (w, s, gam, D) in ligs:
            sU = sU + w * farr[l]
            l += 1
        G = G - beta * np.log(alpha + sU)
    th = np.tanh((rho - ph.rhomax) / ph.cushion)
    cap = ph.maxscale * ph.s2 * (th + 1.0)
    if ph.cap == 'witch':
        cap = cap * (rho / ph.rhomax)
    return G + cap


def dG_of(farr, ph):
    """Partials (dG/drho, [dG/dU_l]) at every point (chain rule through
    log/tanh — what the reference gets from sympy .diff on Gsubs,
    ksfdsym.py:1021-1033, 1094-1100)."""
    rho = farr[0]
    th = np.tanh((rho - ph.rhomax) / ph.cushion)
    sech2 = 1.0 - th * th
    c = ph.maxscale * ph.s2
    if ph.cap == 'witch':
        dcap = c * (sech2 / ph.cushion * (rho / ph.rhomax)
                    + (th + 1.0) / ph.rhomax)
    else:
        dcap = c * sech2 / ph.cushion
    g_rho = ph.s2 / rho + dcap
    g_U = []
    l = 1
    for (alpha, beta, ligs) in ph.groups:
        if not ligs:
            continue
        sU = 0.0
        for k, (w, s, gam, D) in enumerate(ligs):
            sU = sU + w * farr[l + k]
        for (w, s, gam, D) in ligs:
            g_U.append(-beta * w / (alpha + sU))
        l += len(ligs)
    return g_rho, g_U


# --------------------------------------------------------------------------
# stencil helpers on ghosted arrays
# --------------------------------------------------------------------------
def _shift(a, dim, axis, off, n, sw=SW):
    """View of ghosted scalar array a shifted by off along axis, interior
    size n (reference Grid.stencil_slice, ksfdgrid.py:413-434)."""
    sl = [slice(sw, sw + n[d]) for d in range(dim)]
    sl[axis] = slice(sw + off, sw + off + n[axis])
    return a[tuple(sl)]


def d1(a, ph, axis):
    w = ph.w1[axis]
    out = 0.0
    for k, off in enumerate((-2, -1, 0, 1, 2)):
        if w[k] != 0.0:
            out = out + w[k] * _shift(a, ph.dim, axis, off, ph.n)
    return out


def d2(a, ph, axis):
    w = ph.w2[axis]
    out = 0.0
    for k, off in enumerate((-2, -1, 0, 1, 2)):
        out = out + w[k] * _shift(a, ph.dim, axis, off, ph.n)
    return out


def center(a, ph):
    return _shift(a, ph.dim, 0, 0, ph.n)


# --------------------------------------------------------------------------
# residual  (reference Derivatives.dfdt, ksfdsym.py:902-940)
# --------------------------------------------------------------------------
def dfdt(u, ph, sources=None):
    """
    f(u): time derivative of the field vector.
    u: flat or (dof,)+n array (global, no ghosts).  Returns (dof,)+n.
    sources: optional list of dof arrays (shape n) added to each row
    (ksfdsym.py:930-936).
    """
    ua = np.asarray(u, dtype=float).reshape(ph.Vshape, order='F')
    farr = ghost_fill(ua, ph.dim)                     # :919-921
    return dfdt_ghosted(farr, ph, sources)


def dfdt_ghosted(farr, ph, sources=None):
    """f(u) on the interior of an already ghosted array (dof,)+(n+2*SW): the body
    of `dfdt` after the ghost exchange.  With ph.n = the extents of a sub-box and
    farr cut (with wrap) out of a larger periodic field this is what ONE RANK of
    the reference computes on its DMDA patch (ksfdsym.py:919-940) — the tests use
    it to check full-size CUDA results box by box."""
    farr = groom(np.array(farr, dtype=float), ph)     # :922 clamp ghosted COPY
    G = G_of(farr, ph)                                # :797-803 incl. ghosts
    out = np.empty(ph.Vshape)
    # f_rho = grad(rho).grad(G) + rho*lap(G)          # :531-571, :804
    acc = 0.0
    lap = 0.0
    for ax in range(ph.dim):
        acc = acc + d1(farr[0], ph, ax) * d1(G, ph, ax)
        lap = lap + d2(G, ph, ax)
    out[0] = acc + center(farr[0], ph) * lap
    # f_U = -gamma*U + s*rho + D*lap(U)               # :583-613
    for l, lig in enumerate(ph.ligands()):
        lapU = 0.0
        for ax in range(ph.dim):
            lapU = lapU + d2(farr[l + 1], ph, ax)
        out[l + 1] = (-lig['gamma'] * center(farr[l + 1], ph)
                      + lig['s'] * center(farr[0], ph) + lig['D'] * lapU)
    if sources is not None:
        for c in range(ph.dof):
            if sources[c] is not None:
                out[c] = out[c] + sources[c]
    return out


def ifunction(u, udot, ph, sources=None):
    """F = udot - f(u)  (reference implicitIF, ksfdts.py:563-596)."""
    return (np.asarray(udot, dtype=float).reshape(ph.Vshape, order='F')
            - dfdt(u, ph, sources))


# --------------------------------------------------------------------------
# velocity (reference Derivatives.velocity ksfdsym.py:1188-1209)
# --------------------------------------------------------------------------
def velocity(u, ph):
    ua = np.asarray(u, dtype=float).reshape(ph.Vshape, order='F')
    farr = groom(ghost_fill(ua, ph.dim), ph)
    G = G_of(farr, ph)
    return np.stack([d1(G, ph, ax) for ax in range(ph.dim)], axis=0)


def cfl_maxh(u, ph):
    """reference KSFDTS.CFL_step, ksfdts.py:302-319."""
    v = velocity(u, ph)
    hm = []
    for ax in range(ph.dim):
        vm = np.max(np.abs(v[ax]))
        hm.append(np.inf if vm == 0.0 else ph.h[ax] * SW / vm)
    return min(hm)


# --------------------------------------------------------------------------
# Jacobian  (reference Derivatives.Jacobian ksfdsym.py:814-886,
#            rhoJacobian_arrays :675-761, UJacobian_arrays :630-673,
#            insertion ksfdMat.pyx:280-325)
# --------------------------------------------------------------------------
def _point_index(ph):
    """flat point index (x fastest) of every grid point, shape n."""
    return np.arange(ph.npts).reshape(ph.n, order='F')


def jacobian(u, ph):
    """
    Exact derivative of the discrete f(u) w.r.t. the (clamped) unknowns,
    assembled as scipy CSR with the reference's row/column numbering
    idx = c + dof*(i + nx*(j + ny*k)) and periodic wrap of stencil columns.
    The clamp is treated as the identity, as in the reference (the symbolic
    derivative is taken w.r.t. the stencil symbols, ksfdsym.py:1094-1100).

    NOTE: in 3-D the reference pairs the rho-row values with wrongly ordered
    rows (np.meshgrid 'xy' indexing in cartesian_product, ksfdsym.py:81-87);
    this oracle uses the correct x-fastest ordering in every dimension, i.e.
    the exact derivative of dfdt.
    """
    import scipy.sparse as sp
    ua = np.asarray(u, dtype=float).reshape(ph.Vshape, order='F')
    farr = groom(ghost_fill(ua, ph.dim), ph)
    G = G_of(farr, ph)
    g_rho, g_U = dG_of(farr, ph)
    dof, dim, n = ph.dof, ph.dim, ph.n
    pid = _point_index(ph)
    rows, cols, vals = [], [], []

    def add(rdof, cdof, axis, off, val):
        # column point = row point shifted by off along axis (periodic)
        cp = np.roll(pid, -off, axis=axis) if off != 0 else pid
        rows.append((rdof + dof * pid).ravel(order='F'))
        cols.append((cdof + dof * cp).ravel(order='F'))
        v = val if np.ndim(val) else np.full(n, float(val))
        vals.append(np.asarray(v, dtype=float).ravel(order='F'))

    rho0 = center(farr[0], ph)
    for ax in range(dim):
        dG = d1(G, ph, ax)
        drho = d1(farr[0], ph, ax)
        w1, w2 = ph.w1[ax], ph.w2[ax]
        for k, off in enumerate((-2, -1, 0, 1, 2)):
            # d/d rho(p+off):  w1*dG  (direct)  + (drho*w1 + rho0*w2)*g_rho(p+off)
            grs = _shift(g_rho, dim, ax, off, n)
            add(0, 0, ax, off, w1[k] * dG + (drho * w1[k] + rho0 * w2[k]) * grs)
            for l in range(ph.nlig):
                gus = _shift(g_U[l], dim, ax, off, n)
                add(0, l + 1, ax, off, (drho * w1[k] + rho0 * w2[k]) * gus)
    lap = 0.0
    for ax in range(dim):
        lap = lap + d2(G, ph, ax)
    add(0, 0, 0, 0, lap)                               # d(rho0*lapG)/d rho0
    for l, lig in enumerate(ph.ligands()):
        add(l + 1, l + 1, 0, 0, -lig['gamma'])
        add(l + 1, 0, 0, 0, lig['s'])
        for ax in range(dim):
            for k, off in enumerate((-2, -1, 0, 1, 2)):
                add(l + 1, l + 1, ax, off, lig['D'] * ph.w2[ax][k])
    N = dof * ph.npts
    J = sp.coo_matrix((np.concatenate(vals),
                       (np.concatenate(rows), np.concatenate(cols))),
                      shape=(N, N)).tocsr()
    J.sum_duplicates()
    return J


def ijacobian(u, shift, ph):
    """shift*I - df/du  (reference implicitIJ, ksfdts.py:598-640)."""
    import scipy.sparse as sp
    J = jacobian(u, ph)
    return (sp.identity(J.shape[0], format='csr') * shift - J).tocsr()


def jvp(u_lin, v, shift, ph):
    """(shift*I - J(u_lin)) @ v, matrix-free restatement (same linearisation
    the CUDA J.v kernel uses; checked against `ijacobian` in the tests)."""
    ua = np.asarray(u_lin, dtype=float).reshape(ph.Vshape, order='F')
    va = np.asarray(v, dtype=float).reshape(ph.Vshape, order='F')
    return jvp_ghosted(ghost_fill(ua, ph.dim), ghost_fill(va, ph.dim), shift, ph)


def jvp_ghosted(farr, vg, shift, ph):
    """`jvp` on the interior of already ghosted u_lin and v (see dfdt_ghosted)."""
    farr = groom(np.array(farr, dtype=float), ph)
    G = G_of(farr, ph)
    g_rho, g_U = dG_of(farr, ph)
    dGv = g_rho * vg[0]
    for l in range(ph.nlig):
        dGv = dGv + g_U[l] * vg[l + 1]
    out = np.empty(ph.Vshape)
    acc = 0.0
    lapG = 0.0
    lapdG = 0.0
    for ax in range(ph.dim):
        acc = acc + d1(vg[0], ph, ax) * d1(G, ph, ax) \
                  + d1(farr[0], ph, ax) * d1(dGv, ph, ax)
        lapG = lapG + d2(G, ph, ax)
        lapdG = lapdG + d2(dGv, ph, ax)
    Jv0 = acc + center(vg[0], ph) * lapG + center(farr[0], ph) * lapdG
    out[0] = shift * center(vg[0], ph) - Jv0
    for l, lig in enumerate(ph.ligands()):
        lapV = 0.0
        for ax in range(ph.dim):
            lapV = lapV + d2(vg[l + 1], ph, ax)
        JvU = (-lig['gamma'] * center(vg[l + 1], ph)
               + lig['s'] * center(vg[0], ph) + lig['D'] * lapV)
        out[l + 1] = shift * center(vg[l + 1], ph) - JvU
    return out


def cut_box(arr, lo, size, sw=SW):
    """Ghosted sub-box of a periodic global array (dof,)+N: points lo-sw .. lo+size+sw-1
    of every axis, wrapped periodically (what DMDA globalToLocal delivers to the rank
    owning [lo, lo+size))."""
    out = arr
    for d, (l, n) in enumerate(zip(lo, size)):
        idx = np.arange(l - sw, l + n + sw) % arr.shape[d + 1]
        out = np.take(out, idx, axis=d + 1)
    return out


def block_diagonal(u_lin, shift, ph):
    """Per-point dof x dof diagonal block of shift*I - J: array (npts,dof,dof)
    (what the block-Jacobi preconditioner inverts)."""
    A = ijacobian(u_lin, shift, ph).tocsr()
    dof = ph.dof
    out = np.zeros((ph.npts, dof, dof))
    for r in range(dof):
        for c in range(dof):
            idx_r = r + dof * np.arange(ph.npts)
            idx_c = c + dof * np.arange(ph.npts)
            out[:, r, c] = np.asarray(A[idx_r, idx_c]).ravel()
    return out


# --------------------------------------------------------------------------
# Time integration: PETSc TSROSW 'ra34pw2' restated (NOT in the reference
# tree; PETSc src/ts/impls/rosw/rosw.c, Rang & Angermann 2005).
# --------------------------------------------------------------------------
ROSW_GAMMA = 4.3586652150845900e-01
RA34PW2_A = np.array([
    [0, 0, 0, 0],
    [8.7173304301691801e-01, 0, 0, 0],
    [8.4457060015369423e-01, -1.1299064236484185e-01, 0, 0],
    [0, 0, 1., 0]])
RA34PW2_GAMMA = np.array([
    [ROSW_GAMMA, 0, 0, 0],
    [-8.7173304301691801e-01, ROSW_GAMMA, 0, 0],
    [-9.0338057013044082e-01, 5.4180672388095326e-02, ROSW_GAMMA, 0],
    [2.4212380706095346e-01, -1.2232505839045147e+00,
     5.4526025533510214e-01, ROSW_GAMMA]])
RA34PW2_B = np.array([2.4212380706095346e-01, -1.2232505839045147e+00,
                      1.5452602553351020e+00, ROSW_GAMMA])
RA34PW2_BEMBED = np.array([3.7810903145819369e-01, -9.6042292212423178e-02,
                           5.0000000000000000e-01, 2.1793326075422950e-01])


def rosw_transformed(A=RA34PW2_A, Gm=RA34PW2_GAMMA, b=RA34PW2_B,
                     be=RA34PW2_BEMBED):
    """PETSc's transformed tableau: At = A*inv(Gamma), bt = b*inv(Gamma),
    GammaInv, ASum (TSRosWRegister)."""
    Gi = np.linalg.inv(Gm)
    return dict(At=A @ Gi, bt=b @ Gi, bembedt=be @ Gi, GammaInv=Gi,
                ASum=A.sum(axis=1), s=len(b))


def rosw_step(u, t, h, ph, sources_fn=None, linear_solve=None, tab=None):
    """
    One ROSW step as PETSc TSStep_RosW does it with -snes_type ksponly:
    per stage i:  Zstage = u + sum_j At[i,j] Y_j ; Zdot = (1/h) sum_j
    GammaInv[i,j] Y_j ; F = Zdot - f(t_i, Zstage)  (U = 0) ; Jacobian
    shift*I - df/du evaluated ONCE at stage 0 (lagged), shift = 1/(h*gamma);
    Y_i = -J^{-1} F.  Returns (u_new, u_embedded, Y).
    sources_fn(t) -> list of source arrays or None.
    """
    import scipy.sparse.linalg as spla
    tab = tab or rosw_transformed()
    s = tab['s']
    shape = ph.Vshape
    u = np.asarray(u, dtype=float).reshape(shape, order='F')
    Y = []
    lu = None
    for i in range(s):
        ti = t + h * tab['ASum'][i]
        Z = u.copy()
        Zdot = np.zeros(shape)
        for j in range(i):
            Z = Z + tab['At'][i, j] * Y[j]
            Zdot = Zdot + (tab['GammaInv'][i, j] / h) * Y[j]
        src = sources_fn(ti) if sources_fn else None
        F = Zdot - dfdt(Z, ph, src)
        if lu is None:
            shift = 1.0 / (h * ROSW_GAMMA)
            Jm = ijacobian(Z, shift, ph).tocsc()
            lu = (linear_solve(Jm) if linear_solve else
                  spla.splu(Jm, permc_spec='MMD_AT_PLUS_A').solve)
        y = lu(-F.ravel(order='F')).reshape(shape, order='F')
        Y.append(y)
    unew = u.copy()
    uemb = u.copy()
    for j in range(s):
        unew = unew + tab['bt'][j] * Y[j]
        uemb = uemb + tab['bembedt'][j] * Y[j]
    return unew, uemb, Y


def beuler_step(u, t, h, ph, sources_fn=None):
    """Backward Euler with -snes_type ksponly: one Newton step from u_n."""
    import scipy.sparse.linalg as spla
    shape = ph.Vshape
    u = np.asarray(u, dtype=float).reshape(shape, order='F')
    src = sources_fn(t + h) if sources_fn else None
    F = -dfdt(u, ph, src)
    Jm = ijacobian(u, 1.0 / h, ph).tocsc()
    y = spla.splu(Jm, permc_spec='MMD_AT_PLUS_A').solve(
        -F.ravel(order='F')).reshape(shape, order='F')
    return u + y


def wnorm2(u, y, atol, rtol):
    """PETSc TSErrorWeightedNorm, NORM_2: sqrt(mean(((u-y)/tol)^2)),
    tol = atol + rtol*max(|u|,|y|)."""
    tol = atol + rtol * np.maximum(np.abs(u), np.abs(y))
    return float(np.sqrt(np.mean(((u - y) / tol) ** 2)))


def adapt_basic(h, enorm, order=3, safety=0.9, reject_safety=0.5,
                clip=(0.1, 5.0), dt_min=1e-20, dt_max=1e50):
    """PETSc TSAdaptChoose_Basic restated.  Returns (accept, next_h)."""
    accept = not (enorm > 1.0)
    if not accept:
        safety = safety * reject_safety
    hfac = safety * np.inf if enorm == 0 else safety * enorm ** (-1.0 / order)
    hfac = min(max(hfac, clip[0]), clip[1])
    return accept, min(max(h * hfac, dt_min), dt_max)


def integrate(u0, t0, h, nsteps, ph, sources_fn=None, groom_each_step=True,
              adapt=None):
    """
    Reference step loop (KSFDTS.solve, ksfdts.py:202-228) without noise and
    monitors: groom in place, TS.step.  adapt=None -> fixed step
    (-ts_adapt_type none); adapt=dict(atol,rtol,...) -> TSAdapt basic.
    Returns list of (t, u) after each accepted step.
    """
    u = np.array(u0, dtype=float).reshape(ph.Vshape, order='F')
    t = t0
    out = []
    k = 0
    while k < nsteps:
        if groom_each_step:
            u = groom(u, ph)
        unew, uemb, _ = rosw_step(u, t, h, ph, sources_fn)
        if adapt is None:
            u, t, k = unew, t + h, k + 1
            out.append((t, u.copy()))
            continue
        en = wnorm2(unew, uemb, adapt['atol'], adapt['rtol'])
        kw = {a: adapt[a] for a in ('clip', 'dt_min', 'dt_max') if a in adapt}
        ok, hn = adapt_basic(h, en, **kw)
        if ok:
            u, t, k = unew, t + h, k + 1
            out.append((t, u.copy()))
        h = hn
    return out

