// Generic "direct" kernels: one thread per owned point, every stencil point
// re-read from global memory and G recomputed per stencil point (libm
// log/tanh).  Runtime dim (1,2,3) and nlig.  They are the 1-D path, the
// fallback for grids too small to tile, and an independent on-device
// cross-check of the marching kernels (tests/test_gpu_parity.py).  Not the
// performance path.  Also: layout converters, clamp, coefficient / block-Jacobi
// set-up (pointwise, run once per Jacobian).
#pragma once
#include "device_common.cuh"

struct PointIdx {
    int i, j, k;            // x, y index in plane (0 when unused), plane k
    long long pp;           // point offset inside the plane
};

__device__ __forceinline__ PointIdx decode_point(const Geom &g, long long p)
{
    PointIdx q;
    q.k = (int)(p / g.plane_pts);
    q.pp = p - (long long)q.k * g.plane_pts;
    if (g.dim == 3) {
        q.j = (int)(q.pp / g.n0);
        q.i = (int)(q.pp - (long long)q.j * g.n0);
    } else if (g.dim == 2) {
        q.j = 0;
        q.i = (int)q.pp;
    } else {
        q.i = q.j = 0;
    }
    return q;
}

// A point of a plane-SoA vector: field c is at p[c * fs]
struct PtRef {
    const double *p;
    long long fs;           // field stride = plane_pts
    __device__ __forceinline__ double operator[](int c) const { return p[c * fs]; }
};

// neighbour of q shifted by off along ax, in a vector with nf fields
__device__ __forceinline__ PtRef nbr(const Geom &g, const VecRef &v,
                                     const PointIdx &q, int ax, int off, int nf)
{
    int k = q.k;
    long long pp = q.pp;
    if (ax == g.dim - 1) {
        k += off;
    } else if (ax == 0) {
        pp += wrapi(q.i + off, g.n0) - q.i;
    } else {
        pp += (long long)(wrapi(q.j + off, g.n1) - q.j) * g.n0;
    }
    PtRef r;
    r.p = plane_ptr(v, k, g.nloc, g.plane_pts * nf) + pp;
    r.fs = g.plane_pts;
    return r;
}

// flat element index of (point p, field c) in an nf-field plane-SoA vector
__device__ __forceinline__ long long el(const Geom &g, const PointIdx &q, int c, int nf)
{
    return ((long long)q.k * nf + c) * g.plane_pts + q.pp;
}

__device__ __forceinline__ void load_clamped(const DevPhys &P, const PtRef &p,
                                             double &rho, double *U)
{
    rho = clampv(p[0], P.rhomin);
    for (int l = 0; l < P.nlig; ++l) U[l] = clampv(p[1 + l], P.Umin);
}

// ---- layout converters (the boundary) ---------------------------------------
// reference layout: ref[c + nf*p], p = pp + plane_pts*k  (dof fastest)
__global__ void k_to_internal(Geom g, int nf, const double *__restrict__ ref,
                              double *__restrict__ out)
{
    long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (e >= g.npts * nf) return;
    // e enumerates the INTERNAL index so that stores are coalesced
    long long k = e / (g.plane_pts * nf);
    long long r = e - k * g.plane_pts * nf;
    int c = (int)(r / g.plane_pts);
    long long pp = r - (long long)c * g.plane_pts;
    out[e] = ref[(k * g.plane_pts + pp) * nf + c];
}

__global__ void k_from_internal(Geom g, int nf, const double *__restrict__ in,
                                double *__restrict__ ref)
{
    long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (e >= g.npts * nf) return;
    long long k = e / (g.plane_pts * nf);
    long long r = e - k * g.plane_pts * nf;
    int c = (int)(r / g.plane_pts);
    long long pp = r - (long long)c * g.plane_pts;
    ref[(k * g.plane_pts + pp) * nf + c] = in[e];
}

// f_out = udot - (f(u)+src)  or  f(u)+src
__global__ void k_residual_naive(Geom g, DevPhys P, VecRef u,
                                 const double *__restrict__ udot,
                                 const double *__restrict__ src,
                                 double *__restrict__ out)
{
    long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (p >= g.npts) return;
    PointIdx q = decode_point(g, p);
    const int dof = g.dof;
    double rho0, U0[KSFD_MAX_LIGANDS];
    load_clamped(P, nbr(g, u, q, 0, 0, dof), rho0, U0);
    double G0 = G_point_rt(P, rho0, U0);
    double acc = 0.0, lap = 0.0, lapU[KSFD_MAX_LIGANDS];
    for (int l = 0; l < P.nlig; ++l) lapU[l] = 0.0;
    for (int ax = 0; ax < g.dim; ++ax) {
        double d1r = 0.0, d1G = 0.0, d2G = 0.0, d2U[KSFD_MAX_LIGANDS];
        for (int l = 0; l < P.nlig; ++l) d2U[l] = 0.0;
        for (int s = 0; s < 5; ++s) {
            double rn, Un[KSFD_MAX_LIGANDS], Gn;
            if (s == 2) {
                rn = rho0; Gn = G0;
                for (int l = 0; l < P.nlig; ++l) Un[l] = U0[l];
            } else {
                load_clamped(P, nbr(g, u, q, ax, s - 2, dof), rn, Un);
                Gn = G_point_rt(P, rn, Un);
            }
            d1r = fma(P.w1[ax][s], rn, d1r);
            d1G = fma(P.w1[ax][s], Gn, d1G);
            d2G = fma(P.w2[ax][s], Gn, d2G);
            for (int l = 0; l < P.nlig; ++l)
                d2U[l] = fma(P.w2[ax][s], Un[l], d2U[l]);
        }
        acc = fma(d1r, d1G, acc);
        lap += d2G;
        for (int l = 0; l < P.nlig; ++l) lapU[l] += d2U[l];
    }
    double f[KSFD_MAX_LIGANDS + 1];
    f[0] = fma(rho0, lap, acc);
    for (int l = 0; l < P.nlig; ++l)
        f[l + 1] = fma(P.D[l], lapU[l], fma(P.s[l], rho0, -P.gamma[l] * U0[l]));
    for (int c = 0; c < dof; ++c) {
        const long long e = el(g, q, c, dof);
        double v = f[c];
        if (src) v += src[e];
        out[e] = udot ? udot[e] - v : v;
    }
}

// vel(field d) = d/dx_d G (plane-SoA with dim fields); optional per-axis max
__global__ void k_velocity_naive(Geom g, DevPhys P, VecRef u,
                                 double *__restrict__ vel,
                                 double *__restrict__ vmax)
{
    long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    double vm[3] = {0.0, 0.0, 0.0};
    if (p < g.npts) {
        PointIdx q = decode_point(g, p);
        double rho0, U0[KSFD_MAX_LIGANDS];
        load_clamped(P, nbr(g, u, q, 0, 0, g.dof), rho0, U0);
        double G0 = G_point_rt(P, rho0, U0);
        for (int ax = 0; ax < g.dim; ++ax) {
            double d1G = 0.0;
            for (int s = 0; s < 5; ++s) {
                double Gn = G0;
                if (s != 2) {
                    double rn, Un[KSFD_MAX_LIGANDS];
                    load_clamped(P, nbr(g, u, q, ax, s - 2, g.dof), rn, Un);
                    Gn = G_point_rt(P, rn, Un);
                }
                d1G = fma(P.w1[ax][s], Gn, d1G);
            }
            if (vel) vel[el(g, q, ax, g.dim)] = d1G;
            vm[ax] = fabs(d1G);
        }
    }
    if (vmax) {
        for (int ax = 0; ax < g.dim; ++ax) {
            double m = warp_max(vm[ax]);
            if ((threadIdx.x & 31) == 0) atomic_max_nonneg(vmax + ax, m);
        }
    }
}

// clamp in place (KSFD/ksfdts.py:231-237)
__global__ void k_groom(Geom g, double rhomin, double Umin, double *__restrict__ u)
{
    long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (e >= g.npts * g.dof) return;
    int c = (int)((e / g.plane_pts) % g.dof);
    u[e] = clampv(u[e], c == 0 ? rhomin : Umin);
}

// coefficient field over planes -2 .. nloc+1 (ghosted along the last axis),
// plane-SoA with dof+2 fields: rho, G, dG/drho, dG/dU_l
__global__ void k_coef_setup(Geom g, DevPhys P, VecRef u,
                             double *__restrict__ coef)
{
    long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    long long tot = (long long)(g.nloc + 2 * KSFD_SW) * g.plane_pts;
    if (e >= tot) return;
    long long kg = e / g.plane_pts;             // ghosted plane index
    int k = (int)kg - KSFD_SW;
    long long pp = e - kg * g.plane_pts;
    PtRef p;
    p.p = plane_ptr(u, k, g.nloc, g.plane_pts * g.dof) + pp;
    p.fs = g.plane_pts;
    double rho, U[KSFD_MAX_LIGANDS], G, g_rho, g_U[KSFD_MAX_LIGANDS];
    load_clamped(P, p, rho, U);
    G_and_partials_rt(P, rho, U, G, g_rho, g_U);
    const int cs = g.dof + 2;
    double *c = coef + (kg * cs) * g.plane_pts + pp;
    c[0] = rho;
    c[g.plane_pts] = G;
    c[2 * g.plane_pts] = g_rho;
    for (int l = 0; l < P.nlig; ++l) c[(3 + l) * g.plane_pts] = g_U[l];
}

struct InvD {
    double v[KSFD_MAX_LIGANDS];          // 1/d_l, d_l = shift + gamma_l - D_l*w2c
};

// Point-block Jacobi of A = shift*I - J.  The diagonal block is
//   [ a   b_1 .. b_n ]      a   = shift - dJ_rho/drho0
//   [ c_1 d_1        ]      b_l = -dJ_rho/dU_l0      c_l = -s_l
//   [ c_n        d_n ]      d_l = shift + gamma_l - D_l*sum_ax w2c
// The preconditioner M takes b_l = -(rho0*w2c)*dG/dU_l (the exact b_l has an
// extra w1-centre term that is zero up to the last-bit asymmetry of the
// reference weights), so that only ONE field has to be stored per point:
//   pc = 1/(a - sum_l b_l c_l / d_l)
// and b_l/d_l is recomputed from the coefficient field (see pc_point_rt).
// If blocks != NULL the exact dense dof x dof blocks are written too (tests;
// row-major per point in natural point order).
__global__ void k_pc_setup(Geom g, DevPhys P, VecRef coef, double shift, InvD invd,
                           double *__restrict__ pc, double *__restrict__ blocks)
{
    long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (p >= g.npts) return;
    PointIdx q = decode_point(g, p);
    const int cs = g.dof + 2;
    PtRef c0 = nbr(g, coef, q, 0, 0, cs);
    double rho0 = c0[0];
    double fac = 0.0, lap = 0.0, dir = 0.0;
    for (int ax = 0; ax < g.dim; ++ax) {
        double d1r = 0.0, d1G = 0.0, d2G = 0.0;
        for (int s = 0; s < 5; ++s) {
            PtRef cn = (s == 2) ? c0 : nbr(g, coef, q, ax, s - 2, cs);
            d1r = fma(P.w1[ax][s], cn[0], d1r);
            d1G = fma(P.w1[ax][s], cn[1], d1G);
            d2G = fma(P.w2[ax][s], cn[1], d2G);
        }
        fac += d1r * P.w1[ax][2] + rho0 * P.w2[ax][2];
        dir += P.w1[ax][2] * d1G;
        lap += d2G;
    }
    const double Jrr = dir + fac * c0[2] + lap;
    const double a = shift - Jrr;
    const double facm = rho0 * P.w2c;            // M's version of fac
    double schur = a;
    for (int l = 0; l < P.nlig; ++l) {
        const double bd = -facm * c0[3 + l] * invd.v[l];
        schur -= bd * (-P.s[l]);
        if (blocks) {
            double *B = blocks + p * g.dof * g.dof;
            B[0 * g.dof + (1 + l)] = -fac * c0[3 + l];
            B[(1 + l) * g.dof + 0] = -P.s[l];
            for (int m = 0; m < P.nlig; ++m)
                B[(1 + l) * g.dof + (1 + m)] =
                    (m == l) ? shift + P.gamma[l] - P.D[l] * P.w2c : 0.0;
        }
    }
    pc[p] = 1.0 / schur;
    if (blocks) blocks[p * g.dof * g.dof] = a;
}

// z = M^{-1} r for one point (runtime nlig); same arithmetic as pc_solve<NLIG>
// of the marching kernels
__device__ __forceinline__ void pc_point_rt(const DevPhys &P, const InvD &invd, double rho,
                                            const double *gU, double pcinv,
                                            const double *r, double *z)
{
    const double fac = rho * P.w2c;
    double t = r[0];
    for (int l = 0; l < P.nlig; ++l) t = fma(fac * gU[l] * invd.v[l], r[1 + l], t);
    const double zr = pcinv * t;
    z[0] = zr;
    for (int l = 0; l < P.nlig; ++l) z[1 + l] = fma(P.s[l], zr, r[1 + l]) * invd.v[l];
}

__global__ void k_pc_apply(Geom g, DevPhys P, VecRef coef, InvD invd,
                           const double *__restrict__ pc,
                           const double *__restrict__ r, double *__restrict__ z)
{
    long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (p >= g.npts) return;
    PointIdx q = decode_point(g, p);
    PtRef c0 = nbr(g, coef, q, 0, 0, g.dof + 2);
    double rr[KSFD_MAX_LIGANDS + 1], zz[KSFD_MAX_LIGANDS + 1], gU[KSFD_MAX_LIGANDS];
    for (int c = 0; c < g.dof; ++c) rr[c] = r[el(g, q, c, g.dof)];
    for (int l = 0; l < P.nlig; ++l) gU[l] = c0[3 + l];
    pc_point_rt(P, invd, c0[0], gU, pc[p], rr, zz);
    for (int c = 0; c < g.dof; ++c) z[el(g, q, c, g.dof)] = zz[c];
}

// GMRES solution update at the end of a cycle (k, y on the device):
//   x += M^{-1} (sum_{i<k} y_i V_i)   or   x += sum_{i<k} y_i V_i
// gm_y = y, gmi_k = &k, gmi_skip = &no-update flag
// x_zero: x is known to be zero on entry (first cycle) and holds garbage: it is
// written, not read (x = ..., or x = 0 when the update is skipped)
template <int DOF>          // DOF = 0: runtime dof (any), else compile-time (registers)
__global__ void __launch_bounds__(256)
k_gm_update_x(Geom g, DevPhys P, VecRef coef, InvD invd, const double *__restrict__ pc,
              int precond, long long n, const double *__restrict__ V,
              const double *__restrict__ gm_y, const int *__restrict__ gmi_k,
              const int *__restrict__ gmi_skip, int x_zero, double *__restrict__ x)
{
    KSFD_PDL_ENTER();
    if (KSFD_FLAG(gmi_skip)) {
        if (x_zero)
            for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n;
                 e += (long long)gridDim.x * blockDim.x)
                x[e] = 0.0;
        return;
    }
    __shared__ double y[64];
    const int k = gmi_k ? KSFD_FLAG(gmi_k) : 1;     // no k / y: x += [M^-1] V_0
    for (int i = threadIdx.x; i < k; i += blockDim.x) y[i] = gm_y ? KSFD_FLAG(gm_y + i) : 1.0;
    __syncthreads();
    const int fs = (int)g.plane_pts;
    const int dof = DOF ? DOF : g.dof;
    constexpr int MD = DOF ? DOF : KSFD_MAX_LIGANDS + 1;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < (int)g.npts;
         p += gridDim.x * blockDim.x) {
        const int kp = p / fs, pp = p - kp * fs;
        const int e0 = kp * dof * fs + pp;          // field c of this point: e0 + c*fs
        double rr[MD], zz[MD];
#pragma unroll
        for (int c = 0; c < MD; ++c) rr[c] = 0.0;
        for (int i = 0; i < k; ++i) {
            const double *vi = V + (long long)i * n + e0;
            const double yi = y[i];
#pragma unroll
            for (int c = 0; c < MD; ++c)
                if (c < dof) rr[c] = fma(yi, __ldg(vi + c * fs), rr[c]);
        }
        if (precond) {
            // coef is ghosted by KSFD_SW planes along the last axis (coef.base = plane 0)
            const double *c0 = coef.base + ((long long)kp * (dof + 2)) * fs + pp;
            const double fac = __ldg(c0) * P.w2c;
            double t = rr[0];
#pragma unroll
            for (int l = 0; l < MD - 1; ++l)
                if (l < dof - 1) t = fma(fac * __ldg(c0 + (3 + l) * fs) * invd.v[l], rr[1 + l], t);
            const double zr = __ldg(pc + p) * t;
            zz[0] = zr;
#pragma unroll
            for (int l = 0; l < MD - 1; ++l)
                if (l < dof - 1) zz[1 + l] = fma(P.s[l], zr, rr[1 + l]) * invd.v[l];
        } else {
#pragma unroll
            for (int c = 0; c < MD; ++c) zz[c] = rr[c];
        }
#pragma unroll
        for (int c = 0; c < MD; ++c)
            if (c < dof) x[e0 + c * fs] = x_zero ? zz[c] : x[e0 + c * fs] + zz[c];
    }
}

// out = (shift*I - J(u_lin)) * v   (optionally v := M^{-1} v first)
__global__ void k_jvp_naive(Geom g, DevPhys P, VecRef coef, VecRef v, VecRef pc, InvD invd,
                            int precond, double shift, const int *__restrict__ skip,
                            double *__restrict__ out)
{
    if (skip && *skip) return;
    long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (p >= g.npts) return;
    PointIdx q = decode_point(g, p);
    const int dof = g.dof, cs = g.dof + 2;
    double acc = 0.0, lapG = 0.0, lapdG = 0.0, lapV[KSFD_MAX_LIGANDS];
    for (int l = 0; l < P.nlig; ++l) lapV[l] = 0.0;
    double v0[KSFD_MAX_LIGANDS + 1], rho0 = 0.0;
    for (int ax = 0; ax < g.dim; ++ax) {
        double d1r = 0, d1G = 0, d1v = 0, d1dG = 0, d2G = 0, d2dG = 0;
        double d2V[KSFD_MAX_LIGANDS];
        for (int l = 0; l < P.nlig; ++l) d2V[l] = 0.0;
        for (int s = 0; s < 5; ++s) {
            PtRef cn = nbr(g, coef, q, ax, s - 2, cs);
            PtRef vn = nbr(g, v, q, ax, s - 2, dof);
            double vv[KSFD_MAX_LIGANDS + 1];
            if (precond) {
                double rr[KSFD_MAX_LIGANDS + 1], gU[KSFD_MAX_LIGANDS];
                PtRef pn = nbr(g, pc, q, ax, s - 2, 1);
                for (int c = 0; c < dof; ++c) rr[c] = vn[c];
                for (int l = 0; l < P.nlig; ++l) gU[l] = cn[3 + l];
                pc_point_rt(P, invd, cn[0], gU, pn[0], rr, vv);
            } else {
                for (int c = 0; c < dof; ++c) vv[c] = vn[c];
            }
            double dG = cn[2] * vv[0];
            for (int l = 0; l < P.nlig; ++l) dG = fma(cn[3 + l], vv[1 + l], dG);
            if (s == 2) {
                rho0 = cn[0];
                for (int c = 0; c < dof; ++c) v0[c] = vv[c];
            }
            d1r = fma(P.w1[ax][s], cn[0], d1r);
            d1G = fma(P.w1[ax][s], cn[1], d1G);
            d2G = fma(P.w2[ax][s], cn[1], d2G);
            d1v = fma(P.w1[ax][s], vv[0], d1v);
            d1dG = fma(P.w1[ax][s], dG, d1dG);
            d2dG = fma(P.w2[ax][s], dG, d2dG);
            for (int l = 0; l < P.nlig; ++l)
                d2V[l] = fma(P.w2[ax][s], vv[1 + l], d2V[l]);
        }
        acc = fma(d1v, d1G, acc);
        acc = fma(d1r, d1dG, acc);
        lapG += d2G;
        lapdG += d2dG;
        for (int l = 0; l < P.nlig; ++l) lapV[l] += d2V[l];
    }
    double Jv0 = fma(rho0, lapdG, fma(v0[0], lapG, acc));
    out[el(g, q, 0, dof)] = fma(shift, v0[0], -Jv0);
    for (int l = 0; l < P.nlig; ++l) {
        double JvU =
            fma(P.D[l], lapV[l], fma(P.s[l], v0[0], -P.gamma[l] * v0[1 + l]));
        out[el(g, q, 1 + l, dof)] = fma(shift, v0[1 + l], -JvU);
    }
}
