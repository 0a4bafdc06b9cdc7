// Spectral preconditioner: exact inverse (by FFT) of a CONSTANT-COEFFICIENT
// operator A0 close to the Jacobian A = shift*I - J.  The stiff part of A is
//   (A v)_rho = shift v_rho - rho * Lap(g_rho v_rho + sum_l g_l v_l) - (first-order terms)
//   (A v)_l   = (shift + gamma_l - D_l Lap) v_l - s_l v_rho
// with g_rho = dG/drho, g_l = dG/dU_l.  Two diagonal scalings remove most of the
// coefficient variation before anything is frozen: the rho row is divided by
// rho (left scaling) and the rho unknown becomes y_0 = g_rho v_rho (right
// scaling), which turns the row into
//   (shift / k) y_0 - Lap(y_0 + sum_l g_l y_l),      k = rho g_rho
// (k = s2 wherever the density cap is inactive: log-entropy diffusion is
// LINEAR diffusion of v_rho/rho), and the ligand rows into
//   (shift + gamma_l - D_l Lap) y_l - s_l (1/g_rho) y_0.
// A0 freezes k, g_l and 1/g_rho at their grid means.  On the periodic uniform
// grid it is block circulant: in Fourier space one small "arrow" matrix per
// wave number (Lap -> lambda(k), the symbol of the 4th-order stencil), solved
// by a Schur step on rho exactly like the point block.  M^-1 = S_R A0^-1 S_L.
// It captures the stiff cross-diffusion coupling that point-block Jacobi
// misses (at dt ~ 1 GMRES needs hundreds of steps with block Jacobi and a
// handful with this) and, through the scalings, stays effective when rho
// varies over two orders of magnitude (pattern phase of options84: half the
// Arnoldi steps of the unscaled frozen-coefficient inverse).
// The transforms are cuFFT calls (library code, dlopen'ed: one batched D2Z and
// one batched Z2D per application, strided straight from / into the plane-SoA
// vectors); scalings, symbol solve and coefficient means are our kernels.
// One rank only (a slab-decomposed FFT needs all-to-all transposes).
#pragma once
#include "blas1_kernels.cuh"
#include "device_common.cuh"

struct FftSym {                 // what the symbol kernel needs
    int dof, nlig, n0, n1, n2;  // real extents x, y, z (1 when unused); n0h = n0/2+1 complex
    // slab-distributed transform (several ranks): this rank holds the plane wave numbers
    // s in [s0, s0+nsq) (s = kx in 2-D, ky*n0h + kx in 3-D) for ALL NL wave numbers of the
    // last axis, layout [field][s_loc][k]
    int dist, s0, nsq, NL;
    double shift, c2[3];
    double s[KSFD_MAX_LIGANDS], gamma[KSFD_MAX_LIGANDS], D[KSFD_MAX_LIGANDS];
};

// grid means over the owned points of k = rho*dG/drho -> out[0], 1/(dG/drho) ->
// out[1], dG/dU_l -> out[2+l].  Two deterministic stages: grid (fields,
// FFT_MEAN_BLOCKS) partial sums, then one warp per field.
#define FFT_MEAN_BLOCKS 128
__global__ void k_fft_means_partial(Geom g, const double *__restrict__ coef_base,
                                    double *__restrict__ partial)
{
    // coef_base = plane 0 of the ghosted coefficient array, dof+2 fields per plane
    const int nf = g.dof + 2;
    const int which = blockIdx.x;                   // 0: rho*g_rho, 1: 1/g_rho, 2+l: g_l
    __shared__ double sm[32];
    double s = 0.0;
    for (long long p = blockIdx.y * (long long)blockDim.x + threadIdx.x; p < g.npts;
         p += (long long)gridDim.y * blockDim.x) {
        const long long k = p / g.plane_pts, pp = p - k * g.plane_pts;
        const double *cp = coef_base + (k * nf) * g.plane_pts + pp;
        if (which == 0)
            s += cp[0] * cp[2 * g.plane_pts];
        else if (which == 1)
            s += 1.0 / cp[2 * g.plane_pts];
        else
            s += cp[(which + 1) * g.plane_pts];
    }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int q = 0; q < (blockDim.x + 31) / 32; ++q) t += sm[q];
        partial[which * FFT_MEAN_BLOCKS + blockIdx.y] = t;
    }
}

// inv_count = 1 / (GLOBAL number of grid points): with several ranks the per-rank
// results are summed afterwards
__global__ void k_fft_means_final(double inv_count, const double *__restrict__ partial,
                                  double *__restrict__ out)
{
    double s = 0.0;
    for (int q = threadIdx.x; q < FFT_MEAN_BLOCKS; q += 32) s += partial[blockIdx.x * FFT_MEAN_BLOCKS + q];
    s = warp_sum(s);
    if (threadIdx.x == 0) out[blockIdx.x] = s * inv_count;
}

// left scaling S_L: out = in with the rho field divided by rho (in == out allowed)
__global__ void k_fft_prescale(Geom g, const double *__restrict__ coef_base, const double *in,
                               double *out, const int *__restrict__ skip)
{
    if (skip && *skip) return;
    const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (e >= g.npts * g.dof) return;
    const long long ps = g.plane_pts * g.dof;
    const long long k = e / ps, r = e - k * ps;
    double v = in[e];
    if (r < g.plane_pts) v /= coef_base[(k * (g.dof + 2)) * g.plane_pts + r];
    out[e] = v;
}

// right scaling S_R, in place: v_rho = y_0 / g_rho
__global__ void k_fft_postscale(Geom g, const double *__restrict__ coef_base, double *v,
                                const int *__restrict__ skip)
{
    if (skip && *skip) return;
    const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (e >= g.npts) return;
    const long long k = e / g.plane_pts, pp = e - k * g.plane_pts;
    v[k * g.dof * g.plane_pts + pp] /= coef_base[(k * (g.dof + 2) + 2) * g.plane_pts + pp];
}

// in place on the spectra: z^ = A0^(k)^-1 r^ / N   (N = number of grid points:
// cuFFT transforms are unnormalised).  spec layout: [field][k2][k1][k0 < n0/2+1]
// number of wave numbers this rank solves
__host__ __device__ __forceinline__ long long fft_symbol_count(const FftSym &S)
{
    return S.dist ? (long long)S.nsq * S.NL : (long long)(S.n0 / 2 + 1) * S.n1 * S.n2;
}

// one wave number e (host + device: tests/fft_share_check.cu runs it on the host
// against the numpy restatement)
__host__ __device__ __forceinline__ void fft_symbol_elem(const FftSym &S, const double *means,
                                                         double2 *spec, long long e)
{
    const int n0h = S.n0 / 2 + 1;
    const long long nk = fft_symbol_count(S);
    int k0, k1, k2;
    if (S.dist) {
        const int sg = S.s0 + (int)(e / S.NL), kl = (int)(e % S.NL);
        if (S.n2 > 1) {
            k0 = sg % n0h;
            k1 = sg / n0h;
            k2 = kl;
        } else {
            k0 = sg;
            k1 = kl;
            k2 = 0;
        }
    } else {
        k0 = (int)(e % n0h);
        k1 = (int)((e / n0h) % S.n1);
        k2 = (int)(e / ((long long)n0h * S.n1));
    }
    // symbol of the 4th-order Laplacian: sum_ax c2 (32 cos t - 2 cos 2t - 30)
    double lam = 0.0;
    {
        const double c = cospi(2.0 * k0 / S.n0);
        lam += S.c2[0] * (32.0 * c - 2.0 * (2.0 * c * c - 1.0) - 30.0);
    }
    if (S.n1 > 1) {
        const double c = cospi(2.0 * k1 / S.n1);
        lam += S.c2[1] * (32.0 * c - 2.0 * (2.0 * c * c - 1.0) - 30.0);
    }
    if (S.n2 > 1) {
        const double c = cospi(2.0 * k2 / S.n2);
        lam += S.c2[2] * (32.0 * c - 2.0 * (2.0 * c * c - 1.0) - 30.0);
    }
    const double kbar = means[0], cbar = means[1];   // mean rho*g_rho, mean 1/g_rho
    const double inv_n = 1.0 / ((double)S.n0 * S.n1 * S.n2);
    double2 r[KSFD_MAX_LIGANDS + 1];
    for (int c = 0; c < S.dof; ++c) r[c] = spec[(long long)c * nk + e];
    double schur = S.shift / kbar - lam;
    double2 t = r[0];
    double bd[KSFD_MAX_LIGANDS], invd[KSFD_MAX_LIGANDS];
    for (int l = 0; l < S.nlig; ++l) {
        invd[l] = 1.0 / (S.shift + S.gamma[l] - S.D[l] * lam);
        bd[l] = -means[2 + l] * lam * invd[l];              // b_l / d_l
        schur += bd[l] * S.s[l] * cbar;                     // - b_l c_l / d_l, c_l = -s_l*cbar
        t.x -= bd[l] * r[1 + l].x;
        t.y -= bd[l] * r[1 + l].y;
    }
    const double is = inv_n / schur;
    double2 z0;
    z0.x = t.x * is;
    z0.y = t.y * is;
    spec[e] = z0;
    for (int l = 0; l < S.nlig; ++l) {
        double2 z;
        // y_l = (r_l/N + s_l cbar y_0) / d_l
        z.x = (r[1 + l].x * inv_n + S.s[l] * cbar * z0.x) * invd[l];
        z.y = (r[1 + l].y * inv_n + S.s[l] * cbar * z0.y) * invd[l];
        spec[(long long)(1 + l) * nk + e] = z;
    }
}

__global__ void k_fft_symbol_solve(FftSym S, const double *__restrict__ means, double2 *spec,
                                   const int *__restrict__ skip)
{
    if (skip && *skip) return;
    const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (e >= fft_symbol_count(S)) return;
    fft_symbol_elem(S, means, spec, e);
}


// ---------------------------------------------------------------------------
// Slab-distributed transform (several ranks; experimental, KSFD_FFT_MULTI=1):
// local transforms over the plane axes, an all-to-all that gives every rank a
// share [s0_q, s0_q + nsq_q) of the plane wave numbers for ALL planes, a 1-D
// transform along the last axis, and the way back.  s0_q = floor(PS*q/P).
//   A  [(k_loc*dof + c)*PS + s]                     plane spectra of the own planes
//   B  [nloc*dof*s0_q + (k_loc*dof + c)*nsq_q + s-s0_q]   packed per destination q
//   R  [(k*dof + c)*nsq + s_loc]   (k global)       what the all-to-all delivers
//   T  [(c*nsq + s_loc)*NL + k]                     last axis contiguous
// ---------------------------------------------------------------------------
__host__ __device__ __forceinline__ long long fft_share_start(long long PS, int q, int P)
{
    return PS * q / P;
}
__host__ __device__ __forceinline__ int fft_share_owner(long long PS, long long s, int P)
{
    int q = (int)(((s + 1) * P - 1) / PS);
    // (floor rounding: correct the estimate by at most one)
    while (q > 0 && fft_share_start(PS, q, P) > s) --q;
    while (q + 1 < P && fft_share_start(PS, q + 1, P) <= s) ++q;
    return q;
}

// index maps (host + device: tests/fft_share_check.cu dumps them for the numpy emulation)
// A element e = (k_loc*dof + c)*PS + s  ->  its place in the packed buffer B
__host__ __device__ __forceinline__ long long fft_pack_index(int nloc, int dof, long long PS, int P,
                                                             long long e)
{
    const long long kc = e / PS, s = e - kc * PS;
    const int q = fft_share_owner(PS, s, P);
    const long long s0 = fft_share_start(PS, q, P), nsq = fft_share_start(PS, q + 1, P) - s0;
    return (long long)nloc * dof * s0 + kc * nsq + (s - s0);
}
// R element e = (k*dof + c)*nsq + s  ->  its place (c*nsq + s)*NL + k in T
__host__ __device__ __forceinline__ long long fft_transpose_index(int NL, int dof, long long nsq,
                                                                  long long e)
{
    const long long kc = e / nsq, s = e - kc * nsq;
    const long long k = kc / dof, c = kc - k * dof;
    return (c * nsq + s) * NL + k;
}

// to_packed = 1: B <- A;  0: A <- B   (same index map both ways)
__global__ void k_fft_pack(int nloc, int dof, long long PS, int P, int to_packed, double2 *A,
                           double2 *B, const int *__restrict__ skip)
{
    if (skip && *skip) return;
    const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (e >= (long long)nloc * dof * PS) return;
    const long long b = fft_pack_index(nloc, dof, PS, P, e);
    if (to_packed)
        B[b] = A[e];
    else
        A[e] = B[b];
}

// to_T = 1: T[(c*nsq + s)*NL + k] <- R[(k*dof + c)*nsq + s];  0: the way back
__global__ void k_fft_transpose(int NL, int dof, long long nsq, int to_T, double2 *R, double2 *T,
                                const int *__restrict__ skip)
{
    if (skip && *skip) return;
    const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (e >= (long long)NL * dof * nsq) return;
    const long long t = fft_transpose_index(NL, dof, nsq, e);
    if (to_T)
        T[t] = R[e];
    else
        R[e] = T[t];
}
