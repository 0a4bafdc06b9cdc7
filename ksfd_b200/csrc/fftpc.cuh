// Spectral preconditioner: exact inverse (by FFT) of the CONSTANT-COEFFICIENT
// operator obtained by freezing the Jacobian coefficients at their grid means,
//   (A0 v)_rho = shift v_rho - rho0 * Lap(g_rho v_rho + sum_l g_l v_l)
//   (A0 v)_l   = (shift + gamma_l - D_l Lap) v_l - s_l v_rho
// On the periodic uniform grid A0 is block circulant: in Fourier space it is one
// small "arrow" matrix per wave number (Lap -> lambda(k), the symbol of the
// 4th-order stencil), solved by a Schur step on rho exactly like the point
// block.  It captures the stiff cross-diffusion coupling that point-block
// Jacobi misses: at dt ~ 1 (options84 after the start-up phase) GMRES needs
// hundreds of steps with block Jacobi and a handful with this.
// The transforms are cuFFT calls (library code, dlopen'ed: one batched D2Z and
// one batched Z2D per application, strided straight from / into the plane-SoA
// vectors); the symbol solve and the coefficient means are our kernels.
// One rank only (a slab-decomposed FFT needs all-to-all transposes).
#pragma once
#include "blas1_kernels.cuh"
#include "device_common.cuh"

struct FftSym {                 // what the symbol kernel needs
    int dof, nlig, n0, n1, n2;  // real extents x, y, z (1 when unused); n0h = n0/2+1 complex
    double shift, c2[3];
    double s[KSFD_MAX_LIGANDS], gamma[KSFD_MAX_LIGANDS], D[KSFD_MAX_LIGANDS];
};

// means of the coefficient fields rho (0), dG/drho (2), dG/dU_l (3+l) over the
// owned points -> out[0], out[1], out[2+l].  Two deterministic stages:
// grid (fields, FFT_MEAN_BLOCKS) partial sums, then one warp per field.
#define FFT_MEAN_BLOCKS 128
__global__ void k_fft_means_partial(Geom g, const double *__restrict__ coef_base,
                                    double *__restrict__ partial)
{
    // coef_base = plane 0 of the ghosted coefficient array, dof+2 fields per plane
    const int nf = g.dof + 2;
    const int which = blockIdx.x;                   // 0: rho, 1: g_rho, 2+l: g_l
    const int field = which == 0 ? 0 : which + 1;
    __shared__ double sm[32];
    double s = 0.0;
    for (long long p = blockIdx.y * (long long)blockDim.x + threadIdx.x; p < g.npts;
         p += (long long)gridDim.y * blockDim.x) {
        const long long k = p / g.plane_pts, pp = p - k * g.plane_pts;
        s += coef_base[(k * nf + field) * g.plane_pts + pp];
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

__global__ void k_fft_means_final(Geom g, const double *__restrict__ partial,
                                  double *__restrict__ out)
{
    double s = 0.0;
    for (int q = threadIdx.x; q < FFT_MEAN_BLOCKS; q += 32) s += partial[blockIdx.x * FFT_MEAN_BLOCKS + q];
    s = warp_sum(s);
    if (threadIdx.x == 0) out[blockIdx.x] = s / (double)g.npts;
}

// in place on the spectra: z^ = A0^(k)^-1 r^ / N   (N = number of grid points:
// cuFFT transforms are unnormalised).  spec layout: [field][k2][k1][k0 < n0/2+1]
__global__ void k_fft_symbol_solve(FftSym S, const double *__restrict__ means, double2 *spec,
                                   const int *__restrict__ skip)
{
    if (skip && *skip) return;
    const int n0h = S.n0 / 2 + 1;
    const long long nk = (long long)n0h * S.n1 * S.n2;
    const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (e >= nk) return;
    const int k0 = (int)(e % n0h);
    const int k1 = (int)((e / n0h) % S.n1);
    const int k2 = (int)(e / ((long long)n0h * S.n1));
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
    const double rho0 = means[0], grho = means[1];
    const double inv_n = 1.0 / ((double)S.n0 * S.n1 * S.n2);
    double2 r[KSFD_MAX_LIGANDS + 1];
    for (int c = 0; c < S.dof; ++c) r[c] = spec[(long long)c * nk + e];
    double schur = S.shift - rho0 * grho * lam;
    double2 t = r[0];
    double bd[KSFD_MAX_LIGANDS], invd[KSFD_MAX_LIGANDS];
    for (int l = 0; l < S.nlig; ++l) {
        invd[l] = 1.0 / (S.shift + S.gamma[l] - S.D[l] * lam);
        bd[l] = -rho0 * means[2 + l] * lam * invd[l];       // b_l / d_l
        schur += bd[l] * S.s[l];                            // - b_l c_l / d_l, c_l = -s_l
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
        // z_l = (r_l/N + s_l z_rho) / d_l
        z.x = (r[1 + l].x * inv_n + S.s[l] * z0.x) * invd[l];
        z.y = (r[1 + l].y * inv_n + S.s[l] * z0.y) * invd[l];
        spec[(long long)(1 + l) * nk + e] = z;
    }
}
