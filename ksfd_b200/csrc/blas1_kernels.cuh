// Fused BLAS-1 kernels for the device-resident Krylov solver and the ROSW
// stage combinations (replace PETSc VecMDot / VecMAXPY / VecNorm /
// TSErrorWeightedNorm).  All reductions are two-stage and deterministic:
// per-block partial sums (warp shuffle + shared memory), then one small
// kernel that sums the partials in a fixed order.
#pragma once
#include "device_common.cuh"

#define KSFD_RED_BLOCKS 592          // 4 CTAs per SM on 148 SMs
#define KSFD_RED_THREADS 256
#define KSFD_MAXV 8                  // vectors fused per launch

struct VecList {
    const double *v[KSFD_MAXV];
};
struct CoefList {
    double c[KSFD_MAXV];
};

template <int NV>
__device__ __forceinline__ void block_reduce_store(double (&acc)[NV],
                                                   double *partial, int stride)
{
    __shared__ double sm[KSFD_RED_THREADS / 32][KSFD_MAXV + 1];
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        double s = warp_sum(acc[i]);
        if (l == 0) sm[w][i] = s;
    }
    __syncthreads();
    if (threadIdx.x < NV) {
        double s = 0.0;
        for (int q = 0; q < KSFD_RED_THREADS / 32; ++q) s += sm[q][threadIdx.x];
        partial[(size_t)threadIdx.x * stride + blockIdx.x] = s;
    }
}

// partial[i][block] = sum over this block's elements of vs[i]*w   (i < NV)
template <int NV>
__global__ void __launch_bounds__(KSFD_RED_THREADS)
k_mdot(long long n, VecList vs, const double *__restrict__ w,
       double *__restrict__ partial)
{
    double acc[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) acc[i] = 0.0;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n;
         e += (long long)gridDim.x * blockDim.x) {
        const double wv = w[e];
#pragma unroll
        for (int i = 0; i < NV; ++i) acc[i] = fma(__ldg(vs.v[i] + e), wv, acc[i]);
    }
    block_reduce_store<NV>(acc, partial, KSFD_RED_BLOCKS);
}

// out[i] (op)= sum_b partial[i][b];  mode 0: store, 1: add, 2: store sqrt
__global__ void k_reduce_partials(int nv, int nblocks,
                                  const double *__restrict__ partial,
                                  double *__restrict__ out, int mode)
{
    const int i = blockIdx.x;
    if (i >= nv) return;
    __shared__ double sm[32];
    double s = 0.0;
    for (int b = threadIdx.x; b < nblocks; b += blockDim.x)
        s += partial[(size_t)i * KSFD_RED_BLOCKS + b];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int q = 0; q < (blockDim.x + 31) / 32; ++q) t += sm[q];
        if (mode == 1) out[i] += t;
        else if (mode == 2) out[i] = sqrt(t);
        else out[i] = t;
    }
}

// y = ybase_scale*y + sum_i c[i]*x_i ; coefficients by value (host known)
template <int NV>
__global__ void k_maxpy_host(long long n, CoefList c, VecList xs, double yscale,
                             double *__restrict__ y)
{
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n;
         e += (long long)gridDim.x * blockDim.x) {
        double s = (yscale == 0.0) ? 0.0 : yscale * y[e];
#pragma unroll
        for (int i = 0; i < NV; ++i) s = fma(c.c[i], __ldg(xs.v[i] + e), s);
        y[e] = s;
    }
}

// w -= sum_i h[i]*V_i  (h on the device), and partial sums of w_new^2
template <int NV>
__global__ void __launch_bounds__(KSFD_RED_THREADS)
k_orth_update(long long n, VecList vs, const double *__restrict__ h,
              double *__restrict__ w, double *__restrict__ partial, int want_norm)
{
    double hh[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) hh[i] = h[i];
    double acc[1] = {0.0};
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n;
         e += (long long)gridDim.x * blockDim.x) {
        double s = w[e];
#pragma unroll
        for (int i = 0; i < NV; ++i) s = fma(-hh[i], __ldg(vs.v[i] + e), s);
        w[e] = s;
        acc[0] = fma(s, s, acc[0]);
    }
    if (want_norm) block_reduce_store<1>(acc, partial, KSFD_RED_BLOCKS);
}

// Gram-Schmidt bookkeeping of one Arnoldi step, one block:
//   h[i] = sum_b partial[i][b]  for i < nv (the last one is <w,w>);
//   then with ALL h[0..j] (earlier batches already reduced into h):
//   hn2 = <w,w> - sum_i h[i]^2 ;  h[j+1] = sqrt(hn2) ;  aux[0] = 1/h[j+1]
//   aux[1] = 1 if the subtraction cancelled too much (caller recomputes the
//   norm explicitly), else 0.
// nv_batch partial rows are reduced into h[off .. off+nv_batch); j+1 = index of
// the <w,w> entry.
__global__ void k_gs_finalize(int nv_batch, int off, int jp1, int nblocks,
                              const double *__restrict__ partial,
                              double *__restrict__ h, double *__restrict__ aux,
                              double thresh)
{
    __shared__ double sums[KSFD_MAXV + 1];
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    for (int i = w; i < nv_batch; i += blockDim.x >> 5) {
        double s = 0.0;
        for (int b = l; b < nblocks; b += 32) s += partial[(size_t)i * KSFD_RED_BLOCKS + b];
        s = warp_sum(s);
        if (l == 0) sums[i] = s;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 0; i < nv_batch; ++i) h[off + i] = sums[i];
        const double ww = h[jp1];
        double ss = 0.0;
        for (int i = 0; i < jp1; ++i) ss = fma(h[i], h[i], ss);
        const double hn2 = ww - ss;
        const bool bad = !(hn2 > thresh * ww);
        const double hn = bad ? 1.0 : sqrt(hn2);
        h[jp1] = bad ? 0.0 : hn;
        aux[0] = 1.0 / hn;
        aux[1] = bad ? 1.0 : 0.0;
    }
}

// same finalisation when the sums are already in h (multi-rank: after the
// all-reduce of the partial dot products)
__global__ void k_gs_finalize_only(int jp1, double *__restrict__ h,
                                   double *__restrict__ aux, double thresh)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        const double ww = h[jp1];
        double ss = 0.0;
        for (int i = 0; i < jp1; ++i) ss = fma(h[i], h[i], ss);
        const double hn2 = ww - ss;
        const bool bad = !(hn2 > thresh * ww);
        const double hn = bad ? 1.0 : sqrt(hn2);
        h[jp1] = bad ? 0.0 : hn;
        aux[0] = 1.0 / hn;
        aux[1] = bad ? 1.0 : 0.0;
    }
}

// w = (w - sum_i h[i]*V_i) * scale,  scale = *inv (device) or 1 if inv == NULL
template <int NV>
__global__ void __launch_bounds__(KSFD_RED_THREADS)
k_orth_scale(long long n, VecList vs, const double *__restrict__ h,
             const double *__restrict__ inv, double *__restrict__ w)
{
    double hh[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) hh[i] = h[i];
    const double sc = inv ? inv[0] : 1.0;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n;
         e += (long long)gridDim.x * blockDim.x) {
        double s = w[e];
#pragma unroll
        for (int i = 0; i < NV; ++i) s = fma(-hh[i], __ldg(vs.v[i] + e), s);
        w[e] = s * sc;
    }
}

// y = x * (sign / *norm)     (Krylov vector normalisation; norm on device)
__global__ void k_scale_by_inv(long long n, const double *x,
                               const double *__restrict__ norm, double sign,
                               double *y)
{
    const double f = sign / norm[0];
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n;
         e += (long long)gridDim.x * blockDim.x)
        y[e] = x[e] * f;
}

// ROSW stage set-up (PETSc TSStep_RosW):
//   Z = u + sum_j a[j] Y_j ;  Zdot = sum_j g[j] Y_j
template <int NV>
__global__ void k_stage_combine(long long n, const double *__restrict__ u,
                                VecList Y, CoefList a, CoefList gm,
                                double *__restrict__ Z, double *__restrict__ Zdot)
{
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n;
         e += (long long)gridDim.x * blockDim.x) {
        double z = u[e], zd = 0.0;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const double y = __ldg(Y.v[i] + e);
            z = fma(a.c[i], y, z);
            zd = fma(gm.c[i], y, zd);
        }
        Z[e] = z;
        Zdot[e] = zd;
    }
}

// ROSW completion + embedded error estimate (TSEvaluateStep_RosW +
// TSErrorWeightedNorm NORM_2):
//   unew = u + sum b[j] Y_j ; uemb = u + sum be[j] Y_j
//   partial += ((unew-uemb)/(atol + rtol*max(|unew|,|uemb|)))^2
template <int NV>
__global__ void __launch_bounds__(KSFD_RED_THREADS)
k_complete_step(long long n, const double *__restrict__ u, VecList Y, CoefList b,
                CoefList be, double atol, double rtol, double *__restrict__ unew,
                double *__restrict__ partial)
{
    double acc[1] = {0.0};
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n;
         e += (long long)gridDim.x * blockDim.x) {
        const double u0 = u[e];
        double un = u0, ue = u0;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const double y = __ldg(Y.v[i] + e);
            un = fma(b.c[i], y, un);
            ue = fma(be.c[i], y, ue);
        }
        unew[e] = un;
        const double tol = atol + rtol * fmax(fabs(un), fabs(ue));
        const double r = (un - ue) / tol;
        acc[0] = fma(r, r, acc[0]);
    }
    block_reduce_store<1>(acc, partial, KSFD_RED_BLOCKS);
}

// partial sums of field 0 (worm count, KSFD/ksfdts.py:239-246); plane-SoA
__global__ void __launch_bounds__(KSFD_RED_THREADS)
k_sum_dof0(long long npts, long long plane_pts, int dof, const double *__restrict__ u,
           double *__restrict__ partial)
{
    double acc[1] = {0.0};
    for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < npts;
         p += (long long)gridDim.x * blockDim.x) {
        const long long k = p / plane_pts;
        acc[0] += u[k * dof * plane_pts + (p - k * plane_pts)];
    }
    block_reduce_store<1>(acc, partial, KSFD_RED_BLOCKS);
}

__global__ void k_scale_dof0(long long npts, long long plane_pts, int dof, double f,
                             double *__restrict__ u)
{
    for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < npts;
         p += (long long)gridDim.x * blockDim.x) {
        const long long k = p / plane_pts;
        u[k * dof * plane_pts + (p - k * plane_pts)] *= f;
    }
}
