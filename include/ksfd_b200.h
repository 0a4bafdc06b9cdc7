/*
 * ksfd_b200 — C ABI of the B200-native implicit time-stepping hot path of KSFD.
 *
 * This is the drop-in boundary: a plain C interface (pointers + sizes, no torch
 * or C++ types) that a maintainer of leonavery/KSFD would bind from Python with
 * ctypes.  Each entry point names the reference interface
 * it replaces (file:line under the reference tree).
 *
 * Conventions
 *  - every function returns 0 on success, non-zero on error; the message of the
 *    last error on the calling thread is returned by ksfd_last_error().
 *  - all `double *` vector arguments are DEVICE pointers owned by the caller
 *    (torch tensors on the Python side); no ownership is transferred.
 *  - device vectors hold the rank-local part of a field vector (length
 *    dof * n_local_points) in the library's INTERNAL "plane-SoA" layout:
 *        element(k, c, pp) = (k * dof + c) * plane_pts + pp
 *    (k = plane of the last axis, c = dof, pp = x + nx*y inside the plane) so
 *    that kernels read every field fully coalesced.  The reference's layout
 *    (fp64, Fortran order, dof fastest, then x, y, z; KSFD/ksfdgrid.py:10-28)
 *    exists at the boundary only: ksfd_to_internal / ksfd_from_internal
 *    convert between the two on the device (bit-exact permutations; in 1-D
 *    the layouts coincide).
 *  - decomposition: 1-D slabs along the LAST spatial axis (y in 2-D, z in 3-D,
 *    x in 1-D); ownership ranges equal PETSc DMDA's lx[i] = M/P + (M%P > i).
 *  - `stream` is a cudaStream_t passed as void* (0 = default stream).
 *  - a context is thread-compatible: use it from one host thread at a time.
 */
#ifndef KSFD_B200_H
#define KSFD_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KSFD_MAX_LIGANDS 7      /* dof <= 8 */
#define KSFD_MAX_GROUPS 7
#define KSFD_ABI_VERSION 3      /* 2: stream argument on norm2 / sum_dof0 / allreduce_*;
                                   3: ksfd_ksp_opts.ksp_type, ksfd_ksp_solve */

typedef struct ksfd_ctx ksfd_ctx;

/*
 * Plain-number problem description at one time t.  Replaces the constants the
 * reference folds into its generated C ufuncs (KSFD/ksfdufunc.py:126-225) and
 * the time-dependent scalar ufunc arguments (KSFD/ksfdsym.py:1432-1438).
 * Potential: KSFD/ksfdligand.py:527-547,720-746; KSFD/ksfdsoln.py:147-161.
 */
typedef struct ksfd_physics {
    int32_t ngroups;                       /* ligand groups with >= 1 ligand   */
    int32_t nlig;                          /* total ligands (dof = nlig + 1)   */
    int32_t cap_type;                      /* 0 = tophat, 1 = witch            */
    int32_t reserved;
    double s2, rhomax, cushion, maxscale;  /* G = V + s2*log(rho)              */
    double rhomin, Umin;                   /* clamp (KSFD/ksfdsym.py:888-900)  */
    double alpha[KSFD_MAX_GROUPS];
    double beta[KSFD_MAX_GROUPS];
    int32_t lig_group[KSFD_MAX_LIGANDS + 1]; /* group index of each ligand     */
    double weight[KSFD_MAX_LIGANDS];
    double s[KSFD_MAX_LIGANDS];
    double gamma[KSFD_MAX_LIGANDS];
    double D[KSFD_MAX_LIGANDS];
    /* finite-difference weights per axis for offsets -2,-1,0,+1,+2
       (KSFD/ksfdsym.py:391-436), computed on the host the way the reference
       does so last-bit asymmetries are kept */
    double w1[3][5];
    double w2[3][5];
} ksfd_physics;

/* ---- library ---------------------------------------------------------- */
int ksfd_abi_version(void);
const char *ksfd_last_error(void);
/* number of CUDA kernels launched by this library in this process so far */
int64_t ksfd_launch_count(void);

/* ---- context: grid + distribution (replaces KSFD.Grid's three PETSc DMDAs,
 *      KSFD/ksfdgrid.py:61-177,388-411) -------------------------------- */
int ksfd_ctx_create(ksfd_ctx **out, int dim, const int64_t n_global[3],
                    int64_t last_start, int64_t last_count, int dof,
                    int device);
int ksfd_ctx_destroy(ksfd_ctx *ctx);
int ksfd_set_physics(ksfd_ctx *ctx, const ksfd_physics *phys);
/* kernel selection / tuning: key in {"variant","tile","rz","gmres_pipeline",
   "gmres_runahead","gmres_cycle_exp","halo_p2p","profile"};
   variant 0 = auto, 1 = naive direct kernels, 2 = marching kernels (J.v through the
   TMA-fed marcher where the grid is eligible), 3 = marching, register-prefetch kernels only;
   profile 1 = bracket every residual / J.v stencil launch with CUDA events
   (ksfd_profile_fetch), 0 = off (default);
   tile = index of the marching tile shape (-1 = auto), rz = planes per CTA;
   gmres_pipeline 1 = device-decided launch-ahead GMRES (default), 0 = host
   driven; gmres_runahead = Arnoldi steps launched ahead; gmres_cycle_exp = k:
   close a cycle after a 1e-k residual reduction; halo_p2p 0/1 */
int ksfd_set_option(ksfd_ctx *ctx, const char *key, int64_t value);
int64_t ksfd_local_size(const ksfd_ctx *ctx);   /* dof * owned points */
/* In-situ kernel timing (option "profile"): synchronises `stream`, then for each kind
   k < 8 of bracketed launch (0 J.v stencil, 1 residual stencil, 2 multi-dot incl. rank
   sum and Givens update, 3 orthogonalise-and-scale incl. halo push, 4 first Krylov
   vector, 5 start of a GMRES cycle, 6-7 unused):
   out[4k] = ACTIVE launches since the last fetch, out[4k+1] = their summed device time
   in ms, out[4k+2] = all bracketed launches, out[4k+3] = their summed time; resets the
   counters.  Launches the pipelined solver made ahead of a convergence test and that
   returned at once (~2 us) are not real passes: a launch is ACTIVE when it lasted at
   least 4 us. */
int ksfd_profile_fetch(ksfd_ctx *ctx, double out[32], void *stream);

/* ---- layout boundary: reference layout (what PETSc Vec.array / the HDF5
 *      TimeSeries hold, KSFD/ksfdtimeseries.py:485-488) <-> internal ------ */
int ksfd_to_internal(ksfd_ctx *ctx, const double *ref_dev, double *out_dev,
                     int nfields, void *stream);
int ksfd_from_internal(ksfd_ctx *ctx, const double *in_dev, double *ref_dev,
                       int nfields, void *stream);

/* ---- multi-GPU: NCCL ring of slabs (replaces DMDA globalToLocal,
 *      KSFD/ksfdsym.py:704,787,920,1203, and mpi allreduce,
 *      KSFD/ksfdts.py:244,252,310-313) --------------------------------- */
int ksfd_nccl_unique_id(const char *libnccl_path, char id_out[128]);
int ksfd_comm_init(ksfd_ctx *ctx, const char *libnccl_path, int nranks,
                   int rank, const char id[128]);
/* optional: direct peer-to-peer exchange over NVLink peer memory instead of
   NCCL (<= 16 ranks on one node).  export: allocate this rank's IPC-shared
   buffers and return their CUDA IPC handle (64 bytes); the caller gathers the
   handles of ALL ranks in rank order (MPI / torch.distributed) and hands them
   to import (nhandles = number of ranks, 64 bytes each).  Afterwards a halo
   exchange is ONE kernel that stores the boundary planes into the neighbours'
   buffers and waits on flag words, and the small all-reduces of the Krylov
   solver (dot products, norms) run INSIDE its reduction kernels. */
int ksfd_p2p_export(ksfd_ctx *ctx, char handle_out[64]);
int ksfd_p2p_import(ksfd_ctx *ctx, const char *handles, int nhandles);
/* fill this rank's ghost planes of `vec` (kept inside the context, slot 0..3) */
int ksfd_halo_exchange(ksfd_ctx *ctx, const double *vec, int slot,
                       void *stream);

/* ---- operator: replaces Derivatives.groom/dfdt/velocity and implicitIF
 *      (KSFD/ksfdsym.py:888-940,1188-1209; KSFD/ksfdts.py:563-596) ------- */
int ksfd_groom(ksfd_ctx *ctx, double *u, void *stream);
/* f_out = udot - (f(u) + src)  when udot != NULL   (implicitIF)
   f_out =        f(u) + src    when udot == NULL   (Derivatives.dfdt)
   src may be NULL (no source terms); u is clamped on the fly, not modified */
int ksfd_residual(ksfd_ctx *ctx, const double *u, const double *udot,
                  const double *src, double *f_out, void *stream);
/* vmax_out[d] = max |grad_d G| over owned points, d < dim (device pointer);
   caller reduces over ranks (ksfd_allreduce_max)  (KSFD/ksfdts.py:302-319) */
int ksfd_velocity_max(ksfd_ctx *ctx, const double *u, double *vmax_out,
                      void *stream);
int ksfd_velocity(ksfd_ctx *ctx, const double *u, double *vel_out,
                  void *stream);      /* internal layout with dim fields */

/* ---- Jacobian action: replaces implicitIJ / Derivatives.Jacobian /
 *      ksfdMat.setValuesJacobian (KSFD/ksfdts.py:598-640,
 *      KSFD/ksfdsym.py:630-886, cython/ksfdMat/ksfdMat.pyx:55-180) ------- */
/* linearise at u_lin: builds the per-point coefficient field and the
   block-Jacobi preconditioner of A = shift*I - df/du on the device */
int ksfd_jvp_setup(ksfd_ctx *ctx, const double *u_lin, double shift,
                   void *stream);
int ksfd_jvp(ksfd_ctx *ctx, const double *v, double *out, void *stream);
/* out = A * M^{-1} v  (fused right-preconditioned operator) */
int ksfd_jvp_precond(ksfd_ctx *ctx, const double *v, double *out,
                     void *stream);
int ksfd_pc_apply(ksfd_ctx *ctx, const double *r, double *z, void *stream);
/* dense copy of the per-point diagonal blocks of A (dof*dof per point, row
   major), for tests */
int ksfd_block_diagonal(ksfd_ctx *ctx, double *blocks_out, void *stream);

/* ---- fused BLAS-1 (replaces PETSc VecMAXPY/VecMDot/VecNorm) ------------ */
int ksfd_mdot(ksfd_ctx *ctx, int nv, const double *const *vs, const double *w,
              double *out_dev, void *stream);   /* out[i] = <vs[i], w> global */
int ksfd_maxpy(ksfd_ctx *ctx, int nv, const double *coef_host,
               const double *const *vs, double *y, void *stream);
/* global reductions to a HOST scalar: enqueued on `stream`, which is synchronised before
   returning.  All reductions of a context share scratch buffers and the peer-to-peer
   sequence counter: issue them on the stream the context's other work runs on. */
int ksfd_norm2(ksfd_ctx *ctx, const double *x, double *out_host, void *stream);
int ksfd_sum_dof0(ksfd_ctx *ctx, const double *u, double *out_host, void *stream);
int ksfd_scale_dof0(ksfd_ctx *ctx, double *u, double factor, void *stream);
/* u[dof 0] *= exp(sd * z): the lognormal noise injection of KSFDTS.add_variance
   (KSFD/ksfdts.py:268-284).  z_dev: one standard-normal value per owned point (x fastest),
   drawn by the caller from the rank's numpy stream — the reference's stream — and copied
   to the device; the field stays on the device. */
int ksfd_mul_exp_dof0(ksfd_ctx *ctx, double *u, const double *z_dev, double sd, void *stream);

/* ---- linear solve: replaces KSP preonly + PC LU (MUMPS) with a
 *      device-resident restarted GMRES, right-preconditioned by block
 *      Jacobi.  Solves A x = rhs with A from the last ksfd_jvp_setup ------ */
typedef struct ksfd_ksp_opts {
    double rtol, atol, dtol;
    int32_t max_it, restart;
    int32_t reorth;            /* 0 = classical GS once, 1 = twice (CGS2) */
    int32_t precond;           /* 0 = none, 1 = point-block Jacobi, 2 = spectral
                                  (FFT inverse of the frozen-coefficient operator;
                                  one rank, needs cuFFT, else 1), 3 = automatic:
                                  1 until a solve needs >= 16 steps, then 2 */
    int32_t ksp_type;          /* -ksp_type: 0 = gmres, 1 = richardson (stationary sweeps
                                  x += M^-1 r fused into the stencil kernel, block Jacobi
                                  only), 2 = automatic: sweeps while they contract the
                                  residual by >= 65 % per sweep, else GMRES from the iterate
                                  reached; ignored (gmres) where sweeps are not available */
    int32_t reserved;
} ksfd_ksp_opts;
typedef struct ksfd_ksp_result {
    int32_t its, reason;       /* reason > 0 converged, < 0 diverged      */
    double rnorm0, rnorm;
} ksfd_ksp_result;
/* GMRES whatever opts->ksp_type says */
int ksfd_gmres(ksfd_ctx *ctx, const double *rhs, double *x,
               const ksfd_ksp_opts *opts, ksfd_ksp_result *res, void *stream);
/* the solver opts->ksp_type selects (what ksfd_ts_step calls for its stage systems) */
int ksfd_ksp_solve(ksfd_ctx *ctx, const double *rhs, double *x,
                   const ksfd_ksp_opts *opts, ksfd_ksp_result *res, void *stream);

/* one stationary sweep of the block-Jacobi preconditioned iteration, fused into the
   stencil kernel (what ksp_type 1/2 iterate):  r_out = r_in - A M^-1 r_in,
   x = (first ? 0 : x) + M^-1 r_in;  norms (host, may be NULL) = ||r_in||, ||r_out|| over
   all ranks.  The three vectors must be distinct. */
int ksfd_sweep(ksfd_ctx *ctx, const double *r_in, double *x, double *r_out, int first,
               double norms[2], void *stream);

/* ---- time step: replaces PETSc TS.step() for -ts_type rosw (ra34pw2) and
 *      beuler with -snes_type ksponly (call site KSFD/ksfdts.py:211) ------ */
typedef void (*ksfd_time_cb)(double t, void *user);  /* set physics/sources */
typedef struct ksfd_ts_opts {
    int32_t ts_type;           /* 0 = rosw ra34pw2, 1 = beuler */
    int32_t adapt;             /* 0 = none, 1 = basic          */
    double atol, rtol;         /* TSSetTolerances               */
    double clip_lo, clip_hi, dt_min, dt_max, safety, reject_safety;
    int32_t max_reject;
    int32_t flags;             /* KSFD_TS_*: work of the reference's step loop done inside the call */
    ksfd_ksp_opts ksp;
} ksfd_ts_opts;
/* clamp u first (the loop's groom(u), KSFD/ksfdts.py:205,231-237) */
#define KSFD_TS_GROOM 1
/* also return max |grad G| per axis of the accepted state (the loop's CFL_check,
   KSFD/ksfdts.py:287-319): computed on the device and read with the error norm */
#define KSFD_TS_VELOCITY_MAX 2
typedef struct ksfd_ts_result {
    double t_new, h_used, h_next, enorm;
    int32_t accepted, rejections, ksp_its, ksp_fail;
    double vmax[3];            /* valid when have_vmax (KSFD_TS_VELOCITY_MAX, accepted step) */
    int32_t have_vmax, reserved;
} ksfd_ts_result;
/* advance u (in place) from t by one accepted step of size <= h.
   src: device array of source terms or NULL; if cb != NULL it is called with
   each stage time before the stage residual so the host can update physics /
   refill src. */
int ksfd_ts_step(ksfd_ctx *ctx, double *u, double t, double h,
                 const ksfd_ts_opts *opts, const double *src, ksfd_time_cb cb,
                 void *user, ksfd_ts_result *res, void *stream);

/* scalar all-reduce helpers over the context communicator (host values) */
int ksfd_allreduce_max(ksfd_ctx *ctx, double *vals_host, int n, void *stream);   /* 0 <= n <= 64 */
int ksfd_allreduce_sum(ksfd_ctx *ctx, double *vals_host, int n, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* KSFD_B200_H */
