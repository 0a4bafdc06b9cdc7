// Marching residual kernels, one dimension per object file (-DKSFD_MARCH_DIM).
#include "march_launch.cuh"

#ifndef KSFD_MARCH_DIM
#error "compile with -DKSFD_MARCH_DIM=2 or 3"
#endif
#define DIM KSFD_MARCH_DIM

template <int NLIG, bool FIXED>
static int launch_residual_f(ksfd_ctx *c, const HostVec &u, const double *udot, const double *src,
                             double *out, const void *push, cudaStream_t st)
{
    ResidualOp<DIM, NLIG, FIXED> op{u.r, udot, src, out};
    if (push) op.hp = *static_cast<const HaloPush *>(push);     // (else all-null: no push)
    const double cemit = 35.0 * DIM + 30.0;
    // (the TMA-fed marcher is not used here: measured slower for the transcendental-heavy
    // staging of the residual, whose halo points would be one warp's extra pass —
    // profiles/r02_tma_tuner_runs.txt)
#if KSFD_MARCH_DIM == 2
    return launch_op<DIM, ResidualOp<DIM, NLIG, FIXED>, false, 124, 1, 6, 252, 1, 3>(
        c, op, FIXED ? 1 : 5, 150.0, cemit, nullptr, st);
#else
    return launch_op<DIM, ResidualOp<DIM, NLIG, FIXED>, false, 16, 16, 2, 32, 16, 1>(
        c, op, FIXED ? 1 : 5, 150.0, cemit, nullptr, st);
#endif
}

// the implicit-step case (udot given, no sources) has its own instantiation
// without the runtime tests
template <int NLIG>
static int launch_residual(ksfd_ctx *c, const HostVec &u, const double *udot, const double *src,
                           double *out, const void *push, cudaStream_t st)
{
    if (udot && !src) return launch_residual_f<NLIG, true>(c, u, udot, src, out, push, st);
    return launch_residual_f<NLIG, false>(c, u, udot, src, out, push, st);
}

int KSFD_CAT(ksfd_march_residual_d, KSFD_MARCH_DIM)(ksfd_ctx *c, const HostVec &u,
                                                    const double *udot, const double *src,
                                                    double *out, const void *push, cudaStream_t st)
{
    KSFD_DISPATCH_NLIG(launch_residual, c, u, udot, src, out, push, st);
}
