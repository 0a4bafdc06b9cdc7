// Marching J.v kernels, one dimension per object file (-DKSFD_MARCH_DIM).
#include "march_launch.cuh"

#ifndef KSFD_MARCH_DIM
#error "compile with -DKSFD_MARCH_DIM=2 or 3"
#endif
#define DIM KSFD_MARCH_DIM

template <int NLIG, bool PRECOND>
static int launch_jvp_p(ksfd_ctx *c, const HostVec &coef, const HostVec &v, const HostVec &pc,
                        double *out, const int *skip, cudaStream_t st)
{
    JvpOp<DIM, NLIG, PRECOND> op;
    op.coef = coef.r;
    op.v = v.r;
    op.pc = pc.r;
    op.shift = c->shift;
    for (int l = 0; l < NLIG; ++l) op.invd[l] = c->invd[l];
    op.out = out;
    const double cstage = PRECOND ? 45.0 : 30.0, cemit = 40.0 * DIM + 30.0;
    if (ksfd_use_tma(c)) {
        const TmaSrc src[3] = {coef.t, v.t, pc.t};
#if KSFD_MARCH_DIM == 2
        return launch_tma_op<DIM, JvpOp<DIM, NLIG, PRECOND>, true, 256, 1, 2, 4, 128, 1, 4, 3>(
            c, op, src, PRECOND ? 2 : 3, cstage, cemit, skip, st);
#else
        return launch_tma_op<DIM, JvpOp<DIM, NLIG, PRECOND>, true, 16, 16, 2, 2, 32, 16, 1, 2>(
            c, op, src, PRECOND ? 2 : 3, cstage, cemit, skip, st);
#endif
    }
#if KSFD_MARCH_DIM == 2
    return launch_op<DIM, JvpOp<DIM, NLIG, PRECOND>, true, 124, 1, 4, 252, 1, 2>(
        c, op, PRECOND ? 2 : 3, cstage, cemit, skip, st);
#else
    return launch_op<DIM, JvpOp<DIM, NLIG, PRECOND>, true, 32, 8, 1, 16, 16, 1>(
        c, op, PRECOND ? 2 : 3, cstage, cemit, skip, st);
#endif
}

template <int NLIG>
static int launch_jvp(ksfd_ctx *c, const HostVec &coef, const HostVec &v, const HostVec &pc,
                      bool precond, double *out, const int *skip, cudaStream_t st)
{
    if (precond) return launch_jvp_p<NLIG, true>(c, coef, v, pc, out, skip, st);
    return launch_jvp_p<NLIG, false>(c, coef, v, pc, out, skip, st);
}

int KSFD_CAT(ksfd_march_jvp_d, KSFD_MARCH_DIM)(ksfd_ctx *c, const HostVec &coef, const HostVec &v,
                                               const HostVec &pc, bool precond, double *out,
                                               const int *skip, cudaStream_t st)
{
    KSFD_DISPATCH_NLIG(launch_jvp, c, coef, v, pc, precond, out, skip, st);
}
