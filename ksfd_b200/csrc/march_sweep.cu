// Richardson sweep kernels (sweep_op.cuh), one dimension per object file (-DKSFD_MARCH_DIM).
#include "march_launch.cuh"
#include "sweep_op.cuh"
#include <cstring>

#ifndef KSFD_MARCH_DIM
#error "compile with -DKSFD_MARCH_DIM=2 or 3"
#endif
#define DIM KSFD_MARCH_DIM
// plane loop unrolled five times (queue rotates by renaming) or rolled (moves; 5x less code)
#ifndef KSFD_SWEEP_UNR
#define KSFD_SWEEP_UNR 1
#endif

template <int NLIG>
static int launch_sweep(ksfd_ctx *c, const HostVec &coef, const HostVec &r, const HostVec &pc,
                        const SweepHost &a, const int *skip, cudaStream_t st)
{
    SweepOp<DIM, NLIG> op;
    op.coef = coef.r;
    op.v = r.r;
    op.pc = pc.r;
    op.shift = c->shift;
    for (int l = 0; l < NLIG; ++l) op.invd[l] = c->invd[l];
    op.out = nullptr;
    op.rin = r.r.base;
    op.x = a.x;
    op.rout = a.rout;
    op.rsign = a.rsign;
    op.first = a.first;
    op.pad_ = 0;
    op.fin = *static_cast<const SweepFin *>(a.fin);
    const bool push = a.push != nullptr && ksfd_use_tma(c);
    if (push)
        op.hp = *static_cast<const HaloPush *>(a.push);
    else
        memset(&op.hp, 0, sizeof(op.hp));
    const double cstage = 45.0, cemit = 40.0 * DIM + 45.0;
    // (a.partial_cap: CTAs the per-CTA partial-sum buffer has room for)
    if (ksfd_use_tma(c)) {
        const TmaSrc src[3] = {coef.t, r.t, pc.t};
#if KSFD_MARCH_DIM == 2
        return launch_tma_op<DIM, SweepOp<DIM, NLIG>, KSFD_SWEEP_UNR != 0, 256, 1, 2, 4, 128, 1, 4, 3>(
            c, op, src, push ? 5 : 4, cstage, cemit, skip, st, a.partial_cap, push);
#else
        return launch_tma_op<DIM, SweepOp<DIM, NLIG>, KSFD_SWEEP_UNR != 0, 16, 16, 2, 2, 32, 16, 1, 2>(
            c, op, src, push ? 5 : 4, cstage, cemit, skip, st, a.partial_cap, push);
#endif
    }
#if KSFD_MARCH_DIM == 2
    return launch_op<DIM, SweepOp<DIM, NLIG>, KSFD_SWEEP_UNR != 0, 124, 1, 4, 252, 1, 2>(c, op, 4, cstage, cemit,
                                                                          skip, st, a.partial_cap);
#else
    return launch_op<DIM, SweepOp<DIM, NLIG>, KSFD_SWEEP_UNR != 0, 32, 8, 1, 16, 16, 1>(c, op, 4, cstage, cemit,
                                                                         skip, st, a.partial_cap);
#endif
}

int KSFD_CAT(ksfd_march_sweep_d, KSFD_MARCH_DIM)(ksfd_ctx *c, const HostVec &coef, const HostVec &r,
                                                 const HostVec &pc, const SweepHost &a,
                                                 const int *skip, cudaStream_t st)
{
    KSFD_DISPATCH_NLIG(launch_sweep, c, coef, r, pc, a, skip, st);
}
