// Marching velocity(-max) kernels, one dimension per object file.
#include "march_launch.cuh"

#ifndef KSFD_MARCH_DIM
#error "compile with -DKSFD_MARCH_DIM=2 or 3"
#endif
#define DIM KSFD_MARCH_DIM

template <int NLIG>
static int launch_velocity(ksfd_ctx *c, const HostVec &u, double *vel, double *vmax,
                           cudaStream_t st)
{
    VelocityOp<DIM, NLIG> op{u.r, vel, vmax};
    const double cemit = 8.0 * DIM + 10.0;
#if KSFD_MARCH_DIM == 2
    return launch_op<DIM, VelocityOp<DIM, NLIG>, false, 124, 1, 6, 252, 1, 3>(
        c, op, 4, 150.0, cemit, nullptr, st);
#else
    return launch_op<DIM, VelocityOp<DIM, NLIG>, false, 16, 16, 2, 32, 16, 1>(
        c, op, 4, 150.0, cemit, nullptr, st);
#endif
}

int KSFD_CAT(ksfd_march_velocity_d, KSFD_MARCH_DIM)(ksfd_ctx *c, const HostVec &u, double *vel,
                                                    double *vmax, cudaStream_t st)
{
    KSFD_DISPATCH_NLIG(launch_velocity, c, u, vel, vmax, st);
}
