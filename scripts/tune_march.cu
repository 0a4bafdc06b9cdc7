// Stand-alone tuner for the marching kernels (not part of the library):
// times k_march variants (tile, min-blocks, unroll mode, planes per CTA) on
// synthetic options84 data and prints one line per variant.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo
//        -I ksfd_b200/csrc scripts/tune_march.cu -o gpurun_out/tune_march
//   ./tune_march [2d N | 3d N] ...
// -DKSFD_MARCH_VARIANT=bits (1: stage first, 2: column clusters with distributed shared
// memory, CLn lines, 4: register prefetch two planes deep) compiles the experimental code paths of
// march_kernels.cuh (build one binary per variant; the checksums of the outputs
// must agree bit for bit between them).  TUNE_ALL=1 adds the rejected variants
// (cp.async pipeline, other min-blocks) to the library's own configurations.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "march_kernels.cuh"

#define CHECK(x)                                                                    \
    do {                                                                            \
        cudaError_t e_ = (x);                                                       \
        if (e_ != cudaSuccess) {                                                    \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            exit(1);                                                                \
        }                                                                           \
    } while (0)

static DevPhys make_phys(int dim)
{
    DevPhys P{};
    P.ngroups = 2;
    P.nlig = 2;
    P.cap_type = 0;
    P.dim = dim;
    P.s2 = 0.02357 * 0.02357 / 2;
    P.rhomax = 28000.0;
    P.inv_cushion = 1.0 / 2000.0;
    P.capscale = 2.0 * P.s2;
    P.rhomin = 1e-7;
    P.Umin = 1e-7;
    P.inv_rhomax = 1.0 / 28000.0;
    P.alpha[0] = P.alpha[1] = 1500.0;
    P.beta[0] = 5.56e-4;
    P.beta[1] = -5.56e-4;
    P.lig_group[0] = 0;
    P.lig_group[1] = 1;
    P.weight[0] = P.weight[1] = 1.0;
    P.Wgl[0][0] = 1.0;
    P.Wgl[1][1] = 1.0;
    P.s[0] = P.gamma[0] = 0.01;
    P.D[0] = 1e-6;
    P.s[1] = P.gamma[1] = 0.001;
    P.D[1] = 1e-5;
    const double h = 1.0 / 384;
    const double a1[5] = {1, -8, 0, 8, -1}, a2[5] = {-1, 16, -30, 16, -1};
    for (int a = 0; a < dim; ++a)
        for (int s = 0; s < 5; ++s) {
            P.w1[a][s] = a1[s] / (12 * h);
            P.w2[a][s] = a2[s] / (12 * h * h);
        }
    P.w2c = dim * a2[2] / (12 * h * h);
    for (int a = 0; a < dim; ++a) {
        P.c1[a] = 1 / (12 * h);
        P.c2[a] = 1 / (12 * h * h);
        P.c1sq[a] = P.c1[a] * P.c1[a];
    }
    P.sym_ok = 1;
    P.ycap1 = -2.0 * P.inv_cushion;
    P.ycap0 = 2.0 * P.rhomax * P.inv_cushion;
    P.capscale2 = 2.0 * P.capscale;
    P.mk = fastk_default();
    return P;
}

__global__ void k_fill(double *p, long long n, double base, double amp, unsigned seed)
{
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned long long x = (i + 1) * 0x9E3779B97F4A7C15ull + seed * 0xD1B54A32D192ED03ull;
    x ^= x >> 29;
    x *= 0xBF58476D1CE4E5B9ull;
    x ^= x >> 32;
    p[i] = base + amp * ((double)(x & 0xFFFFFF) / 16777216.0 - 0.5);
}

// order-independent checksum of an output vector (integer sum of the bit patterns)
__global__ void k_cks(const double *p, long long n, unsigned long long *out)
{
    unsigned long long s = 0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x)
        s += (unsigned long long)__double_as_longlong(p[i]) * 0x9E3779B97F4A7C15ull + (unsigned long long)i;
    atomicAdd(out, s);
}

struct Problem {
    int dim, n0, n1, nloc;
    long long npts;
    int nrot;
    std::vector<double *> u, v, out;
    double *coef, *pc;
};

static int g_ordinal = 0;
static bool selected()
{
    // TUNE_ONLY="3,7": run only the variants with these ordinals (for ncu)
    const int me = g_ordinal++;
    const char *f = getenv("TUNE_ONLY");
    if (!f || !*f) return true;
    char buf[256];
    strncpy(buf, f, 255);
    buf[255] = 0;
    for (char *t = strtok(buf, ","); t; t = strtok(nullptr, ","))
        if (atoi(t) == me) return true;
    return false;
}

template <int DIM, int TX, int TY, class Op, int MINB, bool UNR, int DEPTH>
static void run_variant(const char *name, const Problem &pb, const DevPhys &P, Op op_proto,
                        void (*bind)(Op &, const Problem &, int), const int *rzs, int nrz)
{
    using T = TileT<DIM, TX, TY>;
    const int ord = g_ordinal;
    if (!selected()) return;
    if (getenv("TUNE_ONLY")) nrz = 1;
    auto kern = k_march<DIM, TX, TY, Op, MINB, UNR, DEPTH>;
    const size_t smem = march_smem_bytes<Op, T::SP, T::NT, DEPTH>();
    CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, T::NT, smem));
    cudaFuncAttributes fa;
    CHECK(cudaFuncGetAttributes(&fa, kern));
    const int ntx = (pb.n0 + TX - 1) / TX, ox = (pb.n0 + ntx - 1) / ntx;
    const int nty = DIM == 3 ? (pb.n1 + TY - 1) / TY : 1, oy = DIM == 3 ? (pb.n1 + nty - 1) / nty : 1;
    cudaEvent_t e0, e1;
    CHECK(cudaEventCreate(&e0));
    CHECK(cudaEventCreate(&e1));
    for (int r = 0; r < nrz; ++r) {
        int rz = rzs[r];
        if (rz <= 0) {
            // auto: fill occ*148 slots in whole waves
            const long long cols = (long long)ntx * nty;
            const int slots = 148 * occ;
            int waves = (int)((cols + slots - 1) / slots);
            int nch = (int)((long long)waves * slots / cols);
            if (nch < 1) nch = 1;
            if (-rz > 1) nch *= -rz;        // rz = -k: k waves
            rz = (pb.nloc + nch - 1) / nch;
            if (rz < 2) rz = 2;
        }
        if (rz > pb.nloc) rz = pb.nloc;
        const int nch = (pb.nloc + rz - 1) / rz;
        MarchArgs a{pb.n0, pb.n1, pb.nloc, pb.n0 * pb.n1, ox, oy, rz};
        dim3 grid(ntx, nty, nch);
        const int reps = getenv("TUNE_ONLY") ? 2 : 20;
        for (int i = 0; i < 3; ++i) {
            Op op = op_proto;
            bind(op, pb, i % pb.nrot);
            kern<<<grid, T::NT, smem>>>(a, P, op, nullptr);
        }
        CHECK(cudaDeviceSynchronize());
        CHECK(cudaEventRecord(e0));
        for (int i = 0; i < reps; ++i) {
            Op op = op_proto;
            bind(op, pb, i % pb.nrot);
            kern<<<grid, T::NT, smem>>>(a, P, op, nullptr);
        }
        CHECK(cudaEventRecord(e1));
        CHECK(cudaDeviceSynchronize());
        float ms = 0;
        CHECK(cudaEventElapsedTime(&ms, e0, e1));
        const double us = ms * 1e3 / reps;
        const double gpts = pb.npts / us / 1e3;
        // checksum of one launch on buffer set 0
        unsigned long long *dck, hck = 0;
        CHECK(cudaMalloc(&dck, 8));
        CHECK(cudaMemset(dck, 0, 8));
        {
            Op op = op_proto;
            bind(op, pb, 0);
            kern<<<grid, T::NT, smem>>>(a, P, op, nullptr);
            k_cks<<<592, 256>>>(pb.out[0], pb.npts * 3, dck);
        }
        CHECK(cudaMemcpy(&hck, dck, 8, cudaMemcpyDeviceToHost));
        CHECK(cudaFree(dck));
        printf("#%02d %-12s V%d TX%3d TY%2d MINB%d UNR%d D%d regs%3d occ%d rz%4d grid %4dx%3dx%4d  %9.2f us  %6.2f Gpts/s  frac %.3f  cks %016llx\n",
               ord, name, KSFD_MARCH_VARIANT, TX, TY, MINB, (int)UNR, DEPTH, fa.numRegs, occ, rz, ntx,
               nty, nch, us, gpts, gpts * 72.0 / 6544.7, hck);
    }
    fflush(stdout);
}

#if KSFD_MARCH_VARIANT & 2
// column-cluster variant (march_kernels.cuh: ClusterMarcher): grid.z is a multiple of CZ,
// every chunk holds the same number of planes
template <int DIM, int TX, int TY, class Op, int MINB, bool UNR>
static void run_variant_cl(const char *name, const Problem &pb, const DevPhys &P, Op op_proto,
                           void (*bind)(Op &, const Problem &, int), int CZ)
{
    using T = TileT<DIM, TX, TY>;
    const int ord = g_ordinal;
    if (!selected()) return;
    auto kern = k_march_cl<DIM, TX, TY, Op, MINB, UNR>;
    const size_t smem = march_cl_smem_bytes<Op, T::SP, T::NT>();
    CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, T::NT, smem));
    cudaFuncAttributes fa;
    CHECK(cudaFuncGetAttributes(&fa, kern));
    const int ntx = (pb.n0 + TX - 1) / TX, ox = (pb.n0 + ntx - 1) / ntx;
    const int nty = DIM == 3 ? (pb.n1 + TY - 1) / TY : 1, oy = DIM == 3 ? (pb.n1 + nty - 1) / nty : 1;
    cudaEvent_t e0, e1;
    CHECK(cudaEventCreate(&e0));
    CHECK(cudaEventCreate(&e1));
    for (int waves = 1; waves <= 2; ++waves) {
        // chunks: fill occ*148 slots `waves` times, rounded to a multiple of CZ that divides nloc
        const long long cols = (long long)ntx * nty;
        int nch = (int)((long long)waves * 148 * occ / cols);
        nch = nch / CZ * CZ;
        while (nch >= CZ && pb.nloc % nch != 0) nch -= CZ;
        if (nch < CZ) continue;
        const int rz = pb.nloc / nch;
        if (rz < 2) continue;
        MarchArgs a{pb.n0, pb.n1, pb.nloc, pb.n0 * pb.n1, ox, oy, rz};
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(ntx, nty, nch);
        cfg.blockDim = dim3(T::NT);
        cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 1;
        at[0].val.clusterDim.y = 1;
        at[0].val.clusterDim.z = CZ;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        const int reps = getenv("TUNE_ONLY") ? 2 : 20;
        for (int i = 0; i < 3; ++i) {
            Op op = op_proto;
            bind(op, pb, i % pb.nrot);
            CHECK(cudaLaunchKernelEx(&cfg, kern, a, P, op));
        }
        CHECK(cudaDeviceSynchronize());
        CHECK(cudaEventRecord(e0));
        for (int i = 0; i < reps; ++i) {
            Op op = op_proto;
            bind(op, pb, i % pb.nrot);
            CHECK(cudaLaunchKernelEx(&cfg, kern, a, P, op));
        }
        CHECK(cudaEventRecord(e1));
        CHECK(cudaDeviceSynchronize());
        float ms = 0;
        CHECK(cudaEventElapsedTime(&ms, e0, e1));
        const double us = ms * 1e3 / reps;
        const double gpts = pb.npts / us / 1e3;
        unsigned long long *dck, hck = 0;
        CHECK(cudaMalloc(&dck, 8));
        CHECK(cudaMemset(dck, 0, 8));
        {
            Op op = op_proto;
            bind(op, pb, 0);
            CHECK(cudaLaunchKernelEx(&cfg, kern, a, P, op));
            k_cks<<<592, 256>>>(pb.out[0], pb.npts * 3, dck);
        }
        CHECK(cudaMemcpy(&hck, dck, 8, cudaMemcpyDeviceToHost));
        CHECK(cudaFree(dck));
        printf("#%02d %-12s CL%d TX%3d TY%2d MINB%d UNR%d regs%3d occ%d rz%4d grid %4dx%3dx%4d  %9.2f us  %6.2f Gpts/s  frac %.3f  cks %016llx\n",
               ord, name, CZ, TX, TY, MINB, (int)UNR, fa.numRegs, occ, rz, ntx, nty, nch, us, gpts,
               gpts * 72.0 / 6544.7, hck);
    }
    fflush(stdout);
}
#endif

template <int DIM>
static void bind_res(ResidualOp<DIM, 2, true> &op, const Problem &pb, int i)
{
    op.u.base = pb.u[i];
    op.u.lo = pb.u[i] + (long long)(pb.nloc - 2) * pb.n0 * pb.n1 * 3;
    op.u.hi = pb.u[i];
    op.u.par = nullptr;
    op.u.pstride = 0;
    op.udot = pb.v[i];
    op.src = nullptr;
    op.out = pb.out[i];
}
template <int DIM, bool PC>
static void bind_jvp(JvpOp<DIM, 2, PC> &op, const Problem &pb, int i)
{
    const long long ps = (long long)pb.n0 * pb.n1;
    op.coef.lo = pb.coef;
    op.coef.base = pb.coef + 2 * ps * 5;
    op.coef.hi = pb.coef + (2 + pb.nloc) * ps * 5;
    op.coef.par = op.v.par = op.pc.par = nullptr;
    op.coef.pstride = op.v.pstride = op.pc.pstride = 0;
    op.v.base = pb.v[i];
    op.v.lo = pb.v[i] + (long long)(pb.nloc - 2) * ps * 3;
    op.v.hi = pb.v[i];
    op.pc.base = pb.pc;
    op.pc.lo = pb.pc + (long long)(pb.nloc - 2) * ps;
    op.pc.hi = pb.pc;
    op.shift = 2294.0;
    op.invd[0] = 1.0 / 2294.5;
    op.invd[1] = 1.0 / 2294.7;
    op.out = pb.out[i];
}

static Problem make_problem(int dim, int n)
{
    Problem pb;
    pb.dim = dim;
    pb.n0 = n;
    pb.n1 = dim == 3 ? n : 1;
    pb.nloc = n;
    pb.npts = (long long)pb.n0 * pb.n1 * pb.nloc;
    const long long N = pb.npts * 3;
    pb.nrot = (int)(2.5 * 126e6 / (N * 8 * 3)) + 1;
    if (pb.nrot < 2) pb.nrot = 2;
    if (pb.nrot > 12) pb.nrot = 12;
    for (int i = 0; i < pb.nrot; ++i) {
        double *u, *v, *o;
        CHECK(cudaMalloc(&u, N * 8));
        CHECK(cudaMalloc(&v, N * 8));
        CHECK(cudaMalloc(&o, N * 8));
        k_fill<<<(unsigned)((N + 255) / 256), 256>>>(u, N, 9000.0, 300.0, 3 * i);
        k_fill<<<(unsigned)((N + 255) / 256), 256>>>(v, N, 0.0, 2.0, 3 * i + 1);
        pb.u.push_back(u);
        pb.v.push_back(v);
        pb.out.push_back(o);
    }
    const long long NC = (long long)(pb.nloc + 4) * pb.n0 * pb.n1 * 5;
    CHECK(cudaMalloc(&pb.coef, NC * 8));
    k_fill<<<(unsigned)((NC + 255) / 256), 256>>>(pb.coef, NC, 1.0, 0.5, 77);
    CHECK(cudaMalloc(&pb.pc, pb.npts * 8));
    k_fill<<<(unsigned)((pb.npts + 255) / 256), 256>>>(pb.pc, pb.npts, 4e-4, 1e-5, 78);
    CHECK(cudaDeviceSynchronize());
    return pb;
}

static void free_problem(Problem &pb)
{
    for (auto p : pb.u) cudaFree(p);
    for (auto p : pb.v) cudaFree(p);
    for (auto p : pb.out) cudaFree(p);
    cudaFree(pb.coef);
    cudaFree(pb.pc);
}

#define RES2(TX, MINB, UNR, D) \
    run_variant<2, TX, 1, ResidualOp<2, 2, true>, MINB, UNR, D>("residual2d", pb, P, ResidualOp<2, 2, true>{}, bind_res<2>, rzs, nrz)
#define JVP2(TX, MINB, UNR, PC, D) \
    run_variant<2, TX, 1, JvpOp<2, 2, PC>, MINB, UNR, D>(PC ? "jvp_pc2d" : "jvp2d", pb, P, JvpOp<2, 2, PC>{}, bind_jvp<2, PC>, rzs, nrz)
#define RES3(TX, TY, MINB, UNR, D) \
    run_variant<3, TX, TY, ResidualOp<3, 2, true>, MINB, UNR, D>("residual3d", pb, P, ResidualOp<3, 2, true>{}, bind_res<3>, rzs, nrz)
#define JVP3(TX, TY, MINB, UNR, PC, D) \
    run_variant<3, TX, TY, JvpOp<3, 2, PC>, MINB, UNR, D>(PC ? "jvp_pc3d" : "jvp3d", pb, P, JvpOp<3, 2, PC>{}, bind_jvp<3, PC>, rzs, nrz)

#define RES2CL(TX, MINB, CZ) \
    run_variant_cl<2, TX, 1, ResidualOp<2, 2, true>, MINB, false>("residual2d", pb, P, ResidualOp<2, 2, true>{}, bind_res<2>, CZ)
#define JVP2CL(TX, MINB, PC, CZ) \
    run_variant_cl<2, TX, 1, JvpOp<2, 2, PC>, MINB, true>(PC ? "jvp_pc2d" : "jvp2d", pb, P, JvpOp<2, 2, PC>{}, bind_jvp<2, PC>, CZ)

int main(int argc, char **argv)
{
    const int rzs[] = {0, -2, -3};     // 1, 2, 3 waves of CTAs
    const int nrz = 3;
    for (int a = 1; a + 1 < argc; a += 2) {
        const int dim = argv[a][0] == '3' ? 3 : 2;
        const int n = atoi(argv[a + 1]);
        Problem pb = make_problem(dim, n);
        DevPhys P = make_phys(dim);
        printf("== %dD n=%d  (%lld points, %d buffer sets)\n", dim, n, pb.npts, pb.nrot);
        const bool all = getenv("TUNE_ALL") != nullptr;
        if (dim == 2) {
            // the library's configurations (march_res.cu / march_jvp.cu)
            RES2(124, 6, false, 0);
            RES2(252, 3, false, 0);
            JVP2(124, 4, true, true, 0);
            JVP2(252, 2, true, true, 0);
            JVP2(124, 4, true, false, 0);
            JVP2(252, 2, true, false, 0);
#if KSFD_MARCH_VARIANT & 4
            // two planes of register prefetch: more registers, fewer CTAs per SM
            RES2(124, 5, false, 0);
            JVP2(124, 3, true, true, 0);
            JVP2(124, 3, true, false, 0);
#endif
#if KSFD_MARCH_VARIANT & 2
            for (int cz = 2; cz <= 8; cz *= 2) {
                RES2CL(124, 6, cz);
                RES2CL(124, 5, cz);
                JVP2CL(124, 4, true, cz);
                JVP2CL(124, 4, false, cz);
            }
            RES2CL(252, 3, 8);
            JVP2CL(252, 2, true, 8);
#endif
            if (all) {
                RES2(124, 6, false, 3);
                RES2(252, 3, false, 3);
                JVP2(124, 4, true, true, 3);
                JVP2(124, 5, true, true, 3);
                JVP2(124, 6, true, true, 3);
                JVP2(252, 3, true, true, 3);
                JVP2(124, 4, false, true, 0);
            }
        } else {
            RES3(16, 16, 2, false, 0);
            RES3(32, 16, 1, false, 0);
            JVP3(32, 8, 1, true, true, 0);
            JVP3(16, 16, 1, true, true, 0);
            JVP3(32, 8, 1, true, false, 0);
            JVP3(16, 16, 1, true, false, 0);
            if (all) {
                RES3(32, 16, 1, false, 3);
                RES3(16, 16, 2, false, 3);
                JVP3(32, 8, 1, true, true, 3);
                JVP3(32, 8, 2, true, true, 3);
                JVP3(32, 16, 1, true, true, 3);
            }
        }
        free_problem(pb);
    }
    return 0;
}
