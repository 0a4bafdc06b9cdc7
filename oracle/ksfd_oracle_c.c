/*
 * CPU oracle in C (OpenMP): a second, multi-threaded restatement of the reference's
 * algorithm for the implicit time-stepping hot path of leonavery/KSFD.
 *
 * TEST INFRASTRUCTURE ONLY.  Loaded (through oracle/ksfd_oracle_c.py) only by tests/ and by
 * bench.py's cpu_baseline / --impl reference legs, as the checker and as the CPU baseline.
 * The product (ksfd_b200/) never loads it.
 *
 * It follows oracle/ksfd_oracle.py operation by operation (same association order, built with
 * -ffp-contract=off), so the two agree to the last bits of log/tanh; the numpy oracle is the
 * one pinned against the golden vectors of the reference's own code, and
 * tests/test_oracle_c.py pins this file against both.  What each function restates:
 *
 *   oc_groom        Derivatives.groom            KSFD/ksfdsym.py:888-900
 *   G / dG          Guf, Gsubs                   KSFD/ksfdsym.py:983-1033, ksfdligand.py:527-547
 *   oc_dfdt         Derivatives.dfdt / drhodt    KSFD/ksfdsym.py:902-940, 763-812, 531-628
 *   oc_ifunction    implicitTS.implicitIF        KSFD/ksfdts.py:563-596
 *   oc_velocity_max Derivatives.velocity + CFL   KSFD/ksfdsym.py:1188-1209, ksfdts.py:302-319
 *   oc_jvp_setup /  implicitIJ, rhoJacobian_arrays, UJacobian_arrays (applied matrix-free:
 *   oc_jvp          the exact derivative of the discrete f) KSFD/ksfdts.py:598-640,
 *                                                 ksfdsym.py:630-761
 *   oc_beuler_step  PETSc TSBEULER with -snes_type ksponly (one Newton step); not in the tree either
 *   oc_rosw_step    PETSc TSROSW ra34pw2 behind TS.step() (KSFD/ksfdts.py:211), -snes_type
 *                   ksponly; NOT in the reference tree (parity unpinned at that boundary, see
 *                   oracle/ksfd_oracle.py).  The stage systems are solved iteratively
 *                   (point-block-Jacobi Richardson, GMRES(30) when that contracts slowly)
 *                   instead of the reference's MUMPS LU: at the benchmark size (3.1 M unknowns)
 *                   a sparse LU per step takes minutes; tests compare with SuperLU solves.
 *
 * Reproducible to the last bit: reductions are summed over fixed chunks in chunk order, every other
 * loop is elementwise — the result does not depend on the number of threads or on the run.
 *
 * Layout: flat fp64, dof fastest then x, y, z (KSFD/ksfdgrid.py:10-28); one process owns the
 * whole periodic grid; ghost points are the periodic images (DMDA globalToLocal on one rank).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define OC_MAXLIG 8
#define OC_MAXDOF (OC_MAXLIG + 1)
#define OC_ABI 1
/* reductions are summed over fixed chunks in chunk order: results do not depend on the number of
   threads or on the run (a checker should be reproducible to the last bit) */
#define OC_CHUNK 4096

typedef struct {
    int dim, n[3], nlig, ngroups, witch, pad_;
    double s2, rhomax, cushion, maxscale, rhomin, Umin;
    double alpha[OC_MAXLIG], beta[OC_MAXLIG];          /* per group */
    int lig_group[OC_MAXLIG];
    double lig_w[OC_MAXLIG], lig_s[OC_MAXLIG], lig_gamma[OC_MAXLIG], lig_D[OC_MAXLIG];
    double w1[3][5], w2[3][5];                          /* offsets -2..+2 */
} oc_phys;

typedef struct {
    double At[4][4], Gi[4][4], bt[4], bet[4], asum[4], gamma;
} oc_tableau;

typedef struct {
    oc_phys P;
    long npts;
    int dof, ncf;
    int *nb[3][5];              /* periodic neighbour coordinate per axis and offset */
    double *ug, *G;             /* clamped state, G at every point */
    double shift;
    int have_jac;
    double *grho, *gU;          /* dG/drho, dG/dU_l at every point */
    double *cf;                 /* per point: dG[dim], drho[dim], lapG, rho0 */
    double *Minv;               /* inverse diagonal block, dof*dof per point */
    double *dGv;
    double *r, *z, *w, *F, *Z, *Zd, *Y[4], *un;
    double *V;                  /* GMRES basis (allocated when first needed) */
    int vcap;
    double *part;               /* per-chunk partial sums of the reductions */
} oc_ctx;

int oc_abi(void) { return OC_ABI; }
int oc_phys_size(void) { return (int)sizeof(oc_phys); }
int oc_tableau_size(void) { return (int)sizeof(oc_tableau); }
int oc_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

static double *dalloc(long n)
{
    double *p = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
    if (p) memset(p, 0, sizeof(double) * (size_t)(n > 0 ? n : 1));
    return p;
}

void oc_destroy(oc_ctx *c)
{
    if (!c) return;
    for (int a = 0; a < 3; ++a)
        for (int o = 0; o < 5; ++o) free(c->nb[a][o]);
    free(c->ug); free(c->G); free(c->grho); free(c->gU); free(c->cf); free(c->Minv);
    free(c->dGv); free(c->r); free(c->z); free(c->w); free(c->F); free(c->Z); free(c->Zd);
    for (int i = 0; i < 4; ++i) free(c->Y[i]);
    free(c->un);
    free(c->V);
    free(c->part);
    free(c);
}

oc_ctx *oc_create(const oc_phys *P)
{
    if (!P || P->dim < 1 || P->dim > 3 || P->nlig < 0 || P->nlig > OC_MAXLIG) return NULL;
    oc_ctx *c = (oc_ctx *)calloc(1, sizeof(oc_ctx));
    if (!c) return NULL;
    c->P = *P;
    for (int a = P->dim; a < 3; ++a) c->P.n[a] = 1;
    c->dof = P->nlig + 1;
    c->npts = (long)c->P.n[0] * c->P.n[1] * c->P.n[2];
    c->ncf = 2 * P->dim + 2;
    for (int a = 0; a < 3; ++a)
        for (int o = 0; o < 5; ++o) {
            const int n = c->P.n[a];
            c->nb[a][o] = (int *)malloc(sizeof(int) * (size_t)n);
            for (int i = 0; i < n; ++i) c->nb[a][o][i] = (((i + o - 2) % n) + n) % n;
        }
    const long nv = c->npts * c->dof;
    c->ug = dalloc(nv); c->G = dalloc(c->npts); c->grho = dalloc(c->npts);
    c->gU = dalloc(c->npts * (P->nlig > 0 ? P->nlig : 1));
    c->cf = dalloc(c->npts * c->ncf); c->Minv = dalloc(c->npts * c->dof * c->dof);
    c->dGv = dalloc(c->npts);
    c->r = dalloc(nv); c->z = dalloc(nv); c->w = dalloc(nv); c->F = dalloc(nv);
    c->Z = dalloc(nv); c->Zd = dalloc(nv);
    for (int i = 0; i < 4; ++i) c->Y[i] = dalloc(nv);
    c->un = dalloc(nv);
    c->part = dalloc(nv / OC_CHUNK + 2);
    return c;
}

/* flat point indices of the five stencil points of (i,j,k) along axis ax */
static inline void nbr(const oc_ctx *c, int ax, int i, int j, int k, long q[5])
{
    const long nx = c->P.n[0], ny = c->P.n[1];
    for (int o = 0; o < 5; ++o) {
        const int ii = ax == 0 ? c->nb[0][o][i] : i;
        const int jj = ax == 1 ? c->nb[1][o][j] : j;
        const int kk = ax == 2 ? c->nb[2][o][k] : k;
        q[o] = ii + nx * (jj + ny * kk);
    }
}

/* first / second derivative sums in the oracle's order (ksfd_oracle.py d1, d2) */
static inline double d1s(const double w[5], const double *a, long st, const long q[5])
{
    double out = 0.0;
    for (int o = 0; o < 5; ++o)
        if (w[o] != 0.0) out = out + w[o] * a[q[o] * st];
    return out;
}
static inline double d2s(const double w[5], const double *a, long st, const long q[5])
{
    double out = 0.0;
    for (int o = 0; o < 5; ++o) out = out + w[o] * a[q[o] * st];
    return out;
}

/* clamp (Derivatives.groom): NaN and values below the minimum become the minimum */
void oc_groom(const oc_ctx *c, double *u)
{
    const int dof = c->dof;
    const double rmin = c->P.rhomin, umin = c->P.Umin;
#pragma omp parallel for schedule(static)
    for (long p = 0; p < c->npts; ++p) {
        double *x = u + p * dof;
        x[0] = x[0] >= rmin ? x[0] : rmin;
        for (int l = 1; l < dof; ++l) x[l] = x[l] >= umin ? x[l] : umin;
    }
}

/* ug = clamp(u);  G at every point;  with_d: also dG/drho, dG/dU_l */
static void eval_G(oc_ctx *c, const double *u, int with_d)
{
    const oc_phys *P = &c->P;
    const int dof = c->dof, nlig = P->nlig;
    const double cc = P->maxscale * P->s2;
#pragma omp parallel for schedule(static)
    for (long p = 0; p < c->npts; ++p) {
        double *x = c->ug + p * dof;
        const double *s = u + p * dof;
        x[0] = s[0] >= P->rhomin ? s[0] : P->rhomin;
        for (int l = 1; l < dof; ++l) x[l] = s[l] >= P->Umin ? s[l] : P->Umin;
        const double rho = x[0];
        double G = P->s2 * log(rho);
        double sUg[OC_MAXLIG];
        for (int g = 0; g < P->ngroups; ++g) {
            double sU = 0.0;
            int any = 0;
            for (int l = 0; l < nlig; ++l)
                if (P->lig_group[l] == g) {
                    sU = sU + P->lig_w[l] * x[l + 1];
                    any = 1;
                }
            sUg[g] = sU;
            if (any) G = G - P->beta[g] * log(P->alpha[g] + sU);
        }
        const double th = tanh((rho - P->rhomax) / P->cushion);
        double cap = cc * (th + 1.0);
        if (P->witch) cap = cap * (rho / P->rhomax);
        c->G[p] = G + cap;
        if (with_d) {
            const double sech2 = 1.0 - th * th;
            double dcap;
            if (P->witch)
                dcap = cc * (sech2 / P->cushion * (rho / P->rhomax) + (th + 1.0) / P->rhomax);
            else
                dcap = cc * sech2 / P->cushion;
            c->grho[p] = P->s2 / rho + dcap;
            for (int l = 0; l < nlig; ++l) {
                const int g = P->lig_group[l];
                c->gU[(long)l * c->npts + p] = -P->beta[g] * P->lig_w[l] / (P->alpha[g] + sUg[g]);
            }
        }
    }
}

/* out = f(ug) from c->ug, c->G (the body of dfdt after the ghost exchange and the clamp);
   sign/udot: out = udot - f when udot != NULL */
static void eval_f(const oc_ctx *c, const double *src, const double *udot, double *out)
{
    const oc_phys *P = &c->P;
    const int dof = c->dof, nlig = P->nlig, dim = P->dim;
    const int nx = P->n[0], ny = P->n[1], nz = P->n[2];
#pragma omp parallel for collapse(2) schedule(static)
    for (int k = 0; k < nz; ++k)
        for (int j = 0; j < ny; ++j)
            for (int i = 0; i < nx; ++i) {
                const long p = i + (long)nx * (j + (long)ny * k);
                long q[3][5];
                for (int a = 0; a < dim; ++a) nbr(c, a, i, j, k, q[a]);
                double acc = 0.0, lap = 0.0;
                for (int a = 0; a < dim; ++a) {
                    acc = acc + d1s(P->w1[a], c->ug, dof, q[a]) * d1s(P->w1[a], c->G, 1, q[a]);
                    lap = lap + d2s(P->w2[a], c->G, 1, q[a]);
                }
                const double rho0 = c->ug[p * dof];
                double f[OC_MAXDOF];
                f[0] = acc + rho0 * lap;
                for (int l = 0; l < nlig; ++l) {
                    double lapU = 0.0;
                    for (int a = 0; a < dim; ++a)
                        lapU = lapU + d2s(P->w2[a], c->ug + l + 1, dof, q[a]);
                    f[l + 1] = (-P->lig_gamma[l] * c->ug[p * dof + l + 1] + P->lig_s[l] * rho0) +
                               P->lig_D[l] * lapU;
                }
                for (int d = 0; d < dof; ++d) {
                    double v = f[d];
                    if (src) v = v + src[p * dof + d];
                    out[p * dof + d] = udot ? udot[p * dof + d] - v : v;
                }
            }
}

/* f(u); src (same layout as u) is added when not NULL (ksfdsym.py:930-936) */
void oc_dfdt(oc_ctx *c, const double *u, const double *src, double *out)
{
    eval_G(c, u, 0);
    eval_f(c, src, NULL, out);
}

/* F = udot - f(u) */
void oc_ifunction(oc_ctx *c, const double *u, const double *udot, const double *src, double *F)
{
    eval_G(c, u, 0);
    eval_f(c, src, udot, F);
}

/* vmax[a] = max |dG/dx_a| over the grid (velocity = grad G) */
void oc_velocity_max(oc_ctx *c, const double *u, double vmax[3])
{
    const oc_phys *P = &c->P;
    const int dim = P->dim, nx = P->n[0], ny = P->n[1], nz = P->n[2];
    eval_G(c, u, 0);
    double m0 = 0.0, m1 = 0.0, m2 = 0.0;
#pragma omp parallel for collapse(2) schedule(static) reduction(max : m0, m1, m2)
    for (int k = 0; k < nz; ++k)
        for (int j = 0; j < ny; ++j)
            for (int i = 0; i < nx; ++i) {
                long q[5];
                for (int a = 0; a < dim; ++a) {
                    nbr(c, a, i, j, k, q);
                    const double v = fabs(d1s(P->w1[a], c->G, 1, q));
                    if (a == 0) m0 = v > m0 ? v : m0;
                    if (a == 1) m1 = v > m1 ? v : m1;
                    if (a == 2) m2 = v > m2 ? v : m2;
                }
            }
    vmax[0] = m0; vmax[1] = m1; vmax[2] = m2;
}

/* velocity field, (dim,) + n in F order (Derivatives.velocity) */
void oc_velocity(oc_ctx *c, const double *u, double *vel)
{
    const oc_phys *P = &c->P;
    const int dim = P->dim, nx = P->n[0], ny = P->n[1], nz = P->n[2];
    eval_G(c, u, 0);
#pragma omp parallel for collapse(2) schedule(static)
    for (int k = 0; k < nz; ++k)
        for (int j = 0; j < ny; ++j)
            for (int i = 0; i < nx; ++i) {
                const long p = i + (long)nx * (j + (long)ny * k);
                long q[5];
                for (int a = 0; a < dim; ++a) {
                    nbr(c, a, i, j, k, q);
                    vel[p * dim + a] = d1s(P->w1[a], c->G, 1, q);
                }
            }
}

/* in-place inverse of a small dense matrix (partial pivoting); returns 0 if singular */
static int inv_small(int n, double *A)
{
    double B[OC_MAXDOF][2 * OC_MAXDOF];
    for (int r = 0; r < n; ++r)
        for (int s = 0; s < n; ++s) {
            B[r][s] = A[r * n + s];
            B[r][n + s] = r == s ? 1.0 : 0.0;
        }
    for (int col = 0; col < n; ++col) {
        int piv = col;
        for (int r = col + 1; r < n; ++r)
            if (fabs(B[r][col]) > fabs(B[piv][col])) piv = r;
        if (B[piv][col] == 0.0) return 0;
        if (piv != col)
            for (int s = 0; s < 2 * n; ++s) {
                const double t = B[col][s];
                B[col][s] = B[piv][s];
                B[piv][s] = t;
            }
        const double d = 1.0 / B[col][col];
        for (int s = 0; s < 2 * n; ++s) B[col][s] *= d;
        for (int r = 0; r < n; ++r)
            if (r != col) {
                const double m = B[r][col];
                if (m != 0.0)
                    for (int s = 0; s < 2 * n; ++s) B[r][s] -= m * B[col][s];
            }
    }
    for (int r = 0; r < n; ++r)
        for (int s = 0; s < n; ++s) A[r * n + s] = B[r][n + s];
    return 1;
}

/* linearise at u_lin: A = shift*I - df/du(u_lin) (implicitIJ).  Caches G's derivatives, the
   v-independent stencil sums and the inverse diagonal blocks.  Returns 0, or 1 if a diagonal
   block is singular. */
int oc_jvp_setup(oc_ctx *c, const double *u_lin, double shift)
{
    const oc_phys *P = &c->P;
    const int dof = c->dof, nlig = P->nlig, dim = P->dim, ncf = c->ncf;
    const int nx = P->n[0], ny = P->n[1], nz = P->n[2];
    eval_G(c, u_lin, 1);
    c->shift = shift;
    int bad = 0;
#pragma omp parallel for collapse(2) schedule(static) reduction(| : bad)
    for (int k = 0; k < nz; ++k)
        for (int j = 0; j < ny; ++j)
            for (int i = 0; i < nx; ++i) {
                const long p = i + (long)nx * (j + (long)ny * k);
                double *cf = c->cf + p * ncf;
                long q[5];
                double lap = 0.0;
                const double rho0 = c->ug[p * dof];
                double B[OC_MAXDOF * OC_MAXDOF];
                for (int s = 0; s < dof * dof; ++s) B[s] = 0.0;
                double j00 = 0.0, jc = 0.0;     /* d f_rho / d rho(p), common factor of the G terms */
                for (int a = 0; a < dim; ++a) {
                    nbr(c, a, i, j, k, q);
                    const double dG = d1s(P->w1[a], c->G, 1, q);
                    const double dr = d1s(P->w1[a], c->ug, dof, q);
                    lap = lap + d2s(P->w2[a], c->G, 1, q);
                    cf[a] = dG;
                    cf[dim + a] = dr;
                    const double w1c = P->w1[a][2], w2c = P->w2[a][2];
                    j00 += w1c * dG;
                    jc += dr * w1c + rho0 * w2c;
                }
                cf[2 * dim] = lap;
                cf[2 * dim + 1] = rho0;
                B[0] = shift - (j00 + jc * c->grho[p] + lap);
                for (int l = 0; l < nlig; ++l) {
                    double w2sum = 0.0;
                    for (int a = 0; a < dim; ++a) w2sum += P->w2[a][2];
                    B[l + 1] = -(jc * c->gU[(long)l * c->npts + p]);
                    B[(l + 1) * dof] = -P->lig_s[l];
                    B[(l + 1) * dof + l + 1] = shift - (-P->lig_gamma[l] + P->lig_D[l] * w2sum);
                }
                if (!inv_small(dof, B)) bad |= 1;
                memcpy(c->Minv + p * dof * dof, B, sizeof(double) * dof * dof);
            }
    c->have_jac = !bad;
    return bad;
}

/* diagonal block (not inverted) of point p is not stored; the tests get M^-1 */
void oc_get_minv(const oc_ctx *c, double *out)
{
    memcpy(out, c->Minv, sizeof(double) * (size_t)(c->npts * c->dof * c->dof));
}

/* z = M^-1 r (point-block Jacobi) */
void oc_pc_apply(const oc_ctx *c, const double *r, double *z)
{
    const int dof = c->dof;
#pragma omp parallel for schedule(static)
    for (long p = 0; p < c->npts; ++p) {
        const double *M = c->Minv + p * dof * dof;
        double t[OC_MAXDOF];
        for (int a = 0; a < dof; ++a) {
            double s = 0.0;
            for (int b = 0; b < dof; ++b) s += M[a * dof + b] * r[p * dof + b];
            t[a] = s;
        }
        for (int a = 0; a < dof; ++a) z[p * dof + a] = t[a];
    }
}

/* out = A v = (shift*I - J(u_lin)) v, matrix-free (ksfd_oracle.py jvp_ghosted) */
void oc_jvp(oc_ctx *c, const double *v, double *out)
{
    const oc_phys *P = &c->P;
    const int dof = c->dof, nlig = P->nlig, dim = P->dim, ncf = c->ncf;
    const int nx = P->n[0], ny = P->n[1], nz = P->n[2];
    const double shift = c->shift;
#pragma omp parallel for schedule(static)
    for (long p = 0; p < c->npts; ++p) {
        double s = c->grho[p] * v[p * dof];
        for (int l = 0; l < nlig; ++l) s = s + c->gU[(long)l * c->npts + p] * v[p * dof + l + 1];
        c->dGv[p] = s;
    }
#pragma omp parallel for collapse(2) schedule(static)
    for (int k = 0; k < nz; ++k)
        for (int j = 0; j < ny; ++j)
            for (int i = 0; i < nx; ++i) {
                const long p = i + (long)nx * (j + (long)ny * k);
                const double *cf = c->cf + p * ncf;
                long q[3][5];
                for (int a = 0; a < dim; ++a) nbr(c, a, i, j, k, q[a]);
                double acc = 0.0, lapdG = 0.0;
                for (int a = 0; a < dim; ++a) {
                    acc = acc + d1s(P->w1[a], v, dof, q[a]) * cf[a] +
                          cf[dim + a] * d1s(P->w1[a], c->dGv, 1, q[a]);
                    lapdG = lapdG + d2s(P->w2[a], c->dGv, 1, q[a]);
                }
                const double v0 = v[p * dof];
                const double Jv0 = (acc + v0 * cf[2 * dim]) + cf[2 * dim + 1] * lapdG;
                out[p * dof] = shift * v0 - Jv0;
                for (int l = 0; l < nlig; ++l) {
                    double lapV = 0.0;
                    for (int a = 0; a < dim; ++a) lapV = lapV + d2s(P->w2[a], v + l + 1, dof, q[a]);
                    const double vl = v[p * dof + l + 1];
                    const double JvU = (-P->lig_gamma[l] * vl + P->lig_s[l] * v0) + P->lig_D[l] * lapV;
                    out[p * dof + l + 1] = shift * vl - JvU;
                }
            }
}

static double dot(oc_ctx *c, long n, const double *a, const double *b)
{
    const long nch = (n + OC_CHUNK - 1) / OC_CHUNK;
#pragma omp parallel for schedule(static)
    for (long k = 0; k < nch; ++k) {
        const long e = (k + 1) * OC_CHUNK < n ? (k + 1) * OC_CHUNK : n;
        double s = 0.0;
        for (long i = k * OC_CHUNK; i < e; ++i) s += a[i] * b[i];
        c->part[k] = s;
    }
    double s = 0.0;
    for (long k = 0; k < nch; ++k) s += c->part[k];
    return s;
}

/* r -= w, returns ||r||^2 (chunked like dot) */
static double sub_norm2(oc_ctx *c, long n, double *r, const double *w)
{
    const long nch = (n + OC_CHUNK - 1) / OC_CHUNK;
#pragma omp parallel for schedule(static)
    for (long k = 0; k < nch; ++k) {
        const long e = (k + 1) * OC_CHUNK < n ? (k + 1) * OC_CHUNK : n;
        double s = 0.0;
        for (long i = k * OC_CHUNK; i < e; ++i) {
            r[i] -= w[i];
            s += r[i] * r[i];
        }
        c->part[k] = s;
    }
    double s = 0.0;
    for (long k = 0; k < nch; ++k) s += c->part[k];
    return s;
}

static void axpy(long n, double a, const double *x, double *y)
{
#pragma omp parallel for schedule(static)
    for (long i = 0; i < n; ++i) y[i] += a * x[i];
}

/* right-preconditioned GMRES(m) on A M^-1, continuing from x; returns iterations, <0 on failure */
static int gmres(oc_ctx *c, const double *b, double *x, double tol, int m, int max_it, double *rnorm)
{
    const long n = c->npts * c->dof;
    if (m > 64) m = 64;
    if (c->vcap < m + 1) {
        free(c->V);
        c->V = dalloc(n * (m + 1));
        c->vcap = m + 1;
    }
    if (!c->V) return -1;
    double H[(64 + 1) * 64], cs[64], sn[64], g[65], y[64];
    int its = 0;
    for (;;) {
        /* r = b - A x */
        oc_jvp(c, x, c->w);
        double *V0 = c->V;
#pragma omp parallel for schedule(static)
        for (long i = 0; i < n; ++i) V0[i] = b[i] - c->w[i];
        double beta = sqrt(dot(c, n, V0, V0));
        *rnorm = beta;
        if (beta <= tol || its >= max_it) return its;
        {
            const double ib = 1.0 / beta;
#pragma omp parallel for schedule(static)
            for (long i = 0; i < n; ++i) V0[i] *= ib;
        }
        memset(g, 0, sizeof(g));
        g[0] = beta;
        int kk = 0;
        for (int k = 0; k < m && its < max_it; ++k) {
            double *vk = c->V + (long)k * n, *wv = c->V + (long)(k + 1) * n;
            oc_pc_apply(c, vk, c->z);
            oc_jvp(c, c->z, wv);
            for (int i = 0; i <= k; ++i) {          /* modified Gram-Schmidt */
                const double h = dot(c, n, wv, c->V + (long)i * n);
                H[i * 64 + k] = h;
                axpy(n, -h, c->V + (long)i * n, wv);
            }
            const double hn = sqrt(dot(c, n, wv, wv));
            H[(k + 1) * 64 + k] = hn;
            if (hn > 0.0) {
                const double ih = 1.0 / hn;
#pragma omp parallel for schedule(static)
                for (long i = 0; i < n; ++i) wv[i] *= ih;
            }
            for (int i = 0; i < k; ++i) {           /* earlier rotations */
                const double t = cs[i] * H[i * 64 + k] + sn[i] * H[(i + 1) * 64 + k];
                H[(i + 1) * 64 + k] = -sn[i] * H[i * 64 + k] + cs[i] * H[(i + 1) * 64 + k];
                H[i * 64 + k] = t;
            }
            const double a = H[k * 64 + k], bb = H[(k + 1) * 64 + k], rr = hypot(a, bb);
            cs[k] = rr > 0.0 ? a / rr : 1.0;
            sn[k] = rr > 0.0 ? bb / rr : 0.0;
            H[k * 64 + k] = rr;
            H[(k + 1) * 64 + k] = 0.0;
            g[k + 1] = -sn[k] * g[k];
            g[k] = cs[k] * g[k];
            ++its;
            kk = k + 1;
            *rnorm = fabs(g[k + 1]);
            if (*rnorm <= tol || hn == 0.0) break;
        }
        for (int i = kk - 1; i >= 0; --i) {
            double s = g[i];
            for (int j2 = i + 1; j2 < kk; ++j2) s -= H[i * 64 + j2] * y[j2];
            y[i] = s / H[i * 64 + i];
        }
        /* x += M^-1 (V y) */
        memset(c->w, 0, sizeof(double) * (size_t)n);
        for (int i = 0; i < kk; ++i) axpy(n, y[i], c->V + (long)i * n, c->w);
        oc_pc_apply(c, c->w, c->z);
        axpy(n, 1.0, c->z, x);
        /* converged by the recurrence: confirmed on the true residual at the top of the loop */
        if (its >= max_it && *rnorm > tol) {
            oc_jvp(c, x, c->w);
#pragma omp parallel for schedule(static)
            for (long i = 0; i < n; ++i) c->z[i] = b[i] - c->w[i];
            *rnorm = sqrt(dot(c, n, c->z, c->z));
            return *rnorm <= tol ? its : -its;
        }
    }
}

/* Solve A x = b (A from oc_jvp_setup) to ||b - A x|| <= max(rtol*||b||, atol).
   ksp_type 0: Richardson sweeps x += M^-1 r, r -= A M^-1 r while a sweep contracts the
   residual to <= 0.65 of the previous one, else GMRES(restart) from the iterate;
   1: Richardson only; 2: GMRES only.  info[0] = iterations, info[1] = 1 if GMRES ran;
   norms[0] = ||b||, norms[1] = final residual norm.  Returns 0, 1 = not converged. */
int oc_solve(oc_ctx *c, const double *b, double *x, double rtol, double atol, int max_it,
             int restart, int ksp_type, int info[2], double norms[2])
{
    const long n = c->npts * c->dof;
    const int dof = c->dof;
    if (max_it <= 0) max_it = 10000;
    if (restart <= 0) restart = 30;
    double *r = c->r, *z = c->z;
    memset(x, 0, sizeof(double) * (size_t)n);
    memcpy(r, b, sizeof(double) * (size_t)n);
    const double bn = sqrt(dot(c, n, b, b));
    const double tol = fmax(rtol * bn, atol);
    double rn = bn;
    int its = 0, used_gmres = 0;
    norms[0] = bn;
    if (ksp_type != 2) {
        while (rn > tol && its < max_it) {
            /* z = M^-1 r ; x += z */
#pragma omp parallel for schedule(static)
            for (long p = 0; p < c->npts; ++p) {
                const double *M = c->Minv + p * dof * dof;
                for (int a = 0; a < dof; ++a) {
                    double s = 0.0;
                    for (int bb = 0; bb < dof; ++bb) s += M[a * dof + bb] * r[p * dof + bb];
                    z[p * dof + a] = s;
                    x[p * dof + a] += s;
                }
            }
            oc_jvp(c, z, c->w);
            const double rnew = sqrt(sub_norm2(c, n, r, c->w));
            ++its;
            const int slow = !(rnew <= 0.65 * rn);
            rn = rnew;
            if (!(rn == rn)) break;
            if (slow && ksp_type == 0 && rn > tol) {
                used_gmres = 1;
                break;
            }
        }
    } else {
        used_gmres = 1;
    }
    if (used_gmres) {
        if (!(rn == rn) || rn > bn) memset(x, 0, sizeof(double) * (size_t)n);
        const int g = gmres(c, b, x, tol, restart, max_it - its, &rn);
        its += g < 0 ? -g : g;
    }
    info[0] = its;
    info[1] = used_gmres;
    norms[1] = rn;
    return (rn <= tol) ? 0 : 1;
}

/* One ROSW step as PETSc's TSStep_RosW with -snes_type ksponly (ksfd_oracle.py rosw_step):
   per stage Z = u + sum_j At[i][j] Y_j, Zdot = sum_j Gi[i][j]/h Y_j, F = Zdot - f(Z); the
   matrix shift*I - df/du, shift = 1/(h*gamma), is set up once at stage 0; A Y_i = -F.
   unew / uemb = completion with bt / bet.  info[0] = linear iterations, info[1] = solves in
   which GMRES ran.  Returns 0, or 1 + the stage whose solve failed. */
int oc_rosw_step(oc_ctx *c, const double *u, double h, const oc_tableau *T, double rtol, double atol,
                 int max_it, int restart, int ksp_type, double *unew, double *uemb, int info[2])
{
    const long n = c->npts * c->dof;
    info[0] = info[1] = 0;
    for (int i = 0; i < 4; ++i) {
        double *Z = c->Z, *Zd = c->Zd;
        memcpy(Z, u, sizeof(double) * (size_t)n);
        memset(Zd, 0, sizeof(double) * (size_t)n);
        for (int j = 0; j < i; ++j) {
            const double a = T->At[i][j], g = T->Gi[i][j] / h;
            const double *Y = c->Y[j];
#pragma omp parallel for schedule(static)
            for (long q = 0; q < n; ++q) {
                Z[q] = Z[q] + a * Y[q];
                Zd[q] = Zd[q] + g * Y[q];
            }
        }
        oc_ifunction(c, Z, Zd, NULL, c->F);
        if (i == 0 && oc_jvp_setup(c, Z, 1.0 / (h * T->gamma))) return 1;
#pragma omp parallel for schedule(static)
        for (long q = 0; q < n; ++q) c->F[q] = -c->F[q];
        int si[2];
        double nn[2];
        const int rc = oc_solve(c, c->F, c->Y[i], rtol, atol, max_it, restart, ksp_type, si, nn);
        info[0] += si[0];
        info[1] += si[1];
        if (rc) return 1 + i;
    }
    if (unew) memcpy(unew, u, sizeof(double) * (size_t)n);
    if (uemb) memcpy(uemb, u, sizeof(double) * (size_t)n);
    for (int j = 0; j < 4; ++j) {
        const double b = T->bt[j], be = T->bet[j];
        const double *Y = c->Y[j];
#pragma omp parallel for schedule(static)
        for (long q = 0; q < n; ++q) {
            if (unew) unew[q] = unew[q] + b * Y[q];
            if (uemb) uemb[q] = uemb[q] + be * Y[q];
        }
    }
    return 0;
}

/* Backward Euler with -snes_type ksponly: one Newton step from u_n (ksfd_oracle.py beuler_step):
   (I/h - df/du(u)) y = f(u), unew = u + y.  Returns 0, 1 = solve failed. */
int oc_beuler_step(oc_ctx *c, const double *u, double h, double rtol, double atol, int max_it,
                   int restart, int ksp_type, double *unew, int info[2])
{
    const long n = c->npts * c->dof;
    oc_dfdt(c, u, NULL, c->F);
    if (oc_jvp_setup(c, u, 1.0 / h)) return 1;
    double nn[2];
    if (oc_solve(c, c->F, c->Y[0], rtol, atol, max_it, restart, ksp_type, info, nn)) return 1;
    const double *Y = c->Y[0];
#pragma omp parallel for schedule(static)
    for (long q = 0; q < n; ++q) unew[q] = u[q] + Y[q];
    return 0;
}

/* stage increment Y_j of the last step (tests) */
void oc_get_stage(const oc_ctx *c, int j, double *out)
{
    memcpy(out, c->Y[j], sizeof(double) * (size_t)(c->npts * c->dof));
}

/* the body of the reference's step loop (KSFD/ksfdts.py:202-228) without noise and monitors:
   clamp in place, TS.step, CFL maxima of the result.  u is advanced in place. */
int oc_ts_step(oc_ctx *c, double *u, double h, const oc_tableau *T, double rtol, int ksp_type,
               double vmax[3], int info[2])
{
    oc_groom(c, u);
    double *un = c->un;
    const int rc = oc_rosw_step(c, u, h, T, rtol, 0.0, 2000, 30, ksp_type, un, NULL, info);
    if (!rc) {
        memcpy(u, un, sizeof(double) * (size_t)(c->npts * c->dof));
        oc_velocity_max(c, u, vmax);
    }
    return rc;
}
