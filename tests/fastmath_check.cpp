// Host accuracy check of ksfd_b200/csrc/fastmath.cuh against long double
// (x86 80-bit: 64-bit mantissa).  Prints max errors in ulps; driven by
// tests/test_fastmath.py.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>

#include "fastmath.cuh"

static double ulp_of(double y)
{
    y = std::fabs(y);
    if (y == 0) return 4.9e-324;
    int e;
    std::frexp(y, &e);
    return std::ldexp(1.0, e - 53);
}

int main()
{
    static double logt[256], expt[64];
    for (int i = 0; i < 256; ++i) std::memcpy(&logt[i], &KSFD_LOG_TAB_BITS[i], 8);
    for (int i = 0; i < 64; ++i) std::memcpy(&expt[i], &KSFD_EXP_TAB_BITS[i], 8);
    FastTabs T{logt, expt};
    const FastK K = fastk_default();
    std::mt19937_64 rng(12345);
    std::uniform_real_distribution<double> U(0.0, 1.0);
    double elog = 0, elog_abs = 0, eexp = 0, elgs = 0, ercp = 0;
    const int N = 4000000;
    for (int n = 0; n < N; ++n) {
        // log: log-uniform over [1e-300, 1e300], and dense near the physics range
        double x = (n & 1) ? std::exp((U(rng) - 0.5) * 1380.0) : 1000.0 + 30000.0 * U(rng);
        if (n % 16 == 3) x = 1.0 + (U(rng) - 0.5) * 0.02;
        const double y = ksfd_log(x, K, T);
        const long double yr = logl((long double)x);
        const double err = (double)fabsl((long double)y - yr);
        const double allowed_extra = std::ldexp(1.0, -58);
        const double e1 = std::fmax(err - allowed_extra, 0.0) / ulp_of((double)yr);
        if (e1 > elog) elog = e1;
        if (err > elog_abs && std::fabs((double)yr) < 1e-3) elog_abs = err;
        // exp
        const double a = (U(rng) - 0.5) * 1414.0;
        const double ea = ksfd_exp(a, K, T);
        const long double er = expl((long double)a);
        const double e2 = (double)(fabsl((long double)ea - er)) / ulp_of((double)er);
        if (e2 > eexp) eexp = e2;
        // logistic
        const double b = (U(rng) - 0.5) * 120.0;
        const double lb = ksfd_logistic(b, K, T);
        const long double lr = 1.0L / (1.0L + expl((long double)b));
        const double e3 = (double)(fabsl((long double)lb - lr)) / ulp_of((double)lr);
        if (e3 > elgs) elgs = e3;
        // rcp
        const double d = std::exp((U(rng) - 0.5) * 1000.0);
        const double e4 = (double)(fabsl((long double)ksfd_rcp(d) - 1.0L / (long double)d)) /
                          ulp_of(1.0 / d);
        if (e4 > ercp) ercp = e4;
    }
    // special cases
    int bad = 0;
    const double inf = INFINITY;
    if (!(ksfd_log(0.0, K, T) == -inf)) bad |= 1;
    if (!(ksfd_log(inf, K, T) == inf)) bad |= 2;
    if (!std::isnan(ksfd_log(-1.0, K, T))) bad |= 4;
    if (!std::isnan(ksfd_log(NAN, K, T))) bad |= 8;
    const double sub = 1e-310;
    if (std::fabs(ksfd_log(sub, K, T) - std::log(sub)) > 2 * ulp_of(std::log(sub))) bad |= 16;
    if (std::fabs(ksfd_log(1.0, K, T)) > std::ldexp(1.0, -58)) bad |= 32;
    if (ksfd_logistic(1e300, K, T) != 0.0 && ksfd_logistic(1e300, K, T) > 1e-300) bad |= 64;
    if (ksfd_logistic(-1e300, K, T) != 1.0) bad |= 128;
    printf("log_ulp %.4f log_abs_small %.3e exp_ulp %.4f logistic_ulp %.4f rcp_ulp %.4f special %d\n",
           elog, elog_abs, eexp, elgs, ercp, bad);
    return 0;
}
