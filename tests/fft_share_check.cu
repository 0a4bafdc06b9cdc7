#include <cstdio>
#include "fftpc.cuh"
// with arguments: dump an index map for the numpy emulation
//   pack nloc dof PS P       -> fft_pack_index(e) for e < nloc*dof*PS
//   transpose NL dof nsq     -> fft_transpose_index(e) for e < NL*dof*nsq
#include <cstdlib>
#include <cstring>
int main(int argc, char **argv) {
    if (argc == 6 && !strcmp(argv[1], "pack")) {
        const int nloc = atoi(argv[2]), dof = atoi(argv[3]), P = atoi(argv[5]);
        const long long PS = atoll(argv[4]);
        for (long long e = 0; e < (long long)nloc * dof * PS; ++e)
            printf("%lld\n", fft_pack_index(nloc, dof, PS, P, e));
        return 0;
    }
    if (argc == 4 && !strcmp(argv[1], "symbol")) {
        // file: 9 int32 (dof nlig n0 n1 n2 dist s0 nsq NL), doubles shift, c2[3],
        // s[7], gamma[7], D[7], means[9], then the spectra (dof * count complex); solved in
        // place with fft_symbol_elem and written to argv[3]
        FILE *f = fopen(argv[2], "rb");
        if (!f) return 2;
        int hd[9];
        double dd[1 + 3 + 3 * KSFD_MAX_LIGANDS + KSFD_MAX_LIGANDS + 2];
        if (fread(hd, sizeof(int), 9, f) != 9) return 2;
        if (fread(dd, sizeof(double), sizeof(dd) / sizeof(double), f) != sizeof(dd) / sizeof(double)) return 2;
        FftSym S{};
        S.dof = hd[0]; S.nlig = hd[1]; S.n0 = hd[2]; S.n1 = hd[3]; S.n2 = hd[4];
        S.dist = hd[5]; S.s0 = hd[6]; S.nsq = hd[7]; S.NL = hd[8];
        S.shift = dd[0];
        for (int a = 0; a < 3; ++a) S.c2[a] = dd[1 + a];
        for (int l = 0; l < KSFD_MAX_LIGANDS; ++l) {
            S.s[l] = dd[4 + l];
            S.gamma[l] = dd[4 + KSFD_MAX_LIGANDS + l];
            S.D[l] = dd[4 + 2 * KSFD_MAX_LIGANDS + l];
        }
        const double *means = dd + 4 + 3 * KSFD_MAX_LIGANDS;
        const long long nk = fft_symbol_count(S);
        double2 *spec = (double2 *)malloc(sizeof(double2) * nk * S.dof);
        if (fread(spec, sizeof(double2), nk * S.dof, f) != (size_t)(nk * S.dof)) return 2;
        fclose(f);
        for (long long e = 0; e < nk; ++e) fft_symbol_elem(S, means, spec, e);
        f = fopen(argv[3], "wb");
        fwrite(spec, sizeof(double2), nk * S.dof, f);
        fclose(f);
        return 0;
    }
    if (argc == 5 && !strcmp(argv[1], "transpose")) {
        const int NL = atoi(argv[2]), dof = atoi(argv[3]);
        const long long nsq = atoll(argv[4]);
        for (long long e = 0; e < (long long)NL * dof * nsq; ++e)
            printf("%lld\n", fft_transpose_index(NL, dof, nsq, e));
        return 0;
    }
    long long bad = 0, n = 0;
    for (int P = 1; P <= 16; ++P)
        for (long long PS = P; PS <= 700; PS += (PS < 40 ? 1 : 37)) {
            long long prev = -1;
            for (long long s = 0; s < PS; ++s) {
                int q = fft_share_owner(PS, s, P);
                ++n;
                if (!(fft_share_start(PS, q, P) <= s && s < fft_share_start(PS, q + 1, P))) ++bad;
                if (q < prev) ++bad;
                prev = q;
            }
            if (fft_share_start(PS, P, P) != PS) ++bad;
        }
    printf("checked %lld, bad %lld\n", n, bad);
    return bad != 0;
}
