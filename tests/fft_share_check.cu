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
