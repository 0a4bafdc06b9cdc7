#include <cstdio>
#include "fftpc.cuh"
int main() {
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
