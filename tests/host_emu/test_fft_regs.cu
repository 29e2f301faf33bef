// Host-side check of the register FFTs against a naive DFT (compiled with nvcc, run on the CPU).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include "../../spectrogram_inversion_b200/csrc/fft_regs.cuh"
using namespace specinv;

template <int N, bool INV, typename F>
double check(F f) {
    float2 a[N]; double ref_re[N], ref_im[N], in_re[N], in_im[N];
    for (int i = 0; i < N; ++i) { a[i].x = (float)rand() / RAND_MAX - 0.5f; a[i].y = (float)rand() / RAND_MAX - 0.5f; in_re[i] = a[i].x; in_im[i] = a[i].y; }
    for (int k = 0; k < N; ++k) {
        double sr = 0, si = 0;
        for (int n = 0; n < N; ++n) {
            double ang = (INV ? 2 : -2) * M_PI * k * n / N;
            sr += in_re[n] * cos(ang) - in_im[n] * sin(ang);
            si += in_re[n] * sin(ang) + in_im[n] * cos(ang);
        }
        ref_re[k] = sr; ref_im[k] = si;
    }
    f(a);
    double err = 0;
    for (int k = 0; k < N; ++k) err = fmax(err, fmax(fabs(a[k].x - ref_re[k]), fabs(a[k].y - ref_im[k])));
    return err;
}

int main() {
    double e = 0;
    e = fmax(e, check<8, false>([](float2* a) { fft8<false>(a); }));
    e = fmax(e, check<8, true>([](float2* a) { fft8<true>(a); }));
    e = fmax(e, check<16, false>([](float2* a) { fft16<false>(a); }));
    e = fmax(e, check<16, true>([](float2* a) { fft16<true>(a); }));
    printf("max err %.3e\n", e);
    return e < 5e-6 ? 0 : 1;
}
