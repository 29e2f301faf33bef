// Host-side check of the mixed-radix shared-memory FFT passes (csrc/mixed_radix.cuh): the team's threads are
// emulated one after the other, pass by pass, and the result is compared with a double-precision naive DFT at the
// digit-reversed positions; the inverse passes must bring the frames back (times M).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../spectrogram_inversion_b200/csrc/mixed_radix.cuh"
using namespace specinv;

template <typename T, typename C>
double run(int M, int tw_n, int nt) {
    mr::Plan p;
    if (!mr::make_plan(M, tw_n, &p, sizeof(T) == 4)) { printf("M=%d: no plan\n", M); return 1e9; }
    const int nf = 3, Mp = mr::padded_len(M);
    std::vector<C> tw(tw_n), wb((size_t)nf * Mp);
    for (int j = 0; j < tw_n; ++j) { tw[j].x = (T)cos(2 * M_PI * j / tw_n); tw[j].y = (T)(-sin(2 * M_PI * j / tw_n)); }
    std::vector<double> re((size_t)nf * M), im((size_t)nf * M);
    for (int f = 0; f < nf; ++f)
        for (int n = 0; n < M; ++n) {
            re[f * M + n] = (double)rand() / RAND_MAX - 0.5; im[f * M + n] = (double)rand() / RAND_MAX - 0.5;
            wb[f * Mp + mr::padidx(n)].x = (T)re[f * M + n]; wb[f * Mp + mr::padidx(n)].y = (T)im[f * M + n];
        }
    for (int s = 0; s < p.nst; ++s)
        for (int tid = 0; tid < nt; ++tid) mr::pass<T, false>(wb.data(), nf, Mp, p, s, tw.data(), tid, nt);
    double err = 0;
    std::vector<char> seen(M, 0);
    for (int f = 0; f < nf; ++f)
        for (int k = 0; k < M; ++k) {
            double sr = 0, si = 0;
            for (int n = 0; n < M; ++n) {
                const double ang = -2 * M_PI * ((long long)k * n % M) / M;
                sr += re[f * M + n] * cos(ang) - im[f * M + n] * sin(ang);
                si += re[f * M + n] * sin(ang) + im[f * M + n] * cos(ang);
            }
            const int pos = mr::mr_position(p, k);
            if (f == 0) { if (seen[pos]) return 1e9; seen[pos] = 1; }
            const C v = wb[f * Mp + mr::padidx(pos)];
            err = fmax(err, fmax(fabs(v.x - sr), fabs(v.y - si)) / sqrt((double)M));
        }
    for (int s = p.nst - 1; s >= 0; --s)
        for (int tid = 0; tid < nt; ++tid) mr::pass<T, true>(wb.data(), nf, Mp, p, s, tw.data(), tid, nt);
    for (int f = 0; f < nf; ++f)
        for (int n = 0; n < M; ++n) {
            const C v = wb[f * Mp + mr::padidx(n)];
            err = fmax(err, fmax(fabs(v.x / M - re[f * M + n]), fabs(v.y / M - im[f * M + n])));
        }
    return err;
}

int main() {
    const int Ms[] = {8, 16, 26, 48, 55, 56, 60, 110, 125, 200, 250, 256, 300, 500, 512, 768, 1000, 1001, 2048, 4096, 3 * 5 * 7 * 11, 13 * 13 * 8};
    double ef = 0, ed = 0;
    for (int M : Ms) {
        for (int mult = 1; mult <= 2; ++mult) {
            const double a = run<float, float2>(M, mult * M, 32), b = run<double, double2>(M, mult * M, 96);
            printf("M=%5d tw_n=%5d  fp32 %.2e  fp64 %.2e\n", M, mult * M, a, b);
            ef = fmax(ef, a); ed = fmax(ed, b);
        }
    }
    mr::Plan p;
    if (mr::make_plan(17, 34, &p) || mr::make_plan(2 * 101, 4 * 101, &p)) { printf("plan for a large prime?\n"); return 1; }
    printf("max err fp32 %.3e fp64 %.3e\n", ef, ed);
    return (ef < 3e-6 && ed < 1e-14) ? 0 : 1;
}
