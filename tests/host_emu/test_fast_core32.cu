// Host emulation of the warp-per-frame pipeline of gl_fast_core32.cuh (n_fft = 2048): 32 lanes run
// sequentially, the two shuffle exchanges are emulated by reading the partner lane's array.
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../spectrogram_inversion_b200/csrc/gl_fast_core32.cuh"
using namespace specinv;
using namespace specinv::fast32;
typedef std::complex<double> cd;
static double frand() { return (double)rand() / RAND_MAX - 0.5; }

template <int OP>
int run() {
    std::vector<float> x(N), wa(N), ws(N);
    for (int i = 0; i < N; ++i) { x[i] = (float)(4 * frand()); wa[i] = (float)(0.5 - 0.5 * cos(2 * M_PI * i / N)); ws[i] = wa[i] / N; }
    std::vector<float2> tw(TBL), swa(TBL), twr(M);
    for (int n2 = 0; n2 < 32; ++n2)
        for (int k = 0; k < 32; ++k) {
            double ang = -2 * M_PI * ((n2 * k) % M) / (double)M;
            tw[n2 * ROW + k] = f2((float)cos(ang), (float)sin(ang));
            swa[n2 * ROW + k] = f2(0.5f * wa[64 * k + 2 * n2], 0.5f * wa[64 * k + 2 * n2 + 1]);
        }
    for (int k = 0; k < M; ++k) twr[k] = f2((float)cos(-2 * M_PI * k / N), (float)sin(-2 * M_PI * k / N));
    Tables tb{tw.data(), swa.data(), nullptr, twr.data()};
    std::vector<float2> s0_in(M), s0_out(M), s1_in(M), s1_out(M);
    float2 s0n_in, s0n_out, s1n_in, s1n_out;
    std::vector<float> mag(M); float magn = 3.f;
    for (int k = 0; k < M; ++k) { s0_in[k] = f2((float)(20 * frand()), (float)(20 * frand())); s1_in[k] = f2((float)(5 * frand()), (float)(5 * frand())); mag[k] = (float)(10 * fabs(frand())); }
    s0n_in = f2((float)(20 * frand()), (float)(20 * frand())); s1n_in = f2((float)frand(), (float)frand());
    FrameIO io{};
    io.s0_stage = s0_in.data(); io.s0_in_nyq = &s0n_in; io.s0_out = s0_out.data(); io.s0_out_nyq = &s0n_out;
    io.s1_in = s1_in.data(); io.s1_in_nyq = &s1n_in; io.s1_out = s1_out.data(); io.s1_out_nyq = &s1n_out;
    io.mag = mag.data(); io.mag_nyq = &magn; io.s0_nyq_val = s0n_in; io.mag_nyq_val = magn;
    const float coef = OP == fast::OP_GL ? 0.3f : 0.1f;
    io.coef = coef; io.coef2 = 1.f / (1.f + coef); io.owned = true;

    std::vector<float2> exch(TBL);
    static float2 v[32][32], A[32][32], Tmp[32][32];
    float hn[32];
    for (int l = 0; l < 32; ++l) {
        for (int n1 = 0; n1 < 32; ++n1) { float2 w = swa[l * ROW + n1]; v[l][n1] = f2(x[64 * n1 + 2 * l] * w.x, x[64 * n1 + 2 * l + 1] * w.y); }
        fast32::phase1_compute(l, v[l], tb);
        fast32::phase1_write(l, v[l], exch.data());
    }
    for (int l = 0; l < 32; ++l) phase2_read_fft(l, exch.data(), A[l]);
    for (int l = 0; l < 32; ++l) for (int j = 0; j < 32; ++j) Tmp[l][j] = A[(32 - l) & 31][j];     // shuffle exchange 1
    float ds = 0, es = 0;
    for (int l = 0; l < 32; ++l) { hn[l] = 0; pointwise_own<OP, true>(l, A[l], Tmp[l], tb, io, mag.data(), hn[l], ds, es); }
    for (int l = 0; l < 32; ++l) for (int j = 0; j < 32; ++j) Tmp[l][j] = A[(32 - l) & 31][j];     // shuffle exchange 2
    for (int l = 0; l < 32; ++l) pre_own(l, A[l], Tmp[l], tb, hn[l]);
    for (int l = 0; l < 32; ++l) phase2_ifft_write(l, A[l], exch.data());
    for (int l = 0; l < 32; ++l) fast32::phase3(l, v[l], tb, exch.data());

    std::vector<cd> s(M + 1), h(M + 1);
    double dref = 0, eref = 0, err_state = 0, err_x = 0;
    for (int k = 0; k <= M; ++k) {
        cd acc = 0;
        for (int n = 0; n < N; ++n) acc += (double)x[n] * (double)wa[n] * std::polar(1.0, -2 * M_PI * k * n / N);
        s[k] = acc;
        cd a0 = k < M ? cd(s0_in[k].x, s0_in[k].y) : cd(s0n_in.x, s0n_in.y);
        cd a1 = k < M ? cd(s1_in[k].x, s1_in[k].y) : cd(s1n_in.x, s1n_in.y);
        double m = k < M ? mag[k] : magn;
        dref += (std::abs(s[k]) - m) * (std::abs(s[k]) - m); eref += std::norm(s[k]);
        cd o0 = k < M ? cd(s0_out[k].x, s0_out[k].y) : cd(s0n_out.x, s0n_out.y);
        if (OP == fast::OP_GL) {
            cd q = s[k] - (double)coef * a0;
            err_state = fmax(err_state, std::abs(q - o0));
            h[k] = q * m / (std::abs(q) + 1e-16);
        } else {
            cd Z = ((double)coef * (a0 + a1) + s[k]) / (1.0 + coef);
            cd Un = a1 + a0 - Z;
            cd Xn = (Z - Un) * m / (std::abs(Z - Un) + 1e-16);
            cd o1 = k < M ? cd(s1_out[k].x, s1_out[k].y) : cd(s1n_out.x, s1n_out.y);
            err_state = fmax(err_state, fmax(std::abs(Xn - o0), std::abs(Un - o1)));
            h[k] = Xn + Un;
        }
    }
    for (int n = 0; n < N; n += 7) {
        double acc = h[0].real() + h[M].real() * ((n & 1) ? -1 : 1);
        for (int k = 1; k < M; ++k) acc += 2 * (h[k] * std::polar(1.0, 2 * M_PI * k * n / N)).real();
        const int l = (n >> 1) & 31, n1 = n >> 6;
        const double ours = ((n & 1) ? v[l][n1].y : v[l][n1].x) / N;
        err_x = fmax(err_x, fabs(ours - acc / N));
    }
    printf("OP %d: state err %.3e  frame err %.3e  sums rel err %.3e %.3e\n", OP, err_state, err_x, fabs(ds - dref) / dref, fabs(es - eref) / eref);
    return (err_state < 4e-4 && err_x < 2e-5 && fabs(ds - dref) / dref < 1e-4) ? 0 : 1;
}
int main() { return run<fast::OP_GL>() | run<fast::OP_ADMM>(); }
