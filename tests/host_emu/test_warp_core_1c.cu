// Host emulation of the "one residue class per lane" pipeline of gl_warp_core_1c.cuh (n_fft = 1024 over 64 lanes, 8
// values per lane): the lanes run sequentially, phase boundaries stand in for the frame barriers.  State update, metric
// sums and the inverse-transform output against a double-precision naive real DFT of one frame.
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../spectrogram_inversion_b200/csrc/gl_warp_core_1c.cuh"
using namespace specinv;
using namespace specinv::wfast;
typedef std::complex<double> cd;
static double frand() { return (double)rand() / RAND_MAX - 0.5; }
static float2 cis(double num, double den) { const double a = -2 * M_PI * num / den; return f2((float)cos(a), (float)sin(a)); }
static float2 cj(float2 w) { return f2(w.x, -w.y); }

struct Tab { float2 wa[8], tw1[8], tw1c[8], tw2[8], tw2c[8], twr[4], twrc[4]; };

template <int OP>
int run() {
    constexpr int L = 64, N = 1024, M = 512;
    std::vector<float> x(N), wa(N);
    for (int i = 0; i < N; ++i) { x[i] = (float)(4 * frand()); wa[i] = (float)(0.5 - 0.5 * cos(2 * M_PI * i / N)); }
    std::vector<Tab> tb(L);
    for (int l = 0; l < L; ++l) {
        for (int i = 0; i < 8; ++i) {
            tb[l].wa[i] = f2(0.5f * wa[2 * L * i + 2 * l], 0.5f * wa[2 * L * i + 2 * l + 1]);
            tb[l].tw1[i] = cis((l * i) % M, M); tb[l].tw1c[i] = cj(tb[l].tw1[i]);
            tb[l].tw2[i] = cis((l & 7) * i, 64); tb[l].tw2c[i] = cj(tb[l].tw2[i]);
        }
        for (int j = 0; j < 4; ++j) { tb[l].twr[j] = cis(l + 64 * j, N); tb[l].twrc[j] = cj(tb[l].twr[j]); }
    }
    std::vector<float2> s0_in(M + 1), s1_in(M + 1), s0_out(M + 1), s1_out(M + 1);
    std::vector<float> mag(M + 1);
    std::vector<int> touched(M + 1, 0);
    for (int k = 0; k <= M; ++k) {
        s0_in[k] = f2((float)(20 * frand()), (float)(20 * frand())); s1_in[k] = f2((float)(5 * frand()), (float)(5 * frand()));
        mag[k] = (float)(10 * fabs(frand()));
    }
    const float coef = OP == OP_GL ? 0.3f : 0.1f, coef2 = 1.f / (1.f + coef);
    std::vector<float2> e1(M), e2(M), X(256);
    std::vector<std::vector<float2>> v(L, std::vector<float2>(8)), A(L, std::vector<float2>(8)), B(L, std::vector<float2>(4));
    for (int l = 0; l < L; ++l) {
        for (int i = 0; i < 8; ++i) v[l][i] = f2(x[2 * L * i + 2 * l] * tb[l].wa[i].x, x[2 * L * i + 2 * l + 1] * tb[l].wa[i].y);
        fwd1_pass1(l, v[l].data(), tb[l].tw1, tb[l].tw1c, e1.data());
    }
    for (int l = 0; l < L; ++l) fwd1_pass2(l, e1.data(), tb[l].tw2, tb[l].tw2c, e2.data());
    for (int l = 0; l < L; ++l) { fwd1_pass3(l, e2.data(), A[l].data()); pair_publish(l, A[l].data(), X.data()); }
    float ds = 0, es = 0;
    for (int l = 0; l < L; ++l) pair_fetch(l, X.data(), B[l].data());
    for (int l = 0; l < L; ++l) {
        struct IO {
            int l; const float2* s0i; const float2* s1i; const float* mg; float2* s0o; float2* s1o; int* touched;
            SPX_HD float2 s0(int e) const { return s0i[bin1(l, e)]; }
            SPX_HD float2 s1(int e) const { return s1i[bin1(l, e)]; }
            SPX_HD float mag(int e) const { return mg[bin1(l, e)]; }
            SPX_HD void put(int e, float2 o0, float2 o1) { s0o[bin1(l, e)] = o0; s1o[bin1(l, e)] = o1; touched[bin1(l, e)]++; }
        } io{l, s0_in.data(), s1_in.data(), mag.data(), s0_out.data(), s1_out.data(), touched.data()};
        pointwise1<OP, true>(l, A[l].data(), B[l].data(), tb[l].twr, tb[l].twrc, io, coef, coef2, ds, es);
        pair_return(l, B[l].data(), X.data());
    }
    for (int l = 0; l < L; ++l) pair_collect(l, X.data(), A[l].data());
    for (int l = 0; l < L; ++l) inv1_pass3(l, A[l].data(), e2.data());
    for (int l = 0; l < L; ++l) inv1_pass2(l, e2.data(), tb[l].tw2, tb[l].tw2c, e1.data());
    for (int l = 0; l < L; ++l) inv1_pass1(l, e1.data(), tb[l].tw1, tb[l].tw1c, v[l].data());

    int bad_cover = 0;
    for (int k = 0; k <= M; ++k) bad_cover += touched[k] != 1;          // every bin updated exactly once
    std::vector<cd> s(M + 1), h(M + 1);
    double dref = 0, eref = 0, err_state = 0, err_x = 0;
    for (int k = 0; k <= M; ++k) {
        cd acc = 0;
        for (int n = 0; n < N; ++n) acc += (double)x[n] * (double)wa[n] * std::polar(1.0, -2 * M_PI * (double)((long long)k * n % N) / N);
        s[k] = acc;
        cd a0(s0_in[k].x, s0_in[k].y), a1(s1_in[k].x, s1_in[k].y), o0(s0_out[k].x, s0_out[k].y), o1(s1_out[k].x, s1_out[k].y);
        const double m = mag[k];
        dref += (std::abs(s[k]) - m) * (std::abs(s[k]) - m); eref += std::norm(s[k]);
        if (OP == OP_GL) {
            cd q = s[k] - (double)coef * a0;
            err_state = fmax(err_state, std::abs(q - o0));
            h[k] = q * m / (std::abs(q) + 1e-16);
        } else {
            cd Z = ((double)coef * (a0 + a1) + s[k]) / (1.0 + coef);
            cd Un = a1 + a0 - Z;
            cd Xn = (Z - Un) * m / (std::abs(Z - Un) + 1e-16);
            err_state = fmax(err_state, fmax(std::abs(Xn - o0), std::abs(Un - o1)));
            h[k] = Xn + Un;
        }
    }
    for (int n = 0; n < N; ++n) {
        double acc = h[0].real() + h[M].real() * ((n & 1) ? -1 : 1);
        for (int k = 1; k < M; ++k) acc += 2 * (h[k] * std::polar(1.0, 2 * M_PI * (double)((long long)k * n % N) / N)).real();
        const int l = (n >> 1) % L, i = n / (2 * L);
        const double ours = ((n & 1) ? v[l][i].y : v[l][i].x) / N;
        err_x = fmax(err_x, fabs(ours - acc / N));
    }
    printf("one class per lane, OP %d: bins not updated exactly once %d  state err %.3e  frame err %.3e  sums rel err %.3e %.3e\n",
           OP, bad_cover, err_state, err_x, fabs(ds - dref) / dref, fabs(es - eref) / eref);
    return (bad_cover == 0 && err_state < 4e-4 && err_x < 2e-5 && fabs(ds - dref) / dref < 1e-4 && fabs(es - eref) / eref < 1e-4) ? 0 : 1;
}

int main() { return run<OP_GL>() | run<OP_ADMM>(); }
