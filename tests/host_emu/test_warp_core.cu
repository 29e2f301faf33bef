// Host emulation of the "16 values per lane" pipeline of gl_warp_core.cuh for LANES = 32 / 64 / 128
// (n_fft = 1024 / 2048 / 4096): the lanes run sequentially, phase boundaries stand in for the group
// barriers.  Checks the state update and the inverse-transform output against a double-precision naive real
// DFT of one frame, and the bank-conflict freedom of the exchange addressing.
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <set>
#include <vector>
#include "../../spectrogram_inversion_b200/csrc/gl_warp_core.cuh"
using namespace specinv;
using namespace specinv::wfast;
typedef std::complex<double> cd;
static double frand() { return (double)rand() / RAND_MAX - 0.5; }
static float2 cis(double num, double den) { const double a = -2 * M_PI * num / den; return f2((float)cos(a), (float)sin(a)); }

template <int LANES, int VV>
static void make_tables(int l, const std::vector<float>& wa, const std::vector<float>& ws, LaneTables<LANES, VV>& t) {
    using C = Cfg<LANES, VV>;
    for (int i = 0; i < VV; ++i) {
        t.wa[i] = f2(0.5f * wa[2 * LANES * i + 2 * l], 0.5f * wa[2 * LANES * i + 2 * l + 1]);
        t.ws[i] = f2(ws[2 * LANES * i + 2 * l], ws[2 * LANES * i + 2 * l + 1]);
    }
    for (int s = 0; s < C::S1; ++s)
        for (int ka = 0; ka < C::R1; ++ka) t.tw1[C::R1 * s + ka] = cis(((l + LANES * s) * ka) % C::M, C::M);
    for (int kb = 0; kb < C::R2; ++kb) t.tw2[kb] = cis((l & (C::RC - 1)) * kb, C::RC * C::R2);
    for (int j = 0; j < C::RC; ++j) t.twr[j] = cis(slot_bin_rt<LANES, VV>(l, j), C::N);
}

// wavefronts of one warp-wide shared-memory access (float2 addresses, width in 32-bit words)
static int wavefronts(const int* addr_f2, int words) {
    const int group = words == 4 ? 8 : 16;        // lanes served together: 128-bit -> quarter warps, 64-bit -> half warps
    int total = 0;
    for (int g = 0; g < 32; g += group) {
        std::set<int> banks[32];
        int worst = 1;
        for (int l = g; l < g + group; ++l)
            for (int w = 0; w < words; ++w) {
                const int word = addr_f2[l] * 2 + w;
                banks[word & 31].insert(word);
            }
        for (auto& b : banks) worst = std::max(worst, (int)b.size());
        total += worst;
    }
    return total;
}

template <int LANES, int VV>
static int check_conflicts() {
    using C = Cfg<LANES, VV>;
    int bad = 0, addr[32];
    for (int w = 0; w < LANES / 32; ++w) {
        for (int s = 0; s < C::S1; ++s) for (int ka = 0; ka < C::R1; ++ka) {      // E1 scattered 64-bit (pass 1 side)
            for (int i = 0; i < 32; ++i) { const int l = 32 * w + i; addr[i] = ex_addr<C::R2, C::RC>(C::RC * ka + (l & (C::RC - 1)), (l + LANES * s) >> C::LOGRC); }
            bad += wavefronts(addr, 2) != 2;
        }
        for (int r = 0; r < C::S2; ++r) for (int p = 0; p < C::R2 / 2; ++p) {     // E1 rows 128-bit (pass 2 side)
            for (int i = 0; i < 32; ++i) { const int l = 32 * w + i; addr[i] = ex_addr4<C::R2, C::RC>(C::RC * ((l >> C::LOGRC) + LANES / C::RC * r) + (l & (C::RC - 1)), p); }
            bad += wavefronts(addr, 4) != 4;
        }
        for (int r = 0; r < C::S2; ++r) for (int kb = 0; kb < C::R2; ++kb) {      // E2 scattered 64-bit (pass 2 side)
            for (int i = 0; i < 32; ++i) { const int l = 32 * w + i; addr[i] = ex_addr<C::RC>((l >> C::LOGRC) + LANES / C::RC * r + C::R1 * kb, l & (C::RC - 1)); }
            bad += wavefronts(addr, 2) != 2;
        }
        for (int cls = 0; cls < 2; ++cls) for (int p = 0; p < C::RC / 2; ++p) {   // E2 rows 128-bit (pass 3 side)
            for (int i = 0; i < 32; ++i) { const int l = 32 * w + i; addr[i] = ex_addr4<C::RC>(cls ? class_b<LANES>(l) : l, p); }
            bad += wavefronts(addr, 4) != 4;
        }
    }
    printf("LANES %d V %d: exchange accesses with bank conflicts: %d\n", LANES, VV, bad);
    return bad;
}

// HC: the variants that take the conjugate twiddle tables (cmul2t / cmulc2t), as the one-warp-per-frame kernels do
template <int LANES, int OP, int VV = V, bool HC = false>
int run() {
    using C = Cfg<LANES, VV>;
    constexpr int V = VV, RC = C::RC;
    constexpr int N = C::N, M = C::M;
    std::vector<float> x(N), wa(N), ws(N);
    for (int i = 0; i < N; ++i) { x[i] = (float)(4 * frand()); wa[i] = (float)(0.5 - 0.5 * cos(2 * M_PI * i / N)); ws[i] = wa[i] / N; }
    std::vector<LaneTables<LANES, VV>> tb(LANES);
    for (int l = 0; l < LANES; ++l) make_tables<LANES, VV>(l, wa, ws, tb[l]);
    std::vector<LaneTables<LANES, VV>> tc = tb;                     // conjugates of the three twiddle tables
    for (int l = 0; l < LANES; ++l) {
        for (int i = 0; i < VV; ++i) tc[l].tw1[i].y = -tc[l].tw1[i].y;
        for (int i = 0; i < C::R2; ++i) tc[l].tw2[i].y = -tc[l].tw2[i].y;
        for (int i = 0; i < VV / 2; ++i) tc[l].twr[i].y = -tc[l].twr[i].y;
    }
    std::vector<float2> s0_in(M + 1), s1_in(M + 1), s0_out(M + 1), s1_out(M + 1);
    std::vector<float> mag(M + 1);
    for (int k = 0; k <= M; ++k) {
        s0_in[k] = f2((float)(20 * frand()), (float)(20 * frand())); s1_in[k] = f2((float)(5 * frand()), (float)(5 * frand()));
        mag[k] = (float)(10 * fabs(frand()));
    }
    const float coef = OP == OP_GL ? 0.3f : 0.1f, coef2 = 1.f / (1.f + coef);

    std::vector<float2> e1(M), e2(M);
    std::vector<std::vector<float2>> v(LANES, std::vector<float2>(V)), A(LANES, std::vector<float2>(RC)), Bv(LANES, std::vector<float2>(RC));
    for (int l = 0; l < LANES; ++l) {
        for (int i = 0; i < V; ++i) v[l][i] = f2(x[2 * LANES * i + 2 * l] * tb[l].wa[i].x, x[2 * LANES * i + 2 * l + 1] * tb[l].wa[i].y);
        fwd_pass1<LANES, VV, HC>(l, v[l].data(), tb[l].tw1, e1.data(), tc[l].tw1);
    }
    for (int l = 0; l < LANES; ++l) fwd_pass2<LANES, VV, HC>(l, e1.data(), tb[l].tw2, e2.data(), tc[l].tw2);
    float ds = 0, es = 0;
    for (int l = 0; l < LANES; ++l) fwd_pass3<LANES, VV>(l, e2.data(), A[l].data(), Bv[l].data());
    for (int l = 0; l < LANES; ++l) {
        struct IO {
            int l; const float2* s0i; const float2* s1i; const float* mg; float2* s0o; float2* s1o;
            SPX_HD int bin(int e) const {
                if (e < 0) return Cfg<LANES, VV>::M;
                const int kP = slot_bin_rt<LANES, VV>(l, e >> 1);
                return (e & 1) ? ((l == 0 && e == 1) ? Cfg<LANES, VV>::M / 2 : Cfg<LANES, VV>::M - kP) : kP;
            }
            SPX_HD float2 s0(int e) const { return s0i[bin(e)]; }
            SPX_HD float2 s1(int e) const { return s1i[bin(e)]; }
            SPX_HD float mag(int e) const { return mg[bin(e)]; }
            SPX_HD void put(int e, float2 o0, float2 o1) { s0o[bin(e)] = o0; s1o[bin(e)] = o1; }
        } io{l, s0_in.data(), s1_in.data(), mag.data(), s0_out.data(), s1_out.data()};
        pointwise<OP, true, VV, HC>(l, A[l].data(), Bv[l].data(), tb[l].twr, io, coef, coef2, ds, es, tc[l].twr);
    }
    for (int l = 0; l < LANES; ++l) inv_pass3<LANES, VV>(l, A[l].data(), Bv[l].data(), e2.data());
    for (int l = 0; l < LANES; ++l) inv_pass2<LANES, VV, HC>(l, e2.data(), tb[l].tw2, e1.data(), tc[l].tw2);
    for (int l = 0; l < LANES; ++l) inv_pass1<LANES, VV, HC>(l, e1.data(), tb[l].tw1, v[l].data(), tc[l].tw1);

    std::vector<cd> s(M + 1), h(M + 1);
    double dref = 0, eref = 0, err_state = 0, err_x = 0;
    for (int k = 0; k <= M; ++k) {
        cd acc = 0;
        for (int n = 0; n < N; ++n) acc += (double)x[n] * (double)wa[n] * std::polar(1.0, -2 * M_PI * (double)((long long)k * n % N) / N);
        s[k] = acc;
        cd a0(s0_in[k].x, s0_in[k].y), a1(s1_in[k].x, s1_in[k].y), o0(s0_out[k].x, s0_out[k].y), o1(s1_out[k].x, s1_out[k].y);
        double m = mag[k];
        dref += (std::abs(s[k]) - m) * (std::abs(s[k]) - m); eref += std::norm(s[k]);
        if (OP == OP_GLP) {             // plain Griffin-Lim: q = s, no state read or written
            h[k] = s[k] * m / (std::abs(s[k]) + 1e-16);
            err_state = fmax(err_state, std::abs(o0) + std::abs(o1));       // outputs untouched (still zero)
        } else if (OP == OP_GL) {
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
    for (int n = 0; n < N; n += (N > 1024 ? 5 : 1)) {
        double acc = h[0].real() + h[M].real() * ((n & 1) ? -1 : 1);
        for (int k = 1; k < M; ++k) acc += 2 * (h[k] * std::polar(1.0, 2 * M_PI * (double)((long long)k * n % N) / N)).real();
        const int l = (n >> 1) % LANES, i = n / (2 * LANES);
        const double ours = ((n & 1) ? v[l][i].y : v[l][i].x) / N;
        err_x = fmax(err_x, fabs(ours - acc / N));
    }
    if (HC) printf("(conjugate tables) ");
    printf("LANES %d V %d OP %d: state err %.3e  frame err %.3e  sums rel err %.3e %.3e\n", LANES, V, OP, err_state, err_x,
           fabs(ds - dref) / dref, fabs(es - eref) / eref);
    return (err_state < 4e-4 && err_x < 2e-5 && fabs(ds - dref) / dref < 1e-4 && fabs(es - eref) / eref < 1e-4) ? 0 : 1;
}

int main() {
    int rc = check_conflicts<32, 16>() | check_conflicts<64, 16>() | check_conflicts<128, 16>() | check_conflicts<32, 8>();
    rc |= run<32, OP_GL, 8>() | run<32, OP_ADMM, 8>();
    rc |= run<32, OP_GL>() | run<32, OP_ADMM>();
    rc |= run<64, OP_GL>() | run<64, OP_ADMM>();
    rc |= run<128, OP_GL>() | run<128, OP_ADMM>();
    rc |= run<32, OP_GLP, 8>() | run<32, OP_GLP>() | run<64, OP_GLP>() | run<128, OP_GLP>();
    rc |= run<32, OP_GL, 16, true>() | run<32, OP_ADMM, 16, true>() | run<32, OP_GLP, 16, true>();
    rc |= run<32, OP_GL, 8, true>() | run<32, OP_ADMM, 8, true>() | run<64, OP_GL, 16, true>();
    return rc;
}
