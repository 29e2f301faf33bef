// "One residue class per lane": a 512-point complex FFT (n_fft = 1024) spread over TWO warps, 8 complex values per
// lane (LANES = 64, R1 = R2 = R3 = 8) -- the latency-oriented sibling of the 16-values-per-lane scheme of
// gl_warp_core.cuh, used by the RTISI-LA kernel where a signal's inner iterations are strictly sequential and the
// time of ONE frame transform is what counts (half the instructions per lane, twice the warps).
//
// With n = 64 a + 8 b + c and k = ka + 8 kb + 64 kc:
//   pass 1 : lane l = 8 b + c owns z[64 a + l], a < 8: FFT8 over a, times W_512^(l ka)          -- exchange E1 --
//   pass 2 : lane l owns (ka = l >> 3, c = l & 7): FFT8 over b, times W_64^(c kb)               -- exchange E2 --
//   pass 3 : lane l owns the residue class k1 = l: FFT8 over c -> A[kc] = Zh[l + 64 kc].
// The mirror of bin l + 64 kc is bin (64 - l) + 64 (7 - kc): class 64 - l, held by ANOTHER lane (classes 0 and 32 are
// their own mirrors).  Every lane therefore handles the pairs of its LOWER half (A[0..3], bins l + 64 j) and gets the
// partners -- the upper half A[4..7] of lane p = (64 - l) mod 64 -- through a 4-value exchange X ("B[m] = A_p[4 + m]":
// the pair of A[j] is B[3 - j], exactly the (A[j], B[RC-1-j]) pairing of gl_warp_core.cuh with RC = 4), processes its 4
// pair slots, and hands the partner's values back the same way.  Lane 0 (bins 0, 64, ..., 448) is the special one as
// before: slot 0 = DC / Nyquist (A[0]) and bin M/2 (B[0]); slots j = 1..3 pair A[j] with B[4 - j].
//
// Everything is __host__ __device__: tests/host_emu/test_warp_core_1c.cu runs the index logic on the CPU.
#pragma once

#include "gl_warp_core.cuh"

namespace specinv {
namespace wfast {

struct Cfg1 {
    static constexpr int LANES = 64, VV = 8, M = 512, N = 1024, HOP = 256, SLOTS = 4;
};

// exchange addressing: E1 rows (ka, c) -> 8 ka + c of 8 float2, E2 rows k1 of 8 float2 (same swizzles as the 8 x 8 x 8
// case of gl_warp_core.cuh, where a lane does two of these transforms)
SPX_HD void fwd1_pass1(int l, const float2* v, const float2* tw1, const float2* tw1c, float2* e1) {
    float2 t[8];
    static_for<8>([&](auto ac) { constexpr int a = decltype(ac)::value; t[a] = v[a]; });
    fft8<false>(t);
    const int b = l >> 3, c = l & 7;
    static_for<8>([&](auto kc) {
        constexpr int ka = decltype(kc)::value;
        const float2 y = ka == 0 ? t[0] : cmul2t(t[ka], tw1[ka], tw1c[ka]);
        e1[ex_addr<8, 8>(8 * ka + c, b)] = y;
    });
}
SPX_HD void fwd1_pass2(int l, const float2* e1, const float2* tw2, const float2* tw2c, float2* e2) {
    const int ka = l >> 3, c = l & 7;
    float2 t[8];
    static_for<4>([&](auto pc) {
        constexpr int p = decltype(pc)::value;
        const float4 q = *reinterpret_cast<const float4*>(e1 + ex_addr4<8, 8>(8 * ka + c, p));
        t[2 * p] = f2(q.x, q.y); t[2 * p + 1] = f2(q.z, q.w);
    });
    fft8<false>(t);
    static_for<8>([&](auto kc) {
        constexpr int kb = decltype(kc)::value;
        const float2 y = kb == 0 ? t[0] : cmul2t(t[kb], tw2[kb], tw2c[kb]);
        e2[ex_addr<8>(ka + 8 * kb, c)] = y;
    });
}
// A[kc] = Zh[l + 64 kc]
SPX_HD void fwd1_pass3(int l, const float2* e2, float2* A) {
    static_for<4>([&](auto pc) {
        constexpr int p = decltype(pc)::value;
        const float4 q = *reinterpret_cast<const float4*>(e2 + ex_addr4<8>(l, p));
        A[2 * p] = f2(q.x, q.y); A[2 * p + 1] = f2(q.z, q.w);
    });
    fft8<false>(A);
}

// ---- the 4-value exchange with the mirror lane (X: 256 float2, [m][lane]) ---------------------------------------
SPX_HD int mirror_lane(int l) { return (64 - l) & 63; }
SPX_HD void pair_publish(int l, const float2* A, float2* X) {          // my upper half, for whoever mirrors me
    static_for<4>([&](auto mc) { constexpr int m = decltype(mc)::value; X[64 * m + l] = A[4 + m]; });
}
SPX_HD void pair_fetch(int l, const float2* X, float2* B) {            // B[m] = A_p[4 + m]
    const int p = mirror_lane(l);
    static_for<4>([&](auto mc) { constexpr int m = decltype(mc)::value; B[m] = X[64 * m + p]; });
}
SPX_HD void pair_return(int l, const float2* B, float2* X) {           // the partner's new upper half, back in its row
    const int p = mirror_lane(l);
    static_for<4>([&](auto mc) { constexpr int m = decltype(mc)::value; X[64 * m + p] = B[m]; });
}
SPX_HD void pair_collect(int l, const float2* X, float2* A) {          // my own new upper half
    static_for<4>([&](auto mc) { constexpr int m = decltype(mc)::value; A[4 + m] = X[64 * m + l]; });
}

// bin of element e of lane l (e = 2 j: the P bin of slot j, e = 2 j + 1: its mirror; lane 0 slot 0: bins 0 and M/2;
// e = -1: the Nyquist bin, lane 0 only)
SPX_HD int bin1(int l, int e) {
    if (e < 0) return Cfg1::M;
    const int j = e >> 1;
    if (!(e & 1)) return l + 64 * j;
    return (l == 0 && j == 0) ? Cfg1::M / 2 : Cfg1::M - l - 64 * j;
}

// Point-wise stage on the lane's 4 pair slots, in place: A[0..3] (own lower half) and B[0..3] (the partner's upper
// half) become the inputs of the inverse transform.  io as in gl_warp_core.cuh::pointwise.
template <int OP, bool SUMS, typename IO>
SPX_HD void pointwise1(int l, float2* A, float2* B, const float2* twr, const float2* twrc, IO& io, float coef, float coef2,
                       float& dsum, float& esum) {
    const bool l0 = l == 0;
    auto upd = [&](auto ec, float2 sv) {
        constexpr int e = decltype(ec)::value;
        float2 o0 = f2(0.f, 0.f), o1 = f2(0.f, 0.f);
        const float2 h = bin_update<OP, SUMS>(sv, OP == OP_GLP ? f2(0.f, 0.f) : io.s0(e),
                                              OP == OP_ADMM ? io.s1(e) : f2(0.f, 0.f), io.mag(e), coef, coef2, o0, o1, dsum, esum);
        if constexpr (OP != OP_GLP) io.put(e, o0, o1);
        return h;
    };
    if (l0) {
        // A[0] = Zh[0] -> DC and Nyquist (both real); B[0] = Zh[M/2] -> bin M/2 = conj(Z[M/2])
        const float2 z0 = A[0], z4 = B[0];
        const float2 h0 = upd(std::integral_constant<int, 0>{}, f2(2.f * (z0.x + z0.y), 0.f));
        const float2 hM = upd(std::integral_constant<int, -1>{}, f2(2.f * (z0.x - z0.y), 0.f));
        const float2 h4 = upd(std::integral_constant<int, 1>{}, f2(2.f * z4.x, -2.f * z4.y));
        A[0] = f2(h0.x + hM.x, h0.x - hM.x);
        B[0] = f2(2.f * h4.x, -2.f * h4.y);
    } else {
        float2 sP, sQ, P, Q;
        post_pair_t<true>(A[0], B[3], twr[0], twrc[0], sP, sQ);
        const float2 hP = upd(std::integral_constant<int, 0>{}, sP);
        const float2 hQ = upd(std::integral_constant<int, 1>{}, sQ);
        pre_pair_t<true>(hP, hQ, twr[0], twrc[0], P, Q);
        A[0] = P; B[3] = Q;
    }
    static_for<3>([&](auto jc) {
        constexpr int j = decltype(jc)::value + 1;
        // general lanes: (A[j], B[3 - j]); lane 0 (class 0 mirrors itself, bins 64 j and 512 - 64 j): (A[j], B[4 - j])
        float2 P = A[j], Q = l0 ? B[4 - j] : B[3 - j];
        float2 sP, sQ;
        post_pair_t<true>(P, Q, twr[j], twrc[j], sP, sQ);
        const float2 hP = upd(std::integral_constant<int, 2 * j>{}, sP);
        const float2 hQ = upd(std::integral_constant<int, 2 * j + 1>{}, sQ);
        pre_pair_t<true>(hP, hQ, twr[j], twrc[j], P, Q);
        A[j] = P;
        if (l0) B[4 - j] = Q; else B[3 - j] = Q;
    });
}

// Stand-alone inverse transform: the given spectrum h (io.s0) replaces the point-wise stage.
template <typename IO>
SPX_HD void spectrum_pairs1(int l, float2* A, float2* B, const float2* twr, const float2* twrc, IO& io) {
    const bool l0 = l == 0;
    if (l0) {
        const float2 h0 = io.s0(0), hM = io.s0(-1), h4 = io.s0(1);
        A[0] = f2(h0.x + hM.x, h0.x - hM.x);
        B[0] = f2(2.f * h4.x, -2.f * h4.y);
    } else {
        pre_pair_t<true>(io.s0(0), io.s0(1), twr[0], twrc[0], A[0], B[3]);
    }
    static_for<3>([&](auto jc) {
        constexpr int j = decltype(jc)::value + 1;
        float2 P, Q;
        pre_pair_t<true>(io.s0(2 * j), io.s0(2 * j + 1), twr[j], twrc[j], P, Q);
        A[j] = P;
        if (l0) B[4 - j] = Q; else B[3 - j] = Q;
    });
}

// ---- inverse ---------------------------------------------------------------------------------------------------
SPX_HD void inv1_pass3(int l, float2* A, float2* e2) {
    fft8<true>(A);
    static_for<4>([&](auto pc) {
        constexpr int p = decltype(pc)::value;
        *reinterpret_cast<float4*>(e2 + ex_addr4<8>(l, p)) = make_float4(A[2 * p].x, A[2 * p].y, A[2 * p + 1].x, A[2 * p + 1].y);
    });
}
SPX_HD void inv1_pass2(int l, const float2* e2, const float2* tw2, const float2* tw2c, float2* e1) {
    const int ka = l >> 3, c = l & 7;
    float2 t[8];
    static_for<8>([&](auto kc) {
        constexpr int kb = decltype(kc)::value;
        const float2 y = e2[ex_addr<8>(ka + 8 * kb, c)];
        t[kb] = kb == 0 ? y : cmulc2t(y, tw2[kb], tw2c[kb]);
    });
    fft8<true>(t);
    static_for<4>([&](auto pc) {
        constexpr int p = decltype(pc)::value;
        *reinterpret_cast<float4*>(e1 + ex_addr4<8, 8>(8 * ka + c, p)) = make_float4(t[2 * p].x, t[2 * p].y, t[2 * p + 1].x, t[2 * p + 1].y);
    });
}
// v[a] = z'[64 a + l] (unscaled)
SPX_HD void inv1_pass1(int l, const float2* e1, const float2* tw1, const float2* tw1c, float2* v) {
    const int b = l >> 3, c = l & 7;
    float2 t[8];
    static_for<8>([&](auto kc) {
        constexpr int ka = decltype(kc)::value;
        const float2 y = e1[ex_addr<8, 8>(8 * ka + c, b)];
        t[ka] = ka == 0 ? y : cmulc2t(y, tw1[ka], tw1c[ka]);
    });
    fft8<true>(t);
    static_for<8>([&](auto ac) { constexpr int a = decltype(ac)::value; v[a] = t[a]; });
}

}  // namespace wfast
}  // namespace specinv
