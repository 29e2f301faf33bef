// Per-lane building blocks of the fast fused iteration kernel for n_fft = 2048, hop = 512:
// "one warp per frame".  Same structure as gl_fast_core.cuh (n_fft = 1024, half-warp per frame) with
// M = 1024 = 32 x 32:
//   forward : lane n2 (0..31) owns z[32 n1 + n2], 32-point FFT over n1, times W_1024^(n2 k1), ONE exchange
//             through shared memory, then lane p owns the residue class k1 = p: 32-point FFT over n2
//             -> Z[p + 32 k2].
//   The mirror Z[M - k] of lane p's bins lives in lane (32 - p) % 32 (index 31 - k2; lane 0: index
//   (32 - k2) % 32), so the real-FFT post/pre-processing pairs are completed with ONE warp-shuffle
//   exchange per direction; every lane then updates its own 32 bins, which are consecutive across the
//   warp for a fixed k2 (256-byte coalesced rows of q / mag).
//   inverse : the same steps backwards.
// A hop of 512 samples is again 8 of a lane's 32 sample pairs: the register overlap-add, the TMEM carry
// and the input ring of the 1024 kernel carry over unchanged.
#pragma once

#include "gl_fast_core.cuh"

namespace specinv {
namespace fast32 {

using fast::FrameIO;
using fast::Tables;
using fast::ROW;
using fast::cmulf;
using fast::cmulcf;
using fast::bin_update_fast;

constexpr int N = 2048;
constexpr int M = 1024;
constexpr int LANES = 32;
constexpr int TBL = LANES * ROW;

SPX_HD void phase1_compute(int l, float2* v, const Tables& tb) {
    fft32<false>(v);
    static_for<16>([&](auto ic) {
        constexpr int k1 = 2 * decltype(ic)::value;
        const float4 t = *reinterpret_cast<const float4*>(tb.tw + l * ROW + k1);
        if constexpr (k1 != 0) v[k1] = cmulf(v[k1], f2(t.x, t.y));
        v[k1 + 1] = cmulf(v[k1 + 1], f2(t.z, t.w));
    });
}
SPX_HD void phase1_write(int l, const float2* v, float2* exch) {
    static_for<16>([&](auto ic) {
        constexpr int k1 = 2 * decltype(ic)::value;
        *reinterpret_cast<float4*>(exch + l * ROW + k1) = make_float4(v[k1].x, v[k1].y, v[k1 + 1].x, v[k1 + 1].y);
    });
}
// lane p reads its residue class k1 = p from every row and transforms over n2
SPX_HD void phase2_read_fft(int p, const float2* exch, float2* A) {
    static_for<32>([&](auto nc) { constexpr int n2 = decltype(nc)::value; A[n2] = exch[n2 * ROW + p]; });
    fft32<false>(A);      // A[k2] = Zh[p + 32 k2]
}

// index (in the partner lane's array) of the mirror of element k2
template <int K2> SPX_HD int mirror_index(bool l0) { return l0 ? ((32 - K2) & 31) : 31 - K2; }

// Point-wise stage of the lane's own 32 bins.  Zp = the partner lane's A (after the shuffle exchange).
// On return A[k2] holds h[p + 32 k2] (the projected / updated spectrum), ready for the second exchange;
// lane 0 keeps the Nyquist value in `h_nyq`.
template <int OP, bool SUMS>
SPX_HD void pointwise_own(int p, float2* A, const float2* Zp, const Tables& tb, const FrameIO& io, const float* mag_row,
                          float& h_nyq, float& dsum, float& esum) {
    const bool l0 = p == 0;
    static_for<32>([&](auto kc) {
        constexpr int k2 = decltype(kc)::value;
        const int kP = p + 32 * k2;
        const float2 P = A[k2];
        const float2 Q = l0 ? Zp[(32 - k2) & 31] : Zp[31 - k2];
        const float2 w = tb.twr[kP];
        const float er = P.x + Q.x, ei = P.y - Q.y, orr = P.y + Q.y, oi = Q.x - P.x;
        const float wor = w.x * orr - w.y * oi, woi = w.x * oi + w.y * orr;
        const float2 sP = f2(er + wor, ei + woi);
        A[k2] = bin_update_fast<OP, SUMS, false>(io, kP, sP, mag_row[kP], dsum, esum);
        if constexpr (k2 == 0) {
            if (l0) {   // the Nyquist bin s[M] = er - wor (real) rides along with DC
                const float2 hM = bin_update_fast<OP, SUMS, true>(io, 0, f2(er - wor, 0.f), io.mag_nyq_val, dsum, esum);
                h_nyq = hM.x;
            }
        }
    });
}

// inverse pre-processing of the lane's own bins: A = h (own), Hp = partner's h  ->  A = Z'[p + 32 k2]
SPX_HD void pre_own(int p, float2* A, const float2* Hp, const Tables& tb, float h_nyq) {
    const bool l0 = p == 0;
    static_for<32>([&](auto kc) {
        constexpr int k2 = decltype(kc)::value;
        const int kP = p + 32 * k2;
        const float2 hP = A[k2];
        const float2 hQ = l0 ? Hp[(32 - k2) & 31] : Hp[31 - k2];
        const float2 w = tb.twr[kP];
        const float Ar = hP.x + hQ.x, Ai = hP.y - hQ.y, Dr = hP.x - hQ.x, Di = hP.y + hQ.y;
        const float Gr = w.x * Dr + w.y * Di, Gi = w.x * Di - w.y * Dr;
        float2 z = f2(Ar - Gi, Ai + Gr);
        if constexpr (k2 == 0) {
            if (l0) z = f2(hP.x + h_nyq, hP.x - h_nyq);     // C2R ignores Im(DC), Im(Nyquist)
        }
        A[k2] = z;
    });
}

SPX_HD void phase2_ifft_write(int p, float2* A, float2* exch) {
    fft32<true>(A);       // A[n2] = Y_p[n2]
    static_for<32>([&](auto nc) { constexpr int n2 = decltype(nc)::value; exch[n2 * ROW + p] = A[n2]; });
}

SPX_HD void phase3(int l, float2* v, const Tables& tb, const float2* exch) {
    static_for<16>([&](auto ic) {
        constexpr int k1 = 2 * decltype(ic)::value;
        const float4 y = *reinterpret_cast<const float4*>(exch + l * ROW + k1);
        const float4 t = *reinterpret_cast<const float4*>(tb.tw + l * ROW + k1);
        v[k1] = k1 == 0 ? f2(y.x, y.y) : cmulcf(f2(y.x, y.y), f2(t.x, t.y));
        v[k1 + 1] = cmulcf(f2(y.z, y.w), f2(t.z, t.w));
    });
    fft32<true>(v);
}

}  // namespace fast32
}  // namespace specinv
