// Per-lane building blocks of the fast (n_fft = 1024, hop = 256) fused iteration kernel.
//
// Thread mapping ("half-warp per frame"): a frame's real FFT of N = 1024 is a complex FFT of
// M = 512 = 32 x 16 points done by G = 16 lanes holding E = 32 complex values each, with ONE
// exchange through shared memory per direction:
//   forward  : lane n2 owns z[16 n1 + n2] (n1 = 0..31), does a 32-point FFT over n1, multiplies by
//              W_512^(n2 k1) and writes V[k1] to exch[n2][k1];   -- exchange --
//              lane p owns the residue classes k1 in {p, 32-p} ({0, 16} for p = 0), reads
//              exch[n2][k1] for all n2 and does two 16-point FFTs over n2 -> Z[k1 + 32 k2].
//   Because class 32-p is the mirror (k -> M-k) of class p, every (Z[k], Z[M-k]) pair that the
//   real-FFT post/pre-processing needs lives in ONE thread: the magnitude projection and the
//   momentum / ADMM update run in registers, straight on the FFT outputs.
//   inverse  : the same two steps backwards (two inverse 16-point FFTs, exchange, conj twiddle,
//              inverse 32-point FFT) leaves lane n2 with the time samples 32 n1 + 2 n2 + {0,1}.
// A hop of 256 samples is 8 values of n1, so consecutive frames of one signal shift the SAME
// thread's data by 8 registers: overlap-add is a per-thread register shift-accumulate, no atomics,
// no shared-memory traffic, and the input ring is thread-private.
//
// All functions are __host__ __device__ so that tests/host_emu can run the exact index logic on the
// CPU (lanes emulated sequentially, __syncwarp points = phase boundaries).
#pragma once

#include "fft_regs.cuh"

namespace specinv {
namespace fast {

constexpr int N = 1024;
constexpr int M = 512;
constexpr int HOP = 256;
constexpr int ROW = 34;          // padded row length (float2) of every [16][32] lane-major table
constexpr int TBL = 16 * ROW;    // float2 elements of one lane-major table

enum { OP_GL = 0, OP_ADMM = 1 };

struct Tables {
    const float2* tw;    // [16][ROW]  W_512^(n2 k1)
    const float2* wa;    // [16][ROW]  0.5 * analysis window pairs (w[32 n1 + 2 n2], w[32 n1 + 2 n2 + 1])
    const float2* ws;    // [16][ROW]  synthesis window pairs (already scaled by 1/N or N^-1/2)
    const float2* twr;   // [512]      W_1024^k
};

// Global-memory view of the current frame (row pointers already offset to the frame).
struct FrameIO {
    const float2* s0_stage;                           // GL: q_in / ADMM: X_in main row, STAGED in shared memory
    const float2* s0_in_nyq;
    float2* s0_out;       float2* s0_out_nyq;
    const float2* s1_in;  const float2* s1_in_nyq;    // ADMM: U_in
    float2* s1_out;       float2* s1_out_nyq;
    const float* mag;     const float* mag_nyq;       // mag: main row (global, or staged in shared memory)
    float2 s0_nyq_val;    // preloaded s0_in_nyq[0] (only lane 0 uses it)
    float mag_nyq_val;    // preloaded mag_nyq[0]
    float coef;           // lr or rho
    float coef2;          // ADMM: 1/(1+rho)
    bool owned;           // write state / count sums for this frame
};

SPX_HD float2 cmulf(float2 a, float2 b) { return f2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
SPX_HD float2 cmulcf(float2 a, float2 b) { return f2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y); }   // a*conj(b)

#ifdef __CUDA_ARCH__
__device__ __forceinline__ float approx_sqrt(float v) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v)); return r; }
__device__ __forceinline__ float approx_rcp(float v) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v)); return r; }
#define SPX_LDG(p) __ldg(p)
#else
inline float approx_sqrt(float v) { return sqrtf(v); }
inline float approx_rcp(float v) { return 1.0f / v; }
#define SPX_LDG(p) (*(p))
#endif

// q * mag / (|q| + 1e-16)   (methods.py:246-247)
SPX_HD float2 project_fast(float2 q, float mag) {
    const float r = approx_sqrt(q.x * q.x + q.y * q.y);
    const float s = mag * approx_rcp(r + 1e-16f);
    return f2(q.x * s, q.y * s);
}

// Point-wise stage for one bin.  `s` is the STFT bin of the current estimate, `kk` its index in the
// main row (ignored when NYQ: the k = 512 bin lives in the separate nyq arrays).  Returns the value
// fed to the inverse transform.
template <int OP, bool SUMS, bool NYQ>
SPX_HD float2 bin_update_fast(const FrameIO& io, int kk, float2 s, float m, float& dsum, float& esum) {
    if constexpr (SUMS) {
        if (io.owned) {
            const float r = approx_sqrt(s.x * s.x + s.y * s.y);
            dsum += (r - m) * (r - m);
            esum += r * r;
        }
    }
    if constexpr (OP == OP_GL) {
        const float2 qp = NYQ ? io.s0_nyq_val : io.s0_stage[kk];
        const float2 q = f2(s.x - qp.x * io.coef, s.y - qp.y * io.coef);
        if (io.owned) *(NYQ ? io.s0_out_nyq : io.s0_out + kk) = q;
        return project_fast(q, m);
    } else {
        const float2 X = NYQ ? io.s0_nyq_val : io.s0_stage[kk];
        const float2 U = SPX_LDG(NYQ ? io.s1_in_nyq : io.s1_in + kk);
        const float rho = io.coef, inv = io.coef2;
        const float2 Z = f2((rho * (X.x + U.x) + s.x) * inv, (rho * (X.y + U.y) + s.y) * inv);
        const float2 Un = f2(U.x + X.x - Z.x, U.y + X.y - Z.y);
        const float2 Xn = project_fast(f2(Z.x - Un.x, Z.y - Un.y), m);
        if (io.owned) {
            *(NYQ ? io.s0_out_nyq : io.s0_out + kk) = Xn;
            *(NYQ ? io.s1_out_nyq : io.s1_out + kk) = Un;
        }
        return f2(Xn.x + Un.x, Xn.y + Un.y);
    }
}

// ---- phase 1: windowed samples -> FFT32 -> twiddle -> exchange ------------------------------------
// v[n1] on entry = (x[32 n1 + 2 l], x[32 n1 + 2 l + 1]) * 0.5 * wa   (caller applies the window)
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
SPX_HD void phase1(int l, float2* v, const Tables& tb, float2* exch) {
    phase1_compute(l, v, tb);
    phase1_write(l, v, exch);
}

SPX_HD int class_a(int p) { return p; }
SPX_HD int class_b(int p) { return p == 0 ? 16 : 32 - p; }

// ---- phase 2a: read the two residue classes of lane p ----------------------------------------------
SPX_HD void phase2_read(int p, const float2* exch, float2* A, float2* B) {
    const int ka = class_a(p), kb = class_b(p);
    static_for<16>([&](auto nc) {
        constexpr int n2 = decltype(nc)::value;
        A[n2] = exch[n2 * ROW + ka];
        B[n2] = exch[n2 * ROW + kb];
    });
}

// real-FFT post-process of one (P = Zh[k], Q = Zh[M-k]) pair, Zh = Z/2:  s[k], s[M-k]
SPX_HD void post_pair(float2 P, float2 Q, float2 w, float2& sP, float2& sQ) {
    const float er = P.x + Q.x, ei = P.y - Q.y;
    const float orr = P.y + Q.y, oi = Q.x - P.x;
    const float wor = w.x * orr - w.y * oi, woi = w.x * oi + w.y * orr;
    sP = f2(er + wor, ei + woi);
    sQ = f2(er - wor, woi - ei);
}
// inverse pre-process: (h[k], h[M-k]) -> (Z'[k], Z'[M-k])
SPX_HD void pre_pair(float2 hP, float2 hQ, float2 w, float2& P, float2& Q) {
    const float Ar = hP.x + hQ.x, Ai = hP.y - hQ.y;
    const float Dr = hP.x - hQ.x, Di = hP.y + hQ.y;
    const float Gr = w.x * Dr + w.y * Di, Gi = w.x * Di - w.y * Dr;
    P = f2(Ar - Gi, Ai + Gr);
    Q = f2(Ar + Gi, Gr - Ai);
}

// bin index of the P element of pair slot j for lane p (the Q element is bin 512 - kP; lane 0 slot 0
// is the special DC / Nyquist / bin-256 slot and uses kP = 0 -> bins 0 and 256)
template <int J>
SPX_HD int slot_bin(int p) {
    if constexpr (J < 8) return p + 32 * J;
    else return p == 0 ? 16 + 32 * (J - 8) : p + 32 * J;
}

// Issue the magnitude loads of all 16 pair slots (32 bins) of this lane early, so that their latency is
// hidden behind the 16-point FFTs.  mP[j] = mag[kP], mQ[j] = mag[512 - kP]; lane 0, slot 0: mag[0], mag[256].
SPX_HD void load_mags(int p, const float* mag_row, float* mP, float* mQ) {
    static_for<16>([&](auto jc) {
        constexpr int j = decltype(jc)::value;
        const int kP = slot_bin<j>(p);
        mP[j] = mag_row[kP];
        mQ[j] = mag_row[(j == 0 && p == 0) ? 256 : M - kP];
    });
}

// ---- phase 2b: two forward FFT16 ---------------------------------------------------------------------
SPX_HD void phase2_fft(float2* A, float2* B) {
    fft16<false>(A);     // A[k2] = Zh[ka + 32 k2]
    fft16<false>(B);     // B[k2] = Zh[kb + 32 k2]
}

// ---- phase 2c: pair processing with the point-wise update, then two inverse FFT16 ---------------------
template <int OP, bool SUMS>
SPX_HD void phase2_pointwise(int p, float2* A, float2* B, const Tables& tb, const FrameIO& io, const float* mP,
                             const float* mQ, float& dsum, float& esum) {
    const bool l0 = p == 0;
    // All updates are done in place: a slot's outputs go back to the registers its inputs came from
    // (which differ between lane 0 and the other lanes, hence the predicated writes).

    // slot 0
    if (l0) {
        // lane 0: A[0] = Zh[0] -> DC and Nyquist (both real), A[8] = Zh[256] -> bin 256 = conj(Z[256])
        const float2 z0 = A[0], z8 = A[8];
        const float2 h0 = bin_update_fast<OP, SUMS, false>(io, 0, f2(2.f * (z0.x + z0.y), 0.f), mP[0], dsum, esum);
        const float2 hM = bin_update_fast<OP, SUMS, true>(io, 0, f2(2.f * (z0.x - z0.y), 0.f), io.mag_nyq_val, dsum, esum);
        const float2 h8 = bin_update_fast<OP, SUMS, false>(io, 256, f2(2.f * z8.x, -2.f * z8.y), mQ[0], dsum, esum);
        A[0] = f2(h0.x + hM.x, h0.x - hM.x);        // C2R ignores Im(DC), Im(Nyquist)
        A[8] = f2(2.f * h8.x, -2.f * h8.y);
    } else {
        const int kP = p;
        float2 sP, sQ, P, Q;
        const float2 w = tb.twr[kP];
        post_pair(A[0], B[15], w, sP, sQ);
        const float2 hP = bin_update_fast<OP, SUMS, false>(io, kP, sP, mP[0], dsum, esum);
        const float2 hQ = bin_update_fast<OP, SUMS, false>(io, M - kP, sQ, mQ[0], dsum, esum);
        pre_pair(hP, hQ, w, P, Q);
        A[0] = P; B[15] = Q;
    }
    // slots 1..15.  general lanes: (A[j], B[15-j]), bins (p + 32 j, 512 - p - 32 j)
    //               lane 0, j < 8 : (A[j], A[16-j]),   bins (32 j, 512 - 32 j)
    //               lane 0, j >= 8: (B[j-8], B[23-j]), bins (16 + 32 (j-8), 496 - 32 (j-8))
    static_for<15>([&](auto jc) {
        constexpr int j = decltype(jc)::value + 1;
        float2 P, Q;
        const int kP = slot_bin<j>(p);
        if constexpr (j < 8) {
            P = A[j];
            Q = l0 ? A[16 - j] : B[15 - j];
        } else {
            P = l0 ? B[j - 8] : A[j];
            Q = l0 ? B[23 - j] : B[15 - j];
        }
        const float2 w = tb.twr[kP];
        float2 sP, sQ;
        post_pair(P, Q, w, sP, sQ);
        const float2 hP = bin_update_fast<OP, SUMS, false>(io, kP, sP, mP[j], dsum, esum);
        const float2 hQ = bin_update_fast<OP, SUMS, false>(io, M - kP, sQ, mQ[j], dsum, esum);
        pre_pair(hP, hQ, w, P, Q);
        // scatter back (the inverse of the gather above)
        if constexpr (j < 8) {
            A[j] = P;
            if (l0) A[16 - j] = Q; else B[15 - j] = Q;
        } else {
            if (l0) { B[j - 8] = P; B[23 - j] = Q; } else { A[j] = P; B[15 - j] = Q; }
        }
    });
    // General lanes: every A[j] and B[15-j] (j = 0..15) was rewritten exactly once.
    // Lane 0: A[0], A[8] (slot 0), A[1..7], A[9..15] (slots 1..7), B[0..15] (slots 8..15).
    fft16<true>(A);      // A[n2] = Y_ka[n2]
    fft16<true>(B);
}

// ---- phase 2d: write the inverse pass-A outputs back to the exchange buffer -------------------------
SPX_HD void phase2_write(int p, float2* exch, const float2* A, const float2* B) {
    const int ka = class_a(p), kb = class_b(p);
    static_for<16>([&](auto nc) {
        constexpr int n2 = decltype(nc)::value;
        exch[n2 * ROW + ka] = A[n2];
        exch[n2 * ROW + kb] = B[n2];
    });
}

// ---- phase 3: exchange -> conj twiddle -> inverse FFT32; v[n1] = z[16 n1 + l] (unscaled) ------------
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

}  // namespace fast
}  // namespace specinv
