// Register-resident small FFTs (4, 8, 16 points) with compile-time twiddles, on packed FP32x2 arithmetic.
// Everything is fully unrolled over compile-time indices so the arrays live in registers and the
// twiddles become immediates.  __host__ __device__ so that tests/host_emu can check them on the CPU.
#pragma once

#include <cuda_runtime.h>
#include <utility>

namespace specinv {

#define SPX_HD __host__ __device__ __forceinline__

template <int... I, typename F>
SPX_HD void static_for_impl(std::integer_sequence<int, I...>, F&& f) {
    (f(std::integral_constant<int, I>{}), ...);
}
template <int N, typename F>
SPX_HD void static_for(F&& f) {
    static_for_impl(std::make_integer_sequence<int, N>{}, static_cast<F&&>(f));
}

// constexpr sin/cos of 2*pi*k/n (Taylor series in double after octant reduction; |err| < 1e-16)
constexpr double cx_pi = 3.14159265358979323846264338327950288;
constexpr double cx_sin_small(double x) {   // |x| <= pi/4
    double term = x, sum = x;
    for (int i = 1; i < 12; ++i) { term *= -x * x / ((2 * i) * (2 * i + 1)); sum += term; }
    return sum;
}
constexpr double cx_cos_small(double x) {
    double term = 1, sum = 1;
    for (int i = 1; i < 12; ++i) { term *= -x * x / ((2 * i - 1) * (2 * i)); sum += term; }
    return sum;
}
// cos(2 pi k / n), sin(2 pi k / n) for integers, exact symmetries
constexpr double cx_cos2pi(int k, int n) {
    k %= n; if (k < 0) k += n;
    if (8 * k == 0) return 1.0;
    if (4 * k == n) return 0.0;
    if (2 * k == n) return -1.0;
    if (4 * k == 3 * n) return 0.0;
    if (2 * k > n) return cx_cos2pi(n - k, n);            // cos(2pi - a) = cos a
    if (4 * k > n) return -cx_cos2pi(n - 2 * k, 2 * n);   // cos(a) = -cos(pi - a); pi - a = 2pi (n-2k)/(2n)
    if (8 * k > n) return cx_sin_small(2 * cx_pi * (n - 4 * k) / (4.0 * n));   // cos a = sin(pi/2 - a)
    return cx_cos_small(2 * cx_pi * k / n);
}
constexpr double cx_sin2pi(int k, int n) {
    // sin a = cos(a - pi/2) = cos(2 pi (4k - n) / (4n))
    return cx_cos2pi(4 * k - n, 4 * n);
}

SPX_HD float2 f2(float x, float y) { float2 r; r.x = x; r.y = y; return r; }

// ---- packed FP32x2 arithmetic ------------------------------------------------------------------------------
// sm_100 executes add / sub / mul / fma on a PAIR of floats held in a 64-bit register pair as ONE instruction
// (FADD2 / FMUL2 / FFMA2: two pipe cycles, one issue slot); the SASS operands take per-half negation, half swap
// and scalar broadcast modifiers, so complex add is 1 instruction and complex multiply 2 (instead of 2 and 4).
// The fused kernels are instruction-issue limited, so every complex value stays a float2 = one register pair and
// all arithmetic goes through these helpers.  On the host (tests/host_emu) they are plain float code.
#ifdef __CUDA_ARCH__
__device__ __forceinline__ unsigned long long pk2(float lo, float hi) {
    unsigned long long d; asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi)); return d;
}
__device__ __forceinline__ float2 upk2(unsigned long long v) {
    float2 r; asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v)); return r;
}
__device__ __forceinline__ float2 padd(float2 a, float2 b) {
    unsigned long long d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(pk2(a.x, a.y)), "l"(pk2(b.x, b.y))); return upk2(d);
}
__device__ __forceinline__ float2 psub(float2 a, float2 b) {
    unsigned long long d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(pk2(a.x, a.y)), "l"(pk2(b.x, b.y))); return upk2(d);
}
__device__ __forceinline__ float2 pmul(float2 a, float2 b) {
    unsigned long long d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(pk2(a.x, a.y)), "l"(pk2(b.x, b.y))); return upk2(d);
}
__device__ __forceinline__ float2 pfma(float2 a, float2 b, float2 c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(pk2(a.x, a.y)), "l"(pk2(b.x, b.y)), "l"(pk2(c.x, c.y)));
    return upk2(d);
}
#else
inline float2 padd(float2 a, float2 b) { return f2(a.x + b.x, a.y + b.y); }
inline float2 psub(float2 a, float2 b) { return f2(a.x - b.x, a.y - b.y); }
inline float2 pmul(float2 a, float2 b) { return f2(a.x * b.x, a.y * b.y); }
inline float2 pfma(float2 a, float2 b, float2 c) { return f2(a.x * b.x + c.x, a.y * b.y + c.y); }
#endif
SPX_HD float2 operator+(float2 a, float2 b) { return padd(a, b); }
SPX_HD float2 operator-(float2 a, float2 b) { return psub(a, b); }
// a * w and a * conj(w): (a.x, a.x) * w' + (a.y, a.y) * w'' with w', w'' = w with a half swapped / negated
SPX_HD float2 cmul2(float2 a, float2 w) { return pfma(f2(a.x, a.x), w, pmul(f2(a.y, a.y), f2(-w.y, w.x))); }
SPX_HD float2 cmulc2(float2 a, float2 w) { return pfma(f2(a.x, a.x), f2(w.x, -w.y), pmul(f2(a.y, a.y), f2(w.y, w.x))); }
SPX_HD float2 smul2(float2 a, float s) { return pmul(a, f2(s, s)); }
// The same two products with conj(w) = (w.x, -w.y) supplied by the caller (a table): the packed instructions take a
// half-swapped or broadcast operand for free, but not a half-NEGATED one -- cmul2 / cmulc2 on a table twiddle cost a
// scalar negate and a move on top of their two packed instructions (52 + 52 of 1288 instructions per frame at
// n_fft = 1024, profiles/r02_warp_n1024_ncu_summary.txt); with the conjugate beside the twiddle they cost two.
SPX_HD float2 cmul2t(float2 a, float2 w, float2 wc) { return pfma(f2(a.x, a.x), w, pmul(f2(a.y, a.y), f2(wc.y, wc.x))); }
SPX_HD float2 cmulc2t(float2 a, float2 w, float2 wc) { return pfma(f2(a.x, a.x), wc, pmul(f2(a.y, a.y), f2(w.y, w.x))); }

// v * exp(-+ 2 pi i K / N)   (forward: minus sign; INV: plus sign), K, N compile time
template <int K, int N, bool INV>
SPX_HD float2 twiddle_mul(float2 v) {
    constexpr int k = ((K % N) + N) % N;
    if constexpr (k == 0) return v;
    else if constexpr (4 * k == N) return INV ? f2(-v.y, v.x) : f2(v.y, -v.x);       // -+ i
    else if constexpr (2 * k == N) return f2(-v.x, -v.y);
    else if constexpr (4 * k == 3 * N) return INV ? f2(v.y, -v.x) : f2(-v.y, v.x);
    else {
        constexpr float c = (float)cx_cos2pi(k, N);
        constexpr float s = (float)(INV ? cx_sin2pi(k, N) : -cx_sin2pi(k, N));
        return cmul2(v, f2(c, s));                                                    // (v.x + i v.y)(c + i s)
    }
}

// 4-point DFT in place, natural order in and out.
template <bool INV>
SPX_HD void fft4(float2& a0, float2& a1, float2& a2, float2& a3) {
    const float2 s02 = a0 + a2, d02 = a0 - a2, s13 = a1 + a3, d13 = a1 - a3;
    const float2 r = INV ? f2(-d13.y, d13.x) : f2(d13.y, -d13.x);    // (-+ i) * d13
    a0 = s02 + s13; a2 = s02 - s13; a1 = d02 + r; a3 = d02 - r;
}

// 8-point DFT in place (natural order), radix 2 x 4.
template <bool INV>
SPX_HD void fft8(float2* a) {
    float2 e0 = a[0], e1 = a[2], e2 = a[4], e3 = a[6];
    float2 o0 = a[1], o1 = a[3], o2 = a[5], o3 = a[7];
    fft4<INV>(e0, e1, e2, e3);
    fft4<INV>(o0, o1, o2, o3);
    constexpr float h = 0.70710678118654752440f;
    // W8^1 = (1 -+ i)/sqrt2, W8^2 = -+ i, W8^3 = (-1 -+ i)/sqrt2:  o * (1 -+ i) = o + (-+ i) o
    const float2 t1 = smul2(INV ? o1 + f2(-o1.y, o1.x) : o1 + f2(o1.y, -o1.x), h);
    const float2 t2 = INV ? f2(-o2.y, o2.x) : f2(o2.y, -o2.x);
    const float2 t3 = smul2(INV ? f2(-o3.y, o3.x) - o3 : f2(o3.y, -o3.x) - o3, h);
    a[0] = e0 + o0; a[4] = e0 - o0;
    a[1] = e1 + t1; a[5] = e1 - t1;
    a[2] = e2 + t2; a[6] = e2 - t2;
    a[3] = e3 + t3; a[7] = e3 - t3;
}

// 16-point DFT in place (natural order in and out): 4 x 4 with W16 twiddles.
template <bool INV>
SPX_HD void fft16(float2* a) {
    float2 t[16];
    // step 1: DFT4 over n1 of a[4 n1 + n2]  -> t[4 k1 + n2], times W16^(n2 k1)
    static_for<4>([&](auto n2c) {
        constexpr int n2 = decltype(n2c)::value;
        float2 b0 = a[n2], b1 = a[4 + n2], b2 = a[8 + n2], b3 = a[12 + n2];
        fft4<INV>(b0, b1, b2, b3);
        t[0 + n2] = b0;
        t[4 + n2] = twiddle_mul<n2 * 1, 16, INV>(b1);
        t[8 + n2] = twiddle_mul<n2 * 2, 16, INV>(b2);
        t[12 + n2] = twiddle_mul<n2 * 3, 16, INV>(b3);
    });
    // step 2: DFT4 over n2 -> X[k1 + 4 k2]
    static_for<4>([&](auto k1c) {
        constexpr int k1 = decltype(k1c)::value;
        float2 b0 = t[4 * k1], b1 = t[4 * k1 + 1], b2 = t[4 * k1 + 2], b3 = t[4 * k1 + 3];
        fft4<INV>(b0, b1, b2, b3);
        a[k1] = b0; a[k1 + 4] = b1; a[k1 + 8] = b2; a[k1 + 12] = b3;
    });
}

}  // namespace specinv
