// Mixed-radix in-place complex FFT passes in shared memory, for ANY transform length M whose prime factors are
// <= 13 (radices 16 (fp32), 8, 4, 2, 3, 5, 7, 11, 13).  Used by the generic tile kernels (specinv_generic.cu) for n_fft that is
// not a power of two (the reference infers n_fft from the bin count, methods.py:65-68: 400, 600, 1000 ...), and as
// the faster path for the power-of-two sizes the specialised kernels do not cover.
//
// Forward: decimation in frequency, stage s = 0 .. nst-1 with radix r_s on blocks of S_s = r_s * span_s points:
//     y[q] = (sum_m v[e0 + m span] W_r^(mq)) * W_S^(jq),   stored back at e0 + q span        (e0 = blk S + j)
// which leaves bin k = q_0 + r_0 (q_1 + r_1 (q_2 + ...)) at position q_0 span_0 + q_1 span_1 + ... (mixed-radix digit
// reversal, `mr_position`).  Inverse: the transposed flow graph, stages nst-1 .. 0, conjugated twiddle BEFORE the
// butterfly; the conjugated R-point DFT is the forward one with its outputs renamed (m -> (R - m) mod R, free in the
// unrolled code), so there is one butterfly per radix.  Digit-reversed input -> natural order, unnormalised.
//
// A pass is executed by a TEAM of `nt` threads (a warp, a few warps, or a whole CTA) over `nf` frames that live at
// wb + f * Mp with the padded index padidx(n); the caller synchronises the team between passes.  Everything is
// __host__ __device__ so that tests/host_emu/test_mixed_radix.cu runs the index logic on the CPU.
#pragma once

#include <cuda_runtime.h>

#include "fft_regs.cuh"

namespace specinv {
namespace mr {

enum { MAX_STAGES = 12 };

struct Plan {
    int nst;                     // stages
    int M;                       // transform length
    int tw_n;                    // entries of the root table tw[j] = exp(-2 pi i j / tw_n); a multiple of M
    int radix[MAX_STAGES];
    int span[MAX_STAGES];        // distance of the butterfly's inputs = block size after the stage
    unsigned span_magic[MAX_STAGES];   // ceil(2^32 / span) (0: span == 1)
    unsigned nb_magic[MAX_STAGES];     // ceil(2^32 / (M / radix)) (0: M == radix)
    int tw_step[MAX_STAGES];     // W_S^(jq) = tw[tw_step * j * q]
};

inline unsigned magic_of(int d) { return d <= 1 ? 0u : (unsigned)((0x100000000ull + (unsigned)d - 1) / (unsigned)d); }

// Host: factor M into the supported radices.  Returns false when a prime factor > 13 remains.  `allow16`: radix-16
// stages (fp32: the packed 16-point butterfly of fft_regs.cuh) where they save a pass over shared memory.
inline bool make_plan(int M, int tw_n, Plan* p, bool allow16 = false) {
    if (M < 1 || tw_n % M != 0) return false;
    int rad[32]; int n = 0, m = M;
    int twos = 0;
    while (m % 2 == 0) { m /= 2; ++twos; }
    {
        // the fewest power-of-two stages, and among those the fewest radix-16 ones
        const int per = allow16 ? 4 : 3;
        const int stages = (twos + per - 1) / per;
        int n16 = 0;
        if (allow16) while (twos - 4 * n16 > 3 * (stages - n16)) ++n16;
        for (int i = 0; i < n16; ++i) rad[n++] = 16;
        int r = twos - 4 * n16;
        while (r >= 3) { rad[n++] = 8; r -= 3; }
        if (r == 2) rad[n++] = 4;
        if (r == 1) rad[n++] = 2;
    }
    const int odd[5] = {3, 5, 7, 11, 13};
    // largest odd radices first: the last stages (span 1, r) then have the small odd strides
    for (int i = 4; i >= 0; --i)
        while (m % odd[i] == 0) { if (n >= MAX_STAGES) return false; rad[n++] = odd[i]; m /= odd[i]; }
    if (m != 1 || n > MAX_STAGES) return false;
    if (n == 0) { rad[n++] = 1; }      // M == 1: a single identity stage never happens (M >= 8), kept for safety
    p->nst = n; p->M = M; p->tw_n = tw_n;
    int S = M;
    for (int s = 0; s < n; ++s) {
        p->radix[s] = rad[s];
        p->span[s] = S / rad[s];
        p->span_magic[s] = magic_of(p->span[s]);
        p->nb_magic[s] = magic_of(M / rad[s]);
        p->tw_step[s] = tw_n / S;
        S /= rad[s];
    }
    for (int s = n; s < MAX_STAGES; ++s) { p->radix[s] = 1; p->span[s] = 1; p->span_magic[s] = 0; p->nb_magic[s] = 0; p->tw_step[s] = 0; }
    return true;
}

SPX_HD int padidx(int n) { return n + (n >> 4); }
inline int padded_len(int M) { return M + (M >> 4) + 1; }

// n / d for n * d < 2^32 with magic = ceil(2^32 / d); magic == 0 means d == 1
SPX_HD int fdiv(int n, unsigned magic) {
#ifdef __CUDA_ARCH__
    return magic ? (int)__umulhi((unsigned)n, magic) : n;
#else
    return magic ? (int)(((unsigned long long)(unsigned)n * magic) >> 32) : n;
#endif
}

// position of bin k after the forward passes (un-padded)
SPX_HD int mr_position(const Plan& p, int k) {
    int pos = 0;
    for (int s = 0; s < p.nst; ++s) {
        const int q = k % p.radix[s];
        k /= p.radix[s];
        pos += q * p.span[s];
    }
    return pos;
}

template <typename C> SPX_HD C ldro(const C* p) {      // read-only table load (ld.global.nc on the device)
#ifdef __CUDA_ARCH__
    return __ldg(p);
#else
    return *p;
#endif
}
template <typename C> SPX_HD C add(C a, C b) { a.x += b.x; a.y += b.y; return a; }
template <typename C> SPX_HD C sub(C a, C b) { a.x -= b.x; a.y -= b.y; return a; }
template <typename C> SPX_HD C mul(C a, C b) { C r; r.x = a.x * b.x - a.y * b.y; r.y = a.x * b.y + a.y * b.x; return r; }
template <typename C> SPX_HD C mulc(C a, C b) { C r; r.x = a.x * b.x + a.y * b.y; r.y = a.y * b.x - a.x * b.y; return r; }   // a conj(b)
template <typename C> SPX_HD C cwise(C a, C b) { a.x *= b.x; a.y *= b.y; return a; }                 // component-wise
template <typename C> SPX_HD C mul_mi(C a) { C r; r.x = a.y; r.y = -a.x; return r; }     // * (-i)
// c + s * a (real s)
template <typename C, typename T> SPX_HD C axpy(C c, T s, C a) { c.x += s * a.x; c.y += s * a.y; return c; }
// fp32: one packed FP32x2 instruction per complex add / real scale, two per complex product (fft_regs.cuh)
SPX_HD float2 add(float2 a, float2 b) { return padd(a, b); }
SPX_HD float2 sub(float2 a, float2 b) { return psub(a, b); }
SPX_HD float2 mul(float2 a, float2 b) { return cmul2(a, b); }
SPX_HD float2 mulc(float2 a, float2 b) { return cmulc2(a, b); }
SPX_HD float2 cwise(float2 a, float2 b) { return pmul(a, b); }
SPX_HD float2 axpy(float2 c, float s, float2 a) { return pfma(f2(s, s), a, c); }

// ---- in-register forward DFTs (roots exp(-2 pi i / R)) -----------------------------------------------------
template <int R, typename T, typename C> struct Dft;

template <typename T, typename C> struct Dft<2, T, C> {
    SPX_HD static void run(C* x) { const C a = x[0], b = x[1]; x[0] = add(a, b); x[1] = sub(a, b); }
};
template <typename T, typename C> struct Dft<4, T, C> {
    SPX_HD static void run(C* x) {
        const C a0 = add(x[0], x[2]), a1 = sub(x[0], x[2]);
        const C b0 = add(x[1], x[3]), b1 = mul_mi(sub(x[1], x[3]));
        x[0] = add(a0, b0); x[2] = sub(a0, b0);
        x[1] = add(a1, b1); x[3] = sub(a1, b1);
    }
};
template <typename T, typename C> struct Dft<8, T, C> {
    SPX_HD static void run(C* x) {
        // two 4-point transforms of the even / odd inputs, then the W_8^q twist
        C e[4] = {x[0], x[2], x[4], x[6]}, o[4] = {x[1], x[3], x[5], x[7]};
        Dft<4, T, C>::run(e); Dft<4, T, C>::run(o);
        const T h = (T)0.70710678118654752440;
        C o1, o3;
        o1.x = (o[1].x + o[1].y) * h; o1.y = (o[1].y - o[1].x) * h;      // * (1 - i) / sqrt 2
        o3.x = (o[3].y - o[3].x) * h; o3.y = -(o[3].x + o[3].y) * h;     // * (-1 - i) / sqrt 2
        const C o2 = mul_mi(o[2]);
        x[0] = add(e[0], o[0]); x[4] = sub(e[0], o[0]);
        x[1] = add(e[1], o1);   x[5] = sub(e[1], o1);
        x[2] = add(e[2], o2);   x[6] = sub(e[2], o2);
        x[3] = add(e[3], o3);   x[7] = sub(e[3], o3);
    }
};
template <> struct Dft<8, float, float2> { SPX_HD static void run(float2* x) { fft8<false>(x); } };
template <> struct Dft<16, float, float2> { SPX_HD static void run(float2* x) { fft16<false>(x); } };
// odd prime R: with a_m = x[m] + x[R-m], b_m = x[m] - x[R-m] (m = 1 .. h = (R-1)/2)
//   y[q]   = x0 + sum_m cos(2 pi m q / R) a_m - i sum_m sin(2 pi m q / R) b_m,   y[R-q] = conj-side (+ i ...)
template <int R, typename T, typename C> struct Dft {
    static_assert(R % 2 == 1 && R >= 3, "odd radix");
    SPX_HD static void run(C* x) {
        constexpr int H = (R - 1) / 2;
        C a[H], b[H];
        static_for<H>([&](auto m) { a[m] = add(x[m + 1], x[R - 1 - m]); b[m] = sub(x[m + 1], x[R - 1 - m]); });
        const C x0 = x[0];
        C y0 = x0;
        static_for<H>([&](auto m) { y0 = add(y0, a[m]); });
        x[0] = y0;
        static_for<H>([&](auto qq) {
            constexpr int q = decltype(qq)::value + 1;
            C c = x0, d; d.x = T(0); d.y = T(0);
            static_for<H>([&](auto mm) {
                constexpr int m = decltype(mm)::value + 1;
                constexpr T cs = (T)cx_cos2pi(m * q, R), sn = (T)cx_sin2pi(m * q, R);
                c = axpy(c, cs, a[mm]);
                d = axpy(d, sn, b[mm]);
            });
            // -i d = (d.y, -d.x)
            x[q].x = c.x + d.y;     x[q].y = c.y - d.x;
            x[R - q].x = c.x - d.y; x[R - q].y = c.y + d.x;
        });
    }
};

// One pass of stage `s` over `nf` frames by a team of `nt` threads (this thread is `tid`).  `scale` (or nullptr): 2 M
// reals, element n of the result is multiplied by (scale[2n], scale[2n+1]) component-wise on the way out -- the last
// inverse pass applies the synthesis window to the sample pairs it holds in registers anyway.
template <typename T, int R, bool INV, typename C>
SPX_HD void pass_r(C* wb, int nf, int Mp, const Plan& p, int s, const C* __restrict__ tw, int tid, int nt,
                   const T* __restrict__ scale) {
    const int nb = p.M / R, span = p.span[s], twstep = p.tw_step[s];
    const unsigned mg_span = p.span_magic[s], mg_nb = p.nb_magic[s];
    const int total = nf * nb;
    for (int idx = tid; idx < total; idx += nt) {
        const int f = fdiv(idx, mg_nb), b = idx - f * nb;
        const int blk = fdiv(b, mg_span), j = b - blk * span;
        C* v = wb + f * Mp;
        const int e0 = blk * span * R + j;
        C x[R];
        static_for<R>([&](auto m) { x[m] = v[padidx(e0 + m * span)]; });
        if (INV && span > 1) {      // inverse: conjugated stage twiddles BEFORE the butterfly
            const int tj = twstep * j;
            static_for<R - 1>([&](auto qq) { x[qq + 1] = mulc(x[qq + 1], ldro(tw + tj * (qq + 1))); });
        }
        Dft<R, T, C>::run(x);
        if (!INV && span > 1) {
            const int tj = twstep * j;
            static_for<R - 1>([&](auto qq) { x[qq + 1] = mul(x[qq + 1], ldro(tw + tj * (qq + 1))); });
        }
        // inverse: sum_q y[q] conj(W_R^(mq)) is output (R - m) mod R of the FORWARD butterfly -- a compile-time renaming
        static_for<R>([&](auto mm) {
            constexpr int m = decltype(mm)::value;
            C t = x[INV ? (R - m) % R : m];
            if (scale) {
                t = cwise(t, ldro(reinterpret_cast<const C*>(scale) + (e0 + m * span)));
            }
            v[padidx(e0 + m * span)] = t;
        });
    }
}

template <typename T, bool INV, typename C>
SPX_HD void pass(C* wb, int nf, int Mp, const Plan& p, int s, const C* __restrict__ tw, int tid, int nt,
                 const T* __restrict__ scale = nullptr) {
    if constexpr (sizeof(T) == 4) {
        if (p.radix[s] == 16) { pass_r<T, 16, INV>(wb, nf, Mp, p, s, tw, tid, nt, scale); return; }
    }
    switch (p.radix[s]) {
        case 8:  pass_r<T, 8, INV>(wb, nf, Mp, p, s, tw, tid, nt, scale); break;
        case 4:  pass_r<T, 4, INV>(wb, nf, Mp, p, s, tw, tid, nt, scale); break;
        case 2:  pass_r<T, 2, INV>(wb, nf, Mp, p, s, tw, tid, nt, scale); break;
        case 3:  pass_r<T, 3, INV>(wb, nf, Mp, p, s, tw, tid, nt, scale); break;
        case 5:  pass_r<T, 5, INV>(wb, nf, Mp, p, s, tw, tid, nt, scale); break;
        case 7:  pass_r<T, 7, INV>(wb, nf, Mp, p, s, tw, tid, nt, scale); break;
        case 11: pass_r<T, 11, INV>(wb, nf, Mp, p, s, tw, tid, nt, scale); break;
        case 13: pass_r<T, 13, INV>(wb, nf, Mp, p, s, tw, tid, nt, scale); break;
        default: break;
    }
}

}  // namespace mr
}  // namespace specinv
