// In-place shared-memory complex FFT passes shared by the generic tile kernels and the RTISI-LA kernel.
// `nfr` frames of M complex points each live at wb + f*Mp (padded index padidx(n)).  Forward:
// decimation in frequency (natural -> bit-reversed order), inverse: decimation in time (bit-reversed
// -> natural), conj twiddles, unnormalised.  Two radix-2 stages are fused per pass; all threads of
// the CTA cooperate and every pass ends with __syncthreads().
#pragma once

#include "specinv_common.cuh"

namespace specinv {

__device__ __forceinline__ int padidx(int n) { return n + (n >> 4); }

template <typename T>
__device__ __forceinline__ void fft_forward_inplace(cx_t<T>* wb, int nfr, int M, int Mp, const cx_t<T>* __restrict__ tw) {
    using C = cx_t<T>;
    const int tid = threadIdx.x, NT = blockDim.x;
    int S = M;
    for (; S >= 4; S >>= 2) {
        const int q4 = S >> 2, step = M / S;
        for (int idx = tid; idx < nfr * (M >> 2); idx += NT) {
            const int f = idx / (M >> 2), r = idx - f * (M >> 2);
            const int blk = r / q4, j = r - blk * q4;
            C* v = wb + f * Mp;
            const int e0 = blk * S + j;
            const int i0 = padidx(e0), i1 = padidx(e0 + q4), i2 = padidx(e0 + 2 * q4), i3 = padidx(e0 + 3 * q4);
            const C v0 = v[i0], v1 = v[i1], v2 = v[i2], v3 = v[i3];
            const C w1 = tw[j * step], w2 = tw[2 * j * step];
            const C a0 = cadd(v0, v2), a2 = cmul(csub(v0, v2), w1);
            const C a1 = cadd(v1, v3), a3 = mul_mi(cmul(csub(v1, v3), w1));
            v[i0] = cadd(a0, a1);
            v[i1] = cmul(csub(a0, a1), w2);
            v[i2] = cadd(a2, a3);
            v[i3] = cmul(csub(a2, a3), w2);
        }
        __syncthreads();
    }
    if (S == 2) {  // odd log2(M): one plain radix-2 stage of span 2 (trivial twiddle)
        for (int idx = tid; idx < nfr * (M >> 1); idx += NT) {
            const int f = idx / (M >> 1), r = idx - f * (M >> 1);
            C* v = wb + f * Mp;
            const int i0 = padidx(2 * r), i1 = padidx(2 * r + 1);
            const C v0 = v[i0], v1 = v[i1];
            v[i0] = cadd(v0, v1);
            v[i1] = csub(v0, v1);
        }
        __syncthreads();
    }
}

template <typename T>
__device__ __forceinline__ void fft_inverse_inplace(cx_t<T>* wb, int nfr, int M, int logM, int Mp, const cx_t<T>* __restrict__ tw) {
    using C = cx_t<T>;
    const int tid = threadIdx.x, NT = blockDim.x;
    int S = 4;
    if (logM & 1) {
        for (int idx = tid; idx < nfr * (M >> 1); idx += NT) {
            const int f = idx / (M >> 1), r = idx - f * (M >> 1);
            C* v = wb + f * Mp;
            const int i0 = padidx(2 * r), i1 = padidx(2 * r + 1);
            const C v0 = v[i0], v1 = v[i1];
            v[i0] = cadd(v0, v1);
            v[i1] = csub(v0, v1);
        }
        __syncthreads();
        S = 8;
    }
    for (; S <= M; S <<= 2) {
        const int q4 = S >> 2, step = M / S;
        for (int idx = tid; idx < nfr * (M >> 2); idx += NT) {
            const int f = idx / (M >> 2), r = idx - f * (M >> 2);
            const int blk = r / q4, j = r - blk * q4;
            C* v = wb + f * Mp;
            const int e0 = blk * S + j;
            const int i0 = padidx(e0), i1 = padidx(e0 + q4), i2 = padidx(e0 + 2 * q4), i3 = padidx(e0 + 3 * q4);
            const C w1 = tw[j * step], w2 = tw[2 * j * step];
            const C v0 = v[i0], v1 = cmulc(v[i1], w2), v2 = v[i2], v3 = cmulc(v[i3], w2);
            const C a0 = cadd(v0, v1), a1 = csub(v0, v1);
            const C a2 = cmulc(cadd(v2, v3), w1), a3 = mul_pi(cmulc(csub(v2, v3), w1));
            v[i0] = cadd(a0, a2);
            v[i2] = csub(a0, a2);
            v[i1] = cadd(a1, a3);
            v[i3] = csub(a1, a3);
        }
        __syncthreads();
    }
}

// real-FFT post-processing of the pair (za = Z[k], zb = Z[M-k]) -> (s[k], s[M-k]);  w = W_N^k
template <typename T>
__device__ __forceinline__ void rfft_post_pair(cx_t<T> za, cx_t<T> zb, cx_t<T> w, cx_t<T>& sA, cx_t<T>& sB) {
    const T er = T(0.5) * (za.x + zb.x), ei = T(0.5) * (za.y - zb.y);
    const T orr = T(0.5) * (za.y + zb.y), oi = T(-0.5) * (za.x - zb.x);
    const T wor = w.x * orr - w.y * oi, woi = w.x * oi + w.y * orr;
    sA = mk<T>(er + wor, ei + woi);
    sB = mk<T>(er - wor, -(ei - woi));
}
// inverse pre-processing (h[k], h[M-k]) -> (Z'[k], Z'[M-k]) with x = (1/N) IFFT_unnormalised(Z')
template <typename T>
__device__ __forceinline__ void irfft_pre_pair(cx_t<T> hA, cx_t<T> hB, cx_t<T> w, cx_t<T>& zA, cx_t<T>& zB) {
    const T Ar = hA.x + hB.x, Ai = hA.y - hB.y;
    const T Dr = hA.x - hB.x, Di = hA.y + hB.y;
    const T Gr = w.x * Dr + w.y * Di, Gi = w.x * Di - w.y * Dr;   // conj(w) * D
    zA = mk<T>(Ar - Gi, Ai + Gr);
    zB = mk<T>(Ar + Gi, Gr - Ai);
}

}  // namespace specinv
