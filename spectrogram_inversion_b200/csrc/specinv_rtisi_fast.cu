// RTISI-LA (torch_specinv/methods.py:273-412) for n_fft = 1024, hop = 256, look_ahead <= 3, onesided fp32 -- the
// shape of BASELINE.json's cfg3 -- as ONE persistent kernel built from the register FFT pipeline of
// gl_warp_core.cuh (16 complex values per lane, one warp per frame).
//
// A signal is owned by LA+1 warps, one per ACTIVE frame; a frame stays in its warp's registers for its whole
// life (LA+1 outer steps x max_iter inner iterations), together with its momentum spectrum (tensor memory) and
// its magnitude row (tensor memory, fetched once when the frame is born).  Per inner iteration (methods.py:365-398)
// every warp
//   * publishes its synthesis-windowed frame u = frame * w * c in shared memory (rows of 32 lanes),
//   * rebuilds ITS frame of the overlap-add y: a hop is 4 of a lane's 16 sample pairs, so the contributions of the
//     other frames are the same lane's rows shifted by 4 per frame; the kept frames' part is constant over the
//     inner iterations and waits in tensor memory,
//   * windows it (asym_window1/2 for the newest frame when asked), runs the forward FFT, the momentum update
//     q = S - lr * pre and the magnitude projection on the FFT outputs in registers, and the inverse FFT.
// One named barrier per inner iteration (the u exchange, double buffered); the four FFT exchanges stay inside the
// warp.  After max_iter iterations the oldest frame is committed: it joins the kept ring and is overlap-added
// (window w, 1/envelope, centre trimming) into the output (methods.py:401-408).  HBM traffic: the magnitudes once,
// the signal once.
#include "specinv_common.cuh"
#include "gl_warp_core.cuh"

namespace specinv {
namespace rfast {

using namespace wfast;

constexpr int LANES = 32;
constexpr int M = 512, N = 1024, HOP = 256, KEEP = 3, NAMAX = 4;
constexpr int SIGS = 2;                       // signals per CTA (2 x 4 warps)
constexpr int WARPS = SIGS * NAMAX;

struct RArgs {
    const float* mag; const float* mag_nyq;
    float* x_out;
    const float2* tw; const float2* twr;
    const float* wa; const float* ws; const float* inv_env;
    const float* asym1; const float* asym2;   // analysis windows of the newest frame (already x forward scale)
    float coef;                               // c = hop / (w . w)
    float lr;                                 // alpha / (1 + alpha)
    int B, T, P, LA, max_iter, asymmetric;
    long long L;
};

// ---- tensor memory helpers (same conventions as specinv_fastw.cu) -----------------------------------------
__device__ __forceinline__ void tmem_alloc(unsigned* smem_dst, int ncols) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(d), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(unsigned taddr, int ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld16(unsigned taddr, float* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n\t"
                 "tcgen05.wait::ld.sync.aligned;"
                 : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7]),
                   "=f"(r[8]), "=f"(r[9]), "=f"(r[10]), "=f"(r[11]), "=f"(r[12]), "=f"(r[13]), "=f"(r[14]), "=f"(r[15])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld32(unsigned taddr, float* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n\t"
                 "tcgen05.wait::ld.sync.aligned;"
                 : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7]),
                   "=f"(r[8]), "=f"(r[9]), "=f"(r[10]), "=f"(r[11]), "=f"(r[12]), "=f"(r[13]), "=f"(r[14]), "=f"(r[15]),
                   "=f"(r[16]), "=f"(r[17]), "=f"(r[18]), "=f"(r[19]), "=f"(r[20]), "=f"(r[21]), "=f"(r[22]), "=f"(r[23]),
                   "=f"(r[24]), "=f"(r[25]), "=f"(r[26]), "=f"(r[27]), "=f"(r[28]), "=f"(r[29]), "=f"(r[30]), "=f"(r[31])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st16(unsigned taddr, const float* r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
                 ::"r"(taddr), "f"(r[0]), "f"(r[1]), "f"(r[2]), "f"(r[3]), "f"(r[4]), "f"(r[5]), "f"(r[6]), "f"(r[7]),
                   "f"(r[8]), "f"(r[9]), "f"(r[10]), "f"(r[11]), "f"(r[12]), "f"(r[13]), "f"(r[14]), "f"(r[15]) : "memory");
}
__device__ __forceinline__ void tmem_st32(unsigned taddr, const float* r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
                 "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
                 ::"r"(taddr), "f"(r[0]), "f"(r[1]), "f"(r[2]), "f"(r[3]), "f"(r[4]), "f"(r[5]), "f"(r[6]), "f"(r[7]),
                   "f"(r[8]), "f"(r[9]), "f"(r[10]), "f"(r[11]), "f"(r[12]), "f"(r[13]), "f"(r[14]), "f"(r[15]),
                   "f"(r[16]), "f"(r[17]), "f"(r[18]), "f"(r[19]), "f"(r[20]), "f"(r[21]), "f"(r[22]), "f"(r[23]),
                   "f"(r[24]), "f"(r[25]), "f"(r[26]), "f"(r[27]), "f"(r[28]), "f"(r[29]), "f"(r[30]), "f"(r[31]) : "memory");
}

// TMEM columns per lane: constant tables (identical in the four sub-partitions), then per-warp state
constexpr int TC_WA = 0, TC_WSC = 32, TC_TW1 = 64, TC_TW2 = 96, TC_TWR = 112, TC_AS1 = 128, TC_AS2 = 160, TC_WARP = 192;
constexpr int TC_PRE = 0, TC_YK = 32, TC_MAG = 64, TC_PER_WARP = 80;
constexpr int TMEM_COLS = 512;
// float2 of shared memory per signal: u of the active frames (double buffered), u of the kept frames, the output
// carry, and the two FFT exchange buffers of every warp
constexpr int ROWS = V * LANES;                                     // one frame = 16 rows of 32 lanes = 512 float2
constexpr int SIG_F2 = 2 * NAMAX * ROWS + KEEP * ROWS + ROWS + NAMAX * 2 * M;

__device__ __forceinline__ void sig_sync(int bar_id, int threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(threads) : "memory");
}

// bin offsets of the lane's 16 bins (gl_warp_core.cuh: slot j -> bins l + 64 j and 512 - l - 64 j; lane 0 special)
struct Bins {
    int pl, ph, ql, qh, q0;
    __device__ __forceinline__ int operator()(int e) const {
        const int j = e >> 1;
        return (e & 1) ? (j == 0 ? q0 : (j >= 4 ? qh : ql) - 2 * LANES * j) : (j >= 4 ? ph : pl) + 2 * LANES * j;
    }
};

__global__ void __launch_bounds__(WARPS * 32, 1) rtisi_fast_kernel(const RArgs a) {
    extern __shared__ __align__(16) float2 sm[];
    __shared__ unsigned s_tmem_base;
    __shared__ float2 s_ws[ROWS];                  // synthesis window pairs [row][lane] (commit only)
    const int tid = threadIdx.x, warp = tid >> 5, l = tid & 31;
    const int sig = warp >> 2;                     // signal slot inside the CTA
    const int p = warp & 3;                        // physical frame slot of this warp
    const int NA = a.LA + 1;
    if (warp == 0) tmem_alloc(&s_tmem_base, TMEM_COLS);
    for (int i = tid; i < ROWS; i += WARPS * 32) {
        const int row = i >> 5, ll = i & 31;
        s_ws[i] = f2(a.ws[64 * row + 2 * ll], a.ws[64 * row + 2 * ll + 1]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tlane = s_tmem_base + ((unsigned)(32 * (warp & 3)) << 16);
    if (warp < 4) {
        float t[32];
#pragma unroll
        for (int i = 0; i < V; ++i) { t[2 * i] = 0.5f * a.wa[64 * i + 2 * l]; t[2 * i + 1] = 0.5f * a.wa[64 * i + 2 * l + 1]; }
        tmem_st32(tlane + TC_WA, t);
#pragma unroll
        for (int i = 0; i < V; ++i) { t[2 * i] = a.ws[64 * i + 2 * l] * a.coef; t[2 * i + 1] = a.ws[64 * i + 2 * l + 1] * a.coef; }
        tmem_st32(tlane + TC_WSC, t);
#pragma unroll
        for (int i = 0; i < V; ++i) {
            const float2 w = a.tw[((l + 32 * (i >> 3)) * (i & 7)) & (M - 1)];
            t[2 * i] = w.x; t[2 * i + 1] = w.y;
        }
        tmem_st32(tlane + TC_TW1, t);
#pragma unroll
        for (int kb = 0; kb < 8; ++kb) {
            const float2 w = a.tw[(8 * (l & 7) * kb) & (M - 1)];
            t[2 * kb] = w.x; t[2 * kb + 1] = w.y;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int k = slot_bin_rt<LANES>(l, j);
            float2 w;
            if (k <= M / 2) w = a.twr[k];
            else { w = a.twr[M - k]; w.x = -w.x; }
            t[16 + 2 * j] = w.x; t[16 + 2 * j + 1] = w.y;
        }
        tmem_st32(tlane + TC_TW2, t);              // TW2 and TWR are adjacent
        if (a.asymmetric) {
#pragma unroll
            for (int i = 0; i < V; ++i) { t[2 * i] = 0.5f * a.asym1[64 * i + 2 * l]; t[2 * i + 1] = 0.5f * a.asym1[64 * i + 2 * l + 1]; }
            tmem_st32(tlane + TC_AS1, t);
#pragma unroll
            for (int i = 0; i < V; ++i) { t[2 * i] = 0.5f * a.asym2[64 * i + 2 * l]; t[2 * i + 1] = 0.5f * a.asym2[64 * i + 2 * l + 1]; }
            tmem_st32(tlane + TC_AS2, t);
        }
        tmem_wait_st();
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    const int b = blockIdx.x * SIGS + sig;
    if (b < a.B && p < NA) {
        const unsigned twarp = tlane + TC_WARP + TC_PER_WARP * (warp >> 2);
        float2* base = sm + sig * SIG_F2;
        float2* U = base;                              // [2][NAMAX][ROWS]
        float2* Uk = U + 2 * NAMAX * ROWS;             // [KEEP][ROWS]
        float2* carry = Uk + KEEP * ROWS;              // [ROWS]
        float2* e1 = carry + ROWS + p * 2 * M;
        float2* e2 = e1 + M;
        const int bar_id = 1 + sig, bar_threads = 32 * NA;
        const Bins bin{l, l == 0 ? -7 * LANES : l, M - l, M - l + (l == 0 ? 7 * LANES : 0), l == 0 ? M / 2 : M - l};
        float* xo = a.x_out + (long long)b * a.L;

        // ---- everything zero (methods.py:353-358)
        float2 v[V];
#pragma unroll
        for (int i = 0; i < V; ++i) v[i] = f2(0.f, 0.f);
        {
            float z[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) z[i] = 0.f;
            tmem_st32(twarp + TC_PRE, z);
            tmem_st16(twarp + TC_MAG, z);              // frames that precede the spectrogram have zero magnitude (:339)
        }
        float2 pre_nyq = f2(0.f, 0.f);
        for (int i = p * 32 + l; i < KEEP * ROWS + ROWS; i += 32 * NA) Uk[i] = f2(0.f, 0.f);    // kept frames and carry
        sig_sync(bar_id, bar_threads);
        int kslot = 0;                                 // kept ring: logical kept frame f (0 = oldest) = slot (kslot + f) % KEEP
        float mag_nyq = 0.f;

        const int steps = a.T + a.LA;
        for (int i = 0; i < steps; ++i) {
            int la = (p - i) % NA; if (la < 0) la += NA;          // logical index of this warp's frame
            const int t_frame = i + la - a.LA;                    // spectrogram frame it reconstructs
            const bool newest = la == a.LA;
            if (newest) {
                // a new frame is born in this warp: fetch its magnitude row (zero outside the spectrogram, :339)
                float mg[16];
                const bool inside = t_frame >= 0 && t_frame < a.T;
                const float* mrow = a.mag + ((long long)b * a.T + (inside ? t_frame : 0)) * M;
#pragma unroll
                for (int e = 0; e < 16; ++e) mg[e] = inside ? __ldg(mrow + bin(e)) : 0.f;
                mag_nyq = (inside && l == 0) ? __ldg(a.mag_nyq + (long long)b * a.T + t_frame) : 0.f;
                tmem_st16(twarp + TC_MAG, mg);
                if (i == 0) {
                    // zero-phase start: the newest frame = irfft(first magnitude frame + 0j) (:353-358)
                    float2 A[8], Bv[8], twr[8];
                    tmem_ld16(tlane + TC_TWR, reinterpret_cast<float*>(twr));
                    struct IO0 {
                        const float* mg; float mn;
                        __device__ __forceinline__ float2 s0(int e) const { return f2(e < 0 ? mn : mg[e], 0.f); }
                    } io0{mg, mag_nyq};
                    spectrum_pairs(l, A, Bv, twr, io0);
                    inv_pass3<LANES>(l, A, Bv, e2);
                    __syncwarp();
                    float2 tw2[8];
                    tmem_ld16(tlane + TC_TW2, reinterpret_cast<float*>(tw2));
                    inv_pass2<LANES>(l, e2, tw2, e1);
                    __syncwarp();
                    float2 tw1[V];
                    tmem_ld32(tlane + TC_TW1, reinterpret_cast<float*>(tw1));
                    inv_pass1<LANES>(l, e1, tw1, v);
                    __syncwarp();
                }
            }
            // ---- part of this frame's y that comes from the kept frames: constant over the inner iterations.
            // kept frame f (0 = oldest) starts (KEEP - f) hops before active frame 0: row i' of this frame is row
            // i' + 4 (la + KEEP - f) of the kept frame.
            {
                float2 yk[V];
#pragma unroll
                for (int r = 0; r < V; ++r) yk[r] = f2(0.f, 0.f);
#pragma unroll
                for (int f = 0; f < KEEP; ++f) {
                    const int shift = 4 * (la + KEEP - f);           // warp-uniform
                    const float2* src = Uk + ((kslot + f) % KEEP) * ROWS + l;
#pragma unroll
                    for (int r = 0; r < V - 4; ++r)                 // a kept frame is at least one hop older
                        if (r + shift < V) yk[r] = yk[r] + src[(r + shift) * LANES];
                }
                tmem_wait_st();
                tmem_st32(twarp + TC_YK, reinterpret_cast<const float*>(yk));
                tmem_wait_st();
            }

            for (int j = 0; j < a.max_iter; ++j) {
                // ---- publish u = frame * w * c (:365-368), double buffered over j
                float2* Ub = U + (j & 1) * NAMAX * ROWS;
                float2 y[V];
                {
                    float2 w[V];
                    tmem_ld32(tlane + TC_WSC, reinterpret_cast<float*>(w));
#pragma unroll
                    for (int r = 0; r < V; ++r) {
                        y[r] = pmul(v[r], w[r]);
                        Ub[p * ROWS + r * LANES + l] = y[r];
                    }
                }
                sig_sync(bar_id, bar_threads);
                // ---- this frame of the overlap-add: own u + kept part + the other active frames, shifted by 4
                // rows per frame of distance (row r of this frame = row r - 4 d of the frame d positions later)
                {
                    float2 yk[V];
                    tmem_ld32(twarp + TC_YK, reinterpret_cast<float*>(yk));
#pragma unroll
                    for (int r = 0; r < V; ++r) y[r] = y[r] + yk[r];
                }
#pragma unroll
                for (int d = -3; d <= 3; ++d) {
                    if (d == 0) continue;
                    const int lo = la + d;                          // logical index of the other frame (warp-uniform)
                    if (lo < 0 || lo > a.LA) continue;
                    int po = (lo + i) % NA;                         // its physical slot
                    const float2* src = Ub + po * ROWS + l;
#pragma unroll
                    for (int r = 0; r < V; ++r)
                        if (r - 4 * d >= 0 && r - 4 * d < V) y[r] = y[r] + src[(r - 4 * d) * LANES];
                }
                // ---- analysis window (:371-385)
                {
                    float2 w[V];
                    const unsigned wcol = (a.asymmetric && newest) ? (j ? TC_AS2 : TC_AS1) : TC_WA;
                    tmem_ld32(tlane + wcol, reinterpret_cast<float*>(w));
#pragma unroll
                    for (int r = 0; r < V; ++r) y[r] = pmul(y[r], w[r]);
                }
                {
                    float2 tw1[V];
                    tmem_ld32(tlane + TC_TW1, reinterpret_cast<float*>(tw1));
                    fwd_pass1<LANES>(l, y, tw1, e1);
                }
                __syncwarp();
                {
                    float2 tw2[8];
                    tmem_ld16(tlane + TC_TW2, reinterpret_cast<float*>(tw2));
                    fwd_pass2<LANES>(l, e1, tw2, e2);
                }
                __syncwarp();
                float2 A[8], Bv[8];
                fwd_pass3<LANES>(l, e2, A, Bv);
                // ---- momentum (:387-392), pre <- S, projection (:394-396)
                {
                    const bool mom = j > 0 || (i > 0 && !newest);
                    float2 pre[V];
                    float mg[16];
                    tmem_wait_st();
                    tmem_ld32(twarp + TC_PRE, reinterpret_cast<float*>(pre));
                    tmem_ld16(twarp + TC_MAG, mg);
                    struct IO {
                        float2* pre; const float* mg; float2 pn; float mn;
                        __device__ __forceinline__ float2 s0(int e) const { return e < 0 ? pn : pre[e]; }
                        __device__ __forceinline__ float2 s1(int) const { return f2(0.f, 0.f); }
                        __device__ __forceinline__ float mag(int e) const { return e < 0 ? mn : mg[e]; }
                        __device__ __forceinline__ void put(int e, float2 q, float2) { if (e < 0) pn = q; else pre[e] = q; }
                    } io{pre, mg, pre_nyq, mag_nyq};
                    float2 twr[8];
                    tmem_ld16(tlane + TC_TWR, reinterpret_cast<float*>(twr));
                    float ds = 0.f, es = 0.f;
                    pointwise<OP_GL, false>(l, A, Bv, twr, io, mom ? a.lr : 0.f, 0.f, ds, es);
                    pre_nyq = io.pn;
                    tmem_st32(twarp + TC_PRE, reinterpret_cast<const float*>(pre));
                }
                __syncwarp();
                inv_pass3<LANES>(l, A, Bv, e2);
                __syncwarp();
                {
                    float2 tw2[8];
                    tmem_ld16(tlane + TC_TW2, reinterpret_cast<float*>(tw2));
                    inv_pass2<LANES>(l, e2, tw2, e1);
                }
                __syncwarp();
                {
                    float2 tw1[V];
                    tmem_ld32(tlane + TC_TW1, reinterpret_cast<float*>(tw1));
                    inv_pass1<LANES>(l, e1, tw1, v);
                }
                __syncwarp();
            }

            // ---- commit the oldest active frame (:401-404) and fuse the final overlap-add (:406-408)
            if (la == 0) {
                const int t = i - a.LA;                             // index of the committed frame in the output
                if (t >= 0) {
                    float2 c[V];
#pragma unroll
                    for (int r = 0; r < V; ++r) {
                        const float2 w = s_ws[r * LANES + l];
                        c[r] = pfma(v[r], w, carry[r * LANES + l]);
                    }
                    const bool last = t == a.T - 1;                 // the last frame flushes the whole carry
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        if (k == 0 || last) {
                            const long long m0 = (long long)(t + k) * HOP - a.P;
                            if (m0 >= 0 && m0 + HOP <= a.L) {
#pragma unroll
                                for (int r = 0; r < 4; ++r) {
                                    const float2 ie = __ldg(reinterpret_cast<const float2*>(a.inv_env + m0 + 64 * r + 2 * l));
                                    const float2 val = c[4 * k + r];
                                    *reinterpret_cast<float2*>(xo + m0 + 64 * r + 2 * l) = pmul(val, ie);
                                }
                            }
                        }
                    }
#pragma unroll
                    for (int r = 0; r < V; ++r) carry[r * LANES + l] = r + 4 < V ? c[r + 4] : f2(0.f, 0.f);
                }
                // the committed frame replaces the oldest kept frame (stored as u = frame * w * c)
                {
                    float2 w[V];
                    tmem_ld32(tlane + TC_WSC, reinterpret_cast<float*>(w));
                    float2* dst = Uk + kslot * ROWS + l;
#pragma unroll
                    for (int r = 0; r < V; ++r) dst[r * LANES] = pmul(v[r], w[r]);
                }
                // this warp's frame slot becomes the newest (all-zero) active frame of the next step
#pragma unroll
                for (int r = 0; r < V; ++r) v[r] = f2(0.f, 0.f);
            }
            kslot = (kslot + 1) % KEEP;
            sig_sync(bar_id, bar_threads);              // the kept ring and the carry are in place for the next step
        }
        tmem_wait_st();
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) tmem_dealloc(s_tmem_base, TMEM_COLS);
}

}  // namespace rfast

// Returns SPECINV_ERR_UNSUPPORTED when the shape is not the one this kernel is specialised for.
int rtisi_fast(const specinv_desc* d, const Dims& dm, const void* plan, const void* mag_main, const void* mag_nyq,
               void* x_out, const void* asym1, const void* asym2, int look_ahead, int asymmetric, int max_iter,
               double alpha, double synth_coeff, cudaStream_t st) {
    if (d->dtype != SPECINV_F32 || !d->onesided || d->n_fft != 1024 || d->hop != 256) return SPECINV_ERR_UNSUPPORTED;
    const int LA = look_ahead < 0 ? dm.K : look_ahead;
    if (LA > 3 || dm.K != rfast::KEEP) return SPECINV_ERR_UNSUPPORTED;
    rfast::RArgs a{};
    const PlanLayout pl = plan_layout(dm, d->dtype);
    const char* p = (const char*)plan;
    a.tw = (const float2*)(p + pl.tw); a.twr = (const float2*)(p + pl.twr);
    a.wa = (const float*)(p + pl.wa); a.ws = (const float*)(p + pl.ws); a.inv_env = (const float*)(p + pl.inv_env);
    a.asym1 = (const float*)asym1; a.asym2 = (const float*)asym2;
    a.mag = (const float*)mag_main; a.mag_nyq = (const float*)mag_nyq; a.x_out = (float*)x_out;
    a.coef = (float)synth_coeff; a.lr = (float)(alpha / (1.0 + alpha));
    a.B = dm.B; a.T = dm.T; a.P = dm.P; a.LA = LA; a.max_iter = max_iter; a.asymmetric = asymmetric; a.L = dm.L;
    const size_t smem = (size_t)rfast::SIGS * rfast::SIG_F2 * sizeof(float2);
    cudaError_t e = cudaFuncSetAttribute(rfast::rtisi_fast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    rfast::rtisi_fast_kernel<<<(dm.B + rfast::SIGS - 1) / rfast::SIGS, rfast::WARPS * 32, smem, st>>>(a);
    return (int)cudaGetLastError();
}

}  // namespace specinv
