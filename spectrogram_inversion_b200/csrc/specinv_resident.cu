// Small problems (BASELINE.json cfg1: ONE 30 s signal, 1292 frames of 2048): ALL iterations of fast Griffin-Lim in
// ONE persistent kernel with the whole state resident on chip.
//
// The per-iteration kernel (specinv_fastw_kernel.cuh) gives every warp group a frame range and re-computes the
// OV - 1 = 3 frames before it so that ranges need no communication.  With ~1.5 frames per group that halo is 70 % of
// the work, and every iteration pays a launch, a table prologue and a pass of the state through L2: 24.6 us per
// iteration at cfg1, 5x the HBM roofline.  Here instead
//   * a CTA (one per SM, launched cooperatively so that all are co-resident) owns a contiguous range of <= FPC
//     frames of one signal for the WHOLE run: their momentum spectra q and magnitudes live in shared memory
//     (read from HBM once, written once), the signal samples they cover too (x_cur);
//   * per iteration every frame is transformed exactly once (no halo): window -> real FFT -> q = s - lr q,
//     projection -> inverse FFT -> synthesis window, with the register pipeline of gl_warp_core.cuh and ONE exchange
//     buffer per warp group (read - barrier - write instead of two buffers; shared memory is what limits FPC);
//   * the windowed frames of a round (one frame per warp group) are overlap-added into the CTA's accumulator in
//     frame order (deterministic); the partial sums of the 3 hops a CTA shares with each neighbour travel through
//     a small L2-resident exchange area guarded by per-CTA epoch flags (st.release / ld.acquire, double buffered by
//     iteration parity: NEIGHBOUR synchronisation only, no grid-wide barrier); both sides add "left partial + right
//     partial", so they hold bit-identical samples;
//   * 1/envelope, the centre padding (reflect / replicate / constant sources live inside the edge CTA's own span)
//     and the next iteration's input are produced in place; the fused metric sums of evaluating iterations are
//     reduced per warp and added to sums[2 k], sums[2 k + 1].
// HBM traffic of a whole run: state once in, once out.  Reference loop: methods.py:178-190, 237-250.
#include <cstdlib>

#include "specinv_common.cuh"
#include "gl_warp_core.cuh"
#include "sm100_ptx.cuh"

namespace specinv {
namespace resident {

using namespace wfast;

struct RArgs {
    const float* x_in; float* x_out;
    const float2* q_in; const float2* q_in_nyq; float2* q_out; float2* q_out_nyq;
    const float* mag; const float* mag_nyq;
    const float2* tw; const float2* twr;
    const float* wa; const float* ws; const float* inv_env;
    double* sums;            // 2 doubles per evaluating iteration of this run, or nullptr
    float2* xb;              // exchange area [cta][parity][side][3 * HOP / 2] float2
    unsigned* flags;         // [cta] epoch flags (zeroed before the launch) + [grid] status word
    unsigned long long* prof;  // [cta][8] clock cycles per phase, summed over the run (thread 0; tools/resident_profile.py)
    float coef;
    int B, T, P, pad_mode;
    long long L;
    int cps;                 // CTAs per signal
    int fpc;                 // frame capacity of a CTA (shared-memory layout)
    int n_iters, iter0, eva_iter;
    unsigned long long timeout_ns;
};

constexpr int V = 16, WARPS = 12, NTHREADS = WARPS * 32, NR = 3;
constexpr int TC_WA = 0, TC_WS = 32, TC_TW1 = 64, TC_TW2 = 96, TC_TWR = 128, TMEM_COLS = 256;

// bytes of dynamic shared memory for a frame capacity
constexpr size_t smem_bytes(int lanes, int fpc) {
    const size_t M = 16 * lanes, HOP = M / 2, groups = WARPS / (lanes / 32);
    return groups * M * 8 + (size_t)fpc * M * 12 + 2 * (size_t)(fpc + NR) * HOP * 4 + HOP * 4 + (size_t)fpc * 16;
}

__device__ __forceinline__ void st_release_gpu(unsigned* p, unsigned v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
template <int LANES>
__device__ __forceinline__ void group_sync(int bar_id) {
    if constexpr (LANES == 32) __syncwarp();
    else asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(LANES) : "memory");
}

// ---- pass 2 of gl_warp_core.cuh split into its read and its write half, so that E1 and E2 can be the same buffer
template <int LANES>
__device__ __forceinline__ void fwd_pass2_read(int l, const float2* e1, float2* t) {
    using C = Cfg<LANES, V>;
    const int c = l & (C::RC - 1);
    static_for<C::S2>([&](auto rc) {
        constexpr int r = decltype(rc)::value;
        const int ka = (l >> C::LOGRC) + (LANES / C::RC) * r;
        static_for<C::R2 / 2>([&](auto pc) {
            constexpr int p = decltype(pc)::value;
            const float4 q = *reinterpret_cast<const float4*>(e1 + ex_addr4<C::R2, C::RC>(C::RC * ka + c, p));
            t[C::R2 * r + 2 * p] = f2(q.x, q.y); t[C::R2 * r + 2 * p + 1] = f2(q.z, q.w);
        });
    });
}
template <int LANES>
__device__ __forceinline__ void fwd_pass2_write(int l, float2* t, const float2* tw2, float2* e2) {
    using C = Cfg<LANES, V>;
    const int c = l & (C::RC - 1);
    static_for<C::S2>([&](auto rc) {
        constexpr int r = decltype(rc)::value;
        const int ka = (l >> C::LOGRC) + (LANES / C::RC) * r;
        fft_small<C::R2, false>(t + C::R2 * r);
        static_for<C::R2>([&](auto kc) {
            constexpr int kb = decltype(kc)::value;
            const float2 y = kb == 0 ? t[C::R2 * r] : cmulf(t[C::R2 * r + kb], tw2[kb]);
            e2[ex_addr<C::RC>(ka + C::R1 * kb, c)] = y;
        });
    });
}
template <int LANES>
__device__ __forceinline__ void inv_pass2_read(int l, const float2* e2, const float2* tw2, float2* t) {
    using C = Cfg<LANES, V>;
    const int c = l & (C::RC - 1);
    static_for<C::S2>([&](auto rc) {
        constexpr int r = decltype(rc)::value;
        const int ka = (l >> C::LOGRC) + (LANES / C::RC) * r;
        static_for<C::R2>([&](auto kc) {
            constexpr int kb = decltype(kc)::value;
            const float2 y = e2[ex_addr<C::RC>(ka + C::R1 * kb, c)];
            t[C::R2 * r + kb] = kb == 0 ? y : cmulcf(y, tw2[kb]);
        });
    });
}
template <int LANES>
__device__ __forceinline__ void inv_pass2_write(int l, float2* t, float2* e1) {
    using C = Cfg<LANES, V>;
    const int c = l & (C::RC - 1);
    static_for<C::S2>([&](auto rc) {
        constexpr int r = decltype(rc)::value;
        const int ka = (l >> C::LOGRC) + (LANES / C::RC) * r;
        fft_small<C::R2, true>(t + C::R2 * r);
        static_for<C::R2 / 2>([&](auto pc) {
            constexpr int p = decltype(pc)::value;
            *reinterpret_cast<float4*>(e1 + ex_addr4<C::R2, C::RC>(C::RC * ka + c, p)) =
                make_float4(t[C::R2 * r + 2 * p].x, t[C::R2 * r + 2 * p].y, t[C::R2 * r + 2 * p + 1].x, t[C::R2 * r + 2 * p + 1].y);
        });
    });
}

template <int LANES>
__global__ void __launch_bounds__(NTHREADS, 1) resident_gl_kernel(const RArgs a) {
    using C = Cfg<LANES, V>;
    constexpr int M = C::M, HOP = C::N / 4, HP2 = HOP / 2, RC = C::RC;
    constexpr int G = LANES / 32, GROUPS = WARPS / G;
    extern __shared__ __align__(16) float2 sm[];
    __shared__ unsigned s_tmem_base;
    __shared__ __align__(8) unsigned long long s_bar;
    __shared__ int s_wait_ok;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int grp = warp / G, l = tid - grp * LANES, bar_id = 1 + grp;

    // ---- which frames: CTA c of signal b owns frames [t0, t0 + nf) of that signal
    const int b = blockIdx.x / a.cps, c = blockIdx.x - b * a.cps;
    const int t0 = (int)((long long)a.T * c / a.cps);
    const int nf = (int)((long long)a.T * (c + 1) / a.cps) - t0;
    const bool has_left = c > 0, has_right = c < a.cps - 1;
    const int span = (nf + NR) * HOP;             // samples (padded coordinates [t0 HOP, t0 HOP + span))

    // ---- shared memory
    float2* E = sm;                                               // [GROUPS][M]   exchange / windowed output frame
    float2* Q = E + GROUPS * M;                                   // [fpc][M]      momentum spectra (bins 0 .. M-1)
    float* MG = reinterpret_cast<float*>(Q + (size_t)a.fpc * M);  // [fpc][M]      magnitudes
    float* XC = MG + (size_t)a.fpc * M;                           // [(fpc+3) HOP] current signal over the span
    float* XA = XC + (size_t)(a.fpc + NR) * HOP;                  // [(fpc+3) HOP] overlap-add accumulator
    float* IEP = XA + (size_t)(a.fpc + NR) * HOP;                 // [HOP]         1/envelope of an interior hop
    float2* QN = reinterpret_cast<float2*>(IEP + HOP);            // [fpc]         Nyquist bins
    float* MN = reinterpret_cast<float*>(QN + a.fpc);             // [fpc]

    if (warp == 0) tmem_alloc(&s_tmem_base, TMEM_COLS);
    const unsigned bar = (unsigned)__cvta_generic_to_shared(&s_bar);
    if (tid == 0) mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // state rows -> shared memory by TMA bulk copies (one elected thread), overlapped with everything below
    if (warp == 0 && elect_one()) {
        mbar_expect_tx(bar, (unsigned)nf * (M * 8 + M * 4));
        const unsigned q_s = (unsigned)__cvta_generic_to_shared(Q), m_s = (unsigned)__cvta_generic_to_shared(MG);
        for (int f = 0; f < nf; ++f) {
            const long long row = (long long)b * a.T + t0 + f;
            bulk_g2s(q_s + f * (M * 8), a.q_in + row * M, M * 8, bar);
            bulk_g2s(m_s + f * (M * 4), a.mag + row * M, M * 4, bar);
        }
    }
    // lane-constant tables -> tensor memory (same columns for the four sub-partitions; G divides 4)
    const unsigned tlane = s_tmem_base + ((unsigned)(32 * (warp & 3)) << 16);
    if (warp < 4) {
        const int tl = 32 * (warp % G) + (tid & 31);
        float t[32];
#pragma unroll
        for (int i = 0; i < V; ++i) { t[2 * i] = 0.5f * a.wa[2 * LANES * i + 2 * tl]; t[2 * i + 1] = 0.5f * a.wa[2 * LANES * i + 2 * tl + 1]; }
        tmem_stw<2 * V>(tlane + TC_WA, t);
#pragma unroll
        for (int i = 0; i < V; ++i) { t[2 * i] = a.ws[2 * LANES * i + 2 * tl]; t[2 * i + 1] = a.ws[2 * LANES * i + 2 * tl + 1]; }
        tmem_stw<2 * V>(tlane + TC_WS, t);
#pragma unroll
        for (int i = 0; i < V; ++i) {
            const float2 w = a.tw[((tl + LANES * (i / C::R1)) * (i % C::R1)) & (M - 1)];
            t[2 * i] = w.x; t[2 * i + 1] = w.y;
        }
        tmem_stw<2 * V>(tlane + TC_TW1, t);
#pragma unroll
        for (int kb = 0; kb < 16; ++kb) {
            const float2 w = a.tw[(C::R1 * (tl & (RC - 1)) * (kb % C::R2)) & (M - 1)];
            t[2 * kb] = w.x; t[2 * kb + 1] = w.y;
        }
        tmem_st32(tlane + TC_TW2, t);
#pragma unroll
        for (int j = 0; j < RC; ++j) {
            const int k = slot_bin_rt<LANES, V>(tl, j);
            float2 w;
            if (k <= M / 2) w = a.twr[k];
            else { w = a.twr[M - k]; w.x = -w.x; }
            t[2 * j] = w.x; t[2 * j + 1] = w.y;
        }
        tmem_stw<2 * RC>(tlane + TC_TWR, t);
        tmem_wait_st();
    }
    // signal samples of the span (centre padding resolved), empty accumulator, periodic 1/envelope, Nyquist bins
    {
        const float* x = a.x_in + (long long)b * a.L;
        for (int s = tid; s < span; s += NTHREADS) {
            const long long i = pad_index((long long)t0 * HOP + s, a.P, a.L, a.pad_mode);
            XC[s] = i >= 0 ? x[i] : 0.f;
            XA[s] = 0.f;
        }
        if (a.T > NR) for (int s = tid; s < HOP; s += NTHREADS) IEP[s] = a.inv_env[(long long)NR * HOP - a.P + s];
        for (int f = tid; f < nf; f += NTHREADS) {
            const long long row = (long long)b * a.T + t0 + f;
            QN[f] = a.q_in_nyq[row];
            MN[f] = a.mag_nyq[row];
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    mbar_wait(bar, 0);

    const int hi_adj = l == 0 ? -(RC - 1) * LANES : 0;
    const int kq0 = l == 0 ? M / 2 : M - l;
    float2* e = E + grp * M;
    float2* xa2 = reinterpret_cast<float2*>(XA);
    float2* xc2 = reinterpret_cast<float2*>(XC);
    double dacc = 0.0, eacc = 0.0;
    int n_eval = 0;
    bool failed = false;

    constexpr int XH = NR * HP2;                              // float2 per shared region
    // frame schedule (see the loop): head = first 3 frames, tail = last 3 (not counted twice), then the interior
    const int head = min(NR, nf), tail0 = max(head, nf - NR), tailc = nf - tail0;
    const int k_send = ((head + tailc - 1) / GROUPS) * GROUPS;   // first schedule position of the round that completes them
    const int k_last = ((nf - 1) / GROUPS) * GROUPS;             // first schedule position of the last round
    // the last warp is idle in the last round, and the send happened in an earlier one
    const bool early_recv = (has_left || has_right) && k_send < k_last && k_last + GROUPS - 1 >= nf;
    // phase timers (thread 0): frames, overlap-add, exchange write, neighbour wait, exchange read, normalise, padding
    unsigned long long pt[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long tk = clock64();
    auto tick = [&](int phase) {
        if (tid == 0) { const long long now = clock64(); pt[phase] += (unsigned long long)(now - tk); tk = now; }
    };
    tick(7);
    for (int it = 0; it < a.n_iters; ++it) {
        const bool eval = a.sums != nullptr && ((a.iter0 + it) % a.eva_iter) == a.eva_iter - 1;
        for (int k0 = 0; k0 < nf; k0 += GROUPS) {
            // schedule position -> frame: the frames that touch a shared region (first 3, last 3) come first, so that
            // their partial sums are on their way to the neighbours while the interior frames are transformed
            const int k = k0 + grp;
            const int f = k < head ? k : (k - head < tailc ? tail0 + (k - head) : head + (k - head - tailc));
            float2 v[V];
            if (early_recv && k0 == k_last && warp == WARPS - 1) {
                // this warp has no frame in the last round: it takes the neighbours' partial sums in the meantime
                // (sent a round ago: the flags are normally up already) -- the shared regions are not touched by the
                // interior frames, and the barriers of the overlap-add below publish the result to the CTA
                const unsigned epoch = (unsigned)it + 1u;
                const int ln = tid & 31;
                bool ok = !failed;
                if (ok && ((ln == 0 && has_left) || (ln == 1 && has_right))) {
                    const unsigned* fp = a.flags + (ln == 0 ? blockIdx.x - 1 : blockIdx.x + 1);
                    const unsigned long long tstart = global_timer_ns();
                    while (ld_acquire_gpu(fp) < epoch) {
                        if (global_timer_ns() - tstart > a.timeout_ns) { ok = false; break; }
                    }
                }
                ok = __all_sync(0xffffffffu, ok);
                if (!ok && !failed) {
                    failed = true;
                    if (ln == 0) atomicExch(a.flags + gridDim.x, epoch);
                }
                if (!failed) {
                    // all loads of a chunk in flight before the first add (8 x 16 bytes per lane)
                    float4* xa4 = reinterpret_cast<float4*>(XA);
                    constexpr int PER = XH / 2 / 32, CH = 6;                        // float4 per lane and side (6, 12, 24); chunk
                    static_assert(PER % CH == 0, "chunking");
                    for (int side = 0; side < 2; ++side) {
                        if (side == 0 ? !has_left : !has_right) continue;
                        const float4* nb = reinterpret_cast<const float4*>(
                            a.xb + ((size_t)(side == 0 ? blockIdx.x - 1 : blockIdx.x + 1) * 2 + (it & 1)) * 2 * XH + (side == 0 ? XH : 0));
                        float4* dst = xa4 + (side == 0 ? 0 : nf * HP2 / 2);
                        for (int c0 = 0; c0 < PER; c0 += CH) {
                            float4 t[CH];
#pragma unroll
                            for (int u = 0; u < CH; ++u) t[u] = __ldcg(nb + (c0 + u) * 32 + ln);
#pragma unroll
                            for (int u = 0; u < CH; ++u) {
                                const float4 m4 = dst[(c0 + u) * 32 + ln];
                                // left partial + right partial, the same order in both CTAs
                                dst[(c0 + u) * 32 + ln] = side == 0 ? make_float4(t[u].x + m4.x, t[u].y + m4.y, t[u].z + m4.z, t[u].w + m4.w)
                                                                    : make_float4(m4.x + t[u].x, m4.y + t[u].y, m4.z + t[u].z, m4.w + t[u].w);
                            }
                        }
                    }
                }
            }
            if (k < nf) {
#pragma unroll
                for (int i = 0; i < V; ++i) v[i] = xc2[f * HP2 + LANES * i + l];
                {
                    float2 w[V];
                    tmem_ldw<2 * V>(tlane + TC_WA, reinterpret_cast<float*>(w));
#pragma unroll
                    for (int i = 0; i < V; ++i) v[i] = pmul(v[i], w[i]);
                }
                {
                    float2 tw1[V];
                    tmem_ldw<2 * V>(tlane + TC_TW1, reinterpret_cast<float*>(tw1));
                    fwd_pass1<LANES, V>(l, v, tw1, e);
                }
                group_sync<LANES>(bar_id);
                fwd_pass2_read<LANES>(l, e, v);
                group_sync<LANES>(bar_id);                      // every lane has read E1: the buffer becomes E2
                {
                    float2 tw2[C::R2];
                    tmem_ldw<2 * C::R2>(tlane + TC_TW2, reinterpret_cast<float*>(tw2));
                    fwd_pass2_write<LANES>(l, v, tw2, e);
                }
                group_sync<LANES>(bar_id);
                float2 A[RC], Bv[RC];
                fwd_pass3<LANES, V>(l, e, A, Bv);
                {
                    struct IO {
                        float2* q; const float* mg; float2* qn; float mgn;
                        int pl, ph, ql, qh, q0;
                        __device__ __forceinline__ int bin(int ee) const {
                            const int j = ee >> 1;
                            return (ee & 1) ? (j == 0 ? q0 : (j >= RC / 2 ? qh : ql) - 2 * LANES * j)
                                            : (j >= RC / 2 ? ph : pl) + 2 * LANES * j;
                        }
                        __device__ __forceinline__ float2 s0(int ee) const { return ee < 0 ? *qn : q[bin(ee)]; }
                        __device__ __forceinline__ float2 s1(int) const { return f2(0.f, 0.f); }
                        __device__ __forceinline__ float mag(int ee) const { return ee < 0 ? mgn : mg[bin(ee)]; }
                        __device__ __forceinline__ void put(int ee, float2 v0, float2) const {
                            if (ee < 0) *qn = v0; else q[bin(ee)] = v0;           // in place: the lane owns these bins
                        }
                    } io{Q + (size_t)f * M, MG + (size_t)f * M, QN + f, MN[f], l, l + hi_adj, M - l, M - l - hi_adj, kq0};
                    float2 twr[RC];
                    tmem_ldw<2 * RC>(tlane + TC_TWR, reinterpret_cast<float*>(twr));
                    float dsum = 0.f, esum = 0.f;
                    if (eval) pointwise<OP_GL, true, V>(l, A, Bv, twr, io, a.coef, 0.f, dsum, esum);
                    else pointwise<OP_GL, false, V>(l, A, Bv, twr, io, a.coef, 0.f, dsum, esum);
                    dacc += (double)dsum; eacc += (double)esum;
                }
                group_sync<LANES>(bar_id);                      // every lane has read its classes from E2
                inv_pass3<LANES, V>(l, A, Bv, e);
                group_sync<LANES>(bar_id);
                {
                    float2 tw2[C::R2];
                    tmem_ldw<2 * C::R2>(tlane + TC_TW2, reinterpret_cast<float*>(tw2));
                    inv_pass2_read<LANES>(l, e, tw2, v);
                }
                group_sync<LANES>(bar_id);
                inv_pass2_write<LANES>(l, v, e);
                group_sync<LANES>(bar_id);
                {
                    float2 tw1[V];
                    tmem_ldw<2 * V>(tlane + TC_TW1, reinterpret_cast<float*>(tw1));
                    inv_pass1<LANES, V>(l, e, tw1, v);
                }
                {
                    float2 w[V];
                    tmem_ldw<2 * V>(tlane + TC_WS, reinterpret_cast<float*>(w));
#pragma unroll
                    for (int i = 0; i < V; ++i) v[i] = pmul(w[i], v[i]);                  // pair LANES i + l of the frame
                }
            }
            tick(0);
            // ---- overlap-add straight from the registers, in four steps: in step s every frame adds its hop 3 - s, so
            // the frames of a round never meet on a sample within a step and a hop block receives its frames in
            // increasing order (frame b-3 first); a CTA barrier separates the steps.  Deterministic, no atomics.
#pragma unroll
            for (int st = 0; st < 4; ++st) {
                constexpr int HPL = V / 4;                    // sample pairs per hop and lane
                const int j = 3 - st;
                if (k < nf) {
                    float2* dst = xa2 + (f + j) * HP2 + l;
#pragma unroll
                    for (int r = 0; r < HPL; ++r) dst[LANES * r] = dst[LANES * r] + v[HPL * j + r];
                }
                __syncthreads();
            }
            tick(1);
            // ---- the shared regions are complete (as far as this CTA's frames go): send them to the neighbours
            if (k0 == k_send && (has_left || has_right)) {
                float4* mine = reinterpret_cast<float4*>(a.xb + ((size_t)blockIdx.x * 2 + (it & 1)) * 2 * XH);
                const float4* xa4 = reinterpret_cast<const float4*>(XA);
                if (has_left) for (int i = tid; i < XH / 2; i += NTHREADS) __stcg(mine + i, xa4[i]);
                if (has_right) for (int i = tid; i < XH / 2; i += NTHREADS) __stcg(mine + XH / 2 + i, xa4[nf * HP2 / 2 + i]);
                __syncthreads();
                // release by ONE thread after the CTA barrier: cumulative over the other threads' stores.  It waits for
                // those stores to be performed, so it is issued by the last warp, which is idle in the next round
                // whenever the frames do not fill the groups.
                if (tid == NTHREADS - 32) st_release_gpu(a.flags + blockIdx.x, (unsigned)it + 1u);
                tick(2);
            }
        }
        if (eval) {
            double d = dacc, ee = eacc;
            for (int o = 16; o > 0; o >>= 1) {
                d += __shfl_xor_sync(0xffffffffu, d, o);
                ee += __shfl_xor_sync(0xffffffffu, ee, o);
            }
            if ((tid & 31) == 0 && (d != 0.0 || ee != 0.0)) { atomicAdd(a.sums + 2 * n_eval, d); atomicAdd(a.sums + 2 * n_eval + 1, ee); }
            ++n_eval;
        }
        dacc = 0.0; eacc = 0.0;

        // ---- partial sums of the 3 hops shared with each neighbour (L2, epoch flags, parity slots): wait, add
        // (unless the idle warp of the last round has done it already)
        if ((has_left || has_right) && !early_recv) {
            const unsigned epoch = (unsigned)it + 1u;
            if (tid == 0) s_wait_ok = 1;
            __syncthreads();
            if ((tid == 0 && has_left) || (tid == 32 && has_right)) {
                const unsigned* fp = a.flags + (tid == 0 ? blockIdx.x - 1 : blockIdx.x + 1);
                if (!failed) {
                    const unsigned long long tstart = global_timer_ns();
                    while (ld_acquire_gpu(fp) < epoch) {
                        if (global_timer_ns() - tstart > a.timeout_ns) { s_wait_ok = 0; break; }
                    }
                }
            }
            __syncthreads();
            tick(3);
            if (!s_wait_ok && !failed) {
                failed = true;                                // a neighbour never arrived: report, keep going without it
                if (tid == 0) atomicExch(a.flags + gridDim.x, epoch);
            }
            if (!failed) {
                float4* xa4 = reinterpret_cast<float4*>(XA);
                auto add4 = [](float4 p, float4 q) { return make_float4(p.x + q.x, p.y + q.y, p.z + q.z, p.w + q.w); };
                // both sides add "left partial + right partial": bit-identical samples in the two CTAs
                if (has_left) {
                    const float4* nb = reinterpret_cast<const float4*>(a.xb + ((size_t)(blockIdx.x - 1) * 2 + (it & 1)) * 2 * XH + XH);
                    for (int i = tid; i < XH / 2; i += NTHREADS) xa4[i] = add4(__ldcg(nb + i), xa4[i]);                     // its tail
                }
                if (has_right) {
                    const float4* nb = reinterpret_cast<const float4*>(a.xb + ((size_t)(blockIdx.x + 1) * 2 + (it & 1)) * 2 * XH);
                    for (int i = tid; i < XH / 2; i += NTHREADS)
                        xa4[nf * HP2 / 2 + i] = add4(xa4[nf * HP2 / 2 + i], __ldcg(nb + i));                                   // its head
                }
            }
            __syncthreads();
            tick(4);
        }
        // ---- x = sums / envelope -> next input; accumulator cleared
        for (int ps = tid; ps < span / 2; ps += NTHREADS) {
            const int hb = ps / HP2, off = ps - hb * HP2;
            const int u = t0 + hb;                            // hop block of the padded signal
            float2 ie = f2(0.f, 0.f);
            if (u >= NR && u <= a.T - 1) ie = reinterpret_cast<const float2*>(IEP)[off];
            else {
                const long long m = (long long)u * HOP - a.P + 2 * off;
                if (m >= 0 && m + 1 < a.L + 1) ie = __ldg(reinterpret_cast<const float2*>(a.inv_env + m));
            }
            xc2[ps] = pmul(xa2[ps], ie);
            xa2[ps] = f2(0.f, 0.f);
        }
        __syncthreads();
        tick(5);
        // ---- centre padding of the signal's ends (sources inside this CTA's span)
        if (a.P > 0 && (c == 0 || c == a.cps - 1)) {
            for (int j = tid; j < 2 * a.P; j += NTHREADS) {
                const bool left = j < a.P;
                if ((left && c != 0) || (!left && c != a.cps - 1)) continue;
                const long long pp = left ? j : a.P + a.L + (j - a.P);
                const long long src = pad_index(pp, a.P, a.L, a.pad_mode);
                const long long sp = pp - (long long)t0 * HOP;
                XC[sp] = src >= 0 ? XC[src + a.P - (long long)t0 * HOP] : 0.f;
            }
            __syncthreads();
            tick(6);
        }
    }
    if (tid == 0 && a.prof) {
#pragma unroll
        for (int k = 0; k < 8; ++k) a.prof[(size_t)blockIdx.x * 8 + k] = pt[k];
    }

    // ---- results: owned samples (a CTA's last 3 hops belong to its right neighbour), state rows
    {
        float* xo = a.x_out + (long long)b * a.L;
        const int own = has_right ? nf * HOP : span;
        for (int s = tid; s < own; s += NTHREADS) {
            const long long m = (long long)t0 * HOP + s - a.P;
            if (m >= 0 && m < a.L) xo[m] = XC[s];
        }
        const float4* q4 = reinterpret_cast<const float4*>(Q);
        float4* qo = reinterpret_cast<float4*>(a.q_out + ((long long)b * a.T + t0) * M);
        for (int i = tid; i < nf * (M / 2); i += NTHREADS) qo[i] = q4[i];
        for (int f = tid; f < nf; f += NTHREADS) a.q_out_nyq[(long long)b * a.T + t0 + f] = QN[f];
    }
    tmem_wait_st();
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) tmem_dealloc(s_tmem_base, TMEM_COLS);
}

static int g_sms = 0, g_smem_optin = 0;

static int device_limits() {
    if (g_sms) return SPECINV_OK;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return SPECINV_ERR_NO_DEVICE;
    if (cudaDeviceGetAttribute(&g_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess) return SPECINV_ERR_NO_DEVICE;
    if (cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return SPECINV_ERR_NO_DEVICE;
    return SPECINV_OK;
}

struct Shape { int lanes, cps, fpc, grid; size_t smem, xb_bytes, flag_bytes, ws_bytes; };

// How a (B, T) problem maps onto the CTAs, or SPECINV_ERR_UNSUPPORTED when it does not fit on chip.
static int plan_shape(const specinv_desc* d, const Dims& dm, Shape* s) {
    if (d->dtype != SPECINV_F32 || !d->onesided || d->hop * 4 != d->n_fft) return SPECINV_ERR_UNSUPPORTED;
    if (d->n_fft != 1024 && d->n_fft != 2048 && d->n_fft != 4096) return SPECINV_ERR_UNSUPPORTED;
    if (d->center && dm.pad_mode == SPECINV_PAD_CIRCULAR) return SPECINV_ERR_UNSUPPORTED;   // sources on the far end
    int rc = device_limits(); if (rc) return rc;
    s->lanes = d->n_fft / 32;
    const int static_smem = 1024;
    int fpc_max = 0;
    while (smem_bytes(s->lanes, fpc_max + 1) + static_smem <= (size_t)g_smem_optin) ++fpc_max;
    if (fpc_max < NR || dm.B > g_sms || dm.T < NR + 1) return SPECINV_ERR_UNSUPPORTED;
    int cps = g_sms / dm.B;                        // CTAs per signal: as many as there are, at least 3 frames each
    if (cps > dm.T / NR) cps = dm.T / NR;
    if (cps < 1 || (dm.T + cps - 1) / cps > fpc_max) return SPECINV_ERR_UNSUPPORTED;
    s->cps = cps; s->grid = cps * dm.B;
    s->fpc = (dm.T + cps - 1) / cps;
    s->smem = smem_bytes(s->lanes, s->fpc);
    s->xb_bytes = (size_t)s->grid * 2 * 2 * NR * (d->hop / 2) * sizeof(float2);
    s->flag_bytes = (((size_t)s->grid + 1) * sizeof(unsigned) + 15) / 16 * 16;
    s->ws_bytes = s->xb_bytes + s->flag_bytes + (size_t)s->grid * 8 * sizeof(unsigned long long);
    return SPECINV_OK;
}

template <int LANES>
static int launch(const RArgs& a, const Shape& s, cudaStream_t st) {
    cudaError_t e = cudaFuncSetAttribute(resident_gl_kernel<LANES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s.smem);
    if (e != cudaSuccess) return (int)e;
    void* args[] = {(void*)&a};
    // cooperative: every CTA must be resident, they wait for each other's flags
    e = cudaLaunchCooperativeKernel((const void*)resident_gl_kernel<LANES>, dim3(s.grid), dim3(NTHREADS), args, s.smem, st);
    return (int)e;
}

}  // namespace resident
}  // namespace specinv

using namespace specinv;

extern "C" {

int specinv_gl_run_workspace_bytes(const specinv_desc* d, size_t* bytes) {
    Dims dm; int rc = make_dims(d, &dm); if (rc) return rc;
    if (!bytes) return SPECINV_ERR_INVALID;
    resident::Shape s{};
    rc = resident::plan_shape(d, dm, &s); if (rc) return rc;
    *bytes = s.ws_bytes;
    return SPECINV_OK;
}

int specinv_gl_run(const specinv_desc* d, const void* plan, const void* x_in, void* x_out,
                   const void* q_in_main, const void* q_in_nyq, void* q_out_main, void* q_out_nyq,
                   const void* mag_main, const void* mag_nyq, double lr, int n_iters, int iter0, int eva_iter,
                   double* sums, void* workspace, void* stream) {
    if (!d || !plan || !x_in || !x_out || !q_in_main || !q_in_nyq || !q_out_main || !q_out_nyq || !mag_main || !mag_nyq ||
        !workspace || n_iters < 1 || iter0 < 0 || eva_iter < 1 || lr < 0)
        return SPECINV_ERR_INVALID;
    Dims dm; int rc = make_dims(d, &dm); if (rc) return rc;
    resident::Shape s{};
    rc = resident::plan_shape(d, dm, &s); if (rc) return rc;
    if ((((uintptr_t)q_in_main | (uintptr_t)mag_main | (uintptr_t)q_out_main | (uintptr_t)workspace) & 15) != 0)
        return SPECINV_ERR_UNSUPPORTED;
    resident::RArgs a{};
    const PlanLayout pl = plan_layout(dm, d->dtype);
    const char* p = (const char*)plan;
    a.tw = (const float2*)(p + pl.tw); a.twr = (const float2*)(p + pl.twr);
    a.wa = (const float*)(p + pl.wa); a.ws = (const float*)(p + pl.ws); a.inv_env = (const float*)(p + pl.inv_env);
    a.x_in = (const float*)x_in; a.x_out = (float*)x_out;
    a.q_in = (const float2*)q_in_main; a.q_in_nyq = (const float2*)q_in_nyq;
    a.q_out = (float2*)q_out_main; a.q_out_nyq = (float2*)q_out_nyq;
    a.mag = (const float*)mag_main; a.mag_nyq = (const float*)mag_nyq;
    a.sums = sums;
    a.xb = (float2*)workspace;
    a.flags = (unsigned*)((char*)workspace + s.xb_bytes);
    a.prof = (unsigned long long*)((char*)workspace + s.xb_bytes + s.flag_bytes);
    a.coef = (float)lr;
    a.B = dm.B; a.T = dm.T; a.P = dm.P; a.pad_mode = dm.pad_mode; a.L = dm.L;
    a.cps = s.cps; a.fpc = s.fpc;
    a.n_iters = n_iters; a.iter0 = iter0; a.eva_iter = eva_iter;
    const char* e = getenv("SPECINV_RESIDENT_TIMEOUT_MS");
    long long ms = e ? atoll(e) : 0;
    a.timeout_ns = (unsigned long long)(ms > 0 ? ms : 5000) * 1000000ull;
    cudaError_t ce = cudaMemsetAsync(a.flags, 0, ((size_t)s.grid + 1) * sizeof(unsigned), (cudaStream_t)stream);
    if (ce != cudaSuccess) return (int)ce;
    note_other_launch((cudaStream_t)stream);
    switch (s.lanes) {
        case 32: return resident::launch<32>(a, s, (cudaStream_t)stream);
        case 64: return resident::launch<64>(a, s, (cudaStream_t)stream);
        case 128: return resident::launch<128>(a, s, (cudaStream_t)stream);
        default: return SPECINV_ERR_UNSUPPORTED;
    }
}

// 0 when every CTA of the last specinv_gl_run on this workspace found its neighbours, else the iteration (1-based) at
// which a wait timed out (SPECINV_RESIDENT_TIMEOUT_MS, default 5000).  Synchronises `stream`.
int specinv_gl_run_status(const specinv_desc* d, const void* workspace, uint32_t* status, void* stream) {
    Dims dm; int rc = make_dims(d, &dm); if (rc) return rc;
    if (!workspace || !status) return SPECINV_ERR_INVALID;
    resident::Shape s{};
    rc = resident::plan_shape(d, dm, &s); if (rc) return rc;
    const char* p = (const char*)workspace + s.xb_bytes + (size_t)s.grid * sizeof(unsigned);
    cudaError_t e = cudaMemcpyAsync(status, p, sizeof(unsigned), cudaMemcpyDeviceToHost, (cudaStream_t)stream);
    if (e != cudaSuccess) return (int)e;
    return (int)cudaStreamSynchronize((cudaStream_t)stream);
}

}  // extern "C"
