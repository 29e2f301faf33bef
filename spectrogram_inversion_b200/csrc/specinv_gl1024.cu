// Fast fused iteration kernel for n_fft = 1024, hop = 256 (onesided, fp32): the headline shape.
//
// Every half-warp streams through a chunk of consecutive frames of ONE signal:
//   * the new 256 input samples of each frame arrive by cp.async into a thread-private ring in
//     shared memory (a hop is 8 of the lane's 32 sample pairs, see gl_fast_core.cuh);
//   * forward real FFT, point-wise update + projection, inverse real FFT run in registers with one
//     shared-memory exchange per direction, synchronised by __syncwarp only (no block barrier);
//   * overlap-add is a register shift-accumulate; a finished 256-sample block is multiplied by
//     1/envelope and written with coalesced 8-byte stores.
// A chunk re-computes the 3 frames before it as a halo (state not written, output not stored) so
// chunks are independent: no atomics, deterministic.  State arrays are ping-ponged (q_in != q_out)
// because a neighbouring chunk re-reads the old state of its halo frames.
#include <cstdlib>

#include "specinv_common.cuh"
#include "gl_fast_core.cuh"

namespace specinv {
namespace fast {

struct FastArgs {
    const float* x_in; float* x_out;
    const float2* s0_in;  const float2* s0_in_nyq;  float2* s0_out; float2* s0_out_nyq;
    const float2* s1_in;  const float2* s1_in_nyq;  float2* s1_out; float2* s1_out_nyq;
    const float* mag;     const float* mag_nyq;
    const float2* tw512;  const float2* twr1024;
    const float* wa; const float* ws; const float* inv_env;
    double* sums;
    float coef, coef2;
    int B, T, P, pad_mode;
    long long L;
    int chunks_per_signal, chunk_len, n_chunks;
};

__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
// cp.async.wait_all that the compiler cannot hoist above the computation of the given values (it would
// otherwise place the wait right after the copies were issued and expose their whole latency)
__device__ __forceinline__ void cp_async_wait_all_after(float& d0, float& d1, float& d2, float& d3) {
    asm volatile("cp.async.wait_all;" : "+f"(d0), "+f"(d1), "+f"(d2), "+f"(d3)::"memory");
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// Pull the state / magnitude rows of frame `row` into L2 (one 128-byte line per call and lane) so the
// point-wise stage of that frame hits L2 instead of paying a DRAM round trip in the middle of the frame.
template <int OP>
__device__ __forceinline__ void prefetch_rows(const FastArgs& a, long long row, int l) {
    const char* q = reinterpret_cast<const char*>(a.s0_in + row * M);
    prefetch_l2(q + 128 * l);
    prefetch_l2(q + 128 * (l + 16));
    prefetch_l2(reinterpret_cast<const char*>(a.mag + row * M) + 128 * l);
    if constexpr (OP == OP_ADMM) {
        const char* u = reinterpret_cast<const char*>(a.s1_in + row * M);
        prefetch_l2(u + 128 * l);
        prefetch_l2(u + 128 * (l + 16));
    }
}

// Issue the load of block u (padded samples [256 u, 256 u + 256)) of signal x into the lane's ring row.
__device__ __forceinline__ void load_block(const FastArgs& a, const float* __restrict__ x, int u, int l, float2* ring_row) {
    const long long base = (long long)u * HOP - a.P;          // unpadded index of the block's first sample
    const bool interior = base >= 0 && base + HOP <= a.L;
    if (interior) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
            cp_async8(ring_row + ((8 * u + j) & 31), x + base + 32 * j + 2 * l);
    } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const long long pp = (long long)u * HOP + 32 * j + 2 * l;
            const long long i0 = pad_index(pp, a.P, a.L, a.pad_mode), i1 = pad_index(pp + 1, a.P, a.L, a.pad_mode);
            ring_row[(8 * u + j) & 31] = f2(i0 >= 0 ? x[i0] : 0.f, i1 >= 0 ? x[i1] : 0.f);
        }
    }
}

// Block u of the output exists (is not trimmed away by the centre padding)?
__device__ __forceinline__ bool block_valid(const FastArgs& a, int u) {
    const long long base = (long long)u * HOP - a.P;
    return base >= 0 && base + HOP <= a.L;
}
// Load the 1/envelope values of block u (issued early so the latency hides behind the inverse FFT).
__device__ __forceinline__ void load_inv_env(const FastArgs& a, int u, int l, float2* ie) {
    const long long base = (long long)u * HOP - a.P;
#pragma unroll
    for (int j = 0; j < 8; ++j) ie[j] = __ldg(reinterpret_cast<const float2*>(a.inv_env + base + 32 * j + 2 * l));
}
// Store a finished block (8 sample pairs per lane) times 1/envelope.
__device__ __forceinline__ void store_block(const FastArgs& a, float* __restrict__ xo, int u, int l, const float2* blk,
                                            const float2* ie) {
    const long long base = (long long)u * HOP - a.P;
#pragma unroll
    for (int j = 0; j < 8; ++j)
        *reinterpret_cast<float2*>(xo + base + 32 * j + 2 * l) = f2(blk[j].x * ie[j].x, blk[j].y * ie[j].y);
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
// Named barrier over the warps that share one SM sub-partition (warp_id % 4): they run the frame
// pipeline in lockstep so that the (large, fully unrolled) instruction stream is fetched once per
// group instead of once per warp.  It also orders the half-warp exchanges through shared memory.
template <int THREADS>
__device__ __forceinline__ void group_barrier(int id) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(THREADS) : "memory");
}

// ---- tensor memory as a software-managed register extension -------------------------------------------
// The overlap-add carry (24 sample pairs per lane = 48 words) only matters at the end of every frame but
// would otherwise pin 48..64 registers through the FFTs.  It lives in TMEM instead: each warp owns the 32
// TMEM lanes of its sub-partition (32 * (warp % 4)) and a private range of 48 columns, and moves the carry
// with tcgen05.ld / tcgen05.st (32x32b: one lane per thread, consecutive columns).
__device__ __forceinline__ void tmem_alloc(unsigned* smem_dst, int ncols) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(d), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(unsigned taddr, int ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_ld16(unsigned taddr, float* r) {
    unsigned* u = reinterpret_cast<unsigned*>(r);
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]),
                   "=r"(u[8]), "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_st16(unsigned taddr, const float* r) {
    const unsigned* u = reinterpret_cast<const unsigned*>(r);
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
                 ::"r"(taddr), "r"(u[0]), "r"(u[1]), "r"(u[2]), "r"(u[3]), "r"(u[4]), "r"(u[5]), "r"(u[6]), "r"(u[7]),
                   "r"(u[8]), "r"(u[9]), "r"(u[10]), "r"(u[11]), "r"(u[12]), "r"(u[13]), "r"(u[14]), "r"(u[15])
                 : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_load_carry(unsigned taddr, float2* carry) {
    float* c = reinterpret_cast<float*>(carry);
    tmem_ld16(taddr, c); tmem_ld16(taddr + 16, c + 16); tmem_ld16(taddr + 32, c + 32);
    tmem_wait_ld();
}
__device__ __forceinline__ void tmem_store_carry(unsigned taddr, const float2* carry) {
    const float* c = reinterpret_cast<const float*>(carry);
    tmem_st16(taddr, c); tmem_st16(taddr + 16, c + 16); tmem_st16(taddr + 32, c + 32);
    // completion is awaited (tmem_wait_st) right before the next tmem_load_carry, one frame later
}
constexpr int HW_F2 = 2 * TBL + 256;   // float2 elements of shared memory per half-warp: ring, exchange, magnitude row
constexpr int CARRY = 24;            // sample pairs carried from frame to frame (3 hops)
constexpr int TMEM_COLS = 256;       // >= (WARPS / 4) * 2 * CARRY, power of two

// Synchronisation of one half-warp exchange step.  LOCKSTEP: barrier over the whole sub-partition group
// (keeps its warps on the same instructions); otherwise only the 16 lanes that share the frame.
template <bool LOCKSTEP, int THREADS>
__device__ __forceinline__ void exch_sync(int bar_id, unsigned hmask, bool active) {
    if constexpr (LOCKSTEP) group_barrier<THREADS>(bar_id);
    else if (active) __syncwarp(hmask);
}

#ifndef SPX_LOCKSTEP
#define SPX_LOCKSTEP 1
#endif

template <int OP, bool SUMS, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) fast_iter_kernel(const FastArgs a) {
    constexpr bool LOCKSTEP = SPX_LOCKSTEP != 0;
    static_assert((WARPS / 4) * 2 * CARRY <= TMEM_COLS, "TMEM columns");
    static_assert(WARPS % 4 == 0, "warps are grouped by SM sub-partition");
    constexpr int GROUP_THREADS = (WARPS / 4) * 32;
    extern __shared__ __align__(16) float2 sm[];
    float2* s_tw = sm;
    float2* s_wa = s_tw + TBL;
    float2* s_ws = s_wa + TBL;
    float2* s_twr = s_ws + TBL;
    float2* s_hw = s_twr + 512;

    __shared__ unsigned s_tmem_base;
    const int tid = threadIdx.x;
    if (tid < 32) tmem_alloc(&s_tmem_base, TMEM_COLS);
    for (int i = tid; i < 512; i += WARPS * 32) {
        const int n2 = i >> 5, k = i & 31;
        s_tw[n2 * ROW + k] = a.tw512[(n2 * k) & 511];
        s_wa[n2 * ROW + k] = f2(0.5f * a.wa[32 * k + 2 * n2], 0.5f * a.wa[32 * k + 2 * n2 + 1]);
        s_ws[n2 * ROW + k] = f2(a.ws[32 * k + 2 * n2], a.ws[32 * k + 2 * n2 + 1]);
        float2 t;
        if (i <= 256) t = a.twr1024[i];
        else { t = a.twr1024[512 - i]; t.x = -t.x; }              // W^k = -conj(W^(512-k))
        s_twr[i] = t;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // this warp's TMEM window: lanes 32 * (warp % 4) .. +31, columns 48 * (warp / 4) .. +47
    const unsigned taddr = s_tmem_base + ((unsigned)(32 * ((tid >> 5) & 3)) << 16) + (unsigned)(2 * CARRY * (tid >> 7));

    const int l = tid & 15;
    const int hw = tid >> 4;                                       // half-warp slot in the CTA
    const int bar_id = 1 + ((tid >> 5) & 3);                       // one named barrier per sub-partition
    const unsigned hmask = 0xFFFFu << (16 * ((tid >> 4) & 1));
    float2* ring_row = s_hw + hw * HW_F2 + l * ROW;
    float2* exch = s_hw + hw * HW_F2 + TBL;
    float* mstage = reinterpret_cast<float*>(s_hw + hw * HW_F2 + 2 * TBL);   // 512 magnitudes of the current frame
    const Tables tb{s_tw, s_wa, s_ws, s_twr};

    double dacc = 0.0, eacc = 0.0;

    const int stride = gridDim.x * 2 * WARPS;
    const int rounds = (a.n_chunks + stride - 1) / stride;
    for (int r = 0; r < rounds; ++r) {
        const int c = r * stride + hw * gridDim.x + blockIdx.x;
        const bool valid = c < a.n_chunks;
        int b = 0, t0 = 0, t1 = 0;
        if (valid) {
            b = c / a.chunks_per_signal;
            t0 = (c - b * a.chunks_per_signal) * a.chunk_len;
            t1 = min(a.T, t0 + a.chunk_len);
        }
        const int tf0 = max(0, t0 - 3);
        const float* x = a.x_in + (long long)b * a.L;
        float* xo = a.x_out + (long long)b * a.L;

        {
            float2 zero[CARRY];
#pragma unroll
            for (int i = 0; i < CARRY; ++i) zero[i] = f2(0.f, 0.f);
            tmem_store_carry(taddr, zero);
        }

        if (tf0 < t1) {
            prefetch_rows<OP>(a, (long long)b * a.T + tf0, l);
            load_block(a, x, tf0, l, ring_row);
            load_block(a, x, tf0 + 1, l, ring_row);
            load_block(a, x, tf0 + 2, l, ring_row);
            load_block(a, x, tf0 + 3, l, ring_row);
        }

        // every half-warp runs the same number of iterations (3 halo + chunk_len frames) so that the
        // group barriers stay matched; iterations outside [tf0, t1) only hit the barriers
        for (int t = t0 - 3; t < t0 + a.chunk_len; ++t) {
            const bool active = t >= tf0 && t < t1;
            const long long row = (long long)b * a.T + t;
            float2 v[32];
            float2 A[16], Bv[16];
            float mP[16], mQ[16];
            float2 s0n = f2(0.f, 0.f);
            float mgn = 0.f;
            if (active) {
                cp_async_wait_all();
                const int slot0 = (8 * t) & 31;
                static_for<16>([&](auto ic) {
                    constexpr int n1 = 2 * decltype(ic)::value;
                    const float4 rr = *reinterpret_cast<const float4*>(ring_row + ((slot0 + n1) & 31));
                    const float4 w = *reinterpret_cast<const float4*>(s_wa + l * ROW + n1);
                    v[n1] = f2(rr.x * w.x, rr.y * w.y);
                    v[n1 + 1] = f2(rr.z * w.z, rr.w * w.w);
                });
                // the oldest block's slots are free now: fetch the block frame t+1 will need
                if (t + 1 < t1) {
                    load_block(a, x, t + 4, l, ring_row);
                    prefetch_rows<OP>(a, row + 1, l);
                }
                phase1_compute(l, v, tb);
            }
            group_barrier<GROUP_THREADS>(bar_id);      // previous frame's phase-3 reads are done: exch is free
            if (active) phase1_write(l, v, exch);
            exch_sync<LOCKSTEP, GROUP_THREADS>(bar_id, hmask, active);
            if (active) phase2_read(l, exch, A, Bv);
            exch_sync<LOCKSTEP, GROUP_THREADS>(bar_id, hmask, active);      // every lane has read its classes: exch is free
            if (active) {
                // stage this frame's state row (q_in / X_in, 4 KB) in the idle exchange buffer while the
                // 16-point FFTs run: the point-wise stage then reads it from shared memory
                const float2* src = a.s0_in + row * M;
#pragma unroll
                for (int i = 0; i < 16; ++i) cp_async16(exch + 2 * (16 * i + l), src + 2 * (16 * i + l));
                const float* msrc = a.mag + row * M;
#pragma unroll
                for (int i = 0; i < 8; ++i) cp_async16(mstage + 4 * (16 * i + l), msrc + 4 * (16 * i + l));
                s0n = (l == 0) ? __ldg(a.s0_in_nyq + row) : f2(0.f, 0.f);
                mgn = (l == 0) ? __ldg(a.mag_nyq + row) : 0.f;
                phase2_fft(A, Bv);
                cp_async_wait_all_after(A[15].x, A[7].y, Bv[15].x, Bv[11].y);
            }
            exch_sync<LOCKSTEP, GROUP_THREADS>(bar_id, hmask, active);      // staged row visible to all lanes
            if (active) {
                FrameIO io;
                io.s0_stage = exch;             io.s0_in_nyq = a.s0_in_nyq + row;
                io.s0_out = a.s0_out + row * M; io.s0_out_nyq = a.s0_out_nyq + row;
                if constexpr (OP == OP_ADMM) {
                    io.s1_in = a.s1_in + row * M;   io.s1_in_nyq = a.s1_in_nyq + row;
                    io.s1_out = a.s1_out + row * M; io.s1_out_nyq = a.s1_out_nyq + row;
                }
                io.mag = a.mag + row * M; io.mag_nyq = a.mag_nyq + row;
                io.s0_nyq_val = s0n; io.mag_nyq_val = mgn;
                io.coef = a.coef; io.coef2 = a.coef2;
                io.owned = t >= t0;
                load_mags(l, mstage, mP, mQ);
                float dsum = 0.f, esum = 0.f;          // per-frame partial sums, folded into doubles below
                phase2_pointwise<OP, SUMS>(l, A, Bv, tb, io, mP, mQ, dsum, esum);
                if constexpr (SUMS) { dacc += (double)dsum; eacc += (double)esum; }
            }
            exch_sync<LOCKSTEP, GROUP_THREADS>(bar_id, hmask, active);      // staged row consumed: exch may be overwritten
            if (active) phase2_write(l, exch, A, Bv);
            exch_sync<LOCKSTEP, GROUP_THREADS>(bar_id, hmask, active);
            const bool emit = active && t >= t0 && block_valid(a, t);
            float2 ie[8];
            if (emit) load_inv_env(a, t, l, ie);
            if (active) phase3(l, v, tb, exch);
            float2 carry[CARRY];
            tmem_wait_st();
            tmem_load_carry(taddr, carry);             // warp-collective: outside the `active` branch
            if (active) {
                // windowed overlap-add: out = carry (3 hops from earlier frames) + ws * v; the first hop
                // (8 pairs) of `out` is a finished block, the other 24 pairs are the new carry
                float2 blk[8];
                static_for<16>([&](auto ic) {
                    constexpr int n1 = 2 * decltype(ic)::value;
                    const float4 w = *reinterpret_cast<const float4*>(s_ws + l * ROW + n1);
                    float2 o0 = f2(w.x * v[n1].x, w.y * v[n1].y), o1 = f2(w.z * v[n1 + 1].x, w.w * v[n1 + 1].y);
                    if constexpr (n1 < CARRY) { o0 = o0 + carry[n1]; o1 = o1 + carry[n1 + 1]; }
                    if constexpr (n1 < 8) { blk[n1] = o0; blk[n1 + 1] = o1; }
                    else { carry[n1 - 8] = o0; carry[n1 - 7] = o1; }
                });
                if (emit) store_block(a, xo, t, l, blk, ie);
            }
            tmem_store_carry(taddr, carry);
        }
        {
            float2 carry[CARRY];
            tmem_wait_st();
            tmem_load_carry(taddr, carry);
            if (valid && t1 == a.T) {     // tail of the signal: blocks T, T+1, T+2 are complete now
#pragma unroll
                for (int k = 0; k < 3; ++k)
                    if (block_valid(a, a.T + k)) {
                        float2 ie[8];
                        load_inv_env(a, a.T + k, l, ie);
                        store_block(a, xo, a.T + k, l, carry + 8 * k, ie);
                    }
            }
        }
    }

    if constexpr (SUMS) {
        double d = dacc, e = eacc;
        for (int o = 16; o > 0; o >>= 1) {
            d += __shfl_xor_sync(0xffffffffu, d, o);
            e += __shfl_xor_sync(0xffffffffu, e, o);
        }
        if ((tid & 31) == 0) { atomicAdd(a.sums, d); atomicAdd(a.sums + 1, e); }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid < 32) tmem_dealloc(s_tmem_base, TMEM_COLS);
}

static int g_sms = 0;

template <int OP, int WARPS>
static int launch(const FastArgs& a0, cudaStream_t st) {
    FastArgs a = a0;
    if (g_sms == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return SPECINV_ERR_NO_DEVICE;
        if (cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return SPECINV_ERR_NO_DEVICE;
    }
    const int grid = g_sms;
    const int slots = grid * 2 * WARPS;
    int cps = slots / a.B;
    if (cps < 1) cps = 1;
    const int min_len = 24;                       // keep the 3-frame halo below ~12 %
    if (cps > (a.T + min_len - 1) / min_len) cps = (a.T + min_len - 1) / min_len;
    if (cps < 1) cps = 1;
    a.chunk_len = (a.T + cps - 1) / cps;
    a.chunks_per_signal = (a.T + a.chunk_len - 1) / a.chunk_len;
    a.n_chunks = a.B * a.chunks_per_signal;
    const size_t smem = (size_t)(3 * TBL + 512 + 2 * WARPS * HW_F2) * sizeof(float2);
    cudaError_t e;
    if (a.sums) {
        e = cudaFuncSetAttribute(fast_iter_kernel<OP, true, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        fast_iter_kernel<OP, true, WARPS><<<grid, WARPS * 32, smem, st>>>(a);
    } else {
        e = cudaFuncSetAttribute(fast_iter_kernel<OP, false, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        fast_iter_kernel<OP, false, WARPS><<<grid, WARPS * 32, smem, st>>>(a);
    }
    return (int)cudaGetLastError();
}

// warps per CTA: 8 (255 registers / thread, no spills) or 12 (168 registers); SPECINV_FAST_WARPS overrides
template <int OP>
static int launch_cfg(const FastArgs& a, cudaStream_t st) {
    static int warps = 0;
    if (warps == 0) {
        const char* e = getenv("SPECINV_FAST_WARPS");
        warps = (e && atoi(e) == 12) ? 12 : 8;
    }
    return warps == 12 ? launch<OP, 12>(a, st) : launch<OP, 8>(a, st);
}

}  // namespace fast

// Returns SPECINV_ERR_UNSUPPORTED when the shape is not the one this kernel is specialised for.
static bool fast_applicable(const specinv_desc* d) {
    return d->dtype == SPECINV_F32 && d->n_fft == 1024 && d->hop == 256 && d->onesided;
}

static void fill_common(fast::FastArgs& a, const Dims& dm, const specinv_desc* d, const void* plan) {
    const PlanLayout pl = plan_layout(dm, d->dtype);
    const char* p = (const char*)plan;
    a.tw512 = (const float2*)(p + pl.tw); a.twr1024 = (const float2*)(p + pl.twr);
    a.wa = (const float*)(p + pl.wa); a.ws = (const float*)(p + pl.ws); a.inv_env = (const float*)(p + pl.inv_env);
    a.B = dm.B; a.T = dm.T; a.P = dm.P; a.pad_mode = dm.pad_mode; a.L = dm.L;
}

int fast_gl_iter(const specinv_desc* d, const void* plan, const void* x_in, void* x_out,
                 const void* q_in_main, const void* q_in_nyq, void* q_out_main, void* q_out_nyq,
                 const void* mag_main, const void* mag_nyq, double lr, double* sums, void* stream) {
    if (!fast_applicable(d)) return SPECINV_ERR_UNSUPPORTED;
    Dims dm; int rc = make_dims(d, &dm); if (rc) return rc;
    fast::FastArgs a{};
    fill_common(a, dm, d, plan);
    a.x_in = (const float*)x_in; a.x_out = (float*)x_out;
    a.s0_in = (const float2*)q_in_main; a.s0_in_nyq = (const float2*)q_in_nyq;
    a.s0_out = (float2*)q_out_main; a.s0_out_nyq = (float2*)q_out_nyq;
    a.mag = (const float*)mag_main; a.mag_nyq = (const float*)mag_nyq;
    a.coef = (float)lr; a.sums = sums;
    return fast::launch_cfg<fast::OP_GL>(a, (cudaStream_t)stream);
}

int fast_admm_iter(const specinv_desc* d, const void* plan, const void* x_in, void* x_out,
                   const void* X_in_main, const void* X_in_nyq, const void* U_in_main, const void* U_in_nyq,
                   void* X_out_main, void* X_out_nyq, void* U_out_main, void* U_out_nyq,
                   const void* mag_main, const void* mag_nyq, double rho, double* sums, void* stream) {
    if (!fast_applicable(d)) return SPECINV_ERR_UNSUPPORTED;
    Dims dm; int rc = make_dims(d, &dm); if (rc) return rc;
    fast::FastArgs a{};
    fill_common(a, dm, d, plan);
    a.x_in = (const float*)x_in; a.x_out = (float*)x_out;
    a.s0_in = (const float2*)X_in_main; a.s0_in_nyq = (const float2*)X_in_nyq;
    a.s0_out = (float2*)X_out_main; a.s0_out_nyq = (float2*)X_out_nyq;
    a.s1_in = (const float2*)U_in_main; a.s1_in_nyq = (const float2*)U_in_nyq;
    a.s1_out = (float2*)U_out_main; a.s1_out_nyq = (float2*)U_out_nyq;
    a.mag = (const float*)mag_main; a.mag_nyq = (const float*)mag_nyq;
    a.coef = (float)rho; a.coef2 = (float)(1.0 / (1.0 + rho)); a.sums = sums;
    return fast::launch_cfg<fast::OP_ADMM>(a, (cudaStream_t)stream);
}

}  // namespace specinv
