// Microbenchmark: per-SM throughput of tcgen05.ld (TMEM -> registers) vs LDS.128 vs SHFL, 16 warps / SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_bw tmem_bw.cu && ./tmem_bw
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void tmem_alloc(unsigned* smem_dst, int ncols) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(d), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(unsigned taddr, int ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_ld32(unsigned taddr, unsigned* u) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]),
                   "=r"(u[8]), "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15]),
                   "=r"(u[16]), "=r"(u[17]), "=r"(u[18]), "=r"(u[19]), "=r"(u[20]), "=r"(u[21]), "=r"(u[22]), "=r"(u[23]),
                   "=r"(u[24]), "=r"(u[25]), "=r"(u[26]), "=r"(u[27]), "=r"(u[28]), "=r"(u[29]), "=r"(u[30]), "=r"(u[31])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_st32(unsigned taddr, const unsigned* u) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
                 "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
                 ::"r"(taddr), "r"(u[0]), "r"(u[1]), "r"(u[2]), "r"(u[3]), "r"(u[4]), "r"(u[5]), "r"(u[6]), "r"(u[7]),
                   "r"(u[8]), "r"(u[9]), "r"(u[10]), "r"(u[11]), "r"(u[12]), "r"(u[13]), "r"(u[14]), "r"(u[15]),
                   "r"(u[16]), "r"(u[17]), "r"(u[18]), "r"(u[19]), "r"(u[20]), "r"(u[21]), "r"(u[22]), "r"(u[23]),
                   "r"(u[24]), "r"(u[25]), "r"(u[26]), "r"(u[27]), "r"(u[28]), "r"(u[29]), "r"(u[30]), "r"(u[31])
                 : "memory");
}

template <int MODE>   // 0: tcgen05.ld x32   1: LDS.128 x8 (conflict free)   2: SHFL x32   3: tcgen05.st x32
__global__ void __launch_bounds__(512, 1) bw_kernel(int iters, unsigned* out, long long* cycles) {
    __shared__ unsigned s_base;
    __shared__ __align__(16) unsigned s_tab[16 * 32 * 20];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0) tmem_alloc(&s_base, 512);
    for (int i = tid; i < 16 * 32 * 20; i += 512) s_tab[i] = i;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned taddr = s_base + ((unsigned)(32 * (warp & 3)) << 16) + 32u * (warp >> 2);
    unsigned r[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) r[i] = tid * 32 + i;
    tmem_st32(taddr, r);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    unsigned acc = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {
            tmem_ld32(taddr, r);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int i = 0; i < 32; ++i) acc += r[i];
        } else if (MODE == 1) {
            const uint4* p = reinterpret_cast<const uint4*>(s_tab + (warp * 32 + lane) * 20);
#pragma unroll
            for (int i = 0; i < 8; ++i) { uint4 v = p[i & 3]; acc += v.x + v.y + v.z + v.w; }
            asm volatile("" ::: "memory");
        } else if (MODE == 2) {
#pragma unroll
            for (int i = 0; i < 32; ++i) { r[i] = __shfl_sync(0xffffffffu, r[i], (lane + 1 + it) & 31); }
#pragma unroll
            for (int i = 0; i < 32; ++i) acc += r[i];
        } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) r[i] += it;
            tmem_st32(taddr, r);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
    }
    const long long t1 = clock64();
    __syncthreads();
    out[blockIdx.x * 512 + tid] = acc + r[5];
    if (tid == 0) cycles[blockIdx.x] = t1 - t0;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) tmem_dealloc(s_base, 512);
}

template <int MODE>
void run(const char* name, int warps_note) {
    unsigned* out; long long* cyc;
    cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 148 * 8);
    const int iters = 2000;
    bw_kernel<MODE><<<148, 512>>>(iters, out, cyc);
    bw_kernel<MODE><<<148, 512>>>(iters, out, cyc);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double c = (double)h[0] / iters;
    // per iteration every warp moves 32 lanes * 32 words * 4 B = 4 KB; 16 warps -> 64 KB per SM
    printf("%-14s %s: %.1f cycles / iteration (16 warps x 4 KB) -> %.1f B/cycle/SM\n", name, cudaGetErrorString(e), c, 65536.0 / c);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    run<0>("tcgen05.ld x32", 16);
    run<1>("LDS.128 x8", 16);
    run<2>("SHFL x32", 16);
    run<3>("tcgen05.st x32", 16);
    return 0;
}
