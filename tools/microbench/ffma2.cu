// Microbenchmark: issue / pipe throughput of packed FP32 (fma.rn.f32x2, add.rn.f32x2, mul.rn.f32x2) vs scalar FFMA.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2 ffma2.cu && ./ffma2
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) {
    unsigned long long d;
    asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}

template <int MODE>   // 0: scalar FFMA x 16 independent chains, 1: fma.f32x2 x 8 chains (same flops), 2: add.f32x2 x 8
__global__ void __launch_bounds__(512, 1) k(int iters, float* out, long long* cycles) {
    const int tid = threadIdx.x;
    float a[16];
    unsigned long long p[8];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = tid * 0.001f + i;
#pragma unroll
    for (int i = 0; i < 8; ++i) p[i] = ((unsigned long long)__float_as_uint(a[2 * i + 1]) << 32) | __float_as_uint(a[2 * i]);
    const float m = 1.0001f, c = 0.5f;
    const unsigned long long m2 = ((unsigned long long)__float_as_uint(m) << 32) | __float_as_uint(m);
    const unsigned long long c2 = ((unsigned long long)__float_as_uint(c) << 32) | __float_as_uint(c);
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], m, c);
        } else if (MODE == 1) {
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int i = 0; i < 8; ++i) p[i] = fma2(p[i], m2, c2);
        } else {
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int i = 0; i < 8; ++i) p[i] = add2(p[i], c2);
        }
    }
    const long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i];
#pragma unroll
    for (int i = 0; i < 8; ++i) s += __uint_as_float((unsigned)p[i]) + __uint_as_float((unsigned)(p[i] >> 32));
    out[blockIdx.x * 512 + tid] = s;
    if (tid == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name) {
    float* out; long long* cyc;
    cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 148 * 8);
    const int iters = 4000;
    k<MODE><<<148, 512>>>(iters, out, cyc);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    const int reps = 200;                       // ~0.35 s: long enough for the clock to settle where it will
    for (int r = 0; r < reps; ++r) k<MODE><<<148, 512>>>(iters, out, cyc);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    // FMA = 2 flops; add = 1 flop per lane-op
    const double flops = (MODE == 2 ? 1.0 : 2.0) * 64.0 * 512 * 148 * (double)iters * reps;
    printf("%-14s wall clock: %.2f TFLOP/s sustained over %.0f ms (all 148 SMs)\n", name, flops / (ms * 1e-3) / 1e12, ms);
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    // per iteration and thread: 64 scalar FP32 ops (or 32 packed); 512 threads = 16 warps = 4 per scheduler
    const double c = (double)h[0] / iters;
    printf("%-14s %s: %.1f cycles / iteration -> %.1f FP32 lane-ops / cycle / SM\n", name, cudaGetErrorString(e), c, 64.0 * 512 / c);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    run<0>("FFMA");
    run<1>("fma.f32x2");
    run<2>("add.f32x2");
    return 0;
}
