// Frame-range sharding of one long signal (BASELINE.json cfg5): the per-iteration overlap-add halo exchange as ONE
// kernel over NVLink peer memory instead of NCCL send/recv + two adds.
//
// Every rank owns a small receive area allocated with cudaMalloc and exported with cudaIpcGetMemHandle; the two
// neighbours map it (cudaIpcOpenMemHandle, peer access over NVLink / NVSwitch).  Per iteration a rank's kernel
//   1. PUSHES its partial overlap-add sums of the (n_fft - hop) samples it shares with a neighbour straight into
//      that neighbour's receive slot (peer stores), fences, and raises the neighbour's flag to the iteration number
//      (st.release.sys),
//   2. waits for its own flag (ld.acquire.sys) -- the neighbour's push --,
//   3. adds "left partial + right partial" in that fixed order, so both ranks hold bit-identical samples.
// Slots are double buffered by the parity of the iteration number: a rank can only be one exchange ahead of its
// neighbour (it needs the neighbour's flag k to finish exchange k), so the slot of iteration k+2 is free.
// No collective library call, no host synchronisation; a rank that arrives early spins on its flag only.
// The spin is BOUNDED (SPECINV_P2P_TIMEOUT_MS, default 10 s, measured with %globaltimer): a neighbour that died or
// took another code path must not leave this GPU in an uninterruptible kernel.  On a timeout the kernel writes the
// iteration number into the status word of its own receive area, skips the add, and every later exchange returns at
// once; the host reads the word where it synchronises anyway (specinv_halo_status) and raises.
#include <cstdlib>

#include "specinv_common.cuh"

namespace specinv {

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

struct HaloArgs {
    void* x; long long ld; int rows; long long local_len; long long ov;
    char* recv_self;                 // this rank's receive area
    char* recv_peer[2];              // the left / right neighbour's receive area (nullptr: no neighbour)
    unsigned seq;
    unsigned long long timeout_ns;   // bound of the wait for the neighbour's push
};

__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
constexpr int FLAG_STATUS = 2;       // word index (after the two arrival flags): 0 = fine, else the exchange that timed out

// byte offsets inside a receive area: slot(parity, from_side) then the two flags
__host__ __device__ inline size_t slot_off(int parity, int from_side, int rows, long long ov, size_t es) {
    return (size_t)(parity * 2 + from_side) * rows * ov * es;
}
__host__ __device__ inline size_t flags_off(int rows, long long ov, size_t es) { return 4 * (size_t)rows * ov * es; }

template <typename T>
__global__ void __launch_bounds__(1024) halo_exchange_kernel(const HaloArgs a) {
    // The next iteration's kernel is launched with programmatic stream serialization (specinv_fastw_kernel.cuh): let
    // its prologue (tensor-memory allocation, table fill from the plan -- nothing this kernel writes) run under the
    // exchange; its griddepcontrol.wait still blocks until this grid has completed and its stores are visible.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int side = blockIdx.x;                 // 0: exchange with the left neighbour, 1: with the right one
    char* peer = a.recv_peer[side];
    if (!peer) return;
    const int parity = a.seq & 1;
    const long long n = (long long)a.rows * a.ov;
    T* x = (T*)a.x;
    const long long base = side == 0 ? 0 : a.local_len - a.ov;      // my head / my tail
    // 1. push: my head arrives at the left neighbour "from its right" (1), my tail at the right one "from its left" (0)
    T* dst = (T*)(peer + slot_off(parity, side == 0 ? 1 : 0, a.rows, a.ov, sizeof(T)));
    for (long long i = threadIdx.x; i < n; i += blockDim.x) {
        const long long r = i / a.ov, c = i - r * a.ov;
        dst[i] = x[r * a.ld + base + c];
    }
    __threadfence_system();
    __syncthreads();
    unsigned* peer_flags = (unsigned*)(peer + flags_off(a.rows, a.ov, sizeof(T)));
    unsigned* my_flags = (unsigned*)(a.recv_self + flags_off(a.rows, a.ov, sizeof(T)));
    __shared__ int s_ok;
    if (threadIdx.x == 0) {
        st_release_sys(peer_flags + (side == 0 ? 1 : 0), a.seq);
        // 2. wait for the neighbour's push of this iteration -- bounded; a rank that already failed does not wait again
        bool ok = ld_acquire_sys(my_flags + FLAG_STATUS) == 0;
        if (ok) {
            const unsigned long long t0 = global_timer_ns();
            while ((int)(ld_acquire_sys(my_flags + side) - a.seq) < 0) {
                __nanosleep(100);
                if (global_timer_ns() - t0 > a.timeout_ns) { ok = false; break; }
            }
            if (!ok) atomicCAS(my_flags + FLAG_STATUS, 0u, a.seq);
        }
        s_ok = ok;
    }
    __syncthreads();
    if (!s_ok) return;               // the host will see the status word; the samples keep this rank's partial sums
    // 3. left partial + right partial
    const T* src = (const T*)(a.recv_self + slot_off(parity, side, a.rows, a.ov, sizeof(T)));
    for (long long i = threadIdx.x; i < n; i += blockDim.x) {
        const long long r = i / a.ov, c = i - r * a.ov;
        T* p = x + r * a.ld + base + c;
        const T recv = src[i];                   // plain load after the acquire above (same thread block)
        *p = side == 0 ? recv + *p : *p + recv;
    }
}

}  // namespace specinv

using namespace specinv;

extern "C" {

size_t specinv_halo_area_bytes(int dtype, int rows, int64_t ov) {
    const size_t es = dtype == SPECINV_F64 ? 8 : 4;
    return flags_off(rows, ov, es) + 64;
}

// cudaMalloc'ed (zeroed) buffer + its 64-byte IPC handle
int specinv_ipc_alloc(size_t bytes, void** dptr, void* handle64) {
    if (!dptr || !handle64 || bytes == 0) return SPECINV_ERR_INVALID;
    cudaError_t e = cudaMalloc(dptr, bytes);
    if (e != cudaSuccess) return (int)e;
    e = cudaMemset(*dptr, 0, bytes);
    if (e != cudaSuccess) return (int)e;
    e = cudaDeviceSynchronize();
    if (e != cudaSuccess) return (int)e;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    return (int)cudaIpcGetMemHandle((cudaIpcMemHandle_t*)handle64, *dptr);
}
int specinv_ipc_open(const void* handle64, void** dptr) {
    if (!dptr || !handle64) return SPECINV_ERR_INVALID;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof(h));
    return (int)cudaIpcOpenMemHandle(dptr, h, cudaIpcMemLazyEnablePeerAccess);
}
int specinv_ipc_close(void* dptr) { return dptr ? (int)cudaIpcCloseMemHandle(dptr) : SPECINV_OK; }
int specinv_ipc_free(void* dptr) { return dptr ? (int)cudaFree(dptr) : SPECINV_OK; }

static unsigned long long halo_timeout_ns() {
    static unsigned long long ns = 0;
    if (ns == 0) {
        const char* e = getenv("SPECINV_P2P_TIMEOUT_MS");
        long long ms = e ? atoll(e) : 0;
        if (ms <= 0) ms = 10000;
        ns = (unsigned long long)ms * 1000000ull;
    }
    return ns;
}

// *status = 0 while every exchange found its neighbour; otherwise the number of the first exchange that timed out.
// Synchronises `stream` (call it where the host waits anyway: evaluations, the end of the run).
int specinv_halo_status(int dtype, const void* recv_self, int rows, int64_t ov, uint32_t* status, void* stream) {
    if (!recv_self || !status || rows < 1 || ov < 1) return SPECINV_ERR_INVALID;
    const size_t es = dtype == SPECINV_F64 ? 8 : 4;
    const char* p = (const char*)recv_self + flags_off(rows, ov, es) + FLAG_STATUS * sizeof(unsigned);
    cudaError_t e = cudaMemcpyAsync(status, p, sizeof(unsigned), cudaMemcpyDeviceToHost, (cudaStream_t)stream);
    if (e != cudaSuccess) return (int)e;
    return (int)cudaStreamSynchronize((cudaStream_t)stream);
}

int specinv_halo_exchange(int dtype, void* x, int64_t ld, int rows, int64_t local_len, int64_t ov, void* recv_self,
                          void* recv_left_peer, void* recv_right_peer, uint32_t seq, void* stream) {
    if (!x || !recv_self || rows < 1 || ov < 1 || local_len < 2 * ov || seq == 0) return SPECINV_ERR_INVALID;
    if (!recv_left_peer && !recv_right_peer) return SPECINV_OK;
    HaloArgs a{x, ld, rows, local_len, ov, (char*)recv_self, {(char*)recv_left_peer, (char*)recv_right_peer}, seq,
               halo_timeout_ns()};
    if (dtype == SPECINV_F64) halo_exchange_kernel<double><<<2, 1024, 0, (cudaStream_t)stream>>>(a);
    else if (dtype == SPECINV_F32) halo_exchange_kernel<float><<<2, 1024, 0, (cudaStream_t)stream>>>(a);
    else return SPECINV_ERR_INVALID;
    return (int)cudaGetLastError();
}

}  // extern "C"
