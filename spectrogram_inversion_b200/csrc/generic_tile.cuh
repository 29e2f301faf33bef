// Argument block and point-wise stage shared by the generic tile kernels (specinv_generic.cu: radix-2^2 and direct
// DFT; specinv_generic_mr.cu: mixed radix).
#pragma once

#include "specinv_common.cuh"

namespace specinv {

enum { OP_STFT = 0, OP_ISTFT = 1, OP_GL = 2, OP_ADMM = 3 };

struct TileArgs {
    const void* x_in;
    void* x_out;
    const void* s0_in_main;  const void* s0_in_nyq;   // GL: q_in   ADMM: X_in   ISTFT: spectrum
    void* s0_out_main;       void* s0_out_nyq;        // GL: q_out  ADMM: X_out  STFT: spectrum
    const void* s1_in_main;  const void* s1_in_nyq;   // ADMM: U_in
    void* s1_out_main;       void* s1_out_nyq;        // ADMM: U_out
    const void* mag_main;    const void* mag_nyq;
    const void* tw; const void* twr; const void* wa; const void* ws; const void* inv_env;
    double* sums;
    double coef;       // GL: lr = alpha/(1+alpha)   ADMM: rho
    Dims dm;
    int tile_frames;   // owned frames per tile
    int Mp;            // padded complex elements per frame in shared memory
};

template <typename T>
struct BinIO {
    const TileArgs& a;
    long long fr;   // b*T + t
    __device__ BinIO(const TileArgs& a_, long long fr_) : a(a_), fr(fr_) {}
    __device__ __forceinline__ cx_t<T> ldc(const void* main, const void* nyq, int kk) const {
        if (a.dm.onesided && kk == a.dm.M) return ((const cx_t<T>*)nyq)[fr];
        return ((const cx_t<T>*)main)[fr * a.dm.row + kk];
    }
    __device__ __forceinline__ void stc(void* main, void* nyq, int kk, cx_t<T> v) const {
        if (a.dm.onesided && kk == a.dm.M) ((cx_t<T>*)nyq)[fr] = v;
        else ((cx_t<T>*)main)[fr * a.dm.row + kk] = v;
    }
    __device__ __forceinline__ T ldm(int kk) const {
        if (a.dm.onesided && kk == a.dm.M) return ((const T*)a.mag_nyq)[fr];
        return ((const T*)a.mag_main)[fr * a.dm.row + kk];
    }
};

// Point-wise stage for one frequency bin.  s = STFT bin of the current signal estimate.
// Returns the spectrum value that goes into the inverse transform.
template <typename T, int OP>
__device__ __forceinline__ cx_t<T> bin_update(const TileArgs& a, const BinIO<T>& io, int kk, cx_t<T> s,
                                              bool owned, bool want_sums, T& dsum, T& esum) {
    if constexpr (OP == OP_STFT) {
        io.stc(a.s0_out_main, a.s0_out_nyq, kk, s);
        return s;
    } else if constexpr (OP == OP_ISTFT) {
        return io.ldc(a.s0_in_main, a.s0_in_nyq, kk);
    } else if constexpr (OP == OP_GL) {
        // methods.py:243-247
        const T lr = (T)a.coef;
        const bool momentum = a.s0_in_main != nullptr;     // NULL state: plain Griffin-Lim (lr == 0), q_n = s
        T m = io.ldm(kk);
        cx_t<T> q = s;
        if (momentum) {
            cx_t<T> qp = io.ldc(a.s0_in_main, a.s0_in_nyq, kk);
            q = mk<T>(s.x - qp.x * lr, s.y - qp.y * lr);
        }
        if (owned) {
            if (momentum) io.stc(a.s0_out_main, a.s0_out_nyq, kk, q);
            if (want_sums) {
                T r = fast_sqrt(s.x * s.x + s.y * s.y);
                dsum += (r - m) * (r - m);
                esum += r * r;
            }
        }
        return project<T>(q, m);
    } else {
        // methods.py:467-475 with Y = X + U
        const T rho = (T)a.coef;
        const T inv = T(1) / (T(1) + rho);
        cx_t<T> X = io.ldc(a.s0_in_main, a.s0_in_nyq, kk);
        cx_t<T> U = io.ldc(a.s1_in_main, a.s1_in_nyq, kk);
        T m = io.ldm(kk);
        cx_t<T> Z = mk<T>((rho * (X.x + U.x) + s.x) * inv, (rho * (X.y + U.y) + s.y) * inv);
        cx_t<T> Un = mk<T>(U.x + X.x - Z.x, U.y + X.y - Z.y);
        cx_t<T> Xn = project<T>(mk<T>(Z.x - Un.x, Z.y - Un.y), m);
        if (owned) {
            io.stc(a.s0_out_main, a.s0_out_nyq, kk, Xn);
            io.stc(a.s1_out_main, a.s1_out_nyq, kk, Un);
            if (want_sums) {
                T r = fast_sqrt(s.x * s.x + s.y * s.y);
                dsum += (r - m) * (r - m);
                esum += r * r;
            }
        }
        return mk<T>(Xn.x + Un.x, Xn.y + Un.y);
    }
}

// implemented in specinv_generic_mr.cu: SPECINV_ERR_UNSUPPORTED when M = n_fft / 2 has a prime factor > 13
int mr_tile_launch(int dtype, int op, TileArgs& a, cudaStream_t st);

}  // namespace specinv
