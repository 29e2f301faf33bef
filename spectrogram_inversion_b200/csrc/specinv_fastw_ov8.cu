// hop = n_fft/8 instances of the specialised fused kernels (specinv_fastw_kernel.cuh).
#include "specinv_fastw_kernel.cuh"

namespace specinv {
namespace wfast {
SPECINV_FASTW_DEFINE_LAUNCH(8)
}  // namespace wfast
}  // namespace specinv
