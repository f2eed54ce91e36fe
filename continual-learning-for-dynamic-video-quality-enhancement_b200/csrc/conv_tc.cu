// tcgen05 / TMEM / TMA implicit-GEMM convolution -- placeholder until the engine lands.
#include "common.cuh"
#include "conv_internal.cuh"

namespace nv {
bool conv_tc_fwd_supported(const nervecl_conv_params&) { return false; }
int conv_tc_fwd(const nervecl_conv_params&, cudaStream_t) { return NERVECL_EUNSUPPORTED; }
bool conv_tc_wgrad_supported(const void*, int64_t, const void*, int64_t, int, int, int, int, int, int, int) {
  return false;
}
int conv_tc_wgrad(const void*, int64_t, const void*, int64_t, float*, float*, int, int, int, int, int, int, float,
                  cudaStream_t) {
  return NERVECL_EUNSUPPORTED;
}
}  // namespace nv
