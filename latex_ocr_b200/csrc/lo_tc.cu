// tcgen05 + TMA kernels (placeholder until the sm_100a GEMM/conv land in this file).
#include "lo_common.cuh"
namespace lo {
bool tc_available() { return false; }
int tc_gemm_nt(const bf16*, int64_t, const bf16*, int64_t, void*, int, int64_t, int, int, int, const float*, int, int, cudaStream_t) {
  return fail(LO_ENOTSUP, "%s: not built", __func__);
}
int tc_conv3x3(const bf16*, const bf16*, const float*, const bf16*, bf16*, int, int, int, int, int, int, int, cudaStream_t) {
  return fail(LO_ENOTSUP, "%s: not built", __func__);
}
}  // namespace lo
