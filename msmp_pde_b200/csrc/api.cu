// Small element-wise entry points + ABI version.
#include "common.cuh"
#include "msmp_b200.h"

namespace msmp {
// out = g * swish'(z)   (gradient through the final Swish of GNN_Layer.update_net_2, models_gnn.py:56-58)
__global__ void k_mul_dswish(const float4* __restrict__ g, const float4* __restrict__ z, float4* __restrict__ out, size_t n4) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  float4 a = g[i], b = z[i];
  out[i] = make_float4(a.x * dswish(b.x), a.y * dswish(b.y), a.z * dswish(b.z), a.w * dswish(b.w));
}
}  // namespace msmp

extern "C" int msmp_abi_version(void) { return MSMP_B200_ABI_VERSION; }

extern "C" int msmp_mul_dswish(const float* g, const float* z, float* out, size_t n, cudaStream_t stream) {
  if (n & 3) return MSMP_ERR_ARG;
  if (n == 0) return MSMP_OK;
  size_t n4 = n / 4;
  msmp::k_mul_dswish<<<(unsigned)((n4 + 255) / 256), 256, 0, stream>>>(
      reinterpret_cast<const float4*>(g), reinterpret_cast<const float4*>(z), reinterpret_cast<float4*>(out), n4);
  MSMP_CHECK_LAUNCH();
  return MSMP_OK;
}
