// G^2 gate statistic of MP_PDE_Solver2DLEMLinG2 (experiments/models_gnn2D.py:598-603):
//     out[s] = mean over the out-edges e of node s of (t[s] - t[dst e])^2          (torch_scatter.scatter(..., edge_index[0], 'mean'))
// forward and backward, rows of 128 floats, one warp per node (a lane owns four channels), no atomics:
//   forward   walks the node's out-edges in CSC order (the order msmp_segment_reduce over csc_perm used on the [E,128]
//             tensor of squared differences that the framework gathers produced before; the squares are fused into the adds)
//   backward  dt[i] = sum_{e: src e = i} 2 (t[i] - t[dst e]) w[i]  -  sum_{e: dst e = i} 2 (t[src e] - t[i]) w[src e],
//             w[s] = g[s] * inv[s]: one pass over the node's out-edges (CSC) and one over its in-edges (CSR); the framework
//             version scattered both terms with atomic adds (index_put with accumulate).
#include "common.cuh"
#include "msmp_b200.h"

namespace msmp {

__device__ __forceinline__ float4 sqdiff4(float4 a, float4 b) {
  const float x = a.x - b.x, y = a.y - b.y, z = a.z - b.z, w = a.w - b.w;
  return make_float4(x * x, y * y, z * z, w * w);
}

__global__ void __launch_bounds__(256) k_g2_fwd(const float* __restrict__ t, const int* __restrict__ colptr,
                                                const int* __restrict__ perm, const int* __restrict__ dst,
                                                const float* __restrict__ inv, float* __restrict__ out, int N) {
  const int lane = threadIdx.x & 31;
  const int node = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (node >= N) return;
  const int k0 = __ldg(colptr + node), k1 = __ldg(colptr + node + 1);
  const float4 ti = ldg4(t + (size_t)node * 128 + 4 * lane);
  float4 sum = zero4();
  int k = k0;
  for (; k + 4 <= k1; k += 4) {       // four independent row gathers in flight; added left to right (msmp_segment_reduce's order)
    const int d0 = __ldg(dst + __ldg(perm + k)), d1 = __ldg(dst + __ldg(perm + k + 1));
    const int d2 = __ldg(dst + __ldg(perm + k + 2)), d3 = __ldg(dst + __ldg(perm + k + 3));
    const float4 a = ldg4(t + (size_t)d0 * 128 + 4 * lane), b = ldg4(t + (size_t)d1 * 128 + 4 * lane);
    const float4 c = ldg4(t + (size_t)d2 * 128 + 4 * lane), d = ldg4(t + (size_t)d3 * 128 + 4 * lane);
    sum = add4(add4(add4(add4(sum, sqdiff4(ti, a)), sqdiff4(ti, b)), sqdiff4(ti, c)), sqdiff4(ti, d));
  }
  for (; k < k1; ++k) {
    const int d0 = __ldg(dst + __ldg(perm + k));
    sum = add4(sum, sqdiff4(ti, ldg4(t + (size_t)d0 * 128 + 4 * lane)));
  }
  st4(out + (size_t)node * 128 + 4 * lane, scale4(sum, __ldg(inv + node)));
}

__global__ void __launch_bounds__(256) k_g2_bwd(const float* __restrict__ t, const float* __restrict__ g,
                                                const int* __restrict__ colptr, const int* __restrict__ perm,
                                                const int* __restrict__ src, const int* __restrict__ dst,
                                                const int* __restrict__ rowptr, const float* __restrict__ inv,
                                                float* __restrict__ dt, int N) {
  const int lane = threadIdx.x & 31;
  const int node = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (node >= N) return;
  const float4 ti = ldg4(t + (size_t)node * 128 + 4 * lane);
  float4 acc = zero4();
  {  // as the source of its out-edges
    const float4 wi = scale4(ldg4(g + (size_t)node * 128 + 4 * lane), __ldg(inv + node));
    const int k0 = __ldg(colptr + node), k1 = __ldg(colptr + node + 1);
    for (int k = k0; k < k1; ++k) {
      const int d = __ldg(dst + __ldg(perm + k));
      const float4 td = ldg4(t + (size_t)d * 128 + 4 * lane);
      acc.x += 2.f * (ti.x - td.x) * wi.x;
      acc.y += 2.f * (ti.y - td.y) * wi.y;
      acc.z += 2.f * (ti.z - td.z) * wi.z;
      acc.w += 2.f * (ti.w - td.w) * wi.w;
    }
  }
  {  // as the destination of its in-edges (CSR order)
    const int e0 = __ldg(rowptr + node), e1 = __ldg(rowptr + node + 1);
    for (int e = e0; e < e1; ++e) {
      const int s = __ldg(src + e);
      const float4 ts = ldg4(t + (size_t)s * 128 + 4 * lane);
      const float4 ws = scale4(ldg4(g + (size_t)s * 128 + 4 * lane), __ldg(inv + s));
      acc.x -= 2.f * (ts.x - ti.x) * ws.x;
      acc.y -= 2.f * (ts.y - ti.y) * ws.y;
      acc.z -= 2.f * (ts.z - ti.z) * ws.z;
      acc.w -= 2.f * (ts.w - ti.w) * ws.w;
    }
  }
  st4(dt + (size_t)node * 128 + 4 * lane, acc);
}

}  // namespace msmp

extern "C" int msmp_g2_fwd(const float* t, const int* colptr, const int* csc_perm, const int* dst, const float* inv, float* out,
                           int N, cudaStream_t stream) {
  if (!t || !colptr || !csc_perm || !dst || !inv || !out || N < 0) return MSMP_ERR_ARG;
  if (N == 0) return MSMP_OK;
  msmp::k_g2_fwd<<<(N + 7) / 8, 256, 0, stream>>>(t, colptr, csc_perm, dst, inv, out, N);
  MSMP_CHECK_LAUNCH();
  return MSMP_OK;
}

extern "C" int msmp_g2_bwd(const float* t, const float* g, const int* colptr, const int* csc_perm, const int* src, const int* dst,
                           const int* rowptr, const float* inv, float* dt, int N, cudaStream_t stream) {
  if (!t || !g || !colptr || !csc_perm || !src || !dst || !rowptr || !inv || !dt || N < 0) return MSMP_ERR_ARG;
  if (N == 0) return MSMP_OK;
  msmp::k_g2_bwd<<<(N + 7) / 8, 256, 0, stream>>>(t, g, colptr, csc_perm, src, dst, rowptr, inv, dt, N);
  MSMP_CHECK_LAUNCH();
  return MSMP_OK;
}
