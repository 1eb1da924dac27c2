// Shared device helpers for the msmp_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define MSMP_OK 0
#define MSMP_ERR_ARG (-1)
#define MSMP_ERR_CUDA (-2)
#define MSMP_ERR_WORKSPACE (-3)

#define MSMP_H 128            // hidden width every kernel here is specialised for

#define MSMP_CHECK_LAUNCH()                                   \
  do {                                                        \
    cudaError_t e__ = cudaGetLastError();                     \
    if (e__ != cudaSuccess) return MSMP_ERR_CUDA;             \
  } while (0)

namespace msmp {

// accurate expf (not __expf): fp32 parity against the fp64 oracle is held to 1e-5 of max|ref|
__device__ __forceinline__ float sigmoidf_(float x) { return __fdividef(1.0f, 1.0f + expf(-x)); }
// MUFU-based variants (ex2.approx + rcp.approx): ~1e-6 absolute error, ~6x fewer instructions than expf/tanhf.
// Used where the transcendental count makes a kernel ALU bound (LEM gate epilogues: 3 per channel per step).
__device__ __forceinline__ float sigmoid_fast(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float tanh_fast(float x) {
  const float e = __expf(-2.0f * fabsf(x));                 // in (0, 1]: no overflow
  return copysignf(__fdividef(1.0f - e, 1.0f + e), x);
}
// tanh from the accurate expf (2 ulp) and one fast division: ~1e-7 absolute error, about a third of tanhf's cost.
// (the __expf variants above were measured to push a few small bias gradients 3% over the 1e-5 parity allowance)
__device__ __forceinline__ float tanh_acc(float x) {
  const float e = expf(-2.0f * fabsf(x));
  return copysignf(__fdividef(1.0f - e, 1.0f + e), x);
}
// Accurate expf, but the division replaced by rcp.approx (1 ulp; the denominators are in [1, inf]): 9 instead of 14
// instructions per sigmoid, 10 instead of 20 per tanh (tanh x = 1 - 2 / (1 + e^{2x}); saturates correctly at +-inf).  Used by
// the persistent LEM forward kernel, whose gate epilogues are chains of dependent instructions (165 per node-channel-step).
__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float sigmoid_r(float x) { return rcp_approx(1.0f + expf(-x)); }
__device__ __forceinline__ float tanh_r(float x) { return fmaf(-2.0f, rcp_approx(1.0f + expf(2.0f * x)), 1.0f); }
// swish(x) = x * sigmoid(x)          (models_gnn.py:12-21, beta = 1)
__device__ __forceinline__ float swish(float x) { return x * sigmoidf_(x); }
// d/dx swish = s * (1 + x * (1 - s))
__device__ __forceinline__ float dswish(float x) {
  float s = sigmoidf_(x);
  return s * (1.0f + x * (1.0f - s));
}
__device__ __forceinline__ float4 swish4(float4 v) {
  return make_float4(swish(v.x), swish(v.y), swish(v.z), swish(v.w));
}
// Activations on the MUFU pipe: sigmoid(x) = rcp(1 + ex2(-x log2 e)), 5 instructions per swish instead of the 13 of
// the expf-based common.cuh version (the roles of the edge kernel and the converters / epilogues of the node GEMMs are
// instruction-issue bound).  ex2.approx / rcp.approx are 1-2 ulp; the one extra error, the
// rounding of x log2 e (6e-8 |x| relative in e^-x), reaches the result scaled by sigmoid (1 - sigmoid) and stays below
// 1e-7 of the activation: within the fp32 noise of the GEMM next to it (parity tests: tests/test_kernels_gpu.py).
__device__ __forceinline__ float sigmoid_mufu(float x) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * -1.4426950408889634f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
  return r;
}
__device__ __forceinline__ float swish_m(float x) { return x * sigmoid_mufu(x); }
__device__ __forceinline__ float dswish_m(float x) {
  const float sg = sigmoid_mufu(x);
  return sg * (1.0f + x * (1.0f - sg));
}
__device__ __forceinline__ float4 swish4_m(float4 v) { return make_float4(swish_m(v.x), swish_m(v.y), swish_m(v.z), swish_m(v.w)); }
__device__ __forceinline__ float4 dswish4_m(float4 v) { return make_float4(dswish_m(v.x), dswish_m(v.y), dswish_m(v.z), dswish_m(v.w)); }
__device__ __forceinline__ float4 add4(float4 a, float4 b) {
  return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}
__device__ __forceinline__ float4 mul4(float4 a, float4 b) {
  return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w);
}
__device__ __forceinline__ float4 scale4(float4 a, float s) {
  return make_float4(a.x * s, a.y * s, a.z * s, a.w * s);
}
__device__ __forceinline__ float4 zero4() { return make_float4(0.f, 0.f, 0.f, 0.f); }

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 ldcg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

// 16-byte cp.async (LDGSTS); src_bytes = 0 zero-fills the destination.
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src, int src_bytes) {
  uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(gmem_src), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// ---------------------------------------------------------------------------------------------
// 128x128 CTA-tile FFMA micro-kernels, 256 threads, 8x8 accumulators per thread.
// Thread (tm = tid>>4, tn = tid&15) owns rows {tm + 16 i} (i<8) and columns {4 tn + 64 j + c}.
// ---------------------------------------------------------------------------------------------

// A row-major in smem (As[row * lda + k]), W reduction-major (Ws[k * 128 + n]); KC = k extent.
template <int KC>
__device__ __forceinline__ void mma_rowA(const float* __restrict__ As, int lda, const float* __restrict__ Ws,
                                         float (&acc)[8][8], int tm, int tn) {
#pragma unroll 2
  for (int k0 = 0; k0 < KC; k0 += 4) {
    float4 a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = *reinterpret_cast<const float4*>(As + (tm + 16 * i) * lda + k0);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      float4 w0 = *reinterpret_cast<const float4*>(Ws + (k0 + kk) * 128 + tn * 4);
      float4 w1 = *reinterpret_cast<const float4*>(Ws + (k0 + kk) * 128 + 64 + tn * 4);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float av = kk == 0 ? a[i].x : kk == 1 ? a[i].y : kk == 2 ? a[i].z : a[i].w;
        acc[i][0] = fmaf(av, w0.x, acc[i][0]);
        acc[i][1] = fmaf(av, w0.y, acc[i][1]);
        acc[i][2] = fmaf(av, w0.z, acc[i][2]);
        acc[i][3] = fmaf(av, w0.w, acc[i][3]);
        acc[i][4] = fmaf(av, w1.x, acc[i][4]);
        acc[i][5] = fmaf(av, w1.y, acc[i][5]);
        acc[i][6] = fmaf(av, w1.z, acc[i][6]);
        acc[i][7] = fmaf(av, w1.w, acc[i][7]);
      }
    }
  }
}

// Both operands reduction-major: out[p][q] += sum_m Ps[m * ldp + p] * Qs[m * ldq + q].
// Thread (tp = tid>>4, tq = tid&15) owns p in {4 tp + c, 64 + 4 tp + c}, q in {4 tq + c, 64 + 4 tq + c}.
template <int MC>
__device__ __forceinline__ void mma_redmajor(const float* __restrict__ Ps, int ldp, const float* __restrict__ Qs,
                                             int ldq, float (&acc)[8][8], int tp, int tq) {
#pragma unroll 4
  for (int m = 0; m < MC; ++m) {
    float4 p0 = *reinterpret_cast<const float4*>(Ps + m * ldp + tp * 4);
    float4 p1 = *reinterpret_cast<const float4*>(Ps + m * ldp + 64 + tp * 4);
    float4 q0 = *reinterpret_cast<const float4*>(Qs + m * ldq + tq * 4);
    float4 q1 = *reinterpret_cast<const float4*>(Qs + m * ldq + 64 + tq * 4);
    float pv[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
    float qv[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(pv[i], qv[j], acc[i][j]);
  }
}

// out[i] (+)= sum_{s < S} part[s * stride + i]   (fixed order => deterministic second-stage reduction)
static __global__ void k_reduce_partials(const float* __restrict__ part, float* __restrict__ out, int count, int S,
                                         size_t stride, int accumulate) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  float s = 0.f;
  for (int q = 0; q < S; ++q) s += part[(size_t)q * stride + i];
  out[i] = accumulate ? out[i] + s : s;
}

// The same for few outputs and MANY partials (decoder weight gradients: 490 outputs, one partial per 8 nodes): a CTA
// of 8 warps owns 32 consecutive outputs; warp w adds the partials w, w + 8, ... in order (coalesced 128-byte rows, four
// loads in flight), the eight warp sums are then added in warp order.  Fixed order => deterministic.
static __global__ void __launch_bounds__(256) k_reduce_partials_tall(const float* __restrict__ part,
                                                                     float* __restrict__ out, int count, int S,
                                                                     size_t stride) {
  __shared__ float sh[8][32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + lane;
  float s = 0.f;
  if (i < count) {
    int q = w;
    for (; q + 24 < S; q += 32) {
      const float a = part[(size_t)q * stride + i], b = part[(size_t)(q + 8) * stride + i];
      const float c = part[(size_t)(q + 16) * stride + i], d = part[(size_t)(q + 24) * stride + i];
      s += a;
      s += b;
      s += c;
      s += d;
    }
    for (; q < S; q += 8) s += part[(size_t)q * stride + i];
  }
  sh[w][lane] = s;
  __syncthreads();
  if (w == 0 && i < count) {
    float t = sh[0][lane];
#pragma unroll
    for (int k = 1; k < 8; ++k) t += sh[k][lane];
    out[i] = t;
  }
}

// Two partial sets in one launch (main weight tile + side/bias rows of a weight-gradient call).
static __global__ void k_reduce_partials2(const float* __restrict__ part0, float* __restrict__ out0, int count0,
                                          size_t stride0, const float* __restrict__ part1, float* __restrict__ out1,
                                          int count1, size_t stride1, int S, int accumulate) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  const float* part;
  float* out;
  size_t stride;
  if (i < count0) {
    part = part0; out = out0; stride = stride0;
  } else {
    i -= count0;
    if (i >= count1) return;
    part = part1; out = out1; stride = stride1;
  }
  float s = 0.f;
  for (int q = 0; q < S; ++q) s += part[(size_t)q * stride + i];
  out[i] = accumulate ? out[i] + s : s;
}

}  // namespace msmp
