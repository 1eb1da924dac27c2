// LEM recurrence gate math (replaces the point-wise kernels of the absent `lem_cuda` extension;
// call sites experiments/models_gnn.py:290-292,300).  The two GEMMs of every step run through
// msmp_linear_fwd; these kernels fuse everything between them.  gates[4][N][128] per step holds
// a = dt*sigmoid(G0), b = dt*sigmoid(G1), zc = tanh(G2), tL = tanh(L) for the backward pass.
#include "common.cuh"
#include "msmp_b200.h"

namespace msmp {

__global__ void __launch_bounds__(256) k_lem_gate_z(const float* __restrict__ G, const float* __restrict__ z_prev,
                                                    float dt, float* __restrict__ gates, float* __restrict__ z_new,
                                                    int N) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N * 128) return;
  const int row = idx >> 7, c = idx & 127;
  const float* g = G + (size_t)row * 384;
  const float a = dt * sigmoidf_(g[c]);
  const float b = dt * sigmoidf_(g[128 + c]);
  const float zc = tanhf(g[256 + c]);
  const size_t plane = (size_t)N * 128;
  gates[idx] = a;
  gates[plane + idx] = b;
  gates[2 * plane + idx] = zc;
  z_new[idx] = (1.f - b) * z_prev[idx] + b * zc;
}

__global__ void __launch_bounds__(256) k_lem_gate_y(const float* __restrict__ L, const float* __restrict__ y_prev,
                                                    float* __restrict__ gates, float* __restrict__ y_new, int N) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N * 128) return;
  const size_t plane = (size_t)N * 128;
  const float a = gates[idx];
  const float tl = tanhf(L[idx]);
  gates[3 * plane + idx] = tl;
  y_new[idx] = (1.f - a) * y_prev[idx] + a * tl;
}

// d = dy + gy;  dL = d a (1 - tL^2);  dG0 = d (tL - y_prev) a (1 - a/dt);  dy <- d (1 - a)
__global__ void __launch_bounds__(256) k_lem_bwd_y(float* __restrict__ dy, const float* __restrict__ gy,
                                                   const float* __restrict__ y_prev, const float* __restrict__ gates,
                                                   float dt, float* __restrict__ dL, float* __restrict__ dG, int N) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N * 128) return;
  const int row = idx >> 7, c = idx & 127;
  const size_t plane = (size_t)N * 128;
  const float a = gates[idx], tl = gates[3 * plane + idx];
  const float d = dy[idx] + (gy ? gy[idx] : 0.f);
  dL[idx] = d * a * (1.f - tl * tl);
  dG[(size_t)row * 384 + c] = d * (tl - y_prev[idx]) * a * (1.f - a / dt);
  dy[idx] = d * (1.f - a);
}

// d = dz_tot + gz;  dG1 = d (zc - z_prev) b (1 - b/dt);  dG2 = d b (1 - zc^2);  dz <- d (1 - b)
__global__ void __launch_bounds__(256) k_lem_bwd_z(const float* __restrict__ dz_tot, const float* __restrict__ gz,
                                                   const float* __restrict__ z_prev, const float* __restrict__ gates,
                                                   float dt, float* __restrict__ dG, float* __restrict__ dz, int N) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N * 128) return;
  const int row = idx >> 7, c = idx & 127;
  const size_t plane = (size_t)N * 128;
  const float b = gates[plane + idx], zc = gates[2 * plane + idx];
  const float d = dz_tot[idx] + (gz ? gz[idx] : 0.f);
  dG[(size_t)row * 384 + 128 + c] = d * (zc - z_prev[idx]) * b * (1.f - b / dt);
  dG[(size_t)row * 384 + 256 + c] = d * b * (1.f - zc * zc);
  dz[idx] = d * (1.f - b);
}

}  // namespace msmp

using namespace msmp;

extern "C" int msmp_lem_gate_z(const float* G, const float* z_prev, float dt, float* gates, float* z_new, int N,
                               cudaStream_t stream) {
  if (N <= 0) return N == 0 ? MSMP_OK : MSMP_ERR_ARG;
  k_lem_gate_z<<<(N * 128 + 255) / 256, 256, 0, stream>>>(G, z_prev, dt, gates, z_new, N);
  MSMP_CHECK_LAUNCH();
  return MSMP_OK;
}
extern "C" int msmp_lem_gate_y(const float* L, const float* y_prev, float* gates, float* y_new, int N,
                               cudaStream_t stream) {
  if (N <= 0) return N == 0 ? MSMP_OK : MSMP_ERR_ARG;
  k_lem_gate_y<<<(N * 128 + 255) / 256, 256, 0, stream>>>(L, y_prev, gates, y_new, N);
  MSMP_CHECK_LAUNCH();
  return MSMP_OK;
}
extern "C" int msmp_lem_bwd_y(float* dy, const float* gy, const float* y_prev, const float* gates, float dt, float* dL,
                              float* dG, int N, cudaStream_t stream) {
  if (N <= 0) return N == 0 ? MSMP_OK : MSMP_ERR_ARG;
  k_lem_bwd_y<<<(N * 128 + 255) / 256, 256, 0, stream>>>(dy, gy, y_prev, gates, dt, dL, dG, N);
  MSMP_CHECK_LAUNCH();
  return MSMP_OK;
}
extern "C" int msmp_lem_bwd_z(const float* dz_tot, const float* gz, const float* z_prev, const float* gates, float dt,
                              float* dG, float* dz, int N, cudaStream_t stream) {
  if (N <= 0) return N == 0 ? MSMP_OK : MSMP_ERR_ARG;
  k_lem_bwd_z<<<(N * 128 + 255) / 256, 256, 0, stream>>>(dz_tot, gz, z_prev, gates, dt, dG, dz, N);
  MSMP_CHECK_LAUNCH();
  return MSMP_OK;
}
