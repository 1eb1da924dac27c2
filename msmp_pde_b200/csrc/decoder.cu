// Output decoder of the MP-PDE / MSMP-PDE solvers (sm_100a, fp32):
//   diff = Conv1d(C, 8, K1, stride S1) -> Swish -> Conv1d(8, C, K2) on h[N, C, 128]
//   out[n, c, k] = base[n, c, k] + dt[k] * diff[n, c, k]
// (models_gnn.py:208-224,275-279: C = 1, base = u[n, tw-1];  models_gnn2D.py:382-391,448-458: C = 2,
//  base = u[n, c*tw + k]).  Geometry for tw = 25: K1 = 16, S1 = 3, L1 = 38, K2 = 14.
// One CTA handles a tile of nodes entirely in shared memory; the first pre-activation is kept for the
// backward pass.  Weight gradients: per-CTA partials + fixed-order reduction (no atomics).
#include "common.cuh"
#include "msmp_b200.h"

namespace msmp {

constexpr int DEC_LIN = 128;
constexpr int DEC_OC = 8;
constexpr int DEC_MAXL1 = 64;       // >= L1 for every supported geometry (38, 29, 59)
constexpr int DEC_MAXW = 8 * 2 * 16 + 8 + 2 * 8 * 16 + 2;
constexpr int DEC_NB_F = 8;         // nodes per CTA, forward
constexpr int DEC_NB_B = 8;         // nodes per CTA, backward

struct DecGeom {
  int C, K1, S1, L1, K2, TW;
};

// Geometry either from the launch parameters (FC = 0) or fixed at compile time: with the tw = 25 geometry of the
// reference's default runs (K1 = 16, S1 = 3, L1 = 38, K2 = 14) as constants, the inner loops unroll and the index
// divisions fold; the kernels are instruction bound (two shared-memory loads per FMA), not bandwidth bound.
template <int FC, int FK1, int FS1, int FL1, int FK2, int FTW>
struct GeomT {
  int C, K1, S1, L1, K2, TW;
  __device__ __forceinline__ explicit GeomT(const DecGeom& r)
      : C(FC ? FC : r.C), K1(FC ? FK1 : r.K1), S1(FC ? FS1 : r.S1), L1(FC ? FL1 : r.L1), K2(FC ? FK2 : r.K2),
        TW(FC ? FTW : r.TW) {}
};

struct DecFwdParams {
  const float* h;       // [N][C*128]
  const float* w1; const float* b1; const float* w2; const float* b2;
  const float* u; int ldu;      // [N][ldu] fp32 node inputs
  const float* dt;      // [TW]
  float* za;            // [N][8*L1]
  float* out;           // [N][C*TW]
  int N;
  DecGeom g;
};

template <int FC, int FK1, int FS1, int FL1, int FK2, int FTW>
__global__ void __launch_bounds__(256) k_decoder_fwd(const DecFwdParams p) {
  __shared__ __align__(16) float sh[DEC_NB_F * 2 * DEC_LIN];
  __shared__ float sa[DEC_NB_F * DEC_OC * DEC_MAXL1];
  __shared__ float sw1[8 * 2 * 16], sw2[2 * 8 * 16], sb1[8], sb2[2], sdt[64];
  const GeomT<FC, FK1, FS1, FL1, FK2, FTW> g(p.g);
  const int tid = threadIdx.x;
  const int n0 = blockIdx.x * DEC_NB_F;
  const int nb = min(DEC_NB_F, p.N - n0);
  const int CL = g.C * DEC_LIN;
  for (int i = tid; i < DEC_OC * g.C * g.K1; i += 256) sw1[i] = p.w1[i];
  for (int i = tid; i < g.C * DEC_OC * g.K2; i += 256) sw2[i] = p.w2[i];
  if (tid < DEC_OC) sb1[tid] = p.b1[tid];
  if (tid < g.C) sb2[tid] = p.b2[tid];
  if (tid < g.TW) sdt[tid] = p.dt[tid];
  for (int i = tid; i < nb * CL / 4; i += 256)
    st4(sh + 4 * i, ldg4(p.h + (size_t)n0 * CL + 4 * i));
  __syncthreads();
  const int per1 = DEC_OC * g.L1;
  for (int i = tid; i < nb * per1; i += 256) {
    const int n = i / per1, r = i - n * per1;
    const int o = r / g.L1, q = r - o * g.L1;
    float z = sb1[o];
    for (int c = 0; c < g.C; ++c) {
      const float* hp = sh + n * CL + c * DEC_LIN + g.S1 * q;
      const float* wp = sw1 + (o * g.C + c) * g.K1;
      for (int j = 0; j < g.K1; ++j) z = fmaf(wp[j], hp[j], z);
    }
    p.za[(size_t)(n0 + n) * per1 + r] = z;
    sa[n * DEC_OC * DEC_MAXL1 + o * DEC_MAXL1 + q] = swish(z);
  }
  __syncthreads();
  const int per2 = g.C * g.TW;
  for (int i = tid; i < nb * per2; i += 256) {
    const int n = i / per2, r = i - n * per2;
    const int c = r / g.TW, k = r - c * g.TW;
    float d = sb2[c];
    for (int o = 0; o < DEC_OC; ++o) {
      const float* ap = sa + n * DEC_OC * DEC_MAXL1 + o * DEC_MAXL1 + k;
      const float* wp = sw2 + (c * DEC_OC + o) * g.K2;
      for (int j = 0; j < g.K2; ++j) d = fmaf(wp[j], ap[j], d);
    }
    const float* un = p.u + (size_t)(n0 + n) * p.ldu;
    const float base = (g.C == 1) ? un[g.TW - 1] : un[c * g.TW + k];
    p.out[(size_t)(n0 + n) * per2 + r] = base + sdt[k] * d;
  }
}

struct DecBwdParams {
  const float* dout;    // [N][C*TW]
  const float* h;       // [N][C*128]
  const float* za;      // [N][8*L1]
  const float* w1; const float* w2;
  const float* dt;
  float* dh;            // [N][C*128]
  float* part;          // [numCTA][nW]  (w1 | b1 | w2 | b2)
  int N;
  DecGeom g;
};

template <int FC, int FK1, int FS1, int FL1, int FK2, int FTW>
__global__ void __launch_bounds__(256) k_decoder_bwd(const DecBwdParams p) {
  __shared__ __align__(16) float sh[DEC_NB_B * 2 * DEC_LIN];
  __shared__ float sa[DEC_NB_B * DEC_OC * DEC_MAXL1];     // a = swish(za)
  __shared__ float sz[DEC_NB_B * DEC_OC * DEC_MAXL1];     // za, then dza
  __shared__ float sdd[DEC_NB_B * 2 * 64];                // dout * dt
  __shared__ float sw1[8 * 2 * 16], sw2[2 * 8 * 16];
  const GeomT<FC, FK1, FS1, FL1, FK2, FTW> g(p.g);
  const int tid = threadIdx.x;
  const int n0 = blockIdx.x * DEC_NB_B;
  const int nb = min(DEC_NB_B, p.N - n0);
  const int CL = g.C * DEC_LIN;
  const int per1 = DEC_OC * g.L1, per2 = g.C * g.TW;
  for (int i = tid; i < DEC_OC * g.C * g.K1; i += 256) sw1[i] = p.w1[i];
  for (int i = tid; i < g.C * DEC_OC * g.K2; i += 256) sw2[i] = p.w2[i];
  for (int i = tid; i < nb * CL / 4; i += 256) st4(sh + 4 * i, ldg4(p.h + (size_t)n0 * CL + 4 * i));
  for (int i = tid; i < nb * per2; i += 256) {
    const int n = i / per2, r = i - n * per2;
    const int c = r / g.TW, k = r - c * g.TW;
    sdd[(n * 2 + c) * 64 + k] = p.dout[(size_t)(n0 + n) * per2 + r] * p.dt[k];
  }
  for (int i = tid; i < nb * per1; i += 256) {
    const int n = i / per1, r = i - n * per1;
    const int o = r / g.L1, q = r - o * g.L1;
    const float z = p.za[(size_t)(n0 + n) * per1 + r];
    sz[n * DEC_OC * DEC_MAXL1 + o * DEC_MAXL1 + q] = z;
    sa[n * DEC_OC * DEC_MAXL1 + o * DEC_MAXL1 + q] = swish(z);
  }
  __syncthreads();
  // da[n][o][l] = sum_c sum_j dd[n][c][l-j] w2[c][o][j];  dza = da * swish'(za)
  for (int i = tid; i < nb * per1; i += 256) {
    const int n = i / per1, r = i - n * per1;
    const int o = r / g.L1, l = r - o * g.L1;
    float s = 0.f;
    const int j0 = max(0, l - g.TW + 1), j1 = min(g.K2 - 1, l);
    for (int c = 0; c < g.C; ++c) {
      const float* dp = sdd + (n * 2 + c) * 64;
      const float* wp = sw2 + (c * DEC_OC + o) * g.K2;
      for (int j = j0; j <= j1; ++j) s = fmaf(dp[l - j], wp[j], s);
    }
    float* zp = sz + n * DEC_OC * DEC_MAXL1 + o * DEC_MAXL1 + l;
    *zp = s * dswish(*zp);       // each element is read and rewritten by exactly one thread
  }
  __syncthreads();
  // dh[n][c][pp] = sum_o sum_{q : 0 <= pp - S1 q < K1} dza[n][o][q] w1[o][c][pp - S1 q]
  for (int i = tid; i < nb * CL; i += 256) {
    const int n = i / CL, r = i - n * CL;
    const int c = r / DEC_LIN, pp = r - c * DEC_LIN;
    int q_lo = (pp - g.K1 + 1 + g.S1 - 1);
    q_lo = q_lo > 0 ? q_lo / g.S1 : 0;
    const int q_hi = min(g.L1 - 1, pp / g.S1);
    float s = 0.f;
    for (int o = 0; o < DEC_OC; ++o) {
      const float* zp = sz + n * DEC_OC * DEC_MAXL1 + o * DEC_MAXL1;
      const float* wp = sw1 + (o * g.C + c) * g.K1;
      for (int q = q_lo; q <= q_hi; ++q) s = fmaf(zp[q], wp[pp - g.S1 * q], s);
    }
    p.dh[(size_t)(n0 + n) * CL + r] = s;
  }
  // weight-gradient partials of this CTA (one weight element per thread iteration; fixed order)
  const int n_w1 = DEC_OC * g.C * g.K1, n_w2 = g.C * DEC_OC * g.K2;
  const int nW = n_w1 + DEC_OC + n_w2 + g.C;
  float* part = p.part + (size_t)blockIdx.x * nW;
  for (int e = tid; e < nW; e += 256) {
    float s = 0.f;
    if (e < n_w1) {
      const int o = e / (g.C * g.K1), r = e - o * g.C * g.K1;
      const int c = r / g.K1, j = r - c * g.K1;
      for (int n = 0; n < nb; ++n) {
        const float* zp = sz + n * DEC_OC * DEC_MAXL1 + o * DEC_MAXL1;
        const float* hp = sh + n * CL + c * DEC_LIN + j;
        for (int q = 0; q < g.L1; ++q) s = fmaf(zp[q], hp[g.S1 * q], s);
      }
    } else if (e < n_w1 + DEC_OC) {
      const int o = e - n_w1;
      for (int n = 0; n < nb; ++n) {
        const float* zp = sz + n * DEC_OC * DEC_MAXL1 + o * DEC_MAXL1;
        for (int q = 0; q < g.L1; ++q) s += zp[q];
      }
    } else if (e < n_w1 + DEC_OC + n_w2) {
      const int r0 = e - n_w1 - DEC_OC;
      const int c = r0 / (DEC_OC * g.K2), r = r0 - c * DEC_OC * g.K2;
      const int o = r / g.K2, j = r - o * g.K2;
      for (int n = 0; n < nb; ++n) {
        const float* dp = sdd + (n * 2 + c) * 64;
        const float* ap = sa + n * DEC_OC * DEC_MAXL1 + o * DEC_MAXL1 + j;
        for (int k = 0; k < g.TW; ++k) s = fmaf(dp[k], ap[k], s);
      }
    } else {
      const int c = e - n_w1 - DEC_OC - n_w2;
      for (int n = 0; n < nb; ++n) {
        const float* dp = sdd + (n * 2 + c) * 64;
        for (int k = 0; k < g.TW; ++k) s += dp[k];
      }
    }
    part[e] = s;
  }
}

// register-tiled kernels of the tw = 25 geometry (decoder_rt.cu)
int launch_decoder_fwd_rt(const float* h, const float* w1, const float* b1, const float* w2, const float* b2, const float* u,
                          int ldu, const float* dt, float* za, float* out, int N, int C, cudaStream_t stream);
int launch_decoder_bwd_rt(const float* dout, const float* h, const float* za, const float* w1, const float* w2,
                          const float* dt, float* dh, float* part, int N, int C, int* nparts, cudaStream_t stream);

static bool geom_ok(const DecGeom& g) {
  if (g.C < 1 || g.C > 2 || g.K1 < 1 || g.K1 > 16 || g.K2 < 1 || g.K2 > 16 || g.S1 < 1) return false;
  if (g.L1 != (DEC_LIN - g.K1) / g.S1 + 1 || g.L1 > DEC_MAXL1) return false;
  if (g.TW != g.L1 - g.K2 + 1 || g.TW > 64) return false;
  return true;
}

}  // namespace msmp

using namespace msmp;

extern "C" int msmp_decoder_nweights(int C, int K1, int K2) { return 8 * C * K1 + 8 + C * 8 * K2 + C; }

extern "C" size_t msmp_decoder_bwd_workspace(int N, int C, int K1, int K2) {
  size_t ctas = (size_t)(N + DEC_NB_B - 1) / DEC_NB_B;
  return ctas * (size_t)msmp_decoder_nweights(C, K1, K2) * sizeof(float);
}

extern "C" int msmp_decoder_fwd(const float* h, const float* w1, const float* b1, const float* w2, const float* b2,
                                const float* u, int ldu, const float* dt, float* za, float* out, int N, int C, int K1,
                                int S1, int L1, int K2, int TW, cudaStream_t stream) {
  DecGeom g{C, K1, S1, L1, K2, TW};
  if (!geom_ok(g) || N < 0) return MSMP_ERR_ARG;
  if (N == 0) return MSMP_OK;
  if (K1 == 16 && S1 == 3 && L1 == 38 && K2 == 14 && TW == 25) {
    const int rc = launch_decoder_fwd_rt(h, w1, b1, w2, b2, u, ldu, dt, za, out, N, C, stream);
    if (rc <= 0) return rc;
  }
  DecFwdParams p{h, w1, b1, w2, b2, u, ldu, dt, za, out, N, g};
  const int grid = (N + DEC_NB_F - 1) / DEC_NB_F;
  if (K1 == 16 && S1 == 3 && L1 == 38 && K2 == 14 && TW == 25 && C == 1)
    k_decoder_fwd<1, 16, 3, 38, 14, 25><<<grid, 256, 0, stream>>>(p);
  else if (K1 == 16 && S1 == 3 && L1 == 38 && K2 == 14 && TW == 25 && C == 2)
    k_decoder_fwd<2, 16, 3, 38, 14, 25><<<grid, 256, 0, stream>>>(p);
  else
    k_decoder_fwd<0, 0, 0, 0, 0, 0><<<grid, 256, 0, stream>>>(p);
  MSMP_CHECK_LAUNCH();
  return MSMP_OK;
}

// dW = [w1 (8*C*K1) | b1 (8) | w2 (C*8*K2) | b2 (C)] in the parameters' own layouts
extern "C" int msmp_decoder_bwd(const float* dout, const float* h, const float* za, const float* w1, const float* w2,
                                const float* dt, float* dh, float* dW, int N, int C, int K1, int S1, int L1, int K2,
                                int TW, void* workspace, size_t ws_bytes, cudaStream_t stream) {
  DecGeom g{C, K1, S1, L1, K2, TW};
  if (!geom_ok(g) || N < 0) return MSMP_ERR_ARG;
  const int nW = msmp_decoder_nweights(C, K1, K2);
  if (N == 0) {
    if (cudaMemsetAsync(dW, 0, nW * sizeof(float), stream) != cudaSuccess) return MSMP_ERR_CUDA;
    return MSMP_OK;
  }
  if (ws_bytes < msmp_decoder_bwd_workspace(N, C, K1, K2)) return MSMP_ERR_WORKSPACE;
  if (K1 == 16 && S1 == 3 && L1 == 38 && K2 == 14 && TW == 25) {
    int nparts = 0;
    const int rc = launch_decoder_bwd_rt(dout, h, za, w1, w2, dt, dh, reinterpret_cast<float*>(workspace), N, C, &nparts, stream);
    if (rc < 0) return rc;
    if (rc == 0) {
      k_reduce_partials_tall<<<(nW + 31) / 32, 256, 0, stream>>>(reinterpret_cast<float*>(workspace), dW, nW, nparts, (size_t)nW);
      MSMP_CHECK_LAUNCH();
      return MSMP_OK;
    }
  }
  const int ctas = (N + DEC_NB_B - 1) / DEC_NB_B;
  DecBwdParams p{dout, h, za, w1, w2, dt, dh, reinterpret_cast<float*>(workspace), N, g};
  if (K1 == 16 && S1 == 3 && L1 == 38 && K2 == 14 && TW == 25 && C == 1)
    k_decoder_bwd<1, 16, 3, 38, 14, 25><<<ctas, 256, 0, stream>>>(p);
  else if (K1 == 16 && S1 == 3 && L1 == 38 && K2 == 14 && TW == 25 && C == 2)
    k_decoder_bwd<2, 16, 3, 38, 14, 25><<<ctas, 256, 0, stream>>>(p);
  else
    k_decoder_bwd<0, 0, 0, 0, 0, 0><<<ctas, 256, 0, stream>>>(p);
  MSMP_CHECK_LAUNCH();
  k_reduce_partials_tall<<<(nW + 31) / 32, 256, 0, stream>>>(p.part, dW, nW, ctas, (size_t)nW);
  MSMP_CHECK_LAUNCH();
  return MSMP_OK;
}
