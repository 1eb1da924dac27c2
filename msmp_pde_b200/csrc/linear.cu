// Node-level dense layers of the MP-PDE / MSMP-PDE stack, fp32 FFMA path (sm_100a).
//
//   msmp_linear_fwd   : Y = epilogue( [A0 | A1 | A2] * Wt + bias + side * Wside )
//                       covers message_net_1 in its per-node factorised form (P | Q projection),
//                       update_net_1/2 (models_gnn.py:47-58,77-86), the encoder MLPs (:201-206),
//                       the LEM gate GEMMs (:290) and every dgrad of those.
//   msmp_linear_wgrad : dWt[k][n] = sum_m X[m][k] * dY[m][n]  (+ side / bias gradients), split over
//                       row ranges into per-CTA partials, then msmp_reduce_partials sums the
//                       partials in a fixed order => run-to-run bit-identical weight gradients
//                       (no float atomics anywhere).
#include <cstdlib>
#include "common.cuh"
#include "msmp_b200.h"

namespace msmp {

constexpr int LIN_KC = 32;                 // k-chunk
constexpr int LIN_LDA = LIN_KC + 4;        // padded A row in smem (conflict-free float4 rows)
constexpr int LIN_SMEM = (2 * 128 * LIN_LDA + 2 * LIN_KC * 128) * 4;

struct LinParams {
  const float* A[3];
  int lda[3];
  int ka[3];
  int aswish[3];
  int nseg;
  const float* Wt;
  int ldw;
  const float* bias;
  const float* side;
  int lds;
  int r;
  const float* Wside;
  const float* Zmul;
  int ldz;
  float* Ypre;
  int ldpre;
  int act;
  const float* R;
  int ldr;
  float* Y;
  int ldy;
  int M;
  int Nout;
};

__global__ void __launch_bounds__(256, 2) k_linear(const LinParams p) {
  extern __shared__ __align__(16) float smem[];
  float* As = smem;                          // [2][128][LIN_LDA]
  float* Ws = smem + 2 * 128 * LIN_LDA;      // [2][LIN_KC][128]
  const int tid = threadIdx.x;
  const int tm = tid >> 4, tn = tid & 15;
  const int row0 = blockIdx.x * 128;
  const int n0 = blockIdx.y * 128;

  int ktot = 0;
  for (int s = 0; s < p.nseg; ++s) ktot += p.ka[s];
  const int nchunks = ktot / LIN_KC;

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  // chunk c -> (segment, offset inside the segment)
  auto locate = [&](int c, int& seg, int& koff) {
    int k = c * LIN_KC;
    seg = 0;
    while (seg < p.nseg - 1 && k >= p.ka[seg]) {
      k -= p.ka[seg];
      ++seg;
    }
    koff = k;
  };
  auto issue = [&](int c, int buf) {
    int seg, koff;
    locate(c, seg, koff);
    const float* A = p.A[seg];
    const int lda = p.lda[seg];
    float* as = As + buf * 128 * LIN_LDA;
    float* ws = Ws + buf * LIN_KC * 128;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int idx = tid + 256 * i;
      int r = idx >> 3, c4 = idx & 7;
      int grow = row0 + r;
      const float* src = A + (size_t)(grow < p.M ? grow : 0) * lda + koff + 4 * c4;
      cp_async16(as + r * LIN_LDA + 4 * c4, src, grow < p.M ? 16 : 0);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int idx = tid + 256 * i;
      int kr = idx >> 5, c4 = idx & 31;
      const float* src = p.Wt + (size_t)(c * LIN_KC + kr) * p.ldw + n0 + 4 * c4;
      cp_async16(ws + kr * 128 + 4 * c4, src, 16);
    }
    cp_async_commit();
  };

  if (nchunks > 0) issue(0, 0);
  for (int c = 0; c < nchunks; ++c) {
    const int buf = c & 1;
    if (c + 1 < nchunks) {
      issue(c + 1, buf ^ 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    int seg, koff;
    locate(c, seg, koff);
    if (p.aswish[seg]) {      // transform the elements this thread copied itself (visible after wait)
      float* as = As + buf * 128 * LIN_LDA;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        int idx = tid + 256 * i;
        int r = idx >> 3, c4 = idx & 7;
        float4* q = reinterpret_cast<float4*>(as + r * LIN_LDA + 4 * c4);
        *q = swish4(*q);
      }
    }
    __syncthreads();
    mma_rowA<LIN_KC>(As + buf * 128 * LIN_LDA, LIN_LDA, Ws + buf * LIN_KC * 128, acc, tm, tn);
    __syncthreads();
  }

  // ---- epilogue
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = row0 + tm + 16 * i;
    if (row >= p.M) continue;
    float sv[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) sv[q] = (q < p.r) ? __ldg(p.side + (size_t)row * p.lds + q) : 0.f;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int col = n0 + tn * 4 + 64 * j;
      if (col >= p.Nout) continue;
      float4 z = make_float4(acc[i][4 * j], acc[i][4 * j + 1], acc[i][4 * j + 2], acc[i][4 * j + 3]);
      if (p.bias) z = add4(z, ldg4(p.bias + col));
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        if (q >= p.r) break;
        float4 w = ldg4(p.Wside + (size_t)q * p.ldw + col);
        z.x = fmaf(sv[q], w.x, z.x);
        z.y = fmaf(sv[q], w.y, z.y);
        z.z = fmaf(sv[q], w.z, z.z);
        z.w = fmaf(sv[q], w.w, z.w);
      }
      if (p.Zmul) {
        float4 zz = ldg4(p.Zmul + (size_t)row * p.ldz + col);
        z = mul4(z, make_float4(dswish(zz.x), dswish(zz.y), dswish(zz.z), dswish(zz.w)));
      }
      if (p.Ypre) st4(p.Ypre + (size_t)row * p.ldpre + col, z);
      if (p.act) z = swish4(z);
      if (p.R) z = add4(z, ldg4(p.R + (size_t)row * p.ldr + col));
      st4(p.Y + (size_t)row * p.ldy + col, z);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// wgrad: part[s][k][n] = sum_{m in split s} X[m][k] * dY[m][n]
// ---------------------------------------------------------------------------------------------
constexpr int WG_MC = 32;
constexpr int WG_SMEM = (2 * WG_MC * 128 * 2 + 2 * WG_MC * 16) * 4;

struct WgradParams {
  const float* X;
  int ldx;
  int K;
  int xswish;
  const float* dY;
  int lddy;
  int Nout;
  const float* side;      // [M][lds], r columns (nullable)
  int lds;
  int r;
  int has_bias;           // extra implicit all-ones side column (bias gradient)
  float* part;            // [S][K][Nout]
  float* part_side;       // [S][r + has_bias][Nout]
  int M;
  int rows_per_split;     // multiple of WG_MC
};

__global__ void __launch_bounds__(256, 2) k_wgrad(const WgradParams p) {
  extern __shared__ __align__(16) float smem[];
  float* Xs = smem;                              // [2][WG_MC][128]
  float* Ys = smem + 2 * WG_MC * 128;            // [2][WG_MC][128]
  float* Ss = smem + 4 * WG_MC * 128;            // [2][WG_MC][16]
  const int tid = threadIdx.x;
  const int tk = tid >> 4, tn = tid & 15;
  const int split = blockIdx.x;
  const int k0 = blockIdx.y * 128;
  const int n0 = blockIdx.z * 128;
  const int m_begin = split * p.rows_per_split;
  const int m_end = min(p.M, m_begin + p.rows_per_split);
  const int nside = p.r + p.has_bias;
  const bool do_side = (blockIdx.y == 0) && nside > 0;

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  float sacc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) sacc[j] = 0.f;

  const int nchunks = (m_end > m_begin) ? (m_end - m_begin + WG_MC - 1) / WG_MC : 0;

  auto issue = [&](int c, int buf) {
    const int mb = m_begin + c * WG_MC;
    float* xs = Xs + buf * WG_MC * 128;
    float* ys = Ys + buf * WG_MC * 128;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int idx = tid + 256 * i;
      int mm = idx >> 5, c4 = idx & 31;
      int m = mb + mm;
      bool okm = m < m_end;
      bool okx = okm && (k0 + 4 * c4 < p.K);
      bool oky = okm && (n0 + 4 * c4 < p.Nout);
      cp_async16(xs + mm * 128 + 4 * c4, p.X + (okx ? (size_t)m * p.ldx + k0 + 4 * c4 : 0), okx ? 16 : 0);
      cp_async16(ys + mm * 128 + 4 * c4, p.dY + (oky ? (size_t)m * p.lddy + n0 + 4 * c4 : 0), oky ? 16 : 0);
    }
    cp_async_commit();
    if (do_side) {
      // 32 rows x 16 side slots, 2 per thread
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        int idx = tid + 256 * i;
        int mm = idx >> 4, q = idx & 15;
        int m = mb + mm;
        float v = 0.f;
        if (m < m_end) {
          if (q < p.r) v = __ldg(p.side + (size_t)m * p.lds + q);
          else if (q == p.r && p.has_bias) v = 1.0f;
        }
        Ss[buf * WG_MC * 16 + mm * 16 + q] = v;
      }
    }
  };

  if (nchunks > 0) issue(0, 0);
  for (int c = 0; c < nchunks; ++c) {
    const int buf = c & 1;
    if (c + 1 < nchunks) {
      issue(c + 1, buf ^ 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    if (p.xswish) {
      float* xs = Xs + buf * WG_MC * 128;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        int idx = tid + 256 * i;
        float4* q = reinterpret_cast<float4*>(xs + (idx >> 5) * 128 + 4 * (idx & 31));
        *q = swish4(*q);      // zero-filled tail rows stay zero (swish(0) = 0)
      }
    }
    __syncthreads();
    const float* xs = Xs + buf * WG_MC * 128;
    const float* ys = Ys + buf * WG_MC * 128;
    mma_redmajor<WG_MC>(xs, 128, ys, 128, acc, tk, tn);
    if (do_side && tk < nside) {
      const float* ss = Ss + buf * WG_MC * 16;
#pragma unroll 4
      for (int mm = 0; mm < WG_MC; ++mm) {
        float s = ss[mm * 16 + tk];
        float4 y0 = *reinterpret_cast<const float4*>(ys + mm * 128 + tn * 4);
        float4 y1 = *reinterpret_cast<const float4*>(ys + mm * 128 + 64 + tn * 4);
        sacc[0] = fmaf(s, y0.x, sacc[0]);
        sacc[1] = fmaf(s, y0.y, sacc[1]);
        sacc[2] = fmaf(s, y0.z, sacc[2]);
        sacc[3] = fmaf(s, y0.w, sacc[3]);
        sacc[4] = fmaf(s, y1.x, sacc[4]);
        sacc[5] = fmaf(s, y1.y, sacc[5]);
        sacc[6] = fmaf(s, y1.z, sacc[6]);
        sacc[7] = fmaf(s, y1.w, sacc[7]);
      }
    }
    __syncthreads();
  }

  float* out = p.part + (size_t)split * p.K * p.Nout;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int k = k0 + 4 * tk + (i & 3) + 64 * (i >> 2);
    if (k >= p.K) continue;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int n = n0 + 4 * tn + 64 * j;
      if (n >= p.Nout) continue;
      st4(out + (size_t)k * p.Nout + n, make_float4(acc[i][4 * j], acc[i][4 * j + 1], acc[i][4 * j + 2], acc[i][4 * j + 3]));
    }
  }
  if (do_side && tk < nside) {
    float* so = p.part_side + ((size_t)split * nside + tk) * p.Nout;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int n = n0 + 4 * tn + 64 * j;
      if (n >= p.Nout) continue;
      st4(so + n, make_float4(sacc[4 * j], sacc[4 * j + 1], sacc[4 * j + 2], sacc[4 * j + 3]));
    }
  }
}

}  // namespace msmp

using namespace msmp;

extern "C" int msmp_linear_fwd(const float* const* A, const int* lda, const int* ka, const int* aswish, int nseg,
                               const float* Wt, int ldw, const float* bias, const float* side, int lds, int r,
                               const float* Wside, const float* Zmul, int ldz, float* Ypre, int ldpre, int act,
                               const float* R, int ldr, float* Y, int ldy, int M, int Nout, cudaStream_t stream) {
  if (nseg < 1 || nseg > 3 || M < 0 || Nout <= 0 || (Nout & 3) || r < 0 || r > 8) return MSMP_ERR_ARG;
  if (M == 0) return MSMP_OK;
  LinParams p{};
  for (int s = 0; s < nseg; ++s) {
    if (ka[s] <= 0 || ka[s] % LIN_KC || (lda[s] & 3)) return MSMP_ERR_ARG;
    p.A[s] = A[s];
    p.lda[s] = lda[s];
    p.ka[s] = ka[s];
    p.aswish[s] = aswish ? aswish[s] : 0;
  }
  const int ntile = (Nout + 127) / 128;
  if (ldw < ntile * 128 || (ldw & 3)) return MSMP_ERR_ARG;
  p.nseg = nseg; p.Wt = Wt; p.ldw = ldw; p.bias = bias; p.side = side; p.lds = lds; p.r = side ? r : 0;
  p.Wside = Wside; p.Zmul = Zmul; p.ldz = ldz; p.Ypre = Ypre; p.ldpre = ldpre; p.act = act; p.R = R; p.ldr = ldr;
  p.Y = Y; p.ldy = ldy; p.M = M; p.Nout = Nout;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(k_linear, cudaFuncAttributeMaxDynamicSharedMemorySize, LIN_SMEM) != cudaSuccess)
      return MSMP_ERR_CUDA;
    attr_set = true;
  }
  dim3 grid((M + 127) / 128, ntile);
  k_linear<<<grid, 256, LIN_SMEM, stream>>>(p);
  MSMP_CHECK_LAUNCH();
  return MSMP_OK;
}

// Split-M factor of the weight-gradient kernels.  MSMP_WGRAD_CTAS (CTAs aimed at per call) and MSMP_WGRAD_MIN_ROWS
// (rows per split at least) override the defaults for tuning runs.
static int wgrad_env(const char* name, int dflt) {
  const char* v = getenv(name);
  const int x = v ? atoi(v) : 0;
  return x > 0 ? x : dflt;
}
extern "C" int msmp_linear_wgrad_splits(int M, int K, int Nout) {
  static const int want_ctas = wgrad_env("MSMP_WGRAD_CTAS", 2 * 148);
  static const int min_rows = wgrad_env("MSMP_WGRAD_MIN_ROWS", 8 * WG_MC);      // 256: C2 step 3.94 -> 3.84 ms vs 128
  if (M <= 0) return 1;
  int tiles = ((K + 127) / 128) * ((Nout + 127) / 128);
  int want = (want_ctas + tiles - 1) / tiles;
  int max_splits = (M + min_rows - 1) / min_rows;
  int s = want < max_splits ? want : max_splits;
  return s < 1 ? 1 : s;
}

extern "C" size_t msmp_linear_wgrad_workspace(int M, int K, int Nout, int nside) {
  size_t S = (size_t)msmp_linear_wgrad_splits(M, K, Nout);
  return S * ((size_t)K * Nout + (size_t)nside * Nout) * sizeof(float);
}

extern "C" int msmp_linear_wgrad(const float* X, int ldx, int K, int xswish, const float* dY, int lddy, int Nout,
                                 const float* side, int lds, int r, int has_bias, float* dWt, float* dWside,
                                 int accumulate, int M, void* workspace, size_t ws_bytes, cudaStream_t stream) {
  if (K <= 0 || (K & 3) || Nout <= 0 || (Nout & 3) || (ldx & 3) || (lddy & 3) || r < 0 || r + has_bias > 16)
    return MSMP_ERR_ARG;
  const int nside = (side ? r : 0) + (has_bias ? 1 : 0);
  if (ws_bytes < msmp_linear_wgrad_workspace(M, K, Nout, nside)) return MSMP_ERR_WORKSPACE;
  const int S = msmp_linear_wgrad_splits(M, K, Nout);
  WgradParams p{};
  p.X = X; p.ldx = ldx; p.K = K; p.xswish = xswish; p.dY = dY; p.lddy = lddy; p.Nout = Nout;
  p.side = side; p.lds = lds; p.r = side ? r : 0; p.has_bias = has_bias ? 1 : 0;
  p.part = reinterpret_cast<float*>(workspace);
  p.part_side = p.part + (size_t)S * K * Nout;
  p.M = M;
  int rps = (M + S - 1) / S;
  rps = ((rps + WG_MC - 1) / WG_MC) * WG_MC;
  if (rps < WG_MC) rps = WG_MC;
  p.rows_per_split = rps;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(k_wgrad, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM) != cudaSuccess)
      return MSMP_ERR_CUDA;
    attr_set = true;
  }
  dim3 grid(S, (K + 127) / 128, (Nout + 127) / 128);
  k_wgrad<<<grid, 256, WG_SMEM, stream>>>(p);
  MSMP_CHECK_LAUNCH();
  {
    const int c0 = K * Nout, c1 = nside * Nout;
    k_reduce_partials2<<<(c0 + c1 + 255) / 256, 256, 0, stream>>>(p.part, dWt, c0, (size_t)K * Nout, p.part_side, dWside,
                                                                 c1, (size_t)nside * Nout, S, accumulate);
    MSMP_CHECK_LAUNCH();
  }
  return MSMP_OK;
}
