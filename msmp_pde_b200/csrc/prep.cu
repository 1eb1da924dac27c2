// Input assembly of a forward pass in two launches (instead of ~20 framework slice copies, fills and pads):
//
//   k_node_features   upad [N, ldu] = [u | 0]  and  side [N, 8] = [pos_x, variables..., 0]     (layers.NodeFeatures)
//   k_lem_inputs      inp [T, N, 32] = the zero-padded input slab of the LEM recurrence, column c of step t being a
//                     per-node constant, a time-indexed column of u, or the clock column cumsum(dt)_t + pos_t
//                     (I_t of experiments/models_gnn2D.py:421-433 and models_gnn.py:1357-1360)
//
// Pure data movement: every value is a copy of an fp32 input, except the clock column, which is formed in double and
// rounded once -- exactly what the framework expression (dt64 + pos_t.double()).float() does.
#include "common.cuh"
#include "msmp_b200.h"

namespace msmp {

struct LemColsArg {
  msmp_lem_col c[MSMP_LEM_MAX_COLS];
  int ncols;
};

// 8 threads per (t, n) row, one float4 (4 columns) each: a warp writes 4 complete 128-byte rows
__global__ void __launch_bounds__(256) k_lem_inputs(const LemColsArg a, const double* __restrict__ clock,
                                                   const double* __restrict__ node_t, int T, int N, float* __restrict__ inp) {
  const size_t idx = (size_t)blockIdx.x * 256 + threadIdx.x;
  const size_t row = idx >> 3;
  if (row >= (size_t)T * N) return;
  const int q = (int)(idx & 7);
  const int t = (int)(row / N), n = (int)(row - (size_t)t * N);
  float v[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c = 4 * q + j;
    if (c < a.ncols) {
      const msmp_lem_col col = a.c[c];
      if (col.kind == MSMP_LEM_COL_CLOCK) {
        v[j] = (float)(clock[t] + node_t[n]);
      } else {
        const float* src = reinterpret_cast<const float*>(col.src);
        v[j] = __ldg(src + (size_t)n * col.ld + col.off + (col.kind == MSMP_LEM_COL_TIME ? t : 0));
      }
    }
  }
  reinterpret_cast<float4*>(inp)[idx] = make_float4(v[0], v[1], v[2], v[3]);
}

// one thread per float4 of the concatenated output row [upad (ldu floats) | side (8 floats)]
__global__ void __launch_bounds__(256) k_node_features(const float* __restrict__ u, int F_u, const float* __restrict__ pos_x,
                                                      const float* __restrict__ variables, int V, int N,
                                                      float* __restrict__ upad, int ldu, float* __restrict__ side) {
  const int per_row = (ldu + 8) >> 2;
  const size_t idx = (size_t)blockIdx.x * 256 + threadIdx.x;
  const size_t n = idx / per_row;
  if (n >= (size_t)N) return;
  const int q = (int)(idx - n * per_row);
  float v[4];
  if (4 * q < ldu) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = 4 * q + j;
      v[j] = c < F_u ? __ldg(u + n * F_u + c) : 0.f;
    }
    reinterpret_cast<float4*>(upad + n * ldu)[q] = make_float4(v[0], v[1], v[2], v[3]);
  } else {
    const int c0 = 4 * q - ldu;          // 0 or 4
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = c0 + j;
      v[j] = c == 0 ? __ldg(pos_x + n) : (c <= V ? __ldg(variables + n * V + (c - 1)) : 0.f);
    }
    reinterpret_cast<float4*>(side + n * 8)[c0 >> 2] = make_float4(v[0], v[1], v[2], v[3]);
  }
}

}  // namespace msmp

extern "C" int msmp_lem_inputs(const msmp_lem_col* cols, int ncols, const double* clock, const double* node_t, int T, int N,
                               float* inp, cudaStream_t stream) {
  if (!cols || ncols < 1 || ncols > MSMP_LEM_MAX_COLS || T < 0 || N < 0 || !inp) return MSMP_ERR_ARG;
  msmp::LemColsArg a;
  a.ncols = ncols;
  for (int c = 0; c < ncols; ++c) {
    a.c[c] = cols[c];
    if (cols[c].kind == MSMP_LEM_COL_CLOCK) {
      if (!clock || !node_t) return MSMP_ERR_ARG;
    } else if ((cols[c].kind != MSMP_LEM_COL_STATIC && cols[c].kind != MSMP_LEM_COL_TIME) || !cols[c].src) {
      return MSMP_ERR_ARG;
    }
  }
  const size_t threads = (size_t)T * N * 8;
  if (threads == 0) return MSMP_OK;
  const size_t nb = (threads + 255) / 256;
  if (nb > 0x7fffffffu) return MSMP_ERR_ARG;
  msmp::k_lem_inputs<<<(unsigned)nb, 256, 0, stream>>>(a, clock, node_t, T, N, inp);
  MSMP_CHECK_LAUNCH();
  return MSMP_OK;
}

extern "C" int msmp_node_features(const float* u, int F_u, const float* pos_x, const float* variables, int V, int N,
                                  float* upad, int ldu, float* side, cudaStream_t stream) {
  if (!u || !pos_x || (V > 0 && !variables) || !upad || !side || F_u < 1 || V < 0 || V > 7 || ldu < F_u || (ldu & 3) || N < 0)
    return MSMP_ERR_ARG;
  if (N == 0) return MSMP_OK;
  const size_t threads = (size_t)N * ((ldu + 8) >> 2);
  const size_t nb = (threads + 255) / 256;
  if (nb > 0x7fffffffu) return MSMP_ERR_ARG;
  msmp::k_node_features<<<(unsigned)nb, 256, 0, stream>>>(u, F_u, pos_x, variables, V, N, upad, ldu, side);
  MSMP_CHECK_LAUNCH();
  return MSMP_OK;
}
