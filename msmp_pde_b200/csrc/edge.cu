// Edge-level kernels of the message-passing layer (sm_100a, fp32 FFMA path).
//
// message()  (models_gnn.py:69-75 / 132-138):  m_e = sw(W2 sw(W1 f_e + b1) + b2)  with
// f_e = [x_i | x_j | u_i-u_j | pos_i-pos_j | v_i].  Layer 1 is linear in the gathered features, so it is
// evaluated per node (P = W1xi x + W1u u + W1p p + W1v v + b1,  Q = W1xj x - W1u u - W1p p; msmp_linear_fwd)
// and per edge z1_e = P[i] + Q[j]  (SURVEY.md 7.2 "algebraic shortcut"; Appendix A notation).
// aggregate  (aggr='mean', models_gnn.py:42,107): deterministic segmented mean over the dst-sorted (CSR)
// edge list -- one warp per destination segment, no atomics.
//
//   msmp_edge_fwd : 128-edge tiles (persistent CTAs): gather P[dst]+Q[src] -> swish -> x W2^T + b2 -> swish
//                   -> segmented mean into agg[N,128]; segments cut by a tile boundary go through a carry
//                   buffer and msmp's ordered fix-up kernel.
//   msmp_edge_bwd : dz2 = dagg[dst]/deg * sw'(z2); da1 = dz2 W2; dz1 = da1 * sw'(P[dst]+Q[src]);
//                   dP = segmented sum of dz1 by dst (CSR); dW2 / db2 per-CTA partials (fixed-order reduce).
//                   dQ (by source) is msmp_segment_reduce over the CSC permutation.
//   msmp_segment_reduce : standalone deterministic scatter-sum / scatter-mean (either index order).
#include "common.cuh"
#include "msmp_b200.h"

namespace msmp {

constexpr int ET = 128;          // edges per tile
constexpr int E_LD = 132;        // padded smem row (floats)
constexpr int EDGE_FWD_SMEM = (128 * 128 + ET * E_LD) * 4 + (2 * ET + ET + 8) * 4;
constexpr int EDGE_BWD_SMEM = (128 * 128 + 2 * ET * E_LD) * 4 + (2 * ET + ET + 8) * 4;

// Builds the list of destination segments of the current tile: seg_start[0..nseg], from s_dst[0..valid).
// Called by all 256 threads; contains __syncthreads.
__device__ __forceinline__ void build_segments(const int* s_dst, int valid, int* seg_start, int* s_misc) {
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  bool flag = false;
  if (tid < ET) flag = (tid < valid) && (tid == 0 || s_dst[tid] != s_dst[tid - 1]);
  unsigned b = __ballot_sync(0xffffffffu, flag);
  if (warp < 4 && lane == 0) s_misc[warp] = __popc(b);
  __syncthreads();
  if (tid < ET) {
    int base = 0;
    for (int w = 0; w < warp; ++w) base += s_misc[w];
    if (flag) seg_start[base + __popc(b & ((1u << lane) - 1u))] = tid;
  }
  if (tid == 0) {
    int n = s_misc[0] + s_misc[1] + s_misc[2] + s_misc[3];
    s_misc[4] = n;
    seg_start[n] = valid;
  }
  __syncthreads();
}

// Segmented row-sum of a tile held in smem (rows[e * E_LD + c]); one warp per segment.
// Complete segments are written to out[node] (scaled); cut segments go to carry[tile][0|1].
__device__ __forceinline__ void reduce_segments_smem(const float* rows, const int* s_dst, const int* seg_start, int nseg,
                                                     int e0, const int* __restrict__ rowptr,
                                                     const float* __restrict__ scale, float* __restrict__ out,
                                                     float* __restrict__ carry, int tile) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int s = warp; s < nseg; s += 8) {
    const int r0 = seg_start[s], r1 = seg_start[s + 1];
    float4 sum = zero4();
    for (int r = r0; r < r1; ++r) sum = add4(sum, *reinterpret_cast<const float4*>(rows + r * E_LD + 4 * lane));
    const int node = s_dst[r0];
    const bool left = (__ldg(rowptr + node) == e0 + r0);
    const bool right = (__ldg(rowptr + node + 1) == e0 + r1);
    if (left && right) {
      float sc = scale ? __ldg(scale + node) : 1.0f;
      st4(out + (size_t)node * 128 + 4 * lane, scale4(sum, sc));
    } else {
      st4(carry + ((size_t)tile * 2 + (left ? 1 : 0)) * 128 + 4 * lane, sum);
    }
  }
}

// ------------------------------------------------------------------------------------------ forward
struct EdgeFwdParams {
  const float* P; const float* Q; int ldpq;     // per-node projections (row stride ldpq)
  const int* src; const int* dst; const int* rowptr; const float* inv_deg;
  const float* W2t; const float* b2;            // W2t[k][n] = W2[n][k]
  float* z2;                                    // [E][128] pre-activation (saved for backward); nullable
  float* agg;                                   // [N][128] (pre-zeroed by the host wrapper)
  float* carry;                                 // [T][2][128]
  int E; int T;
};

__global__ void __launch_bounds__(256, 1) k_edge_fwd(const EdgeFwdParams p) {
  extern __shared__ __align__(16) float smem[];
  float* Ws = smem;                      // [128][128]
  float* At = smem + 128 * 128;          // [ET][E_LD]
  int* s_src = reinterpret_cast<int*>(At + ET * E_LD);
  int* s_dst = s_src + ET;
  int* seg_start = s_dst + ET;           // [ET + 1] (uses ET+... see size)
  int* s_misc = seg_start + ET + 1;      // [5]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tm = tid >> 4, tn = tid & 15;

  for (int i = tid; i < 128 * 32; i += 256) st4(Ws + 4 * i, ldg4(p.W2t + 4 * i));

  for (int tile = blockIdx.x; tile < p.T; tile += gridDim.x) {
    const int e0 = tile * ET;
    const int valid = min(ET, p.E - e0);
    __syncthreads();                     // previous tile fully consumed (also covers the Ws fill)
    if (tid < ET) {
      s_src[tid] = tid < valid ? __ldg(p.src + e0 + tid) : 0;
      s_dst[tid] = tid < valid ? __ldg(p.dst + e0 + tid) : -1;
    }
    __syncthreads();
    // gather + first activation: At[e][:] = sw(P[dst_e] + Q[src_e])
#pragma unroll 4
    for (int r = warp; r < ET; r += 8) {
      float4 v = zero4();
      if (r < valid) {
        float4 a = ldg4(p.P + (size_t)s_dst[r] * p.ldpq + 4 * lane);
        float4 b = ldg4(p.Q + (size_t)s_src[r] * p.ldpq + 4 * lane);
        v = swish4(add4(a, b));
      }
      st4(At + r * E_LD + 4 * lane, v);
    }
    build_segments(s_dst, valid, seg_start, s_misc);     // syncs => At complete
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    mma_rowA<128>(At, E_LD, Ws, acc, tm, tn);
    __syncthreads();                     // everyone done reading At
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = tm + 16 * i;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int col = tn * 4 + 64 * j;
        float4 z = add4(make_float4(acc[i][4 * j], acc[i][4 * j + 1], acc[i][4 * j + 2], acc[i][4 * j + 3]),
                        ldg4(p.b2 + col));
        if (p.z2 && r < valid) st4(p.z2 + (size_t)(e0 + r) * 128 + col, z);
        st4(At + r * E_LD + col, swish4(z));
      }
    }
    __syncthreads();
    reduce_segments_smem(At, s_dst, seg_start, s_misc[4], e0, p.rowptr, p.inv_deg, p.agg, p.carry, tile);
  }
}

// Ordered fix-up of segments cut by tile boundaries: the tile in which a node's segment starts owns it.
__global__ void k_carry_fix(const float* __restrict__ carry, const int* __restrict__ dst,
                            const int* __restrict__ rowptr, const float* __restrict__ scale,
                            float* __restrict__ out, int E, int T) {
  const int lane = threadIdx.x & 31;
  const int tile = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (tile >= T) return;
  const int e0 = tile * ET;
  const int e_last = min(E, e0 + ET) - 1;
  const int node = dst[e_last];
  const int seg_begin = rowptr[node], seg_end = rowptr[node + 1];
  if (seg_end <= e_last + 1) return;           // last segment of this tile is complete on the right
  if (seg_begin < e0) return;                  // started in an earlier tile: that tile owns it
  float4 sum = ldcg4(carry + ((size_t)tile * 2 + 1) * 128 + 4 * lane);
  for (int t = tile + 1; t < T && t * ET < seg_end; ++t)
    sum = add4(sum, ldcg4(carry + ((size_t)t * 2 + 0) * 128 + 4 * lane));
  float sc = scale ? scale[node] : 1.0f;
  st4(out + (size_t)node * 128 + 4 * lane, scale4(sum, sc));
}

// ------------------------------------------------------------------------------------------ backward
struct EdgeBwdParams {
  const float* P; const float* Q; int ldpq;
  const int* src; const int* dst; const int* rowptr; const float* inv_deg;
  const float* W2;                 // original layout [n][k]
  const float* z2;                 // [E][128]
  const float* dagg; int lddagg;   // [N][.]
  float* dz1;                      // [E][128]
  float* dP; int lddp;             // [N][.] (pre-zeroed)
  float* carry;                    // [T][2][128]
  float* dW2_part;                 // [G][128][128]  ([n][k])
  float* db2_part;                 // [G][128]
  int E; int T;
};

__global__ void __launch_bounds__(256, 1) k_edge_bwd(const EdgeBwdParams p) {
  extern __shared__ __align__(16) float smem[];
  float* Ws = smem;                      // W2[n][k]
  float* Dt = smem + 128 * 128;          // dz2 tile [e][n]
  float* A1 = Dt + ET * E_LD;            // a1 tile  [e][k]
  int* s_src = reinterpret_cast<int*>(A1 + ET * E_LD);
  int* s_dst = s_src + ET;
  int* seg_start = s_dst + ET;
  int* s_misc = seg_start + ET + 1;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tm = tid >> 4, tn = tid & 15;

  for (int i = tid; i < 128 * 32; i += 256) st4(Ws + 4 * i, ldg4(p.W2 + 4 * i));

  float dw[8][8];                        // dW2[n][k] accumulator, persistent across tiles
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) dw[i][j] = 0.f;
  float db = 0.f;                        // threads < 128: column sums of dz2

  for (int tile = blockIdx.x; tile < p.T; tile += gridDim.x) {
    const int e0 = tile * ET;
    const int valid = min(ET, p.E - e0);
    __syncthreads();
    if (tid < ET) {
      s_src[tid] = tid < valid ? __ldg(p.src + e0 + tid) : 0;
      s_dst[tid] = tid < valid ? __ldg(p.dst + e0 + tid) : -1;
    }
    __syncthreads();
    // dz2[e][:] = dagg[dst_e] * inv_deg[dst_e] * sw'(z2[e])
#pragma unroll 4
    for (int r = warp; r < ET; r += 8) {
      float4 v = zero4();
      if (r < valid) {
        const int d = s_dst[r];
        float4 g = ldg4(p.dagg + (size_t)d * p.lddagg + 4 * lane);
        float4 z = ldg4(p.z2 + (size_t)(e0 + r) * 128 + 4 * lane);
        float s = __ldg(p.inv_deg + d);
        v = make_float4(g.x * s * dswish(z.x), g.y * s * dswish(z.y), g.z * s * dswish(z.z), g.w * s * dswish(z.w));
      }
      st4(Dt + r * E_LD + 4 * lane, v);
    }
    build_segments(s_dst, valid, seg_start, s_misc);
    if (tid < 128) {
      float s = 0.f;
      for (int r = 0; r < valid; ++r) s += Dt[r * E_LD + tid];
      db += s;
    }
    // da1 = dz2 * W2   (reduction over n)
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    mma_rowA<128>(Dt, E_LD, Ws, acc, tm, tn);
    // epilogue: z1 = P[dst]+Q[src]; a1 -> smem; dz1 = da1 * sw'(z1) -> global
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = tm + 16 * i;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int col = tn * 4 + 64 * j;
        float4 a1 = zero4();
        if (r < valid) {
          float4 z1 = add4(ldg4(p.P + (size_t)s_dst[r] * p.ldpq + col), ldg4(p.Q + (size_t)s_src[r] * p.ldpq + col));
          a1 = swish4(z1);
          float4 d = make_float4(acc[i][4 * j] * dswish(z1.x), acc[i][4 * j + 1] * dswish(z1.y),
                                 acc[i][4 * j + 2] * dswish(z1.z), acc[i][4 * j + 3] * dswish(z1.w));
          st4(p.dz1 + (size_t)(e0 + r) * 128 + col, d);
        }
        st4(A1 + r * E_LD + col, a1);
      }
    }
    __syncthreads();
    // dW2[n][k] += sum_e dz2[e][n] * a1[e][k]
    mma_redmajor<ET>(Dt, E_LD, A1, E_LD, dw, tm, tn);
    // dP: segmented sum of the dz1 rows this CTA just wrote (L2 hits; bypass L1)
    {
      const int nseg = s_misc[4];
      for (int s = warp; s < nseg; s += 8) {
        const int r0 = seg_start[s], r1 = seg_start[s + 1];
        float4 sum = zero4();
        for (int r = r0; r < r1; ++r) sum = add4(sum, ldcg4(p.dz1 + (size_t)(e0 + r) * 128 + 4 * lane));
        const int node = s_dst[r0];
        const bool left = (__ldg(p.rowptr + node) == e0 + r0);
        const bool right = (__ldg(p.rowptr + node + 1) == e0 + r1);
        if (left && right) st4(p.dP + (size_t)node * p.lddp + 4 * lane, sum);
        else st4(p.carry + ((size_t)tile * 2 + (left ? 1 : 0)) * 128 + 4 * lane, sum);
      }
    }
  }
  // per-CTA partials
  float* o = p.dW2_part + (size_t)blockIdx.x * 128 * 128;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int n = 4 * tm + (i & 3) + 64 * (i >> 2);
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int k = 4 * tn + 64 * j;
      st4(o + n * 128 + k, make_float4(dw[i][4 * j], dw[i][4 * j + 1], dw[i][4 * j + 2], dw[i][4 * j + 3]));
    }
  }
  if (tid < 128) p.db2_part[(size_t)blockIdx.x * 128 + tid] = db;
}

// carry fix-up with an output row stride (dP lives inside a wider [N][ld] buffer)
__global__ void k_carry_fix_ld(const float* __restrict__ carry, const int* __restrict__ dst,
                               const int* __restrict__ rowptr, float* __restrict__ out, int ldo, int E, int T) {
  const int lane = threadIdx.x & 31;
  const int tile = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (tile >= T) return;
  const int e0 = tile * ET;
  const int e_last = min(E, e0 + ET) - 1;
  const int node = dst[e_last];
  const int seg_begin = rowptr[node], seg_end = rowptr[node + 1];
  if (seg_end <= e_last + 1 || seg_begin < e0) return;
  float4 sum = ldcg4(carry + ((size_t)tile * 2 + 1) * 128 + 4 * lane);
  for (int t = tile + 1; t < T && t * ET < seg_end; ++t)
    sum = add4(sum, ldcg4(carry + ((size_t)t * 2 + 0) * 128 + 4 * lane));
  st4(out + (size_t)node * ldo + 4 * lane, sum);
}

// ------------------------------------------------------------------------ standalone segmented reduce
// out[n][:] = scale[n] * sum_{k in [ptr[n], ptr[n+1])} src[perm ? perm[k] : k][:]      (C = 128)
__global__ void __launch_bounds__(256) k_segment_reduce(const float* __restrict__ src, int lds,
                                                        const int* __restrict__ perm, const int* __restrict__ ptr,
                                                        const float* __restrict__ scale, float* __restrict__ out,
                                                        int ldo, int N) {
  const int lane = threadIdx.x & 31;
  const int node = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (node >= N) return;
  const int k0 = __ldg(ptr + node), k1 = __ldg(ptr + node + 1);
  float4 sum = zero4();
  int k = k0;
  for (; k + 4 <= k1; k += 4) {       // 4 independent row loads in flight
    int r0 = perm ? __ldg(perm + k) : k, r1 = perm ? __ldg(perm + k + 1) : k + 1;
    int r2 = perm ? __ldg(perm + k + 2) : k + 2, r3 = perm ? __ldg(perm + k + 3) : k + 3;
    float4 a = ldg4(src + (size_t)r0 * lds + 4 * lane), b = ldg4(src + (size_t)r1 * lds + 4 * lane);
    float4 c = ldg4(src + (size_t)r2 * lds + 4 * lane), d = ldg4(src + (size_t)r3 * lds + 4 * lane);
    sum = add4(add4(add4(add4(sum, a), b), c), d);    // fixed left-to-right order
  }
  for (; k < k1; ++k) {
    int r = perm ? __ldg(perm + k) : k;
    sum = add4(sum, ldg4(src + (size_t)r * lds + 4 * lane));
  }
  float sc = scale ? __ldg(scale + node) : 1.0f;
  st4(out + (size_t)node * ldo + 4 * lane, scale4(sum, sc));
}

}  // namespace msmp

using namespace msmp;

static int edge_grid(int T) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return T < sms ? T : sms;
}

extern "C" int msmp_edge_tiles(int E) { return (E + ET - 1) / ET; }
extern "C" int msmp_edge_grid(int E) { return edge_grid(msmp_edge_tiles(E)); }

extern "C" size_t msmp_edge_fwd_workspace(int E) { return (size_t)msmp_edge_tiles(E) * 2 * 128 * sizeof(float); }

extern "C" int msmp_edge_fwd(const float* P, const float* Q, int ldpq, const int* src, const int* dst,
                             const int* rowptr, const float* inv_deg, const float* W2t, const float* b2, float* z2,
                             float* agg, int E, int N, void* workspace, size_t ws_bytes, cudaStream_t stream) {
  if (E < 0 || N < 0 || (ldpq & 3)) return MSMP_ERR_ARG;
  if (cudaMemsetAsync(agg, 0, (size_t)N * 128 * sizeof(float), stream) != cudaSuccess) return MSMP_ERR_CUDA;
  if (E == 0) return MSMP_OK;
  if (ws_bytes < msmp_edge_fwd_workspace(E)) return MSMP_ERR_WORKSPACE;
  const int T = msmp_edge_tiles(E);
  EdgeFwdParams p{P, Q, ldpq, src, dst, rowptr, inv_deg, W2t, b2, z2, agg, reinterpret_cast<float*>(workspace), E, T};
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(k_edge_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, EDGE_FWD_SMEM) != cudaSuccess)
      return MSMP_ERR_CUDA;
    attr_set = true;
  }
  k_edge_fwd<<<edge_grid(T), 256, EDGE_FWD_SMEM, stream>>>(p);
  MSMP_CHECK_LAUNCH();
  k_carry_fix<<<(T + 7) / 8, 256, 0, stream>>>(p.carry, dst, rowptr, inv_deg, agg, E, T);
  MSMP_CHECK_LAUNCH();
  return MSMP_OK;
}

extern "C" size_t msmp_edge_bwd_workspace(int E) {
  size_t T = (size_t)msmp_edge_tiles(E);
  size_t G = (size_t)msmp_edge_grid(E);
  return (T * 2 * 128 + G * 128 * 128 + G * 128) * sizeof(float);
}

extern "C" int msmp_edge_bwd(const float* P, const float* Q, int ldpq, const int* src, const int* dst,
                             const int* rowptr, const float* inv_deg, const float* W2, const float* z2,
                             const float* dagg, int lddagg, float* dz1, float* dP, int lddp, float* dW2, float* db2,
                             int E, int N, void* workspace, size_t ws_bytes, cudaStream_t stream) {
  if (E < 0 || N < 0 || (ldpq & 3) || (lddagg & 3) || (lddp & 3)) return MSMP_ERR_ARG;
  if (cudaMemset2DAsync(dP, (size_t)lddp * sizeof(float), 0, 128 * sizeof(float), N, stream) != cudaSuccess)
    return MSMP_ERR_CUDA;
  if (E == 0) {
    if (cudaMemsetAsync(dW2, 0, 128 * 128 * sizeof(float), stream) != cudaSuccess) return MSMP_ERR_CUDA;
    if (cudaMemsetAsync(db2, 0, 128 * sizeof(float), stream) != cudaSuccess) return MSMP_ERR_CUDA;
    return MSMP_OK;
  }
  if (ws_bytes < msmp_edge_bwd_workspace(E)) return MSMP_ERR_WORKSPACE;
  const int T = msmp_edge_tiles(E);
  const int G = edge_grid(T);
  float* carry = reinterpret_cast<float*>(workspace);
  float* dW2_part = carry + (size_t)T * 2 * 128;
  float* db2_part = dW2_part + (size_t)G * 128 * 128;
  EdgeBwdParams p{P, Q, ldpq, src, dst, rowptr, inv_deg, W2, z2, dagg, lddagg, dz1, dP, lddp, carry, dW2_part, db2_part, E, T};
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(k_edge_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, EDGE_BWD_SMEM) != cudaSuccess)
      return MSMP_ERR_CUDA;
    attr_set = true;
  }
  k_edge_bwd<<<G, 256, EDGE_BWD_SMEM, stream>>>(p);
  MSMP_CHECK_LAUNCH();
  k_carry_fix_ld<<<(T + 7) / 8, 256, 0, stream>>>(carry, dst, rowptr, dP, lddp, E, T);
  MSMP_CHECK_LAUNCH();
  k_reduce_partials<<<(128 * 128 + 255) / 256, 256, 0, stream>>>(dW2_part, dW2, 128 * 128, G, (size_t)128 * 128, 0);
  MSMP_CHECK_LAUNCH();
  k_reduce_partials<<<1, 128, 0, stream>>>(db2_part, db2, 128, G, (size_t)128, 0);
  MSMP_CHECK_LAUNCH();
  return MSMP_OK;
}

extern "C" int msmp_segment_reduce(const float* src, int lds, const int* perm, const int* ptr, const float* scale,
                                   float* out, int ldo, int N, cudaStream_t stream) {
  if (N < 0 || (lds & 3) || (ldo & 3)) return MSMP_ERR_ARG;
  if (N == 0) return MSMP_OK;
  k_segment_reduce<<<(N + 7) / 8, 256, 0, stream>>>(src, lds, perm, ptr, scale, out, ldo, N);
  MSMP_CHECK_LAUNCH();
  return MSMP_OK;
}
