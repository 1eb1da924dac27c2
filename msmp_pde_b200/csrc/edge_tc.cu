// Edge kernels on the tensor cores (tcgen05.mma kind::tf32, 3xTF32, accumulator in TMEM).
//
// Same contracts as msmp_edge_fwd / msmp_edge_bwd (edge.cu) -- see there for the math.  Persistent CTAs walk
// 128-edge tiles of the destination-sorted edge list:
//   stage   : threads gather the operand rows (fwd: P[dst] + Q[src] -> swish;  bwd: dagg[dst]/deg * sw'(z2)),
//             32 columns at a time, split into tf32 hi/lo and store in the UMMA 128B-swizzle layout (2-stage ring,
//             next chunk's gathers in flight in registers)
//   weights : W2 (fwd: W2[n][k];  bwd: W2^T) as pre-split pre-swizzled images, resident in shared memory for the
//             whole kernel (128 KiB, four bulk copies at kernel start)
//   MMA     : 4 chunks x 4 k-steps x 3 products per tile, one issuing thread, tcgen05.commit per chunk
//   epilogue: TMEM -> registers (one thread per edge row): bias/swish (fwd) or sw'(P+Q) (bwd), row written to
//             global (z2 | dz1, a1) and to a swizzled shared tile; one warp per destination segment sums it
//             (fixed order, no atomics); segments cut by a tile boundary use the carry buffer + ordered fix-up.
// Backward materialises dz2 and a1 so that dW2 / db2 come from the generic msmp_linear_wgrad_tc.
#include "umma.cuh"
#include "msmp_b200.h"

namespace msmp {

constexpr int ETC_TILE = 128;
constexpr int ETC_W_BYTES = 4 * 2 * IMG_BYTES;         // 4 k-chunks x (hi | lo)
constexpr int ETC_A_BYTES = 2 * IMG_BYTES;             // one stage: A_hi, A_lo
constexpr int ETC_SMEM = ETC_W_BYTES + 2 * ETC_A_BYTES + 1024 + 256 + (3 * ETC_TILE + 16) * 4;

struct EdgeTcParams {
  const float* P; const float* Q; int ldpq;
  const int* src; const int* dst; const int* rowptr; const float* inv_deg;
  const float* Wimg;            // [4][2][4096] images (fwd: of W2t;  bwd: of W2)
  const float* b2;              // fwd only
  float* z2;                    // fwd: out (nullable);  bwd: in
  const float* dagg; int lddagg;     // bwd
  float* dz2; float* a1; float* dz1; // bwd outs [E][128]
  float* out; int ldo;          // fwd: agg [N][128] (scaled by inv_deg);  bwd: dP [N][ldo]
  float* carry;                 // [T][2][128]
  int E; int T;
};

// swizzled fp32 tile in shared memory: logical 16-byte chunk j of row r lives at chunk slot j ^ (r & 31)
__device__ __forceinline__ float* mt_ptr(float* mt, int r, int j) { return mt + r * 128 + ((j ^ (r & 31)) << 2); }

template <bool BWD>
__global__ void __launch_bounds__(256, 1) k_edge_tc(const EdgeTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smW = smem;
  uint8_t* smA = smem + ETC_W_BYTES;
  float* mt = reinterpret_cast<float*>(smA);                         // [128][128] swizzled tile (aliases the A ring)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smA + 2 * ETC_A_BYTES);   // wfull, done[2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
  int* s_src = reinterpret_cast<int*>(smA + 2 * ETC_A_BYTES + 256);
  int* s_dst = s_src + ETC_TILE;
  int* seg_start = s_dst + ETC_TILE;                                 // [129]
  int* s_misc = seg_start + ETC_TILE + 1;                            // [5]
  const int tid = threadIdx.x, warp = warp_index_uniform(), lane = tid & 31;

  if (warp == 0) tmem_alloc(tmem_slot, 128);
  if (tid == 32) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_init(&bars[2], 1);
    fence_barrier_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  constexpr uint32_t IDESC = umma_idesc_tf32(128, 128, 0, 0);
  if (tid == 0) {
    mbar_expect_tx(&bars[0], ETC_W_BYTES);
    for (int c = 0; c < 4; ++c) bulk_g2s(smW + c * 2 * IMG_BYTES, p.Wimg + (size_t)c * 2 * (IMG_BYTES / 4), 2 * IMG_BYTES, &bars[0]);
  }
  uint32_t nchunk = 0;            // running chunk counter (stage / parity bookkeeping across tiles)
  bool w_ready = false;

  for (int tile = blockIdx.x; tile < p.T; tile += gridDim.x) {
    const int e0 = tile * ETC_TILE;
    const int valid = min(ETC_TILE, p.E - e0);
    if (tid < ETC_TILE) {
      s_src[tid] = tid < valid ? __ldg(p.src + e0 + tid) : 0;
      s_dst[tid] = tid < valid ? __ldg(p.dst + e0 + tid) : -1;
    }
    __syncthreads();
    // ---- operand staging + MMAs, 4 chunks of 32 columns
    float4 ga[4], gb[4];
    auto gather = [&](int c) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int idx = tid + 256 * i;
        const int r = idx >> 3, col = 32 * c + 4 * (idx & 7);
        if (r < valid) {
          if (!BWD) {
            ga[i] = ldg4(p.P + (size_t)s_dst[r] * p.ldpq + col);
            gb[i] = ldg4(p.Q + (size_t)s_src[r] * p.ldpq + col);
          } else {
            ga[i] = ldg4(p.dagg + (size_t)s_dst[r] * p.lddagg + col);
            gb[i] = ldg4(p.z2 + (size_t)(e0 + r) * 128 + col);
          }
        } else {
          ga[i] = zero4();
          gb[i] = zero4();
        }
      }
    };
    gather(0);
    for (int c = 0; c < 4; ++c, ++nchunk) {
      const int s = nchunk & 1;
      uint8_t* st = smA + s * ETC_A_BYTES;
      if (nchunk >= 2) mbar_wait_warp(&bars[1 + s], ((nchunk >> 1) - 1) & 1);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int idx = tid + 256 * i;
        const int r = idx >> 3, c16 = idx & 7;
        float4 v;
        if (!BWD) {
          v = swish4(add4(ga[i], gb[i]));
          if (r >= valid) v = zero4();
        } else {
          const float sc = (r < valid) ? __ldg(p.inv_deg + s_dst[r]) : 0.f;
          v = make_float4(ga[i].x * sc * dswish(gb[i].x), ga[i].y * sc * dswish(gb[i].y), ga[i].z * sc * dswish(gb[i].z),
                          ga[i].w * sc * dswish(gb[i].w));
          if (r < valid) st4(p.dz2 + (size_t)(e0 + r) * 128 + 32 * c + 4 * c16, v);
        }
        store_split4(st, st + IMG_BYTES, img_off(r, c16), v);
      }
      if (c + 1 < 4) gather(c + 1);
      fence_proxy_async();
      __syncthreads();
      if (warp == 0) {      // the whole warp runs the issue code convergently, one elected lane issues (see elect_one())
        if (!w_ready) {
          mbar_wait(&bars[0], 0);
          w_ready = true;
        }
        tc_fence_after();
        const uint32_t tm = __shfl_sync(0xffffffffu, tmem, 0);
        const uint32_t a_hi = smem_u32(st), a_lo = a_hi + IMG_BYTES;
        const uint32_t b_hi = smem_u32(smW + c * 2 * IMG_BYTES), b_lo = b_hi + IMG_BYTES;
        const bool leader = elect_one();
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t dah = umma_desc(a_hi + 32 * k, 16, 1024), dal = umma_desc(a_lo + 32 * k, 16, 1024);
          const uint64_t dbh = umma_desc(b_hi + 32 * k, 16, 1024), dbl = umma_desc(b_lo + 32 * k, 16, 1024);
          if (leader) {
            umma_tf32(tm, dah, dbh, IDESC, (c | k) ? 1u : 0u);
            umma_tf32(tm, dal, dbh, IDESC, 1u);
            umma_tf32(tm, dah, dbl, IDESC, 1u);
          }
        }
        if (leader) umma_commit(&bars[1 + s]);
        __syncwarp();
      }
    }
    // ---- wait for the accumulator: both stages' last commits (also frees the A ring for the tile below)
    {
      const uint32_t l1 = nchunk - 1, l0 = nchunk - 2;
      mbar_wait_warp(&bars[1 + (l0 & 1)], (l0 >> 1) & 1);
      mbar_wait_warp(&bars[1 + (l1 & 1)], (l1 >> 1) & 1);
      tc_fence_after();
    }
    // ---- epilogue: thread = edge row r, 64 columns (two TMEM loads of 32)
    const int r = 32 * (warp & 3) + lane;
    const bool live = r < valid;
    const int my_dst = live ? s_dst[r] : 0, my_src = live ? s_src[r] : 0;
#pragma unroll 1
    for (int cb = 0; cb < 2; ++cb) {
      const int colbase = 64 * (warp >> 2) + 32 * cb;
      float v[32];
      __syncwarp();
      tmem_ld32(tmem + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)colbase, v);
      float4 pz[8];          // bwd: z1 = P[dst] + Q[src] of this 32-column block, loaded up front
      if (BWD && live) {
        float4 pa[8], qa[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          pa[j] = ldg4(p.P + (size_t)my_dst * p.ldpq + colbase + 4 * j);
          qa[j] = ldg4(p.Q + (size_t)my_src * p.ldpq + colbase + 4 * j);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) pz[j] = add4(pa[j], qa[j]);
      }
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const int col = colbase + j;
        float4 acc = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        float4 o = zero4();
        if (live) {
          if (!BWD) {
            float4 z = add4(acc, ldg4(p.b2 + col));
            if (p.z2) st4(p.z2 + (size_t)(e0 + r) * 128 + col, z);
            o = swish4(z);
          } else {
            const float4 z1 = pz[j >> 2];
            st4(p.a1 + (size_t)(e0 + r) * 128 + col, swish4(z1));
            o = make_float4(acc.x * dswish(z1.x), acc.y * dswish(z1.y), acc.z * dswish(z1.z), acc.w * dswish(z1.w));
            st4(p.dz1 + (size_t)(e0 + r) * 128 + col, o);
          }
        }
        st4(mt_ptr(mt, r, col >> 2), o);
      }
    }
    tc_fence_before();
    __syncthreads();
    // ---- destination segments of this tile
    {
      bool flag = false;
      if (tid < ETC_TILE) flag = (tid < valid) && (tid == 0 || s_dst[tid] != s_dst[tid - 1]);
      unsigned b = __ballot_sync(0xffffffffu, flag);
      if (warp < 4 && lane == 0) s_misc[warp] = __popc(b);
      __syncthreads();
      if (tid < ETC_TILE) {
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
    const int nseg = s_misc[4];
    for (int sg = warp; sg < nseg; sg += 8) {
      const int r0 = seg_start[sg], r1 = seg_start[sg + 1];
      float4 sum = zero4();
      for (int rr = r0; rr < r1; ++rr) sum = add4(sum, *reinterpret_cast<const float4*>(mt_ptr(mt, rr, lane)));
      const int node = s_dst[r0];
      const bool left = (__ldg(p.rowptr + node) == e0 + r0);
      const bool right = (__ldg(p.rowptr + node + 1) == e0 + r1);
      if (left && right) {
        const float sc = BWD ? 1.0f : __ldg(p.inv_deg + node);
        st4(p.out + (size_t)node * p.ldo + 4 * lane, scale4(sum, sc));
      } else {
        st4(p.carry + ((size_t)tile * 2 + (left ? 1 : 0)) * 128 + 4 * lane, sum);
      }
    }
    __syncthreads();          // the shared tile / index arrays are rewritten by the next tile
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 128);
}

// ordered fix-up of segments cut by tile boundaries (same rule as edge.cu: the tile where a segment starts owns it)
__global__ void k_carry_fix_tc(const float* __restrict__ carry, const int* __restrict__ dst,
                               const int* __restrict__ rowptr, const float* __restrict__ scale, float* __restrict__ out,
                               int ldo, int E, int T) {
  const int lane = threadIdx.x & 31;
  const int tile = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (tile >= T) return;
  const int e0 = tile * ETC_TILE;
  const int e_last = min(E, e0 + ETC_TILE) - 1;
  const int node = dst[e_last];
  const int seg_begin = rowptr[node], seg_end = rowptr[node + 1];
  if (seg_end <= e_last + 1 || seg_begin < e0) return;
  float4 sum = ldcg4(carry + ((size_t)tile * 2 + 1) * 128 + 4 * lane);
  for (int t = tile + 1; t < T && t * ETC_TILE < seg_end; ++t)
    sum = add4(sum, ldcg4(carry + ((size_t)t * 2 + 0) * 128 + 4 * lane));
  const float sc = scale ? scale[node] : 1.0f;
  st4(out + (size_t)node * ldo + 4 * lane, scale4(sum, sc));
}

static int sm_count() {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return sms;
}

template <bool BWD>
static int launch_edge_tc(const EdgeTcParams& p, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(k_edge_tc<BWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, ETC_SMEM) != cudaSuccess)
      return MSMP_ERR_CUDA;
    attr_set = true;
  }
  const int sms = sm_count();
  k_edge_tc<BWD><<<p.T < sms ? p.T : sms, 256, ETC_SMEM, stream>>>(p);
  MSMP_CHECK_LAUNCH();
  return MSMP_OK;
}

}  // namespace msmp

using namespace msmp;

extern "C" int msmp_edge_tc_fwd(const float* P, const float* Q, int ldpq, const int* src, const int* dst,
                                const int* rowptr, const float* inv_deg, const float* W2t_img, const float* b2,
                                float* z2, float* agg, int E, int N, void* workspace, size_t ws_bytes,
                                cudaStream_t stream) {
  if (E < 0 || N < 0 || (ldpq & 3)) return MSMP_ERR_ARG;
  if (cudaMemsetAsync(agg, 0, (size_t)N * 128 * sizeof(float), stream) != cudaSuccess) return MSMP_ERR_CUDA;
  if (E == 0) return MSMP_OK;
  if (ws_bytes < msmp_edge_fwd_workspace(E)) return MSMP_ERR_WORKSPACE;
  EdgeTcParams p{};
  p.P = P; p.Q = Q; p.ldpq = ldpq; p.src = src; p.dst = dst; p.rowptr = rowptr; p.inv_deg = inv_deg;
  p.Wimg = W2t_img; p.b2 = b2; p.z2 = z2; p.out = agg; p.ldo = 128;
  p.carry = reinterpret_cast<float*>(workspace); p.E = E; p.T = msmp_edge_tiles(E);
  int rc = launch_edge_tc<false>(p, stream);
  if (rc) return rc;
  k_carry_fix_tc<<<(p.T + 7) / 8, 256, 0, stream>>>(p.carry, dst, rowptr, inv_deg, agg, 128, E, p.T);
  MSMP_CHECK_LAUNCH();
  return MSMP_OK;
}

// Writes dz2, a1, dz1 [E,128] and dP (segmented sum of dz1 by destination).  dW2/db2: msmp_linear_wgrad_tc(a1, dz2).
extern "C" int msmp_edge_tc_bwd(const float* P, const float* Q, int ldpq, const int* src, const int* dst,
                                const int* rowptr, const float* inv_deg, const float* W2_img, const float* z2,
                                const float* dagg, int lddagg, float* dz2, float* a1, float* dz1, float* dP, int lddp,
                                int E, int N, void* workspace, size_t ws_bytes, cudaStream_t stream) {
  if (E < 0 || N < 0 || (ldpq & 3) || (lddagg & 3) || (lddp & 3)) return MSMP_ERR_ARG;
  if (cudaMemset2DAsync(dP, (size_t)lddp * sizeof(float), 0, 128 * sizeof(float), N, stream) != cudaSuccess)
    return MSMP_ERR_CUDA;
  if (E == 0) return MSMP_OK;
  if (ws_bytes < msmp_edge_fwd_workspace(E)) return MSMP_ERR_WORKSPACE;
  EdgeTcParams p{};
  p.P = P; p.Q = Q; p.ldpq = ldpq; p.src = src; p.dst = dst; p.rowptr = rowptr; p.inv_deg = inv_deg;
  p.Wimg = W2_img; p.z2 = const_cast<float*>(z2); p.dagg = dagg; p.lddagg = lddagg; p.dz2 = dz2; p.a1 = a1; p.dz1 = dz1;
  p.out = dP; p.ldo = lddp; p.carry = reinterpret_cast<float*>(workspace); p.E = E; p.T = msmp_edge_tiles(E);
  int rc = launch_edge_tc<true>(p, stream);
  if (rc) return rc;
  k_carry_fix_tc<<<(p.T + 7) / 8, 256, 0, stream>>>(p.carry, dst, rowptr, nullptr, dP, lddp, E, p.T);
  MSMP_CHECK_LAUNCH();
  return MSMP_OK;
}
