// Weight gradients on the tensor cores: dWt[k][n] = sum_m X[m][k] * dY[m][n]  (tcgen05 kind::tf32, 3xTF32).
//
// The reduction index m is the MMA K dimension, so BOTH operands are MN-major.  For 32-bit MN-major operands the
// only shared-memory layout the tensor core accepts is SWIZZLE_128B_BASE32B (CUTLASS sm100_common.inl:92): an atom
// is 4 K-rows x 128 B (32 MN-contiguous floats per row) and the four 32-byte units of a row are XOR-permuted with
// (row & 3) (Swizzle<2,5,2> on byte addresses; verified on hardware by scripts/diag history).  A chunk of 32 rows of X
// ([32 x 128] fp32, rows = MMA-K, columns = MMA-M) is stored as four column blocks of [32 rows x 128 B]:
//     descriptor start = base + kstep * 1024 (8 rows),  LBO = 4096 (next 32 columns),  SBO = 512 (next 4 rows).
// Per CTA: one 128 (k) x 128 (n) tile of dWt over a contiguous range of rows (split-M), accumulated in TMEM;
// the per-split partials are summed in a fixed order afterwards (deterministic, no atomics).  Side / bias
// gradients ([side | 1]^T dY) are accumulated on the CUDA cores by the threads that stage the dY operand.
#include "umma.cuh"
#include "msmp_b200.h"

namespace msmp {

constexpr int WT_CHUNK = 32;                               // rows (MMA-K) per stage
constexpr int WT_OP_BYTES = WT_CHUNK * 128 * 4;            // one [32 x 128] operand image = 16 KiB
constexpr int WT_STAGE_BYTES = 4 * WT_OP_BYTES;            // X_hi, X_lo, dY_hi, dY_lo
constexpr int WT_TAIL_COLS = 64;                           // a narrow second column block (K1 <= 64) rides along, see below
constexpr int WT_TAIL_OP_BYTES = WT_CHUNK * WT_TAIL_COLS * 4;     // one [32 x 64] operand image = 8 KiB
constexpr int WT_TAIL_STAGE_BYTES = 2 * WT_TAIL_OP_BYTES;  // X1_hi, X1_lo
constexpr int WT_SMEM = 2 * WT_STAGE_BYTES + 2 * WT_TAIL_STAGE_BYTES + 1024 + 256 + 8 * 8 * 128 * 4;

struct WgradTcParams {
  const float* X; int ldx; int K; int xswish;     // K = rows of dWt in total
  const float* X1; int ldx1; int K0;              // optional second column block of X: columns K0.. come from X1
  const float* dY; int lddy; int Nout;
  const float* side; int lds; int r; int has_bias;
  float* part;         // [S][K][Nout]
  float* part_side;    // [S][nside][Nout]
  int M; int rows_per_split;
  int tail;            // > 0: X1 has `tail` (<= 64) columns and is NOT a k-block of the grid: the CTAs of k-block 0 also
                       // compute D'[n][k1] = dY^T X1 (operands swapped: M = 128 outputs, N = 64) into TMEM columns 128..191
};

// byte offset of (row m in 0..31, 16-byte chunk c4 in 0..31) of a [32 x 128] MN-major (BASE32B) operand image:
// column block (c4 >> 3), row m, and inside the 128-byte row the 32-byte unit index is XORed with (m & 3).
__device__ __forceinline__ uint32_t mn_off(int m, int c4) {
  const int c = c4 & 7;
  const int cs = ((((c >> 1) ^ m) & 3) << 1) | (c & 1);
  return (uint32_t)((c4 >> 3) * 4096 + m * 128 + cs * 16);
}

__global__ void __launch_bounds__(256, 1) k_wgrad_tc(const WgradTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_tail = smem + 2 * WT_STAGE_BYTES;                              // 2 stages of (X1_hi | X1_lo)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_tail + 2 * WT_TAIL_STAGE_BYTES);     // free[2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
  float* sred = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);        // [8 warps][8 side rows][128]
  const int tid = threadIdx.x, warp = warp_index_uniform(), lane = tid & 31;
  const int split = blockIdx.x;
  const int k0 = blockIdx.y * 128, n0 = blockIdx.z * 128;
  // column block of X handled by this CTA: [X | X1] concatenated at K0 (K0 % 128 == 0 when X1 is given)
  const bool second = p.X1 != nullptr && k0 >= p.K0;
  const float* Xb = second ? p.X1 : p.X;
  const int ldxb = second ? p.ldx1 : p.ldx;
  const int kloc = second ? k0 - p.K0 : k0;                 // first column inside the selected block
  const int kvalid = (second ? p.K - p.K0 : (p.X1 ? p.K0 : p.K)) - kloc;   // valid columns from kloc on
  const int m_begin = split * p.rows_per_split;
  const int m_end = min(p.M, m_begin + p.rows_per_split);
  const int nchunks = (m_end > m_begin) ? (m_end - m_begin + WT_CHUNK - 1) / WT_CHUNK : 0;
  const int nside = p.r + p.has_bias;
  const bool do_side = (blockIdx.y == 0) && nside > 0;
  const bool do_tail = p.tail > 0 && blockIdx.y == 0;

  if (warp == 0) tmem_alloc(tmem_slot, 256);
  if (tid == 32) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_barrier_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  constexpr uint32_t IDESC = umma_idesc_tf32(128, 128, 1, 1);      // both operands MN-major

  // each thread stages 4 float4 of X and 4 of dY per chunk: element idx -> (row mm = idx >> 5, chunk c4 = idx & 31)
  // Two register sets: the rows of chunks c+1 and c+2 are in flight while chunk c is split / stored and its MMAs are
  // issued (with a single set the loads were issued a few hundred cycles before they were needed).
  float4 pxA[4], pyA[4], pxB[4], pyB[4];
  float4 ptA[2], ptB[2];      // the narrow tail block: 32 rows x 64 columns = 2 float4 per thread
  float sacc[8][4];
#pragma unroll
  for (int q = 0; q < 8; ++q) sacc[q][0] = sacc[q][1] = sacc[q][2] = sacc[q][3] = 0.f;
  auto prefetch = [&](int c, float4 (&px)[4], float4 (&py)[4], float4 (&pt)[2]) {
    const int mb = m_begin + c * WT_CHUNK;
    if (do_tail) {
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int idx = tid + 256 * i;
        const int m = mb + (idx >> 4), c4 = idx & 15;
        pt[i] = (m < m_end && 4 * c4 < p.tail) ? ldg4(p.X1 + (size_t)m * p.ldx1 + 4 * c4) : zero4();
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = tid + 256 * i;
      const int m = mb + (idx >> 5), c4 = idx & 31;
      const bool okm = m < m_end;
      px[i] = (okm && 4 * c4 < kvalid) ? ldg4(Xb + (size_t)m * ldxb + kloc + 4 * c4) : zero4();
      py[i] = (okm && n0 + 4 * c4 < p.Nout) ? ldg4(p.dY + (size_t)m * p.lddy + n0 + 4 * c4) : zero4();
    }
  };
  if (nchunks > 0) prefetch(0, pxA, pyA, ptA);
  if (nchunks > 1) prefetch(1, pxB, pyB, ptB);
  constexpr uint32_t IDESC_TAIL = umma_idesc_tf32(128, WT_TAIL_COLS, 1, 1);
  auto chunk = [&](int c, float4 (&px)[4], float4 (&py)[4], float4 (&pt)[2]) {
    const int s = c & 1, use = c >> 1;
    uint8_t* st = smem + s * WT_STAGE_BYTES;
    uint8_t* stt = smem_tail + s * WT_TAIL_STAGE_BYTES;
    if (c >= 2) mbar_wait_warp(&bars[s], (use - 1) & 1);
    const int mb = m_begin + c * WT_CHUNK;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = tid + 256 * i;
      const int mm = idx >> 5, c4 = idx & 31;
      float4 x = px[i];
      if (p.xswish) x = swish4(x);
      const uint32_t off = mn_off(mm, c4);
      store_split4(st, st + WT_OP_BYTES, off, x);
      store_split4(st + 2 * WT_OP_BYTES, st + 3 * WT_OP_BYTES, off, py[i]);
      if (do_side) {
        const int m = mb + mm;
        if (m < m_end) {
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            if (q >= nside) break;
            const float sv = (q < p.r) ? __ldg(p.side + (size_t)m * p.lds + q) : 1.0f;
            sacc[q][0] = fmaf(sv, py[i].x, sacc[q][0]);
            sacc[q][1] = fmaf(sv, py[i].y, sacc[q][1]);
            sacc[q][2] = fmaf(sv, py[i].z, sacc[q][2]);
            sacc[q][3] = fmaf(sv, py[i].w, sacc[q][3]);
          }
        }
      }
    }
    if (do_tail) {
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int idx = tid + 256 * i;
        float4 x = pt[i];
        if (p.xswish) x = swish4(x);
        store_split4(stt, stt + WT_TAIL_OP_BYTES, mn_off(idx >> 4, idx & 15), x);
      }
    }
    if (c + 2 < nchunks) prefetch(c + 2, px, py, pt);
    fence_proxy_async();
    __syncthreads();
    if (warp == 0) {      // the whole warp runs the issue code convergently, one elected lane issues (see elect_one())
      tc_fence_after();
      const uint32_t tm = __shfl_sync(0xffffffffu, tmem, 0);
      const uint32_t xh = smem_u32(st), xl = xh + WT_OP_BYTES, yh = xh + 2 * WT_OP_BYTES, yl = xh + 3 * WT_OP_BYTES;
      const bool leader = elect_one();
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t ko = 1024 * k;      // 8 rows per k-step; LBO = 4096 (next 32 columns), SBO = 512 (next 4 rows)
        const uint64_t dxh = umma_desc(xh + ko, 4096, 512, 1), dxl = umma_desc(xl + ko, 4096, 512, 1);
        const uint64_t dyh = umma_desc(yh + ko, 4096, 512, 1), dyl = umma_desc(yl + ko, 4096, 512, 1);
        if (leader) {
          umma_tf32(tm, dxh, dyh, IDESC, (c | k) ? 1u : 0u);
          umma_tf32(tm, dxl, dyh, IDESC, 1u);
          umma_tf32(tm, dxh, dyl, IDESC, 1u);
        }
        if (do_tail) {      // D'[n][k1] += dY^T X1: the dY images are the A operand here, the X1 images the (64-wide) B operand
          const uint32_t th = smem_u32(stt), tl = th + WT_TAIL_OP_BYTES;
          const uint64_t dth = umma_desc(th + ko, 4096, 512, 1), dtl = umma_desc(tl + ko, 4096, 512, 1);
          if (leader) {
            umma_tf32(tm + 128, dyh, dth, IDESC_TAIL, (c | k) ? 1u : 0u);
            umma_tf32(tm + 128, dyl, dth, IDESC_TAIL, 1u);
            umma_tf32(tm + 128, dyh, dtl, IDESC_TAIL, 1u);
          }
        }
      }
      if (leader) umma_commit(&bars[s]);
      __syncwarp();
    }
  };
  for (int c = 0; c < nchunks; c += 2) {
    chunk(c, pxA, pyA, ptA);
    if (c + 1 < nchunks) chunk(c + 1, pxB, pyB, ptB);
  }
  float* out = p.part + (size_t)split * p.K * p.Nout;
  if (nchunks > 0) {
    const int last = nchunks - 1;
    mbar_wait_warp(&bars[last & 1], (last >> 1) & 1);
    tc_fence_after();
  }
  // epilogue: warp w -> TMEM lanes 32*(w&3).. (= k rows), columns 64*(w>>2)..
  {
    const int krow = k0 + 32 * (warp & 3) + lane;
#pragma unroll 1
    for (int cb = 0; cb < 2; ++cb) {
      const int colbase = 64 * (warp >> 2) + 32 * cb;
      float v[32];
      if (nchunks > 0) {
        __syncwarp();
        tmem_ld32(tmem + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)colbase, v);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0.f;
      }
      if (krow < p.K) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const int n = n0 + colbase + j;
          if (n < p.Nout) st4(out + (size_t)krow * p.Nout + n, make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]));
        }
      }
    }
  }
  // narrow tail block: TMEM lanes = output columns n, columns = k1; rows K0 + k1 of the partial dWt
  if (do_tail) {
    const int n = n0 + 32 * (warp & 3) + lane;
    const int kb = 32 * (warp >> 2);
    float v[32];
    if (nchunks > 0) {
      __syncwarp();
      tmem_ld32(tmem + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(128 + kb), v);
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = 0.f;
    }
    if (n < p.Nout) {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (kb + j < p.tail) out[(size_t)(p.K0 + kb + j) * p.Nout + n] = v[j];
    }
  }
  // side / bias partials: fixed-order reduction over the 8 warps (thread lane owns columns 4*lane..4*lane+3)
  if (do_side) {
#pragma unroll
    for (int q = 0; q < 8; ++q)
      if (q < nside) st4(sred + (warp * 8 + q) * 128 + 4 * lane, make_float4(sacc[q][0], sacc[q][1], sacc[q][2], sacc[q][3]));
    __syncthreads();
    for (int i = tid; i < nside * 128; i += 256) {
      const int q = i >> 7, col = i & 127;
      float s = 0.f;
      for (int w = 0; w < 8; ++w) s += sred[(w * 8 + q) * 128 + col];
      if (n0 + col < p.Nout) p.part_side[((size_t)split * nside + q) * p.Nout + n0 + col] = s;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

}  // namespace msmp

using namespace msmp;

extern "C" int msmp_linear_wgrad_tc2(const float* X, int ldx, int K0, const float* X1, int ldx1, int K1, int xswish,
                                     const float* dY, int lddy, int Nout, const float* side, int lds, int r,
                                     int has_bias, float* dWt, float* dWside, int accumulate, int M, void* workspace,
                                     size_t ws_bytes, cudaStream_t stream);

extern "C" int msmp_linear_wgrad_tc(const float* X, int ldx, int K, int xswish, const float* dY, int lddy, int Nout,
                                    const float* side, int lds, int r, int has_bias, float* dWt, float* dWside,
                                    int accumulate, int M, void* workspace, size_t ws_bytes, cudaStream_t stream) {
  return msmp_linear_wgrad_tc2(X, ldx, K, nullptr, 0, 0, xswish, dY, lddy, Nout, side, lds, r, has_bias, dWt, dWside,
                               accumulate, M, workspace, ws_bytes, stream);
}

// dWt[K0 + K1, Nout] = [X | X1]^T dY : the two column blocks of the (virtual) concatenation come from different
// tensors (e.g. [h | agg] for update_net_1, [h | u] for the P|Q projection); K0 % 128 == 0 when X1 is given.
extern "C" int msmp_linear_wgrad_tc2(const float* X, int ldx, int K0, const float* X1, int ldx1, int K1, int xswish,
                                     const float* dY, int lddy, int Nout, const float* side, int lds, int r,
                                     int has_bias, float* dWt, float* dWside, int accumulate, int M, void* workspace,
                                     size_t ws_bytes, cudaStream_t stream) {
  const int K = K0 + (X1 ? K1 : 0);
  if (K <= 0 || (K & 3) || Nout <= 0 || (Nout & 3) || (ldx & 3) || (lddy & 3) || r < 0 || r + has_bias > 8)
    return MSMP_ERR_ARG;
  if (X1 && ((K0 & 127) || (K1 & 3) || (ldx1 & 3) || K1 <= 0)) return MSMP_ERR_ARG;
  const int nside = (side ? r : 0) + (has_bias ? 1 : 0);
  if (ws_bytes < msmp_linear_wgrad_workspace(M, K, Nout, nside)) return MSMP_ERR_WORKSPACE;
  int S = msmp_linear_wgrad_splits(M, K, Nout);
  const int tail = (X1 && K1 <= WT_TAIL_COLS) ? K1 : 0;
  if (tail) {      // fewer CTAs per split than the workspace query assumed: stay within one wave of 148 SMs
    const int tiles = ((K0 + 127) / 128) * ((Nout + 127) / 128);
    const int one_wave = 148 / tiles > 0 ? 148 / tiles : 1;
    if (S > one_wave) S = one_wave;
  }
  WgradTcParams p{};
  p.X = X; p.ldx = ldx; p.K = K; p.xswish = xswish; p.dY = dY; p.lddy = lddy; p.Nout = Nout;
  p.X1 = X1; p.ldx1 = ldx1; p.K0 = K0;
  p.side = side; p.lds = lds; p.r = side ? r : 0; p.has_bias = has_bias ? 1 : 0;
  p.part = reinterpret_cast<float*>(workspace);
  p.part_side = p.part + (size_t)S * K * Nout;
  p.M = M;
  // a narrow second block (the zero-padded inputs of the LEM maps, the u columns of the P|Q projection) does not get
  // k-block CTAs of its own -- three quarters of their MMA rows and of their staging would be padding
  p.tail = tail;
  int rps = (M + S - 1) / S;
  rps = ((rps + WT_CHUNK - 1) / WT_CHUNK) * WT_CHUNK;
  if (rps < WT_CHUNK) rps = WT_CHUNK;
  p.rows_per_split = rps;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(k_wgrad_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, WT_SMEM) != cudaSuccess)
      return MSMP_ERR_CUDA;
    attr_set = true;
  }
  dim3 grid(S, ((p.tail ? K0 : K) + 127) / 128, (Nout + 127) / 128);
  k_wgrad_tc<<<grid, 256, WT_SMEM, stream>>>(p);
  MSMP_CHECK_LAUNCH();
  {
    const int c0 = K * Nout, c1 = nside * Nout;
    k_reduce_partials2<<<(c0 + c1 + 255) / 256, 256, 0, stream>>>(p.part, dWt, c0, (size_t)K * Nout, p.part_side, dWside,
                                                                 c1, (size_t)nside * Nout, S, accumulate);
    MSMP_CHECK_LAUNCH();
  }
  return MSMP_OK;
}
