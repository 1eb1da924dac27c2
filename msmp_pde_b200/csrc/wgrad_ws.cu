// Weight gradients, second generation: persistent, warp-specialised, every operand read from HBM exactly once.
//
//     dWt[k][n]    = sum_m [X0 | X1 | X2][m][k] * dY[m][n]           (k < KB)
//     dWside[q][n] = sum_m [side[m][c0..c0+r) | 1][q] * dY[m][n]     (q < r + has_bias)
//
// (dW4 / dW3 / dW2 / dWpq of a message-passing layer, experiments/models_gnn.py:47-58 via autograd; the LEM maps,
// models_gnn.py:310-313; the embedding / output Linear layers.)  k_wgrad_tc (wgrad_tc.cu) gave every 128 x 128 output
// tile its own CTAs, so X and dY were re-read once per tile, each CTA ran load -> split -> MMA back to back, and a second
// launch summed the split-M partials.  Here ONE CTA owns a contiguous range of rows m for ALL output tiles it can hold
// in tensor memory:
//
//   loader warp    : the fp32 rows of a 16-row chunk (dY, up to three X segments, the side columns) arrive in a raw
//                    ring by 1-D bulk copies (TMA engine; one copy per operand when its rows are contiguous) -- loads
//                    never wait on registers, up to six chunks are in flight per SM
//   16 producer warps: raw rows -> optional swish -> bf16 operand images (MN-major, SWIZZLE_128B) in an image ring.
//                    fp32-parity mode: every fp32 value is split EXACTLY into three bf16 pieces x = b0 + b1 + b2 (+ 2^-24 x)
//                    -> three images per operand; bf16 mode: one image (x rounded to bf16).  The side columns and the
//                    bias "ones" column are 32 extra columns of the B operand, so side / bias gradients come out of the
//                    same MMAs (no CUDA-core accumulation)
//   MMA warp       : D'[n][k] += dY^T [X | side | 1]  -- dY blocks (128 columns) are the A operand (M = 128), the whole
//                    X row (up to 256 + 32 columns, N <= 256 per instruction) is the B operand, K = 16 rows per MMA
//                    (kind::f16, bf16 inputs, fp32 accumulation).  fp32-parity mode issues the six products
//                    a0 b0, a0 b1, a1 b0, a1 b1, a0 b2, a2 b0 (every bf16 x bf16 product is exact in fp32; the dropped terms
//                    are <= 2^-24 relative, i.e. fp32 rounding level: measured 2e-7 of max|ref| against float64) -- the same
//                    tensor-pipe time as 3xTF32 (6 x K16 bf16 = 6 x K8 tf32 per 16 rows) with 25 % fewer operand bytes, and
//                    it only needs the 16-bit MN-major layout (the 32-bit one, SWIZZLE_128B_BASE32B, is used by
//                    wgrad_tc.cu with 32-row blocks; with 16-row blocks its B operand read wrong columns on hardware).
//                    bf16 mode issues the a0 b0 product only.  Accumulators (lanes = n, columns = k) stay in TMEM for the
//                    CTA's whole row range
//   epilogue       : the producer warps drain TMEM into the CTA's partial [K][Nout] (coalesced 128-byte lines)
//
// The partials of the S row ranges are summed in a fixed order -- by k_reduce_partials2 (plain autograd use), or for free
// inside the one k_unpack launch that ends the captured training step's backward pass (gradsink.py): then a weight
// gradient is ONE launch.  Deterministic, no atomics.
#include <cstdlib>
#include <cstring>
#include <cuda.h>
#include <cuda_bf16.h>
#include "umma.cuh"
#include "msmp_b200.h"

namespace msmp {

constexpr int WW_R = 16;                 // rows (the MMA K dimension) per chunk
constexpr int WW_PROD_WARPS = 16;        // the fp32 -> bf16-piece conversion is ALU / latency bound: 8 warps left the SM idle
constexpr int WW_PROD_THREADS = 32 * WW_PROD_WARPS;
constexpr int WW_MMA_WARP = WW_PROD_WARPS, WW_LOAD_WARP = WW_PROD_WARPS + 1;
constexpr int WW_THREADS = 32 * (WW_PROD_WARPS + 2);
constexpr int WW_MAX_RAW = 6, WW_MAX_IMG = 4;
constexpr int WW_SMEM_LIMIT = 232448;    // 227 KiB

struct WgradWsParams {
  const float* X[3];
  int ldx[3];
  int kx[3];
  int xsw[3];
  int nseg;
  int KB;              // sum of kx: rows of dWt
  const float* dY;
  int lddy;
  int Nout;
  const float* side;   // base of the [M][lds] side array (16-byte aligned rows)
  int lds;
  int side_c0;         // first side column used
  int r;
  int has_bias;
  int KBS;             // B operand width = KB + 32 when side / bias rows are wanted: accumulator columns per dY block
  int npc;             // dY columns per CTA (128 or 256)
  float* part;         // [S][KB][Nout]
  float* part_side;    // [S][r + has_bias][Nout]
  int M;
  int rows_per_split;
  int nraw, nimg;      // ring depths
  int raw_bytes, img_bytes;
  int tmem_cols;
};

__device__ __forceinline__ void mbar_wait_all(uint64_t* bar, uint32_t parity) {
  // lane 0 polls with back-off, then every lane observes the completed phase itself (acquire for TMA-written data)
  mbar_wait_backoff(bar, parity);
  while (!mbar_try_wait(bar, parity)) {
  }
}
// byte offset of (row m, 16-byte chunk c8 = 8 bf16 columns) in an MN-major SWIZZLE_128B image (64 columns per block)
__device__ __forceinline__ uint32_t ww_off16(int m, int c8) {
  return (uint32_t)((c8 >> 3) * (WW_R * 128) + m * 128 + (((c8 ^ m) & 7) << 4));
}
__device__ __forceinline__ uint4 pack_bf16x8(float4 a, float4 b) {
  __nv_bfloat162 p0 = __floats2bfloat162_rn(a.x, a.y), p1 = __floats2bfloat162_rn(a.z, a.w);
  __nv_bfloat162 p2 = __floats2bfloat162_rn(b.x, b.y), p3 = __floats2bfloat162_rn(b.z, b.w);
  uint4 o;
  o.x = *reinterpret_cast<uint32_t*>(&p0);
  o.y = *reinterpret_cast<uint32_t*>(&p1);
  o.z = *reinterpret_cast<uint32_t*>(&p2);
  o.w = *reinterpret_cast<uint32_t*>(&p3);
  return o;
}
// x = b0 + b1 + b2 (+ <= 2^-25 |x|): b0 = bf16_rn(x) (one cvt.rn.bf16x2 per pair; |x - b0| <= 2^-9 |x|, either sign),
// b1 = the upper 16 bits of r1 = x - b0 (truncation: one LOP, |r1 - b1| < 2^-7 |r1|), b2 = bf16_rn(r1 - b1).  Both
// subtractions are exact.  The products the MMA warp leaves out (a1 b2, a2 b1, a2 b2) are then <= 2^-24 of |a b| and of
// random sign.
__device__ __forceinline__ uint32_t hi16x2(float lo_elem, float hi_elem) {      // {bf16_trunc(hi_elem), bf16_trunc(lo_elem)}
  return __byte_perm(__float_as_uint(lo_elem), __float_as_uint(hi_elem), 0x7632);
}
__device__ __forceinline__ uint32_t rn16x2(float lo_elem, float hi_elem) {
  __nv_bfloat162 p = __floats2bfloat162_rn(lo_elem, hi_elem);
  return *reinterpret_cast<uint32_t*>(&p);
}
__device__ __forceinline__ void split_pair(float x0, float x1, uint32_t& p0, uint32_t& p1, uint32_t& p2) {
  p0 = rn16x2(x0, x1);
  const float r0 = x0 - __uint_as_float(p0 << 16), r1 = x1 - __uint_as_float(p0 & 0xffff0000u);
  p1 = hi16x2(r0, r1);
  p2 = rn16x2(r0 - __uint_as_float(__float_as_uint(r0) & 0xffff0000u), r1 - __uint_as_float(__float_as_uint(r1) & 0xffff0000u));
}
// the NP (1 or 3) bf16 images of 8 consecutive fp32 values -> one 16-byte chunk per image
template <int NP>
__device__ __forceinline__ void store_pieces(uint8_t* img, uint32_t piece_bytes, uint32_t off, float4 v0, float4 v1) {
  if (NP == 1) {
    *reinterpret_cast<uint4*>(img + off) = pack_bf16x8(v0, v1);
  } else {
    uint4 q0, q1, q2;
    split_pair(v0.x, v0.y, q0.x, q1.x, q2.x);
    split_pair(v0.z, v0.w, q0.y, q1.y, q2.y);
    split_pair(v1.x, v1.y, q0.z, q1.z, q2.z);
    split_pair(v1.z, v1.w, q0.w, q1.w, q2.w);
    *reinterpret_cast<uint4*>(img + off) = q0;
    *reinterpret_cast<uint4*>(img + piece_bytes + off) = q1;
    *reinterpret_cast<uint4*>(img + 2 * piece_bytes + off) = q2;
  }
}
// Instruction descriptor, kind::f16 with bf16 operands, fp32 accumulate
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N, int a_major, int b_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_major << 15) | ((uint32_t)b_major << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

template <int MODE>      // 0: fp32 parity (three bf16 pieces per operand, six products), 1: bf16 operands
__global__ void __launch_bounds__(WW_THREADS, 1) k_wgrad_ws(const WgradWsParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* img_ring = smem;
  uint8_t* raw_ring = img_ring + p.nimg * p.img_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(raw_ring + p.nraw * p.raw_bytes);
  uint64_t* raw_full = bars;                       // [6] TMA -> producers
  uint64_t* raw_empty = bars + WW_MAX_RAW;         // [6] producers -> loader
  uint64_t* img_full = bars + 2 * WW_MAX_RAW;      // [4] producers -> MMA
  uint64_t* img_empty = img_full + WW_MAX_IMG;     // [4] MMA -> producers
  uint64_t* acc_full = img_empty + WW_MAX_IMG;     // [1] MMA -> epilogue
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);
  const int tid = threadIdx.x, warp = warp_index_uniform(), lane = tid & 31;
  const int split = blockIdx.x;
  const int n0 = blockIdx.y * p.npc;
  const int npc_here = min(p.npc, p.Nout - n0);
  const int nb = npc_here >> 7;                                   // 128-column dY blocks of this CTA
  const int m_begin = split * p.rows_per_split;
  const int m_end = min(p.M, m_begin + p.rows_per_split);
  const int nchunks = (m_end > m_begin) ? (m_end - m_begin + WW_R - 1) / WW_R : 0;
  const int nside = p.r + p.has_bias;

  if (warp == WW_MMA_WARP) tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
  if (tid == 0) {
    for (int i = 0; i < WW_MAX_RAW; ++i) {
      mbar_init(&raw_full[i], 1);
      mbar_init(&raw_empty[i], WW_PROD_WARPS);
    }
    for (int i = 0; i < WW_MAX_IMG; ++i) {
      mbar_init(&img_full[i], WW_PROD_WARPS);
      mbar_init(&img_empty[i], 1);
    }
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  // raw stage layout (floats): [dY: R x npc][X0: R x kx0][X1][X2][side: R x lds]
  const int raw_x0 = WW_R * p.npc;
  const int raw_side = raw_x0 + WW_R * p.KB;
  // image stage layout: NP bf16 images of A (dY), then NP bf16 images of B ([X | side])
  constexpr int NP = (MODE == 0) ? 3 : 1;
  const int a_bytes = WW_R * p.npc * 2;                          // one A image
  const int b_bytes = WW_R * ((p.KBS + 63) & ~63) * 2;           // one B image

  if (warp == WW_LOAD_WARP) {
    // ============================================================================= loader: bulk copies of raw fp32 rows
    int i = 0;
    uint32_t ph = 0;
    const uint32_t row_bytes = (uint32_t)(npc_here + p.KB + (p.r > 0 ? p.lds : 0)) * 4u;
#pragma unroll 1
    for (int c = 0; c < nchunks; ++c) {
      if (c >= p.nraw) mbar_wait_backoff(&raw_empty[i], ph ^ 1);
      const int m0 = m_begin + c * WW_R;
      const int nrows = min(WW_R, m_end - m0);
      float* rs = reinterpret_cast<float*>(raw_ring + i * p.raw_bytes);
      if (lane == 0) mbar_expect_tx(&raw_full[i], (uint32_t)nrows * row_bytes);
      __syncwarp();
      auto copy_rows = [&](float* dst, const float* src, int ld, int width, int pitch) {
        if (ld == width && pitch == width) {
          if (lane == 0) bulk_g2s(dst, src, (uint32_t)(nrows * width) * 4u, &raw_full[i]);
        } else if (lane < nrows) {
          bulk_g2s(dst + lane * pitch, src + (size_t)lane * ld, (uint32_t)width * 4u, &raw_full[i]);
        }
      };
      copy_rows(rs, p.dY + (size_t)m0 * p.lddy + n0, p.lddy, npc_here, p.npc);
      int off = raw_x0;
      for (int s = 0; s < p.nseg; ++s) {
        copy_rows(rs + off, p.X[s] + (size_t)m0 * p.ldx[s], p.ldx[s], p.kx[s], p.kx[s]);
        off += WW_R * p.kx[s];
      }
      if (p.r > 0) copy_rows(rs + raw_side, p.side + (size_t)m0 * p.lds, p.lds, p.lds, p.lds);
      if (++i == p.nraw) {
        i = 0;
        ph ^= 1;
      }
    }
  } else if (warp == WW_MMA_WARP) {
    // ============================================================================= MMA warp
    const uint32_t tm = __shfl_sync(0xffffffffu, tmem, 0);
    const bool leader = elect_one();
    const int w0 = p.KBS < 256 ? p.KBS : 256, w1 = p.KBS - w0;          // B operand pieces (N <= 256 per MMA)
    const uint32_t id0 = umma_idesc_bf16(128, w0, 1, 1);
    const uint32_t id1 = umma_idesc_bf16(128, w1 ? w1 : 16, 1, 1);
    constexpr uint32_t LBO = WW_R * 128;
    // products (A piece, B piece) in issue order
    constexpr int NPROD = (MODE == 0) ? 6 : 1;
    constexpr int PA[6] = {0, 0, 1, 1, 0, 2}, PB[6] = {0, 1, 0, 1, 2, 0};
    int j = 0;
    uint32_t ph = 0;
#pragma unroll 1
    for (int c = 0; c < nchunks; ++c) {
      mbar_wait_backoff(&img_full[j], ph);
      tc_fence_after();
      const uint32_t a_img = smem_u32(img_ring + j * p.img_bytes), b_img = a_img + NP * a_bytes;
      // (product loop outside, dY-block loop inside: with the loops nested the other way round nvcc 12.9 emitted a
      // uniform-predicate conversion of the accumulate flag that dropped accumulations -- caught by tests/test_wgrad_ws_gpu.py)
#pragma unroll
      for (int q = 0; q < NPROD; ++q) {
        const uint32_t acc = (c | q) ? 1u : 0u;
        const uint32_t ap = a_img + PA[q] * a_bytes, bp = b_img + PB[q] * b_bytes;
        for (int jb = 0; jb < nb; ++jb) {
          const uint32_t d = tm + (uint32_t)(jb * p.KBS);
          const uint64_t da = umma_desc(ap + (uint32_t)jb * (2 * LBO), LBO, 1024, 2);
          const uint64_t db0 = umma_desc(bp, LBO, 1024, 2);
          if (leader) umma_bf16(d, da, db0, id0, acc);
          if (w1) {
            const uint64_t db1 = umma_desc(bp + 4 * LBO, LBO, 1024, 2);
            if (leader) umma_bf16(d + 256, da, db1, id1, acc);
          }
        }
      }
      if (leader) {
        umma_commit(&img_empty[j]);
        if (c == nchunks - 1) umma_commit(acc_full);
      }
      __syncwarp();
      if (++j == p.nimg) {
        j = 0;
        ph ^= 1;
      }
    }
  } else {
    // ============================================================================= producers, then epilogue
    const int pt = tid;                                  // 0 .. WW_PROD_THREADS - 1
    int i = 0, j = 0;
    uint32_t phr = 0, phi = 0;
    const int ash8 = (npc_here == 256) ? 5 : 4;          // 16-byte bf16 chunks (8 columns) per dY row: 32 or 16
#pragma unroll 1
    for (int c = 0; c < nchunks; ++c) {
      const int m0 = m_begin + c * WW_R;
      const int nrows = min(WW_R, m_end - m0);
      mbar_wait_all(&raw_full[i], phr);
      if (c >= p.nimg) mbar_wait_backoff(&img_empty[j], phi ^ 1);
      const float* rs = reinterpret_cast<const float*>(raw_ring + i * p.raw_bytes);
      uint8_t* a_img = img_ring + j * p.img_bytes;
      uint8_t* b_img = a_img + NP * a_bytes;
      for (int idx = pt; idx < (WW_R << ash8); idx += WW_PROD_THREADS) {
        const int m = idx >> ash8, c8 = idx & ((1 << ash8) - 1);
        float4 v0 = zero4(), v1 = zero4();
        if (m < nrows) {
          v0 = *reinterpret_cast<const float4*>(rs + m * p.npc + 8 * c8);
          v1 = *reinterpret_cast<const float4*>(rs + m * p.npc + 8 * c8 + 4);
        }
        store_pieces<NP>(a_img, a_bytes, ww_off16(m, c8), v0, v1);
      }
      int off = raw_x0, cb8 = 0;
      for (int s = 0; s < p.nseg; ++s) {
        const int k8 = p.kx[s] >> 3;
        const bool sw = p.xsw[s] != 0;
        const int sh = (k8 & (k8 - 1)) ? -1 : 31 - __clz(k8);          // power-of-two widths (all of the models'): shifts
        for (int idx = pt; idx < WW_R * k8; idx += WW_PROD_THREADS) {
          const int m = sh >= 0 ? (idx >> sh) : idx / k8, c8 = idx - m * k8;
          float4 v0 = zero4(), v1 = zero4();
          if (m < nrows) {
            v0 = *reinterpret_cast<const float4*>(rs + off + m * p.kx[s] + 8 * c8);
            v1 = *reinterpret_cast<const float4*>(rs + off + m * p.kx[s] + 8 * c8 + 4);
          }
          if (sw) {
            v0 = swish4(v0);
            v1 = swish4(v1);
          }
          store_pieces<NP>(b_img, b_bytes, ww_off16(m, cb8 + c8), v0, v1);
        }
        off += WW_R * p.kx[s];
        cb8 += k8;
      }
      // [side | 1 | 0...] -> the last 32 columns of the B images (the last two warps: they have the fewest units above)
      if (nside && pt >= WW_PROD_THREADS - WW_R * 4) {
        const int q4 = pt - (WW_PROD_THREADS - WW_R * 4);
        const int m = q4 >> 2, c8 = q4 & 3;
        float e[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int col = 8 * c8 + q;
          e[q] = (m >= nrows) ? 0.f : (col < p.r) ? rs[raw_side + m * p.lds + p.side_c0 + col]
                                     : (col == p.r && p.has_bias) ? 1.0f : 0.f;
        }
        store_pieces<NP>(b_img, b_bytes, ww_off16(m, cb8 + c8), make_float4(e[0], e[1], e[2], e[3]),
                         make_float4(e[4], e[5], e[6], e[7]));
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&img_full[j]);
        mbar_arrive(&raw_empty[i]);
      }
      if (++i == p.nraw) {
        i = 0;
        phr ^= 1;
      }
      if (++j == p.nimg) {
        j = 0;
        phi ^= 1;
      }
    }
    // ---- epilogue: TMEM lanes = n (32 per warp quadrant), columns = k; 32-column groups are dealt to the four
    // warps of a quadrant; every store instruction writes one 128-byte line of the partial
    if (nchunks > 0) {
      mbar_wait_backoff(acc_full, 0);
      tc_fence_after();
    }
    const int quad = warp & 3, part = warp >> 2;                  // 4 warps per TMEM lane quadrant
    float* out = p.part + (size_t)split * p.KB * p.Nout;
    float* out_side = p.part_side + (size_t)split * nside * p.Nout;
    const int ngroups = p.KBS >> 5;
#pragma unroll 1
    for (int jb = 0; jb < nb; ++jb) {
      const int n = n0 + jb * 128 + 32 * quad + lane;
#pragma unroll 1
      for (int g = part; g < ngroups; g += WW_PROD_WARPS / 4) {
        float v[32];
        if (nchunks > 0) {
          __syncwarp();
          tmem_ld32(tmem + ((uint32_t)(32 * quad) << 16) + (uint32_t)(jb * p.KBS + 32 * g), v);
        } else {
#pragma unroll
          for (int q = 0; q < 32; ++q) v[q] = 0.f;
        }
        const int k0 = 32 * g;
        if (k0 < p.KB) {
#pragma unroll
          for (int q = 0; q < 32; ++q) out[(size_t)(k0 + q) * p.Nout + n] = v[q];
        } else {
#pragma unroll
          for (int q = 0; q < 8; ++q)
            if (q < nside) out_side[(size_t)q * p.Nout + n] = v[q];
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == WW_MMA_WARP) tmem_dealloc(tmem, (uint32_t)p.tmem_cols);
}

// =====================================================================================================================
// k_wgrad_ts: fp32-parity weight gradients for products with at most 160 operand columns (the LEM maps on T x N rows, the
// edge dW2 / node dW4 products), dY^T in TENSOR MEMORY.
//
// Both k_wgrad_tc and k_wgrad_ws read two MN-major operands from shared memory and are bound by that fetch (~45 B/clk; see
// above).  Here the dY block is the A operand and lives in tensor memory: a converter thread owns an output column n (its
// TMEM lane), reads dY[m][n] of the chunk's 32 rows from the raw stage (stride-one across the warp), splits it into tf32
// hi + exact lo and writes both with tcgen05.st -- the transposition costs nothing and the tensor pipe fetches only the X
// operand (MN-major SWIZZLE_128B_BASE32B hi | lo images, written by eight more converter warps) from shared memory.
// 3xTF32: (hi, hi), (lo, hi), (hi, lo), four K = 8 steps per 32-row chunk.
//
// ACCUMULATION.  The tensor core's fp32 accumulation truncates: over a long chain the error of D'[n][k] grows like
// rows^1.5 (measured 3.2e-8 rows^1.5 per CTA on unit-variance data, scripts/wgrad_error_growth.py: 4.6e-4 of max|ref| for a
// 3.3 Mi-row LEM product on 49 CTAs with k_wgrad_tc and with the first version of this kernel, where an FFMA GEMM has 5e-6).
// So the MMA accumulator only ever holds one PERIOD of `flush` chunks (8 x 32 rows): four drain warps then add it with
// round-to-nearest fp32 adds into a second, outer accumulator (also in tensor memory: tcgen05.ld both, add, tcgen05.st) and
// the MMA warp starts the next period from zero.  The converters never wait for this; the MMA warp waits for the read-out.
// Error with 8 chunks per period: 2e-6 of max|ref| at every row count.
// Side / bias gradients ([side | 1]^T dY) do not go through the tensor pipe at all: the dY converter thread of column n has
// every dY[m][n] in a register and accumulates them itself.
// The partials of the S row ranges are summed in split order exactly as for k_wgrad_ws (same layout, same callers).
constexpr int WT_R = 32;                                   // rows per chunk
constexpr int WT_DY_WARPS = 4, WT_X_WARPS = 8;
constexpr int WT_CV_WARPS = WT_DY_WARPS + WT_X_WARPS;      // 12; warp 12 = MMA, warp 13 = loader, warps 14..17 = drain
constexpr int WT_THREADS2 = 32 * (WT_CV_WARPS + 2 + 4);
constexpr int WT_MAX_TA = 4;                               // dY operand stages in tensor memory (64 columns each)
constexpr int WT_MAX_KBS = 160;                            // operand columns (no side block: see above)

// byte offset of (row m in 0..31, 16-byte chunk c4) of a [32 x 32 nblk] MN-major (BASE32B) image: column block c4 >> 3,
// row m, and inside the 128-byte row the 32-byte unit index is XORed with (m & 3)   (same layout as wgrad_tc.cu)
__device__ __forceinline__ uint32_t wt_mn_off(int m, int c4) {
  const int c = c4 & 7;
  const int cs = ((((c >> 1) ^ m) & 3) << 1) | (c & 1);
  return (uint32_t)((c4 >> 3) * 4096 + m * 128 + cs * 16);
}

__global__ void __launch_bounds__(WT_THREADS2, 1) k_wgrad_ts(const WgradWsParams p, const __grid_constant__ CUtensorMap tm_dy,
                                                           const int dy_tma, const int flush) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* img_ring = smem;                                        // nimg stages of [X hi | X lo]
  uint8_t* raw_ring = img_ring + p.nimg * p.img_bytes;             // nraw stages of raw fp32 rows
  uint64_t* bars = reinterpret_cast<uint64_t*>(raw_ring + p.nraw * p.raw_bytes);
  uint64_t* raw_full = bars;                       // [6] TMA -> converters
  uint64_t* raw_empty = bars + WW_MAX_RAW;         // [6] converters -> loader
  uint64_t* img_full = bars + 2 * WW_MAX_RAW;      // [4] X converters -> MMA
  uint64_t* img_empty = img_full + WW_MAX_IMG;     // [4] MMA -> X converters
  uint64_t* a_full = img_empty + WW_MAX_IMG;       // [4] dY converters -> MMA
  uint64_t* a_empty = a_full + WT_MAX_TA;          // [4] MMA -> dY converters
  uint64_t* acc_full = a_empty + WT_MAX_TA;        // [1] MMA -> drain warps (the accumulator of a period is complete)
  uint64_t* acc_drained = acc_full + 1;            // [1] drain warps -> MMA (added to the outer accumulator)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_drained + 1);
  const int tid = threadIdx.x, warp = warp_index_uniform(), lane = tid & 31;
  const int split = blockIdx.x;
  const int n0 = blockIdx.y * 128;
  const int m_begin = split * p.rows_per_split;
  const int m_end = min(p.M, m_begin + p.rows_per_split);
  const int nchunks = (m_end > m_begin) ? (m_end - m_begin + WT_R - 1) / WT_R : 0;
  const int nper = (nchunks + flush - 1) / flush;                 // accumulator periods
  const int nside = p.r + p.has_bias;
  const int half_bytes = p.img_bytes >> 1;                         // one X image (hi or lo): KB / 32 column blocks of 4 KiB
  // tensor memory: MMA accumulator at column 0, outer accumulator at KP, dY operand stages from 2 KP on
  const uint32_t KP = p.KB <= 128 ? 128u : 160u;
  const uint32_t acol = 2u * KP;
  const int nta = (512 - (int)acol) / 64 < WT_MAX_TA ? (512 - (int)acol) / 64 : WT_MAX_TA;

  if (warp == WT_CV_WARPS) tmem_alloc(tmem_slot, 512);
  if (tid == 0) {
    for (int i = 0; i < WW_MAX_RAW; ++i) {
      mbar_init(&raw_full[i], 1);
      mbar_init(&raw_empty[i], WT_CV_WARPS);
    }
    for (int i = 0; i < WW_MAX_IMG; ++i) {
      mbar_init(&img_full[i], WT_X_WARPS);
      mbar_init(&img_empty[i], 1);
    }
    for (int i = 0; i < WT_MAX_TA; ++i) {
      mbar_init(&a_full[i], WT_DY_WARPS);
      mbar_init(&a_empty[i], 1);
    }
    mbar_init(acc_full, 1);
    mbar_init(acc_drained, 4);
    fence_barrier_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  // raw stage layout (floats): [dY: R x 128][X0: R x kx0][X1][X2][side: R x lds]
  const int raw_x0 = WT_R * 128;
  const int raw_side = raw_x0 + WT_R * p.KB;

  if (warp == WT_CV_WARPS + 1) {
    // ============================================================================= loader: bulk copies of raw fp32 rows
    int i = 0;
    uint32_t ph = 0;
    const uint32_t row_bytes = (uint32_t)(128 + p.KB + (p.r > 0 ? p.lds : 0)) * 4u;
#pragma unroll 1
    for (int c = 0; c < nchunks; ++c) {
      if (c >= p.nraw) mbar_wait_backoff(&raw_empty[i], ph ^ 1);
      const int m0 = m_begin + c * WT_R;
      const int nrows = min(WT_R, m_end - m0);
      float* rs = reinterpret_cast<float*>(raw_ring + i * p.raw_bytes);
      // a strided dY block (Nout > 128) arrives by ONE tensor-map copy of the [32 x 128] box (rows past M are zero filled, the
      // box always counts 16 KiB) instead of 32 row copies of 512 bytes
      if (lane == 0)
        mbar_expect_tx(&raw_full[i], dy_tma ? (uint32_t)(WT_R * 128 * 4) + (uint32_t)nrows * (row_bytes - 512u) : (uint32_t)nrows * row_bytes);
      __syncwarp();
      auto copy_rows = [&](float* dst, const float* src, int ld, int width) {
        if (ld == width) {
          if (lane == 0) bulk_g2s(dst, src, (uint32_t)(nrows * width) * 4u, &raw_full[i]);
        } else if (lane < nrows) {
          bulk_g2s(dst + lane * width, src + (size_t)lane * ld, (uint32_t)width * 4u, &raw_full[i]);
        }
      };
      if (dy_tma) {
        if (lane == 0)
          asm volatile(
              "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                  smem_u32(rs)),
              "l"(reinterpret_cast<uint64_t>(&tm_dy)), "r"(n0), "r"(m0), "r"(smem_u32(&raw_full[i]))
              : "memory");
      } else {
        copy_rows(rs, p.dY + (size_t)m0 * p.lddy + n0, p.lddy, 128);
      }
      int off = raw_x0;
      for (int s = 0; s < p.nseg; ++s) {
        copy_rows(rs + off, p.X[s] + (size_t)m0 * p.ldx[s], p.ldx[s], p.kx[s]);
        off += WT_R * p.kx[s];
      }
      if (p.r > 0) copy_rows(rs + raw_side, p.side + (size_t)m0 * p.lds, p.lds, p.lds);
      if (++i == p.nraw) {
        i = 0;
        ph ^= 1;
      }
    }
  } else if (warp == WT_CV_WARPS) {
    // ============================================================================= MMA warp
    const uint32_t tm = __shfl_sync(0xffffffffu, tmem, 0);
    const bool leader = elect_one();
    const uint32_t idesc = umma_idesc_tf32(128, p.KB, 0, 1);           // A from tensor memory, B MN-major
    int j = 0, ts = 0, cf = 0, per = 0;
    uint32_t phj = 0, tph = 0;
#pragma unroll 1
    for (int c = 0; c < nchunks; ++c) {
      if (cf == 0 && per > 0) mbar_wait_backoff(acc_drained, (uint32_t)(per - 1) & 1);      // previous period read out
      mbar_wait_backoff(&img_full[j], phj);
      mbar_wait_backoff(&a_full[ts], tph);
      tc_fence_after();
      const uint32_t xh = smem_u32(img_ring + j * p.img_bytes), xl = xh + (uint32_t)half_bytes;
      const uint32_t a_hi = tm + acol + 64u * (uint32_t)ts, a_lo = a_hi + 32;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t ko = 1024 * k;        // 8 rows per k-step; LBO = 4096 (next 32 columns), SBO = 512 (next 4 rows)
        const uint64_t dxh = umma_desc(xh + ko, 4096, 512, 1), dxl = umma_desc(xl + ko, 4096, 512, 1);
        if (leader) {
          umma_tf32_ts(tm, a_hi + 8 * k, dxh, idesc, (cf | k) ? 1u : 0u);
          umma_tf32_ts(tm, a_lo + 8 * k, dxh, idesc, 1u);
          umma_tf32_ts(tm, a_hi + 8 * k, dxl, idesc, 1u);
        }
      }
      const bool period_end = cf == flush - 1 || c == nchunks - 1;
      if (leader) {
        umma_commit(&img_empty[j]);
        umma_commit(&a_empty[ts]);
        if (period_end) umma_commit(acc_full);
      }
      __syncwarp();
      if (period_end) {
        cf = 0;
        ++per;
      } else {
        ++cf;
      }
      if (++j == p.nimg) {
        j = 0;
        phj ^= 1;
      }
      if (++ts == nta) {
        ts = 0;
        tph ^= 1;
      }
    }
  } else if (warp > WT_CV_WARPS + 1) {
    // ============================================================================= drain warps (one per TMEM lane quadrant)
    const int quad = warp & 3;
    const uint32_t tq = tmem + ((uint32_t)(32 * quad) << 16);
    const int ng = p.KB >> 5;
#pragma unroll 1
    for (int per = 0; per < nper; ++per) {
      mbar_wait_backoff(acc_full, (uint32_t)per & 1);
      tc_fence_after();
#pragma unroll 1
      for (int g = 0; g < ng; ++g) {
        float v[32], o[32];
        __syncwarp();
        tmem_ld32(tq + (uint32_t)(32 * g), v);
        if (per > 0) {
          tmem_ld32(tq + KP + (uint32_t)(32 * g), o);
#pragma unroll
          for (int q = 0; q < 32; ++q) v[q] += o[q];
        }
#pragma unroll
        for (int q = 0; q < 32; q += 8) {
          float w8[8];
#pragma unroll
          for (int t = 0; t < 8; ++t) w8[t] = v[q + t];
          tmem_st8(tq + KP + (uint32_t)(32 * g + q), w8);
        }
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_drained);
    }
    // the CTA's partial: every store instruction writes one 128-byte line (lanes = consecutive n)
    float* out = p.part + (size_t)split * p.KB * p.Nout;
    const int n = n0 + 32 * quad + lane;
#pragma unroll 1
    for (int g = 0; g < ng; ++g) {
      float v[32];
      if (nper > 0) {
        __syncwarp();
        tmem_ld32(tq + KP + (uint32_t)(32 * g), v);
      } else {
#pragma unroll
        for (int q = 0; q < 32; ++q) v[q] = 0.f;
      }
      if (n < p.Nout) {
#pragma unroll
        for (int q = 0; q < 32; ++q) out[(size_t)(32 * g + q) * p.Nout + n] = v[q];
      }
    }
  } else {
    // ============================================================================= converters
    int i = 0;
    uint32_t phr = 0;
    if (warp < WT_DY_WARPS) {
      // ---- dY^T -> tensor memory: thread = output column n (TMEM lane), 32 rows of the chunk = 32 k-columns (hi | lo);
      // the same thread accumulates the side / bias gradients of its column
      const int nl = 32 * warp + lane;
      const uint32_t lane_off = (uint32_t)(32 * warp) << 16;
      const bool ncol_ok = n0 + nl < p.Nout;
      float sacc[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) sacc[q] = 0.f;
      const bool bias_only = p.r == 0 && p.has_bias != 0;
      int ts = 0;
      uint32_t tph = 0;
#pragma unroll 1
      for (int c = 0; c < nchunks; ++c) {
        const int nrows = min(WT_R, m_end - (m_begin + c * WT_R));
        mbar_wait_all(&raw_full[i], phr);
        const float* rs = reinterpret_cast<const float*>(raw_ring + i * p.raw_bytes);
        if (c >= nta) mbar_wait_backoff(&a_empty[ts], tph ^ 1);
        tc_fence_after();
        const uint32_t col = tmem + lane_off + acol + 64u * (uint32_t)ts;
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          float hi[8], lo[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const int m = 8 * h + q;
            const float x = (m < nrows && ncol_ok) ? rs[m * 128 + nl] : 0.f;
            split_tf32(x, hi[q], lo[q]);
            if (bias_only) {                                 // (every product this kernel takes in the models: one add)
              sacc[0] += x;
            } else if (nside) {
#pragma unroll 1
              for (int sq = 0; sq < nside; ++sq) {
                const float sv = (sq < p.r) ? rs[raw_side + m * p.lds + p.side_c0 + sq] : 1.0f;      // warp-uniform address
                sacc[sq] = fmaf(sv, x, sacc[sq]);
              }
            }
          }
          tmem_st8(col + 8 * h, hi);
          tmem_st8(col + 32 + 8 * h, lo);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&a_full[ts]);
          mbar_arrive(&raw_empty[i]);
        }
        if (++i == p.nraw) {
          i = 0;
          phr ^= 1;
        }
        if (++ts == nta) {
          ts = 0;
          tph ^= 1;
        }
      }
      if (nside && ncol_ok) {
        float* out_side = p.part_side + (size_t)split * nside * p.Nout;
#pragma unroll
        for (int q = 0; q < 8; ++q)
          if (q < nside) out_side[(size_t)q * p.Nout + n0 + nl] = sacc[q];
      }
    } else {
      // ---- X -> MN-major hi | lo images
      const int xt = tid - 32 * WT_DY_WARPS;               // 0 .. 255
      const int kb4 = p.KB >> 2;                           // float4 units per row
      const int kx0 = p.kx[0], kx01 = kx0 + (p.nseg > 1 ? p.kx[1] : 0);
      int j = 0;
      uint32_t phj = 0;
#pragma unroll 1
      for (int c = 0; c < nchunks; ++c) {
        const int nrows = min(WT_R, m_end - (m_begin + c * WT_R));
        mbar_wait_all(&raw_full[i], phr);
        if (c >= p.nimg) mbar_wait_backoff(&img_empty[j], phj ^ 1);
        const float* rs = reinterpret_cast<const float*>(raw_ring + i * p.raw_bytes);
        uint8_t* hi_img = img_ring + j * p.img_bytes;
        uint8_t* lo_img = hi_img + half_bytes;
        // thread -> row m = xt / 8 of the chunk and the 16-byte units c4 = xt % 8 + 8 i (one column block per i): all loads
        // of a chunk are issued before the first split / store (a chain of dependent load -> split -> store steps made the
        // conversion latency, not the MMAs, set the pace)
        {
          const int m = xt >> 3, cl = xt & 7;
          const int nit = kb4 >> 3;                          // column blocks: KB / 32 <= 5
          float4 v[5];
#pragma unroll
          for (int it = 0; it < 5; ++it) {
            v[it] = zero4();
            const int col = 4 * (cl + 8 * it);
            if (it < nit && m < nrows) {
              if (col < kx0) {
                v[it] = *reinterpret_cast<const float4*>(rs + raw_x0 + m * kx0 + col);
                if (p.xsw[0]) v[it] = swish4(v[it]);
              } else if (col < kx01) {
                v[it] = *reinterpret_cast<const float4*>(rs + raw_x0 + WT_R * kx0 + m * p.kx[1] + (col - kx0));
                if (p.xsw[1]) v[it] = swish4(v[it]);
              } else {
                v[it] = *reinterpret_cast<const float4*>(rs + raw_x0 + WT_R * kx01 + m * p.kx[2] + (col - kx01));
                if (p.xsw[2]) v[it] = swish4(v[it]);
              }
            }
          }
#pragma unroll
          for (int it = 0; it < 5; ++it)
            if (it < nit) store_split4(hi_img, lo_img, wt_mn_off(m, cl + 8 * it), v[it]);
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&img_full[j]);
          mbar_arrive(&raw_empty[i]);
        }
        if (++i == p.nraw) {
          i = 0;
          phr ^= 1;
        }
        if (++j == p.nimg) {
          j = 0;
          phj ^= 1;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == WT_CV_WARPS) tmem_dealloc(tmem, 512);
}

bool tensor_map_2d(CUtensorMap* tm, const float* A, int ld, int cols, int rows, int box_cols, int box_rows, bool swizzle128);      // linear_tma.cu

static int ww_env(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

// geometry of one call: dY columns per CTA, ring depths, splits.  Returns false when the shape is not supported.
static bool ww_plan(int M, int KB, int Nout, int nside, int lds, int mode, WgradWsParams& p, int& S, int& ny, int& smem,
                    bool* use_ts = nullptr) {
  if (KB <= 0 || (KB & 31) || Nout <= 0 || (Nout & 127) || nside < 0 || nside > 8) return false;
  const int KBS = KB + (nside ? 32 : 0);
  if (KBS > 512) return false;
  const int nbtot = Nout >> 7;
  // products with at most 160 operand columns (the LEM maps, dW2 / dW4): k_wgrad_ts in fp32-parity mode.  The row split is
  // computed for its grid (one 128-column dY block per CTA) in BOTH modes, so that msmp_wgrad_ws_splits -- which sizes the
  // callers' partial buffers -- does not depend on the mode.
  static const bool ts_on = ww_env("MSMP_WGRAD_TS", 1) != 0;
  // (k_wgrad_ts keeps the side / bias sums out of the MMAs: one add per element for a bias, a slow loop for side columns, so
  // products with side columns -- the P | Q projection of the one-field models, 3.6 against 1.3 ms at 1 Mi rows -- stay on k_wgrad_ws)
  const bool ts_shape = ts_on && KB <= WT_MAX_KBS && nside <= 1;
  const bool ts = ts_shape && mode == 0;
  if (use_ts) *use_ts = ts;
  const int nb = ts ? 1 : ((nbtot >= 2 && 2 * KBS <= 512) ? 2 : 1);
  p.KB = KB;
  p.KBS = KBS;
  p.npc = nb * 128;
  ny = (Nout + p.npc - 1) / p.npc;
  int cols = 32;
  while (cols < nb * KBS) cols <<= 1;
  p.tmem_cols = cols;
  const int R = ts ? WT_R : WW_R;
  const int raw = R * (p.npc + KB + lds) * 4;      // lds = 0 when no side columns are read
  p.raw_bytes = (raw + 127) & ~127;
  p.img_bytes = ts ? KB * 256 : (mode == 0 ? 3 : 1) * WW_R * (p.npc + ((KBS + 63) & ~63)) * 2;
  const int budget = WW_SMEM_LIMIT - 1024 - 512;
  int nraw = 2, nimg = 2;
  if (nraw * p.raw_bytes + nimg * p.img_bytes > budget) return false;
  if (ts) {
    // raw stages (bytes in flight) first, then image stages
    while (nraw < WW_MAX_RAW && (nraw + 1) * p.raw_bytes + nimg * p.img_bytes <= budget) ++nraw;
    while (nimg < WW_MAX_IMG && nraw * p.raw_bytes + (nimg + 1) * p.img_bytes <= budget) ++nimg;
  } else {
    // deeper rings while they fit: first a third image stage, then raw stages (loads in flight)
    if (nraw * p.raw_bytes + (nimg + 1) * p.img_bytes <= budget) ++nimg;
    while (nraw < WW_MAX_RAW && (nraw + 1) * p.raw_bytes + nimg * p.img_bytes <= budget) ++nraw;
    if (nimg < WW_MAX_IMG && nraw == WW_MAX_RAW && nraw * p.raw_bytes + (nimg + 1) * p.img_bytes <= budget) ++nimg;
  }
  p.nraw = nraw;
  p.nimg = nimg;
  smem = 1024 + nraw * p.raw_bytes + nimg * p.img_bytes + 512;
  static const int sms = [] { int d = 0, n = 148; cudaGetDevice(&d); cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, d); return n > 0 ? n : 148; }();
  static const int min_rows = ww_env("MSMP_WGRAD_WS_MIN_ROWS", 256);
  const int ny_split = ts_shape ? nbtot : ny;
  const int rq = ts_shape ? WT_R : WW_R;              // row quantum of a split
  int max_s = sms / ny_split;
  if (max_s < 1) max_s = 1;
  int s = M / (min_rows > rq ? min_rows : rq);
  if (s < 1) s = 1;
  if (s > max_s) s = max_s;
  // k_wgrad_ws keeps a CTA's whole row range in the tensor-core accumulator, whose truncating adds make the error grow like
  // rows^1.5 (see k_wgrad_ts): more, shorter row ranges than SMs from 1 Ki rows per CTA on (1 Mi-row products: 1024 CTAs)
  static const int max_rows = ww_env("MSMP_WGRAD_WS_MAX_ROWS", 1024);
  if (!ts_shape && max_rows > 0 && (M + s - 1) / s > max_rows) s = (M + max_rows - 1) / max_rows;
  int rps = (M + s - 1) / s;
  rps = (rps + rq - 1) / rq * rq;
  if (rps < rq) rps = rq;
  p.rows_per_split = rps;
  S = M > 0 ? (M + rps - 1) / rps : 1;
  p.M = M;
  return true;
}

}  // namespace msmp

using namespace msmp;

extern "C" int msmp_wgrad_ws_splits(int M, int KB, int Nout, int nside) {
  WgradWsParams p{};
  int S = 0, ny = 0, smem = 0;
  if (!ww_plan(M, KB, Nout, nside, 8, 0, p, S, ny, smem)) return -1;
  return S;
}

extern "C" size_t msmp_wgrad_ws_workspace(int M, int KB, int Nout, int nside) {
  const int S = msmp_wgrad_ws_splits(M, KB, Nout, nside);
  if (S < 0) return 0;
  return (size_t)S * ((size_t)KB + nside) * Nout * sizeof(float);
}

extern "C" int msmp_wgrad_ws(const float* const* X, const int* ldx, const int* kx, const int* xswish, int nseg,
                             const float* dY, int lddy, int Nout, const float* side, int lds, int side_c0, int r,
                             int has_bias, float* part, float* part_side, float* dWt, float* dWside, int accumulate,
                             int M, int mode, cudaStream_t stream) {
  if (nseg < 1 || nseg > 3 || M < 0 || (lddy & 3) || r < 0 || (mode != 0 && mode != 1)) return MSMP_ERR_ARG;
  WgradWsParams p{};
  int KB = 0;
  for (int s = 0; s < nseg; ++s) {
    if (kx[s] <= 0 || (kx[s] & 31) || (ldx[s] & 3)) return MSMP_ERR_ARG;
    p.X[s] = X[s];
    p.ldx[s] = ldx[s];
    p.kx[s] = kx[s];
    p.xsw[s] = xswish ? xswish[s] : 0;
    KB += kx[s];
  }
  p.nseg = nseg;
  const int rr = side ? r : 0;
  const int nside = rr + (has_bias ? 1 : 0);
  if (nside && side && ((lds & 3) || lds > 16 || side_c0 < 0 || side_c0 + rr > lds)) return MSMP_ERR_ARG;
  int S = 0, ny = 0, smem = 0;
  bool use_ts = false;
  if (!ww_plan(M, KB, Nout, nside, side ? lds : 0, mode, p, S, ny, smem, &use_ts)) return MSMP_ERR_ARG;
  p.dY = dY; p.lddy = lddy; p.Nout = Nout;
  p.side = side; p.lds = side ? lds : 0; p.side_c0 = side_c0; p.r = rr; p.has_bias = has_bias ? 1 : 0;
  p.part = part; p.part_side = part_side;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(k_wgrad_ws<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, WW_SMEM_LIMIT) != cudaSuccess ||
        cudaFuncSetAttribute(k_wgrad_ws<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, WW_SMEM_LIMIT) != cudaSuccess ||
        cudaFuncSetAttribute(k_wgrad_ts, cudaFuncAttributeMaxDynamicSharedMemorySize, WW_SMEM_LIMIT) != cudaSuccess)
      return MSMP_ERR_CUDA;
    attr_set = true;
  }
  dim3 grid(S, ny);
  if (use_ts) {
    alignas(64) CUtensorMap tm;
    const int dy_tma = (lddy != 128 && tensor_map_2d(&tm, dY, lddy, Nout, M, 128, WT_R, false)) ? 1 : 0;
    if (!dy_tma) memset(&tm, 0, sizeof(tm));
    static const int flush = [] { const int f = ww_env("MSMP_WGRAD_TS_FLUSH", 8); return f < 1 ? 1 : f; }();      // chunks per accumulator period
    k_wgrad_ts<<<grid, WT_THREADS2, smem, stream>>>(p, tm, dy_tma, flush);
  }
  else if (mode == 0)
    k_wgrad_ws<0><<<grid, WW_THREADS, smem, stream>>>(p);
  else
    k_wgrad_ws<1><<<grid, WW_THREADS, smem, stream>>>(p);
  MSMP_CHECK_LAUNCH();
  if (dWt) {
    const int c0 = KB * Nout, c1 = nside * Nout;
    k_reduce_partials2<<<(c0 + c1 + 255) / 256, 256, 0, stream>>>(part, dWt, c0, (size_t)KB * Nout, part_side, dWside,
                                                                 c1, (size_t)nside * Nout, S, accumulate);
    MSMP_CHECK_LAUNCH();
  }
  return MSMP_OK;
}
