// Shared by the node-level GEMM kernels (linear_tc.cu, linear_tma.cu): parameters, the coalescing epilogue, hi-only staging.
#pragma once
#include "umma.cuh"

namespace msmp {

constexpr int TC_A_BYTES = 2 * IMG_BYTES, TC_B_BYTES = 2 * IMG_BYTES;

struct LinTcParams {
  const float* A[3];
  int lda[3];
  int ka[3];
  int aswish[3];
  int nseg;
  const float* Bimg;
  const float* bias;
  const float* side;
  int lds;
  int r;
  const float* Wside;
  int ldws;
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

// ---- epilogue of one 32-row x 32-column accumulator block --------------------------------------------------------------
// tcgen05.ld hands every thread 32 consecutive columns of ONE row; reading / writing global memory in that shape makes
// each warp instruction touch 32 different rows 16 bytes at a time (32 sectors per request, half of every written sector
// unused: ncu of the first role-split kernel).  The block therefore goes through a warp-private shared-memory tile in two
// [32 rows x 16 columns] halves (pitch 20 floats; the (row, row + 4) pairing keeps every quarter-warp access conflict
// free) so that a warp instruction covers 8 rows x 64 contiguous bytes: every sector fully used, 16 sectors per request,
// and bias / side weights are per-thread constants of a half block.  Same arithmetic per output element as before.
constexpr int EPI_TILE_FLOATS = 32 * 20;

// FAST (reduced-precision mode): operands rounded to tf32 once, ONE MMA pass (hi x hi): no lo images are written, only
// the hi half of every weight chunk is copied (16 instead of 32 KiB from L2) and 4 instead of 12 MMAs run per chunk.
// Error of a K = 256 product: ~3e-4 of max|ref| (tf32 keeps 10 mantissa bits, three more than bf16).
__device__ __forceinline__ void store_hi4(uint8_t* hi_img, uint32_t off, float4 v) {
  float4 h;
  float l;
  split_tf32(v.x, h.x, l);
  split_tf32(v.y, h.y, l);
  split_tf32(v.z, h.z, l);
  split_tf32(v.w, h.w, l);
  *reinterpret_cast<float4*>(hi_img + off) = h;
}

__device__ __forceinline__ float4 dswish4(float4 z) { return make_float4(dswish(z.x), dswish(z.y), dswish(z.z), dswish(z.w)); }

// Tile-constant epilogue operands staged in shared memory: the 128 rows' side values [128][8], the side weights of the
// tile's 128 columns [8][128] and its bias [128].  Read from global memory inside the per-row loop they were 44 % of the
// kernel's stall samples (ncu source page, round 2: every FMA of the side term waited on an L2 round trip); staged once per
// tile -- before the wait for the accumulator, so the latency hides behind the main loop -- they are short shared-memory
// reads.  All pointers null: read global memory (k_linear_ws).
struct EpiStage {
  const float* side;
  const float* ws;
  const float* bias;
};
constexpr int EPI_STAGE_FLOATS = 128 * 8 + 8 * 128 + 128;

// nthreads cooperating threads (index et) fill the stage for the tile at (row0, n0); the caller synchronises them before use
__device__ __forceinline__ void epi_stage_fill(const LinTcParams& p, float* st, int row0, int n0, int et, int nthreads) {
  float* s_side = st;
  float* s_ws = st + 1024;
  float* s_bias = st + 2048;
  if (p.r > 0) {                                                  // no side term: the epilogue never reads these two
    for (int idx = et; idx < 1024; idx += nthreads) {
      const int row = row0 + (idx >> 3), q = idx & 7;
      s_side[idx] = (q < p.r && row < p.M) ? __ldg(p.side + (size_t)row * p.lds + q) : 0.f;
    }
    for (int idx = et; idx < 128 * p.r; idx += nthreads) {
      const int q = idx >> 7, col = n0 + (idx & 127);
      s_ws[idx] = (col < p.Nout) ? __ldg(p.Wside + (size_t)q * p.ldws + col) : 0.f;
    }
  }
  for (int idx = et; idx < 128; idx += nthreads) s_bias[idx] = (p.bias && n0 + idx < p.Nout) ? __ldg(p.bias + n0 + idx) : 0.f;
}

__device__ __forceinline__ void lin_epilogue32(const LinTcParams& p, float* tb, const float (&v)[32], int row_base,
                                               int colb, int lane, const EpiStage es = EpiStage{nullptr, nullptr, nullptr},
                                               int row0 = 0, int n0 = 0) {
  const int c = lane & 3, rl0 = (lane >> 3) + 4 * ((lane >> 2) & 1);
#pragma unroll
  for (int hb = 0; hb < 32; hb += 16) {
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 4; ++j)
      st4(tb + lane * 20 + 4 * j, make_float4(v[hb + 4 * j], v[hb + 4 * j + 1], v[hb + 4 * j + 2], v[hb + 4 * j + 3]));
    __syncwarp();
    const int col = colb + hb + 4 * c;
    if (col >= p.Nout) continue;
    const float4 b4 = es.bias ? *reinterpret_cast<const float4*>(es.bias + (col - n0)) : (p.bias ? ldg4(p.bias + col) : zero4());
    float4 zm[4], rr[4];
    if (p.Zmul) {
#pragma unroll
      for (int ps = 0; ps < 4; ++ps) {
        const int row = row_base + 8 * ps + rl0;
        zm[ps] = row < p.M ? ldg4(p.Zmul + (size_t)row * p.ldz + col) : zero4();
      }
    }
    if (p.R) {
#pragma unroll
      for (int ps = 0; ps < 4; ++ps) {
        const int row = row_base + 8 * ps + rl0;
        rr[ps] = row < p.M ? ldg4(p.R + (size_t)row * p.ldr + col) : zero4();
      }
    }
#pragma unroll
    for (int ps = 0; ps < 4; ++ps) {
      const int rl = 8 * ps + rl0;
      const int row = row_base + rl;
      if (row >= p.M) continue;
      float4 z = *reinterpret_cast<const float4*>(tb + rl * 20 + 4 * c);
      if (p.bias) z = add4(z, b4);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        if (q >= p.r) break;
        float sv;
        float4 w;
        if (es.side) {
          sv = es.side[(row - row0) * 8 + q];
          w = *reinterpret_cast<const float4*>(es.ws + q * 128 + (col - n0));
        } else {
          sv = __ldg(p.side + (size_t)row * p.lds + q);
          w = ldg4(p.Wside + (size_t)q * p.ldws + col);
        }
        z.x = fmaf(sv, w.x, z.x);
        z.y = fmaf(sv, w.y, z.y);
        z.z = fmaf(sv, w.z, z.z);
        z.w = fmaf(sv, w.w, z.w);
      }
      if (p.Zmul) z = mul4(z, dswish4(zm[ps]));
      if (p.Ypre) st4(p.Ypre + (size_t)row * p.ldpre + col, z);
      if (p.act) z = swish4(z);
      if (p.R) z = add4(z, rr[ps]);
      st4(p.Y + (size_t)row * p.ldy + col, z);
    }
  }
}

}  // namespace msmp
