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
  int dbg;              // experiment switches of k_linear_tma (MSMP_LIN_DBG; 0 in production): 1 no epilogue memory traffic, 2 no conversion, 4 no MMAs, 8 no weight copies, 16 no A copies
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

// explicit shared-space accesses: through the generic pointers of EpiStage / the tile the compiler emitted generic LD / ST
// (ncu source page, round 2), and the per-row `continue`s cut the block into 8 reconvergence regions of dependent code:
// 684 instructions per 32 x 32 block at ~10 cycles each made the EPILOGUE the bound of every large node GEMM
// (MSMP_LIN_DBG ablation: 0.63 ms per layer with it, 0.29 ms without).  Now: LDS / STS, four independent row streams
// (predicated accesses instead of branches), addresses formed once per block.
__device__ __forceinline__ float4 lds4(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ float lds1(uint32_t a) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void sts4(uint32_t a, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// STAGED: side values / side weights / bias come from the EpiStage in shared memory (every kernel but k_linear_ws)
template <bool STAGED>
__device__ __forceinline__ void lin_epilogue32(const LinTcParams& p, float* tb, const float (&v)[32], int row_base,
                                               int colb, int lane, const EpiStage es = EpiStage{nullptr, nullptr, nullptr},
                                               int row0 = 0, int n0 = 0) {
  const int c = lane & 3, rl0 = (lane >> 3) + 4 * ((lane >> 2) & 1);
  const uint32_t tbs = smem_u32(tb);
  const uint32_t wr = tbs + (uint32_t)lane * 80u;                        // this lane's row of the transposition tile
  const uint32_t rd = tbs + (uint32_t)rl0 * 80u + (uint32_t)c * 16u;     // (row rl0 + 8 ps, columns 4c..4c+3): + 640 ps
  const int r0 = row_base + rl0;
  bool ok[4];
#pragma unroll
  for (int ps = 0; ps < 4; ++ps) ok[ps] = r0 + 8 * ps < p.M;
  constexpr bool staged = STAGED;
  const bool bias_staged = staged && es.bias != nullptr;      // (launches without a side term do not fill the stage)
  const uint32_t s_side = bias_staged ? smem_u32(es.side) + (uint32_t)(r0 - row0) * 32u : 0u;
  const uint32_t s_ws = bias_staged ? smem_u32(es.ws) : 0u, s_bias = bias_staged ? smem_u32(es.bias) : 0u;
  const bool has_bias = p.bias != nullptr, has_z = p.Zmul != nullptr, has_r = p.R != nullptr, has_pre = p.Ypre != nullptr;
#pragma unroll
  for (int hb = 0; hb < 32; hb += 16) {
    if (p.dbg & 128) continue;
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 4; ++j)
      sts4(wr + 16 * j, make_float4(v[hb + 4 * j], v[hb + 4 * j + 1], v[hb + 4 * j + 2], v[hb + 4 * j + 3]));
    __syncwarp();
    const int col = colb + hb + 4 * c;
    if (col >= p.Nout || (p.dbg & 1)) continue;
    float4 zm[4], rr[4], z[4];
    if (has_z) {
#pragma unroll
      for (int ps = 0; ps < 4; ++ps)
        zm[ps] = (ok[ps] && !(p.dbg & 64)) ? ldg4(p.Zmul + (size_t)(r0 + 8 * ps) * p.ldz + col) : zero4();
    }
    if (has_r) {
#pragma unroll
      for (int ps = 0; ps < 4; ++ps)
        rr[ps] = (ok[ps] && !(p.dbg & 64)) ? ldg4(p.R + (size_t)(r0 + 8 * ps) * p.ldr + col) : zero4();
    }
    const float4 b4 = bias_staged ? lds4(s_bias + (uint32_t)(col - n0) * 4u) : (has_bias ? ldg4(p.bias + col) : zero4());
#pragma unroll
    for (int ps = 0; ps < 4; ++ps) {
      z[ps] = lds4(rd + 640u * ps);
      if (has_bias) z[ps] = add4(z[ps], b4);
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      if (q >= p.r) break;
      const float4 w = staged ? lds4(s_ws + (uint32_t)(q * 128 + (col - n0)) * 4u) : ldg4(p.Wside + (size_t)q * p.ldws + col);
#pragma unroll
      for (int ps = 0; ps < 4; ++ps) {
        const float sv = staged ? lds1(s_side + (uint32_t)(64 * ps + q) * 4u)
                                : (ok[ps] ? __ldg(p.side + (size_t)(r0 + 8 * ps) * p.lds + q) : 0.f);
        z[ps].x = fmaf(sv, w.x, z[ps].x);
        z[ps].y = fmaf(sv, w.y, z[ps].y);
        z[ps].z = fmaf(sv, w.z, z[ps].z);
        z[ps].w = fmaf(sv, w.w, z[ps].w);
      }
    }
    if (has_z) {
#pragma unroll
      for (int ps = 0; ps < 4; ++ps) z[ps] = mul4(z[ps], dswish4_m(zm[ps]));
    }
    if (has_pre && !(p.dbg & 32)) {
#pragma unroll
      for (int ps = 0; ps < 4; ++ps)
        if (ok[ps]) st4(p.Ypre + (size_t)(r0 + 8 * ps) * p.ldpre + col, z[ps]);
    }
    if (p.act) {
#pragma unroll
      for (int ps = 0; ps < 4; ++ps) z[ps] = swish4_m(z[ps]);
    }
    if (has_r) {
#pragma unroll
      for (int ps = 0; ps < 4; ++ps) z[ps] = add4(z[ps], rr[ps]);
    }
#pragma unroll
    for (int ps = 0; ps < 4; ++ps)
      if (ok[ps] && (!(p.dbg & 32) || z[ps].x == 123.456f)) st4(p.Y + (size_t)(r0 + 8 * ps) * p.ldy + col, z[ps]);
  }
}

}  // namespace msmp
