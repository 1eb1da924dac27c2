// Node-level dense layers on the 5th-gen tensor cores (tcgen05.mma kind::tf32, accumulators in TMEM).
//
//   msmp_linear_tc_fwd : same contract as msmp_linear_fwd (Y = epilogue([A0|A1|A2] W^T + bias + side Wside))
//                        with the weight given as pre-split (hi | lo), pre-swizzled tile images:
//                        Bimg[ntile][kchunk][2][128 x 32 fp32]  (row n of tile, column k of chunk).
// Per CTA: one 128 x 128 output tile.  K is consumed in 32-wide chunks through a 2-stage ring:
//   weights  : one 32 KiB cp.async.bulk (TMA engine, 1-D) per chunk, completion on an mbarrier
//   A operand: threads load fp32 rows (coalesced), apply the optional swish, split into tf32 hi/lo and
//              store both in the UMMA 128B-swizzled layout (generic proxy -> fence.proxy.async)
//   MMA      : one thread issues 4 k-steps x 3 products (hi*hi, lo*hi, hi*lo) per chunk, tcgen05.commit
//              releases the stage; production of chunk c+1 overlaps the MMAs of chunk c
//   epilogue : tcgen05.ld (32 lanes x 32 columns per warp) -> bias / side / swish / residual -> global
#include <cstdlib>
#include "linear_common.cuh"
#include "msmp_b200.h"

namespace msmp {

constexpr int TC_A_STAGES = 2;                                // A_hi, A_lo per stage (32 KiB)
constexpr int TC_B_STAGES = 4;                                // B_hi, B_lo per stage (32 KiB), prefetched ahead
constexpr int TC_SMEM = TC_A_STAGES * TC_A_BYTES + TC_B_STAGES * TC_B_BYTES + 1024 /*align*/ + 256 /*barriers*/ +
                        EPI_STAGE_FLOATS * 4 /*side values, side weights, bias of the tile*/;

template <bool FAST>
__global__ void __launch_bounds__(256, 1) k_linear_tc(const LinTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smemB = smem + TC_A_STAGES * TC_A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smemB + TC_B_STAGES * TC_B_BYTES);     // bfull[4], done[2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  const int tid = threadIdx.x, warp = warp_index_uniform(), lane = tid & 31;
  const int row0 = blockIdx.x * 128;
  const int ntile = blockIdx.y;

  int ktot = 0;
  for (int s = 0; s < p.nseg; ++s) ktot += p.ka[s];
  const int nchunks = ktot >> 5;

  if (warp == 0) tmem_alloc(tmem_slot, 128);
  if (tid == 32) {
    for (int s = 0; s < 6; ++s) mbar_init(&bars[s], 1);
    fence_barrier_init();
  }
  // the epilogue's tile constants (side values, side weights, bias) -> shared memory now: their latency hides behind the
  // main loop instead of stalling every row of the epilogue
  float* epi_stage = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);
  epi_stage_fill(p, epi_stage, row0, ntile * 128, tid, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  constexpr uint32_t IDESC = umma_idesc_tf32(128, 128, 0, 0);

  // chunk c -> (segment pointer, row stride, column offset, swish flag)
  auto locate = [&](int c, const float*& A, int& lda, int& koff, bool& sw) {
    int seg = 0;
    koff = c * 32;
    while (seg < p.nseg - 1 && koff >= p.ka[seg]) {
      koff -= p.ka[seg];
      ++seg;
    }
    A = p.A[seg];
    lda = p.lda[seg];
    sw = p.aswish[seg] != 0;
  };
  // software pipeline: the fp32 rows of chunk c+1 are in flight (registers) while chunk c is split/stored and
  // its MMAs run; weights of chunk c arrive by bulk copy while the A operand is being staged.
  // Two register sets: chunks c+1 and c+2 are in flight while chunk c is staged (a single set issued the loads only a
  // few hundred cycles before they were needed).
  float4 preA[4], preB[4];
  bool swA, swB;
  auto prefetch = [&](int c, float4 (&pre)[4], bool& pre_sw) {
    const float* A;
    int lda, koff;
    locate(c, A, lda, koff, pre_sw);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = tid + 256 * i;
      const int grow = row0 + (idx >> 3);
      pre[i] = (grow < p.M) ? ldg4(A + (size_t)grow * lda + koff + 4 * (idx & 7)) : zero4();
    }
  };
  prefetch(0, preA, swA);
  if (nchunks > 1) prefetch(1, preB, swB);
  const float* bsrc = p.Bimg + (size_t)ntile * nchunks * (TC_B_BYTES / 4);
  auto issue_b = [&](int c) {      // thread 0: bulk copy of the (hi | lo) weight images of chunk c
    uint64_t* bar = &bars[c & 3];
    constexpr uint32_t NB = FAST ? IMG_BYTES : TC_B_BYTES;
    mbar_expect_tx(bar, NB);
    bulk_g2s(smemB + (c & 3) * TC_B_BYTES, bsrc + (size_t)c * (TC_B_BYTES / 4), NB, bar);
  };
  if (tid == 0)
    for (int c = 0; c < nchunks && c < TC_B_STAGES; ++c) issue_b(c);
  auto chunk = [&](int c, float4 (&pre)[4], bool& pre_sw) {
    const int s = c & 1, use = c >> 1;
    uint8_t* st = smem + s * TC_A_BYTES;
    if (c >= 2) {
      mbar_wait_warp(&bars[4 + s], (use - 1) & 1);   // MMAs of chunk c-2 done: A stage s and B stage (c-2)&3 free
      if (tid == 0 && c + 2 < nchunks) issue_b(c + 2);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = tid + 256 * i;
      float4 v = pre[i];
      if (pre_sw) v = swish4_m(v);
      if (FAST)
        store_hi4(st, img_off(idx >> 3, idx & 7), v);
      else
        store_split4(st, st + IMG_BYTES, img_off(idx >> 3, idx & 7), v);
    }
    if (c + 2 < nchunks) prefetch(c + 2, pre, pre_sw);
    fence_proxy_async();
    __syncthreads();
    if (warp == 0) {      // the whole warp runs the issue code convergently, one elected lane issues (see elect_one())
      mbar_wait(&bars[c & 3], (c >> 2) & 1);                         // weights of this chunk have landed
      tc_fence_after();
      const uint32_t tm = __shfl_sync(0xffffffffu, tmem, 0);
      const uint32_t a_hi = smem_u32(st), a_lo = a_hi + IMG_BYTES;
      const uint32_t b_hi = smem_u32(smemB + (c & 3) * TC_B_BYTES), b_lo = b_hi + IMG_BYTES;
      const bool leader = elect_one();
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint64_t dah = umma_desc(a_hi + 32 * k, 16, 1024), dal = umma_desc(a_lo + 32 * k, 16, 1024);
        const uint64_t dbh = umma_desc(b_hi + 32 * k, 16, 1024), dbl = umma_desc(b_lo + 32 * k, 16, 1024);
        if (leader) {
          umma_tf32(tm, dah, dbh, IDESC, (c | k) ? 1u : 0u);
          if (!FAST) {
            umma_tf32(tm, dal, dbh, IDESC, 1u);
            umma_tf32(tm, dah, dbl, IDESC, 1u);
          }
        }
      }
      if (leader) umma_commit(&bars[4 + s]);
      __syncwarp();
    }
  };
  for (int c = 0; c < nchunks; c += 2) {
    chunk(c, preA, swA);
    if (c + 1 < nchunks) chunk(c + 1, preB, swB);
  }
  // accumulator complete when the last commit arrives
  {
    const int last = nchunks - 1;
    mbar_wait_warp(&bars[4 + (last & 1)], (last >> 1) & 1);
    tc_fence_after();
  }
  // ---- epilogue: warp w reads lanes 32*(w&3).., columns 64*(w>>2)..; the operand ring is free now (the last commit
  // covers every MMA) and serves as the warps' transposition tiles
  {
    float* tb = reinterpret_cast<float*>(smem) + warp * EPI_TILE_FLOATS;
    const int n0 = ntile * 128;
#pragma unroll 1
    for (int cb = 0; cb < 2; ++cb) {
      const int colbase = 64 * (warp >> 2) + 32 * cb;
      float v[32];
      __syncwarp();
      tmem_ld32(tmem + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)colbase, v);
      lin_epilogue32<true>(p, tb, v, row0 + 32 * (warp & 3), n0 + colbase, lane, EpiStage{epi_stage, epi_stage + 1024, epi_stage + 2048},
                     row0, n0);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 128);
}


// ================================================================================================================
// Warp-specialised variant for many row tiles (large graphs): one persistent CTA per SM walks (row tile, column tile)
// pairs; the phases that k_linear_tc runs back to back inside a tile overlap ACROSS tiles here:
//   producers (8 warps) : fp32 rows of the A operand (up to three column segments, optional swish), two chunks in flight
//                         per thread, tf32 hi | lo split into a 3-stage ring
//   weight loader (1)   : one 32 KiB bulk copy per chunk into the matching ring stage
//   MMA warp            : 12 MMAs per chunk, two TMEM accumulators alternate between tiles
//   epilogue (8 warps)  : the epilogue of k_linear_tc on the accumulator of the previous tile
// Same operands, same images, same epilogue options, same arithmetic per output as k_linear_tc.
constexpr int LW_STAGES = 3;
constexpr int LW_STAGE_BYTES = TC_A_BYTES + TC_B_BYTES;                // A hi | lo, B hi | lo of one chunk
constexpr int LW_EPI_WARPS = 8, LW_PROD_WARPS = 8;
constexpr int LW_MMA_WARP = LW_EPI_WARPS + LW_PROD_WARPS;              // 16; warp 17 = weight loader; 18, 19 idle
constexpr int LW_THREADS = 32 * (LW_MMA_WARP + 4);
constexpr int LW_SMEM = 1024 + LW_STAGES * LW_STAGE_BYTES + 256 + LW_EPI_WARPS * EPI_TILE_FLOATS * 4;

template <bool FAST>
__global__ void __launch_bounds__(LW_THREADS, 1) k_linear_ws(const LinTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + LW_STAGES * LW_STAGE_BYTES);
  uint64_t* a_full = bars;              // [3] producers -> MMA
  uint64_t* w_full = bars + 3;          // [3] bulk copy -> MMA
  uint64_t* empty = bars + 6;           // [3] MMA -> producers and weight loader
  uint64_t* acc_full = bars + 9;        // [2] MMA -> epilogue
  uint64_t* acc_empty = bars + 11;      // [2] epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);
  const int tid = threadIdx.x, warp = warp_index_uniform(), lane = tid & 31;

  int ktot = 0;
  for (int s = 0; s < p.nseg; ++s) ktot += p.ka[s];
  const int nchunks = ktot >> 5;
  const int nct = (p.Nout + 127) / 128;
  const int ntiles = ((p.M + 127) / 128) * nct;
  // tiles of this CTA: blockIdx.x, blockIdx.x + gridDim.x, ...  (the column tiles of a row tile are neighbours: the rows
  // are re-read from L2)
  const int my_tiles = (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (warp == LW_MMA_WARP) tmem_alloc(tmem_slot, 256);
  if (tid == 0) {
    for (int i = 0; i < LW_STAGES; ++i) {
      mbar_init(&a_full[i], LW_PROD_WARPS);
      mbar_init(&w_full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], LW_EPI_WARPS);
    }
    fence_barrier_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp < LW_EPI_WARPS) {
    // =========================================================================== epilogue warps
#pragma unroll 1
    for (int i = 0; i < my_tiles; ++i) {
      const int t = blockIdx.x + i * gridDim.x;
      const int row0 = (t / nct) * 128, ntile = t % nct;
      const int buf = i & 1;
      const int n0 = ntile * 128;
      float* tb = reinterpret_cast<float*>(smem + LW_STAGES * LW_STAGE_BYTES + 256) + warp * EPI_TILE_FLOATS;
      mbar_wait_backoff(&acc_full[buf], (i >> 1) & 1);
      tc_fence_after();
      const uint32_t ta = tmem + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(128 * buf + 64 * (warp >> 2));
#pragma unroll 1
      for (int cb = 0; cb < 2; ++cb) {
        float v[32];
        __syncwarp();
        tmem_ld32(ta + 32 * cb, v);
        if (cb == 1) {          // accumulator drained: the MMA warp may start the tile after next
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc_empty[buf]);
        }
        lin_epilogue32<false>(p, tb, v, row0 + 32 * (warp & 3), n0 + 64 * (warp >> 2) + 32 * cb, lane);
      }
    }
  } else if (warp >= LW_MMA_WARP) {
    reg_dec<40>();
    if (warp == LW_MMA_WARP) {
      // ========================================================================= MMA warp
      constexpr uint32_t IDESC = umma_idesc_tf32(128, 128, 0, 0);
      const uint32_t tm = __shfl_sync(0xffffffffu, tmem, 0);
      const bool leader = elect_one();
      uint32_t s = 0, ph = 0;
#pragma unroll 1
      for (int i = 0; i < my_tiles; ++i) {
        const int buf = i & 1;
        if (i >= 2) mbar_wait_backoff(&acc_empty[buf], ((i >> 1) - 1) & 1);
        tc_fence_after();
        const uint32_t acc = tm + 128 * buf;
#pragma unroll 1
        for (int c = 0; c < nchunks; ++c) {
          mbar_wait_backoff(&a_full[s], ph);
          mbar_wait_backoff(&w_full[s], ph);
          tc_fence_after();
          const uint32_t a_hi = smem_u32(smem + s * LW_STAGE_BYTES), a_lo = a_hi + IMG_BYTES;
          const uint32_t b_hi = a_hi + TC_A_BYTES, b_lo = b_hi + IMG_BYTES;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t dah = umma_desc(a_hi + 32 * k, 16, 1024), dal = umma_desc(a_lo + 32 * k, 16, 1024);
            const uint64_t dbh = umma_desc(b_hi + 32 * k, 16, 1024), dbl = umma_desc(b_lo + 32 * k, 16, 1024);
            if (leader) {
              umma_tf32(acc, dah, dbh, IDESC, (c | k) ? 1u : 0u);
              if (!FAST) {
                umma_tf32(acc, dal, dbh, IDESC, 1u);
                umma_tf32(acc, dah, dbl, IDESC, 1u);
              }
            }
          }
          if (leader) {
            umma_commit(&empty[s]);
            if (c == nchunks - 1) umma_commit(&acc_full[buf]);
          }
          __syncwarp();
          if (++s == LW_STAGES) {
            s = 0;
            ph ^= 1;
          }
        }
      }
    } else if (warp == LW_MMA_WARP + 1) {
      // ========================================================================= weight loader
      if (elect_one()) {
        uint32_t s = 0, ph = 0, n = 0;
        for (int i = 0; i < my_tiles; ++i) {
          const int t = blockIdx.x + i * gridDim.x;
          const float* bsrc = p.Bimg + (size_t)(t % nct) * nchunks * (TC_B_BYTES / 4);
          for (int c = 0; c < nchunks; ++c, ++n) {
            if (n >= LW_STAGES) mbar_wait(&empty[s], ph ^ 1);
            constexpr uint32_t NB = FAST ? IMG_BYTES : TC_B_BYTES;
            mbar_expect_tx(&w_full[s], NB);
            bulk_g2s(smem + s * LW_STAGE_BYTES + TC_A_BYTES, bsrc + (size_t)c * (TC_B_BYTES / 4), NB, &w_full[s]);
            if (++s == LW_STAGES) {
              s = 0;
              ph ^= 1;
            }
          }
        }
      }
      __syncwarp();
    }
  } else {
    // =========================================================================== producer warps
    reg_inc<120>();
    const int pt = tid - 32 * LW_EPI_WARPS;
    const int c16 = pt & 7;
    auto gather = [&](int i, int c, float4 (&pre)[4], bool& sw) {
      const int t = blockIdx.x + i * gridDim.x;
      const int row0 = (t / nct) * 128;
      int seg = 0, koff = c * 32;
      while (seg < p.nseg - 1 && koff >= p.ka[seg]) {
        koff -= p.ka[seg];
        ++seg;
      }
      const float* A = p.A[seg];
      const int lda = p.lda[seg];
      sw = p.aswish[seg] != 0;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int grow = row0 + (pt >> 3) + 32 * q;
        pre[q] = (grow < p.M) ? ldg4(A + (size_t)grow * lda + koff + 4 * c16) : zero4();
      }
    };
    uint32_t s = 0, ph = 0, n = 0;
    auto stage = [&](const float4 (&pre)[4], bool sw) {
      if (n >= LW_STAGES) mbar_wait_backoff(&empty[s], ph ^ 1);
      uint8_t* st = smem + s * LW_STAGE_BYTES;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float4 v = pre[q];
        if (sw) v = swish4_m(v);
        if (FAST)
          store_hi4(st, img_off((pt >> 3) + 32 * q, c16), v);
        else
          store_split4(st, st + IMG_BYTES, img_off((pt >> 3) + 32 * q, c16), v);
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&a_full[s]);
      ++n;
      if (++s == LW_STAGES) {
        s = 0;
        ph ^= 1;
      }
    };
    // two register sets: the loads of the next chunk are in flight while this one is split and stored
    float4 pa[4], pb[4];
    bool swa = false, swb = false;
    const int total = my_tiles * nchunks;
    int i = 0, c = 0;                 // (tile, chunk) of the NEXT gather
    auto advance = [&]() {
      if (++c == nchunks) {
        c = 0;
        ++i;
      }
    };
    if (total > 0) {
      gather(i, c, pa, swa);
      advance();
    }
#pragma unroll 1
    for (int w = 0; w < total; w += 2) {
      if (w + 1 < total) {
        gather(i, c, pb, swb);
        advance();
      }
      stage(pa, swa);
      if (w + 2 < total) {
        gather(i, c, pa, swa);
        advance();
      }
      if (w + 1 < total) stage(pb, swb);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == LW_MMA_WARP) tmem_dealloc(tmem, 256);
}

int launch_linear_tma(const LinTcParams& p, int mode, int grid, cudaStream_t stream);      // linear_tma.cu
int launch_linear_ts(const LinTcParams& p, int mode, int grid, cudaStream_t stream);       // linear_ts.cu

}  // namespace msmp

using namespace msmp;

extern "C" size_t msmp_linear_tc_image_floats(int K, int Nout) {
  return (size_t)((Nout + 127) / 128) * (size_t)(K / 32) * 2 * (IMG_BYTES / 4);
}

extern "C" int msmp_linear_tc_fwd(const float* const* A, const int* lda, const int* ka, const int* aswish, int nseg,
                                  const float* Bimg, const float* bias, const float* side, int lds, int r,
                                  const float* Wside, int ldws, const float* Zmul, int ldz, float* Ypre, int ldpre,
                                  int act, const float* R, int ldr, float* Y, int ldy, int M, int Nout, int mode,
                                  cudaStream_t stream) {
  if (nseg < 1 || nseg > 3 || M < 0 || Nout <= 0 || (Nout & 3) || r < 0 || r > 8) return MSMP_ERR_ARG;
  if (M == 0) return MSMP_OK;
  LinTcParams p{};
  for (int s = 0; s < nseg; ++s) {
    if (ka[s] <= 0 || (ka[s] & 31) || (lda[s] & 3)) return MSMP_ERR_ARG;
    p.A[s] = A[s];
    p.lda[s] = lda[s];
    p.ka[s] = ka[s];
    p.aswish[s] = aswish ? aswish[s] : 0;
  }
  p.nseg = nseg; p.Bimg = Bimg; p.bias = bias; p.side = side; p.lds = lds; p.r = side ? r : 0; p.Wside = Wside;
  p.ldws = ldws; p.Zmul = Zmul; p.ldz = ldz; p.Ypre = Ypre; p.ldpre = ldpre; p.act = act; p.R = R; p.ldr = ldr;
  p.Y = Y; p.ldy = ldy; p.M = M; p.Nout = Nout;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(k_linear_tc<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM) != cudaSuccess ||
        cudaFuncSetAttribute(k_linear_ws<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, LW_SMEM) != cudaSuccess ||
        cudaFuncSetAttribute(k_linear_tc<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM) != cudaSuccess ||
        cudaFuncSetAttribute(k_linear_ws<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, LW_SMEM) != cudaSuccess)
      return MSMP_ERR_CUDA;
    attr_set = true;
  }
  dim3 grid((M + 127) / 128, (Nout + 127) / 128);
  // two or more tiles per SM: the persistent warp-specialised kernel (MSMP_LINEAR_WS_MIN_TILES overrides the threshold,
  // 0 disables it).  Measured: 1 Mi-node layer 7.5 -> 6.35 ms, C4 step with 8 lattices per GPU 49.0 -> 46.9 ms; at the C2
  // size (100 tiles) the single-tile kernel is used.
  static const int ws_min_tiles = [] { const char* e = getenv("MSMP_LINEAR_WS_MIN_TILES"); return e ? atoi(e) : 296; }();
  static const int sms = [] { int d = 0, n = 148; cudaGetDevice(&d); cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, d); return n; }();
  const int ntiles = (int)(grid.x * grid.y);
  if (ws_min_tiles > 0 && ntiles >= ws_min_tiles) {
    // A operand by tensor-map TMA (linear_tma.cu); falls back to the register-staged kernel when a map cannot be made
    // activation operand in tensor memory (linear_ts.cu), then the shared-memory operand variant (linear_tma.cu)
    int rc = launch_linear_ts(p, mode, ntiles < sms ? ntiles : sms, stream);
    if (rc <= 0) return rc;
    rc = launch_linear_tma(p, mode, ntiles < sms ? ntiles : sms, stream);
    if (rc <= 0) return rc;
    if (mode)
      k_linear_ws<true><<<ntiles < sms ? ntiles : sms, LW_THREADS, LW_SMEM, stream>>>(p);
    else
      k_linear_ws<false><<<ntiles < sms ? ntiles : sms, LW_THREADS, LW_SMEM, stream>>>(p);
    MSMP_CHECK_LAUNCH();
    return MSMP_OK;
  }
  if (mode)
    k_linear_tc<true><<<grid, 256, TC_SMEM, stream>>>(p);
  else
    k_linear_tc<false><<<grid, 256, TC_SMEM, stream>>>(p);
  MSMP_CHECK_LAUNCH();
  return MSMP_OK;
}
