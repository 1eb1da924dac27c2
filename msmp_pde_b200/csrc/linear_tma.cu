// Node-level dense layers for many row tiles, third generation: the A operand arrives by TENSOR-MAP TMA.
//
// Same contract, operands, weight images and epilogue as k_linear_ws (linear_tc.cu) -- Y = epilogue([A0|A1|A2] W^T + ...),
// the update / projection / dgrad GEMMs of experiments/models_gnn.py:61-86,124-149 -- but no thread loads an operand:
//
//   loader lane      : per 32-column chunk one cp.async.bulk.tensor.2d (box 128 rows x 32 fp32 columns, SWIZZLE_128B: the
//                      TMA engine writes exactly the UMMA K-major tile image, rows past M are zero filled) for the A operand
//                      and one 1-D bulk copy for the pre-swizzled weight images, both completing on one mbarrier; it runs
//                      up to a ring of stages ahead of the MMA warp, so HBM / L2 latency is hidden by the ring, not by
//                      registers (k_linear_ws kept 32 KB per SM in flight through its producers' registers and measured
//                      ~5000 cycles per chunk against 768 cycles of MMAs)
//   4 converter warps: fp32-parity mode: read the landed fp32 tile, (swish), write tf32-rounded hi in place and the exact
//                      remainder lo next to it (no global access at all).  Reduced-precision mode: the landed tile IS the
//                      operand (kind::tf32 reads the upper 19 bits) -- nothing to do unless the segment wants a swish
//   MMA warp         : 12 (4) MMAs per chunk, two TMEM accumulators alternate between tiles
//   8 epilogue warps : lin_epilogue32 (coalesced through a shared-memory transposition)
#include <cstdlib>
#include <cuda.h>
#include "linear_common.cuh"
#include "msmp_b200.h"

namespace msmp {

constexpr int LM_EPI_WARPS = 8, LM_CV_WARPS = 4;
constexpr int LM_MMA_WARP = LM_EPI_WARPS + LM_CV_WARPS;       // 12; warp 13 = loader
constexpr int LM_THREADS = 32 * (LM_MMA_WARP + 2);
constexpr int LM_RING_BYTES = 3 * (TC_A_BYTES + TC_B_BYTES);  // fp32 mode: 3 stages of 64 KiB; reduced precision: 6 of 32 KiB
constexpr int LM_SMEM = 1024 + LM_RING_BYTES + 512 + LM_EPI_WARPS * EPI_TILE_FLOATS * 4 + EPI_STAGE_FLOATS * 4;
constexpr int LM_MAX_STAGES = 6;

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* tm, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(smem_u32(bar))
      : "memory");
}

template <bool FAST>
__global__ void __launch_bounds__(LM_THREADS, 1) k_linear_tma(const LinTcParams p, const __grid_constant__ CUtensorMap tm0,
                                                             const __grid_constant__ CUtensorMap tm1,
                                                             const __grid_constant__ CUtensorMap tm2) {
  constexpr int STAGES = FAST ? 6 : 3;
  constexpr int A_BYTES = FAST ? IMG_BYTES : TC_A_BYTES;        // raw (= hi) tile [, lo tile]
  constexpr int B_BYTES = FAST ? IMG_BYTES : TC_B_BYTES;        // weight hi image [, lo image]
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + LM_RING_BYTES);
  uint64_t* ld_full = bars;                          // [6] TMA (A tile + weight images) -> converters
  uint64_t* cv_full = bars + LM_MAX_STAGES;          // [6] converters -> MMA
  uint64_t* empty = bars + 2 * LM_MAX_STAGES;        // [6] MMA -> loader
  uint64_t* acc_full = bars + 3 * LM_MAX_STAGES;     // [2] MMA -> epilogue
  uint64_t* acc_empty = acc_full + 2;                // [2] epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  float* epi_tiles = reinterpret_cast<float*>(smem + LM_RING_BYTES + 512);
  float* epi_stage = epi_tiles + LM_EPI_WARPS * EPI_TILE_FLOATS;
  const int tid = threadIdx.x, warp = warp_index_uniform(), lane = tid & 31;

  int ktot = 0;
  for (int s = 0; s < p.nseg; ++s) ktot += p.ka[s];
  const int nchunks = ktot >> 5;
  const int nct = (p.Nout + 127) / 128;
  const int ntiles = ((p.M + 127) / 128) * nct;
  const int my_tiles = (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int k0 = p.ka[0] >> 5, k1 = p.nseg > 1 ? (p.ka[1] >> 5) : (1 << 30);      // chunks of segments 0 and 1

  if (warp == LM_MMA_WARP) tmem_alloc(tmem_slot, 256);
  if (tid == 0) {
    for (int i = 0; i < LM_MAX_STAGES; ++i) {
      mbar_init(&ld_full[i], 1);
      mbar_init(&cv_full[i], LM_CV_WARPS);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], LM_EPI_WARPS);
    }
    fence_barrier_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp < LM_EPI_WARPS) {
    // =========================================================================== epilogue warps
    float* tb = epi_tiles + warp * EPI_TILE_FLOATS;
#pragma unroll 1
    for (int i = 0; i < my_tiles; ++i) {
      const int t = blockIdx.x + i * gridDim.x;
      const int row0 = (t / nct) * 128, n0 = (t % nct) * 128;
      const int buf = i & 1;
      // this tile's side values / side weights / bias -> shared memory, while its MMAs are still running
      // (launches without a side term keep their epilogue warps independent: the bias alone is a per-thread constant)
      EpiStage es{nullptr, nullptr, nullptr};
      if (p.r > 0) {
        asm volatile("bar.sync 1, %0;" ::"n"(32 * LM_EPI_WARPS) : "memory");        // the previous tile's readers are done
        epi_stage_fill(p, epi_stage, row0, n0, tid, 32 * LM_EPI_WARPS);
        asm volatile("bar.sync 1, %0;" ::"n"(32 * LM_EPI_WARPS) : "memory");
        es = EpiStage{epi_stage, epi_stage + 1024, epi_stage + 2048};
      }
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
        lin_epilogue32<true>(p, tb, v, row0 + 32 * (warp & 3), n0 + 64 * (warp >> 2) + 32 * cb, lane, es, row0, n0);
      }
    }
  } else if (warp < LM_MMA_WARP) {
    // =========================================================================== converter warps
    const int ct = tid - 32 * LM_EPI_WARPS;            // 0..127
    const bool sw0 = p.aswish[0] != 0, sw1 = p.aswish[1] != 0, sw2 = p.aswish[2] != 0;
    uint32_t s = 0, ph = 0;
    const int total = my_tiles * nchunks;
    int c = 0;
#pragma unroll 1
    for (int w = 0; w < total; ++w) {
      const bool sw = c < k0 ? sw0 : (c - k0 < k1 ? sw1 : sw2);
      mbar_wait_backoff(&ld_full[s], ph);
      if ((!FAST || sw) && !(p.dbg & 2)) {
        while (!mbar_try_wait(&ld_full[s], ph)) {          // every lane observes the TMA completion itself
        }
        uint8_t* a_hi = smem + s * STAGE_BYTES;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int idx = ct + 128 * q;
          const uint32_t off = img_off(idx >> 3, idx & 7);
          float4 v = *reinterpret_cast<const float4*>(a_hi + off);
          if (sw) v = swish4_m(v);
          if (FAST) {
            *reinterpret_cast<float4*>(a_hi + off) = v;
          } else {
            store_split4(a_hi, a_hi + IMG_BYTES, off, v);
          }
        }
        fence_proxy_async();
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&cv_full[s]);
      if (++c == nchunks) c = 0;
      if (++s == STAGES) {
        s = 0;
        ph ^= 1;
      }
    }
  } else if (warp == LM_MMA_WARP) {
    // =========================================================================== MMA warp
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
        mbar_wait_backoff(&ld_full[s], ph);         // the TMA writes themselves (async proxy -> async proxy)
        mbar_wait_backoff(&cv_full[s], ph);
        tc_fence_after();
        const uint32_t a_hi = smem_u32(smem + s * STAGE_BYTES), a_lo = a_hi + IMG_BYTES;
        const uint32_t b_hi = a_hi + A_BYTES, b_lo = b_hi + IMG_BYTES;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t dah = umma_desc(a_hi + 32 * k, 16, 1024), dal = umma_desc(a_lo + 32 * k, 16, 1024);
          const uint64_t dbh = umma_desc(b_hi + 32 * k, 16, 1024), dbl = umma_desc(b_lo + 32 * k, 16, 1024);
          if (leader && !(p.dbg & 4)) {
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
        if (++s == STAGES) {
          s = 0;
          ph ^= 1;
        }
      }
    }
  } else {
    // =========================================================================== loader (one lane)
    if (elect_one()) {
      uint32_t s = 0, ph = 0, n = 0;
      for (int i = 0; i < my_tiles; ++i) {
        const int t = blockIdx.x + i * gridDim.x;
        const int row0 = (t / nct) * 128;
        const float* bsrc = p.Bimg + (size_t)(t % nct) * nchunks * (TC_B_BYTES / 4);
        for (int c = 0; c < nchunks; ++c, ++n) {
          if (n >= (uint32_t)STAGES) mbar_wait(&empty[s], ph ^ 1);
          uint8_t* st = smem + s * STAGE_BYTES;
          mbar_expect_tx(&ld_full[s], ((p.dbg & 16) ? 0 : IMG_BYTES) + ((p.dbg & 8) ? 0 : B_BYTES));
          const CUtensorMap* tmap = c < k0 ? &tm0 : (c - k0 < k1 ? &tm1 : &tm2);
          const int kc = c < k0 ? c : (c - k0 < k1 ? c - k0 : c - k0 - k1);
          if (!(p.dbg & 16)) tma_load_2d(st, tmap, 32 * kc, row0, &ld_full[s]);
          if (!(p.dbg & 8)) bulk_g2s(st + A_BYTES, bsrc + (size_t)c * (TC_B_BYTES / 4), B_BYTES, &ld_full[s]);
          if (++s == STAGES) {
            s = 0;
            ph ^= 1;
          }
        }
      }
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == LM_MMA_WARP) tmem_dealloc(tmem, 256);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled() {
  static EncodeTiledFn fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      f = nullptr;
    return reinterpret_cast<EncodeTiledFn>(f);
  }();
  return fn;
}

// [rows x cols] fp32 row-major (row stride ld floats) as a 2-D tensor map with a box_cols x box_rows box; swizzle128: the
// 128-byte swizzle of the UMMA K-major tile image (box_cols = 32), else plain row-major boxes
bool tensor_map_2d(CUtensorMap* tm, const float* A, int ld, int cols, int rows, int box_cols, int box_rows, bool swizzle128) {
  EncodeTiledFn enc = encode_tiled();
  if (!enc) return false;
  if ((reinterpret_cast<uintptr_t>(A) & 15) || (ld & 3)) return false;
  const cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t gstride[1] = {(cuuint64_t)ld * 4};
  const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  static const int promo = [] { const char* e = getenv("MSMP_TMA_PROMO"); return e ? atoi(e) : 128; }();
  return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(A), gdim, gstride, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
             promo == 256 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : (promo == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : (promo == 64 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B : CU_TENSOR_MAP_L2_PROMOTION_L2_128B)),
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// [M x K] fp32 row-major (row stride lda floats): 32-column x 128-row boxes, 128-byte swizzle (the node GEMMs' A operand)
bool linear_make_map(CUtensorMap* tm, const float* A, int lda, int K, int M) {
  return tensor_map_2d(tm, A, lda, K, M, 32, 128, true);
}

// Returns MSMP_OK when the launch was made, 1 when this path cannot take the call (the caller falls back to k_linear_ws).
int launch_linear_tma(const LinTcParams& p_in, int mode, int grid, cudaStream_t stream) {
  static const bool enabled = [] { const char* e = getenv("MSMP_LINEAR_TMA"); return !(e && atoi(e) == 0); }();
  if (!enabled) return 1;
  static const int dbg = [] { const char* e = getenv("MSMP_LIN_DBG"); return e ? atoi(e) : 0; }();
  LinTcParams p = p_in;
  p.dbg = dbg;
  alignas(64) CUtensorMap tm[3];
  for (int s = 0; s < 3; ++s) {
    const int q = s < p.nseg ? s : 0;
    if (!linear_make_map(&tm[s], p.A[q], p.lda[q], p.ka[q], p.M)) return 1;
  }
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(k_linear_tma<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, LM_SMEM) != cudaSuccess ||
        cudaFuncSetAttribute(k_linear_tma<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, LM_SMEM) != cudaSuccess)
      return MSMP_ERR_CUDA;
    attr_set = true;
  }
  if (mode)
    k_linear_tma<true><<<grid, LM_THREADS, LM_SMEM, stream>>>(p, tm[0], tm[1], tm[2]);
  else
    k_linear_tma<false><<<grid, LM_THREADS, LM_SMEM, stream>>>(p, tm[0], tm[1], tm[2]);
  MSMP_CHECK_LAUNCH();
  return MSMP_OK;
}

}  // namespace msmp
