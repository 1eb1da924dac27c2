// Node-level dense layers for many row tiles, fourth generation: the activation operand lives in TENSOR MEMORY.
//
// Same contract, weight images and epilogue as k_linear_tma (linear_tma.cu) -- Y = epilogue([A0|A1|A2] W^T + ...), the update /
// projection / dgrad GEMMs of experiments/models_gnn.py:61-86,124-149.  Ablation of k_linear_tma on 131 072 rows (round 2,
// MSMP_LIN_DBG): its three 64 KiB stages keep only 48 KiB of activations in flight per SM (the A stream alone ran at
// 3.6 TB/s), and every chunk crosses shared memory four times (TMA write, converter read, hi | lo write, 12 MMAs reading
// both operands: 192 KiB of port traffic per 16 KiB of activations).  Here
//
//   A loader lane     : cp.async.bulk.tensor.2d of the fp32 [128 rows x 32 columns] box into a RAW ring of 6 x 16 KiB
//                       (twice the bytes in flight; a slot is free again as soon as the converters have read it)
//   B loader lane     : the pre-swizzled weight chunk images (hi | lo, 32 KiB) from L2 into their own ring of 3
//   4 converter warps : one per TMEM lane quadrant: a thread owns one row, reads its 32 fp32 values from the landed tile,
//                       (swish), splits them into tf32 hi + exact remainder lo and writes both with tcgen05.st into one of
//                       four 64-column operand stages of tensor memory -- no shared-memory write at all
//   MMA warp          : tcgen05.mma with the A operand FROM TENSOR MEMORY (3xTF32: hi*hi + lo*hi + hi*lo): the tensor pipe
//                       reads only the weight images from shared memory (half the operand bytes per MMA)
//   16 epilogue warps : lin_epilogue32 out of two alternating 128-column accumulators, one 32 x 32 block per warp and tile
//                       (with 8 warps the epilogue -- ~10 cycles per dependent instruction at two warps per scheduler --
//                       was the bound of every launch: MSMP_LIN_DBG ablation, 0.54 ms per layer with, 0.30 without it)
// Registers: 768 threads start with 80; the converter and MMA / loader warpgroups hand theirs (setmaxnreg 56 / 40) to the
// four epilogue warpgroups (96).
//
// Tensor memory: columns 0..255 accumulators, 256..511 the four operand stages [hi 32 | lo 32].
// Reduced-precision mode (FAST): the raw values are the operand (kind::tf32 reads the upper 19 bits), one product.
#include <cstdlib>
#include <cuda.h>
#include "linear_common.cuh"
#include "msmp_b200.h"

namespace msmp {

constexpr int LS_EPI_WARPS = 16, LS_CV_WARPS = 4;
constexpr int LS_MMA_WARP = LS_EPI_WARPS + LS_CV_WARPS;       // 20; 21 = A loader, 22 = B loader, 23 idle (warpgroup padding)
constexpr int LS_THREADS = 32 * (LS_MMA_WARP + 4);
// Ring depths (-DLS_RA_N / -DLS_RB_N for experiments).  Measured at 131 072 rows, one layer's six launches: (5, 3) 0.428 ms,
// (7, 2) 0.433, (3, 4) 0.426 -- with the operand in tensor memory the kernel is not sensitive to the ring depths any more.
#ifndef LS_RA_N
#define LS_RA_N 5
#define LS_RB_N 3
#endif
constexpr int LS_RA = LS_RA_N;                                // raw activation ring: 5 x 16 KiB
constexpr int LS_RB = LS_RB_N;                                // weight ring: 3 x 32 KiB (FAST: 6 x 16 KiB)
constexpr int LS_TA = 4;                                      // operand stages in tensor memory
constexpr int LS_RING_BYTES = LS_RA * IMG_BYTES + LS_RB * TC_B_BYTES;
constexpr int LS_SMEM = 1024 + LS_RING_BYTES + 512 + LS_EPI_WARPS * EPI_TILE_FLOATS * 4 + EPI_STAGE_FLOATS * 4;
constexpr uint32_t LS_ACOL = 256;                             // first operand-stage column

__device__ __forceinline__ void tma_load_2d_ls(void* dst, const CUtensorMap* tm, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(smem_u32(bar))
      : "memory");
}

// 32 lanes x 16 consecutive columns written from registers (thread i of the warp writes lane base + i)
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::
          "r"(taddr),
      "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
      "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
      "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
      "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
      : "memory");
}

template <bool FAST>
__global__ void __launch_bounds__(LS_THREADS, 1) k_linear_ts(const LinTcParams p, const __grid_constant__ CUtensorMap tm0,
                                                            const __grid_constant__ CUtensorMap tm1,
                                                            const __grid_constant__ CUtensorMap tm2) {
  constexpr int RB = FAST ? 2 * LS_RB : LS_RB;
  constexpr int B_BYTES = FAST ? IMG_BYTES : TC_B_BYTES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* raw_ring = smem;
  uint8_t* b_ring = smem + LS_RA * IMG_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + LS_RING_BYTES);
  uint64_t* raw_full = bars;                         // [6] TMA -> converters
  uint64_t* raw_empty = raw_full + LS_RA;            // [6] converters -> A loader
  uint64_t* a_full = raw_empty + LS_RA;              // [4] converters -> MMA
  uint64_t* a_empty = a_full + LS_TA;                // [4] MMA -> converters
  uint64_t* b_full = a_empty + LS_TA;                // [6] bulk copy -> MMA
  uint64_t* b_empty = b_full + 2 * LS_RB;            // [6] MMA -> B loader
  uint64_t* acc_full = b_empty + 2 * LS_RB;          // [2] MMA -> epilogue
  uint64_t* acc_empty = acc_full + 2;                // [2] epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  float* epi_tiles = reinterpret_cast<float*>(smem + LS_RING_BYTES + 512);
  float* epi_stage = epi_tiles + LS_EPI_WARPS * EPI_TILE_FLOATS;
  const int tid = threadIdx.x, warp = warp_index_uniform(), lane = tid & 31;

  int ktot = 0;
  for (int s = 0; s < p.nseg; ++s) ktot += p.ka[s];
  const int nchunks = ktot >> 5;
  const int nct = (p.Nout + 127) / 128;
  const int ntiles = ((p.M + 127) / 128) * nct;
  const int my_tiles = (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int k0 = p.ka[0] >> 5, k1 = p.nseg > 1 ? (p.ka[1] >> 5) : (1 << 30);      // chunks of segments 0 and 1
  const int total = my_tiles * nchunks;

  if (warp == LS_MMA_WARP) tmem_alloc(tmem_slot, 512);
  if (tid == 0) {
    for (int i = 0; i < LS_RA; ++i) {
      mbar_init(&raw_full[i], 1);
      mbar_init(&raw_empty[i], LS_CV_WARPS);
    }
    for (int i = 0; i < LS_TA; ++i) {
      mbar_init(&a_full[i], LS_CV_WARPS);
      mbar_init(&a_empty[i], 1);
    }
    for (int i = 0; i < 2 * LS_RB; ++i) {
      mbar_init(&b_full[i], 1);
      mbar_init(&b_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], LS_EPI_WARPS);
    }
    fence_barrier_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp < LS_EPI_WARPS) {
    // =========================================================================== epilogue warps
    reg_inc<96>();
    float* tb = epi_tiles + warp * EPI_TILE_FLOATS;
#pragma unroll 1
    for (int i = 0; i < my_tiles; ++i) {
      const int t = blockIdx.x + i * gridDim.x;
      const int row0 = (t / nct) * 128, n0 = (t % nct) * 128;
      const int buf = i & 1;
      EpiStage es{nullptr, nullptr, nullptr};
      if (p.r > 0) {                                   // side values / side weights / bias of this tile -> shared memory
        asm volatile("bar.sync 1, %0;" ::"n"(32 * LS_EPI_WARPS) : "memory");
        epi_stage_fill(p, epi_stage, row0, n0, tid, 32 * LS_EPI_WARPS);
        asm volatile("bar.sync 1, %0;" ::"n"(32 * LS_EPI_WARPS) : "memory");
        es = EpiStage{epi_stage, epi_stage + 1024, epi_stage + 2048};
      }
      mbar_wait_backoff(&acc_full[buf], (i >> 1) & 1);
      tc_fence_after();
      float v[32];
      tmem_ld32(tmem + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(128 * buf + 32 * (warp >> 2)), v);
      tc_fence_before();          // accumulator block drained: the MMA warp may start the tile after next
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);
      lin_epilogue32<true>(p, tb, v, row0 + 32 * (warp & 3), n0 + 32 * (warp >> 2), lane, es, row0, n0);
    }
  } else if (warp < LS_MMA_WARP) {
    // =========================================================================== converter warps (one per lane quadrant)
    reg_dec<56>();
    const int row = 32 * (warp & 3) + lane;            // this thread's row = its TMEM lane
    const uint32_t lane_off = (uint32_t)(32 * (warp & 3)) << 16;
    const bool sw0 = p.aswish[0] != 0, sw1 = p.aswish[1] != 0, sw2 = p.aswish[2] != 0;
#pragma unroll 1
    for (int w = 0; w < total; ++w) {
      const int c = w % nchunks;
      const bool sw = c < k0 ? sw0 : (c - k0 < k1 ? sw1 : sw2);
      const uint32_t rs = (uint32_t)w % LS_RA, rph = ((uint32_t)w / LS_RA) & 1;
      const uint32_t ts = (uint32_t)w % LS_TA, tu = (uint32_t)w / LS_TA;
      while (!mbar_try_wait(&raw_full[rs], rph)) __nanosleep(40);          // every lane observes the TMA completion itself
      const uint8_t* raw = raw_ring + rs * IMG_BYTES;
      if (tu > 0) mbar_wait_backoff(&a_empty[ts], (tu - 1) & 1);           // the MMAs that read this stage are complete
      tc_fence_after();
      if (!(p.dbg & 2)) {
        const uint32_t col = tmem + lane_off + LS_ACOL + 64 * ts;
#pragma unroll
        for (int h = 0; h < 4; ++h) {                                       // 8 columns at a time (56 registers per thread)
          float4 x0 = *reinterpret_cast<const float4*>(raw + img_off(row, 2 * h));
          float4 x1 = *reinterpret_cast<const float4*>(raw + img_off(row, 2 * h + 1));
          if (sw) {
            x0 = swish4_m(x0);
            x1 = swish4_m(x1);
          }
          float hi[8], lo[8];
          if (FAST) {
            hi[0] = x0.x; hi[1] = x0.y; hi[2] = x0.z; hi[3] = x0.w; hi[4] = x1.x; hi[5] = x1.y; hi[6] = x1.z; hi[7] = x1.w;
          } else {
            split_tf32(x0.x, hi[0], lo[0]);
            split_tf32(x0.y, hi[1], lo[1]);
            split_tf32(x0.z, hi[2], lo[2]);
            split_tf32(x0.w, hi[3], lo[3]);
            split_tf32(x1.x, hi[4], lo[4]);
            split_tf32(x1.y, hi[5], lo[5]);
            split_tf32(x1.z, hi[6], lo[6]);
            split_tf32(x1.w, hi[7], lo[7]);
          }
          tmem_st8(col + 8 * h, hi);
          if (!FAST) tmem_st8(col + 32 + 8 * h, lo);
        }
        tmem_st_wait();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&raw_empty[rs]);
        mbar_arrive(&a_full[ts]);
      }
    }
  } else if (warp == LS_MMA_WARP) {
    // =========================================================================== MMA warp
    reg_dec<40>();
    constexpr uint32_t IDESC = umma_idesc_tf32(128, 128, 0, 0);
    const uint32_t tm = __shfl_sync(0xffffffffu, tmem, 0);
    const bool leader = elect_one();
    uint32_t w = 0;
#pragma unroll 1
    for (int i = 0; i < my_tiles; ++i) {
      const int buf = i & 1;
      if (i >= 2) mbar_wait_backoff(&acc_empty[buf], ((i >> 1) - 1) & 1);
      tc_fence_after();
      const uint32_t acc = tm + 128 * buf;
#pragma unroll 1
      for (int c = 0; c < nchunks; ++c, ++w) {
        const uint32_t ts = w % LS_TA, tph = (w / LS_TA) & 1;
        const uint32_t bs = w % RB, bph = (w / RB) & 1;
        mbar_wait_backoff(&b_full[bs], bph);
        mbar_wait_backoff(&a_full[ts], tph);
        tc_fence_after();
        const uint32_t a_hi = tm + LS_ACOL + 64 * ts, a_lo = a_hi + 32;
        const uint32_t b_hi = smem_u32(b_ring + bs * B_BYTES), b_lo = b_hi + IMG_BYTES;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t dbh = umma_desc(b_hi + 32 * k, 16, 1024), dbl = umma_desc(b_lo + 32 * k, 16, 1024);
          if (leader && !(p.dbg & 4)) {
            umma_tf32_ts(acc, a_hi + 8 * k, dbh, IDESC, (c | k) ? 1u : 0u);
            if (!FAST) {
              umma_tf32_ts(acc, a_lo + 8 * k, dbh, IDESC, 1u);
              umma_tf32_ts(acc, a_hi + 8 * k, dbl, IDESC, 1u);
            }
          }
        }
        if (leader) {
          umma_commit(&a_empty[ts]);
          umma_commit(&b_empty[bs]);
          if (c == nchunks - 1) umma_commit(&acc_full[buf]);
        }
        __syncwarp();
      }
    }
  } else if (warp == LS_MMA_WARP + 1) {
    // =========================================================================== A loader (one lane)
    reg_dec<40>();
    if (elect_one()) {
      uint32_t w = 0;
      for (int i = 0; i < my_tiles; ++i) {
        const int t = blockIdx.x + i * gridDim.x;
        const int row0 = (t / nct) * 128;
        for (int c = 0; c < nchunks; ++c, ++w) {
          const uint32_t rs = w % LS_RA, ru = w / LS_RA;
          if (ru > 0) mbar_wait(&raw_empty[rs], (ru - 1) & 1);
          mbar_expect_tx(&raw_full[rs], (p.dbg & 16) ? 0 : IMG_BYTES);
          const CUtensorMap* tmap = c < k0 ? &tm0 : (c - k0 < k1 ? &tm1 : &tm2);
          const int kc = c < k0 ? c : (c - k0 < k1 ? c - k0 : c - k0 - k1);
          if (!(p.dbg & 16)) tma_load_2d_ls(raw_ring + rs * IMG_BYTES, tmap, 32 * kc, row0, &raw_full[rs]);
        }
      }
    }
    __syncwarp();
  } else if (warp == LS_MMA_WARP + 2) {
    // =========================================================================== B loader (one lane)
    reg_dec<40>();
    if (elect_one()) {
      uint32_t w = 0;
      for (int i = 0; i < my_tiles; ++i) {
        const int t = blockIdx.x + i * gridDim.x;
        const float* bsrc = p.Bimg + (size_t)(t % nct) * nchunks * (TC_B_BYTES / 4);
        for (int c = 0; c < nchunks; ++c, ++w) {
          const uint32_t bs = w % RB, bu = w / RB;
          if (bu > 0) mbar_wait(&b_empty[bs], (bu - 1) & 1);
          mbar_expect_tx(&b_full[bs], (p.dbg & 8) ? 0 : B_BYTES);
          if (!(p.dbg & 8)) bulk_g2s(b_ring + bs * B_BYTES, bsrc + (size_t)c * (TC_B_BYTES / 4), B_BYTES, &b_full[bs]);
        }
      }
    }
    __syncwarp();
  } else {
    reg_dec<40>();                                     // fourth warp of the MMA / loader warpgroup: nothing to do
  }
  tc_fence_before();
  __syncthreads();
  if (warp == LS_MMA_WARP) tmem_dealloc(tmem, 512);
}

bool linear_make_map(CUtensorMap* tm, const float* A, int lda, int K, int M);      // linear_tma.cu

// Returns MSMP_OK when the launch was made, 1 when this path cannot take the call (the caller falls back to k_linear_tma).
int launch_linear_ts(const LinTcParams& p_in, int mode, int grid, cudaStream_t stream) {
  static const bool enabled = [] { const char* e = getenv("MSMP_LINEAR_TS"); return !(e && atoi(e) == 0); }();
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
    if (cudaFuncSetAttribute(k_linear_ts<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, LS_SMEM) != cudaSuccess ||
        cudaFuncSetAttribute(k_linear_ts<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, LS_SMEM) != cudaSuccess)
      return MSMP_ERR_CUDA;
    attr_set = true;
  }
  if (mode)
    k_linear_ts<true><<<grid, LS_THREADS, LS_SMEM, stream>>>(p, tm[0], tm[1], tm[2]);
  else
    k_linear_ts<false><<<grid, LS_THREADS, LS_SMEM, stream>>>(p, tm[0], tm[1], tm[2]);
  MSMP_CHECK_LAUNCH();
  return MSMP_OK;
}

}  // namespace msmp
