// Warp-specialised tensor-core edge kernels (tcgen05 / TMEM, 3xTF32) -- the pipelined successor of edge_tc.cu.
//
// One persistent CTA per SM walks a contiguous range of 128-edge tiles (edges are destination sorted).  Three roles
// run concurrently and hand tiles to each other through mbarriers, so the gathers of tile t+1, the MMAs of tile t and
// the epilogue of tile t-1 overlap (edge_tc.cu ran them back to back: 28 k cycles per tile on 1 Mi x 6 Mi graphs):
//
//   producers (8 warps) : gather P[dst] + Q[src] (fwd) / dagg[dst], z2 (bwd) with LDG.128, two 32-column chunks in
//                         flight per thread, swish / sw', tf32 hi|lo split, store into a 3-stage ring of 128B-swizzled
//                         operand images; indices of the next two tiles are prefetched into an 8-slot ring.
//                         Backward only: a second pass per tile gathers z1 = P[dst] + Q[src], writes a1 = sw(z1) and
//                         leaves sw'(z1) in a shared [128 x 128] tile for the epilogue.
//   MMA warp            : D^T[channel][edge] += W . tile^T.  The weight matrix is the A operand and is RESIDENT IN
//                         TENSOR MEMORY (tf32 hi | lo, 2 x 128 columns, written once with tcgen05.st); the staged edge
//                         tile is the B operand (N = 128 edges).  No weight bytes in shared memory, half the shared-
//                         memory operand traffic per MMA, and no pre-packed weight images: the kernel reads the fp32
//                         parameter itself.  Two accumulators (2 x 128 columns) alternate between tiles: 512 columns.
//   epilogue (4 warps)  : a thread owns one output channel (its TMEM lane) and walks the 128 edges of the tile:
//                         bias + swish (fwd) or * sw'(z1) (bwd), one coalesced 128-byte store per warp and edge row,
//                         destination-segment sums as running sums in registers (segment boundaries are warp-uniform;
//                         fixed order, no atomics).  Segments cut by a tile boundary use the carry buffer + ordered
//                         fix-up exactly like edge_tc.cu.
#include <cstdlib>
#include "umma.cuh"
#include "msmp_b200.h"

namespace msmp {

constexpr int EW_TILE = 128;
constexpr int EW_STAGES = 3;
constexpr int EW_STAGE_BYTES = 2 * IMG_BYTES;          // hi | lo image of one [128 edges x 32 columns] chunk
constexpr int EW_EH = 2;                               // epilogue warps per TMEM lane quadrant
constexpr int EW_UNIT = EW_TILE / EW_EH;               // edges walked by one epilogue thread = carry granularity
constexpr int EW_EPI_WARPS = 4 * EW_EH, EW_PROD_WARPS = 8;
constexpr int EW_PROD_T0 = 32 * EW_EPI_WARPS;          // first producer thread
constexpr int EW_PROD_THREADS = 32 * EW_PROD_WARPS;
constexpr int EW_MMA_WARP = EW_EPI_WARPS + EW_PROD_WARPS;
// Warps 0..7 epilogue, 8..15 producers, 16 MMA issue (17..19 only fill its warpgroup).  640 threads start with 96
// registers each; the MMA warpgroup hands registers to the producers with setmaxnreg (40 / 120), whose two float4
// register sets would otherwise spill.
constexpr int EW_THREADS = 32 * (EW_MMA_WARP + 4);
constexpr int EW_SLOTS = 8;                            // index ring (tiles): a slot is rewritten 7 tiles later, when the
                                                       // epilogue that read it has long finished
constexpr int EW_DST_LD = 132;                         // dst[-1 .. 128] of a tile (+ padding)
constexpr int EW_S_BYTES = EW_TILE * 128 * 4;          // backward: sw'(z1) tile
constexpr int EW_IDX_BYTES = EW_SLOTS * (EW_TILE + EW_DST_LD + EW_TILE) * 4 + 132 * 4;      // + 1 / n table
constexpr uint32_t EW_ACC = 0, EW_WHI = 256, EW_WLO = 384;      // TMEM columns
template <bool BWD>
constexpr int ew_smem() { return 1024 + EW_STAGES * EW_STAGE_BYTES + (BWD ? EW_S_BYTES : 0) + 256 + EW_IDX_BYTES; }

struct EdgeWsParams {
  const float* P; const float* Q; int ldpq;
  const int* src; const int* dst;
  const float* inv_deg_e;           // bwd: 1 / max(deg(dst e), 1) per edge
  const float* W; int w_rs, w_cs;   // A operand [m][k] = W[m * w_rs + k * w_cs]  (fwd: W2[n][k];  bwd: W2^T)
  const float* b2;                  // fwd
  float* z2;                        // fwd: out [E][128] (may be null);  bwd: in
  const float* dagg; int lddagg;    // bwd
  float* dz2; float* a1; float* dz1;      // bwd outs [E][128]
  float* out; int ldo;              // fwd: agg [N][128] (mean);  bwd: dP [N][ldo] (sum)
  float* carry;                     // [T * EW_EH][2][128], per unit of EW_UNIT edges
  int E; int T; int N;
};

// sigmoid_mufu / swish_m / dswish_m (activations on the MUFU pipe) live in common.cuh
__device__ __forceinline__ void producer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(EW_PROD_THREADS) : "memory"); }

// -DMSMP_EW_TICKS: CTA 0 prints the cycles each role spent waiting on its barriers (diagnostic builds only)
#ifdef MSMP_EW_TICKS
#define EW_TICK_DECL(n) long long n = 0
#define EW_TIMED(acc, stmt) do { const long long t_ = clock64(); stmt; acc += clock64() - t_; } while (0)
#else
#define EW_TICK_DECL(n)
#define EW_TIMED(acc, stmt) do { stmt; } while (0)
#endif

template <bool BWD, bool ZFILL>
__global__ void __launch_bounds__(EW_THREADS, 1) k_edge_ws(const EdgeWsParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smB = smem;                                                       // operand ring
  float* smS = reinterpret_cast<float*>(smem + EW_STAGES * EW_STAGE_BYTES);    // bwd only
  uint8_t* after = smem + EW_STAGES * EW_STAGE_BYTES + (BWD ? EW_S_BYTES : 0);
  uint64_t* bars = reinterpret_cast<uint64_t*>(after);
  uint64_t* full = bars;                    // [3] producers -> MMA
  uint64_t* empty = bars + 3;               // [3] MMA -> producers
  uint64_t* acc_full = bars + 6;            // [2] MMA -> epilogue
  uint64_t* acc_empty = bars + 8;           // [2] epilogue -> MMA
  uint64_t* s_full = bars + 10;             // producers -> epilogue (bwd)
  uint64_t* s_empty = bars + 11;            // epilogue -> producers (bwd)
  uint64_t* w_full = bars + 12;             // weights are in tensor memory
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);
  int* s_src = reinterpret_cast<int*>(after + 256);         // [SLOTS][128]
  int* s_dst = s_src + EW_SLOTS * EW_TILE;                  // [SLOTS][132]: entry 0 = dst[e0 - 1], 1 + r = row r
  float* s_sc = reinterpret_cast<float*>(s_dst + EW_SLOTS * EW_DST_LD);      // [SLOTS][128]
  float* s_inv = s_sc + EW_SLOTS * EW_TILE;                 // [129]: 1.0f / n
  const int tid = threadIdx.x, warp = warp_index_uniform(), lane = tid & 31;
  // contiguous tile range of this CTA
  const int tb = (int)((long long)blockIdx.x * p.T / gridDim.x);
  const int ntile = (int)((long long)(blockIdx.x + 1) * p.T / gridDim.x) - tb;

  if (warp == EW_MMA_WARP) tmem_alloc(tmem_slot, 512);
  if (tid == 0) {
    for (int i = 0; i < EW_STAGES; ++i) {
      mbar_init(&full[i], EW_PROD_WARPS);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], EW_EPI_WARPS);
    }
    mbar_init(s_full, EW_PROD_WARPS);
    mbar_init(s_empty, EW_EPI_WARPS);
    mbar_init(w_full, EW_EPI_WARPS);
    fence_barrier_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp < EW_EPI_WARPS) {
    // =========================================================================== epilogue warps
    const int ch = 32 * (warp & 3) + lane;
    const int uh = warp >> 2;                                  // which EW_UNIT-edge part of every tile
    const uint32_t lane_off = (uint32_t)(32 * (warp & 3)) << 16;
    // ---- weights -> tensor memory (row ch of the A operand, tf32 hi | lo); the warps of a quadrant share the columns
#pragma unroll 1
    for (int kb = 32 * uh; kb < 128; kb += 32 * EW_EH) {
      float w[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) w[j] = __ldg(p.W + (size_t)ch * p.w_rs + (size_t)(kb + j) * p.w_cs);
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        float hi[8], lo[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) split_tf32(w[j + q], hi[q], lo[q]);
        tmem_st8(tmem + lane_off + EW_WHI + kb + j, hi);
        tmem_st8(tmem + lane_off + EW_WLO + kb + j, lo);
      }
    }
    tmem_st_wait();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(w_full);
    const float bias = BWD ? 0.f : __ldg(p.b2 + ch);
    if (uh == 0) s_inv[ch + 1] = 1.0f / (float)(ch + 1);
    asm volatile("bar.sync 2, %0;" ::"n"(32 * EW_EPI_WARPS) : "memory");
    EW_TICK_DECL(t_acc); EW_TICK_DECL(t_s); EW_TICK_DECL(t_all);
#ifdef MSMP_EW_TICKS
    const long long t_begin = clock64();
#endif

#pragma unroll 1
    for (int i = 0; i < ntile; ++i) {
      const int tile = tb + i;
      const int e0 = tile * EW_TILE;
      const int valid = min(EW_TILE, p.E - e0);
      const int buf = i & 1;
      const int* sd = s_dst + (i & (EW_SLOTS - 1)) * EW_DST_LD + 1;
      EW_TIMED(t_acc, mbar_wait_backoff(&acc_full[buf], (i >> 1) & 1));
      if (BWD) EW_TIMED(t_s, mbar_wait_backoff(s_full, i & 1));
      tc_fence_after();
      // Destination-segment sums: running sums in registers.  The segment boundaries are warp-uniform bit masks
      // (ballots over the tile's destination indices incl. the two neighbouring edges), the element-wise math of a
      // 32-edge block is branch-free (32 independent chains in flight), and only a segment's last edge branches.
      const int u0 = EW_UNIT * uh;                     // this warp's unit = tile rows [u0, min(u0 + EW_UNIT, valid))
      const int uvalid = min(u0 + EW_UNIT, valid);
      const int unit = tile * EW_EH + uh;
      float sum = 0.f;
      int seg0 = u0;
      bool left = sd[u0 - 1] != sd[u0];          // does the unit's first segment start here?
#pragma unroll 1
      for (int cb = 0; cb < 4 / EW_EH; ++cb) {
        float v[32];
        __syncwarp();
        tmem_ld32(tmem + lane_off + EW_ACC + (uint32_t)(128 * buf + u0 + 32 * cb), v);
        if (cb == 4 / EW_EH - 1) {          // accumulator drained: the MMA warp may start the tile after next
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc_empty[buf]);
        }
        const int eb = u0 + 32 * cb;
        const int nb = min(32, valid - eb);          // live edges of this block
        if (nb <= 0) continue;
        const int dj = sd[eb + lane];
        const uint32_t live = nb >= 32 ? 0xffffffffu : ((1u << nb) - 1u);
        // (the shuffles from lane 0 let the compiler see the masks as warp-uniform: plain branches, no BSSY / BSYNC)
        const uint32_t smask = __shfl_sync(0xffffffffu, __ballot_sync(0xffffffffu, dj != sd[eb + lane - 1]) & live, 0);
        const uint32_t emask = __shfl_sync(0xffffffffu, __ballot_sync(0xffffffffu, dj != sd[eb + lane + 1]) & live, 0);
        const uint32_t fmask = emask | (eb + nb == uvalid ? (1u << (nb - 1)) : 0u);      // last edges of segments + of the unit
        // Nodes without in-edges never get a segment: their output rows are the gaps between consecutive destinations
        // (plus the rows before the first and after the last edge's destination).  ZFILL: zeroed here, no memset launch
        // (small graphs, where a launch costs as much as the kernel; on large ones the memset is cheaper than these
        // extra instructions in the epilogue, the busiest role: 1 Mi x 6 Mi forward 1.97 ms against 2.26 ms).
        if (ZFILL) {
          const int prev = (e0 + eb + lane == 0) ? -1 : sd[eb + lane - 1];
          uint32_t gm = __shfl_sync(0xffffffffu, __ballot_sync(0xffffffffu, dj - prev > 1) & live, 0);
          for (; gm; gm &= gm - 1) {
            const int e = eb + __ffs(gm) - 1;
            for (int n = (e0 + e == 0) ? 0 : sd[e - 1] + 1; n < sd[e]; ++n) p.out[(size_t)n * p.ldo + ch] = 0.f;
          }
          if (e0 + eb + nb == p.E)
            for (int n = sd[eb + nb - 1] + 1; n < p.N; ++n) p.out[(size_t)n * p.ldo + ch] = 0.f;
        }
        if (!BWD) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] += bias;
          if (p.z2) {
            float* zp = p.z2 + (size_t)(e0 + eb) * 128 + ch;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j < nb) zp[j * 128] = v[j];
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = swish_m(v[j]);
        } else {
          const float* sp = smS + eb * 128 + ch;
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] *= sp[j * 128];
          float* dp = p.dz1 + (size_t)(e0 + eb) * 128 + ch;
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (j < nb) dp[j * 128] = v[j];
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const bool st = (smask >> j) & 1u;
          sum = st ? v[j] : sum + v[j];
          seg0 = st ? eb + j : seg0;
          if ((fmask >> j) & 1u) {
            // Segment rows [seg0, e] of node sd[e].  It lies inside the unit iff it started here (left) and ends
            // here (right); its in-degree is then e + 1 - seg0, and s_inv[] holds the same IEEE 1.0f / deg that
            // built inv_deg.  Otherwise the partial sum goes to the carry buffer (slot 1: starts here, continues).
            const int e = eb + j;
            const bool right = (emask >> j) & 1u;
            if (left && right) {
              p.out[(size_t)sd[e] * p.ldo + ch] = BWD ? sum : sum * s_inv[e + 1 - seg0];
            } else {
              p.carry[((size_t)unit * 2 + (left ? 1 : 0)) * 128 + ch] = sum;
            }
            left = true;
          }
        }
      }
      if (BWD) {
        __syncwarp();
        if (lane == 0) mbar_arrive(s_empty);
      }
    }
#ifdef MSMP_EW_TICKS
    t_all = clock64() - t_begin;
    if (blockIdx.x == 0 && tid == 0)
      printf("edge_ws<%d> epilogue: tiles %d total %lld wait_acc %lld wait_s %lld\n", (int)BWD, ntile, t_all, t_acc, t_s);
#endif
  } else if (warp >= EW_MMA_WARP) {
    // =========================================================================== MMA warp (+ 3 idle warps)
    reg_dec<40>();
    if (warp == EW_MMA_WARP) {
    constexpr uint32_t IDESC = umma_idesc_tf32(128, 128, 0, 0);
    mbar_wait_backoff(w_full, 0);
    tc_fence_after();
    const uint32_t tm = __shfl_sync(0xffffffffu, tmem, 0);
    const bool leader = elect_one();
    uint32_t s = 0, ph = 0;             // stage and its phase parity
    EW_TICK_DECL(t_full); EW_TICK_DECL(t_ae);
#ifdef MSMP_EW_TICKS
    const long long t_begin = clock64();
#endif
#pragma unroll 1
    for (int i = 0; i < ntile; ++i) {
      const int buf = i & 1;
      if (i >= 2) EW_TIMED(t_ae, mbar_wait_backoff(&acc_empty[buf], ((i >> 1) - 1) & 1));
      tc_fence_after();
      const uint32_t acc = tm + EW_ACC + 128 * buf;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        EW_TIMED(t_full, mbar_wait_backoff(&full[s], ph));
        tc_fence_after();
        const uint32_t b_hi = smem_u32(smB + s * EW_STAGE_BYTES), b_lo = b_hi + IMG_BYTES;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t dbh = umma_desc(b_hi + 32 * k, 16, 1024), dbl = umma_desc(b_lo + 32 * k, 16, 1024);
          const uint32_t a_hi = tm + EW_WHI + 32 * c + 8 * k, a_lo = tm + EW_WLO + 32 * c + 8 * k;
          if (leader) {
            umma_tf32_ts(acc, a_hi, dbh, IDESC, (c | k) ? 1u : 0u);
            umma_tf32_ts(acc, a_lo, dbh, IDESC, 1u);
            umma_tf32_ts(acc, a_hi, dbl, IDESC, 1u);
          }
        }
        if (leader) {
          umma_commit(&empty[s]);
          if (c == 3) umma_commit(&acc_full[buf]);
        }
        __syncwarp();
        if (++s == EW_STAGES) {
          s = 0;
          ph ^= 1;
        }
      }
    }
#ifdef MSMP_EW_TICKS
    if (blockIdx.x == 0 && lane == 0)
      printf("edge_ws<%d> mma: total %lld wait_full %lld wait_acc_empty %lld\n", (int)BWD, clock64() - t_begin, t_full, t_ae);
#endif
    }
  } else {
    // =========================================================================== producer warps
    reg_inc<120>();
    const int pt = tid - EW_PROD_T0;
    // ---- index ring: (src, dst, scale) of a tile's 128 rows + the two neighbouring destinations
    auto idx_issue = [&](int i, int& rs, int& rd, float& rsc) {
      rs = 0;
      rd = -1;
      rsc = 0.f;
      if (i < ntile) {
        const int e0 = (tb + i) * EW_TILE;
        if (pt < EW_TILE) {
          const int e = e0 + pt;
          if (e < p.E) {
            rs = __ldg(p.src + e);
            rd = __ldg(p.dst + e);
            if (BWD) rsc = __ldg(p.inv_deg_e + e);
          }
        } else if (pt == EW_TILE) {
          rd = e0 > 0 ? __ldg(p.dst + e0 - 1) : -2;
        } else if (pt == EW_TILE + 1) {
          rd = e0 + EW_TILE < p.E ? __ldg(p.dst + e0 + EW_TILE) : -2;
        }
      }
    };
    auto idx_store = [&](int i, int rs, int rd, float rsc) {
      const int slot = i & (EW_SLOTS - 1);
      if (pt < EW_TILE) {
        s_src[slot * EW_TILE + pt] = rs;
        s_dst[slot * EW_DST_LD + 1 + pt] = rd;
        if (BWD) s_sc[slot * EW_TILE + pt] = rsc;
      } else if (pt == EW_TILE) {
        s_dst[slot * EW_DST_LD] = rd;
      } else if (pt == EW_TILE + 1) {
        s_dst[slot * EW_DST_LD + 1 + EW_TILE] = rd;
      }
    };
    // ---- work items: fwd 4 per tile (operand chunks); bwd 8 per tile (4 operand chunks, then 4 z1 chunks)
    constexpr int IPT = BWD ? 8 : 4;
    const int c16 = pt & 7;
    auto gather = [&](int i, int j, float4 (&ga)[4], float4 (&gb)[4]) {
      const int c = j & 3;
      const int e0 = (tb + i) * EW_TILE;
      const int valid = min(EW_TILE, p.E - e0);
      const int slot = i & (EW_SLOTS - 1);
      const int col = 32 * c + 4 * c16;
      // Rows past the end of the edge list (last tile only) re-read the last live row: they become extra COLUMNS of
      // D^T that nobody reads, so they need no zero fill, only in-bounds addresses.
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int r = min((pt >> 3) + 32 * q, valid - 1);
        const int d = s_dst[slot * EW_DST_LD + 1 + r];
        if (!BWD || j >= 4) {
          ga[q] = ldg4(p.P + (size_t)d * p.ldpq + col);
          gb[q] = ldg4(p.Q + (size_t)s_src[slot * EW_TILE + r] * p.ldpq + col);
        } else {
          ga[q] = ldg4(p.dagg + (size_t)d * p.lddagg + col);
          gb[q] = ldg4(p.z2 + (size_t)(e0 + r) * 128 + col);
        }
      }
    };
    uint32_t s = 0, ph = 0, nfill = 0;      // ring stage, its phase parity, chunks staged so far
    EW_TICK_DECL(t_empty); EW_TICK_DECL(t_se); EW_TICK_DECL(t_proc);
#ifdef MSMP_EW_TICKS
    const long long t_begin = clock64();
#endif
    auto process = [&](int i, int j, const float4 (&ga)[4], const float4 (&gb)[4]) {
      const int c = j & 3;
      const int e0 = (tb + i) * EW_TILE;
      const int valid = min(EW_TILE, p.E - e0);
      const int slot = i & (EW_SLOTS - 1);
      if (!BWD || j < 4) {
        // ---- operand chunk -> ring stage
        if (nfill >= EW_STAGES) EW_TIMED(t_empty, mbar_wait_backoff(&empty[s], ph ^ 1));
        uint8_t* st = smB + s * EW_STAGE_BYTES;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int r = (pt >> 3) + 32 * q;
          float4 v;
          if (!BWD) {
            const float4 z1 = add4(ga[q], gb[q]);
            v = make_float4(swish_m(z1.x), swish_m(z1.y), swish_m(z1.z), swish_m(z1.w));
          } else {
            const float sc = s_sc[slot * EW_TILE + min(r, valid - 1)];
            v = make_float4(ga[q].x * sc * dswish_m(gb[q].x), ga[q].y * sc * dswish_m(gb[q].y),
                            ga[q].z * sc * dswish_m(gb[q].z), ga[q].w * sc * dswish_m(gb[q].w));
            if (r < valid) st4(p.dz2 + (size_t)(e0 + r) * 128 + 32 * c + 4 * c16, v);
          }
          store_split4(st, st + IMG_BYTES, img_off(r, c16), v);
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&full[s]);
        ++nfill;
        if (++s == EW_STAGES) {
          s = 0;
          ph ^= 1;
        }
      } else {
        // ---- backward: z1 chunk -> a1 (global), sw'(z1) (shared tile for the epilogue)
        if (j == 4 && i >= 1) EW_TIMED(t_se, mbar_wait_backoff(s_empty, (i - 1) & 1));
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int r = (pt >> 3) + 32 * q;
          const float4 z1 = add4(ga[q], gb[q]);
          if (r < valid)
            st4(p.a1 + (size_t)(e0 + r) * 128 + 32 * c + 4 * c16,
                make_float4(swish_m(z1.x), swish_m(z1.y), swish_m(z1.z), swish_m(z1.w)));
          st4(smS + r * 128 + 32 * c + 4 * c16,
              make_float4(dswish_m(z1.x), dswish_m(z1.y), dswish_m(z1.z), dswish_m(z1.w)));
        }
        if (j == 7) {
          __syncwarp();
          if (lane == 0) mbar_arrive(s_full);
        }
      }
    };

    int rs0, rd0, rs1, rd1;
    float rc0, rc1;
    idx_issue(0, rs0, rd0, rc0);
    idx_issue(1, rs1, rd1, rc1);
    idx_store(0, rs0, rd0, rc0);
    producer_sync();
    float4 ga0[4], gb0[4], ga1[4], gb1[4];
    gather(0, 0, ga0, gb0);
#pragma unroll 1
    for (int i = 0; i < ntile; ++i) {
      // publish the indices of tile i + 1 (loaded one tile ago), start loading those of tile i + 2
      idx_store(i + 1, rs1, rd1, rc1);
      idx_issue(i + 2, rs1, rd1, rc1);
      producer_sync();
#pragma unroll
      for (int j = 0; j < IPT; j += 2) {      // two register sets: the next item's loads are in flight during this one
        gather(i, j + 1, ga1, gb1);
        EW_TIMED(t_proc, process(i, j, ga0, gb0));
        if (j + 2 < IPT) gather(i, j + 2, ga0, gb0);
        else if (i + 1 < ntile) gather(i + 1, 0, ga0, gb0);
        EW_TIMED(t_proc, process(i, j + 1, ga1, gb1));
      }
    }
#ifdef MSMP_EW_TICKS
    if (blockIdx.x == 0 && pt == 0)
      printf("edge_ws<%d> producer: total %lld in process %lld (wait_empty %lld wait_s_empty %lld)\n", (int)BWD,
             clock64() - t_begin, t_proc, t_empty, t_se);
#endif
  }
  tc_fence_before();
  __syncthreads();
  if (warp == EW_MMA_WARP) tmem_dealloc(tmem, 512);
}

// ordered fix-up of segments cut by tile boundaries: the tile where a segment starts owns it
__global__ void k_carry_fix_ws(const float* __restrict__ carry, const int* __restrict__ dst,
                               const int* __restrict__ rowptr, const float* __restrict__ scale, float* __restrict__ out,
                               int ldo, int E, int T) {
  const int lane = threadIdx.x & 31;
  const int tile = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (tile >= T) return;
  const int e0 = tile * EW_UNIT;            // ("tile" = unit of EW_UNIT edges here, T = number of units)
  if (e0 >= E) return;
  const int e_last = min(E, e0 + EW_UNIT) - 1;
  const int node = dst[e_last];
  const int seg_begin = rowptr[node], seg_end = rowptr[node + 1];
  if (seg_end <= e_last + 1 || seg_begin < e0) return;
  float4 sum = ldcg4(carry + ((size_t)tile * 2 + 1) * 128 + 4 * lane);
  for (int t = tile + 1; t < T && t * EW_UNIT < seg_end; ++t)
    sum = add4(sum, ldcg4(carry + ((size_t)t * 2 + 0) * 128 + 4 * lane));
  const float sc = scale ? scale[node] : 1.0f;
  st4(out + (size_t)node * ldo + 4 * lane, scale4(sum, sc));
}

static int ws_sm_count() {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return sms;
}

// in-kernel zero fill of the rows of in-degree-0 nodes (instead of a memset in front of the kernel) up to this many tiles
static int ws_zfill_tiles() {
  static const int v = [] { const char* e = getenv("MSMP_EDGE_WS_ZFILL_TILES"); return e ? atoi(e) : 16 * 148; }();
  return v;
}

template <bool BWD, bool ZFILL>
static int launch_edge_ws(const EdgeWsParams& p, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(k_edge_ws<BWD, ZFILL>, cudaFuncAttributeMaxDynamicSharedMemorySize, ew_smem<BWD>()) !=
        cudaSuccess)
      return MSMP_ERR_CUDA;
    attr_set = true;
  }
  // One persistent CTA per SM, but at least `tpc` tiles per CTA: a CTA that walks several tiles overlaps its roles, and
  // on small graphs the SMs left free run the other stream's kernels (MSMP_EDGE_WS_TPC overrides for tuning runs).
  static const int tpc = [] { const char* v = getenv("MSMP_EDGE_WS_TPC"); const int x = v ? atoi(v) : 0; return x > 0 ? x : 1; }();
  const int sms = ws_sm_count();
  int grid = (p.T + tpc - 1) / tpc;
  if (grid > sms) grid = sms;
  k_edge_ws<BWD, ZFILL><<<grid, EW_THREADS, ew_smem<BWD>(), stream>>>(p);
  MSMP_CHECK_LAUNCH();
  return MSMP_OK;
}

}  // namespace msmp

using namespace msmp;

extern "C" size_t msmp_edge_ws_workspace(int E) { return (size_t)msmp_edge_tiles(E) * EW_EH * 2 * 128 * sizeof(float); }

extern "C" int msmp_edge_ws_fwd(const float* P, const float* Q, int ldpq, const int* src, const int* dst,
                                const int* rowptr, const float* inv_deg, const float* W, int w_rs, int w_cs,
                                const float* b2, float* z2, float* agg, int E, int N, int no_isolated, void* workspace,
                                size_t ws_bytes, cudaStream_t stream) {
  if (E < 0 || N < 0 || (ldpq & 3)) return MSMP_ERR_ARG;
  // no_isolated: the caller guarantees that every node has an in-edge, so every output row is written by a segment
  // flush or the carry fix-up and no zero fill is needed at all (large graphs: the 67 MB memset in front of every call
  // of the C4 step was 4 % of its critical path)
  const bool zfill = E > 0 && !no_isolated && msmp_edge_tiles(E) <= ws_zfill_tiles();
  if (!zfill && !(no_isolated && E > 0) && cudaMemsetAsync(agg, 0, (size_t)N * 128 * sizeof(float), stream) != cudaSuccess) return MSMP_ERR_CUDA;
  if (E == 0) return MSMP_OK;
  if (ws_bytes < msmp_edge_ws_workspace(E)) return MSMP_ERR_WORKSPACE;
  EdgeWsParams p{};
  p.P = P; p.Q = Q; p.ldpq = ldpq; p.src = src; p.dst = dst; p.W = W; p.w_rs = w_rs; p.w_cs = w_cs;
  p.b2 = b2; p.z2 = z2; p.out = agg; p.ldo = 128;
  p.carry = reinterpret_cast<float*>(workspace); p.E = E; p.T = msmp_edge_tiles(E); p.N = N;
  int rc = zfill ? launch_edge_ws<false, true>(p, stream) : launch_edge_ws<false, false>(p, stream);
  if (rc) return rc;
  k_carry_fix_ws<<<(p.T * EW_EH + 7) / 8, 256, 0, stream>>>(p.carry, dst, rowptr, inv_deg, agg, 128, E, p.T * EW_EH);
  MSMP_CHECK_LAUNCH();
  return MSMP_OK;
}

extern "C" int msmp_edge_ws_bwd(const float* P, const float* Q, int ldpq, const int* src, const int* dst,
                                const int* rowptr, const float* inv_deg_e, const float* W, int w_rs, int w_cs,
                                const float* z2, const float* dagg, int lddagg, float* dz2, float* a1, float* dz1,
                                float* dP, int lddp, int E, int N, int no_isolated, void* workspace, size_t ws_bytes,
                                cudaStream_t stream) {
  if (E < 0 || N < 0 || (ldpq & 3) || (lddagg & 3) || (lddp & 3)) return MSMP_ERR_ARG;
  const bool zfill = E > 0 && !no_isolated && msmp_edge_tiles(E) <= ws_zfill_tiles();
  if (!zfill && !(no_isolated && E > 0) &&
      cudaMemset2DAsync(dP, (size_t)lddp * sizeof(float), 0, 128 * sizeof(float), N, stream) != cudaSuccess)
    return MSMP_ERR_CUDA;
  if (E == 0) return MSMP_OK;
  if (ws_bytes < msmp_edge_ws_workspace(E)) return MSMP_ERR_WORKSPACE;
  EdgeWsParams p{};
  p.P = P; p.Q = Q; p.ldpq = ldpq; p.src = src; p.dst = dst; p.inv_deg_e = inv_deg_e;
  p.W = W; p.w_rs = w_rs; p.w_cs = w_cs; p.z2 = const_cast<float*>(z2); p.dagg = dagg; p.lddagg = lddagg;
  p.dz2 = dz2; p.a1 = a1; p.dz1 = dz1; p.out = dP; p.ldo = lddp;
  p.carry = reinterpret_cast<float*>(workspace); p.E = E; p.T = msmp_edge_tiles(E); p.N = N;
  int rc = zfill ? launch_edge_ws<true, true>(p, stream) : launch_edge_ws<true, false>(p, stream);
  if (rc) return rc;
  k_carry_fix_ws<<<(p.T * EW_EH + 7) / 8, 256, 0, stream>>>(p.carry, dst, rowptr, nullptr, dP, lddp, E, p.T * EW_EH);
  MSMP_CHECK_LAUNCH();
  return MSMP_OK;
}
