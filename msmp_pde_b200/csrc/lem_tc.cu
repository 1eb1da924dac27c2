// Persistent LEM recurrence on the tensor cores (replaces lem_cuda.forward / lem_cuda.backward,
// experiments/models_gnn.py:290-292,300).  The recurrence is independent per node, so a tile of 128 nodes walks all
// T time steps inside ONE launch and its state never leaves the chip.
//
// A tile is owned by a CLUSTER OF TWO CTAs (two SMs): CTA `rank` computes hidden channels [64 rank, 64 rank + 64) of
// every gate / state for all 128 rows.  At the reference's graph sizes (6400 nodes = 50 tiles) this puts the
// recurrence on 100 of the 148 SMs instead of 50, halves the per-SM epilogue work and halves the weight bytes each
// SM streams per step.  Both CTAs keep the FULL state operand (K = 128) in their own shared memory: after a gate
// epilogue each thread writes its values into its own state tile and, through distributed shared memory
// (st.shared::cluster), into the peer's; two mbarriers per CTA order the exchange (see Xchg).
//
// forward, per step t.  The input part of both affine maps is hoisted out of the recurrence:
// pre[t][n][0:512] = [b | bz] + I_t [W_in | Wz_in]^T (k_lem_inproj, memory bound, exact fp32), then per CTA
//   G[128 x 3*64] = y_{t-1} W_h^T (own columns)   12 half weight chunks (3 gates x 4 k-chunks), TMEM columns 0..191
//   gate_z      : a = dt sig(G0 + pre), b = dt sig(G1 + pre), zc = tanh(G2 + pre), z_t = (1-b) z_{t-1} + b zc
//   L[128 x 64]   = z_t Wz_h^T (own columns)      4 half weight chunks, TMEM columns 192..255
//   gate_y      : tL = tanh(L + pre), y_t = (1-a) y_{t-1} + a tL
// The weights (pre-split tf32 hi | lo, pre-swizzled [128 n x 32 k] images) are streamed from L2: a CTA needs rows
// [64 rank, +64) of every image (8 KiB hi + 8 KiB lo per chunk, 256 KiB per step).  Four dedicated loader warps run
// ahead of the MMA-issuing thread through a 6-stage ring, across phase boundaries: the weights do not depend on the
// step, so the first chunks of the next GEMM are already in shared memory while the gate epilogue runs.  One thread
// issues the MMAs (3xTF32: hi*hi + lo*hi + hi*lo), eight warps run the gate epilogues.
//
// backward, per step t = t_end-1 .. t_begin (dy, dz carried in global scratch, owned row-wise by the same thread):
//   bwd_y : d = dy + gY[t]; dL = d a (1-tL^2); dG0 = d (tL - y_{t-1}) a (1 - a/dt); dy = d (1-a)
//   acc1  = dL Wz[:, :128]  (own 64 columns)      4 half chunks
//   bwd_z : d = dz + gZ[t] + acc1; dG1 = d (zc - z_{t-1}) b (1 - b/dt); dG2 = d b (1-zc^2); dz = d (1-b)
//   acc2  = [dG1 | dG2 | dG0] W[:, :128]          3 x 4 half chunks (the state tile is restaged per 128-row k-block)
//   dy   += acc2
// dG [T,N,384] and dL [T,N,128] are written for the four weight-gradient GEMMs (msmp_linear_wgrad_tc).
//
// Memory access.  The gate epilogues own one node row per thread (a TMEM lane), so row-major global arrays would be
// touched 16 bytes per thread at a 512-byte stride (32 lines per warp request; measured: 77 us per step).  Every array
// private to the recurrence (pre, gates, the y/z history used for y_{t-1}/z_{t-1}, the carried dy/dz, the dG0/dG2
// scratch) is therefore kept LANE-MAJOR: element (row n, channel c) of a C-channel array lives at
//     ((n / 32) * C + c) * 32 + n % 32
// so the 32 lanes of a warp (32 consecutive rows) read/write one contiguous 128-byte line per channel.  Arrays
// that other kernels consume row-major (Y, Z, dL, dG) are written by a cooperative, coalesced copy-out of the CTA's
// own 64 columns of the state tile image (hi + lo reconstructs the fp32 value exactly); it overlaps the next GEMM.
#include "umma.cuh"
#include "msmp_b200.h"

namespace msmp {

constexpr int LT_A_BYTES = 4 * 2 * IMG_BYTES;      // state tile: 4 k-chunks x (hi | lo) = 128 KiB
constexpr int LT_STAGE_BYTES = IMG_BYTES;          // half a weight chunk: 64 rows hi (8 KiB) | 64 rows lo (8 KiB)
constexpr int LT_HALF = IMG_BYTES / 2;
constexpr int LT_STAGES = 6;
constexpr int LT_SMEM = LT_A_BYTES + LT_STAGES * LT_STAGE_BYTES + 1024 + 256;
constexpr int LT_EPI = 256;                        // warps 0..7: gate epilogues (thread 0 also issues the MMAs)
constexpr int LT_LOADERS = 128;                    // warps 8..11: weight ring producers
constexpr int LT_THREADS = LT_EPI + LT_LOADERS;

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- cluster plumbing ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// shared::cluster address of `saddr` (a shared::cta address of this CTA's window) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster4(uint32_t caddr, float4 v) {
  asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(caddr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}
__device__ __forceinline__ void st_shared4(uint32_t saddr, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t caddr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(caddr) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait_cluster(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait_cluster(bar, parity)) {
    __nanosleep(20);
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }      // the 8 epilogue warps
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

// State-tile exchange between the two CTAs of a cluster.  A CTA's 64 columns are k-chunks 2 rank, 2 rank + 1 of the
// tile: one contiguous 64 KiB block (hi | lo images), written locally by the gate epilogue and then pushed into the
// peer's tile by ONE bulk shared-to-shared-cluster copy.  Each CTA owns two mbarriers:
//   xfree: the PEER arrives when its GEMM (which reads the peer's whole tile) has completed: this CTA may overwrite
//          its own columns (the previous push out of them has been consumed) and push into the peer's tile again;
//   xfull: armed by this CTA with expect_tx = 64 KiB, completed by the peer's bulk copy landing in this CTA's tile.
// A GEMM multiplies the CTA's own k-chunks first and waits for xfull only before the peer's k-chunks, so the push
// overlaps half of the MMAs.  Every thread keeps its own phase counters, so no bookkeeping is communicated.
struct Xchg {
  uint64_t* xfull;
  uint64_t* xfree;
  uint32_t peer_xfull;
  uint32_t peer_xfree;
  uint32_t nfull;
  uint32_t nfree;
  uint32_t pending_free;      // GEMMs issued since the launch started (nothing to wait for before the first one)
};

// loader thread lt (0..127): copy rows [64 rank, +64) of weight chunk image `src` (hi | lo, 4096 floats each) into stage i
struct Ring {
  uint8_t* smB;        // LT_STAGES stages
  uint64_t* bfull;     // [LT_STAGES], LT_LOADERS arrivals
  uint64_t* bfree;     // [LT_STAGES], one arrival (tcgen05.commit)
};

__device__ __forceinline__ void ring_load(const Ring& rg, uint32_t i, const float* src, int rank, int lt) {
  const uint32_t s = i % LT_STAGES, use = i / LT_STAGES;
  if (use > 0) mbar_wait_warp(&rg.bfree[s], (use - 1) & 1);       // MMAs that read this stage are complete
  // 8 x 16 B per thread with cp.async: nothing is held in registers, so all LT_STAGES stages (96 KiB) can be in
  // flight per SM -- the stream is latency bound (LDG + STS with one chunk in flight reached 10 B/clk per SM)
  const float4* g = reinterpret_cast<const float4*>(src) + rank * (LT_HALF / 16) + lt;
  const uint32_t d = smem_u32(rg.smB + s * LT_STAGE_BYTES) + 16u * (uint32_t)lt;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d + 16u * LT_LOADERS * q), "l"(g + q * LT_LOADERS) : "memory");
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d + LT_HALF + 16u * LT_LOADERS * q),
                 "l"(g + (IMG_BYTES / 16) + q * LT_LOADERS)
                 : "memory");
  }
  // the barrier receives this thread's arrival when all of its copies above have landed
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&rg.bfull[s])) : "memory");
}

// MMA-issuing thread: chunk i of the ring times state-tile chunk at a_img -> 64 TMEM columns at tmem_d
__device__ __forceinline__ void ring_mma(const Ring& rg, uint32_t i, uint32_t a_img, uint32_t tmem_d, bool accumulate) {
  const uint32_t s = i % LT_STAGES, use = i / LT_STAGES;
  mbar_wait(&rg.bfull[s], use & 1);
  fence_proxy_async();         // cp.async wrote the stage through the generic proxy
  tc_fence_after();
  constexpr uint32_t IDESC = umma_idesc_tf32(128, 64, 0, 0);
  const uint32_t a_hi = a_img, a_lo = a_img + IMG_BYTES;
  const uint32_t b_hi = smem_u32(rg.smB + s * LT_STAGE_BYTES), b_lo = b_hi + LT_HALF;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const uint64_t dah = umma_desc(a_hi + 32 * k, 16, 1024), dal = umma_desc(a_lo + 32 * k, 16, 1024);
    const uint64_t dbh = umma_desc(b_hi + 32 * k, 16, 1024), dbl = umma_desc(b_lo + 32 * k, 16, 1024);
    umma_tf32(tmem_d, dah, dbh, IDESC, (accumulate || k) ? 1u : 0u);
    umma_tf32(tmem_d, dal, dbh, IDESC, 1u);
    umma_tf32(tmem_d, dah, dbl, IDESC, 1u);
  }
  umma_commit(&rg.bfree[s]);
}

// Per-thread context of the epilogue warps.
struct Epi {
  Ring rg;
  Xchg x;
  uint64_t* acc;
  uint32_t nacc;       // accumulator barrier phases consumed
  uint32_t nchunk;     // ring chunks consumed (MMA thread)
  uint32_t smA;        // shared::cta address of this CTA's state tile
  uint32_t smA_peer;   // shared::cluster address of the peer's state tile
  uint32_t tmem;
};

constexpr uint32_t LT_SLICE_BYTES = 2 * 2 * IMG_BYTES;      // two k-chunks x (hi | lo) = 64 KiB

// Thread 0: issue one GEMM phase of `ngroups` 64-column accumulator blocks (TMEM columns dcol + 64 g).  Ring chunk
// order (mirrored by the loader warps, see ring_schedule): for every block the CTA's own two k-chunks, then -- after
// the peer's half of the state tile has landed (need_full) -- for every block the peer's two k-chunks.  The first
// chunk of a block overwrites the accumulator unless acc_first.  Returns right after the commit.
__device__ __forceinline__ void gemm_issue(Epi& e, int rank, bool need_full, uint32_t ngroups, uint32_t dcol, bool acc_first) {
  if (need_full) mbar_expect_tx(e.x.xfull, LT_SLICE_BYTES);       // this CTA's single arrival of the phase + the byte count
  fence_proxy_async_all();
  tc_fence_after();
  for (uint32_t g = 0; g < ngroups; ++g)
    for (uint32_t kk = 0; kk < 2; ++kk)
      ring_mma(e.rg, e.nchunk++, e.smA + (2 * rank + kk) * 2 * IMG_BYTES, e.tmem + dcol + 64 * g, acc_first || kk != 0);
  if (need_full) {
    mbar_wait_cluster(e.x.xfull, e.x.nfull & 1);
    ++e.x.nfull;
    tc_fence_after();
  }
  for (uint32_t g = 0; g < ngroups; ++g)
    for (uint32_t kk = 0; kk < 2; ++kk)
      ring_mma(e.rg, e.nchunk++, e.smA + (2 * (rank ^ 1) + kk) * 2 * IMG_BYTES, e.tmem + dcol + 64 * g, true);
  umma_commit(e.acc);
}

// Loader warps: the ring order of one GEMM phase over weight chunks w0 + 4 g + kc (g < ngroups, kc = k-chunk)
template <class F>
__device__ __forceinline__ void ring_schedule(int rank, int w0, int ngroups, F&& load) {
  for (int g = 0; g < ngroups; ++g)
    for (int kk = 0; kk < 2; ++kk) load(w0 + 4 * g + 2 * rank + kk);
  for (int g = 0; g < ngroups; ++g)
    for (int kk = 0; kk < 2; ++kk) load(w0 + 4 * g + 2 * (rank ^ 1) + kk);
}

// All epilogue threads: wait for the GEMM issued last; thread 0 also tells the peer that this CTA's tile is free again.
__device__ __forceinline__ void gemm_wait(Epi& e) {
  if (threadIdx.x == 0) {
    mbar_wait(e.acc, e.nacc & 1);
    mbar_arrive_remote(e.x.peer_xfree);
  }
  ++e.nacc;
  ++e.x.pending_free;
  __syncwarp();
  epi_bar();
  tc_fence_after();
}

// All epilogue threads, before the first store into the state tiles after a GEMM: the peer's GEMM has completed too.
__device__ __forceinline__ void wait_peer_free(Epi& e) {
  if (e.x.pending_free == 0) return;
  e.x.pending_free = 0;      // (one wait per GEMM round; gemm_wait sets it again)
  if ((threadIdx.x & 31) == 0) mbar_wait_cluster(e.x.xfree, e.x.nfree & 1);
  ++e.x.nfree;
  __syncwarp();
}

// All epilogue threads, after the last store of a state-tile refill: make it visible to the async proxy, then thread 0
// pushes the CTA's 64 columns into the peer's tile (completion is counted on the peer's xfull barrier).
__device__ __forceinline__ void publish(Epi& e, int rank) {
  fence_proxy_async_all();
  tc_fence_before();
  epi_bar();
  if (threadIdx.x == 0) {
    const uint32_t off = (uint32_t)(2 * rank) * 2 * IMG_BYTES;
    asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     e.smA_peer + off),
                 "r"(e.smA + off), "r"(LT_SLICE_BYTES), "r"(e.x.peer_xfull)
                 : "memory");
  }
}

// write 4 consecutive values of row r, columns col..col+3 (col % 4 == 0, col < 128) into this CTA's state tile image
__device__ __forceinline__ void state_store4(const Epi& e, int r, int col, float4 v) {
  const uint32_t off = (uint32_t)(col >> 5) * (2 * IMG_BYTES) + img_off(r, (col & 31) >> 2);
  float4 h, l;
  split_tf32(v.x, h.x, l.x);
  split_tf32(v.y, h.y, l.y);
  split_tf32(v.z, h.z, l.z);
  split_tf32(v.w, h.w, l.w);
  st_shared4(e.smA + off, h);
  st_shared4(e.smA + off + IMG_BYTES, l);
}

// clock64() phase stamps of CTA 0 at step 2 (scripts/lem_ticks.py); compiled in only with -DMSMP_LEM_TICKS.
#ifdef MSMP_LEM_TICKS
__device__ long long g_lem_dbg[64];
#define LEM_TICK(i) do { if (blockIdx.x == 0 && threadIdx.x == 0 && t == 2) g_lem_dbg[i] = clock64(); } while (0)
#else
#define LEM_TICK(i) do { } while (0)
#endif

// L2 prefetch of `nch` consecutive lane-major channel lines (128 B each) of row-tile gt, starting at channel c_begin
__device__ __forceinline__ void prefetch_lm(const float* base, size_t gt, int C, int c_begin, int nch, int lane) {
  for (int c = c_begin + lane; c < c_begin + nch; c += 32)
    asm volatile("prefetch.global.L2 [%0];" ::"l"(base + (gt * C + c) * 32));
}

// lane-major address of (row-tile gt, channel c, lane l) in a C-channel array
__device__ __forceinline__ size_t lm(size_t gt, int C, int c, int l) { return (gt * C + c) * 32 + l; }

// pre (lane-major, 512 channels) = [bias | bias_z] + inp[:, 0:ninp] * [Wt_in | Wzt_in]; rows = T * Npad
__global__ void __launch_bounds__(256) k_lem_inproj(const float* __restrict__ inp, const float* __restrict__ Wt_in,
                                                    const float* __restrict__ Wzt_in, const float* __restrict__ bias,
                                                    const float* __restrict__ bias_z, float* __restrict__ pre, int T,
                                                    int N, int Npad, int ninp) {
  // one warp = a quarter (128 channels) of one 32-row tile; x values live in registers
  const int lane = threadIdx.x & 31;
  const size_t wid = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);      // over 4 * T * Npad / 32
  const size_t ntiles = (size_t)T * (Npad / 32);
  const size_t tile = wid >> 2;
  const int cq = (int)(wid & 3) * 128;
  if (tile >= ntiles) return;
  const int t = (int)(tile / (Npad / 32));
  const int n = (int)(tile % (Npad / 32)) * 32 + lane;
  float x[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) x[q] = (q < ninp && n < N) ? __ldg(inp + ((size_t)t * N + n) * 32 + q) : 0.f;
  float* o = pre + tile * 512 * 32 + lane;
  const bool g = cq < 384;
  const float* Wb = g ? Wt_in + cq : Wzt_in;
  const float* bb = g ? bias + cq : bias_z;
  const int ldw = g ? 384 : 128;
#pragma unroll 4
  for (int c = 0; c < 128; ++c) {
    float acc = __ldg(bb + c);
#pragma unroll
    for (int q = 0; q < 8; ++q)
      if (q < ninp) acc = fmaf(x[q], __ldg(Wb + (size_t)q * ldw + c), acc);
    o[(size_t)(cq + c) * 32] = acc;
  }
}

// cooperative coalesced copy of this CTA's 64 columns [64 rank, +64) of the state tile (hi + lo) to a row-major array:
// dst[(row0 + r) * ld + 64 rank + c], c < 64  (256 epilogue threads)
__device__ __forceinline__ void image_to_global(uint32_t smA, int rank, float* dst, int ld, int row0, int N) {
  const int tid = threadIdx.x;
#pragma unroll 4
  for (int i = 0; i < 8; ++i) {
    const int idx = tid + 256 * i;
    const int rr = idx >> 4, c4 = idx & 15;
    const uint32_t a = smA + (uint32_t)(2 * rank + (c4 >> 3)) * (2 * IMG_BYTES) + img_off(rr, c4 & 7);
    float4 h, l;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(h.x), "=f"(h.y), "=f"(h.z), "=f"(h.w) : "r"(a));
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(l.x), "=f"(l.y), "=f"(l.z), "=f"(l.w) : "r"(a + IMG_BYTES));
    if (row0 + rr < N) st4(dst + (size_t)(row0 + rr) * ld + 64 * rank + 4 * c4, add4(h, l));
  }
}

// common prologue: barriers, TMEM, cluster addresses.  Returns false for the loader warps (after they finished).
struct LemSmem {
  uint8_t* smA;
  uint8_t* smB;
  uint64_t* bars;      // bfull[6], bfree[6], acc, xfull, xfree
  uint32_t* tmem_slot;
};

__device__ __forceinline__ LemSmem lem_smem(uint8_t* smem_raw) {
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  LemSmem m;
  m.smA = smem;
  m.smB = smem + LT_A_BYTES;
  m.bars = reinterpret_cast<uint64_t*>(m.smB + LT_STAGES * LT_STAGE_BYTES);
  m.tmem_slot = reinterpret_cast<uint32_t*>(m.bars + 16);
  return m;
}

__device__ __forceinline__ void lem_init(const LemSmem& m, uint32_t tmem_cols) {
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) tmem_alloc(m.tmem_slot, tmem_cols);
  if (tid == 32) {
    for (int i = 0; i < LT_STAGES; ++i) {
      mbar_init(&m.bars[i], LT_LOADERS);
      mbar_init(&m.bars[LT_STAGES + i], 1);
    }
    mbar_init(&m.bars[2 * LT_STAGES], 1);
    mbar_init(&m.bars[2 * LT_STAGES + 1], 1);
    mbar_init(&m.bars[2 * LT_STAGES + 2], 1);
    fence_barrier_init();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();      // the peer's barriers exist before anything is signalled remotely
  tc_fence_after();
}

__device__ __forceinline__ Epi lem_epi(const LemSmem& m, int rank) {
  Epi e;
  e.rg = Ring{m.smB, &m.bars[0], &m.bars[LT_STAGES]};
  e.acc = &m.bars[2 * LT_STAGES];
  e.x.xfull = &m.bars[2 * LT_STAGES + 1];
  e.x.xfree = &m.bars[2 * LT_STAGES + 2];
  e.x.peer_xfull = mapa_u32(smem_u32(e.x.xfull), rank ^ 1);
  e.x.peer_xfree = mapa_u32(smem_u32(e.x.xfree), rank ^ 1);
  e.x.nfull = e.x.nfree = e.x.pending_free = 0;
  e.nacc = e.nchunk = 0;
  e.smA = smem_u32(m.smA);
  e.smA_peer = mapa_u32(e.smA, rank ^ 1);
  e.tmem = *m.tmem_slot;
  return e;
}

struct LemFwdParams {
  const float* pre;      // lane-major [T][Npad/32][512][32]
  const float* Wimg;     // images of Wt[:128]  [128 x 384]: [3 ntiles][4 chunks][2][4096]
  const float* Wzimg;    // images of Wzt[:128] [128 x 128]: [1][4][2][4096]
  float* Y;              // row-major [T+1][N][128]  (Y[0] = y0 on entry)
  float* Z;              // row-major [T+1][N][128]  (Z[0] = z0 on entry)
  float* Yt;             // lane-major [T+1][Npad/32][128][32]  (Yt[0] = y0 on entry)
  float* Zt;             // lane-major [T+1][Npad/32][128][32]
  float* gates;          // lane-major [T][Npad/32][512][32]  a | b | zc | tL
  float dt;
  int T; int N; int Npad;
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(LT_THREADS, 1) k_lem_fwd_tc(const LemFwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  const LemSmem m = lem_smem(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int rank = (int)cluster_ctarank();
  const int tile = blockIdx.x >> 1;
  const int row0 = tile * 128;
  const size_t plane = (size_t)p.N * 128;
  const size_t ntile = p.Npad / 32;
  lem_init(m, 256);

  if (warp >= 8) {
    // ---- weight ring producers: 12 G chunks + 4 L chunks per step, running ahead of the MMA thread
    const Ring rg{m.smB, &m.bars[0], &m.bars[LT_STAGES]};
    const int lt = tid - LT_EPI;
    uint32_t n = 0;
    for (int t = 0; t < p.T; ++t) {
      ring_schedule(rank, 0, 3, [&](int c) { ring_load(rg, n++, p.Wimg + (size_t)c * 2 * (IMG_BYTES / 4), rank, lt); });
      ring_schedule(rank, 0, 1, [&](int c) { ring_load(rg, n++, p.Wzimg + (size_t)c * 2 * (IMG_BYTES / 4), rank, lt); });
    }
  } else {
    Epi e = lem_epi(m, rank);
    // epilogue ownership: thread = row r (TMEM lane), 32 channels [c0, c0+32) of this CTA's 64; gt = its 32-row tile
    const int r = 32 * (warp & 3) + lane;
    const int ch0 = 32 * (warp >> 2);           // column inside the CTA's 64-wide accumulator blocks
    const int c0 = 64 * rank + ch0;             // global hidden channel
    const size_t gt = (size_t)tile * 4 + (warp & 3);
    const uint32_t tlane = (uint32_t)(32 * (warp & 3)) << 16;

    // y_{-1}: both CTAs load the full tile from the row-major Y[0]
    for (int i = 0; i < 16; ++i) {
      const int idx = tid + 256 * i;
      const int rr = idx >> 5, c4 = idx & 31;
      const int g = row0 + rr;
      float4 v = (g < p.N) ? ldg4(p.Y + (size_t)g * 128 + 4 * c4) : zero4();
      const uint32_t off = (uint32_t)(c4 >> 3) * (2 * IMG_BYTES) + img_off(rr, c4 & 7);
      float4 h, l;
      split_tf32(v.x, h.x, l.x);
      split_tf32(v.y, h.y, l.y);
      split_tf32(v.z, h.z, l.z);
      split_tf32(v.w, h.w, l.w);
      st_shared4(e.smA + off, h);
      st_shared4(e.smA + off + IMG_BYTES, l);
    }
    fence_proxy_async_all();
    tc_fence_before();
    epi_bar();

    for (int t = 0; t < p.T; ++t) {
      const float* pre_t = p.pre + ((size_t)t * ntile) * 512 * 32;
      float* g_t = p.gates + ((size_t)t * ntile) * 512 * 32;
      // ---- G = y W_h^T (own columns): 3 gates x 4 chunks -> TMEM columns 0..191
      LEM_TICK(0);
      if (tid == 0) gemm_issue(e, rank, t > 0, 3, 0, false);
      // pull this step's input-projection lines (HBM) into L2 while the GEMM runs
#pragma unroll
      for (int q = 0; q < 4; ++q) prefetch_lm(pre_t, gt, 512, 128 * q + c0, 32, lane);
      gemm_wait(e);
      LEM_TICK(1);
      // ---- gate_z
      const float* zprev = p.Zt + ((size_t)t * ntile) * 128 * 32;
      float* znext = p.Zt + ((size_t)(t + 1) * ntile) * 128 * 32;
      {
        float v0[32], v1[32], v2[32];
        tmem_ld32(e.tmem + tlane + (uint32_t)ch0, v0);
        tmem_ld32(e.tmem + tlane + (uint32_t)(64 + ch0), v1);
        tmem_ld32(e.tmem + tlane + (uint32_t)(128 + ch0), v2);
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          // batch the 32 loads of 8 channels before any dependent math (the epilogue is latency bound otherwise)
          float p0[8], p1[8], p2[8], zp[8], zn[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const int c = c0 + j + q;
            p0[q] = __ldg(pre_t + lm(gt, 512, c, lane));
            p1[q] = __ldg(pre_t + lm(gt, 512, 128 + c, lane));
            p2[q] = __ldg(pre_t + lm(gt, 512, 256 + c, lane));
            zp[q] = __ldcg(zprev + lm(gt, 128, c, lane));
          }
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const int c = c0 + j + q;
            const float a = p.dt * sigmoidf_(v0[j + q] + p0[q]);
            const float b = p.dt * sigmoidf_(v1[j + q] + p1[q]);
            const float zc = tanh_acc(v2[j + q] + p2[q]);
            zn[q] = (1.f - b) * zp[q] + b * zc;
            g_t[lm(gt, 512, c, lane)] = a;
            g_t[lm(gt, 512, 128 + c, lane)] = b;
            g_t[lm(gt, 512, 256 + c, lane)] = zc;
            znext[lm(gt, 128, c, lane)] = zn[q];
          }
          if (j == 0) { LEM_TICK(10); wait_peer_free(e); LEM_TICK(11); }
          state_store4(e, r, c0 + j, make_float4(zn[0], zn[1], zn[2], zn[3]));    // z_t: A operand of the L GEMM
          state_store4(e, r, c0 + j + 4, make_float4(zn[4], zn[5], zn[6], zn[7]));
        }
      }
      LEM_TICK(12);
      publish(e, rank);
      LEM_TICK(2);
      // ---- L = z Wz_h^T (own columns): 4 chunks -> TMEM columns 192..255; the Z copy-out overlaps it
      if (tid == 0) gemm_issue(e, rank, true, 1, 192, false);
      LEM_TICK(13);
      __syncwarp();
      image_to_global(e.smA, rank, p.Z + (size_t)(t + 1) * plane, 128, row0, p.N);
      LEM_TICK(3);
      gemm_wait(e);
      LEM_TICK(4);
      // ---- gate_y
      const float* yprev = p.Yt + ((size_t)t * ntile) * 128 * 32;
      float* ynext = p.Yt + ((size_t)(t + 1) * ntile) * 128 * 32;
      {
        float v[32];
        tmem_ld32(e.tmem + tlane + (uint32_t)(192 + ch0), v);
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          float pz[8], av[8], yp[8], yn[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const int c = c0 + j + q;
            pz[q] = __ldg(pre_t + lm(gt, 512, 384 + c, lane));
            av[q] = __ldcg(g_t + lm(gt, 512, c, lane));
            yp[q] = __ldcg(yprev + lm(gt, 128, c, lane));
          }
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const int c = c0 + j + q;
            const float tl = tanh_acc(v[j + q] + pz[q]);
            yn[q] = (1.f - av[q]) * yp[q] + av[q] * tl;
            g_t[lm(gt, 512, 384 + c, lane)] = tl;
            ynext[lm(gt, 128, c, lane)] = yn[q];
          }
          if (j == 0) wait_peer_free(e);
          state_store4(e, r, c0 + j, make_float4(yn[0], yn[1], yn[2], yn[3]));    // y_t: A operand of the next G GEMM
          state_store4(e, r, c0 + j + 4, make_float4(yn[4], yn[5], yn[6], yn[7]));
        }
      }
      publish(e, rank);
      LEM_TICK(5);
      image_to_global(e.smA, rank, p.Y + (size_t)(t + 1) * plane, 128, row0, p.N);
      LEM_TICK(6);
    }
    if (tid == 0 && p.T > 0) {      // the peer's last push (y_T) has no GEMM to consume it: drain it before leaving
      mbar_expect_tx(e.x.xfull, LT_SLICE_BYTES);
      mbar_wait_cluster(e.x.xfull, e.x.nfull & 1);
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();      // no CTA leaves while its peer may still store into its shared memory
  if (warp == 0) tmem_dealloc(*m.tmem_slot, 256);
}

// ------------------------------------------------------------------------------------------------ backward
struct LemBwdParams {
  const float* Wzh_img;  // images of Wt := Wz[:, :128]  ([K = n][N = k]) : [1][4][2][4096]
  const float* Wh_img;   // images of Wt := W[:, :128]   ([K = 384][N = 128]) : [1][12][2][4096]
  const float* Yt;       // lane-major [T+1][Npad/32][128][32]
  const float* Zt;       // lane-major [T+1][Npad/32][128][32]
  const float* gates;    // lane-major [T][Npad/32][512][32]
  const float* gYt;      // lane-major external gradients: [T][..] or, if g_last_only, one slab applied at t = T-1
  const float* gZt;      // (either may be NULL)
  int g_last_only;
  float* dG;             // row-major [T][N][384]
  float* dL;             // row-major [T][N][128]
  float* dyt;            // lane-major [Npad/32][128][32] carried gradient (zero on entry; d/dy0 on exit)
  float* dzt;            // lane-major
  float* s0;             // lane-major scratch [Npad/32][128][32]  (dG0 of the current step)
  float* s2;             // lane-major scratch                      (dG2 of the current step)
  float dt;
  int T; int N; int Npad;
  int t_begin; int t_end;   // this launch walks t = t_end-1 .. t_begin (the carried dy/dz live in dyt/dzt between launches)
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(LT_THREADS, 1) k_lem_bwd_tc(const LemBwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  const LemSmem m = lem_smem(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int rank = (int)cluster_ctarank();
  const int tile = blockIdx.x >> 1;
  const int row0 = tile * 128;
  const size_t ntile = p.Npad / 32;
  lem_init(m, 128);

  if (warp >= 8) {
    // ---- weight ring producers: per step Wz chunks 0..3, then W chunks 4..7, 8..11, 0..3
    const Ring rg{m.smB, &m.bars[0], &m.bars[LT_STAGES]};
    const int lt = tid - LT_EPI;
    uint32_t n = 0;
    for (int t = p.t_end - 1; t >= p.t_begin; --t) {
      ring_schedule(rank, 0, 1, [&](int c) { ring_load(rg, n++, p.Wzh_img + (size_t)c * 2 * (IMG_BYTES / 4), rank, lt); });
      for (int w0 = 4; w0 != 16; w0 += 4)      // dG1 (k-rows 128..255), dG2 (256..383), dG0 (0..127)
        ring_schedule(rank, w0 % 12, 1, [&](int c) { ring_load(rg, n++, p.Wh_img + (size_t)c * 2 * (IMG_BYTES / 4), rank, lt); });
    }
  } else {
    Epi e = lem_epi(m, rank);
    const int r = 32 * (warp & 3) + lane;
    const int ch0 = 32 * (warp >> 2);
    const int c0 = 64 * rank + ch0;
    const size_t gt = (size_t)tile * 4 + (warp & 3);
    const uint32_t tlane = (uint32_t)(32 * (warp & 3)) << 16;
    const float inv_dt = 1.0f / p.dt;

    // stage this thread's 32 channels of a lane-major 128-channel scratch slab into both state tiles
    auto stage_lm = [&](const float* slab) {
      float g[32];
#pragma unroll
      for (int q = 0; q < 32; ++q) g[q] = __ldcg(slab + lm(gt, 128, c0 + q, lane));
      wait_peer_free(e);
#pragma unroll
      for (int q = 0; q < 32; q += 4) state_store4(e, r, c0 + q, make_float4(g[q], g[q + 1], g[q + 2], g[q + 3]));
    };

    for (int t = p.t_end - 1; t >= p.t_begin; --t) {
      const float* g_t = p.gates + ((size_t)t * ntile) * 512 * 32;
      const float* yprev = p.Yt + ((size_t)t * ntile) * 128 * 32;
      const float* zprev = p.Zt + ((size_t)t * ntile) * 128 * 32;
      const bool ext = !p.g_last_only || t == p.T - 1;
      const float* gy = (p.gYt && ext) ? p.gYt + (p.g_last_only ? 0 : (size_t)t * ntile * 128 * 32) : nullptr;
      const float* gz = (p.gZt && ext) ? p.gZt + (p.g_last_only ? 0 : (size_t)t * ntile * 128 * 32) : nullptr;
      float* dG_t = p.dG + (size_t)t * p.N * 384;
      if (t > p.t_begin) {      // next step's saved activations (written by the forward pass, now in HBM) -> L2
        const float* g_n = p.gates + ((size_t)(t - 1) * ntile) * 512 * 32;
#pragma unroll
        for (int q = 0; q < 4; ++q) prefetch_lm(g_n, gt, 512, 128 * q + c0, 32, lane);
        prefetch_lm(p.Yt + ((size_t)(t - 1) * ntile) * 128 * 32, gt, 128, c0, 32, lane);
        prefetch_lm(p.Zt + ((size_t)(t - 1) * ntile) * 128 * 32, gt, 128, c0, 32, lane);
      }
      // ---- bwd_y : dL -> state tiles, dG0 -> s0, dy <- d (1 - a)
#pragma unroll 1
      for (int j = 0; j < 32; j += 8) {
        float dv[8], av[8], tv[8], yv[8], dl[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int c = c0 + j + q;
          dv[q] = __ldcg(p.dyt + lm(gt, 128, c, lane));
          if (gy) dv[q] += __ldg(gy + lm(gt, 128, c, lane));
          av[q] = __ldg(g_t + lm(gt, 512, c, lane));
          tv[q] = __ldg(g_t + lm(gt, 512, 384 + c, lane));
          yv[q] = __ldg(yprev + lm(gt, 128, c, lane));
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int c = c0 + j + q;
          const float d = dv[q], a = av[q], tl = tv[q];
          dl[q] = d * a * (1.f - tl * tl);
          p.s0[lm(gt, 128, c, lane)] = d * (tl - yv[q]) * a * (1.f - a * inv_dt);
          p.dyt[lm(gt, 128, c, lane)] = d * (1.f - a);
        }
        if (j == 0) wait_peer_free(e);
        state_store4(e, r, c0 + j, make_float4(dl[0], dl[1], dl[2], dl[3]));
        state_store4(e, r, c0 + j + 4, make_float4(dl[4], dl[5], dl[6], dl[7]));
      }
      publish(e, rank);
      // ---- acc1 = dL Wz[:, :128] (own columns) -> TMEM columns 0..63; the dL copy-out overlaps it
      if (tid == 0) gemm_issue(e, rank, true, 1, 0, false);
      __syncwarp();
      image_to_global(e.smA, rank, p.dL + (size_t)t * p.N * 128, 128, row0, p.N);
      gemm_wait(e);
      // ---- bwd_z : dG1 -> state tiles, dG2 -> s2, dz <- d (1 - b)
      {
        float v[32];
        tmem_ld32(e.tmem + tlane + (uint32_t)ch0, v);
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          float dv[8], bv[8], zcv[8], zpv[8], g1[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const int c = c0 + j + q;
            dv[q] = __ldcg(p.dzt + lm(gt, 128, c, lane)) + v[j + q];
            if (gz) dv[q] += __ldg(gz + lm(gt, 128, c, lane));
            bv[q] = __ldg(g_t + lm(gt, 512, 128 + c, lane));
            zcv[q] = __ldg(g_t + lm(gt, 512, 256 + c, lane));
            zpv[q] = __ldg(zprev + lm(gt, 128, c, lane));
          }
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const int c = c0 + j + q;
            const float d = dv[q], b = bv[q], zc = zcv[q];
            g1[q] = d * (zc - zpv[q]) * b * (1.f - b * inv_dt);
            p.s2[lm(gt, 128, c, lane)] = d * b * (1.f - zc * zc);
            p.dzt[lm(gt, 128, c, lane)] = d * (1.f - b);
          }
          if (j == 0) wait_peer_free(e);
          state_store4(e, r, c0 + j, make_float4(g1[0], g1[1], g1[2], g1[3]));
          state_store4(e, r, c0 + j + 4, make_float4(g1[4], g1[5], g1[6], g1[7]));
        }
      }
      // ---- acc2 = [dG1 | dG2 | dG0] W[:, :128] (own columns) -> TMEM columns 64..127 (weight chunks 4..7, 8..11, 0..3)
      publish(e, rank);
      if (tid == 0) gemm_issue(e, rank, true, 1, 64, false);
      __syncwarp();
      image_to_global(e.smA, rank, dG_t + 128, 384, row0, p.N);
      gemm_wait(e);
      stage_lm(p.s2);
      publish(e, rank);
      if (tid == 0) gemm_issue(e, rank, true, 1, 64, true);
      __syncwarp();
      image_to_global(e.smA, rank, dG_t + 256, 384, row0, p.N);
      gemm_wait(e);
      stage_lm(p.s0);
      publish(e, rank);
      if (tid == 0) gemm_issue(e, rank, true, 1, 64, true);
      __syncwarp();
      image_to_global(e.smA, rank, dG_t, 384, row0, p.N);
      gemm_wait(e);
      // ---- dy += acc2
      {
        float v[32];
        tmem_ld32(e.tmem + tlane + (uint32_t)(64 + ch0), v);
        float cur[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) cur[j] = __ldcg(p.dyt + lm(gt, 128, c0 + j, lane));
#pragma unroll
        for (int j = 0; j < 32; ++j) p.dyt[lm(gt, 128, c0 + j, lane)] = cur[j] + v[j];
      }
      tc_fence_before();
      epi_bar();
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 0) tmem_dealloc(*m.tmem_slot, 128);
}

}  // namespace msmp

using namespace msmp;

#ifdef MSMP_LEM_TICKS
extern "C" int msmp_lem_debug_ticks(long long* host_out) {
  return cudaMemcpyFromSymbol(host_out, g_lem_dbg, sizeof(long long) * 64) == cudaSuccess ? 0 : -2;
}
#endif

extern "C" int msmp_lem_tc_fwd(const float* inp, int ninp, const float* Wt_in, const float* Wzt_in, const float* Wimg,
                               const float* Wzimg, const float* bias, const float* bias_z, float* pre, float* Y,
                               float* Z, float* Yt, float* Zt, float* gates, float dt, int T, int N, int Npad,
                               cudaStream_t stream) {
  if (T < 0 || N < 0 || ninp < 0 || ninp > 8 || Npad < N || (Npad & 127)) return MSMP_ERR_ARG;
  if (T == 0 || N == 0) return MSMP_OK;
  const size_t tiles = (size_t)T * (Npad / 32);
  k_lem_inproj<<<(unsigned)((4 * tiles + 7) / 8), 256, 0, stream>>>(inp, Wt_in, Wzt_in, bias, bias_z, pre, T, N, Npad, ninp);
  MSMP_CHECK_LAUNCH();
  LemFwdParams p{pre, Wimg, Wzimg, Y, Z, Yt, Zt, gates, dt, T, N, Npad};
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(k_lem_fwd_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, LT_SMEM) != cudaSuccess)
      return MSMP_ERR_CUDA;
    attr_set = true;
  }
  k_lem_fwd_tc<<<2 * (Npad / 128), LT_THREADS, LT_SMEM, stream>>>(p);
  MSMP_CHECK_LAUNCH();
  return MSMP_OK;
}

extern "C" int msmp_lem_tc_bwd(const float* Wzh_img, const float* Wh_img, const float* Yt, const float* Zt,
                               const float* gates, const float* gYt, const float* gZt, int g_last_only, float* dG,
                               float* dL, float* dyt, float* dzt, float* s0, float* s2, float dt, int T, int t_begin,
                               int t_end, int N, int Npad, cudaStream_t stream) {
  if (T < 0 || N < 0 || Npad < N || (Npad & 127) || t_begin < 0 || t_end > T || t_begin > t_end) return MSMP_ERR_ARG;
  if (t_begin == t_end || N == 0) return MSMP_OK;
  LemBwdParams p{Wzh_img, Wh_img, Yt, Zt, gates, gYt, gZt, g_last_only, dG, dL, dyt, dzt, s0, s2, dt, T, N, Npad,
                 t_begin, t_end};
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(k_lem_bwd_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, LT_SMEM) != cudaSuccess)
      return MSMP_ERR_CUDA;
    attr_set = true;
  }
  k_lem_bwd_tc<<<2 * (Npad / 128), LT_THREADS, LT_SMEM, stream>>>(p);
  MSMP_CHECK_LAUNCH();
  return MSMP_OK;
}
