// Persistent LEM recurrence on the tensor cores (replaces lem_cuda.forward / lem_cuda.backward,
// experiments/models_gnn.py:290-292,300).  The recurrence is independent per node, so a CTA owns a tile of 64 nodes
// and walks all T time steps inside ONE launch: no inter-CTA synchronisation, the state tile never leaves the SM.
//
// The GEMMs are issued TRANSPOSED: D^T[channel][node] = W[channel][k] * state[node][k]^T, i.e. the weight images
// are the A operand (M = 128 channels = TMEM lanes) and the state tile is the B operand (N = 64 nodes = TMEM
// columns).  Consequences:
//   * the node count per CTA is the MMA's N, not its M: 64-node tiles put the reference's 6400-node batches on 100
//     of the 148 SMs (128-node tiles: 50) without padding the MMA and without any exchange between CTAs
//     (a 2-CTA column split of 128-node tiles was measured first: its state exchange is bound by the 17 B/clk
//     distributed-shared-memory path and cancelled the gain);
//   * an epilogue thread owns one hidden channel (its TMEM lane) for 32 nodes, so the 32 lanes of a warp touch 32
//     consecutive channels of one node: every global array (pre, gates, Y, Z, dG, dL, carried dy/dz) is plain
//     row-major [node][channel] and every access is one full 128-byte line per warp -- no lane-major copies, no
//     staged copy-out.
//
// Roles (320 threads): warps 0..7 run the gate epilogues, one lane of warp 8 streams the weights, warp 9 issues the
// MMAs convergently (3xTF32: hi*hi + lo*hi + hi*lo; see elect_one() in umma.cuh).  The epilogue warps hand the state
// tile to the MMA warp through a named barrier (bar.arrive / bar.sync) and wait for the accumulators on an mbarrier
// that receives the tcgen05.commit.  The weights (pre-split, pre-swizzled [128 x 32] images, 512 KiB per step) are
// streamed from L2 through a ring of 32 KiB bulk copies; the producer runs ahead of the MMA warp across phase
// boundaries (the weights do not depend on the step), so the next GEMM's first chunks land while a gate epilogue runs.
// The tensor pipe reads both operands from shared memory (6 KiB per 128x64x8 MMA, ~65 cycles measured): the GEMM phases
// are bound by that, not by the stream.  An epilogue thread keeps its channel for all steps, so its recurrent state
// (y, z forward; dy, dz backward: 64 registers) never leaves the register file.
//
// forward, per step t:
//   G^T[3 x 128][64] = W_h y_{t-1}^T            12 weight chunks (3 m-tiles x 4 k-chunks), TMEM columns 0..191
//   gate_z      : a = dt sig(G0 + pre), b = dt sig(G1 + pre), zc = tanh(G2 + pre), z_t = (1-b) z_{t-1} + b zc
//   L^T[128][64]     = Wz_h z_t^T               4 weight chunks, TMEM columns 192..255
//   gate_y      : tL = tanh(L + pre), y_t = (1-a) y_{t-1} + a tL
// The input part of both affine maps never touches memory: a thread's channel is fixed, so it keeps its <= 8 input
// weights per gate and the bias in registers; while the G GEMM occupies the tensor pipe it evaluates
// pre = b + I_t[node] . w_in (exact fp32 FMAs, I_t[node] is a warp-uniform 32-byte load) for the four gates into
// TMEM columns 256..511, from where the gate epilogues add it to the accumulators.  The gate epilogues issue no
// global loads at all; per node-step they write the four gate activations (for the backward) and y, z.
// Shared memory: one state tile [64 nodes x 128 k] (tf32 hi | lo images, 64 KiB) + 5 ring stages.
//
// backward, per step t = t_end-1 .. t_begin (the carried dy, dz enter and leave through global memory once per launch):
//   bwd_y : d = dy + gY[t]; dL = d a (1-tL^2); dG0 = d (tL - y_{t-1}) a (1 - a/dt); dy = d (1-a)
//   acc1  = Wz[:, :128]^T dL^T                    4 chunks
//   bwd_z : d = dz + gZ[t] + acc1; dG1 = d (zc - z_{t-1}) b (1 - b/dt); dG2 = d b (1-zc^2); dz = d (1-b)
//   acc2  = W[:, :128]^T [dG1 | dG2 | dG0]^T      3 x 4 chunks
//   dy   += acc2
// Shared memory: TWO state tiles + 3 ring stages.  bwd_y puts dL into tile X (and stashes dG0 in TMEM); bwd_z puts dG1
// into X and dG2 into Y, so the MMA warp runs the dG1 and dG2 k-blocks back to back; as soon as the dG1 block has been
// read (its own commit) the epilogue restages dG0 from TMEM into X for the third k-block.
// dG [T,N,384] and dL [T,N,128] are written (coalesced, by the epilogue itself) for the four weight-gradient GEMMs
// (msmp_linear_wgrad_tc2).  Measured per step (CTA 0, clock64): forward 31.4 k cycles (input projection 5.4 k inside
// the 13.9 k G GEMM, gate_z 8.5 k, L GEMM 4.3 k, gate_y 4.7 k); backward 49 k (bwd_y 16 k, bwd_z 16 k, GEMM waits 16 k):
// the backward epilogues are bound by the number of 4-byte-per-lane global load/store instructions, not by HBM.
#include "umma.cuh"
#include "msmp_b200.h"

namespace msmp {

constexpr int LT_NODES = 64;                       // nodes per CTA = MMA N
constexpr int LT_SCHUNK = LT_NODES * 128;          // one hi (or lo) state k-chunk image: 64 rows x 128 B = 8 KiB
constexpr int LT_S_BYTES = 4 * 2 * LT_SCHUNK;      // state tile: 4 k-chunks x (hi | lo) = 64 KiB
constexpr int LT_STAGE_BYTES = 2 * IMG_BYTES;      // one weight chunk: [128 x 32] hi | lo = 32 KiB
constexpr int LT_SMEM = 7 * LT_STAGE_BYTES + 1024 + 256;      // forward: 1 state tile + 5 ring stages; backward: 2 + 3
constexpr int LT_EPI = 256;                        // warps 0..7: gate epilogues; warp 8: ring producer; warp 9: MMA issue

#ifdef MSMP_LEM_TICKS
__device__ long long g_lem_dbg[64];
__device__ int g_lem_chunk_tick = -1;      // >= 0: ring_mma stamps (data ready, issued) of the next chunks
#endif

struct Ring {
  uint8_t* smB;        // nst stages
  uint64_t* bfull;     // [nst], one arrival (expect_tx) + 32 KiB of bulk-copy bytes
  uint64_t* bfree;     // [nst], one arrival (tcgen05.commit)
  uint32_t nst;
};

// producer lane: one 32 KiB bulk copy of weight chunk image `src` (hi | lo, contiguous) into ring stage i
// FAST (reduced-precision mode, see linear_tc.cu): only the hi image is copied, only the hi x hi product is issued and the
// epilogues store only the tf32-rounded state -- half the weight stream from L2 and a third of the MMAs.
template <bool FAST>
__device__ __forceinline__ void ring_load(const Ring& rg, uint32_t i, const float* src) {
  const uint32_t s = i % rg.nst, use = i / rg.nst;
  if (use > 0) mbar_wait(&rg.bfree[s], (use - 1) & 1);            // MMAs that read this stage are complete
  constexpr uint32_t NB = FAST ? IMG_BYTES : LT_STAGE_BYTES;      // = the stage size (reduced precision: hi image only)
  mbar_expect_tx(&rg.bfull[s], NB);
  bulk_g2s(rg.smB + s * NB, src, NB, &rg.bfull[s]);
}

// MMA-issuing warp (all lanes, convergent): weight chunk in ring slot `slot` (A, 128 channels x 32 k) times state k-chunk at
// s_img (B, 64 nodes x 32 k).  The warp's own instruction stream used to set the pace of every GEMM phase (~1.1 k cycles per
// chunk whether it issued 12 MMAs or 4, scripts/lem_ticks.py): sixteen descriptors were assembled from scratch per chunk and
// the slot came from a modulo by a run-time ring depth.  Now: one base descriptor per operand image and chunk, the k-steps and
// the lo images are 32-bit adds on its address field (+2 units of 16 bytes per k-step; smem addresses stay below 2^18), and
// the caller carries slot / phase counters.
template <bool FAST>
__device__ __forceinline__ void ring_mma(const Ring& rg, uint32_t slot, uint32_t phase, uint32_t s_img, uint32_t tmem_d,
                                         bool accumulate) {
  mbar_wait(&rg.bfull[slot], phase);                              // (bulk copies write through the async proxy)
#ifdef MSMP_LEM_TICKS
  if (blockIdx.x == 0 && g_lem_chunk_tick >= 0 && g_lem_chunk_tick < 16) g_lem_dbg[16 + 2 * g_lem_chunk_tick] = clock64();
#endif
  tc_fence_after();
  constexpr uint32_t IDESC = umma_idesc_tf32(128, LT_NODES, 0, 0);
  const uint64_t dw = umma_desc(smem_u32(rg.smB + slot * (FAST ? IMG_BYTES : LT_STAGE_BYTES)), 16, 1024);
  const uint64_t ds = umma_desc(s_img, 16, 1024);
  const uint32_t dw_lo = (uint32_t)dw, dw_hi = (uint32_t)(dw >> 32), ds_lo = (uint32_t)ds, ds_hi = (uint32_t)(ds >> 32);
  const bool leader = elect_one();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const uint64_t dwh = ((uint64_t)dw_hi << 32) | (dw_lo + 2 * k), dwl = ((uint64_t)dw_hi << 32) | (dw_lo + 2 * k + (IMG_BYTES >> 4));
    const uint64_t dsh = ((uint64_t)ds_hi << 32) | (ds_lo + 2 * k), dsl = ((uint64_t)ds_hi << 32) | (ds_lo + 2 * k + (LT_SCHUNK >> 4));
    if (leader) {
      umma_tf32(tmem_d, dwh, dsh, IDESC, (accumulate || k) ? 1u : 0u);
      if (!FAST) {
        umma_tf32(tmem_d, dwl, dsh, IDESC, 1u);
        umma_tf32(tmem_d, dwh, dsl, IDESC, 1u);
      }
    }
  }
  if (leader) umma_commit(&rg.bfree[slot]);
  __syncwarp();
#ifdef MSMP_LEM_TICKS
  if (blockIdx.x == 0 && g_lem_chunk_tick >= 0 && g_lem_chunk_tick < 16) g_lem_dbg[17 + 2 * g_lem_chunk_tick++] = clock64();
#endif
}

// Per-thread context of the epilogue warps.
struct Epi {
  Ring rg;
  uint64_t* acc;
  uint32_t nacc;       // accumulator barrier phases consumed
  uint32_t slot;       // ring slot of the next chunk (MMA warp)
  uint32_t sphase;     // ... and its barrier phase
  uint32_t smS;        // shared::cta address of the state tile
  uint32_t tmem;
};

// MMA warp (all lanes): issue one GEMM phase of `nchunks` ring chunks; chunk j multiplies k-chunk (j & 3) of the state
// tile at `sbase` into TMEM columns dcol + 64 * (j >> 2) (the first chunk of every 64-column block overwrites unless
// acc_first), then commits to `done` (if not null).
template <bool FAST>
__device__ __forceinline__ void gemm_issue(Epi& e, uint32_t sbase, uint32_t nchunks, uint32_t dcol, bool acc_first,
                                           uint64_t* done) {
  tc_fence_after();
  for (uint32_t j = 0; j < nchunks; ++j) {
    ring_mma<FAST>(e.rg, e.slot, e.sphase, sbase + (j & 3) * 2 * LT_SCHUNK, e.tmem + dcol + LT_NODES * (j >> 2), acc_first || (j & 3) != 0);
    if (++e.slot == e.rg.nst) {
      e.slot = 0;
      e.sphase ^= 1;
    }
  }
  if (done != nullptr && elect_one()) umma_commit(done);
  __syncwarp();
}

// state tile element (node j, k): k-chunk k >> 5, row j, column k & 31 of a [64 x 32] 128B-swizzled image (hi, then lo)
__device__ __forceinline__ uint32_t state_off(int j, int k) {
  return (uint32_t)(k >> 5) * (2 * LT_SCHUNK) + img_off(j, (k & 31) >> 2) + 4u * (uint32_t)(k & 3);
}
template <bool FAST>
__device__ __forceinline__ void state_store(uint32_t smS, uint32_t off, float v) {
  float h, l;
  split_tf32(v, h, l);
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(smS + off), "f"(h) : "memory");
  if (!FAST) asm volatile("st.shared.f32 [%0], %1;" ::"r"(smS + off + LT_SCHUNK), "f"(l) : "memory");
}

// clock64() phase stamps of CTA 0 at step 2 (scripts/lem_ticks.py); compiled in only with -DMSMP_LEM_TICKS.
#ifdef MSMP_LEM_TICKS
#define LEM_TICK(i) do { if (blockIdx.x == 0 && threadIdx.x == 0 && t == 2) g_lem_dbg[i] = clock64(); } while (0)
#else
#define LEM_TICK(i) do { } while (0)
#endif

struct LemSmem {
  uint8_t* smS;        // nbuf state tiles of LT_S_BYTES
  uint8_t* smB;        // nst ring stages
  uint64_t* bars;      // bfull[8], bfree[8], acc, acc_mid, forward: accG[3], accL
  uint32_t* tmem_slot;
  uint32_t nst;
};

__device__ __forceinline__ LemSmem lem_smem(uint8_t* smem_raw, int nbuf, int nst) {      // 2 * nbuf + nst * 2 == 14 half-stages
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  LemSmem m;
  m.smS = smem;
  m.smB = smem + nbuf * LT_S_BYTES;
  m.bars = reinterpret_cast<uint64_t*>(m.smB + nst * LT_STAGE_BYTES);
  m.tmem_slot = reinterpret_cast<uint32_t*>(m.bars + 24);
  m.nst = (uint32_t)nst;
  return m;
}

__device__ __forceinline__ void lem_init(const LemSmem& m, uint32_t tmem_cols) {
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) tmem_alloc(m.tmem_slot, tmem_cols);
  if (tid == 32) {
    for (int i = 0; i < 8; ++i) {
      mbar_init(&m.bars[i], 1);
      mbar_init(&m.bars[8 + i], 1);
    }
    mbar_init(&m.bars[16], 1);
    mbar_init(&m.bars[17], 1);
    for (int i = 18; i < 22; ++i) mbar_init(&m.bars[i], 1);      // forward: one barrier per gate block of G and one for L
    fence_barrier_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
}

__device__ __forceinline__ Epi lem_epi(const LemSmem& m) {
  Epi e;
  e.rg = Ring{m.smB, &m.bars[0], &m.bars[8], m.nst};
  e.acc = &m.bars[16];
  e.nacc = e.slot = e.sphase = 0;
  e.smS = smem_u32(m.smS);
  e.tmem = __shfl_sync(0xffffffffu, *m.tmem_slot, 0);
  return e;
}

struct LemFwdParams {
  const float* inp;      // row-major [T][N][32]: inputs, columns >= ninp are zero
  const float* Wt_in;    // [>= ninp][384] input rows of the k-major W^T
  const float* Wzt_in;   // [>= ninp][128] input rows of the k-major Wz^T
  const float* bias;     // [384]
  const float* bias_z;   // [128]
  const float* Wimg;     // images of W[:, :128]  rows = gate channel: [3 m-tiles][4 k-chunks][2][4096]
  const float* Wzimg;    // images of Wz[:, :128]: [1][4][2][4096]
  float* Y;              // row-major [T+1][N][128]  (Y[0] = y0 on entry)
  float* Z;              // row-major [T+1][N][128]  (Z[0] = z0 on entry)
  float* gates;          // row-major [T][Npad][512]  a | b | zc | tL
  float dt;
  int T; int N; int Npad; int ninp;
};

// Forward kernel roles: warps 0..7 gate epilogues, warp 8 weight ring producer, warp 9 MMA issue.
constexpr int LF_THREADS = LT_EPI + 64;
// TMEM columns (512 allocated): the G accumulator (3 gates x 64 nodes), the L accumulator, and the input projections
// of the four gates (bias + I_t . w_in), written by the epilogue warps while the G GEMM runs and added to the
// accumulators in the gate epilogues.  (Initialising the accumulators with them instead was measured slightly less
// accurate: every MMA then adds onto a full-magnitude partial sum.)
constexpr uint32_t LF_G = 0, LF_L = 192, LF_PRE = 256;

// producer / consumer named barrier between the 256 epilogue threads (arrive) and the MMA warp (sync)
__device__ __forceinline__ void state_ready_arrive() { asm volatile("bar.arrive 2, 288;" ::: "memory"); }
__device__ __forceinline__ void state_ready_wait() { asm volatile("bar.sync 2, 288;" ::: "memory"); }

template <bool FAST>
__global__ void __launch_bounds__(LF_THREADS, 1) k_lem_fwd_tc(const LemFwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  // (reduced precision: 16 KiB stages, so the same ring memory holds 8 of them -- the GEMM phases are bound by the refill
  // round trip of the ring, not by the MMAs; the barrier arrays hold 8 stages)
  LemSmem m = lem_smem(smem_raw, 1, 5);
  if (FAST) m.nst = 8;
  const int tid = threadIdx.x, warp = warp_index_uniform(), lane = tid & 31;
  const int row0 = blockIdx.x * LT_NODES;
  const size_t plane = (size_t)p.N * 128;
  lem_init(m, 512);

  if (warp == 8) {
    // ---- weight ring producer: 12 G chunks + 4 L chunks per step, running ahead of the MMA warp
    const Ring rg{m.smB, &m.bars[0], &m.bars[8], m.nst};
    if (elect_one()) {
      uint32_t n = 0;
      for (int t = 0; t < p.T; ++t) {
        for (int j = 0; j < 12; ++j) ring_load<FAST>(rg, n++, p.Wimg + (size_t)j * 2 * (IMG_BYTES / 4));
        for (int j = 0; j < 4; ++j) ring_load<FAST>(rg, n++, p.Wzimg + (size_t)j * 2 * (IMG_BYTES / 4));
      }
    }
    __syncwarp();
  } else if (warp == 9) {
    // ---- MMA issue (whole warp convergent, elected lane issues)
    Epi e = lem_epi(m);
    for (int t = 0; t < p.T; ++t) {
      state_ready_wait();                                   // y_{t-1} tile written, G accumulator of step t-1 consumed
      // one commit per gate block (4 chunks = 64 accumulator columns): the epilogue warps turn block g into its gate
      // while blocks g + 1.. are still on the tensor pipe.  Every barrier completes exactly once per step.
      gemm_issue<FAST>(e, e.smS, 4, LF_G, false, &m.bars[18]);
      gemm_issue<FAST>(e, e.smS, 4, LF_G + 64, false, &m.bars[19]);
      gemm_issue<FAST>(e, e.smS, 4, LF_G + 128, false, &m.bars[20]);
      state_ready_wait();                                   // z_t tile written, L accumulator of step t-1 consumed
      gemm_issue<FAST>(e, e.smS, 4, LF_L, false, &m.bars[21]);
    }
  } else {
    Epi e = lem_epi(m);
    // epilogue ownership: thread = hidden channel c (TMEM lane), nodes [j0, j0 + 32) of the tile, for all T steps:
    // its y / z state lives in registers.
    const int c = 32 * (warp & 3) + lane;
    const int j0 = 32 * (warp >> 2);
    const uint32_t tbase = e.tmem + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)j0;

    // this thread's input weights and biases: gates 0..2 (W) and the y-gate (Wz)
    float win[4][8], bia[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      bia[g] = g < 3 ? __ldg(p.bias + 128 * g + c) : __ldg(p.bias_z + c);
#pragma unroll
      for (int q = 0; q < 8; ++q)
        win[g][q] = q < p.ninp ? (g < 3 ? __ldg(p.Wt_in + (size_t)q * 384 + 128 * g + c) : __ldg(p.Wzt_in + (size_t)q * 128 + c))
                               : 0.f;
    }
    // input projections of the four gates for step t: bias + I_t[node] . w_in (exact fp32) -> TMEM columns LF_PRE + 64 g;
    // I_t[node] is a warp-uniform 32-byte load.  Runs while the G GEMM of the same step occupies the tensor pipe.
    auto inproj = [&](int t) {
      const float* x_t = p.inp + ((size_t)t * p.N + row0) * 32;
#pragma unroll 1
      for (int jj = 0; jj < 32; jj += 4) {
        float4 xa[4], xb[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int j = j0 + jj + q;
          const bool ok = row0 + j < p.N;
          xa[q] = ok ? ldg4(x_t + (size_t)j * 32) : zero4();
          xb[q] = ok ? ldg4(x_t + (size_t)j * 32 + 4) : zero4();
        }
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          float v[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float acc = bia[g];
            acc = fmaf(xa[q].x, win[g][0], acc);
            acc = fmaf(xa[q].y, win[g][1], acc);
            acc = fmaf(xa[q].z, win[g][2], acc);
            acc = fmaf(xa[q].w, win[g][3], acc);
            acc = fmaf(xb[q].x, win[g][4], acc);
            acc = fmaf(xb[q].y, win[g][5], acc);
            acc = fmaf(xb[q].z, win[g][6], acc);
            acc = fmaf(xb[q].w, win[g][7], acc);
            v[q] = acc;
          }
          tmem_st4(tbase + LF_PRE + 64 * g + jj, v);
        }
      }
      tmem_st_wait();
    };
    // make state-tile stores (generic proxy) and TMEM stores visible to the MMA warp, then signal it
    auto publish_to_mma = [&]() {
      fence_proxy_async();
      tc_fence_before();
      state_ready_arrive();
    };
    auto wait_acc = [&](int b, int t) {          // b: 0..2 = gate block of G, 3 = L; each completes once per step
      mbar_wait_warp(&m.bars[18 + b], (uint32_t)t & 1u);
      tc_fence_after();
    };

    // initial state: registers + the y_{-1} tile
    float yreg[32], zreg[32], hold[32];          // hold: zc / tanh(L) of the step between their use and their global store
#pragma unroll
    for (int q = 0; q < 32; ++q) {
      const int g = row0 + j0 + q;
      yreg[q] = g < p.N ? __ldg(p.Y + (size_t)g * 128 + c) : 0.f;
      zreg[q] = g < p.N ? __ldg(p.Z + (size_t)g * 128 + c) : 0.f;
      state_store<FAST>(e.smS, state_off(j0 + q, c), yreg[q]);
    }
    publish_to_mma();

    for (int t = 0; t < p.T; ++t) {
      float* __restrict__ g_t = p.gates + ((size_t)t * p.Npad + row0) * 512;
      float* __restrict__ znext = p.Z + (size_t)(t + 1) * plane + (size_t)row0 * 128;
      float* __restrict__ ynext = p.Y + (size_t)(t + 1) * plane + (size_t)row0 * 128;
      const uint32_t gbuf = tbase + LF_G;
      LEM_TICK(0);
      // ---- while G^T = W_h y^T runs: this step's input projections -> TMEM
      inproj(t);
      LEM_TICK(1);
      // ---- gate_z, block by block behind the G GEMM (no global loads: accumulators from TMEM, z_{t-1} from registers).
      // a and b are parked in the projection columns they consume (gate 0 / gate 1) until gate_y / the z update needs them.
      wait_acc(0, t);
      LEM_TICK(2);
#pragma unroll
      for (int jj = 0; jj < 32; jj += 8) {
        uint32_t r0[8], q0[8];
        tmem_ld8_nowait(gbuf + jj, r0);
        tmem_ld8_nowait(tbase + LF_PRE + jj, q0);
        tmem_ld_wait();
        float av[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float a = p.dt * sigmoid_r(__uint_as_float(r0[q]) + __uint_as_float(q0[q]));
          av[q] = a;
          g_t[(size_t)(j0 + jj + q) * 512 + c] = a;
        }
        tmem_st8(tbase + LF_PRE + jj, av);
      }
      wait_acc(1, t);
#pragma unroll
      for (int jj = 0; jj < 32; jj += 8) {
        uint32_t r1[8], q1[8];
        tmem_ld8_nowait(gbuf + 64 + jj, r1);
        tmem_ld8_nowait(tbase + LF_PRE + 64 + jj, q1);
        tmem_ld_wait();
        float bv[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float b = p.dt * sigmoid_r(__uint_as_float(r1[q]) + __uint_as_float(q1[q]));
          bv[q] = b;
          g_t[(size_t)(j0 + jj + q) * 512 + 128 + c] = b;
        }
        tmem_st8(tbase + LF_PRE + 64 + jj, bv);
      }
      tmem_st_wait();
      wait_acc(2, t);
#pragma unroll
      for (int jj = 0; jj < 32; jj += 8) {
        uint32_t r2[8], q2[8], rb[8];
        tmem_ld8_nowait(gbuf + 128 + jj, r2);
        tmem_ld8_nowait(tbase + LF_PRE + 128 + jj, q2);
        tmem_ld8_nowait(tbase + LF_PRE + 64 + jj, rb);
        tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int j = j0 + jj + q;
          const float b = __uint_as_float(rb[q]);
          const float zc = tanh_r(__uint_as_float(r2[q]) + __uint_as_float(q2[q]));
          const float zn = (1.f - b) * zreg[jj + q] + b * zc;
          zreg[jj + q] = zn;
          hold[jj + q] = zc;
          state_store<FAST>(e.smS, state_off(j, c), zn);      // z_t: B operand of the L GEMM
        }
      }
      publish_to_mma();
      LEM_TICK(3);
      // the global stores of zc and z_t are off the recurrence's critical path: they go out while the L GEMM runs
#pragma unroll
      for (int q = 0; q < 32; ++q) {
        const int j = j0 + q;
        g_t[(size_t)j * 512 + 256 + c] = hold[q];
        if (row0 + j < p.N) znext[(size_t)j * 128 + c] = zreg[q];
      }
      // ---- L^T = Wz_h z^T
      wait_acc(3, t);
      LEM_TICK(4);
      // ---- gate_y
#pragma unroll
      for (int jj = 0; jj < 32; jj += 8) {
        uint32_t r3[8], q3[8], ra[8];
        tmem_ld8_nowait(tbase + LF_L + jj, r3);
        tmem_ld8_nowait(tbase + LF_PRE + 192 + jj, q3);
        tmem_ld8_nowait(tbase + LF_PRE + jj, ra);
        tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int j = j0 + jj + q;
          const float a = __uint_as_float(ra[q]);
          const float tl = tanh_r(__uint_as_float(r3[q]) + __uint_as_float(q3[q]));
          const float yn = (1.f - a) * yreg[jj + q] + a * tl;
          yreg[jj + q] = yn;
          hold[jj + q] = tl;
          state_store<FAST>(e.smS, state_off(j, c), yn);      // y_t: B operand of the next G GEMM
        }
      }
      if (t + 1 < p.T) publish_to_mma();
      LEM_TICK(5);
      // ... and those of tanh(L) and y_t while the next step's G GEMM runs
#pragma unroll
      for (int q = 0; q < 32; ++q) {
        const int j = j0 + q;
        g_t[(size_t)j * 512 + 384 + c] = hold[q];
        if (row0 + j < p.N) ynext[(size_t)j * 128 + c] = yreg[q];
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(*m.tmem_slot, 512);
}

// ------------------------------------------------------------------------------------------------ backward
struct LemBwdParams {
  const float* Wzh_img;  // images of Wz[:, :128]^T (rows = hidden channel, k = gate channel): [1][4][2][4096]
  const float* Wh_img;   // images of W[:, :128]^T  (rows = hidden channel, k = 384 gate channels): [1][12][2][4096]
  const float* Y;        // row-major [T+1][N][128]
  const float* Z;        // row-major [T+1][N][128]
  const float* gates;    // row-major [T][Npad][512]
  const float* gY;       // row-major external gradients: [T][N][128] or, if g_last_only, [N][128] applied at t = T-1
  const float* gZ;       // (either may be NULL)
  int g_last_only;
  float* dG;             // row-major [T][N][384]
  float* dL;             // row-major [T][N][128]
  float* dy;             // row-major [Npad][128] carried gradient (zero on entry; d/dy0 on exit)
  float* dz;             // row-major [Npad][128]
  float dt;
  int T; int N; int Npad;
  int t_begin; int t_end;   // this launch walks t = t_end-1 .. t_begin (the carried dy/dz live in dy/dz between launches)
};

// TMEM columns of the backward kernel (256 allocated): acc1, acc2, and the dG2 / dG0 blocks of the current step
// (stashed by the thread that computed them until their k-block of the acc2 GEMM is staged).
constexpr uint32_t LB_ACC1 = 0, LB_ACC2 = 64, LB_S2 = 128, LB_S0 = 192;
// Round 2: the other half of tensor memory holds Wz[:, :128]^T (tf32 hi at columns 256.., lo at 384..; lane = hidden channel,
// column = gate channel k), copied once per launch from the packed images: the acc1 GEMM takes its A operand from tensor
// memory and its 4 weight chunks per step (128 KiB of the 512 KiB streamed from L2, through a ring of only 3 stages) are gone.
constexpr uint32_t LB_WZ_HI = 256, LB_WZ_LO = 384;
// The backward epilogues are bound by memory latency (saved activations in, dG / dL out), not by issue slots: sixteen
// epilogue warps (16 nodes per thread instead of 32) keep twice as many loads in flight.
constexpr int LB_EPI_WARPS = 16;
constexpr int LB_NPT = LT_NODES / (LB_EPI_WARPS / 4);       // nodes per thread
constexpr int LB_THREADS = 32 * LB_EPI_WARPS + 128;         // + ring producer warp + MMA warp + two idle warps: a full warpgroup,
                                                            // which hands its registers to the epilogue warps (setmaxnreg 40 / 104)

__device__ __forceinline__ void bwd_ready_arrive() { asm volatile("bar.arrive 2, %0;" ::"n"(32 * LB_EPI_WARPS + 32) : "memory"); }
__device__ __forceinline__ void bwd_ready_wait() { asm volatile("bar.sync 2, %0;" ::"n"(32 * LB_EPI_WARPS + 32) : "memory"); }

template <bool FAST>
__global__ void __launch_bounds__(LB_THREADS, 1) k_lem_bwd_tc(const LemBwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  LemSmem m = lem_smem(smem_raw, 2, 3);      // two state tiles (X, Y) + 3 ring stages (reduced precision: 6 of 16 KiB)
  if (FAST) m.nst = 6;
  const int tid = threadIdx.x, warp = warp_index_uniform(), lane = tid & 31;
  const int row0 = blockIdx.x * LT_NODES;
  const size_t plane = (size_t)p.N * 128;
  lem_init(m, 512);
  uint64_t* acc_mid = &m.bars[17];

  if (warp >= LB_EPI_WARPS + 2) {
    reg_dec<40>();
  } else if (warp == LB_EPI_WARPS) {
    reg_dec<40>();
    // ---- weight ring producer: per step W chunks 4..7 (dG1), 8..11 (dG2), 0..3 (dG0)   (Wz lives in tensor memory)
    const Ring rg{m.smB, &m.bars[0], &m.bars[8], m.nst};
    if (elect_one()) {
      uint32_t n = 0;
      for (int t = p.t_end - 1; t >= p.t_begin; --t) {
        for (int j = 0; j < 12; ++j) ring_load<FAST>(rg, n++, p.Wh_img + (size_t)((j + 4) % 12) * 2 * (IMG_BYTES / 4));
      }
    }
    __syncwarp();
  } else if (warp == LB_EPI_WARPS + 1) {
    reg_dec<40>();
    // ---- MMA issue (whole warp convergent, elected lane issues)
    Epi e = lem_epi(m);
    const uint32_t X = e.smS, Y = e.smS + LT_S_BYTES;
    for (int t = p.t_end - 1; t >= p.t_begin; --t) {
      bwd_ready_wait();                                         // dL in X
      {                                                         // acc1 = Wz[:, :128]^T dL^T, A operand from tensor memory
        tc_fence_after();
        constexpr uint32_t IDESC = umma_idesc_tf32(128, LT_NODES, 0, 0);
        const bool leader = elect_one();
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t s_hi = X + j * 2 * LT_SCHUNK, s_lo = s_hi + LT_SCHUNK;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t dsh = umma_desc(s_hi + 32 * k, 16, 1024), dsl = umma_desc(s_lo + 32 * k, 16, 1024);
            const uint32_t a_hi = e.tmem + LB_WZ_HI + 32 * j + 8 * k, a_lo = e.tmem + LB_WZ_LO + 32 * j + 8 * k;
            if (leader) {
              umma_tf32_ts(e.tmem + LB_ACC1, a_hi, dsh, IDESC, (j | k) ? 1u : 0u);
              if (!FAST) {
                umma_tf32_ts(e.tmem + LB_ACC1, a_lo, dsh, IDESC, 1u);
                umma_tf32_ts(e.tmem + LB_ACC1, a_hi, dsl, IDESC, 1u);
              }
            }
          }
        }
        if (leader) umma_commit(e.acc);
        __syncwarp();
      }
      bwd_ready_wait();                                         // dG1 in X, dG2 in Y
      gemm_issue<FAST>(e, X, 4, LB_ACC2, false, acc_mid);               // acc2  = W[128:256, :128]^T dG1^T   (X free afterwards)
      gemm_issue<FAST>(e, Y, 4, LB_ACC2, true, nullptr);                // acc2 += W[256:384, :128]^T dG2^T
      bwd_ready_wait();                                         // dG0 in X
      gemm_issue<FAST>(e, X, 4, LB_ACC2, true, e.acc);                  // acc2 += W[0:128, :128]^T dG0^T
    }
  } else {
    reg_inc<104>();
    Epi e = lem_epi(m);
    // epilogue ownership: thread = hidden channel c (TMEM lane), nodes [j0, j0 + LB_NPT): its carried dy / dz live in registers
    const int c = 32 * (warp & 3) + lane;
    const int j0 = LB_NPT * (warp >> 2);
    const uint32_t tbase = e.tmem + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)j0;
    const uint32_t X = e.smS, Y = e.smS + LT_S_BYTES;
    const float inv_dt = 1.0f / p.dt;
    float* dyc = p.dy + (size_t)row0 * 128 + c;
    float* dzc = p.dz + (size_t)row0 * 128 + c;
    uint32_t nmid = 0;

    auto publish_to_mma = [&]() {
      fence_proxy_async();
      tc_fence_before();
      bwd_ready_arrive();
    };
    auto wait_bar = [&](uint64_t* bar, uint32_t& n) {
      mbar_wait_warp(bar, n & 1);
      ++n;
      tc_fence_after();
    };

    {  // Wz[:, :128]^T -> tensor memory: row c of k-chunk (warp >> 2) of the packed hi | lo images (already split)
      const int jk = warp >> 2;
      const float* img = p.Wzh_img + (size_t)jk * 2 * (IMG_BYTES / 4);
      const uint32_t lane_base = e.tmem + ((uint32_t)(32 * (warp & 3)) << 16);
#pragma unroll
      for (int half = 0; half < (FAST ? 1 : 2); ++half) {
#pragma unroll
        for (int q = 0; q < 8; q += 2) {
          const float4 w0 = ldg4(img + half * (IMG_BYTES / 4) + img_off(c, q) / 4);
          const float4 w1 = ldg4(img + half * (IMG_BYTES / 4) + img_off(c, q + 1) / 4);
          const float w8[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
          tmem_st8(lane_base + (half ? LB_WZ_LO : LB_WZ_HI) + 32 * jk + 4 * q, w8);
        }
      }
      tmem_st_wait();
    }
    float dyreg[LB_NPT], dzreg[LB_NPT];
#pragma unroll
    for (int q = 0; q < LB_NPT; ++q) {
      dyreg[q] = __ldcg(dyc + (size_t)(j0 + q) * 128);
      dzreg[q] = __ldcg(dzc + (size_t)(j0 + q) * 128);
    }

    for (int t = p.t_end - 1; t >= p.t_begin; --t) {
      const float* __restrict__ g_t = p.gates + ((size_t)t * p.Npad + row0) * 512;
      const float* __restrict__ yprev = p.Y + (size_t)t * plane + (size_t)row0 * 128;
      const float* __restrict__ zprev = p.Z + (size_t)t * plane + (size_t)row0 * 128;
      const bool ext = !p.g_last_only || t == p.T - 1;
      const float* __restrict__ gy = (p.gY && ext) ? p.gY + (p.g_last_only ? 0 : (size_t)t * plane) + (size_t)row0 * 128 : nullptr;
      const float* __restrict__ gz = (p.gZ && ext) ? p.gZ + (p.g_last_only ? 0 : (size_t)t * plane) + (size_t)row0 * 128 : nullptr;
      float* __restrict__ dG_t = p.dG + (size_t)t * p.N * 384 + (size_t)row0 * 384;
      float* __restrict__ dL_t = p.dL + (size_t)t * plane + (size_t)row0 * 128;
      // ---- bwd_y : dL -> tile X + global, dG0 -> global + TMEM stash, dy <- d (1 - a)
      LEM_TICK(32);
#pragma unroll
      for (int jj = 0; jj < LB_NPT; jj += 8) {
        float gv[8], av[8], tv[8], yv[8], g0[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int j = j0 + jj + q;
          const bool ok = row0 + j < p.N;
          gv[q] = (gy && ok) ? __ldg(gy + (size_t)j * 128 + c) : 0.f;
          av[q] = __ldg(g_t + (size_t)j * 512 + c);
          tv[q] = __ldg(g_t + (size_t)j * 512 + 384 + c);
          yv[q] = ok ? __ldg(yprev + (size_t)j * 128 + c) : 0.f;
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int j = j0 + jj + q;
          const bool ok = row0 + j < p.N;
          const float d = dyreg[jj + q] + gv[q], a = av[q], tl = tv[q];
          const float dl = ok ? d * a * (1.f - tl * tl) : 0.f;
          g0[q] = ok ? d * (tl - yv[q]) * a * (1.f - a * inv_dt) : 0.f;
          if (ok) {
            dL_t[(size_t)j * 128 + c] = dl;
            dG_t[(size_t)j * 384 + c] = g0[q];
          }
          dyreg[jj + q] = d * (1.f - a);
          state_store<FAST>(X, state_off(j, c), dl);
        }
        tmem_st8(tbase + LB_S0 + jj, g0);
        LEM_TICK(40 + jj / 8);
      }
      tmem_st_wait();
      publish_to_mma();
      LEM_TICK(33);
      // ---- acc1^T = Wz[:, :128]^T dL^T
      wait_bar(e.acc, e.nacc);
      LEM_TICK(34);
      // ---- bwd_z : dG1 -> tile X + global, dG2 -> tile Y + global, dz <- d (1 - b)
#pragma unroll
      for (int jj = 0; jj < LB_NPT; jj += 8) {
        uint32_t r1[8];
        tmem_ld8_nowait(tbase + LB_ACC1 + jj, r1);
        float gv[8], bv[8], zcv[8], zpv[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int j = j0 + jj + q;
          const bool ok = row0 + j < p.N;
          gv[q] = (gz && ok) ? __ldg(gz + (size_t)j * 128 + c) : 0.f;
          bv[q] = __ldg(g_t + (size_t)j * 512 + 128 + c);
          zcv[q] = __ldg(g_t + (size_t)j * 512 + 256 + c);
          zpv[q] = ok ? __ldg(zprev + (size_t)j * 128 + c) : 0.f;
        }
        tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int j = j0 + jj + q;
          const bool ok = row0 + j < p.N;
          const float d = dzreg[jj + q] + gv[q] + __uint_as_float(r1[q]), b = bv[q], zc = zcv[q];
          const float g1 = ok ? d * (zc - zpv[q]) * b * (1.f - b * inv_dt) : 0.f;
          const float g2 = ok ? d * b * (1.f - zc * zc) : 0.f;
          if (ok) {
            dG_t[(size_t)j * 384 + 128 + c] = g1;
            dG_t[(size_t)j * 384 + 256 + c] = g2;
          }
          dzreg[jj + q] = d * (1.f - b);
          state_store<FAST>(X, state_off(j, c), g1);
          state_store<FAST>(Y, state_off(j, c), g2);
        }
      }
      publish_to_mma();
      LEM_TICK(35);
      if (t > p.t_begin) {
        // the activations saved by the forward pass are in HBM by now: while the acc2 GEMMs run (the memory pipe is idle)
        // pull the next step's rows of this thread's node half into L2 (gates 2 KiB, y and z 512 B per node; one
        // 128-byte line per prefetch).  Issued in front of bwd_y the prefetches queued ahead of its demand loads.
        const float* g_n = p.gates + ((size_t)(t - 1) * p.Npad + row0 + j0) * 512;
        const float* y_n = p.Y + (size_t)(t - 1) * plane + (size_t)(row0 + j0) * 128;
        const float* z_n = p.Z + (size_t)(t - 1) * plane + (size_t)(row0 + j0) * 128;
        const int part = warp & 3;      // the four warps of a node half share the work
        for (int q = lane + 32 * part; q < LB_NPT * 16; q += 128)
          asm volatile("prefetch.global.L2 [%0];" ::"l"(g_n + (size_t)(q >> 4) * 512 + 32 * (q & 15)));
        for (int q = lane + 32 * part; q < LB_NPT * 4; q += 128) {
          const int j = q >> 2;
          if (row0 + j0 + j < p.N) {
            asm volatile("prefetch.global.L2 [%0];" ::"l"(y_n + (size_t)j * 128 + 32 * (q & 3)));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(z_n + (size_t)j * 128 + 32 * (q & 3)));
          }
        }
      }
      // ---- acc2^T = W[:, :128]^T [dG1 | dG2 | dG0]^T: the dG0 block is staged into X as soon as the dG1 GEMM has read it
      wait_bar(acc_mid, nmid);
      LEM_TICK(36);
#pragma unroll
      for (int jj = 0; jj < LB_NPT; jj += 8) {
        uint32_t r0[8];
        tmem_ld8_nowait(tbase + LB_S0 + jj, r0);
        tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 8; ++q) state_store<FAST>(X, state_off(j0 + jj + q, c), __uint_as_float(r0[q]));
      }
      publish_to_mma();
      LEM_TICK(37);
      wait_bar(e.acc, e.nacc);
      LEM_TICK(38);
      // ---- dy += acc2
#pragma unroll
      for (int jj = 0; jj < LB_NPT; jj += 8) {
        uint32_t r2[8];
        tmem_ld8_nowait(tbase + LB_ACC2 + jj, r2);
        tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 8; ++q) dyreg[jj + q] += __uint_as_float(r2[q]);
      }
      tc_fence_before();
      LEM_TICK(39);
    }
#pragma unroll
    for (int q = 0; q < LB_NPT; ++q) {
      dyc[(size_t)(j0 + q) * 128] = dyreg[q];
      dzc[(size_t)(j0 + q) * 128] = dzreg[q];
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(*m.tmem_slot, 512);
}

}  // namespace msmp

using namespace msmp;

#ifdef MSMP_LEM_TICKS
extern "C" int msmp_lem_debug_ticks(long long* host_out) {
  return cudaMemcpyFromSymbol(host_out, g_lem_dbg, sizeof(long long) * 64) == cudaSuccess ? 0 : -2;
}
#endif

extern "C" int msmp_lem_tc_fwd(const float* inp, int ninp, const float* Wt_in, const float* Wzt_in, const float* Wimg,
                               const float* Wzimg, const float* bias, const float* bias_z, float* Y, float* Z,
                               float* gates, float dt, int T, int N, int Npad, int mode, cudaStream_t stream) {
  if (T < 0 || N < 0 || ninp < 0 || ninp > 8 || Npad < N || (Npad % LT_NODES)) return MSMP_ERR_ARG;
  if (T == 0 || N == 0) return MSMP_OK;
  LemFwdParams p{inp, Wt_in, Wzt_in, bias, bias_z, Wimg, Wzimg, Y, Z, gates, dt, T, N, Npad, ninp};
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(k_lem_fwd_tc<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, LT_SMEM) != cudaSuccess ||
        cudaFuncSetAttribute(k_lem_fwd_tc<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, LT_SMEM) != cudaSuccess)
      return MSMP_ERR_CUDA;
    attr_set = true;
  }
  if (mode)
    k_lem_fwd_tc<true><<<Npad / LT_NODES, LF_THREADS, LT_SMEM, stream>>>(p);
  else
    k_lem_fwd_tc<false><<<Npad / LT_NODES, LF_THREADS, LT_SMEM, stream>>>(p);
  MSMP_CHECK_LAUNCH();
  return MSMP_OK;
}

extern "C" int msmp_lem_tc_bwd(const float* Wzh_img, const float* Wh_img, const float* Y, const float* Z,
                               const float* gates, const float* gY, const float* gZ, int g_last_only, float* dG,
                               float* dL, float* dy, float* dz, float dt, int T, int t_begin, int t_end, int N,
                               int Npad, int mode, cudaStream_t stream) {
  if (T < 0 || N < 0 || Npad < N || (Npad % LT_NODES) || t_begin < 0 || t_end > T || t_begin > t_end) return MSMP_ERR_ARG;
  if (t_begin == t_end || N == 0) return MSMP_OK;
  LemBwdParams p{Wzh_img, Wh_img, Y, Z, gates, gY, gZ, g_last_only, dG, dL, dy, dz, dt, T, N, Npad, t_begin, t_end};
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(k_lem_bwd_tc<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, LT_SMEM) != cudaSuccess ||
        cudaFuncSetAttribute(k_lem_bwd_tc<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, LT_SMEM) != cudaSuccess)
      return MSMP_ERR_CUDA;
    attr_set = true;
  }
  if (mode)
    k_lem_bwd_tc<true><<<Npad / LT_NODES, LB_THREADS, LT_SMEM, stream>>>(p);
  else
    k_lem_bwd_tc<false><<<Npad / LT_NODES, LB_THREADS, LT_SMEM, stream>>>(p);
  MSMP_CHECK_LAUNCH();
  return MSMP_OK;
}
