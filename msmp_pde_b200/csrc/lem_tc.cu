// Persistent LEM recurrence on the tensor cores (replaces lem_cuda.forward / lem_cuda.backward,
// experiments/models_gnn.py:290-292,300).  The recurrence is independent per node, so one CTA owns 128 nodes and
// walks all T time steps inside ONE launch: no inter-CTA synchronisation, the state tiles never leave the SM.
//
// forward, per step t.  The input part of both affine maps is hoisted out of the recurrence:
// pre[t][n][0:512] = [b | bz] + I_t [W_in | Wz_in]^T (k_lem_inproj, memory bound, exact fp32), then
//   G[128 x 384] = y_{t-1} W_h^T                12 weight chunks (3 n-tiles x 4 k-chunks), TMEM columns 0..383
//   gate_z      : a = dt sig(G0 + pre), b = dt sig(G1 + pre), zc = tanh(G2 + pre), z_t = (1-b) z_{t-1} + b zc
//   L[128 x 128] = z_t Wz_h^T                   4 weight chunks, TMEM columns 384..511
//   gate_y      : tL = tanh(L + pre), y_t = (1-a) y_{t-1} + a tL
// The state operand (y, then z, then y again) lives in shared memory as a tf32 hi/lo tile image written by the
// gate epilogues.  The weights (pre-split, pre-swizzled images, 512 KiB per step) are streamed from L2 through a
// 3-stage ring: four loader warps copy each 32 KiB chunk with 16 x LDG.128 in flight per thread (a single bulk copy
// per chunk left the tensor pipe waiting ~2.4 us per chunk on copy latency), one thread issues the MMAs, all eight
// warps run the gate epilogues.
//
// backward, per step t = T-1..0 (dy, dz carried in global scratch, owned row-wise by the same thread):
//   bwd_y : d = dy + gY[t]; dL = d a (1-tL^2); dG0 = d (tL - y_{t-1}) a (1 - a/dt); dy = d (1-a)
//   acc1  = dL Wz[:, :128]                        4 chunks
//   bwd_z : d = dz + gZ[t] + acc1; dG1 = d (zc - z_{t-1}) b (1 - b/dt); dG2 = d b (1-zc^2); dz = d (1-b)
//   acc2  = [dG0 | dG1 | dG2] W[:, :128]          3 x 4 chunks (the A tile is restaged per 128-column block)
//   dy   += acc2
// dG [T,N,384] and dL [T,N,128] are written for the four weight-gradient GEMMs (msmp_linear_wgrad_tc).
//
// Memory access.  The gate epilogues own one node row per thread (a TMEM lane), so row-major global arrays would be
// touched 16 bytes per thread at a 512-byte stride (32 lines per warp request; measured: 77 us per step).  Every array
// private to the recurrence (pre, gates, the y/z history used for y_{t-1}/z_{t-1}, the carried dy/dz, the dG0/dG2
// scratch) is therefore kept LANE-MAJOR: element (row n, channel c) of a C-channel array lives at
//     ((n / 32) * C + c) * 32 + n % 32
// so the 32 lanes of a warp (32 consecutive rows) read/write one contiguous 128-byte line per channel.  Arrays
// that other kernels consume row-major (Y, Z, dL, dG) are written by a cooperative, fully coalesced copy-out of the
// state tile image (hi + lo reconstructs the fp32 value exactly).
#include "umma.cuh"
#include "msmp_b200.h"

namespace msmp {

constexpr int LT_A_BYTES = 4 * 2 * IMG_BYTES;      // state tile: 4 k-chunks x (hi | lo) = 128 KiB
constexpr int LT_B_BYTES = 2 * IMG_BYTES;          // one weight chunk (hi | lo) = 32 KiB
constexpr int LT_STAGES = 3;
constexpr int LT_SMEM = LT_A_BYTES + LT_STAGES * LT_B_BYTES + 1024 + 256;
constexpr int LT_LOADERS = 128;                    // warps 4..7

// Weight ring shared by the loader warps (producers) and the MMA-issuing thread (consumer).  Every thread keeps
// the same running chunk counter `n`, so stage / parity bookkeeping needs no communication.
struct Ring {
  uint8_t* smB;        // LT_STAGES stages
  uint64_t* bfull;     // [LT_STAGES], LT_LOADERS arrivals
  uint64_t* bfree;     // [LT_STAGES], one arrival (tcgen05.commit)
  uint32_t n;
};

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// loader thread lt (0..127): copy chunk `i` (32 KiB at src) into its ring stage
__device__ __forceinline__ void ring_load(const Ring& rg, uint32_t i, const float* src, int lt) {
  const uint32_t s = i % LT_STAGES, use = i / LT_STAGES;
  if (use > 0) mbar_wait_warp(&rg.bfree[s], (use - 1) & 1);       // MMAs that read this stage are complete
  float4 v[16];
  const float4* g = reinterpret_cast<const float4*>(src) + lt;
#pragma unroll
  for (int q = 0; q < 16; ++q) v[q] = __ldg(g + q * LT_LOADERS);
  float4* d = reinterpret_cast<float4*>(rg.smB + s * LT_B_BYTES) + lt;
#pragma unroll
  for (int q = 0; q < 16; ++q) d[q * LT_LOADERS] = v[q];
  fence_proxy_async();
  mbar_arrive(&rg.bfull[s]);
}

__device__ __forceinline__ void ring_mma(const Ring& rg, uint32_t i, uint32_t a_img /* smem addr of (hi|lo) A chunk */,
                                         uint32_t tmem_d, bool accumulate) {
  const uint32_t s = i % LT_STAGES, use = i / LT_STAGES;
  mbar_wait(&rg.bfull[s], use & 1);
  tc_fence_after();
  constexpr uint32_t IDESC = umma_idesc_tf32(128, 128, 0, 0);
  const uint32_t a_hi = a_img, a_lo = a_img + IMG_BYTES;
  const uint32_t b_hi = smem_u32(rg.smB + s * LT_B_BYTES), b_lo = b_hi + IMG_BYTES;
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

// One GEMM phase: `nchunks` weight chunks (image index w0 + j); chunk j multiplies state-tile chunk (j & 3) into TMEM
// columns dcol + 128 * (j >> 2) (first chunk of every 128-column block overwrites unless `acc_first`).  Called by ALL
// threads; returns after the accumulator is complete.
__device__ __forceinline__ void gemm_phase(Ring& rg, const float* wimg, uint32_t w0, uint32_t nchunks, uint8_t* smA,
                                           uint32_t tmem, uint32_t dcol, bool acc_first, uint64_t* acc, uint32_t& nacc) {
  const int tid = threadIdx.x;
  const uint32_t base = rg.n;
  if (tid >= 256 - LT_LOADERS) {
    const int lt = tid - (256 - LT_LOADERS);
    for (uint32_t j = 0; j < nchunks; ++j) ring_load(rg, base + j, wimg + (size_t)(w0 + j) * (LT_B_BYTES / 4), lt);
  } else if (tid == 0) {
    tc_fence_after();
    for (uint32_t j = 0; j < nchunks; ++j)
      ring_mma(rg, base + j, smem_u32(smA + (j & 3) * 2 * IMG_BYTES), tmem + dcol + 128 * (j >> 2), acc_first || (j & 3) != 0);
    umma_commit(acc);
    mbar_wait(acc, nacc & 1);          // only this thread polls; everyone else parks at the block barrier below
  }
  rg.n = base + nchunks;
  ++nacc;
  __syncthreads();
  tc_fence_after();
}

// write 4 consecutive values of row r, columns col..col+3 (col % 4 == 0, col < 128) into the state tile image
__device__ __forceinline__ void state_store4(uint8_t* smA, int r, int col, float4 v) {
  uint8_t* chunk = smA + (col >> 5) * (2 * IMG_BYTES);
  store_split4(chunk, chunk + IMG_BYTES, img_off(r, (col & 31) >> 2), v);
}

// clock64() phase stamps of CTA 0 at step 2 (scripts/lem_ticks.py); compiled in only with -DMSMP_LEM_TICKS.
#ifdef MSMP_LEM_TICKS
__device__ long long g_lem_dbg[64];
#define LEM_TICK(i) do { if (blockIdx.x == 0 && tid == 0 && t == 2) g_lem_dbg[i] = clock64(); } while (0)
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

// cooperative coalesced copy of the state tile (hi + lo) to a row-major array: dst[(row0 + r) * ld + c], c < 128
__device__ __forceinline__ void image_to_global(const uint8_t* smA, float* dst, int ld, int row0, int N) {
  const int tid = threadIdx.x;
#pragma unroll 4
  for (int i = 0; i < 16; ++i) {
    const int idx = tid + 256 * i;
    const int rr = idx >> 5, c4 = idx & 31;
    const uint8_t* chunk = smA + (c4 >> 3) * (2 * IMG_BYTES);
    const uint32_t off = img_off(rr, c4 & 7);
    const float4 h = *reinterpret_cast<const float4*>(chunk + off);
    const float4 l = *reinterpret_cast<const float4*>(chunk + IMG_BYTES + off);
    if (row0 + rr < N) st4(dst + (size_t)(row0 + rr) * ld + 4 * c4, add4(h, l));
  }
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

__global__ void __launch_bounds__(256, 1) k_lem_fwd_tc(const LemFwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smA = smem;
  uint8_t* smB = smem + LT_A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smB + LT_STAGES * LT_B_BYTES);   // bfull[3], bfree[3], acc
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row0 = blockIdx.x * 128;
  const size_t plane = (size_t)p.N * 128;
  const size_t ntile = p.Npad / 32;

  if (warp == 0) tmem_alloc(tmem_slot, 512);
  if (tid == 32) {
    for (int i = 0; i < LT_STAGES; ++i) {
      mbar_init(&bars[i], LT_LOADERS);
      mbar_init(&bars[LT_STAGES + i], 1);
    }
    mbar_init(&bars[2 * LT_STAGES], 1);
    fence_barrier_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  Ring rg{smB, &bars[0], &bars[LT_STAGES], 0};
  uint64_t* acc = &bars[2 * LT_STAGES];
  uint32_t nacc = 0;

  // epilogue ownership: thread = row r (TMEM lane), 64 channels [c0, c0+64); gt = its 32-row tile
  const int r = 32 * (warp & 3) + lane;
  const int c0 = 64 * (warp >> 2);
  const size_t gt = (size_t)blockIdx.x * 4 + (warp & 3);
  const uint32_t tlane = (uint32_t)(32 * (warp & 3)) << 16;

  // y_{-1} tile from the row-major Y[0]
  for (int i = 0; i < 16; ++i) {
    const int idx = tid + 256 * i;
    const int rr = idx >> 5, c4 = idx & 31;
    const int g = row0 + rr;
    float4 v = (g < p.N) ? ldg4(p.Y + (size_t)g * 128 + 4 * c4) : zero4();
    state_store4(smA, rr, 4 * c4, v);
  }

  for (int t = 0; t < p.T; ++t) {
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    // pull this step's input-projection lines (HBM) into L2 while the G GEMM runs
    {
      const float* pre_pf = p.pre + ((size_t)t * ntile) * 512 * 32;
#pragma unroll
      for (int q = 0; q < 4; ++q) prefetch_lm(pre_pf, gt, 512, 128 * q + c0, 64, lane);
    }
    // ---- G = y W_h^T : 3 n-tiles x 4 chunks -> TMEM columns 0..383
    LEM_TICK(0);
    gemm_phase(rg, p.Wimg, 0, 12, smA, tmem, 0, false, acc, nacc);
    LEM_TICK(1);
    // ---- gate_z
    const float* pre_t = p.pre + ((size_t)t * ntile) * 512 * 32;
    float* g_t = p.gates + ((size_t)t * ntile) * 512 * 32;
    const float* zprev = p.Zt + ((size_t)t * ntile) * 128 * 32;
    float* znext = p.Zt + ((size_t)(t + 1) * ntile) * 128 * 32;
#pragma unroll 1
    for (int cb = 0; cb < 2; ++cb) {
      const int cc = c0 + 32 * cb;
      float v0[32], v1[32], v2[32];
      __syncwarp();
      tmem_ld32(tmem + tlane + (uint32_t)cc, v0);
      tmem_ld32(tmem + tlane + (uint32_t)(128 + cc), v1);
      tmem_ld32(tmem + tlane + (uint32_t)(256 + cc), v2);
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        // batch the 32 loads of 8 channels before any dependent math (the epilogue is latency bound otherwise)
        float p0[8], p1[8], p2[8], zp[8], zn[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int c = cc + j + e;
          p0[e] = __ldg(pre_t + lm(gt, 512, c, lane));
          p1[e] = __ldg(pre_t + lm(gt, 512, 128 + c, lane));
          p2[e] = __ldg(pre_t + lm(gt, 512, 256 + c, lane));
          zp[e] = __ldcg(zprev + lm(gt, 128, c, lane));
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int c = cc + j + e;
          const float a = p.dt * sigmoidf_(v0[j + e] + p0[e]);
          const float b = p.dt * sigmoidf_(v1[j + e] + p1[e]);
          const float zc = tanh_acc(v2[j + e] + p2[e]);
          zn[e] = (1.f - b) * zp[e] + b * zc;
          g_t[lm(gt, 512, c, lane)] = a;
          g_t[lm(gt, 512, 128 + c, lane)] = b;
          g_t[lm(gt, 512, 256 + c, lane)] = zc;
          znext[lm(gt, 128, c, lane)] = zn[e];
        }
        state_store4(smA, r, cc + j, make_float4(zn[0], zn[1], zn[2], zn[3]));    // z_t: A operand of the L GEMM
        state_store4(smA, r, cc + j + 4, make_float4(zn[4], zn[5], zn[6], zn[7]));
      }
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    LEM_TICK(2);
    image_to_global(smA, p.Z + (size_t)(t + 1) * plane, 128, row0, p.N);
    LEM_TICK(3);
    // ---- L = z Wz_h^T : 4 chunks -> TMEM columns 384..511
    gemm_phase(rg, p.Wzimg, 0, 4, smA, tmem, 384, false, acc, nacc);
    LEM_TICK(4);
    // ---- gate_y
    const float* yprev = p.Yt + ((size_t)t * ntile) * 128 * 32;
    float* ynext = p.Yt + ((size_t)(t + 1) * ntile) * 128 * 32;
#pragma unroll 1
    for (int cb = 0; cb < 2; ++cb) {
      const int cc = c0 + 32 * cb;
      float v[32];
      __syncwarp();
      tmem_ld32(tmem + tlane + (uint32_t)(384 + cc), v);
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        float pz[8], av[8], yp[8], yn[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int c = cc + j + e;
          pz[e] = __ldg(pre_t + lm(gt, 512, 384 + c, lane));
          av[e] = __ldcg(g_t + lm(gt, 512, c, lane));
          yp[e] = __ldcg(yprev + lm(gt, 128, c, lane));
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int c = cc + j + e;
          const float tl = tanh_acc(v[j + e] + pz[e]);
          yn[e] = (1.f - av[e]) * yp[e] + av[e] * tl;
          g_t[lm(gt, 512, 384 + c, lane)] = tl;
          ynext[lm(gt, 128, c, lane)] = yn[e];
        }
        state_store4(smA, r, cc + j, make_float4(yn[0], yn[1], yn[2], yn[3]));    // y_t: A operand of the next G GEMM
        state_store4(smA, r, cc + j + 4, make_float4(yn[4], yn[5], yn[6], yn[7]));
      }
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    LEM_TICK(5);
    image_to_global(smA, p.Y + (size_t)(t + 1) * plane, 128, row0, p.N);
    LEM_TICK(6);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
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

__global__ void __launch_bounds__(256, 1) k_lem_bwd_tc(const LemBwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smA = smem;
  uint8_t* smB = smem + LT_A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smB + LT_STAGES * LT_B_BYTES);    // bfull[3], bfree[3], acc
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row0 = blockIdx.x * 128;
  const size_t ntile = p.Npad / 32;

  if (warp == 0) tmem_alloc(tmem_slot, 256);
  if (tid == 32) {
    for (int i = 0; i < LT_STAGES; ++i) {
      mbar_init(&bars[i], LT_LOADERS);
      mbar_init(&bars[LT_STAGES + i], 1);
    }
    mbar_init(&bars[2 * LT_STAGES], 1);
    fence_barrier_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  Ring rg{smB, &bars[0], &bars[LT_STAGES], 0};
  uint64_t* acc = &bars[2 * LT_STAGES];
  uint32_t nacc = 0;

  const int r = 32 * (warp & 3) + lane;
  const int c0 = 64 * (warp >> 2);
  const size_t gt = (size_t)blockIdx.x * 4 + (warp & 3);
  const uint32_t tlane = (uint32_t)(32 * (warp & 3)) << 16;
  const float inv_dt = 1.0f / p.dt;

  // stage a lane-major 128-channel scratch slab (this thread's row, its 64 channels) into the state tile
  auto stage_lm = [&](const float* slab) {
#pragma unroll 1
    for (int j = 0; j < 64; j += 16) {
      float g[16];
#pragma unroll
      for (int e = 0; e < 16; ++e) g[e] = __ldcg(slab + lm(gt, 128, c0 + j + e, lane));
#pragma unroll
      for (int e = 0; e < 16; e += 4) state_store4(smA, r, c0 + j + e, make_float4(g[e], g[e + 1], g[e + 2], g[e + 3]));
    }
  };
  auto publish = [&]() {       // make the freshly written state tile visible to the async proxy / other threads
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
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
      for (int q = 0; q < 4; ++q) prefetch_lm(g_n, gt, 512, 128 * q + c0, 64, lane);
      prefetch_lm(p.Yt + ((size_t)(t - 1) * ntile) * 128 * 32, gt, 128, c0, 64, lane);
      prefetch_lm(p.Zt + ((size_t)(t - 1) * ntile) * 128 * 32, gt, 128, c0, 64, lane);
    }
    // ---- bwd_y : dL -> state tile, dG0 -> s0, dy <- d (1 - a)
#pragma unroll 1
    for (int j = 0; j < 64; j += 8) {
      float dv[8], av[8], tv[8], yv[8], dl[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int c = c0 + j + e;
        dv[e] = __ldcg(p.dyt + lm(gt, 128, c, lane));
        if (gy) dv[e] += __ldg(gy + lm(gt, 128, c, lane));
        av[e] = __ldg(g_t + lm(gt, 512, c, lane));
        tv[e] = __ldg(g_t + lm(gt, 512, 384 + c, lane));
        yv[e] = __ldg(yprev + lm(gt, 128, c, lane));
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int c = c0 + j + e;
        const float d = dv[e], a = av[e], tl = tv[e];
        dl[e] = d * a * (1.f - tl * tl);
        p.s0[lm(gt, 128, c, lane)] = d * (tl - yv[e]) * a * (1.f - a * inv_dt);
        p.dyt[lm(gt, 128, c, lane)] = d * (1.f - a);
      }
      state_store4(smA, r, c0 + j, make_float4(dl[0], dl[1], dl[2], dl[3]));
      state_store4(smA, r, c0 + j + 4, make_float4(dl[4], dl[5], dl[6], dl[7]));
    }
    publish();
    image_to_global(smA, p.dL + (size_t)t * p.N * 128, 128, row0, p.N);
    // ---- acc1 = dL Wz[:, :128]   -> TMEM columns 0..127
    gemm_phase(rg, p.Wzh_img, 0, 4, smA, tmem, 0, false, acc, nacc);
    // ---- bwd_z : dG1 -> state tile, dG2 -> s2, dz <- d (1 - b)
#pragma unroll 1
    for (int cb = 0; cb < 2; ++cb) {
      const int cc = c0 + 32 * cb;
      float v[32];
      __syncwarp();
      tmem_ld32(tmem + tlane + (uint32_t)cc, v);
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        float dv[8], bv[8], zcv[8], zpv[8], g1[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int c = cc + j + e;
          dv[e] = __ldcg(p.dzt + lm(gt, 128, c, lane)) + v[j + e];
          if (gz) dv[e] += __ldg(gz + lm(gt, 128, c, lane));
          bv[e] = __ldg(g_t + lm(gt, 512, 128 + c, lane));
          zcv[e] = __ldg(g_t + lm(gt, 512, 256 + c, lane));
          zpv[e] = __ldg(zprev + lm(gt, 128, c, lane));
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int c = cc + j + e;
          const float d = dv[e], b = bv[e], zc = zcv[e];
          g1[e] = d * (zc - zpv[e]) * b * (1.f - b * inv_dt);
          p.s2[lm(gt, 128, c, lane)] = d * b * (1.f - zc * zc);
          p.dzt[lm(gt, 128, c, lane)] = d * (1.f - b);
        }
        state_store4(smA, r, cc + j, make_float4(g1[0], g1[1], g1[2], g1[3]));
        state_store4(smA, r, cc + j + 4, make_float4(g1[4], g1[5], g1[6], g1[7]));
      }
    }
    // ---- acc2 = [dG1 | dG2 | dG0] W[:, :128]  -> TMEM columns 128..255 (weight chunks 4..7, 8..11, 0..3)
    publish();
    image_to_global(smA, dG_t + 128, 384, row0, p.N);
    gemm_phase(rg, p.Wh_img, 4, 4, smA, tmem, 128, false, acc, nacc);
    stage_lm(p.s2);
    publish();
    image_to_global(smA, dG_t + 256, 384, row0, p.N);
    gemm_phase(rg, p.Wh_img, 8, 4, smA, tmem, 128, true, acc, nacc);
    stage_lm(p.s0);
    publish();
    image_to_global(smA, dG_t, 384, row0, p.N);
    gemm_phase(rg, p.Wh_img, 0, 4, smA, tmem, 128, true, acc, nacc);
    // ---- dy += acc2
#pragma unroll 1
    for (int cb = 0; cb < 2; ++cb) {
      const int cc = c0 + 32 * cb;
      float v[32];
      __syncwarp();
      tmem_ld32(tmem + tlane + (uint32_t)(128 + cc), v);
      float cur[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) cur[j] = __ldcg(p.dyt + lm(gt, 128, cc + j, lane));
#pragma unroll
      for (int j = 0; j < 32; ++j) p.dyt[lm(gt, 128, cc + j, lane)] = cur[j] + v[j];
    }
    tc_fence_before();
    __syncthreads();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
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
  k_lem_fwd_tc<<<Npad / 128, 256, LT_SMEM, stream>>>(p);
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
  k_lem_bwd_tc<<<Npad / 128, 256, LT_SMEM, stream>>>(p);
  MSMP_CHECK_LAUNCH();
  return MSMP_OK;
}
