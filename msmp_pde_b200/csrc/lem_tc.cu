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

// pre[m][0:512] = [bias | bias_z] + inp[m][0:ninp] * [Wt_in | Wzt_in]   (m over T*N rows; exact fp32, memory bound)
__global__ void __launch_bounds__(256) k_lem_inproj(const float* __restrict__ inp, const float* __restrict__ Wt_in,
                                                    const float* __restrict__ Wzt_in, const float* __restrict__ bias,
                                                    const float* __restrict__ bias_z, float* __restrict__ pre,
                                                    size_t rows, int ninp) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;      // one float4 of one row
  if (idx >= rows * 128) return;
  const size_t m = idx >> 7;
  const int c = (int)(idx & 127) * 4;                                    // column 0..508
  const bool g = c < 384;
  const float* W = g ? Wt_in + c : Wzt_in + (c - 384);
  const int ldw = g ? 384 : 128;
  float4 acc = g ? ldg4(bias + c) : ldg4(bias_z + (c - 384));
  const float* x = inp + m * 32;
  for (int q = 0; q < ninp; ++q) {
    const float xv = __ldg(x + q);
    const float4 w = ldg4(W + (size_t)q * ldw);
    acc.x = fmaf(xv, w.x, acc.x);
    acc.y = fmaf(xv, w.y, acc.y);
    acc.z = fmaf(xv, w.z, acc.z);
    acc.w = fmaf(xv, w.w, acc.w);
  }
  st4(pre + m * 512 + c, acc);
}

struct LemFwdParams {
  const float* pre;      // [T][N][512]  bias + input projection (G0 | G1 | G2 | L)
  const float* Wimg;     // images of Wt[:128]  [128 x 384]: [3 ntiles][4 chunks][2][4096]
  const float* Wzimg;    // images of Wzt[:128] [128 x 128]: [1][4][2][4096]
  float* Y;              // [T+1][N][128]  (Y[0] = y0 on entry)
  float* Z;              // [T+1][N][128]  (Z[0] = z0 on entry)
  float* gates;          // [T][4][N][128]  a, b, zc, tL
  float dt;
  int T; int N;
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

  // epilogue ownership: thread = row r, 64 channels [c0, c0+64)
  const int r = 32 * (warp & 3) + lane;
  const int grow = row0 + r;
  const bool live = grow < p.N;
  const int c0 = 64 * (warp >> 2);
  const uint32_t tlane = (uint32_t)(32 * (warp & 3)) << 16;

  // y_{-1} tile from Y[0]
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
    // ---- G = y W_h^T : 3 n-tiles x 4 chunks -> TMEM columns 0..383
    gemm_phase(rg, p.Wimg, 0, 12, smA, tmem, 0, false, acc, nacc);
    // ---- gate_z: thread (row r, channels c0..c0+63)
    float* g_t = p.gates + (size_t)t * 4 * plane;
    const float* pre_t = p.pre + ((size_t)t * p.N + (live ? grow : 0)) * 512;
#pragma unroll 1
    for (int cb = 0; cb < 2; ++cb) {
      const int cc = c0 + 32 * cb;
      float v0[32], v1[32], v2[32];
      __syncwarp();
      tmem_ld32(tmem + tlane + (uint32_t)cc, v0);
      tmem_ld32(tmem + tlane + (uint32_t)(128 + cc), v1);
      tmem_ld32(tmem + tlane + (uint32_t)(256 + cc), v2);
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const int col = cc + j;
        float4 zn = zero4();
        if (live) {
          const float4 b0 = ldg4(pre_t + col), b1 = ldg4(pre_t + 128 + col), b2 = ldg4(pre_t + 256 + col);
          const float4 zp = ldcg4(p.Z + (size_t)t * plane + (size_t)grow * 128 + col);
          float4 a, b, zc;
          a.x = p.dt * sigmoidf_(v0[j] + b0.x); a.y = p.dt * sigmoidf_(v0[j + 1] + b0.y);
          a.z = p.dt * sigmoidf_(v0[j + 2] + b0.z); a.w = p.dt * sigmoidf_(v0[j + 3] + b0.w);
          b.x = p.dt * sigmoidf_(v1[j] + b1.x); b.y = p.dt * sigmoidf_(v1[j + 1] + b1.y);
          b.z = p.dt * sigmoidf_(v1[j + 2] + b1.z); b.w = p.dt * sigmoidf_(v1[j + 3] + b1.w);
          zc.x = tanhf(v2[j] + b2.x); zc.y = tanhf(v2[j + 1] + b2.y);
          zc.z = tanhf(v2[j + 2] + b2.z); zc.w = tanhf(v2[j + 3] + b2.w);
          zn = make_float4((1.f - b.x) * zp.x + b.x * zc.x, (1.f - b.y) * zp.y + b.y * zc.y,
                           (1.f - b.z) * zp.z + b.z * zc.z, (1.f - b.w) * zp.w + b.w * zc.w);
          const size_t o = (size_t)grow * 128 + col;
          st4(g_t + o, a);
          st4(g_t + plane + o, b);
          st4(g_t + 2 * plane + o, zc);
          st4(p.Z + (size_t)(t + 1) * plane + o, zn);
        }
        state_store4(smA, r, col, zn);          // z_t becomes the A operand of the L GEMM
      }
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    // ---- L = z Wz_h^T : 4 chunks -> TMEM columns 384..511
    gemm_phase(rg, p.Wzimg, 0, 4, smA, tmem, 384, false, acc, nacc);
    // ---- gate_y
#pragma unroll 1
    for (int cb = 0; cb < 2; ++cb) {
      const int cc = c0 + 32 * cb;
      float v[32];
      __syncwarp();
      tmem_ld32(tmem + tlane + (uint32_t)(384 + cc), v);
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const int col = cc + j;
        float4 yn = zero4();
        if (live) {
          const size_t o = (size_t)grow * 128 + col;
          const float4 bz = ldg4(pre_t + 384 + col);
          const float4 a = ldcg4(g_t + o);
          const float4 yp = ldcg4(p.Y + (size_t)t * plane + o);
          float4 tl = make_float4(tanhf(v[j] + bz.x), tanhf(v[j + 1] + bz.y), tanhf(v[j + 2] + bz.z), tanhf(v[j + 3] + bz.w));
          yn = make_float4((1.f - a.x) * yp.x + a.x * tl.x, (1.f - a.y) * yp.y + a.y * tl.y,
                           (1.f - a.z) * yp.z + a.z * tl.z, (1.f - a.w) * yp.w + a.w * tl.w);
          st4(g_t + 3 * plane + o, tl);
          st4(p.Y + (size_t)(t + 1) * plane + o, yn);
        }
        state_store4(smA, r, col, yn);          // y_t is the A operand of the next step's G GEMM
      }
    }
    // the loop head fences + syncs before the next MMAs read the state tile / overwrite TMEM
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------------------ backward
struct LemBwdParams {
  const float* Wzh_img;  // images of Wt := Wz[:, :128]  ([K = n][N = k]) : [1][4][2][4096]
  const float* Wh_img;   // images of Wt := W[:, :128]   ([K = 384][N = 128]) : [1][12][2][4096]
  const float* Y;        // [T+1][N][128]
  const float* Z;        // [T+1][N][128]
  const float* gates;    // [T][4][N][128]
  const float* gY;       // [T][N][128] external gradients (nullable)
  const float* gZ;       // [T][N][128] (nullable)
  float* dG;             // [T][N][384]
  float* dL;             // [T][N][128]
  float* dy;             // [N][128] carried gradient (zero on entry; d/dy0 on exit)
  float* dz;             // [N][128]
  float dt;
  int T; int N;
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
  const size_t plane = (size_t)p.N * 128;

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
  const int grow = row0 + r;
  const bool live = grow < p.N;
  const int c0 = 64 * (warp >> 2);
  const uint32_t tlane = (uint32_t)(32 * (warp & 3)) << 16;

  for (int t = p.T - 1; t >= 0; --t) {
    const float* g_t = p.gates + (size_t)t * 4 * plane;
    float* dG_t = p.dG + (size_t)t * p.N * 384;
    // ---- bwd_y (thread = row r, channels c0..c0+63)
#pragma unroll 1
    for (int j = 0; j < 64; j += 4) {
      const int col = c0 + j;
      float4 dl = zero4();
      if (live) {
        const size_t o = (size_t)grow * 128 + col;
        float4 d = ldcg4(p.dy + o);
        if (p.gY) d = add4(d, ldg4(p.gY + (size_t)t * plane + o));
        const float4 a = ldg4(g_t + o), tl = ldg4(g_t + 3 * plane + o);
        const float4 yp = ldg4(p.Y + (size_t)t * plane + o);
        dl = make_float4(d.x * a.x * (1.f - tl.x * tl.x), d.y * a.y * (1.f - tl.y * tl.y),
                         d.z * a.z * (1.f - tl.z * tl.z), d.w * a.w * (1.f - tl.w * tl.w));
        const float4 dg0 = make_float4(d.x * (tl.x - yp.x) * a.x * (1.f - a.x / p.dt), d.y * (tl.y - yp.y) * a.y * (1.f - a.y / p.dt),
                                       d.z * (tl.z - yp.z) * a.z * (1.f - a.z / p.dt), d.w * (tl.w - yp.w) * a.w * (1.f - a.w / p.dt));
        st4(p.dL + (size_t)t * plane + o, dl);
        st4(dG_t + (size_t)grow * 384 + col, dg0);
        st4(p.dy + o, make_float4(d.x * (1.f - a.x), d.y * (1.f - a.y), d.z * (1.f - a.z), d.w * (1.f - a.w)));
      }
      state_store4(smA, r, col, dl);
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    // ---- acc1 = dL Wz[:, :128]   -> TMEM columns 0..127
    gemm_phase(rg, p.Wzh_img, 0, 4, smA, tmem, 0, false, acc, nacc);
    // ---- bwd_z, and stage dG0 for the first block of the dy GEMM
#pragma unroll 1
    for (int cb = 0; cb < 2; ++cb) {
      const int cc = c0 + 32 * cb;
      float v[32];
      __syncwarp();
      tmem_ld32(tmem + tlane + (uint32_t)cc, v);
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const int col = cc + j;
        float4 g0 = zero4();
        if (live) {
          const size_t o = (size_t)grow * 128 + col;
          float4 d = add4(ldcg4(p.dz + o), make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]));
          if (p.gZ) d = add4(d, ldg4(p.gZ + (size_t)t * plane + o));
          const float4 b = ldg4(g_t + plane + o), zc = ldg4(g_t + 2 * plane + o);
          const float4 zp = ldg4(p.Z + (size_t)t * plane + o);
          st4(dG_t + (size_t)grow * 384 + 128 + col,
              make_float4(d.x * (zc.x - zp.x) * b.x * (1.f - b.x / p.dt), d.y * (zc.y - zp.y) * b.y * (1.f - b.y / p.dt),
                          d.z * (zc.z - zp.z) * b.z * (1.f - b.z / p.dt), d.w * (zc.w - zp.w) * b.w * (1.f - b.w / p.dt)));
          st4(dG_t + (size_t)grow * 384 + 256 + col,
              make_float4(d.x * b.x * (1.f - zc.x * zc.x), d.y * b.y * (1.f - zc.y * zc.y), d.z * b.z * (1.f - zc.z * zc.z),
                          d.w * b.w * (1.f - zc.w * zc.w)));
          st4(p.dz + o, make_float4(d.x * (1.f - b.x), d.y * (1.f - b.y), d.z * (1.f - b.z), d.w * (1.f - b.w)));
          g0 = ldcg4(dG_t + (size_t)grow * 384 + col);      // written by this thread in bwd_y
        }
        state_store4(smA, r, col, g0);
      }
    }
    // ---- acc2 = [dG0 | dG1 | dG2] W[:, :128]  -> TMEM columns 128..255, A tile restaged per block
    for (int blk = 0; blk < 3; ++blk) {
      if (blk > 0) {
#pragma unroll 1
        for (int j = 0; j < 64; j += 4) {
          const int col = c0 + j;
          float4 g = live ? ldcg4(dG_t + (size_t)grow * 384 + 128 * blk + col) : zero4();
          state_store4(smA, r, col, g);
        }
      }
      fence_proxy_async();
      tc_fence_before();
      __syncthreads();
      gemm_phase(rg, p.Wh_img, 4 * blk, 4, smA, tmem, 128, blk != 0, acc, nacc);
    }
    // ---- dy += acc2
#pragma unroll 1
    for (int cb = 0; cb < 2; ++cb) {
      const int cc = c0 + 32 * cb;
      float v[32];
      __syncwarp();
      tmem_ld32(tmem + tlane + (uint32_t)(128 + cc), v);
      if (live) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const size_t o = (size_t)grow * 128 + cc + j;
          st4(p.dy + o, add4(ldcg4(p.dy + o), make_float4(v[j], v[j + 1], v[j + 2], v[j + 3])));
        }
      }
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

// inp [T][N][32]; Wt_in = rows 128.. of the k-major W^T pack ([ninp..][384]); Wzt_in likewise ([..][128]);
// pre [T][N][512] scratch; Wimg / Wzimg = images of the STATE rows only (Wt[:128], Wzt[:128]).
extern "C" int msmp_lem_tc_fwd(const float* inp, int ninp, const float* Wt_in, const float* Wzt_in, const float* Wimg,
                               const float* Wzimg, const float* bias, const float* bias_z, float* pre, float* Y,
                               float* Z, float* gates, float dt, int T, int N, cudaStream_t stream) {
  if (T < 0 || N < 0 || ninp < 0 || ninp > 32) return MSMP_ERR_ARG;
  if (T == 0 || N == 0) return MSMP_OK;
  const size_t rows = (size_t)T * N;
  k_lem_inproj<<<(unsigned)((rows * 128 + 255) / 256), 256, 0, stream>>>(inp, Wt_in, Wzt_in, bias, bias_z, pre, rows, ninp);
  MSMP_CHECK_LAUNCH();
  LemFwdParams p{pre, Wimg, Wzimg, Y, Z, gates, dt, T, N};
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(k_lem_fwd_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, LT_SMEM) != cudaSuccess)
      return MSMP_ERR_CUDA;
    attr_set = true;
  }
  k_lem_fwd_tc<<<(N + 127) / 128, 256, LT_SMEM, stream>>>(p);
  MSMP_CHECK_LAUNCH();
  return MSMP_OK;
}

extern "C" int msmp_lem_tc_bwd(const float* Wzh_img, const float* Wh_img, const float* Y, const float* Z,
                               const float* gates, const float* gY, const float* gZ, float* dG, float* dL, float* dy,
                               float* dz, float dt, int T, int N, cudaStream_t stream) {
  if (T < 0 || N < 0) return MSMP_ERR_ARG;
  if (T == 0 || N == 0) return MSMP_OK;
  LemBwdParams p{Wzh_img, Wh_img, Y, Z, gates, gY, gZ, dG, dL, dy, dz, dt, T, N};
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(k_lem_bwd_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, LT_SMEM) != cudaSuccess)
      return MSMP_ERR_CUDA;
    attr_set = true;
  }
  k_lem_bwd_tc<<<(N + 127) / 128, 256, LT_SMEM, stream>>>(p);
  MSMP_CHECK_LAUNCH();
  return MSMP_OK;
}
