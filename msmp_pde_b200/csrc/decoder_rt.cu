// Output decoder for the tw = 25 geometry (K1 = 16, S1 = 3, L1 = 38, K2 = 14), register-tiled.
//
// Same contract as k_decoder_fwd / k_decoder_bwd (decoder.cu; models_gnn.py:208-224,275-279, models_gnn2D.py:382-391,448-458):
//   za[n][o][q]  = b1[o] + sum_c sum_j w1[o][c][j] h[n][c][3q + j]                   (Conv1d(C, 8, 16, stride 3))
//   out[n][c][k] = base + dt[k] (b2[c] + sum_o sum_j w2[c][o][j] swish(za)[n][o][k + j])   (Conv1d(8, C, 14))
// The first-generation kernels computed one output element per thread with two shared-memory loads per FMA (1/8 of the FFMA
// rate at best; C4 step: 0.60 + 1.76 ms, 5 TFLOP/s) and wrote one weight-gradient partial per 8 nodes (16 384 x 490 floats).
// Here every thread keeps a ROW of outputs (or a weight row's gradient) in registers, slides a register window along the
// input row (all indices compile-time constants after unrolling: 4 - 14 FMAs per shared-memory load), a CTA walks 32-node
// tiles persistently and carries its weight-gradient sums in registers across tiles: one partial per CTA, summed in a
// fixed order (deterministic).  Activations on the MUFU pipe (common.cuh).
#include <cstdlib>
#include "common.cuh"
#include "msmp_b200.h"

namespace msmp {

constexpr int DR_NB = 32;                  // nodes per tile
constexpr int DR_L1 = 38, DR_K1 = 16, DR_S1 = 3, DR_K2 = 14, DR_TW = 25, DR_OC = 8;
constexpr int DR_ZP = 305;                 // pitch of a node's [8][38] block (odd: lanes = nodes hit distinct banks)
constexpr int DR_HP = 132;                 // pitch of one (node, channel) row of h (128 + 4: rows 4 banks apart, 16-byte aligned)
constexpr int DR_DP = 27;                  // pitch of one (node, channel) row of dout * dt

struct DecRtFwd {
  const float* h; const float* w1; const float* b1; const float* w2; const float* b2;
  const float* u; int ldu; const float* dt; float* za; float* out; int N;
};
struct DecRtBwd {
  const float* dout; const float* h; const float* za; const float* w1; const float* w2; const float* dt;
  float* dh; float* part; int N;
};

template <int C> __host__ __device__ constexpr int dr_fwd_smem() { return (DR_NB * C * DR_HP + DR_NB * DR_ZP + DR_OC * C * DR_K1 + C * DR_OC * DR_K2 + 64) * 4; }
template <int C> __host__ __device__ constexpr int dr_nw() { return DR_OC * C * DR_K1 + DR_OC + C * DR_OC * DR_K2 + C; }
template <int C> __host__ __device__ constexpr int dr_bwd_smem() {
  return (DR_NB * C * DR_HP + 2 * DR_NB * DR_ZP + DR_NB * C * DR_DP + DR_OC * C * DR_K1 + C * DR_OC * DR_K2 + 32) * 4;
}

// h tile -> shared memory rows of pitch DR_HP (zero rows past N)
template <int C>
__device__ __forceinline__ void dr_load_h(float* sh, const float* h, int n0, int N, int tid) {
  for (int i = tid; i < DR_NB * C * 32; i += 256) {
    const int row = i >> 5, q4 = i & 31;                      // row = node * C + channel
    const int n = n0 + row / C;
    const float4 v = n < N ? ldg4(h + (size_t)n0 * C * 128 + (size_t)row * 128 + 4 * q4) : zero4();
    *reinterpret_cast<float4*>(sh + row * DR_HP + 4 * q4) = v;
  }
}

template <int C>
__global__ void __launch_bounds__(256, 1) k_decoder_fwd_rt(const DecRtFwd p, int ntiles) {
  extern __shared__ __align__(16) float dsm[];
  float* sh = dsm;                                 // [NB * C][HP]; later the staged outputs [NB][C * 25 + 1]
  float* sa = sh + DR_NB * C * DR_HP;              // [NB][ZP]: za, then swish(za)
  float* sw1 = sa + DR_NB * DR_ZP;                 // [8][C][16]
  float* sw2 = sw1 + DR_OC * C * DR_K1;            // [C][8][14]
  float* sdt = sw2 + C * DR_OC * DR_K2;            // [25] | b2 [C] at 32
  const int tid = threadIdx.x;
  for (int i = tid; i < DR_OC * C * DR_K1; i += 256) sw1[i] = p.w1[i];
  for (int i = tid; i < C * DR_OC * DR_K2; i += 256) sw2[i] = p.w2[i];
  if (tid < DR_TW) sdt[tid] = p.dt[tid];
  if (tid < C) sdt[32 + tid] = p.b2[tid];
  __syncthreads();
  // phase-1 role: (node n1, first-conv channel o1), weights of that channel in registers
  const int n1 = tid >> 3, o1 = tid & 7;
  float w1r[C][DR_K1];
#pragma unroll
  for (int c = 0; c < C; ++c)
#pragma unroll
    for (int j = 0; j < DR_K1; ++j) w1r[c][j] = sw1[(o1 * C + c) * DR_K1 + j];
  const float b1r = p.b1[o1];
  // phase-2 role: (node n2 = lane, channel c2, quarter kq of the 25 outputs); a warp shares (c2, kq)
  const int n2 = tid & 31, c2 = (tid >> 5) % C, kq = (tid >> 5) / C;
  const int k0 = kq == 0 ? 0 : 1 + 6 * kq;        // 0, 7, 13, 19 (7 + 6 + 6 + 6 outputs)

#pragma unroll 1
  for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int n0 = t * DR_NB;
    dr_load_h<C>(sh, p.h, n0, p.N, tid);
    __syncthreads();
    {  // ---- first convolution: a row of 38 outputs per thread, sliding 16-wide register window over h
      float acc[DR_L1];
#pragma unroll
      for (int q = 0; q < DR_L1; ++q) acc[q] = b1r;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const float* hp = sh + (n1 * C + c) * DR_HP;
        float win[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) win[j] = hp[j];
#pragma unroll
        for (int q = 0; q < DR_L1; ++q) {
#pragma unroll
          for (int j = 0; j < DR_K1; ++j) acc[q] = fmaf(w1r[c][j], win[(3 * q + j) & 15], acc[q]);
          if (q + 1 < DR_L1) {
#pragma unroll
            for (int j = 0; j < 3; ++j) win[(3 * q + j) & 15] = hp[3 * q + 16 + j];
          }
        }
      }
      float* ap = sa + n1 * DR_ZP + o1 * DR_L1;
#pragma unroll
      for (int q = 0; q < DR_L1; ++q) ap[q] = acc[q];
    }
    __syncthreads();
    // za -> global (kept for the backward pass), coalesced; then swish in place
    for (int i = tid; i < DR_NB * (DR_OC * DR_L1); i += 256) {
      const int n = i / (DR_OC * DR_L1), e = i - n * (DR_OC * DR_L1);
      const float z = sa[n * DR_ZP + e];
      if (n0 + n < p.N) p.za[(size_t)(n0 + n) * (DR_OC * DR_L1) + e] = z;
      sa[n * DR_ZP + e] = swish_m(z);
    }
    __syncthreads();
    // ---- second convolution: up to 7 outputs per thread, 20-wide window of a per first-conv channel
    if (kq < 4) {
      float acc[7];
#pragma unroll
      for (int i = 0; i < 7; ++i) acc[i] = sdt[32 + c2];
#pragma unroll
      for (int o = 0; o < DR_OC; ++o) {
        const float* ap = sa + n2 * DR_ZP + o * DR_L1 + k0;
        const float* wp = sw2 + (c2 * DR_OC + o) * DR_K2;
        float a[20];
#pragma unroll
        for (int i = 0; i < 20; ++i) a[i] = (k0 + i < DR_L1) ? ap[i] : 0.f;
#pragma unroll
        for (int j = 0; j < DR_K2; ++j) {
          const float w = wp[j];
#pragma unroll
          for (int i = 0; i < 7; ++i) acc[i] = fmaf(w, a[i + j], acc[i]);
        }
      }
      __syncthreads();                                      // every thread is done with sh (first convolution) ...
      float* so = sh + n2 * (C * DR_TW + 1) + c2 * DR_TW;   // ... which now stages the outputs
#pragma unroll
      for (int i = 0; i < 7; ++i)
        if (i < (kq == 0 ? 7 : 6)) so[k0 + i] = acc[i];
    } else {
      __syncthreads();
    }
    __syncthreads();
    for (int i = tid; i < DR_NB * C * DR_TW; i += 256) {
      const int n = i / (C * DR_TW), r = i - n * (C * DR_TW);
      if (n0 + n >= p.N) continue;
      const int c = r / DR_TW, k = r - c * DR_TW;
      const float* un = p.u + (size_t)(n0 + n) * p.ldu;
      const float base = (C == 1) ? un[DR_TW - 1] : un[r];
      p.out[(size_t)(n0 + n) * (C * DR_TW) + r] = base + sdt[k] * sh[n * (C * DR_TW + 1) + r];
    }
    __syncthreads();
  }
}

template <int C>
__global__ void __launch_bounds__(256, 1) k_decoder_bwd_rt(const DecRtBwd p, int ntiles) {
  extern __shared__ __align__(16) float dsm[];
  constexpr int NW = dr_nw<C>();
  constexpr int N_W1 = DR_OC * C * DR_K1, N_W2 = C * DR_OC * DR_K2;
  float* sh = dsm;                                 // [NB * C][HP]
  float* sz = sh + DR_NB * C * DR_HP;              // [NB][ZP]: za, then dza
  float* sa = sz + DR_NB * DR_ZP;                  // [NB][ZP]: swish(za); later the staged dh [NB][C * 128 + 1]
  float* sdd = sa + DR_NB * DR_ZP;                 // [NB * C][DP]: dout * dt
  float* sw1 = sdd + DR_NB * C * DR_DP;            // [8][C][16]
  float* sw2 = sw1 + N_W1;                         // [C][8][14]
  float* sdt = sw2 + N_W2;                         // [25]
  const int tid = threadIdx.x;
  for (int i = tid; i < N_W1; i += 256) sw1[i] = p.w1[i];
  for (int i = tid; i < N_W2; i += 256) sw2[i] = p.w2[i];
  if (tid < DR_TW) sdt[tid] = p.dt[tid];
  __syncthreads();
  // role A: (node nA, channel oA): da row; its second-conv weights in registers
  const int nA = tid >> 3, oA = tid & 7;
  float w2r[C][DR_K2];
#pragma unroll
  for (int c = 0; c < C; ++c)
#pragma unroll
    for (int j = 0; j < DR_K2; ++j) w2r[c][j] = sw2[(c * DR_OC + oA) * DR_K2 + j];
  // role B: (node nB = lane, channel cB, residue rB of the position): dh[3m + rB]; six taps w1[o][cB][3t + rB]
  const int nB = tid & 31, crB = tid >> 5, cB = crB / 3, rB = crB - 3 * cB;
  const bool doB = crB < 3 * C;
  // role C / D: (node pair gW, channel oW, channel cW): weight-gradient rows, carried across tiles
  const int gW = tid >> 4, oW = (tid & 15) >> 1, cW = tid & 1;
  const bool doW = cW < C;
  float acc1[DR_K1], acc2[DR_K2], accb1 = 0.f, accb2 = 0.f;
#pragma unroll
  for (int j = 0; j < DR_K1; ++j) acc1[j] = 0.f;
#pragma unroll
  for (int j = 0; j < DR_K2; ++j) acc2[j] = 0.f;

#pragma unroll 1
  for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int n0 = t * DR_NB;
    dr_load_h<C>(sh, p.h, n0, p.N, tid);
    for (int i = tid; i < DR_NB * (DR_OC * DR_L1); i += 256) {
      const int n = i / (DR_OC * DR_L1), e = i - n * (DR_OC * DR_L1);
      sz[n * DR_ZP + e] = n0 + n < p.N ? __ldg(p.za + (size_t)(n0 + n) * (DR_OC * DR_L1) + e) : 0.f;
    }
    for (int i = tid; i < DR_NB * C * DR_TW; i += 256) {
      const int row = i / DR_TW, k = i - row * DR_TW;       // row = node * C + channel
      const int n = n0 + row / C;
      sdd[row * DR_DP + k] = n < p.N ? __ldg(p.dout + (size_t)n0 * (C * DR_TW) + i) * sdt[k] : 0.f;
    }
    __syncthreads();
    {  // ---- A: da[l] = sum_c sum_k dd[c][k] w2[c][o][l - k]; dza = da * swish'(za); a = swish(za)
      float s[DR_L1];
#pragma unroll
      for (int l = 0; l < DR_L1; ++l) s[l] = 0.f;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const float* dp = sdd + (nA * C + c) * DR_DP;
#pragma unroll
        for (int k = 0; k < DR_TW; ++k) {
          const float d = dp[k];
#pragma unroll
          for (int j = 0; j < DR_K2; ++j) s[k + j] = fmaf(d, w2r[c][j], s[k + j]);
        }
      }
      float* zp = sz + nA * DR_ZP + oA * DR_L1;
      float* ap = sa + nA * DR_ZP + oA * DR_L1;
      float sb = 0.f;
#pragma unroll
      for (int l = 0; l < DR_L1; ++l) {
        const float z = zp[l];
        const float sg = sigmoid_mufu(z);
        const float dz = s[l] * (sg * (1.0f + z * (1.0f - sg)));
        zp[l] = dz;
        ap[l] = z * sg;
        sb += dz;
      }
      accb1 += sb;
    }
    __syncthreads();
    if (doW) {  // ---- C: dw2[c][o][j] += sum_k dd[n][c][k] a[n][o][k + j]  (two nodes per thread)
#pragma unroll 1
      for (int i = 0; i < 2; ++i) {
        const int n = 2 * gW + i;
        const float* ap = sa + n * DR_ZP + oW * DR_L1;
        const float* dp = sdd + (n * C + cW) * DR_DP;
        float a[DR_L1];
#pragma unroll
        for (int l = 0; l < DR_L1; ++l) a[l] = ap[l];
        float sb = 0.f;
#pragma unroll
        for (int k = 0; k < DR_TW; ++k) {
          const float d = dp[k];
          sb += d;
#pragma unroll
          for (int j = 0; j < DR_K2; ++j) acc2[j] = fmaf(d, a[k + j], acc2[j]);
        }
        if (oW == 0) accb2 += sb;
      }
    }
    __syncthreads();                                        // sa is free: it stages dh from here on
    if (doB) {  // ---- B: dh[3m + r] = sum_o sum_t dza[o][m - t] w1[o][c][3t + r]
      float acc[43];
#pragma unroll
      for (int m = 0; m < 43; ++m) acc[m] = 0.f;
#pragma unroll 1
      for (int o = 0; o < DR_OC; ++o) {
        const float* wp = sw1 + (o * C + cB) * DR_K1 + rB;
        float tap[6];
#pragma unroll
        for (int tt = 0; tt < 6; ++tt) tap[tt] = (3 * tt + rB < DR_K1) ? wp[3 * tt] : 0.f;
        const float* zp = sz + nB * DR_ZP + o * DR_L1;
#pragma unroll
        for (int q = 0; q < DR_L1; ++q) {
          const float d = zp[q];
#pragma unroll
          for (int tt = 0; tt < 6; ++tt) acc[q + tt] = fmaf(d, tap[tt], acc[q + tt]);
        }
      }
      float* st = sa + nB * (C * 128 + 1) + cB * 128 + rB;
#pragma unroll
      for (int m = 0; m < 43; ++m)
        if (3 * m + rB < 128) st[3 * m] = acc[m];
    }
    if (doW) {  // ---- D: dw1[o][c][j] += sum_q dza[n][o][q] h[n][c][3q + j]
#pragma unroll 1
      for (int i = 0; i < 2; ++i) {
        const int n = 2 * gW + i;
        const float* zp = sz + n * DR_ZP + oW * DR_L1;
        const float* hp = sh + (n * C + cW) * DR_HP;
        float win[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) win[j] = hp[j];
#pragma unroll
        for (int q = 0; q < DR_L1; ++q) {
          const float d = zp[q];
#pragma unroll
          for (int j = 0; j < DR_K1; ++j) acc1[j] = fmaf(d, win[(3 * q + j) & 15], acc1[j]);
          if (q + 1 < DR_L1) {
#pragma unroll
            for (int j = 0; j < 3; ++j) win[(3 * q + j) & 15] = hp[3 * q + 16 + j];
          }
        }
      }
    }
    __syncthreads();
    for (int i = tid; i < DR_NB * C * 32; i += 256) {       // staged dh -> global, 16 bytes per thread
      const int n = i / (C * 32), q4 = i - n * (C * 32);
      if (n0 + n >= p.N) continue;
      const float* st = sa + n * (C * 128 + 1) + 4 * q4;
      st4(p.dh + (size_t)(n0 + n) * (C * 128) + 4 * q4, make_float4(st[0], st[1], st[2], st[3]));
    }
    __syncthreads();
  }
  // ---- this CTA's weight-gradient partial: the 16 node-pair groups (32 nodes for the first bias) summed in a fixed order
  float* red = sz;                                          // [16][NW] (NW <= 490, 16 * 490 <= NB * ZP)
  float* redb = sa;                                         // [32][8]
  if (doW) {
#pragma unroll
    for (int j = 0; j < DR_K1; ++j) red[gW * NW + (oW * C + cW) * DR_K1 + j] = acc1[j];
#pragma unroll
    for (int j = 0; j < DR_K2; ++j) red[gW * NW + N_W1 + DR_OC + (cW * DR_OC + oW) * DR_K2 + j] = acc2[j];
    if (oW == 0) red[gW * NW + N_W1 + DR_OC + N_W2 + cW] = accb2;
  }
  redb[nA * 8 + oA] = accb1;
  __syncthreads();
  float* part = p.part + (size_t)blockIdx.x * NW;
  for (int e = tid; e < NW; e += 256) {
    float s = 0.f;
    if (e >= N_W1 && e < N_W1 + DR_OC) {
      for (int n = 0; n < DR_NB; ++n) s += redb[n * 8 + (e - N_W1)];
    } else {
      for (int g = 0; g < 16; ++g) s += red[g * NW + e];
    }
    part[e] = s;
  }
}

static bool dr_enabled() {
  static const bool on = [] { const char* e = getenv("MSMP_DECODER_RT"); return !(e && atoi(e) == 0); }();
  return on;
}
static int dr_grid(int ntiles) {
  static const int sms = [] { int d = 0, n = 148; cudaGetDevice(&d); cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, d); return n > 0 ? n : 148; }();
  return ntiles < sms ? ntiles : sms;
}

// 0: launched; 1: not taken (the caller uses the generic kernels)
int launch_decoder_fwd_rt(const float* h, const float* w1, const float* b1, const float* w2, const float* b2, const float* u,
                          int ldu, const float* dt, float* za, float* out, int N, int C, cudaStream_t stream) {
  if (!dr_enabled() || (C != 1 && C != 2)) return 1;
  const int ntiles = (N + DR_NB - 1) / DR_NB;
  DecRtFwd p{h, w1, b1, w2, b2, u, ldu, dt, za, out, N};
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(k_decoder_fwd_rt<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, dr_fwd_smem<1>()) != cudaSuccess ||
        cudaFuncSetAttribute(k_decoder_fwd_rt<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, dr_fwd_smem<2>()) != cudaSuccess)
      return MSMP_ERR_CUDA;
    attr_set = true;
  }
  if (C == 1)
    k_decoder_fwd_rt<1><<<dr_grid(ntiles), 256, dr_fwd_smem<1>(), stream>>>(p, ntiles);
  else
    k_decoder_fwd_rt<2><<<dr_grid(ntiles), 256, dr_fwd_smem<2>(), stream>>>(p, ntiles);
  MSMP_CHECK_LAUNCH();
  return MSMP_OK;
}

// 0: launched, *nparts = number of partials written to `part`; 1: not taken
int launch_decoder_bwd_rt(const float* dout, const float* h, const float* za, const float* w1, const float* w2,
                          const float* dt, float* dh, float* part, int N, int C, int* nparts, cudaStream_t stream) {
  if (!dr_enabled() || (C != 1 && C != 2)) return 1;
  const int ntiles = (N + DR_NB - 1) / DR_NB;
  DecRtBwd p{dout, h, za, w1, w2, dt, dh, part, N};
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(k_decoder_bwd_rt<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, dr_bwd_smem<1>()) != cudaSuccess ||
        cudaFuncSetAttribute(k_decoder_bwd_rt<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, dr_bwd_smem<2>()) != cudaSuccess)
      return MSMP_ERR_CUDA;
    attr_set = true;
  }
  const int grid = dr_grid(ntiles);
  if (C == 1)
    k_decoder_bwd_rt<1><<<grid, 256, dr_bwd_smem<1>(), stream>>>(p, ntiles);
  else
    k_decoder_bwd_rt<2><<<grid, 256, dr_bwd_smem<2>(), stream>>>(p, ntiles);
  MSMP_CHECK_LAUNCH();
  *nparts = grid;
  return MSMP_OK;
}

}  // namespace msmp
