// InstanceNorm (PyG semantics, affine=False; models_gnn.py:59,66,122,129) and the MSMP-PDE gate blend
// (models_gnn.py:1365-1368 / models_gnn2D.py:438-441), forward and backward, fp32 (sm_100a).
//
// Per graph g and channel c:  mu = mean_n y,  var = mean_n (y - mu)^2 (biased),  o = (y - mu) * rsqrt(var + eps).
// Statistics are computed per graph-aligned node chunk with a two-pass (centred) sum, chunks are merged with
// Chan's formula in a fixed order (bit-stable; no E[x^2]-mu^2 cancellation; no atomics).
//   mode 0 (plain):  out = o                                           (MP_PDE_Solver stack)
//   mode 1 (gated):  out = (1 - tau) * h + tau * sw(o_main), tau = sigmoid(o_gate)
#include "common.cuh"
#include "msmp_b200.h"

namespace msmp {

// stats layout per tensor: stat[g][0][c] = mu, stat[g][1][c] = rstd
__global__ void __launch_bounds__(256) k_in_chunk_stats(const float* __restrict__ y0, const float* __restrict__ y1,
                                                        int ld, const int* __restrict__ chunk_begin,
                                                        const int* __restrict__ chunk_end, float* __restrict__ cstat,
                                                        int nchunks) {
  __shared__ float red[2][128];
  const float* y = blockIdx.y == 0 ? y0 : y1;
  const int chunk = blockIdx.x;
  const int c = threadIdx.x & 127, g = threadIdx.x >> 7;
  const int b = chunk_begin[chunk], e = chunk_end[chunk];
  const float n = (float)(e - b);
  float s = 0.f;
  for (int r = b + g; r < e; r += 2) s += __ldg(y + (size_t)r * ld + c);
  red[g][c] = s;
  __syncthreads();
  const float mean = (red[0][c] + red[1][c]) / fmaxf(n, 1.f);
  __syncthreads();
  float q = 0.f;
  for (int r = b + g; r < e; r += 2) {
    float d = __ldg(y + (size_t)r * ld + c) - mean;
    q = fmaf(d, d, q);
  }
  red[g][c] = q;
  __syncthreads();
  if (g == 0) {
    float* o = cstat + ((size_t)blockIdx.y * nchunks + chunk) * 256;
    o[c] = mean;
    o[128 + c] = red[0][c] + red[1][c];
  }
}

// Chan merge of a graph's chunk statistics in a fixed order: 8 slices of 128 threads merge every 8th chunk each (a lattice
// graph of the C4 batch has 128 chunks: one serial chain of dependent L2 loads took 71 us per launch), then slice 0 merges
// the 8 partial results in slice order.  With one chunk per graph the result is that chunk's statistics, bit for bit.
constexpr int IN_FIN_SLICES = 8;
__global__ void __launch_bounds__(128 * IN_FIN_SLICES) k_in_finalize(const float* __restrict__ cstat,
                                                                     const int* __restrict__ chunk_begin,
                                                                     const int* __restrict__ chunk_end,
                                                                     const int* __restrict__ graph_chunk_ptr,
                                                                     float* __restrict__ stat, int nchunks, int B, float eps) {
  __shared__ float sh[IN_FIN_SLICES][3][128];
  const int g = blockIdx.x, c = threadIdx.x & 127, sl = threadIdx.x >> 7;
  const float* cs = cstat + (size_t)blockIdx.y * nchunks * 256;
  float n = 0.f, mean = 0.f, m2 = 0.f;
  for (int k = graph_chunk_ptr[g] + sl; k < graph_chunk_ptr[g + 1]; k += IN_FIN_SLICES) {
    const float nb = (float)(chunk_end[k] - chunk_begin[k]);
    if (nb <= 0.f) continue;
    const float mb = cs[(size_t)k * 256 + c], qb = cs[(size_t)k * 256 + 128 + c];
    const float nt = n + nb;
    const float d = mb - mean;
    mean += d * (nb / nt);
    m2 += qb + d * d * (n * nb / nt);
    n = nt;
  }
  sh[sl][0][c] = n;
  sh[sl][1][c] = mean;
  sh[sl][2][c] = m2;
  __syncthreads();
  if (sl != 0) return;
  for (int s2 = 1; s2 < IN_FIN_SLICES; ++s2) {
    const float nb = sh[s2][0][c];
    if (nb <= 0.f) continue;
    const float mb = sh[s2][1][c], qb = sh[s2][2][c];
    const float nt = n + nb;
    const float d = mb - mean;
    mean += d * (nb / nt);
    m2 += qb + d * d * (n * nb / nt);
    n = nt;
  }
  float* o = stat + ((size_t)blockIdx.y * B + g) * 256;
  o[c] = mean;
  o[128 + c] = rsqrtf(m2 / fmaxf(n, 1.f) + eps);
}

// d(out)/d(o_gate), d(out)/d(o_main) for one element of the gated blend
// (activations on the MUFU pipe, common.cuh: these element-wise kernels were bound by the three expf / divisions per element)
__device__ __forceinline__ void blend_grads(float dout, float og, float om, float h, float& dog, float& dom, float& dh) {
  const float t = sigmoid_mufu(og);
  const float sm = sigmoid_mufu(om);
  dog = dout * (om * sm - h) * t * (1.f - t);
  dom = dout * t * (sm * (1.0f + om * (1.0f - sm)));
  dh = dout * (1.f - t);
}
// out = (1 - sigmoid(o_gate)) h + sigmoid(o_gate) swish(o_main)      (models_gnn.py:1365-1368)
__device__ __forceinline__ float blend_fwd(float og, float om, float h) {
  const float t = sigmoid_mufu(og);
  return (1.f - t) * h + t * swish_m(om);
}

// forward apply. y1/stat1/h only used in gated mode.
__global__ void __launch_bounds__(256) k_in_apply(const float* __restrict__ y0, const float* __restrict__ y1, int ld,
                                                  const float* __restrict__ stat, const int* __restrict__ node_graph,
                                                  const float* __restrict__ h, float* __restrict__ out, int N, int B,
                                                  int mode) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N * 32) return;
  const int row = idx >> 5, c4 = (idx & 31) * 4;
  const int g = __ldg(node_graph + row);
  const float* st0 = stat + (size_t)g * 256;
  float4 mu = ldg4(st0 + c4), rs = ldg4(st0 + 128 + c4);
  float4 y = ldg4(y0 + (size_t)row * ld + c4);
  float4 o = make_float4((y.x - mu.x) * rs.x, (y.y - mu.y) * rs.y, (y.z - mu.z) * rs.z, (y.w - mu.w) * rs.w);
  if (mode == 0) {
    st4(out + (size_t)row * 128 + c4, o);
    return;
  }
  const float* st1 = stat + ((size_t)B + g) * 256;
  float4 mu1 = ldg4(st1 + c4), rs1 = ldg4(st1 + 128 + c4);
  float4 ym = ldg4(y1 + (size_t)row * ld + c4);
  float4 om = make_float4((ym.x - mu1.x) * rs1.x, (ym.y - mu1.y) * rs1.y, (ym.z - mu1.z) * rs1.z, (ym.w - mu1.w) * rs1.w);
  float4 hh = ldg4(h + (size_t)row * 128 + c4);
  float4 r;
  {
    float t;
    r.x = blend_fwd(o.x, om.x, hh.x);
    r.y = blend_fwd(o.y, om.y, hh.y);
    r.z = blend_fwd(o.z, om.z, hh.z);
    r.w = blend_fwd(o.w, om.w, hh.w);
  }
  st4(out + (size_t)row * 128 + c4, r);
}

// ---- backward -------------------------------------------------------------------------------------
// per chunk: part[chunk][q][c], q = 0: sum do0, 1: sum do0*o0, 2: sum do1, 3: sum do1*o1
__global__ void __launch_bounds__(1024) k_in_bwd_chunk(const float* __restrict__ dout, const float* __restrict__ y0,
                                                       const float* __restrict__ y1, int ld,
                                                       const float* __restrict__ stat, const float* __restrict__ h,
                                                       const int* __restrict__ chunk_begin,
                                                       const int* __restrict__ chunk_end,
                                                       const int* __restrict__ node_graph, float* __restrict__ part,
                                                       int B, int mode) {
  // 8 row groups x 128 channels; the groups are combined in a fixed order (bit-stable)
  __shared__ float red[8][4][128];
  const int chunk = blockIdx.x;
  const int c = threadIdx.x & 127, grp = threadIdx.x >> 7;
  const int b = chunk_begin[chunk], e = chunk_end[chunk];
  float s[4] = {0.f, 0.f, 0.f, 0.f};
  if (e > b) {
    const int g = node_graph[b];
    const float mu0 = stat[(size_t)g * 256 + c], rs0 = stat[(size_t)g * 256 + 128 + c];
    float mu1 = 0.f, rs1 = 0.f;
    if (mode == 1) {
      mu1 = stat[((size_t)B + g) * 256 + c];
      rs1 = stat[((size_t)B + g) * 256 + 128 + c];
    }
    for (int r = b + grp; r < e; r += 8) {
      const float d = __ldg(dout + (size_t)r * 128 + c);
      const float o0 = (__ldg(y0 + (size_t)r * ld + c) - mu0) * rs0;
      if (mode == 0) {
        s[0] += d;
        s[1] = fmaf(d, o0, s[1]);
      } else {
        const float o1 = (__ldg(y1 + (size_t)r * ld + c) - mu1) * rs1;
        float dog, dom, dh;
        blend_grads(d, o0, o1, __ldg(h + (size_t)r * 128 + c), dog, dom, dh);
        s[0] += dog;
        s[1] = fmaf(dog, o0, s[1]);
        s[2] += dom;
        s[3] = fmaf(dom, o1, s[3]);
      }
    }
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) red[grp][q][c] = s[q];
  __syncthreads();
  if (threadIdx.x < 512) {
    const int q = threadIdx.x >> 7;
    float t = 0.f;
#pragma unroll
    for (int gI = 0; gI < 8; ++gI) t += red[gI][q][c];
    part[((size_t)chunk * 4 + q) * 128 + c] = t;
  }
}

// gm[g][q][c] = (sum over the graph's chunks, fixed order) / n_g.  Eight slices of 128 threads sum every 8th chunk each (all
// four sums of a channel in one thread: four independent chains of L2 loads), then the eight partial results are added in slice
// order -- one serial chain over the 128 chunks of a C4 lattice graph took 24 us per launch.  With one chunk per graph the result
// is that chunk's sum, bit for bit (the fused single-chunk path below relies on it).
__global__ void __launch_bounds__(128 * IN_FIN_SLICES) k_in_bwd_finalize(const float* __restrict__ part,
                                                                         const int* __restrict__ chunk_begin,
                                                                         const int* __restrict__ chunk_end,
                                                                         const int* __restrict__ graph_chunk_ptr,
                                                                         float* __restrict__ gm) {
  __shared__ float sh[IN_FIN_SLICES][5][128];
  const int g = blockIdx.x, sl = threadIdx.x >> 7, c = threadIdx.x & 127;
  float s[4] = {0.f, 0.f, 0.f, 0.f}, n = 0.f;
  for (int k = graph_chunk_ptr[g] + sl; k < graph_chunk_ptr[g + 1]; k += IN_FIN_SLICES) {
#pragma unroll
    for (int q = 0; q < 4; ++q) s[q] += part[((size_t)k * 4 + q) * 128 + c];
    n += (float)(chunk_end[k] - chunk_begin[k]);
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) sh[sl][q][c] = s[q];
  sh[sl][4][c] = n;
  __syncthreads();
  if (sl >= 4) return;
  const int q = sl;
  float t = sh[0][q][c], nt = sh[0][4][c];
  for (int s2 = 1; s2 < IN_FIN_SLICES; ++s2) {
    t += sh[s2][q][c];
    nt += sh[s2][4][c];
  }
  gm[((size_t)g * 4 + q) * 128 + c] = t / fmaxf(nt, 1.f);
}

// dy = rstd * (do - mean(do) - o * mean(do * o)); four channels per thread (16-byte accesses; same arithmetic per element)
__global__ void __launch_bounds__(256) k_in_bwd_apply(const float* __restrict__ dout, const float* __restrict__ y0,
                                                      const float* __restrict__ y1, int ld,
                                                      const float* __restrict__ stat, const float* __restrict__ h,
                                                      const float* __restrict__ gm,
                                                      const int* __restrict__ node_graph, float* __restrict__ dy0,
                                                      float* __restrict__ dy1, int lddy, float* __restrict__ dh, int N,
                                                      int B, int mode) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N * 32) return;
  const int row = idx >> 5, c = (idx & 31) * 4;
  const int g = __ldg(node_graph + row);
  const float4 mu0 = ldg4(stat + (size_t)g * 256 + c), rs0 = ldg4(stat + (size_t)g * 256 + 128 + c);
  const float4 d = ldg4(dout + (size_t)row * 128 + c);
  const float4 v0 = ldg4(y0 + (size_t)row * ld + c);
  const float4 o0 = make_float4((v0.x - mu0.x) * rs0.x, (v0.y - mu0.y) * rs0.y, (v0.z - mu0.z) * rs0.z, (v0.w - mu0.w) * rs0.w);
  const float* m = gm + (size_t)g * 4 * 128;
  const float4 m0 = ldg4(m + c), m1 = ldg4(m + 128 + c);
  if (mode == 0) {
    st4(dy0 + (size_t)row * lddy + c, make_float4(rs0.x * (d.x - m0.x - o0.x * m1.x), rs0.y * (d.y - m0.y - o0.y * m1.y),
                                                  rs0.z * (d.z - m0.z - o0.z * m1.z), rs0.w * (d.w - m0.w - o0.w * m1.w)));
    return;
  }
  const float4 mu1 = ldg4(stat + ((size_t)B + g) * 256 + c), rs1 = ldg4(stat + ((size_t)B + g) * 256 + 128 + c);
  const float4 v1 = ldg4(y1 + (size_t)row * ld + c);
  const float4 o1 = make_float4((v1.x - mu1.x) * rs1.x, (v1.y - mu1.y) * rs1.y, (v1.z - mu1.z) * rs1.z, (v1.w - mu1.w) * rs1.w);
  const float4 hh = ldg4(h + (size_t)row * 128 + c);
  const float4 m2 = ldg4(m + 256 + c), m3 = ldg4(m + 384 + c);
  float4 dog, dom, dhv;
  blend_grads(d.x, o0.x, o1.x, hh.x, dog.x, dom.x, dhv.x);
  blend_grads(d.y, o0.y, o1.y, hh.y, dog.y, dom.y, dhv.y);
  blend_grads(d.z, o0.z, o1.z, hh.z, dog.z, dom.z, dhv.z);
  blend_grads(d.w, o0.w, o1.w, hh.w, dog.w, dom.w, dhv.w);
  st4(dy0 + (size_t)row * lddy + c, make_float4(rs0.x * (dog.x - m0.x - o0.x * m1.x), rs0.y * (dog.y - m0.y - o0.y * m1.y),
                                                rs0.z * (dog.z - m0.z - o0.z * m1.z), rs0.w * (dog.w - m0.w - o0.w * m1.w)));
  st4(dy1 + (size_t)row * lddy + c, make_float4(rs1.x * (dom.x - m2.x - o1.x * m3.x), rs1.y * (dom.y - m2.y - o1.y * m3.y),
                                                rs1.z * (dom.z - m2.z - o1.z * m3.z), rs1.w * (dom.w - m2.w - o1.w * m3.w)));
  st4(dh + (size_t)row * 128 + c, dhv);
}

// ---- one launch per direction when every graph is exactly one chunk (graphs of <= 128 nodes: the reference's 100-node
// grids).  Chunk g is graph g; the arithmetic and its order are those of the three-kernel path above (two strided row
// groups for the statistics, eight for the backward sums, Chan merge of a single chunk = identity), so both paths give
// the same bits; what goes away is two launches and two trips through global memory per call.
__global__ void __launch_bounds__(1024) k_in_fused_fwd(const float* __restrict__ y0, const float* __restrict__ y1, int ld,
                                                       const int* __restrict__ chunk_begin,
                                                       const int* __restrict__ chunk_end, const float* __restrict__ h,
                                                       float* __restrict__ stat, float* __restrict__ out, int B, int mode,
                                                       float eps) {
  // The graph's rows of both tensors are staged in shared memory once (all 1024 threads, float4), the statistics
  // (512 threads: tensor x row group x channel, the chunked path's two row groups and order) and the apply pass read
  // them from there.
  extern __shared__ __align__(16) float tile[];      // [mode + 1][128 rows][128]
  __shared__ float red[2][2][128];
  __shared__ __align__(16) float st[2][2][128];      // [tensor][mu | rstd][c]
  const int gI = blockIdx.x, tid = threadIdx.x;
  const int b = chunk_begin[gI], e = chunk_end[gI];
  const int nrows = e - b;
  const float n = (float)nrows;
  for (int t = 0; t <= mode; ++t) {
    const float* y = t == 0 ? y0 : y1;
    for (int i = tid; i < nrows * 32; i += 1024)
      st4(tile + (t * 128 + (i >> 5)) * 128 + 4 * (i & 31), ldg4(y + (size_t)(b + (i >> 5)) * ld + 4 * (i & 31)));
  }
  __syncthreads();
  const int t = tid >> 8, g = (tid >> 7) & 1, c = tid & 127;
  const bool worker = tid < 512 && t <= mode;
  const float* col = tile + t * 128 * 128 + c;
  float mean = 0.f;
  if (worker) {
    float s = 0.f;
    for (int r = g; r < nrows; r += 2) s += col[r * 128];
    red[t][g][c] = s;
  }
  __syncthreads();
  if (worker) mean = (red[t][0][c] + red[t][1][c]) / fmaxf(n, 1.f);
  __syncthreads();
  if (worker) {
    float q = 0.f;
    for (int r = g; r < nrows; r += 2) {
      const float d = col[r * 128] - mean;
      q = fmaf(d, d, q);
    }
    red[t][g][c] = q;
  }
  __syncthreads();
  if (worker && g == 0) {
    // k_in_finalize with one chunk: mean = 0 + (mb - 0) * (nb / nb) = mb;  m2 = 0 + qb + d * d * 0 = qb
    const float rs = rsqrtf((red[t][0][c] + red[t][1][c]) / fmaxf(n, 1.f) + eps);
    st[t][0][c] = mean;
    st[t][1][c] = rs;
    float* o = stat + ((size_t)t * B + gI) * 256;
    o[c] = mean;
    o[128 + c] = rs;
  }
  __syncthreads();
  for (int idx = tid; idx < nrows * 32; idx += 1024) {
    const int rl = idx >> 5, row = b + rl, c4 = (idx & 31) * 4;
    const float4 mu = *reinterpret_cast<const float4*>(&st[0][0][c4]), rs = *reinterpret_cast<const float4*>(&st[0][1][c4]);
    const float4 y = *reinterpret_cast<const float4*>(tile + rl * 128 + c4);
    const float4 o = make_float4((y.x - mu.x) * rs.x, (y.y - mu.y) * rs.y, (y.z - mu.z) * rs.z, (y.w - mu.w) * rs.w);
    if (mode == 0) {
      st4(out + (size_t)row * 128 + c4, o);
      continue;
    }
    const float4 mu1 = *reinterpret_cast<const float4*>(&st[1][0][c4]), rs1 = *reinterpret_cast<const float4*>(&st[1][1][c4]);
    const float4 ym = *reinterpret_cast<const float4*>(tile + (128 + rl) * 128 + c4);
    const float4 om = make_float4((ym.x - mu1.x) * rs1.x, (ym.y - mu1.y) * rs1.y, (ym.z - mu1.z) * rs1.z, (ym.w - mu1.w) * rs1.w);
    const float4 hh = ldg4(h + (size_t)row * 128 + c4);
    float4 r;
    float tt;
    r.x = blend_fwd(o.x, om.x, hh.x);
    r.y = blend_fwd(o.y, om.y, hh.y);
    r.z = blend_fwd(o.z, om.z, hh.z);
    r.w = blend_fwd(o.w, om.w, hh.w);
    st4(out + (size_t)row * 128 + c4, r);
  }
}

__global__ void __launch_bounds__(1024) k_in_fused_bwd(const float* __restrict__ dout, const float* __restrict__ y0,
                                                       const float* __restrict__ y1, int ld,
                                                       const float* __restrict__ stat, const float* __restrict__ h,
                                                       const int* __restrict__ chunk_begin,
                                                       const int* __restrict__ chunk_end, float* __restrict__ dy0,
                                                       float* __restrict__ dy1, int lddy, float* __restrict__ dh, int B,
                                                       int mode) {
  __shared__ float red[8][4][128];
  __shared__ float gm[4][128];
  const int gI = blockIdx.x;
  const int c = threadIdx.x & 127, grp = threadIdx.x >> 7;
  const int b = chunk_begin[gI], e = chunk_end[gI];
  const float mu0 = stat[(size_t)gI * 256 + c], rs0 = stat[(size_t)gI * 256 + 128 + c];
  float mu1 = 0.f, rs1 = 0.f;
  if (mode == 1) {
    mu1 = stat[((size_t)B + gI) * 256 + c];
    rs1 = stat[((size_t)B + gI) * 256 + 128 + c];
  }
  float s[4] = {0.f, 0.f, 0.f, 0.f};
  for (int r = b + grp; r < e; r += 8) {
    const float d = __ldg(dout + (size_t)r * 128 + c);
    const float o0 = (__ldg(y0 + (size_t)r * ld + c) - mu0) * rs0;
    if (mode == 0) {
      s[0] += d;
      s[1] = fmaf(d, o0, s[1]);
    } else {
      const float o1 = (__ldg(y1 + (size_t)r * ld + c) - mu1) * rs1;
      float dog, dom, dhv;
      blend_grads(d, o0, o1, __ldg(h + (size_t)r * 128 + c), dog, dom, dhv);
      s[0] += dog;
      s[1] = fmaf(dog, o0, s[1]);
      s[2] += dom;
      s[3] = fmaf(dom, o1, s[3]);
    }
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) red[grp][q][c] = s[q];
  __syncthreads();
  if (threadIdx.x < 512) {
    const int q = threadIdx.x >> 7;
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += red[k][q][c];
    gm[q][c] = t / fmaxf((float)(e - b), 1.f);          // k_in_bwd_finalize with one chunk
  }
  __syncthreads();
  for (int r = b + grp; r < e; r += 8) {
    const size_t idx = (size_t)r * 128 + c;
    const float d = dout[idx];
    const float o0 = (y0[(size_t)r * ld + c] - mu0) * rs0;
    if (mode == 0) {
      dy0[(size_t)r * lddy + c] = rs0 * (d - gm[0][c] - o0 * gm[1][c]);
      continue;
    }
    const float o1 = (y1[(size_t)r * ld + c] - mu1) * rs1;
    float dog, dom, dhv;
    blend_grads(d, o0, o1, h[idx], dog, dom, dhv);
    dy0[(size_t)r * lddy + c] = rs0 * (dog - gm[0][c] - o0 * gm[1][c]);
    dy1[(size_t)r * lddy + c] = rs1 * (dom - gm[2][c] - o1 * gm[3][c]);
    dh[idx] = dhv;
  }
}

}  // namespace msmp

using namespace msmp;

extern "C" size_t msmp_instnorm_workspace(int nchunks, int B) {
  // chunk stats [2][nchunks][256] (fwd) or chunk partials [nchunks][4][128] (bwd) + per-graph means [B][4][128]
  return ((size_t)nchunks * 512 + (size_t)B * 512) * sizeof(float);
}

extern "C" int msmp_instnorm_fwd(const float* y0, const float* y1, int ld, const float* h, const int* chunk_begin,
                                 const int* chunk_end, const int* graph_chunk_ptr, const int* node_graph, int nchunks,
                                 int B, int N, int mode, float eps, float* stat, float* out, void* workspace,
                                 size_t ws_bytes, cudaStream_t stream) {
  if (N < 0 || B < 0 || (mode != 0 && mode != 1) || (ld & 3)) return MSMP_ERR_ARG;
  if (N == 0) return MSMP_OK;
  if (ws_bytes < msmp_instnorm_workspace(nchunks, B)) return MSMP_ERR_WORKSPACE;
  float* cstat = reinterpret_cast<float*>(workspace);
  const int nt = mode + 1;
  k_in_chunk_stats<<<dim3(nchunks, nt), 256, 0, stream>>>(y0, y1, ld, chunk_begin, chunk_end, cstat, nchunks);
  MSMP_CHECK_LAUNCH();
  k_in_finalize<<<dim3(B, nt), 128 * IN_FIN_SLICES, 0, stream>>>(cstat, chunk_begin, chunk_end, graph_chunk_ptr, stat, nchunks, B, eps);
  MSMP_CHECK_LAUNCH();
  k_in_apply<<<(N * 32 + 255) / 256, 256, 0, stream>>>(y0, y1, ld, stat, node_graph, h, out, N, B, mode);
  MSMP_CHECK_LAUNCH();
  return MSMP_OK;
}

extern "C" int msmp_instnorm_bwd(const float* dout, const float* y0, const float* y1, int ld, const float* h,
                                 const float* stat, const int* chunk_begin, const int* chunk_end,
                                 const int* graph_chunk_ptr, const int* node_graph, int nchunks, int B, int N, int mode,
                                 float* dy0, float* dy1, int lddy, float* dh, void* workspace, size_t ws_bytes,
                                 cudaStream_t stream) {
  if (N < 0 || B < 0 || (mode != 0 && mode != 1) || (ld & 3) || (lddy & 3)) return MSMP_ERR_ARG;
  if (N == 0) return MSMP_OK;
  if (ws_bytes < msmp_instnorm_workspace(nchunks, B)) return MSMP_ERR_WORKSPACE;
  float* part = reinterpret_cast<float*>(workspace);
  float* gm = part + (size_t)nchunks * 512;
  k_in_bwd_chunk<<<nchunks, 1024, 0, stream>>>(dout, y0, y1, ld, stat, h, chunk_begin, chunk_end, node_graph, part, B, mode);
  MSMP_CHECK_LAUNCH();
  k_in_bwd_finalize<<<B, 128 * IN_FIN_SLICES, 0, stream>>>(part, chunk_begin, chunk_end, graph_chunk_ptr, gm);
  MSMP_CHECK_LAUNCH();
  k_in_bwd_apply<<<(N * 32 + 255) / 256, 256, 0, stream>>>(dout, y0, y1, ld, stat, h, gm, node_graph, dy0, dy1, lddy, dh,
                                                            N, B, mode);
  MSMP_CHECK_LAUNCH();
  return MSMP_OK;
}

/* One chunk per graph (chunk g = graph g; the caller guarantees it: every graph has 1..128 nodes): one launch each. */
extern "C" int msmp_instnorm1_fwd(const float* y0, const float* y1, int ld, const float* h, const int* chunk_begin,
                                  const int* chunk_end, int B, int N, int mode, float eps, float* stat, float* out,
                                  cudaStream_t stream) {
  if (N < 0 || B < 0 || (mode != 0 && mode != 1) || (ld & 3)) return MSMP_ERR_ARG;
  if (N == 0 || B == 0) return MSMP_OK;
  const int smem = (mode + 1) * 128 * 128 * (int)sizeof(float);
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(k_in_fused_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 128 * 128 * 4) != cudaSuccess)
      return MSMP_ERR_CUDA;
    attr_set = true;
  }
  k_in_fused_fwd<<<B, 1024, smem, stream>>>(y0, y1, ld, chunk_begin, chunk_end, h, stat, out, B, mode, eps);
  MSMP_CHECK_LAUNCH();
  return MSMP_OK;
}

extern "C" int msmp_instnorm1_bwd(const float* dout, const float* y0, const float* y1, int ld, const float* h,
                                  const float* stat, const int* chunk_begin, const int* chunk_end, int B, int N,
                                  int mode, float* dy0, float* dy1, int lddy, float* dh, cudaStream_t stream) {
  if (N < 0 || B < 0 || (mode != 0 && mode != 1)) return MSMP_ERR_ARG;
  if (N == 0 || B == 0) return MSMP_OK;
  k_in_fused_bwd<<<B, 1024, 0, stream>>>(dout, y0, y1, ld, stat, h, chunk_begin, chunk_end, dy0, dy1, lddy, dh, B, mode);
  MSMP_CHECK_LAUNCH();
  return MSMP_OK;
}
