// Multi-tensor AdamW in ONE launch, with every hyper-parameter read from device memory.
//
// The reference's scripts create `optim.AdamW(model.parameters(), lr=args.lr)` and a MultiStepLR scheduler that only
// rewrites param_group['lr'] (experiments/train.py:410-411,437).  torch's own step() cannot be captured into a CUDA
// graph unless the optimizer was built with capturable=True, and even then a Python-float lr is baked into the
// captured kernels.  This kernel is the optimizer step of the captured training step: it updates the optimizer's own
// state tensors (exp_avg, exp_avg_sq) in place and takes lr, betas, eps, weight decay and the two bias corrections
// from a small device array that the host refreshes before every replay -- so schedulers keep working, and a later
// eager optimizer.step() continues from the same state.
//
// Semantics = torch.optim.AdamW (decoupled weight decay, no amsgrad, no maximize):
//     g   = grad * gscale                       (gscale: device scalar, 1 / (2 sqrt(sum of squared errors)) of the step)
//     p  *= 1 - lr * wd
//     m   = m + (1 - b1) (g - m);   v = b2 v + (1 - b2) g g
//     p  -= (lr / bc1) * m / (sqrt(v) / sqrt(bc2) + eps)
// The scaled gradient is written back, so `param.grad` holds d loss / d param after the step.
#include "common.cuh"
#include "msmp_b200.h"

namespace msmp {

struct AdamJob {
  float* p;
  float* g;
  float* m;
  float* v;
  int n;
  int group;
};

constexpr int ADAM_CHUNK = 1024;       // elements per CTA: 256 threads x 4 (strided, coalesced)
constexpr int ADAM_HYPER = 8;          // floats per param group: lr, 1 - beta1, beta2, eps, wd, bc1, sqrt(bc2), 1 - beta2
                                       // (the complements are formed in double on the host, as torch does: 1 - 0.999f != 0.001f)

__global__ void __launch_bounds__(256) k_adamw(const AdamJob* __restrict__ jobs, const int2* __restrict__ chunks,
                                              const float* __restrict__ hyper, const float* __restrict__ gscale) {
  const int2 c = chunks[blockIdx.x];
  const AdamJob j = jobs[c.x];
  const float* h = hyper + j.group * ADAM_HYPER;
  const float lr = h[0], omb1 = h[1], b2 = h[2], eps = h[3], wd = h[4], bc1 = h[5], sbc2 = h[6], omb2 = h[7];
  const float gs = gscale ? *gscale : 1.0f;
  const float decay = 1.0f - lr * wd, step = lr / bc1;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int i = c.y + threadIdx.x + 256 * q;
    if (i >= j.n) break;
    const float g = j.g[i] * gs;
    float m = j.m[i], v = j.v[i], p = j.p[i];
    p *= decay;
    m = m + omb1 * (g - m);
    v = b2 * v + omb2 * g * g;
    p -= step * m / (sqrtf(v) / sbc2 + eps);
    j.p[i] = p;
    j.m[i] = m;
    j.v[i] = v;
    j.g[i] = g;
  }
}

// dst[i] = scale * src[i] for a few scalars plus the derived step scalars of the captured training step:
//   out[0] = sqrt(sse)  (the loss),  out[1] = 0.5 / sqrt(sse)  (gradient scale)   from sse = hi + lo (two floats)
__global__ void k_loss_scalars(const float* __restrict__ sse_hi_lo, double* __restrict__ loss, float* __restrict__ gscale) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    const double sse = (double)sse_hi_lo[0] + (double)sse_hi_lo[1];
    const double l = sqrt(sse);
    *loss = l;
    *gscale = (float)(0.5 / l);
  }
}


// ---- summed squared error of the training criterion and its gradient -----------------------------------------------
// experiments/train_helper.py:126,138: loss = sqrt(MSELoss(reduction='sum')(pred, y)) with float64 labels.  The captured
// step used to spell (pred.double() - y) ** 2 -> sum -> backward with eight framework kernels between the decoder's forward
// and backward launches (all of them alone on the GPU); here: one partial-sum launch + one single-CTA finish, one gradient
// launch.  Deterministic: fixed 4096-element blocks, per-thread sums in index order, fixed shuffle / shared-memory trees.
constexpr int SSE_BLOCK = 4096;        // elements per CTA: 256 threads x 16 (strided, coalesced)

__device__ __forceinline__ double block_sum_256(double v, double* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int w = 0; w < 8; ++w) t += sh[w];
  }
  return t;          // valid in thread 0
}

__global__ void __launch_bounds__(256) k_sse_partial(const float* __restrict__ pred, const double* __restrict__ y, size_t n,
                                                    double* __restrict__ partials) {
  __shared__ double sh[8];
  const size_t base = (size_t)blockIdx.x * SSE_BLOCK;
  double acc = 0.0;
#pragma unroll
  for (int q = 0; q < SSE_BLOCK / 256; ++q) {
    const size_t i = base + threadIdx.x + 256 * q;
    if (i < n) {
      const double d = (double)pred[i] - y[i];
      acc += d * d;
    }
  }
  const double t = block_sum_256(acc, sh);
  if (threadIdx.x == 0) partials[blockIdx.x] = t;
}

__global__ void __launch_bounds__(256) k_sse_finish(const double* __restrict__ partials, int nblocks, double* __restrict__ sse,
                                                   float* __restrict__ sse_hi_lo) {
  __shared__ double sh[8];
  double acc = 0.0;
  for (int b = threadIdx.x; b < nblocks; b += 256) acc += partials[b];
  const double t = block_sum_256(acc, sh);
  if (threadIdx.x == 0) {
    *sse = t;
    if (sse_hi_lo) {          // the float pair that rides in the gradient bucket (exact split of the double's leading 48 bits)
      const float hi = (float)t;
      sse_hi_lo[0] = hi;
      sse_hi_lo[1] = (float)(t - (double)hi);
    }
  }
}

__global__ void __launch_bounds__(256) k_sse_grad(const float* __restrict__ pred, const double* __restrict__ y,
                                                 const double* __restrict__ g, size_t n, float* __restrict__ dpred) {
  const double two_g = 2.0 * (g ? *g : 1.0);
  const size_t base = (size_t)blockIdx.x * 1024;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const size_t i = base + threadIdx.x + 256 * q;
    if (i < n) dpred[i] = (float)(two_g * ((double)pred[i] - y[i]));
  }
}

}  // namespace msmp

extern "C" int msmp_adamw_job_bytes(void) { return (int)sizeof(msmp::AdamJob); }
extern "C" int msmp_adamw_chunk(void) { return msmp::ADAM_CHUNK; }
extern "C" int msmp_adamw_hyper_floats(void) { return msmp::ADAM_HYPER; }

extern "C" int msmp_adamw_run(const void* jobs_dev, const void* chunks_dev, int nchunks, const float* hyper_dev,
                              const float* gscale_dev, cudaStream_t stream) {
  if (nchunks < 0) return MSMP_ERR_ARG;
  if (nchunks == 0) return MSMP_OK;
  msmp::k_adamw<<<nchunks, 256, 0, stream>>>(reinterpret_cast<const msmp::AdamJob*>(jobs_dev),
                                             reinterpret_cast<const int2*>(chunks_dev), hyper_dev, gscale_dev);
  MSMP_CHECK_LAUNCH();
  return MSMP_OK;
}

extern "C" int msmp_loss_scalars(const float* sse_hi_lo, double* loss, float* gscale, cudaStream_t stream) {
  msmp::k_loss_scalars<<<1, 32, 0, stream>>>(sse_hi_lo, loss, gscale);
  MSMP_CHECK_LAUNCH();
  return MSMP_OK;
}

extern "C" size_t msmp_sse_workspace(size_t n) { return ((n + msmp::SSE_BLOCK - 1) / msmp::SSE_BLOCK + 1) * sizeof(double); }

extern "C" int msmp_sse_fwd(const float* pred, const double* y, size_t n, void* workspace, size_t ws_bytes, double* sse,
                            float* sse_hi_lo, cudaStream_t stream) {
  if (!pred || !y || !sse || !workspace || ws_bytes < msmp_sse_workspace(n)) return MSMP_ERR_ARG;
  const size_t nb = (n + msmp::SSE_BLOCK - 1) / msmp::SSE_BLOCK;
  if (nb > 0x7fffffffu) return MSMP_ERR_ARG;
  double* partials = reinterpret_cast<double*>(workspace);
  if (nb) {
    msmp::k_sse_partial<<<(unsigned)nb, 256, 0, stream>>>(pred, y, n, partials);
    MSMP_CHECK_LAUNCH();
  }
  msmp::k_sse_finish<<<1, 256, 0, stream>>>(partials, (int)nb, sse, sse_hi_lo);
  MSMP_CHECK_LAUNCH();
  return MSMP_OK;
}

extern "C" int msmp_sse_bwd(const float* pred, const double* y, const double* g, size_t n, float* dpred, cudaStream_t stream) {
  if (!pred || !y || !dpred) return MSMP_ERR_ARG;
  if (n == 0) return MSMP_OK;
  const size_t nb = (n + 1023) / 1024;
  if (nb > 0x7fffffffu) return MSMP_ERR_ARG;
  msmp::k_sse_grad<<<(unsigned)nb, 256, 0, stream>>>(pred, y, g, n, dpred);
  MSMP_CHECK_LAUNCH();
  return MSMP_OK;
}
