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
