// Weight packing for the tensor-core kernels: ONE launch converts every parameter of the model into the layouts
// the kernels consume (pre-split tf32 hi | lo, UMMA 128B-swizzled tile images; small plain side / bias arrays).
// Parameters keep the reference's layout and names (state_dict compatibility); they change once per optimizer
// step, so this runs once per step -- it replaces several hundred tiny framework ops per step.
//
// A job describes one tile-image block  dst[chunk][hi|lo][4096]  of the k-major weight  Wt[k][n]:
//     Wt[k][n] = sign * (transpose ? src[n * ld + k] : src[k * ld + n])   for k < kvalid, n < nvalid, else 0
// (kind 0), or a plain row block  dst[q * ldd + n] = sign * src[n * ld + q]  (q < kvalid rows, n < nvalid; kind 1).
#include "umma.cuh"
#include "msmp_b200.h"

namespace msmp {

struct PackJob {
  const float* src;
  float* dst;
  int ld;
  int transpose;
  float sign;
  int kvalid;
  int nvalid;
  int nchunks;     // kind 0: number of 32-row k-chunks;  kind 1: unused
  int kind;
  int ldd;
};

__global__ void __launch_bounds__(256) k_pack(const PackJob* __restrict__ jobs) {
  const PackJob j = jobs[blockIdx.y];
  const int tid = threadIdx.x;
  if (j.kind == 1) {
    if (blockIdx.x != 0) return;
    for (int i = tid; i < j.kvalid * j.nvalid; i += 256) {
      const int q = i / j.nvalid, n = i - q * j.nvalid;
      j.dst[(size_t)q * j.ldd + n] = j.sign * j.src[(size_t)n * j.ld + q];
    }
    return;
  }
  const int chunk = blockIdx.x;
  if (chunk >= j.nchunks) return;
  uint8_t* hi = reinterpret_cast<uint8_t*>(j.dst + (size_t)chunk * 2 * (IMG_BYTES / 4));
  uint8_t* lo = hi + IMG_BYTES;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int idx = tid + 256 * i;
    const int n = idx >> 3, c16 = idx & 7;
    const int k0 = chunk * 32 + 4 * c16;
    float v[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int k = k0 + e;
      float x = 0.f;
      if (n < j.nvalid && k < j.kvalid) x = j.sign * (j.transpose ? j.src[(size_t)n * j.ld + k] : j.src[(size_t)k * j.ld + n]);
      v[e] = x;
    }
    store_split4(hi, lo, img_off(n, c16), make_float4(v[0], v[1], v[2], v[3]));
  }
}

// ---------------------------------------------------------------------------------------------------------------
// The inverse direction, once per backward pass: the weight-gradient kernels leave their results in the kernels'
// k-major layouts inside one raw buffer ([K][N] blocks, side / bias rows); ONE launch scatters them into the
// parameter-layout gradients (param.grad), overwriting them:
//     dst[n * ldd + k] = sum_s src0[s * sstride0 + k * ld0 + n] (+ sign1 * sum_s src1[s * sstride1 + k * ld1 + n])
//                                                                            n < rows, k < cols     (or 0 if `zero`)
// i.e. a transposed copy with an optional second signed source (the +/- rows of the factorised first message layer).  A
// source may be the `nsplit` split-M partials of a weight-gradient launch (wgrad_ws.cu): they are summed here in split
// order (fixed => deterministic), which replaces one reduction launch per weight gradient.
struct UnpackJob {
  float* dst;
  const float* src0;
  const float* src1;
  int ldd;
  int ld0;
  int ld1;
  int rows;
  int cols;
  float sign1;
  int zero;
  int nsplit;          // >= 1
  int sstride0;        // floats between consecutive partials of src0 / src1
  int sstride1;
};

__global__ void __launch_bounds__(256) k_unpack(const UnpackJob* __restrict__ jobs) {
  __shared__ float tile[32][33];
  const UnpackJob j = jobs[blockIdx.y];
  const int tn = (j.rows + 31) >> 5, tk = (j.cols + 31) >> 5;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int ti = blockIdx.x; ti < tn * tk; ti += gridDim.x) {
    const int n0 = (ti % tn) << 5, k0 = (ti / tn) << 5;
    // the four rows (k0 + ty + 8 q) of this thread are walked together: four independent chains of loads per split step
    // (a row at a time left a thread with one chain of nsplit dependent rounds: 67 us alone at the end of the C2 step);
    // every element is still summed in split order, so the result does not depend on the grouping
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    const int n = n0 + tx;
    if (!j.zero && n < j.rows) {
      const float* s0[4];
      bool ok[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int k = k0 + ty + 8 * q;
        ok[q] = k < j.cols;
        s0[q] = j.src0 + (size_t)(ok[q] ? k : k0) * j.ld0 + n;
      }
      int s = 0;
      for (; s + 2 <= j.nsplit; s += 2) {
        float a[4], b[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          a[q] = ok[q] ? s0[q][(size_t)s * j.sstride0] : 0.f;
          b[q] = ok[q] ? s0[q][(size_t)(s + 1) * j.sstride0] : 0.f;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          v[q] += a[q];
          v[q] += b[q];
        }
      }
      if (s < j.nsplit) {
#pragma unroll
        for (int q = 0; q < 4; ++q) v[q] += ok[q] ? s0[q][(size_t)s * j.sstride0] : 0.f;
      }
      if (j.src1) {
        float w[4] = {0.f, 0.f, 0.f, 0.f};
        for (int t = 0; t < j.nsplit; ++t) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int k = k0 + ty + 8 * q;
            w[q] += ok[q] ? j.src1[(size_t)k * j.ld1 + n + (size_t)t * j.sstride1] : 0.f;
          }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) v[q] += j.sign1 * w[q];
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) tile[ty + 8 * q][tx] = v[q];
    __syncthreads();
#pragma unroll
    for (int y = ty; y < 32; y += 8) {
      const int n = n0 + y, k = k0 + tx;
      if (n < j.rows && k < j.cols) j.dst[(size_t)n * j.ldd + k] = tile[tx][y];
    }
    __syncthreads();
  }
}

}  // namespace msmp

extern "C" int msmp_unpack_job_bytes(void) { return (int)sizeof(msmp::UnpackJob); }

// jobs_dev: device array of `njobs` UnpackJob records (layout above; see msmp_pde_b200/gradsink.py)
extern "C" int msmp_unpack_run(const void* jobs_dev, int njobs, int max_tiles, cudaStream_t stream) {
  if (njobs < 0 || max_tiles < 1) return MSMP_ERR_ARG;
  if (njobs == 0) return MSMP_OK;
  dim3 grid(max_tiles < 32 ? max_tiles : 32, njobs);
  msmp::k_unpack<<<grid, 256, 0, stream>>>(reinterpret_cast<const msmp::UnpackJob*>(jobs_dev));
  MSMP_CHECK_LAUNCH();
  return MSMP_OK;
}

extern "C" int msmp_pack_job_bytes(void) { return (int)sizeof(msmp::PackJob); }

// jobs_dev: device array of `njobs` PackJob records (layout above; see msmp_pde_b200/packing.py)
extern "C" int msmp_pack_run(const void* jobs_dev, int njobs, int max_chunks, cudaStream_t stream) {
  if (njobs < 0 || max_chunks < 1) return MSMP_ERR_ARG;
  if (njobs == 0) return MSMP_OK;
  dim3 grid(max_chunks, njobs);
  msmp::k_pack<<<grid, 256, 0, stream>>>(reinterpret_cast<const msmp::PackJob*>(jobs_dev));
  MSMP_CHECK_LAUNCH();
  return MSMP_OK;
}
