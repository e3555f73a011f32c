// Step bookkeeping kernels: BatchNorm running-statistics update (one launch for every BatchNorm of the
// network) and the fused RMSprop step on the flat parameter / gradient buffers.
#include "common.cuh"
#include "../../include/cunet_b200.h"
#include "host_util.h"

namespace cunet {

// nn.BatchNorm2d train mode: running = (1-m)*running + m*batch, with the UNBIASED batch variance.
// reps == 2 reproduces the reference's double update of checkpointed BatchNorms.
__global__ void bn_running_update_kernel(const cunet_bn_update_desc* __restrict__ descs) {
  const cunet_bn_update_desc d = descs[blockIdx.x];
  int cin = 0;
  for (int s = 0; s < d.nseg; ++s) cin += d.C[s];
  for (int k = threadIdx.x; k < cin; k += blockDim.x) {
    int s = 0, base = 0;
    while (s + 1 < d.nseg && k >= base + d.C[s]) {
      base += d.C[s];
      ++s;
    }
    const int c = k - base;
    const double mean = d.stats[s][c] * d.inv_count[s];
    double var = d.stats[s][d.C[s] + c] * d.inv_count[s] - mean * mean;
    if (var < 0.0) var = 0.0;
    const double unbiased = d.n_elems > 1.0 ? var * d.n_elems / (d.n_elems - 1.0) : var;
    float rm = d.rmean[k], rv = d.rvar[k];
    for (int r = 0; r < d.reps; ++r) {
      rm = (1.f - d.momentum) * rm + d.momentum * (float)mean;
      rv = (1.f - d.momentum) * rv + d.momentum * (float)unbiased;
    }
    d.rmean[k] = rm;
    d.rvar[k] = rv;
  }
}

// torch.optim.RMSprop (momentum 0, not centered):  v = alpha*v + (1-alpha)*g*g ; p -= lr * g / (sqrt(v) + eps)
__global__ void rmsprop_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ v, long n,
                               const float* __restrict__ lr_dev, float alpha, float eps) {
  const float lr = *lr_dev;
  const long n4 = n >> 2;
  const long stride = (long)gridDim.x * blockDim.x;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 pv = reinterpret_cast<float4*>(p)[i];
    const float4 gv = reinterpret_cast<const float4*>(g)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    vv.x = alpha * vv.x + (1.f - alpha) * gv.x * gv.x;
    vv.y = alpha * vv.y + (1.f - alpha) * gv.y * gv.y;
    vv.z = alpha * vv.z + (1.f - alpha) * gv.z * gv.z;
    vv.w = alpha * vv.w + (1.f - alpha) * gv.w * gv.w;
    pv.x -= lr * gv.x / (sqrtf(vv.x) + eps);
    pv.y -= lr * gv.y / (sqrtf(vv.y) + eps);
    pv.z -= lr * gv.z / (sqrtf(vv.z) + eps);
    pv.w -= lr * gv.w / (sqrtf(vv.w) + eps);
    reinterpret_cast<float4*>(p)[i] = pv;
    reinterpret_cast<float4*>(v)[i] = vv;
  }
  for (long i = (n4 << 2) + (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float gi = g[i];
    const float vi = alpha * v[i] + (1.f - alpha) * gi * gi;
    v[i] = vi;
    p[i] -= lr * gi / (sqrtf(vi) + eps);
  }
}

}  // namespace cunet
using namespace cunet;

extern "C" int cunet_bn_running_update(const cunet_bn_update_desc* descs_dev, int ndesc, void* stream) {
  if (ndesc <= 0) return 0;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  bn_running_update_kernel<<<ndesc, 128, 0, st>>>(descs_dev);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cunet_fail_cuda("bn_running_update launch", e);
  return 0;
}

extern "C" int cunet_rmsprop_step(float* params, const float* grads, float* square_avg, long n, const float* lr_dev,
                                  float alpha, float eps, void* stream) {
  if (n <= 0) return 0;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  rmsprop_kernel<<<148 * 4, 256, 0, st>>>(params, grads, square_avg, n, lr_dev, alpha, eps);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cunet_fail_cuda("rmsprop launch", e);
  return 0;
}
