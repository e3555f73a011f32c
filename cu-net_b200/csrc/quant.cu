// Weight / activation quantizers: QuanOp (utils/quantize.py:77-175), BinOp
// (models/cu_net_prev_version.py:17-92), QuanInput (utils/quantize.py:47-63).
// The reference loops in Python over ~80 tensors with ~10 tiny kernels each per call; here every phase is ONE
// launch: a thread block per filter, filter staged in shared memory, reductions in fp64 (block tree).
#include "common.cuh"
#include "../../include/cunet_b200.h"
#include "host_util.h"

namespace cunet {

constexpr int QT = 128;         // threads per filter block
constexpr int QMAX = 384 * 9;   // largest filter (elements)

__device__ __forceinline__ float q_scale(int bits) { return exp2f((float)(bits - 1)); }       // S(bits)
__device__ __forceinline__ float q_clip(float x, int bits) {                                   // C(x, bits)
  const float delta = (bits > 15 || bits == 1 || bits == 2) ? 0.f : 1.f / q_scale(bits);
  return fminf(fmaxf(x, -1.f + delta), 1.f - delta);
}
__device__ __forceinline__ float q_sign(float x) { return x > 0.f ? 1.f : (x < 0.f ? -1.f : 0.f); }
__device__ __forceinline__ float q_round(float x, int bits) {                                  // Q(x, bits)
  if (bits > 15) return x;
  if (bits == 1) return q_sign(x);
  if (bits == 2) return rintf(x);               // torch.round: half to even
  const float s = q_scale(bits);
  return rintf(x * s) / s;
}

__device__ __forceinline__ double block_sum(double v, double* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  for (int i = 0; i < QT / 32; ++i) t += red[i];
  return t;
}

__device__ __forceinline__ const cunet_quant_desc* find_desc(const cunet_quant_desc* d, int nd, int b) {
  int lo = 0, hi = nd - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (d[mid].first_block <= b) lo = mid; else hi = mid - 1;
  }
  return d + lo;
}

__global__ void __launch_bounds__(QT) quant_forward_kernel(const cunet_quant_desc* descs, int nd, int mode, int bw, int bg) {
  __shared__ float f[QMAX];
  __shared__ float tmean[16];
  __shared__ double red[QT / 32];
  const cunet_quant_desc d = *find_desc(descs, nd, blockIdx.x);
  const int co = blockIdx.x - d.first_block;
  const int n = d.Cin * d.taps;
  float* w = d.w + (long)co * n;
  for (int i = threadIdx.x; i < n; i += QT) f[i] = w[i];
  __syncthreads();
  // mean over the input-channel dimension, per kernel tap (w.mean(1, True))
  for (int t = threadIdx.x >> 5; t < d.taps; t += QT / 32) {
    double s = 0.0;
    for (int ci = threadIdx.x & 31; ci < d.Cin; ci += 32) s += (double)f[ci * d.taps + t];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) tmean[t] = (float)(s / (double)d.Cin);
  }
  __syncthreads();
  double asum = 0.0;
  for (int i = threadIdx.x; i < n; i += QT) {
    float v = f[i] + (-tmean[i % d.taps]);
    v = mode == 1 ? fminf(fmaxf(v, -1.f), 1.f) : q_clip(v, bg);
    f[i] = v;
    d.saved[(long)co * n + i] = mode == 1 ? v : q_round(v, bg);
    asum += (double)fabsf(v);
  }
  const float m = (float)(block_sum(asum, red) / (double)n);   // mean_filter |w|
  for (int i = threadIdx.x; i < n; i += QT) {
    const float v = f[i];
    float o;
    if (mode == 1) {
      o = q_sign(v) * m;
    } else if (bw == 1) {
      // sign(w)*Q(m) and then the fall-through Q(C(., 1), 1) of quantize.py:148-149  ->  values in {-1, 0, +1}
      o = q_sign(q_clip(q_sign(v) * q_round(m, bg), 1));
    } else if (bw == 2) {
      const float dd = m * 0.7f;
      o = (v > dd ? 1.f : 0.f) - (v < -dd ? 1.f : 0.f);
    } else {
      o = q_round(q_clip(v, bw), bw);
    }
    w[i] = o;
  }
}

__global__ void __launch_bounds__(QT) quant_restore_kernel(const cunet_quant_desc* descs, int nd) {
  const cunet_quant_desc d = *find_desc(descs, nd, blockIdx.x);
  const int co = blockIdx.x - d.first_block;
  const int n = d.Cin * d.taps;
  for (int i = threadIdx.x; i < n; i += QT) d.w[(long)co * n + i] = d.saved[(long)co * n + i];
}

__global__ void __launch_bounds__(QT) quant_grad_kernel(const cunet_quant_desc* descs, int nd, int mode, int bw, int bg) {
  __shared__ double red[QT / 32];
  const cunet_quant_desc d = *find_desc(descs, nd, blockIdx.x);
  if (!d.grad) return;
  const int co = blockIdx.x - d.first_block;
  const int n = d.Cin * d.taps;
  const float* w = d.w + (long)co * n;
  float* g = d.grad + (long)co * n;
  if (mode == 0 && bw != 1) {
    for (int i = threadIdx.x; i < n; i += QT) g[i] = q_round(q_clip(g[i], bg), bg);
    return;
  }
  double asum = 0.0, ssum = 0.0;
  for (int i = threadIdx.x; i < n; i += QT) {
    asum += (double)fabsf(w[i]);
    ssum += (double)(q_sign(w[i]) * g[i]);
  }
  const float m = (float)(block_sum(asum, red) / (double)n);
  const float sg = (float)(block_sum(ssum, red) / (double)n);
  const float fac = (float)(1.0 - 1.0 / (double)d.Cin);
  for (int i = threadIdx.x; i < n; i += QT) {
    const float wi = w[i];
    float mm = (wi < -1.f || wi > 1.f) ? 0.f : m;
    if (mode == 0) mm = q_round(mm, bg);
    float o = (mm * g[i] + sg * q_sign(wi)) * fac * (float)n;
    if (mode == 0) o = q_round(q_clip(o, bg), bg);
    g[i] = o;
  }
}

__global__ void quant_input_fwd_kernel(const float* x, float* y, long n, int bits) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
    y[i] = q_round(q_clip(x[i], bits), bits);
}
__global__ void quant_input_bwd_kernel(const float* x, const float* dy, float* dx, long n) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
    dx[i] = (x[i] >= 1.f || x[i] <= -1.f) ? 0.f : dy[i];
}

}  // namespace cunet
using namespace cunet;

static int launch_check(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cunet_fail_cuda(what, e);
  return 0;
}

extern "C" int cunet_quant_forward(const cunet_quant_desc* descs_dev, int ndesc, int nblocks, int mode, int bits_w,
                                   int bits_g, void* stream) {
  if (ndesc <= 0 || nblocks <= 0) return 0;
  quant_forward_kernel<<<nblocks, QT, 0, reinterpret_cast<cudaStream_t>(stream)>>>(descs_dev, ndesc, mode, bits_w, bits_g);
  return launch_check("quant_forward launch");
}
extern "C" int cunet_quant_restore(const cunet_quant_desc* descs_dev, int ndesc, int nblocks, void* stream) {
  if (ndesc <= 0 || nblocks <= 0) return 0;
  quant_restore_kernel<<<nblocks, QT, 0, reinterpret_cast<cudaStream_t>(stream)>>>(descs_dev, ndesc);
  return launch_check("quant_restore launch");
}
extern "C" int cunet_quant_grad(const cunet_quant_desc* descs_dev, int ndesc, int nblocks, int mode, int bits_w,
                                int bits_g, void* stream) {
  if (ndesc <= 0 || nblocks <= 0) return 0;
  quant_grad_kernel<<<nblocks, QT, 0, reinterpret_cast<cudaStream_t>(stream)>>>(descs_dev, ndesc, mode, bits_w, bits_g);
  return launch_check("quant_grad launch");
}
extern "C" int cunet_quant_input_fwd(const float* x, float* y, long n, int bits, void* stream) {
  if (n <= 0) return 0;
  quant_input_fwd_kernel<<<148 * 4, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, y, n, bits);
  return launch_check("quant_input_fwd launch");
}
extern "C" int cunet_quant_input_bwd(const float* x, const float* dy, float* dx, long n, void* stream) {
  if (n <= 0) return 0;
  quant_input_bwd_kernel<<<148 * 4, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, dy, dx, n);
  return launch_check("quant_input_bwd launch");
}
