// Multi-loss MSE + gradient + argmax landmark decode in one pass over the heads.
//   loss = sum_k sum((out_k - heatmap)^2) / numel          (cu-net.py:175-178)
//   d loss / d out_k = 2 (out_k - heatmap) / numel
//   decode (last head): argmax over H*W, first maximum, (x, y) 1-based, zero where max <= 0
//                                                            (pylib/Evaluation.py:6-23)
// One block per (n, h) row of the heatmap: the NCHW target row-tile is staged in shared memory once and
// reused by all heads; head rows are read / gradient rows written fully coalesced (NHWC).
#include "common.cuh"
#include "../../include/cunet_b200.h"
#include "host_util.h"

namespace cunet {

__device__ __forceinline__ unsigned int float_orderable(float f) {
  unsigned int u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float float_from_orderable(unsigned int u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u);
}

template <typename T>
__global__ void __launch_bounds__(256) mse_decode_kernel(const cunet_mse_params p) {
  extern __shared__ float sm[];
  float* tgt = sm;                                                       // [C][W+1]
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(sm + ((p.C * (p.W + 1) + 1) & ~1));  // [C]
  __shared__ float wsum[8];
  const int n = blockIdx.x / p.H, h = blockIdx.x - n * p.H;
  const int W = p.W, C = p.C, ld = p.ld;
  for (int i = threadIdx.x; i < C * W; i += blockDim.x) {
    const int c = i / W, w = i - c * W;
    tgt[c * (W + 1) + w] = p.target[(((long)n * C + c) * p.H + h) * W + w];
  }
  for (int c = threadIdx.x; c < C; c += blockDim.x) keys[c] = 0ull;
  __syncthreads();
  const long row0 = ((long)n * p.H + h) * W;
  const float gscale = 2.f * p.grad_scale / (float)((long)p.N * C * p.H * W);
  for (int k = 0; k < p.nheads; ++k) {
    const float* o = p.heads[k] + row0 * ld;
    T* d = p.dheads[k] ? reinterpret_cast<T*>(p.dheads[k]) + row0 * ld : nullptr;
    float acc = 0.f;
    const bool last = (k == p.nheads - 1) && p.keys != nullptr;
    for (int i = threadIdx.x; i < W * ld; i += blockDim.x) {
      const int w = i / ld, c = i - w * ld;
      float g = 0.f;
      if (c < C) {
        const float v = o[i];
        const float diff = v - tgt[c * (W + 1) + w];
        acc += diff * diff;
        g = diff * gscale;
        if (last) {
          const unsigned long long key =
              ((unsigned long long)float_orderable(v) << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)(h * W + w));
          atomicMax(&keys[c], key);
        }
      }
      if (d) d[i] = from_f<T>(g);
    }
    // block reduction of the squared error
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.f;
      for (int i = 0; i < 8; ++i) t += wsum[i];
      const double v = (double)t / (double)((long)p.N * C * p.H * W);
      atomicAdd(p.loss, v);
      atomicAdd(p.loss + 1 + k, v);
    }
    __syncthreads();
  }
  if (p.keys) {
    for (int c = threadIdx.x; c < C; c += blockDim.x) atomicMax(p.keys + (long)n * C + c, keys[c]);
  }
}

// Vector form for ld % 4 == 0 (the network's heads: ld = C = 68): one pass over the row tile with ALL heads in the inner
// loop -- 16-byte loads of eight independent heads per position instead of one scalar load per element, the target read
// once per position, one block reduction and one set of loss atomics per block instead of one per head (the scalar
// kernel ran 234 us for 321 MB at batch 24; 1536 blocks x 8 heads of double atomics on the single total-loss word).
template <typename T>
__global__ void __launch_bounds__(256) mse_decode_v4_kernel(const __grid_constant__ cunet_mse_params p) {
  extern __shared__ float sm[];
  float* tgt = sm;                                                       // [C][W+1]
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(sm + ((p.C * (p.W + 1) + 1) & ~1));  // [C]
  __shared__ float wsum[8][16];
  __shared__ double hsum[16];
  const int n = blockIdx.x / p.H, h = blockIdx.x - n * p.H;
  const int W = p.W, C = p.C, ld = p.ld, ld4 = p.ld >> 2, nh = p.nheads;
  for (int i = threadIdx.x; i < C * W; i += blockDim.x) {
    const int c = i / W, w = i - c * W;
    tgt[c * (W + 1) + w] = p.target[(((long)n * C + c) * p.H + h) * W + w];
  }
  for (int c = threadIdx.x; c < C; c += blockDim.x) keys[c] = 0ull;
  __syncthreads();
  const long row0 = ((long)n * p.H + h) * W;
  const float gscale = 2.f * p.grad_scale / (float)((long)p.N * C * p.H * W);
  float acc[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) acc[k] = 0.f;
  for (int i = threadIdx.x; i < W * ld4; i += blockDim.x) {
    const int w = i / ld4, c = (i - w * ld4) * 4;
    float t[4];
    bool ok[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      ok[j] = c + j < C;
      t[j] = ok[j] ? tgt[(c + j) * (W + 1) + w] : 0.f;
    }
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      if (k >= nh) break;
      const float4 v4 = *reinterpret_cast<const float4*>(p.heads[k] + row0 * ld + (long)i * 4);
      const float v[4] = {v4.x, v4.y, v4.z, v4.w};
      float g[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float diff = ok[j] ? v[j] - t[j] : 0.f;
        acc[k] = fmaf(diff, diff, acc[k]);
        g[j] = diff * gscale;
      }
      if (p.dheads[k]) {
        T* d = reinterpret_cast<T*>(p.dheads[k]) + row0 * ld + (long)i * 4;
        if (sizeof(T) == 2) {
          uint2 o;
          const __nv_bfloat162 lo = __floats2bfloat162_rn(g[0], g[1]), hi = __floats2bfloat162_rn(g[2], g[3]);
          o.x = *reinterpret_cast<const uint32_t*>(&lo);
          o.y = *reinterpret_cast<const uint32_t*>(&hi);
          *reinterpret_cast<uint2*>(d) = o;
        } else {
          *reinterpret_cast<float4*>(d) = make_float4(g[0], g[1], g[2], g[3]);
        }
      }
      if (k == nh - 1 && p.keys != nullptr) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (ok[j]) {
            const unsigned long long key = ((unsigned long long)float_orderable(v[j]) << 32) |
                                           (unsigned long long)(0xFFFFFFFFu - (unsigned)(h * W + w));
            atomicMax(&keys[c + j], key);
          }
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    if (k >= nh) break;
    float a = acc[k];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) a += __shfl_xor_sync(0xffffffffu, a, off);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5][k] = a;
  }
  __syncthreads();
  if ((int)threadIdx.x < nh) {
    float tsum = 0.f;
    for (int i = 0; i < 8; ++i) tsum += wsum[i][threadIdx.x];
    const double v = (double)tsum / (double)((long)p.N * C * p.H * W);
    hsum[threadIdx.x] = v;
    atomicAdd(p.loss + 1 + threadIdx.x, v);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int k = 0; k < nh; ++k) tot += hsum[k];
    atomicAdd(p.loss, tot);
  }
  if (p.keys) {
    for (int c = threadIdx.x; c < C; c += blockDim.x) atomicMax(p.keys + (long)n * C + c, keys[c]);
  }
}

__global__ void decode_finalize_kernel(const unsigned long long* keys, float* preds, int NC, int W) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= NC) return;
  const unsigned long long key = keys[i];
  const unsigned int idx = 0xFFFFFFFFu - (unsigned int)(key & 0xFFFFFFFFull);
  const float v = float_from_orderable((unsigned int)(key >> 32));
  float x = (float)(idx % W + 1), y = (float)(idx / W + 1);  // Evaluation.py:18-19 (square maps)
  if (!(v > 0.f)) { x = 0.f; y = 0.f; }
  preds[2 * i] = x;
  preds[2 * i + 1] = y;
}

}  // namespace cunet
using namespace cunet;

extern "C" int cunet_mse_decode(const cunet_mse_params* p, void* stream) {
  if (!p) return cunet_fail("mse_decode: null params");
  if (p->nheads < 1 || p->nheads > 16) return cunet_fail("mse_decode: 1..16 heads");
  if (p->C > p->ld) return cunet_fail("mse_decode: ld < C");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const size_t smem = (((size_t)p->C * (p->W + 1) + 1) & ~(size_t)1) * 4 + (size_t)p->C * 8;
  if (smem > 200 * 1024) return cunet_fail("mse_decode: tile too large");
  cudaError_t e;
  bool vec = p->ld % 4 == 0 && p->W * (p->ld / 4) >= 1;
  for (int k = 0; k < p->nheads && vec; ++k)
    vec = (reinterpret_cast<uintptr_t>(p->heads[k]) & 15) == 0 && (reinterpret_cast<uintptr_t>(p->dheads[k]) & 15) == 0;
  if (vec) {
    if (p->dtype == CUNET_BF16) {
      e = cudaFuncSetAttribute(mse_decode_v4_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return cunet_fail_cuda("mse_decode attr", e);
      mse_decode_v4_kernel<bf16><<<p->N * p->H, 256, smem, st>>>(*p);
    } else {
      e = cudaFuncSetAttribute(mse_decode_v4_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return cunet_fail_cuda("mse_decode attr", e);
      mse_decode_v4_kernel<float><<<p->N * p->H, 256, smem, st>>>(*p);
    }
  } else if (p->dtype == CUNET_BF16) {
    e = cudaFuncSetAttribute(mse_decode_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cunet_fail_cuda("mse_decode attr", e);
    mse_decode_kernel<bf16><<<p->N * p->H, 256, smem, st>>>(*p);
  } else {
    e = cudaFuncSetAttribute(mse_decode_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cunet_fail_cuda("mse_decode attr", e);
    mse_decode_kernel<float><<<p->N * p->H, 256, smem, st>>>(*p);
  }
  e = cudaGetLastError();
  if (e != cudaSuccess) return cunet_fail_cuda("mse_decode launch", e);
  return 0;
}

extern "C" int cunet_decode_finalize(const unsigned long long* keys, float* preds, int NC, int W, void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  decode_finalize_kernel<<<(NC + 255) / 256, 256, 0, st>>>(keys, preds, NC, W);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cunet_fail_cuda("decode_finalize launch", e);
  return 0;
}
