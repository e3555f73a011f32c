// Stem: conv0 (7x7 s2 p3, 3->128) -> norm0 -> relu0 -> pool0   (models/cu_net.py:299-304).
// conv0 itself runs on the tensor cores via cunet_conv_fwd / cunet_conv_wgrad over an im2col matrix built
// here; the BatchNorm+ReLU+MaxPool that follows and their backward are bandwidth-bound elementwise kernels.
#include "loaders.cuh"
#include "host_util.h"

namespace cunet {

constexpr int STEM_K = 147;      // 3 * 7 * 7
constexpr int STEM_KPAD = 160;

// K index -> (input offset relative to the window's centre row/column, dy, dx), built once per block: the first version
// decomposed k with two integer divisions per ELEMENT (186 us per step at batch 24 for a 126 MB matrix).
template <typename T>
__global__ void __launch_bounds__(256) stem_im2col_kernel(const float* __restrict__ img, T* __restrict__ cols, int N,
                                                          int Hi, int Wi) {
  using E = Elem<T>;
  __shared__ int tab_off[STEM_KPAD];
  __shared__ int tab_yx[STEM_KPAD];   // (dy + 3) << 8 | (dx + 3), or -1 for the padding columns
  for (int k = threadIdx.x; k < STEM_KPAD; k += blockDim.x) {
    if (k < STEM_K) {
      const int c = k / 49, r = k - c * 49, kh = r / 7, kw = r - kh * 7;
      tab_off[k] = (c * Hi + kh - 3) * Wi + kw - 3;
      tab_yx[k] = (kh << 8) | kw;
    } else {
      tab_off[k] = 0;
      tab_yx[k] = -1;
    }
  }
  __syncthreads();
  const int Ho = Hi >> 1, Wo = Wi >> 1;
  constexpr int CPR = STEM_KPAD / E::EPC;
  const long total = (long)N * Ho * Wo * CPR;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int j = (int)(i % CPR);
    const long px = i / CPR;
    const int ox = (int)(px % Wo);
    const long t = px / Wo;
    const int oy = (int)(t % Ho), n = (int)(t / Ho);
    const float* base = img + ((long)n * 3 * Hi + 2 * oy) * Wi + 2 * ox;
    float f[E::EPC];
#pragma unroll
    for (int e = 0; e < E::EPC; ++e) {
      const int k = j * E::EPC + e;
      const int yx = tab_yx[k];
      const int iy = 2 * oy + (yx >> 8) - 3, ix = 2 * ox + (yx & 255) - 3;
      float v = 0.f;
      if (yx >= 0 && (unsigned)iy < (unsigned)Hi && (unsigned)ix < (unsigned)Wi) v = base[tab_off[k]];
      f[e] = v;
    }
    // two column blocks, each a dense matrix: [px][128] followed by [px][32] -- the concat-segment widths the
    // persistent conv kernels take (a 160-wide row would straddle their 128-channel chunks)
    constexpr int C0V = 128 / E::EPC;   // 16-byte vectors per row of block 0
    const long npx = (long)N * Ho * Wo;
    const long vec = j < C0V ? px * C0V + j : npx * C0V + px * (CPR - C0V) + (j - C0V);
    *reinterpret_cast<uint4*>(reinterpret_cast<char*>(cols) + vec * 16) = Chunk<T>::pack(f);
  }
}

// per-channel BatchNorm coefficients of norm0 into shared memory
struct StemBn {
  float scale[128], shift[128], mean[128], istd[128];
};
__device__ __forceinline__ void stem_bn_coefs(StemBn* b, const double* stats, const float* gamma, const float* beta,
                                              const float* rmean, const float* rvar, int bn_train, double inv_n,
                                              float eps) {
  for (int c = threadIdx.x; c < 128; c += blockDim.x) {
    double mean, var;
    if (bn_train) {
      mean = stats[c] * inv_n;
      var = stats[128 + c] * inv_n - mean * mean;
      if (var < 0.0) var = 0.0;
    } else {
      mean = rmean[c];
      var = rvar[c];
    }
    const double istd = 1.0 / sqrt(var + (double)eps);
    b->scale[c] = (float)(gamma[c] * istd);
    b->shift[c] = (float)((double)beta[c] - mean * gamma[c] * istd);
    b->mean[c] = (float)mean;
    b->istd[c] = (float)istd;
  }
}

// block: 256 threads = 8 row lanes x 32 channel quads; 64 pooled pixels per block
template <typename T>
__global__ void __launch_bounds__(256) stem_pool_fwd_kernel(const cunet_stem_pool_params p) {
  __shared__ StemBn bn;
  __shared__ float red[256 * 8];
  const long n_y = (long)p.N * p.H * p.W;
  stem_bn_coefs(&bn, p.y_stats, p.gamma, p.beta, p.rmean, p.rvar, p.bn_train, 1.0 / (double)n_y, p.eps);
  __syncthreads();
  const int quad = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int Hp = p.H >> 1, Wp = p.W >> 1;
  const long npool = (long)p.N * Hp * Wp;
  float sc[4], sh[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    sc[e] = bn.scale[quad * 4 + e];
    sh[e] = bn.shift[quad * 4 + e];
  }
  const T* y = reinterpret_cast<const T*>(p.y);
  T* x = reinterpret_cast<T*>(p.x);
  float s1[4] = {0, 0, 0, 0}, s2[4] = {0, 0, 0, 0};
  for (int i = rl; i < 64; i += 8) {
    const long q = (long)blockIdx.x * 64 + i;
    if (q >= npool) break;
    const int wp = (int)(q % Wp);
    const long t = q / Wp;
    const int hp = (int)(t % Hp), n = (int)(t / Hp);
    float m[4] = {0.f, 0.f, 0.f, 0.f};  // relu output is >= 0, so 0 is the identity of the max
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const long row = ((long)n * p.H + 2 * hp + (k >> 1)) * p.W + 2 * wp + (k & 1);
      float v[4];
      load4<T>(y + row * 128 + quad * 4, v);
#pragma unroll
      for (int e = 0; e < 4; ++e) m[e] = fmaxf(m[e], fmaf(v[e], sc[e], sh[e]));
    }
    store4<T>(x + q * 128 + quad * 4, m);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      s1[e] += m[e];
      s2[e] += m[e] * m[e];
    }
  }
  if (p.x_stats) {
    float4* rp = reinterpret_cast<float4*>(red + threadIdx.x * 8);
    rp[0] = make_float4(s1[0], s1[1], s1[2], s1[3]);
    rp[1] = make_float4(s2[0], s2[1], s2[2], s2[3]);
    __syncthreads();
    if (threadIdx.x < 128) {
      const int q = threadIdx.x >> 2, e = threadIdx.x & 3;
      float a = 0.f, b = 0.f;
      for (int t = q; t < 256; t += 32) {
        a += red[t * 8 + e];
        b += red[t * 8 + 4 + e];
      }
      atomicAdd(p.x_stats + threadIdx.x, (double)a);
      atomicAdd(p.x_stats + 128 + threadIdx.x, (double)b);
    }
  }
}

// backward: phase 0 = parameter-gradient reduction, phase 1 = dy
template <typename T>
__global__ void __launch_bounds__(256) stem_bwd_kernel(const cunet_stem_bwd_params p) {
  __shared__ StemBn bn;
  __shared__ GradSmem gc;
  __shared__ float red[256 * 8];
  const long n_y = (long)p.N * p.H * p.W;
  stem_bn_coefs(&bn, p.y_stats, p.gamma, p.beta, nullptr, nullptr, 1, 1.0 / (double)n_y, p.eps);
  compute_grad_coefs(p.dx, &gc, threadIdx.x, blockDim.x);
  __syncthreads();
  const int quad = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int Hp = p.H >> 1, Wp = p.W >> 1;
  const long npool = (long)p.N * Hp * Wp;
  const T* y = reinterpret_cast<const T*>(p.y);
  const T* G = reinterpret_cast<const T*>(p.dx.g);
  const T* X = reinterpret_cast<const T*>(p.dx.t);
  T* dy = reinterpret_cast<T*>(p.dy);
  float sc[4], sh[4], mu[4], is[4], ca[4], cb[4], cc[4];
  const float inv_n = (float)(1.0 / (double)n_y);
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int c = quad * 4 + e;
    sc[e] = bn.scale[c]; sh[e] = bn.shift[c]; mu[e] = bn.mean[c]; is[e] = bn.istd[c];
    if (p.phase == 1) {
      ca[e] = p.gamma[c] * bn.istd[c];      // gamma * istd
      cb[e] = p.dbeta[c] * inv_n;           // mean(dz)
      cc[e] = p.dgamma[c] * inv_n;          // mean(dz * yhat)
    }
  }
  float s1[4] = {0, 0, 0, 0}, s2[4] = {0, 0, 0, 0};
  for (int i = rl; i < 64; i += 8) {
    const long q = (long)blockIdx.x * 64 + i;
    if (q >= npool) break;
    const int wp = (int)(q % Wp);
    const long t = q / Wp;
    const int hp = (int)(t % Hp), n = (int)(t / Hp);
    float g[4], xv[4], dxv[4];
    load4<T>(G + q * p.dx.ld + quad * 4, g);
    load4<T>(X + q * p.dx.ld + quad * 4, xv);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int c = quad * 4 + e;
      dxv[e] = fmaf(gc.a[c], g[e], fmaf(gc.b[c], xv[e] - gc.mu[c], gc.d[c]));
    }
    float v[4][4], z[4][4];
    int am[4] = {0, 0, 0, 0};
    float best[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const long row = ((long)n * p.H + 2 * hp + (k >> 1)) * p.W + 2 * wp + (k & 1);
      load4<T>(y + row * 128 + quad * 4, v[k]);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        z[k][e] = fmaf(v[k][e], sc[e], sh[e]);
        const float a = fmaxf(z[k][e], 0.f);
        if (k == 0) best[e] = a;
        else if (a > best[e]) { best[e] = a; am[e] = k; }  // nn.MaxPool2d: first maximum wins
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float o[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float dz = (am[e] == k && z[k][e] > 0.f) ? dxv[e] : 0.f;
        const float yh = (v[k][e] - mu[e]) * is[e];
        if (p.phase == 0) {
          s1[e] += dz;
          s2[e] += dz * yh;
        } else {
          o[e] = ca[e] * (dz - cb[e] - yh * cc[e]);
        }
      }
      if (p.phase == 1) {
        const long row = ((long)n * p.H + 2 * hp + (k >> 1)) * p.W + 2 * wp + (k & 1);
        store4<T>(dy + row * 128 + quad * 4, o);
      }
    }
  }
  if (p.phase == 0) {
    float4* rp = reinterpret_cast<float4*>(red + threadIdx.x * 8);
    rp[0] = make_float4(s1[0], s1[1], s1[2], s1[3]);
    rp[1] = make_float4(s2[0], s2[1], s2[2], s2[3]);
    __syncthreads();
    if (threadIdx.x < 128) {
      const int q = threadIdx.x >> 2, e = threadIdx.x & 3;
      float a = 0.f, b = 0.f;
      for (int t = q; t < 256; t += 32) {
        a += red[t * 8 + e];
        b += red[t * 8 + 4 + e];
      }
      atomicAdd(p.dbeta + threadIdx.x, a);
      atomicAdd(p.dgamma + threadIdx.x, b);
    }
  }
}

}  // namespace cunet
using namespace cunet;

extern "C" int cunet_stem_im2col(const float* img, void* cols, int N, int Hi, int Wi, int dtype, void* stream) {
  if ((Hi | Wi) & 1) return cunet_fail("stem_im2col: odd input size");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const long chunks = (long)N * (Hi / 2) * (Wi / 2) * (STEM_KPAD / (dtype == CUNET_BF16 ? 8 : 4));
  const int blocks = (int)((chunks + 255) / 256 > 148 * 16 ? 148 * 16 : (chunks + 255) / 256);
  if (dtype == CUNET_BF16)
    stem_im2col_kernel<bf16><<<blocks, 256, 0, st>>>(img, reinterpret_cast<bf16*>(cols), N, Hi, Wi);
  else
    stem_im2col_kernel<float><<<blocks, 256, 0, st>>>(img, reinterpret_cast<float*>(cols), N, Hi, Wi);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cunet_fail_cuda("stem_im2col launch", e);
  return 0;
}

extern "C" int cunet_stem_pool_fwd(const cunet_stem_pool_params* p, void* stream) {
  if (!p) return cunet_fail("stem_pool_fwd: null params");
  if ((p->H | p->W) & 1) return cunet_fail("stem_pool_fwd: odd size");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const long npool = (long)p->N * (p->H / 2) * (p->W / 2);
  const int blocks = (int)((npool + 63) / 64);
  if (p->dtype == CUNET_BF16)
    stem_pool_fwd_kernel<bf16><<<blocks, 256, 0, st>>>(*p);
  else
    stem_pool_fwd_kernel<float><<<blocks, 256, 0, st>>>(*p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cunet_fail_cuda("stem_pool_fwd launch", e);
  return 0;
}

extern "C" int cunet_stem_bwd(const cunet_stem_bwd_params* p, void* stream) {
  if (!p) return cunet_fail("stem_bwd: null params");
  if (p->dx.mode != 1 || p->dx.C != 128) return cunet_fail("stem_bwd: dx must be the 128-channel batch-norm form");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const long npool = (long)p->N * (p->H / 2) * (p->W / 2);
  const int blocks = (int)((npool + 63) / 64);
  if (p->dtype == CUNET_BF16)
    stem_bwd_kernel<bf16><<<blocks, 256, 0, st>>>(*p);
  else
    stem_bwd_kernel<float><<<blocks, 256, 0, st>>>(*p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cunet_fail_cuda("stem_bwd launch", e);
  return 0;
}
