// Operand gathers shared by the forward / dgrad / wgrad GEMM kernels.
//
//   activation operand : relu(scale_k * x + shift_k) over a virtual concat (cunet_concat), optional 3x3 tap
//                        shift with zero padding, optional nearest-x2 upsampled source.
//   gradient operand   : dT = p_c*G + q_c*T + r_c (batch-norm backward form) or plain G, optional routing
//                        through the 2x2 max-pool argmax, optional 3x3 tap shift.
// Both produce 16-byte chunks ready for the SWIZZLE_128B row tiles of common.cuh.
#pragma once
#include "common.cuh"
#include "../../include/cunet_b200.h"

namespace cunet {

constexpr int MAX_CIN = 512;

// Pipeline-stage layout shared by the GEMM kernels.  bf16: [A 16K][B 16K].  fp32 (3xTF32 split):
// [A_hi 16K][A_lo 16K][B_hi ..][B_lo ..] (B hi/lo contiguous: one bulk copy brings both).
template <typename T> struct StageGeom {
  static constexpr bool SPLIT = Elem<T>::SPLIT;
  static constexpr int BYTES = SPLIT ? 65536 : 32768;
  static constexpr int A_LO = 16384;
  static constexpr int B_OFF = SPLIT ? 32768 : 16384;
  static constexpr int MIN_CTAS = SPLIT ? 1 : 2;
};

struct BnSmem {
  float scale[MAX_CIN];
  float shift[MAX_CIN];
  float mean[MAX_CIN];
  float istd[MAX_CIN];
  int seg_start[CUNET_MAX_SEG + 1];
  int cin;
  float floor;  // 0: ReLU;  -inf: identity input (bn_train == 2, the stem's im2col operand)
};

__device__ __forceinline__ int concat_cin(const cunet_concat& in) {
  int c = 0;
  for (int s = 0; s < in.nseg; ++s) c += in.seg[s].C;
  return c;
}

// BatchNorm coefficients of the op's own BatchNorm over the concat (nn.BatchNorm2d: biased batch variance
// in train mode, running statistics in eval mode).  kpad: channels to fill (zeros beyond Cin).
__device__ __forceinline__ void compute_bn_coefs(const cunet_concat& in, BnSmem* b, int kpad, int tid, int nthreads) {
  if (tid == 0) {
    int acc = 0;
    for (int s = 0; s < in.nseg; ++s) {
      b->seg_start[s] = acc;
      acc += in.seg[s].C;
    }
    for (int s = in.nseg; s <= CUNET_MAX_SEG; ++s) b->seg_start[s] = acc;
    b->cin = acc;
    b->floor = in.bn_train == 2 ? -__int_as_float(0x7f800000) : 0.f;
  }
  const int Cin = concat_cin(in);
  for (int k = tid; k < kpad; k += nthreads) {
    float sc = 0.f, sh = 0.f, mu = 0.f, is = 0.f;
    if (k < Cin && in.bn_train == 2) {
      sc = 1.f;
      is = 1.f;
    } else if (k < Cin) {
      double mean, var;
      if (in.bn_train) {
        int s = 0, base = 0;
        while (s + 1 < in.nseg && k >= base + in.seg[s].C) {
          base += in.seg[s].C;
          ++s;
        }
        const int c = k - base;
        const double su = in.seg[s].stats[c], sq = in.seg[s].stats[in.seg[s].C + c];
        mean = su * in.seg[s].inv_count;
        var = sq * in.seg[s].inv_count - mean * mean;
        if (var < 0.0) var = 0.0;
      } else {
        mean = in.rmean[k];
        var = in.rvar[k];
      }
      const double istd = 1.0 / sqrt(var + (double)in.eps);
      const double g = in.gamma[k];
      sc = (float)(g * istd);
      sh = (float)((double)in.beta[k] - mean * g * istd);
      mu = (float)mean;
      is = (float)istd;
    }
    b->scale[k] = sc;
    b->shift[k] = sh;
    b->mean[k] = mu;
    b->istd[k] = is;
  }
}

// ------------------------------------------------------------------------------------------------
// activation chunk: concat channel `ch` (multiple of EPC) of pixel (n, h+dy, w+dx); false -> zero chunk
template <typename T>
__device__ __forceinline__ bool act_issue(const cunet_concat& in, const BnSmem* b, int H, int W, int ch, int n,
                                          int h, int w, int dy, int dx, uint4& raw) {
  using E = Elem<T>;
  if (ch >= b->cin) return false;
  const int hh = h + dy, ww = w + dx;
  if (hh < 0 || hh >= H || ww < 0 || ww >= W) return false;
  int s = 0;
  while (ch >= b->seg_start[s + 1]) ++s;
  const cunet_seg& sg = in.seg[s];
  long row;
  if (sg.up)
    row = ((long)n * (H >> 1) + (hh >> 1)) * (W >> 1) + (ww >> 1);
  else
    row = ((long)n * H + hh) * W + ww;
  raw = ldg128(reinterpret_cast<const char*>(sg.ptr) + (row * sg.ld + (ch - b->seg_start[s])) * E::ESZ);
  return true;
}

template <typename T>
__device__ __forceinline__ uint4 act_transform(const BnSmem* b, int ch, const uint4& raw, uint4& lo) {
  using E = Elem<T>;
  float f[E::EPC];
  Chunk<T>::unpack(raw, f);
  const float fl = b->floor;
#pragma unroll
  for (int e = 0; e < E::EPC; ++e) f[e] = fmaxf(fmaf(f[e], b->scale[ch + e], b->shift[ch + e]), fl);
  if (E::SPLIT) lo = Chunk<T>::pack_lo(f);
  return Chunk<T>::pack_mma(f);
}

// ------------------------------------------------------------------------------------------------
// gradient operand
// dT = istd * (G - c1 - (T - mu) * c2),  c1 = mean(G),  c2 = istd * mean(G * xhat)   (centered form: no
// cancellation between large q*T and r terms; see cunet_grad_src in the header)
struct GradSmem {
  float istd[128];
  float c1[128];
  float mu[128];
  float c2[128];
};

// gstats layout: [0,C) = sum G, [C,2C) = sum G * xhat  (xhat = (T - mean) * istd, accumulated centered by the
// consumers' dgrad epilogues)
__device__ __forceinline__ void compute_grad_coefs(const cunet_grad_src& gs, GradSmem* g, int tid, int nthreads) {
  for (int c = tid; c < 128; c += nthreads) {
    float istd = 1.f, c1 = 0.f, mu = 0.f, c2 = 0.f;
    if (gs.mode == 1 && c < gs.C) {
      const double n_inv = gs.inv_count;
      const double mean = gs.stats[c] * n_inv;
      double var = gs.stats[gs.C + c] * n_inv - mean * mean;
      if (var < 0.0) var = 0.0;
      const double is = 1.0 / sqrt(var + (double)gs.eps);
      istd = (float)is;
      c1 = (float)(gs.gstats[c] * n_inv);
      mu = (float)mean;
      c2 = (float)(is * gs.gstats[gs.C + c] * n_inv);
    }
    g->istd[c] = istd;
    g->c1[c] = c1;
    g->mu[c] = mu;
    g->c2[c] = c2;
  }
}

template <typename T> struct GradRaw {
  uint4 g, t;
  uint32_t idx[2];
  uint32_t pos;
};

// channel `co` (multiple of EPC) of the gradient at output pixel (n, h, w) [already tap-shifted by caller]
template <typename T>
__device__ __forceinline__ bool grad_issue(const cunet_grad_src& gs, int H, int W, int co, int n, int h, int w,
                                           GradRaw<T>& raw) {
  using E = Elem<T>;
  if (co >= gs.C) return false;
  if (h < 0 || h >= H || w < 0 || w >= W) return false;
  long row;
  if (gs.pooled)
    row = ((long)n * (H >> 1) + (h >> 1)) * (W >> 1) + (w >> 1);
  else
    row = ((long)n * H + h) * W + w;
  const long off = (row * gs.ld + co) * E::ESZ;
  raw.g = ldg128(reinterpret_cast<const char*>(gs.g) + off);
  if (gs.mode == 1) raw.t = ldg128(reinterpret_cast<const char*>(gs.t) + off);
  if (gs.pooled) {
    const uint8_t* ip = gs.pool_idx + row * gs.C + co;
    raw.idx[0] = *reinterpret_cast<const uint32_t*>(ip);
    if (E::EPC == 8) raw.idx[1] = *reinterpret_cast<const uint32_t*>(ip + 4);
    raw.pos = (uint32_t)(((h & 1) << 1) | (w & 1));
  }
  return true;
}

template <typename T>
__device__ __forceinline__ uint4 grad_transform(const cunet_grad_src& gs, const GradSmem* gc, int co,
                                                const GradRaw<T>& raw, uint4& lo) {
  using E = Elem<T>;
  float g[E::EPC];
  Chunk<T>::unpack(raw.g, g);
  if (gs.mode == 1) {
    float t[E::EPC];
    Chunk<T>::unpack(raw.t, t);
#pragma unroll
    for (int e = 0; e < E::EPC; ++e)
      g[e] = gc->istd[co + e] * (g[e] - gc->c1[co + e] - (t[e] - gc->mu[co + e]) * gc->c2[co + e]);
  }
  if (gs.pooled) {
#pragma unroll
    for (int e = 0; e < E::EPC; ++e) {
      const uint32_t id = (raw.idx[e >> 2] >> ((e & 3) * 8)) & 0xFFu;
      if (id != raw.pos) g[e] = 0.f;
    }
  }
  if (E::SPLIT) lo = Chunk<T>::pack_lo(g);
  return Chunk<T>::pack_mma(g);
}

}  // namespace cunet
