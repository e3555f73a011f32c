// Operand gathers shared by the forward / dgrad / wgrad GEMM kernels.
//
//   activation operand : relu(scale_k * x + shift_k) over a virtual concat (cunet_concat), optional 3x3 tap
//                        shift with zero padding, optional nearest-x2 upsampled source.
//   gradient operand   : dT = istd*(G - c1 - (T - mu)*c2) (batch-norm backward form) or plain G, optional
//                        routing through the 2x2 max-pool argmax, optional 3x3 tap shift.
// Both produce 16-byte chunks ready for the swizzled row tiles of common.cuh.
//
// The gathers are written for instruction economy (the round-1 ncu captures showed the loaders were
// issue-bound, not bandwidth-bound): all per-row index math is done once per tile (RowCtx), the segment lookup
// once per K step (ActStep), and the bf16 path transforms two elements per instruction (HFMA2.BF16 / HMNMX2).
#pragma once
#include "common.cuh"
#include "../../include/cunet_b200.h"

namespace cunet {

constexpr int MAX_CIN = 384;  // order-3 up-block adapter: 256 + 4*32

// Pipeline-stage layout shared by the GEMM kernels.  bf16: [A 16K][B 16K].  fp32 (3xTF32 split):
// [A_hi 16K][A_lo 16K][B_hi ..][B_lo ..] (B hi/lo contiguous: one bulk copy brings both).
template <typename T> struct StageGeom {
  static constexpr bool SPLIT = Elem<T>::SPLIT;
  static constexpr int BYTES = SPLIT ? 65536 : 32768;
  static constexpr int A_LO = 16384;
  static constexpr int B_OFF = SPLIT ? 32768 : 16384;
  static constexpr int MIN_CTAS = SPLIT ? 1 : 2;
};

struct alignas(16) BnSmem {
  float scale[MAX_CIN];
  float shift[MAX_CIN];
  float mean[MAX_CIN];
  float istd[MAX_CIN];
  uint32_t sc2[MAX_CIN / 2];  // bf16x2 copies of scale / shift for the packed bf16 transform
  uint32_t sh2[MAX_CIN / 2];
  int seg_start[CUNET_MAX_SEG + 1];
  int cin;
  int relu;  // 1: ReLU;  0: identity input (bn_train == 2, the stem's im2col operand)
  int act_bits;  // QuanInput between the ReLU and the conv (0: none), see ActQuant
};

// QuanInput (utils/quantize.py:47-63) on the activated operand a = relu(bn(x)) >= 0:
//   a <- round_half_even(min(a, 1 - 2^-(b-1)) * 2^(b-1)) / 2^(b-1)
// bf16: adding and subtracting 2^(8-b) rounds a bf16 in [0, 1) to that grid exactly (the sum has ulp 2^-(b-1) and the
// hardware add rounds to nearest even); fp32: rintf.  Backward (dgrad epilogues): straight-through, zero where a >= 1.
struct ActQuant {
  int bits;
  float qmax, scale, inv;   // 1 - 2^-(b-1), 2^(b-1), 2^-(b-1)
  uint32_t qmax2, magic2;   // bf16x2
  __device__ __forceinline__ void init(int b) {
    bits = (b >= 2 && b <= 8) ? b : 0;
    scale = (float)(1 << (bits > 0 ? bits - 1 : 0));
    inv = 1.f / scale;
    qmax = 1.f - inv;
    const float magic = (float)(1 << (bits > 0 ? 8 - bits : 0));
    qmax2 = pack_bf16x2_raw(qmax, qmax);
    magic2 = pack_bf16x2_raw(magic, magic);
  }
  static __device__ __forceinline__ uint32_t pack_bf16x2_raw(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
};

// 1 / sqrt(v) in double: fp32 rsqrt as the seed, two Newton steps in fp64 (relative error < 1e-15).  The straight
// `1.0 / sqrt(v)` is a double-precision square root plus a double-precision division -- two long software sequences on
// a GPU whose fp64 rate is a small fraction of fp32 -- and EVERY conv launch computes it per input channel in its
// prologue, after griddepcontrol.wait, i.e. on the critical path of the ~630-launch chain of a step.
__device__ __forceinline__ double inv_sqrt_f64(double v) {
  double y = (double)rsqrtf((float)v);
  y = y * (1.5 - 0.5 * v * y * y);
  y = y * (1.5 - 0.5 * v * y * y);
  return y;
}

__device__ __forceinline__ int concat_cin(const cunet_concat& in) {
  int c = 0;
  for (int s = 0; s < in.nseg; ++s) c += in.seg[s].C;
  return c;
}

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// BatchNorm coefficients of the op's own BatchNorm over the concat (nn.BatchNorm2d: biased batch variance
// in train mode, running statistics in eval mode).  kpad: channels to fill (zeros beyond Cin), even.
__device__ __forceinline__ void compute_bn_coefs(const cunet_concat& in, BnSmem* b, int kpad, int tid, int nthreads) {
  if (tid == 0) {
    int acc = 0;
    for (int s = 0; s < in.nseg; ++s) {
      b->seg_start[s] = acc;
      acc += in.seg[s].C;
    }
    for (int s = in.nseg; s <= CUNET_MAX_SEG; ++s) b->seg_start[s] = acc;
    b->cin = acc;
    b->relu = in.bn_train == 2 ? 0 : 1;
    b->act_bits = in.act_bits;
  }
  const int Cin = concat_cin(in);
  for (int k2 = tid; k2 < kpad / 2; k2 += nthreads) {
    float scv[2], shv[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int k = 2 * k2 + j;
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
        const double istd = inv_sqrt_f64(var + (double)in.eps);
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
      scv[j] = sc;
      shv[j] = sh;
    }
    b->sc2[k2] = pack_bf16x2(scv[0], scv[1]);
    b->sh2[k2] = pack_bf16x2(shv[0], shv[1]);
  }
}

// ------------------------------------------------------------------------------------------------
// per-thread row context: the (up to 4) tile rows a loader thread gathers
struct RowCtx {
  int rd[4];  // full-resolution row index (n*H + h)*W + w, or -1 when the tile row is past the end
  int ru[4];  // half-resolution row index (n*(H/2) + h/2)*(W/2) + w/2  (upsampled / pooled sources)
  int hw[4];  // (h << 16) | w   (3x3 bounds checks, pool position)
};

__device__ __forceinline__ void rowctx_set(RowCtx& rc, int q, bool valid, int n, int h, int w, int H, int W) {
  rc.rd[q] = valid ? (n * H + h) * W + w : -1;
  rc.ru[q] = valid ? (n * (H >> 1) + (h >> 1)) * (W >> 1) + (w >> 1) : 0;
  rc.hw[q] = (h << 16) | w;
}

// pixel index -> (n, h, w); shifts when H and W are powers of two (always, for 256x256 inputs)
struct PixDiv {
  int H, W, lh, lw;  // lh / lw < 0: not a power of two
  __device__ __forceinline__ void init(int H_, int W_) {
    H = H_; W = W_;
    lw = (W & (W - 1)) == 0 ? __ffs(W) - 1 : -1;
    lh = (H & (H - 1)) == 0 ? __ffs(H) - 1 : -1;
  }
  __device__ __forceinline__ void split(int m, int& n, int& h, int& w) const {
    if ((lw | lh) >= 0) {
      w = m & (W - 1);
      const int t = m >> lw;
      h = t & (H - 1);
      n = t >> lh;
    } else {
      w = m % W;
      const int t = m / W;
      h = t % H;
      n = t / H;
    }
  }
};

// one K step of the activation operand as seen by one thread (its fixed 16-byte chunk column)
struct ActStep {
  const char* base;  // segment base + channel byte offset
  int ldb;           // row stride in bytes
  int up;
  int ch;            // concat channel of the chunk's first element
  bool ok;
};

template <typename T>
__device__ __forceinline__ ActStep act_step(const cunet_concat& in, const BnSmem* b, int ch) {
  using E = Elem<T>;
  ActStep st;
  st.ch = ch;
  st.ok = ch < b->cin;
  st.base = nullptr;
  st.ldb = 0;
  st.up = 0;
  if (st.ok) {
    int s = 0;
    while (ch >= b->seg_start[s + 1]) ++s;
    const cunet_seg& sg = in.seg[s];
    st.base = reinterpret_cast<const char*>(sg.ptr) + (size_t)(ch - b->seg_start[s]) * E::ESZ;
    st.ldb = sg.ld * E::ESZ;
    st.up = sg.up;
  }
  return st;
}

// (dy, dx): 3x3 tap offset, zero for 1x1
__device__ __forceinline__ bool act_load(const ActStep& st, const RowCtx& rc, int q, int H, int W, int dy, int dx,
                                         uint4& raw) {
  if (!st.ok || rc.rd[q] < 0) return false;
  int row;
  if (st.up) {
    row = rc.ru[q];
  } else {
    row = rc.rd[q];
    if ((dy | dx) != 0) {
      const int h = (rc.hw[q] >> 16) + dy, w = (rc.hw[q] & 0xffff) + dx;
      if ((unsigned)h >= (unsigned)H || (unsigned)w >= (unsigned)W) return false;
      row += dy * W + dx;
    }
  }
  raw = ldg128(st.base + (long)row * st.ldb);
  return true;
}

// coefficients of one chunk column for one K step, loaded once and reused for the thread's rows
template <typename T> struct ActCoef;
template <> struct ActCoef<bf16> {
  uint4 sc, sh;
  int relu;
  uint32_t qmax2 = 0, magic2 = 0;   // magic2 != 0: QuanInput on the activated operand (ActQuant)
  __device__ __forceinline__ void set_quant(int act_bits) {
    ActQuant q;
    q.init(act_bits);
    qmax2 = q.bits ? q.qmax2 : 0u;
    magic2 = q.bits ? q.magic2 : 0u;
  }
  __device__ __forceinline__ void load(const BnSmem* b, int ch) {
    sc = *reinterpret_cast<const uint4*>(&b->sc2[ch >> 1]);
    sh = *reinterpret_cast<const uint4*>(&b->sh2[ch >> 1]);
    relu = b->relu;
    set_quant(b->act_bits);
  }
  __device__ __forceinline__ uint4 apply(const uint4& raw, uint4& lo) const {
    const uint32_t x[4] = {raw.x, raw.y, raw.z, raw.w};
    const uint32_t s[4] = {sc.x, sc.y, sc.z, sc.w};
    const uint32_t t[4] = {sh.x, sh.y, sh.z, sh.w};
    uint32_t o[4];
    const __nv_bfloat162 zero = __floats2bfloat162_rn(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __nv_bfloat162 v = __hfma2(*reinterpret_cast<const __nv_bfloat162*>(&x[i]),
                                 *reinterpret_cast<const __nv_bfloat162*>(&s[i]),
                                 *reinterpret_cast<const __nv_bfloat162*>(&t[i]));
      if (relu) v = __hmax2(v, zero);
      if (magic2) {
        const __nv_bfloat162 m = *reinterpret_cast<const __nv_bfloat162*>(&magic2);
        v = __hmin2(v, *reinterpret_cast<const __nv_bfloat162*>(&qmax2));
        v = __hsub2(__hadd2(v, m), m);
      }
      o[i] = *reinterpret_cast<uint32_t*>(&v);
    }
    (void)lo;
    return make_uint4(o[0], o[1], o[2], o[3]);
  }
};
template <> struct ActCoef<float> {
  float4 sc, sh;
  float fl;
  float qmax = 0.f, qscale = 0.f, qinv = 0.f;   // qscale != 0: QuanInput on the activated operand (ActQuant)
  __device__ __forceinline__ void load(const BnSmem* b, int ch) {
    sc = *reinterpret_cast<const float4*>(&b->scale[ch]);
    sh = *reinterpret_cast<const float4*>(&b->shift[ch]);
    fl = b->relu ? 0.f : -__int_as_float(0x7f800000);
    ActQuant q;
    q.init(b->act_bits);
    qmax = q.qmax;
    qscale = q.bits ? q.scale : 0.f;
    qinv = q.inv;
  }
  __device__ __forceinline__ uint4 apply(const uint4& raw, uint4& lo) const {
    float f[4];
    f[0] = fmaxf(fmaf(__uint_as_float(raw.x), sc.x, sh.x), fl);
    f[1] = fmaxf(fmaf(__uint_as_float(raw.y), sc.y, sh.y), fl);
    f[2] = fmaxf(fmaf(__uint_as_float(raw.z), sc.z, sh.z), fl);
    f[3] = fmaxf(fmaf(__uint_as_float(raw.w), sc.w, sh.w), fl);
    if (qscale != 0.f) {
#pragma unroll
      for (int e = 0; e < 4; ++e) f[e] = rintf(fminf(f[e], qmax) * qscale) * qinv;
    }
    lo = Chunk<float>::pack_lo(f);
    return Chunk<float>::pack_mma(f);
  }
};

// ------------------------------------------------------------------------------------------------
// gradient operand
// dT = istd * (G - c1 - (T - mu) * c2) = a*G + b*(T - mu) + d   with a = istd, b = -istd*c2, d = -istd*c1,
// c1 = mean(G), c2 = istd * mean(G * xhat)   (centered form; see cunet_grad_src in the header).
// gstats layout: [0,C) = sum G, [C,2C) = sum G * xhat (accumulated centered by the consumers' dgrad epilogues).
struct alignas(16) GradSmem {
  float a[128];
  float b[128];
  float mu[128];
  float d[128];
  uint32_t a2[64], b2[64], mu2[64], d2[64];  // bf16x2 copies
};

__device__ __forceinline__ void compute_grad_coefs(const cunet_grad_src& gs, GradSmem* g, int tid, int nthreads) {
  for (int c2i = tid; c2i < 64; c2i += nthreads) {
    float av[2], bv[2], mv[2], dv[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int c = 2 * c2i + j;
      float a = 1.f, b = 0.f, mu = 0.f, d = 0.f;
      if (gs.mode == 1 && c < gs.C) {
        const double n_inv = gs.inv_count;
        const double mean = gs.stats[c] * n_inv;
        double var = gs.stats[gs.C + c] * n_inv - mean * mean;
        if (var < 0.0) var = 0.0;
        const double is = inv_sqrt_f64(var + (double)gs.eps);
        const double c1 = gs.gstats[c] * n_inv;
        const double c2 = is * gs.gstats[gs.C + c] * n_inv;
        a = (float)is;
        b = (float)(-is * c2);
        mu = (float)mean;
        d = (float)(-is * c1);
      }
      g->a[c] = a; g->b[c] = b; g->mu[c] = mu; g->d[c] = d;
      av[j] = a; bv[j] = b; mv[j] = mu; dv[j] = d;
    }
    g->a2[c2i] = pack_bf16x2(av[0], av[1]);
    g->b2[c2i] = pack_bf16x2(bv[0], bv[1]);
    g->mu2[c2i] = pack_bf16x2(mv[0], mv[1]);
    g->d2[c2i] = pack_bf16x2(dv[0], dv[1]);
  }
}

template <typename T> struct GradRaw {
  uint4 g, t;
  uint32_t idx[2];
  uint32_t pos;
};

// channel `co` (multiple of EPC) of the gradient for tile row q, optionally tap-shifted by (dy, dx)
template <typename T>
__device__ __forceinline__ bool grad_load(const cunet_grad_src& gs, const RowCtx& rc, int q, int H, int W, int co,
                                          int dy, int dx, GradRaw<T>& raw) {
  using E = Elem<T>;
  if (co >= gs.C || rc.rd[q] < 0) return false;
  int row;
  if (gs.pooled) {
    row = rc.ru[q];
    raw.pos = (uint32_t)((((rc.hw[q] >> 16) & 1) << 1) | (rc.hw[q] & 1));
  } else {
    row = rc.rd[q];
    if ((dy | dx) != 0) {
      const int h = (rc.hw[q] >> 16) + dy, w = (rc.hw[q] & 0xffff) + dx;
      if ((unsigned)h >= (unsigned)H || (unsigned)w >= (unsigned)W) return false;
      row += dy * W + dx;
    }
  }
  const long off = ((long)row * gs.ld + co) * E::ESZ;
  raw.g = ldg128(reinterpret_cast<const char*>(gs.g) + off);
  if (gs.mode == 1) raw.t = ldg128(reinterpret_cast<const char*>(gs.t) + off);
  if (gs.pooled) {
    const uint8_t* ip = gs.pool_idx + (long)row * gs.C + co;
    raw.idx[0] = *reinterpret_cast<const uint32_t*>(ip);
    if (E::EPC == 8) raw.idx[1] = *reinterpret_cast<const uint32_t*>(ip + 4);
  }
  return true;
}

template <typename T> struct GradCoef;
template <> struct GradCoef<bf16> {
  uint4 a, b, mu, d;
  __device__ __forceinline__ void load(const GradSmem* g, int co) {
    a = *reinterpret_cast<const uint4*>(&g->a2[co >> 1]);
    b = *reinterpret_cast<const uint4*>(&g->b2[co >> 1]);
    mu = *reinterpret_cast<const uint4*>(&g->mu2[co >> 1]);
    d = *reinterpret_cast<const uint4*>(&g->d2[co >> 1]);
  }
  __device__ __forceinline__ uint4 apply(const cunet_grad_src& gs, const GradRaw<bf16>& raw, uint4& lo) const {
    uint32_t o[4] = {raw.g.x, raw.g.y, raw.g.z, raw.g.w};
    if (gs.mode == 1) {
      const uint32_t t[4] = {raw.t.x, raw.t.y, raw.t.z, raw.t.w};
      const uint32_t av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
      const uint32_t mv[4] = {mu.x, mu.y, mu.z, mu.w}, dv[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const __nv_bfloat162 tc = __hsub2(*reinterpret_cast<const __nv_bfloat162*>(&t[i]),
                                          *reinterpret_cast<const __nv_bfloat162*>(&mv[i]));
        const __nv_bfloat162 u = __hfma2(*reinterpret_cast<const __nv_bfloat162*>(&bv[i]), tc,
                                         *reinterpret_cast<const __nv_bfloat162*>(&dv[i]));
        __nv_bfloat162 v = __hfma2(*reinterpret_cast<const __nv_bfloat162*>(&av[i]),
                                   *reinterpret_cast<const __nv_bfloat162*>(&o[i]), u);
        o[i] = *reinterpret_cast<uint32_t*>(&v);
      }
    }
    if (gs.pooled) {
      const uint32_t posw = raw.pos * 0x01010101u;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint32_t eq = __vcmpeq4(raw.idx[i >> 1], posw);  // 0xFF per matching byte
        const uint32_t m = (i & 1) ? __byte_perm(eq, 0, 0x3322) : __byte_perm(eq, 0, 0x1100);
        o[i] &= m;
      }
    }
    (void)lo;
    return make_uint4(o[0], o[1], o[2], o[3]);
  }
};
template <> struct GradCoef<float> {
  float4 a, b, mu, d;
  __device__ __forceinline__ void load(const GradSmem* g, int co) {
    a = *reinterpret_cast<const float4*>(&g->a[co]);
    b = *reinterpret_cast<const float4*>(&g->b[co]);
    mu = *reinterpret_cast<const float4*>(&g->mu[co]);
    d = *reinterpret_cast<const float4*>(&g->d[co]);
  }
  __device__ __forceinline__ uint4 apply(const cunet_grad_src& gs, const GradRaw<float>& raw, uint4& lo) const {
    float f[4] = {__uint_as_float(raw.g.x), __uint_as_float(raw.g.y), __uint_as_float(raw.g.z),
                  __uint_as_float(raw.g.w)};
    if (gs.mode == 1) {
      f[0] = fmaf(a.x, f[0], fmaf(b.x, __uint_as_float(raw.t.x) - mu.x, d.x));
      f[1] = fmaf(a.y, f[1], fmaf(b.y, __uint_as_float(raw.t.y) - mu.y, d.y));
      f[2] = fmaf(a.z, f[2], fmaf(b.z, __uint_as_float(raw.t.z) - mu.z, d.z));
      f[3] = fmaf(a.w, f[3], fmaf(b.w, __uint_as_float(raw.t.w) - mu.w, d.w));
    }
    if (gs.pooled) {
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (((raw.idx[0] >> (8 * e)) & 0xFFu) != raw.pos) f[e] = 0.f;
    }
    lo = Chunk<float>::pack_lo(f);
    return Chunk<float>::pack_mma(f);
  }
};

}  // namespace cunet
