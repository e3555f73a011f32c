// Fused backward of the 1x1 fused conv (cat -> BN -> ReLU -> conv1x1 [-> pool]; adapters, bottleneck conv1,
// intermedia adapters, heat-map heads; bf16): backward-data AND backward-filter in ONE persistent kernel.
// Same math as cunet_conv_dgrad + cunet_conv_wgrad for taps == 1 (see conv_dgrad.cu / conv_wgrad.cu):
//
//   dgrad : dA[p][k]   = sum_co dT[p][co] * W[co][k]      then per source: dz = dA * [bn(x) > 0], dgamma / dbeta,
//                                                          G_src += gamma * dz (4-child sum for an upsampled source)
//   wgrad : dW[co][k] += sum_p  a[p][k]  * dT[p][co],      a = relu(bn(x))
//
// Both contractions consume the SAME gradient operand dT (the SWIZZLE_128B row tile [64 px][128 co] is a K-major
// operand for the first and an MN-major operand for the second -- common.cuh) and the SAME landed input x, which the
// round-1 pair of kernels (conv_dgrad_v2 + conv_wgrad_v2) each fetched from HBM and transformed on their own, in two
// launches that both wanted all 148 SMs.  Here, per stage of 64 pixels:
//   * G / T (/ pool argmax) of the output and every source piece of the virtual concat are contiguous blocks of NHWC
//     tensors: 1-D TMA bulk copies land them in shared memory (x double buffered, landed ONCE for both contractions);
//   * 8 transformer warps build dT = a*G + b*(T - mu) + d (batch-norm backward form, pool routing) and, per
//     128-channel chunk, a = relu(bn(x)), smem -> smem, as swizzled operand tiles;
//   * one thread issues  D1_c[128 k][64 px]   = Wimg_c[128 k][Cout] * dT^T       (fresh per (stage, chunk), TMEM
//                                                                                 double buffered)
//                        D2_c[128 k][Cout]   += A_c^T[128 k][64 px] * dT          (TMEM resident for the CTA's whole
//                                                                                 pixel range, one per chunk);
//     the dgrad weight image of the whole conv (<= 96 KB) stays resident in shared memory;
//   * 8 epilogue warps (thread = input channel = TMEM lane, wide tcgen05.ld) apply the ReLU mask, keep dbeta / dgamma
//     in registers and overwrite x IN PLACE with gamma*dz; one thread bulk-stores (or L2-reduce-adds) the pieces;
//   * at the end D2 is added to the fp32 weight gradient with coalesced red.global.add.
// Stage geometry: 64 consecutive raster pixels, except for 64-wide images with an upsampled source, where a stage is
// 2 rows x 32 columns (two runs of 32 pixels) so that every 2x2 window of the half-resolution source is complete
// inside one stage (its four children are summed in the epilogue).
// HBM traffic per launch = G, T of the output + every source once + every source gradient once.
#include "loaders.cuh"
#include "host_util.h"
#include <stdlib.h>

namespace cunet {

// warp 0 landing producer | 1 store issuer | 2 dgrad MMA issuer | 3 wgrad MMA issuer | 4-15 transformers | 16-31 epilogue.
// 16 epilogue warps (four per TMEM lane quarter, 16 pixels each): the epilogue is a chain of dependent shared-memory
// round trips per pixel, and with only two warps per scheduler the first version of this kernel spent 1.7 us per
// 128-channel chunk of a 64-pixel stage waiting on instruction latency (in-kernel timeline, tools/time_bwd1x1.py).
constexpr int F1_THREADS = 1024;
constexpr int F1_TRW = 12;      // transformer warps
constexpr int F1_R = 64;         // pixels per stage
constexpr int F1_SUB = F1_R * 128;  // one operand sub-tile: 64 rows x 128 B
constexpr int F1_DWSTG = 16 * 32 * 144;  // final epilogue: one [32 co][32 k] fp32 tile (144-byte pitch) per epilogue warp
constexpr int F1_GT_BYTES = 32768;  // one G/T landing set: G at +0, T at +16384, pool argmax bytes at +24576
constexpr int F1_MAXCH = 3;      // 128-channel chunks (Cin <= 384)

struct F1Layout {
  int w_off, dt_off, a_off, gt_off, x_off, tail_off;
  int x_bytes;      // one x buffer (all pieces of a stage)
  int gt_bufs;      // 1 or 2 G/T landing sets
  int dt_bufs;      // 1 or 2 gradient-operand buffers
  int a_slots;      // 2 or 3 activation-operand slots (ring over the (stage, chunk) items)
  int split;        // 1: stage = 2 rows x 32 columns (64-wide image with an upsampled source)
  int per;          // stages per CTA
  int npad;         // gradient operand columns (Cout rounded up to 64)
  int dwc;          // row length of dw
  float* dw;
};

struct F1Tail {
  alignas(16) uint32_t sc2[MAX_CIN / 2];   // bf16x2 BatchNorm scale / shift of the concat (transformers)
  alignas(16) uint32_t sh2[MAX_CIN / 2];
  alignas(16) uint32_t ga2[64], gb2[64], gmu2[64], gd2[64];   // bf16x2 gradient-form coefficients (GradSmem packed)
  // per concat channel, for the epilogue thread that owns it (kept here instead of in registers: with 896 threads the
  // register file allows 72 per thread, and 3 chunks x 7 constants spilled inside the per-pixel loop):
  //   x: thr, y: thr1 -- ReLU / QuanInput mask  (x > thr) != neg  [&& (x < thr1) != neg]
  //   z: gamma,  w: bit 31 neg | bit 30 up | bit 29 valid | bits 16..17 log2(C/32) | bits 0..15 byte offset of the
  //   channel inside an x buffer
  alignas(16) float4 ech[MAX_CIN];
  uint64_t w_full, done, dt_ready[2], dt_free[2];
  uint64_t gt_full[2], gt_free[2], x_full[2], x_free[2], d1_full[4], d1_free[4], a_full[3], a_free[3];
  uint64_t a_done[2][F1_MAXCH], g_ready[2][F1_MAXCH];
  uint32_t tmem_base;
  int seg_start[CUNET_MAX_SEG + 1];
  int xoff[CUNET_MAX_SEG];     // byte offset of segment s's piece inside an x buffer
  int woff[F1_MAXCH + 1];      // byte offset of chunk c's weight rows inside the resident image
  int lowmap[F1_R];            // stage row -> row of the half-resolution run
  int rowpos[F1_R];            // stage row -> position inside its 2x2 window ((h & 1) * 2 + (w & 1))
};

// the largest op of an order-1 network (320-channel up-block adapter: weights 80 KB + dT 16 KB + A slots 32 KB + one
// G/T set 32 KB + two 28 KB x buffers = 221184 bytes) must still fit beside the tail
static_assert(232448 - 1024 - (int)sizeof(F1Tail) >= 221184, "conv_bwd1x1: tail too large for the 320-channel op");

__device__ __forceinline__ void f1_bulk_s2g(void* dst, const void* src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src_smem)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void f1_bulk_red_add_bf16(void* dst, const void* src_smem, uint32_t bytes) {
  asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.noftz.bf16 [%0], [%1], %2;" ::"l"(dst),
               "r"(smem_u32(src_smem)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void f1_bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void f1_bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ float f1_lds_bf16(uint32_t saddr) {
  uint16_t v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(saddr));
  return __uint_as_float((uint32_t)v << 16);
}
__device__ __forceinline__ void f1_sts_u16(uint32_t saddr, uint16_t v) {
  asm volatile("st.shared.u16 [%0], %1;" ::"r"(saddr), "h"(v) : "memory");
}
// same with a compile-time byte offset folded into the instruction (no address arithmetic per access)
template <int OFF> __device__ __forceinline__ float f1_lds_bf16_o(uint32_t saddr) {
  uint16_t v;
  asm volatile("ld.shared.u16 %0, [%1+%2];" : "=h"(v) : "r"(saddr), "n"(OFF));
  return __uint_as_float((uint32_t)v << 16);
}
template <int OFF> __device__ __forceinline__ void f1_sts_u16_o(uint32_t saddr, uint16_t v) {
  asm volatile("st.shared.u16 [%0+%1], %2;" ::"r"(saddr), "n"(OFF), "h"(v) : "memory");
}
// 4 consecutive fp32 added to global memory with one reduction (sm_90+)
__device__ __forceinline__ void f1_red_add_v4(float* dst, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(dst), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
// 16 pixels of one channel (the thread's TMEM lane): dz = dA * [x on the active side of the ReLU threshold],
// accumulates sum dz and sum dz*x, overwrites x in place with gamma*dz.  CP2 = bytes between consecutive pixels.
template <int CP2, bool QUANT>
__device__ __forceinline__ void f1_ep16(uint32_t a0, float* v, float thr, float thr1, bool neg, float gm, float& db0,
                                        float& db1, float& dx0, float& dx1) {
  float x[16];
#define F1_LD(q) x[q] = f1_lds_bf16_o<(q) * CP2>(a0);
  F1_LD(0) F1_LD(1) F1_LD(2) F1_LD(3) F1_LD(4) F1_LD(5) F1_LD(6) F1_LD(7)
  F1_LD(8) F1_LD(9) F1_LD(10) F1_LD(11) F1_LD(12) F1_LD(13) F1_LD(14) F1_LD(15)
#undef F1_LD
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int q = 0; q < 16; q += 2) {
    const float dz0 = (((x[q] > thr) != neg) && (!QUANT || ((x[q] < thr1) != neg))) ? v[q] : 0.f;
    const float dz1 = (((x[q + 1] > thr) != neg) && (!QUANT || ((x[q + 1] < thr1) != neg))) ? v[q + 1] : 0.f;
    db0 += dz0;
    db1 += dz1;
    dx0 = fmaf(dz0, x[q], dx0);
    dx1 = fmaf(dz1, x[q + 1], dx1);
    v[q] = gm * dz0;
    v[q + 1] = gm * dz1;
  }
#define F1_ST(q) f1_sts_u16_o<(q) * CP2>(a0, __bfloat16_as_ushort(__float2bfloat16_rn(v[q])));
  F1_ST(0) F1_ST(1) F1_ST(2) F1_ST(3) F1_ST(4) F1_ST(5) F1_ST(6) F1_ST(7)
  F1_ST(8) F1_ST(9) F1_ST(10) F1_ST(11) F1_ST(12) F1_ST(13) F1_ST(14) F1_ST(15)
#undef F1_ST
}

__device__ __forceinline__ uint4 f1_lds128(uint32_t saddr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr));
  return v;
}
__device__ __forceinline__ uint2 f1_lds64(uint32_t saddr) {
  uint2 v;
  asm volatile("ld.shared.v2.b32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(saddr));
  return v;
}
// 32 lanes x 16 consecutive 32-bit columns -> 16 registers per thread; the wait is separate so that two loads can be
// in flight before the first use
__device__ __forceinline__ void f1_tmem_ld16_nowait(uint32_t taddr, float* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]), "=f"(v[8]),
        "=f"(v[9]), "=f"(v[10]), "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15])
      : "r"(taddr));
}
__device__ __forceinline__ void f1_tmem_ld8_nowait(uint32_t taddr, float* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
               : "r"(taddr));
}
__device__ __forceinline__ void f1_tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void f1_named_bar(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// geometry of one stage: first pixel of run 0 / run 1 (split only), valid pixel count, half-resolution run
struct F1Geo {
  int p0, p1, nv, low0, nlow;
};
__device__ __forceinline__ F1Geo f1_geo(int st, int M, int W, int split, int need_low) {
  F1Geo g;
  if (split) {                       // W == 64: row pair a = st >> 1, column half b = st & 1
    const int a = st >> 1, b = st & 1;
    g.p0 = a * 128 + 32 * b;
    g.p1 = g.p0 + 64;
    g.nv = 64;
    g.low0 = a * 32 + 16 * b;
    g.nlow = 16;
  } else {
    g.p0 = st * F1_R;
    g.p1 = -1;
    g.nv = min(F1_R, M - g.p0);
    g.low0 = 0;
    g.nlow = 0;
    if (need_low) {
      if (W == 64) {                 // one image row: its half-resolution row (H is even)
        g.low0 = ((g.p0 >> 6) >> 1) * 32;
        g.nlow = 32;
      } else {                       // whole row pairs (W <= 32): low pixels of the block are consecutive
        g.low0 = g.p0 >> 2;
        g.nlow = g.nv >> 2;
      }
    }
  }
  return g;
}

// timeline slots (CUNET_TRACE builds), stage i < 12 of CTA 0: producer 0+2i (G/T issued, x issued), transformer 32+4i
// (start, G/T landed + operand free, dT done, all chunk operands done), MMA 96+3i (dT ready, first D1 issued, stage
// issued), epilogue 144+4i (chunk 0 accumulator full, chunk 0 done, last chunk done), store 200+2i; 230 / 231: dW
// epilogue start / end
CUNET_TRACE_DECL(g_f1_trace)

__global__ void __launch_bounds__(F1_THREADS, 1) conv_bwd1x1_kernel(const __grid_constant__ cunet_conv_dgrad_params p,
                                                                     const __grid_constant__ F1Layout L) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  F1Tail* tail = reinterpret_cast<F1Tail*>(smem + L.tail_off);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int H = p.H, W = p.W;
  const int M = p.N * H * W;
  const int Cin = concat_cin(p.in);
  const int nchunk = (Cin + 127) >> 7;
  const int nkb = (p.CoutPad + 63) >> 6;
  const int total = (M + F1_R - 1) / F1_R;
  const int st0 = (int)blockIdx.x * L.per;
  const int st1 = min(total, st0 + L.per);
  const int ns = max(0, st1 - st0);
  const int ldo = p.dy.ld * 2;
  const int split = L.split;
  // TMEM: wgrad accumulators D2_c at columns c*128 (the last chunk only as wide as its real channels), then a ring of
  // 64-column dgrad accumulators D1 in what is left (2 buffers for 384 input channels, 3 for 320, 4 below)
  const int d2cols = ((Cin + 63) >> 6) << 6;
  const uint32_t d1_base = (uint32_t)d2cols;
  const uint32_t d1_nbuf = (uint32_t)min(4, (512 - d2cols) >> 6);
  int need_low = p.dy.pooled;
  for (int s = 0; s < p.in.nseg; ++s) need_low |= p.in.seg[s].up;
  CUNET_TRACE_LOAD(trace, g_f1_trace)

  if (tid == 0) {
    mbar_init(&tail->w_full, 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tail->dt_ready[b], F1_TRW);
      mbar_init(&tail->dt_free[b], 2);   // the dgrad and the wgrad MMA issuers both release the gradient operand
    }
    mbar_init(&tail->a_full[2], F1_TRW);
    mbar_init(&tail->a_free[2], 1);
    mbar_init(&tail->done, 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tail->gt_full[b], 1);
      mbar_init(&tail->gt_free[b], F1_TRW);
      mbar_init(&tail->x_full[b], 1);
      mbar_init(&tail->x_free[b], 1);
      mbar_init(&tail->d1_full[b], 1);
      mbar_init(&tail->d1_free[b], 16);
      mbar_init(&tail->d1_full[b + 2], 1);
      mbar_init(&tail->d1_free[b + 2], 16);
      mbar_init(&tail->a_full[b], F1_TRW);
      mbar_init(&tail->a_free[b], 1);
      for (int c = 0; c < F1_MAXCH; ++c) {
        mbar_init(&tail->a_done[b][c], F1_TRW);
        mbar_init(&tail->g_ready[b][c], 16);
      }
    }
    fence_mbar_init();
    // static tables
    int acc = 0, xo = 0;
    for (int s = 0; s < p.in.nseg; ++s) {
      tail->seg_start[s] = acc;
      acc += p.in.seg[s].C;
      tail->xoff[s] = xo;
      xo += (p.in.seg[s].up ? 16 : F1_R) * p.in.seg[s].C * 2;   // an upsampled source contributes 16 low pixels
    }
    for (int s = p.in.nseg; s <= CUNET_MAX_SEG; ++s) tail->seg_start[s] = acc;
    int wo = 0;
    for (int c = 0; c <= F1_MAXCH; ++c) {
      tail->woff[c] = wo;
      if (c < nchunk) wo += min(128, Cin - c * 128) * 128 * nkb;
    }
  }
  if (tid < F1_R) {
    // stage row -> half-resolution row of the stage's low run, and position inside the 2x2 window
    const int r = tid;
    int lm = 0, ps = 0;
    if (split) {
      lm = (r & 31) >> 1;
      ps = ((r >> 5) << 1) | (r & 1);
    } else if (W == 64) {
      lm = r >> 1;
      ps = r & 1;                     // + 2 * (image row parity), added per stage
    } else {
      const int lw = 31 - __clz(W);
      const int hl = r >> lw, w = r & (W - 1);
      lm = (hl >> 1) * (W >> 1) + (w >> 1);
      ps = ((hl & 1) << 1) | (w & 1);
    }
    tail->lowmap[r] = lm;
    tail->rowpos[r] = ps;
  }
  if (warp == 2) tmem_alloc(&tail->tmem_base, 512);
  __syncthreads();   // barrier inits and the static tables above are visible to every warp (warp 0 takes no part in the
                     // coefficient-phase barriers below); still before griddepcontrol.wait, i.e. under the previous kernel
  griddep_wait();    // everything above overlaps the tail of the previous kernel (programmatic dependent launch)
  griddep_launch();
  // landing of one stage (thread 0 only): G / T (/ argmax) of the output, then the source pieces
  auto land_stage = [&](int i) {
    const char* gsrc = reinterpret_cast<const char*>(p.dy.g);
    const char* tsrc = reinterpret_cast<const char*>(p.dy.t);
      const F1Geo g = f1_geo(st0 + i, M, W, split, need_low);
      // ---- G / T (/ argmax) of the output
      const int gbuf = L.gt_bufs == 2 ? (i & 1) : 0;
      const uint32_t gu = (uint32_t)(L.gt_bufs == 2 ? (i >> 1) : i);
      mbar_wait(&tail->gt_free[gbuf], (gu & 1u) ^ 1u);
      if (i < 12) CUNET_TRACE_MARK(trace, 0 + 2 * i);
      uint8_t* gdst = smem + L.gt_off + gbuf * F1_GT_BYTES;
      if (p.dy.pooled) {
        const uint32_t gb = (uint32_t)(g.nlow * ldo), ib = (uint32_t)(g.nlow * p.dy.C);
        mbar_arrive_expect_tx(&tail->gt_full[gbuf], 2u * gb + ib);
        bulk_g2s(gdst, gsrc + (long)g.low0 * ldo, gb, &tail->gt_full[gbuf]);
        bulk_g2s(gdst + 16384, tsrc + (long)g.low0 * ldo, gb, &tail->gt_full[gbuf]);
        bulk_g2s(gdst + 24576, p.dy.pool_idx + (long)g.low0 * p.dy.C, ib, &tail->gt_full[gbuf]);
      } else {
        const uint32_t nt = p.dy.mode == 1 ? 2u : 1u;
        if (split) {
          const uint32_t rb32 = (uint32_t)(32 * ldo);
          mbar_arrive_expect_tx(&tail->gt_full[gbuf], 2u * nt * rb32);
          bulk_g2s(gdst, gsrc + (long)g.p0 * ldo, rb32, &tail->gt_full[gbuf]);
          bulk_g2s(gdst + rb32, gsrc + (long)g.p1 * ldo, rb32, &tail->gt_full[gbuf]);
          if (nt == 2) {
            bulk_g2s(gdst + 16384, tsrc + (long)g.p0 * ldo, rb32, &tail->gt_full[gbuf]);
            bulk_g2s(gdst + 16384 + rb32, tsrc + (long)g.p1 * ldo, rb32, &tail->gt_full[gbuf]);
          }
        } else {
          const uint32_t gb = (uint32_t)(g.nv * ldo);
          mbar_arrive_expect_tx(&tail->gt_full[gbuf], nt * gb);
          bulk_g2s(gdst, gsrc + (long)g.p0 * ldo, gb, &tail->gt_full[gbuf]);
          if (nt == 2) bulk_g2s(gdst + 16384, tsrc + (long)g.p0 * ldo, gb, &tail->gt_full[gbuf]);
        }
      }
      // ---- the source pieces
      const uint32_t b = (uint32_t)i & 1u;
      mbar_wait(&tail->x_free[b], (((uint32_t)i >> 1) & 1u) ^ 1u);
      if (i < 12) CUNET_TRACE_MARK(trace, 1 + 2 * i);
      uint32_t xtot = 0;
      for (int s = 0; s < p.in.nseg; ++s) {
        const cunet_seg& sg = p.in.seg[s];
        xtot += (uint32_t)((sg.up ? g.nlow : g.nv) * sg.C * 2);
      }
      mbar_arrive_expect_tx(&tail->x_full[b], xtot);
      uint8_t* xdst = smem + L.x_off + b * L.x_bytes;
      for (int s = 0; s < p.in.nseg; ++s) {
        const cunet_seg& sg = p.in.seg[s];
        const char* src = reinterpret_cast<const char*>(sg.ptr);
        const int Cp2 = sg.C * 2;
        if (sg.up) {
          bulk_g2s(xdst + tail->xoff[s], src + (long)g.low0 * Cp2, (uint32_t)(g.nlow * Cp2), &tail->x_full[b]);
        } else if (split) {
          bulk_g2s(xdst + tail->xoff[s], src + (long)g.p0 * Cp2, (uint32_t)(32 * Cp2), &tail->x_full[b]);
          bulk_g2s(xdst + tail->xoff[s] + 32 * Cp2, src + (long)g.p1 * Cp2, (uint32_t)(32 * Cp2), &tail->x_full[b]);
        } else {
          bulk_g2s(xdst + tail->xoff[s], src + (long)g.p0 * Cp2, (uint32_t)(g.nv * Cp2), &tail->x_full[b]);
        }
      }

  };
  // Warp 0 is the landing producer and needs no coefficient: its thread issues the weight image and then runs its whole
  // landing loop right away; warps 1.. compute the coefficients behind NAMED barriers that do not include warp 0.
  if (tid == 0 && ns > 0) {
    // resident dgrad weight image: per chunk and K block only the rows that hold real input channels
    uint32_t wtot = 0;
    for (int c = 0; c < nchunk; ++c) wtot += (uint32_t)(min(128, Cin - c * 128) * 128 * nkb);
    mbar_arrive_expect_tx(&tail->w_full, wtot);
    for (int c = 0; c < nchunk; ++c) {
      const int rows = min(128, Cin - c * 128);
      for (int kb = 0; kb < nkb; ++kb)
        bulk_g2s(smem + L.w_off + tail->woff[c] + kb * rows * 128,
                 reinterpret_cast<const char*>(p.wpack_dgrad) + ((size_t)c * nkb + kb) * 16384, (uint32_t)(rows * 128),
                 &tail->w_full);
    }
    for (int i = 0; i < ns; ++i) land_stage(i);
  }
  // BatchNorm / gradient coefficients: computed in the (not yet used) activation-operand slots, then copied into the
  // tail tables.  NOT in the landing areas: thread 0 issues the weight image and stage 0's landings before this
  // phase (land_stage above), so that their HBM latency overlaps it.
  BnSmem* bn = reinterpret_cast<BnSmem*>(smem + L.a_off);
  GradSmem* gc = reinterpret_cast<GradSmem*>(smem + L.a_off + 8192);
  const int tq = tid - 32;                 // thread index among warps 1.. (the warps that share the coefficient phase)
  constexpr int TQN = F1_THREADS - 32;
  if (warp > 0) {
    compute_bn_coefs(p.in, bn, nchunk * 128, tq, TQN);
    compute_grad_coefs(p.dy, gc, tq, TQN);
    tc_fence_before();
    f1_named_bar(1, TQN);
    tc_fence_after();
  }
  const uint32_t tmem = warp > 0 ? tail->tmem_base : 0u;

  const bool is_tr = warp >= 4 && warp < 4 + F1_TRW, is_ep = warp >= 4 + F1_TRW;
  const int t = tid - 128;                 // transformer thread index (0..383)
  const int cc = t & 15, rb = t >> 4;      // 16-byte column x rows rb + 24q (rb < 24)
  const int e = warp - (4 + F1_TRW);
  const int qd = warp & 3, pq = (e >> 2) & 3;  // epilogue: TMEM lane quarter (hardware: warp % 4), pixel quarter
  const int k = qd * 32 + lane;                // epilogue: channel inside the chunk
  // transformer constants, one packed word per chunk (coefficients stay in shared memory: tail->sc2 / sh2 / ga2 ...):
  // bit 31 valid | bit 30 upsampled source | bits 16..17 log2(C/32) | bits 0..15 byte offset of this thread's 16-byte
  // column inside an x buffer.  (Four separate ints per chunk were spilled to local memory and reloaded in the hot loop.)
  uint32_t cs_pk[F1_MAXCH];
#pragma unroll
  for (int c = 0; c < F1_MAXCH; ++c) cs_pk[c] = 0u;
  // packed coefficient tables -> tail (they outlive the prologue area)
  for (int i = tq; warp > 0 && i < nchunk * 64; i += TQN) {
    tail->sc2[i] = bn->sc2[i];
    tail->sh2[i] = bn->sh2[i];
  }
  if (warp > 0 && tq < 64) {
    tail->ga2[tq] = gc->a2[tq];
    tail->gb2[tq] = gc->b2[tq];
    tail->gmu2[tq] = gc->mu2[tq];
    tail->gd2[tq] = gc->d2[tq];
  }
  if (is_tr) {
#pragma unroll
    for (int c = 0; c < F1_MAXCH; ++c) {
      const int ch = c * 128 + cc * 8;
      if (c < nchunk && ch < Cin) {
        int s = 0;
        while (ch >= tail->seg_start[s + 1]) ++s;
        const int C = p.in.seg[s].C;
        cs_pk[c] = 0x80000000u | (p.in.seg[s].up ? 0x40000000u : 0u) |
                   ((uint32_t)(C == 128 ? 2 : (C == 64 ? 1 : 0)) << 16) |
                   (uint32_t)(tail->xoff[s] + (ch - tail->seg_start[s]) * 2);
      }
    }
  }
  // epilogue table.  ReLU mask of a channel: bn(x) = sc*x + sh > 0  <=>  (x > thr) != neg  with thr = -sh/sc, neg = sc < 0;
  // with QuanInput between the ReLU and the conv (act_bits != 0) the straight-through gradient is also zero where
  // bn(x) >= 1:  (x < thr1) != neg  with thr1 = (1 - sh)/sc  (thr1 = -+inf otherwise, so that test is always true)
  for (int kg = tq; warp > 0 && kg < nchunk * 128; kg += TQN) {
    float4 ec = make_float4(0.f, 0.f, 0.f, 0.f);
    if (kg < Cin) {
      int sgi = 0;
      while (kg >= tail->seg_start[sgi + 1]) ++sgi;
      const float sc = bn->scale[kg], sh = bn->shift[kg];
      const float inf = __int_as_float(0x7f800000);
      const bool neg = sc < 0.f;
      ec.x = sc != 0.f ? -sh / sc : (sh > 0.f ? -inf : inf);   // sc == 0: the mask is the sign of the shift
      if (p.in.act_bits && sc != 0.f) ec.y = (1.f - sh) / sc;
      else if (p.in.act_bits) ec.y = sh < 1.f ? inf : -inf;
      else ec.y = neg ? -inf : inf;
      ec.z = p.in.gamma[kg];
      const int C = p.in.seg[sgi].C;
      const uint32_t w = (neg ? 0x80000000u : 0u) | (p.in.seg[sgi].up ? 0x40000000u : 0u) | 0x20000000u |
                         ((uint32_t)(C == 128 ? 2 : (C == 64 ? 1 : 0)) << 16) |
                         (uint32_t)(tail->xoff[sgi] + (kg - tail->seg_start[sgi]) * 2);
      ec.w = __uint_as_float(w);
    }
    tail->ech[kg] = ec;
  }
  if (warp > 0) f1_named_bar(1, TQN);    // tables complete; the operand slots (coefficient scratch) may be overwritten

  if (warp == 0) {
    // landing producer: done above
  } else if (warp == 1) {
    // ============================================================== G store issuer
    if (lane == 0) {
      for (int i = 0; i < ns; ++i) {
        const F1Geo g = f1_geo(st0 + i, M, W, split, need_low);
        const uint32_t b = (uint32_t)i & 1u, upar = ((uint32_t)i >> 1) & 1u;
        const uint8_t* xsrc = smem + L.x_off + b * L.x_bytes;
        for (int c = 0; c < nchunk; ++c) {
          mbar_wait(&tail->g_ready[b][c], upar);
          for (int s = 0; s < p.in.nseg; ++s) {
            if ((tail->seg_start[s] >> 7) != c || p.gacc[s].G == nullptr) continue;
            const cunet_seg& sg = p.in.seg[s];
            const int Cp2 = sg.C * 2;
            char* G = reinterpret_cast<char*>(p.gacc[s].G);
            const uint8_t* src = xsrc + tail->xoff[s];
            const int acc = p.gacc[s].accumulate;
            if (sg.up) {
              if (acc) f1_bulk_red_add_bf16(G + (long)g.low0 * Cp2, src, (uint32_t)(g.nlow * Cp2));
              else f1_bulk_s2g(G + (long)g.low0 * Cp2, src, (uint32_t)(g.nlow * Cp2));
            } else if (split) {
              if (acc) {
                f1_bulk_red_add_bf16(G + (long)g.p0 * Cp2, src, (uint32_t)(32 * Cp2));
                f1_bulk_red_add_bf16(G + (long)g.p1 * Cp2, src + 32 * Cp2, (uint32_t)(32 * Cp2));
              } else {
                f1_bulk_s2g(G + (long)g.p0 * Cp2, src, (uint32_t)(32 * Cp2));
                f1_bulk_s2g(G + (long)g.p1 * Cp2, src + 32 * Cp2, (uint32_t)(32 * Cp2));
              }
            } else {
              if (acc) f1_bulk_red_add_bf16(G + (long)g.p0 * Cp2, src, (uint32_t)(g.nv * Cp2));
              else f1_bulk_s2g(G + (long)g.p0 * Cp2, src, (uint32_t)(g.nv * Cp2));
            }
          }
          f1_bulk_commit();
        }
        f1_bulk_wait_read0();   // the x buffer may be overwritten once the stores have read it
        if (i < 12) CUNET_TRACE_MARK(trace, 200 + 2 * i);
        mbar_arrive(&tail->x_free[b]);
      }
    }
  } else if (warp == 2) {
    // ============================================================== dgrad MMA issuer: D1[(stage, chunk)] = W_c * dT^T
    if (lane == 0 && ns > 0) {
      const uint32_t idesc_d = make_idesc(Elem<bf16>::FMT, 128, 64, 0, 0);
      const uint32_t wA = smem_u32(smem + L.w_off);
      const uint64_t dtdesc00 = make_sdesc(smem_u32(smem + L.dt_off), 16, 1024);   // see sdesc_advance (common.cuh)
      mbar_wait(&tail->w_full, 0);
      uint32_t buf = 0, bph = 0;     // ring position of the next D1 accumulator and its use parity
      for (int i = 0; i < ns; ++i) {
        const uint32_t db = L.dt_bufs == 2 ? ((uint32_t)i & 1u) : 0u;
        const uint32_t du = (uint32_t)(L.dt_bufs == 2 ? (i >> 1) : i);
        const uint64_t dtdesc0 = sdesc_advance(dtdesc00, db * 2u * F1_SUB);
        mbar_wait(&tail->dt_ready[db], du & 1u);
        if (i < 12) CUNET_TRACE_MARK(trace, 96 + 3 * i);
        for (int c = 0; c < nchunk; ++c) {
          const int rows = min(128, Cin - c * 128);
          mbar_wait(&tail->d1_free[buf], bph ^ 1u);
          tc_fence_after();
          const uint32_t d1 = tmem + d1_base + buf * 64u;
          const uint64_t wdesc0 = make_sdesc(wA + (uint32_t)tail->woff[c], 16, 1024);
#pragma unroll
          for (int kb = 0; kb < 2; ++kb) {
            if (kb >= nkb) break;
            const uint64_t wd = sdesc_advance(wdesc0, (uint32_t)(kb * rows * 128));
            const uint64_t dd = sdesc_advance(dtdesc0, (uint32_t)(kb * F1_SUB));
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              umma<bf16>(d1, sdesc_advance(wd, kk * 32), sdesc_advance(dd, kk * 32), idesc_d, (uint32_t)((kb | kk) != 0));
          }
          tc_commit(&tail->d1_full[buf]);
          if (c == 0 && i < 12) CUNET_TRACE_MARK(trace, 97 + 3 * i);
          if (++buf == d1_nbuf) {
            buf = 0;
            bph ^= 1u;
          }
        }
        tc_commit(&tail->dt_free[db]);
        if (i < 12) CUNET_TRACE_MARK(trace, 98 + 3 * i);
      }
    }
  } else if (warp == 3) {
    // ============================================================== wgrad MMA issuer: D2_c[co][k] += dT^T * A_c
    // (output channel = TMEM lane, input channel = column: a thread of the final epilogue then holds consecutive k of
    //  one co, i.e. consecutive floats of dW[co][:], and adds them with 16-byte reductions)
    if (lane == 0 && ns > 0) {
      const uint64_t dtdesc00 = make_sdesc_mn<bf16>(smem_u32(smem + L.dt_off), F1_SUB);
      const uint64_t adesc0 = make_sdesc_mn<bf16>(smem_u32(smem + L.a_off), F1_SUB);
      uint32_t slot = 0, sph = 0;     // ring position of the next activation-operand slot and its use parity
      for (int i = 0; i < ns; ++i) {
        const uint32_t db = L.dt_bufs == 2 ? ((uint32_t)i & 1u) : 0u;
        const uint32_t du = (uint32_t)(L.dt_bufs == 2 ? (i >> 1) : i);
        const uint64_t dtdesc0 = sdesc_advance(dtdesc00, db * 2u * F1_SUB);
        mbar_wait(&tail->dt_ready[db], du & 1u);
        for (int c = 0; c < nchunk; ++c) {
          mbar_wait(&tail->a_full[slot], sph);
          tc_fence_after();
          const uint64_t ad = sdesc_advance(adesc0, slot * 16384u);
          const uint32_t idesc_w = make_idesc(Elem<bf16>::FMT, 128, (uint32_t)min(128, Cin - c * 128), 1, 1);  // MN-major both
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma<bf16>(tmem + (uint32_t)c * 128u, sdesc_advance(dtdesc0, kk * 2048), sdesc_advance(ad, kk * 2048),
                       idesc_w, (uint32_t)((i | kk) != 0));
          tc_commit(&tail->a_free[slot]);
          if (++slot == (uint32_t)L.a_slots) {
            slot = 0;
            sph ^= 1u;
          }
        }
        tc_commit(&tail->dt_free[db]);
      }
      tc_commit(&tail->done);
      CUNET_TRACE_MARK(trace, 232);
    }
  } else if (is_tr) {
    // ============================================================== transformers (256 threads)
    const bool gcol_ok = cc * 8 < p.dy.C;
    const bool gcol_used = cc * 8 < L.npad;
    const uint32_t dtb0 = smem_u32(smem + L.dt_off), ab = smem_u32(smem + L.a_off);
    // this thread's rows are the same in every stage: their half-resolution row / 2x2 position are registers, not a
    // dependent shared-memory load in front of every operand load
    int lowr[3], rpos[3];
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      const int r = min(rb + 24 * q, F1_R - 1);
      lowr[q] = tail->lowmap[r];
      rpos[q] = tail->rowpos[r];
    }
    GradCoef<bf16> gcf;
    {
      const int co2 = ((cc * 8) & 127) >> 1;
      gcf.a = *reinterpret_cast<const uint4*>(&tail->ga2[co2]);
      gcf.b = *reinterpret_cast<const uint4*>(&tail->gb2[co2]);
      gcf.mu = *reinterpret_cast<const uint4*>(&tail->gmu2[co2]);
      gcf.d = *reinterpret_cast<const uint4*>(&tail->gd2[co2]);
    }
    // 32-bit shared-space bases, opaque to the compiler (it otherwise rebuilds the aligned base and the thread index at
    // every use -- a dozen uniform-datapath instructions each time -- and reads the tables with generic loads)
    uint32_t tr_sbase = smem_u32(smem);
    uint32_t tr_coef = smem_u32(tail) + (uint32_t)offsetof(F1Tail, sc2) + (uint32_t)cc * 16u;   // + 256 B per chunk
    uint32_t tr_stoff = (uint32_t)(cc >> 3) * F1_SUB + tile_off(rb, cc & 7);   // rows rb + 24 q: + q * 24 * 128 bytes
    asm volatile("" : "+r"(tr_sbase), "+r"(tr_coef), "+r"(tr_stoff));
    const int relu_on = p.in.bn_train == 2 ? 0 : 1;
    ActCoef<bf16> acf_q;              // QuanInput constants (cunet_concat.act_bits), computed once
    acf_q.set_quant(p.in.act_bits);
    uint32_t slot = 0, sph = 0;       // ring position of the next activation-operand slot and its use parity
    for (int i = 0; i < ns; ++i) {
      const int st = st0 + i;
      const uint32_t db = L.dt_bufs == 2 ? ((uint32_t)i & 1u) : 0u;
      const uint32_t du = (uint32_t)(L.dt_bufs == 2 ? (i >> 1) : i);
      const uint32_t dtb = dtb0 + db * 2u * F1_SUB;
      const int nv = split ? F1_R : min(F1_R, M - st * F1_R);
      const uint32_t b = (uint32_t)i & 1u, upar = ((uint32_t)i >> 1) & 1u;
      const int gbuf = L.gt_bufs == 2 ? (i & 1) : 0;
      const uint32_t gu = (uint32_t)(L.gt_bufs == 2 ? (i >> 1) : i);
      const uint32_t posadd = (!split && W == 64) ? (uint32_t)((st & 1) << 1) : 0u;   // image row parity (H even)
      if (t == 0 && i < 12) CUNET_TRACE_MARK(trace, 32 + 4 * i);
      mbar_wait(&tail->gt_full[gbuf], gu & 1u);
      mbar_wait(&tail->dt_free[db], (du & 1u) ^ 1u);   // the MMAs that read this operand buffer before have completed
      if (t == 0 && i < 12) CUNET_TRACE_MARK(trace, 33 + 4 * i);
      if (gcol_used) {
        const uint32_t rg = smem_u32(smem + L.gt_off + gbuf * F1_GT_BYTES) + (uint32_t)cc * 16u;
        const uint32_t ri = smem_u32(smem + L.gt_off + gbuf * F1_GT_BYTES + 24576) + (uint32_t)cc * 8u;
#pragma unroll
        for (int q = 0; q < 3; ++q) {
          const int r = rb + 24 * q;
          if (r >= F1_R) break;
          uint4 o = make_uint4(0, 0, 0, 0), lo_unused;
          if (gcol_ok && r < nv) {
            GradRaw<bf16> raw;
            const int loc = p.dy.pooled ? lowr[q] : r;
            raw.g = f1_lds128(rg + (uint32_t)(loc * ldo));
            if (p.dy.mode == 1) raw.t = f1_lds128(rg + 16384u + (uint32_t)(loc * ldo));
            if (p.dy.pooled) {
              const uint2 iv = f1_lds64(ri + (uint32_t)(loc * p.dy.C));
              raw.idx[0] = iv.x;
              raw.idx[1] = iv.y;
              raw.pos = (uint32_t)rpos[q] + posadd;
            }
            o = gcf.apply(p.dy, raw, lo_unused);
          }
          sts128(dtb + (uint32_t)(cc >> 3) * F1_SUB + tile_off(r, cc & 7), o);
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&tail->dt_ready[db]);
        mbar_arrive(&tail->gt_free[gbuf]);
      }
      if (t == 0 && i < 12) CUNET_TRACE_MARK(trace, 34 + 4 * i);
      // ---- activation operand per chunk -> A slots
      mbar_wait(&tail->x_full[b], upar);
      const uint32_t xb = tr_sbase + (uint32_t)(L.x_off + (int)b * L.x_bytes);
      const bool v0 = rb < nv, v1 = rb + 24 < nv, v2 = rb + 48 < nv;   // this thread's three rows inside the stage
      // NOT unrolled over the chunks: unrolling tripled the code of this role (and of the epilogue below); the kernel
      // was 198 KB of SASS and its top stall reason in ncu was "no instruction" (instruction-cache misses).  The
      // per-chunk constants rotate through three registers instead of being indexed by a compile-time chunk number.
      // The loop is written for the fewest instructions per warp: this role executes ~10 clocks per instruction (ncu:
      // 183 SASS instructions and 1 us per chunk), every address is a 32-bit shared-space register + immediate, loads
      // are unconditional (an invalid row reads in-range garbage that is replaced by zero), nothing branches.
#pragma unroll 1
      for (int c = 0; c < nchunk; ++c) {
        ActCoef<bf16> acf;
        acf.sc = f1_lds128(tr_coef + (uint32_t)c * 256u);
        acf.sh = f1_lds128(tr_coef + (uint32_t)(offsetof(F1Tail, sh2) - offsetof(F1Tail, sc2)) + (uint32_t)c * 256u);
        acf.relu = relu_on;
        acf.qmax2 = acf_q.qmax2;
        acf.magic2 = acf_q.magic2;
        const uint32_t pk = cs_pk[0];
        {                                     // rotate: the next chunk's word moves to the front
          const uint32_t t0 = cs_pk[0];
          if (nchunk == 3) { cs_pk[0] = cs_pk[1]; cs_pk[1] = cs_pk[2]; cs_pk[2] = t0; }
          else if (nchunk == 2) { cs_pk[0] = cs_pk[1]; cs_pk[1] = t0; }
        }
        const bool cvalid = (pk & 0x80000000u) != 0u, cup = (pk & 0x40000000u) != 0u;
        const uint32_t rx = xb + (pk & 0xFFFFu);
        const uint32_t lsh = 6u + ((pk >> 16) & 3u);      // log2(bytes per source row)
        const uint4 raw0 = f1_lds128(rx + ((uint32_t)(cup ? lowr[0] : rb) << lsh));
        const uint4 raw1 = f1_lds128(rx + ((uint32_t)(cup ? lowr[1] : rb + 24) << lsh));
        const uint4 raw2 = f1_lds128(rx + ((uint32_t)(cup ? lowr[2] : min(rb + 48, F1_R - 1)) << lsh));
        mbar_wait(&tail->a_free[slot], sph ^ 1u);
        const uint32_t abase = tr_sbase + (uint32_t)L.a_off + slot * 16384u + tr_stoff;
        const uint4 zero4 = make_uint4(0, 0, 0, 0);
        uint4 lo_unused;
        const uint4 o0 = acf.apply(raw0, lo_unused), o1 = acf.apply(raw1, lo_unused), o2 = acf.apply(raw2, lo_unused);
        sts128(abase, (cvalid && v0) ? o0 : zero4);
        sts128(abase + 24u * 128u, (cvalid && v1) ? o1 : zero4);
        if (rb + 48 < F1_R) sts128(abase + 48u * 128u, (cvalid && v2) ? o2 : zero4);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&tail->a_full[slot]);
          mbar_arrive(&tail->a_done[b][c]);   // the epilogue may now overwrite this chunk's x with gamma*dz
        }
        if (++slot == (uint32_t)L.a_slots) {
          slot = 0;
          sph ^= 1u;
        }
      }
      if (t == 0 && i < 12) CUNET_TRACE_MARK(trace, 35 + 4 * i);
    }
  } else if (is_ep) {
    // ============================================================== epilogue (512 threads)
    float a_db[F1_MAXCH], a_dx[F1_MAXCH];   // per chunk: sum dz, sum dz*x of this thread's channel
#pragma unroll
    for (int c = 0; c < F1_MAXCH; ++c) a_db[c] = a_dx[c] = 0.f;
    // upsampled source: the two 8-column TMEM windows that hold the 16 children of this thread's 4 low pixels
    uint32_t upA = 0, upB = 0;
    if (W >= 32) {                 // split, or two whole rows of 32: rows 2j.. of both image rows
      upA = (uint32_t)(8 * pq);
      upB = (uint32_t)(32 + 8 * pq);
    } else if (W == 16) {
      upA = (uint32_t)(32 * (pq >> 1) + 8 * (pq & 1));
      upB = upA + 16u;
    } else {                       // W == 8: two rows of 8; W == 4: four rows of 4
      upA = (uint32_t)(16 * pq);
      upB = upA + 8u;
    }
    uint32_t buf = 0, bph = 0;       // ring position of the next D1 accumulator and its use parity
    for (int i = 0; i < ns; ++i) {
      const int nv = split ? F1_R : min(F1_R, M - (st0 + i) * F1_R);
      const uint32_t b = (uint32_t)i & 1u, upar = ((uint32_t)i >> 1) & 1u;
      const uint32_t xb = smem_u32(smem + L.x_off + b * L.x_bytes);
#pragma unroll 1
      for (int c = 0; c < nchunk; ++c) {     // not unrolled, accumulators rotate (see the transformers)
        mbar_wait(&tail->d1_full[buf], bph);
        mbar_wait(&tail->x_full[b], upar);        // completed long ago: acquires the landed x for this thread
        mbar_wait(&tail->a_done[b][c], upar);     // the transformers have read this chunk's x
        if (tid == 128 + 32 * F1_TRW && i < 12 && c == 0) CUNET_TRACE_MARK(trace, 144 + 4 * i);
        tc_fence_after();
        const float4 ec = tail->ech[c * 128 + k];
        const uint32_t ew = __float_as_uint(ec.w);
        if (ew & 0x20000000u) {
          const uint32_t Cp2 = 64u << ((ew >> 16) & 3u);
          const uint32_t xa = xb + (ew & 0xFFFFu);
          const float thr = ec.x, thr1 = ec.y, gm = ec.z;
          const bool neg = (ew & 0x80000000u) != 0u, is_up = (ew & 0x40000000u) != 0u;
          const uint32_t tb = tmem + d1_base + buf * 64u + ((uint32_t)(qd * 32) << 16);
          float db0 = 0.f, db1 = 0.f, dx0 = 0.f, dx1 = 0.f;   // sum dz, sum dz*x (xhat applied once at the end)
          float v[16];
          if (!is_up) {
            // this thread's 16 pixels: stage rows 16 pq .. 16 pq + 15
            const int r0 = 16 * pq;
            const int nvl = nv - r0;
            if (nvl > 0) {
              f1_tmem_ld16_nowait(tb + (uint32_t)r0, v);
              const uint32_t a0 = xa + (uint32_t)r0 * Cp2;
              if (nvl >= 16) {
                if (p.in.act_bits) {   // only convs behind a QuanInput2d: heads, one 128-channel source (host checks)
                  f1_ep16<256, true>(a0, v, thr, thr1, neg, gm, db0, db1, dx0, dx1);
                } else if (Cp2 == 256u) f1_ep16<256, false>(a0, v, thr, thr1, neg, gm, db0, db1, dx0, dx1);
                else if (Cp2 == 64u) f1_ep16<64, false>(a0, v, thr, thr1, neg, gm, db0, db1, dx0, dx1);
                else f1_ep16<128, false>(a0, v, thr, thr1, neg, gm, db0, db1, dx0, dx1);
              } else {
                float x[16];
#pragma unroll
                for (int q = 0; q < 16; ++q) x[q] = q < nvl ? f1_lds_bf16(a0 + (uint32_t)q * Cp2) : 0.f;
                f1_tmem_wait_ld();
#pragma unroll
                for (int q = 0; q < 16; ++q) {
                  const float dz = (q < nvl && ((x[q] > thr) != neg) && ((x[q] < thr1) != neg)) ? v[q] : 0.f;
                  db0 += dz;
                  dx0 = fmaf(dz, x[q], dx0);
                  if (q < nvl) f1_sts_u16(a0 + (uint32_t)q * Cp2, __bfloat16_as_ushort(__float2bfloat16_rn(gm * dz)));
                }
              }
            }
          } else {
            // upsampled source: this thread's 4 half-resolution pixels (rows 4 pq .. 4 pq + 3 of the stage's low run)
            // and their 16 children, gathered as a 2 x 8 image (W == 4: a 4 x 4 image)
            const int nlow = split ? 16 : (nv >> 2);
            const int l0 = 4 * pq;
            if (l0 < nlow) {
              f1_tmem_ld8_nowait(tb + upA, v);
              f1_tmem_ld8_nowait(tb + upB, v + 8);
              const uint32_t a0 = xa + (uint32_t)l0 * Cp2;
              float x[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) x[j] = (l0 + j < nlow) ? f1_lds_bf16(a0 + (uint32_t)j * Cp2) : 0.f;
              f1_tmem_wait_ld();
              float d4[4];
              if (W != 4) {
#pragma unroll
                for (int j = 0; j < 4; ++j) d4[j] = (v[2 * j] + v[2 * j + 1]) + (v[8 + 2 * j] + v[9 + 2 * j]);
              } else {
                d4[0] = (v[0] + v[1]) + (v[4] + v[5]);
                d4[1] = (v[2] + v[3]) + (v[6] + v[7]);
                d4[2] = (v[8] + v[9]) + (v[12] + v[13]);
                d4[3] = (v[10] + v[11]) + (v[14] + v[15]);
              }
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const bool ok = l0 + j < nlow;
                const float dz = (ok && ((x[j] > thr) != neg) && ((x[j] < thr1) != neg)) ? d4[j] : 0.f;
                db0 += dz;
                dx0 = fmaf(dz, x[j], dx0);
                if (ok) f1_sts_u16(a0 + (uint32_t)j * Cp2, __bfloat16_as_ushort(__float2bfloat16_rn(gm * dz)));
              }
            }
          }
          a_dx[0] += dx0 + dx1;
          a_db[0] += db0 + db1;
        }
        {                                     // rotate the per-chunk accumulators: the next chunk's pair moves to the front
          const float tb0 = a_db[0], tx0 = a_dx[0];
          if (nchunk == 3) {
            a_db[0] = a_db[1]; a_db[1] = a_db[2]; a_db[2] = tb0;
            a_dx[0] = a_dx[1]; a_dx[1] = a_dx[2]; a_dx[2] = tx0;
          } else if (nchunk == 2) {
            a_db[0] = a_db[1]; a_db[1] = tb0;
            a_dx[0] = a_dx[1]; a_dx[1] = tx0;
          }
        }
        fence_proxy_async();  // G written over x -> visible to the bulk store
        tc_fence_before();
        __syncwarp();
        if (tid == 128 + 32 * F1_TRW && i < 12 && c == 0) CUNET_TRACE_MARK(trace, 145 + 4 * i);
        if (lane == 0) {
          mbar_arrive(&tail->g_ready[b][c]);
          mbar_arrive(&tail->d1_free[buf]);
        }
        if (++buf == d1_nbuf) {
          buf = 0;
          bph ^= 1u;
        }
      }
      if (tid == 128 + 32 * F1_TRW && i < 12) CUNET_TRACE_MARK(trace, 146 + 4 * i);
    }
    if (ns > 0) {
      if (tid == 128 + 32 * F1_TRW) CUNET_TRACE_MARK(trace, 228);
      // ---- per-channel sums.  The four pixel quarters of a channel meet in shared memory (the landed G/T set is dead
      // once the last stage's D1 has been read), then ONE thread per concat channel adds to the global sums: the
      // (dbeta, dgamma, gstats) arrays are a few dozen cache lines, and 16 warps x 3 chunks x 4 atomics per CTA queued
      // on their L2 slices for 7-10 us per launch (in-kernel timeline, round 2) -- a quarter of that now.
      {
        // (placed behind the dW staging tiles of the warps that are already past this point: see below)
        float2* red = reinterpret_cast<float2*>(smem + max(L.gt_off, F1_DWSTG));
#pragma unroll
        for (int c = 0; c < F1_MAXCH; ++c)
          if (c < nchunk) red[pq * MAX_CIN + c * 128 + k] = make_float2(a_db[c], a_dx[c]);
        f1_named_bar(2, F1_THREADS - 128 - 32 * F1_TRW);
        const int kg = tid - (128 + 32 * F1_TRW);
        if (kg < Cin) {
          const float2 r0 = red[kg], r1 = red[MAX_CIN + kg], r2 = red[2 * MAX_CIN + kg], r3 = red[3 * MAX_CIN + kg];
          const float s_db = (r0.x + r1.x) + (r2.x + r3.x), s_dx = (r0.y + r1.y) + (r2.y + r3.y);
          int ps = 0;
          while (kg >= tail->seg_start[ps + 1]) ++ps;
          const int kl = kg - tail->seg_start[ps], Cp = p.in.seg[ps].C;
          // dgamma = sum dz*xhat = istd * (sum dz*x - mean * sum dz); mean / istd of the channel as compute_bn_coefs has them
          const cunet_seg& sg = p.in.seg[ps];
          const double mean = sg.stats[kl] * sg.inv_count;
          double var = sg.stats[Cp + kl] * sg.inv_count - mean * mean;
          if (var < 0.0) var = 0.0;
          const float is = (float)inv_sqrt_f64(var + (double)p.in.eps);
          const float dg = is * (s_dx - (float)mean * s_db);
          const float gm = tail->ech[kg].z;
          atomicAdd(p.dbeta + kg, s_db);
          atomicAdd(p.dgamma + kg, dg);
          if (p.gacc[ps].gstats) {
            // this consumer's share of (sum G, sum G*xhat) = gamma * (dbeta, dgamma)  (see conv_dgrad_v2.cu)
            atomicAdd(p.gacc[ps].gstats + kl, (double)(gm * s_db));
            atomicAdd(p.gacc[ps].gstats + Cp + kl, (double)(gm * dg));
          }
        }
      }
      // ---- weight gradient: D2_c[co][k] -> dW[co][k]: lane = output channel, 32 consecutive k per thread and chunk
      if (tid == 128 + 32 * F1_TRW) CUNET_TRACE_MARK(trace, 229);
      mbar_wait(&tail->done, 0);
      tc_fence_after();
      if (tid == 128 + 32 * F1_TRW) CUNET_TRACE_MARK(trace, 230);
      // Every warp owns a [32 co][32 k] block per chunk.  Out of TMEM a lane holds one co (row of dW) -- reductions issued
      // like that touch 32 half-used sectors per instruction, and the L2 retires reductions per sector (one-stage
      // launch: 5.2 us for 15 MB).  The block is turned in a private shared-memory tile (weights are dead after
      // `done`; 144-byte pitch: conflict-free both ways) so that 8 lanes cover one full 128-byte line of a dW row.
      const uint32_t stg = smem_u32(smem + L.w_off) + (uint32_t)(warp - 4 - F1_TRW) * (32u * 144u);   // 16 tiles = F1_DWSTG bytes
      const int co0 = qd * 32, rsub = lane >> 3, c16 = lane & 7;
      for (int c = 0; c < nchunk; ++c) {
        const int kvalid = min(min(128, Cin - c * 128), L.dwc - c * 128);   // real input channels of this chunk
        const int col0 = 32 * pq;
        if (col0 >= kvalid) continue;
        float v[32];
        f1_tmem_ld16_nowait(tmem + (uint32_t)c * 128u + ((uint32_t)co0 << 16) + (uint32_t)col0, v);
        f1_tmem_ld16_nowait(tmem + (uint32_t)c * 128u + ((uint32_t)co0 << 16) + (uint32_t)(col0 + 16), v + 16);
        f1_tmem_wait_ld();
#pragma unroll
        for (int q = 0; q < 32; q += 4)
          sts128(stg + (uint32_t)lane * 144u + (uint32_t)q * 4u,
                 make_uint4(__float_as_uint(v[q]), __float_as_uint(v[q + 1]), __float_as_uint(v[q + 2]), __float_as_uint(v[q + 3])));
        __syncwarp();
        const bool kok = col0 + c16 * 4 < kvalid;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          const int row = r * 4 + rsub;
          const uint4 u = f1_lds128(stg + (uint32_t)row * 144u + (uint32_t)c16 * 16u);
          if (kok && co0 + row < p.Cout)
            f1_red_add_v4(L.dw + (long)(co0 + row) * L.dwc + c * 128 + col0 + c16 * 4, __uint_as_float(u.x),
                          __uint_as_float(u.y), __uint_as_float(u.z), __uint_as_float(u.w));
        }
        __syncwarp();
      }
      if (tid == 128 + 32 * F1_TRW) CUNET_TRACE_MARK(trace, 231);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem, 512);
}

}  // namespace cunet
using namespace cunet;

CUNET_TRACE_SETTER(cunet_debug_trace_bwd1x1, g_f1_trace)

// 1: handled by the fused kernel; 0: not eligible (caller runs the two separate kernels); <0: error
static int conv_bwd1x1_try(const cunet_conv_dgrad_params* d, const cunet_conv_wgrad_params* w, cudaStream_t st) {
  static const bool off = getenv("CUNET_BWD1X1_OFF") != nullptr;
  if (off) return 0;
  if (d->dtype != CUNET_BF16 || w->dtype != CUNET_BF16 || d->taps != 1 || w->taps != 1) return 0;
  if (d->in.bn_train != 1) return 0;
  if (d->dy.ld != d->dy.C || d->dy.C > 128 || (d->dy.C & 7) || d->CoutPad > 128 || d->CoutPad != d->dy.C) return 0;
  if (w->Cout > 128 || w->Cout != d->Cout) return 0;
  // same op on both sides
  if (w->in.nseg != d->in.nseg || w->dy.g != d->dy.g || w->dy.t != d->dy.t || w->N != d->N || w->H != d->H ||
      w->W != d->W || w->dy.mode != d->dy.mode || w->dy.pooled != d->dy.pooled)
    return 0;
  int cin = 0, up = 0, xbytes = 0;
  const int W = d->W, H = d->H;
  const int split_geo = (W == 64);
  for (int s = 0; s < d->in.nseg; ++s) {
    const cunet_seg& sg = d->in.seg[s];
    if (sg.ld != sg.C || (sg.C != 32 && sg.C != 64 && sg.C != 128)) return 0;
    if ((cin >> 7) != ((cin + sg.C - 1) >> 7)) return 0;   // a segment must not straddle a 128-channel chunk
    if (w->in.seg[s].ptr != sg.ptr || w->in.seg[s].up != sg.up) return 0;
    if (d->gacc[s].G && d->gacc[s].ld != sg.C) return 0;
    cin += sg.C;
    up |= sg.up;
  }
  if (cin > MAX_CIN || cin > 128 * F1_MAXCH) return 0;
  if ((W & (W - 1)) || (H & (H - 1)) || W > 64 || W < 4 || H < 4) return 0;
  if (up && d->dy.pooled) return 0;
  if (d->in.act_bits && (d->in.nseg != 1 || d->in.seg[0].C != 128)) return 0;   // quantized operand: heads only
  const int split = (up && split_geo) ? 1 : 0;
  const long M = (long)d->N * H * W;
  if (M <= 0) return 1;
  if (M > (1L << 30)) return 0;
  if ((up || d->dy.pooled) && (M % 64) && ((M % 64) % (2 * W))) return 0;
  if (w->dw_cin != 0 && w->dw_cin < cin) return 0;
  // piece sizes inside an x buffer (must match the kernel's xoff table): 64 rows, or the <= 16 low rows of an
  // upsampled source
  for (int s = 0; s < d->in.nseg; ++s) xbytes += (d->in.seg[s].up ? 16 : F1_R) * d->in.seg[s].C * 2;
  const int nkb = (d->CoutPad + 63) / 64;
  F1Layout L;
  L.w_off = 0;
  const int wbytes = cin * 128 * nkb;
  // the last chunk's weight rows are read as a full 128-row operand: keep 16 KB of readable slack behind them
  L.x_bytes = (xbytes + 1023) & ~1023;
  if (L.x_bytes < 16384) L.x_bytes = 16384;   // the coefficient tables live there during the prologue
  const int limit = 232448 - 1024 - (int)sizeof(F1Tail);
  // spare shared memory buys decoupling, in this order: an activation-operand slot per chunk (the transform of chunk c
  // then waits for the wgrad MMAs of the PREVIOUS stage, not of two items ago), a second gradient-operand buffer (dT of
  // stage i+1 is built while the MMAs of stage i still read theirs), a second G/T landing set
  const int nchunk_h = (cin + 127) / 128;
  const int base = ((wbytes + 1023) & ~1023) + 2 * L.x_bytes;
  L.a_slots = 2;
  L.dt_bufs = 1;
  L.gt_bufs = 1;
  auto need = [&]() { return base + L.dt_bufs * 2 * F1_SUB + L.a_slots * 16384 + L.gt_bufs * F1_GT_BYTES; };
  if (need() > limit) return 0;
  if (nchunk_h == 3) {
    L.a_slots = 3;
    if (need() > limit) L.a_slots = 2;
  }
  L.dt_bufs = 2;
  if (need() > limit) L.dt_bufs = 1;
  L.gt_bufs = 2;
  if (need() > limit) L.gt_bufs = 1;
  L.dt_off = (wbytes + 1023) & ~1023;
  L.a_off = L.dt_off + L.dt_bufs * 2 * F1_SUB;
  L.gt_off = L.a_off + L.a_slots * 16384;
  L.x_off = L.gt_off + L.gt_bufs * F1_GT_BYTES;
  L.tail_off = L.x_off + 2 * L.x_bytes;
  if (L.tail_off > limit) return 0;
  L.split = split;
  L.npad = ((d->dy.C + 63) / 64) * 64;
  L.dwc = w->dw_cin > 0 ? w->dw_cin : cin;
  L.dw = w->dw;
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  const int total = (int)((M + F1_R - 1) / F1_R);
  L.per = (total + sms - 1) / sms;
  const int grid = (total + L.per - 1) / L.per;
  const size_t smem = (size_t)L.tail_off + sizeof(F1Tail) + 1024;
  cudaError_t e = cudaFuncSetAttribute(conv_bwd1x1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
  if (e != cudaSuccess) return cunet_fail_cuda("conv_bwd1x1 attr", e);
  e = cunet_launch(conv_bwd1x1_kernel, dim3(grid), dim3(F1_THREADS), smem, st, *d, L);
  if (e != cudaSuccess) return cunet_fail_cuda("conv_bwd1x1 launch", e);
  return 1;
}

extern "C" int cunet_conv_bwd1x1(const cunet_conv_dgrad_params* d, const cunet_conv_wgrad_params* w, void* stream) {
  if (!d || !w) return cunet_fail("conv_bwd1x1: null params");
  const int r = conv_bwd1x1_try(d, w, reinterpret_cast<cudaStream_t>(stream));
  if (r != 0) return r < 0 ? r : 0;
  const int rc = cunet_conv_dgrad(d, stream);
  if (rc != 0) return rc;
  return cunet_conv_wgrad(w, stream);
}
