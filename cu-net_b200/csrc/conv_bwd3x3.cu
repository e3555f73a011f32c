// Fused backward of the dense-layer 3x3 conv (models/cu_net.py:47-48 conv2, bf16): backward-data AND
// backward-filter in ONE persistent kernel.  Same math as cunet_conv_dgrad + cunet_conv_wgrad (taps == 9):
//
//   dgrad : dA[p][ci]        = sum_{tap,co} dT[p - off(tap)][co] * W[co][ci][tap]
//   wgrad : dW[co][ci][tap] += sum_p        a[p][ci]           * dT[p - off(tap)][co],    a = relu(bn(x))
//
// Both contractions consume the SAME operand: the im2col of the output gradient, I[p][(tap, co)] = dT[p - off(tap)][co]
// (zero outside the image).  The round-1 kernels built it twice, each time with nine scattered global gathers per
// pixel and no prefetch (ncu, batch 24: dgrad 139 us + wgrad 124 us at 64x64, 50 + 42 us even for 4x4 images).
// Here, per stage of 64 consecutive pixels:
//   * G/T of the output for the 64 pixels + a one-row halo (W + 1 pixels each side) and the raw input x are
//     CONTIGUOUS blocks of NHWC tensors: three 1-D TMA bulk copies land them in shared memory;
//   * 8 transformer warps evaluate dT = a*G + b*(T - mu) + d (batch-norm backward form) while gathering smem -> smem
//     into the swizzled operand image I[64 px][320] (5 sub-tiles of 64 rows x 128 B) and a = relu(bn(x)) into
//     A[64 px][128] (2 sub-tiles);
//   * one thread issues  D1[128 ci][64 px]     = Wimg[128 ci][320] * I^T        (K-major x K-major, fresh per stage)
//                        D2[128 ci][320 (tap,co)] += A^T[128 ci][64 px] * I     (MN-major x MN-major, TMEM-resident
//                                                                                for the CTA's whole pixel range);
//   * 8 epilogue warps (thread = input channel = TMEM lane) apply the ReLU mask, accumulate dbeta / dgamma in
//     registers and overwrite x IN PLACE with gamma*dz; one thread bulk-stores (or L2-reduce-adds) the block to G_x;
//   * at the end D2 is transposed through shared memory and added to the fp32 weight gradient with coalesced
//     red.global.add (reference layout [co][ci][tap]).
// HBM traffic per stage = G, T (8 KB) + x (16 KB) + G_x (16 KB): everything is read once and written once.
#include "loaders.cuh"
#include "host_util.h"
#include <stdlib.h>

namespace cunet {

constexpr int B3_THREADS = 640;  // warp 0 landing producer | 1 store issuer | 2 MMA | 3 idle | 4-11 transformers | 12-19 epilogue
constexpr int B3_R = 64;                            // pixels per stage
constexpr int B3_SUB = B3_R * 128;                  // one sub-tile: 64 rows x 128 B
constexpr int B3_W_BYTES = 5 * 16384;               // dgrad weight image: 5 K blocks x [128 rows][128 B]
constexpr int B3_W_OFF = 0;
constexpr int B3_B_OFF = B3_W_OFF + B3_W_BYTES;     // im2col operand: 5 sub-tiles
constexpr int B3_A_OFF = B3_B_OFF + 5 * B3_SUB;     // activation operand: 2 sub-tiles
constexpr int B3_X_OFF = B3_A_OFF + 2 * B3_SUB;     // raw x / outgoing G, double buffered: 2 x 16 KB
constexpr int B3_GT_BYTES = 12544;                  // >= (64 + 2*(64 + 1)) rows x 64 B
constexpr int B3_G_OFF = B3_X_OFF + 2 * 16384;      // landed G / T with halo, double buffered: [buf][G | T]
constexpr int B3_TAIL_OFF = B3_G_OFF + 4 * B3_GT_BYTES;
constexpr int B3_COEF_OFF = B3_A_OFF;               // BnSmem + GradSmem live in the activation-operand area during the prologue only
constexpr int B3_STG_FLOATS = 32 * 128 * 9;         // dW staging in destination order [32 co][128 ci][9 taps]
constexpr uint32_t B3_D1_COL = 320;                 // TMEM: D2 in columns [0, 320), D1 buffers at 320 and 384
static_assert(sizeof(BnSmem) <= 8192 && 8192 + sizeof(GradSmem) <= 2 * B3_SUB, "coefficient overlay");
static_assert(B3_STG_FLOATS * 4 <= B3_TAIL_OFF, "dW staging must fit the (dead) weight / operand / landing regions");

struct B3Tail {
  uint64_t w_full, ops_ready, ops_free, done, stores_done;
  uint64_t gt_full[2], gt_free[2], x_full[2], x_free[2], d1_full[2], d1_free[2], g_ready[2];
  uint32_t tmem_base;
  float2 csum[128];          // final epilogue: (dbeta, dgamma) partial of the second pixel half
};

__device__ __forceinline__ void b3_bulk_s2g(void* dst, const void* src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src_smem)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void b3_bulk_red_add_bf16(void* dst, const void* src_smem, uint32_t bytes) {
  asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.noftz.bf16 [%0], [%1], %2;" ::"l"(dst),
               "r"(smem_u32(src_smem)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void b3_bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void b3_bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ float b3_lds_bf16(uint32_t saddr) {
  uint16_t v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(saddr));
  return __uint_as_float((uint32_t)v << 16);
}
__device__ __forceinline__ void b3_sts_u16(uint32_t saddr, uint16_t v) {
  asm volatile("st.shared.u16 [%0], %1;" ::"r"(saddr), "h"(v) : "memory");
}
__device__ __forceinline__ uint4 b3_lds128(uint32_t saddr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr));
  return v;
}

// timeline slots (CUNET_TRACE builds): stage i < 16 -> producer 0+2i, transformer 32+4i, MMA 96+3i, epilogue 144+2i,
// store issuer 176+2i; 208: dW staging starts, 209: staged, 210: added to the gradient
CUNET_TRACE_DECL(g_b3_trace)

__device__ __forceinline__ void b3_named_bar(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

__global__ void __launch_bounds__(B3_THREADS, 1) conv_bwd3x3_kernel(const __grid_constant__ cunet_conv_dgrad_params p,
                                                                     float* __restrict__ dw, int per) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  B3Tail* tail = reinterpret_cast<B3Tail*>(smem + B3_TAIL_OFF);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int H = p.H, W = p.W;
  const int M = p.N * H * W;
  const int total = (M + B3_R - 1) / B3_R;
  const int st0 = (int)blockIdx.x * per;
  const int st1 = min(total, st0 + per);
  const int ns = max(0, st1 - st0);
  const int halo = W + 1;
  CUNET_TRACE_LOAD(trace, g_b3_trace)

  if (tid == 0) {
    mbar_init(&tail->w_full, 1);
    mbar_init(&tail->ops_ready, 8);
    mbar_init(&tail->ops_free, 1);
    mbar_init(&tail->done, 1);
    mbar_init(&tail->stores_done, 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tail->gt_full[b], 1);
      mbar_init(&tail->gt_free[b], 8);
      mbar_init(&tail->x_full[b], 1);
      mbar_init(&tail->x_free[b], 1);
      mbar_init(&tail->d1_full[b], 1);
      mbar_init(&tail->d1_free[b], 8);
      mbar_init(&tail->g_ready[b], 8);
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(&tail->tmem_base, 512);
  {
    // the operand image is zeroed once: its padding columns (K index >= 288) are never written again
    const uint32_t b0 = smem_u32(smem + B3_B_OFF);
    for (int i = tid; i < 5 * B3_SUB / 16; i += B3_THREADS) sts128(b0 + (uint32_t)i * 16u, make_uint4(0, 0, 0, 0));
    fence_proxy_async();
  }
  __syncthreads();   // barrier inits / zero fill visible to every warp; still under the previous kernel
  griddep_wait();    // everything above overlaps the tail of the previous kernel (programmatic dependent launch)
  griddep_launch();
  // warp 0 (landing producer) needs no coefficient: it starts its copies at once; the other warps compute them in the
  // activation-operand area (not in the x landing buffers) behind named barriers that do not include warp 0
  BnSmem* bn = reinterpret_cast<BnSmem*>(smem + B3_COEF_OFF);
  GradSmem* gc = reinterpret_cast<GradSmem*>(smem + B3_COEF_OFF + 8192);
  if (warp > 0) {
    compute_bn_coefs(p.in, bn, 128, tid - 32, B3_THREADS - 32);
    compute_grad_coefs(p.dy, gc, tid - 32, B3_THREADS - 32);
    tc_fence_before();
    b3_named_bar(1, B3_THREADS - 32);
    tc_fence_after();
  }
  const uint32_t tmem = warp > 0 ? tail->tmem_base : 0u;

  // per-role coefficients go to registers; after the next barrier the coefficient area is the activation operand
  const bool is_tr = warp >= 4 && warp < 12, is_ep = warp >= 12;
  const int t = tid - 128;                 // transformer thread index (0..255)
  const int c4 = t & 3, rI = t >> 2;       // im2col: 16-byte column (8 output channels) x pixel row rI
  const int cc = t & 15, rb = t >> 4;      // activation: 16-byte column (8 input channels) x rows rb + 16q
  const int e = warp - 12;
  const int qd = warp & 3, hf = (e >> 2) & 1;  // epilogue: TMEM lane quarter (hardware: warp % 4), pixel-column half
  const int k = qd * 32 + lane;                // epilogue: input channel
  GradCoef<bf16> gcf;
  ActCoef<bf16> acf;
  float sc = 0.f, sh = 0.f, is = 0.f, nmi = 0.f;
  if (is_tr) {
    gcf.load(gc, c4 * 8);
    acf.load(bn, cc * 8);
  }
  if (is_ep) {
    sc = bn->scale[k];
    sh = bn->shift[k];
    is = bn->istd[k];
    nmi = -bn->mean[k] * is;  // xhat = x * istd - mean * istd
  }
  if (warp > 0) b3_named_bar(1, B3_THREADS - 32);   // the coefficient area becomes the activation operand again

  if (warp == 0) {
    // ============================================================== landing producer
    if (lane == 0 && ns > 0) {
      mbar_arrive_expect_tx(&tail->w_full, (uint32_t)B3_W_BYTES);
      bulk_g2s(smem + B3_W_OFF, p.wpack_dgrad, (uint32_t)B3_W_BYTES, &tail->w_full);
      const char* xsrc = reinterpret_cast<const char*>(p.in.seg[0].ptr);
      const char* gsrc = reinterpret_cast<const char*>(p.dy.g);
      const char* tsrc = reinterpret_cast<const char*>(p.dy.t);
      for (int i = 0; i < ns; ++i) {
        const int m0 = (st0 + i) * B3_R, nv = min(B3_R, M - m0);
        const uint32_t b = (uint32_t)i & 1u, fpar = (((uint32_t)i >> 1) & 1u) ^ 1u;
        const int lo = max(0, m0 - halo), hi = min(M, m0 + nv + halo);
        const uint32_t gb = (uint32_t)((hi - lo) * 64);
        mbar_wait(&tail->gt_free[b], fpar);
        if (i < 16) CUNET_TRACE_MARK(trace, 0 + 2 * i);
        mbar_arrive_expect_tx(&tail->gt_full[b], 2u * gb);
        bulk_g2s(smem + B3_G_OFF + b * 2 * B3_GT_BYTES, gsrc + (long)lo * 64, gb, &tail->gt_full[b]);
        bulk_g2s(smem + B3_G_OFF + b * 2 * B3_GT_BYTES + B3_GT_BYTES, tsrc + (long)lo * 64, gb, &tail->gt_full[b]);
        mbar_wait(&tail->x_free[b], fpar);
        if (i < 16) CUNET_TRACE_MARK(trace, 1 + 2 * i);
        mbar_arrive_expect_tx(&tail->x_full[b], (uint32_t)(nv * 256));
        bulk_g2s(smem + B3_X_OFF + b * 16384, xsrc + (long)m0 * 256, (uint32_t)(nv * 256), &tail->x_full[b]);
      }
    }
  } else if (warp == 1) {
    // ============================================================== G store issuer
    if (lane == 0) {
      char* G = reinterpret_cast<char*>(p.gacc[0].G);
      for (int i = 0; i < ns; ++i) {
        const int m0 = (st0 + i) * B3_R, nv = min(B3_R, M - m0);
        const uint32_t b = (uint32_t)i & 1u;
        mbar_wait(&tail->g_ready[b], ((uint32_t)i >> 1) & 1u);
        if (i < 16) CUNET_TRACE_MARK(trace, 176 + 2 * i);
        const uint8_t* src = smem + B3_X_OFF + b * 16384;
        if (p.gacc[0].accumulate) b3_bulk_red_add_bf16(G + (long)m0 * 256, src, (uint32_t)(nv * 256));
        else b3_bulk_s2g(G + (long)m0 * 256, src, (uint32_t)(nv * 256));
        b3_bulk_commit();
        b3_bulk_wait_read0();
        if (i < 16) CUNET_TRACE_MARK(trace, 177 + 2 * i);
        mbar_arrive(&tail->x_free[b]);
      }
      mbar_arrive(&tail->stores_done);   // every G store has read its staging block: the dW epilogue may reuse the area
    }
  } else if (warp == 2) {
    // ============================================================== MMA issuer
    if (lane == 0 && ns > 0) {
      const uint32_t idesc_d = make_idesc(Elem<bf16>::FMT, 128, 64, 0, 0);
      const uint32_t idesc_w1 = make_idesc(Elem<bf16>::FMT, 128, 192, 1, 1);
      const uint32_t idesc_w2 = make_idesc(Elem<bf16>::FMT, 128, 128, 1, 1);
      const uint32_t wA = smem_u32(smem + B3_W_OFF), bB = smem_u32(smem + B3_B_OFF), aA = smem_u32(smem + B3_A_OFF);
      // operand descriptors are advanced from these bases (sdesc_advance, common.cuh): one thread issues the 26 MMAs of a
      // stage, and rebuilding every descriptor from scratch was a measurable part of its time
      const uint64_t wdesc0 = make_sdesc(wA, 16, 1024), bdesc0 = make_sdesc(bB, 16, 1024);
      const uint64_t amn0 = make_sdesc_mn<bf16>(aA, B3_SUB), bmn0 = make_sdesc_mn<bf16>(bB, B3_SUB);
      mbar_wait(&tail->w_full, 0);
      for (int i = 0; i < ns; ++i) {
        const uint32_t b = (uint32_t)i & 1u;
        mbar_wait(&tail->ops_ready, (uint32_t)i & 1u);
        if (i < 16) CUNET_TRACE_MARK(trace, 96 + 3 * i);
        mbar_wait(&tail->d1_free[b], (((uint32_t)i >> 1) & 1u) ^ 1u);
        if (i < 16) CUNET_TRACE_MARK(trace, 97 + 3 * i);
        tc_fence_after();
        const uint32_t d1 = tmem + B3_D1_COL + b * 64u;
#pragma unroll
        for (int kb = 0; kb < 5; ++kb) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            if (kb == 4 && kk >= 2) break;  // K index >= 288: padding
            umma<bf16>(d1, sdesc_advance(wdesc0, kb * 16384 + kk * 32), sdesc_advance(bdesc0, kb * B3_SUB + kk * 32),
                       idesc_d, (uint32_t)((kb | kk) != 0));
          }
        }
        tc_commit(&tail->d1_full[b]);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const uint32_t acc = (uint32_t)((i | kk) != 0);
          const uint64_t ad = sdesc_advance(amn0, kk * 2048);
          umma<bf16>(tmem, ad, sdesc_advance(bmn0, kk * 2048), idesc_w1, acc);
          umma<bf16>(tmem + 192, ad, sdesc_advance(bmn0, 3 * B3_SUB + kk * 2048), idesc_w2, acc);
        }
        tc_commit(&tail->ops_free);
        if (i < 16) CUNET_TRACE_MARK(trace, 98 + 3 * i);
      }
      tc_commit(&tail->done);
    }
  } else if (is_tr) {
    // ============================================================== transformers (256 threads)
    const uint32_t bbase = smem_u32(smem + B3_B_OFF), abase = smem_u32(smem + B3_A_OFF);
    for (int i = 0; i < ns; ++i) {
      const int m0 = (st0 + i) * B3_R, nv = min(B3_R, M - m0);
      const int lo = max(0, m0 - halo);
      const uint32_t b = (uint32_t)i & 1u, upar = ((uint32_t)i >> 1) & 1u;
      const uint32_t gl = smem_u32(smem + B3_G_OFF + b * 2 * B3_GT_BYTES) + (uint32_t)c4 * 16u, tl = gl + B3_GT_BYTES;
      if (t == 0 && i < 16) CUNET_TRACE_MARK(trace, 32 + 4 * i);
      mbar_wait(&tail->gt_full[b], upar);
      mbar_wait(&tail->x_full[b], upar);
      if (t == 0 && i < 16) CUNET_TRACE_MARK(trace, 33 + 4 * i);
      mbar_wait(&tail->ops_free, ((uint32_t)i & 1u) ^ 1u);  // MMAs of the previous stage no longer read the operands
      if (t == 0 && i < 16) CUNET_TRACE_MARK(trace, 34 + 4 * i);
      {
        const int r = rI;
        const int m = m0 + r;
        int n = 0, h = 0, w = 0;
        const bool rv = r < nv;
        if (rv) pix_split(m, H, W, n, h, w);
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
          const int dy = tap / 3 - 1, dx = tap % 3 - 1;
          // out(q) reads in(q + off)  =>  in(p) receives from out(p - off)
          const int hs = h - dy, ws = w - dx;
          uint4 o = make_uint4(0, 0, 0, 0), lo_unused;
          if (rv && (unsigned)hs < (unsigned)H && (unsigned)ws < (unsigned)W) {
            const uint32_t loc = (uint32_t)(m - dy * W - dx - lo) * 64u;
            GradRaw<bf16> raw;
            raw.g = b3_lds128(gl + loc);
            raw.t = b3_lds128(tl + loc);
            o = gcf.apply(p.dy, raw, lo_unused);
          }
          sts128(bbase + (uint32_t)(tap >> 1) * B3_SUB + tile_off(r, ((tap & 1) << 2) | c4), o);
        }
      }
      {
        const uint32_t xb = smem_u32(smem + B3_X_OFF + b * 16384) + (uint32_t)cc * 16u;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int r = rb + 16 * q;
          uint4 o = make_uint4(0, 0, 0, 0), lo_unused;
          if (r < nv) o = acf.apply(b3_lds128(xb + (uint32_t)r * 256u), lo_unused);
          sts128(abase + (uint32_t)(cc >> 3) * B3_SUB + tile_off(r, cc & 7), o);
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (t == 0 && i < 16) CUNET_TRACE_MARK(trace, 35 + 4 * i);
      if (lane == 0) {
        mbar_arrive(&tail->ops_ready);
        mbar_arrive(&tail->gt_free[b]);
      }
    }
  } else if (is_ep) {
    // ============================================================== epilogue (256 threads)
    const float gm = p.in.gamma[k];
    // QuanInput between the ReLU and the conv (cunet_concat.act_bits): straight-through, zero where the activation >= 1
    const float zmax = p.in.act_bits ? 1.f : __int_as_float(0x7f800000);
    float a_db = 0.f, a_dg = 0.f;
    for (int i = 0; i < ns; ++i) {
      const int m0 = (st0 + i) * B3_R, nv = min(B3_R, M - m0);
      const uint32_t b = (uint32_t)i & 1u;
      mbar_wait(&tail->d1_full[b], ((uint32_t)i >> 1) & 1u);
      mbar_wait(&tail->x_full[b], ((uint32_t)i >> 1) & 1u);  // completed long ago: acquires the landed x for this thread
      if (tid == 384 && i < 16) CUNET_TRACE_MARK(trace, 144 + 2 * i);
      tc_fence_after();
      const uint32_t xa = smem_u32(smem + B3_X_OFF + b * 16384) + (uint32_t)k * 2u;
      const uint32_t tb = tmem + B3_D1_COL + b * 64u + ((uint32_t)(qd * 32) << 16);
#pragma unroll 1
      for (int g8 = 0; g8 < 4; ++g8) {
        const int col0 = hf * 32 + g8 * 8;
        const int nval = nv - col0;
        if (nval <= 0) break;
        float v[8];
        tmem_ld8(tb + (uint32_t)col0, v);
        const uint32_t a0 = xa + (uint32_t)col0 * 256u;
        float x[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) x[q] = q < nval ? b3_lds_bf16(a0 + (uint32_t)q * 256u) : 0.f;
        uint16_t gb[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float z = fmaf(x[q], sc, sh);
          const bool on = z > 0.f && z < zmax && q < nval;
          const float dz = on ? v[q] : 0.f;
          a_db += dz;
          a_dg = fmaf(dz, fmaf(x[q], is, nmi), a_dg);
          gb[q] = __bfloat16_as_ushort(__float2bfloat16_rn(gm * dz));
        }
#pragma unroll
        for (int q = 0; q < 8; ++q)
          if (q < nval) b3_sts_u16(a0 + (uint32_t)q * 256u, gb[q]);
      }
      fence_proxy_async();  // G written over x -> visible to the bulk store
      tc_fence_before();
      __syncwarp();
      if (tid == 384 && i < 16) CUNET_TRACE_MARK(trace, 145 + 2 * i);
      if (lane == 0) {
        mbar_arrive(&tail->g_ready[b]);
        mbar_arrive(&tail->d1_free[b]);
      }
    }
    {
      // the two pixel halves of a channel meet in shared memory: one thread per channel and CTA adds to the global sums
      // (same-line atomics queue on a few L2 slices -- conv_bwd1x1.cu's tail lost 9 us per launch to that)
      if (hf) tail->csum[k] = make_float2(a_db, a_dg);
      b3_named_bar(2, 256);
      if (!hf && ns > 0) {
        const float2 o = tail->csum[k];
        a_db += o.x;
        a_dg += o.y;
        atomicAdd(p.dbeta + k, a_db);
        atomicAdd(p.dgamma + k, a_dg);
        if (p.gacc[0].gstats) {
          // this consumer's share of (sum G, sum G*xhat) = gamma * (dbeta, dgamma)  (see conv_dgrad_v2.cu)
          atomicAdd(p.gacc[0].gstats + k, (double)(gm * a_db));
          atomicAdd(p.gacc[0].gstats + 128 + k, (double)(gm * a_dg));
        }
      }
    }
    // ---- weight gradient: D2[128 ci][(tap, co)] -> dW[co][ci][tap].  A thread owns one ci (TMEM lane), but the
    // reference layout has ci * 9 + tap contiguous per co: adding straight from registers put 32 different sectors
    // under every warp-wide red (ncu, round 1: ~30 us of the kernel, even for a 6-CTA launch).  The accumulator is
    // transposed through the (now dead) weight / operand / landing regions in ONE pass over all 32 output channels
    // (round 2: wide tcgen05.ld, 16-byte reductions, one barrier instead of four: the epilogue of a one-stage CTA took
    // 6.8 of its ~11 us).  All CTAs run this (uniform barriers); ns == 0 CTAs do not exist (host: grid).
    float* stg = reinterpret_cast<float*>(smem);
    if (ns > 0) {
      mbar_wait(&tail->done, 0);
      tc_fence_after();
    }
    mbar_wait(&tail->stores_done, 0);
    if (tid == 384) CUNET_TRACE_MARK(trace, 208);
    if (ns > 0) {
      // this thread's taps: hf == 0 -> 0..4, hf == 1 -> 5..8; per tap 32 output channels = two 16-column loads
      const int t0 = hf ? 5 : 0, t1 = hf ? 9 : 5;
      for (int tap = t0; tap < t1; ++tap) {
        float v[32];
        const uint32_t ta = tmem + ((uint32_t)(qd * 32) << 16) + (uint32_t)(tap * 32);
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
            : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]), "=f"(v[8]),
              "=f"(v[9]), "=f"(v[10]), "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15])
            : "r"(ta));
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
            : "=f"(v[16]), "=f"(v[17]), "=f"(v[18]), "=f"(v[19]), "=f"(v[20]), "=f"(v[21]), "=f"(v[22]), "=f"(v[23]),
              "=f"(v[24]), "=f"(v[25]), "=f"(v[26]), "=f"(v[27]), "=f"(v[28]), "=f"(v[29]), "=f"(v[30]), "=f"(v[31])
            : "r"(ta + 16u));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int q = 0; q < 32; ++q) stg[q * 1152 + k * 9 + tap] = v[q];   // destination order [co][ci][tap]; lanes 9 words apart
      }
    }
    if (tid == 384) CUNET_TRACE_MARK(trace, 209);
  }

  // every role is done and the staged dW is complete: ALL 640 threads add it to the fp32 gradient (the add loop is
  // instruction-bound -- 36864 floats per CTA -- and ran 7 us on the 8 epilogue warps alone)
  tc_fence_before();
  __syncthreads();
  if (ns > 0) {
    const float* stg = reinterpret_cast<const float*>(smem);
    // the staged block is in destination order: 16-byte loads, 16-byte reductions (the first version of this loop rebuilt
    // (co, ci, tap) from the index with two integer divisions per element and took 6 us per CTA)
    for (int idx = tid * 4; idx < 32 * 1152; idx += B3_THREADS * 4) {
      const float4 o = *reinterpret_cast<const float4*>(stg + idx);
      asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(dw + idx), "f"(o.x), "f"(o.y), "f"(o.z), "f"(o.w)
                   : "memory");
    }
  }
  if (tid == 384) CUNET_TRACE_MARK(trace, 210);
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem, 512);
}

}  // namespace cunet
using namespace cunet;

CUNET_TRACE_SETTER(cunet_debug_trace_bwd3x3, g_b3_trace)

// 1: handled by the fused kernel; 0: not eligible (caller runs the two generic kernels); <0: error
static int conv_bwd3x3_try(const cunet_conv_dgrad_params* d, const cunet_conv_wgrad_params* w, cudaStream_t st) {
  static const bool off = getenv("CUNET_BWD3X3_OFF") != nullptr;
  if (off) return 0;
  if (d->dtype != CUNET_BF16 || w->dtype != CUNET_BF16 || d->taps != 9 || w->taps != 9) return 0;
  if (d->Cout != 32 || d->CoutPad != 32 || w->Cout != 32) return 0;
  if (d->dy.C != 32 || d->dy.ld != 32 || d->dy.mode != 1 || d->dy.pooled) return 0;
  if (d->in.nseg != 1 || d->in.bn_train != 1) return 0;
  const cunet_seg& sg = d->in.seg[0];
  if (sg.C != 128 || sg.ld != 128 || sg.up) return 0;
  if (!d->gacc[0].G || d->gacc[0].ld != 128) return 0;
  if (w->dw_cin != 0 && w->dw_cin != 128) return 0;
  if (d->W > 64 || d->W < 1 || d->H < 1) return 0;
  // same op on both sides
  if (w->in.seg[0].ptr != sg.ptr || w->dy.g != d->dy.g || w->dy.t != d->dy.t || w->N != d->N || w->H != d->H ||
      w->W != d->W)
    return 0;
  const long M = (long)d->N * d->H * d->W;
  if (M <= 0) return 1;
  if (M > (1L << 30)) return 0;
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  const int total = (int)((M + B3_R - 1) / B3_R);
  const int per = (total + sms - 1) / sms;
  const int grid = (total + per - 1) / per;
  const size_t smem = B3_TAIL_OFF + sizeof(B3Tail) + 1024;
  static_assert(B3_TAIL_OFF + sizeof(B3Tail) + 1024 <= 232448, "conv_bwd3x3 shared memory");
  cudaError_t e = cudaFuncSetAttribute(conv_bwd3x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return cunet_fail_cuda("conv_bwd3x3 attr", e);
  e = cunet_launch(conv_bwd3x3_kernel, dim3(grid), dim3(B3_THREADS), smem, st, *d, w->dw, per);
  if (e != cudaSuccess) return cunet_fail_cuda("conv_bwd3x3 launch", e);
  return 1;
}

extern "C" int cunet_conv_bwd3x3(const cunet_conv_dgrad_params* d, const cunet_conv_wgrad_params* w, void* stream) {
  if (!d || !w) return cunet_fail("conv_bwd3x3: null params");
  const int r = conv_bwd3x3_try(d, w, reinterpret_cast<cudaStream_t>(stream));
  if (r != 0) return r < 0 ? r : 0;
  const int rc = cunet_conv_dgrad(d, stream);
  if (rc != 0) return rc;
  return cunet_conv_wgrad(w, stream);
}
