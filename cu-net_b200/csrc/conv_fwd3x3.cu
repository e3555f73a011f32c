// Forward of the dense-layer 3x3 conv (models/cu_net.py:45-48: norm2 -> relu -> conv2, 128 -> 32), bf16, second
// generation: persistent, bulk-landed input, NO im2col and no per-tap re-gather.
//
// The round-1 kernel (conv_fwd.cu, taps == 9) re-gathered and re-transformed the 128-channel input once per tap
// (18 K steps per tile, nine global gathers per pixel): 82 us at 64x64 / batch 24 against a 5 us HBM floor.
// Here the nine taps are split into a ROW shift (dy) and a COLUMN shift (dx):
//
//   out[q][co] = sum_dx Y_dx[q + dx][co] * [0 <= w + dx < W],     Y_dx[p][co] = sum_dy sum_ci a[p + dy*W][ci] * Wt[co][ci][dy][dx]
//
//   * the activated input a = relu(bn(x)) of a tile and its one-row halo is written ONCE into shared memory as a
//     K-major SWIZZLE_128B operand over "padded raster" positions (every image gets a zero row above and below, so
//     a shift by dy image rows is a shift by W buffer rows everywhere, image borders included);
//   * a shift by W rows (W >= 8 a power of two) is a multiple of the 1024-byte swizzle period: the operand of tap
//     row dy is the SAME buffer with the descriptor start address advanced by dy*W*128 bytes -- zero copies;
//   * the three dx taps become 96 output columns of one GEMM  D[128 pos][(dx, co)] = sum_{dy, half} A_dy * B_dy^T
//     (6 K blocks, 24 tcgen05.mma of N = 96 instead of 72 of N = 32), and the epilogue adds the column-shifted
//     partials: lane q takes Y_-1 from lane q-1 and Y_+1 from lane q+1 (warp shuffles; tiles start at w == 0, so a
//     masked neighbour is never in another tile).
// Weights stay resident in shared memory for the CTA's whole tile list (18 bulk copies of the ordinary forward
// image, re-stacked as [dy][half][(dx, co)] rows); x is landed with <= 5 bulk copies per tile (one per image run).
#include "loaders.cuh"
#include "host_util.h"
#include <stdlib.h>

namespace cunet {

constexpr int F3_THREADS = 512;  // warp 0 landing producer | 1 MMA | 2-3 idle | 4-11 transformers | 12-15 epilogue
constexpr int F3_W_BYTES = 18 * 4096;            // 72 KB: [dy][half][96 rows][128 B]
constexpr int F3_W_OFF = 0;
constexpr int F3_A_OFF = F3_W_OFF + F3_W_BYTES;  // activated operand: [half][256 rows][128 B] = 64 KB
constexpr int F3_RAW_OFF = F3_A_OFF + 65536;     // landed raw x: [256 positions][256 B] = 64 KB
constexpr int F3_TAIL_OFF = F3_RAW_OFF + 65536;

struct F3Tail {
  uint64_t w_full, raw_full, raw_free, a_ready, a_free;
  uint64_t acc_full[2], acc_free[2];
  uint32_t tmem_base;
  int rowv[2][256];          // transformers: 1 when the window row is a real pixel
  float xch[2][4][2][32];    // epilogue: [tile parity][warp][0: lane 31's Y_-1, 1: lane 0's Y_+1][co]
  BnSmem bn;
};

__device__ __forceinline__ uint4 f3_lds128(uint32_t saddr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr));
  return v;
}
__device__ __forceinline__ void f3_named_bar(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// padded raster: position -> (image, padded row hp in [0, H+2), w); real pixel iff 1 <= hp <= H
// Maps narrower than 8 pixels (the 4x4 neck of the hourglass) use a raster of width 8 whose columns >= Wr are padding
// like the rows above and below an image: a shift by one raster row then stays a multiple of the swizzle period.
struct F3Geom {
  int H, W, lw, S, P;  // W = raster width (>= 8), S = (H + 2) * W positions per image, P = N * S
  int Wr, lwr;         // real width of the map and its log2
  __device__ __forceinline__ bool real(int pos, int& grow, int& w) const {
    w = pos & (W - 1);
    if (pos < 0 || pos >= P) return false;
    const int img = pos / S;
    const int hp = (pos - img * S) >> lw;
    grow = ((img * H + hp - 1) << lwr) + w;
    return hp >= 1 && hp <= H && w < Wr;
  }
};

// timeline slots (CUNET_TRACE builds), tile i < 16: producer 0+i (landing issued), transformer 16+3i (start, raw landed +
// operand free, done), MMA 64+2i (operand ready + accumulator free, issued), epilogue 96+3i (accumulator full, TMEM
// drained, stores + statistics done)
CUNET_TRACE_DECL(g_f3_trace)

__global__ void __launch_bounds__(F3_THREADS, 1) conv_fwd3x3_kernel(const __grid_constant__ cunet_conv_fwd_params p,
                                                                     int ntiles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  F3Tail* tail = reinterpret_cast<F3Tail*>(smem + F3_TAIL_OFF);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  F3Geom g;
  const int W = max(p.W, 8), Wr = p.W;   // raster width, real width
  g.H = p.H; g.W = W; g.lw = 31 - __clz(W); g.S = (p.H + 2) * W; g.P = p.N * g.S;
  g.Wr = Wr; g.lwr = 31 - __clz(Wr);
  const int tile0 = (int)blockIdx.x, tstride = (int)gridDim.x;
  CUNET_TRACE_LOAD(trace, g_f3_trace)

  if (tid == 0) {
    mbar_init(&tail->w_full, 1);
    mbar_init(&tail->raw_full, 1);
    mbar_init(&tail->raw_free, 8);
    mbar_init(&tail->a_ready, 8);
    mbar_init(&tail->a_free, 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tail->acc_full[b], 1);
      mbar_init(&tail->acc_free[b], 4);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(&tail->tmem_base, 256);
  __syncthreads();   // barrier inits visible to every warp; still under the previous kernel (before griddepcontrol.wait)
  griddep_wait();    // everything above overlaps the tail of the previous kernel (programmatic dependent launch)
  griddep_launch();
  // warp 0 (landing producer) needs no BatchNorm coefficient: it starts its copies at once, the other warps compute
  // the coefficients behind a named barrier that does not include it (round 2, as in conv_fwd_v3 / conv_bwd1x1)
  if (warp > 0) {
    compute_bn_coefs(p.in, &tail->bn, 128, tid - 32, F3_THREADS - 32);
    tc_fence_before();
    f3_named_bar(2, F3_THREADS - 32);
    tc_fence_after();
  }
  const uint32_t tmem = warp > 0 ? tail->tmem_base : 0u;

  if (warp == 0) {
    // ============================================================== landing producer
    if (lane == 0 && tile0 < ntiles) {
      // weights: forward image block (tap = dy*3 + dx, half) of [32 co rows][128 B] -> row group dx of tile (dy, half)
      mbar_arrive_expect_tx(&tail->w_full, (uint32_t)F3_W_BYTES);
      const char* wsrc = reinterpret_cast<const char*>(p.wpack);
      for (int dy = 0; dy < 3; ++dy)
        for (int dx = 0; dx < 3; ++dx)
          for (int kb = 0; kb < 2; ++kb)
            bulk_g2s(smem + F3_W_OFF + (dy * 2 + kb) * 12288 + dx * 4096, wsrc + ((dy * 3 + dx) * 2 + kb) * 4096, 4096u,
                     &tail->w_full);
      const char* xsrc = reinterpret_cast<const char*>(p.in.seg[0].ptr);
      const int HW = p.H * W;
      uint32_t i = 0;
      for (int tile = tile0; tile < ntiles; tile += tstride, ++i) {
        const int p0w = tile * 128 - W;  // position of window row 0
        const int ws = max(0, p0w), we = min(g.P, tile * 128 + 128 + W);
        if (Wr < W) {
          // narrow map: one copy per real image row of the window (window bounds are multiples of the raster width)
          const int r0 = ws >> g.lw, r1 = we >> g.lw;
          uint32_t nreal = 0;
          for (int rr = r0; rr < r1; ++rr) {
            const int hp = rr % (p.H + 2);
            nreal += (hp >= 1 && hp <= p.H) ? 1u : 0u;
          }
          mbar_wait(&tail->raw_free, (i & 1u) ^ 1u);
          if (i < 16) CUNET_TRACE_MARK(trace, 0 + i);
          if (nreal) mbar_arrive_expect_tx(&tail->raw_full, nreal * (uint32_t)Wr * 256u);
          else mbar_arrive(&tail->raw_full);
          for (int rr = r0; rr < r1; ++rr) {
            const int img = rr / (p.H + 2), hp = rr - img * (p.H + 2);
            if (hp >= 1 && hp <= p.H)
              bulk_g2s(smem + F3_RAW_OFF + (rr * W - p0w) * 256, xsrc + ((long)(img * p.H + hp - 1) * Wr) * 256,
                       (uint32_t)Wr * 256u, &tail->raw_full);
          }
          continue;
        }
        const int i0 = ws / g.S, i1 = (we - 1) / g.S;
        uint32_t total = 0;
        for (int img = i0; img <= i1; ++img) {
          const int a = max(ws, img * g.S + W), b = min(we, img * g.S + (p.H + 1) * W);
          if (a < b) total += (uint32_t)(b - a) * 256u;
        }
        mbar_wait(&tail->raw_free, (i & 1u) ^ 1u);
        if (i < 16) CUNET_TRACE_MARK(trace, 0 + i);
        if (total) mbar_arrive_expect_tx(&tail->raw_full, total);
        else mbar_arrive(&tail->raw_full);
        for (int img = i0; img <= i1; ++img) {
          const int a = max(ws, img * g.S + W), b = min(we, img * g.S + (p.H + 1) * W);
          if (a < b)
            bulk_g2s(smem + F3_RAW_OFF + (a - p0w) * 256, xsrc + ((long)img * HW + (a - img * g.S - W)) * 256,
                     (uint32_t)(b - a) * 256u, &tail->raw_full);
        }
      }
    }
  } else if (warp == 1) {
    // ============================================================== MMA issuer
    if (lane == 0 && tile0 < ntiles) {
      const uint32_t idesc = make_idesc(Elem<bf16>::FMT, 128, 96, 0, 0);
      const uint32_t wB = smem_u32(smem + F3_W_OFF), aA = smem_u32(smem + F3_A_OFF);
      mbar_wait(&tail->w_full, 0);
      uint32_t i = 0;
      for (int tile = tile0; tile < ntiles; tile += tstride, ++i) {
        const uint32_t b = i & 1u;
        mbar_wait(&tail->a_ready, i & 1u);
        mbar_wait(&tail->acc_free[b], ((i >> 1) & 1u) ^ 1u);
        if (i < 16) CUNET_TRACE_MARK(trace, 64 + 2 * i);
        tc_fence_after();
        // descriptors are advanced from two bases (sdesc_advance, common.cuh): the issuer is one thread and rebuilding
        // 48 descriptors per tile from scratch was a measurable part of its time
        const uint64_t adesc0 = make_sdesc(aA, 16, 1024), wdesc0 = make_sdesc(wB, 16, 1024);
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
          const uint64_t ady = sdesc_advance(adesc0, (uint32_t)(dy * W * 128));
#pragma unroll
          for (int kb = 0; kb < 2; ++kb) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              umma<bf16>(tmem + b * 128u, sdesc_advance(ady, kb * 32768 + kk * 32),
                         sdesc_advance(wdesc0, (dy * 2 + kb) * 12288 + kk * 32), idesc,
                         (uint32_t)((dy | kb | kk) != 0));
          }
        }
        tc_commit(&tail->a_free);
        tc_commit(&tail->acc_full[b]);
        if (i < 16) CUNET_TRACE_MARK(trace, 65 + 2 * i);
      }
    }
  } else if (warp >= 4 && warp < 12) {
    // ============================================================== transformers (256 threads)
    const int t = tid - 128;
    const int cc = t & 15, rbase = t >> 4;  // 16-byte column (8 channels) x window rows rbase + 16 j
    ActCoef<bf16> acf;
    acf.load(&tail->bn, cc * 8);
    const uint32_t rawb = smem_u32(smem + F3_RAW_OFF) + (uint32_t)cc * 16u;
    const uint32_t ab = smem_u32(smem + F3_A_OFF) + (uint32_t)(cc >> 3) * 32768u;
    uint32_t i = 0;
    for (int tile = tile0; tile < ntiles; tile += tstride, ++i) {
      const int p0w = tile * 128 - W;
      int* rv = tail->rowv[i & 1];
      {
        int grow, w;
        rv[t] = g.real(p0w + t, grow, w) ? 1 : 0;
      }
      f3_named_bar(1, 256);
      if (t == 0 && i < 16) CUNET_TRACE_MARK(trace, 16 + 3 * i);
      mbar_wait(&tail->raw_full, i & 1u);
      mbar_wait(&tail->a_free, (i & 1u) ^ 1u);  // MMAs of the previous tile no longer read the operand
      if (t == 0 && i < 16) CUNET_TRACE_MARK(trace, 17 + 3 * i);
#pragma unroll 4
      for (int j = 0; j < 16; ++j) {
        const int r = rbase + 16 * j;
        uint4 o = make_uint4(0, 0, 0, 0), lo_unused;
        if (rv[r]) o = acf.apply(f3_lds128(rawb + (uint32_t)r * 256u), lo_unused);
        sts128(ab + tile_off(r, cc & 7), o);
      }
      fence_proxy_async();
      __syncwarp();
      if (t == 0 && i < 16) CUNET_TRACE_MARK(trace, 18 + 3 * i);
      if (lane == 0) {
        mbar_arrive(&tail->a_ready);
        mbar_arrive(&tail->raw_free);
      }
    }
  } else if (warp >= 12) {
    // ============================================================== epilogue (128 threads, thread = tile position)
    const int qd = warp & 3;  // TMEM lane quarter (hardware: warp % 4)
    const int row = qd * 32 + lane;
    double s1 = 0.0, s2 = 0.0;  // lane l: channel l
    char* outp = reinterpret_cast<char*>(p.out);
    uint32_t i = 0;
    for (int tile = tile0; tile < ntiles; tile += tstride, ++i) {
      const uint32_t b = i & 1u;
      int grow = 0, w = 0;
      const bool valid = g.real(tile * 128 + row, grow, w);
      mbar_wait(&tail->acc_full[b], (i >> 1) & 1u);
      if (tid == 384 && i < 16) CUNET_TRACE_MARK(trace, 96 + 3 * i);
      tc_fence_after();
      const uint32_t tb = tmem + b * 128u + ((uint32_t)(qd * 32) << 16);
      float ym[32], yp[32];
#pragma unroll
      for (int c8 = 0; c8 < 4; ++c8) {
        tmem_ld8(tb + (uint32_t)(c8 * 8), ym + c8 * 8);
        tmem_ld8(tb + (uint32_t)(64 + c8 * 8), yp + c8 * 8);
      }
      float(*xc)[2][32] = tail->xch[i & 1];
      if (lane == 31) {
#pragma unroll
        for (int j = 0; j < 32; ++j) xc[qd][0][j] = ym[j];
      }
      if (lane == 0) {
#pragma unroll
        for (int j = 0; j < 32; ++j) xc[qd][1][j] = yp[j];
      }
      f3_named_bar(2, 128);
      const bool take_l = valid && w > 0, take_r = valid && w < Wr - 1;
      float o[32];
#pragma unroll
      for (int c8 = 0; c8 < 4; ++c8) {
        float y0[8];
        tmem_ld8(tb + (uint32_t)(32 + c8 * 8), y0);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int c = c8 * 8 + j;
          float up = __shfl_up_sync(0xffffffffu, ym[c], 1);
          float dn = __shfl_down_sync(0xffffffffu, yp[c], 1);
          if (lane == 0) up = qd > 0 ? xc[qd - 1][0][c] : 0.f;     // row 0 of a tile has w == 0: masked anyway
          if (lane == 31) dn = qd < 3 ? xc[qd + 1][1][c] : 0.f;    // row 127 is the last raster column: masked anyway
          o[c] = y0[j] + (take_l ? up : 0.f) + (take_r ? dn : 0.f);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tail->acc_free[b]);  // TMEM buffer drained by this warp
      if (tid == 384 && i < 16) CUNET_TRACE_MARK(trace, 97 + 3 * i);
      // round to the storage type, store this pixel's 32 channels (64 contiguous bytes)
      uint32_t pk[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        __nv_bfloat162 h = __floats2bfloat162_rn(o[2 * j], o[2 * j + 1]);
        pk[j] = *reinterpret_cast<uint32_t*>(&h);
        // statistics of the values as stored; pad positions contribute nothing
        o[2 * j] = valid ? __uint_as_float(pk[j] << 16) : 0.f;
        o[2 * j + 1] = valid ? __uint_as_float(pk[j] & 0xFFFF0000u) : 0.f;
      }
      if (valid) {
        uint4* dst = reinterpret_cast<uint4*>(outp + (long)grow * 64);
        dst[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        dst[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        dst[2] = make_uint4(pk[8], pk[9], pk[10], pk[11]);
        dst[3] = make_uint4(pk[12], pk[13], pk[14], pk[15]);
      }
      if (p.out_stats != nullptr) {
        // transpose-reduce over the warp: after 5 halving steps lane l holds the 32-position sum of channel l
        float q2[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) q2[j] = o[j] * o[j];
#pragma unroll
        for (int s = 16; s >= 1; s >>= 1) {
          const bool hi = (lane & s) != 0;
#pragma unroll
          for (int j = 0; j < s; ++j) {
            const float send1 = hi ? o[j] : o[j + s], keep1 = hi ? o[j + s] : o[j];
            const float send2 = hi ? q2[j] : q2[j + s], keep2 = hi ? q2[j + s] : q2[j];
            o[j] = keep1 + __shfl_xor_sync(0xffffffffu, send1, s);
            q2[j] = keep2 + __shfl_xor_sync(0xffffffffu, send2, s);
          }
        }
        s1 += (double)o[0];
        s2 += (double)q2[0];
      }
      if (tid == 384 && i < 16) CUNET_TRACE_MARK(trace, 98 + 3 * i);
    }
    if (p.out_stats != nullptr && tile0 < ntiles) {
      atomicAdd(p.out_stats + lane, s1);
      atomicAdd(p.out_stats + 32 + lane, s2);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 256);
}

}  // namespace cunet
using namespace cunet;

CUNET_TRACE_SETTER(cunet_debug_trace_fwd3x3, g_f3_trace)

// Returns 1 when this kernel handled the call, 0 when the caller must use the generic kernel, <0 on error.
int cunet_conv_fwd3x3_try(const cunet_conv_fwd_params* p, cudaStream_t st) {
  static const bool off = getenv("CUNET_FWD3X3_OFF") != nullptr;
  if (off) return 0;
  if (p->dtype != CUNET_BF16 || p->taps != 9 || p->pool || p->out_fp32) return 0;
  if (p->Cout != 32 || p->CoutPad != 32 || p->out_ld != 32) return 0;
  if (p->in.nseg != 1 || p->in.bn_train == 2) return 0;
  const cunet_seg& sg = p->in.seg[0];
  if (sg.C != 128 || sg.ld != 128 || sg.up) return 0;
  const int W = p->W, H = p->H;
  if ((W & (W - 1)) || W < 2 || W > 64 || H < 1) return 0;
  const long P = (long)p->N * (H + 2) * (W < 8 ? 8 : W);   // narrow maps: raster of width 8 (F3Geom)
  if (P <= 0) return 1;
  if (P > (1L << 30)) return 0;
  const int ntiles = (int)((P + 127) / 128);
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  const int grid = ntiles < sms ? ntiles : sms;
  const size_t smem = F3_TAIL_OFF + sizeof(F3Tail) + 1024;
  cudaError_t e = cudaFuncSetAttribute(conv_fwd3x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return cunet_fail_cuda("conv_fwd3x3 attr", e);
  e = cunet_launch(conv_fwd3x3_kernel, dim3(grid), dim3(F3_THREADS), smem, st, *p, ntiles);
  if (e != cudaSuccess) return cunet_fail_cuda("conv_fwd3x3 launch", e);
  return 1;
}
