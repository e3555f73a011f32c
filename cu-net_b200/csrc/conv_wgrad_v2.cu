// Backward-filter of the fused 1x1 conv, second generation (bf16): persistent, every global read a 1-D TMA bulk
// copy, every concat chunk and the gradient operand handled by ONE CTA per pixel range.
// Same contract as conv_wgrad.cu (cunet_conv_wgrad_params).
//
//   dW[co][k] += sum_px dY[px][co] * A[px][k],   A = relu(bn(concat(x)))
//
// Round-1 structure: one CTA per (pixel range, 128-channel chunk); every chunk CTA re-gathered and re-transformed
// the gradient operand dY, and each 64-pixel stage was a serial load -> wait -> transform -> store chain with no
// prefetch (121 us for 320->128 at 64x64, batch 24 = 12 % of the HBM roofline).  Here:
//   * 64 consecutive pixels of an NHWC tensor are a contiguous block: a landing producer streams the raw blocks of
//     every source tensor and of G/T of the output into an 8-slot smem ring with cp.async.bulk, several jobs ahead;
//   * 8 transformer warps turn raw blocks into MN-major SWIZZLE_128B operand tiles smem -> smem (dT once per stage,
//     shared by all chunks; one activation tile per chunk);
//   * one accumulator D_c[128 k][128 co] per chunk stays in TMEM (up to 3 x 128 columns) for the CTA's whole pixel
//     range; a single reduction into the fp32 gradient at the end.
#include "loaders.cuh"
#include "host_util.h"

namespace cunet {

constexpr int W2_THREADS = 320;  // warp 0 landing producer, warp 1 MMA, warps 2-9 transformers
constexpr int W2_NSLOT = 8;
constexpr int W2_SLOT = 16384;
constexpr int W2_R = 64;                               // pixels per stage
constexpr int W2_SUB = W2_R * 128;                     // one MN-major sub-tile: 64 rows x 128 B
constexpr int W2_B_OFF = W2_NSLOT * W2_SLOT;           // dT operand, double buffered: 2 x 16 KB
constexpr int W2_A_OFF = W2_B_OFF + 2 * 16384;         // activation operand ring: 3 x 16 KB
constexpr int W2_TAIL_OFF = W2_A_OFF + 3 * 16384;

struct W2Tail {
  uint64_t slot_full[W2_NSLOT], slot_empty[W2_NSLOT];
  uint64_t b_full[2], b_free[2];
  uint64_t a_full[3], a_free[3];
  uint64_t done;
  uint32_t tmem_base;
  int low_loc[2][W2_R];   // per stage parity: half-resolution row of tile row r, relative to the stage's first
  int row_pos[2][W2_R];   // position inside the 2x2 window, or -1 when the row is past the end
  BnSmem bn;
  GradSmem gc;
};

__device__ __forceinline__ void w2_named_bar(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// first half-resolution row and count covered by the 64 raster pixels starting at m0 (power-of-two H, W <= 64)
__device__ __forceinline__ void w2_low_span(int m0, int nvalid, int H, int W, int& low0, int& nlow) {
  int n, h, w;
  pix_split(m0, H, W, n, h, w);
  low0 = (n * (H >> 1) + (h >> 1)) * (W >> 1) + (w >> 1);
  nlow = W >= 64 ? (nvalid >> 1) : (nvalid >> 2);   // one image row: 2 px per window; whole row pairs: 4 px
}

// timeline slots (CUNET_TRACE builds), stage i < 16 of CTA 0: producer 0+i (stage's first landing issued), transformer
// 16+4i (start, dT slots landed + B buffer free, dT operand done, all chunk operands done), MMA 80+2i (B ready, stage
// issued); 120: epilogue starts, 121: epilogue done
CUNET_TRACE_DECL(g_w2_trace)

__global__ void __launch_bounds__(W2_THREADS, 1) conv_wgrad_v2_kernel(const __grid_constant__ cunet_conv_wgrad_params p,
                                                                       int npad) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  W2Tail* tail = reinterpret_cast<W2Tail*>(smem + W2_TAIL_OFF);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int Cin = concat_cin(p.in);
  const int nchunk = (Cin + 127) >> 7;
  const int M = p.N * p.H * p.W;
  const int total = (M + W2_R - 1) / W2_R;
  const int per = (total + (int)gridDim.x - 1) / (int)gridDim.x;
  const int st0 = (int)blockIdx.x * per, st1 = min(total, st0 + per);
  const int ldo = p.dy.ld * 2;
  const int dTn = 1 + (p.dy.mode == 1 ? 1 : 0) + (p.dy.pooled ? 1 : 0);
  int need_low = p.dy.pooled;
  for (int s = 0; s < p.in.nseg; ++s) need_low |= p.in.seg[s].up;
  CUNET_TRACE_LOAD(trace, g_w2_trace)

  if (tid == 0) {
    for (int s = 0; s < W2_NSLOT; ++s) {
      mbar_init(&tail->slot_full[s], 1);
      mbar_init(&tail->slot_empty[s], 8);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tail->b_full[b], 8);
      mbar_init(&tail->b_free[b], 1);
    }
    for (int b = 0; b < 3; ++b) {
      mbar_init(&tail->a_full[b], 8);
      mbar_init(&tail->a_free[b], 1);
    }
    mbar_init(&tail->done, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(&tail->tmem_base, 512);
  griddep_wait();    // everything above overlaps the tail of the previous kernel (programmatic dependent launch)
  griddep_launch();
  compute_bn_coefs(p.in, &tail->bn, nchunk * 128, tid, W2_THREADS);
  compute_grad_coefs(p.dy, &tail->gc, tid, W2_THREADS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tail->tmem_base;

  if (warp == 0) {
    // ============================================================== landing producer
    if (lane == 0) {
      uint32_t jn = 0;
      auto land = [&](const void* src, uint32_t bytes) {
        const int slot = jn & 7;
        mbar_wait(&tail->slot_empty[slot], ((jn >> 3) & 1) ^ 1);
        if (bytes) {
          mbar_arrive_expect_tx(&tail->slot_full[slot], bytes);
          bulk_g2s(smem + slot * W2_SLOT, src, bytes, &tail->slot_full[slot]);
        } else {
          mbar_arrive(&tail->slot_full[slot]);
        }
        ++jn;
      };
      for (int st = st0; st < st1; ++st) {
        const int m0 = st * W2_R, nv = min(W2_R, M - m0);
        int low0 = 0, nlow = 0;
        if (need_low) w2_low_span(m0, nv, p.H, p.W, low0, nlow);
        const int r0 = p.dy.pooled ? low0 : m0, nr = p.dy.pooled ? nlow : nv;
        land(reinterpret_cast<const char*>(p.dy.g) + (long)r0 * ldo, (uint32_t)(nr * ldo));
        if (st - st0 < 16) CUNET_TRACE_MARK(trace, 0 + (st - st0));
        if (p.dy.mode == 1) land(reinterpret_cast<const char*>(p.dy.t) + (long)r0 * ldo, (uint32_t)(nr * ldo));
        if (p.dy.pooled) land(p.dy.pool_idx + (long)r0 * p.dy.C, (uint32_t)(nr * p.dy.C));
        for (int s = 0; s < p.in.nseg; ++s) {
          const cunet_seg& sg = p.in.seg[s];
          const int x0 = sg.up ? low0 : m0, nx = sg.up ? nlow : nv;
          land(reinterpret_cast<const char*>(sg.ptr) + (long)x0 * sg.C * 2, (uint32_t)(nx * sg.C * 2));
        }
      }
    }
  } else if (warp == 1) {
    // ============================================================== MMA issuer
    if (lane == 0) {
      const uint32_t idesc = make_idesc(Elem<bf16>::FMT, 128, (uint32_t)npad, 1, 1);  // both operands MN-major
      uint32_t ai = 0, si = 0;
      for (int st = st0; st < st1; ++st, ++si) {
        const uint32_t bb = si & 1;
        mbar_wait(&tail->b_full[bb], (si >> 1) & 1);
        if (si < 16) CUNET_TRACE_MARK(trace, 80 + 2 * si);
        const uint32_t b = smem_u32(smem + W2_B_OFF + bb * 16384);
        for (int c = 0; c < nchunk; ++c, ++ai) {
          const uint32_t ab = ai % 3;
          mbar_wait(&tail->a_full[ab], (ai / 3) & 1);
          tc_fence_after();
          const uint32_t a = smem_u32(smem + W2_A_OFF + ab * 16384);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma<bf16>(tmem + c * 128, make_sdesc_mn<bf16>(a + kk * 2048, W2_SUB), make_sdesc_mn<bf16>(b + kk * 2048, W2_SUB),
                       idesc, (uint32_t)((si | kk) != 0));
          tc_commit(&tail->a_free[ab]);
        }
        tc_commit(&tail->b_free[bb]);
        if (si < 16) CUNET_TRACE_MARK(trace, 81 + 2 * si);
      }
      tc_commit(&tail->done);
    }
  } else {
    // ============================================================== transformers (warps 2-9, 256 threads)
    const int t = tid - 64;
    const int cc = t & 15, rb = t >> 4;  // chunk column (16 B) x 16 rows per pass, 4 passes
    GradCoef<bf16> gcf;
    gcf.load(&tail->gc, (cc * 8) & 127);
    const bool gcol_ok = cc * 8 < p.dy.C;
    const bool gcol_used = cc * 8 < npad;
    uint32_t jn = 0, ai = 0, si = 0;
    // per-chunk constants of this thread's 16-byte column, hoisted out of the stage loop (the per-chunk segment
    // search and coefficient loads used to cost more instructions than the four transform passes they served)
    ActCoef<bf16> acf[3];
    int cs_s[3], cs_lo[3], cs_hi[3], cs_cl2[3], cs_ldx[3], cs_up[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const int ch = c * 128 + cc * 8;
      acf[c].load(&tail->bn, ch < MAX_CIN ? ch : 0);
      cs_s[c] = -1; cs_lo[c] = p.in.nseg; cs_hi[c] = -1; cs_cl2[c] = 0; cs_ldx[c] = 0; cs_up[c] = 0;
      if (c < nchunk) {
        if (ch < Cin) {
          int s = 0;
          while (ch >= tail->bn.seg_start[s + 1]) ++s;
          cs_s[c] = s;
          cs_cl2[c] = (ch - tail->bn.seg_start[s]) * 2;
          cs_ldx[c] = p.in.seg[s].C * 2;
          cs_up[c] = p.in.seg[s].up;
        }
        for (int q = 0; q < p.in.nseg; ++q)
          if ((tail->bn.seg_start[q] >> 7) == c) {
            cs_lo[c] = min(cs_lo[c], q);
            cs_hi[c] = max(cs_hi[c], q);
          }
      }
    }
    for (int st = st0; st < st1; ++st, ++si) {
      const int m0 = st * W2_R, nv = min(W2_R, M - m0);
      const uint32_t par = si & 1;
      if (need_low && t < W2_R) {
        int loc = 0, pos = -1;
        if (t < nv) {
          int n, h, w, low0, nlow;
          w2_low_span(m0, nv, p.H, p.W, low0, nlow);
          pix_split(m0 + t, p.H, p.W, n, h, w);
          loc = (n * (p.H >> 1) + (h >> 1)) * (p.W >> 1) + (w >> 1) - low0;
          pos = ((h & 1) << 1) | (w & 1);
        }
        tail->low_loc[par][t] = loc;
        tail->row_pos[par][t] = pos;
      }
      if (need_low) w2_named_bar(1, 256);
      const int* lowloc = tail->low_loc[par];
      const int* rpos = tail->row_pos[par];
      // ---- gradient operand dT -> B[par]
      const uint32_t jg = jn, jt = jn + 1, jx = jn + (p.dy.mode == 1 ? 2 : 1);
      if (t == 0 && si < 16) CUNET_TRACE_MARK(trace, 16 + 4 * si);
      mbar_wait(&tail->slot_full[jg & 7], (jg >> 3) & 1);
      if (p.dy.mode == 1) mbar_wait(&tail->slot_full[jt & 7], (jt >> 3) & 1);
      if (p.dy.pooled) mbar_wait(&tail->slot_full[jx & 7], (jx >> 3) & 1);
      mbar_wait(&tail->b_free[par], ((si >> 1) & 1) ^ 1);
      if (t == 0 && si < 16) CUNET_TRACE_MARK(trace, 17 + 4 * si);
      {
        const uint8_t* rg = smem + (jg & 7) * W2_SLOT;
        const uint8_t* rt = smem + (jt & 7) * W2_SLOT;
        const uint8_t* ri = smem + (jx & 7) * W2_SLOT;
        const uint32_t bbase = smem_u32(smem + W2_B_OFF + par * 16384);
        if (gcol_used) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int r = rb + 16 * q;
            uint4 o = make_uint4(0, 0, 0, 0), lo;
            if (gcol_ok && r < nv) {
              GradRaw<bf16> raw;
              const int loc = p.dy.pooled ? lowloc[r] : r;
              raw.g = *reinterpret_cast<const uint4*>(rg + loc * ldo + cc * 16);
              if (p.dy.mode == 1) raw.t = *reinterpret_cast<const uint4*>(rt + loc * ldo + cc * 16);
              if (p.dy.pooled) {
                const uint2 iv = *reinterpret_cast<const uint2*>(ri + loc * p.dy.C + cc * 8);
                raw.idx[0] = iv.x;
                raw.idx[1] = iv.y;
                raw.pos = (uint32_t)rpos[r];
              }
              o = gcf.apply(p.dy, raw, lo);
            }
            sts128(bbase + (cc >> 3) * W2_SUB + tile_off_mn<bf16>(r, cc & 7), o);
          }
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&tail->b_full[par]);
        mbar_arrive(&tail->slot_empty[jg & 7]);
        if (p.dy.mode == 1) mbar_arrive(&tail->slot_empty[jt & 7]);
        if (p.dy.pooled) mbar_arrive(&tail->slot_empty[jx & 7]);
      }
      if (t == 0 && si < 16) CUNET_TRACE_MARK(trace, 18 + 4 * si);
      jn += dTn;
      // ---- activation operand per chunk -> A ring
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        if (c >= nchunk) break;
        const uint32_t ab = ai % 3;
        const int s = cs_s[c], s_lo = cs_lo[c], s_hi = cs_hi[c];
        // every warp must observe every segment job of this chunk (single-party ring discipline)
        for (int q = s_lo; q <= s_hi; ++q) mbar_wait(&tail->slot_full[(jn + q) & 7], ((jn + q) >> 3) & 1);
        mbar_wait(&tail->a_free[ab], ((ai / 3) & 1) ^ 1);
        const uint32_t abase = smem_u32(smem + W2_A_OFF + ab * 16384);
        const uint8_t* rx = smem + ((jn + (s < 0 ? 0 : s)) & 7) * W2_SLOT + cs_cl2[c];
        const int ldx = cs_ldx[c];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int r = rb + 16 * q;
          uint4 o = make_uint4(0, 0, 0, 0), lo;
          if (s >= 0 && r < nv) {
            const int loc = cs_up[c] ? lowloc[r] : r;
            o = acf[c].apply(*reinterpret_cast<const uint4*>(rx + loc * ldx), lo);
          }
          sts128(abase + (cc >> 3) * W2_SUB + tile_off_mn<bf16>(r, cc & 7), o);
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&tail->a_full[ab]);
          for (int q = s_lo; q <= s_hi; ++q) mbar_arrive(&tail->slot_empty[(jn + q) & 7]);
        }
        ++ai;
      }
      jn += p.in.nseg;
      if (t == 0 && si < 16) CUNET_TRACE_MARK(trace, 19 + 4 * si);
    }
    // ============================================================== epilogue: TMEM -> red.global.add
    if (st1 > st0) {
      mbar_wait(&tail->done, 0);
      if (t == 0) CUNET_TRACE_MARK(trace, 120);
      tc_fence_after();
      const int e = warp - 2;
      const int qd = warp & 3, hf = e >> 2;
      const int dwc = p.dw_cin > 0 ? p.dw_cin : Cin;
      const int nh = npad >> 1;
      for (int c = 0; c < nchunk; ++c) {
        const int k = c * 128 + qd * 32 + lane;
        for (int j = 0; j < nh; j += 8) {
          float v[8];
          const int col = hf * nh + j;
          tmem_ld8(tmem + c * 128 + ((uint32_t)(qd * 32) << 16) + (uint32_t)col, v);
          if (k < dwc && k < Cin) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const int co = col + q;
              if (co < p.Cout) atomicAdd(p.dw + (long)co * dwc + k, v[q]);
            }
          }
        }
      }
      if (t == 0) CUNET_TRACE_MARK(trace, 121);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

}  // namespace cunet
using namespace cunet;

CUNET_TRACE_SETTER(cunet_debug_trace_wgrad_v2, g_w2_trace)

// Returns 1 when the v2 kernel handled the call, 0 when the caller must use the generic kernel, <0 on error.
int cunet_conv_wgrad_v2_try(const cunet_conv_wgrad_params* p, cudaStream_t st) {
  if (p->dtype != CUNET_BF16 || p->taps != 1) return 0;
  if (p->dy.ld != p->dy.C || p->dy.C > 128 || (p->dy.C & 7) || p->Cout > 128) return 0;
  int cin = 0, low = p->dy.pooled;
  for (int s = 0; s < p->in.nseg; ++s) {
    const cunet_seg& sg = p->in.seg[s];
    if (sg.ld != sg.C || (sg.C != 32 && sg.C != 64 && sg.C != 128)) return 0;
    if ((cin >> 7) != ((cin + sg.C - 1) >> 7)) return 0;
    cin += sg.C;
    low |= sg.up;
  }
  if (cin > MAX_CIN || cin > 384) return 0;
  if (p->in.nseg + 3 > W2_NSLOT) return 0;
  const int W = p->W, H = p->H;
  if (low) {
    if ((W & (W - 1)) || (H & (H - 1)) || W > 64 || W < 2 || H < 2) return 0;
    if (W < 64 && (64 % (2 * W))) return 0;   // a stage must cover whole row pairs
  }
  const long M = (long)p->N * p->H * p->W;
  if (M <= 0) return 1;
  if (low && (M % 64) && W < 64 && ((M % 64) % (2 * W))) return 0;
  const int npad = ((p->dy.C + 63) / 64) * 64;
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  const int total = (int)((M + W2_R - 1) / W2_R);
  int grid = p->nsplit > 0 ? p->nsplit : sms;
  if (grid > total) grid = total;
  const size_t smem = W2_TAIL_OFF + sizeof(W2Tail) + 1024;
  cudaError_t e = cudaFuncSetAttribute(conv_wgrad_v2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return cunet_fail_cuda("conv_wgrad_v2 attr", e);
  e = cunet_launch(conv_wgrad_v2_kernel, dim3(grid), dim3(W2_THREADS), smem, st, *p, npad);
  if (e != cudaSuccess) return cunet_fail_cuda("conv_wgrad_v2 launch", e);
  return 1;
}
