// Forward of the 1x1 fused conv, third generation (bf16): cat -> BatchNorm -> ReLU -> conv1x1 [-> 2x2 max-pool]
// (adapters, bottleneck conv1, intermedia adapters, heat-map heads).  Same contract as conv_fwd.cu for taps == 1.
//
// Built like the fused backward (conv_bwd1x1.cu), whose structure it shares:
//   * persistent CTAs, 64-pixel stages; every source piece of the virtual concat is a contiguous block of an NHWC
//     tensor and lands by 1-D TMA bulk copy (double buffered); an upsampled source contributes its half-resolution
//     run (nearest x2 = index math in the transform);
//   * the weight image of the whole conv (<= 96 KB) stays resident in shared memory;
//   * 8 transformer warps apply BatchNorm + ReLU smem -> smem into the K-major SWIZZLE_128B operand, one 128-channel
//     chunk at a time, and the MMA issuer consumes chunk c while chunk c+1 is being transformed;
//   * TRANSPOSED GEMM  D[128 cout][64 px] = Wimg[cout][Cin] * A[px][Cin]^T : an epilogue thread owns ONE output channel
//     (its TMEM lane), so the per-channel sum / sum of squares the consumers' BatchNorms need are register accumulators
//     (the second-generation kernel reduced them with 248 shuffles per tile and was no faster than generation 1);
//     16 epilogue warps round, write the [px][cout] staging block with 2-byte stores and one thread bulk-stores it;
//   * the 2x2 max-pool of a down-block adapter is four TMEM columns of the same lane (a stage of a 64-wide image is
//     2 rows x 32 columns so that every pooling window is complete): max, first-maximum argmax byte, pooled statistics.
#include "loaders.cuh"
#include "host_util.h"
#include <stdlib.h>

namespace cunet {

// warp 0 landing producer | 1 store issuer | 2 MMA | 3 idle | 4-19 transformers | 20-27 epilogue.  The operand transform
// is the long pole of the forward (in-kernel timeline, tools/time_fwd_v3.py: 2.1 us per 64-pixel stage of a 320-channel
// concat on 8 warps, against 0.5 us of epilogue on 16), so it gets 16 warps and the epilogue 8 (32 pixels per thread).
constexpr int F3_THREADS = 896;
constexpr int F3_R = 64;
constexpr int F3_SUB = F3_R * 128;
constexpr int F3_MAXCH = 3;
constexpr int F3_DBUF = 4;       // TMEM accumulator ring (64 columns each)

struct F3Layout {
  int w_off, a_off, x_off, o_off, tail_off;
  int x_bytes;     // one x buffer
  int o_bytes;     // one output staging buffer
  int split;       // 1: stage = 2 rows x 32 columns (64-wide image with pooling)
  int low_rows;    // rows of an upsampled source's piece per stage (16, or 32 for a raster stage of a 64-wide image)
  int per;         // stages per CTA
};

struct F3Tail {
  alignas(16) uint32_t sc2[MAX_CIN / 2];
  alignas(16) uint32_t sh2[MAX_CIN / 2];
  uint64_t w_full;
  uint64_t x_full[2], x_free[2], a_full[F3_MAXCH], a_free[F3_MAXCH], d_full[F3_DBUF], d_free[F3_DBUF], o_ready[2],
      o_free[2];
  uint32_t tmem_base;
  int seg_start[CUNET_MAX_SEG + 1];
  int xoff[CUNET_MAX_SEG];
  int lowmap[F3_R];
};

__device__ __forceinline__ void f3_bulk_s2g(void* dst, const void* src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src_smem)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void f3_bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void f3_bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ uint4 f3_lds128(uint32_t saddr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr));
  return v;
}
template <int OFF> __device__ __forceinline__ void f3_sts_u16_o(uint32_t saddr, uint16_t v) {
  asm volatile("st.shared.u16 [%0+%1], %2;" ::"r"(saddr), "n"(OFF), "h"(v) : "memory");
}
__device__ __forceinline__ void f3_sts_u16(uint32_t saddr, uint16_t v) {
  asm volatile("st.shared.u16 [%0], %1;" ::"r"(saddr), "h"(v) : "memory");
}
__device__ __forceinline__ void f3_sts_u8(uint32_t saddr, uint32_t v) {
  asm volatile("st.shared.u8 [%0], %1;" ::"r"(saddr), "r"(v) : "memory");
}
__device__ __forceinline__ void f3_sts_f32(uint32_t saddr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(saddr), "f"(v) : "memory");
}
__device__ __forceinline__ void f3_tmem_ld16_nowait(uint32_t taddr, float* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]), "=f"(v[8]),
        "=f"(v[9]), "=f"(v[10]), "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15])
      : "r"(taddr));
}
__device__ __forceinline__ void f3_tmem_ld8_nowait(uint32_t taddr, float* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
               : "r"(taddr));
}
__device__ __forceinline__ void f3_named_bar(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void f3_tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

struct F3Geo {
  int p0, p1, nv, low0, nlow;
};
__device__ __forceinline__ F3Geo f3_geo(int st, int M, int W, int split, int need_low) {
  F3Geo g;
  if (split) {                       // W == 64: row pair a = st >> 1, column half b = st & 1
    const int a = st >> 1, b = st & 1;
    g.p0 = a * 128 + 32 * b;
    g.p1 = g.p0 + 64;
    g.nv = 64;
    g.low0 = a * 32 + 16 * b;        // pooled output pixels of the stage
    g.nlow = 16;
  } else {
    g.p0 = st * F3_R;
    g.p1 = -1;
    g.nv = min(F3_R, M - g.p0);
    g.low0 = 0;
    g.nlow = 0;
    if (need_low) {
      if (W == 64) {                 // one image row: its half-resolution row (H is even)
        g.low0 = ((g.p0 >> 6) >> 1) * 32;
        g.nlow = 32;
      } else {                       // whole row pairs (W <= 32)
        g.low0 = g.p0 >> 2;
        g.nlow = g.nv >> 2;
      }
    }
  }
  return g;
}

// 16 pixels of one output channel: round to bf16, statistics of the values as stored, 2-byte stores into the staging
// block.  CP2 = bytes between consecutive pixels of the staging block (= 2 * Cout).
template <int CP2>
__device__ __forceinline__ void f3_ep16(uint32_t a0, const float* v, float& s1, float& s2) {
  uint16_t h[16];
#pragma unroll
  for (int q = 0; q < 16; ++q) {
    h[q] = __bfloat16_as_ushort(__float2bfloat16_rn(v[q]));
    const float r = __uint_as_float((uint32_t)h[q] << 16);
    s1 += r;
    s2 = fmaf(r, r, s2);
  }
#define F3_ST(q) f3_sts_u16_o<(q) * CP2>(a0, h[q]);
  F3_ST(0) F3_ST(1) F3_ST(2) F3_ST(3) F3_ST(4) F3_ST(5) F3_ST(6) F3_ST(7)
  F3_ST(8) F3_ST(9) F3_ST(10) F3_ST(11) F3_ST(12) F3_ST(13) F3_ST(14) F3_ST(15)
#undef F3_ST
}

CUNET_TRACE_DECL(g_f3_trace)

__global__ void __launch_bounds__(F3_THREADS, 1) conv_fwd_v3_kernel(const __grid_constant__ cunet_conv_fwd_params p,
                                                                     const __grid_constant__ F3Layout L) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  F3Tail* tail = reinterpret_cast<F3Tail*>(smem + L.tail_off);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int H = p.H, W = p.W;
  const int M = p.N * H * W;
  const int Cin = concat_cin(p.in);
  const int nchunk = (Cin + 127) >> 7;
  const int nkb = (Cin + 63) >> 6;
  const int total = (M + F3_R - 1) / F3_R;
  const int st0 = (int)blockIdx.x * L.per;
  const int st1 = min(total, st0 + L.per);
  const int ns = max(0, st1 - st0);
  const int split = L.split;
  const uint32_t wblk = (uint32_t)p.CoutPad * 128u;   // bytes of one K block of the weight image
  int any_up = 0;
  for (int s = 0; s < p.in.nseg; ++s) any_up |= p.in.seg[s].up;
  const int need_low = any_up | p.pool;
  CUNET_TRACE_LOAD(trace, g_f3_trace)
  if (tid == 0) CUNET_TRACE_MARK(trace, 296);

  if (tid == 0) {
    mbar_init(&tail->w_full, 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tail->x_full[b], 1);
      mbar_init(&tail->x_free[b], 16);
      mbar_init(&tail->o_ready[b], 8);
      mbar_init(&tail->o_free[b], 1);
    }
    for (int c = 0; c < F3_MAXCH; ++c) {
      mbar_init(&tail->a_full[c], 16);
      mbar_init(&tail->a_free[c], 1);
    }
    for (int b = 0; b < F3_DBUF; ++b) {
      mbar_init(&tail->d_full[b], 1);
      mbar_init(&tail->d_free[b], 8);
    }
    fence_mbar_init();
    int acc = 0, xo = 0;
    for (int s = 0; s < p.in.nseg; ++s) {
      tail->seg_start[s] = acc;
      acc += p.in.seg[s].C;
      tail->xoff[s] = xo;
      xo += (p.in.seg[s].up ? L.low_rows : F3_R) * p.in.seg[s].C * 2;
    }
    for (int s = p.in.nseg; s <= CUNET_MAX_SEG; ++s) tail->seg_start[s] = acc;
  }
  if (tid < F3_R) {
    const int r = tid;
    int lm = 0;
    if (split) lm = (r & 31) >> 1;
    else if (W == 64) lm = r >> 1;
    else {
      const int lw = 31 - __clz(W);
      const int hl = r >> lw, w = r & (W - 1);
      lm = (hl >> 1) * (W >> 1) + (w >> 1);
    }
    tail->lowmap[r] = lm;
  }
  if (warp == 2) tmem_alloc(&tail->tmem_base, 256);
  __syncthreads();   // barrier inits and the static tables above are visible to every warp (warp 0 takes no part in the
                     // coefficient-phase barriers below); still before griddepcontrol.wait, i.e. under the previous kernel
  griddep_wait();    // everything above overlaps the tail of the previous kernel (programmatic dependent launch)
  griddep_launch();
  if (tid == 0) CUNET_TRACE_MARK(trace, 297);
  // landing of one stage's source pieces (thread 0 only)
  auto land_stage = [&](int i) {
    const F3Geo g = f3_geo(st0 + i, M, W, split, need_low);
    const uint32_t b = (uint32_t)i & 1u;
    mbar_wait(&tail->x_free[b], (((uint32_t)i >> 1) & 1u) ^ 1u);
    if (i < 12) CUNET_TRACE_MARK(trace, 0 + i);
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
  // Warp 0 is the landing producer and needs no BatchNorm coefficient: its thread starts issuing the weight image
  // and the source pieces right away, while warps 1.. compute the coefficients behind a NAMED barrier that does not
  // include warp 0 (with a CTA-wide barrier here everybody waited ~2 us for thread 0's serial issue work; the
  // coefficient scratch lives in the operand area, not in the x landing area, for the same reason).
  uint32_t tmem = 0;
  int relu_on = 1;
  if (warp == 0) {
    if (lane == 0 && ns > 0) {
      mbar_arrive_expect_tx(&tail->w_full, (uint32_t)nkb * wblk);
      bulk_g2s(smem + L.w_off, p.wpack, (uint32_t)nkb * wblk, &tail->w_full);
      for (int i = 0; i < ns; ++i) land_stage(i);
    }
  } else {
    BnSmem* bn = reinterpret_cast<BnSmem*>(smem + L.a_off);   // the operand area is not in use yet
    compute_bn_coefs(p.in, bn, nchunk * 128, tid - 32, F3_THREADS - 32);
    tc_fence_before();
    f3_named_bar(1, F3_THREADS - 32);
    tc_fence_after();
    if (tid == 32) CUNET_TRACE_MARK(trace, 298);
    tmem = tail->tmem_base;
    for (int i = tid - 32; i < nchunk * 64; i += F3_THREADS - 32) {
      tail->sc2[i] = bn->sc2[i];
      tail->sh2[i] = bn->sh2[i];
    }
    relu_on = bn->relu;
    f3_named_bar(1, F3_THREADS - 32);
  }

  const bool is_tr = warp >= 4 && warp < 20, is_ep = warp >= 20;
  const int t = tid - 128;                     // transformer thread index (0..511)
  const int cc = t & 15, rb = t >> 4;          // 16-byte column x rows rb + 32q
  const int e = warp - 20;
  const int qd = warp & 3, ph = (e >> 2) & 1;  // epilogue: TMEM lane quarter (hardware: warp % 4), pixel half
  const int co = qd * 32 + lane;               // epilogue: output channel = TMEM lane

  if (warp == 0) {
    // landing producer: done above
  } else if (warp == 1) {
    // ============================================================== output store issuer
    if (lane == 0) {
      const int esz = p.out_fp32 ? 4 : 2;
      const int rowb = p.out_ld * esz;              // bytes of one output row
      char* out = reinterpret_cast<char*>(p.out);
      for (int i = 0; i < ns; ++i) {
        const F3Geo g = f3_geo(st0 + i, M, W, split, need_low);
        const uint32_t b = (uint32_t)i & 1u;
        mbar_wait(&tail->o_ready[b], ((uint32_t)i >> 1) & 1u);
        const uint8_t* src = smem + L.o_off + b * L.o_bytes;
        if (p.pool) {
          f3_bulk_s2g(out + (long)g.low0 * rowb, src, (uint32_t)(g.nlow * rowb));
          if (p.pool_idx)
            f3_bulk_s2g(p.pool_idx + (long)g.low0 * p.Cout, src + 16 * rowb, (uint32_t)(g.nlow * p.Cout));
        } else if (split) {
          f3_bulk_s2g(out + (long)g.p0 * rowb, src, (uint32_t)(32 * rowb));
          f3_bulk_s2g(out + (long)g.p1 * rowb, src + 32 * rowb, (uint32_t)(32 * rowb));
        } else {
          f3_bulk_s2g(out + (long)g.p0 * rowb, src, (uint32_t)(g.nv * rowb));
        }
        f3_bulk_commit();
        f3_bulk_wait_read0();
        if (i < 12) CUNET_TRACE_MARK(trace, 200 + i);
        mbar_arrive(&tail->o_free[b]);
      }
    }
  } else if (warp == 2) {
    // ============================================================== MMA issuer: D[cout][px] = Wimg * A^T
    if (lane == 0 && ns > 0) {
      const uint32_t idesc = make_idesc(Elem<bf16>::FMT, 128, 64, 0, 0);
      // descriptors of K block 0 / K step 0; every other operand slice is a constant byte offset away
      const uint64_t wdesc0 = make_sdesc(smem_u32(smem + L.w_off), 16, 1024);
      const uint64_t adesc0 = make_sdesc(smem_u32(smem + L.a_off), 16, 1024);
      mbar_wait(&tail->w_full, 0);
      uint32_t buf = 0, bph = 0;
      for (int i = 0; i < ns; ++i) {
        mbar_wait(&tail->d_free[buf], bph ^ 1u);
        if (i < 12) CUNET_TRACE_MARK(trace, 96 + 2 * i);
        const uint32_t d = tmem + buf * 64u;
#pragma unroll 1
        for (int c = 0; c < nchunk; ++c) {
          mbar_wait(&tail->a_full[c], (uint32_t)i & 1u);
          tc_fence_after();
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int kb = 2 * c + h;
            if (kb >= nkb) break;
            const int nkk = min(4, (Cin - kb * 64 + 15) >> 4);   // 16-channel steps that hold real channels
            const uint64_t wd = sdesc_advance(wdesc0, (uint32_t)kb * wblk);
            const uint64_t ad = sdesc_advance(adesc0, (uint32_t)(kb * F3_SUB));
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              if (kk >= nkk) break;
              umma<bf16>(d, sdesc_advance(wd, kk * 32), sdesc_advance(ad, kk * 32), idesc, (uint32_t)((kb | kk) != 0));
            }
          }
          tc_commit(&tail->a_free[c]);
        }
        tc_commit(&tail->d_full[buf]);
        if (i < 12) CUNET_TRACE_MARK(trace, 97 + 2 * i);
        if (++buf == F3_DBUF) {
          buf = 0;
          bph ^= 1u;
        }
      }
    }
  } else if (is_tr) {
    // ============================================================== transformers (512 threads)
    const uint32_t ab = smem_u32(smem + L.a_off);
    int lowr[2];   // this thread's rows are the same in every stage: their half-resolution rows live in registers
#pragma unroll
    for (int q = 0; q < 2; ++q) lowr[q] = tail->lowmap[rb + 32 * q];
    // one packed word per chunk: bit 31 valid | bit 30 upsampled source | bits 16..17 log2(C/32) | bits 0..15 byte
    // offset of this thread's 16-byte column inside an x buffer
    uint32_t cs_pk[F3_MAXCH];
#pragma unroll
    for (int c = 0; c < F3_MAXCH; ++c) {
      cs_pk[c] = 0u;
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
    ActCoef<bf16> acf_q;              // QuanInput constants (cunet_concat.act_bits), computed once
    acf_q.set_quant(p.in.act_bits);
    for (int i = 0; i < ns; ++i) {
      const int nv = split ? F3_R : min(F3_R, M - (st0 + i) * F3_R);
      const uint32_t b = (uint32_t)i & 1u;
      if (t == 0 && i < 12) CUNET_TRACE_MARK(trace, 32 + 2 * i);
      mbar_wait(&tail->x_full[b], ((uint32_t)i >> 1) & 1u);
      if (t == 0 && i == 4) CUNET_TRACE_MARK(trace, 299);
      const uint32_t xb = smem_u32(smem + L.x_off + b * L.x_bytes);
#pragma unroll 1
      for (int c = 0; c < nchunk; ++c) {       // not unrolled (code size); the per-chunk word rotates through 3 registers
        ActCoef<bf16> acf;
        acf.sc = *reinterpret_cast<const uint4*>(&tail->sc2[(c * 128 + cc * 8) >> 1]);
        acf.sh = *reinterpret_cast<const uint4*>(&tail->sh2[(c * 128 + cc * 8) >> 1]);
        acf.relu = relu_on;
        acf.qmax2 = acf_q.qmax2;
        acf.magic2 = acf_q.magic2;
        if (t == 0 && i == 4) CUNET_TRACE_MARK(trace, 300 + 4 * c);
        mbar_wait(&tail->a_free[c], ((uint32_t)i & 1u) ^ 1u);   // MMAs of the previous stage have read this chunk
        if (t == 0 && i == 4) CUNET_TRACE_MARK(trace, 301 + 4 * c);
        const uint32_t pk = cs_pk[0];
        {
          const uint32_t t0 = cs_pk[0];
          if (nchunk == 3) { cs_pk[0] = cs_pk[1]; cs_pk[1] = cs_pk[2]; cs_pk[2] = t0; }
          else if (nchunk == 2) { cs_pk[0] = cs_pk[1]; cs_pk[1] = t0; }
        }
        const bool cvalid = (pk & 0x80000000u) != 0u, cup = (pk & 0x40000000u) != 0u;
        const uint32_t rx = xb + (pk & 0xFFFFu);
        const uint32_t lsh = 6u + ((pk >> 16) & 3u);      // log2(bytes per source row)
        // K block of this thread's 16-byte column: channels c*128 + cc*8 .. -> block 2c + (cc >> 3), chunk cc & 7
        const uint32_t abase = ab + (uint32_t)(2 * c + (cc >> 3)) * F3_SUB;
        if (c * 128 + (cc >> 3) * 64 < nkb * 64) {
          uint4 raw[2];
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const int r = rb + 32 * q;
            raw[q] = make_uint4(0, 0, 0, 0);
            if (cvalid && r < nv) raw[q] = f3_lds128(rx + ((uint32_t)(cup ? lowr[q] : r) << lsh));
          }
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const int r = rb + 32 * q;
            uint4 o = make_uint4(0, 0, 0, 0), lo_unused;
            if (cvalid && r < nv) o = acf.apply(raw[q], lo_unused);
            sts128(abase + tile_off(r, cc & 7), o);
          }
        }
        if (t == 0 && i == 4) CUNET_TRACE_MARK(trace, 302 + 4 * c);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tail->a_full[c]);
        if (t == 0 && i == 4) CUNET_TRACE_MARK(trace, 303 + 4 * c);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&tail->x_free[b]);   // this warp has read everything it needs of x buffer b
      if (t == 0 && i < 12) CUNET_TRACE_MARK(trace, 33 + 2 * i);
    }
  } else if (is_ep) {
    // ============================================================== epilogue (256 threads, thread = output channel)
    const bool do_stats = p.out_stats != nullptr && co < p.Cout;
    const bool co_ok = co < (p.out_fp32 ? p.out_ld : p.Cout);
    double s1 = 0., s2 = 0.;
    const uint32_t rowb = (uint32_t)(p.out_ld * (p.out_fp32 ? 4 : 2));
    uint32_t buf = 0, bph = 0;
    for (int i = 0; i < ns; ++i) {
      const int nv = split ? F3_R : min(F3_R, M - (st0 + i) * F3_R);
      const uint32_t b = (uint32_t)i & 1u;
      mbar_wait(&tail->d_full[buf], bph);
      mbar_wait(&tail->o_free[b], (((uint32_t)i >> 1) & 1u) ^ 1u);   // the store of two stages ago has read the buffer
      if (tid == 640 && i < 12) CUNET_TRACE_MARK(trace, 144 + 2 * i);
      tc_fence_after();
      const uint32_t tb = tmem + buf * 64u + ((uint32_t)(qd * 32) << 16);
      const uint32_t ob = smem_u32(smem + L.o_off + b * L.o_bytes);
      float f1 = 0.f, f2 = 0.f;
#pragma unroll 1
      for (int sub = 0; sub < 2; ++sub) {
      const int pq = 2 * ph + sub;     // pixel quarter handled in this pass (16 pixels / 4 pooling windows)
      float v[16];
      // pooling: the two 8-column TMEM windows that hold the 16 pixels of this pass's 4 pooling windows
      uint32_t upA, upB;
      if (W >= 32) {
        upA = (uint32_t)(8 * pq);
        upB = (uint32_t)(32 + 8 * pq);
      } else if (W == 16) {
        upA = (uint32_t)(32 * (pq >> 1) + 8 * (pq & 1));
        upB = upA + 16u;
      } else {
        upA = (uint32_t)(16 * pq);
        upB = upA + 8u;
      }
      if (!p.pool) {
        const int r0 = 16 * pq;
        const int nvl = nv - r0;
        if (nvl > 0) {
          f3_tmem_ld16_nowait(tb + (uint32_t)r0, v);
          f3_tmem_wait_ld();
          if (co_ok) {
            const uint32_t a0 = ob + (uint32_t)r0 * rowb + (uint32_t)co * (p.out_fp32 ? 4u : 2u);
            if (p.out_fp32) {
#pragma unroll
              for (int q = 0; q < 16; ++q)
                if (q < nvl) f3_sts_f32(a0 + (uint32_t)q * rowb, v[q]);
            } else if (nvl >= 16 && rowb == 256u) {
              f3_ep16<256>(a0, v, f1, f2);
            } else if (nvl >= 16 && rowb == 64u) {
              f3_ep16<64>(a0, v, f1, f2);
            } else {
#pragma unroll
              for (int q = 0; q < 16; ++q) {
                if (q < nvl) {
                  const uint16_t h = __bfloat16_as_ushort(__float2bfloat16_rn(v[q]));
                  const float r = __uint_as_float((uint32_t)h << 16);
                  f1 += r;
                  f2 = fmaf(r, r, f2);
                  f3_sts_u16(a0 + (uint32_t)q * rowb, h);
                }
              }
            }
          }
        }
      } else {
        // 2x2 max-pool (nn.MaxPool2d(2, 2): first maximum in row-major window order wins): this thread's 4 windows
        const int nlow = split ? 16 : (nv >> 2);
        const int l0 = 4 * pq;
        if (l0 < nlow) {
          f3_tmem_ld8_nowait(tb + upA, v);
          f3_tmem_ld8_nowait(tb + upB, v + 8);
          f3_tmem_wait_ld();
          if (co_ok) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float c0, c1, c2, c3;
              if (W != 4) {
                c0 = v[2 * j]; c1 = v[2 * j + 1]; c2 = v[8 + 2 * j]; c3 = v[9 + 2 * j];
              } else {          // a 4 x 4 image per 16 columns: rows of 4
                const int base = (j >> 1) * 8 + (j & 1) * 2;
                c0 = v[base]; c1 = v[base + 1]; c2 = v[base + 4]; c3 = v[base + 5];
              }
              float m = c0;
              uint32_t k = 0;
              if (c1 > m) { m = c1; k = 1; }
              if (c2 > m) { m = c2; k = 2; }
              if (c3 > m) { m = c3; k = 3; }
              if (l0 + j < nlow) {
                const uint16_t h = __bfloat16_as_ushort(__float2bfloat16_rn(m));
                const float r = __uint_as_float((uint32_t)h << 16);
                f1 += r;
                f2 = fmaf(r, r, f2);
                f3_sts_u16(ob + (uint32_t)(l0 + j) * rowb + (uint32_t)co * 2u, h);
                f3_sts_u8(ob + 16u * rowb + (uint32_t)(l0 + j) * (uint32_t)p.Cout + (uint32_t)co, k);
              }
            }
          }
        }
      }
      }   // sub
      s1 += (double)f1;
      s2 += (double)f2;
      fence_proxy_async();  // staging block -> visible to the bulk store
      tc_fence_before();
      __syncwarp();
      if (tid == 640 && i < 12) CUNET_TRACE_MARK(trace, 145 + 2 * i);
      if (lane == 0) {
        mbar_arrive(&tail->o_ready[b]);
        mbar_arrive(&tail->d_free[buf]);
      }
      if (++buf == F3_DBUF) {
        buf = 0;
        bph ^= 1u;
      }
    }
    if (do_stats && ns > 0) {
      atomicAdd(p.out_stats + co, s1);
      atomicAdd(p.out_stats + p.Cout + co, s2);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem, 256);
}

}  // namespace cunet
using namespace cunet;

CUNET_TRACE_SETTER(cunet_debug_trace_fwd_v3, g_f3_trace)

// pixel rows from which a bf16 1x1 forward runs in this kernel (default: always -- it is at least as fast as the
// generation-1 kernel at every size of the network, and whole-step A/B runs are 0.35 ms faster with it everywhere than
// with it on the 64x64 maps only); CUNET_FWD_V3_MIN_ROWS overrides, a negative value disables
static long g_fwd_v3_min_rows = [] {
  const char* e = getenv("CUNET_FWD_V3_MIN_ROWS");
  return e ? atol(e) : 0L;
}();

// Returns 1 when this kernel handled the call, 0 when the caller must use another kernel, <0 on error.
int cunet_conv_fwd_v3_try(const cunet_conv_fwd_params* p, cudaStream_t st) {
  if (g_fwd_v3_min_rows < 0) return 0;
  if (p->dtype != CUNET_BF16 || p->taps != 1) return 0;
  if (p->in.bn_train < 0 || p->in.bn_train > 2) return 0;     // 2: identity input (the stem's im2col blocks)
  if (p->CoutPad % 16 || p->CoutPad < 16 || p->CoutPad > 128) return 0;
  if (p->out_fp32) {
    if (p->pool || p->out_ld % 4 || p->out_ld > p->CoutPad) return 0;
  } else {
    if (p->Cout != p->CoutPad || p->out_ld != p->Cout || (p->Cout & 7)) return 0;
  }
  if (p->pool && !p->pool_idx) return 0;
  int cin = 0, up = 0;
  for (int s = 0; s < p->in.nseg; ++s) {
    const cunet_seg& sg = p->in.seg[s];
    if (sg.ld != sg.C || (sg.C != 32 && sg.C != 64 && sg.C != 128)) return 0;
    if ((cin >> 7) != ((cin + sg.C - 1) >> 7)) return 0;   // a segment must not straddle a 128-channel chunk
    cin += sg.C;
    up |= sg.up;
  }
  if (cin > MAX_CIN || cin > 128 * F3_MAXCH) return 0;
  const int W = p->W, H = p->H;
  if ((W & (W - 1)) || (H & (H - 1)) || W > 64 || W < 4 || H < 4) return 0;
  if (up && p->pool) return 0;
  const long M = (long)p->N * H * W;
  if (M <= 0) return 1;
  if (M > (1L << 30)) return 0;
  if (M < g_fwd_v3_min_rows) return 0;
  if ((up || p->pool) && (M % 64) && ((M % 64) % (2 * W))) return 0;
  F3Layout L;
  L.split = (p->pool && W == 64) ? 1 : 0;
  L.low_rows = (W == 64 && !L.split) ? 32 : 16;
  int xbytes = 0;
  for (int s = 0; s < p->in.nseg; ++s) xbytes += (p->in.seg[s].up ? L.low_rows : F3_R) * p->in.seg[s].C * 2;
  const int nkb = (cin + 63) / 64;
  L.w_off = 0;
  // a short last K block / CoutPad < 128 is read as a full 128-row operand: keep readable bytes behind the image
  L.a_off = (nkb * p->CoutPad * 128 + 1023) & ~1023;
  L.x_off = L.a_off + nkb * F3_SUB;
  L.x_bytes = (xbytes + 1023) & ~1023;
  if (L.x_bytes < 8192) L.x_bytes = 8192;       // the BatchNorm coefficient table lives there during the prologue
  L.o_off = L.x_off + 2 * L.x_bytes;
  const int rowb = p->out_ld * (p->out_fp32 ? 4 : 2);
  L.o_bytes = p->pool ? 16 * rowb + 16 * p->Cout : F3_R * rowb;
  L.o_bytes = (L.o_bytes + 1023) & ~1023;
  L.tail_off = L.o_off + 2 * L.o_bytes;
  const size_t smem = (size_t)L.tail_off + sizeof(F3Tail) + 1024;
  if (smem > 232448) return 0;
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  const int total = (int)((M + F3_R - 1) / F3_R);
  L.per = (total + sms - 1) / sms;
  const int grid = (total + L.per - 1) / L.per;
  cudaError_t e = cudaFuncSetAttribute(conv_fwd_v3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
  if (e != cudaSuccess) return cunet_fail_cuda("conv_fwd_v3 attr", e);
  e = cunet_launch(conv_fwd_v3_kernel, dim3(grid), dim3(F3_THREADS), smem, st, *p, L);
  if (e != cudaSuccess) return cunet_fail_cuda("conv_fwd_v3 launch", e);
  return 1;
}

// debug / experiment switch (tests): route bf16 1x1 forward calls with at least min_rows pixel rows to this kernel;
// negative disables.  Returns the previous setting.
extern "C" long cunet_debug_fwd_v3_min_rows(long min_rows) {
  const long old = g_fwd_v3_min_rows;
  g_fwd_v3_min_rows = min_rows;
  return old;
}
