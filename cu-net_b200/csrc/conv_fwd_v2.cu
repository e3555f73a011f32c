// Forward of the fused 1x1 conv, second generation (bf16): persistent, bulk-landed sources, resident weights.
// Same contract as conv_fwd.cu for taps == 1 without pooling (cunet_conv_fwd_params); see that file for the math.
//
// The round-1 kernel is one CTA per 128-pixel tile: every CTA pays the prologue (BatchNorm coefficients, TMEM
// allocation, weight fetch) and gathers its A operand with per-thread 16-byte global loads, prefetch depth 1
// (ncu, 320 -> 128 at 64x64 / batch 24: 52 us, issue-slot utilisation 44 %, long-scoreboard bound, 10 % of DRAM peak).
// Here:
//   * a 128-pixel raster tile of every source tensor of the virtual concat is a CONTIGUOUS block (an upsampled
//     source contributes its 32 half-resolution pixels): the producer lands it with 1-D TMA bulk copies into a ring of
//     16 KB slots, up to three jobs ahead of the transformers;
//   * 8 transformer warps apply BatchNorm + ReLU smem -> smem and write the K-major SWIZZLE_128B operand
//     (nearest x2 upsampling = index math on the landed block);
//   * the weight image (<= 80 KB) is fetched once per CTA and stays resident for its whole tile list;
//   * D[128 px][CoutPad] accumulates in a double-buffered TMEM region, so the MMAs of tile i+1 overlap the epilogue of
//     tile i; the epilogue (thread = pixel) rounds, stores each pixel's channel row, and reduces the per-channel
//     sum / sum-of-squares over the warp with a transpose-reduce (31 shuffles per 32 channels), accumulating in fp64
//     registers across the CTA's tiles (one atomic per channel per warp at the end).
#include "loaders.cuh"
#include "host_util.h"
#include <stdlib.h>

namespace cunet {

constexpr int F2_THREADS = 512;  // warp 0 landing producer | 1 MMA | 2-3 idle | 4-11 transformers | 12-15 epilogue
constexpr int F2_NSLOT = 3;
constexpr int F2_SLOT = 16384;

struct F2Tail {
  uint64_t w_full, a_ready, a_free;
  uint64_t slot_full[F2_NSLOT], slot_free[F2_NSLOT], acc_full[2], acc_free[2];
  uint32_t tmem_base;
  int relu;
  int seg_start[CUNET_MAX_SEG + 1];
  alignas(16) uint32_t sc2[MAX_CIN / 2];  // bf16x2 BatchNorm scale / shift of the concat (copied out of the prologue area)
  alignas(16) uint32_t sh2[MAX_CIN / 2];
};

__device__ __forceinline__ uint4 f2_lds128(uint32_t saddr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr));
  return v;
}

// landing jobs of one tile (identical in the producer and the transformers): per segment, one job per 16 KB of rows
__device__ __forceinline__ int f2_jobs_of_seg(const cunet_seg& sg) { return (!sg.up && sg.C == 128) ? 2 : 1; }

// timeline slots (CUNET_TRACE builds), tile i < 16: producer 0+i (tile's first landing issued), transformer 16+3i (start,
// operand free, all jobs transformed), MMA 64+2i (operand ready + accumulator free, issued), epilogue 96+2i (accumulator
// full, tile done)
CUNET_TRACE_DECL(g_f2_trace)

__global__ void __launch_bounds__(F2_THREADS, 1) conv_fwd_v2_kernel(const __grid_constant__ cunet_conv_fwd_params p,
                                                                     int ntiles, int a_off, int raw_off, int tail_off) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  F2Tail* tail = reinterpret_cast<F2Tail*>(smem + tail_off);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int Cin = concat_cin(p.in);
  const int nkb = (Cin + 63) >> 6;
  const int M = p.N * p.H * p.W;
  const int W = p.W, lw = 31 - __clz(p.W);
  const int tile0 = (int)blockIdx.x, tstride = (int)gridDim.x;
  const uint32_t wblk = (uint32_t)p.CoutPad * 128u;  // bytes of one K block of the weight image
  CUNET_TRACE_LOAD(trace, g_f2_trace)

  if (tid == 0) {
    mbar_init(&tail->w_full, 1);
    mbar_init(&tail->a_ready, 8);
    mbar_init(&tail->a_free, 1);
    for (int s = 0; s < F2_NSLOT; ++s) {
      mbar_init(&tail->slot_full[s], 1);
      mbar_init(&tail->slot_free[s], 8);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tail->acc_full[b], 1);
      mbar_init(&tail->acc_free[b], 4);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(&tail->tmem_base, 256);
  {
    // operand buffer zeroed once: chunk columns past Cin (a 160-channel concat fills 2.5 K blocks) are never written
    const uint32_t a0 = smem_u32(smem + a_off);
    for (int i = tid; i < nkb * 1024; i += F2_THREADS) sts128(a0 + (uint32_t)i * 16u, make_uint4(0, 0, 0, 0));
    fence_proxy_async();
  }
  griddep_wait();    // everything above overlaps the tail of the previous kernel (programmatic dependent launch)
  griddep_launch();
  {
    // BatchNorm coefficients: computed in the (not yet used) landing area, the packed pair copied to the tail
    BnSmem* bn = reinterpret_cast<BnSmem*>(smem + raw_off);
    compute_bn_coefs(p.in, bn, nkb * 64, tid, F2_THREADS);
    __syncthreads();
    for (int i = tid; i < nkb * 32; i += F2_THREADS) {
      tail->sc2[i] = bn->sc2[i];
      tail->sh2[i] = bn->sh2[i];
    }
    if (tid <= CUNET_MAX_SEG) tail->seg_start[tid] = bn->seg_start[tid];
    if (tid == 0) tail->relu = bn->relu;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tail->tmem_base;

  if (warp == 0) {
    // ============================================================== landing producer
    if (lane == 0 && tile0 < ntiles) {
      mbar_arrive_expect_tx(&tail->w_full, (uint32_t)nkb * wblk);
      bulk_g2s(smem, p.wpack, (uint32_t)nkb * wblk, &tail->w_full);
      int slot = 0;
      uint32_t ph = 0, ti = 0;
      for (int tile = tile0; tile < ntiles; tile += tstride, ++ti) {
        const int m0 = tile * 128, nv = min(128, M - m0);
        if (ti < 16) CUNET_TRACE_MARK(trace, 0 + ti);
        for (int s = 0; s < p.in.nseg; ++s) {
          const cunet_seg& sg = p.in.seg[s];
          const int nj = f2_jobs_of_seg(sg);
          for (int j = 0; j < nj; ++j) {
            // rows of the source block this job brings in
            long src_row;
            int nrows;
            if (sg.up) {
              src_row = (long)tile * 32;
              nrows = nv >> 2;
            } else {
              const int r0 = j * 64, r1 = nj == 2 ? min(nv, r0 + 64) : nv;
              src_row = (long)m0 + r0;
              nrows = max(0, r1 - r0);
            }
            const uint32_t bytes = (uint32_t)(nrows * sg.C * 2);
            mbar_wait(&tail->slot_free[slot], ph ^ 1u);
            if (bytes) {
              mbar_arrive_expect_tx(&tail->slot_full[slot], bytes);
              bulk_g2s(smem + raw_off + slot * F2_SLOT, reinterpret_cast<const char*>(sg.ptr) + src_row * sg.C * 2, bytes,
                       &tail->slot_full[slot]);
            } else {
              mbar_arrive(&tail->slot_full[slot]);
            }
            if (++slot == F2_NSLOT) {
              slot = 0;
              ph ^= 1u;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ============================================================== MMA issuer
    if (lane == 0 && tile0 < ntiles) {
      const uint32_t idesc = make_idesc(Elem<bf16>::FMT, 128, (uint32_t)p.CoutPad, 0, 0);
      const uint32_t wB = smem_u32(smem), aA = smem_u32(smem + a_off);
      mbar_wait(&tail->w_full, 0);
      uint32_t i = 0;
      for (int tile = tile0; tile < ntiles; tile += tstride, ++i) {
        const uint32_t b = i & 1u;
        mbar_wait(&tail->a_ready, i & 1u);
        mbar_wait(&tail->acc_free[b], ((i >> 1) & 1u) ^ 1u);
        if (i < 16) CUNET_TRACE_MARK(trace, 64 + 2 * i);
        tc_fence_after();
        for (int kb = 0; kb < nkb; ++kb) {
          const int nkk = min(4, (Cin - kb * 64 + 15) >> 4);  // 16-channel steps that hold real channels
          for (int kk = 0; kk < nkk; ++kk)
            umma<bf16>(tmem + b * 128u, make_sdesc(aA + kb * 16384 + kk * 32, 16, 1024),
                       make_sdesc(wB + kb * wblk + kk * 32, 16, 1024), idesc, (uint32_t)((kb | kk) != 0));
        }
        tc_commit(&tail->a_free);
        tc_commit(&tail->acc_full[b]);
        if (i < 16) CUNET_TRACE_MARK(trace, 65 + 2 * i);
      }
    }
  } else if (warp >= 4 && warp < 12) {
    // ============================================================== transformers (256 threads)
    const int t = tid - 128;
    const uint32_t abase = smem_u32(smem + a_off);
    int slot = 0;
    uint32_t ph = 0, i = 0;
    for (int tile = tile0; tile < ntiles; tile += tstride, ++i) {
      const int m0 = tile * 128, nv = min(128, M - m0);
      if (t == 0 && i < 16) CUNET_TRACE_MARK(trace, 16 + 3 * i);
      mbar_wait(&tail->a_free, (i & 1u) ^ 1u);  // MMAs of the previous tile no longer read the operand
      if (t == 0 && i < 16) CUNET_TRACE_MARK(trace, 17 + 3 * i);
      for (int s = 0; s < p.in.nseg; ++s) {
        const cunet_seg& sg = p.in.seg[s];
        const int nj = f2_jobs_of_seg(sg);
        const int lc = sg.C == 128 ? 4 : (sg.C == 64 ? 3 : 2);  // log2(16-byte chunks per source row)
        const int cc = t & ((1 << lc) - 1), rstep = 256 >> lc;
        const int col = tail->seg_start[s] + cc * 8;              // concat channel of this thread's chunk
        ActCoef<bf16> acf;
        acf.sc = *reinterpret_cast<const uint4*>(&tail->sc2[col >> 1]);
        acf.sh = *reinterpret_cast<const uint4*>(&tail->sh2[col >> 1]);
        acf.relu = tail->relu;
        const uint32_t dstb = abase + (uint32_t)(col >> 6) * 16384u;
        const int dchunk = (col >> 3) & 7;
        const uint32_t pitch = (uint32_t)sg.C * 2u;
        for (int j = 0; j < nj; ++j) {
          const int r0 = nj == 2 ? j * 64 : 0, r1 = nj == 2 ? r0 + 64 : 128;
          mbar_wait(&tail->slot_full[slot], ph);
          const uint32_t rawb = smem_u32(smem + raw_off + slot * F2_SLOT) + (uint32_t)cc * 16u;
          for (int r = r0 + (t >> lc); r < r1; r += rstep) {
            uint4 o = make_uint4(0, 0, 0, 0), lo_unused;
            if (r < nv) {
              int srow = r - r0;
              if (sg.up) srow = ((r >> lw) >> 1) * (W >> 1) + ((r & (W - 1)) >> 1);
              o = acf.apply(f2_lds128(rawb + (uint32_t)srow * pitch), lo_unused);
            }
            sts128(dstb + tile_off(r, dchunk), o);
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&tail->slot_free[slot]);
          if (++slot == F2_NSLOT) {
            slot = 0;
            ph ^= 1u;
          }
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (t == 0 && i < 16) CUNET_TRACE_MARK(trace, 18 + 3 * i);
      if (lane == 0) mbar_arrive(&tail->a_ready);
    }
  } else if (warp >= 12) {
    // ============================================================== epilogue (128 threads, thread = tile pixel)
    const int qd = warp & 3;  // TMEM lane quarter (hardware: warp % 4)
    const int row = qd * 32 + lane;
    const bool do_stats = p.out_stats != nullptr;
    double s1[4] = {0., 0., 0., 0.}, s2[4] = {0., 0., 0., 0.};  // lane l, group g: channel g*32 + l
    const int ngrp = p.CoutPad >> 5, rem8 = (p.CoutPad & 31) >> 3;  // full 32-column groups, leftover 8-column chunks
    uint32_t i = 0;
    for (int tile = tile0; tile < ntiles; tile += tstride, ++i) {
      const uint32_t b = i & 1u;
      const int m0 = tile * 128, nv = min(128, M - m0);
      const bool valid = row < nv;
      const long grow = (long)m0 + row;
      mbar_wait(&tail->acc_full[b], (i >> 1) & 1u);
      if (tid == 384 && i < 16) CUNET_TRACE_MARK(trace, 96 + 2 * i);
      tc_fence_after();
      const uint32_t tb = tmem + b * 128u + ((uint32_t)(qd * 32) << 16);
      if (p.out_fp32) {
        // heat-map heads: fp32 rows of out_ld floats, no statistics
        float* orow = reinterpret_cast<float*>(p.out) + grow * p.out_ld;
        for (int c8 = 0; c8 < (p.CoutPad >> 3); ++c8) {
          float v[8];
          tmem_ld8(tb + (uint32_t)(c8 * 8), v);
          if (valid && c8 * 8 < p.out_ld) {
            *reinterpret_cast<float4*>(orow + c8 * 8) = make_float4(v[0], v[1], v[2], v[3]);
            *reinterpret_cast<float4*>(orow + c8 * 8 + 4) = make_float4(v[4], v[5], v[6], v[7]);
          }
        }
      } else {
        char* orow = reinterpret_cast<char*>(p.out) + grow * p.out_ld * 2;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          if (g >= ngrp) break;
          float o[32];
#pragma unroll
          for (int c8 = 0; c8 < 4; ++c8) tmem_ld8(tb + (uint32_t)(g * 32 + c8 * 8), o + c8 * 8);
          uint32_t pk[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            __nv_bfloat162 h = __floats2bfloat162_rn(o[2 * j], o[2 * j + 1]);
            pk[j] = *reinterpret_cast<uint32_t*>(&h);
            // statistics of the values as stored; rows past the end contribute nothing
            o[2 * j] = valid ? __uint_as_float(pk[j] << 16) : 0.f;
            o[2 * j + 1] = valid ? __uint_as_float(pk[j] & 0xFFFF0000u) : 0.f;
          }
          if (valid) {
            uint4* dst = reinterpret_cast<uint4*>(orow + g * 64);
            dst[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            dst[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
            dst[2] = make_uint4(pk[8], pk[9], pk[10], pk[11]);
            dst[3] = make_uint4(pk[12], pk[13], pk[14], pk[15]);
          }
          if (do_stats) {
            // transpose-reduce over the warp: after 5 halving steps lane l holds the 32-pixel sum of channel g*32 + l
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
            s1[g] += (double)o[0];
            s2[g] += (double)q2[0];
          }
        }
        // leftover 8-column chunks (CoutPad not a multiple of 32; host: no statistics in that case)
        for (int c8 = 0; c8 < rem8; ++c8) {
          float v[8];
          tmem_ld8(tb + (uint32_t)(ngrp * 32 + c8 * 8), v);
          if (valid) {
            uint32_t w4[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
              w4[j] = *reinterpret_cast<uint32_t*>(&h);
            }
            *reinterpret_cast<uint4*>(orow + (ngrp * 32 + c8 * 8) * 2) = make_uint4(w4[0], w4[1], w4[2], w4[3]);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (tid == 384 && i < 16) CUNET_TRACE_MARK(trace, 97 + 2 * i);
      if (lane == 0) mbar_arrive(&tail->acc_free[b]);
    }
    if (do_stats && tile0 < ntiles) {
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        if (g >= ngrp) break;
        atomicAdd(p.out_stats + g * 32 + lane, s1[g]);
        atomicAdd(p.out_stats + p.Cout + g * 32 + lane, s2[g]);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 256);
}

}  // namespace cunet
using namespace cunet;

CUNET_TRACE_SETTER(cunet_debug_trace_fwd_v2, g_f2_trace)

// Opt-in for now: on the bench shapes it matches the round-1 kernel but does not beat it yet, so
// cunet_conv_fwd keeps the round-1 kernel unless CUNET_FWD_V2_MIN_TILES >= 0 (or the debug setter) enables this one
// for calls with at least that many 128-pixel tiles.
static int g_fwd_v2_min_tiles = [] {
  const char* e = getenv("CUNET_FWD_V2_MIN_TILES");
  return e ? atoi(e) : -1;
}();
extern "C" int cunet_debug_fwd_v2_min_tiles(int min_tiles) {
  const int old = g_fwd_v2_min_tiles;
  g_fwd_v2_min_tiles = min_tiles;
  return old;
}

// Returns 1 when this kernel handled the call, 0 when the caller must use the generic kernel, <0 on error.
int cunet_conv_fwd_v2_try(const cunet_conv_fwd_params* p, cudaStream_t st) {
  const int min_tiles = g_fwd_v2_min_tiles;
  if (min_tiles < 0) return 0;
  if (p->dtype != CUNET_BF16 || p->taps != 1 || p->pool) return 0;
  if (p->in.bn_train == 2) return 0;  // identity (im2col) input of the stem: generic kernel
  if (p->in.act_bits) return 0;       // activation-quantized operand: the other kernels
  if (p->CoutPad % 16 || p->CoutPad < 16 || p->CoutPad > 128) return 0;
  if (p->out_fp32) {
    if (p->out_ld % 8 || p->out_ld > p->CoutPad) return 0;
  } else {
    if (p->Cout != p->CoutPad || p->out_ld != p->Cout || (p->Cout & 7)) return 0;
    if (p->out_stats && (p->Cout & 31)) return 0;
  }
  int cin = 0, up = 0;
  for (int s = 0; s < p->in.nseg; ++s) {
    const cunet_seg& sg = p->in.seg[s];
    if (sg.ld != sg.C || (sg.C != 32 && sg.C != 64 && sg.C != 128)) return 0;
    if ((cin >> 6) != ((cin + sg.C - 1) >> 6) && (cin & 63)) return 0;  // a segment starts on a 64-boundary or fits its block
    cin += sg.C;
    up |= sg.up;
  }
  if (cin > MAX_CIN || cin > 320) return 0;
  const int W = p->W, H = p->H;
  if (up && ((W & (W - 1)) || (H & 1) || W > 64 || W < 2)) return 0;
  const long M = (long)p->N * H * W;
  if (M <= 0) return 1;
  if (M > (1L << 30)) return 0;
  if (up && (M % (2 * W))) return 0;
  const int ntiles = (int)((M + 127) / 128);
  if (ntiles < min_tiles) return 0;
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  const int nkb = (cin + 63) / 64;
  const int a_off = ((nkb * p->CoutPad * 128) + 1023) & ~1023;
  const int raw_off = a_off + nkb * 16384;
  const int tail_off = raw_off + F2_NSLOT * F2_SLOT;
  const size_t smem = (size_t)tail_off + sizeof(F2Tail) + 1024;
  if (smem > 232448) return 0;
  const int grid = ntiles < sms ? ntiles : sms;
  cudaError_t e = cudaFuncSetAttribute(conv_fwd_v2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
  if (e != cudaSuccess) return cunet_fail_cuda("conv_fwd_v2 attr", e);
  e = cunet_launch(conv_fwd_v2_kernel, dim3(grid), dim3(F2_THREADS), smem, st, *p, ntiles, a_off, raw_off, tail_off);
  if (e != cudaSuccess) return cunet_fail_cuda("conv_fwd_v2 launch", e);
  return 1;
}
