// Backward-data of the fused conv, sm_100a.  Replaces what autograd does for one
// cat -> BatchNorm -> ReLU -> conv call of the reference (models/cu_net.py:11-17):
// conv backward-data, ReLU backward, the per-op BatchNorm parameter gradients, and the split of the concat
// gradient back onto the (shared, multi-consumer) source tensors, including the 2x2 sum that is the
// backward of nearest-x2 upsampling.
//
// GEMM view (per CTA: one 128-pixel tile x one chunk of 128 concat channels):
//   dA[128 px][128 k] = dY[128 px][Kg] * Wt[128 k][Kg]^T,   Kg = taps * Cout (tap-major)
//   A operand : the gradient of the conv output, evaluated on the fly as p*G + q*T + r (batch-norm
//               backward form of the consumer chain, see cunet_grad_src), 3x3: gathered at px - tap.
//   B operand : "dgrad image" of the weights, 1-D TMA bulk copy per K block.
// Epilogue (per source segment of the concat):  x = source value, z = scale*x + shift,
//   dz = dA * [z > 0];  dbeta += sum dz;  dgamma += sum dz * xhat;  G_src (+)= gamma * dz
//   (for an upsampled source the four children are summed first), and for the last consumer of a
//   source the per-channel (sum G, sum G*xhat) that the producer's backward needs.
#include "loaders.cuh"
#include "host_util.h"
#include <stdlib.h>

namespace cunet {

constexpr int DG_STAGES = 3;
constexpr int DG_THREADS = 320;

struct DgSmemTail {
  uint64_t full[DG_STAGES];
  uint64_t empty[DG_STAGES];
  uint64_t accum;
  uint32_t tmem_base;
  BnSmem bn;
  GradSmem gc;
  TileRowTable rows;
};

#ifndef DG_MIN_CTAS
#define DG_MIN_CTAS StageGeom<T>::MIN_CTAS
#endif
template <typename T>
__global__ void __launch_bounds__(DG_THREADS, DG_MIN_CTAS) conv_dgrad_kernel(const __grid_constant__ cunet_conv_dgrad_params p) {
  using E = Elem<T>;
  using SG = StageGeom<T>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  DgSmemTail* tail = reinterpret_cast<DgSmemTail*>(smem + DG_STAGES * SG::BYTES);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tile = blockIdx.x, chunk = blockIdx.y;
  const int Cin = concat_cin(p.in);
  int grouped = 0;
  for (int s = 0; s < p.in.nseg; ++s) grouped |= p.in.seg[s].up;

  const int coutk = p.taps == 9 ? p.Cout : p.CoutPad;
  const int Kg = p.taps * coutk;
  const int nsteps = (Kg + E::KBE - 1) / E::KBE;

  if (tid == 0) {
    for (int s = 0; s < DG_STAGES; ++s) {
      mbar_init(&tail->full[s], 9);
      mbar_init(&tail->empty[s], 1);
    }
    mbar_init(&tail->accum, 1);
    fence_mbar_init();
  }
  if (warp == 9) tmem_alloc(&tail->tmem_base, 128);
  compute_bn_coefs(p.in, &tail->bn, ((Cin + 127) / 128) * 128, tid, DG_THREADS);
  compute_grad_coefs(p.dy, &tail->gc, tid, DG_THREADS);
  PixGeom geom;
  geom.N = p.N; geom.H = p.H; geom.W = p.W; geom.M = p.N * p.H * p.W;
  tile_rows_init(&tail->rows, geom, tile, grouped, tid);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tail->tmem_base;

  if (warp < 8) {
    // ============================================================== A loaders (gradient operand)
    const int c = tid & 7;
    RowCtx rc;
    uint32_t soff[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int r = (tid >> 3) + 32 * q;
      rc.rd[q] = tail->rows.rd[r];
      rc.ru[q] = tail->rows.ru[r];
      rc.hw[q] = tail->rows.hw[r];
      soff[q] = tile_off(r, c);
    }
    GradRaw<T> cur[4], nxt[4];
    uint32_t cmask = 0, nmask = 0;
    int cco = 0, nco = 0;

    auto issue = [&](int it, GradRaw<T>* dst, uint32_t& mask, int& co_out) {
      mask = 0;
      const int kg = it * E::KBE + c * E::EPC;
      const int tap = kg / coutk, co = kg - tap * coutk;
      co_out = co;
      if (tap >= p.taps) return;
      int dy = 0, dx = 0;
      if (p.taps == 9) {
        dy = tap / 3 - 1;
        dx = tap - (tap / 3) * 3 - 1;
      }
      // out(px) reads in(px + off)  =>  in(px) receives from out(px - off)
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (grad_load<T>(p.dy, rc, q, p.H, p.W, co, -dy, -dx, dst[q])) mask |= 1u << q;
    };

    issue(0, cur, cmask, cco);
    for (int it = 0; it < nsteps; ++it) {
      const int s = it % DG_STAGES;
      const uint32_t ph = (it / DG_STAGES) & 1;
      if (it + 1 < nsteps) issue(it + 1, nxt, nmask, nco);
      GradCoef<T> cf;
      cf.load(&tail->gc, cco & 127);
      mbar_wait(&tail->empty[s], ph ^ 1);
      const uint32_t abase = smem_u32(smem + s * SG::BYTES);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint4 o = make_uint4(0, 0, 0, 0), lo = make_uint4(0, 0, 0, 0);
        if ((cmask >> q) & 1) o = cf.apply(p.dy, cur[q], lo);
        sts128(abase + soff[q], o);
        if (SG::SPLIT) sts128(abase + SG::A_LO + soff[q], lo);
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tail->full[s]);
#pragma unroll
      for (int q = 0; q < 4; ++q) cur[q] = nxt[q];
      cmask = nmask;
      cco = nco;
    }
  } else if (warp == 8) {
    if (lane == 0) {
      const uint32_t bbytes = 128u * 128u * (SG::SPLIT ? 2u : 1u);
      const char* w = reinterpret_cast<const char*>(p.wpack_dgrad) + (size_t)chunk * nsteps * bbytes;
      for (int it = 0; it < nsteps; ++it) {
        const int s = it % DG_STAGES;
        const uint32_t ph = (it / DG_STAGES) & 1;
        mbar_wait(&tail->empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&tail->full[s], bbytes);
        bulk_g2s(smem + s * SG::BYTES + SG::B_OFF, w + (size_t)it * bbytes, bbytes, &tail->full[s]);
      }
    }
  } else {
    if (lane == 0) {
      const uint32_t idesc = make_idesc(E::FMT, 128, 128, 0, 0);
      for (int it = 0; it < nsteps; ++it) {
        const int s = it % DG_STAGES;
        const uint32_t ph = (it / DG_STAGES) & 1;
        mbar_wait(&tail->full[s], ph);
        tc_fence_after();
        const uint32_t a = smem_u32(smem + s * SG::BYTES);
        const uint32_t b = a + SG::B_OFF;
        const uint32_t alo = a + SG::A_LO, blo = b + 16384u;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          umma<T>(tmem, make_sdesc(a + kk * 32, 16, 1024), make_sdesc(b + kk * 32, 16, 1024), idesc,
                  (uint32_t)((it | kk) != 0));
          if (SG::SPLIT) {
            umma<T>(tmem, make_sdesc(alo + kk * 32, 16, 1024), make_sdesc(b + kk * 32, 16, 1024), idesc, 1u);
            umma<T>(tmem, make_sdesc(a + kk * 32, 16, 1024), make_sdesc(blo + kk * 32, 16, 1024), idesc, 1u);
          }
        }
        tc_commit(&tail->empty[s]);
      }
      tc_commit(&tail->accum);
    }
  }

  // ================================================================== epilogue
  constexpr int LD_EP = 132;
  float* ep = reinterpret_cast<float*>(smem);
  float* red = reinterpret_cast<float*>(smem + 128 * LD_EP * 4);  // [256][16] partial sums

  // thread -> channel quad (32 quads per chunk) x 8 row groups.  Each quad lies inside one segment.
  float a_db[4] = {0, 0, 0, 0}, a_dg[4] = {0, 0, 0, 0};
  const int quad = tid & 31;
  const int rg = tid >> 5;  // 0..7 for the epilogue warps
  const int k0 = chunk * 128 + quad * 4;
  bool active = warp < 8 && k0 < Cin;
  int sidx = 0;
  if (active) {
    while (k0 >= tail->bn.seg_start[sidx + 1]) ++sidx;
    if (p.gacc[sidx].G == nullptr) active = false;
  }
  const cunet_seg& sg = p.in.seg[sidx];
  const cunet_gacc& ga = p.gacc[sidx];
  const int cl = k0 - tail->bn.seg_start[sidx];
  const T* src = reinterpret_cast<const T*>(sg.ptr);
  T* G = reinterpret_cast<T*>(ga.G);
  const int nrow_it = (active && sg.up) ? 4 : 16;  // up: 4 windows per thread, else 16 rows per thread

  if (warp < 8) {
    mbar_wait(&tail->accum, 0);
    tc_fence_after();
    const int lq = warp & 3, half = warp >> 2;
    const int row = lq * 32 + lane;
    for (int j = 0; j < 64; j += 8) {
      float v[8];
      const int col = half * 64 + j;
      tmem_ld8(tmem + ((uint32_t)(lq * 32) << 16) + (uint32_t)col, v);
      float4* dst = reinterpret_cast<float4*>(ep + row * LD_EP + col);
      dst[0] = make_float4(v[0], v[1], v[2], v[3]);
      dst[1] = make_float4(v[4], v[5], v[6], v[7]);
    }
    tc_fence_before();
  }
  __syncthreads();

  if (active) {
    float sc[4], sh[4], mu[4], is[4], gm[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      sc[e] = tail->bn.scale[k0 + e];
      sh[e] = tail->bn.shift[k0 + e];
      mu[e] = tail->bn.mean[k0 + e];
      is[e] = tail->bn.istd[k0 + e];
      gm[e] = p.in.gamma[k0 + e];
    }
    // one row (or one 2x2 window of an upsampled source) per iteration; kept deliberately light on registers:
    // at 2 CTAs / SM a deeper register prefetch spills and measured slower ("dgrad epilogue")
    // software pipeline of depth 1: the global loads of row j+1 are issued before row j is processed
    auto row_of = [&](int j, long& row, int& r) -> bool {
      r = sg.up ? 4 * (rg + 8 * j) : rg + 8 * j;
      const int rd = tail->rows.rd[r];
      row = sg.up ? tail->rows.ru[r] : rd;
      return rd >= 0;
    };
    Raw4<T> xn, on;
    long row_n = 0;
    int r_n = 0;
    bool ok_n = row_of(0, row_n, r_n);
    if (ok_n) {
      xn.load(src + row_n * sg.ld + cl);
      if (ga.accumulate) on.load(G + row_n * ga.ld + cl);
    }
    // QuanInput between the ReLU and the conv (cunet_concat.act_bits): straight-through, zero where the activation >= 1
    const float zmax = p.in.act_bits ? 1.f : __int_as_float(0x7f800000);
    for (int j = 0; j < nrow_it; ++j) {
      const Raw4<T> xc = xn, oc = on;
      const long row = row_n;
      const int r = r_n;
      const bool ok = ok_n;
      if (j + 1 < nrow_it) {
        ok_n = row_of(j + 1, row_n, r_n);
        if (ok_n) {
          xn.load(src + row_n * sg.ld + cl);
          if (ga.accumulate) on.load(G + row_n * ga.ld + cl);
        }
      }
      if (!ok) continue;
      float gv[4] = {0, 0, 0, 0}, xv[4];
      xc.get(xv);
      if (!sg.up) {
        const float4 a = *reinterpret_cast<const float4*>(ep + r * LD_EP + quad * 4);
        const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float z = fmaf(xv[e], sc[e], sh[e]);
          const float dz = (z > 0.f && z < zmax) ? av[e] : 0.f;
          a_db[e] += dz;
          a_dg[e] += dz * (xv[e] - mu[e]) * is[e];
          gv[e] = gm[e] * dz;
        }
      } else {
        // upsampled source: tile rows r..r+3 are the four children of one low-resolution pixel
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          const float4 a = *reinterpret_cast<const float4*>(ep + (r + ch) * LD_EP + quad * 4);
          const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float z = fmaf(xv[e], sc[e], sh[e]);
            const float dz = (z > 0.f && z < zmax) ? av[e] : 0.f;
            a_db[e] += dz;
            a_dg[e] += dz * (xv[e] - mu[e]) * is[e];
            gv[e] += gm[e] * dz;
          }
        }
      }
      if (ga.accumulate) {
        float ov[4];
        oc.get(ov);
#pragma unroll
        for (int e = 0; e < 4; ++e) gv[e] += ov[e];
      }
      store4<T>(G + row * ga.ld + cl, gv);  // gv <- values as stored
    }
  }
  if (warp < 8) {
    float4* rp = reinterpret_cast<float4*>(red + tid * 16);
    rp[0] = make_float4(a_db[0], a_db[1], a_db[2], a_db[3]);
    rp[1] = make_float4(a_dg[0], a_dg[1], a_dg[2], a_dg[3]);
  }
  __syncthreads();
  if (tid < 128) {
    const int k = chunk * 128 + tid;
    if (k < Cin) {
      const int q = tid >> 2, e = tid & 3;
      float s0 = 0, s1 = 0;
      for (int t = q; t < 256; t += 32) {
        s0 += red[t * 16 + e];
        s1 += red[t * 16 + 4 + e];
      }
      int si = 0;
      while (k >= tail->bn.seg_start[si + 1]) ++si;
      if (p.gacc[si].G != nullptr) {
        atomicAdd(p.dbeta + k, s0);
        atomicAdd(p.dgamma + k, s1);
        if (p.gacc[si].gstats) {
          const int cl = k - tail->bn.seg_start[si];
          // this consumer's share of (sum G, sum G*xhat) = gamma * (dbeta, dgamma): every consumer of a tensor
          // normalises it with the same batch statistics (see conv_dgrad_v2.cu)
          const float gmk = p.in.gamma[k];
          atomicAdd(p.gacc[si].gstats + cl, (double)(gmk * s0));
          atomicAdd(p.gacc[si].gstats + p.in.seg[si].C + cl, (double)(gmk * s1));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc(tmem, 128);
}

}  // namespace cunet
using namespace cunet;

int cunet_conv_dgrad_v2_try(const cunet_conv_dgrad_params* p, cudaStream_t st);  // conv_dgrad_v2.cu

extern "C" int cunet_conv_dgrad(const cunet_conv_dgrad_params* p, void* stream) {
  if (!p) return cunet_fail("conv_dgrad: null params");
  if (p->in.nseg < 1 || p->in.nseg > CUNET_MAX_SEG) return cunet_fail("conv_dgrad: bad nseg");
  if (p->taps != 1 && p->taps != 9) return cunet_fail("conv_dgrad: taps must be 1 or 9");
  if (p->in.bn_train != 1) return cunet_fail("conv_dgrad: backward requires train-mode BN statistics");
  int cin = 0, up = 0;
  for (int s = 0; s < p->in.nseg; ++s) {
    if (p->in.seg[s].C % 32) return cunet_fail("conv_dgrad: segment channels must be a multiple of 32");
    cin += p->in.seg[s].C;
    up |= p->in.seg[s].up;
  }
  if (cin > MAX_CIN) return cunet_fail("conv_dgrad: too many input channels");
  if (p->dy.C > 128) return cunet_fail("conv_dgrad: Cout > 128");
  if (up && ((p->H | p->W) & 1)) return cunet_fail("conv_dgrad: upsampled source needs even H, W");
  const long M = (long)p->N * p->H * p->W;
  if (M <= 0) return 0;
  {
    // bf16 1x1: persistent bulk-copy kernel; everything else (fp32 split mode, 3x3, unusual layouts): this file
    static const bool v1_only = getenv("CUNET_DGRAD_V1") != nullptr;
    if (!v1_only) {
      const int r = cunet_conv_dgrad_v2_try(p, reinterpret_cast<cudaStream_t>(stream));
      if (r != 0) return r < 0 ? r : 0;
    }
  }
  const long tiles = up ? (M / 4 + 31) / 32 : (M + 127) / 128;
  dim3 grid((unsigned)tiles, (unsigned)((cin + 127) / 128));
  const size_t smem = DG_STAGES * (p->dtype == CUNET_BF16 ? StageGeom<bf16>::BYTES : StageGeom<float>::BYTES) +
                      sizeof(DgSmemTail) + 1024;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  cudaError_t e;
  if (p->dtype == CUNET_BF16) {
    e = cudaFuncSetAttribute(conv_dgrad_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cunet_fail_cuda("conv_dgrad attr", e);
    conv_dgrad_kernel<bf16><<<grid, DG_THREADS, smem, st>>>(*p);
  } else {
    e = cudaFuncSetAttribute(conv_dgrad_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cunet_fail_cuda("conv_dgrad attr", e);
    conv_dgrad_kernel<float><<<grid, DG_THREADS, smem, st>>>(*p);
  }
  e = cudaGetLastError();
  if (e != cudaSuccess) return cunet_fail_cuda("conv_dgrad launch", e);
  return 0;
}
