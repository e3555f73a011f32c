// Fused  virtual-concat -> BatchNorm -> ReLU -> conv (1x1 | 3x3)  [-> maxpool 2x2]  forward, sm_100a.
//
// Replaces, per call, the reference's torch.cat + nn.BatchNorm2d + nn.ReLU + nn.Conv2d (+ nn.MaxPool2d)
// sequence (models/cu_net.py:11-17, 22-25, 41-48, 195-198, 249, 260).
//
// GEMM view: D[128 pixels][CoutPad] = A[128 pixels][K] * W^T, K = taps * CinPad.
//   A operand : gathered from <= 8 source tensors (no cat tensor exists), BN scale/shift + ReLU applied
//               on the fly by 8 loader warps (global -> registers -> swizzled smem row tile).
//               Nearest-x2 upsample of a source = index math in the gather (models/cu_net.py:250,265).
//   B operand : pre-packed weight image, one 1-D TMA bulk copy per K block.
//   MMA       : tcgen05.mma cta_group::1, kind::f16 (bf16) or kind::tf32, M=128, N=CoutPad, fp32 accum in TMEM.
//   Epilogue  : TMEM -> smem (fp32) -> {round to storage dtype, optional 2x2 maxpool + argmax index,
//               per-channel sum / sum-of-squares for the consumers' BatchNorm} -> global (coalesced).
//
// Warp roles (320 threads): warps 0-7 A loaders then epilogue, warp 8 weight-copy producer,
// warp 9 TMEM allocator + MMA issuer.  3-stage mbarrier ring; 2 CTAs / SM.
#include "loaders.cuh"
#include "host_util.h"

namespace cunet {

constexpr int FWD_STAGES = 3;
constexpr int FWD_THREADS = 320;

struct FwdSmemTail {
  uint64_t full[FWD_STAGES];
  uint64_t empty[FWD_STAGES];
  uint64_t accum;
  uint32_t tmem_base;
  BnSmem bn;
  TileRowTable rows;
};

template <typename T>
__global__ void __launch_bounds__(FWD_THREADS, StageGeom<T>::MIN_CTAS) conv_fwd_kernel(const __grid_constant__ cunet_conv_fwd_params p) {
  using E = Elem<T>;
  using SG = StageGeom<T>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  FwdSmemTail* tail = reinterpret_cast<FwdSmemTail*>(smem + FWD_STAGES * SG::BYTES);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tile = blockIdx.x;
  const int grouped = p.pool;

  const int Cin = concat_cin(p.in);
  const int nkb = (Cin + E::KBE - 1) / E::KBE;  // K blocks per tap
  const int nsteps = p.taps * nkb;
  const uint32_t tmem_cols = p.CoutPad <= 32 ? 32 : (p.CoutPad <= 64 ? 64 : 128);

  // ---------------------------------------------------------------- one-time setup
  if (tid == 0) {
    for (int s = 0; s < FWD_STAGES; ++s) {
      mbar_init(&tail->full[s], 9);   // 8 loader warps + the producer's expect_tx arrive
      mbar_init(&tail->empty[s], 1);  // one tcgen05.commit
    }
    mbar_init(&tail->accum, 1);
    fence_mbar_init();
  }
  if (warp == 9) tmem_alloc(&tail->tmem_base, tmem_cols);
  griddep_wait();    // everything above overlaps the tail of the previous kernel (programmatic dependent launch)
  griddep_launch();
  compute_bn_coefs(p.in, &tail->bn, nkb * E::KBE, tid, FWD_THREADS);
  PixGeom geom;
  geom.N = p.N; geom.H = p.H; geom.W = p.W; geom.M = p.N * p.H * p.W;
  tile_rows_init(&tail->rows, geom, tile, grouped, tid);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tail->tmem_base;

  if (warp < 8) {
    // ============================================================== A loaders
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
    uint4 cur[4], nxt[4];
    uint32_t cmask = 0, nmask = 0;

    auto issue = [&](int it, uint4* dst, uint32_t& mask) {
      mask = 0;
      const int tap = it / nkb, kb = it - tap * nkb;
      int dy = 0, dx = 0;
      if (p.taps == 9) {
        dy = tap / 3 - 1;
        dx = tap - (tap / 3) * 3 - 1;
      }
      const ActStep st = act_step<T>(p.in, &tail->bn, kb * E::KBE + c * E::EPC);
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (act_load(st, rc, q, p.H, p.W, dy, dx, dst[q])) mask |= 1u << q;
    };

    issue(0, cur, cmask);
    for (int it = 0; it < nsteps; ++it) {
      const int s = it % FWD_STAGES;
      const uint32_t ph = (it / FWD_STAGES) & 1;
      if (it + 1 < nsteps) issue(it + 1, nxt, nmask);
      ActCoef<T> cf;
      cf.load(&tail->bn, (it % nkb) * E::KBE + c * E::EPC);
      mbar_wait(&tail->empty[s], ph ^ 1);
      const uint32_t abase = smem_u32(smem + s * SG::BYTES);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint4 o = make_uint4(0, 0, 0, 0), lo = make_uint4(0, 0, 0, 0);
        if ((cmask >> q) & 1) o = cf.apply(cur[q], lo);
        sts128(abase + soff[q], o);
        if (SG::SPLIT) sts128(abase + SG::A_LO + soff[q], lo);
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tail->full[s]);
#pragma unroll
      for (int q = 0; q < 4; ++q) cur[q] = nxt[q];
      cmask = nmask;
    }
  } else if (warp == 8) {
    // ============================================================== weight producer (TMA bulk)
    if (lane == 0) {
      const uint32_t bbytes = (uint32_t)p.CoutPad * 128u * (SG::SPLIT ? 2u : 1u);
      const char* w = reinterpret_cast<const char*>(p.wpack);
      for (int it = 0; it < nsteps; ++it) {
        const int s = it % FWD_STAGES;
        const uint32_t ph = (it / FWD_STAGES) & 1;
        mbar_wait(&tail->empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&tail->full[s], bbytes);
        bulk_g2s(smem + s * SG::BYTES + SG::B_OFF, w + (size_t)it * bbytes, bbytes, &tail->full[s]);
      }
    }
  } else {
    // ============================================================== MMA issuer
    if (lane == 0) {
      const uint32_t idesc = make_idesc(E::FMT, 128, (uint32_t)p.CoutPad, 0, 0);
      for (int it = 0; it < nsteps; ++it) {
        const int s = it % FWD_STAGES;
        const uint32_t ph = (it / FWD_STAGES) & 1;
        mbar_wait(&tail->full[s], ph);
        tc_fence_after();
        const uint32_t a = smem_u32(smem + s * SG::BYTES);
        const uint32_t b = a + SG::B_OFF;
        const uint32_t alo = a + SG::A_LO, blo = b + (uint32_t)p.CoutPad * 128u;
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

  // ================================================================== epilogue (warps 0-7)
  const int ld_ep = p.CoutPad + 4;  // fp32 words per staged row (bank-conflict-free float4 rows)
  float* ep = reinterpret_cast<float*>(smem);
  double* red = reinterpret_cast<double*>(smem + 128 * (128 + 4) * 4);  // [256][8] partial stats (fp64)
  if (warp < 8) {
    mbar_wait(&tail->accum, 0);
    tc_fence_after();
    const int lq = warp & 3, half = warp >> 2;
    const int row = lq * 32 + lane;
    const int ncol_half = p.CoutPad >> 1;
    for (int j = 0; j < ncol_half; j += 8) {
      float v[8];
      const int col = half * ncol_half + j;
      tmem_ld8(tmem + ((uint32_t)(lq * 32) << 16) + (uint32_t)col, v);
      float4* dst = reinterpret_cast<float4*>(ep + row * ld_ep + col);
      dst[0] = make_float4(v[0], v[1], v[2], v[3]);
      dst[1] = make_float4(v[4], v[5], v[6], v[7]);
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp < 8) {
    const int nq = p.CoutPad >> 2;  // channel quads per row
    const bool do_stats = p.out_stats != nullptr;
    double s1[4] = {0., 0., 0., 0.}, s2[4] = {0., 0., 0., 0.};
    const int nrows_out = grouped ? 32 : 128;
    const long rows_total = grouped ? (long)geom.M / 4 : (long)geom.M;
    for (int idx = tid; idx < nrows_out * nq; idx += 256) {
      const int r = idx / nq, q = idx - r * nq;
      const long grow = (long)tile * nrows_out + r;
      if (grow >= rows_total) continue;
      float4 v;
      uint32_t pidx = 0;
      if (!grouped) {
        v = *reinterpret_cast<const float4*>(ep + r * ld_ep + 4 * q);
      } else {
        // nn.MaxPool2d(2,2): first maximum in row-major window order wins
        v = *reinterpret_cast<const float4*>(ep + (4 * r) * ld_ep + 4 * q);
#pragma unroll
        for (int k = 1; k < 4; ++k) {
          const float4 u = *reinterpret_cast<const float4*>(ep + (4 * r + k) * ld_ep + 4 * q);
          if (u.x > v.x) { v.x = u.x; pidx = (pidx & ~0xFFu) | (uint32_t)k; }
          if (u.y > v.y) { v.y = u.y; pidx = (pidx & ~0xFF00u) | ((uint32_t)k << 8); }
          if (u.z > v.z) { v.z = u.z; pidx = (pidx & ~0xFF0000u) | ((uint32_t)k << 16); }
          if (u.w > v.w) { v.w = u.w; pidx = (pidx & ~0xFF000000u) | ((uint32_t)k << 24); }
        }
      }
      const int ch = 4 * q;
      if (ch >= p.Cout && !p.out_fp32) continue;
      if (p.out_fp32) {
        if (ch < p.out_ld)
          *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + grow * p.out_ld + ch) = v;
      } else {
        T* o = reinterpret_cast<T*>(p.out) + grow * p.out_ld + ch;
        T t0 = from_f<T>(v.x), t1 = from_f<T>(v.y), t2 = from_f<T>(v.z), t3 = from_f<T>(v.w);
        if (sizeof(T) == 4) {
          *reinterpret_cast<float4*>(o) = v;
        } else {
          uint2 pk;
          __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
          pk.x = *reinterpret_cast<uint32_t*>(&a);
          pk.y = *reinterpret_cast<uint32_t*>(&b);
          *reinterpret_cast<uint2*>(o) = pk;
        }
        // statistics of the values as stored (what the consumers will read back)
        v.x = to_f<T>(t0); v.y = to_f<T>(t1); v.z = to_f<T>(t2); v.w = to_f<T>(t3);
      }
      if (grouped && p.pool_idx) *reinterpret_cast<uint32_t*>(p.pool_idx + grow * p.Cout + ch) = pidx;
      if (do_stats) {
        s1[0] += (double)v.x; s1[1] += (double)v.y; s1[2] += (double)v.z; s1[3] += (double)v.w;
        s2[0] += (double)v.x * v.x; s2[1] += (double)v.y * v.y; s2[2] += (double)v.z * v.z; s2[3] += (double)v.w * v.w;
      }
    }
    if (do_stats) {
      // host guarantees 256 % nq == 0 when out_stats != NULL -> every thread owns one channel quad
      double* rp = red + tid * 8;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        rp[e] = s1[e];
        rp[4 + e] = s2[e];
      }
    }
  }
  __syncthreads();
  if (warp < 8 && p.out_stats != nullptr && tid < p.Cout) {
    const int nq = p.CoutPad >> 2;
    const int q = tid >> 2, e = tid & 3;
    double a = 0., b = 0.;
    for (int t = q; t < 256; t += nq) {
      a += red[t * 8 + e];
      b += red[t * 8 + 4 + e];
    }
    atomicAdd(p.out_stats + tid, a);
    atomicAdd(p.out_stats + p.Cout + tid, b);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc(tmem, tmem_cols);
}

}  // namespace cunet

using namespace cunet;

int cunet_conv_fwd3x3_try(const cunet_conv_fwd_params* p, cudaStream_t st);  // conv_fwd3x3.cu
int cunet_conv_fwd_v2_try(const cunet_conv_fwd_params* p, cudaStream_t st);   // conv_fwd_v2.cu
int cunet_conv_fwd_v3_try(const cunet_conv_fwd_params* p, cudaStream_t st);   // conv_fwd_v3.cu

extern "C" int cunet_conv_fwd(const cunet_conv_fwd_params* p, void* stream) {
  if (!p) return cunet_fail("conv_fwd: null params");
  if (p->in.nseg < 1 || p->in.nseg > CUNET_MAX_SEG) return cunet_fail("conv_fwd: bad nseg");
  if (p->taps != 1 && p->taps != 9) return cunet_fail("conv_fwd: taps must be 1 or 9");
  if (p->CoutPad % 16 || p->CoutPad < 16 || p->CoutPad > 128 || p->Cout > p->CoutPad)
    return cunet_fail("conv_fwd: CoutPad must be a multiple of 16 in [16,128]");
  int cin = 0;
  for (int s = 0; s < p->in.nseg; ++s) {
    if (p->in.seg[s].C % 32) return cunet_fail("conv_fwd: segment channels must be a multiple of 32");
    if (p->in.bn_train == 1 && !p->in.seg[s].stats) return cunet_fail("conv_fwd: train-mode BN needs seg stats");
    cin += p->in.seg[s].C;
  }
  if (cin > MAX_CIN) return cunet_fail("conv_fwd: too many input channels");
  if (p->out_stats && (256 % (p->CoutPad / 4) || p->Cout != p->CoutPad))
    return cunet_fail("conv_fwd: out_stats needs Cout == CoutPad in {32, 64, 128}");
  if (p->pool && ((p->H | p->W) & 1)) return cunet_fail("conv_fwd: pool needs even H, W");
  if (p->Cout % 4 && !p->out_fp32) return cunet_fail("conv_fwd: Cout must be a multiple of 4");
  const long M = (long)p->N * p->H * p->W;
  if (M <= 0) return 0;
  if (p->taps == 1) {
    // bf16 1x1 on large maps: third-generation persistent kernel (resident weights, transposed GEMM, pooling fused)
    int r = cunet_conv_fwd_v3_try(p, reinterpret_cast<cudaStream_t>(stream));
    if (r != 0) return r < 0 ? r : 0;
    // opt-in second-generation kernel (kept for A/B runs)
    r = cunet_conv_fwd_v2_try(p, reinterpret_cast<cudaStream_t>(stream));
    if (r != 0) return r < 0 ? r : 0;
  }
  if (p->taps == 9) {
    // bf16 dense-layer 3x3 (128 -> 32, W in {2..64}): persistent shifted-descriptor kernel; everything else: this file
    const int r = cunet_conv_fwd3x3_try(p, reinterpret_cast<cudaStream_t>(stream));
    if (r != 0) return r < 0 ? r : 0;
  }
  const long tiles = p->pool ? (M / 4 + 31) / 32 : (M + 127) / 128;
  const size_t smem = FWD_STAGES * (p->dtype == CUNET_BF16 ? StageGeom<bf16>::BYTES : StageGeom<float>::BYTES) +
                      sizeof(FwdSmemTail) + 1024;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  cudaError_t e;
  if (p->dtype == CUNET_BF16) {
    e = cudaFuncSetAttribute(conv_fwd_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cunet_fail_cuda("conv_fwd attr", e);
    e = cunet_launch(conv_fwd_kernel<bf16>, dim3((unsigned)tiles), dim3(FWD_THREADS), smem, st, *p);
    if (e != cudaSuccess) return cunet_fail_cuda("conv_fwd launch", e);
  } else {
    e = cudaFuncSetAttribute(conv_fwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cunet_fail_cuda("conv_fwd attr", e);
    e = cunet_launch(conv_fwd_kernel<float>, dim3((unsigned)tiles), dim3(FWD_THREADS), smem, st, *p);
    if (e != cudaSuccess) return cunet_fail_cuda("conv_fwd launch", e);
  }
  e = cudaGetLastError();
  if (e != cudaSuccess) return cunet_fail_cuda("conv_fwd launch", e);
  return 0;
}
