// Backward-filter of the 3x3 conv (dense-layer conv2, models/cu_net.py:47-48), all nine taps in one CTA.
//
//   dW[co][c][tap] = sum_px dY[px][co] * A[px + off(tap)][c]  =  sum_px' A[px'][c] * dY[px' - off(tap)][co]
//
// The second form keeps the WIDE operand (A = relu(bn2(bottleneck)), 128 channels) unshifted -- it is gathered
// once per pixel row -- and shifts the NARROW one (dY has 32 channels): the nine shifted dY copies form the
// N dimension of one GEMM  D[128 c][(tap, co)] += A^T[128 c][R px] * B[R px][(tap, co)]  (N = 288, padded to 320
// for bf16 so that it is a whole number of 64-element MN-major groups).  The round-1 kernel launched one CTA per
// tap, each re-gathering both operands (9x the traffic and 9x the transform instructions): 184 us at 64x64, batch 24.
#include "loaders.cuh"
#include "host_util.h"

namespace cunet {

constexpr int W3_THREADS = 320;

template <typename T> struct W3Geom {
  using E = Elem<T>;
  static constexpr bool SPLIT = E::SPLIT;
  static constexpr int R = 4 * E::MMA_K;                  // pixel rows per stage (4 MMAs along K)
  static constexpr int SUB = R * 128;                     // one sub-tile: R rows x 128 B
  static constexpr int A_SUBS = 128 / E::KBE;             // 2 (bf16) / 4 (fp32)
  static constexpr int TAPS_PAD = SPLIT ? 9 : 10;         // bf16: two taps per 64-element group -> pad to 10
  static constexpr int B_SUBS = TAPS_PAD * 32 / E::KBE;   // 5 (bf16) / 9 (fp32)
  static constexpr int NCOL = TAPS_PAD * 32;              // 320 / 288 accumulator columns
  static constexpr int N1_SUBS = SPLIT ? 5 : 3;           // first MMA: 160 (fp32) / 192 (bf16) columns
  static constexpr int N1 = N1_SUBS * E::KBE;
  static constexpr int N2 = NCOL - N1;                    // 128
  static constexpr int A_BYTES = A_SUBS * SUB;            // 16 KB
  static constexpr int B_BYTES = B_SUBS * SUB;            // 40 KB / 36 KB
  static constexpr int STAGE = (A_BYTES + B_BYTES) * (SPLIT ? 2 : 1);
  static constexpr int STAGES = SPLIT ? 2 : 3;
  static constexpr int A_LO = A_BYTES;                    // split mode: [A_hi][A_lo][B_hi][B_lo]
  static constexpr int B_OFF = SPLIT ? 2 * A_BYTES : A_BYTES;
  static constexpr int B_LO = B_OFF + B_BYTES;
  static constexpr int ACPR = 128 / E::EPC;               // activation chunks per pixel row
  static constexpr int ARPP = 256 / ACPR;                 // rows per loader pass
  static constexpr int BCPR = B_SUBS * 8;                 // gradient chunk columns per pixel row (incl. padding)
  static constexpr int BPT = (R * BCPR + 255) / 256;      // gradient chunks per thread per stage (10 / 9)
};

struct W3SmemTail {
  uint64_t full[3];
  uint64_t empty[3];
  uint64_t accum;
  uint32_t tmem_base;
  BnSmem bn;
  GradSmem gc;
};

template <typename T>
__global__ void __launch_bounds__(W3_THREADS, 1) conv_wgrad3x3_kernel(const __grid_constant__ cunet_conv_wgrad_params p,
                                                                        int nsplit) {
  using E = Elem<T>;
  using G = W3Geom<T>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  W3SmemTail* tail = reinterpret_cast<W3SmemTail*>(smem + G::STAGES * G::STAGE);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int split = blockIdx.x;
  const int Cin = concat_cin(p.in);  // 128
  const int M = p.N * p.H * p.W;
  const int total_steps = (M + G::R - 1) / G::R;
  const int per = (total_steps + nsplit - 1) / nsplit;
  const int st0 = split * per;
  const int st1 = min(total_steps, st0 + per);
  const int nsteps = max(0, st1 - st0);

  if (tid == 0) {
    for (int s = 0; s < G::STAGES; ++s) {
      mbar_init(&tail->full[s], 8);
      mbar_init(&tail->empty[s], 1);
    }
    mbar_init(&tail->accum, 1);
    fence_mbar_init();
  }
  if (warp == 9) tmem_alloc(&tail->tmem_base, 512);
  compute_bn_coefs(p.in, &tail->bn, 128, tid, W3_THREADS);
  compute_grad_coefs(p.dy, &tail->gc, tid, W3_THREADS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tail->tmem_base;

  if (warp < 8) {
    // ============================================================== loaders
    const int acc = tid % G::ACPR, ar0 = tid / G::ACPR;
    const int ach = acc * E::EPC;
    const ActStep ast = act_step<T>(p.in, &tail->bn, ach);
    ActCoef<T> acf;
    acf.load(&tail->bn, ach);
    uint32_t aoff[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) aoff[q] = (acc >> 3) * G::SUB + tile_off_mn<T>(ar0 + G::ARPP * q, acc & 7);
    // gradient chunks of this thread: idx = tid + 256*j -> (row, chunk column) ; chunk column -> (tap, co)
    PixDiv pd;
    pd.init(p.H, p.W);

    for (int it = 0; it < nsteps; ++it) {
      const int s = it % G::STAGES;
      const uint32_t ph = (it / G::STAGES) & 1;
      const int m0 = (st0 + it) * G::R;
      uint4 araw[4];
      uint32_t amask = 0;
      RowCtx arc;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int m = m0 + ar0 + G::ARPP * q;
        int n = 0, h = 0, w = 0;
        const bool valid = m < M;
        if (valid) pd.split(m, n, h, w);
        rowctx_set(arc, q, valid, n, h, w, p.H, p.W);
        if (act_load(ast, arc, q, p.H, p.W, 0, 0, araw[q])) amask |= 1u << q;
      }
      GradRaw<T> graw[G::BPT];
      uint32_t gmask = 0;
#pragma unroll
      for (int j = 0; j < G::BPT; ++j) {
        const int idx = tid + 256 * j;
        if (idx < G::R * G::BCPR) {
          const int r = idx / G::BCPR, cc = idx - r * G::BCPR;
          const int kg = cc * E::EPC, tap = kg >> 5, co = kg & 31;
          const int m = m0 + r;
          if (tap < 9 && m < M) {
            int n, h, w;
            pd.split(m, n, h, w);
            RowCtx rc;
            rowctx_set(rc, 0, true, n, h, w, p.H, p.W);
            const int dy = tap / 3 - 1, dx = tap - (tap / 3) * 3 - 1;
            if (grad_load<T>(p.dy, rc, 0, p.H, p.W, co, -dy, -dx, graw[j])) gmask |= 1u << j;
          }
        }
      }
      mbar_wait(&tail->empty[s], ph ^ 1);
      const uint32_t abase = smem_u32(smem + s * G::STAGE);
      const uint32_t bbase = abase + G::B_OFF;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint4 o = make_uint4(0, 0, 0, 0), lo = make_uint4(0, 0, 0, 0);
        if ((amask >> q) & 1) o = acf.apply(araw[q], lo);
        sts128(abase + aoff[q], o);
        if (G::SPLIT) sts128(abase + G::A_LO + aoff[q], lo);
      }
#pragma unroll
      for (int j = 0; j < G::BPT; ++j) {
        const int idx = tid + 256 * j;
        if (idx < G::R * G::BCPR) {
          const int r = idx / G::BCPR, cc = idx - r * G::BCPR;
          uint4 o = make_uint4(0, 0, 0, 0), lo = make_uint4(0, 0, 0, 0);
          if ((gmask >> j) & 1) {
            GradCoef<T> gcf;
            gcf.load(&tail->gc, (cc * E::EPC) & 31);
            o = gcf.apply(p.dy, graw[j], lo);
          }
          const uint32_t off = (cc >> 3) * G::SUB + tile_off_mn<T>(r, cc & 7);
          sts128(bbase + off, o);
          if (G::SPLIT) sts128(abase + G::B_LO + off, lo);
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tail->full[s]);
    }
  } else if (warp == 9) {
    // ============================================================== MMA issuer
    if (lane == 0) {
      const uint32_t idesc1 = make_idesc(E::FMT, 128, (uint32_t)G::N1, 1, 1);
      const uint32_t idesc2 = make_idesc(E::FMT, 128, (uint32_t)G::N2, 1, 1);
      constexpr uint32_t KSTEP = E::MMA_K * 128;
      for (int it = 0; it < nsteps; ++it) {
        const int s = it % G::STAGES;
        const uint32_t ph = (it / G::STAGES) & 1;
        mbar_wait(&tail->full[s], ph);
        tc_fence_after();
        const uint32_t a = smem_u32(smem + s * G::STAGE);
        const uint32_t b1 = a + G::B_OFF, b2 = b1 + G::N1_SUBS * G::SUB;
        const uint32_t alo = a + G::A_LO, b1lo = a + G::B_LO, b2lo = b1lo + G::N1_SUBS * G::SUB;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const uint32_t acc = (uint32_t)((it | kk) != 0);
          const uint64_t ad = make_sdesc_mn<T>(a + kk * KSTEP, G::SUB);
          umma<T>(tmem, ad, make_sdesc_mn<T>(b1 + kk * KSTEP, G::SUB), idesc1, acc);
          umma<T>(tmem + G::N1, ad, make_sdesc_mn<T>(b2 + kk * KSTEP, G::SUB), idesc2, acc);
          if (G::SPLIT) {
            const uint64_t adl = make_sdesc_mn<T>(alo + kk * KSTEP, G::SUB);
            umma<T>(tmem, adl, make_sdesc_mn<T>(b1 + kk * KSTEP, G::SUB), idesc1, 1u);
            umma<T>(tmem + G::N1, adl, make_sdesc_mn<T>(b2 + kk * KSTEP, G::SUB), idesc2, 1u);
            umma<T>(tmem, ad, make_sdesc_mn<T>(b1lo + kk * KSTEP, G::SUB), idesc1, 1u);
            umma<T>(tmem + G::N1, ad, make_sdesc_mn<T>(b2lo + kk * KSTEP, G::SUB), idesc2, 1u);
          }
        }
        tc_commit(&tail->empty[s]);
      }
      tc_commit(&tail->accum);
    }
  }

  // ================================================================== epilogue: TMEM -> red.global.add
  if (warp < 8 && nsteps > 0) {
    mbar_wait(&tail->accum, 0);
    tc_fence_after();
    const int lq = warp & 3, half = warp >> 2;
    const int c = lq * 32 + lane;  // input channel (TMEM lane)
    for (int col = half * 8; col < 288; col += 16) {
      float v[8];
      tmem_ld8(tmem + ((uint32_t)(lq * 32) << 16) + (uint32_t)col, v);
      const int tap = col >> 5, co0 = col & 31;
      if (c < Cin) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int co = co0 + e;
          if (co < p.Cout) atomicAdd(p.dw + ((long)co * Cin + c) * 9 + tap, v[e]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc(tmem, 512);
}

}  // namespace cunet
using namespace cunet;

// called by cunet_conv_wgrad for taps == 9 (conv_wgrad.cu)
int cunet_conv_wgrad3x3_launch(const cunet_conv_wgrad_params* p, cudaStream_t st) {
  const long M = (long)p->N * p->H * p->W;
  const int R = p->dtype == CUNET_BF16 ? W3Geom<bf16>::R : W3Geom<float>::R;
  const int total_steps = (int)((M + R - 1) / R);
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  int nsplit = p->nsplit > 0 ? p->nsplit : sms;
  if (nsplit > total_steps) nsplit = total_steps;
  if (nsplit < 1) nsplit = 1;
  cudaError_t e;
  if (p->dtype == CUNET_BF16) {
    const size_t smem = W3Geom<bf16>::STAGES * W3Geom<bf16>::STAGE + sizeof(W3SmemTail) + 1024;
    e = cudaFuncSetAttribute(conv_wgrad3x3_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cunet_fail_cuda("conv_wgrad3x3 attr", e);
    conv_wgrad3x3_kernel<bf16><<<nsplit, W3_THREADS, smem, st>>>(*p, nsplit);
  } else {
    const size_t smem = W3Geom<float>::STAGES * W3Geom<float>::STAGE + sizeof(W3SmemTail) + 1024;
    e = cudaFuncSetAttribute(conv_wgrad3x3_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cunet_fail_cuda("conv_wgrad3x3 attr", e);
    conv_wgrad3x3_kernel<float><<<nsplit, W3_THREADS, smem, st>>>(*p, nsplit);
  }
  e = cudaGetLastError();
  if (e != cudaSuccess) return cunet_fail_cuda("conv_wgrad3x3 launch", e);
  return 0;
}
