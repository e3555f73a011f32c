// Shared device-side helpers for the CU-Net sm_100a kernels: mbarrier / TMA-bulk / tcgen05 PTX
// wrappers, UMMA descriptor construction and the SWIZZLE_128B shared-memory tile layout that
// every GEMM-shaped kernel in this library uses.
//
// Tile layout ("row tile"): R rows x 128 bytes, row r at byte r*128, the eight 16-byte chunks of a
// row XOR-swizzled with (r & 7).  Base address 1024-byte aligned.  The same physical image is
//   * a K-major  SWIZZLE_128B operand (rows = M/N index, 128 B = one K block)         -> fwd, dgrad
//   * an MN-major SWIZZLE_128B operand (rows = K index, 128 B = one group of MN elems) -> wgrad
// (canonical layouts: cute/atom/mma_traits_sm100.hpp, make_umma_desc<Major::K / Major::MN>).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace cunet {

typedef __nv_bfloat16 bf16;

// ------------------------------------------------------------------------------------------------
// element traits
template <typename T> struct Elem;
template <> struct Elem<float> {
  static constexpr int ESZ = 4;      // bytes per element
  static constexpr int EPC = 4;      // elements per 16-byte chunk
  static constexpr int KBE = 32;     // elements per 128-byte row (one K block)
  static constexpr uint32_t FMT = 2; // UMMA F16F32Format::TF32
  static constexpr int MMA_K = 8;    // K per tcgen05.mma (32 bytes)
  // fp32 mode = 3xTF32: every operand is split x = hi + lo (both tf32) and D += Ahi*Bhi + Alo*Bhi + Ahi*Blo,
  // which restores ~fp32 accuracy on the tf32 tensor-core path (needed for the 1e-3 parity config).
  static constexpr bool SPLIT = true;
};
template <> struct Elem<bf16> {
  static constexpr int ESZ = 2;
  static constexpr int EPC = 8;
  static constexpr int KBE = 64;
  static constexpr uint32_t FMT = 1; // BF16
  static constexpr int MMA_K = 16;
  static constexpr bool SPLIT = false;
};

// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
// Bounded wait: a protocol bug becomes a trap (launch error) instead of a hung GPU.  The suspend-time hint lets
// the hardware park the thread until the phase completes (without it try_wait returns at once and the waiting
// warps hot-spin: the round-1 profile of the persistent dgrad kernel showed ~60 % of all issued instructions
// were these polls, stealing issue slots from the working warps).
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t addr = smem_u32(bar);
  uint32_t done = 0;
#pragma unroll 1
  for (uint32_t spin = 0; spin < 2000u; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity), "r"(0x989680u)
        : "memory");
    if (done) return;
  }
  __trap();
}

// Programmatic dependent launch (launch attribute set by cunet_launch, host_util.h).  wait: block until every
// prerequisite grid has completed and its memory is visible (no-op for a normal launch); launch: let the next
// kernel in the stream start its prologue (barrier init, TMEM allocation) while this one is still running.
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// Optional in-kernel timeline (only in builds with -DCUNET_TRACE, see tools/build_trace.sh; the default build compiles
// these to nothing): CTA 0 stores clock64() at role milestones into a device buffer handed over by a per-kernel
// cunet_debug_trace_*(buf) setter; tools/trace_kernels.py prints the timelines.
#ifdef CUNET_TRACE
#define CUNET_TRACE_DECL(sym) __device__ long long* sym = nullptr;
#define CUNET_TRACE_LOAD(var, sym) long long* var = sym;
#define CUNET_TRACE_MARK(var, slot)                                          \
  do {                                                                       \
    if ((var) != nullptr && blockIdx.x == 0) (var)[(slot)] = clock64();      \
  } while (0)
#define CUNET_TRACE_SETTER(fn, sym)                                          \
  extern "C" int fn(void* buf) {                                             \
    long long* b = reinterpret_cast<long long*>(buf);                        \
    return cudaMemcpyToSymbol(sym, &b, sizeof(b)) == cudaSuccess ? 0 : -1;   \
  }
#else
#define CUNET_TRACE_DECL(sym)
#define CUNET_TRACE_LOAD(var, sym)
#define CUNET_TRACE_MARK(var, slot) \
  do {                              \
  } while (0)
#define CUNET_TRACE_SETTER(fn, sym)
#endif

// generic-proxy smem writes -> visible to the async proxy (tcgen05.mma / bulk copies)
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// 1-D TMA bulk copy global -> shared, completion on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst_smem)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// ------------------------------------------------------------------------------------------------
// tcgen05 / TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// all previously issued tcgen05.mma of this thread -> arrive on an mbarrier when complete
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}

template <typename T>
__device__ __forceinline__ void umma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                     uint32_t accumulate);
template <>
__device__ __forceinline__ void umma<bf16>(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
template <>
__device__ __forceinline__ void umma<float>(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// 32 lanes x 8 consecutive 32-bit columns -> 8 registers per thread (thread i <-> lane base+i)
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
  uint32_t r0, r1, r2, r3, r4, r5, r6, r7;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3), "=r"(r4), "=r"(r5), "=r"(r6), "=r"(r7)
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  v[0] = __uint_as_float(r0); v[1] = __uint_as_float(r1); v[2] = __uint_as_float(r2); v[3] = __uint_as_float(r3);
  v[4] = __uint_as_float(r4); v[5] = __uint_as_float(r5); v[6] = __uint_as_float(r6); v[7] = __uint_as_float(r7);
}

// UMMA instruction descriptor (cute/arch/mma_sm100_desc.hpp, union InstrDescriptor):
//  [4,6) c_format=1 (F32)  [7,10) a_format  [10,13) b_format  [15] a_major  [16] b_major
//  [17,23) N>>3  [24,29) M>>4
__host__ __device__ constexpr uint32_t make_idesc(uint32_t fmt, uint32_t M, uint32_t N, uint32_t a_mn_major,
                                                  uint32_t b_mn_major) {
  return (1u << 4) | (fmt << 7) | (fmt << 10) | (a_mn_major << 15) | (b_mn_major << 16) | ((N >> 3) << 17) |
         ((M >> 4) << 24);
}

// UMMA shared-memory descriptor (union SmemDescriptor): start>>4 [0,14), LBO>>4 [16,30),
// SBO>>4 [32,46), version=1 [46,48), layout_type [61,64) (2 = SWIZZLE_128B).
__device__ __forceinline__ uint64_t make_sdesc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46) | (2ull << 61);
}

// Advance a shared-memory descriptor by a byte offset (a multiple of 16 that keeps the start address inside the
// 256 KB window, so only the 14-bit start-address field changes): the MMA issuer is ONE thread, and rebuilding a
// descriptor from scratch for every tcgen05.mma (~10 dependent integer instructions, twice per MMA) made that thread
// the long pole of the 64-pixel-stage kernels -- 2 us to issue the 20 MMAs of a stage (in-kernel timeline of
// conv_fwd_v3, round 2).  With this the per-MMA cost is two adds.
__device__ __forceinline__ uint64_t sdesc_advance(uint64_t desc, uint32_t byte_off) {
  return desc + (uint64_t)(byte_off >> 4);
}

// byte offset of 16-byte chunk `c` (0..7) of row `r` inside a row tile (tile base 1024-aligned)
__device__ __forceinline__ uint32_t tile_off(int r, int c) {
  return (uint32_t)(r * 128 + ((c ^ (r & 7)) << 4));
}

// MN-major operands (wgrad).  16-bit types use the same SWIZZLE_128B image as above.  32-bit types (tf32)
// need SWIZZLE_128B_BASE32B (layout_type 1): the swizzle permutes 32-byte units with (row & 3) and the K
// atom is 4 rows, SBO = stride between 4-row groups (measured with tools/umma_probe.cu on B200; the other
// layout types return zeros for MN-major tf32).
template <typename T> __device__ __forceinline__ uint32_t tile_off_mn(int r, int c);
template <> __device__ __forceinline__ uint32_t tile_off_mn<bf16>(int r, int c) { return tile_off(r, c); }
template <> __device__ __forceinline__ uint32_t tile_off_mn<float>(int r, int c) {
  return (uint32_t)(r * 128 + (((((c >> 1) ^ (r & 3)) << 1) | (c & 1)) << 4));
}
template <typename T> __device__ __forceinline__ uint64_t make_sdesc_mn(uint32_t saddr, uint32_t lbo_bytes);
template <> __device__ __forceinline__ uint64_t make_sdesc_mn<bf16>(uint32_t saddr, uint32_t lbo_bytes) {
  return make_sdesc(saddr, lbo_bytes, 1024);
}
template <> __device__ __forceinline__ uint64_t make_sdesc_mn<float>(uint32_t saddr, uint32_t lbo_bytes) {
  return (make_sdesc(saddr, lbo_bytes, 512) & ~(7ull << 61)) | (1ull << 61);
}

__device__ __forceinline__ void sts128(uint32_t saddr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(saddr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}
__device__ __forceinline__ uint4 ldg128(const void* p) {
  uint4 v;
  asm volatile("ld.global.nc.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
// coherent variant (buffers that were written by an earlier kernel on the same stream are fine with
// .nc as well; this one is for buffers read-modify-written inside the same kernel)
__device__ __forceinline__ uint4 ldg128_c(const void* p) {
  uint4 v;
  asm volatile("ld.global.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}

__device__ __forceinline__ float tf32_round(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// ------------------------------------------------------------------------------------------------
// chunk <-> float conversion (a chunk is 16 bytes: 4 floats or 8 bf16)
template <typename T> struct Chunk;
template <> struct Chunk<float> {
  static constexpr int N = 4;
  static __device__ __forceinline__ void unpack(const uint4& v, float* f) {
    f[0] = __uint_as_float(v.x); f[1] = __uint_as_float(v.y); f[2] = __uint_as_float(v.z); f[3] = __uint_as_float(v.w);
  }
  // operand rounding for the tensor core (tf32, round-to-nearest)
  static __device__ __forceinline__ uint4 pack_mma(const float* f) {
    return make_uint4(__float_as_uint(tf32_round(f[0])), __float_as_uint(tf32_round(f[1])),
                      __float_as_uint(tf32_round(f[2])), __float_as_uint(tf32_round(f[3])));
  }
  static __device__ __forceinline__ uint4 pack_lo(const float* f) {
    return make_uint4(__float_as_uint(tf32_round(f[0] - tf32_round(f[0]))), __float_as_uint(tf32_round(f[1] - tf32_round(f[1]))),
                      __float_as_uint(tf32_round(f[2] - tf32_round(f[2]))), __float_as_uint(tf32_round(f[3] - tf32_round(f[3]))));
  }
  static __device__ __forceinline__ uint4 pack(const float* f) {
    return make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]), __float_as_uint(f[3]));
  }
};
template <> struct Chunk<bf16> {
  static constexpr int N = 8;
  static __device__ __forceinline__ void unpack(const uint4& v, float* f) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      f[2 * i] = __uint_as_float(w[i] << 16);
      f[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
    }
  }
  static __device__ __forceinline__ uint4 pack(const float* f) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&h);
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
  }
  static __device__ __forceinline__ uint4 pack_mma(const float* f) { return pack(f); }
  static __device__ __forceinline__ uint4 pack_lo(const float*) { return make_uint4(0, 0, 0, 0); }
};

template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<bf16>(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f<bf16>(float v) { return __float2bfloat16_rn(v); }

// 4 consecutive elements (16 B fp32 / 8 B bf16, naturally aligned)
template <typename T> __device__ __forceinline__ void load4(const T* p, float* f);
template <> __device__ __forceinline__ void load4<float>(const float* p, float* f) {
  const float4 v = *reinterpret_cast<const float4*>(p);
  f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
}
template <> __device__ __forceinline__ void load4<bf16>(const bf16* p, float* f) {
  const uint2 v = *reinterpret_cast<const uint2*>(p);
  f[0] = __uint_as_float(v.x << 16); f[1] = __uint_as_float(v.x & 0xFFFF0000u);
  f[2] = __uint_as_float(v.y << 16); f[3] = __uint_as_float(v.y & 0xFFFF0000u);
}
// raw (still packed) 4-element load, for prefetching many rows without spending fp32 registers on them
template <typename T> struct Raw4;
template <> struct Raw4<float> {
  float4 v;
  __device__ __forceinline__ void load(const float* p) { v = *reinterpret_cast<const float4*>(p); }
  __device__ __forceinline__ void get(float* f) const { f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w; }
};
template <> struct Raw4<bf16> {
  uint2 v;
  __device__ __forceinline__ void load(const bf16* p) { v = *reinterpret_cast<const uint2*>(p); }
  __device__ __forceinline__ void get(float* f) const {
    f[0] = __uint_as_float(v.x << 16); f[1] = __uint_as_float(v.x & 0xFFFF0000u);
    f[2] = __uint_as_float(v.y << 16); f[3] = __uint_as_float(v.y & 0xFFFF0000u);
  }
};

// 8 consecutive elements (one 16-byte bf16 vector / two fp32 vectors), kept packed until used
template <typename T> struct Raw8;
template <> struct Raw8<float> {
  float4 a, b;
  __device__ __forceinline__ void load(const float* p) {
    a = *reinterpret_cast<const float4*>(p);
    b = *reinterpret_cast<const float4*>(p + 4);
  }
  __device__ __forceinline__ void get(float* f) const {
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
  }
  // stores f (8 values) and leaves in f the values as stored
  static __device__ __forceinline__ void store(float* p, float* f) {
    *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(f[4], f[5], f[6], f[7]);
  }
};
template <> struct Raw8<bf16> {
  uint4 v;
  __device__ __forceinline__ void load(const bf16* p) { v = *reinterpret_cast<const uint4*>(p); }
  __device__ __forceinline__ void get(float* f) const {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      f[2 * i] = __uint_as_float(w[i] << 16);
      f[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
    }
  }
  static __device__ __forceinline__ void store(bf16* p, float* f) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&h);
      f[2 * i] = __uint_as_float(w[i] << 16);
      f[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
    }
    *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
  }
};

// store 4 elements; returns the values as stored (after rounding to T) in f
template <typename T> __device__ __forceinline__ void store4(T* p, float* f);
template <> __device__ __forceinline__ void store4<float>(float* p, float* f) {
  *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
}
template <> __device__ __forceinline__ void store4<bf16>(bf16* p, float* f) {
  __nv_bfloat162 a = __floats2bfloat162_rn(f[0], f[1]), b = __floats2bfloat162_rn(f[2], f[3]);
  uint2 pk;
  pk.x = *reinterpret_cast<uint32_t*>(&a);
  pk.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = pk;
  f[0] = __uint_as_float(pk.x << 16); f[1] = __uint_as_float(pk.x & 0xFFFF0000u);
  f[2] = __uint_as_float(pk.y << 16); f[3] = __uint_as_float(pk.y & 0xFFFF0000u);
}

// ------------------------------------------------------------------------------------------------
// tile row -> pixel.  Raster order, or 2x2-window-grouped order (rows 4q..4q+3 = one pooling
// window / the four children of one low-resolution pixel), used by pooling / upsample epilogues.
struct PixGeom {
  int N, H, W;      // full-resolution dims of this op
  int M;            // N*H*W
};
// 32-bit index math only (64-bit division is ~100 instructions on the GPU); shifts when W, H are powers of two.
__device__ __forceinline__ void pix_split(int m, int H, int W, int& n, int& h, int& w) {
  if (((W & (W - 1)) | (H & (H - 1))) == 0) {
    const int lw = __ffs(W) - 1, lh = __ffs(H) - 1;
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
__device__ __forceinline__ bool tile_row_pixel(const PixGeom& g, int tile, int r, int grouped, int& n, int& h, int& w) {
  if (!grouped) {
    const int m = tile * 128 + r;
    if (m >= g.M) return false;
    pix_split(m, g.H, g.W, n, h, w);
    return true;
  } else {
    const int win = tile * 32 + (r >> 2);
    const int Wh = g.W >> 1, Hh = g.H >> 1;
    if (win >= g.N * Hh * Wh) return false;
    int hh, ww;
    pix_split(win, Hh, Wh, n, hh, ww);
    h = hh * 2 + ((r >> 1) & 1);
    w = ww * 2 + (r & 1);
    return true;
  }
}

// per-CTA table of the tile's 128 rows (computed once by 128 threads, read by loaders and epilogues)
struct TileRowTable {
  int rd[128];  // full-resolution row index or -1
  int ru[128];  // half-resolution row index
  int hw[128];  // (h << 16) | w
};
__device__ __forceinline__ void tile_rows_init(TileRowTable* t, const PixGeom& g, int tile, int grouped, int tid) {
  if (tid < 128) {
    int n = 0, h = 0, w = 0;
    const bool valid = tile_row_pixel(g, tile, tid, grouped, n, h, w);
    t->rd[tid] = valid ? (n * g.H + h) * g.W + w : -1;
    t->ru[tid] = valid ? (n * (g.H >> 1) + (h >> 1)) * (g.W >> 1) + (w >> 1) : 0;
    t->hw[tid] = (h << 16) | w;
  }
}

}  // namespace cunet
