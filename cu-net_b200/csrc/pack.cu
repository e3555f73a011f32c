// Weight packing: reference-layout fp32 master weights [Cout][Cin][kh][kw] -> tensor-core operand
// images in the SWIZZLE_128B row-tile layout of common.cuh (so a K block is one 1-D TMA bulk copy).
// Runs once per optimizer step for all convs in one launch (weights change every step).
#include "common.cuh"
#include "../../include/cunet_b200.h"
#include "host_util.h"

namespace cunet {

template <typename T> struct PackGeom {
  using E = Elem<T>;
  static __host__ __device__ int nkb_fwd(int Cin) { return (Cin + E::KBE - 1) / E::KBE; }
  static __host__ __device__ int coutk(int taps, int Cout, int CoutPad) { return taps == 9 ? Cout : CoutPad; }
  static __host__ __device__ int nkb_dgrad(int taps, int Cout, int CoutPad) {
    return (taps * coutk(taps, Cout, CoutPad) + E::KBE - 1) / E::KBE;
  }
  static __host__ __device__ int nchunk(int Cin) { return (Cin + 127) / 128; }
};

template <typename T>
__global__ void pack_weights_kernel(const cunet_pack_desc* __restrict__ descs, int max_chunks) {
  using E = Elem<T>;
  using G = PackGeom<T>;
  const cunet_pack_desc d = descs[blockIdx.y];
  const int nkb = G::nkb_fwd(d.Cin);
  const long n_fwd = d.fwd ? (long)d.taps * nkb * d.CoutPad * 8 : 0;
  const int nkbg = G::nkb_dgrad(d.taps, d.Cout, d.CoutPad);
  const int coutk = G::coutk(d.taps, d.Cout, d.CoutPad);
  const long n_dg = d.dgrad ? (long)G::nchunk(d.Cin) * nkbg * 128 * 8 : 0;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n_fwd + n_dg; i += (long)gridDim.x * blockDim.x) {
    float f[E::EPC];
    char* dst;
    long lo_off = 0;
    if (i < n_fwd) {
      // fwd image [tap][kb][co][8 chunks]
      const int cphys = (int)(i & 7);
      long t = i >> 3;
      const int co = (int)(t % d.CoutPad);
      t /= d.CoutPad;
      const int kb = (int)(t % nkb);
      const int tap = (int)(t / nkb);
      const int c = cphys ^ (co & 7);
#pragma unroll
      for (int e = 0; e < E::EPC; ++e) {
        const int k = kb * E::KBE + c * E::EPC + e;
        f[e] = (co < d.Cout && k < d.Cin) ? d.w[((long)co * d.Cin + k) * d.taps + tap] : 0.f;
      }
      // split mode: per (tap, kb) block the hi image (CoutPad rows) is followed by the lo image
      const long blk = (long)tap * nkb + kb;
      const long in_blk = ((long)co * 8 + cphys) * 16;
      dst = reinterpret_cast<char*>(d.fwd) + blk * d.CoutPad * 128 * (E::SPLIT ? 2 : 1) + in_blk;
      lo_off = (long)d.CoutPad * 128;
    } else {
      // dgrad image [chunk][kb][128 rows = input channel][8 chunks], K index = tap*coutk + co
      const long j = i - n_fwd;
      const int cphys = (int)(j & 7);
      long t = j >> 3;
      const int row = (int)(t & 127);
      t >>= 7;
      const int kb = (int)(t % nkbg);
      const int chunk = (int)(t / nkbg);
      const int c = cphys ^ (row & 7);
      const int k = chunk * 128 + row;
#pragma unroll
      for (int e = 0; e < E::EPC; ++e) {
        const int kg = kb * E::KBE + c * E::EPC + e;
        const int tap = kg / coutk, co = kg - tap * coutk;
        f[e] = (tap < d.taps && co < d.Cout && k < d.Cin) ? d.w[((long)co * d.Cin + k) * d.taps + tap] : 0.f;
      }
      const long blk = (long)chunk * nkbg + kb;
      const long in_blk = ((long)row * 8 + cphys) * 16;
      dst = reinterpret_cast<char*>(d.dgrad) + blk * 16384 * (E::SPLIT ? 2 : 1) + in_blk;
      lo_off = 16384;
    }
    *reinterpret_cast<uint4*>(dst) = Chunk<T>::pack_mma(f);
    if (E::SPLIT) *reinterpret_cast<uint4*>(dst + lo_off) = Chunk<T>::pack_lo(f);
  }
}

}  // namespace cunet
using namespace cunet;

extern "C" long cunet_pack_fwd_bytes(int Cin, int taps, int CoutPad, int dtype) {
  const int nkb = dtype == CUNET_BF16 ? PackGeom<bf16>::nkb_fwd(Cin) : PackGeom<float>::nkb_fwd(Cin);
  return (long)taps * nkb * CoutPad * 128 * (dtype == CUNET_BF16 ? 1 : 2);
}
extern "C" long cunet_pack_dgrad_bytes(int Cin, int taps, int CoutPad, int dtype) {
  // Cout is needed only for 3x3 (coutk = Cout = 32 in this network); pass CoutPad == Cout there.
  const int nkbg = dtype == CUNET_BF16 ? PackGeom<bf16>::nkb_dgrad(taps, CoutPad, CoutPad)
                                       : PackGeom<float>::nkb_dgrad(taps, CoutPad, CoutPad);
  return (long)((Cin + 127) / 128) * nkbg * 128 * 128 * (dtype == CUNET_BF16 ? 1 : 2);
}

extern "C" int cunet_pack_weights(const cunet_pack_desc* descs_dev, int ndesc, int dtype, int max_chunks, void* stream) {
  if (ndesc <= 0) return 0;
  dim3 grid(64, ndesc);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (dtype == CUNET_BF16)
    pack_weights_kernel<bf16><<<grid, 256, 0, st>>>(descs_dev, max_chunks);
  else
    pack_weights_kernel<float><<<grid, 256, 0, st>>>(descs_dev, max_chunks);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cunet_fail_cuda("pack_weights launch", e);
  return 0;
}
