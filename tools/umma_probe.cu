// Standalone probe: which shared-memory byte does tcgen05.mma read for element (mn, k) of an MN-major
// SWIZZLE_128B A operand, as a function of (LBO, SBO)?  Build: nvcc -arch=sm_100a -o umma_probe umma_probe.cu
// Method: smem is filled with bf16 codes of the element position; B (K-major, known-good) is one-hot at
// k = k0 so D[m][0] = A(m, k0).  Two passes recover low / high 7 bits of the position.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../cu-net_b200/csrc/common.cuh"
using namespace cunet;

__global__ void probe(float* out, int lbo, int sbo, int dtype_is_f32, int a_mn_major, int ltype) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  uint8_t* A = smem;                 // 64 KB probe area
  uint8_t* B = smem + 65536;         // 16 rows x 128 B K-major tile
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(&tmem_base, 32);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = tmem_base;
  const int esz = dtype_is_f32 ? 4 : 2;
  const int kk = dtype_is_f32 ? 8 : 16;
  uint32_t parity = 0;
  for (int pass = 0; pass < 2; ++pass) {
    // fill A area: element index e (in units of esz) -> code
    for (int e = tid; e < 65536 / esz; e += blockDim.x) {
      int code = pass == 0 ? (e & 127) : ((e >> 7) & 127) + ((e >> 14) ? 0 : 0);
      if (dtype_is_f32) ((float*)A)[e] = (float)code; else ((bf16*)A)[e] = __float2bfloat16((float)code);
    }
    for (int k0 = 0; k0 < kk; ++k0) {
      for (int e = tid; e < 16 * 128 / esz; e += blockDim.x) {
        if (dtype_is_f32) ((float*)B)[e] = 0.f; else ((bf16*)B)[e] = __float2bfloat16(0.f);
      }
      __syncthreads();
      if (tid == 0) {
        // row 0, k = k0 : chunk = k0*esz/16, swizzle with row 0 = identity
        if (dtype_is_f32) ((float*)B)[k0] = 1.f; else ((bf16*)B)[k0] = __float2bfloat16(1.f);
      }
      fence_proxy_async();
      __syncthreads();
      if (tid == 0) {
        tc_fence_after();
        uint32_t fmt = dtype_is_f32 ? 2 : 1;
        uint32_t idesc = make_idesc(fmt, 128, 16, a_mn_major, 0);
        uint64_t ad = (make_sdesc(smem_u32(A), lbo, sbo) & ~(7ull << 61)) | ((uint64_t)ltype << 61);
        uint64_t bd = make_sdesc(smem_u32(B), 16, 1024);
        if (dtype_is_f32) umma<float>(tmem, ad, bd, idesc, 0); else umma<bf16>(tmem, ad, bd, idesc, 0);
        tc_commit(&bar);
      }
      mbar_wait(&bar, parity);
      parity ^= 1;
      tc_fence_after();
      if (warp < 4) {
        float v[8];
        tmem_ld8(tmem + ((uint32_t)(warp * 32) << 16), v);
        out[(pass * kk + k0) * 128 + tid] = v[0];
      }
      tc_fence_before();
      __syncthreads();
    }
  }
  if (warp == 0) tmem_dealloc(tmem, 32);
}

int main(int argc, char** argv) {
  int cfgs[][5] = {
      // lbo, sbo, f32, a_mn_major, layout_type
      {4096, 1024, 1, 1, 1}, {4096, 512, 1, 1, 1}, {8192, 2048, 1, 1, 1}, {4096, 1024, 1, 1, 2}, {4096, 1024, 1, 1, 4}, {4096, 1024, 1, 1, 6}, {4096, 1024, 1, 1, 0}};
  float* d;
  cudaMalloc(&d, 2 * 16 * 128 * sizeof(float));
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 2048 + 1024);
  for (auto& c : cfgs) {
    cudaMemset(d, 0, 2 * 16 * 128 * sizeof(float));
    probe<<<1, 128, 65536 + 2048 + 1024>>>(d, c[0], c[1], c[2], c[3], c[4]);
    cudaError_t e = cudaDeviceSynchronize();
    printf("== lbo=%d sbo=%d f32=%d a_mn_major=%d ltype=%d : %s\n", c[0], c[1], c[2], c[3], c[4], cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    int kk = c[2] ? 8 : 16, esz = c[2] ? 4 : 2;
    std::vector<float> h(2 * 16 * 128);
    cudaMemcpy(h.data(), d, h.size() * 4, cudaMemcpyDeviceToHost);
    // print byte offset read for selected (m, k)
    int ms[] = {0, 1, 3, 4, 5, 7, 8, 9, 15, 16, 31, 32, 33, 63, 64, 65, 96, 127};
    for (int k0 = 0; k0 < kk; ++k0) {
      printf(" k=%2d:", k0);
      for (int m : ms) {
        int lo = (int)h[(0 * kk + k0) * 128 + m], hi = (int)h[(1 * kk + k0) * 128 + m];
        printf(" m%d->%d", m, (hi * 128 + lo) * esz);
      }
      printf("\n");
    }
  }
  return 0;
}
