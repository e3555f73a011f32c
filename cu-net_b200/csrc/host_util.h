// Host-side error plumbing and the kernel-launch helper shared by all translation units of libcunet_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <utility>

int cunet_fail(const char* msg);
int cunet_fail_cuda(const char* where, cudaError_t e);
// 1 when kernels are chained with programmatic dependent launch (env CUNET_PDL, see host_util.cu)
int cunet_pdl_enabled();

// kernel<<<grid, block, smem, st>>>(args...) with, optionally, the programmatic-stream-serialization attribute: the
// kernel may then start while its stream predecessor is still running and MUST execute griddep_wait() (common.cuh)
// before it touches global memory.  Works under stream capture (the edge becomes a programmatic graph dependency).
template <typename... KArgs, typename... Args>
inline cudaError_t cunet_launch(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  if (cunet_pdl_enabled()) {
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
  }
  return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}
