// Host-side error plumbing shared by all translation units of libcunet_b200.so.
#pragma once
#include <cuda_runtime.h>

int cunet_fail(const char* msg);
int cunet_fail_cuda(const char* where, cudaError_t e);
