#include "host_util.h"
#include "../../include/cunet_b200.h"
#include <stdio.h>
#include <string.h>

static thread_local char g_err[512] = "";

int cunet_fail(const char* msg) {
  snprintf(g_err, sizeof(g_err), "%s", msg);
  return -1;
}
int cunet_fail_cuda(const char* where, cudaError_t e) {
  snprintf(g_err, sizeof(g_err), "%s: %s", where, cudaGetErrorString(e));
  return -2;
}
extern "C" const char* cunet_last_error(void) { return g_err; }
extern "C" int cunet_abi_version(void) { return 1; }
