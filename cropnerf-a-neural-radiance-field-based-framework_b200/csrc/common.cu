// Error plumbing + version for the C ABI (include/cropnerf_b200.h).
#include <cstdarg>
#include <cstdio>

#include "cnb_common.cuh"

static thread_local char g_err[512] = "";

void cnb_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cnb_check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    cnb_set_error("%s: %s", what, cudaGetErrorString(e));
    return CNB_ERR_CUDA;
  }
  return CNB_OK;
}

extern "C" int cnb_version(void) { return CNB_VERSION; }
extern "C" const char* cnb_last_error(void) { return g_err; }
extern "C" int cnb_device_count(void) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    cnb_set_error("cudaGetDeviceCount: %s", cudaGetErrorString(e));
    (void)cudaGetLastError();
    return CNB_ERR_CUDA;
  }
  return n;
}
