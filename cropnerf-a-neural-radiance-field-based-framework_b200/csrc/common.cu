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

// Host -> device staging for the graphed training step (engine._GraphedStep): n async copies on `stream` with one call, no
// framework dispatch in between.  src[i] should be pinned host memory (otherwise the copy is staged synchronously by the driver).
extern "C" int cnb_upload(void* const* dst, const void* const* src, const int64_t* bytes, int32_t n, cnb_stream_t stream) {
  CNB_REQUIRE(n >= 0 && (n == 0 || (dst && src && bytes)), "upload: null arrays");
  for (int i = 0; i < n; ++i) {
    if (bytes[i] == 0) continue;
    CNB_REQUIRE(dst[i] && src[i] && bytes[i] > 0, "upload: null pointer / negative size in entry %d", i);
    if (cudaMemcpyAsync(dst[i], src[i], (size_t)bytes[i], cudaMemcpyHostToDevice, stream) != cudaSuccess) return cnb_check_launch("upload");
  }
  return CNB_OK;
}
