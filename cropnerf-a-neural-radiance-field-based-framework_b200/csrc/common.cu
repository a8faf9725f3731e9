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

// Device -> device staging of one step's inputs + the optimiser's per-step scalars in ONE kernel launch: the graphed training step reads
// its rays / targets from static buffers, so a device-resident batch has to be copied in every step.  Five framework copies and one tiny
// H2D copy of the scalars cost ~45 us of stream time per step (measured: graph A + graph B = 0.796 ms inside a 0.845 ms step); here the
// buffers are copied by one grid and the scalars travel as kernel parameters.
namespace {
constexpr int STAGE_MAX = 8, STAGE_SCALARS = 32;
struct StageArgs {
  void* dst[STAGE_MAX];
  const void* src[STAGE_MAX];
  int64_t bytes[STAGE_MAX];
  int n;
  float* scalars_dst;
  int n_scalars;
  float scalars[STAGE_SCALARS];
};
__global__ void __launch_bounds__(256) k_stage_inputs(const __grid_constant__ StageArgs a) {
  const int b = blockIdx.y;
  if (b < a.n) {
    const int64_t n16 = a.bytes[b] >> 4;
    const uint4* s = reinterpret_cast<const uint4*>(a.src[b]);
    uint4* d = reinterpret_cast<uint4*>(a.dst[b]);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n16; i += (int64_t)gridDim.x * blockDim.x) d[i] = __ldg(s + i);
    if (blockIdx.x == 0) {  // tail bytes
      const unsigned char* sb = reinterpret_cast<const unsigned char*>(a.src[b]);
      unsigned char* db = reinterpret_cast<unsigned char*>(a.dst[b]);
      for (int64_t i = (n16 << 4) + threadIdx.x; i < a.bytes[b]; i += blockDim.x) db[i] = sb[i];
    }
  } else if (blockIdx.x == 0 && threadIdx.x < a.n_scalars) {
    a.scalars_dst[threadIdx.x] = a.scalars[threadIdx.x];
  }
}
}  // namespace

extern "C" int cnb_stage_inputs(void* const* dst, const void* const* src, const int64_t* bytes, int32_t n, float* scalars_dst, const float* scalars,
                                int32_t n_scalars, cnb_stream_t stream) {
  CNB_REQUIRE(n >= 0 && n <= STAGE_MAX && (n == 0 || (dst && src && bytes)), "stage_inputs: 0..%d buffers", STAGE_MAX);
  CNB_REQUIRE(n_scalars >= 0 && n_scalars <= STAGE_SCALARS && (n_scalars == 0 || (scalars_dst && scalars)), "stage_inputs: 0..%d scalars", STAGE_SCALARS);
  if (n == 0 && n_scalars == 0) return CNB_OK;
  StageArgs a;
  a.n = n; a.scalars_dst = scalars_dst; a.n_scalars = n_scalars;
  int64_t most = 0;
  for (int i = 0; i < STAGE_MAX; ++i) {
    a.dst[i] = i < n ? dst[i] : nullptr; a.src[i] = i < n ? src[i] : nullptr; a.bytes[i] = i < n ? bytes[i] : 0;
    if (i < n) {
      CNB_REQUIRE(bytes[i] >= 0 && (bytes[i] == 0 || (dst[i] && src[i])), "stage_inputs: null pointer / negative size in entry %d", i);
      CNB_REQUIRE(((((uintptr_t)dst[i]) | ((uintptr_t)src[i])) & 15) == 0, "stage_inputs: buffers must be 16-byte aligned (entry %d)", i);
      if (bytes[i] > most) most = bytes[i];
    }
  }
  for (int i = 0; i < STAGE_SCALARS; ++i) a.scalars[i] = i < n_scalars ? scalars[i] : 0.0f;
  int64_t bx = (most / 16 + 255) / 256;
  if (bx < 1) bx = 1;
  if (bx > 64) bx = 64;
  k_stage_inputs<<<dim3((unsigned)bx, (unsigned)(n + (n_scalars > 0 ? 1 : 0))), 256, 0, stream>>>(a);
  return cnb_check_launch("stage_inputs");
}
