// Warp-level scans for the per-ray kernels (one warp owns one ray; its samples live in shared memory).
//
// The torch CPU path the oracle restates accumulates float cumsums in double (at::acc_type<float,false>) and
// rounds every prefix to float; these helpers do the same (double partials, float results) so cdf / cumulative
// weights agree with the oracle to the last bit except for double-rounding ties.
#pragma once
#include <cuda_runtime.h>

#define CNB_FULL 0xffffffffu

// inclusive cumsum of in[0..n) -> out[0..n) (out may alias in). Lane owns a contiguous chunk.
__device__ __forceinline__ double cnb_warp_cumsum(const float* in, float* out, int n, int lane) {
  const int per = (n + 31) >> 5;
  const int b = min(n, lane * per), e = min(n, b + per);
  double s = 0.0;
  for (int j = b; j < e; ++j) s += (double)in[j];
  double incl = s;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const double t = __shfl_up_sync(CNB_FULL, incl, off);
    if (lane >= off) incl += t;
  }
  double run = incl - s;
  for (int j = b; j < e; ++j) { run += (double)in[j]; out[j] = (float)run; }
  return __shfl_sync(CNB_FULL, incl, 31);  // total
}

// suffix sums: out[j] = sum_{k > j} in[k]  (exclusive reverse cumsum), double partials
__device__ __forceinline__ void cnb_warp_suffix_excl(const float* in, float* out, int n, int lane) {
  const int per = (n + 31) >> 5;
  const int b = min(n, lane * per), e = min(n, b + per);
  double s = 0.0;
  for (int j = b; j < e; ++j) s += (double)in[j];
  double incl = s;  // inclusive scan from the top lane down
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const double t = __shfl_down_sync(CNB_FULL, incl, off);
    if (lane + off < 32) incl += t;
  }
  double run = incl - s;  // sum of all chunks above this lane
  for (int j = e - 1; j >= b; --j) { const float v = in[j]; out[j] = (float)run; run += (double)v; }
}

__device__ __forceinline__ float cnb_warp_sum(float v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(CNB_FULL, v, off);
  return v;
}
__device__ __forceinline__ double cnb_warp_sum_d(double v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(CNB_FULL, v, off);
  return v;
}

// torch.searchsorted(sorted[0..n), v, side="right"): number of entries <= v
__device__ __forceinline__ int cnb_search_right(const float* a, int n, float v) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (!(v < a[mid])) lo = mid + 1; else hi = mid;
  }
  return lo;
}
// side="left": number of entries < v
__device__ __forceinline__ int cnb_search_left(const float* a, int n, float v) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (a[mid] < v) lo = mid + 1; else hi = mid;
  }
  return lo;
}
