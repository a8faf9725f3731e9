// Warp-level scans for the per-ray kernels (one warp owns one ray; its samples live in shared memory).
//
// The torch CPU path the oracle restates accumulates float cumsums in double (at::acc_type<float,false>) and
// rounds every prefix to float; these helpers do the same (double partials, float results) so cdf / cumulative
// weights agree with the oracle to the last bit except for double-rounding ties.
#pragma once
#include <cuda_runtime.h>

#define CNB_FULL 0xffffffffu

// ---- 8 elements per lane (n in 225..256: the 256-sample proposal level) -------------------------------------------------------
// A lane's chunk in[8 lane .. 8 lane + 8) starts in bank 8 (lane % 4): walking the chunks in step makes the eight lanes with equal
// lane % 4 hit the same bank (8-way conflict on every access, three passes per scan).  Here every lane starts its walk at element
// q = lane / 4 of its chunk (bank 8 (lane % 4) + (k + q) % 8: all 32 distinct), keeps the chunk in registers, and a 3-stage barrel
// rotation puts the values back into element order, so the additions happen in exactly the order of the generic code below.
__device__ __forceinline__ void cnb_rot8_right(float (&v)[8], int q) {  // v[j] <- v[(j - q) mod 8]
#pragma unroll
  for (int bit = 1; bit < 8; bit <<= 1) {
    const bool on = (q & bit) != 0;
    float t[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) t[j] = v[(j - bit) & 7];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = on ? t[j] : v[j];
  }
}
__device__ __forceinline__ void cnb_chunk8_load(const float* in, int n, int lane, float (&v)[8]) {
  const int b = lane * 8, q = lane >> 2;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int idx = b + ((k + q) & 7);
    v[k] = idx < n ? in[idx] : 0.0f;  // v[k] = element (k + q) mod 8
  }
  cnb_rot8_right(v, q);               // v[j] = element j
}
__device__ __forceinline__ void cnb_chunk8_store(float* out, int n, int lane, float (&r)[8]) {
  const int b = lane * 8, q = lane >> 2;
  cnb_rot8_right(r, (8 - q) & 7);     // r[k] = result of element (k + q) mod 8
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int idx = b + ((k + q) & 7);
    if (idx < n) out[idx] = r[k];
  }
}
static __device__ __noinline__ double cnb_warp_cumsum8(const float* in, float* out, int n, int lane) {
  float v[8];
  cnb_chunk8_load(in, n, lane, v);
  const int cnt = min(8, max(0, n - lane * 8));  // elements this lane owns (adding the zero padding would be exact too; keep the op count equal)
  double s = 0.0;
#pragma unroll
  for (int j = 0; j < 8; ++j) if (j < cnt) s += (double)v[j];
  double incl = s;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const double t = __shfl_up_sync(CNB_FULL, incl, off);
    if (lane >= off) incl += t;
  }
  double run = incl - s;
  float r[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { if (j < cnt) run += (double)v[j]; r[j] = (float)run; }
  cnb_chunk8_store(out, n, lane, r);
  return __shfl_sync(CNB_FULL, incl, 31);
}
static __device__ __noinline__ void cnb_warp_suffix_excl8(const float* in, float* out, int n, int lane) {
  float v[8];
  cnb_chunk8_load(in, n, lane, v);
  const int cnt = min(8, max(0, n - lane * 8));
  double s = 0.0;
#pragma unroll
  for (int j = 0; j < 8; ++j) if (j < cnt) s += (double)v[j];
  double incl = s;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const double t = __shfl_down_sync(CNB_FULL, incl, off);
    if (lane + off < 32) incl += t;
  }
  double run = incl - s;
  float r[8];
#pragma unroll
  for (int j = 7; j >= 0; --j) { r[j] = (float)run; if (j < cnt) run += (double)v[j]; }
  cnb_chunk8_store(out, n, lane, r);
}

// inclusive cumsum of in[0..n) -> out[0..n) (out may alias in). Lane owns a contiguous chunk.
// WIDE = the caller may see the 256-sample proposal level: take the conflict-free 8-per-lane path there (same results either way; kernels
// that only ever scan the field's 48 samples instantiate WIDE = false and keep the callee out of their register allocation).
template <bool WIDE>
__device__ __forceinline__ double cnb_warp_cumsum_t(const float* in, float* out, int n, int lane) {
  const int per = (n + 31) >> 5;
  if (WIDE && per == 8) return cnb_warp_cumsum8(in, out, n, lane);
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
template <bool WIDE>
__device__ __forceinline__ void cnb_warp_suffix_excl_t(const float* in, float* out, int n, int lane) {
  const int per = (n + 31) >> 5;
  if (WIDE && per == 8) { cnb_warp_suffix_excl8(in, out, n, lane); return; }
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

__device__ __forceinline__ double cnb_warp_cumsum(const float* in, float* out, int n, int lane) { return cnb_warp_cumsum_t<true>(in, out, n, lane); }
__device__ __forceinline__ double cnb_warp_cumsum_chunked(const float* in, float* out, int n, int lane) { return cnb_warp_cumsum_t<false>(in, out, n, lane); }
__device__ __forceinline__ void cnb_warp_suffix_excl(const float* in, float* out, int n, int lane) { cnb_warp_suffix_excl_t<true>(in, out, n, lane); }
__device__ __forceinline__ void cnb_warp_suffix_excl_chunked(const float* in, float* out, int n, int lane) { cnb_warp_suffix_excl_t<false>(in, out, n, lane); }

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
