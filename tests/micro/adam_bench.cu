// micro-benchmark: variants of the fused Adam + grad-clear pass (dev tool, not part of the library)
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("err %s line %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ void upd(float4& gi, float4& mi, float4& vi, float4& pi, float step_size, float b1, float b2, float eps, float bc2s, float inv) {
  float* gp = &gi.x; float* mp = &mi.x; float* vp = &vi.x; float* pp = &pi.x;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float gk = gp[k] * inv;
    mp[k] = mp[k] + (1.0f - b1) * (gk - mp[k]);
    vp[k] = b2 * vp[k] + (1.0f - b2) * gk * gk;
    pp[k] -= step_size * (mp[k] / (sqrtf(vp[k]) / bc2s + eps));
  }
}

__global__ void __launch_bounds__(256) k_a(float4* __restrict__ p, float4* __restrict__ g, float4* __restrict__ m, float4* __restrict__ v, long n4) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
    float4 gi = g[i], mi = m[i], vi = v[i], pi = p[i];
    g[i] = make_float4(0, 0, 0, 0);
    upd(gi, mi, vi, pi, 1e-2f, 0.9f, 0.999f, 1e-15f, 0.5f, 1.f);
    m[i] = mi; v[i] = vi; p[i] = pi;
  }
}
// streaming hints
__global__ void __launch_bounds__(256) k_b(float4* __restrict__ p, float4* __restrict__ g, float4* __restrict__ m, float4* __restrict__ v, long n4) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
    float4 gi = __ldcs(g + i), mi = __ldcs(m + i), vi = __ldcs(v + i), pi = __ldcs(p + i);
    __stcs(g + i, make_float4(0, 0, 0, 0));
    upd(gi, mi, vi, pi, 1e-2f, 0.9f, 0.999f, 1e-15f, 0.5f, 1.f);
    __stcs(m + i, mi); __stcs(v + i, vi); __stcs(p + i, pi);
  }
}
// unroll 2: 8 loads in flight
__global__ void __launch_bounds__(256) k_c(float4* __restrict__ p, float4* __restrict__ g, float4* __restrict__ m, float4* __restrict__ v, long n4) {
  const long stride = (long)gridDim.x * blockDim.x;
  long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  for (; i + stride < n4; i += 2 * stride) {
    float4 g0 = g[i], m0 = m[i], v0 = v[i], p0 = p[i];
    float4 g1 = g[i + stride], m1 = m[i + stride], v1 = v[i + stride], p1 = p[i + stride];
    g[i] = make_float4(0, 0, 0, 0); g[i + stride] = make_float4(0, 0, 0, 0);
    upd(g0, m0, v0, p0, 1e-2f, 0.9f, 0.999f, 1e-15f, 0.5f, 1.f);
    upd(g1, m1, v1, p1, 1e-2f, 0.9f, 0.999f, 1e-15f, 0.5f, 1.f);
    m[i] = m0; v[i] = v0; p[i] = p0; m[i + stride] = m1; v[i + stride] = v1; p[i + stride] = p1;
  }
  if (i < n4) {
    float4 gi = g[i], mi = m[i], vi = v[i], pi = p[i];
    g[i] = make_float4(0, 0, 0, 0);
    upd(gi, mi, vi, pi, 1e-2f, 0.9f, 0.999f, 1e-15f, 0.5f, 1.f);
    m[i] = mi; v[i] = vi; p[i] = pi;
  }
}
// fast math (approx division / rsqrt)
__global__ void __launch_bounds__(256) k_d(float4* __restrict__ p, float4* __restrict__ g, float4* __restrict__ m, float4* __restrict__ v, long n4) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
    float4 gi = g[i], mi = m[i], vi = v[i], pi = p[i];
    g[i] = make_float4(0, 0, 0, 0);
    float* gp = &gi.x; float* mp = &mi.x; float* vp = &vi.x; float* pp = &pi.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float gk = gp[k];
      mp[k] = mp[k] + 0.1f * (gk - mp[k]);
      vp[k] = 0.999f * vp[k] + 0.001f * gk * gk;
      pp[k] -= 1e-2f * __fdividef(mp[k], __fsqrt_rn(vp[k]) * 2.0f + 1e-15f);
    }
    m[i] = mi; v[i] = vi; p[i] = pi;
  }
}
// g cleared LATE: the zero store carries a fake dependency on the computed parameter, so it cannot be hoisted above the math
__global__ void __launch_bounds__(256) k_f(float4* __restrict__ p, float4* __restrict__ g, float4* __restrict__ m, float4* __restrict__ v, long n4) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
    float4 gi = g[i], mi = m[i], vi = v[i], pi = p[i];
    upd(gi, mi, vi, pi, 1e-2f, 0.9f, 0.999f, 1e-15f, 0.5f, 1.f);
    float z;
    asm volatile("mov.f32 %0, 0f00000000;" : "=f"(z) : "f"(pi.w));
    m[i] = mi; v[i] = vi; p[i] = pi; g[i] = make_float4(z, z, z, z);
  }
}
__global__ void __launch_bounds__(256) k_g(float4* __restrict__ p, float4* __restrict__ g, float4* __restrict__ m, float4* __restrict__ v, long n4) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
    float4 gi = g[i], mi = m[i], vi = v[i], pi = p[i];
    upd(gi, mi, vi, pi, 1e-2f, 0.9f, 0.999f, 1e-15f, 0.5f, 1.f);
    float z;
    asm volatile("mov.f32 %0, 0f00000000;" : "=f"(z) : "f"(pi.w));
    __stwt(m + i, mi); __stwt(v + i, vi); __stwt(p + i, pi); __stwt(g + i, make_float4(z, z, z, z));
  }
}
// pure copy-like baseline: read 4 write 4, no math
__global__ void __launch_bounds__(256) k_e(float4* __restrict__ p, float4* __restrict__ g, float4* __restrict__ m, float4* __restrict__ v, long n4) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
    float4 gi = g[i], mi = m[i], vi = v[i], pi = p[i];
    g[i] = make_float4(0, 0, 0, 0);
    mi.x += gi.x; vi.x += gi.y; pi.x += mi.x * vi.x;
    m[i] = mi; v[i] = vi; p[i] = pi;
  }
}

int main() {
  const long n = 16777216 + 65536;
  const long n4 = n / 4;
  float4 *p, *g, *m, *v; char* flush;
  CK(cudaMalloc(&p, n * 4)); CK(cudaMalloc(&g, n * 4)); CK(cudaMalloc(&m, n * 4)); CK(cudaMalloc(&v, n * 4)); CK(cudaMalloc(&flush, 256 << 20));
  CK(cudaMemset(p, 0, n * 4)); CK(cudaMemset(g, 0, n * 4)); CK(cudaMemset(m, 0, n * 4)); CK(cudaMemset(v, 0, n * 4));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const char* names[] = {"a_current", "b_streaming", "c_unroll2", "d_fastmath", "e_nomath", "f_late_zero", "g_late_zero_wt"};
  for (int grid_mul : {16, 64}) {
    for (int var = 0; var < 7; ++var) {
      float best = 1e9, sum = 0;
      for (int it = 0; it < 6; ++it) {
        CK(cudaMemsetAsync(flush, it, 256 << 20));
        cudaEventRecord(e0);
        const int blocks = 148 * grid_mul;
        switch (var) {
          case 0: k_a<<<blocks, 256>>>(p, g, m, v, n4); break;
          case 1: k_b<<<blocks, 256>>>(p, g, m, v, n4); break;
          case 2: k_c<<<blocks, 256>>>(p, g, m, v, n4); break;
          case 3: k_d<<<blocks, 256>>>(p, g, m, v, n4); break;
          case 4: k_e<<<blocks, 256>>>(p, g, m, v, n4); break;
          case 5: k_f<<<blocks, 256>>>(p, g, m, v, n4); break;
          case 6: k_g<<<blocks, 256>>>(p, g, m, v, n4); break;
        }
        cudaEventRecord(e1);
        CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (it > 0) { best = ms < best ? ms : best; sum += ms; }
      }
      printf("grid %2dx148 %-12s best %.1f us  mean %.1f us  -> %.2f TB/s\n", grid_mul, names[var], best * 1e3, sum / 5 * 1e3, n * 32.0 / (best * 1e-3) / 1e12);
    }
  }
  // one block per chunk (non-persistent): n4/256 blocks
  {
    float best = 1e9;
    for (int it = 0; it < 6; ++it) {
      CK(cudaMemsetAsync(flush, it, 256 << 20));
      cudaEventRecord(e0);
      k_a<<<(int)((n4 + 255) / 256), 256>>>(p, g, m, v, n4);
      cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (it > 0) best = ms < best ? ms : best;
    }
    printf("one-iteration-per-thread a_current best %.1f us -> %.2f TB/s\n", best * 1e3, n * 32.0 / (best * 1e-3) / 1e12);
  }
  return 0;
}
