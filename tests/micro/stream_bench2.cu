// micro-benchmark 2: in-place 4-array update, address-layout variants (dev tool)
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("err %s line %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)
__global__ void __launch_bounds__(256) k_inplace4(float4* __restrict__ p, float4* __restrict__ g, float4* __restrict__ m, float4* __restrict__ v, long n4) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
    float4 gi = g[i], mi = m[i], vi = v[i], pi = p[i];
    mi.x += gi.x; vi.x += gi.y; pi.x += mi.x * vi.x; gi.x = 0;
    g[i] = gi; m[i] = mi; v[i] = vi; p[i] = pi;
  }
}
// each block owns contiguous chunks of CH float4 per array
template <int CH>
__global__ void __launch_bounds__(256) k_chunk(float4* __restrict__ p, float4* __restrict__ g, float4* __restrict__ m, float4* __restrict__ v, long n4) {
  const long nchunks = (n4 + CH - 1) / CH;
  for (long c = blockIdx.x; c < nchunks; c += gridDim.x) {
    const long base = c * CH;
#pragma unroll 1
    for (int j = threadIdx.x; j < CH; j += 256) {
      const long i = base + j;
      if (i < n4) {
        float4 gi = g[i], mi = m[i], vi = v[i], pi = p[i];
        mi.x += gi.x; vi.x += gi.y; pi.x += mi.x * vi.x; gi.x = 0;
        g[i] = gi; m[i] = mi; v[i] = vi; p[i] = pi;
      }
    }
  }
}
// interleaved layout: one array of struct {p,g,m,v} float4 x4 (64 B per 4 params)
__global__ void __launch_bounds__(256) k_aos(float4* __restrict__ a, long n4) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
    float4 pi = a[4 * i], gi = a[4 * i + 1], mi = a[4 * i + 2], vi = a[4 * i + 3];
    mi.x += gi.x; vi.x += gi.y; pi.x += mi.x * vi.x; gi.x = 0;
    a[4 * i] = pi; a[4 * i + 1] = gi; a[4 * i + 2] = mi; a[4 * i + 3] = vi;
  }
}
// 2 in-place + 2 in-place in two passes
__global__ void __launch_bounds__(256) k_inplace2(float4* __restrict__ a, float4* __restrict__ b, long n4) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
    float4 x = a[i], y = b[i];
    x.x += y.x; y.y += x.x;
    a[i] = x; b[i] = y;
  }
}
__global__ void __launch_bounds__(256) k_inplace3(float4* __restrict__ a, float4* __restrict__ b, float4* __restrict__ c, long n4) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
    float4 x = a[i], y = b[i], z = c[i];
    x.x += y.x; y.y += x.x; z.x += y.y;
    a[i] = x; b[i] = y; c[i] = z;
  }
}
int main() {
  const long n = 16777216 + 65536;
  const long n4 = n / 4;
  char* big; char* flush;
  const size_t slot = (size_t)n * 4 + (8 << 20);
  CK(cudaMalloc(&big, slot * 4 + (64 << 20))); CK(cudaMemset(big, 0, slot * 4 + (64 << 20)));
  CK(cudaMalloc(&flush, 256 << 20));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto run = [&](const char* name, auto launch, double bytes_per_param) {
    float best = 1e9;
    for (int it = 0; it < 6; ++it) {
      cudaMemsetAsync(flush, it, 256 << 20);
      cudaEventRecord(e0);
      launch();
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (it > 0) best = ms < best ? ms : best;
    }
    printf("%-34s best %.1f us -> %.2f TB/s\n", name, best * 1e3, n * bytes_per_param / (best * 1e-3) / 1e12);
  };
  const size_t pads[][4] = {{0, 0, 0, 0}, {0, 4096, 8192, 12288}, {0, 1 << 20, 2 << 20, 3 << 20}, {0, 256, 512, 768}, {0, 65536 + 1024, 131072 + 2048, 196608 + 3072}, {0, 3 << 19, 5 << 19, 7 << 19}};
  for (auto& pd : pads) {
    float4* a[4];
    for (int i = 0; i < 4; ++i) a[i] = (float4*)(big + slot * i + pd[i]);
    char nm[96]; snprintf(nm, sizeof(nm), "inplace4 pad {%zu,%zu,%zu,%zu}", pd[0], pd[1], pd[2], pd[3]);
    run(nm, [&] { k_inplace4<<<148 * 16, 256>>>(a[0], a[1], a[2], a[3], n4); }, 32.0);
  }
  float4* a[4];
  for (int i = 0; i < 4; ++i) a[i] = (float4*)(big + slot * i);
  run("chunk 1024 float4 (16 KB)", [&] { k_chunk<1024><<<148 * 16, 256>>>(a[0], a[1], a[2], a[3], n4); }, 32.0);
  run("chunk 4096 float4 (64 KB)", [&] { k_chunk<4096><<<148 * 16, 256>>>(a[0], a[1], a[2], a[3], n4); }, 32.0);
  run("chunk 256 float4 (4 KB)", [&] { k_chunk<256><<<148 * 16, 256>>>(a[0], a[1], a[2], a[3], n4); }, 32.0);
  run("array-of-structs interleaved", [&] { k_aos<<<148 * 16, 256>>>((float4*)big, n4); }, 32.0);
  run("inplace2", [&] { k_inplace2<<<148 * 16, 256>>>(a[0], a[1], n4); }, 16.0);
  run("inplace3", [&] { k_inplace3<<<148 * 16, 256>>>(a[0], a[1], a[2], n4); }, 24.0);
  run("inplace2 x2 passes", [&] { k_inplace2<<<148 * 16, 256>>>(a[0], a[1], n4); k_inplace2<<<148 * 16, 256>>>(a[2], a[3], n4); }, 32.0);
  return 0;
}
