// Micro-benchmark (dev tool, NOT part of the library or the tests): the hardware ceilings of the hash-table access pattern.
//
// DESIGN.md section 4 argues that the forward gathers are bound by L1TEX sectors/clk/SM and the backward scatter by the L2
// reduction rate, both far from the HBM roofline because the 74 MiB of tables are L2-resident.  This program measures those two
// ceilings directly, with the library's own access shapes, so that `roofline.frac` of the gather / scatter kernels can be quoted
// against what the memory system delivers for this pattern:
//   gather : 8-byte rows (float2) at random row indices of a table of `rows` rows, 8 independent loads in flight per thread
//   scatter: red.global.add.v2.f32 / .v4.f32 to random rows (the two shapes cnb_scatter_cell issues)
// for three table footprints: one hash level that fits L1 (2^12 rows = 32 KiB), one field level (2^19 rows = 4 MiB) and the whole
// field table (16 x 2^19 rows = 64 MiB), plus a sequential pattern as the upper reference.
//
// Build + run on a B200:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o table_access_bench table_access_bench.cu && ./table_access_bench
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t mix(uint32_t x) {  // cheap integer hash (indices must not cost more than the access)
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}

template <bool RANDOM>
__global__ void __launch_bounds__(256) k_gather(const float2* __restrict__ table, uint32_t mask, int64_t n, float* __restrict__ sink) {
  float acc = 0.f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float2 v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const uint32_t row = RANDOM ? (mix((uint32_t)i * 8u + k) & mask) : (((uint32_t)i * 8u + k) & mask);
      v[k] = __ldg(table + row);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) acc += v[k].x + v[k].y;
  }
  if (acc == 123.456f) *sink = acc;
}

// the fp16-table question (SURVEY.md section 7 hard part (c)): the same gather with 4-byte rows (__half2) -- half the footprint, the same
// number of 32-byte sectors per warp request
template <bool RANDOM>
__global__ void __launch_bounds__(256) k_gather4(const uint32_t* __restrict__ table, uint32_t mask, int64_t n, float* __restrict__ sink) {
  uint32_t acc = 0u;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    uint32_t v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const uint32_t row = RANDOM ? (mix((uint32_t)i * 8u + k) & mask) : (((uint32_t)i * 8u + k) & mask);
      v[k] = __ldg(table + row);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) acc += v[k];
  }
  if (acc == 0x12345678u) *sink = (float)acc;
}

template <bool RANDOM, bool V4>
__global__ void __launch_bounds__(256) k_scatter(float2* __restrict__ table, uint32_t mask, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
#pragma unroll
    for (int k = 0; k < (V4 ? 4 : 8); ++k) {
      uint32_t row = RANDOM ? (mix((uint32_t)i * 8u + k) & mask) : (((uint32_t)i * 8u + k) & mask);
      if (V4) {
        row &= ~1u;
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(table + row), "f"(1.f), "f"(2.f), "f"(3.f), "f"(4.f) : "memory");
      } else {
        asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(table + row), "f"(1.f), "f"(2.f) : "memory");
      }
    }
  }
}

int main() {
  int dev = 0, sms = 0, khz = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
  const uint32_t max_rows = 16u << 19;
  float2* table; float* sink; char* flush;
  cudaMalloc(&table, (size_t)max_rows * 8); cudaMemset(table, 0, (size_t)max_rows * 8);
  cudaMalloc(&sink, 4); cudaMalloc(&flush, 256u << 20);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int64_t n = 4 << 20;  // threads' worth of work items: 8 accesses each = 33.5 M accesses (the field backward issues 16.4 M reds)
  const int blocks = sms * 8;
  auto time_it = [&](auto launch) {
    float best = 1e9f;
    for (int rep = 0; rep < 5; ++rep) {
      if (rep == 0) cudaMemsetAsync(flush, rep, 256u << 20);
      cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (rep > 0 && ms < best) best = ms;  // rep 0 = warm-up (cold L2)
    }
    return best;
  };
  printf("device: %d SMs, %.0f MHz nominal\n", sms, khz / 1e3);
  printf("%-34s %10s %12s %14s %16s\n", "pattern", "ms", "G access/s", "GB/s (8 B)", "access/clk/SM");
  const uint32_t sizes[3] = {1u << 12, 1u << 19, 16u << 19};
  const char* names[3] = {"32 KiB (L1)", "4 MiB (one level)", "64 MiB (field table)"};
  auto report = [&](const char* what, const char* size, float ms, double accesses) {
    const double rate = accesses / (ms * 1e-3);
    printf("%-14s %-19s %10.4f %12.2f %14.1f %16.3f\n", what, size, ms, rate / 1e9, rate * 8 / 1e9, rate / (khz * 1e3) / sms);
  };
  for (int s = 0; s < 3; ++s) {
    const uint32_t mask = sizes[s] - 1u;
    report("gather rand", names[s], time_it([&] { k_gather<true><<<blocks, 256>>>(table, mask, n, sink); }), 8.0 * n);
  }
  report("gather seq", names[2], time_it([&] { k_gather<false><<<blocks, 256>>>(table, max_rows - 1u, n, sink); }), 8.0 * n);
  // fp16 rows: the same ROW counts, half the bytes (the 16 x 2^19-row field table is 32 MiB); the big preset's 16 x 2^21 rows are 256 MiB in
  // fp32 (beyond the 126 MB L2) and 128 MiB in fp16 -- measured with the 8-byte gather on a 256 MiB table and the 4-byte one on 128 MiB
  const char* names4[3] = {"16 KiB fp16 (L1)", "2 MiB fp16 level", "32 MiB fp16 table"};
  for (int s = 0; s < 3; ++s) {
    const uint32_t mask = sizes[s] - 1u;
    report("gather4 rand", names4[s], time_it([&] { k_gather4<true><<<blocks, 256>>>(reinterpret_cast<const uint32_t*>(table), mask, n, sink); }), 8.0 * n);
  }
  {
    float2* big; cudaMalloc(&big, (size_t)(16u << 21) * 8); cudaMemset(big, 0, (size_t)(16u << 21) * 8);
    const uint32_t mask = (16u << 21) - 1u;
    report("gather rand", "256 MiB (big preset)", time_it([&] { k_gather<true><<<blocks, 256>>>(big, mask, n, sink); }), 8.0 * n);
    report("gather4 rand", "128 MiB fp16 (big)", time_it([&] { k_gather4<true><<<blocks, 256>>>(reinterpret_cast<const uint32_t*>(big), mask, n, sink); }), 8.0 * n);
    cudaFree(big);
  }
  for (int s = 0; s < 3; ++s) {
    const uint32_t mask = sizes[s] - 1u;
    report("red.v2 rand", names[s], time_it([&] { k_scatter<true, false><<<blocks, 256>>>(table, mask, n); }), 8.0 * n);
    report("red.v4 rand", names[s], time_it([&] { k_scatter<true, true><<<blocks, 256>>>(table, mask, n); }), 4.0 * n);
  }
  report("red.v2 seq", names[2], time_it([&] { k_scatter<false, false><<<blocks, 256>>>(table, max_rows - 1u, n); }), 8.0 * n);
  const cudaError_t err = cudaDeviceSynchronize();
  if (err != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(err)); return 1; }
  return 0;
}
