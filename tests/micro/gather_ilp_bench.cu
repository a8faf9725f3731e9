// Micro-benchmark (dev tool, NOT part of the library or the tests): how many corner loads must a thread have in flight for the
// proposal-network gather at the TRAINING batch size?
//
// Evidence it follows (profiles/r1_final_stalls_top_kernels.txt): at 4096 rays k_density_fwd spends 52 % of its stall
// samples on long-scoreboard waits and its SASS shows the reason -- the level loop is serialised: 8 LDG.64 of one level, the blend that
// consumes them, then the next level's 8 loads.  A sample therefore pays five dependent L2 round trips.  At the 32 768-ray render size the
// same kernel is at the L1 wavefront limit instead, so more loads in flight can only help the small grids of a training step.
//   variant 0 : one level's 8 loads in flight (the library's order)
//   variant 1 : the same loop behind 40 `prefetch.global.L1` of every corner row of the sample (no registers, cannot be sunk by ptxas;
//               costs a second L1 lookup per row, which is why it is a training-size measure only)
//   variant 2 : two levels' 16 loads pinned in front of their blends with a warp barrier (64 registers instead of 40).  ptxas otherwise
//               sinks grouped loads back to their consumers whatever the source order -- every plain-C++ grouping compiled to 40
//               registers and the library's load order (checked in the SASS here)
// Same arithmetic (cnb_cell / cnb_corner_rows / cnb_blend of the library) and bit-identical features in every variant (checked).
//
// Build + run on a B200 (from tests/micro):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o gather_ilp_bench gather_ilp_bench.cu && ./gather_ilp_bench
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../cropnerf-a-neural-radiance-field-based-framework_b200/csrc/cnb_common.cuh"

constexpr int LEVELS = 5;

struct Grid5 {
  const float* table;
  uint32_t mask, T;
  float scalings[LEVELS];
};

template <int GROUP, bool PREFETCH>
__global__ void __launch_bounds__(128) k_gather(const __grid_constant__ Grid5 g, const float* __restrict__ pos, int64_t n, float* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float x = __ldg(pos + 3 * i), y = __ldg(pos + 3 * i + 1), z = __ldg(pos + 3 * i + 2);
    float feat[2 * LEVELS];
    if (PREFETCH) {
#pragma unroll
      for (int l = 0; l < LEVELS; ++l) {
        const CnbCell c = cnb_cell(x, y, z, g.scalings[l]);
        uint32_t h[8];
        cnb_corner_rows(c, g.mask, (uint32_t)l * g.T, h);
#pragma unroll
        for (int k = 0; k < 8; ++k) asm volatile("prefetch.global.L1 [%0];" ::"l"(reinterpret_cast<const float2*>(g.table) + h[k]));
      }
    }
#pragma unroll
    for (int l0 = 0; l0 < LEVELS; l0 += GROUP) {
      CnbCell c[GROUP];
      float2 v[GROUP][8];
#pragma unroll
      for (int q = 0; q < GROUP; ++q) {   // every load of the group is issued ...
        if (l0 + q < LEVELS) {
          c[q] = cnb_cell(x, y, z, g.scalings[l0 + q]);
          uint32_t h[8];
          cnb_corner_rows(c[q], g.mask, (uint32_t)(l0 + q) * g.T, h);
#pragma unroll
          for (int k = 0; k < 8; ++k) v[q][k] = cnb_ldg2(g.table, h[k]);
        }
      }
      // ptxas sinks loads back to their consumers to save registers whatever the PTX order (40 registers for every variant without this line);
      // it does not move them across a warp barrier, so the barrier pins "all loads of the group first"
      if (GROUP > 1) __syncwarp();
#pragma unroll
      for (int q = 0; q < GROUP; ++q) {   // ... before the first blend consumes one
        if (l0 + q < LEVELS) {
          float f0[8], f1[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) { f0[k] = v[q][k].x; f1[k] = v[q][k].y; }
          feat[2 * (l0 + q)] = cnb_blend(f0, c[q].ox, c[q].oy, c[q].oz);
          feat[2 * (l0 + q) + 1] = cnb_blend(f1, c[q].ox, c[q].oy, c[q].oz);
        }
      }
    }
    // level-major, coalesced (the layout cnb_density_field_fwd_keep writes)
#pragma unroll
    for (int l = 0; l < LEVELS; ++l) reinterpret_cast<float2*>(out)[(int64_t)l * n + i] = make_float2(feat[2 * l], feat[2 * l + 1]);
  }
}

static float frand() { return (float)rand() / (float)RAND_MAX; }

int main() {
  const int S = 256, LOG2T = 17;
  const int64_t sizes[2] = {4096ll * S, 32768ll * S};
  const int64_t nmax = sizes[1];
  std::vector<float> pos(3 * nmax);
  srand(2);
  for (int64_t r = 0; r < nmax / S; ++r) {  // 256 samples marching along a ray through the unit cube
    float o[3] = {frand(), frand(), frand()}, d[3] = {frand() - 0.5f, frand() - 0.5f, frand() - 0.5f};
    for (int s = 0; s < S; ++s)
      for (int k = 0; k < 3; ++k) {
        float v = o[k] + d[k] * (float)s / S;
        v -= floorf(v);
        pos[3 * (r * S + s) + k] = fminf(fmaxf(v, 1e-4f), 1.0f - 1e-4f);
      }
  }
  Grid5 g;
  g.T = 1u << LOG2T; g.mask = g.T - 1u;
  const float sc[LEVELS] = {16.f, 26.f, 45.f, 76.f, 128.f};  // proposal network 0 of the fruit_nerf preset
  for (int l = 0; l < LEVELS; ++l) g.scalings[l] = sc[l];
  const size_t tab_floats = (size_t)LEVELS * g.T * 2;
  std::vector<float> tab(tab_floats);
  for (auto& v : tab) v = frand() - 0.5f;
  float *d_pos, *d_tab, *d_out[5];
  char* flush;
  cudaMalloc(&d_pos, pos.size() * 4); cudaMalloc(&d_tab, tab_floats * 4); cudaMalloc(&flush, 256u << 20);
  for (int k = 0; k < 5; ++k) cudaMalloc(&d_out[k], (size_t)nmax * LEVELS * 2 * 4);
  cudaMemcpy(d_pos, pos.data(), pos.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(d_tab, tab.data(), tab_floats * 4, cudaMemcpyHostToDevice);
  g.table = d_tab;
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto run = [&](int which, int64_t n) {
    int64_t blocks = (n + 127) / 128;
    if (blocks > (int64_t)sms * 16) blocks = (int64_t)sms * 16;
    float best = 1e9f;
    for (int rep = 0; rep < 6; ++rep) {
      cudaMemsetAsync(flush, rep, 256u << 20);  // tables out of L2 before every repetition, as between training steps
      cudaEventRecord(e0);
      if (which == 0) k_gather<1, false><<<(int)blocks, 128>>>(g, d_pos, n, d_out[0]);
      else if (which == 1) k_gather<1, true><<<(int)blocks, 128>>>(g, d_pos, n, d_out[1]);
      else if (which == 2) k_gather<2, false><<<(int)blocks, 128>>>(g, d_pos, n, d_out[2]);
      else if (which == 3) k_gather<3, false><<<(int)blocks, 128>>>(g, d_pos, n, d_out[3]);
      else k_gather<5, false><<<(int)blocks, 128>>>(g, d_pos, n, d_out[4]);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (rep > 0 && ms < best) best = ms;
    }
    return best;
  };
  const char* names[5] = {"0 library order (8 loads)", "1 + 40 L1 prefetches up front", "2 two levels pinned (16 loads)", "3 three levels pinned (24 loads)",
                          "4 five levels pinned (40 loads)"};
  for (int s = 0; s < 2; ++s) {
    printf("---- %lld samples (%lld rays x %d) ----\n", (long long)sizes[s], (long long)(sizes[s] / S), S);
    for (int w = 0; w < 5; ++w) {
      const float ms = run(w, sizes[s]);
      printf("%-32s %8.4f ms   %7.2f G corner fetches/s\n", names[w], ms, sizes[s] * 40.0 / (ms * 1e-3) / 1e9);
    }
  }
  // bit-identical features
  const size_t bytes = (size_t)nmax * LEVELS * 2 * 4;
  std::vector<float> h0(bytes / 4), hk(bytes / 4);
  cudaMemcpy(h0.data(), d_out[0], bytes, cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int k = 1; k < 5; ++k) {
    cudaMemcpy(hk.data(), d_out[k], bytes, cudaMemcpyDeviceToHost);
    if (memcmp(h0.data(), hk.data(), bytes) != 0) { printf("variant %d differs from variant 0\n", k); bad = 1; }
  }
  if (!bad) printf("all variants bit-identical\n");
  const cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
  return bad;
}
