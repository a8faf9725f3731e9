// Micro-benchmark (dev tool, NOT part of the library or the tests): candidate restructurings of the hash-grid gradient scatter
// (k_hashgrid_bwd, csrc/hashgrid.cu), evaluated stand-alone on field-shaped data before anything is changed in the library.
//
// Evidence it follows (ncu source page of the round-1 final build, profiles/r1_final_changed_kernels_ncu_full.csv + DESIGN.md section 7):
// 41 % of the kernel's stall samples are long-scoreboard waits -- 14 % on the first use of the d(features) load alone -- because a warp
// has ONE (32 samples, level) work item in flight: load d -> test -> load the position -> cell -> segmented scan -> reds, then the next
// item.  (Another 10 % sit on LDC reloads of the d_table pointer in front of every reduction: ptxas re-materialises the kernel parameter
// instead of keeping it in a register, also when the source copies it into a local first -- checked in the SASS of this file -- so that
// one needs a different parameter path and is not tried here.)
//   variant A  = the library's loop, verbatim
//   variant B  = A with the NEXT item's d(features) and position requested before the current item is processed (+8 registers),
//                32-bit work-item division
// Both use the library's own cnb_cell / cnb_scatter_cell, so the reductions issued are identical; the resulting tables are compared
// (differences = fp32 atomic ordering only).
//
// Build + run on a B200 (from tests/micro):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o scatter_pipeline_bench scatter_pipeline_bench.cu && ./scatter_pipeline_bench
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../cropnerf-a-neural-radiance-field-based-framework_b200/csrc/cnb_common.cuh"

struct GridArgs {
  float* d_table;
  int32_t L;
  uint32_t mask;
  uint32_t T;
  float scalings[CNB_MAX_LEVELS];
};

// ---- A: the library's kernel (level-major d_out) ------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_scatter_a(const __grid_constant__ GridArgs g, const float* __restrict__ pos, const float* __restrict__ d_out, int64_t n) {
  const int lane = threadIdx.x & 31;
  const int64_t nblk = (n + 31) >> 5;
  const int64_t nwork = nblk * g.L;
  const int64_t wstride = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t w = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5); w < nwork; w += wstride) {
    const int64_t sb = w / g.L;
    const int l = (int)(w - sb * g.L);
    const int64_t s = sb * 32 + lane;
    bool active = s < n;
    float2 d = make_float2(0.f, 0.f);
    if (active) {
      d = __ldg(reinterpret_cast<const float2*>(d_out) + (int64_t)l * n + s);
      active = d.x != 0.0f || d.y != 0.0f;
    }
    CnbCell c = {};
    if (active) c = cnb_cell(__ldg(pos + 3 * s), __ldg(pos + 3 * s + 1), __ldg(pos + 3 * s + 2), g.scalings[l]);
    cnb_scatter_cell(g.d_table, c, g.mask, (uint32_t)l * g.T, d.x, d.y, active);
  }
}

// ---- B: next item's inputs in flight while the current one is scattered -----------------------------------------------------
struct Item {
  float2 d;
  float x, y, z;
  int l;
  bool in;
};
__device__ __forceinline__ Item fetch(const GridArgs& g, const float* __restrict__ pos, const float* __restrict__ d_out, int64_t n, int64_t w, int64_t nwork,
                                      int lane) {
  Item it;
  it.d = make_float2(0.f, 0.f);
  it.x = it.y = it.z = 0.f;
  it.l = 0;
  it.in = false;
  if (w < nwork) {
    const uint32_t w32 = (uint32_t)w, L = (uint32_t)g.L;  // nwork < 2^32 for every batch this library sees (checked by the host below)
    const uint32_t sb = w32 / L;
    it.l = (int)(w32 - sb * L);
    const int64_t s = (int64_t)sb * 32 + lane;
    if (s < n) {
      it.in = true;
      it.d = __ldg(reinterpret_cast<const float2*>(d_out) + (int64_t)it.l * n + s);
      it.x = __ldg(pos + 3 * s); it.y = __ldg(pos + 3 * s + 1); it.z = __ldg(pos + 3 * s + 2);  // unconditional: no dependent round trip
    }
  }
  return it;
}
__global__ void __launch_bounds__(256) k_scatter_pf(const __grid_constant__ GridArgs g, const float* __restrict__ pos, const float* __restrict__ d_out, int64_t n) {
  const int lane = threadIdx.x & 31;
  const int64_t nblk = (n + 31) >> 5;
  const int64_t nwork = nblk * g.L;
  const int64_t wstride = (int64_t)gridDim.x * (blockDim.x >> 5);
  int64_t w = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
  Item cur = fetch(g, pos, d_out, n, w, nwork, lane);
  for (; w < nwork; w += wstride) {
    const Item nxt = fetch(g, pos, d_out, n, w + wstride, nwork, lane);
    const bool active = cur.in && (cur.d.x != 0.0f || cur.d.y != 0.0f);
    CnbCell c = {};
    if (active) c = cnb_cell(cur.x, cur.y, cur.z, g.scalings[cur.l]);
    cnb_scatter_cell(g.d_table, c, g.mask, (uint32_t)cur.l * g.T, cur.d.x, cur.d.y, active);
    cur = nxt;
  }
}

static float frand() { return (float)rand() / (float)RAND_MAX; }

int main() {
  const int R = 4096, S = 48, L = 16, LOG2T = 19;
  const int64_t n = (int64_t)R * S;
  if (((n + 31) / 32) * L >= (1ll << 32)) { printf("work count exceeds 32 bits\n"); return 1; }
  // field-shaped samples: 48 samples clustered around a surface point on each ray through the unit cube (what PDF resampling produces)
  std::vector<float> pos(3 * n), dout(2 * (size_t)L * n);
  srand(1);
  for (int r = 0; r < R; ++r) {
    float o[3] = {frand(), frand(), frand()}, d[3] = {frand() - 0.5f, frand() - 0.5f, frand() - 0.5f};
    const float nd = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]) + 1e-6f;
    const float t0 = 0.1f * frand();
    for (int s = 0; s < S; ++s) {
      const float t = t0 + 0.004f * s * (0.5f + frand());
      for (int k = 0; k < 3; ++k) {
        float v = o[k] + d[k] / nd * t;
        v = v - floorf(v);
        pos[3 * ((size_t)r * S + s) + k] = fminf(fmaxf(v, 1e-4f), 1.0f - 1e-4f);
      }
    }
  }
  for (auto& v : dout) v = (frand() < 0.05f) ? 0.0f : (frand() - 0.5f);
  GridArgs g;
  g.L = L; g.T = 1u << LOG2T; g.mask = g.T - 1u;
  const double growth = exp((log(2048.0) - log(16.0)) / (L - 1));
  for (int l = 0; l < CNB_MAX_LEVELS; ++l) g.scalings[l] = l < L ? (float)floor(16.0 * pow(growth, l)) : 0.f;
  const size_t tab_bytes = (size_t)L * g.T * 2 * sizeof(float);
  float *d_pos, *d_dout, *tab[2];
  cudaMalloc(&d_pos, pos.size() * 4); cudaMalloc(&d_dout, dout.size() * 4);
  for (int k = 0; k < 2; ++k) cudaMalloc(&tab[k], tab_bytes);
  cudaMemcpy(d_pos, pos.data(), pos.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(d_dout, dout.data(), dout.size() * 4, cudaMemcpyHostToDevice);
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int64_t threads = ((n + 31) / 32) * 32 * L;
  const int full = (int)((threads + 255) / 256);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto run = [&](const char* name, int which, int blocks, float* table) {
    g.d_table = table;
    float best = 1e9f;
    for (int rep = 0; rep < 8; ++rep) {
      cudaMemsetAsync(table, 0, tab_bytes);
      cudaEventRecord(e0);
      if (which == 0) k_scatter_a<<<blocks, 256>>>(g, d_pos, d_dout, n);
      else k_scatter_pf<<<blocks, 256>>>(g, d_pos, d_dout, n);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (rep > 1 && ms < best) best = ms;
    }
    printf("%-44s blocks %6d  %8.4f ms\n", name, blocks, best);
  };
  // the library caps the grid at 16 blocks per SM (grid_for): ~5 items per warp; also try one resident wave so that B has more items to prefetch
  const int cap = sms * 16 < full ? sms * 16 : full;
  run("A library loop", 0, cap, tab[0]);
  run("B next item prefetched", 1, cap, tab[1]);
  run("A library loop, 8 blocks/SM", 0, sms * 8, tab[0]);
  run("B next item prefetched, 8 blocks/SM", 1, sms * 8, tab[1]);
  // correctness: same reductions, different order
  std::vector<float> h[2];
  double ref = 0.0, err = 0.0;
  for (int k = 0; k < 2; ++k) { h[k].resize(tab_bytes / 4); cudaMemcpy(h[k].data(), tab[k], tab_bytes, cudaMemcpyDeviceToHost); }
  for (size_t i = 0; i < h[0].size(); ++i) {
    ref = fmax(ref, fabs((double)h[0][i]));
    err = fmax(err, fabs((double)h[1][i] - (double)h[0][i]));
  }
  printf("max |grad| %.4g   max |B - A| %.3g   (fp32 atomic ordering: expect ~1e-6 of max)\n", ref, err);
  const cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
  return err < 1e-4 * ref ? 0 : 2;
}
