// Multiresolution hash-grid encoding, forward and backward (row a2 of SURVEY.md section 8).
// Replaces nerfstudio/field_components/encodings.py HashEncoding.pytorch_fwd + its autograd
// (constructed at fruit_field.py:125-132 and, through HashMLPDensityField, fruit_nerf.py:124-141).
//
// One thread per (sample, level): 8 independent 8-byte gathers in flight per thread, coalesced 8-byte
// stores of the [n, 2L] feature rows.  HBM/L2-bound: 8 corners * 8 B per (sample, level).
#include "cnb_common.cuh"

namespace {

struct GridArgs {
  const float* table;
  float* d_table;
  int32_t L;
  uint32_t mask;
  uint32_t T;
  float scalings[CNB_MAX_LEVELS];
};

GridArgs make_args(const cnb_grid* g) {
  GridArgs a;
  a.table = g->table;
  a.d_table = g->d_table;
  a.L = g->num_levels;
  a.T = 1u << g->log2_hashmap_size;
  a.mask = a.T - 1u;
  for (int i = 0; i < CNB_MAX_LEVELS; ++i) a.scalings[i] = g->scalings[i];
  return a;
}

// Forward: a warp owns 32 CONSECUTIVE samples of one level, so one load instruction touches few distinct 128-byte lines at
// the coarse / middle levels (the gathers are bound by the L1TEX t-stage: one wavefront per distinct line per request).
__global__ void __launch_bounds__(256) k_hashgrid_fwd(const __grid_constant__ GridArgs g, const float* __restrict__ pos, int64_t n, float* __restrict__ out,
                                                       int32_t* __restrict__ indices) {
  const int lane = threadIdx.x & 31;
  const int64_t nblk = (n + 31) >> 5;
  const int64_t nwork = nblk * g.L;
  const int64_t wstride = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t w = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5); w < nwork; w += wstride) {
    const int64_t sb = w / g.L;
    const int l = (int)(w - sb * g.L);
    const int64_t s = sb * 32 + lane;
    if (s >= n) continue;
    const int64_t i = s * g.L + l;
    const float x = __ldg(pos + 3 * s), y = __ldg(pos + 3 * s + 1), z = __ldg(pos + 3 * s + 2);
    const CnbCell c = cnb_cell(x, y, z, g.scalings[l]);
    uint32_t h[8];
    cnb_corner_rows(c, g.mask, (uint32_t)l * g.T, h);
    float2 v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = cnb_ldg2(g.table, h[k]);
    float f0[8], f1[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { f0[k] = v[k].x; f1[k] = v[k].y; }
    float2 r;
    r.x = cnb_blend(f0, c.ox, c.oy, c.oz);
    r.y = cnb_blend(f1, c.ox, c.oy, c.oz);
    reinterpret_cast<float2*>(out)[i] = r;
    if (indices) {
      int4* dst = reinterpret_cast<int4*>(indices + i * 8);
      dst[0] = make_int4((int)h[0], (int)h[1], (int)h[2], (int)h[3]);
      dst[1] = make_int4((int)h[4], (int)h[5], (int)h[6], (int)h[7]);
    }
  }
}

// Backward: a warp owns 32 CONSECUTIVE samples of one level (samples are ray-major, so neighbouring lanes are
// neighbouring samples of a ray and often share a cell): cnb_scatter_cell aggregates those runs before the reductions.
__global__ void __launch_bounds__(256) k_hashgrid_bwd(const __grid_constant__ GridArgs g, const float* __restrict__ pos, const float* __restrict__ d_out, int64_t n) {
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
      d = __ldg(reinterpret_cast<const float2*>(d_out) + s * g.L + l);
      active = d.x != 0.0f || d.y != 0.0f;  // zero gradients add nothing (masked samples, App. B-3)
    }
    CnbCell c = {};
    if (active) c = cnb_cell(__ldg(pos + 3 * s), __ldg(pos + 3 * s + 1), __ldg(pos + 3 * s + 2), g.scalings[l]);
    cnb_scatter_cell(g.d_table, c, g.mask, (uint32_t)l * g.T, d.x, d.y, active);
  }
}

int grid_for(int64_t work_items, int block) {
  int64_t blocks = (work_items + block - 1) / block;
  int64_t cap = (int64_t)cnb_num_sms() * 16;  // multiple of the SM count; grid-stride beyond that
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

}  // namespace

static int check_grid(const cnb_grid* g, bool need_grad) {
  CNB_REQUIRE(g != nullptr && g->table != nullptr, "hashgrid: null grid/table");
  CNB_REQUIRE(g->num_levels >= 1 && g->num_levels <= CNB_MAX_LEVELS, "hashgrid: num_levels %d outside 1..%d", g->num_levels, CNB_MAX_LEVELS);
  CNB_REQUIRE(g->log2_hashmap_size >= 1 && g->log2_hashmap_size <= 24, "hashgrid: log2_hashmap_size %d outside 1..24", g->log2_hashmap_size);
  CNB_REQUIRE(!need_grad || g->d_table != nullptr, "hashgrid: backward needs d_table");
  return CNB_OK;
}

extern "C" int cnb_hashgrid_fwd(const cnb_grid* g, const float* positions, int64_t n, float* out, int32_t* indices, cnb_stream_t stream) {
  int rc = check_grid(g, false);
  if (rc) return rc;
  CNB_REQUIRE(n >= 0 && (n == 0 || (positions && out)), "hashgrid_fwd: null positions/out");
  if (n == 0) return CNB_OK;
  GridArgs a = make_args(g);
  k_hashgrid_fwd<<<grid_for(((n + 31) / 32) * 32 * a.L, 256), 256, 0, stream>>>(a, positions, n, out, indices);
  return cnb_check_launch("hashgrid_fwd");
}

extern "C" int cnb_hashgrid_bwd(const cnb_grid* g, const float* positions, const float* d_out, int64_t n, cnb_stream_t stream) {
  int rc = check_grid(g, true);
  if (rc) return rc;
  CNB_REQUIRE(n >= 0 && (n == 0 || (positions && d_out)), "hashgrid_bwd: null positions/d_out");
  if (n == 0) return CNB_OK;
  GridArgs a = make_args(g);
  for (int i = 0; i < a.L; ++i) CNB_REQUIRE(a.scalings[i] < 65535.0f, "hashgrid_bwd: level resolution %g too large for the aggregated scatter", a.scalings[i]);
  k_hashgrid_bwd<<<grid_for(((n + 31) / 32) * 32 * a.L, 256), 256, 0, stream>>>(a, positions, d_out, n);
  return cnb_check_launch("hashgrid_bwd");
}
