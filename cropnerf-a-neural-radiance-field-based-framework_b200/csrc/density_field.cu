// Proposal-network density, fused: position warp -> hash grid (L levels) -> MLP 2L->H->1 -> trunc_exp * selector.
// Row a7 of SURVEY.md section 8: replaces nerfstudio/fields/density_fields.py HashMLPDensityField.get_density /
// Field.density_fn as built at fruit_nerf.py:118-142 and called by the ProposalNetworkSampler (fruit_nerf.py:549).
// 70 % of all corner fetches of a training step happen here (256+96 of 400 samples per ray).
//
// One thread per sample; everything between the ray description and the density stays in registers.  The tiny
// MLP (193 parameters for H=16) is read from shared memory as warp-wide broadcasts.  Backward recomputes the
// forward, scatters the table gradient with vector reductions (red.global.add.v2.f32) and reduces the MLP
// parameter gradients per CTA through a shared-memory tile before one atomic flush per CTA.
#include "cnb_common.cuh"

namespace {

constexpr int BLOCK = 128;

struct DfArgs {
  const float* table;
  float* d_table;
  int L;
  uint32_t mask, T;
  float scalings[CNB_MAX_LEVELS];
  cnb_warp warp;
  float avg;
  const float *W1, *b1, *W2, *b2;
  float *dW1, *db1, *dW2, *db2;
  int in;  // 2L
  cnb_samples sm;
};

template <int LMAX>
__device__ __forceinline__ void encode(const DfArgs& a, float x, float y, float z, float (&feat)[2 * LMAX]) {
#pragma unroll
  for (int l = 0; l < LMAX; ++l) {
    if (l < a.L) {
      const CnbCell c = cnb_cell(x, y, z, a.scalings[l]);
      uint32_t h[8];
      cnb_corner_rows(c, a.mask, (uint32_t)l * a.T, h);
      float f0[8], f1[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) { const float2 v = cnb_ldg2(a.table, h[k]); f0[k] = v.x; f1[k] = v.y; }
      feat[2 * l] = cnb_blend(f0, c.ox, c.oy, c.oz);
      feat[2 * l + 1] = cnb_blend(f1, c.ox, c.oy, c.oz);
    } else {
      feat[2 * l] = 0.0f;
      feat[2 * l + 1] = 0.0f;
    }
  }
}

// smem weights: W1 [H][2*LMAX] (zero padded), b1 [H], W2 [H], b2 [4]
template <int LMAX, int H>
__device__ __forceinline__ void load_weights(const DfArgs& a, float* Ws) {
  constexpr int INP = 2 * LMAX;
  for (int e = threadIdx.x; e < H * INP; e += BLOCK) {
    const int j = e / INP, k = e - j * INP;
    Ws[e] = k < a.in ? __ldg(a.W1 + j * a.in + k) : 0.0f;
  }
  for (int j = threadIdx.x; j < H; j += BLOCK) {
    Ws[H * INP + j] = __ldg(a.b1 + j);
    Ws[H * INP + H + j] = __ldg(a.W2 + j);
  }
  if (threadIdx.x == 0) Ws[H * INP + 2 * H] = __ldg(a.b2);
}

template <int LMAX, int H>
__device__ __forceinline__ float mlp_forward(const float* Ws, const float (&feat)[2 * LMAX], float (&hid)[H]) {
  constexpr int INP = 2 * LMAX;
  float out = Ws[H * INP + 2 * H];
#pragma unroll
  for (int j = 0; j < H; ++j) {
    float acc = Ws[H * INP + j];
#pragma unroll
    for (int k = 0; k < INP; k += 4) {
      const float4 w = *reinterpret_cast<const float4*>(Ws + j * INP + k);
      acc = fmaf(w.x, feat[k], acc); acc = fmaf(w.y, feat[k + 1], acc);
      acc = fmaf(w.z, feat[k + 2], acc); acc = fmaf(w.w, feat[k + 3], acc);
    }
    hid[j] = fmaxf(acc, 0.0f);
    out = fmaf(Ws[H * INP + H + j], hid[j], out);
  }
  return out;
}

template <int LMAX, int H>
__global__ void __launch_bounds__(BLOCK) k_density_fwd(DfArgs a, float* __restrict__ density, float* __restrict__ pos_out) {
  constexpr int INP = 2 * LMAX;
  __shared__ __align__(16) float Ws[H * INP + 2 * H + 4];
  load_weights<LMAX, H>(a, Ws);
  __syncthreads();
  const int S = a.sm.samples_per_ray;
  const int64_t total = a.sm.num_rays * S;
  for (int64_t i = blockIdx.x * (int64_t)BLOCK + threadIdx.x; i < total; i += (int64_t)gridDim.x * BLOCK) {
    const int64_t r = i / S;
    const int s = (int)(i - r * S);
    float x, y, z;
    const bool sel = cnb_sample_position(a.sm, a.warp, r, s, x, y, z);
    if (pos_out) { pos_out[3 * i] = x; pos_out[3 * i + 1] = y; pos_out[3 * i + 2] = z; }
    float feat[INP], hid[H];
    encode<LMAX>(a, x, y, z, feat);
    const float out = mlp_forward<LMAX, H>(Ws, feat, hid);
    density[i] = sel ? a.avg * expf(out) : 0.0f;
  }
}

// Backward.  Per tile of BLOCK samples the per-sample factors are staged in shared memory as
//   U_s = [dh_s (H) | g_s]      V_s = [x_s (in) | 1 | h_s (H)]
// and the parameter gradients are the pair sums (dh_j,x_k) (dh_j,1) (g,h_j) (g,1).
template <int LMAX, int H>
__global__ void __launch_bounds__(BLOCK, (H <= 16 && LMAX <= 8) ? 5 : 1) k_density_bwd(DfArgs a, const float* __restrict__ d_density) {
  constexpr int INP = 2 * LMAX;
  constexpr int LD = BLOCK + 1;
  extern __shared__ float4 smem4[];
  float* Ws = reinterpret_cast<float*>(smem4);       // H*INP + 2H + 4
  float* U = Ws + (H * INP + 2 * H + 4);             // (H+1) x LD
  float* V = U + (H + 1) * LD;                       // (in+1+H) x LD
  const int nv = a.in + 1 + H;
  float* acc = V + nv * LD;                          // nout
  const int nout = H * a.in + 2 * H + 1;
  load_weights<LMAX, H>(a, Ws);
  for (int o = threadIdx.x; o < nout; o += BLOCK) acc[o] = 0.0f;
  __syncthreads();
  const int tid = threadIdx.x;
  const int S = a.sm.samples_per_ray;
  const int64_t total = a.sm.num_rays * S;
  const int64_t ntiles = (total + BLOCK - 1) / BLOCK;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t i = tile * BLOCK + tid;
    float g = 0.0f;
    float feat[INP], hid[H];
    float x = 0.f, y = 0.f, z = 0.f;
    bool active = false;
    if (i < total) {
      const float dd = __ldg(d_density + i);
      if (dd != 0.0f) {
        const int64_t r = i / S;
        const int s = (int)(i - r * S);
        const bool sel = cnb_sample_position(a.sm, a.warp, r, s, x, y, z);
        if (sel) {
          encode<LMAX>(a, x, y, z, feat);
          const float out = mlp_forward<LMAX, H>(Ws, feat, hid);
          g = dd * a.avg * cnb_trunc_exp_grad(out);
          active = (g != 0.0f);
        }
      }
    }
    float dfeat[INP];
#pragma unroll
    for (int k = 0; k < INP; ++k) dfeat[k] = 0.0f;
    if (active) {
#pragma unroll
      for (int j = 0; j < H; ++j) {
        const float dh = hid[j] > 0.0f ? g * Ws[H * INP + H + j] : 0.0f;
        U[j * LD + tid] = dh;
        V[(a.in + 1 + j) * LD + tid] = hid[j];
#pragma unroll
        for (int k = 0; k < INP; k += 4) {
          const float4 w = *reinterpret_cast<const float4*>(Ws + j * INP + k);
          dfeat[k] = fmaf(dh, w.x, dfeat[k]); dfeat[k + 1] = fmaf(dh, w.y, dfeat[k + 1]);
          dfeat[k + 2] = fmaf(dh, w.z, dfeat[k + 2]); dfeat[k + 3] = fmaf(dh, w.w, dfeat[k + 3]);
        }
      }
      U[H * LD + tid] = g;
#pragma unroll
      for (int k = 0; k < INP; ++k)
        if (k < a.in) V[k * LD + tid] = feat[k];
      V[a.in * LD + tid] = 1.0f;
    } else {
      for (int j = 0; j <= H; ++j) U[j * LD + tid] = 0.0f;
      for (int k = 0; k < nv; ++k) V[k * LD + tid] = 0.0f;
    }
    // table gradient: the lanes of a warp are consecutive samples of a ray -> warp-aggregated scatter (all lanes take part)
#pragma unroll
    for (int l = 0; l < LMAX; ++l) {
      if (l < a.L) {
        const float d0 = dfeat[2 * l], d1 = dfeat[2 * l + 1];
        const bool on = active && (d0 != 0.0f || d1 != 0.0f);
        CnbCell c = {};
        if (on) c = cnb_cell(x, y, z, a.scalings[l]);
        cnb_scatter_cell(a.d_table, c, a.mask, (uint32_t)l * a.T, d0, d1, on);
      }
    }
    __syncthreads();
    for (int o = tid; o < nout; o += BLOCK) {
      int u, v;
      if (o < H * a.in) { u = o / a.in; v = o - u * a.in; }
      else if (o < H * a.in + H) { u = o - H * a.in; v = a.in; }
      else if (o < H * a.in + 2 * H) { u = H; v = a.in + 1 + (o - H * a.in - H); }
      else { u = H; v = a.in; }
      const float* up = U + u * LD;
      const float* vp = V + v * LD;
      float sacc = 0.0f;
#pragma unroll 8
      for (int s = 0; s < BLOCK; ++s) sacc = fmaf(up[s], vp[s], sacc);
      acc[o] += sacc;
    }
    __syncthreads();
  }
  for (int o = tid; o < nout; o += BLOCK) {
    const float v = acc[o];
    if (v == 0.0f) continue;
    if (o < H * a.in) { if (a.dW1) atomicAdd(a.dW1 + o, v); }
    else if (o < H * a.in + H) { if (a.db1) atomicAdd(a.db1 + (o - H * a.in), v); }
    else if (o < H * a.in + 2 * H) { if (a.dW2) atomicAdd(a.dW2 + (o - H * a.in - H), v); }
    else { if (a.db2) atomicAdd(a.db2, v); }
  }
}

int make_args(const cnb_density_field* f, const cnb_samples* s, bool bwd, DfArgs& a) {
  CNB_REQUIRE(f && s, "density_field: null descriptor");
  const cnb_grid& g = f->grid;
  CNB_REQUIRE(g.table != nullptr, "density_field: null table");
  CNB_REQUIRE(g.num_levels >= 1 && g.num_levels <= CNB_MAX_LEVELS, "density_field: num_levels %d unsupported", g.num_levels);
  CNB_REQUIRE(g.log2_hashmap_size >= 1 && g.log2_hashmap_size <= 24, "density_field: log2_hashmap_size %d unsupported", g.log2_hashmap_size);
  const cnb_mlp& m = f->mlp;
  if (m.num_layers != 2 || m.dims[2] != 1 || m.out_activation != CNB_ACT_NONE) {
    cnb_set_error("density_field: fused kernel supports MLP 2L->H->1 only (got %d layers); compose cnb_hashgrid_* + cnb_mlp_* instead", m.num_layers);
    return CNB_ERR_UNSUPPORTED;
  }
  CNB_REQUIRE(m.dims[0] == 2 * g.num_levels, "density_field: mlp in dim %d != 2*num_levels", m.dims[0]);
  CNB_REQUIRE(m.W[0] && m.b[0] && m.W[1] && m.b[1], "density_field: null mlp weights");
  CNB_REQUIRE(s->origins && s->directions && s->starts && s->ends, "density_field: null sample arrays");
  CNB_REQUIRE(s->samples_per_ray >= 1 && s->num_rays >= 0, "density_field: bad sample counts");
  CNB_REQUIRE(!bwd || g.d_table != nullptr, "density_field_bwd: d_table required");
  if (bwd)
    for (int i = 0; i < g.num_levels; ++i) CNB_REQUIRE(g.scalings[i] < 65535.0f, "density_field_bwd: level resolution %g too large for the aggregated scatter", g.scalings[i]);
  a.table = g.table; a.d_table = g.d_table; a.L = g.num_levels;
  a.T = 1u << g.log2_hashmap_size; a.mask = a.T - 1u;
  for (int i = 0; i < CNB_MAX_LEVELS; ++i) a.scalings[i] = g.scalings[i];
  a.warp = f->warp; a.avg = f->average_init_density;
  a.W1 = m.W[0]; a.b1 = m.b[0]; a.W2 = m.W[1]; a.b2 = m.b[1];
  a.dW1 = m.dW[0]; a.db1 = m.db[0]; a.dW2 = m.dW[1]; a.db2 = m.db[1];
  a.in = m.dims[0];
  a.sm = *s;
  return CNB_OK;
}

template <int LMAX, int H>
int launch_fwd(const DfArgs& a, float* density, float* pos_out, cudaStream_t st) {
  const int64_t total = a.sm.num_rays * a.sm.samples_per_ray;
  int64_t blocks = (total + BLOCK - 1) / BLOCK;
  const int64_t cap = (int64_t)cnb_num_sms() * 16;
  if (blocks > cap) blocks = cap;
  k_density_fwd<LMAX, H><<<(int)blocks, BLOCK, 0, st>>>(a, density, pos_out);
  return cnb_check_launch("density_field_fwd");
}

template <int LMAX, int H>
int launch_bwd(const DfArgs& a, const float* d_density, cudaStream_t st) {
  constexpr int INP = 2 * LMAX;
  const int64_t total = a.sm.num_rays * a.sm.samples_per_ray;
  const int nv = a.in + 1 + H, nout = H * a.in + 2 * H + 1;
  const size_t smem = sizeof(float) * ((H * INP + 2 * H + 4) + (size_t)(H + 1 + nv) * (BLOCK + 1) + nout + 4);
  static size_t configured = 0;
  if (smem > configured) {
    if (cudaFuncSetAttribute(k_density_bwd<LMAX, H>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
      return cnb_check_launch("density_field_bwd attr");
    configured = smem;
  }
  int64_t blocks = (total + BLOCK - 1) / BLOCK;
  const int64_t cap = (int64_t)cnb_num_sms() * 8;  // 8 resident CTAs/SM (smem 25 KB, 76 regs): the gathers and reds need the warps
  if (blocks > cap) blocks = cap;
  k_density_bwd<LMAX, H><<<(int)blocks, BLOCK, smem, st>>>(a, d_density);
  return cnb_check_launch("density_field_bwd");
}

#define DISPATCH(FN, ...)                                                                   \
  do {                                                                                      \
    const int H = f->mlp.dims[1];                                                           \
    const bool small = a.L <= 8;                                                            \
    if (small && H == 16) return FN<8, 16>(__VA_ARGS__);                                    \
    if (small && H == 32) return FN<8, 32>(__VA_ARGS__);                                    \
    if (small && H == 64) return FN<8, 64>(__VA_ARGS__);                                    \
    if (!small && H == 16) return FN<16, 16>(__VA_ARGS__);                                  \
    if (!small && H == 64) return FN<16, 64>(__VA_ARGS__);                                  \
    cnb_set_error("density_field: hidden width %d not compiled (16/32/64)", H);             \
    return CNB_ERR_UNSUPPORTED;                                                             \
  } while (0)

}  // namespace

extern "C" int cnb_density_field_fwd(const cnb_density_field* f, const cnb_samples* s, float* density, float* positions_out, cnb_stream_t stream) {
  DfArgs a;
  int rc = make_args(f, s, false, a);
  if (rc) return rc;
  if (a.sm.num_rays == 0) return CNB_OK;
  CNB_REQUIRE(density != nullptr, "density_field_fwd: null output");
  DISPATCH(launch_fwd, a, density, positions_out, stream);
}

extern "C" int cnb_density_field_bwd(const cnb_density_field* f, const cnb_samples* s, const float* d_density, cnb_stream_t stream) {
  DfArgs a;
  int rc = make_args(f, s, true, a);
  if (rc) return rc;
  if (a.sm.num_rays == 0) return CNB_OK;
  CNB_REQUIRE(d_density != nullptr, "density_field_bwd: null d_density");
  DISPATCH(launch_bwd, a, d_density, stream);
}
