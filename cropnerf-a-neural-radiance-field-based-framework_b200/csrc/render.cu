// Volume-rendering compositing (rows a10-a14 of SURVEY.md section 8).
//  - cnb_weights_fwd/bwd : nerfstudio rays.py RaySamples.get_weights (fruit_nerf.py:556,508,442,341)
//  - cnb_render_fwd/bwd  : nerfstudio renderers.py RGBRenderer ("last_sample"/override background),
//    DepthRenderer(method="median"), AccumulationRenderer, SemanticRenderer (fruit_nerf.py:170-174,560-591)
// One warp per ray: a transmittance prefix scan forward, a matching suffix scan backward.  HBM-bound:
// a ray reads S*(density, start, end, rgb[3], sem) once and writes S weights + 6 scalars.
#include "cnb_common.cuh"
#include "warp_scan.cuh"

namespace {

constexpr int WARPS = 4;

__global__ void __launch_bounds__(WARPS * 32) k_weights_fwd(const float* __restrict__ density, const float* __restrict__ starts,
                                                            const float* __restrict__ ends, int64_t row_stride, int64_t R, int S,
                                                            float* __restrict__ weights) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* dd = smem + (size_t)warp * 2 * S;
  float* cs = dd + S;
  for (int64_t r = blockIdx.x * (int64_t)WARPS + warp; r < R; r += (int64_t)gridDim.x * WARPS) {
    for (int j = lane; j < S; j += 32) {
      const float delta = __fsub_rn(__ldg(ends + r * row_stride + j), __ldg(starts + r * row_stride + j));
      dd[j] = __fmul_rn(delta, __ldg(density + r * S + j));
    }
    __syncwarp();
    cnb_warp_cumsum(dd, cs, S, lane);  // inclusive; exclusive prefix of j is cs[j-1]
    __syncwarp();
    for (int j = lane; j < S; j += 32) {
      const float alpha = __fsub_rn(1.0f, expf(-dd[j]));
      const float T = expf(-(j > 0 ? cs[j - 1] : 0.0f));
      weights[r * S + j] = cnb_nan_to_num(__fmul_rn(alpha, T));
    }
    __syncwarp();
  }
}

// dL/d dd_j = d_w_j * T_j * exp(-dd_j) - sum_{k>j} d_w_k * w_k ;  d_density_j = delta_j * dL/d dd_j
__global__ void __launch_bounds__(WARPS * 32) k_weights_bwd(const float* __restrict__ density, const float* __restrict__ starts,
                                                            const float* __restrict__ ends, int64_t row_stride, int64_t R, int S,
                                                            const float* __restrict__ d_weights, float* __restrict__ d_density) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* dd = smem + (size_t)warp * 3 * S;
  float* cs = dd + S;
  float* gw = cs + S;
  for (int64_t r = blockIdx.x * (int64_t)WARPS + warp; r < R; r += (int64_t)gridDim.x * WARPS) {
    for (int j = lane; j < S; j += 32) {
      const float delta = __fsub_rn(__ldg(ends + r * row_stride + j), __ldg(starts + r * row_stride + j));
      dd[j] = __fmul_rn(delta, __ldg(density + r * S + j));
    }
    __syncwarp();
    cnb_warp_cumsum(dd, cs, S, lane);
    __syncwarp();
    for (int j = lane; j < S; j += 32) {
      const float T = expf(-(j > 0 ? cs[j - 1] : 0.0f));
      const float w = __fmul_rn(__fsub_rn(1.0f, expf(-dd[j])), T);
      const float g = __ldg(d_weights + r * S + j);
      gw[j] = isfinite(w) ? g * w : 0.0f;
    }
    __syncwarp();
    cnb_warp_suffix_excl(gw, gw, S, lane);
    __syncwarp();
    for (int j = lane; j < S; j += 32) {
      const float T = expf(-(j > 0 ? cs[j - 1] : 0.0f));
      const float e = expf(-dd[j]);
      const float w = __fmul_rn(__fsub_rn(1.0f, e), T);
      const float g = isfinite(w) ? __ldg(d_weights + r * S + j) : 0.0f;
      const float delta = __fsub_rn(__ldg(ends + r * row_stride + j), __ldg(starts + r * row_stride + j));
      d_density[r * S + j] = delta * (g * T * e - gw[j]);
    }
    __syncwarp();
  }
}

__global__ void __launch_bounds__(WARPS * 32) k_render_fwd(const float* __restrict__ weights, const float* __restrict__ rgb,
                                                           const float* __restrict__ sem, const float* __restrict__ starts,
                                                           const float* __restrict__ ends, int64_t row_stride, int64_t R, int S, int bg_mode,
                                                           float bg0, float bg1, float bg2, int eval_mode, float* __restrict__ rgb_out,
                                                           float* __restrict__ depth_out, float* __restrict__ acc_out, float* __restrict__ sem_out,
                                                           int32_t* __restrict__ median_index) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* wsm = smem + (size_t)warp * S;
  for (int64_t r = blockIdx.x * (int64_t)WARPS + warp; r < R; r += (int64_t)gridDim.x * WARPS) {
    float a = 0.f, c0 = 0.f, c1 = 0.f, c2 = 0.f, sm = 0.f;
    for (int j = lane; j < S; j += 32) {
      const float w = __ldg(weights + r * S + j);
      wsm[j] = w;
      a += w;
      if (rgb) {
        float x = __ldg(rgb + (r * S + j) * 3), y = __ldg(rgb + (r * S + j) * 3 + 1), z = __ldg(rgb + (r * S + j) * 3 + 2);
        if (eval_mode) { x = cnb_nan_to_num(x); y = cnb_nan_to_num(y); z = cnb_nan_to_num(z); }
        c0 = fmaf(w, x, c0); c1 = fmaf(w, y, c1); c2 = fmaf(w, z, c2);
      }
      if (sem) sm = fmaf(w, __ldg(sem + r * S + j), sm);
    }
    a = cnb_warp_sum(a);
    if (acc_out && lane == 0) acc_out[r] = a;
    if (sem && sem_out) { sm = cnb_warp_sum(sm); if (lane == 0) sem_out[r] = sm; }
    if (rgb && rgb_out) {
      c0 = cnb_warp_sum(c0); c1 = cnb_warp_sum(c1); c2 = cnb_warp_sum(c2);
      if (lane == 0) {
        if (bg_mode == CNB_BG_LAST_SAMPLE) {
          const float* last = rgb + (r * S + S - 1) * 3;
          bg0 = last[0]; bg1 = last[1]; bg2 = last[2];
          if (eval_mode) { bg0 = cnb_nan_to_num(bg0); bg1 = cnb_nan_to_num(bg1); bg2 = cnb_nan_to_num(bg2); }
        }
        if (bg_mode != CNB_BG_NONE) {
          const float rem = 1.0f - a;
          c0 += bg0 * rem; c1 += bg1 * rem; c2 += bg2 * rem;
        }
        if (eval_mode) { c0 = fminf(fmaxf(c0, 0.f), 1.f); c1 = fminf(fmaxf(c1, 0.f), 1.f); c2 = fminf(fmaxf(c2, 0.f), 1.f); }
        rgb_out[3 * r] = c0; rgb_out[3 * r + 1] = c1; rgb_out[3 * r + 2] = c2;
      }
    }
    if (depth_out || median_index) {
      __syncwarp();
      cnb_warp_cumsum(wsm, wsm, S, lane);
      __syncwarp();
      if (lane == 0) {
        int idx = cnb_search_left(wsm, S, 0.5f);
        idx = min(max(idx, 0), S - 1);
        if (median_index) median_index[r] = idx;
        if (depth_out)
          depth_out[r] = __fmul_rn(__fadd_rn(__ldg(starts + r * row_stride + idx), __ldg(ends + r * row_stride + idx)), 0.5f);
      }
    }
    __syncwarp();
  }
}

// one thread per sample: gradients of comp_rgb / accumulation / semantics wrt weights, rgb, sem
__global__ void __launch_bounds__(WARPS * 32) k_render_bwd(const float* __restrict__ weights, const float* __restrict__ rgb,
                                                           const float* __restrict__ sem, int64_t R, int S, int bg_mode, float bg0, float bg1,
                                                           float bg2, const float* __restrict__ d_rgb_out, const float* __restrict__ d_acc_out,
                                                           const float* __restrict__ d_sem_out, int sem_weight_grad, float* __restrict__ d_weights,
                                                           float* __restrict__ d_rgb, float* __restrict__ d_sem) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int64_t r = blockIdx.x * (int64_t)WARPS + warp; r < R; r += (int64_t)gridDim.x * WARPS) {
    float g0 = 0.f, g1 = 0.f, g2 = 0.f;
    if (d_rgb_out) { g0 = __ldg(d_rgb_out + 3 * r); g1 = __ldg(d_rgb_out + 3 * r + 1); g2 = __ldg(d_rgb_out + 3 * r + 2); }
    const float ga = d_acc_out ? __ldg(d_acc_out + r) : 0.0f;
    const float gs = d_sem_out ? __ldg(d_sem_out + r) : 0.0f;
    float b0 = bg0, b1 = bg1, b2 = bg2;
    float rem = 0.0f;
    if (rgb && bg_mode == CNB_BG_LAST_SAMPLE) {
      const float* last = rgb + (r * S + S - 1) * 3;
      b0 = __ldg(last); b1 = __ldg(last + 1); b2 = __ldg(last + 2);
      if (d_rgb) {
        float a = 0.f;
        for (int j = lane; j < S; j += 32) a += __ldg(weights + r * S + j);
        rem = 1.0f - cnb_warp_sum(a);
      }
    }
    if (bg_mode == CNB_BG_NONE) { b0 = b1 = b2 = 0.0f; }
    for (int j = lane; j < S; j += 32) {
      const int64_t i = r * S + j;
      const float w = __ldg(weights + i);
      float gw = ga;
      if (rgb) {
        const float x = __ldg(rgb + 3 * i), y = __ldg(rgb + 3 * i + 1), z = __ldg(rgb + 3 * i + 2);
        gw += g0 * (x - b0) + g1 * (y - b1) + g2 * (z - b2);
        if (d_rgb) {
          float e = (bg_mode == CNB_BG_LAST_SAMPLE && j == S - 1) ? rem : 0.0f;
          d_rgb[3 * i] = g0 * (w + e); d_rgb[3 * i + 1] = g1 * (w + e); d_rgb[3 * i + 2] = g2 * (w + e);
        }
      }
      if (sem) {
        if (sem_weight_grad) gw += gs * __ldg(sem + i);
        if (d_sem) d_sem[i] = gs * w;
      }
      if (d_weights) d_weights[i] = gw;
    }
  }
}

int ray_grid(int64_t R) {
  int64_t blocks = (R + WARPS - 1) / WARPS;
  const int64_t cap = (int64_t)cnb_num_sms() * 16;
  if (blocks > cap) blocks = cap;
  return (int)(blocks < 1 ? 1 : blocks);
}

template <typename K>
int ensure_smem(K kernel, size_t smem, size_t& configured, const char* what) {
  if (smem > configured) {
    if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return cnb_check_launch(what);
    configured = smem;
  }
  return CNB_OK;
}

}  // namespace

extern "C" int cnb_weights_fwd(const float* density, const float* starts, const float* ends, int64_t row_stride, int64_t R, int32_t S,
                               float* weights, cnb_stream_t stream) {
  CNB_REQUIRE(R >= 0 && S >= 1 && S <= 8192, "weights_fwd: bad sizes R=%lld S=%d", (long long)R, S);
  if (R == 0) return CNB_OK;
  CNB_REQUIRE(density && starts && ends && weights && row_stride >= S, "weights_fwd: null pointer / bad stride");
  const size_t smem = sizeof(float) * WARPS * 2 * (size_t)S;
  static size_t configured = 48 * 1024;
  int rc = ensure_smem(k_weights_fwd, smem, configured, "weights_fwd attr");
  if (rc) return rc;
  k_weights_fwd<<<ray_grid(R), WARPS * 32, smem, stream>>>(density, starts, ends, row_stride, R, S, weights);
  return cnb_check_launch("weights_fwd");
}

extern "C" int cnb_weights_bwd(const float* density, const float* starts, const float* ends, int64_t row_stride, int64_t R, int32_t S,
                               const float* d_weights, float* d_density, cnb_stream_t stream) {
  CNB_REQUIRE(R >= 0 && S >= 1 && S <= 8192, "weights_bwd: bad sizes R=%lld S=%d", (long long)R, S);
  if (R == 0) return CNB_OK;
  CNB_REQUIRE(density && starts && ends && d_weights && d_density && row_stride >= S, "weights_bwd: null pointer / bad stride");
  const size_t smem = sizeof(float) * WARPS * 3 * (size_t)S;
  static size_t configured = 48 * 1024;
  int rc = ensure_smem(k_weights_bwd, smem, configured, "weights_bwd attr");
  if (rc) return rc;
  k_weights_bwd<<<ray_grid(R), WARPS * 32, smem, stream>>>(density, starts, ends, row_stride, R, S, d_weights, d_density);
  return cnb_check_launch("weights_bwd");
}

extern "C" int cnb_render_fwd(const float* weights, const float* rgb, const float* sem, const float* starts, const float* ends,
                              int64_t row_stride, int64_t R, int32_t S, int32_t bg_mode, const float* bg_color, int32_t eval_mode,
                              float* rgb_out, float* depth_out, float* acc_out, float* sem_out, int32_t* median_index, cnb_stream_t stream) {
  CNB_REQUIRE(R >= 0 && S >= 1 && S <= 8192, "render_fwd: bad sizes R=%lld S=%d", (long long)R, S);
  if (R == 0) return CNB_OK;
  CNB_REQUIRE(weights != nullptr, "render_fwd: null weights");
  CNB_REQUIRE(!(depth_out || median_index) || (starts && ends && row_stride >= S), "render_fwd: depth needs starts/ends");
  CNB_REQUIRE(bg_mode != CNB_BG_CONSTANT || bg_color != nullptr, "render_fwd: constant background needs bg_color (host pointer, 3 floats)");
  float b0 = 0.f, b1 = 0.f, b2 = 0.f;
  if (bg_mode == CNB_BG_CONSTANT) { b0 = bg_color[0]; b1 = bg_color[1]; b2 = bg_color[2]; }
  const size_t smem = sizeof(float) * WARPS * (size_t)S;
  static size_t configured = 48 * 1024;
  int rc = ensure_smem(k_render_fwd, smem, configured, "render_fwd attr");
  if (rc) return rc;
  k_render_fwd<<<ray_grid(R), WARPS * 32, smem, stream>>>(weights, rgb, sem, starts, ends, row_stride, R, S, bg_mode, b0, b1, b2, eval_mode, rgb_out,
                                                          depth_out, acc_out, sem_out, median_index);
  return cnb_check_launch("render_fwd");
}

extern "C" int cnb_render_bwd(const float* weights, const float* rgb, const float* sem, int64_t R, int32_t S, int32_t bg_mode,
                              const float* bg_color, const float* d_rgb_out, const float* d_acc_out, const float* d_sem_out,
                              int32_t sem_weight_grad, float* d_weights, float* d_rgb, float* d_sem, cnb_stream_t stream) {
  CNB_REQUIRE(R >= 0 && S >= 1, "render_bwd: bad sizes R=%lld S=%d", (long long)R, S);
  if (R == 0) return CNB_OK;
  CNB_REQUIRE(weights != nullptr, "render_bwd: null weights");
  CNB_REQUIRE(bg_mode != CNB_BG_CONSTANT || bg_color != nullptr, "render_bwd: constant background needs bg_color (host pointer, 3 floats)");
  float b0 = 0.f, b1 = 0.f, b2 = 0.f;
  if (bg_mode == CNB_BG_CONSTANT) { b0 = bg_color[0]; b1 = bg_color[1]; b2 = bg_color[2]; }
  k_render_bwd<<<ray_grid(R), WARPS * 32, 0, stream>>>(weights, rgb, sem, R, S, bg_mode, b0, b1, b2, d_rgb_out, d_acc_out, d_sem_out,
                                                       sem_weight_grad, d_weights, d_rgb, d_sem);
  return cnb_check_launch("render_bwd");
}
