// Fused per-ray stages of the render loop (rows a8, a10-a14, a16 of SURVEY.md section 8): the same arithmetic as the
// per-primitive kernels of sampler.cu / render.cu / losses.cu, chained inside ONE warp-per-ray kernel per pipeline
// stage so a ray's samples make one trip through shared memory instead of one kernel launch per nerfstudio call:
//   cnb_level_resample      = RaySamples.get_weights -> DepthRenderer(median) -> PDFSampler          (per proposal level)
//   cnb_final_composite     = RaySamples.get_weights -> RGB / accumulation / semantic / depth renderers (final level)
//   cnb_final_composite_bwd = MSE + BCE-with-logits gradients -> renderers' backward -> get_weights backward
//   cnb_interlevel_fused    = interlevel_loss forward (+ backward -> get_weights backward) of one proposal level
// At 4096 rays every one of the original kernels is a ~7-30 us latency-bound launch; together they were 22 % of the
// training step and 25 % of a 32768-ray render.  Results are bit-identical to the unfused kernels (tests compare).
#include "cnb_common.cuh"
#include "warp_scan.cuh"

namespace {

constexpr int WARPS = 4;
constexpr float EPS7 = 1e-7f;

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

// 4-byte asynchronous global -> shared copies: a ray's inputs are all put in flight at once (no register staging, no load->store stall of the
// in-order warp) and waited for together; the per-ray kernels are bound by the number of dependent global round trips, not by bytes
__device__ __forceinline__ void cp_async4(float* smem_dst, const float* gmem_src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// weights of one ray (k_weights_fwd's arithmetic) from STAGED inputs: dens_s[0..S) and edges_s[0..S] already in shared memory; wsm may alias dens_s
// WIDE: the caller sees the 256-sample proposal level (conflict-free 8-per-lane scan, warp_scan.cuh); the field-level kernels pass false
template <bool WIDE>
__device__ __forceinline__ void ray_weights_staged(const float* dens_s, const float* edges_s, int S, float* dd, float* cs, float* wsm, int lane) {
  for (int j = lane; j < S; j += 32) dd[j] = __fmul_rn(__fsub_rn(edges_s[j + 1], edges_s[j]), dens_s[j]);
  __syncwarp();
  if (WIDE) cnb_warp_cumsum(dd, cs, S, lane);
  else cnb_warp_cumsum_chunked(dd, cs, S, lane);
  __syncwarp();
  for (int j = lane; j < S; j += 32) {
    const float alpha = __fsub_rn(1.0f, expf(-dd[j]));
    const float T = expf(-(j > 0 ? cs[j - 1] : 0.0f));
    wsm[j] = cnb_nan_to_num(__fmul_rn(alpha, T));
  }
  __syncwarp();
}

// smem per warp: dd [Sp], cs [Sp], w [Sp], cdf [Sp+1], bins [Sp+1]
__global__ void __launch_bounds__(WARPS * 32, 8) k_level_resample(const float* __restrict__ density, const float* __restrict__ eu_prev, const float* __restrict__ sp_prev,
                                                               const float* __restrict__ nears, const float* __restrict__ fars, int kind, float anneal,
                                                               const float* __restrict__ u_base, const float* __restrict__ rand, int rand_stride, int64_t R,
                                                               int Sp, int S, float hist_pad, float eps, float* __restrict__ weights_out,
                                                               float* __restrict__ depth_out, float* __restrict__ sp_bins, float* __restrict__ eu_bins,
                                                               int32_t* __restrict__ inds_out) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* dd = smem + (size_t)warp * (5 * Sp + 2);
  float* cs = dd + Sp;
  float* wsm = cs + Sp;
  float* cdf = wsm + Sp;
  float* bins = cdf + (Sp + 1);
  const int nb = S + 1;
  const float inv_nb_half = (float)(1.0 / (2.0 * (double)nb));
  for (int64_t r = blockIdx.x * (int64_t)WARPS + warp; r < R; r += (int64_t)gridDim.x * WARPS) {
    // every global input of the ray in flight at once: density -> wsm, euclidean edges -> cdf, spacing edges -> bins (async copies), plus the
    // per-ray scalars; one wait instead of four dependent round trips
    for (int j = lane; j < Sp; j += 32) cp_async4(wsm + j, density + r * Sp + j);
    for (int j = lane; j <= Sp; j += 32) { cp_async4(cdf + j, eu_prev + r * (Sp + 1) + j); cp_async4(bins + j, sp_prev + r * (Sp + 1) + j); }
    const float near_v = __ldg(nears + r), far_v = __ldg(fars + r);
    const float rand_ray = (rand != nullptr && rand_stride == 1) ? __ldg(rand + r) : 0.0f;
    cp_async_wait_all();
    __syncwarp();
    ray_weights_staged<true>(wsm, cdf, Sp, dd, cs, wsm, lane);
    if (weights_out)
      for (int j = lane; j < Sp; j += 32) weights_out[r * Sp + j] = wsm[j];
    if (depth_out) {  // DepthRenderer "median" on the staged edges (ray_median_depth's arithmetic)
      cnb_warp_cumsum(wsm, cs, Sp, lane);
      __syncwarp();
      if (lane == 0) {
        int idx = cnb_search_left(cs, Sp, 0.5f);
        idx = min(max(idx, 0), Sp - 1);
        depth_out[r] = __fmul_rn(__fadd_rn(cdf[idx], cdf[idx + 1]), 0.5f);
      }
      __syncwarp();
    }
    // ---- PDFSampler (k_sample_pdf's arithmetic) ------------------------------------------------------------------------------
    double part = 0.0;
    for (int j = lane; j < Sp; j += 32) {
      float w = wsm[j];
      if (anneal == 0.0f) w = 1.0f;
      else if (anneal != 1.0f) w = powf(w, anneal);
      w = __fadd_rn(w, hist_pad);
      cdf[1 + j] = w;
      part += (double)w;
    }
    float wsum = (float)cnb_warp_sum_d(part);
    const float padding = fmaxf(__fsub_rn(eps, wsum), 0.0f);
    const float padj = __fdiv_rn(padding, (float)Sp);
    wsum = __fadd_rn(wsum, padding);
    __syncwarp();
    for (int j = lane; j < Sp; j += 32) cdf[1 + j] = __fdiv_rn(__fadd_rn(cdf[1 + j], padj), wsum);
    __syncwarp();
    cnb_warp_cumsum(cdf + 1, cdf + 1, Sp, lane);
    __syncwarp();
    for (int j = lane; j < Sp; j += 32) cdf[1 + j] = fminf(1.0f, cdf[1 + j]);
    if (lane == 0) cdf[0] = 0.0f;
    __syncwarp();
    const float s_near = cnb_spacing_fn(kind, near_v);
    const float s_far = cnb_spacing_fn(kind, far_v);
    for (int k = lane; k < nb; k += 32) {
      float u = __ldg(u_base + k);
      if (rand != nullptr) u = __fadd_rn(u, __fdiv_rn(rand_stride == 1 ? rand_ray : __ldg(rand + r * rand_stride + k), (float)nb));
      else u = __fadd_rn(u, inv_nb_half);
      const int ind = cnb_search_right(cdf, Sp + 1, u);
      const int below = min(max(ind - 1, 0), Sp);
      const int above = min(max(ind, 0), Sp);
      const float c0 = cdf[below], c1 = cdf[above];
      float t = __fdiv_rn(__fsub_rn(u, c0), __fsub_rn(c1, c0));
      if (isnan(t)) t = 0.0f;
      t = fminf(fmaxf(t, 0.0f), 1.0f);
      const float b0 = bins[below], b1 = bins[above];
      const float nbv = __fadd_rn(b0, __fmul_rn(t, __fsub_rn(b1, b0)));
      sp_bins[r * nb + k] = nbv;
      eu_bins[r * nb + k] = cnb_spacing_to_euclid(kind, nbv, s_near, s_far);
      if (inds_out) inds_out[r * nb + k] = ind;
    }
    __syncwarp();
  }
}

// smem per warp: dd [S], cs [S], w [S], edges [S+1], rgb [3S], sem [S]
__global__ void __launch_bounds__(WARPS * 32) k_final_composite(const float* __restrict__ density, const float* __restrict__ rgb, const float* __restrict__ sem,
                                                                const float* __restrict__ eu, int64_t R, int S, int bg_mode, float bg0, float bg1, float bg2,
                                                                int eval_mode, float* __restrict__ weights_out, float* __restrict__ rgb_out,
                                                                float* __restrict__ depth_out, float* __restrict__ acc_out, float* __restrict__ sem_out) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* dd = smem + (size_t)warp * (8 * S + 1);
  float* cs = dd + S;
  float* wsm = cs + S;
  float* ed = wsm + S;        // [S+1]
  float* rg = ed + (S + 1);   // [3S]
  float* se = rg + 3 * S;     // [S]
  for (int64_t r = blockIdx.x * (int64_t)WARPS + warp; r < R; r += (int64_t)gridDim.x * WARPS) {
    // all global inputs of the ray in flight at once (see k_level_resample)
    for (int j = lane; j < S; j += 32) { cp_async4(wsm + j, density + r * S + j); cp_async4(se + j, sem + r * S + j); }
    for (int j = lane; j <= S; j += 32) cp_async4(ed + j, eu + r * (S + 1) + j);
    for (int j = lane; j < 3 * S; j += 32) cp_async4(rg + j, rgb + r * S * 3 + j);
    cp_async_wait_all();
    __syncwarp();
    ray_weights_staged<false>(wsm, ed, S, dd, cs, wsm, lane);
    float a = 0.f, c0 = 0.f, c1 = 0.f, c2 = 0.f, sm = 0.f;
    for (int j = lane; j < S; j += 32) {  // k_render_fwd's arithmetic
      const float w = wsm[j];
      if (weights_out) weights_out[r * S + j] = w;
      a += w;
      float x = rg[3 * j], y = rg[3 * j + 1], z = rg[3 * j + 2];
      if (eval_mode) { x = cnb_nan_to_num(x); y = cnb_nan_to_num(y); z = cnb_nan_to_num(z); }
      c0 = fmaf(w, x, c0); c1 = fmaf(w, y, c1); c2 = fmaf(w, z, c2);
      sm = fmaf(w, se[j], sm);
    }
    a = cnb_warp_sum(a);
    if (acc_out && lane == 0) acc_out[r] = a;
    sm = cnb_warp_sum(sm);
    if (sem_out && lane == 0) sem_out[r] = sm;
    c0 = cnb_warp_sum(c0); c1 = cnb_warp_sum(c1); c2 = cnb_warp_sum(c2);
    if (rgb_out && lane == 0) {
      float b0 = bg0, b1 = bg1, b2 = bg2;
      if (bg_mode == CNB_BG_LAST_SAMPLE) {
        b0 = rg[3 * (S - 1)]; b1 = rg[3 * (S - 1) + 1]; b2 = rg[3 * (S - 1) + 2];
        if (eval_mode) { b0 = cnb_nan_to_num(b0); b1 = cnb_nan_to_num(b1); b2 = cnb_nan_to_num(b2); }
      }
      if (bg_mode != CNB_BG_NONE) {
        const float rem = 1.0f - a;
        c0 += b0 * rem; c1 += b1 * rem; c2 += b2 * rem;
      }
      if (eval_mode) { c0 = fminf(fmaxf(c0, 0.f), 1.f); c1 = fminf(fmaxf(c1, 0.f), 1.f); c2 = fminf(fmaxf(c2, 0.f), 1.f); }
      rgb_out[3 * r] = c0; rgb_out[3 * r + 1] = c1; rgb_out[3 * r + 2] = c2;
    }
    if (depth_out) {  // DepthRenderer "median" on the staged edges (ray_median_depth's arithmetic)
      __syncwarp();
      cnb_warp_cumsum_chunked(wsm, cs, S, lane);
      __syncwarp();
      if (lane == 0) {
        int idx = cnb_search_left(cs, S, 0.5f);
        idx = min(max(idx, 0), S - 1);
        depth_out[r] = __fmul_rn(__fadd_rn(ed[idx], ed[idx + 1]), 0.5f);
      }
    }
    __syncwarp();
  }
}

// pixel-loss gradients -> renderers' backward -> get_weights backward, one warp per ray.
// smem per warp: dd [S], cs [S], gw [S], gd [S]
__global__ void __launch_bounds__(WARPS * 32) k_final_composite_bwd(const float* __restrict__ density, const float* __restrict__ rgb, const float* __restrict__ sem,
                                                                    const float* __restrict__ eu, const float* __restrict__ weights,
                                                                    const float* __restrict__ rgb_out, const float* __restrict__ sem_out,
                                                                    const float* __restrict__ image, const float* __restrict__ mask, int64_t R, int S,
                                                                    int bg_mode, float bg0, float bg1, float bg2, float sem_weight, float grad_scale,
                                                                    int sem_weight_grad, float* __restrict__ losses_out, float* __restrict__ d_density,
                                                                    float* __restrict__ d_rgb, float* __restrict__ d_sem) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* dd = smem + (size_t)warp * 4 * S;
  float* cs = dd + S;
  float* gw = cs + S;
  float* gd = gw + S;
  const float inv3r = 1.0f / (3.0f * (float)R), invr = 1.0f / (float)R;
  float mse = 0.0f, bce = 0.0f;
  for (int64_t r = blockIdx.x * (int64_t)WARPS + warp; r < R; r += (int64_t)gridDim.x * WARPS) {
    // ---- k_pixel_losses ---------------------------------------------------------------------------------------------------------
    float g0, g1, g2, gs;
    {
      const float d0 = __ldg(rgb_out + 3 * r) - __ldg(image + 3 * r), d1 = __ldg(rgb_out + 3 * r + 1) - __ldg(image + 3 * r + 1),
                  d2 = __ldg(rgb_out + 3 * r + 2) - __ldg(image + 3 * r + 2);
      g0 = 2.0f * d0 * inv3r * grad_scale; g1 = 2.0f * d1 * inv3r * grad_scale; g2 = 2.0f * d2 * inv3r * grad_scale;
      const float x = __ldg(sem_out + r), y = __ldg(mask + r);
      gs = (1.0f / (1.0f + expf(-x)) - y) * invr * sem_weight * grad_scale;
      if (lane == 0) {
        mse = fmaf(d0, d0, mse); mse = fmaf(d1, d1, mse); mse = fmaf(d2, d2, mse);
        bce += fmaxf(x, 0.0f) - x * y + log1pf(expf(-fabsf(x)));
      }
    }
    // ---- k_render_bwd -----------------------------------------------------------------------------------------------------------
    float b0 = bg0, b1 = bg1, b2 = bg2, rem = 0.0f;
    if (bg_mode == CNB_BG_LAST_SAMPLE) {
      const float* last = rgb + (r * S + S - 1) * 3;
      b0 = __ldg(last); b1 = __ldg(last + 1); b2 = __ldg(last + 2);
      float a = 0.f;
      for (int j = lane; j < S; j += 32) a += __ldg(weights + r * S + j);
      rem = 1.0f - cnb_warp_sum(a);
    }
    if (bg_mode == CNB_BG_NONE) { b0 = b1 = b2 = 0.0f; }
    const float* e = eu + r * (S + 1);
    for (int j = lane; j < S; j += 32) {
      const int64_t i = r * S + j;
      const float w = __ldg(weights + i);
      const float x = __ldg(rgb + 3 * i), y = __ldg(rgb + 3 * i + 1), z = __ldg(rgb + 3 * i + 2);
      float gwj = g0 * (x - b0) + g1 * (y - b1) + g2 * (z - b2);
      const float ex = (bg_mode == CNB_BG_LAST_SAMPLE && j == S - 1) ? rem : 0.0f;
      d_rgb[3 * i] = g0 * (w + ex); d_rgb[3 * i + 1] = g1 * (w + ex); d_rgb[3 * i + 2] = g2 * (w + ex);
      if (sem_weight_grad) gwj += gs * __ldg(sem + i);
      d_sem[i] = gs * w;
      gd[j] = gwj;  // d_weights
      const float delta = __fsub_rn(__ldg(e + j + 1), __ldg(e + j));
      dd[j] = __fmul_rn(delta, __ldg(density + i));
    }
    __syncwarp();
    // ---- k_weights_bwd ------------------------------------------------------------------------------------------------------------
    cnb_warp_cumsum_chunked(dd, cs, S, lane);
    __syncwarp();
    for (int j = lane; j < S; j += 32) {
      const float T = expf(-(j > 0 ? cs[j - 1] : 0.0f));
      const float w = __fmul_rn(__fsub_rn(1.0f, expf(-dd[j])), T);
      gw[j] = isfinite(w) ? gd[j] * w : 0.0f;
    }
    __syncwarp();
    cnb_warp_suffix_excl_chunked(gw, gw, S, lane);
    __syncwarp();
    for (int j = lane; j < S; j += 32) {
      const float T = expf(-(j > 0 ? cs[j - 1] : 0.0f));
      const float ev = expf(-dd[j]);
      const float w = __fmul_rn(__fsub_rn(1.0f, ev), T);
      const float g = isfinite(w) ? gd[j] : 0.0f;
      const float delta = __fsub_rn(__ldg(e + j + 1), __ldg(e + j));
      d_density[r * S + j] = delta * (g * T * ev - gw[j]);
    }
    __syncwarp();
  }
  if (lane == 0 && losses_out != nullptr) {
    if (mse != 0.0f) atomicAdd(losses_out, mse * inv3r);
    if (bce != 0.0f) atomicAdd(losses_out + 1, bce * invr * sem_weight);
  }
}

// interlevel loss of one proposal level, forward (+ backward through get_weights when d_density != nullptr).
// smem per warp: cps [Sp+1], cy1 [Sp+1], dcy [Sp+1], dd [Sp], cs [Sp]
__global__ void __launch_bounds__(WARPS * 32, 8) k_interlevel_fused(const float* __restrict__ c, const float* __restrict__ w, const float* __restrict__ cp,
                                                                 const float* __restrict__ wp, const float* __restrict__ density_p,
                                                                 const float* __restrict__ eu_p, int64_t R, int Sc, int Sp, float grad_scale,
                                                                 float* __restrict__ loss_out, float* __restrict__ d_density_p) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* cps = smem + (size_t)warp * (5 * Sp + 3);
  float* cy1 = cps + (Sp + 1);
  float* dcy = cy1 + (Sp + 1);
  float* dd = dcy + (Sp + 1);
  float* cs = dd + Sp;
  const float norm = 1.0f / (float)((double)R * (double)Sc);
  const bool bwd = d_density_p != nullptr;
  float loss_acc = 0.0f;
  for (int64_t r = blockIdx.x * (int64_t)WARPS + warp; r < R; r += (int64_t)gridDim.x * WARPS) {
    for (int j = lane; j <= Sp; j += 32) { cps[j] = __ldg(cp + r * (Sp + 1) + j); dcy[j] = 0.0f; }
    for (int j = lane; j < Sp; j += 32) cy1[1 + j] = __ldg(wp + r * Sp + j);
    if (lane == 0) cy1[0] = 0.0f;
    __syncwarp();
    cnb_warp_cumsum(cy1 + 1, cy1 + 1, Sp, lane);
    __syncwarp();
    for (int i = lane; i < Sc; i += 32) {
      const float t0s = __ldg(c + r * (Sc + 1) + i), t0e = __ldg(c + r * (Sc + 1) + i + 1);
      int lo = cnb_search_right(cps, Sp, t0s) - 1;
      lo = min(max(lo, 0), Sp - 1);
      int hi = cnb_search_right(cps + 1, Sp, t0e);
      hi = min(max(hi, 0), Sp - 1);
      const float w_outer = __fsub_rn(cy1[hi + 1], cy1[lo]);
      const float wi = __ldg(w + r * Sc + i);
      const float diff = fmaxf(__fsub_rn(wi, w_outer), 0.0f);
      const float denom = __fadd_rn(wi, EPS7);
      loss_acc += diff * diff / denom;
      if (bwd && diff > 0.0f) {
        const float g = -2.0f * diff / denom * norm * grad_scale;
        atomicAdd(dcy + hi + 1, g);
        atomicAdd(dcy + lo, -g);
      }
    }
    if (bwd) {
      __syncwarp();
      cnb_warp_suffix_excl(dcy, dcy, Sp + 1, lane);  // dcy[k] = d loss / d wp_k  (k < Sp)
      // ---- get_weights backward of the proposal level (k_weights_bwd's arithmetic) -----------------------------------------------
      const float* e = eu_p + r * (Sp + 1);
      for (int j = lane; j < Sp; j += 32) {
        const float delta = __fsub_rn(__ldg(e + j + 1), __ldg(e + j));
        dd[j] = __fmul_rn(delta, __ldg(density_p + r * Sp + j));
      }
      __syncwarp();
      cnb_warp_cumsum(dd, cs, Sp, lane);
      __syncwarp();
      for (int j = lane; j < Sp; j += 32) {
        const float T = expf(-(j > 0 ? cs[j - 1] : 0.0f));
        const float wv = __fmul_rn(__fsub_rn(1.0f, expf(-dd[j])), T);
        cy1[j] = isfinite(wv) ? dcy[j] * wv : 0.0f;  // cy1 reused as the g*w array
      }
      __syncwarp();
      cnb_warp_suffix_excl(cy1, cy1, Sp, lane);
      __syncwarp();
      for (int j = lane; j < Sp; j += 32) {
        const float T = expf(-(j > 0 ? cs[j - 1] : 0.0f));
        const float ev = expf(-dd[j]);
        const float wv = __fmul_rn(__fsub_rn(1.0f, ev), T);
        const float g = isfinite(wv) ? dcy[j] : 0.0f;
        const float delta = __fsub_rn(__ldg(e + j + 1), __ldg(e + j));
        d_density_p[r * Sp + j] = delta * (g * T * ev - cy1[j]);
      }
    }
    __syncwarp();
  }
  if (loss_out != nullptr) {
    loss_acc = cnb_warp_sum(loss_acc);
    if (lane == 0 && loss_acc != 0.0f) atomicAdd(loss_out, loss_acc * norm);
  }
}

}  // namespace

extern "C" int cnb_level_resample(const float* density, const float* euclid_bins_prev, const float* spacing_bins_prev, const float* nears,
                                  const float* fars, int32_t spacing, float anneal, const float* u_base, const float* rand, int32_t rand_stride,
                                  int64_t R, int32_t Sp, int32_t S, float histogram_padding, float eps, float* weights_out, float* depth_out,
                                  float* spacing_bins, float* euclid_bins, int32_t* inds, cnb_stream_t stream) {
  CNB_REQUIRE(R >= 0 && S >= 1 && Sp >= 1 && Sp <= 2048, "level_resample: bad sizes R=%lld Sp=%d S=%d", (long long)R, Sp, S);
  if (R == 0) return CNB_OK;
  CNB_REQUIRE(density && euclid_bins_prev && spacing_bins_prev && nears && fars && u_base && spacing_bins && euclid_bins, "level_resample: null pointer");
  CNB_REQUIRE(rand == nullptr || rand_stride == 1 || rand_stride == S + 1, "level_resample: rand_stride must be 1 or S+1");
  const size_t smem = sizeof(float) * WARPS * (size_t)(5 * Sp + 2);
  static size_t configured = 48 * 1024;
  int rc = ensure_smem(k_level_resample, smem, configured, "level_resample attr");
  if (rc) return rc;
  k_level_resample<<<ray_grid(R), WARPS * 32, smem, stream>>>(density, euclid_bins_prev, spacing_bins_prev, nears, fars, spacing, anneal, u_base, rand,
                                                              rand_stride, R, Sp, S, histogram_padding, eps, weights_out, depth_out, spacing_bins,
                                                              euclid_bins, inds);
  return cnb_check_launch("level_resample");
}

extern "C" int cnb_final_composite(const float* density, const float* rgb, const float* sem, const float* euclid_bins, int64_t R, int32_t S,
                                   int32_t bg_mode, const float* bg_color, int32_t eval_mode, float* weights_out, float* rgb_out, float* depth_out,
                                   float* acc_out, float* sem_out, cnb_stream_t stream) {
  CNB_REQUIRE(R >= 0 && S >= 1 && S <= 4096, "final_composite: bad sizes R=%lld S=%d", (long long)R, S);
  if (R == 0) return CNB_OK;
  CNB_REQUIRE(density && rgb && sem && euclid_bins, "final_composite: null pointer");
  CNB_REQUIRE(bg_mode != CNB_BG_CONSTANT || bg_color != nullptr, "final_composite: constant background needs bg_color (host pointer, 3 floats)");
  float b0 = 0.f, b1 = 0.f, b2 = 0.f;
  if (bg_mode == CNB_BG_CONSTANT) { b0 = bg_color[0]; b1 = bg_color[1]; b2 = bg_color[2]; }
  const size_t smem = sizeof(float) * WARPS * (8 * (size_t)S + 1);
  static size_t configured = 48 * 1024;
  int rc = ensure_smem(k_final_composite, smem, configured, "final_composite attr");
  if (rc) return rc;
  k_final_composite<<<ray_grid(R), WARPS * 32, smem, stream>>>(density, rgb, sem, euclid_bins, R, S, bg_mode, b0, b1, b2, eval_mode, weights_out, rgb_out,
                                                               depth_out, acc_out, sem_out);
  return cnb_check_launch("final_composite");
}

extern "C" int cnb_final_composite_bwd(const float* density, const float* rgb, const float* sem, const float* euclid_bins, const float* weights,
                                       const float* rgb_out, const float* sem_out, const float* image, const float* mask, int64_t R, int32_t S,
                                       int32_t bg_mode, const float* bg_color, float sem_weight, float grad_scale, int32_t sem_weight_grad,
                                       float* losses_out, float* d_density, float* d_rgb, float* d_sem, cnb_stream_t stream) {
  CNB_REQUIRE(R >= 0 && S >= 1 && S <= 4096, "final_composite_bwd: bad sizes R=%lld S=%d", (long long)R, S);
  if (R == 0) return CNB_OK;
  CNB_REQUIRE(density && rgb && sem && euclid_bins && weights && rgb_out && sem_out && image && mask && d_density && d_rgb && d_sem,
              "final_composite_bwd: null pointer");
  CNB_REQUIRE(bg_mode != CNB_BG_CONSTANT || bg_color != nullptr, "final_composite_bwd: constant background needs bg_color (host pointer, 3 floats)");
  float b0 = 0.f, b1 = 0.f, b2 = 0.f;
  if (bg_mode == CNB_BG_CONSTANT) { b0 = bg_color[0]; b1 = bg_color[1]; b2 = bg_color[2]; }
  const size_t smem = sizeof(float) * WARPS * 4 * (size_t)S;
  static size_t configured = 48 * 1024;
  int rc = ensure_smem(k_final_composite_bwd, smem, configured, "final_composite_bwd attr");
  if (rc) return rc;
  k_final_composite_bwd<<<ray_grid(R), WARPS * 32, smem, stream>>>(density, rgb, sem, euclid_bins, weights, rgb_out, sem_out, image, mask, R, S, bg_mode, b0,
                                                                   b1, b2, sem_weight, grad_scale, sem_weight_grad, losses_out, d_density, d_rgb, d_sem);
  return cnb_check_launch("final_composite_bwd");
}

extern "C" int cnb_interlevel_fused(const float* c, const float* w, const float* cp, const float* wp, const float* density_p, const float* euclid_bins_p,
                                    int64_t R, int32_t Sc, int32_t Sp, float grad_scale, float* loss_out, float* d_density_p, cnb_stream_t stream) {
  CNB_REQUIRE(R >= 0 && Sc >= 1 && Sp >= 1 && Sp <= 2048, "interlevel_fused: bad sizes R=%lld Sc=%d Sp=%d", (long long)R, Sc, Sp);
  if (R == 0) return CNB_OK;
  CNB_REQUIRE(c && w && cp && wp, "interlevel_fused: null pointer");
  CNB_REQUIRE(d_density_p == nullptr || (density_p && euclid_bins_p), "interlevel_fused: backward needs the proposal density and bin edges");
  const size_t smem = sizeof(float) * WARPS * (size_t)(5 * Sp + 3);
  static size_t configured = 48 * 1024;
  int rc = ensure_smem(k_interlevel_fused, smem, configured, "interlevel_fused attr");
  if (rc) return rc;
  k_interlevel_fused<<<ray_grid(R), WARPS * 32, smem, stream>>>(c, w, cp, wp, density_p, euclid_bins_p, R, Sc, Sp, grad_scale, loss_out, d_density_p);
  return cnb_check_launch("interlevel_fused");
}
