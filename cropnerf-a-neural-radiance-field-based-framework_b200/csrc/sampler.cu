// Ray samplers (rows a8/a9 of SURVEY.md section 8).
//  - cnb_sample_spaced : nerfstudio ray_samplers.py SpacedSampler.generate_ray_samples, i.e. the routine the
//    reference copies at components/ray_samplers.py:54-104 (UniformSamplerWithNoise) and the
//    UniformLinDispPiecewiseSampler that ProposalNetworkSampler starts from (fruit_nerf.py:157-164).
//  - cnb_sample_pdf    : nerfstudio PDFSampler.generate_ray_samples (include_original=False) with the
//    `weights ** anneal` of ProposalNetworkSampler.generate_ray_samples folded in.
// The PDF resampler is one warp per ray: annealed weights and the previous bin edges are staged in shared
// memory, the cdf is a warp prefix scan (double partials, see warp_scan.cuh), and each lane inverts the cdf for
// its bin edges with a binary search (torch.searchsorted side="right").
#include "cnb_common.cuh"
#include "warp_scan.cuh"

namespace {

__global__ void __launch_bounds__(256) k_sample_spaced(const float* __restrict__ nears, const float* __restrict__ fars,
                                                        const float* __restrict__ lin_bins, const float* __restrict__ t_rand, int rand_stride,
                                                        int kind, int64_t R, int S, float* __restrict__ sp_bins, float* __restrict__ eu_bins,
                                                        float near_plane, float far_plane, float* __restrict__ nears_out, float* __restrict__ fars_out) {
  // nears / fars may be NULL: the collider's planes apply (NearFarCollider); nears_out / fars_out (optional) receive the per-ray values
  const int nb = S + 1;
  const int64_t total = R * nb;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = cnb_ray_of(i, nb);
    const int j = (int)(i - r * nb);
    float b = __ldg(lin_bins + j);
    if (t_rand != nullptr) {
      // bin_centers = (bins[1:] + bins[:-1]) / 2 ; upper = cat(centers, bins[-1]) ; lower = cat(bins[0], centers)
      const float upper = j < S ? __fmul_rn(__fadd_rn(__ldg(lin_bins + j + 1), b), 0.5f) : b;
      const float lower = j > 0 ? __fmul_rn(__fadd_rn(b, __ldg(lin_bins + j - 1)), 0.5f) : b;
      const float t = __ldg(t_rand + r * rand_stride + (rand_stride == 1 ? 0 : j));
      b = __fadd_rn(lower, __fmul_rn(__fsub_rn(upper, lower), t));
    }
    const float near_v = nears != nullptr ? __ldg(nears + r) : near_plane;
    const float far_v = fars != nullptr ? __ldg(fars + r) : far_plane;
    if (j == 0 && nears_out != nullptr) { nears_out[r] = near_v; fars_out[r] = far_v; }
    const float s_near = cnb_spacing_fn(kind, near_v);
    const float s_far = cnb_spacing_fn(kind, far_v);
    sp_bins[i] = b;
    eu_bins[i] = cnb_spacing_to_euclid(kind, b, s_near, s_far);
  }
}

constexpr int PDF_WARPS = 4;

__global__ void __launch_bounds__(PDF_WARPS * 32) k_sample_pdf(const float* __restrict__ weights, float anneal, const float* __restrict__ prev_bins,
                                                               const float* __restrict__ nears, const float* __restrict__ fars, int kind,
                                                               const float* __restrict__ u_base, const float* __restrict__ rand, int rand_stride,
                                                               int64_t R, int Sp, int S, float hist_pad, float eps, float* __restrict__ sp_bins,
                                                               float* __restrict__ eu_bins, int32_t* __restrict__ inds_out) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* cdf = smem + (size_t)warp * 2 * (Sp + 1);  // [Sp+1]
  float* bins = cdf + (Sp + 1);                     // [Sp+1]
  const int nb = S + 1;
  const float inv_nb_half = (float)(1.0 / (2.0 * (double)nb));
  for (int64_t r = blockIdx.x * (int64_t)PDF_WARPS + warp; r < R; r += (int64_t)gridDim.x * PDF_WARPS) {
    // weights = weights**anneal + histogram_padding
    double part = 0.0;
    for (int j = lane; j < Sp; j += 32) {
      float w = __ldg(weights + r * Sp + j);
      if (anneal == 0.0f) w = 1.0f;
      else if (anneal != 1.0f) w = powf(w, anneal);
      w = __fadd_rn(w, hist_pad);
      cdf[1 + j] = w;
      part += (double)w;
    }
    for (int j = lane; j <= Sp; j += 32) bins[j] = __ldg(prev_bins + r * (Sp + 1) + j);
    float wsum = (float)cnb_warp_sum_d(part);
    const float padding = fmaxf(__fsub_rn(eps, wsum), 0.0f);
    const float padj = __fdiv_rn(padding, (float)Sp);
    wsum = __fadd_rn(wsum, padding);
    __syncwarp();
    for (int j = lane; j < Sp; j += 32) cdf[1 + j] = __fdiv_rn(__fadd_rn(cdf[1 + j], padj), wsum);  // pdf
    __syncwarp();
    cnb_warp_cumsum(cdf + 1, cdf + 1, Sp, lane);
    __syncwarp();
    for (int j = lane; j < Sp; j += 32) cdf[1 + j] = fminf(1.0f, cdf[1 + j]);
    if (lane == 0) cdf[0] = 0.0f;
    __syncwarp();
    const float s_near = cnb_spacing_fn(kind, __ldg(nears + r));
    const float s_far = cnb_spacing_fn(kind, __ldg(fars + r));
    for (int k = lane; k < nb; k += 32) {
      float u = __ldg(u_base + k);
      if (rand != nullptr) u = __fadd_rn(u, __fdiv_rn(__ldg(rand + r * rand_stride + (rand_stride == 1 ? 0 : k)), (float)nb));
      else u = __fadd_rn(u, inv_nb_half);
      const int ind = cnb_search_right(cdf, Sp + 1, u);
      const int below = min(max(ind - 1, 0), Sp);
      const int above = min(max(ind, 0), Sp);
      const float c0 = cdf[below], c1 = cdf[above];
      float t = __fdiv_rn(__fsub_rn(u, c0), __fsub_rn(c1, c0));
      if (isnan(t)) t = 0.0f;
      t = fminf(fmaxf(t, 0.0f), 1.0f);  // +-inf clip to the ends like nan_to_num + clip
      const float b0 = bins[below], b1 = bins[above];
      const float nbv = __fadd_rn(b0, __fmul_rn(t, __fsub_rn(b1, b0)));
      sp_bins[r * nb + k] = nbv;
      eu_bins[r * nb + k] = cnb_spacing_to_euclid(kind, nbv, s_near, s_far);
      if (inds_out) inds_out[r * nb + k] = ind;
    }
    __syncwarp();
  }
}

}  // namespace

extern "C" int cnb_sample_spaced(const float* nears, const float* fars, const float* lin_bins, const float* t_rand, int32_t rand_stride,
                                 int32_t spacing, int64_t R, int32_t S, float* spacing_bins, float* euclid_bins, cnb_stream_t stream) {
  CNB_REQUIRE(R >= 0 && S >= 1, "sample_spaced: bad sizes R=%lld S=%d", (long long)R, S);
  if (R == 0) return CNB_OK;
  CNB_REQUIRE(nears && fars && lin_bins && spacing_bins && euclid_bins, "sample_spaced: null pointer");
  CNB_REQUIRE(t_rand == nullptr || rand_stride == 1 || rand_stride == S + 1, "sample_spaced: rand_stride must be 1 or S+1");
  CNB_REQUIRE(spacing == CNB_SPACING_UNIFORM || spacing == CNB_SPACING_LINDISP_PIECEWISE, "sample_spaced: unknown spacing %d", spacing);
  int64_t blocks = (R * (S + 1) + 255) / 256;
  const int64_t cap = (int64_t)cnb_num_sms() * 8;
  if (blocks > cap) blocks = cap;
  k_sample_spaced<<<(int)blocks, 256, 0, stream>>>(nears, fars, lin_bins, t_rand, rand_stride, spacing, R, S, spacing_bins, euclid_bins, 0.0f, 0.0f, nullptr,
                                                   nullptr);
  return cnb_check_launch("sample_spaced");
}

// the collider folded in (what cnb_render_rays / cnb_train_step start with): per-ray nears / fars when given, else the planes; the values used
// are also written to nears_out / fars_out for the PDF resampling levels
extern "C" int cnb_sample_spaced_collide(const float* ray_nears, const float* ray_fars, float near_plane, float far_plane, const float* lin_bins,
                                         const float* t_rand, int32_t rand_stride, int32_t spacing, int64_t R, int32_t S, float* nears_out, float* fars_out,
                                         float* spacing_bins, float* euclid_bins, cnb_stream_t stream) {
  CNB_REQUIRE(R >= 0 && S >= 1, "sample_spaced_collide: bad sizes R=%lld S=%d", (long long)R, S);
  if (R == 0) return CNB_OK;
  CNB_REQUIRE(lin_bins && spacing_bins && euclid_bins && nears_out && fars_out, "sample_spaced_collide: null pointer");
  CNB_REQUIRE((ray_nears == nullptr) == (ray_fars == nullptr), "sample_spaced_collide: nears and fars come together");
  CNB_REQUIRE(t_rand == nullptr || rand_stride == 1 || rand_stride == S + 1, "sample_spaced_collide: rand_stride must be 1 or S+1");
  CNB_REQUIRE(spacing == CNB_SPACING_UNIFORM || spacing == CNB_SPACING_LINDISP_PIECEWISE, "sample_spaced_collide: unknown spacing %d", spacing);
  int64_t blocks = (R * (S + 1) + 255) / 256;
  const int64_t cap = (int64_t)cnb_num_sms() * 8;
  if (blocks > cap) blocks = cap;
  k_sample_spaced<<<(int)blocks, 256, 0, stream>>>(ray_nears, ray_fars, lin_bins, t_rand, rand_stride, spacing, R, S, spacing_bins, euclid_bins, near_plane,
                                                   far_plane, nears_out, fars_out);
  return cnb_check_launch("sample_spaced_collide");
}

extern "C" int cnb_sample_pdf(const float* weights, float anneal, const float* prev_spacing_bins, const float* nears, const float* fars,
                              int32_t spacing, const float* u_base, const float* rand, int32_t rand_stride, int64_t R, int32_t Sp, int32_t S,
                              float histogram_padding, float eps, float* spacing_bins, float* euclid_bins, int32_t* inds, cnb_stream_t stream) {
  CNB_REQUIRE(R >= 0 && S >= 1 && Sp >= 1 && Sp <= 4096, "sample_pdf: bad sizes R=%lld Sp=%d S=%d", (long long)R, Sp, S);
  if (R == 0) return CNB_OK;
  CNB_REQUIRE(weights && prev_spacing_bins && nears && fars && u_base && spacing_bins && euclid_bins, "sample_pdf: null pointer");
  CNB_REQUIRE(rand == nullptr || rand_stride == 1 || rand_stride == S + 1, "sample_pdf: rand_stride must be 1 or S+1");
  const size_t smem = sizeof(float) * PDF_WARPS * 2 * (size_t)(Sp + 1);
  static size_t configured = 48 * 1024;
  if (smem > configured) {
    if (cudaFuncSetAttribute(k_sample_pdf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return cnb_check_launch("sample_pdf attr");
    configured = smem;
  }
  int64_t blocks = (R + PDF_WARPS - 1) / PDF_WARPS;
  const int64_t cap = (int64_t)cnb_num_sms() * 16;
  if (blocks > cap) blocks = cap;
  k_sample_pdf<<<(int)blocks, PDF_WARPS * 32, smem, stream>>>(weights, anneal, prev_spacing_bins, nears, fars, spacing, u_base, rand, rand_stride, R,
                                                              Sp, S, histogram_padding, eps, spacing_bins, euclid_bins, inds);
  return cnb_check_launch("sample_pdf");
}
