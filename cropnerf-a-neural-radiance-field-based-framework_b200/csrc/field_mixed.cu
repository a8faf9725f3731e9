// FruitField on tensor cores, mixed precision (rows a1/a5/a6 of SURVEY.md section 8; the hot-path version of
// fruit_field.py:169-302): hash-grid gather -> base MLP -> trunc_exp density -> semantic MLP + head -> SH | geo |
// appearance -> RGB MLP, ONE kernel, activations never leave registers.
//
// Mapping.  A warp owns an m-tile of 16 samples; lane (g = lane/4, t = lane%4) is wired exactly like the
// mma.sync.m16n8k16 fragments:
//   * the multiresolution gather lands directly in the A-operand registers of the first GEMM: feature pair of level
//     8*kt + t (+4) of sample g (+8) IS fragment register a0..a3 of k-tile kt, so the 16 levels x 16 samples of a tile
//     are spread over the 32 lanes with no shuffle and no shared-memory staging (64 independent 8-byte gathers/lane);
//   * each layer's fp32 accumulator fragment (C layout) is bias-added, ReLU'd, packed to fp16 and re-used as the next
//     layer's A fragment (two adjacent n-tiles of C == one k-tile of A);
//   * weights live in shared memory as fp16 [out][in] rows padded by 8 halves (conflict-free 32-bit B-fragment loads).
// Column bookkeeping: the 16 base-MLP outputs [dba | geo15] are fed unchanged as one k-tile to the semantic MLP and to
// the RGB MLP (whose input becomes [SH16 | dba,geo15 | emb32] = 64 = 4 k-tiles) with the weight column of the dba
// slot zeroed and the slot itself cleared, which is arithmetically the reference's concat([d, geo, emb]).
//
// Why mma.sync and not tcgen05 here: the dense work is ~8 K FLOP per sample in 7 dependent 16x64x64-sized steps; it
// needs ~5 % of the tensor peak at the gather roofline.  Register-resident fragment chaining
// has no smem/TMEM round trip between layers, whereas tcgen05 would need TMEM->reg->smem->fence per layer for the
// activation.  DESIGN.md ("tensor-core path") records the measurement this choice rests on.
#include "field_mixed.cuh"

using namespace cnbmix;

namespace {

constexpr int TILE = WARPS * 16;   // samples per CTA tile
constexpr int XS = 40;             // row stride (halves) of the staged feature tile: conflict-free A-fragment reads
constexpr size_t SMEM_FWD2 = SMEM_FWD + (size_t)TILE * XS * sizeof(__half);

// gather + trilinear blend of one (sample, level) -> packed fp16 feature pair, in two halves so that the loads of TWO levels can be issued
// before the first blend consumes one.  When floor(x) is even the two x-neighbours of each corner pair are adjacent table rows (hash prime
// of x is 1): one 16-byte load instead of two 8-byte loads.
struct LevelLoads { CnbCell c; float2 v[8]; };

__device__ __forceinline__ void gather_issue(const float* __restrict__ table, uint32_t mask, uint32_t level_offset, float scale, float x, float y, float z,
                                             LevelLoads& o) {
  o.c = cnb_cell(x, y, z, scale);
  uint32_t h[8];
  cnb_corner_rows(o.c, mask, level_offset, h);
  cnb_gather8(table, o.c, h, o.v);
}

__device__ __forceinline__ uint32_t gather_blend(const LevelLoads& o) {
  const CnbCell& c = o.c;
  const float2 (&v)[8] = o.v;
  const float mx = 1.f - c.ox, my = 1.f - c.oy, mz = 1.f - c.oz;
  // same pairing as the reference blend (f03,f12,f56,f47 -> f0312,f4756), FMA-contracted
  float2 f03, f12, f56, f47;
  f03.x = v[0].x * c.ox + v[3].x * mx; f03.y = v[0].y * c.ox + v[3].y * mx;
  f12.x = v[1].x * c.ox + v[2].x * mx; f12.y = v[1].y * c.ox + v[2].y * mx;
  f56.x = v[5].x * c.ox + v[6].x * mx; f56.y = v[5].y * c.ox + v[6].y * mx;
  f47.x = v[4].x * c.ox + v[7].x * mx; f47.y = v[4].y * c.ox + v[7].y * mx;
  const float a0 = (f03.x * c.oy + f12.x * my) * c.oz + (f47.x * c.oy + f56.x * my) * mz;
  const float a1 = (f03.y * c.oy + f12.y * my) * c.oz + (f47.y * c.oy + f56.y * my) * mz;
  return pack_h2(a0, a1);
}

// A warp owns a 16-sample m-tile end to end (no block-level barriers: the 16 resident warps of an SM drift apart, so the
// L1-bound gathers of some overlap the tensor-pipe work of others).
//   gather: lane = (level parity, sample 0..15), so each half-warp request covers 16 CONSECUTIVE samples of one level
//     (neighbouring samples of a ray share cells / sectors at the coarse and middle levels) and x-neighbour rows go out
//     as one 16-byte load.  The forward gathers are bound by the L1TEX data pipe (one wavefront per distinct sector per
//     request; ncu: l1tex__data_pipe_lsu_wavefronts 85 % busy, L2 16-42 %), so wavefronts are what is minimised.
//   features are staged (fp16) in a warp-private 16x32 shared-memory tile and re-read as mma A fragments.
// KEEP: the training forward, which also writes what the backward needs (encoded features, positions, stash, ReLU flags); the eval / export
// instantiation carries none of that code (it cost 3-6 % of the render rate as run-time branches: registers)
template <bool KEEP>
__global__ void __launch_bounds__(THREADS, 2) k_field_mixed_fwd(const __grid_constant__ MixArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __half* Wsm = reinterpret_cast<__half*>(smem_raw);
  float* Bf = reinterpret_cast<float*>(smem_raw + HALVES * sizeof(__half));
  load_weights(a, Wsm, Bf);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int S = a.sm.samples_per_ray;
  const int64_t N = a.sm.num_rays * S;
  const int64_t ntiles = (N + 15) >> 4;
  __half* X0 = reinterpret_cast<__half*>(smem_raw + SMEM_FWD) + warp * 16 * XS;
  uint32_t* X32 = reinterpret_cast<uint32_t*>(X0);
  const int sidx = lane & 15, lpar = lane >> 4;
  for (int64_t tile = (int64_t)blockIdx.x * WARPS + warp; tile < ntiles; tile += (int64_t)gridDim.x * WARPS) {
    const int64_t base = tile * 16;
    // ---- position of this lane's gather sample -------------------------------------------------------------------------------
    float px = 0.f, py = 0.f, pz = 0.f;
    bool my_sel = false;
    {
      const int64_t i = base + sidx;
      if (i < N) {
        const int64_t r = cnb_ray_of(i, S);
        my_sel = cnb_sample_position(a.sm, a.warp, r, (int)(i - r * S), px, py, pz);
        if (a.pos_out && lpar == 0) { a.pos_out[3 * i] = px; a.pos_out[3 * i + 1] = py; a.pos_out[3 * i + 2] = pz; }
      }
    }
    __syncwarp();  // the previous tile's fragment reads of X0 are complete
    // two levels per lane in flight: 16 corner loads issued, THEN blended (the warp barrier keeps ptxas from sinking the second level's
    // loads behind the first level's blend; tests/micro/gather_ilp_bench: two levels in flight is the optimum)
#pragma unroll 1
    for (int it = 0; it < 8; it += 2) {
      const int l0 = 2 * it + lpar, l1 = l0 + 2;
      LevelLoads g0, g1;
      if (l0 < a.L) gather_issue(a.table, a.mask, (uint32_t)l0 * a.T, a.scalings[l0], px, py, pz, g0);
      if (l1 < a.L) gather_issue(a.table, a.mask, (uint32_t)l1 * a.T, a.scalings[l1], px, py, pz, g1);
      __syncwarp();
      X32[sidx * (XS / 2) + l0] = l0 < a.L ? gather_blend(g0) : 0u;
      X32[sidx * (XS / 2) + l1] = l1 < a.L ? gather_blend(g1) : 0u;
    }
    __syncwarp();
    int64_t row[2] = {base + g, base + g + 8};
    bool valid[2], sel[2];
    int64_t ray[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      valid[h] = row[h] < N;
      ray[h] = (valid[h] ? row[h] : N - 1) / S;
      sel[h] = __shfl_sync(0xffffffffu, my_sel ? 1 : 0, g + 8 * h) != 0;
    }
    uint32_t A0[2][4];
#pragma unroll
    for (int kt = 0; kt < 2; ++kt)
#pragma unroll
      for (int q = 0; q < 2; ++q)
#pragma unroll
        for (int h = 0; h < 2; ++h) A0[kt][2 * q + h] = X32[(g + 8 * h) * (XS / 2) + 8 * kt + 4 * q + t];
    if (KEEP && a.x0_out) {  // encoded features kept for the backward: the tile is one contiguous 1 KB block of the [N,32] fp16 array
#pragma unroll
      for (int e = lane; e < 64; e += 32) {
        const int rw = e >> 2, q = e & 3;
        if (base + rw < N) reinterpret_cast<uint4*>(a.x0_out + (base + rw) * 32)[q] = reinterpret_cast<const uint4*>(X0 + rw * XS)[q];
      }
    }
    // ---- base MLP: 32 -> 64 -> 16 --------------------------------------------------------------------------------
    uint32_t Abo[1][4];
    uint32_t mk[2][2] = {{0u, 0u}, {0u, 0u}};   // ReLU flags [fragment row][word]: word 0 = base hidden | semantic hidden 1 << 8, word 1 = rgb hidden 1 | 2 << 8
    {
      float acc[8][4];
      init_bias<8>(acc, Bf + F_BB1, t);
      layer<8, 2, S32>(Wsm + O_WB1, A0, acc, g, t);
      uint32_t AH[4][4];
      to_afrag<4, true>(acc, AH);
      if (KEEP && a.mask_out) relu_flags(AH, mk[0][0], mk[1][0]);
      float acc2[2][4];
      init_bias<2>(acc2, Bf + F_BB2, t);
      layer<2, 4, S64>(Wsm + O_WB2, AH, acc2, g, t);
      if (a.geo_out) {  // get_density's (density_before_activation, embedding) pair, fp32 accumulators as they are (fruit_field.py:183-187)
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
          if (valid[0]) *reinterpret_cast<float2*>(a.geo_out + row[0] * 16 + 8 * nt + 2 * t) = make_float2(acc2[nt][0], acc2[nt][1]);
          if (valid[1]) *reinterpret_cast<float2*>(a.geo_out + row[1] * 16 + 8 * nt + 2 * t) = make_float2(acc2[nt][2], acc2[nt][3]);
        }
      }
      if (t == 0) {  // column 0 = density before activation (fruit_field.py:185-193: trunc_exp in fp32, times the selector)
        if (valid[0]) a.density[row[0]] = sel[0] ? expf(acc2[0][0]) : 0.f;
        if (valid[1]) a.density[row[1]] = sel[1] ? expf(acc2[0][2]) : 0.f;
        if (KEEP && a.stash_out) {  // the backward's d(density)/d(pre-activation): trunc_exp' in fp32, times the selector
          if (valid[0]) a.stash_out[4 * row[0]] = sel[0] ? cnb_trunc_exp_grad(acc2[0][0]) : 0.f;
          if (valid[1]) a.stash_out[4 * row[1]] = sel[1] ? cnb_trunc_exp_grad(acc2[0][2]) : 0.f;
        }
        acc2[0][0] = 0.f; acc2[0][2] = 0.f;  // clear the dba slot: its weight columns downstream are zero
      }
      to_afrag<1, false>(acc2, Abo);
    }
    // ---- semantic MLP 15 -> 64 -> 64 and head 64 -> 1 (fruit_field.py:264-269) ------------------------------------
    if (a.sem) {
      float acc[8][4];
      init_bias<8>(acc, Bf + F_BS1, t);
      layer<8, 1, S16>(Wsm + O_WS1, Abo, acc, g, t);
      uint32_t AS1[4][4];
      to_afrag<4, true>(acc, AS1);
      if (KEEP && a.mask_out) {
        uint32_t f0, f1;
        relu_flags(AS1, f0, f1);
        mk[0][0] |= f0 << 8; mk[1][0] |= f1 << 8;
      }
      init_bias<8>(acc, Bf + F_BS2, t);
      layer<8, 4, S64>(Wsm + O_WS2, AS1, acc, g, t);
      float s0 = 0.f, s1 = 0.f;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const float2 w = *reinterpret_cast<const float2*>(Bf + F_WH + nt * 8 + 2 * t);
        s0 = fmaf(acc[nt][0], w.x, fmaf(acc[nt][1], w.y, s0));
        s1 = fmaf(acc[nt][2], w.x, fmaf(acc[nt][3], w.y, s1));
      }
      s0 += __shfl_xor_sync(0xffffffffu, s0, 1); s0 += __shfl_xor_sync(0xffffffffu, s0, 2);
      s1 += __shfl_xor_sync(0xffffffffu, s1, 1); s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
      if (t == 0) {
        const float bh = Bf[F_BH];
        if (valid[0]) a.sem[row[0]] = s0 + bh;
        if (valid[1]) a.sem[row[1]] = s1 + bh;
      }
    }
    // ---- RGB MLP [SH16 | geo15 | emb32] -> 64 -> 64 -> 3, sigmoid (fruit_field.py:271-280) ---------------------------
    if (a.rgb) {
      uint32_t Ain[4][4];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float c[16];
        cnb_sh16(__ldg(a.sm.directions + 3 * ray[h]), __ldg(a.sm.directions + 3 * ray[h] + 1), __ldg(a.sm.directions + 3 * ray[h] + 2), c);
        Ain[0][h] = pack_h2(pick4(t, c[0], c[2], c[4], c[6]), pick4(t, c[1], c[3], c[5], c[7]));
        Ain[0][2 + h] = pack_h2(pick4(t, c[8], c[10], c[12], c[14]), pick4(t, c[9], c[11], c[13], c[15]));
        const float* e = nullptr;
        if (a.app_mode == CNB_APP_PER_CAMERA) e = a.embedding + (int64_t)__ldg(a.sm.camera_indices + ray[h]) * 32;
        else if (a.app_mode == CNB_APP_MEAN) e = a.embedding;
#pragma unroll
        for (int kt = 0; kt < 2; ++kt) {
          float2 lo = make_float2(0.f, 0.f), hi = make_float2(0.f, 0.f);
          if (e) { lo = __ldg(reinterpret_cast<const float2*>(e + 16 * kt + 2 * t)); hi = __ldg(reinterpret_cast<const float2*>(e + 16 * kt + 8 + 2 * t)); }
          Ain[2 + kt][h] = pack_h2(lo.x, lo.y);
          Ain[2 + kt][2 + h] = pack_h2(hi.x, hi.y);
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) Ain[1][i] = Abo[0][i];
      float acc[8][4];
      init_bias<8>(acc, Bf + F_BR1, t);
      layer<8, 4, S64>(Wsm + O_WR1, Ain, acc, g, t);
      uint32_t AR[4][4];
      to_afrag<4, true>(acc, AR);
      if (KEEP && a.mask_out) relu_flags(AR, mk[0][1], mk[1][1]);
      init_bias<8>(acc, Bf + F_BR2, t);
      layer<8, 4, S64>(Wsm + O_WR2, AR, acc, g, t);
      to_afrag<4, true>(acc, AR);
      if (KEEP && a.mask_out) {
        uint32_t f0, f1;
        relu_flags(AR, f0, f1);
        mk[0][1] |= f0 << 8; mk[1][1] |= f1 << 8;
#pragma unroll
        for (int h = 0; h < 2; ++h)
          if (valid[h]) a.mask_out[row[h] * 4 + t] = make_uint2(mk[h][0], mk[h][1]);
      }
      float acc3[1][4];
      init_bias<1>(acc3, Bf + F_BR3, t);
      layer<1, 4, S64>(Wsm + O_WR3, AR, acc3, g, t);
      if (t < 2) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          if (!valid[h]) continue;
          const float v0 = 1.f / (1.f + __expf(-acc3[0][2 * h])), v1 = 1.f / (1.f + __expf(-acc3[0][2 * h + 1]));
          if (t == 0) { a.rgb[3 * row[h]] = v0; a.rgb[3 * row[h] + 1] = v1; }
          else a.rgb[3 * row[h] + 2] = v0;
          if (KEEP && a.stash_out) {
            if (t == 0) { a.stash_out[4 * row[h] + 1] = v0; a.stash_out[4 * row[h] + 2] = v1; }
            else a.stash_out[4 * row[h] + 3] = v0;
          }
        }
      }
    }
  }
}

}  // namespace

bool cnb_field_mixed_supported(const cnb_field* f) {
  const cnb_mlp &b = f->base, &s = f->sem, &r = f->rgb;
  return f->grid.num_levels <= 16 && b.num_layers == 2 && b.dims[0] == 2 * f->grid.num_levels && b.dims[1] == H && b.dims[2] == 16 &&
         s.num_layers == 2 && s.dims[0] == 15 && s.dims[1] == H && s.dims[2] == H && f->sem_head.dims[0] == H && f->sem_head.dims[1] == 1 &&
         r.num_layers == 3 && r.dims[0] == 63 && r.dims[1] == H && r.dims[2] == H && r.dims[3] == 3 && f->appearance_dim == 32 && f->geo_feat_dim == 15;
}

int64_t cnb_field_mixed_ctx_floats(int64_t n, int training) { return training ? ctx_total(n) : 0; }
float* cnb_field_mixed_dx0(float* ctx, int64_t n) { return ctx + ctx_dx0_off(n); }

int cnb_field_mixed_fwd(const cnb_field* f, const cnb_samples* s, float* density, float* geo, float* rgb, float* sem, float* positions_out, float* ctx,
                        int training, cudaStream_t stream) {
  MixArgs a;
  fill_args(f, s, a);
  const int64_t N = s->num_rays * s->samples_per_ray;
  a.density = density; a.rgb = rgb; a.sem = sem; a.pos_out = positions_out; a.x0_out = nullptr; a.geo_out = geo;
  if (training) {
    a.x0_out = reinterpret_cast<__half*>(ctx);
    a.pos_out = ctx + ctx_pos_off(N);
    a.stash_out = ctx + ctx_stash_off(N);
    a.mask_out = reinterpret_cast<uint2*>(ctx + ctx_mask_off(N));
  }
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(k_field_mixed_fwd<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_FWD2) != cudaSuccess ||
        cudaFuncSetAttribute(k_field_mixed_fwd<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_FWD2) != cudaSuccess)
      return cnb_check_launch("field_mixed_fwd attr");
    configured = true;
  }
  int64_t blocks = (N + TILE - 1) / TILE;  // one 16-sample m-tile per warp and round
  const int64_t cap = (int64_t)cnb_num_sms() * 2;
  if (blocks > cap) blocks = cap;
  if (training) k_field_mixed_fwd<true><<<(int)blocks, THREADS, SMEM_FWD2, stream>>>(a);
  else k_field_mixed_fwd<false><<<(int)blocks, THREADS, SMEM_FWD2, stream>>>(a);
  int rc = cnb_check_launch("field_mixed_fwd");
  if (rc) return rc;
  if (training && positions_out != nullptr &&
      cudaMemcpyAsync(positions_out, a.pos_out, sizeof(float) * 3 * N, cudaMemcpyDeviceToDevice, stream) != cudaSuccess)
    return cnb_check_launch("field_mixed_fwd positions copy");
  return CNB_OK;
}
