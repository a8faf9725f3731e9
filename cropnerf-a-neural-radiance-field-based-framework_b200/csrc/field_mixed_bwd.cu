// FruitField backward on tensor cores (rows a1/a5/a6 of SURVEY.md section 8, the autograd of fruit_field.py:169-302).
//
// One kernel does, per 128-sample batch of a persistent CTA (8 warps, one 16-sample m-tile each):
//   1. re-run the forward MLPs from the fp16 encoded features the forward kept (no second table gather); the fp16
//      activations stay in registers as mma A fragments, exactly as in field_mixed.cu, so the ReLU masks are the
//      forward's;
//   2. input gradients layer by layer: dX = dY * W with dY re-packed from the accumulator fragments (bf16 -- the range of
//      fp32, so no loss scaling is needed for the 1e-7-sized pixel gradients) and W^T read from a transposed bf16 copy
//      of the weights in shared memory;
//   3. weight gradients: dW = dY^T * X is a contraction over SAMPLES, so each layer's (X, dY) pair of the 128 samples
//      is staged once in shared memory and read back with ldmatrix.trans; every warp owns a fixed 1/8 slice of every
//      dW and accumulates it (fp32) across all batches of the CTA -- no atomics until the one flush per CTA at the end;
//      bias gradients are the same contraction against a fragment of ones;
//   4. d(encoded features), level-major [16][N][2] fp32, goes to a scratch array; the appearance embedding gradient is reduced over the
//      m-tile (all 16 samples share the ray when S % 16 == 0) before its atomics;
//   5. (FUSED, the default) the hash-table scatter of those d(features) runs INSIDE this kernel on four extra "scatter" warps: the MLP
//      part is a latency chain on the tensor pipe at 8 warps/SM (1 CTA/SM: 228 KB of shared memory), the scatter is bound by the L2
//      reduction rate and needs no shared memory -- run back to back as two kernels they cost 0.164 + 0.104 ms, but no second kernel can
//      co-reside with a 228 KB CTA, so the overlap has to happen inside the CTA.  The MLP warps publish each finished 128-sample batch
//      through a ring of full/empty mbarriers; scatter warp w takes samples [32 w, 32 w + 32) of the batch, reads its 16 levels'
//      d(features) back (L2-resident, written microseconds earlier by the same SM) and issues the aggregated vector reductions of
//      cnb_scatter_cell.  Registers are re-split with setmaxnreg (MLP warpgroups 224, scatter warpgroup 56; 384 threads x 168 at launch).
#include <cstdlib>

#include "field_mixed.cuh"

using namespace cnbmix;

namespace {

constexpr int ST = 72;        // staging row stride (halves)
constexpr int BATCH = WARPS * 16;
// transposed bf16 weights, [in][out] with padded rows (half offsets)
constexpr int T_R3 = 0;                   // [64][24]   out 0..2 of 16 used
constexpr int T_R2 = T_R3 + 64 * 24;      // [64][72]
constexpr int T_R1 = T_R2 + 64 * 72;      // [64][72]   rows = padded rgb input [SH16 | 0,geo15 | emb32]
constexpr int T_S2 = T_R1 + 64 * 72;      // [64][72]
constexpr int T_B2 = T_S2 + 64 * 72;      // [64][24]
constexpr int T_B1 = T_B2 + 64 * 24;      // [32][72]
constexpr int T_HALVES = T_B1 + 32 * 72;
// per-warp dW accumulator tiles (16x8 fp32 each): r3, r2 x4, r1 x4, head, s2 x4, s1, b2, b1 x2
constexpr int A_R3 = 0, A_R2 = 1, A_R1 = 5, A_H = 9, A_S2 = 10, A_S1 = 14, A_B2 = 15, A_B1 = 16, A_TILES = 18;
// bias accumulators (floats)
constexpr int B_R3 = 0, B_R2 = 16, B_R1 = 80, B_H = 144, B_S2 = 160, B_S1 = 224, B_B2 = 288, B_B1 = 304, B_FLOATS = 368;

constexpr size_t SMEM_BWD = (size_t)(HALVES + T_HALVES + 4 * BATCH * ST) * 2 + (size_t)(FLOATS + B_FLOATS + WARPS * A_TILES * 128) * 4;  // 228 272 B
// fused scatter: 4 scatter warps behind the 8 MLP warps, NSLOT-deep full/empty mbarrier ring (after SMEM_BWD)
constexpr int SC_WARPS = 4;
constexpr int THREADS_FUSED = THREADS + SC_WARPS * 32;
constexpr int NSLOT = 4;
constexpr size_t SMEM_BWD_FUSED = SMEM_BWD + 2 * NSLOT * sizeof(uint64_t);
static_assert(BATCH == SC_WARPS * 32, "one scatter warp per 32 samples of a batch");
static_assert(SMEM_BWD_FUSED <= 232448, "227 KB of dynamic shared memory per CTA");

struct BwdArgs {
  MixArgs m;
  float* d_table;               // FUSED: gradient table of the hash grid
  const __half* x0;
  const float* pos;
  const float *d_density, *d_rgb, *d_sem;
  float* d_x0;
  float *dWb1, *dbb1, *dWb2, *dbb2, *dWs1, *dbs1, *dWs2, *dbs2, *dWh, *dbh, *dWr1, *dbr1, *dWr2, *dbr2, *dWr3, *dbr3;
  float* d_embedding;
};

__device__ __forceinline__ uint32_t pack_bf2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float2 unpack_h2(uint32_t v) { return __half22float2(*reinterpret_cast<__half2*>(&v)); }
__device__ __forceinline__ uint32_t h2_to_bf2(uint32_t v) {
  const float2 f = unpack_h2(v);
  return pack_bf2(f.x, f.y);
}

__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// dX: acc[NT][4] += dY[KT][4] x WT^T, WT = smem bf16 [8*NT rows (= layer inputs)][STRIDE] indexed [in][out]
template <int NT, int KT, int STRIDE>
__device__ __forceinline__ void layer_bf(const __nv_bfloat16* __restrict__ W, const uint32_t (&A)[KT][4], float (&acc)[NT][4], int g, int t) {
  (void)g; (void)t;
  layer_ldsm<NT, KT, STRIDE>(W, A, acc, [](float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) { mma_bf16(c, a, b0, b1); });
}

template <int NT>
__device__ __forceinline__ void zero_acc(float (&acc)[NT][4]) {
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) { acc[nt][0] = 0.f; acc[nt][1] = 0.f; acc[nt][2] = 0.f; acc[nt][3] = 0.f; }
}

// dY (bf16 A fragments) = acc masked by the ReLU of the fp16 activations Act (same fragment geometry)
template <int KT>
__device__ __forceinline__ void relu_mask_pack(const float (&acc)[2 * KT][4], const uint32_t (&Act)[KT][4], uint32_t (&D)[KT][4]) {
#pragma unroll
  for (int kt = 0; kt < KT; ++kt) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const float2 lo = unpack_h2(Act[kt][2 * h]), hi = unpack_h2(Act[kt][2 * h + 1]);
      D[kt][2 * h] = pack_bf2(lo.x > 0.f ? acc[2 * kt + h][0] : 0.f, lo.y > 0.f ? acc[2 * kt + h][1] : 0.f);
      D[kt][2 * h + 1] = pack_bf2(hi.x > 0.f ? acc[2 * kt + h][2] : 0.f, hi.y > 0.f ? acc[2 * kt + h][3] : 0.f);
    }
  }
}

// write KT k-tiles of an A fragment into the staging matrix [BATCH][ST] (32-bit view, row stride ST/2 words)
template <int KT, bool CVT>
__device__ __forceinline__ void stage(uint32_t* st32, int row0, const uint32_t (&A)[KT][4], int g, int t) {
  uint32_t* r0 = st32 + (row0 + g) * (ST / 2) + t;
  uint32_t* r1 = r0 + 8 * (ST / 2);
#pragma unroll
  for (int kt = 0; kt < KT; ++kt) {
    r0[8 * kt] = CVT ? h2_to_bf2(A[kt][0]) : A[kt][0];
    r1[8 * kt] = CVT ? h2_to_bf2(A[kt][1]) : A[kt][1];
    r0[8 * kt + 4] = CVT ? h2_to_bf2(A[kt][2]) : A[kt][2];
    r1[8 * kt + 4] = CVT ? h2_to_bf2(A[kt][3]) : A[kt][3];
  }
}

__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x2_t(uint32_t (&r)[2], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];\n" : "=r"(r[0]), "=r"(r[1]) : "r"(addr));
}

// dW[16*MT, 8*NTT] += dY^T X over the BATCH staged samples.  Warp w owns m-tile w / WPM and NT consecutive n-tiles.
template <int MT, int NTT>
__device__ __forceinline__ void dw_gemm(uint32_t dout_s, uint32_t in_s, float* acc_tiles, float* bias_acc, int warp, int lane) {
  constexpr int WPM = WARPS / MT;
  constexpr int NT = NTT / WPM;
  static_assert(NT == 1 || NT % 2 == 0, "n-tiles per warp");
  const int mi = warp / WPM, nj0 = (warp % WPM) * NT;
  const bool do_bias = (warp % WPM) == 0;
  float c[NT][4], cb[4] = {0.f, 0.f, 0.f, 0.f};
  zero_acc<NT>(c);
  const int j = lane >> 3, r = lane & 7;
  const uint32_t a_addr = dout_s + 2u * (((j >> 1) * 8 + r) * ST + 16 * mi + (j & 1) * 8);
  const uint32_t b_addr = in_s + 2u * (((j & 1) * 8 + r) * ST + 8 * (nj0 + (j >> 1)));
  constexpr uint32_t ONES = 0x3F803F80u;  // bf16 (1, 1)
#pragma unroll
  for (int ks = 0; ks < BATCH / 16; ++ks) {
    uint32_t A[4];
    ldsm_x4_t(A, a_addr + 2u * ks * 16 * ST);
    if constexpr (NT == 1) {
      uint32_t B[2];
      ldsm_x2_t(B, b_addr + 2u * ks * 16 * ST);
      mma_bf16(c[0], A, B[0], B[1]);
    } else {
#pragma unroll
      for (int p = 0; p < NT / 2; ++p) {
        uint32_t B[4];
        ldsm_x4_t(B, b_addr + 2u * (ks * 16 * ST + 16 * p));
        mma_bf16(c[2 * p], A, B[0], B[1]);
        mma_bf16(c[2 * p + 1], A, B[2], B[3]);
      }
    }
    if (do_bias) mma_bf16(cb, A, ONES, ONES);
  }
#pragma unroll
  for (int nt = 0; nt < NT; ++nt)
#pragma unroll
    for (int q = 0; q < 4; ++q) acc_tiles[(nt * 4 + q) * 32 + lane] += c[nt][q];
  if (do_bias && (lane & 3) == 0) {
    bias_acc[16 * mi + (lane >> 2)] += cb[0];
    bias_acc[16 * mi + (lane >> 2) + 8] += cb[2];
  }
}

// flush one warp-owned group of accumulator tiles: element (n, k) of the padded dW goes through `put`
template <int MT, int NTT, typename Put>
__device__ __forceinline__ void flush_tiles(const float* acc_tiles, int warp, int lane, Put put) {
  constexpr int WPM = WARPS / MT;
  constexpr int NT = NTT / WPM;
  const int mi = warp / WPM, nj0 = (warp % WPM) * NT;
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int nt = 0; nt < NT; ++nt)
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float v = acc_tiles[(nt * 4 + q) * 32 + lane];
      if (v != 0.f) put(16 * mi + g + (q >> 1) * 8, 8 * (nj0 + nt) + 2 * t + (q & 1), v);
    }
}

// block barrier over the 8 MLP warps only (the scatter warps of the fused kernel never take part)
__device__ __forceinline__ void bar_mlp() { asm volatile("bar.sync 1, %0;" ::"n"(THREADS) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  } while (!done);
}

// Scatter role of the fused kernel: warp sw owns samples [32 sw, 32 sw + 32) of every batch of this CTA.
__device__ __forceinline__ void scatter_role(const MixArgs& a, float* __restrict__ d_table, const float* __restrict__ pos, const float* __restrict__ d_x0,
                                             uint32_t mbar_s, int sw, int lane, int64_t N, int64_t nbatches) {
  int j = 0;
  for (int64_t batch = blockIdx.x; batch < nbatches; batch += gridDim.x, ++j) {
    const int slot = j % NSLOT;
    const uint32_t par = (uint32_t)(j / NSLOT) & 1u;
    const int64_t s = batch * BATCH + sw * 32 + lane;
    const bool in = s < N;
    // the positions come from the forward: requested before the wait
    float px = 0.f, py = 0.f, pz = 0.f;
    if (in) { px = __ldg(pos + 3 * s); py = __ldg(pos + 3 * s + 1); pz = __ldg(pos + 3 * s + 2); }
    mbar_wait(mbar_s + 8 * slot, par);  // "full": all 256 MLP threads have stored this batch's d(features)
#pragma unroll 1
    for (int l0 = 0; l0 < a.L; l0 += 4) {
      float2 d[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        d[q] = make_float2(0.f, 0.f);
        if (in && l0 + q < a.L) d[q] = __ldcg(reinterpret_cast<const float2*>(d_x0) + (int64_t)(l0 + q) * N + s);
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int l = l0 + q;
        if (l < a.L) {  // warp-uniform
          const bool active = d[q].x != 0.0f || d[q].y != 0.0f;  // zero gradients add nothing (masked samples, App. B-3)
          CnbCell c = {};
          if (active) c = cnb_cell(px, py, pz, a.scalings[l]);
          cnb_scatter_cell(d_table, c, a.mask, (uint32_t)l * a.T, d[q].x, d[q].y, active);
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(mbar_s + 8 * (NSLOT + slot));  // "empty"
  }
}

__device__ inline void load_weights_t(const MixArgs& a, __nv_bfloat16* WT) {
  const int tid = threadIdx.x;
  for (int e = tid; e < 64 * 24; e += THREADS) {
    const int in = e / 24, out = e - in * 24;
    WT[T_R3 + e] = __float2bfloat16_rn(out < 3 ? __ldg(a.Wr3 + out * 64 + in) : 0.f);
    WT[T_B2 + e] = __float2bfloat16_rn(out < 16 ? __ldg(a.Wb2 + out * 64 + in) : 0.f);
  }
  for (int e = tid; e < 64 * 72; e += THREADS) {
    const int in = e / 72, out = e - in * 72;
    const bool ok = out < 64;
    WT[T_R2 + e] = __float2bfloat16_rn(ok ? __ldg(a.Wr2 + out * 64 + in) : 0.f);
    WT[T_S2 + e] = __float2bfloat16_rn(ok ? __ldg(a.Ws2 + out * 64 + in) : 0.f);
    float w1 = 0.f;
    if (ok) {
      if (in < 16) w1 = __ldg(a.Wr1 + out * 63 + in);
      else if (in >= 17) w1 = __ldg(a.Wr1 + out * 63 + (in - 1));
    }
    WT[T_R1 + e] = __float2bfloat16_rn(w1);
  }
  for (int e = tid; e < 32 * 72; e += THREADS) {
    const int in = e / 72, out = e - in * 72;
    WT[T_B1 + e] = __float2bfloat16_rn((out < 64 && in < a.in0) ? __ldg(a.Wb1 + out * a.in0 + in) : 0.f);
  }
}

template <bool FUSED>
__global__ void __launch_bounds__(FUSED ? THREADS_FUSED : THREADS, 1) k_field_mixed_bwd(const __grid_constant__ BwdArgs b) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const MixArgs& a = b.m;
  __half* Wsm = reinterpret_cast<__half*>(smem_raw);
  __nv_bfloat16* WT = reinterpret_cast<__nv_bfloat16*>(Wsm + HALVES);
  // two staging buffer pairs (X, dY), used alternately: a warp may stage GEMM k+1 while others still contract GEMM k, so one
  // block barrier per GEMM suffices (the buffer being overwritten was last read two GEMMs ago, i.e. before the previous barrier)
  __nv_bfloat16* st_base = WT + T_HALVES;
  float* Bf = reinterpret_cast<float*>(st_base + 4 * BATCH * ST);
  float* bias_acc = Bf + FLOATS;
  float* acc_all = bias_acc + B_FLOATS;
  const uint32_t mbar_s = (uint32_t)__cvta_generic_to_shared(smem_raw + SMEM_BWD);  // FUSED: NSLOT "full" then NSLOT "empty" barriers
  if (!FUSED || threadIdx.x < THREADS) {
    load_weights(a, Wsm, Bf);
    load_weights_t(a, WT);
    for (int e = threadIdx.x; e < B_FLOATS + WARPS * A_TILES * 128; e += THREADS) bias_acc[e] = 0.f;
  } else if (threadIdx.x == THREADS) {
    for (int i = 0; i < NSLOT; ++i) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar_s + 8 * i), "r"(THREADS));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar_s + 8 * (NSLOT + i)), "r"(SC_WARPS));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // the pad columns of the staging rows are never written by `stage`; ldmatrix.x2 never consumes them either
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (FUSED) {
    const int64_t N_ = a.sm.num_rays * a.sm.samples_per_ray;
    if (warp >= WARPS) {
      asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
      scatter_role(a, b.d_table, b.pos, b.d_x0, mbar_s, warp - WARPS, lane, N_, (N_ + BATCH - 1) / BATCH);
      return;
    }
    asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
  }
  const int g = lane >> 2, t = lane & 3;
  float* acc_w = acc_all + warp * A_TILES * 128;
  uint32_t* const st32 = reinterpret_cast<uint32_t*>(st_base);
  const uint32_t st_s = (uint32_t)__cvta_generic_to_shared(st_base);
  constexpr int MATW = BATCH * ST / 2;  // 32-bit words per staged matrix
  int buf = 0;
  const int row0 = warp * 16;
  const int S = a.sm.samples_per_ray;
  const int64_t N = a.sm.num_rays * S;
  const int64_t nbatches = (N + BATCH - 1) / BATCH;

  // Software pipeline over the batches of this CTA: the encoded features and camera indices of the NEXT tile are requested
  // while the current one is processed, and every other global input of the current tile (directions, embedding row,
  // incoming gradients) is requested at the top of the iteration, long before its first use -- with one CTA (2 warps per
  // scheduler) on the SM nothing else hides these latencies (ncu: long-scoreboard was 32 % of all issue stalls).
  struct TileIn {           // every global input of one m-tile, as this lane needs it
    uint32_t A0[2][4];      // encoded features (A fragments)
    float dirv[2][3];
    float2 embv[2][2][2];
    float drgb[2][2], dsem[2], ddens[2], posx[2];
    int cam[2];
  };
  auto load_cam = [&](int64_t tl, int (&cam)[2]) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int64_t rw = tl * 16 + g + 8 * h;
      cam[h] = 0;
      if (a.app_mode == CNB_APP_PER_CAMERA) cam[h] = __ldg(a.sm.camera_indices + (rw < N ? rw : N - 1) / S);
    }
  };
  auto load_tile = [&](int64_t tl, const int (&cam)[2], TileIn& in) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int64_t rw = tl * 16 + g + 8 * h;
      const bool ok = rw < N;
      const int64_t ry = (ok ? rw : N - 1) / S;
      in.cam[h] = cam[h];
#pragma unroll
      for (int kt = 0; kt < 2; ++kt)
#pragma unroll
        for (int q = 0; q < 2; ++q) in.A0[kt][2 * q + h] = ok ? __ldg(reinterpret_cast<const uint32_t*>(b.x0 + rw * 32 + 16 * kt + 8 * q + 2 * t)) : 0u;
      in.dirv[h][0] = __ldg(a.sm.directions + 3 * ry); in.dirv[h][1] = __ldg(a.sm.directions + 3 * ry + 1); in.dirv[h][2] = __ldg(a.sm.directions + 3 * ry + 2);
      const float* e = nullptr;
      if (a.app_mode == CNB_APP_PER_CAMERA) e = a.embedding + (int64_t)cam[h] * 32;
      else if (a.app_mode == CNB_APP_MEAN) e = a.embedding;
#pragma unroll
      for (int kt = 0; kt < 2; ++kt) {
        in.embv[h][kt][0] = e ? __ldg(reinterpret_cast<const float2*>(e + 16 * kt + 2 * t)) : make_float2(0.f, 0.f);
        in.embv[h][kt][1] = e ? __ldg(reinterpret_cast<const float2*>(e + 16 * kt + 8 + 2 * t)) : make_float2(0.f, 0.f);
      }
      in.drgb[h][0] = 0.f; in.drgb[h][1] = 0.f; in.dsem[h] = 0.f; in.ddens[h] = 0.f; in.posx[h] = 0.f;
      if (ok) {
        if (b.d_rgb && t < 2) {
          in.drgb[h][0] = __ldg(b.d_rgb + 3 * rw + 2 * t);
          if (t == 0) in.drgb[h][1] = __ldg(b.d_rgb + 3 * rw + 1);
        }
        if (b.d_sem) in.dsem[h] = __ldg(b.d_sem + rw);
        if (t == 0 && b.d_density) { in.ddens[h] = __ldg(b.d_density + rw); in.posx[h] = __ldg(b.pos + 3 * rw); }
      }
    }
  };
  // Software pipeline over the batches of this CTA (one CTA, i.e. 2 warps per scheduler, per SM: nothing else hides global
  // latency; ncu: long-scoreboard was 32 % of all issue stalls): all inputs of tile i+1 and the camera indices of tile i+2
  // are requested while tile i is processed; they are loop-carried registers, so the compiler cannot sink the loads.
  const int64_t tstride = (int64_t)gridDim.x * WARPS;
  TileIn nxt;
  int cam2[2];
  {
    const int64_t t0 = (int64_t)blockIdx.x * WARPS + warp;
    int cam1[2];
    load_cam(t0, cam1);
    load_tile(t0, cam1, nxt);
    load_cam(t0 + tstride, cam2);
  }

  int jb = 0;  // FUSED: index of this CTA's batch in the full / empty barrier ring
  for (int64_t batch = blockIdx.x; batch < nbatches; batch += gridDim.x, ++jb) {
    const int64_t tile = batch * WARPS + warp;
    const int64_t row[2] = {tile * 16 + g, tile * 16 + g + 8};
    bool valid[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) valid[h] = row[h] < N;
    const TileIn cur = nxt;
    load_tile(tile + tstride, cam2, nxt);
    load_cam(tile + 2 * tstride, cam2);
    uint32_t A0[2][4];
#pragma unroll
    for (int kt = 0; kt < 2; ++kt)
#pragma unroll
      for (int q = 0; q < 4; ++q) A0[kt][q] = cur.A0[kt][q];
    const int cam[2] = {cur.cam[0], cur.cam[1]};
    const float (&dirv)[2][3] = cur.dirv;
    const float2 (&embv)[2][2][2] = cur.embv;
    const float (&drgb)[2][2] = cur.drgb;
    const float (&dsem)[2] = cur.dsem;
    const float (&ddens)[2] = cur.ddens;
    const float (&posx)[2] = cur.posx;
    // ---- base MLP forward ---------------------------------------------------------------------------------------------
    uint32_t AH[4][4], Abo[1][4];
    float dba[2] = {0.f, 0.f};
    {
      float acc[8][4];
      init_bias<8>(acc, Bf + F_BB1, t);
      layer<8, 2, S32>(Wsm + O_WB1, A0, acc, g, t);
      to_afrag<4, true>(acc, AH);
      float acc2[2][4];
      init_bias<2>(acc2, Bf + F_BB2, t);
      layer<2, 4, S64>(Wsm + O_WB2, AH, acc2, g, t);
      if (t == 0) { dba[0] = acc2[0][0]; dba[1] = acc2[0][2]; acc2[0][0] = 0.f; acc2[0][2] = 0.f; }
      to_afrag<1, false>(acc2, Abo);
    }
    // gradient that reaches the base-MLP output [d_dba | d_geo15]; filled by the rgb branch, col 0 by the density
    float dbo[2][4];
    zero_acc<2>(dbo);

    // ===================================== RGB branch ===================================================================
    {
      uint32_t Ain[4][4], AR1[4][4], AR2[4][4];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float c[16];
        cnb_sh16(dirv[h][0], dirv[h][1], dirv[h][2], c);
        Ain[0][h] = pack_h2(pick4(t, c[0], c[2], c[4], c[6]), pick4(t, c[1], c[3], c[5], c[7]));
        Ain[0][2 + h] = pack_h2(pick4(t, c[8], c[10], c[12], c[14]), pick4(t, c[9], c[11], c[13], c[15]));
#pragma unroll
        for (int kt = 0; kt < 2; ++kt) {
          Ain[2 + kt][h] = pack_h2(embv[h][kt][0].x, embv[h][kt][0].y);
          Ain[2 + kt][2 + h] = pack_h2(embv[h][kt][1].x, embv[h][kt][1].y);
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) Ain[1][i] = Abo[0][i];
      float acc[8][4];
      init_bias<8>(acc, Bf + F_BR1, t);
      layer<8, 4, S64>(Wsm + O_WR1, Ain, acc, g, t);
      to_afrag<4, true>(acc, AR1);
      init_bias<8>(acc, Bf + F_BR2, t);
      layer<8, 4, S64>(Wsm + O_WR2, AR1, acc, g, t);
      to_afrag<4, true>(acc, AR2);
      float acc3[1][4];
      init_bias<1>(acc3, Bf + F_BR3, t);
      layer<1, 4, S64>(Wsm + O_WR3, AR2, acc3, g, t);
      // d(rgb pre-activation) = d_rgb * sigmoid'  (columns 0,1 on t==0, column 2 on t==1)
      uint32_t D3[1][4] = {{0u, 0u, 0u, 0u}};
      if (t < 2 && b.d_rgb) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          if (!valid[h]) continue;
          const float s0 = 1.f / (1.f + __expf(-acc3[0][2 * h])), s1 = 1.f / (1.f + __expf(-acc3[0][2 * h + 1]));
          float d0, d1 = 0.f;
          if (t == 0) { d0 = drgb[h][0] * s0 * (1.f - s0); d1 = drgb[h][1] * s1 * (1.f - s1); }
          else d0 = drgb[h][0] * s0 * (1.f - s0);
          D3[0][h] = pack_bf2(d0, d1);
        }
      }
      // ---- layer 3: dW = D3^T r2 ; d_r2 = D3 W3 -------------------------------------------------------------------------
      stage<1, false>(st32 + (2 * buf + 1) * MATW, row0, D3, g, t);
      stage<4, true>(st32 + (2 * buf) * MATW, row0, AR2, g, t);
      bar_mlp();
      dw_gemm<1, 8>(st_s + (2 * buf + 1) * MATW * 4, st_s + (2 * buf) * MATW * 4, acc_w + A_R3 * 128, bias_acc + B_R3, warp, lane);
      buf ^= 1;
      uint32_t D[4][4];
      zero_acc<8>(acc);
      layer_bf<8, 1, 24>(WT + T_R3, D3, acc, g, t);
      relu_mask_pack<4>(acc, AR2, D);
      // ---- layer 2 ----------------------------------------------------------------------------------------------------------
      stage<4, false>(st32 + (2 * buf + 1) * MATW, row0, D, g, t);
      stage<4, true>(st32 + (2 * buf) * MATW, row0, AR1, g, t);
      bar_mlp();
      dw_gemm<4, 8>(st_s + (2 * buf + 1) * MATW * 4, st_s + (2 * buf) * MATW * 4, acc_w + A_R2 * 128, bias_acc + B_R2, warp, lane);
      buf ^= 1;
      zero_acc<8>(acc);
      layer_bf<8, 4, 72>(WT + T_R2, D, acc, g, t);
      relu_mask_pack<4>(acc, AR1, D);
      // ---- layer 1 ----------------------------------------------------------------------------------------------------------
      stage<4, false>(st32 + (2 * buf + 1) * MATW, row0, D, g, t);
      stage<4, true>(st32 + (2 * buf) * MATW, row0, Ain, g, t);
      bar_mlp();
      dw_gemm<4, 8>(st_s + (2 * buf + 1) * MATW * 4, st_s + (2 * buf) * MATW * 4, acc_w + A_R1 * 128, bias_acc + B_R1, warp, lane);
      buf ^= 1;
      float din[6][4];  // d(rgb input) columns 16..63: [0, geo15 | emb32]
      zero_acc<6>(din);
      layer_bf<6, 4, 72>(WT + T_R1 + 16 * 72, D, din, g, t);
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int q = 0; q < 4; ++q) dbo[nt][q] = din[nt][q];
      // appearance-embedding gradient (fruit_field.py:251-258: per-camera rows in training)
      if (a.app_mode == CNB_APP_PER_CAMERA && b.d_embedding != nullptr) {
        const int64_t first = tile * 16, last = tile * 16 + 15;
        if (last < N && first / S == last / S) {  // whole m-tile on one ray: reduce over its 16 samples first
          float* dst = b.d_embedding + (int64_t)cam[0] * 32;
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) {
            float v0 = din[2 + nt][0] + din[2 + nt][2], v1 = din[2 + nt][1] + din[2 + nt][3];
#pragma unroll
            for (int o = 4; o < 32; o <<= 1) { v0 += __shfl_xor_sync(0xffffffffu, v0, o); v1 += __shfl_xor_sync(0xffffffffu, v1, o); }
            if (g == 0) { atomicAdd(dst + 8 * nt + 2 * t, v0); atomicAdd(dst + 8 * nt + 2 * t + 1, v1); }
          }
        } else {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            if (!valid[h]) continue;
            float* dst = b.d_embedding + (int64_t)cam[h] * 32;
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) { atomicAdd(dst + 8 * nt + 2 * t, din[2 + nt][2 * h]); atomicAdd(dst + 8 * nt + 2 * t + 1, din[2 + nt][2 * h + 1]); }
          }
        }
      }
    }

    // ===================================== semantic branch (input detached: fruit_field.py:264-266) ======================
    {
      uint32_t AS1[4][4], D[4][4];
      float acc[8][4];
      init_bias<8>(acc, Bf + F_BS1, t);
      layer<8, 1, S16>(Wsm + O_WS1, Abo, acc, g, t);
      to_afrag<4, true>(acc, AS1);
      init_bias<8>(acc, Bf + F_BS2, t);
      layer<8, 4, S64>(Wsm + O_WS2, AS1, acc, g, t);
      uint32_t S2f[4][4];
#pragma unroll
      for (int kt = 0; kt < 4; ++kt)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          S2f[kt][2 * h] = pack_bf2(acc[2 * kt + h][0], acc[2 * kt + h][1]);
          S2f[kt][2 * h + 1] = pack_bf2(acc[2 * kt + h][2], acc[2 * kt + h][3]);
        }
      // ---- head: dWh = d_sem^T s2 ------------------------------------------------------------------------------------------
      uint32_t Dh[1][4] = {{0u, 0u, 0u, 0u}};
      if (t == 0) { Dh[0][0] = pack_bf2(dsem[0], 0.f); Dh[0][1] = pack_bf2(dsem[1], 0.f); }
      stage<1, false>(st32 + (2 * buf + 1) * MATW, row0, Dh, g, t);
      stage<4, false>(st32 + (2 * buf) * MATW, row0, S2f, g, t);
      bar_mlp();
      dw_gemm<1, 8>(st_s + (2 * buf + 1) * MATW * 4, st_s + (2 * buf) * MATW * 4, acc_w + A_H * 128, bias_acc + B_H, warp, lane);
      buf ^= 1;
      // d_s2 = d_sem * Wh (no activation after the last semantic layer)
#pragma unroll
      for (int kt = 0; kt < 4; ++kt)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const float2 w = *reinterpret_cast<const float2*>(Bf + F_WH + (2 * kt + h) * 8 + 2 * t);
          D[kt][2 * h] = pack_bf2(dsem[0] * w.x, dsem[0] * w.y);
          D[kt][2 * h + 1] = pack_bf2(dsem[1] * w.x, dsem[1] * w.y);
        }
      // ---- semantic layer 2 ------------------------------------------------------------------------------------------------
      stage<4, false>(st32 + (2 * buf + 1) * MATW, row0, D, g, t);
      stage<4, true>(st32 + (2 * buf) * MATW, row0, AS1, g, t);
      bar_mlp();
      dw_gemm<4, 8>(st_s + (2 * buf + 1) * MATW * 4, st_s + (2 * buf) * MATW * 4, acc_w + A_S2 * 128, bias_acc + B_S2, warp, lane);
      buf ^= 1;
      zero_acc<8>(acc);
      layer_bf<8, 4, 72>(WT + T_S2, D, acc, g, t);
      relu_mask_pack<4>(acc, AS1, D);
      // ---- semantic layer 1 (input = [0 | geo15]) -----------------------------------------------------------------------------
      stage<4, false>(st32 + (2 * buf + 1) * MATW, row0, D, g, t);
      stage<1, true>(st32 + (2 * buf) * MATW, row0, Abo, g, t);
      bar_mlp();
      dw_gemm<4, 2>(st_s + (2 * buf + 1) * MATW * 4, st_s + (2 * buf) * MATW * 4, acc_w + A_S1 * 128, bias_acc + B_S1, warp, lane);
      buf ^= 1;
    }

    // ===================================== base MLP ========================================================================
    {
      if (t == 0 && b.d_density) {  // trunc_exp backward (fp32) times the selector (fruit_field.py:185-193)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          float gd = 0.f;
          if (valid[h] && posx[h] > 0.f) gd = ddens[h] * cnb_trunc_exp_grad(dba[h]);
          dbo[0][2 * h] = gd;
        }
      } else if (t == 0) { dbo[0][0] = 0.f; dbo[0][2] = 0.f; }
      uint32_t Dbo[1][4];
      Dbo[0][0] = pack_bf2(dbo[0][0], dbo[0][1]); Dbo[0][1] = pack_bf2(dbo[0][2], dbo[0][3]);
      Dbo[0][2] = pack_bf2(dbo[1][0], dbo[1][1]); Dbo[0][3] = pack_bf2(dbo[1][2], dbo[1][3]);
      stage<1, false>(st32 + (2 * buf + 1) * MATW, row0, Dbo, g, t);
      stage<4, true>(st32 + (2 * buf) * MATW, row0, AH, g, t);
      bar_mlp();
      dw_gemm<1, 8>(st_s + (2 * buf + 1) * MATW * 4, st_s + (2 * buf) * MATW * 4, acc_w + A_B2 * 128, bias_acc + B_B2, warp, lane);
      buf ^= 1;
      float acc[8][4];
      zero_acc<8>(acc);
      layer_bf<8, 1, 24>(WT + T_B2, Dbo, acc, g, t);
      uint32_t D[4][4];
      relu_mask_pack<4>(acc, AH, D);
      stage<4, false>(st32 + (2 * buf + 1) * MATW, row0, D, g, t);
      stage<2, true>(st32 + (2 * buf) * MATW, row0, A0, g, t);
      bar_mlp();
      dw_gemm<4, 4>(st_s + (2 * buf + 1) * MATW * 4, st_s + (2 * buf) * MATW * 4, acc_w + A_B1 * 128, bias_acc + B_B1, warp, lane);
      buf ^= 1;
      float dx[4][4];
      zero_acc<4>(dx);
      layer_bf<4, 4, 72>(WT + T_B1, D, dx, g, t);
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        // level-major [16][N][2]: this thread's accumulator pair is level 4 nt + t of its two rows
        float2* const dlev = reinterpret_cast<float2*>(b.d_x0) + (int64_t)(4 * nt + t) * N;
        if (valid[0]) dlev[row[0]] = make_float2(dx[nt][0], dx[nt][1]);
        if (valid[1]) dlev[row[1]] = make_float2(dx[nt][2], dx[nt][3]);
      }
      if (FUSED) {
        // publish the batch to the scatter warps.  The slot's previous phase (batch jb - NSLOT) must have been consumed before the
        // barrier may complete again; the data itself is never overwritten (every batch has its own rows of d_x0)
        const int slot = jb % NSLOT;
        if (jb >= NSLOT) mbar_wait(mbar_s + 8 * (NSLOT + slot), (uint32_t)(jb / NSLOT - 1) & 1u);
        mbar_arrive(mbar_s + 8 * slot);  // release at CTA scope: orders this thread's d_x0 stores before the scatter warps' loads
      }
    }
  }

  // ---- one flush per CTA ------------------------------------------------------------------------------------------------------
  const int in0 = a.in0;
  if (b.dWr3) flush_tiles<1, 8>(acc_w + A_R3 * 128, warp, lane, [&](int n, int k, float v) { if (n < 3) atomicAdd(b.dWr3 + n * 64 + k, v); });
  if (b.dWr2) flush_tiles<4, 8>(acc_w + A_R2 * 128, warp, lane, [&](int n, int k, float v) { atomicAdd(b.dWr2 + n * 64 + k, v); });
  if (b.dWr1) flush_tiles<4, 8>(acc_w + A_R1 * 128, warp, lane, [&](int n, int k, float v) { if (k != 16) atomicAdd(b.dWr1 + n * 63 + (k < 16 ? k : k - 1), v); });
  if (b.dWh) flush_tiles<1, 8>(acc_w + A_H * 128, warp, lane, [&](int n, int k, float v) { if (n == 0) atomicAdd(b.dWh + k, v); });
  if (b.dWs2) flush_tiles<4, 8>(acc_w + A_S2 * 128, warp, lane, [&](int n, int k, float v) { atomicAdd(b.dWs2 + n * 64 + k, v); });
  if (b.dWs1) flush_tiles<4, 2>(acc_w + A_S1 * 128, warp, lane, [&](int n, int k, float v) { if (k >= 1) atomicAdd(b.dWs1 + n * 15 + (k - 1), v); });
  if (b.dWb2) flush_tiles<1, 8>(acc_w + A_B2 * 128, warp, lane, [&](int n, int k, float v) { atomicAdd(b.dWb2 + n * 64 + k, v); });
  if (b.dWb1) flush_tiles<4, 4>(acc_w + A_B1 * 128, warp, lane, [&](int n, int k, float v) { if (k < in0) atomicAdd(b.dWb1 + n * in0 + k, v); });
  bar_mlp();
  for (int e = threadIdx.x; e < B_FLOATS; e += THREADS) {
    const float v = bias_acc[e];
    if (v == 0.f) continue;
    float* dst = nullptr;
    if (e < B_R2) { if (e - B_R3 < 3) dst = b.dbr3 ? b.dbr3 + (e - B_R3) : nullptr; }
    else if (e < B_R1) dst = b.dbr2 ? b.dbr2 + (e - B_R2) : nullptr;
    else if (e < B_H) dst = b.dbr1 ? b.dbr1 + (e - B_R1) : nullptr;
    else if (e < B_S2) { if (e == B_H) dst = b.dbh; }
    else if (e < B_S1) dst = b.dbs2 ? b.dbs2 + (e - B_S2) : nullptr;
    else if (e < B_B2) dst = b.dbs1 ? b.dbs1 + (e - B_S1) : nullptr;
    else if (e < B_B1) dst = b.dbb2 ? b.dbb2 + (e - B_B2) : nullptr;
    else dst = b.dbb1 ? b.dbb1 + (e - B_B1) : nullptr;
    if (dst) atomicAdd(dst, v);
  }
}

}  // namespace

int cnb_field_mixed_bwd_umma(const cnb_field* f, const cnb_samples* s, const float* d_density, const float* d_rgb, const float* d_sem, float* ctx,
                             cudaStream_t stream);  // field_mixed_bwd_umma.cu: tcgen05 / TMEM variant of the dW contraction

int cnb_field_mixed_bwd_tc5(const cnb_field* f, const cnb_samples* s, const float* d_density, const float* d_rgb, const float* d_sem, float* ctx,
                            cudaStream_t stream);   // field_mixed_bwd_tc5.cu: every layer a tcgen05.mma tile, row-parallel epilogues

int cnb_field_mixed_bwd(const cnb_field* f, const cnb_samples* s, const float* d_density, const float* d_rgb, const float* d_sem, float* ctx,
                        cudaStream_t stream) {
  // default: the all-tcgen05 kernel (field_mixed_bwd_tc5.cu).  CNB_FIELD_BWD=mma selects this file's mma.sync kernel, CNB_FIELD_BWD=umma the
  // mma.sync kernel with the dW contraction on tcgen05 (A/B measurements; same numerics contract, same tests).
  static const int mode = [] {
    const char* e = getenv("CNB_FIELD_BWD");
    if (e != nullptr && e[0] == 'm') return 1;
    if (e != nullptr && e[0] == 'u') return 2;
    const char* u = getenv("CNB_FIELD_BWD_UMMA");
    return (u != nullptr && u[0] == '1') ? 2 : 0;
  }();
  if (mode == 0) return cnb_field_mixed_bwd_tc5(f, s, d_density, d_rgb, d_sem, ctx, stream);
  if (mode == 2) return cnb_field_mixed_bwd_umma(f, s, d_density, d_rgb, d_sem, ctx, stream);
  BwdArgs b;
  fill_args(f, s, b.m);
  const int64_t N = s->num_rays * s->samples_per_ray;
  b.x0 = reinterpret_cast<const __half*>(ctx);
  b.pos = ctx + ctx_pos_off(N);
  b.d_x0 = ctx + ctx_dx0_off(N);
  b.d_density = d_density; b.d_rgb = d_rgb; b.d_sem = d_sem;
  b.dWb1 = f->base.dW[0]; b.dbb1 = f->base.db[0]; b.dWb2 = f->base.dW[1]; b.dbb2 = f->base.db[1];
  b.dWs1 = f->sem.dW[0]; b.dbs1 = f->sem.db[0]; b.dWs2 = f->sem.dW[1]; b.dbs2 = f->sem.db[1];
  b.dWh = f->sem_head.dW[0]; b.dbh = f->sem_head.db[0];
  b.dWr1 = f->rgb.dW[0]; b.dbr1 = f->rgb.db[0]; b.dWr2 = f->rgb.dW[1]; b.dbr2 = f->rgb.db[1]; b.dWr3 = f->rgb.dW[2]; b.dbr3 = f->rgb.db[2];
  b.d_embedding = f->appearance_mode == CNB_APP_PER_CAMERA ? f->d_embedding : nullptr;
  b.d_table = f->grid.d_table;
  // CNB_FIELD_BWD_FUSED=1: the table scatter on four extra warps of this kernel instead of a second kernel (same arithmetic).  Measured on
  // B200 (4096-ray step): 0.298 ms fused vs 0.270 ms as two kernels -- the reductions occupy the SM's LSU for ~1.6 cycles per lane-red
  // (tests/micro/table_access_bench: the red ceiling is an SM-side rate), and the MLP warps' shared-memory fragment loads queue behind
  // them, so the latency chain gets longer than the overlap saves.  Not the default.
  static const bool want_fused = [] { const char* e = getenv("CNB_FIELD_BWD_FUSED"); return e != nullptr && e[0] == '1'; }();
  bool fused = want_fused && b.d_table != nullptr;
  for (int i = 0; i < f->grid.num_levels && fused; ++i) fused = f->grid.scalings[i] < 65535.0f;  // cnb_scatter_cell's 16-bit cell keys
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(k_field_mixed_bwd<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BWD) != cudaSuccess ||
        cudaFuncSetAttribute(k_field_mixed_bwd<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BWD_FUSED) != cudaSuccess)
      return cnb_check_launch("field_mixed_bwd attr");
    configured = true;
  }
  const int64_t nbatches = (N + BATCH - 1) / BATCH;
  int64_t blocks = nbatches < (int64_t)cnb_num_sms() ? nbatches : (int64_t)cnb_num_sms();
  if (fused) {
    k_field_mixed_bwd<true><<<(int)blocks, THREADS_FUSED, SMEM_BWD_FUSED, stream>>>(b);
    return cnb_check_launch("field_mixed_bwd (fused scatter)");
  }
  k_field_mixed_bwd<false><<<(int)blocks, THREADS, SMEM_BWD, stream>>>(b);
  int rc = cnb_check_launch("field_mixed_bwd");
  if (rc) return rc;
  return cnb_hashgrid_bwd_level_major(&f->grid, b.pos, b.d_x0, N, stream);
}
