// FruitField backward, tcgen05 variant (rows a1/a5/a6 of SURVEY.md section 8, the autograd of fruit_field.py:169-302).
//
// ALTERNATIVE to field_mixed_bwd.cu, selected with CNB_FIELD_BWD_UMMA=1.  Same numerics contract and tests.  Measured on B200
// (4096-ray training step): 0.325 ms for the field backward stage versus 0.285 ms for the pure mma.sync kernel, so it is NOT
// the default: the dW contraction it moves to tcgen05/TMEM was only ~14 % of the instructions, the kernel is bound by the
// register-chained mma.sync forward-recompute / dX work, and the asynchronous UTCHMMA stream competes with those HMMAs for
// the same tensor pipe (math-pipe-throttle stalls 0.38 -> 0.86 per issue).  Kept as the working tcgen05 reference for this
// path (tests/micro/umma_dw_test.cu is the stand-alone descriptor / TMEM-layout test it grew from).
//
// One kernel does, per 128-sample batch of a persistent CTA (8 warps, one 16-sample m-tile each):
//   1. re-run the forward MLPs from the fp16 encoded features the forward kept (no second table gather); the fp16
//      activations stay in registers as mma A fragments, exactly as in field_mixed.cu, so the ReLU masks are the
//      forward's;
//   2. input gradients layer by layer: dX = dY * W with dY re-packed from the accumulator fragments (bf16 -- the range of
//      fp32, so no loss scaling is needed for the 1e-7-sized pixel gradients) and W^T read from a transposed bf16 copy
//      of the weights in shared memory;
//   3. weight gradients: dW = dY^T * X is a contraction over SAMPLES (K = the 128 samples of the batch) whose operands are
//      shared-memory resident -- the shape tcgen05 is made for.  Each layer's (dY, X) pair is staged once in the canonical
//      no-swizzle MN-major UMMA layout, ONE thread issues 8 tcgen05.mma (K = 16 each) per layer, and the eight dW
//      accumulators (328 of the 512 TMEM columns) live in TENSOR MEMORY across all batches of the persistent CTA: no
//      ldmatrix, no per-warp accumulator slices, no atomics until the one read-out (tcgen05.ld) + flush per CTA.  The MMAs run
//      asynchronously behind the warps' mma.sync work of the next layer; a tcgen05.commit -> mbarrier per staging buffer
//      (4 buffers) tells the warps when a buffer may be overwritten.  Bias gradients come from the same MMAs: a block of
//      ones sits right behind every staged X matrix, so it is one more 8-column block of the B operand (or 8 more rows of
//      the A operand for the three narrow layers, which are issued transposed with M = 128 so that N can stay 16);
//   4. d(encoded features), level-major [16][N][2] fp32, goes to the scratch the hash-grid scatter (cnb_hashgrid_bwd_level_major) reads; the appearance
//      embedding gradient is reduced over the m-tile (all 16 samples share the ray when S % 16 == 0) before its atomics.
#include "field_mixed.cuh"

using namespace cnbmix;

namespace {

constexpr int BATCH = WARPS * 16;
// transposed bf16 weights, [in][out] with padded rows (half offsets)
constexpr int T_R3 = 0;                   // [64][24]   out 0..2 of 16 used
constexpr int T_R2 = T_R3 + 64 * 24;      // [64][72]
constexpr int T_R1 = T_R2 + 64 * 72;      // [64][72]   rows = padded rgb input [SH16 | 0,geo15 | emb32]
constexpr int T_S2 = T_R1 + 64 * 72;      // [64][72]
constexpr int T_B2 = T_S2 + 64 * 72;      // [64][24]
constexpr int T_B1 = T_B2 + 64 * 24;      // [32][72]
constexpr int T_HALVES = T_B1 + 32 * 72;
// ---- tcgen05 (UMMA) dW contraction: staging buffers in the canonical no-swizzle MN-major layout -----------------------------------
//   byte(feature f, sample k) = (f/8)*SBO + (k/8)*LBO + (k%8)*16 + (f%8)*2      (cute/atom/mma_traits_sm100.hpp, Major::MN, INTERLEAVE)
constexpr uint32_t U_LBO = 128;                  // between 8-sample k-blocks
constexpr uint32_t U_SBO = (BATCH / 8) * 128;    // between 8-feature mn-blocks (2048)
constexpr uint32_t U_MAT = 8 * U_SBO;            // one 64-feature x 128-sample matrix (16 KB)
constexpr uint32_t U_BUF = 2 * U_MAT + U_SBO;    // [dY | X | ones block]
constexpr int NBUF = 4;
// 8 worker warps (warpgroups 0, 1) + one more warpgroup whose first warp is the MMA issuer.  A 9-warp CTA would cap every thread at 168
// registers (three warps on one SM sub-partition) and spill the workers' fragment chains; with a full third warpgroup the registers are
// re-split after launch (setmaxnreg: workers 224, issuer warpgroup 56; 3 x 168 = 224 + 224 + 56 per sub-partition lane).
constexpr int BTHREADS = THREADS + 128;
// GEMM table: swapped = issued as (dW)^T = [X | ones]^T dY with M = 128 (narrow dY, N = 16); otherwise dW = dY^T [X | ones], M = 64.
// xblk0 = first 8-feature block of the X region that holds data (narrow X is right-aligned so that the ones block follows it).
struct GemmSpec { bool swapped; int xblk0; int n; int col; };
enum { G_R3 = 0, G_R2, G_R1, G_H, G_S2, G_S1, G_B2, G_B1, NGEMM };
__device__ constexpr GemmSpec GEMMS[NGEMM] = {
    {true, 0, 16, 0},     // dWr3^T [64(+bias row) x 3]
    {false, 0, 72, 16},   // dWr2   [64 x 64 | bias]
    {false, 0, 72, 88},   // dWr1   [64 x 64 | bias]   (column 16 = the cleared dba slot)
    {true, 0, 16, 160},   // dWh^T  [64(+bias row) x 1]
    {false, 0, 72, 176},  // dWs2
    {false, 6, 24, 248},  // dWs1   [64 x 16 | bias]   (X = [0 | geo15] in blocks 6,7)
    {true, 0, 16, 272},   // dWb2^T [64(+bias row) x 16]
    {false, 4, 40, 288},  // dWb1   [64 x 32 | bias]   (X = encoded features in blocks 4..7)
};
constexpr int TMEM_COLS_USED = 328;

constexpr size_t SMEM_BWD = (size_t)NBUF * U_BUF + (size_t)(HALVES + T_HALVES) * 2 + (size_t)FLOATS * 4 + 128;

struct BwdArgs {
  MixArgs m;
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

// write KT k-tiles (16 features each) of an A fragment into a staged matrix (canonical UMMA layout), features starting at
// 8-feature block blk0; this warp's 16 samples are k-blocks 2*warp, 2*warp+1.  Conflict-free 32-bit stores (g*16 + t*4).
template <int KT, bool CVT>
__device__ __forceinline__ void stage_u(unsigned char* mat, int blk0, int warp, const uint32_t (&A)[KT][4], int g, int t) {
  unsigned char* p = mat + (uint32_t)blk0 * U_SBO + (uint32_t)(2 * warp) * U_LBO + g * 16 + t * 4;
#pragma unroll
  for (int kt = 0; kt < KT; ++kt)
#pragma unroll
    for (int q = 0; q < 2; ++q)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const uint32_t v = A[kt][2 * q + h];
        *reinterpret_cast<uint32_t*>(p + (uint32_t)(2 * kt + q) * U_SBO + (uint32_t)h * U_LBO) = CVT ? h2_to_bf2(v) : v;
      }
}

__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((uint64_t)(U_LBO >> 4) << 16) | ((uint64_t)(U_SBO >> 4) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ uint32_t umma_idesc(int M, int N) {  // bf16 x bf16 -> f32, both operands MN-major
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_issue(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d), "l"(da),
               "l"(db), "r"(idesc), "r"(accumulate)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
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

__global__ void __launch_bounds__(BTHREADS, 1) k_field_mixed_bwd_umma(const __grid_constant__ BwdArgs b) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const MixArgs& a = b.m;
  // staging buffers first: the M = 128 operand of the transposed GEMMs reads 7 junk feature blocks past its ones block, which
  // must stay inside the allocation (they land in the next buffer / the weights and only feed accumulator rows nobody reads)
  unsigned char* stg = smem_raw;
  __half* Wsm = reinterpret_cast<__half*>(smem_raw + NBUF * U_BUF);
  __nv_bfloat16* WT = reinterpret_cast<__nv_bfloat16*>(Wsm + HALVES);
  float* Bf = reinterpret_cast<float*>(WT + T_HALVES);
  uint64_t* mbar = reinterpret_cast<uint64_t*>(Bf + FLOATS);   // NBUF "empty" (MMAs drained the buffer) + NBUF "full" (all 8 warps staged)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar + 2 * NBUF);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  load_weights(a, Wsm, Bf);
  load_weights_t(a, WT);
  for (int bb = 0; bb < NBUF; ++bb) {  // the ones block behind every X matrix (bias gradients), written once
    uint32_t* ones = reinterpret_cast<uint32_t*>(stg + bb * U_BUF + 2 * U_MAT);
    for (int e = threadIdx.x; e < (int)(U_SBO / 4); e += THREADS) ones[e] = 0x3F803F80u;  // bf16 (1, 1)
  }
  if (threadIdx.x == 0) {
    for (int bb = 0; bb < NBUF; ++bb) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(mbar + bb)));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(mbar + NBUF + bb)), "r"(WARPS));
    }
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (warp == 0) {  // all 512 TMEM columns: the CTA owns the SM (217 KB of shared memory)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(tmem_slot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = *tmem_slot;
  const uint32_t stg_s = (uint32_t)__cvta_generic_to_shared(stg), mbar_s = (uint32_t)__cvta_generic_to_shared(mbar);
  uint32_t buf = 0, used = 0, phase = 0;   // staging buffer ring + per-buffer "empty" phase bits (same sequence in every worker warp)
  // Workers never meet at a block barrier inside the batch loop.  acquire(): wait until the MMAs of the buffer's previous use
  // have drained it (tcgen05.commit -> "empty" mbarrier).  publish(): make this warp's rows visible to the async proxy and
  // arrive on the buffer's "full" mbarrier (8 arrivals); the issuer warp below turns a full buffer into 8 tcgen05.mma.
  auto acquire = [&]() -> unsigned char* {
    if (used & (1u << buf)) { mbar_wait(mbar_s + 8 * buf, (phase >> buf) & 1u); phase ^= 1u << buf; }
    used |= 1u << buf;
    return stg + buf * U_BUF;
  };
  auto launch = [&](int) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(mbar_s + 8 * (NBUF + buf)) : "memory");
    buf = (buf + 1) % NBUF;
  };
  const int S = a.sm.samples_per_ray;
  const int64_t N = a.sm.num_rays * S;
  const int64_t nbatches = (N + BATCH - 1) / BATCH;

  if (warp >= WARPS) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    if (warp > WARPS) return;  // the rest of the third warpgroup only donates its registers
    // ===== MMA issuer warp: one lane issues the 8 K-steps of every layer's dW GEMM as soon as its buffer is full =====
    uint32_t fphase = 0, ibuf = 0;
    bool first_batch = true;
    for (int64_t batch = blockIdx.x; batch < nbatches; batch += gridDim.x) {
      for (int gi = 0; gi < NGEMM; ++gi) {
        mbar_wait(mbar_s + 8 * (NBUF + ibuf), (fphase >> ibuf) & 1u);
        fphase ^= 1u << ibuf;
        if (lane == 0) {
          asm volatile("tcgen05.fence::after_thread_sync;");
          const GemmSpec sp = GEMMS[gi];
          const uint32_t dy_s = stg_s + ibuf * U_BUF, x_s = dy_s + U_MAT;
          const uint32_t a_s = sp.swapped ? x_s : dy_s;
          const uint32_t b_s = sp.swapped ? dy_s : x_s + (uint32_t)sp.xblk0 * U_SBO;
          const uint32_t idesc = umma_idesc(sp.swapped ? 128 : 64, sp.n);
#pragma unroll
          for (int ks = 0; ks < BATCH / 16; ++ks)
            umma_issue(tmem + sp.col, umma_desc(a_s + ks * 2 * U_LBO), umma_desc(b_s + ks * 2 * U_LBO), idesc, (first_batch && ks == 0) ? 0u : 1u);
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.b64 [%0];" ::"l"((uint64_t)(mbar_s + 8 * ibuf)) : "memory");
        }
        __syncwarp();
        ibuf = (ibuf + 1) % NBUF;
      }
      first_batch = false;
    }
    return;
  }
  asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");

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

  for (int64_t batch = blockIdx.x; batch < nbatches; batch += gridDim.x) {
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
      {
        unsigned char* sb = acquire();
        stage_u<1, false>(sb, 0, warp, D3, g, t);
        stage_u<4, true>(sb + U_MAT, 0, warp, AR2, g, t);
        launch(G_R3);
      }
      uint32_t D[4][4];
      zero_acc<8>(acc);
      layer_bf<8, 1, 24>(WT + T_R3, D3, acc, g, t);
      relu_mask_pack<4>(acc, AR2, D);
      // ---- layer 2 ----------------------------------------------------------------------------------------------------------
      {
        unsigned char* sb = acquire();
        stage_u<4, false>(sb, 0, warp, D, g, t);
        stage_u<4, true>(sb + U_MAT, 0, warp, AR1, g, t);
        launch(G_R2);
      }
      zero_acc<8>(acc);
      layer_bf<8, 4, 72>(WT + T_R2, D, acc, g, t);
      relu_mask_pack<4>(acc, AR1, D);
      // ---- layer 1 ----------------------------------------------------------------------------------------------------------
      {
        unsigned char* sb = acquire();
        stage_u<4, false>(sb, 0, warp, D, g, t);
        stage_u<4, true>(sb + U_MAT, 0, warp, Ain, g, t);
        launch(G_R1);
      }
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
      {
        unsigned char* sb = acquire();
        stage_u<1, false>(sb, 0, warp, Dh, g, t);
        stage_u<4, false>(sb + U_MAT, 0, warp, S2f, g, t);
        launch(G_H);
      }
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
      {
        unsigned char* sb = acquire();
        stage_u<4, false>(sb, 0, warp, D, g, t);
        stage_u<4, true>(sb + U_MAT, 0, warp, AS1, g, t);
        launch(G_S2);
      }
      zero_acc<8>(acc);
      layer_bf<8, 4, 72>(WT + T_S2, D, acc, g, t);
      relu_mask_pack<4>(acc, AS1, D);
      // ---- semantic layer 1 (input = [0 | geo15]) -----------------------------------------------------------------------------
      {
        unsigned char* sb = acquire();
        stage_u<4, false>(sb, 0, warp, D, g, t);
        stage_u<1, true>(sb + U_MAT, 6, warp, Abo, g, t);
        launch(G_S1);
      }
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
      {
        unsigned char* sb = acquire();
        stage_u<1, false>(sb, 0, warp, Dbo, g, t);
        stage_u<4, true>(sb + U_MAT, 0, warp, AH, g, t);
        launch(G_B2);
      }
      float acc[8][4];
      zero_acc<8>(acc);
      layer_bf<8, 1, 24>(WT + T_B2, Dbo, acc, g, t);
      uint32_t D[4][4];
      relu_mask_pack<4>(acc, AH, D);
      {
        unsigned char* sb = acquire();
        stage_u<4, false>(sb, 0, warp, D, g, t);
        stage_u<2, true>(sb + U_MAT, 4, warp, A0, g, t);
        launch(G_B1);
      }
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
    }
  }

  // ---- one read-out + flush per CTA: wait for the last MMAs, tcgen05.ld the accumulators, reduce into the global gradients ------
  for (int bb = 0; bb < NBUF; ++bb)
    if (used & (1u << bb)) mbar_wait(mbar_s + 8 * bb, (phase >> bb) & 1u);
  asm volatile("bar.sync 1, %0;" ::"n"(THREADS) : "memory");  // the 8 worker warps (the issuer warpgroup has left)
  asm volatile("tcgen05.fence::after_thread_sync;");
  // Gradient image in shared memory (the staging buffers are free now): every tensor in its GLOBAL element order, so the
  // flush below is coalesced; straight from the TMEM read-out a warp instruction would scatter over 16-32 rows, and all
  // 148 CTAs would hit the same addresses in the same order.
  const int in0 = a.in0;
  float* img = reinterpret_cast<float*>(stg);
  constexpr int I_WR3 = 0, I_BR3 = 192, I_WR2 = 196, I_BR2 = I_WR2 + 4096, I_WR1 = I_BR2 + 64, I_BR1 = I_WR1 + 4032, I_WH = I_BR1 + 64, I_BH = I_WH + 64,
                I_WS2 = I_BH + 4, I_BS2 = I_WS2 + 4096, I_WS1 = I_BS2 + 64, I_BS1 = I_WS1 + 960, I_WB2 = I_BS1 + 64, I_BB2 = I_WB2 + 1024,
                I_WB1 = I_BB2 + 16, I_BB1 = I_WB1 + 2048, I_END = I_BB1 + 64;
  static_assert(I_END * 4 <= NBUF * U_BUF, "gradient image must fit the staging buffers");
  if (warp < 4) {
    // TMEM lane of this thread: 32*warp + lane.  M = 64 accumulators keep row m in lane (m%16) + 32*(m/16) -> lanes 0..15 of
    // each warp hold rows 16*warp + lane; M = 128 accumulators (transposed GEMMs) keep row r in lane r.
    const int m64 = 16 * warp + lane, r128 = 32 * warp + lane;
    const bool has64 = lane < 16;
    const uint32_t taddr = tmem + ((uint32_t)(32 * warp) << 16);
    for (int c0 = 0; c0 < TMEM_COLS_USED; c0 += 8) {
      uint32_t v[8];
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                   : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                   : "r"(taddr + c0));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int col = c0 + j;
        const float val = __uint_as_float(v[j]);
        int dst = -1;
        if (col < 16) {                                   // dWr3^T (M = 128): row = input feature / bias row 64, column = output
          if (col < 3) { if (r128 < 64) dst = I_WR3 + col * 64 + r128; else if (r128 == 64) dst = I_BR3 + col; }
        } else if (col < 88) {                            // dWr2 | bias
          const int k = col - 16;
          if (has64) dst = k < 64 ? I_WR2 + m64 * 64 + k : (k == 64 ? I_BR2 + m64 : -1);
        } else if (col < 160) {                           // dWr1 | bias (padded input column 16 = dba slot)
          const int k = col - 88;
          if (has64) {
            if (k < 64) { if (k != 16) dst = I_WR1 + m64 * 63 + (k < 16 ? k : k - 1); }
            else if (k == 64) dst = I_BR1 + m64;
          }
        } else if (col < 176) {                           // dWh^T (M = 128)
          if (col == 160) { if (r128 < 64) dst = I_WH + r128; else if (r128 == 64) dst = I_BH; }
        } else if (col < 248) {                           // dWs2 | bias
          const int k = col - 176;
          if (has64) dst = k < 64 ? I_WS2 + m64 * 64 + k : (k == 64 ? I_BS2 + m64 : -1);
        } else if (col < 272) {                           // dWs1 | bias (input column 0 = dba slot)
          const int k = col - 248;
          if (has64) {
            if (k >= 1 && k < 16) dst = I_WS1 + m64 * 15 + (k - 1);
            else if (k == 16) dst = I_BS1 + m64;
          }
        } else if (col < 288) {                           // dWb2^T (M = 128)
          const int n = col - 272;
          if (r128 < 64) dst = I_WB2 + n * 64 + r128; else if (r128 == 64) dst = I_BB2 + n;
        } else {                                          // dWb1 | bias
          const int k = col - 288;
          if (has64) {
            if (k < 32) { if (k < in0) dst = I_WB1 + m64 * in0 + k; }
            else if (k == 32) dst = I_BB1 + m64;
          }
        }
        if (dst >= 0) img[dst] = val;
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  asm volatile("bar.sync 1, %0;" ::"n"(THREADS) : "memory");
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
  // coalesced flush, CTAs start at staggered offsets so they do not all hammer the same addresses at the same moment
  {
    auto flush = [&](float* g, int off, int n) {
      if (g == nullptr) return;
      const int rot = (int)((blockIdx.x * 97u) % (unsigned)n);
      for (int e = threadIdx.x; e < n; e += THREADS) {
        int idx = e + rot;
        if (idx >= n) idx -= n;
        const float v = img[off + idx];
        if (v != 0.f) atomicAdd(g + idx, v);
      }
    };
    flush(b.dWr3, I_WR3, 192); flush(b.dbr3, I_BR3, 3);
    flush(b.dWr2, I_WR2, 4096); flush(b.dbr2, I_BR2, 64);
    flush(b.dWr1, I_WR1, 4032); flush(b.dbr1, I_BR1, 64);
    flush(b.dWh, I_WH, 64); flush(b.dbh, I_BH, 1);
    flush(b.dWs2, I_WS2, 4096); flush(b.dbs2, I_BS2, 64);
    flush(b.dWs1, I_WS1, 960); flush(b.dbs1, I_BS1, 64);
    flush(b.dWb2, I_WB2, 1024); flush(b.dbb2, I_BB2, 16);
    flush(b.dWb1, I_WB1, 64 * in0); flush(b.dbb1, I_BB1, 64);
  }
}

}  // namespace

int cnb_field_mixed_bwd_umma(const cnb_field* f, const cnb_samples* s, const float* d_density, const float* d_rgb, const float* d_sem, float* ctx,
                        cudaStream_t stream) {
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
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(k_field_mixed_bwd_umma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BWD) != cudaSuccess) return cnb_check_launch("field_mixed_bwd_umma attr");
    configured = true;
  }
  const int64_t nbatches = (N + BATCH - 1) / BATCH;
  int64_t blocks = nbatches < (int64_t)cnb_num_sms() ? nbatches : (int64_t)cnb_num_sms();
  k_field_mixed_bwd_umma<<<(int)blocks, BTHREADS, SMEM_BWD, stream>>>(b);
  int rc = cnb_check_launch("field_mixed_bwd_umma");
  if (rc) return rc;
  return cnb_hashgrid_bwd_level_major(&f->grid, b.pos, b.d_x0, N, stream);
}
