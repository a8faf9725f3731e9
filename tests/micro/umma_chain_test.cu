// Dev test for the all-tcgen05 field backward (csrc/field_mixed_bwd_tc5.cu): the three GEMM shapes of one MLP layer on ONE pair of shared
// buffers, with no-swizzle canonical layouts and TMEM accumulators.
//   activations X / dY : [128 samples][64 features], stored  byte(s, f) = (f/8)*2048 + s*16 + (f%8)*2   (one 16-byte chunk per sample and
//                        feature block: a thread that owns sample s writes 8 conflict-free 16-byte stores)
//   weights W          : [out][in] fp16, stored              byte(o, i) = (i/8)*(OUT*16) + o*16 + (i%8)*2
//   (1) forward   Y  = X  W^T : A = X  K-major (SBO 128, LBO 2048), B = W K-major  (SBO 128, LBO OUT*16),        M 128, N OUT, K IN
//   (2) dX        dX = dY W   : A = dY K-major,                     B = W MN-major (SBO OUT*16, LBO 128),         M 128, N IN,  K OUT
//   (3) dW        dW = dY^T X : A = dY MN-major (SBO 2048, LBO 128), B = X MN-major (same),                         M 64,  N IN,  K 128 samples
// Question 1: do these descriptors give the right numbers.  Question 2: does kind::f16 accept a bf16 A operand (dY) with an fp16 B operand
// (W, X) -- the instruction descriptor has separate a_format / b_format fields.  argv[1] = "mixed" (default) or "same" (everything bf16).
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

constexpr int S = 128, F = 64;

__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
// fmt: 0 = f16, 1 = bf16 ; major: 0 = K, 1 = MN
__device__ __forceinline__ uint32_t make_idesc(int M, int N, int afmt, int bfmt, int amajor, int bmajor) {
  uint32_t d = 0;
  d |= 1u << 4;
  d |= (uint32_t)afmt << 7;
  d |= (uint32_t)bfmt << 10;
  d |= (uint32_t)amajor << 15;
  d |= (uint32_t)bmajor << 16;
  d |= (uint32_t)(N >> 3) << 17;
  d |= (uint32_t)(M >> 4) << 24;
  return d;
}
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void ld64(uint32_t taddr, float* v) {
  uint32_t r[64];
#pragma unroll
  for (int c = 0; c < 64; c += 8)
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[c]), "=r"(r[c + 1]), "=r"(r[c + 2]), "=r"(r[c + 3]), "=r"(r[c + 4]), "=r"(r[c + 5]), "=r"(r[c + 6]), "=r"(r[c + 7])
                 : "r"(taddr + c));
  asm volatile("tcgen05.wait::ld.sync.aligned;");
#pragma unroll
  for (int c = 0; c < 64; ++c) v[c] = __uint_as_float(r[c]);
}

// X, dY: [S][F] row-major 16-bit raw ; W: [OUT][F] raw (OUT = 64 or 16)
__global__ void __launch_bounds__(128, 1) k_test(const uint16_t* __restrict__ X, const uint16_t* __restrict__ dY, const uint16_t* __restrict__ W, int OUT,
                                                 int dy_fmt, int xw_fmt, float* __restrict__ Y, float* __restrict__ dX, float* __restrict__ dW) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* sX = smem;                 // 16 KB
  unsigned char* sD = sX + S * F * 2;       // 16 KB (only OUT feature columns used)
  unsigned char* sW = sD + S * F * 2;       // OUT*F*2
  __shared__ __align__(8) uint64_t mbar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(&mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"((uint32_t)__cvta_generic_to_shared(&tmem_base_s)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  // stage: thread s owns sample row s
  {
    const int s = tid;
    for (int fb = 0; fb < F / 8; ++fb) *reinterpret_cast<uint4*>(sX + fb * 2048 + s * 16) = *reinterpret_cast<const uint4*>(X + s * F + fb * 8);
    for (int fb = 0; fb < OUT / 8; ++fb) *reinterpret_cast<uint4*>(sD + fb * 2048 + s * 16) = *reinterpret_cast<const uint4*>(dY + s * OUT + fb * 8);
    for (int e = tid; e < OUT * F / 8; e += 128) {
      const int o = e / (F / 8), ib = e % (F / 8);
      *reinterpret_cast<uint4*>(sW + ib * (OUT * 16) + o * 16) = *reinterpret_cast<const uint4*>(W + o * F + ib * 8);
    }
  }
  asm volatile("fence.proxy.async.shared::cta;");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = tmem_base_s;
  const uint32_t sX_s = (uint32_t)__cvta_generic_to_shared(sX), sD_s = (uint32_t)__cvta_generic_to_shared(sD), sW_s = (uint32_t)__cvta_generic_to_shared(sW);
  if (tid == 0) {
    // (1) forward: D cols [0, OUT)
    const uint32_t i1 = make_idesc(128, OUT, xw_fmt, xw_fmt, 0, 0);
    for (int ks = 0; ks < F / 16; ++ks)
      umma(tmem, make_desc(sX_s + ks * 2 * 2048, 2048, 128), make_desc(sW_s + ks * 2 * (OUT * 16), OUT * 16, 128), i1, ks > 0);
    // (2) dX: D cols [64, 128) : K = OUT
    const uint32_t i2 = make_idesc(128, F, dy_fmt, xw_fmt, 0, 1);
    for (int ks = 0; ks < OUT / 16; ++ks)
      umma(tmem + 64, make_desc(sD_s + ks * 2 * 2048, 2048, 128), make_desc(sW_s + ks * 2 * 16 * 8, 128, OUT * 16), i2, ks > 0);
    // (3) dW: M = OUT rows (M = 64; for OUT = 16 the upper rows of a 64-row tile read garbage columns -> only test OUT = 64 here), cols [128, 192)
    if (OUT == 64) {
      const uint32_t i3 = make_idesc(64, F, dy_fmt, xw_fmt, 1, 1);
      for (int ks = 0; ks < S / 16; ++ks)
        umma(tmem + 128, make_desc(sD_s + ks * 2 * 128, 128, 2048), make_desc(sX_s + ks * 2 * 128, 128, 2048), i3, ks > 0);
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.b64 [%0];" ::"l"((uint64_t)__cvta_generic_to_shared(&mbar)) : "memory");
  }
  {
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&mbar);
    uint32_t done = 0;
    while (!done)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(0u) : "memory");
  }
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t taddr = tmem + ((uint32_t)(32 * warp) << 16);
  float v[64];
  ld64(taddr, v);
  for (int c = 0; c < OUT; ++c) Y[tid * OUT + c] = v[c];
  ld64(taddr + 64, v);
  for (int c = 0; c < F; ++c) dX[tid * F + c] = v[c];
  if (OUT == 64) {
    ld64(taddr + 128, v);
    const int lane = tid & 31;
    if (lane < 16) for (int c = 0; c < F; ++c) dW[(16 * warp + lane) * F + c] = v[c];
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem));
}

static uint16_t enc(float v, int fmt) {
  if (fmt == 0) { __half h = __float2half(v); uint16_t r; memcpy(&r, &h, 2); return r; }
  __nv_bfloat16 h = __float2bfloat16(v); uint16_t r; memcpy(&r, &h, 2); return r;
}

static int run(int OUT, int dy_fmt, int xw_fmt) {
  static uint16_t hX[S * F], hD[S * F], hW[F * F];
  static float fX[S * F], fD[S * F], fW[F * F];
  srand(7 + OUT);
  for (int i = 0; i < S * F; ++i) { fX[i] = (rand() % 17 - 8) * 0.125f; hX[i] = enc(fX[i], xw_fmt); }
  for (int i = 0; i < S * OUT; ++i) { fD[i] = (rand() % 13 - 6) * 0.25f; hD[i] = enc(fD[i], dy_fmt); }
  for (int i = 0; i < OUT * F; ++i) { fW[i] = (rand() % 11 - 5) * 0.5f; hW[i] = enc(fW[i], xw_fmt); }
  uint16_t *dXi, *dDi, *dWi; float *oY, *oX, *oW;
  CK(cudaMalloc(&dXi, sizeof(hX))); CK(cudaMalloc(&dDi, sizeof(hD))); CK(cudaMalloc(&dWi, sizeof(hW)));
  CK(cudaMalloc(&oY, S * F * 4)); CK(cudaMalloc(&oX, S * F * 4)); CK(cudaMalloc(&oW, F * F * 4));
  CK(cudaMemcpy(dXi, hX, sizeof(hX), cudaMemcpyHostToDevice)); CK(cudaMemcpy(dDi, hD, sizeof(hD), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dWi, hW, sizeof(hW), cudaMemcpyHostToDevice));
  const size_t smem = (size_t)S * F * 2 * 2 + F * F * 2 + 1024;
  CK(cudaFuncSetAttribute(k_test, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_test<<<1, 128, smem>>>(dXi, dDi, dWi, OUT, dy_fmt, xw_fmt, oY, oX, oW);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  static float rY[S * F], rX[S * F], rW[F * F];
  CK(cudaMemcpy(rY, oY, S * OUT * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(rX, oX, S * F * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(rW, oW, F * F * 4, cudaMemcpyDeviceToHost));
  double e1 = 0, e2 = 0, e3 = 0;
  for (int s = 0; s < S; ++s) {
    for (int o = 0; o < OUT; ++o) { double r = 0; for (int i = 0; i < F; ++i) r += (double)fX[s * F + i] * fW[o * F + i]; e1 = fmax(e1, fabs(r - rY[s * OUT + o])); }
    for (int i = 0; i < F; ++i) { double r = 0; for (int o = 0; o < OUT; ++o) r += (double)fD[s * OUT + o] * fW[o * F + i]; e2 = fmax(e2, fabs(r - rX[s * F + i])); }
  }
  if (OUT == 64)
    for (int o = 0; o < OUT; ++o)
      for (int i = 0; i < F; ++i) { double r = 0; for (int s = 0; s < S; ++s) r += (double)fD[s * OUT + o] * fX[s * F + i]; e3 = fmax(e3, fabs(r - rW[o * F + i])); }
  printf("OUT %2d  dY %s  X/W %s :  forward max|err| %g   dX %g   dW %g  -> %s\n", OUT, dy_fmt ? "bf16" : "f16", xw_fmt ? "bf16" : "f16", e1, e2, e3,
         (e1 < 1e-3 && e2 < 1e-3 && e3 < 1e-3) ? "PASS" : "FAIL");
  return 0;
}

int main(int argc, char** argv) {
  const bool same = argc > 1 && strcmp(argv[1], "same") == 0;
  if (same) { run(64, 1, 1); run(16, 1, 1); }
  else { run(64, 1, 0); run(16, 1, 0); run(64, 0, 0); }
  return 0;
}
