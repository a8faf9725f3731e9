// Dev test: dW = dY^T X on tcgen05 (UMMA) with MN-major, no-swizzle shared-memory operands and a TMEM accumulator.
//   A = dY  [K = 128 samples][M = 64 features]   (MN-major: features contiguous)
//   B = X   [K = 128 samples][N = 64 features]   (MN-major)
//   D[m][n] = sum_k A[k][m] * B[k][n], accumulated over two "batches" to test the accumulate flag; plus a bias column
//   block: Dbias[m][0..7] = sum_k A[k][m] via an all-ones B operand with N = 8.
// Canonical no-swizzle MN-major layout (cute/atom/mma_traits_sm100.hpp): in 16-byte units ((1,n),(8,k)):((X,SBO),(1,LBO)):
//   byte(mn, k) = (mn/8)*SBO + (k/8)*LBO + (k%8)*16 + (mn%8)*2
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

constexpr int KS = 128;              // samples per batch
constexpr uint32_t LBO = 128;        // bytes between k-blocks of 8 samples
constexpr uint32_t SBO = 16 * 128;   // bytes between mn-blocks of 8 features

__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((LBO >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((SBO >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // version = 1 (Blackwell)
  return d;                // layout_type = 0 (no swizzle), base_offset = 0, lbo_mode = 0
}

__device__ __forceinline__ uint32_t make_idesc(int M, int N) {
  uint32_t d = 0;
  d |= 1u << 4;                    // c_format = F32
  d |= 1u << 7;                    // a_format = BF16
  d |= 1u << 10;                   // b_format = BF16
  d |= 1u << 15;                   // a_major = MN
  d |= 1u << 16;                   // b_major = MN
  d |= (uint32_t)(N >> 3) << 17;   // n_dim
  d |= (uint32_t)(M >> 4) << 24;   // m_dim
  return d;
}

__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}

__global__ void __launch_bounds__(128, 1) k_test(const __nv_bfloat16* __restrict__ A, const __nv_bfloat16* __restrict__ B, float* __restrict__ D,
                                                 float* __restrict__ Dbias, int batches) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __nv_bfloat16* sA = reinterpret_cast<__nv_bfloat16*>(smem);            // 64 x 128 -> 16 KB
  __nv_bfloat16* sB = sA + 64 * KS;                                      // 16 KB
  __nv_bfloat16* sOnes = sB + 64 * KS;                                   // 8 x 128 -> 2 KB
  __shared__ __align__(8) uint64_t mbar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int e = tid; e < 8 * KS; e += 128) sOnes[e] = __float2bfloat16(1.0f);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(&mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"((uint32_t)__cvta_generic_to_shared(&tmem_base_s)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = tmem_base_s;
  const uint32_t sA_s = (uint32_t)__cvta_generic_to_shared(sA), sB_s = (uint32_t)__cvta_generic_to_shared(sB),
                 sO_s = (uint32_t)__cvta_generic_to_shared(sOnes);
  uint32_t phase = 0;
  for (int b = 0; b < batches; ++b) {
    // stage this batch: global [k][f] row-major -> canonical MN-major layout
    for (int e = tid; e < KS * 64; e += 128) {
      const int k = e / 64, f = e % 64;
      const uint32_t off = (f / 8) * SBO + (k / 8) * LBO + (k % 8) * 16 + (f % 8) * 2;
      *reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<unsigned char*>(sA) + off) = A[(size_t)b * KS * 64 + e];
      *reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<unsigned char*>(sB) + off) = B[(size_t)b * KS * 64 + e];
    }
    asm volatile("fence.proxy.async.shared::cta;");
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;");
      const uint32_t idesc = make_idesc(64, 64), idesc_b = make_idesc(64, 8);
      for (int ks = 0; ks < KS / 16; ++ks) {
        const uint32_t koff = ks * 2 * LBO;
        umma(tmem, make_desc(sA_s + koff), make_desc(sB_s + koff), idesc, (b > 0 || ks > 0) ? 1u : 0u);
        umma(tmem + 64, make_desc(sA_s + koff), make_desc(sO_s + koff), idesc_b, (b > 0 || ks > 0) ? 1u : 0u);
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.b64 [%0];" ::"l"((uint64_t)__cvta_generic_to_shared(&mbar)) : "memory");
    }
    // everyone waits for the MMAs of this batch before the staging buffers are overwritten
    {
      const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&mbar);
      uint32_t done = 0;
      while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(phase) : "memory");
      }
      phase ^= 1;
    }
    __syncthreads();
  }
  asm volatile("tcgen05.fence::after_thread_sync;");
  // read back: M = 64 -> row m lives in TMEM lane (m % 16) + 32 * (m / 16); warp w reads lanes 32w..32w+31
  {
    const int lane = tid & 31;
    const uint32_t taddr = tmem + ((uint32_t)(32 * warp) << 16);
    uint32_t v[64];
#pragma unroll
    for (int c = 0; c < 64; c += 8) {
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                   : "=r"(v[c]), "=r"(v[c + 1]), "=r"(v[c + 2]), "=r"(v[c + 3]), "=r"(v[c + 4]), "=r"(v[c + 5]), "=r"(v[c + 6]), "=r"(v[c + 7])
                   : "r"(taddr + c));
    }
    uint32_t vb[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(vb[0]), "=r"(vb[1]), "=r"(vb[2]), "=r"(vb[3]), "=r"(vb[4]), "=r"(vb[5]), "=r"(vb[6]), "=r"(vb[7])
                 : "r"(taddr + 64));
    asm volatile("tcgen05.wait::ld.sync.aligned;");
    if (lane < 16) {
      const int m = 16 * warp + lane;
      for (int c = 0; c < 64; ++c) D[m * 64 + c] = __uint_as_float(v[c]);
      for (int c = 0; c < 8; ++c) Dbias[m * 8 + c] = __uint_as_float(vb[c]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem));
}

int main() {
  const int batches = 2;
  const size_t n = (size_t)batches * KS * 64;
  __nv_bfloat16 *hA = new __nv_bfloat16[n], *hB = new __nv_bfloat16[n];
  srand(1);
  for (size_t i = 0; i < n; ++i) {
    hA[i] = __float2bfloat16((rand() % 17 - 8) * 0.125f);
    hB[i] = __float2bfloat16((rand() % 13 - 6) * 0.25f);
  }
  __nv_bfloat16 *dA, *dB; float *dD, *dDb;
  CK(cudaMalloc(&dA, n * 2)); CK(cudaMalloc(&dB, n * 2)); CK(cudaMalloc(&dD, 64 * 64 * 4)); CK(cudaMalloc(&dDb, 64 * 8 * 4));
  CK(cudaMemcpy(dA, hA, n * 2, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, hB, n * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(dD, 0, 64 * 64 * 4)); CK(cudaMemset(dDb, 0, 64 * 8 * 4));
  const size_t smem = (size_t)(64 * KS * 2 + 8 * KS) * 2 + 1024;
  CK(cudaFuncSetAttribute(k_test, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_test<<<1, 128, smem>>>(dA, dB, dD, dDb, batches);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  float hD[64 * 64], hDb[64 * 8];
  CK(cudaMemcpy(hD, dD, sizeof(hD), cudaMemcpyDeviceToHost)); CK(cudaMemcpy(hDb, dDb, sizeof(hDb), cudaMemcpyDeviceToHost));
  double maxerr = 0, maxerr_b = 0;
  int bad = 0;
  for (int m = 0; m < 64; ++m) {
    double bs = 0;
    for (size_t k = 0; k < (size_t)batches * KS; ++k) bs += __bfloat162float(hA[k * 64 + m]);
    for (int c = 0; c < 8; ++c) maxerr_b = fmax(maxerr_b, fabs(hDb[m * 8 + c] - bs));
    for (int nn = 0; nn < 64; ++nn) {
      double ref = 0;
      for (size_t k = 0; k < (size_t)batches * KS; ++k) ref += (double)__bfloat162float(hA[k * 64 + m]) * (double)__bfloat162float(hB[k * 64 + nn]);
      const double err = fabs(hD[m * 64 + nn] - ref);
      if (err > 1e-3 && bad < 5) { printf("mismatch D[%d][%d] = %f ref %f\n", m, nn, hD[m * 64 + nn], ref); ++bad; }
      maxerr = fmax(maxerr, err);
    }
  }
  printf("umma dW test: max |err| = %g, bias max |err| = %g -> %s\n", maxerr, maxerr_b, (maxerr < 1e-3 && maxerr_b < 1e-3) ? "PASS" : "FAIL");
  return 0;
}
