// Layer-by-layer path of the fp32 nn.Linear stack for layers wider than CNB_MAX_WIDTH (rows a3/a4 of SURVEY.md section 8).
//
// The fused MLP operators (mlp.cu) keep every layer's weights in shared memory and are compiled for widths <= 64, which is what the
// fruit_nerf preset forwards to FruitField.  The reference's other two presets forward wider layers (fruit_nerf_config.py:86-98,141-150:
// hidden_dim_semantics = 128, num_layers_semantic = 3, geo_feat_dim = 30, i.e. a 78-wide RGB input): those run here, one tiled GEMM per
// layer, activations through global memory -- plain FFMA, fp32, the exact-mode contract (1e-5 of the output scale against torch).  This
// is the functional path for those presets, not a tuned one: nothing on it is benchmarked.
//   forward  : Y_l = act(X_l W_l^T + b_l)                      k_wide_gemm<true>
//   backward : dZ_L = dY * act'(Y)                             k_wide_act_grad
//              dW_l += dZ_l^T X_l , db_l += colsum(dZ_l)       k_wide_dw  (row range split over blockIdx.z, atomics)
//              dX_l = (dZ_l W_l) * [X_l > 0]                    k_wide_gemm<false> with the ReLU mask of the producing layer
// Rows are processed in chunks of WIDE_CHUNK so that the two dZ ping-pong buffers are a fixed 2 x 32 MiB per device.
#include "cnb_common.cuh"

namespace {

constexpr int WIDE_MAX = CNB_WIDE_MAX_WIDTH;
constexpr int64_t WIDE_CHUNK = 32768;
constexpr int TM = 64, TN = 64, TK = 16;

__device__ __forceinline__ float wide_act(float v, int act) {
  if (act == CNB_ACT_RELU) return fmaxf(v, 0.0f);
  if (act == CNB_ACT_SIGMOID) return 1.0f / (1.0f + expf(-v));
  return v;
}

// C[m][n] = epilogue(sum_k A[m][k] * Bop[k][n]);  B_TRANS: Bop[k][n] = B[n * ldb + k] (nn.Linear weight [out][in]), else B[k * ldb + n]
template <bool B_TRANS>
__global__ void __launch_bounds__(256) k_wide_gemm(const float* __restrict__ A, int64_t lda, const float* __restrict__ B, int ldb, float* __restrict__ C,
                                                   int64_t ldc, int64_t M, int N, int K, const float* __restrict__ bias, int act,
                                                   const float* __restrict__ relu_mask, int64_t ldmask) {
  __shared__ float As[TK][TM + 4];
  __shared__ float Bs[TK][TN + 4];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int64_t m0 = (int64_t)blockIdx.y * TM;
  const int n0 = blockIdx.x * TN;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += TK) {
    for (int e = threadIdx.x; e < TM * TK; e += 256) {
      const int kk = e % TK, i = e / TK;
      const int64_t m = m0 + i;
      As[kk][i] = (m < M && k0 + kk < K) ? __ldg(A + m * lda + k0 + kk) : 0.0f;
    }
    for (int e = threadIdx.x; e < TN * TK; e += 256) {
      int kk, j;
      if (B_TRANS) { kk = e % TK; j = e / TK; } else { j = e % TN; kk = e / TN; }
      const int n = n0 + j, k = k0 + kk;
      Bs[kk][j] = (n < N && k < K) ? __ldg(B_TRANS ? B + (int64_t)n * ldb + k : B + (int64_t)k * ldb + n) : 0.0f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < TK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = acc[i][j];
      if (bias) v += __ldg(bias + n);
      v = wide_act(v, act);
      if (relu_mask && !(__ldg(relu_mask + m * ldmask + n) > 0.0f)) v = 0.0f;
      C[m * ldc + n] = v;
    }
  }
}

// dW[i][j] += sum_m dZ[m][i] * X[m][j] over this block's row range; db[i] += sum_m dZ[m][i] (blocks of the first input tile)
__global__ void __launch_bounds__(256) k_wide_dw(const float* __restrict__ dZ, int64_t ldz, const float* __restrict__ X, int64_t ldx, int64_t M, int OUT, int IN,
                                                 int64_t rows_per_block, float* __restrict__ dW, float* __restrict__ db) {
  __shared__ float Zs[TK][TM + 4];
  __shared__ float Xs[TK][TN + 4];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int i0 = blockIdx.y * TM, j0 = blockIdx.x * TN;
  const int64_t r0 = (int64_t)blockIdx.z * rows_per_block, r1 = min(M, r0 + rows_per_block);
  float acc[4][4] = {};
  float bsum[4] = {0.f, 0.f, 0.f, 0.f};
  for (int64_t m0 = r0; m0 < r1; m0 += TK) {
    for (int e = threadIdx.x; e < TM * TK; e += 256) {
      const int i = e % TM, kk = e / TM;
      const int64_t m = m0 + kk;
      Zs[kk][i] = (m < r1 && i0 + i < OUT) ? __ldg(dZ + m * ldz + i0 + i) : 0.0f;
    }
    for (int e = threadIdx.x; e < TN * TK; e += 256) {
      const int j = e % TN, kk = e / TN;
      const int64_t m = m0 + kk;
      Xs[kk][j] = (m < r1 && j0 + j < IN) ? __ldg(X + m * ldx + j0 + j) : 0.0f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < TK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = Zs[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Xs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        bsum[i] += a[i];
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int o = i0 + ty * 4 + i;
    if (o >= OUT) continue;
    if (dW != nullptr) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = j0 + tx * 4 + j;
        if (c < IN && acc[i][j] != 0.0f) atomicAdd(dW + (int64_t)o * IN + c, acc[i][j]);
      }
    }
    if (db != nullptr && blockIdx.x == 0 && tx == 0 && bsum[i] != 0.0f) atomicAdd(db + o, bsum[i]);
  }
}

// dZ = dY * act'(Y)
__global__ void __launch_bounds__(256) k_wide_act_grad(const float* __restrict__ dy, const float* __restrict__ y, int64_t count, int act, float* __restrict__ dz) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
    float v = __ldg(dy + i);
    if (act == CNB_ACT_SIGMOID) { const float yy = __ldg(y + i); v *= yy * (1.0f - yy); }
    else if (act == CNB_ACT_RELU) { if (!(__ldg(y + i) > 0.0f)) v = 0.0f; }
    dz[i] = v;
  }
}

// two ping-pong buffers of WIDE_CHUNK x WIDE_MAX floats per device, allocated on first use (never while a stream is capturing: the engine's
// eager warm-up step comes first)
float* g_scratch[16] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};

int wide_scratch(float*& a, float*& b) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) { cnb_set_error("mlp_wide: no CUDA device"); return CNB_ERR_CUDA; }
  if (g_scratch[dev] == nullptr) {
    if (cudaMalloc(&g_scratch[dev], sizeof(float) * 2 * (size_t)WIDE_CHUNK * WIDE_MAX) != cudaSuccess) {
      cudaGetLastError();
      cnb_set_error("mlp_wide: cannot allocate the 64 MiB layer scratch (first call must not be inside a stream capture)");
      return CNB_ERR_CUDA;
    }
  }
  a = g_scratch[dev];
  b = a + WIDE_CHUNK * WIDE_MAX;
  return CNB_OK;
}

inline dim3 gemm_grid(int64_t M, int N) { return dim3((unsigned)((N + TN - 1) / TN), (unsigned)((M + TM - 1) / TM), 1); }

int64_t hidden_off(const cnb_mlp* m, int64_t n, int layer) {  // offset of layer `layer`'s output in the hidden array (layer-major, as mlp.cu)
  int64_t off = 0;
  for (int q = 0; q < layer; ++q) off += n * (int64_t)m->dims[q + 1];
  return off;
}

// ---- single Linear(in -> 1): the semantic head (components/field_heads.py:29-40).  As a one-layer "MLP" on the tiled tensor-core operator it
// cost 48 us forward / 165 us backward per step at 196 608 samples (a 16-wide output tile for one column); it is a dot product per row:
// warp per row, lanes stride over the columns (coalesced), memory-bound.
constexpr int LIN1_MAXC = WIDE_MAX / 32;   // columns per lane

__global__ void __launch_bounds__(256) k_lin1_fwd(const float* __restrict__ x, int64_t x_stride, int64_t n, int in, const float* __restrict__ W,
                                                  const float* __restrict__ b, int act, float* __restrict__ y) {
  const int lane = threadIdx.x & 31;
  float w[LIN1_MAXC];
#pragma unroll
  for (int k = 0; k < LIN1_MAXC; ++k) w[k] = (lane + 32 * k < in) ? __ldg(W + lane + 32 * k) : 0.0f;
  const float bias = __ldg(b);
  const int64_t wstride = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t r = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5); r < n; r += wstride) {
    float acc = 0.0f;
#pragma unroll
    for (int k = 0; k < LIN1_MAXC; ++k)
      if (32 * k < in) acc = fmaf((lane + 32 * k < in) ? __ldg(x + r * x_stride + lane + 32 * k) : 0.0f, w[k], acc);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if (lane == 0) y[r] = wide_act(acc + bias, act);
  }
}

__global__ void __launch_bounds__(256) k_lin1_bwd(const float* __restrict__ x, int64_t x_stride, const float* __restrict__ y, const float* __restrict__ dy,
                                                  int64_t n, int in, const float* __restrict__ W, int act, float* __restrict__ dx, int64_t dx_stride,
                                                  float* __restrict__ dW, float* __restrict__ db) {
  __shared__ float red[8][WIDE_MAX + 1];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float w[LIN1_MAXC], gw[LIN1_MAXC];
#pragma unroll
  for (int k = 0; k < LIN1_MAXC; ++k) { w[k] = (lane + 32 * k < in) ? __ldg(W + lane + 32 * k) : 0.0f; gw[k] = 0.0f; }
  float gb = 0.0f;
  const int64_t wstride = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t r = blockIdx.x * (int64_t)(blockDim.x >> 5) + warp; r < n; r += wstride) {
    float dz = __ldg(dy + r);
    if (act == CNB_ACT_SIGMOID) { const float yy = __ldg(y + r); dz *= yy * (1.0f - yy); }
    else if (act == CNB_ACT_RELU) { if (!(__ldg(y + r) > 0.0f)) dz = 0.0f; }
    gb += dz;
#pragma unroll
    for (int k = 0; k < LIN1_MAXC; ++k) {
      const int c = lane + 32 * k;
      if (c < in) {
        gw[k] = fmaf(dz, __ldg(x + r * x_stride + c), gw[k]);
        if (dx != nullptr) dx[r * dx_stride + c] = dz * w[k];
      }
    }
  }
#pragma unroll
  for (int k = 0; k < LIN1_MAXC; ++k)
    if (lane + 32 * k < in) red[warp][lane + 32 * k] = gw[k];
  if (lane == 0) red[warp][WIDE_MAX] = gb;
  __syncthreads();
  for (int c = threadIdx.x; c <= in; c += blockDim.x) {
    const int col = c < in ? c : WIDE_MAX;
    float v = 0.0f;
    for (int q = 0; q < 8; ++q) v += red[q][col];
    if (v == 0.0f) continue;
    if (c < in) { if (dW != nullptr) atomicAdd(dW + c, v); }
    else if (db != nullptr) atomicAdd(db, v);
  }
}

}  // namespace

// Linear(in -> 1) heads: dedicated dot-product kernels (any in <= CNB_WIDE_MAX_WIDTH)
bool cnb_mlp_is_lin1(const cnb_mlp* m) { return m && m->num_layers == 1 && m->dims[1] == 1 && m->dims[0] >= 1 && m->dims[0] <= WIDE_MAX; }

int cnb_mlp_lin1_fwd(const cnb_mlp* m, const float* x, int64_t x_stride, int64_t n, float* y, cudaStream_t stream) {
  CNB_REQUIRE(m->W[0] && m->b[0], "mlp_fwd: null W/b");
  const int64_t warps = (n + 3) / 4;   // ~4 rows per warp before the grid-stride wraps
  const int grid = (int)min((warps + 7) / 8, (int64_t)cnb_num_sms() * 8);
  k_lin1_fwd<<<grid < 1 ? 1 : grid, 256, 0, stream>>>(x, x_stride, n, m->dims[0], m->W[0], m->b[0], m->out_activation, y);
  return cnb_check_launch("mlp lin1 fwd");
}

int cnb_mlp_lin1_bwd(const cnb_mlp* m, const float* x, int64_t x_stride, const float* y, const float* dy, int64_t n, float* dx, int64_t dx_stride,
                     cudaStream_t stream) {
  CNB_REQUIRE(m->W[0] != nullptr, "mlp_bwd: null W");
  const int64_t warps = (n + 15) / 16;
  const int grid = (int)min((warps + 7) / 8, (int64_t)cnb_num_sms() * 4);
  k_lin1_bwd<<<grid < 1 ? 1 : grid, 256, 0, stream>>>(x, x_stride, y, dy, n, m->dims[0], m->W[0], m->out_activation, dx, dx_stride, m->dW[0], m->db[0]);
  return cnb_check_launch("mlp lin1 bwd");
}

bool cnb_mlp_is_wide(const cnb_mlp* m) {
  if (!m || m->num_layers < 1 || m->num_layers > CNB_MAX_LAYERS) return false;
  bool wide = false;
  for (int l = 0; l <= m->num_layers; ++l) {
    if (m->dims[l] < 1 || m->dims[l] > WIDE_MAX) return false;
    wide = wide || m->dims[l] > CNB_MAX_WIDTH;
  }
  return wide;
}

int cnb_mlp_wide_fwd(const cnb_mlp* m, const float* x, int64_t x_stride, int64_t n, float* y, float* hidden, cudaStream_t stream) {
  const int nl = m->num_layers;
  float *sa = nullptr, *sb = nullptr;
  if (hidden == nullptr && nl > 1) { int rc = wide_scratch(sa, sb); if (rc) return rc; }
  for (int l = 0; l < nl; ++l) CNB_REQUIRE(m->W[l] && m->b[l], "mlp_wide: null W/b for layer %d", l);
  for (int64_t r0 = 0; r0 < n; r0 += WIDE_CHUNK) {
    const int64_t cnt = min(WIDE_CHUNK, n - r0);
    const float* in = x + r0 * x_stride;
    int64_t ldin = x_stride;
    for (int l = 0; l < nl; ++l) {
      const int K = m->dims[l], N = m->dims[l + 1];
      const bool last = l == nl - 1;
      float* out;
      int64_t ldout = N;
      if (last) out = y + r0 * N;
      else if (hidden) out = hidden + hidden_off(m, n, l) + r0 * N;
      else out = (l & 1) ? sb : sa;
      k_wide_gemm<true><<<gemm_grid(cnt, N), 256, 0, stream>>>(in, ldin, m->W[l], K, out, ldout, cnt, N, K, m->b[l], last ? m->out_activation : CNB_ACT_RELU, nullptr, 0);
      int rc = cnb_check_launch("mlp_wide_fwd");
      if (rc) return rc;
      in = out; ldin = ldout;
    }
  }
  return CNB_OK;
}

int cnb_mlp_wide_bwd(const cnb_mlp* m, const float* x, int64_t x_stride, const float* hidden, const float* y, const float* dy, int64_t n, float* dx,
                     int64_t dx_stride, cudaStream_t stream) {
  const int nl = m->num_layers;
  float *sa = nullptr, *sb = nullptr;
  int rc = wide_scratch(sa, sb);
  if (rc) return rc;
  const int out_last = m->dims[nl];
  for (int64_t r0 = 0; r0 < n; r0 += WIDE_CHUNK) {
    const int64_t cnt = min(WIDE_CHUNK, n - r0);
    // dZ of the last layer
    const float* dz = dy + r0 * out_last;
    float* nxt = sa;
    if (m->out_activation != CNB_ACT_NONE) {
      const int64_t count = cnt * out_last;
      k_wide_act_grad<<<(unsigned)min((count + 255) / 256, (int64_t)cnb_num_sms() * 8), 256, 0, stream>>>(dy + r0 * out_last, y + r0 * out_last, count, m->out_activation, sa);
      if ((rc = cnb_check_launch("mlp_wide act grad"))) return rc;
      dz = sa;
      nxt = sb;
    }
    for (int l = nl - 1; l >= 0; --l) {
      const int IN = m->dims[l], OUT = m->dims[l + 1];
      const float* in;
      int64_t ldin;
      if (l == 0) { in = x + r0 * x_stride; ldin = x_stride; }
      else { in = hidden + hidden_off(m, n, l - 1) + r0 * IN; ldin = IN; }
      if (m->dW[l] != nullptr || m->db[l] != nullptr) {
        const int64_t rows_per_block = 2048;
        const dim3 grid((unsigned)((IN + TN - 1) / TN), (unsigned)((OUT + TM - 1) / TM), (unsigned)((cnt + rows_per_block - 1) / rows_per_block));
        k_wide_dw<<<grid, 256, 0, stream>>>(dz, OUT, in, ldin, cnt, OUT, IN, rows_per_block, m->dW[l], m->db[l]);
        if ((rc = cnb_check_launch("mlp_wide dW"))) return rc;
      }
      if (l > 0) {
        // dX = dZ W, masked by the ReLU of the layer that produced X (= this layer's input)
        k_wide_gemm<false><<<gemm_grid(cnt, IN), 256, 0, stream>>>(dz, OUT, m->W[l], IN, nxt, IN, cnt, IN, OUT, nullptr, CNB_ACT_NONE, in, ldin);
        if ((rc = cnb_check_launch("mlp_wide dX"))) return rc;
        dz = nxt;
        nxt = (nxt == sa) ? sb : sa;
      } else if (dx != nullptr) {
        k_wide_gemm<false><<<gemm_grid(cnt, IN), 256, 0, stream>>>(dz, OUT, m->W[l], IN, dx + r0 * dx_stride, dx_stride, cnt, IN, OUT, nullptr, CNB_ACT_NONE, nullptr, 0);
        if ((rc = cnb_check_launch("mlp_wide dX0"))) return rc;
      }
    }
  }
  return CNB_OK;
}
