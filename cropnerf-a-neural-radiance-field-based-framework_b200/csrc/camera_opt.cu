// Camera-optimizer hook of the training step on the device (row a17 of SURVEY.md section 8; fruit_nerf.py:114-116,547,614):
// nerfstudio's CameraOptimizer(mode="SO3xR3") -- nerfacto's default, which the fruit_nerf preset inherits -- keeps one 6-vector per
// training camera (translation | axis-angle), turns it into a [R|t] correction with exp_map_SO3xR3 (cameras/lie_groups.py) and applies it
// to every ray of that camera (apply_to_raybundle: origins += t, directions = R directions); its regulariser
// (get_loss_dict: mean ||t|| * trans_l2_penalty + mean ||w|| * rot_l2_penalty) joins the training loss.
//
// With these two kernels the whole thing lives inside cnb_train_step (and therefore inside the replayed CUDA graph):
//   cnb_camera_opt_apply : per ray, exp-map of its camera's 6-vector -> adjusted origin / direction        (before the samplers)
//   cnb_camera_opt_bwd   : dLoss/d(adjusted rays) (what the hash-grid input-gradient kernels return) -> per-camera dL/dt and dL/dR = sum d_dir' (x) dir
//                          (atomics into a 12-float accumulator per camera), then one thread per camera chains dL/dR through the
//                          Rodrigues formula (incl. the clamp at |w|^2 = 1e-4) and adds the regulariser's gradient.
#include "cnb_common.cuh"

namespace {

struct Pose { float R[9]; float t[3]; };

// exp_map_SO3xR3: theta = sqrt(clamp(|w|^2, 1e-4)); R = I + sin(theta)/theta K + (1 - cos(theta))/theta^2 K^2, K = skew(w)
__device__ __forceinline__ Pose exp_map(const float* __restrict__ tv) {
  Pose p;
  const float wx = __ldg(tv + 3), wy = __ldg(tv + 4), wz = __ldg(tv + 5);
  const float theta = sqrtf(fmaxf(wx * wx + wy * wy + wz * wz, 1e-4f));
  const float inv = 1.0f / theta;
  const float f1 = inv * sinf(theta), f2 = inv * inv * (1.0f - cosf(theta));
  const float K[9] = {0.f, -wz, wy, wz, 0.f, -wx, -wy, wx, 0.f};
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      float k2 = 0.f;
#pragma unroll
      for (int q = 0; q < 3; ++q) k2 = fmaf(K[3 * i + q], K[3 * q + j], k2);
      p.R[3 * i + j] = f1 * K[3 * i + j] + f2 * k2 + (i == j ? 1.0f : 0.0f);
    }
  p.t[0] = __ldg(tv); p.t[1] = __ldg(tv + 1); p.t[2] = __ldg(tv + 2);
  return p;
}

__global__ void __launch_bounds__(256) k_camera_opt_apply(const float* __restrict__ pose, const int32_t* __restrict__ cam, const float* __restrict__ o,
                                                          const float* __restrict__ d, int64_t R, int32_t C, float* __restrict__ o_out,
                                                          float* __restrict__ d_out) {
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < R; r += (int64_t)gridDim.x * blockDim.x) {
    const int c = min(max(__ldg(cam + r), 0), C - 1);
    const Pose p = exp_map(pose + 6 * c);
    const float dx = __ldg(d + 3 * r), dy = __ldg(d + 3 * r + 1), dz = __ldg(d + 3 * r + 2);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      o_out[3 * r + i] = __ldg(o + 3 * r + i) + p.t[i];
      d_out[3 * r + i] = p.R[3 * i] * dx + p.R[3 * i + 1] * dy + p.R[3 * i + 2] * dz;
    }
  }
}

// acc[c] = [dL/dR (9, row-major) | dL/dt (3)]
__global__ void __launch_bounds__(256) k_camera_opt_reduce(const int32_t* __restrict__ cam, const float* __restrict__ d, const float* __restrict__ g_o,
                                                           const float* __restrict__ g_d, int64_t R, int32_t C, float* __restrict__ acc) {
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < R; r += (int64_t)gridDim.x * blockDim.x) {
    const int c = min(max(__ldg(cam + r), 0), C - 1);
    float* a = acc + 12 * c;
    const float dv[3] = {__ldg(d + 3 * r), __ldg(d + 3 * r + 1), __ldg(d + 3 * r + 2)};
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const float gi = __ldg(g_d + 3 * r + i);
      if (gi != 0.0f) {
#pragma unroll
        for (int j = 0; j < 3; ++j) atomicAdd(a + 3 * i + j, gi * dv[j]);
      }
      const float go = __ldg(g_o + 3 * r + i);
      if (go != 0.0f) atomicAdd(a + 9 + i, go);
    }
  }
}

__global__ void __launch_bounds__(128) k_camera_opt_finish(const float* __restrict__ pose, const float* __restrict__ acc, int32_t C, float trans_pen,
                                                           float rot_pen, float grad_scale, float* __restrict__ d_pose, float* __restrict__ reg_loss) {
  float reg = 0.0f;
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < C; c += gridDim.x * blockDim.x) {
    const float* tv = pose + 6 * c;
    const float* G = acc + 12 * c;
    const float w[3] = {__ldg(tv + 3), __ldg(tv + 4), __ldg(tv + 5)};
    const float nrm2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
    const bool free_angle = nrm2 >= 1e-4f;           // gradient of clamp(nrm2, 1e-4)
    const float theta = sqrtf(fmaxf(nrm2, 1e-4f));
    const float s = sinf(theta), co = cosf(theta);
    const float f1 = s / theta, f2 = (1.0f - co) / (theta * theta);
    const float K[9] = {0.f, -w[2], w[1], w[2], 0.f, -w[0], -w[1], w[0], 0.f};
    float K2[9];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) K2[3 * i + j] = K[3 * i] * K[j] + K[3 * i + 1] * K[3 + j] + K[3 * i + 2] * K[6 + j];
    float gK = 0.f, gK2 = 0.f;                        // dL/df1, dL/df2
#pragma unroll
    for (int q = 0; q < 9; ++q) { gK = fmaf(G[q], K[q], gK); gK2 = fmaf(G[q], K2[q], gK2); }
    const float df1 = (theta * co - s) / (theta * theta), df2 = (theta * s - 2.0f * (1.0f - co)) / (theta * theta * theta);
    const float gtheta = gK * df1 + gK2 * df2;
    float dw[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      // E = skew(e_i): dK/dw_i ; d(K^2)/dw_i = E K + K E
      float E[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (i == 0) { E[5] = -1.f; E[7] = 1.f; } else if (i == 1) { E[2] = 1.f; E[6] = -1.f; } else { E[1] = -1.f; E[3] = 1.f; }
      float acc1 = 0.f, acc2 = 0.f;
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) {
          float ek = 0.f;
#pragma unroll
          for (int q = 0; q < 3; ++q) ek += E[3 * a + q] * K[3 * q + b] + K[3 * a + q] * E[3 * q + b];
          acc1 = fmaf(G[3 * a + b], E[3 * a + b], acc1);
          acc2 = fmaf(G[3 * a + b], ek, acc2);
        }
      dw[i] = f1 * acc1 + f2 * acc2 + (free_angle ? gtheta * w[i] / theta : 0.0f);
    }
    // regulariser (get_loss_dict): mean_c ||t_c|| * trans_pen + mean_c ||w_c|| * rot_pen; d||x||/dx = x / ||x|| (0 at x = 0, like torch)
    const float tn = sqrtf(tv[0] * tv[0] + tv[1] * tv[1] + tv[2] * tv[2]), wn = sqrtf(nrm2);
    reg += (tn * trans_pen + wn * rot_pen) / (float)C;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const float rt = tn > 0.f ? grad_scale * trans_pen / (float)C * __ldg(tv + i) / tn : 0.f;
      const float rw = wn > 0.f ? grad_scale * rot_pen / (float)C * w[i] / wn : 0.f;
      d_pose[6 * c + i] += G[9 + i] + rt;
      d_pose[6 * c + 3 + i] += dw[i] + rw;
    }
  }
  if (reg_loss != nullptr) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) reg += __shfl_xor_sync(0xffffffffu, reg, off);
    if ((threadIdx.x & 31) == 0 && reg != 0.0f) atomicAdd(reg_loss, reg);
  }
}

int grid_for(int64_t n, int block) {
  int64_t b = (n + block - 1) / block;
  const int64_t cap = (int64_t)cnb_num_sms() * 8;
  return (int)(b > cap ? cap : (b < 1 ? 1 : b));
}

}  // namespace

extern "C" int cnb_camera_opt_apply(const float* pose_adjustment, const int32_t* camera_indices, const float* origins, const float* directions, int64_t R,
                                    int32_t num_cameras, float* origins_out, float* directions_out, cnb_stream_t stream) {
  CNB_REQUIRE(R >= 0 && num_cameras >= 1, "camera_opt_apply: bad sizes");
  if (R == 0) return CNB_OK;
  CNB_REQUIRE(pose_adjustment && camera_indices && origins && directions && origins_out && directions_out, "camera_opt_apply: null pointer");
  k_camera_opt_apply<<<grid_for(R, 256), 256, 0, stream>>>(pose_adjustment, camera_indices, origins, directions, R, num_cameras, origins_out, directions_out);
  return cnb_check_launch("camera_opt_apply");
}

extern "C" int cnb_camera_opt_bwd(const float* pose_adjustment, const int32_t* camera_indices, const float* directions, const float* d_origins,
                                  const float* d_directions, int64_t R, int32_t num_cameras, float trans_l2_penalty, float rot_l2_penalty, float grad_scale,
                                  float* scratch, float* d_pose_adjustment, float* reg_loss, cnb_stream_t stream) {
  CNB_REQUIRE(R >= 0 && num_cameras >= 1, "camera_opt_bwd: bad sizes");
  CNB_REQUIRE(pose_adjustment && scratch && d_pose_adjustment, "camera_opt_bwd: null pointer");
  if (cudaMemsetAsync(scratch, 0, sizeof(float) * 12 * (size_t)num_cameras, stream) != cudaSuccess) return cnb_check_launch("camera_opt_bwd memset");
  if (R > 0) {
    CNB_REQUIRE(camera_indices && directions && d_origins && d_directions, "camera_opt_bwd: null ray arrays");
    k_camera_opt_reduce<<<grid_for(R, 256), 256, 0, stream>>>(camera_indices, directions, d_origins, d_directions, R, num_cameras, scratch);
    int rc = cnb_check_launch("camera_opt_bwd reduce");
    if (rc) return rc;
  }
  k_camera_opt_finish<<<grid_for(num_cameras, 128), 128, 0, stream>>>(pose_adjustment, scratch, num_cameras, trans_l2_penalty, rot_l2_penalty,
                                                                      grad_scale == 0.0f ? 1.0f : grad_scale, d_pose_adjustment, reg_loss);
  return cnb_check_launch("camera_opt_bwd finish");
}
