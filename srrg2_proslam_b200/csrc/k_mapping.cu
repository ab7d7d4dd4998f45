// k_mapping.cu -- N3: per-landmark structure-only filtering after the alignment (sm_100a).
//
// Reference path replaced (all in-tree):
//   LandmarkEstimatorEKF_::compute        .../mapping/landmarks/landmark_estimator_ekf_impl.cpp:17-82
//   PointEKFBase::_predict / _correct     .../mapping/landmarks/filters/point_ekf_base.hpp:62-131
//   ProjectivePointEKF / ProjectiveDepthPointEKF / StereoProjectivePointEKF::_computeMeasurementPrediction
//                                         .../filters/{projective,projective_depth,stereo_projective}_point_ekf_impl.cpp
// The merger (mapping/mergers/merger_projective_impl.cpp:61-135) calls the estimator once per correspondence; the
// correspondences of a frame are bijective, so every landmark is updated independently: one thread per landmark,
// double precision inside like the reference ("we locally operate in double precision", landmark_estimator_ekf.h:21),
// fp32 landmark statistics in and out.  Operation order = oracle/pslam_oracle_mapping.hpp.
#include "pslam_internal.cuh"
#include "pslam_kernels.cuh"

namespace {

struct EkfParams {
  double fx, fy, cx, cy, bx, by;
  double min_cov, max_cov_norm2;
  float max_dist2;
  double Rw[9], tw[3];          // world_in_sensor (fp32 values, LandmarkEstimatorBase_::setTransforms)
  float sensor_in_world[12];    // row-major 3x4
  float Rl[9], tl[3];           // world_in_local_map
};

template <int E>
__device__ __forceinline__ void spd_inverse(double* A) {
  double I[E * E];
#pragma unroll
  for (int i = 0; i < E; ++i)
#pragma unroll
    for (int j = 0; j < E; ++j) I[i * E + j] = i == j ? 1.0 : 0.0;
#pragma unroll
  for (int k = 0; k < E; ++k) {
    const double inv = 1.0 / A[k * E + k];
#pragma unroll
    for (int j = 0; j < E; ++j) {
      A[k * E + j] *= inv;
      I[k * E + j] *= inv;
    }
#pragma unroll
    for (int i = 0; i < E; ++i) {
      if (i == k) continue;
      const double f = A[i * E + k];
#pragma unroll
      for (int j = 0; j < E; ++j) {
        A[i * E + j] -= f * A[k * E + j];
        I[i * E + j] -= f * I[k * E + j];
      }
    }
  }
#pragma unroll
  for (int i = 0; i < E * E; ++i) A[i] = I[i];
}

template <int E>
__global__ void __launch_bounds__(128, 4)
landmarks_ekf_kernel(const EkfParams p, long long n, float* __restrict__ state_world, float* __restrict__ covariance,
                     const float* __restrict__ meas, float* __restrict__ coords_in_local_map, uint8_t* __restrict__ inlier,
                     int* __restrict__ n_inliers) {
  const long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  bool ok = false;
  if (i < n) {
    double cov[9], st[3], z[E];
#pragma unroll
    for (int k = 0; k < 9; ++k) cov[k] = covariance[9 * i + k];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      st[k] = state_world[3 * i + k];
      cov[4 * k] = cov[4 * k] > p.min_cov ? cov[4 * k] : p.min_cov;  // landmark_estimator_ekf_impl.cpp:47-49
    }
#pragma unroll
    for (int k = 0; k < E; ++k) z[k] = meas[(long long) E * i + k];
    // ---- _predict (point_ekf_base.hpp:62-76): P = R cov R^T (+ 0), s = T state
    double RC[9], P[9], s[3];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = 0; b < 3; ++b) RC[3 * a + b] = (p.Rw[3 * a] * cov[b] + p.Rw[3 * a + 1] * cov[3 + b]) + p.Rw[3 * a + 2] * cov[6 + b];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = 0; b < 3; ++b) P[3 * a + b] = (RC[3 * a] * p.Rw[3 * b] + RC[3 * a + 1] * p.Rw[3 * b + 1]) + RC[3 * a + 2] * p.Rw[3 * b + 2];
#pragma unroll
    for (int a = 0; a < 3; ++a) s[a] = ((p.Rw[3 * a] * st[0] + p.Rw[3 * a + 1] * st[1]) + p.Rw[3 * a + 2] * st[2]) + p.tw[a];
    // ---- _computeMeasurementPrediction
    double h[E], J[E * 3];
#pragma unroll
    for (int k = 0; k < E * 3; ++k) J[k] = 0.0;
    {
      const double x = s[0], y = s[1], zz = s[2];
      const double z_2 = zz * zz, fx_x = p.fx * x, fy_y = p.fy * y, fx_by_z = p.fx / zz, fy_by_z = p.fy / zz;
      if (E == 4) {  // stereo_projective_point_ekf_impl.cpp:22-47
        const double x_h = fx_x + p.cx * zz, y_h = fy_y + p.cy * zz;
        h[0] = x_h / zz;
        h[1] = y_h / zz;
        h[E - 2] = (x_h - p.bx) / zz;
        h[E - 1] = (y_h - p.by) / zz;
        J[0] = fx_by_z;
        J[2] = -fx_x / z_2;
        J[4] = fy_by_z;
        J[5] = -fy_y / z_2;
        J[3 * (E - 2)] = fx_by_z;
        J[3 * (E - 2) + 2] = -(fx_x - p.bx) / z_2;
        J[3 * (E - 1) + 1] = fy_by_z;
        J[3 * (E - 1) + 2] = -(fy_y - p.by) / z_2;
      } else {  // projective_point_ekf_impl.cpp:24-43, projective_depth_point_ekf_impl.cpp:14-36
        h[0] = fx_by_z * x + p.cx;
        h[1] = fy_by_z * y + p.cy;
        J[0] = fx_by_z;
        J[2] = -fx_x / z_2;
        J[4] = fy_by_z;
        J[5] = -fy_y / z_2;
        if (E == 3) {
          h[E - 1] = zz;
          J[3 * (E - 1) + 2] = 1.0;
        }
      }
    }
    // ---- _correct (point_ekf_base.hpp:79-131): measurement covariance = min_cov * I (estimator :26-28)
    double PJt[3 * E], S[E * E], G[3 * E];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int k = 0; k < E; ++k) PJt[a * E + k] = (P[3 * a] * J[3 * k] + P[3 * a + 1] * J[3 * k + 1]) + P[3 * a + 2] * J[3 * k + 2];
#pragma unroll
    for (int a = 0; a < E; ++a)
#pragma unroll
      for (int b = 0; b < E; ++b)
        S[a * E + b] = (a == b ? p.min_cov : 0.0) + ((J[3 * a] * PJt[b] + J[3 * a + 1] * PJt[E + b]) + J[3 * a + 2] * PJt[2 * E + b]);
    spd_inverse<E>(S);
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = 0; b < E; ++b) {
        double acc = 0;
#pragma unroll
        for (int k = 0; k < E; ++k) acc += PJt[a * E + k] * S[k * E + b];
        G[a * E + b] = acc;
      }
    double ns[3], nc[9];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      double acc = 0;
#pragma unroll
      for (int k = 0; k < E; ++k) acc += G[a * E + k] * (z[k] - h[k]);
      ns[a] = s[a] + acc;
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      double ikj[3];
#pragma unroll
      for (int b = 0; b < 3; ++b) {
        double acc = 0;
#pragma unroll
        for (int k = 0; k < E; ++k) acc += G[a * E + k] * J[3 * k + b];
        ikj[b] = (a == b ? 1.0 : 0.0) - acc;
      }
#pragma unroll
      for (int b = 0; b < 3; ++b) nc[3 * a + b] = (ikj[0] * P[b] + ikj[1] * P[3 + b]) + ikj[2] * P[6 + b];
    }
    double norm2 = 0;
#pragma unroll
    for (int k = 0; k < 9; ++k) norm2 += nc[k] * nc[k];
    if (!(ns[2] <= 0 || norm2 > p.max_cov_norm2)) {  // estimator :59-63
      const float sf0 = (float) ns[0], sf1 = (float) ns[1], sf2 = (float) ns[2];
      float w[3];
#pragma unroll
      for (int a = 0; a < 3; ++a)
        w[a] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(p.sensor_in_world[4 * a], sf0), __fmul_rn(p.sensor_in_world[4 * a + 1], sf1)),
                                   __fmul_rn(p.sensor_in_world[4 * a + 2], sf2)), p.sensor_in_world[4 * a + 3]);
      const float d0 = __fsub_rn(w[0], state_world[3 * i]), d1 = __fsub_rn(w[1], state_world[3 * i + 1]), d2 = __fsub_rn(w[2], state_world[3 * i + 2]);
      const float dist2 = __fadd_rn(__fadd_rn(__fmul_rn(d0, d0), __fmul_rn(d1, d1)), __fmul_rn(d2, d2));
      if (!(dist2 > p.max_dist2)) {  // estimator :69-73
        ok = true;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          state_world[3 * i + a] = w[a];
          coords_in_local_map[3 * i + a] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(p.Rl[3 * a], w[0]), __fmul_rn(p.Rl[3 * a + 1], w[1])),
                                                               __fmul_rn(p.Rl[3 * a + 2], w[2])), p.tl[a]);
        }
#pragma unroll
        for (int k = 0; k < 9; ++k) covariance[9 * i + k] = (float) nc[k];
      }
    }
    inlier[i] = ok ? 1 : 0;
  }
  const unsigned bal = __ballot_sync(0xffffffffu, ok);
  if ((threadIdx.x & 31) == 0 && bal) atomicAdd(n_inliers, __popc(bal));
}

// LandmarkEstimatorWeightedMean_::compute (.../mapping/landmarks/landmark_estimator_weighted_mean_impl.cpp:7-41), fp32,
// operation order of the oracle (bit exact): one thread per landmark
struct WmParams {
  float sw[12];         // sensor_in_world, row-major 3x4
  float Rl[9], tl[3];   // world_in_local_map
  float max_dist2;
};
__global__ void __launch_bounds__(256)
landmarks_weighted_mean_kernel(const WmParams p, long long n, float* __restrict__ state_world, const int* __restrict__ n_opt,
                               const float* __restrict__ landmark_in_sensor, float* __restrict__ coords_in_local_map,
                               uint8_t* __restrict__ inlier, int* __restrict__ n_inliers) {
  const long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  bool ok = false;
  if (i < n) {
    const float lx = landmark_in_sensor[3 * i], ly = landmark_in_sensor[3 * i + 1], lz = landmark_in_sensor[3 * i + 2];
    const float s0 = state_world[3 * i], s1 = state_world[3 * i + 1], s2 = state_world[3 * i + 2];
    const float st[3] = {s0, s1, s2};
    const float n1 = (float) (n_opt[i] + 1), n2 = __fadd_rn(n1, 1.0f);
    float w[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const float upd = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(p.sw[4 * a], lx), __fmul_rn(p.sw[4 * a + 1], ly)), __fmul_rn(p.sw[4 * a + 2], lz)), p.sw[4 * a + 3]);
      w[a] = __fdiv_rn(__fadd_rn(__fmul_rn(n1, st[a]), upd), n2);
    }
    const float d0 = __fsub_rn(w[0], s0), d1 = __fsub_rn(w[1], s1), d2 = __fsub_rn(w[2], s2);
    const float dist2 = __fadd_rn(__fadd_rn(__fmul_rn(d0, d0), __fmul_rn(d1, d1)), __fmul_rn(d2, d2));
    ok = !(dist2 > p.max_dist2);
    if (ok) {
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        state_world[3 * i + a] = w[a];
        coords_in_local_map[3 * i + a] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(p.Rl[3 * a], w[0]), __fmul_rn(p.Rl[3 * a + 1], w[1])),
                                                             __fmul_rn(p.Rl[3 * a + 2], w[2])), p.tl[a]);
      }
    }
    inlier[i] = ok ? 1 : 0;
  }
  const unsigned bal = __ballot_sync(0xffffffffu, ok);
  if ((threadIdx.x & 31) == 0 && bal) atomicAdd(n_inliers, __popc(bal));
}

}  // namespace

// all pointers are device pointers; d_n_inliers: one int, zeroed here
int pslam_k_landmarks_ekf(pslam_ctx* ctx, const pslam_ekf_cfg* cfg, long long n, float* d_state_world, float* d_covariance,
                          const float* d_meas, float* d_local, uint8_t* d_inlier, int* d_n_inliers) {
  EkfParams p;
  p.fx = cfg->K[0];
  p.fy = cfg->K[4];
  p.cx = cfg->K[2];
  p.cy = cfg->K[5];
  p.bx = cfg->baseline_pixels[0];
  p.by = cfg->baseline_pixels[1];
  p.min_cov = cfg->minimum_state_element_covariance;
  p.max_cov_norm2 = cfg->maximum_covariance_norm_squared;
  p.max_dist2 = cfg->maximum_distance_geometry_meters_squared;
  // LandmarkEstimatorBase_::setTransforms (landmark_estimator_base.hpp:49-58), fp32, operation order of the oracle's Pose
  const float* A = cfg->sensor_in_world;
  const float* L = cfg->sensor_in_local_map;
  float Rw[9], tw[3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) Rw[3 * i + j] = A[4 * j + i];
  for (int i = 0; i < 3; ++i) tw[i] = -((Rw[3 * i] * A[3] + Rw[3 * i + 1] * A[7]) + Rw[3 * i + 2] * A[11]);
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) {
      p.Rw[3 * i + j] = Rw[3 * i + j];
      p.Rl[3 * i + j] = (L[4 * i] * Rw[j] + L[4 * i + 1] * Rw[3 + j]) + L[4 * i + 2] * Rw[6 + j];  // sensor_in_local_map * world_in_sensor
    }
    p.tw[i] = tw[i];
    p.tl[i] = ((L[4 * i] * tw[0] + L[4 * i + 1] * tw[1]) + L[4 * i + 2] * tw[2]) + L[4 * i + 3];
  }
  for (int i = 0; i < 12; ++i) p.sensor_in_world[i] = A[i];
  PSLAM_CUDA_TRY(ctx, cudaMemsetAsync(d_n_inliers, 0, sizeof(int), ctx->stream));
  if (n == 0) return PSLAM_OK;
  const unsigned grid = (unsigned) ((n + 127) / 128);
  if (cfg->kind == 0) landmarks_ekf_kernel<2><<<grid, 128, 0, ctx->stream>>>(p, n, d_state_world, d_covariance, d_meas, d_local, d_inlier, d_n_inliers);
  else if (cfg->kind == 1) landmarks_ekf_kernel<3><<<grid, 128, 0, ctx->stream>>>(p, n, d_state_world, d_covariance, d_meas, d_local, d_inlier, d_n_inliers);
  else if (cfg->kind == 2) landmarks_ekf_kernel<4><<<grid, 128, 0, ctx->stream>>>(p, n, d_state_world, d_covariance, d_meas, d_local, d_inlier, d_n_inliers);
  else return pslam_set_error(ctx, PSLAM_E_INVALID, "landmarks_ekf: unknown filter kind", cudaSuccess);
  PSLAM_LAUNCH_CHECK(ctx, "landmarks_ekf_kernel");
  return PSLAM_OK;
}

int pslam_k_landmarks_weighted_mean(pslam_ctx* ctx, const float* sensor_in_world12, const float* sensor_in_local_map12, float max_dist2,
                                    long long n, float* d_state_world, const int* d_n_opt, const float* d_landmark_in_sensor,
                                    float* d_local, uint8_t* d_inlier, int* d_n_inliers) {
  WmParams p;
  const float* A = sensor_in_world12;
  const float* L = sensor_in_local_map12;
  float Rw[9], tw[3];  // world_in_sensor, fp32 (LandmarkEstimatorBase_::setTransforms)
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) Rw[3 * i + j] = A[4 * j + i];
  for (int i = 0; i < 3; ++i) tw[i] = -((Rw[3 * i] * A[3] + Rw[3 * i + 1] * A[7]) + Rw[3 * i + 2] * A[11]);
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) p.Rl[3 * i + j] = (L[4 * i] * Rw[j] + L[4 * i + 1] * Rw[3 + j]) + L[4 * i + 2] * Rw[6 + j];
    p.tl[i] = ((L[4 * i] * tw[0] + L[4 * i + 1] * tw[1]) + L[4 * i + 2] * tw[2]) + L[4 * i + 3];
  }
  for (int i = 0; i < 12; ++i) p.sw[i] = A[i];
  p.max_dist2 = max_dist2;
  PSLAM_CUDA_TRY(ctx, cudaMemsetAsync(d_n_inliers, 0, sizeof(int), ctx->stream));
  if (n == 0) return PSLAM_OK;
  landmarks_weighted_mean_kernel<<<(unsigned) ((n + 255) / 256), 256, 0, ctx->stream>>>(p, n, d_state_world, d_n_opt, d_landmark_in_sensor, d_local,
                                                                                      d_inlier, d_n_inliers);
  PSLAM_LAUNCH_CHECK(ctx, "landmarks_weighted_mean_kernel");
  return PSLAM_OK;
}
