// pslam_oracle_mapping.hpp -- TEST INFRASTRUCTURE (CPU oracle), not the product.
//
// CPU restatement of the per-landmark structure-only filters that follow the frontend path:
//   PointEKFBase::_predict / _correct          .../mapping/landmarks/filters/point_ekf_base.hpp:62-131
//   ProjectivePointEKF                          .../filters/projective_point_ekf_impl.cpp:15-44
//   ProjectiveDepthPointEKF                     .../filters/projective_depth_point_ekf_impl.cpp:5-37
//   StereoProjectivePointEKF                    .../filters/stereo_projective_point_ekf_impl.cpp:13-48
//   LandmarkEstimatorEKF_::setTransforms/compute .../mapping/landmarks/landmark_estimator_ekf_impl.cpp:6-82
//   LandmarkEstimatorBase_::setTransforms        .../mapping/landmarks/landmark_estimator_base.hpp:49-58
//   LandmarkEstimatorWeightedMean_::compute      .../mapping/landmarks/landmark_estimator_weighted_mean_impl.cpp:7-41
// All of this code is IN the reference tree (no external arithmetic except Eigen's fixed-size inverse(), restated as a
// Gauss-Jordan elimination of the symmetric positive definite innovation covariance) and is pinned through the
// scenarios of tests/test_{projective,projective_depth,stereo_projective}_point_ekf.cpp (tests/test_oracle_ekf.py).
// The filter runs in double like the reference ("we locally operate in double precision", landmark_estimator_ekf.h:21).
//   LandmarkEstimatorPoseBasedSmoother_::compute .../mapping/landmarks/landmark_estimator_pose_based_smoother_impl.cpp:6-148
//     (pinned through the scenario of tests/test_landmark_estimators.cpp:210-258, tests/test_oracle_ekf.py)
//   MergerProjective_::compute / _addPoints binning .../mapping/mergers/merger_projective_impl.cpp:61-135,205-253
//     (the sequential map-of-maps walk; pinned on the reference's constant tests/test_mergers.cpp:337 -- ICL 00 -> 01 grows
//     the scene from 321 to 337 points -- and :286-287, tests/test_oracle_merger.py)
#pragma once
#include <cmath>
#include <cstring>
#include <unordered_map>
#include <utility>
#include <vector>

namespace pslam_oracle {

enum EkfKind { EKF_PROJECTIVE = 0, EKF_PROJECTIVE_DEPTH = 1, EKF_STEREO = 2 };
static inline int ekf_measurement_dim(int kind) { return kind == EKF_PROJECTIVE ? 2 : (kind == EKF_PROJECTIVE_DEPTH ? 3 : 4); }

struct EkfCamera {
  double fx = 1, fy = 1, cx = 0, cy = 0;  // ProjectivePointEKF::setCameraMatrix (projective_point_ekf_impl.cpp:6-12)
  double bx = 0, by = 0;                  // StereoProjectivePointEKF::setBaseline (stereo_projective_point_ekf_impl.cpp:6-10)
};

// in-place inverse of a symmetric positive definite E x E matrix (row major), Gauss-Jordan without pivoting
template <int E>
static inline void spd_inverse(double* A) {
  double I[E * E];
  for (int i = 0; i < E; ++i)
    for (int j = 0; j < E; ++j) I[i * E + j] = i == j ? 1.0 : 0.0;
  for (int k = 0; k < E; ++k) {
    const double inv = 1.0 / A[k * E + k];
    for (int j = 0; j < E; ++j) {
      A[k * E + j] *= inv;
      I[k * E + j] *= inv;
    }
    for (int i = 0; i < E; ++i) {
      if (i == k) continue;
      const double f = A[i * E + k];
      for (int j = 0; j < E; ++j) {
        A[i * E + j] -= f * A[k * E + j];
        I[i * E + j] -= f * I[k * E + j];
      }
    }
  }
  std::memcpy(A, I, sizeof(I));
}

// h(state) and its 3-column Jacobian (row major E x 3) for the three filters: E = 2 projective, 3 projective depth,
// 4 rectified stereo
template <int E>
static inline void ekf_predict_measurement(const EkfCamera& c, const double* s, double* h, double* J) {
  const double x = s[0], y = s[1], z = s[2];
  const double z_2 = z * z, fx_x = c.fx * x, fy_y = c.fy * y, fx_by_z = c.fx / z, fy_by_z = c.fy / z;
  for (int i = 0; i < E * 3; ++i) J[i] = 0.0;
  if constexpr (E == 4) {  // stereo_projective_point_ekf_impl.cpp:22-47
    const double x_h = fx_x + c.cx * z, y_h = fy_y + c.cy * z;
    h[0] = x_h / z;
    h[1] = y_h / z;
    h[2] = (x_h - c.bx) / z;
    h[3] = (y_h - c.by) / z;
    J[0] = fx_by_z; J[2] = -fx_x / z_2;
    J[4] = fy_by_z; J[5] = -fy_y / z_2;
    J[6] = fx_by_z; J[8] = -(fx_x - c.bx) / z_2;
    J[10] = fy_by_z; J[11] = -(fy_y - c.by) / z_2;
  } else {  // projective_point_ekf_impl.cpp:24-43, projective_depth_point_ekf_impl.cpp:14-36
    h[0] = fx_by_z * x + c.cx;
    h[1] = fy_by_z * y + c.cy;
    J[0] = fx_by_z; J[2] = -fx_x / z_2;
    J[4] = fy_by_z; J[5] = -fy_y / z_2;
    if constexpr (E == 3) {
      h[2] = z;
      J[8] = 1.0;
    }
  }
}

// PointEKFBase::compute = _predict + _correct (point_ekf_base.hpp:48-131).  T = world_in_sensor (R row major 3x3, t),
// Q = transition covariance, Rm = measurement covariance (E x E); state / cov are updated in place.
template <int E>
static inline void ekf_compute(const EkfCamera& cam, const double* R, const double* t, const double* Q,
                               const double* meas, const double* Rm, double* state, double* cov) {
  // _predict: cov = R cov R^T + Q; state = T state
  double RC[9], P[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) RC[3 * i + j] = (R[3 * i] * cov[j] + R[3 * i + 1] * cov[3 + j]) + R[3 * i + 2] * cov[6 + j];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      P[3 * i + j] = ((RC[3 * i] * R[3 * j] + RC[3 * i + 1] * R[3 * j + 1]) + RC[3 * i + 2] * R[3 * j + 2]) + Q[3 * i + j];
  double s[3];
  for (int i = 0; i < 3; ++i) s[i] = ((R[3 * i] * state[0] + R[3 * i + 1] * state[1]) + R[3 * i + 2] * state[2]) + t[i];
  // _correct
  double h[E], J[E * 3];
  ekf_predict_measurement<E>(cam, s, h, J);
  double PJt[3 * E];  // P J^T (3 x E)
  for (int i = 0; i < 3; ++i)
    for (int k = 0; k < E; ++k) PJt[i * E + k] = (P[3 * i] * J[3 * k] + P[3 * i + 1] * J[3 * k + 1]) + P[3 * i + 2] * J[3 * k + 2];
  double S[E * E];  // Rm + J P J^T
  for (int a = 0; a < E; ++a)
    for (int b = 0; b < E; ++b)
      S[a * E + b] = Rm[a * E + b] + ((J[3 * a] * PJt[b] + J[3 * a + 1] * PJt[E + b]) + J[3 * a + 2] * PJt[2 * E + b]);
  spd_inverse<E>(S);
  double G[3 * E];  // Kalman gain
  for (int i = 0; i < 3; ++i)
    for (int b = 0; b < E; ++b) {
      double acc = 0;
      for (int a = 0; a < E; ++a) acc += PJt[i * E + a] * S[a * E + b];
      G[i * E + b] = acc;
    }
  for (int i = 0; i < 3; ++i) {
    double acc = 0;
    for (int a = 0; a < E; ++a) acc += G[i * E + a] * (meas[a] - h[a]);
    state[i] = s[i] + acc;
  }
  double IKJ[9];  // I - G J
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double acc = 0;
      for (int a = 0; a < E; ++a) acc += G[i * E + a] * J[3 * a + j];
      IKJ[3 * i + j] = (i == j ? 1.0 : 0.0) - acc;
    }
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) cov[3 * i + j] = (IKJ[3 * i] * P[j] + IKJ[3 * i + 1] * P[3 + j]) + IKJ[3 * i + 2] * P[6 + j];
}

struct LandmarkEkfConfig {
  int kind = EKF_STEREO;
  float K[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  double baseline_pixels[2] = {0, 0};
  double minimum_state_element_covariance = 0.01;           // landmark_estimator_ekf.h:33-37
  double maximum_covariance_norm_squared = 1;               // :38-42
  float maximum_distance_geometry_meters_squared = 1;       // landmark_estimator_base.hpp:22-26
  float sensor_in_world[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};      // measurement_in_world
  float sensor_in_local_map[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};  // measurement_in_scene
};

// LandmarkEstimatorEKF_::compute for ONE landmark (landmark_estimator_ekf_impl.cpp:17-82) with the transforms of
// LandmarkEstimatorBase_::setTransforms (fp32, landmark_estimator_base.hpp:49-58).  state_world / covariance are the
// landmark statistics (fp32); on success (return true = isInlier) they receive the filtered values that
// addOptimizationResult stores (:74-75) and coords_in_local_map the landmark's new local coordinates (:79-80).
template <int E>
static inline bool landmark_ekf_update(const LandmarkEkfConfig& cfg, const float* world_in_sensor_R, const float* world_in_sensor_t,
                                       const float* world_in_local_map_R, const float* world_in_local_map_t, float* state_world,
                                       float* covariance, const float* measurement, float* coords_in_local_map) {
  EkfCamera cam;
  cam.fx = cfg.K[0];
  cam.fy = cfg.K[4];
  cam.cx = cfg.K[2];
  cam.cy = cfg.K[5];
  cam.bx = cfg.baseline_pixels[0];
  cam.by = cfg.baseline_pixels[1];
  double Rm[E * E], Q[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, R[9], t[3], state[3], cov[9], meas[E];
  for (int i = 0; i < E * E; ++i) Rm[i] = 0;
  for (int i = 0; i < E; ++i) Rm[i * E + i] = cfg.minimum_state_element_covariance;  // :26-28
  for (int i = 0; i < 9; ++i) {
    R[i] = world_in_sensor_R[i];
    cov[i] = covariance[i];
  }
  for (int i = 0; i < 3; ++i) {
    t[i] = world_in_sensor_t[i];
    state[i] = state_world[i];
    cov[4 * i] = cov[4 * i] > cfg.minimum_state_element_covariance ? cov[4 * i] : cfg.minimum_state_element_covariance;  // :47-49
  }
  for (int i = 0; i < E; ++i) meas[i] = measurement[i];
  ekf_compute<E>(cam, R, t, Q, meas, Rm, state, cov);
  double norm2 = 0;
  for (int i = 0; i < 9; ++i) norm2 += cov[i] * cov[i];
  if (state[2] <= 0 || norm2 > cfg.maximum_covariance_norm_squared) return false;  // :59-63
  // coordinates_in_world = sensor_in_world * state.cast<float>()  (:67-68)
  const float sf[3] = {(float) state[0], (float) state[1], (float) state[2]};
  float w[3];
  for (int i = 0; i < 3; ++i)
    w[i] = ((cfg.sensor_in_world[4 * i] * sf[0] + cfg.sensor_in_world[4 * i + 1] * sf[1]) + cfg.sensor_in_world[4 * i + 2] * sf[2]) +
           cfg.sensor_in_world[4 * i + 3];
  const float d0 = w[0] - state_world[0], d1 = w[1] - state_world[1], d2 = w[2] - state_world[2];
  if ((d0 * d0 + d1 * d1) + d2 * d2 > cfg.maximum_distance_geometry_meters_squared) return false;  // :69-73
  for (int i = 0; i < 3; ++i) {
    state_world[i] = w[i];
    coords_in_local_map[i] = ((world_in_local_map_R[3 * i] * w[0] + world_in_local_map_R[3 * i + 1] * w[1]) + world_in_local_map_R[3 * i + 2] * w[2]) +
                             world_in_local_map_t[i];
  }
  for (int i = 0; i < 9; ++i) covariance[i] = (float) cov[i];
  return true;
}

// LandmarkEstimatorWeightedMean_::compute for ONE landmark, fp32
// (.../mapping/landmarks/landmark_estimator_weighted_mean_impl.cpp:7-41).  sensor_in_world / world_in_local_map: R row
// major 3x3 + t.  Returns isInlier; on success state_world (what addOptimizationResult stores, :35) and the local
// coordinates (:39) are written.
static inline bool landmark_weighted_mean_update(const float* sw_R, const float* sw_t, const float* wl_R, const float* wl_t,
                                                 float max_dist2, int number_of_optimizations, const float* landmark_in_sensor,
                                                 float* state_world, float* coords_in_local_map) {
  float upd[3], w[3];
  for (int i = 0; i < 3; ++i)
    upd[i] = ((sw_R[3 * i] * landmark_in_sensor[0] + sw_R[3 * i + 1] * landmark_in_sensor[1]) + sw_R[3 * i + 2] * landmark_in_sensor[2]) + sw_t[i];  // :20-21
  const float n1 = (float) (number_of_optimizations + 1);  // :23
  for (int i = 0; i < 3; ++i) w[i] = (n1 * state_world[i] + upd[i]) / (n1 + 1);  // :25-27
  const float d0 = w[0] - state_world[0], d1 = w[1] - state_world[1], d2 = w[2] - state_world[2];
  if ((d0 * d0 + d1 * d1) + d2 * d2 > max_dist2) return false;  // :30-34
  for (int i = 0; i < 3; ++i) state_world[i] = w[i];
  for (int i = 0; i < 3; ++i) coords_in_local_map[i] = ((wl_R[3 * i] * w[0] + wl_R[3 * i + 1] * w[1]) + wl_R[3 * i + 2] * w[2]) + wl_t[i];
  return true;
}

// -----------------------------------------------------------------------------------------------------------------
// LandmarkEstimatorPoseBasedSmoother_::compute  (.../mapping/landmarks/landmark_estimator_pose_based_smoother_impl.cpp:6-148)
// fp32 like the reference.  The landmark's measurement history lives in srrg2_core's PointStatisticsField3D (external,
// un-vendored); what the reference code shows of it is restated here as plain arrays:
//   CameraMeasurement(point_in_image, point_in_camera, sensor_in_world, world_in_sensor)  (:16-19) with the accessors
//   camera_from_world * point_in_camera -> world (:142) and world_from_camera * world -> camera (:60), i.e. one
//   sensor pose per frame: `frame[k]` indexes the tables frames_sensor_in_world / frames_world_in_sensor (R row major
//   3x3 + t, 12 floats).  The history handed in already contains the current measurement (addMeasurement, :15).
//   [upstream, not verifiable here] addOptimizationResult(x) is taken to store x as the state and to increment
//   numberOfOptimizations by one (consistent with the asserts at :127,135).
// Eigen's Matrix3f::fullPivLu().solve (:112) is restated as full-pivot Gaussian elimination with Eigen's rank rule.
// -----------------------------------------------------------------------------------------------------------------
struct SmootherConfig {
  float K[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  unsigned maximum_number_of_iterations = 100;                     // landmark_estimator_pose_based_smoother.h:14-18
  float convergence_criterion_minimum_chi2_delta = 1e-5f;          // :19-23
  float maximum_reprojection_error_pixels_squared = 100;           // :24-28
  unsigned minimum_number_of_measurements_for_optimization = 3;    // :29-33
  float maximum_distance_geometry_meters_squared = 1;              // landmark_estimator_base.hpp:22-26
};

static inline void smoother_apply(const float* T12, const float* p, float* o) {  // Isometry3f * Vector3f
  for (int i = 0; i < 3; ++i) o[i] = ((T12[4 * i] * p[0] + T12[4 * i + 1] * p[1]) + T12[4 * i + 2] * p[2]) + T12[4 * i + 3];
}

// x = A^-1 rhs for a 3x3 system with full pivoting (Eigen::FullPivLU::solve semantics, rank threshold eps * 3)
static inline void full_piv_lu_solve3(const float* A_in, const float* rhs, float* x) {
  float A[9];
  for (int i = 0; i < 9; ++i) A[i] = A_in[i];
  int rp[3] = {0, 1, 2}, cp[3] = {0, 1, 2};  // row / column transpositions applied so far
  float c[3] = {rhs[0], rhs[1], rhs[2]};
  float maxpivot = 0;
  int nonzero = 3;
  for (int k = 0; k < 3; ++k) {
    int br = k, bc = k;
    float best = -1;
    for (int j = k; j < 3; ++j)       // column-major scan, the first maximum wins (Eigen's maxCoeff visitor)
      for (int i = k; i < 3; ++i) {
        const float v = std::fabs(A[3 * i + j]);
        if (v > best) {
          best = v;
          br = i;
          bc = j;
        }
      }
    if (best == 0.0f) {
      nonzero = k;
      break;
    }
    if (best > maxpivot) maxpivot = best;
    if (br != k) {
      for (int j = 0; j < 3; ++j) std::swap(A[3 * k + j], A[3 * br + j]);
      std::swap(c[k], c[br]);
      std::swap(rp[k], rp[br]);
    }
    if (bc != k) {
      for (int i = 0; i < 3; ++i) std::swap(A[3 * i + k], A[3 * i + bc]);
      std::swap(cp[k], cp[bc]);
    }
    for (int i = k + 1; i < 3; ++i) A[3 * i + k] /= A[3 * k + k];
    for (int i = k + 1; i < 3; ++i)
      for (int j = k + 1; j < 3; ++j) A[3 * i + j] -= A[3 * i + k] * A[3 * k + j];
  }
  // rank: pivots above maxpivot * epsilon * 3
  const float thr = maxpivot * (1.1920929e-7f * 3.0f);
  int rank = 0;
  for (int k = 0; k < nonzero; ++k) rank += std::fabs(A[3 * k + k]) > thr;
  // forward substitution with the unit lower factor
  for (int i = 1; i < 3; ++i)
    for (int j = 0; j < i; ++j) c[i] -= A[3 * i + j] * c[j];
  // back substitution on the leading rank x rank block of U
  float y[3] = {0, 0, 0};
  for (int i = rank - 1; i >= 0; --i) {
    float acc = c[i];
    for (int j = i + 1; j < rank; ++j) acc -= A[3 * i + j] * y[j];
    y[i] = acc / A[3 * i + i];
  }
  for (int k = 0; k < 3; ++k) x[cp[k]] = y[k];
  (void) rp;
}

// One landmark.  frames_*: tables of 12 floats per frame; hist_*: this landmark's n_meas measurements.
// world_in_local_map12: LandmarkEstimatorBase_::setTransforms.  Returns isInlier; state_world / number_of_optimizations
// / coords_in_local_map are updated exactly where the reference updates them.
static inline bool landmark_smoother_update(const SmootherConfig& cfg, const float* frames_sensor_in_world,
                                            const float* frames_world_in_sensor, const float* world_in_local_map12, int n_meas,
                                            const int* hist_frame, const float* hist_uv, const float* hist_point_in_camera,
                                            float* state_world, int* number_of_optimizations, float* coords_in_local_map) {
  const float initial[3] = {state_world[0], state_world[1], state_world[2]};
  float w[3] = {initial[0], initial[1], initial[2]};
  auto mean_in_world = [&](float* out) {  // _setMeanCoordinatesInWorld (:137-146)
    float acc[3] = {0, 0, 0};
    for (int k = 0; k < n_meas; ++k) {
      float p[3];
      smoother_apply(frames_sensor_in_world + 12 * hist_frame[k], hist_point_in_camera + 3 * k, p);
      for (int i = 0; i < 3; ++i) acc[i] += p[i];
    }
    for (int i = 0; i < 3; ++i) out[i] = acc[i] / (float) n_meas;
  };
  if ((unsigned) n_meas < cfg.minimum_number_of_measurements_for_optimization) {  // :29-43
    mean_in_world(w);
    const float d0 = w[0] - initial[0], d1 = w[1] - initial[1], d2 = w[2] - initial[2];
    if ((d0 * d0 + d1 * d1) + d2 * d2 < cfg.maximum_distance_geometry_meters_squared) {
      smoother_apply(world_in_local_map12, w, coords_in_local_map);
      for (int i = 0; i < 3; ++i) state_world[i] = w[i];
      *number_of_optimizations = n_meas;
      return true;
    }
    return false;
  }
  const float* K = cfg.K;
  float prev = 0;
  int n_inliers = 0;
  for (unsigned it = 0; it < cfg.maximum_number_of_iterations; ++it) {  // :49-123
    float H[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, b[3] = {0, 0, 0}, total = 0;
    int n_outliers = 0;
    for (int k = 0; k < n_meas; ++k) {
      const float* T = frames_world_in_sensor + 12 * hist_frame[k];
      float cam[3];
      smoother_apply(T, w, cam);
      if (cam[2] <= 0) {
        ++n_outliers;
        continue;
      }
      float ph[3];
      for (int i = 0; i < 3; ++i) ph[i] = (K[3 * i] * cam[0] + K[3 * i + 1] * cam[1]) + K[3 * i + 2] * cam[2];
      const float c = ph[2], inv_c = 1.0f / c, inv_c2 = inv_c * inv_c;
      const float e[3] = {ph[0] / c - hist_uv[2 * k], ph[1] / c - hist_uv[2 * k + 1], c - hist_point_in_camera[3 * k + 2]};
      float om[3] = {1.0f, 1.0f, 10.0f};  // :56-57
      const float e2 = (e[0] * om[0] * e[0] + e[1] * om[1] * e[1]) + e[2] * om[2] * e[2];
      total += e2;
      if (e2 > cfg.maximum_reprojection_error_pixels_squared) {  // saturated kernel (:80-84)
        const float s = cfg.maximum_reprojection_error_pixels_squared / e2;
        for (int i = 0; i < 3; ++i) om[i] *= s;
        ++n_outliers;
      }
      float Jl[9], J[9];
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) Jl[3 * i + j] = (K[3 * i] * T[j] + K[3 * i + 1] * T[4 + j]) + K[3 * i + 2] * T[8 + j];  // K * R (:87)
      const float a = -ph[0] * inv_c2, bb = -ph[1] * inv_c2;
      for (int j = 0; j < 3; ++j) {
        J[j] = inv_c * Jl[j] + a * Jl[6 + j];
        J[3 + j] = inv_c * Jl[3 + j] + bb * Jl[6 + j];
        J[6 + j] = Jl[6 + j];
      }
      for (int i = 0; i < 3; ++i) {
        const float t0 = J[i] * om[0], t1 = J[3 + i] * om[1], t2 = J[6 + i] * om[2];  // row i of J^T Omega
        for (int j = 0; j < 3; ++j) H[3 * i + j] += (t0 * J[j] + t1 * J[3 + j]) + t2 * J[6 + j];
        b[i] += (t0 * e[0] + t1 * e[1]) + t2 * e[2];
      }
    }
    const float nb[3] = {-b[0], -b[1], -b[2]};
    float dx[3];
    full_piv_lu_solve3(H, nb, dx);
    for (int i = 0; i < 3; ++i) w[i] += dx[i];
    n_inliers = n_meas - n_outliers;
    if (std::fabs(total - prev) < cfg.convergence_criterion_minimum_chi2_delta) break;
    prev = total;
  }
  bool inlier = false;
  if (n_inliers > *number_of_optimizations) {  // :126-131
    for (int i = 0; i < 3; ++i) state_world[i] = w[i];
    *number_of_optimizations += 1;
    inlier = true;
  } else {  // :134-139
    mean_in_world(w);
    for (int i = 0; i < 3; ++i) state_world[i] = w[i];
  }
  smoother_apply(world_in_local_map12, w, coords_in_local_map);  // :142
  return inlier;
}

// -----------------------------------------------------------------------------
// MergerProjective_::compute binning  (.../mapping/mergers/merger_projective_impl.cpp)
//   the sequential walk of the reference, map of maps included: which correspondences reach _updatePoint (:61-135)
//   and which measurements form points_in_image_to_add (:205-253).  Measurements are [n][dim] floats, dim 4 = stereo
//   (uL, vL, uR, vR), dim 3 = (u, v, depth).
// -----------------------------------------------------------------------------
enum MergerKind { MERGER_BASE = 0, MERGER_STEREO = 1, MERGER_DEPTH = 2 };

struct MergerConfig {
  int canvas_rows = 0, canvas_cols = 0;                   // param_projector (merger_projective.h:36-40)
  unsigned number_of_row_bins = 10, number_of_col_bins = 30;  // merger_projective.h:46-55
  float maximum_distance_appearance = 50;                 // merger_projective.h:41-45
  bool enable_binning = true;                             // MergerCorrespondence_ (srrg2_slam_interfaces)
  int kind = MERGER_STEREO;
};

using MergerBinMap = std::unordered_map<size_t, std::unordered_map<size_t, size_t>>;  // merger_projective.h:11-12

// merger_projective.h:89-92 / merger_projective_rigid_stereo_impl.cpp:42-52 / merger_projective_depth_ekf_impl.cpp:44-52
static inline bool merger_is_better_for_addition(int kind, const float* a, const float* b) {
  if (kind == MERGER_STEREO) return (a[0] - a[2]) > (b[0] - b[2]);
  if (kind == MERGER_DEPTH) return a[2] < b[2];
  return false;
}

// update pass (:61-135): selected[c] = the reference reaches _updatePoint for correspondence c.  Returns their number.
static inline int merger_select_updates(const MergerConfig& cfg, const float* meas, int dim, const int* corr_moving,
                                        const float* corr_response, int n_corr, unsigned char* selected, MergerBinMap& occupied) {
  const float row_w = static_cast<float>(cfg.canvas_rows) / static_cast<float>(cfg.number_of_row_bins);  // :31-34
  const float col_w = static_cast<float>(cfg.canvas_cols) / static_cast<float>(cfg.number_of_col_bins);
  int n = 0;
  for (int c = 0; c < n_corr; ++c) {
    selected[c] = 0;
    if (corr_response[c] > cfg.maximum_distance_appearance) continue;  // :72-75
    const size_t index_measurement = (size_t) corr_moving[c];
    const float* m = meas + index_measurement * dim;
    const size_t bin_row = std::round(m[1] / row_w), bin_col = std::round(m[0] / col_w);  // :82-83
    if (cfg.enable_binning) {  // :88-121
      auto it_row = occupied.find(bin_row);
      if (it_row != occupied.end()) {
        if (it_row->second.find(bin_col) == it_row->second.end()) it_row->second.insert(std::make_pair(bin_col, index_measurement));
        else continue;  // skip multiple merges in the same bin
      } else {
        std::unordered_map<size_t, size_t> column_candidates;
        column_candidates.insert(std::make_pair(bin_col, index_measurement));
        occupied.insert(std::make_pair(bin_row, column_candidates));
      }
    }
    selected[c] = 1;  // :125 _updatePoint
    ++n;
  }
  return n;
}

// addition pass (:205-253): the source measurement of every entry of points_in_image_to_add, in its order
static inline void merger_select_additions(const MergerConfig& cfg, const float* meas, int dim, int n_meas, const MergerBinMap& occupied,
                                           std::vector<int>& winners) {
  winners.clear();
  if (!cfg.enable_binning) {  // :250-253
    for (int i = 0; i < n_meas; ++i) winners.push_back(i);
    return;
  }
  const float row_w = static_cast<float>(cfg.canvas_rows) / static_cast<float>(cfg.number_of_row_bins);
  const float col_w = static_cast<float>(cfg.canvas_cols) / static_cast<float>(cfg.number_of_col_bins);
  MergerBinMap addition;
  for (int i = 0; i < n_meas; ++i) {
    const float* m = meas + (size_t) i * dim;
    const size_t bin_row = std::round(m[1] / row_w), bin_col = std::round(m[0] / col_w);  // :215-216
    auto it_tracked = occupied.find(bin_row);  // :221-227
    if (it_tracked != occupied.end() && it_tracked->second.find(bin_col) != it_tracked->second.end()) continue;
    auto it_row = addition.find(bin_row);
    if (it_row != addition.end()) {
      auto it_col = it_row->second.find(bin_col);
      if (it_col != it_row->second.end()) {
        const size_t slot = it_col->second;  // :233-240: replace the occupant when the candidate is better
        if (merger_is_better_for_addition(cfg.kind, m, meas + (size_t) winners[slot] * dim)) winners[slot] = i;
      } else {
        it_row->second.insert(std::make_pair(bin_col, winners.size()));
        winners.push_back(i);
      }
    } else {
      std::unordered_map<size_t, size_t> column_candidates;
      column_candidates.insert(std::make_pair(bin_col, winners.size()));
      addition.insert(std::make_pair(bin_row, column_candidates));
      winners.push_back(i);
    }
  }
}

}  // namespace pslam_oracle
