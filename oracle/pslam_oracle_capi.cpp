// =============================================================================
// pslam_oracle_capi.cpp -- C entry points of the CPU ORACLE (TEST INFRASTRUCTURE).
// Loaded with ctypes by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs ONLY.  See pslam_oracle.hpp for the restatement and citations.
// Build: make -C oracle   (g++ -O2 -ffp-contract=off; -ffp-contract=off keeps the fp32
// projection arithmetic free of FMA contraction so that it is the same on every host).
// =============================================================================
#include <atomic>
#include <chrono>
#include <cstdio>
#include <thread>

#include "pslam_oracle.hpp"
#include "pslam_oracle_solver.hpp"
#include "pslam_oracle_mapping.hpp"

using namespace pslam_oracle;

namespace {

Cloud make_cloud(int n, const float* coords, int dim, const uint8_t* desc, const float* intensity) {
  Cloud c((size_t) n);
  for (int i = 0; i < n; ++i) {
    c[i].x = coords[dim * i];
    c[i].y = coords[dim * i + 1];
    if (dim > 2) c[i].z = coords[dim * i + 2];
    if (dim > 3) c[i].w = coords[dim * i + 3];
    if (intensity) c[i].intensity = intensity[i];
    if (desc) std::memcpy(c[i].desc.b, desc + 32 * (size_t) i, 32);
  }
  return c;
}

int write_corr(const CorrespondenceVector& v, int cap, int* fi, int* mi, float* d) {
  const int n = (int) v.size();
  for (int i = 0; i < n && i < cap; ++i) {
    fi[i] = v[i].fixed_idx;
    mi[i] = v[i].moving_idx;
    d[i] = v[i].response;
  }
  return n;
}

ExtractConfig make_extract_cfg(const float* c) {
  ExtractConfig e;
  e.detector_threshold = c[0];
  e.enable_nms = (int) c[1];
  e.target_number_of_keypoints = (int) c[2];
  e.detectors_horizontal = (int) c[3];
  e.detectors_vertical = (int) c[4];
  return e;
}

int write_cloud(const Cloud& f, int cap, int dim, float* coords, float* intensity, uint8_t* desc) {
  const int n = (int) f.size();
  for (int i = 0; i < n && i < cap; ++i) {
    coords[dim * i] = f[i].x;
    coords[dim * i + 1] = f[i].y;
    if (dim > 2) coords[dim * i + 2] = f[i].z;
    if (dim > 3) coords[dim * i + 3] = f[i].w;
    if (intensity) intensity[i] = f[i].intensity;
    if (desc) std::memcpy(desc + 32 * (size_t) i, f[i].desc.b, 32);
  }
  return n;
}

template <typename S>
Pose<S> pose_from(const S* m12) {  // row-major 3x4 [R|t]
  Pose<S> p;
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) p.R[3 * i + j] = m12[4 * i + j];
    p.t[i] = m12[4 * i + 3];
  }
  return p;
}
template <typename S>
void pose_to(const Pose<S>& p, S* m12) {
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) m12[4 * i + j] = p.R[3 * i + j];
    m12[4 * i + 3] = p.t[i];
  }
}

// lcfg = {kind, K[9], cols, rows, baseline[3], mean_disparity, robustifier, chi_threshold} (18)
template <typename S>
LinearizeConfig<S> make_lcfg(const double* l) {
  LinearizeConfig<S> c;
  c.kind = (int) l[0];
  for (int i = 0; i < 9; ++i) c.K[i] = (S) l[1 + i];
  c.image_cols = (S) l[10];
  c.image_rows = (S) l[11];
  for (int i = 0; i < 3; ++i) c.baseline[i] = (S) l[12 + i];
  c.mean_disparity = (S) l[15];
  c.robustifier = (int) l[16];
  c.chi_threshold = (S) l[17];
  return c;
}

}  // namespace

extern "C" {

int orc_fast_detect(const uint8_t* img, int rows, int cols, int stride, int thr, int nms,
                    const uint8_t* mask, int cap, float* xy, float* response) {
  std::vector<KeyPoint> k;
  fast_detect(img, rows, cols, stride, thr, nms != 0, mask, cols, k);
  for (size_t i = 0; i < k.size() && (int) i < cap; ++i) {
    xy[2 * i] = k[i].x;
    xy[2 * i + 1] = k[i].y;
    response[i] = k[i].response;
  }
  return (int) k.size();
}

int orc_blur7(const uint8_t* img, int rows, int cols, int stride, uint8_t* out) {
  std::vector<uint8_t> b;
  blur7(img, rows, cols, stride, b);
  std::memcpy(out, b.data(), b.size());
  return 0;
}

// cfg5 = {detector_threshold, enable_nms, target_number_of_keypoints, nh, nv}
int orc_extract_binned(const uint8_t* img, int rows, int cols, int stride, const float* cfg5,
                       const uint8_t* mask, int cap, float* xy, float* response, float* intensity,
                       uint8_t* desc) {
  Cloud f;
  std::vector<float> resp;
  extract_binned(img, rows, cols, stride, make_extract_cfg(cfg5), mask, cols, f, &resp);
  write_cloud(f, cap, 2, xy, intensity, desc);
  for (size_t i = 0; i < resp.size() && (int) i < cap; ++i) response[i] = resp[i];
  return (int) f.size();
}

// keypoints after binning but before the ORB border filter (for stage-level parity)
int orc_detect_binned(const uint8_t* img, int rows, int cols, int stride, const float* cfg5,
                      int cap, float* xy, float* response) {
  ExtractConfig cfg = make_extract_cfg(cfg5);
  BinGrid grid;
  grid.init(rows, cols, cfg.detectors_horizontal, cfg.detectors_vertical,
            cfg.target_number_of_keypoints);
  std::vector<KeyPoint> k;
  fast_detect(img, rows, cols, stride, (int) cfg.detector_threshold, cfg.enable_nms != 0, nullptr,
              0, k);
  bin_select(grid, k);
  for (size_t i = 0; i < k.size() && (int) i < cap; ++i) {
    xy[2 * i] = k[i].x;
    xy[2 * i + 1] = k[i].y;
    response[i] = k[i].response;
  }
  return (int) k.size();
}

int orc_bin_lut(int rows, int cols, int nh, int nv, int target, int32_t* lut, int64_t* quota) {
  BinGrid g;
  g.init(rows, cols, nh, nv, target);
  for (int r = 0; r < rows; ++r)
    for (int c = 0; c < cols; ++c) lut[(size_t) r * cols + c] = (int32_t) g.region(r, c);
  *quota = (int64_t) g.quota;
  return 0;
}

// libstdc++ std::sort with the binned extractor's comparator (response descending) applied to
// (key,payload) items; returns the permuted payloads.  Test helper for the device introsort.
int orc_std_sort_desc(int n, const float* keys, int32_t* payload_inout) {
  struct It {
    float k;
    int32_t p;
  };
  std::vector<It> v((size_t) n);
  for (int i = 0; i < n; ++i) v[i] = It{keys[i], payload_inout[i]};
  std::sort(v.begin(), v.end(), [](const It& a, const It& b) { return a.k > b.k; });
  for (int i = 0; i < n; ++i) payload_inout[i] = v[i].p;
  return 0;
}
int orc_std_sort_asc(int n, const float* keys, int32_t* payload_inout) {
  struct It {
    float k;
    int32_t p;
  };
  std::vector<It> v((size_t) n);
  for (int i = 0; i < n; ++i) v[i] = It{keys[i], payload_inout[i]};
  std::sort(v.begin(), v.end(), [](const It& a, const It& b) { return a.k < b.k; });
  for (int i = 0; i < n; ++i) payload_inout[i] = v[i].p;
  return 0;
}

int orc_hamming_matrix(int nf, const uint8_t* df, int nm, const uint8_t* dm, int32_t* out) {
  for (int f = 0; f < nf; ++f)
    for (int m = 0; m < nm; ++m) {
      Descriptor a, b;
      std::memcpy(a.b, df + 32 * (size_t) f, 32);
      std::memcpy(b.b, dm + 32 * (size_t) m, 32);
      out[(size_t) f * nm + m] = hamming256(a, b);
    }
  return 0;
}

// per fixed row: (best distance, second distance, argmin) over all moving, first index wins ties.
// second == INT32_MAX / best_idx == -1 when absent.  Sharded by rows over `threads`.
int orc_bf_best2(int nf, const uint8_t* df, int nm, const uint8_t* dm, int threads,
                 int32_t* best, int32_t* second, int32_t* best_idx) {
  auto work = [&](int f0, int f1) {
    for (int f = f0; f < f1; ++f) {
      Descriptor a;
      std::memcpy(a.b, df + 32 * (size_t) f, 32);
      int b1 = INT32_MAX, b2 = INT32_MAX, bi = -1;
      for (int m = 0; m < nm; ++m) {
        Descriptor b;
        std::memcpy(b.b, dm + 32 * (size_t) m, 32);
        const int d = hamming256(a, b);
        if (d < b1) {
          b2 = b1;
          b1 = d;
          bi = m;
        } else if (d < b2) {
          b2 = d;
        }
      }
      best[f] = b1;
      second[f] = b2;
      best_idx[f] = bi;
    }
  };
  if (threads <= 1) {
    work(0, nf);
  } else {
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t)
      pool.emplace_back(work, (int) ((int64_t) nf * t / threads), (int) ((int64_t) nf * (t + 1) / threads));
    for (auto& th : pool) th.join();
  }
  return 0;
}

int orc_match_epipolar(int nf, const float* xy_f, const uint8_t* desc_f, int nm, const float* xy_m,
                       const uint8_t* desc_m, float max_dist, float ratio, int max_disp,
                       int thickness, int cap, int* fi, int* mi, float* dist) {
  Cloud F = make_cloud(nf, xy_f, 2, desc_f, nullptr), M = make_cloud(nm, xy_m, 2, desc_m, nullptr);
  EpipolarConfig c;
  c.maximum_descriptor_distance = max_dist;
  c.maximum_distance_ratio_to_second_best = ratio;
  c.maximum_disparity_pixels = max_disp;
  c.epipolar_line_thickness_pixels = thickness;
  CorrespondenceVector out;
  match_epipolar(F, M, c, out);
  return write_corr(out, cap, fi, mi, dist);
}

int orc_match_bruteforce(int nf, const uint8_t* desc_f, int nm, const uint8_t* desc_m,
                         float max_dist, float ratio, int cap, int* fi, int* mi, float* dist) {
  std::vector<float> zf(2 * (size_t) std::max(nf, 1), 0.f), zm(2 * (size_t) std::max(nm, 1), 0.f);
  Cloud F = make_cloud(nf, zf.data(), 2, desc_f, nullptr), M = make_cloud(nm, zm.data(), 2, desc_m, nullptr);
  BruteforceConfig c;
  c.maximum_descriptor_distance = max_dist;
  c.maximum_distance_ratio_to_second_best = ratio;
  CorrespondenceVector out;
  match_bruteforce(F, M, c, out);
  return write_corr(out, cap, fi, mi, dist);
}

// RawDataPreprocessorStereoProjective::compute; matcher: 0 epipolar, 1 bruteforce.
// mcfg4 = {max_dist, ratio, max_disparity, thickness}.  Output: 4 floats per point.
int orc_stereo_adaptor(const uint8_t* left, const uint8_t* right, int rows, int cols, int stride,
                       const float* cfg5, int matcher, const float* mcfg4, int cap, float* uvuv,
                       float* intensity, uint8_t* desc, int* n_left, int* n_right, int* n_matches) {
  Cloud L, R, meas;
  ExtractConfig ec = make_extract_cfg(cfg5);
  extract_binned(left, rows, cols, stride, ec, nullptr, 0, L);
  extract_binned(right, rows, cols, stride, ec, nullptr, 0, R);
  CorrespondenceVector m;
  if (matcher == 0) {
    EpipolarConfig c;
    c.maximum_descriptor_distance = mcfg4[0];
    c.maximum_distance_ratio_to_second_best = mcfg4[1];
    c.maximum_disparity_pixels = (unsigned) mcfg4[2];
    c.epipolar_line_thickness_pixels = (unsigned) mcfg4[3];
    match_epipolar(L, R, c, m);
  } else {
    BruteforceConfig c;
    c.maximum_descriptor_distance = mcfg4[0];
    c.maximum_distance_ratio_to_second_best = mcfg4[1];
    match_bruteforce(L, R, c, m);
  }
  assemble_stereo_points(L, R, m, meas);
  if (n_left) *n_left = (int) L.size();
  if (n_right) *n_right = (int) R.size();
  if (n_matches) *n_matches = (int) m.size();
  return write_cloud(meas, cap, 4, uvuv, intensity, desc);
}

// The frame-independent part of the frontend over a batch of stereo pairs (images 2p, 2p+1 of pair p at
// images + i * image_pitch), sharded over `threads` std::threads by pair -- the CPU baseline of bench.py
// (frame-level sharding over the host cores).  Per pair it is exactly
// orc_stereo_adaptor with the epipolar matcher.  counts[p] = stereo points of pair p; checksum[p] =
// FNV-1a over the pair's output floats + descriptors (lets bench.py / tests compare whole batches cheaply).
int orc_stereo_frontend_batch(const uint8_t* images, int n_pairs, int rows, int cols, int stride,
                              long long image_pitch, const float* cfg5, const float* mcfg4, int threads,
                              int* counts, unsigned long long* checksum) {
  const ExtractConfig ec = make_extract_cfg(cfg5);
  EpipolarConfig c;
  c.maximum_descriptor_distance = mcfg4[0];
  c.maximum_distance_ratio_to_second_best = mcfg4[1];
  c.maximum_disparity_pixels = (unsigned) mcfg4[2];
  c.epipolar_line_thickness_pixels = (unsigned) mcfg4[3];
  std::atomic<int> next(0);
  auto work = [&]() {
    for (;;) {
      const int p = next.fetch_add(1);
      if (p >= n_pairs) break;
      Cloud L, R, meas;
      extract_binned(images + (size_t) (2 * p) * image_pitch, rows, cols, stride, ec, nullptr, 0, L);
      extract_binned(images + (size_t) (2 * p + 1) * image_pitch, rows, cols, stride, ec, nullptr, 0, R);
      CorrespondenceVector m;
      match_epipolar(L, R, c, m);
      assemble_stereo_points(L, R, m, meas);
      if (counts) counts[p] = (int) meas.size();
      if (checksum) {
        unsigned long long h = 1469598103934665603ULL;
        auto mix = [&h](const void* d, size_t n) {
          const uint8_t* b = (const uint8_t*) d;
          for (size_t i = 0; i < n; ++i) h = (h ^ b[i]) * 1099511628211ULL;
        };
        for (const auto& q : meas) {
          const float v[5] = {q.x, q.y, q.z, q.w, q.intensity};
          mix(v, sizeof(v));
          mix(q.desc.b, 32);
        }
        checksum[p] = h;
      }
    }
  };
  if (threads <= 1) {
    work();
  } else {
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t) pool.emplace_back(work);
    for (auto& t : pool) t.join();
  }
  return 0;
}

// RawDataPreprocessorMonocularDepth::compute; depth image u16 (depth_is_float=0) or f32.
int orc_mono_depth_adaptor(const uint8_t* img, int rows, int cols, int stride, const void* depth,
                           int depth_is_float, int depth_stride, float depth_scale,
                           const float* cfg5, int cap, float* uvd, float* intensity, uint8_t* desc,
                           int* n_features) {
  Cloud meas;
  extract_binned(img, rows, cols, stride, make_extract_cfg(cfg5), nullptr, 0, meas);
  if (n_features) *n_features = (int) meas.size();
  if (depth_is_float)
    read_depth(meas, (const float*) depth, depth_stride, depth_scale);
  else
    read_depth(meas, (const uint16_t*) depth, depth_stride, depth_scale);
  return write_cloud(meas, cap, 3, uvd, intensity, desc);
}

int orc_triangulate(int n, const float* uvuv, const float* K9, float b_x, float min_disparity,
                    float infinity_depth, float* xyz, int* n_invalid) {
  Cloud s = make_cloud(n, uvuv, 4, nullptr, nullptr), out;
  std::vector<int> inv;
  triangulate_rectified(s, K9, b_x, min_disparity, infinity_depth, out, &inv);
  for (int i = 0; i < n; ++i) {
    xyz[3 * i] = out[i].x;
    xyz[3 * i + 1] = out[i].y;
    xyz[3 * i + 2] = out[i].z;
  }
  if (n_invalid) *n_invalid = (int) inv.size();
  return n;
}

// PointProjectorPinhole_::compute.  pose12 = moving_in_camera (== local_map_in_sensor).
int orc_project(int n, const float* xyz, const float* pose12, const float* K9, int rows, int cols,
                float rmin, float rmax, float* uvz, int* indices) {
  ProjectorConfig pc;
  std::memcpy(pc.K, K9, sizeof(pc.K));
  pc.canvas_rows = rows;
  pc.canvas_cols = cols;
  pc.range_min = rmin;
  pc.range_max = rmax;
  Pose<float> T = pose_from(pose12);
  int m = 0;
  for (int i = 0; i < n; ++i) {
    float o[3];
    if (!project_point(pc, T, xyz + 3 * i, o)) continue;
    uvz[3 * m] = o[0];
    uvz[3 * m + 1] = o[1];
    uvz[3 * m + 2] = o[2];
    indices[m] = i;
    ++m;
  }
  return m;
}

// SceneClipperProjective3D::compute  (.../mapping/scene_clipper_projective_3d.cpp:9-67):
//   projector->setCameraPose(robot_in_local_map * sensor_in_robot) (:46); projector->compute(full_scene,
//   clipped_scene, projections, global_indices) (:52); clipped_scene.transformInPlace(sensor_in_robot) when that is
//   not the identity (:60-62).  camera_in_map12 is the product of :46 (the caller multiplies, like the reference).
//   out_xyz: survivors in the sensor / robot frame, out_uvz: (u, v, depth), out_index: index into the full scene.
int orc_scene_clip(int n, const float* xyz, const float* camera_in_map12, const float* sensor_in_robot12_or_null,
                   const float* K9, int rows, int cols, float rmin, float rmax, float* out_xyz, float* out_uvz,
                   int* out_index) {
  ProjectorConfig pc;
  std::memcpy(pc.K, K9, sizeof(pc.K));
  pc.canvas_rows = rows;
  pc.canvas_cols = cols;
  pc.range_min = rmin;
  pc.range_max = rmax;
  const Pose<float> map_in_camera = pose_from(camera_in_map12).inverse();
  Pose<float> S = Pose<float>::identity();
  if (sensor_in_robot12_or_null) S = pose_from(sensor_in_robot12_or_null);
  int m = 0;
  for (int i = 0; i < n; ++i) {
    float o[3], c[3];
    if (!project_point(pc, map_in_camera, xyz + 3 * i, o)) continue;
    map_in_camera.apply(xyz + 3 * i, c);  // the point in the camera frame (second output of the projector)
    if (sensor_in_robot12_or_null) {
      float r[3];
      S.apply(c, r);
      c[0] = r[0];
      c[1] = r[1];
      c[2] = r[2];
    }
    for (int k = 0; k < 3; ++k) {
      out_xyz[3 * m + k] = c[k];
      out_uvz[3 * m + k] = o[k];
    }
    out_index[m] = i;
    ++m;
  }
  return m;
}

// ---- N3: point EKFs and LandmarkEstimatorEKF (pslam_oracle_mapping.hpp) ------------------------------------
// PointEKFBase::compute for one point, double precision (tests/test_*_point_ekf.cpp drive the filter like this).
// cam6 = fx, fy, cx, cy, bx, by; world_in_sensor12 row-major 3x4; Q 3x3; Rm E x E.
int orc_point_ekf(int kind, const double* cam6, const double* world_in_sensor12, const double* Q9, const double* meas,
                  const double* Rm, double* state3, double* cov9) {
  EkfCamera cam;
  cam.fx = cam6[0];
  cam.fy = cam6[1];
  cam.cx = cam6[2];
  cam.cy = cam6[3];
  cam.bx = cam6[4];
  cam.by = cam6[5];
  double R[9], t[3];
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) R[3 * i + j] = world_in_sensor12[4 * i + j];
    t[i] = world_in_sensor12[4 * i + 3];
  }
  if (kind == EKF_PROJECTIVE) ekf_compute<2>(cam, R, t, Q9, meas, Rm, state3, cov9);
  else if (kind == EKF_PROJECTIVE_DEPTH) ekf_compute<3>(cam, R, t, Q9, meas, Rm, state3, cov9);
  else if (kind == EKF_STEREO) ekf_compute<4>(cam, R, t, Q9, meas, Rm, state3, cov9);
  else return -1;
  return 0;
}

// LandmarkEstimatorEKF_::compute over n landmarks that share the frame's transforms (one merger pass).
int orc_landmarks_ekf_update(int n, int kind, const float* K9, const double* baseline2, double min_cov, double max_cov_norm2,
                             float max_dist2, const float* sensor_in_world12, const float* sensor_in_local_map12,
                             float* state_world, float* covariance, const float* meas, float* coords_in_local_map,
                             unsigned char* inlier) {
  LandmarkEkfConfig cfg;
  cfg.kind = kind;
  std::memcpy(cfg.K, K9, sizeof(cfg.K));
  cfg.baseline_pixels[0] = baseline2[0];
  cfg.baseline_pixels[1] = baseline2[1];
  cfg.minimum_state_element_covariance = min_cov;
  cfg.maximum_covariance_norm_squared = max_cov_norm2;
  cfg.maximum_distance_geometry_meters_squared = max_dist2;
  std::memcpy(cfg.sensor_in_world, sensor_in_world12, sizeof(cfg.sensor_in_world));
  std::memcpy(cfg.sensor_in_local_map, sensor_in_local_map12, sizeof(cfg.sensor_in_local_map));
  // LandmarkEstimatorBase_::setTransforms (landmark_estimator_base.hpp:49-58), fp32
  const Pose<float> sensor_in_world = pose_from(sensor_in_world12), sensor_in_local_map = pose_from(sensor_in_local_map12);
  const Pose<float> world_in_sensor = sensor_in_world.inverse();
  const Pose<float> world_in_local_map = sensor_in_local_map * world_in_sensor;
  const int E = ekf_measurement_dim(kind);
  int n_inliers = 0;
  for (int i = 0; i < n; ++i) {
    bool ok = false;
    float* s = state_world + 3 * i;
    float* c = covariance + 9 * i;
    float* l = coords_in_local_map + 3 * i;
    const float* m = meas + (size_t) E * i;
    if (kind == EKF_PROJECTIVE) ok = landmark_ekf_update<2>(cfg, world_in_sensor.R, world_in_sensor.t, world_in_local_map.R, world_in_local_map.t, s, c, m, l);
    else if (kind == EKF_PROJECTIVE_DEPTH) ok = landmark_ekf_update<3>(cfg, world_in_sensor.R, world_in_sensor.t, world_in_local_map.R, world_in_local_map.t, s, c, m, l);
    else ok = landmark_ekf_update<4>(cfg, world_in_sensor.R, world_in_sensor.t, world_in_local_map.R, world_in_local_map.t, s, c, m, l);
    inlier[i] = ok ? 1 : 0;
    n_inliers += ok;
  }
  return n_inliers;
}

// LandmarkEstimatorWeightedMean_::compute over n landmarks of one merger pass (fp32).
int orc_landmarks_weighted_mean_update(int n, float max_dist2, const float* sensor_in_world12, const float* sensor_in_local_map12,
                                       float* state_world, const int* number_of_optimizations, const float* landmark_in_sensor,
                                       float* coords_in_local_map, unsigned char* inlier) {
  const Pose<float> sensor_in_world = pose_from(sensor_in_world12), sensor_in_local_map = pose_from(sensor_in_local_map12);
  const Pose<float> world_in_local_map = sensor_in_local_map * sensor_in_world.inverse();
  int k = 0;
  for (int i = 0; i < n; ++i) {
    const bool ok = landmark_weighted_mean_update(sensor_in_world.R, sensor_in_world.t, world_in_local_map.R, world_in_local_map.t, max_dist2,
                                                  number_of_optimizations[i], landmark_in_sensor + 3 * i, state_world + 3 * i,
                                                  coords_in_local_map + 3 * i);
    inlier[i] = ok ? 1 : 0;
    k += ok;
  }
  return k;
}

// LandmarkEstimatorPoseBasedSmoother_::compute over n landmarks; histories in CSR form (offsets[n + 1]).
// params5 = max iterations, chi2 delta, max reprojection error^2, min measurements, max distance^2
int orc_landmarks_smoother_update(int n, const float* K9, const float* params5, int n_frames, const float* frames_sensor_in_world,
                                  const float* sensor_in_world12, const float* sensor_in_local_map12, const int* offsets,
                                  const int* hist_frame, const float* hist_uv, const float* hist_point_in_camera, float* state_world,
                                  int* number_of_optimizations, float* coords_in_local_map, unsigned char* inlier) {
  SmootherConfig cfg;
  std::memcpy(cfg.K, K9, sizeof(cfg.K));
  cfg.maximum_number_of_iterations = (unsigned) params5[0];
  cfg.convergence_criterion_minimum_chi2_delta = params5[1];
  cfg.maximum_reprojection_error_pixels_squared = params5[2];
  cfg.minimum_number_of_measurements_for_optimization = (unsigned) params5[3];
  cfg.maximum_distance_geometry_meters_squared = params5[4];
  std::vector<float> wis((size_t) 12 * n_frames);
  for (int f = 0; f < n_frames; ++f) {
    const Pose<float> inv = pose_from(frames_sensor_in_world + 12 * f).inverse();
    for (int i = 0; i < 3; ++i) {
      for (int j = 0; j < 3; ++j) wis[12 * f + 4 * i + j] = inv.R[3 * i + j];
      wis[12 * f + 4 * i + 3] = inv.t[i];
    }
  }
  const Pose<float> wl = pose_from(sensor_in_local_map12) * pose_from(sensor_in_world12).inverse();
  float wl12[12];
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) wl12[4 * i + j] = wl.R[3 * i + j];
    wl12[4 * i + 3] = wl.t[i];
  }
  int k = 0;
  for (int i = 0; i < n; ++i) {
    const int o = offsets[i], m = offsets[i + 1] - o;
    const bool ok = landmark_smoother_update(cfg, frames_sensor_in_world, wis.data(), wl12, m, hist_frame + o, hist_uv + 2 * (size_t) o,
                                             hist_point_in_camera + 3 * (size_t) o, state_world + 3 * i, number_of_optimizations + i,
                                             coords_in_local_map + 3 * i);
    inlier[i] = ok ? 1 : 0;
    k += ok;
  }
  return k;
}

// ---- stateful projective finder ----------------------------------------------
struct OrcFinder {
  ProjectiveFinder f;
  Cloud fixed, moving;
};

// fcfg12 = {max_desc_dist, ratio, min_matching_ratio, min_desc_dist, desc_step, max_radius,
//           min_radius, radius_step, min_iterations, max_change_norm, iters_per_projection, shape}
// pcfg13 = {K[9], rows, cols, range_min, range_max}
void* orc_pf_create(const float* fcfg12, const float* pcfg13) {
  OrcFinder* h = new OrcFinder();
  ProjectiveFinderConfig& c = h->f.cfg;
  c.maximum_descriptor_distance = fcfg12[0];
  c.maximum_distance_ratio_to_second_best = fcfg12[1];
  c.minimum_matching_ratio = fcfg12[2];
  c.minimum_descriptor_distance = fcfg12[3];
  c.descriptor_distance_step_size_pixels = fcfg12[4];
  c.maximum_search_radius_pixels = (unsigned) fcfg12[5];
  c.minimum_search_radius_pixels = (unsigned) fcfg12[6];
  c.search_radius_step_size_pixels = (unsigned) fcfg12[7];
  c.minimum_number_of_iterations = (unsigned) fcfg12[8];
  c.maximum_estimate_change_norm_for_convergence = fcfg12[9];
  c.number_of_solver_iterations_per_projection = (unsigned) fcfg12[10];
  c.shape = (int) fcfg12[11];
  std::memcpy(h->f.projector.K, pcfg13, 9 * sizeof(float));
  h->f.projector.canvas_rows = (int) pcfg13[9];
  h->f.projector.canvas_cols = (int) pcfg13[10];
  h->f.projector.range_min = pcfg13[11];
  h->f.projector.range_max = pcfg13[12];
  return h;
}
void orc_pf_destroy(void* h) { delete (OrcFinder*) h; }
void orc_pf_set_fixed(void* hh, int n, const float* coords, int dim, const uint8_t* desc) {
  OrcFinder* h = (OrcFinder*) hh;
  h->fixed = make_cloud(n, coords, dim, desc, nullptr);
  h->f.setFixed(&h->fixed);
}
void orc_pf_set_moving(void* hh, int n, const float* xyz, const uint8_t* desc) {
  OrcFinder* h = (OrcFinder*) hh;
  h->moving = make_cloud(n, xyz, 3, desc, nullptr);
  h->f.setMoving(&h->moving);
}
void orc_pf_set_estimate(void* hh, const float* pose12) {
  ((OrcFinder*) hh)->f.setEstimate(pose_from(pose12));
}
void orc_pf_get_estimate(void* hh, float* pose12) {
  pose_to(((OrcFinder*) hh)->f.local_map_in_sensor, pose12);
}
void orc_pf_set_radius(void* hh, int r) { ((OrcFinder*) hh)->f.setSearchRadiusPixels(r); }
void orc_pf_set_descriptor_distance(void* hh, float d) {
  ((OrcFinder*) hh)->f.setDescriptorDistance(d);
}
int orc_pf_compute(void* hh, int cap, int* fi, int* mi, float* dist) {
  OrcFinder* h = (OrcFinder*) hh;
  h->f.compute();
  return write_corr(h->f.correspondences, cap, fi, mi, dist);
}
// state6 = {radius, descriptor_distance, iteration, converged, searches, n_projected}
void orc_pf_state(void* hh, float* state6) {
  OrcFinder* h = (OrcFinder*) hh;
  state6[0] = (float) h->f.search_radius_pixels;
  state6[1] = h->f.descriptor_distance;
  state6[2] = (float) h->f.current_iteration;
  state6[3] = h->f.has_converged ? 1.f : 0.f;
  state6[4] = (float) h->f.number_of_searches;
  state6[5] = (float) h->f.points_in_image.size();
}
// window candidates of the last full search: 5 ints per projected point
// {moving_idx, fixed_best, dist_best, fixed_second, dist_second}; lattice order too.
int orc_pf_candidates(void* hh, int cap, int* out5) {
  OrcFinder* h = (OrcFinder*) hh;
  const auto& v = h->f.last_candidates;
  for (size_t i = 0; i < v.size() && (int) i < cap; ++i) {
    out5[5 * i] = v[i].moving_idx;
    out5[5 * i + 1] = v[i].fixed_best;
    out5[5 * i + 2] = (int) v[i].dist_best;
    out5[5 * i + 3] = v[i].fixed_second;
    out5[5 * i + 4] = (int) v[i].dist_second;
  }
  return (int) v.size();
}
int orc_pf_lattice(void* hh, int cap, int* index) {
  OrcFinder* h = (OrcFinder*) hh;
  const auto& v = h->f.database_fixed;
  for (size_t i = 0; i < v.size() && (int) i < cap; ++i) index[i] = v[i].index;
  return (int) v.size();
}

// ---- SE3 factors, H/b, GN -----------------------------------------------------
// fp64 linearisation.  stats4 = {chi_total, inliers, outliers, suppressed}
int orc_linearize_f64(const double* lcfg18, const double* pose12, int n_moving,
                      const double* moving_xyz, int n_fixed, const double* fixed_meas,
                      int fixed_dim, int n_corr, const int* cf, const int* cm,
                      const double* info_diag, double* H36, double* b6, double* stats4) {
  (void) n_moving;
  (void) n_fixed;
  LinearSystem<double> sys;
  linearize(make_lcfg<double>(lcfg18), pose_from(pose12), moving_xyz, fixed_meas, fixed_dim, cf,
            cm, n_corr, info_diag, sys);
  std::memcpy(H36, sys.H, sizeof(sys.H));
  std::memcpy(b6, sys.b, sizeof(sys.b));
  stats4[0] = sys.chi_total;
  stats4[1] = sys.inliers;
  stats4[2] = sys.outliers;
  stats4[3] = sys.suppressed;
  return 0;
}

// as orc_linearize_f64, plus: prior48 (12 doubles predicted pose + 36 doubles information; NULL = no prior
// factor) summed into H, b (its chi goes to stats5[4], not into the slice's chi), and the per-correspondence
// factor status (0 inlier, 1 kernelized, 2 suppressed; NULL = not wanted)
int orc_linearize_ex_f64(const double* lcfg18, const double* pose12, int n_moving,
                         const double* moving_xyz, int n_fixed, const double* fixed_meas,
                         int fixed_dim, int n_corr, const int* cf, const int* cm,
                         const double* info_diag, const double* prior48, uint8_t* status,
                         double* H36, double* b6, double* stats5) {
  (void) n_moving;
  (void) n_fixed;
  LinearSystem<double> sys;
  const Pose<double> X = pose_from(pose12);
  linearize(make_lcfg<double>(lcfg18), X, moving_xyz, fixed_meas, fixed_dim, cf, cm, n_corr,
            info_diag, sys, status);
  stats5[4] = prior48 ? pose_prior_accumulate(pose_from(prior48), prior48 + 12, X, sys) : 0.0;
  std::memcpy(H36, sys.H, sizeof(sys.H));
  std::memcpy(b6, sys.b, sizeof(sys.b));
  stats5[0] = sys.chi_total;
  stats5[1] = sys.inliers;
  stats5[2] = sys.outliers;
  stats5[3] = sys.suppressed;
  return 0;
}

// the reference's own precision (fp32 accumulate), for reporting the fp32-vs-fp64 gap
int orc_linearize_f32(const double* lcfg18, const double* pose12, int n_moving,
                      const double* moving_xyz, int n_fixed, const double* fixed_meas,
                      int fixed_dim, int n_corr, const int* cf, const int* cm,
                      const double* info_diag, double* H36, double* b6, double* stats4) {
  std::vector<float> mv(3 * (size_t) n_moving), fx((size_t) fixed_dim * n_fixed),
    inf(3 * (size_t) n_fixed);
  for (size_t i = 0; i < mv.size(); ++i) mv[i] = (float) moving_xyz[i];
  for (size_t i = 0; i < fx.size(); ++i) fx[i] = (float) fixed_meas[i];
  for (size_t i = 0; i < inf.size(); ++i) inf[i] = (float) info_diag[i];
  float p[12];
  for (int i = 0; i < 12; ++i) p[i] = (float) pose12[i];
  LinearSystem<float> sys;
  linearize(make_lcfg<float>(lcfg18), pose_from(p), mv.data(), fx.data(), fixed_dim, cf, cm,
            n_corr, inf.data(), sys);
  for (int i = 0; i < 36; ++i) H36[i] = sys.H[i];
  for (int i = 0; i < 6; ++i) b6[i] = sys.b[i];
  stats4[0] = sys.chi_total;
  stats4[1] = sys.inliers;
  stats4[2] = sys.outliers;
  stats4[3] = sys.suppressed;
  return 0;
}

// (H + damping I) dx = -b ; pose <- pose * v2t(dx).  returns 0 ok, -1 not SPD
int orc_gn_step_f64(const double* H36, const double* b6, double damping, double* pose12,
                    double* dx6) {
  LinearSystem<double> sys;
  std::memcpy(sys.H, H36, sizeof(sys.H));
  std::memcpy(sys.b, b6, sizeof(sys.b));
  Pose<double> X = pose_from(pose12);
  if (!gn_step(sys, damping, X, dx6)) return -1;
  pose_to(X, pose12);
  return 0;
}

void orc_t2tnq_f64(const double* pose12, double* v6) { t2tnq(pose_from(pose12), v6); }
void orc_pose_inverse_f64(const double* pose12, double* out12) {
  pose_to(pose_from(pose12).inverse(), out12);
}
void orc_pose_mul_f64(const double* a12, const double* b12, double* out12) {
  pose_to(pose_from(a12) * pose_from(b12), out12);
}


// MergerProjective_::compute binning.  cfg7 = canvas rows, canvas cols, row bins, col bins, enable_binning, kind, (unused);
// the blocked bins travel as a bitmap over the (row bins + 1) x (col bins + 1) grid, bit = bin_row * (col bins + 1) + bin_col
// (-1 when a bin leaves that grid: the reference asserts it cannot, merger_projective_impl.cpp:84-85)
static MergerConfig merger_cfg_from(const int* cfg6, float max_dist) {
  MergerConfig m;
  m.canvas_rows = cfg6[0];
  m.canvas_cols = cfg6[1];
  m.number_of_row_bins = (unsigned) cfg6[2];
  m.number_of_col_bins = (unsigned) cfg6[3];
  m.enable_binning = cfg6[4] != 0;
  m.kind = cfg6[5];
  m.maximum_distance_appearance = max_dist;
  return m;
}

int orc_merger_select_updates(const int* cfg6, float max_dist, const float* meas, int dim, const int* corr_moving,
                              const float* corr_response, int n_corr, unsigned char* selected, unsigned* occupied_words, int n_words) {
  const MergerConfig cfg = merger_cfg_from(cfg6, max_dist);
  MergerBinMap occupied;
  const int n = merger_select_updates(cfg, meas, dim, corr_moving, corr_response, n_corr, selected, occupied);
  for (int w = 0; w < n_words; ++w) occupied_words[w] = 0;
  for (const auto& row : occupied)
    for (const auto& col : row.second) {
      if (row.first > cfg.number_of_row_bins || col.first > cfg.number_of_col_bins) return -1;
      const size_t b = row.first * (cfg.number_of_col_bins + 1) + col.first;
      occupied_words[b >> 5] |= 1u << (b & 31);
    }
  return n;
}

int orc_merger_select_additions(const int* cfg6, const float* meas, int dim, int n_meas, const unsigned* occupied_words, int* winners) {
  const MergerConfig cfg = merger_cfg_from(cfg6, 0.f);
  MergerBinMap occupied;
  if (occupied_words)
    for (size_t r = 0; r <= cfg.number_of_row_bins; ++r)
      for (size_t c = 0; c <= cfg.number_of_col_bins; ++c) {
        const size_t b = r * (cfg.number_of_col_bins + 1) + c;
        if ((occupied_words[b >> 5] >> (b & 31)) & 1u) occupied[r][c] = 0;
      }
  std::vector<int> w;
  merger_select_additions(cfg, meas, dim, n_meas, occupied, w);
  for (size_t k = 0; k < w.size(); ++k) winners[k] = w[k];
  return (int) w.size();
}

}  // extern "C"
