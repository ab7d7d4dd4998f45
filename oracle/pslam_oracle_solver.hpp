// =============================================================================
// pslam_oracle_solver.hpp -- CPU ORACLE (TEST INFRASTRUCTURE, NOT THE PRODUCT)
//
// Projective correspondence finders (in-tree reference code) and the stage 3/4
// arithmetic that lives in the reference's un-vendored dependencies:
//   srrg2_core   : PointProjectorPinhole_, geometry3d::t2tnq / v2t
//   srrg2_solver : SE3RectifiedStereoProjectiveErrorFactor, SE3ProjectiveDepthErrorFactor,
//                  SE3ProjectiveErrorFactor, RobustifierSaturated/Clamp,
//                  FactorCorrespondenceDriven_::compute (H,b accumulation), IterationAlgorithmGN
// Neither dependency is pinned by the reference (catkin workspace HEADs,
// srrg2_proslam/readme.md:15-20).  Their arithmetic is restated here from the
// reference's call sites, in-tree analogues and tests (cited per function).
//
// PARITY UNPINNED (value level) for: factor error/Jacobian, H, b, chi, GN step.
// The reference's tests hold only pose tolerances against ground truth
// (tests/test_aligners.cpp:632-637 ...); tests/test_oracle_solver.py checks those
// tolerances, and the CUDA path is compared against THIS fp64 restatement.
// The projector semantics are pinned indirectly by one reference constant
// (90 correspondences, tests/test_correspondence_finders.cpp:499,509).
// =============================================================================
#pragma once
#include <cmath>
#include <cstdint>
#include <unordered_map>
#include <vector>

#include "pslam_oracle.hpp"

namespace pslam_oracle {

// ---- small fixed-size linear algebra (row-major) ------------------------------
template <typename S>
struct Pose {  // Isometry3: p' = R p + t
  S R[9];
  S t[3];
  static Pose identity() {
    Pose p;
    for (int i = 0; i < 9; ++i) p.R[i] = (i % 4 == 0) ? S(1) : S(0);
    p.t[0] = p.t[1] = p.t[2] = S(0);
    return p;
  }
  void apply(const S* p, S* out) const {
    for (int i = 0; i < 3; ++i)
      out[i] = ((R[3 * i] * p[0] + R[3 * i + 1] * p[1]) + R[3 * i + 2] * p[2]) + t[i];
  }
  Pose inverse() const {
    Pose q;
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) q.R[3 * i + j] = R[3 * j + i];
    for (int i = 0; i < 3; ++i)
      q.t[i] = -((q.R[3 * i] * t[0] + q.R[3 * i + 1] * t[1]) + q.R[3 * i + 2] * t[2]);
    return q;
  }
  Pose operator*(const Pose& o) const {
    Pose q;
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j)
        q.R[3 * i + j] = (R[3 * i] * o.R[j] + R[3 * i + 1] * o.R[3 + j]) + R[3 * i + 2] * o.R[6 + j];
    apply(o.t, q.t);
    return q;
  }
};

// geometry3d::t2tnq -- translation + vector part of the normalised quaternion (w >= 0)
// (used at .../correspondence_finder_projective_base_impl.cpp:182, tests/test_aligners.cpp:132)
template <typename S>
static inline void t2tnq(const Pose<S>& T, S* v6) {
  v6[0] = T.t[0];
  v6[1] = T.t[1];
  v6[2] = T.t[2];
  const S* R = T.R;
  S w, x, y, z;
  const S tr = R[0] + R[4] + R[8];
  if (tr > 0) {
    S s = std::sqrt(tr + S(1)) * 2;
    w = S(0.25) * s;
    x = (R[7] - R[5]) / s;
    y = (R[2] - R[6]) / s;
    z = (R[3] - R[1]) / s;
  } else if (R[0] > R[4] && R[0] > R[8]) {
    S s = std::sqrt(S(1) + R[0] - R[4] - R[8]) * 2;
    w = (R[7] - R[5]) / s;
    x = S(0.25) * s;
    y = (R[1] + R[3]) / s;
    z = (R[2] + R[6]) / s;
  } else if (R[4] > R[8]) {
    S s = std::sqrt(S(1) + R[4] - R[0] - R[8]) * 2;
    w = (R[2] - R[6]) / s;
    x = (R[1] + R[3]) / s;
    y = S(0.25) * s;
    z = (R[5] + R[7]) / s;
  } else {
    S s = std::sqrt(S(1) + R[8] - R[0] - R[4]) * 2;
    w = (R[3] - R[1]) / s;
    x = (R[2] + R[6]) / s;
    y = (R[5] + R[7]) / s;
    z = S(0.25) * s;
  }
  const S n = std::sqrt(w * w + x * x + y * y + z * z);
  const S sgn = (w < 0) ? S(-1) : S(1);
  v6[3] = sgn * x / n;
  v6[4] = sgn * y / n;
  v6[5] = sgn * z / n;
}

// geometry3d::v2t for the (t, normalised-quaternion vector part) chart
template <typename S>
static inline Pose<S> v2t(const S* v6) {
  Pose<S> T;
  T.t[0] = v6[0];
  T.t[1] = v6[1];
  T.t[2] = v6[2];
  S x = v6[3], y = v6[4], z = v6[5];
  const S n2 = x * x + y * y + z * z;
  S w;
  if (n2 < S(1)) {
    w = std::sqrt(S(1) - n2);
  } else {
    const S n = std::sqrt(n2);
    x /= n;
    y /= n;
    z /= n;
    w = 0;
  }
  T.R[0] = 1 - 2 * (y * y + z * z);
  T.R[1] = 2 * (x * y - z * w);
  T.R[2] = 2 * (x * z + y * w);
  T.R[3] = 2 * (x * y + z * w);
  T.R[4] = 1 - 2 * (x * x + z * z);
  T.R[5] = 2 * (y * z - x * w);
  T.R[6] = 2 * (x * z - y * w);
  T.R[7] = 2 * (y * z + x * w);
  T.R[8] = 1 - 2 * (x * x + y * y);
  return T;
}

// -----------------------------------------------------------------------------
// PointProjectorPinhole_::compute (srrg2_core).  fp32 like the
// reference; the operation order below is the definition the CUDA kernel follows.
// camera_in_world^-1 == local_map_in_sensor (projective_base_impl.cpp:158).
// Output: (u, v, z_cam), descriptor copied, index into the moving cloud.
// -----------------------------------------------------------------------------
struct ProjectorConfig {
  float K[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  int canvas_rows = 0, canvas_cols = 0;  // kitti.conf:164-170 (filled at runtime)
  float range_min = 0.1f, range_max = 1000.0f;  // kitti.conf:172-179
};

static inline bool project_point(const ProjectorConfig& pc, const Pose<float>& moving_in_camera,
                                 const float* p, float* uvz) {
  float c[3];
  moving_in_camera.apply(p, c);
  if (c[2] < pc.range_min || c[2] > pc.range_max) return false;
  const float* K = pc.K;
  const float hx = (K[0] * c[0] + K[1] * c[1]) + K[2] * c[2];
  const float hy = (K[3] * c[0] + K[4] * c[1]) + K[5] * c[2];
  const float hz = (K[6] * c[0] + K[7] * c[1]) + K[8] * c[2];
  const float u = hx / hz, v = hy / hz;
  if (u < 0 || u > (float) pc.canvas_cols || v < 0 || v > (float) pc.canvas_rows) return false;
  uvz[0] = u;
  uvz[1] = v;
  uvz[2] = c[2];
  return true;
}

static inline void project_cloud(const ProjectorConfig& pc, const Pose<float>& moving_in_camera,
                                 const Cloud& moving, Cloud& in_image, std::vector<int>& indices) {
  in_image.clear();
  indices.clear();
  for (size_t i = 0; i < moving.size(); ++i) {
    const float p[3] = {moving[i].x, moving[i].y, moving[i].z};
    float uvz[3];
    if (!project_point(pc, moving_in_camera, p, uvz)) continue;
    Feature2 f = moving[i];
    f.x = uvz[0];
    f.y = uvz[1];
    f.z = uvz[2];
    in_image.push_back(f);
    indices.push_back((int) i);
  }
}

// -----------------------------------------------------------------------------
// TriangulatorRigidStereo::triangulateRectifiedMidpoint
//   .../mapping/triangulator_rigid_stereo.cpp:59-85  (used by the test fixture chain
//   tests/fixtures.hpp:926-952 that produces the moving cloud of the projective tests)
// -----------------------------------------------------------------------------
static inline void triangulate_rectified(const Cloud& stereo, const float* K, float b_x,
                                         float min_disparity, float infinity_depth, Cloud& out,
                                         std::vector<int>* invalid = nullptr) {
  out.clear();
  const float fx = K[0], fy = K[4], cx = K[2], cy = K[5];
  for (size_t i = 0; i < stereo.size(); ++i) {
    Feature2 p = stereo[i];
    const float xL = stereo[i].x, yL = stereo[i].y, xR = stereo[i].z, yR = stereo[i].w;
    p.w = 0;
    if (xL - xR < min_disparity) {
      if (invalid) invalid->push_back((int) i);
      p.x = p.y = p.z = 0;
      out.push_back(p);
      continue;
    }
    float depth = infinity_depth;
    if (xL > xR) depth = b_x / (xL - xR);
    p.z = depth;
    p.x = 1 / fx * (xL - cx) * depth;
    p.y = 1 / fy * ((yL + yR) / 2 - cy) * depth;
    out.push_back(p);
  }
}

// -----------------------------------------------------------------------------
// Projective correspondence finders (state machine + window search + filter)
//   .../correspondence_finders/correspondence_finder_projective_base_impl.cpp:39-293
//   ..._square_impl.cpp:7-118, ..._circle_impl.cpp:7-94, ..._rhombus_impl.cpp:7-93
// -----------------------------------------------------------------------------
enum WindowShape { WINDOW_SQUARE = 0, WINDOW_CIRCLE = 1, WINDOW_RHOMBUS = 2, WINDOW_KDTREE = 3 };

struct ProjectiveFinderConfig {
  // inherited from the bruteforce base (bruteforce.h:23-37)
  float maximum_descriptor_distance = 50.0f;
  float maximum_distance_ratio_to_second_best = 0.9f;
  float minimum_matching_ratio = 0.25f;
  // projective_base.h:30-75
  float minimum_descriptor_distance = 25.0f;
  float descriptor_distance_step_size_pixels = 5;
  unsigned maximum_search_radius_pixels = 100;
  unsigned minimum_search_radius_pixels = 10;
  unsigned search_radius_step_size_pixels = 5;
  unsigned minimum_number_of_iterations = 10;
  float maximum_estimate_change_norm_for_convergence = 1e-5f;
  unsigned number_of_solver_iterations_per_projection = 25;
  int shape = WINDOW_CIRCLE;
};

struct LatticeElement {  // square.h:37-47
  int16_t row, col, index;
};

struct CandidatePair {  // per projected point: best and second best with their fixed indices
  int moving_idx = -1;
  int fixed_best = -1, fixed_second = -1;
  float dist_best = 0, dist_second = 0;
};

using CandidateMap = std::unordered_map<size_t, CorrespondenceVector>;

struct ProjectiveFinder {
  ProjectiveFinderConfig cfg;
  ProjectorConfig projector;
  // state (projective_base.h:130-154)
  bool config_changed = true;
  bool fixed_changed = true, moving_changed = true;
  size_t search_radius_pixels = 0;
  float descriptor_distance = 0;
  Pose<float> local_map_in_sensor = Pose<float>::identity();
  Pose<float> local_map_in_sensor_previous = Pose<float>::identity();
  bool has_converged = false;
  size_t current_iteration = 0;
  Cloud points_in_image;
  std::vector<int> indices_projected_to_moving;
  std::vector<LatticeElement> database_fixed;
  const Cloud* fixed = nullptr;
  const Cloud* moving = nullptr;
  CorrespondenceVector correspondences;
  // trace for parity tests: window candidates of the last full search
  std::vector<CandidatePair> last_candidates;
  int number_of_searches = 0;

  void setFixed(const Cloud* f) {
    fixed = f;
    fixed_changed = true;
  }
  void setMoving(const Cloud* m) {
    moving = m;
    moving_changed = true;
  }
  void setEstimate(const Pose<float>& T) { local_map_in_sensor = T; }
  void setSearchRadiusPixels(size_t r) {  // projective_base.h:82-85
    search_radius_pixels = r;
    config_changed = false;
  }
  void setDescriptorDistance(float d) {  // :94-97
    descriptor_distance = d;
    config_changed = false;
  }

  void initializeDatabase() {  // square_impl.cpp:7-31
    database_fixed.clear();
    database_fixed.reserve(fixed->size());
    for (size_t i = 0; i < fixed->size(); ++i) {
      database_fixed.push_back(
        LatticeElement{(int16_t)(*fixed)[i].y, (int16_t)(*fixed)[i].x, (int16_t) i});
    }
    std::sort(database_fixed.begin(), database_fixed.end(),
              [](const LatticeElement& a, const LatticeElement& b) { return a.row < b.row; });
  }

  // CorrespondenceFinderProjectiveKDTree::_findNearestNeighbors (..._kdtree_impl.cpp:28-79).  PARITY UNPINNED for this
  // variant: the candidates come from srrg2_core's KDTree<float, 2>::findNeighbors (external, approximate: it only visits
  // the leaf cluster of the query), in an order that tree defines.  Restated here as the EXACT radius query -- every fixed
  // point with squared fp32 distance < radius^2 to the projection, the superset of what the tree returns -- visited in
  // lattice order; everything after the candidate list follows the file: best initialised to maximum_descriptor_distance
  // (the PARAM, not the adaptive threshold), strict "<" updates, ONLY the best candidate is recorded (:72-78).
  CandidatePair findNearestNeighborsKDTree(const Feature2& query, int query_index) const {
    CandidatePair out;
    out.moving_idx = query_index;
    const float maximum_distance_squared = (float) (search_radius_pixels * search_radius_pixels);  // :41-42
    size_t index_best = 0;
    float best = cfg.maximum_descriptor_distance;
    float second = std::numeric_limits<float>::max();
    for (const LatticeElement& e : database_fixed) {
      const Feature2& f = (*fixed)[e.index];
      const float dx = f.x - query.x, dy = f.y - query.y;
      if (!(dx * dx + dy * dy < maximum_distance_squared)) continue;
      const float d = hamming256(f.desc, query.desc);
      if (d < best) {  // :61-68
        second = best;
        best = d;
        index_best = e.index;
      } else if (d < second) {
        second = d;
      }
    }
    if (best < cfg.maximum_descriptor_distance) {  // :72-78
      out.fixed_best = (int) index_best;
      out.dist_best = best;
    }
    return out;
  }

  // one query against the row-sorted lattice; returns best/second with fixed indices
  CandidatePair findNearestNeighbors(const Feature2& query, int query_index) const {
    if (cfg.shape == WINDOW_KDTREE) return findNearestNeighborsKDTree(query, query_index);
    CandidatePair out;
    out.moving_idx = query_index;
    const int16_t row = std::round(query.y);
    const int16_t col = std::round(query.x);
    const int16_t r = (int16_t) search_radius_pixels;
    const int16_t row_min = row - r;
    const int16_t row_max = row + r + 1;
    const int16_t col_min = col - r - 1;  // square
    const int16_t col_max = col + r + 1;
    const int32_t radius_squared = (int32_t)(search_radius_pixels * search_radius_pixels);
    size_t index_best = 0, index_second = 0;
    float best = std::numeric_limits<float>::max();
    float second = std::numeric_limits<float>::max();
    auto it = database_fixed.begin();
    while (it != database_fixed.end() && it->row < row_min) ++it;
    while (it != database_fixed.end() && it->row < row_max) {
      bool in_window;
      if (cfg.shape == WINDOW_SQUARE) {
        in_window = it->col > col_min && it->col < col_max;  // square_impl.cpp:80
      } else if (cfg.shape == WINDOW_CIRCLE) {
        const int32_t height = it->row - row;  // circle_impl.cpp:51-56
        const int32_t width = std::sqrt(radius_squared - height * height) + 1;
        in_window = it->col > col - width && it->col < col + width;
      } else {
        int16_t width = it->row - row_min + 1;  // rhombus_impl.cpp:49-55
        if (width > r) width = row_max - it->row;
        in_window = it->col > col - width && it->col < col + width;
      }
      if (in_window) {
        const float d = hamming256((*fixed)[it->index].desc, query.desc);
        if (d < best) {
          second = best;
          best = d;
          index_second = index_best;
          index_best = it->index;
        } else if (d < second) {
          second = d;
          index_second = it->index;
        }
      }
      ++it;
    }
    if (best < std::numeric_limits<float>::max()) {
      out.fixed_best = (int) index_best;
      out.dist_best = best;
      if (second < std::numeric_limits<float>::max()) {
        out.fixed_second = (int) index_second;
        out.dist_second = second;
      }
    }
    return out;
  }

  static void addCandidate(int f, int m, float d, CandidateMap& by_fixed, CandidateMap& by_moving) {
    by_fixed[f].push_back(Correspondence{f, m, d});  // projective_base_impl.cpp:7-37
    by_moving[m].push_back(Correspondence{f, m, d});
  }

  static void filter(CandidateMap& by_fixed, CandidateMap& by_moving, float max_dist,
                     float max_ratio, CorrespondenceVector& out) {  // :39-102
    out.clear();
    out.reserve(by_fixed.size());
    for (auto& kv : by_fixed) {
      CorrespondenceVector& cands = kv.second;
      size_t index_best = 0;
      float lowest = std::numeric_limits<float>::max();
      float second = std::numeric_limits<float>::max();
      for (size_t i = 0; i < cands.size(); ++i) {
        const float r = cands[i].response;
        if (r < lowest) {
          second = lowest;
          lowest = r;
          index_best = i;
        } else if (r < second) {
          second = r;
        }
      }
      if (lowest < max_dist && lowest / second < max_ratio) {
        const Correspondence& best = cands[index_best];
        const CorrespondenceVector& mc = by_moving.at(best.moving_idx);
        float l2 = std::numeric_limits<float>::max();
        size_t ib = 0;
        for (size_t i = 0; i < mc.size(); ++i)
          if (mc[i].response < l2) {
            l2 = mc[i].response;
            ib = i;
          }
        if (mc[ib].fixed_idx == best.fixed_idx) out.push_back(best);
      }
    }
  }

  // returns false on the "no recompute" paths as well; result is in `correspondences`
  void compute() {  // projective_base_impl.cpp:104-293
    if (fixed_changed || moving_changed || config_changed) {
      fixed_changed = false;
      moving_changed = false;
      if ((search_radius_pixels == 0 && descriptor_distance == 0) || config_changed) {
        search_radius_pixels = cfg.maximum_search_radius_pixels;
        descriptor_distance = cfg.minimum_descriptor_distance;
      }
      has_converged = false;
      current_iteration = 0;
      local_map_in_sensor_previous = Pose<float>::identity();
      initializeDatabase();
      config_changed = false;
    }
    if (has_converged) return;
    // projector->setCameraPose(local_map_in_sensor^-1): camera pose inverse == local_map_in_sensor
    if (current_iteration % cfg.number_of_solver_iterations_per_projection == 0 ||
        current_iteration == 1) {
      project_cloud(projector, local_map_in_sensor, *moving, points_in_image,
                    indices_projected_to_moving);
    } else {
      local_map_in_sensor_previous = local_map_in_sensor;
      ++current_iteration;
      return;
    }
    // estimate change: t2tnq(cameraPose * previous) with cameraPose = local_map_in_sensor^-1
    float v6[6];
    t2tnq(local_map_in_sensor.inverse() * local_map_in_sensor_previous, v6);
    float n2 = 0;
    for (float v : v6) n2 += v * v;
    const float estimate_change_norm = std::sqrt(n2);
    local_map_in_sensor_previous = local_map_in_sensor;

    CandidateMap by_fixed, by_moving;
    by_fixed.reserve(fixed->size());
    by_moving.reserve(moving->size());
    last_candidates.clear();
    ++number_of_searches;
    for (size_t i = 0; i < points_in_image.size(); ++i) {
      CandidatePair c = findNearestNeighbors(points_in_image[i], indices_projected_to_moving[i]);
      last_candidates.push_back(c);
      if (c.fixed_best >= 0) {
        addCandidate(c.fixed_best, c.moving_idx, c.dist_best, by_fixed, by_moving);
        if (c.fixed_second >= 0)
          addCandidate(c.fixed_second, c.moving_idx, c.dist_second, by_fixed, by_moving);
      }
    }
    CorrespondenceVector filtered;
    filter(by_fixed, by_moving, descriptor_distance, cfg.maximum_distance_ratio_to_second_best,
           filtered);
    const float matching_ratio = static_cast<float>(filtered.size()) / fixed->size();
    if (matching_ratio < cfg.minimum_matching_ratio) {
      if (search_radius_pixels < cfg.maximum_search_radius_pixels ||
          descriptor_distance > cfg.minimum_descriptor_distance) {
        search_radius_pixels = cfg.maximum_search_radius_pixels;
        descriptor_distance = cfg.minimum_descriptor_distance;
        if (matching_ratio == 0) {
          local_map_in_sensor = Pose<float>::identity();
          current_iteration = 0;
        } else {
          ++current_iteration;
        }
        return compute();
      }
    }
    correspondences.swap(filtered);
    if (estimate_change_norm < cfg.maximum_estimate_change_norm_for_convergence &&
        current_iteration > cfg.minimum_number_of_iterations) {
      has_converged = true;
      if (matching_ratio > cfg.minimum_matching_ratio) {
        search_radius_pixels =
          std::max<size_t>(search_radius_pixels - cfg.search_radius_step_size_pixels,
                           cfg.minimum_search_radius_pixels);
        descriptor_distance =
          std::min(descriptor_distance + cfg.descriptor_distance_step_size_pixels,
                   cfg.maximum_descriptor_distance);
      }
    }
    ++current_iteration;
  }
};

// -----------------------------------------------------------------------------
// SE3 projective factors + robustifier + H/b accumulation + GN step
// (srrg2_solver, external).  Templated on the scalar so the
// oracle offers the reference's fp32 and the new build's fp64.
// -----------------------------------------------------------------------------
enum FactorKind { FACTOR_STEREO = 0, FACTOR_DEPTH = 1, FACTOR_MONO = 2 };
enum RobustifierKind { ROBUST_NONE = 0, ROBUST_SATURATED = 1, ROBUST_CLAMP = 2 };
// per-factor outcome of one linearisation (srrg2_solver FactorStats::Status analogue): what the aligner's
// inlier-only runs / keep_only_inlier_correspondences read (configurations/icl.conf:55-58)
enum FactorStatus { FACTOR_INLIER = 0, FACTOR_KERNELIZED = 1, FACTOR_SUPPRESSED = 2 };

template <typename S>
struct LinearizeConfig {
  int kind = FACTOR_STEREO;
  S K[9];
  S image_cols = 0, image_rows = 0;           // setImageDim(cols, rows) aligner_slice_processor_projective.cpp:38-39
  S baseline[3] = {0, 0, 0};                  // K * t_left_in_right (:96-101); kitti: (-386.1448,0,0)
  S mean_disparity = 0;                       // >0 enables inverse-depth weighting (:107-112)
  int robustifier = ROBUST_SATURATED;
  S chi_threshold = 25;                       // kitti.conf:137-142
};

template <typename S>
struct LinearSystem {
  S H[36];
  S b[6];
  S chi_total = 0;      // sum of (robustified) chi over accumulated factors
  int inliers = 0;      // chi <= threshold
  int outliers = 0;     // kernelized
  int suppressed = 0;   // invalid projection (not accumulated)
  void clear() {
    for (S& v : H) v = 0;
    for (S& v : b) v = 0;
    chi_total = 0;
    inliers = outliers = suppressed = 0;
  }
};

// error and Jacobian of one correspondence.  fixed = measurement (uL,vL,uR,vR | u,v,depth | u,v),
// moving = 3-D point in the local map; X maps moving into the (left) camera.
// J is 3x6 row-major: columns 0..2 translation, 3..5 normalised-quaternion vector part,
// right perturbation X <- X * v2t(dx)  (App. E.2).  Returns false when the factor is invalid.
template <typename S>
static inline bool error_and_jacobian(const LinearizeConfig<S>& c, const Pose<S>& X, const S* pm,
                                      const S* z, S* e, S* J) {
  S pc[3];
  X.apply(pm, pc);
  if (pc[2] <= 0) return false;
  const S* K = c.K;
  const S hx = (K[0] * pc[0] + K[1] * pc[1]) + K[2] * pc[2];
  const S hy = (K[3] * pc[0] + K[4] * pc[1]) + K[5] * pc[2];
  const S hz = (K[6] * pc[0] + K[7] * pc[1]) + K[8] * pc[2];
  const S iz = S(1) / hz;
  const S u = hx * iz, v = hy * iz;
  if (u < 0 || u > c.image_cols || v < 0 || v > c.image_rows) return false;
  // d p_c / d dx = [ R | -2 R [p_m]x ]
  S Jx[18];
  const S* R = X.R;
  for (int i = 0; i < 3; ++i) {
    Jx[6 * i + 0] = R[3 * i + 0];
    Jx[6 * i + 1] = R[3 * i + 1];
    Jx[6 * i + 2] = R[3 * i + 2];
    // R [p]x : column j of [p]x
    // [p]x = [[0,-pz,py],[pz,0,-px],[-py,px,0]]
    Jx[6 * i + 3] = -2 * (R[3 * i + 1] * pm[2] - R[3 * i + 2] * pm[1]);
    Jx[6 * i + 4] = -2 * (-R[3 * i + 0] * pm[2] + R[3 * i + 2] * pm[0]);
    Jx[6 * i + 5] = -2 * (R[3 * i + 0] * pm[1] - R[3 * i + 1] * pm[0]);
  }
  // KJ = K * Jx  (3x6)
  S KJ[18];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 6; ++j)
      KJ[6 * i + j] = (K[3 * i] * Jx[j] + K[3 * i + 1] * Jx[6 + j]) + K[3 * i + 2] * Jx[12 + j];
  const S iz2 = iz * iz;
  for (int j = 0; j < 6; ++j) {
    J[j] = KJ[j] * iz - hx * iz2 * KJ[12 + j];
    J[6 + j] = KJ[6 + j] * iz - hy * iz2 * KJ[12 + j];
  }
  e[0] = u - z[0];
  e[1] = v - z[1];
  if (c.kind == FACTOR_STEREO) {
    const S hxr = hx + c.baseline[0];
    e[2] = hxr * iz - z[2];
    for (int j = 0; j < 6; ++j) J[12 + j] = KJ[j] * iz - hxr * iz2 * KJ[12 + j];
    if (c.mean_disparity > 0) {
      // inverse-depth weighting policy "min(1, 0.01 + d/mean_d)" on the translation block
      // (aligner_slice_processor_projective.cpp:110 comment; exact upstream clamp unknown)
      S w = S(0.01) + (z[0] - z[2]) / c.mean_disparity;
      if (w > 1) w = 1;
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) J[6 * i + j] *= w;
    }
  } else if (c.kind == FACTOR_DEPTH) {
    e[2] = pc[2] - z[2];  // landmark_estimator_pose_based_smoother_impl.cpp:65-97 analogue
    for (int j = 0; j < 6; ++j) J[12 + j] = Jx[12 + j];
  } else {
    e[2] = 0;
    for (int j = 0; j < 6; ++j) J[12 + j] = 0;
  }
  return true;
}

// FactorCorrespondenceDriven_::compute analogue: one pass over the correspondences.
// info_diag: 3 entries per FIXED index (aligner_slice_processor_projective.cpp:46-57).
template <typename S>
static inline void linearize(const LinearizeConfig<S>& c, const Pose<S>& X, const S* moving_xyz,
                             const S* fixed_meas, int fixed_dim, const int* corr_fixed,
                             const int* corr_moving, int n_corr, const S* info_diag,
                             LinearSystem<S>& sys, uint8_t* status = nullptr) {
  sys.clear();
  const int edim = (c.kind == FACTOR_MONO) ? 2 : 3;
  for (int k = 0; k < n_corr; ++k) {
    const int fi = corr_fixed[k], mi = corr_moving[k];
    S e[3], J[18];
    if (!error_and_jacobian(c, X, moving_xyz + 3 * mi, fixed_meas + fixed_dim * fi, e, J)) {
      ++sys.suppressed;
      if (status) status[k] = FACTOR_SUPPRESSED;
      continue;
    }
    S om[3] = {info_diag[3 * fi], info_diag[3 * fi + 1], info_diag[3 * fi + 2]};
    S chi = 0;
    for (int i = 0; i < edim; ++i) chi += e[i] * om[i] * e[i];
    S scale = 1;
    if (c.robustifier != ROBUST_NONE && chi > c.chi_threshold) {
      ++sys.outliers;
      scale = (c.robustifier == ROBUST_SATURATED) ? c.chi_threshold / chi : S(0);
      if (status) status[k] = FACTOR_KERNELIZED;
    } else {
      ++sys.inliers;
      if (status) status[k] = FACTOR_INLIER;
    }
    sys.chi_total += chi * scale;
    for (int i = 0; i < edim; ++i) {
      const S w = om[i] * scale;
      for (int a = 0; a < 6; ++a) {
        const S Jw = J[6 * i + a] * w;
        sys.b[a] += Jw * e[i];
        for (int bcol = 0; bcol < 6; ++bcol) sys.H[6 * a + bcol] += Jw * J[6 * i + bcol];
      }
    }
  }
}

// SE3 pose-prior factor = the second slice of the shipped aligners, AlignerSliceMotionModel3D
// (configurations/kitti.conf:747-772, icl.conf:268-293, euroc.conf:94-119: fixed and moving slice
// "trajectory_chunk", a MotionModelConstantVelocity3D, no robustifier).  [upstream, srrg2_slam_interfaces +
// srrg2_solver, not verifiable here]  The motion model predicts the estimate Z; the factor's error is the
// 6-vector chart of the deviation, e = t2tnq(Z^-1 X), with a constant information matrix, summed into the
// same 6x6 system before the solve.  For the right perturbation X <- X v2t(dx),
// E = Z^-1 X = (R_E, t_E), q_E = (w, v) with w >= 0:
//   d e_t / d dt = R_E      d e_t / d dq = 0      d e_q / d dt = 0      d e_q / d dq = w I + [v]x
// (first order in dq: q_E (x) (1, dq) has vector part w dq + v x dq + v).  Returns chi = e' Omega e.
template <typename S>
static inline S pose_prior_accumulate(const Pose<S>& Z, const S* Omega36, const Pose<S>& X,
                                      LinearSystem<S>& sys) {
  const Pose<S> E = Z.inverse() * X;
  S e[6];
  t2tnq(E, e);
  const S n2 = e[3] * e[3] + e[4] * e[4] + e[5] * e[5];
  const S w = std::sqrt(n2 < S(1) ? S(1) - n2 : S(0));
  S J[36];
  for (S& v : J) v = 0;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) J[6 * i + j] = E.R[3 * i + j];
  const S vx = e[3], vy = e[4], vz = e[5];
  const S Q[9] = {w, -vz, vy, vz, w, -vx, -vy, vx, w};
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) J[6 * (3 + i) + 3 + j] = Q[3 * i + j];
  S OJ[36], Oe[6];
  for (int i = 0; i < 6; ++i) {
    S s = 0;
    for (int k = 0; k < 6; ++k) s += Omega36[6 * i + k] * e[k];
    Oe[i] = s;
    for (int j = 0; j < 6; ++j) {
      S a = 0;
      for (int k = 0; k < 6; ++k) a += Omega36[6 * i + k] * J[6 * k + j];
      OJ[6 * i + j] = a;
    }
  }
  S chi = 0;
  for (int i = 0; i < 6; ++i) chi += e[i] * Oe[i];
  for (int a = 0; a < 6; ++a) {
    S bs = 0;
    for (int k = 0; k < 6; ++k) bs += J[6 * k + a] * Oe[k];
    sys.b[a] += bs;
    for (int c = 0; c < 6; ++c) {
      S hs = 0;
      for (int k = 0; k < 6; ++k) hs += J[6 * k + a] * OJ[6 * k + c];
      sys.H[6 * a + c] += hs;
    }
  }
  return chi;
}

// (H + lambda I) dx = -b by Cholesky; X <- X * v2t(dx).  Returns false if not SPD.
template <typename S>
static inline bool gn_step(const LinearSystem<S>& sys, S damping, Pose<S>& X, S* dx_out) {
  S A[36];
  for (int i = 0; i < 36; ++i) A[i] = sys.H[i];
  for (int i = 0; i < 6; ++i) A[7 * i] += damping;
  S L[36] = {0};
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j <= i; ++j) {
      S s = A[6 * i + j];
      for (int k = 0; k < j; ++k) s -= L[6 * i + k] * L[6 * j + k];
      if (i == j) {
        if (!(s > 0)) return false;
        L[6 * i + i] = std::sqrt(s);
      } else {
        L[6 * i + j] = s / L[6 * j + j];
      }
    }
  S y[6], dx[6];
  for (int i = 0; i < 6; ++i) {
    S s = -sys.b[i];
    for (int k = 0; k < i; ++k) s -= L[6 * i + k] * y[k];
    y[i] = s / L[6 * i + i];
  }
  for (int i = 5; i >= 0; --i) {
    S s = y[i];
    for (int k = i + 1; k < 6; ++k) s -= L[6 * k + i] * dx[k];
    dx[i] = s / L[6 * i + i];
  }
  X = X * v2t(dx);
  if (dx_out)
    for (int i = 0; i < 6; ++i) dx_out[i] = dx[i];
  return true;
}

}  // namespace pslam_oracle
