// pslam_plugin.cpp -- see pslam_plugin.hpp.  All arithmetic of the hot path runs in libpslam_cuda.so.
#include "pslam_plugin.hpp"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <mutex>

namespace pslam_host {

// ---- small isometry helpers (control flow only: convergence test of the projective finder) --------------------
Isometry3f Isometry3f::inverse() const {
  Isometry3f r;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) r.m[4 * i + j] = m[4 * j + i];
  for (int i = 0; i < 3; ++i) r.m[4 * i + 3] = -(r.m[4 * i] * m[3] + r.m[4 * i + 1] * m[7] + r.m[4 * i + 2] * m[11]);
  return r;
}
Isometry3f Isometry3f::operator*(const Isometry3f& o) const {
  Isometry3f r;
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j)
      r.m[4 * i + j] = m[4 * i] * o.m[j] + m[4 * i + 1] * o.m[4 + j] + m[4 * i + 2] * o.m[8 + j];
    r.m[4 * i + 3] = m[4 * i] * o.m[3] + m[4 * i + 1] * o.m[7] + m[4 * i + 2] * o.m[11] + m[4 * i + 3];
  }
  return r;
}
void t2tnq(const Isometry3f& T, float v6[6]) {
  v6[0] = T.m[3];
  v6[1] = T.m[7];
  v6[2] = T.m[11];
  // rotation matrix -> unit quaternion (Eigen's branch structure), then w >= 0 and the vector part
  const float r00 = T.m[0], r01 = T.m[1], r02 = T.m[2], r10 = T.m[4], r11 = T.m[5], r12 = T.m[6], r20 = T.m[8],
              r21 = T.m[9], r22 = T.m[10];
  float w, x, y, z;
  float t = r00 + r11 + r22;
  if (t > 0.f) {
    t = std::sqrt(t + 1.f);
    w = 0.5f * t;
    t = 0.5f / t;
    x = (r21 - r12) * t;
    y = (r02 - r20) * t;
    z = (r10 - r01) * t;
  } else {
    int i = 0;
    if (r11 > r00) i = 1;
    if (r22 > (i == 0 ? r00 : r11)) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    auto R = [&](int a, int b) { return T.m[4 * a + b]; };
    t = std::sqrt(R(i, i) - R(j, j) - R(k, k) + 1.f);
    float q[3];
    q[i] = 0.5f * t;
    t = 0.5f / t;
    w = (R(k, j) - R(j, k)) * t;
    q[j] = (R(j, i) + R(i, j)) * t;
    q[k] = (R(k, i) + R(i, k)) * t;
    x = q[0];
    y = q[1];
    z = q[2];
  }
  const float n = std::sqrt(w * w + x * x + y * y + z * z);
  const float s = (w < 0.f ? -1.f : 1.f) / n;
  v6[3] = x * s;
  v6[4] = y * s;
  v6[5] = z * s;
}

// ---- device context ---------------------------------------------------------------------------------------------
namespace {
std::mutex g_mutex;
pslam_ctx* g_ctx = nullptr;
int g_device = 0, g_rows = 0, g_cols = 0;
constexpr int kMaxFeatures = 8192;
}  // namespace

void PslamDevice::setDevice(int device) {
  std::lock_guard<std::mutex> lock(g_mutex);
  g_device = device;
}

void PslamDevice::release() {
  std::lock_guard<std::mutex> lock(g_mutex);
  if (g_ctx) pslam_destroy(g_ctx);
  g_ctx = nullptr;
  g_rows = g_cols = 0;
}

pslam_ctx* PslamDevice::context(int rows, int cols) {
  std::lock_guard<std::mutex> lock(g_mutex);
  if (g_ctx && rows <= g_rows && cols <= g_cols) return g_ctx;
  const int nr = std::max({rows, g_rows, 480}), nc = std::max({cols, g_cols, 1280});
  if (g_ctx) pslam_destroy(g_ctx);
  g_ctx = nullptr;
  pslam_limits lim{};
  lim.max_images = 2;
  lim.max_rows = nr;
  lim.max_cols = nc;
  lim.max_features = kMaxFeatures;
  lim.max_raw_per_bin = 32768;
  lim.max_bins = 64;
  lim.work_images = 2;
  pslam_ctx* ctx = nullptr;
  const int rc = pslam_create(g_device, &lim, &ctx);
  if (rc != PSLAM_OK) {
    std::string msg = std::string("PslamDevice::context|ERROR: no usable CUDA device (there is no CPU fallback): ") +
                      (ctx ? pslam_last_error(ctx) : "pslam_create failed");
    if (ctx) pslam_destroy(ctx);
    throw std::runtime_error(msg);
  }
  g_ctx = ctx;
  g_rows = nr;
  g_cols = nc;
  return g_ctx;
}

void PslamDevice::check(int rc, const char* where) {
  if (rc >= 0) return;
  const char* e = g_ctx ? pslam_last_error(g_ctx) : "";
  // PSLAM_E_INVALID carries the reference's own std::runtime_error text
  if (rc == PSLAM_E_INVALID && e && *e) throw std::runtime_error(e);
  throw std::runtime_error(std::string(where) + "|ERROR: " + (e && *e ? e : "CUDA path failed") + " (code " + std::to_string(rc) + ")");
}

float Solver::damping() const {
  auto gn = std::dynamic_pointer_cast<IterationAlgorithmGN>(param_algorithm.value());
  return gn ? gn->param_damping.value() : 0.f;
}

// ---- feature extractors -----------------------------------------------------------------------------------------
void IntensityFeatureExtractorBaseCUDA::checkDetectorAndDescriptor() const {
  // base.cpp:105-176: the CUDA path implements the detector / descriptor pair every shipped configuration uses
  const std::string& det = param_detector_type.value();
  if (det != "FAST") {
    if (det == "BRISK-512" || det == "ORB-256" || det == "MSER" || det == "GFTT")
      throw std::runtime_error("IntensityFeatureExtractor::_setDetector|ERROR: detector type not available in the CUDA path: " + det);
    throw std::runtime_error("IntensityFeatureExtractor::_setDetector|ERROR: unknown detector type chosen: " + det);
  }
  const std::string& desc = param_descriptor_type.value();
  if (desc != "ORB-256") {
    if (desc == "BRIEF-256" || desc == "BRISK-512" || desc == "FREAK-512")
      throw std::runtime_error("IntensityFeatureExtractor::_setDescriptor|ERROR: descriptor type not available in the CUDA path: " + desc);
    throw std::runtime_error("IntensityFeatureExtractor::_setDescriptor|ERROR: unknown descriptor type chosen: " + desc);
  }
}

pslam_extract_cfg IntensityFeatureExtractorBaseCUDA::cudaConfig() const {
  pslam_extract_cfg c{};
  c.detector_threshold = param_detector_threshold.value();
  c.enable_non_maximum_suppression = param_enable_non_maximum_suppression.value() ? 1 : 0;
  c.target_number_of_keypoints = param_target_number_of_keypoints.value();
  c.number_of_detectors_horizontal = 1;
  c.number_of_detectors_vertical = 1;
  return c;
}

void IntensityFeatureExtractorBaseCUDA::prepare(int rows, int cols) {
  if ((size_t) rows != _image_rows || (size_t) cols != _image_cols) {  // base.cpp:60-66: re-init on new dimensions
    _image_rows = rows;
    _image_cols = cols;
    _config_changed = true;
  }
  init();
}

void IntensityFeatureExtractorBaseCUDA::compute(const ImageView& image) {
  if (!_features) throw std::runtime_error("IntensityFeatureExtractor::compute|ERROR: features not set");
  prepare(image.rows, image.cols);
  pslam_ctx* ctx = PslamDevice::context(image.rows, image.cols);
  std::vector<float> xy(2 * (size_t) kMaxFeatures), response(kMaxFeatures);
  PointIntensityDescriptorCloud& out = *_features;
  out.dim = _point_dim;
  out.number_of_optimizations.clear();
  out.resize(kMaxFeatures);
  const int n = extract(ctx, image, kMaxFeatures, xy.data(), response.data(), out.intensity.data(), out.descriptor.data());
  _mask_set = false;  // the detection mask is consumed by one call (selective.cpp:203, binned.cpp:206)
  if (n < 0) out.resize(0);
  PslamDevice::check(n, "IntensityFeatureExtractor::compute");
  if (n > kMaxFeatures) {
    out.resize(0);
    throw std::runtime_error("IntensityFeatureExtractor::compute|ERROR: more features than the device context holds");
  }
  out.resize(n);
  for (int i = 0; i < n; ++i) {  // AoS fill of base.cpp:73-84: coordinates(0) = x, (1) = y, remaining dimensions 0
    float* p = out.point(i);
    p[0] = xy[2 * i];
    p[1] = xy[2 * i + 1];
    for (int d = 2; d < _point_dim; ++d) p[d] = 0.f;
  }
}

void IntensityFeatureExtractorBinnedCUDA::init() {
  if (!_config_changed) return;
  // binned.cpp:13-28
  if (param_number_of_detectors_vertical.value() <= 0)
    throw std::runtime_error("IntensityFeatureExtractor::init|ERROR: invalid number of vertical detectors (check configuration!)");
  if (param_number_of_detectors_horizontal.value() <= 0)
    throw std::runtime_error("IntensityFeatureExtractor::init|ERROR: invalid number of horizontal detectors (check configuration!)");
  if (_image_rows == 0)
    throw std::runtime_error("IntensityFeatureExtractor::init|ERROR: invalid number of image rows (check configuration!)");
  if (_image_cols == 0)
    throw std::runtime_error("IntensityFeatureExtractor::init|ERROR: invalid number of image cols (check configuration!)");
  checkDetectorAndDescriptor();
  std::cerr << "IntensityFeatureExtractorBinned_::init|initialized for image sources [" << _image_rows << " x " << _image_cols
            << "] with detector grid [" << param_number_of_detectors_vertical.value() << " x "
            << param_number_of_detectors_horizontal.value() << "]" << std::endl;
  _config_changed = false;
}

pslam_extract_cfg IntensityFeatureExtractorBinnedCUDA::cudaConfig() const {
  pslam_extract_cfg c = IntensityFeatureExtractorBaseCUDA::cudaConfig();
  c.number_of_detectors_horizontal = param_number_of_detectors_horizontal.value();
  c.number_of_detectors_vertical = param_number_of_detectors_vertical.value();
  return c;
}

int IntensityFeatureExtractorBinnedCUDA::extract(pslam_ctx* ctx, const ImageView& image, int capacity, float* xy, float* response,
                                                 float* intensity, uint8_t* desc) {
  const pslam_extract_cfg cfg = cudaConfig();
  return pslam_extract_binned(ctx, image.data, image.rows, image.cols, image.stride, &cfg, _mask_set ? _mask.data : nullptr,
                              capacity, xy, response, intensity, desc);
}

void IntensityFeatureExtractorSelectiveCUDA::init() {
  if (!_config_changed) return;
  if (_image_rows == 0)
    throw std::runtime_error("IntensityFeatureExtractorSelective_::init|ERROR: invalid number of image rows (check configuration!)");
  if (_image_cols == 0)
    throw std::runtime_error("IntensityFeatureExtractorSelective_::init|ERROR: invalid number of image cols (check configuration!)");
  checkDetectorAndDescriptor();
  std::cerr << "IntensityFeatureExtractor::init|initialized for image sources [" << _image_rows << " x " << _image_cols << "]" << std::endl;
  _config_changed = false;
}

// detection_mask_tracking of selective.cpp:63-144: one rectangle per projection, half-size = radius + base_radius (10),
// clipped at the top / left border, its extent cut at the bottom / right border; four variants (full bars to the left /
// right image border).  Host-side integer work; pinned through the reference's own constants (94 seeds -> 719 / 581 /
// 294 / 237 candidates, tests/test_feature_extractors.cpp:182-213) in tests/test_plugin_cpu.py.
void IntensityFeatureExtractorSelectiveCUDA::paintTrackingMask(int rows, int cols, const PointIntensityDescriptorCloud& projections,
                                                               int projection_detection_radius, std::vector<uint8_t>& mask) const {
  constexpr int16_t base_radius = 10;
  const int16_t radius_pixels = (int16_t) (projection_detection_radius + base_radius);
  const int16_t radius_x2 = 2 * radius_pixels;
  mask.assign((size_t) rows * cols, 0);
  const bool to_left = param_enable_full_distance_to_left.value(), to_right = param_enable_full_distance_to_right.value();
  for (size_t k = 0; k < projections.size(); ++k) {
    const float* c = projections.point(k);
    const int16_t row = (int16_t) std::round(c[1]), col = (int16_t) std::round(c[0]);
    const int16_t tl_row = std::max<int>(row - radius_pixels, 0);
    const int16_t height = std::min<int16_t>(radius_x2, (int16_t) (rows - tl_row));
    int x0, w;
    if (to_left && to_right) {
      x0 = 0;
      w = cols;
    } else if (to_left) {
      x0 = 0;
      w = col;
    } else if (to_right) {
      x0 = col;
      w = (int16_t) (cols - col);
    } else {
      const int16_t tl_col = std::max<int>(col - radius_pixels, 0);
      x0 = tl_col;
      w = std::min<int16_t>(radius_x2, (int16_t) (cols - tl_col));
    }
    // cv::Mat(rect).setTo(1); a rectangle reaching outside the image asserts in the reference
    for (int r = tl_row; r < tl_row + height && r < rows; ++r)
      for (int x = std::max(x0, 0); x < x0 + w && x < cols; ++x) mask[(size_t) r * cols + x] = 1;
  }
}

int IntensityFeatureExtractorSelectiveCUDA::extract(pslam_ctx* ctx, const ImageView& image, int capacity, float* xy,
                                                    float* response, float* intensity, uint8_t* desc) {
  const pslam_extract_cfg cfg = cudaConfig();
  _n_tracking = 0;
  if (_projections && !_projections->empty()) {
    // tracking phase (selective.cpp:63-178): rectangles of half-size radius + 10 around the projections
    const int rows = image.rows, cols = image.cols;
    paintTrackingMask(rows, cols, *_projections, _projection_detection_radius, _tracking_mask);
    int n_tracking = 0;
    const int n = pslam_extract_selective(ctx, image.data, rows, cols, image.stride, &cfg, _tracking_mask.data(),
                                          param_enable_seeding_when_tracking.value() ? 1 : 0, capacity, xy, response, intensity,
                                          desc, &n_tracking);
    if (n >= 0) {
      _n_tracking = n_tracking;
      if (n_tracking == 0)
        std::cerr << "IntensityFeatureExtractorSelective_::computeKeypoints|WARNING: no keypoints detected for projected regions: "
                  << _projections->size() << std::endl;
    }
    _projections = nullptr;  // :177
    return n;
  }
  // seeding phase (selective.cpp:181-198): every FAST keypoint (optionally inside the external mask), no selection
  std::vector<uint8_t> all;
  const uint8_t* mask = _mask_set ? _mask.data : nullptr;
  if (!mask) {
    all.assign((size_t) image.rows * image.cols, 1);
    mask = all.data();
  }
  const int n = pslam_extract_binned(ctx, image.data, image.rows, image.cols, image.stride, &cfg, mask, capacity, xy, response,
                                     intensity, desc);
  if (n == 0) std::cerr << "IntensityFeatureExtractorSelective_::computeKeypoints|WARNING: unable to seed keypoints" << std::endl;
  return n;
}

// ---- descriptor based finders -----------------------------------------------------------------------------------
void CorrespondenceFinderDescriptorBasedBruteforceCUDA::_preCompute() {
  if (!_fixed) throw std::runtime_error("CorrespondenceFinderDescriptorBased::compute|ERROR: fixed not set");
  if (!_moving) throw std::runtime_error("CorrespondenceFinderDescriptorBased::compute|ERROR: moving not set");
  if (!_correspondences) throw std::runtime_error("CorrespondenceFinderDescriptorBased::compute|ERROR: correspondences not set");
  if (_fixed->empty()) std::cerr << "CorrespondenceFinderDescriptorBased::compute|WARNING: no points in fixed" << std::endl;
  if (_moving->empty()) std::cerr << "CorrespondenceFinderDescriptorBased::compute|WARNING: no points in moving" << std::endl;
}

void CorrespondenceFinderDescriptorBasedBruteforceCUDA::_postCompute() {
  _fixed_changed_flag = false;
  _moving_changed_flag = false;
  if (_correspondences->empty())
    std::cerr << "CorrespondenceFinderDescriptorBased::compute|WARNING: no correspondences found" << std::endl;
}

namespace {
void fill(CorrespondenceVector& out, int n, const std::vector<int>& f, const std::vector<int>& m, const std::vector<float>& d) {
  out.resize(n);
  for (int i = 0; i < n; ++i) out[i] = Correspondence{f[i], m[i], d[i]};
}
// 2-D image coordinates of a cloud (epipolar finder: truncated (row, col) = (y, x), epipolar_impl.cpp:8-20)
std::vector<float> xy_of(const PointIntensityDescriptorCloud& c) {
  std::vector<float> xy(2 * c.size());
  for (size_t i = 0; i < c.size(); ++i) {
    xy[2 * i] = c.point(i)[0];
    xy[2 * i + 1] = c.point(i)[1];
  }
  return xy;
}
}  // namespace

void CorrespondenceFinderDescriptorBasedBruteforceCUDA::compute() {
  _preCompute();
  if (!_fixed_changed_flag && !_moving_changed_flag) return;  // bruteforce_impl.cpp:13-15
  pslam_ctx* ctx = PslamDevice::context();
  pslam_match_cfg cfg{};
  cfg.maximum_descriptor_distance = param_maximum_descriptor_distance.value();
  cfg.maximum_distance_ratio_to_second_best = param_maximum_distance_ratio_to_second_best.value();
  const int cap = (int) std::min(_fixed->size(), _moving->size()) + 1;
  std::vector<int> f(cap), m(cap);
  std::vector<float> d(cap);
  const int n = pslam_match_bruteforce(ctx, (int) _fixed->size(), _fixed->descriptor.data(), (int) _moving->size(),
                                       _moving->descriptor.data(), &cfg, cap, f.data(), m.data(), d.data());
  PslamDevice::check(n, "CorrespondenceFinderDescriptorBasedBruteforce::compute");
  fill(*_correspondences, n, f, m, d);
  _postCompute();
}

pslam_match_cfg CorrespondenceFinderDescriptorBasedEpipolarCUDA::cudaConfig() const {
  pslam_match_cfg cfg{};
  cfg.maximum_descriptor_distance = param_maximum_descriptor_distance.value();
  cfg.maximum_distance_ratio_to_second_best = param_maximum_distance_ratio_to_second_best.value();
  cfg.maximum_disparity_pixels = (int) param_maximum_disparity_pixels.value();
  cfg.epipolar_line_thickness_pixels = (int) param_epipolar_line_thickness_pixels.value();
  return cfg;
}

void CorrespondenceFinderDescriptorBasedEpipolarCUDA::compute() {
  _preCompute();
  if (!_fixed_changed_flag && !_moving_changed_flag) return;  // epipolar_impl.cpp:49-51
  pslam_ctx* ctx = PslamDevice::context();
  const pslam_match_cfg cfg = cudaConfig();
  const std::vector<float> xf = xy_of(*_fixed), xm = xy_of(*_moving);
  const int cap = (int) _fixed->size() + 1;
  std::vector<int> f(cap), m(cap);
  std::vector<float> d(cap);
  const int n = pslam_match_epipolar(ctx, (int) _fixed->size(), xf.data(), _fixed->descriptor.data(), (int) _moving->size(),
                                     xm.data(), _moving->descriptor.data(), &cfg, cap, f.data(), m.data(), d.data());
  PslamDevice::check(n, "CorrespondenceFinderDescriptorBasedEpipolar::compute");
  fill(*_correspondences, n, f, m, d);
  _postCompute();
}

// ---- projective finder: host state machine (projective_base_impl.cpp:104-293), device search + filter ---------
void CorrespondenceFinderProjectiveCUDA::_prepare() {
  _preCompute();
  pslam_ctx* ctx = PslamDevice::context();
  // the device cache is shared by every finder instance of the process (and dropped when the context is re-created):
  // upload again when it no longer carries the stamps of OUR uploads
  unsigned long long fixed_epoch = 0, moving_epoch = 0;
  pslam_projective_cache_epochs(ctx, &fixed_epoch, &moving_epoch);
  const bool fixed_stale = fixed_epoch != _device_fixed_epoch, moving_stale = moving_epoch != _device_moving_epoch;
  if (!(_fixed_changed_flag || _moving_changed_flag || _config_changed) && (fixed_stale || moving_stale)) {
    if (fixed_stale)
      PslamDevice::check(pslam_projective_set_fixed(ctx, (int) _fixed->size(), _fixed->coordinates.data(), _fixed->dim,
                                                    _fixed->descriptor.data()),
                         "CorrespondenceFinderProjective::compute");
    if (moving_stale)
      PslamDevice::check(pslam_projective_set_moving(ctx, (int) _moving->size(), _moving->coordinates.data(),
                                                     _moving->descriptor.data()),
                         "CorrespondenceFinderProjective::compute");
    pslam_projective_cache_epochs(ctx, &_device_fixed_epoch, &_device_moving_epoch);
  }
  if (_fixed_changed_flag || _moving_changed_flag || _config_changed) {
    const bool fixed_changed = _fixed_changed_flag || fixed_stale, moving_changed = _moving_changed_flag || moving_stale;
    _fixed_changed_flag = false;
    _moving_changed_flag = false;
    if ((_search_radius_pixels == 0 && _descriptor_distance == 0) || _config_changed) {
      _search_radius_pixels = param_maximum_search_radius_pixels.value();
      _descriptor_distance = param_minimum_descriptor_distance.value();
      std::cerr << "CorrespondenceFinderProjective::compute|initialized search radius (px): " << _search_radius_pixels
                << " descriptor distance: " << _descriptor_distance
                << " maximum distance ratio: " << param_maximum_distance_ratio_to_second_best.value() << std::endl;
    }
    _has_converged = false;
    _current_iteration = 0;
    _local_map_in_sensor_previous.setIdentity();
    // _initializeDatabase (square_impl.cpp:7-31): the row-sorted lattice is built and cached on the device
    if (fixed_changed || _config_changed)
      PslamDevice::check(pslam_projective_set_fixed(ctx, (int) _fixed->size(), _fixed->coordinates.data(), _fixed->dim,
                                                    _fixed->descriptor.data()),
                         "CorrespondenceFinderProjective::compute");
    if (moving_changed || _config_changed) {
      if (_moving->dim != 3) throw std::runtime_error("CorrespondenceFinderProjective::compute|ERROR: moving cloud must be 3D");
      PslamDevice::check(pslam_projective_set_moving(ctx, (int) _moving->size(), _moving->coordinates.data(),
                                                     _moving->descriptor.data()),
                         "CorrespondenceFinderProjective::compute");
    }
    pslam_projective_cache_epochs(ctx, &_device_fixed_epoch, &_device_moving_epoch);
    _config_changed = false;
  }
}

pslam_projective_cfg CorrespondenceFinderProjectiveCUDA::_deviceConfig() const {
  const ProjectorPinhole& projector = *param_projector.value();
  pslam_projective_cfg cfg{};
  for (int i = 0; i < 9; ++i) cfg.K[i] = projector.cameraMatrix()[i];
  cfg.canvas_rows = (int) projector.param_canvas_rows.value();
  cfg.canvas_cols = (int) projector.param_canvas_cols.value();
  cfg.range_min = projector.param_range_min.value();
  cfg.range_max = projector.param_range_max.value();
  cfg.shape = _shape;
  cfg.search_radius_pixels = (int) _search_radius_pixels;
  cfg.descriptor_distance = _descriptor_distance;
  cfg.maximum_distance_ratio_to_second_best = param_maximum_distance_ratio_to_second_best.value();
  cfg.maximum_descriptor_distance = param_maximum_descriptor_distance.value();
  return cfg;
}

void CorrespondenceFinderProjectiveCUDA::_uploadWeights(pslam_ctx* ctx, const float* moving_scale) {
  if (!moving_scale) return;
  unsigned long long fe = 0, me = 0;
  pslam_projective_cache_epochs(ctx, &fe, &me);
  if (_device_weights_of != moving_scale || _device_weights_epoch != me) {
    PslamDevice::check(pslam_projective_set_moving_weights(ctx, (int) _moving->size(), moving_scale),
                       "CorrespondenceFinderProjective::compute");
    _device_weights_of = moving_scale;
    _device_weights_epoch = me;
  }
}

void CorrespondenceFinderProjectiveCUDA::_adaptAfterConvergence() {
  _search_radius_pixels = std::max<size_t>(_search_radius_pixels - param_search_radius_step_size_pixels.value(),
                                           param_minimum_search_radius_pixels.value());
  _descriptor_distance = std::min(_descriptor_distance + param_descriptor_distance_step_size_pixels.value(),
                                  param_maximum_descriptor_distance.value());
}

bool CorrespondenceFinderProjectiveCUDA::alignFused(FusedSolveRequest& r, int min_num_correspondences) {
  if (!_fixed || !_moving || !_correspondences || _fixed->empty() || _moving->empty()) return false;
  if (!r.factor || r.max_fused <= 0 || !param_projector.value() || _moving->dim != 3) return false;
  const size_t N = param_number_of_solver_iterations_per_projection.value();
  if (N == 0 || r.max_fused > (1 << 20)) return false;
  _prepare();
  if (_has_converged || !(_current_iteration % N == 0 || _current_iteration == 1)) return false;
  pslam_ctx* ctx = PslamDevice::context();
  _uploadWeights(ctx, r.moving_scale);
  const pslam_projective_cfg cfg = _deviceConfig();
  pslam_align al{};
  al.max_iterations = r.max_fused;
  al.solver_iterations_per_projection = (int) N;
  al.minimum_number_of_iterations = (int) param_minimum_number_of_iterations.value();
  al.maximum_estimate_change_norm_for_convergence = param_maximum_estimate_change_norm_for_convergence.value();
  al.minimum_matching_ratio = param_minimum_matching_ratio.value();
  al.can_widen_search = (_search_radius_pixels < param_maximum_search_radius_pixels.value() ||
                         _descriptor_distance > param_minimum_descriptor_distance.value()) ? 1 : 0;
  al.min_num_correspondences = min_num_correspondences;
  al.current_iteration = (int) _current_iteration;
  al.has_converged = 0;
  for (int i = 0; i < 12; ++i) al.previous12[i] = _local_map_in_sensor_previous.m[i];
  pslam_fused_gn g{};
  g.factor = r.factor;
  for (int i = 0; i < 3; ++i) g.diagonal_info[i] = r.diagonal_info[i];
  g.n_iterations = r.max_fused;
  g.damping = r.damping;
  for (int i = 0; i < 12; ++i) g.pose12[i] = r.estimate[i];
  g.prior = r.prior;
  const int cap = (int) _fixed->size() + 1;
  r.poses.resize(12 * (size_t) r.max_fused);
  r.stats.resize(4 * (size_t) r.max_fused);
  r.status.assign((size_t) cap, 0);
  g.poses12 = r.poses.data();
  g.stats4 = r.stats.data();
  g.factor_status = r.status.data();
  std::vector<int> f(cap), m(cap);
  std::vector<float> d(cap);
  const int n = pslam_projective_align(ctx, (int) _fixed->size(), (int) _moving->size(), &cfg, &al, cap, f.data(), m.data(), d.data(), &g);
  PslamDevice::check(n, "CorrespondenceFinderProjective::compute");
  r.executed = true;
  r.done = al.iterations_done;
  r.spd = g.spd != 0;
  r.stop_reason = al.stop_reason;
  r.phase_log.assign(al.phase_log, al.phase_log + 3 * std::min(al.n_phases, (int) PSLAM_ALIGN_MAX_PHASES));
  r.status.resize((size_t) std::max(n, 0));
  if (al.n_phases == 0) return true;  // the first search already needs the caller's loop: nothing has changed
  // the state the call-by-call loop would have left behind
  _number_of_searches += al.n_phases;
  _current_iteration = (size_t) al.current_iteration;
  _has_converged = al.has_converged != 0;
  for (int i = 0; i < 12; ++i) _local_map_in_sensor_previous.m[i] = al.previous12[i];
  {  // _local_map_in_sensor = the pose handed to the last compute() of the loop
    const int last = al.n_phases - 1, start = al.phase_log[3 * last], done = al.phase_log[3 * last + 1];
    const int row = done >= 2 ? start + done - 2 : start - 1;  // pose after that solver iteration (-1: the initial estimate)
    for (int i = 0; i < 12; ++i) _local_map_in_sensor.m[i] = (float) (row >= 0 ? r.poses[12 * (size_t) row + i] : r.estimate[i]);
  }
  for (int p = 0; p < al.n_phases && p < (int) PSLAM_ALIGN_MAX_PHASES; ++p) {
    const float matching_ratio = static_cast<float>(al.phase_log[3 * p + 2]) / _fixed->size();
    if (matching_ratio < param_minimum_matching_ratio.value())
      std::cerr << "CorrespondenceFinderProjective::compute|low matching ratio: " << matching_ratio << " (" << al.phase_log[3 * p + 2]
                << "/" << _fixed->size() << ") target: " << param_minimum_matching_ratio.value() << std::endl;
  }
  if (al.n_projected == 0) std::cerr << "CorrespondenceFinderProjective::compute|WARNING: all projections failed" << std::endl;
  if (al.converged_with_good_ratio) _adaptAfterConvergence();
  if (al.stop_reason != 2 && al.stop_reason != 3) fill(*_correspondences, n, f, m, d);
  _postCompute();
  return true;
}

void CorrespondenceFinderProjectiveCUDA::compute() {
  _prepare();
  pslam_ctx* ctx = PslamDevice::context();
  if (_has_converged) {  // correspondences are not touched (:137-141)
    _postCompute();
    return;
  }
  if (!param_projector.value()) throw std::runtime_error("CorrespondenceFinderProjective::compute|ERROR: projector not set");

  // periodically re-project; always for iterations 0 and 1 (:162-178)
  if (!(_current_iteration % param_number_of_solver_iterations_per_projection.value() == 0 || _current_iteration == 1)) {
    _local_map_in_sensor_previous = _local_map_in_sensor;
    ++_current_iteration;
    _postCompute();
    return;
  }
  float v6[6];
  t2tnq(_local_map_in_sensor.inverse() * _local_map_in_sensor_previous, v6);  // cameraPose() * previous (:181-183)
  float n2 = 0;
  for (float v : v6) n2 += v * v;
  const float estimate_change_norm = std::sqrt(n2);
  _local_map_in_sensor_previous = _local_map_in_sensor;

  // projector->compute + _findNearestNeighbors + _filterCorrespondences: one device call
  const pslam_projective_cfg cfg = _deviceConfig();
  const int cap = (int) _fixed->size() + 1;
  std::vector<int> f(cap), m(cap);
  std::vector<float> d(cap);
  int n_projected = 0;
  int n;
  if (_fused && _fused->factor && _fused->max_fused > 0) {
    // the aligner runs this iteration and the following "quiet" ones (calls that keep these correspondences, :162-178) in the
    // same device round trip as the search.  State AFTER this call, as the lines below will set it:
    const bool converged_after = estimate_change_norm < param_maximum_estimate_change_norm_for_convergence.value() &&
                                 _current_iteration > param_minimum_number_of_iterations.value();
    int quiet = 0;
    if (converged_after) {
      quiet = 1 << 30;
    } else {
      const size_t N = param_number_of_solver_iterations_per_projection.value();
      for (size_t it = _current_iteration + 1; !(it % N == 0 || it == 1); ++it) ++quiet;
    }
    const int n_fused = quiet >= _fused->max_fused ? _fused->max_fused : quiet + 1;
    _uploadWeights(ctx, _fused->moving_scale);
    pslam_fused_gn g{};
    g.factor = _fused->factor;
    for (int i = 0; i < 3; ++i) g.diagonal_info[i] = _fused->diagonal_info[i];
    g.n_iterations = n_fused;
    g.damping = _fused->damping;
    for (int i = 0; i < 12; ++i) g.pose12[i] = _fused->estimate[i];
    g.prior = _fused->prior;
    _fused->poses.resize(12 * (size_t) n_fused);
    _fused->stats.resize(4 * (size_t) n_fused);
    _fused->status.assign((size_t) cap, 0);
    g.poses12 = _fused->poses.data();
    g.stats4 = _fused->stats.data();
    g.factor_status = _fused->status.data();
    n = pslam_projective_match_gn(ctx, (int) _fixed->size(), (int) _moving->size(), _local_map_in_sensor.m, &cfg, cap, f.data(),
                                  m.data(), d.data(), &n_projected, &g);
    PslamDevice::check(n, "CorrespondenceFinderProjective::compute");
    _fused->executed = true;
    _fused->done = g.iterations_done;
    _fused->spd = g.spd != 0;
    _fused->status.resize((size_t) std::max(n, 0));
  } else {
    n = pslam_projective_match(ctx, (int) _fixed->size(), (int) _moving->size(), _local_map_in_sensor.m, &cfg, cap, f.data(),
                               m.data(), d.data(), &n_projected);
    PslamDevice::check(n, "CorrespondenceFinderProjective::compute");
  }
  ++_number_of_searches;
  if (n_projected == 0) std::cerr << "CorrespondenceFinderProjective::compute|WARNING: all projections failed" << std::endl;

  const float matching_ratio = static_cast<float>(n) / _fixed->size();
  if (matching_ratio < param_minimum_matching_ratio.value()) {
    std::cerr << "CorrespondenceFinderProjective::compute|low matching ratio: " << matching_ratio << " (" << n << "/"
              << _fixed->size() << ") target: " << param_minimum_matching_ratio.value() << std::endl;
    if (_search_radius_pixels < param_maximum_search_radius_pixels.value() ||
        _descriptor_distance > param_minimum_descriptor_distance.value()) {
      _search_radius_pixels = param_maximum_search_radius_pixels.value();
      _descriptor_distance = param_minimum_descriptor_distance.value();
      std::cerr << "CorrespondenceFinderProjective|WARNING: bad initial guess - triggering internal repeat with increased search radius"
                << std::endl;
      if (matching_ratio == 0) {
        std::cerr << "CorrespondenceFinderProjective|WARNING: complete track loss - fallback to identity motion guess" << std::endl;
        _local_map_in_sensor.setIdentity();
        _current_iteration = 0;
      } else {
        ++_current_iteration;
      }
      if (_fused) _fused->executed = false;  // those solver iterations ran on rejected correspondences
      return compute();  // recursion terminates as soon as the thresholds are saturated
    }
  }
  fill(*_correspondences, n, f, m, d);
  if (estimate_change_norm < param_maximum_estimate_change_norm_for_convergence.value() &&
      _current_iteration > param_minimum_number_of_iterations.value()) {
    _has_converged = true;
    if (matching_ratio > param_minimum_matching_ratio.value()) _adaptAfterConvergence();
  }
  ++_current_iteration;
  _postCompute();
}

int CorrespondenceFinderProjectiveCUDA::callsWithoutNewCorrespondences() const {
  if (_fixed_changed_flag || _moving_changed_flag || _config_changed) return 0;
  if (_has_converged) return 1 << 30;  // :137-141
  // the next call sees _current_iteration; it re-projects iff iteration % N == 0 || iteration == 1 (:162-164)
  const size_t N = param_number_of_solver_iterations_per_projection.value();
  int n = 0;
  for (size_t it = _current_iteration; !(it % N == 0 || it == 1); ++it) ++n;
  return n;
}

// ---- stereo adaptor ---------------------------------------------------------------------------------------------
RawDataPreprocessorStereoProjectiveCUDA::RawDataPreprocessorStereoProjectiveCUDA() {}

bool RawDataPreprocessorStereoProjectiveCUDA::setRawData(const ImageView& left, const ImageView& right) {
  _raw_set = false;
  _status = Error;
  if (!left.data)
    throw std::runtime_error("RawDataPreprocessorStereoProjective::setMeasurement|ERROR, ImageMessage not found - topic [ " +
                             param_topic_camera_left.value() + " ]");
  if (!right.data)
    throw std::runtime_error("RawDataPreprocessorStereoProjective::setMeasurement|ERROR, ImageMessage not found - topic [ " +
                             param_topic_camera_right.value() + " ]");
  _left = left;
  _right = right;
  _raw_set = true;
  _raw_data_changed_flag = true;
  _status = Initializing;
  return true;
}

void RawDataPreprocessorStereoProjectiveCUDA::compute() {
  _status = Error;
  if (!_raw_set) throw std::runtime_error("RawDataPreprocessorStereoProjective::compute|measurement not set");
  if (!_raw_data_changed_flag) {
    _status = Ready;
    return;
  }
  if (!_meas) throw std::runtime_error("RawDataPreprocessorStereoProjective::compute|destination buffer not set");
  _meas->clear();
  if (!param_feature_extractor.value() || !param_feature_extractor_right.value() || !param_correspondence_finder.value())
    throw std::runtime_error("RawDataPreprocessorStereoProjective::compute|ERROR: feature extractor / correspondence finder not set");
  auto epipolar = std::dynamic_pointer_cast<CorrespondenceFinderDescriptorBasedEpipolarCUDA>(param_correspondence_finder.value());
  IntensityFeatureExtractorBaseCUDA& ex_l = *param_feature_extractor.value();
  IntensityFeatureExtractorBaseCUDA& ex_r = *param_feature_extractor_right.value();
  const pslam_extract_cfg cl = ex_l.cudaConfig(), cr = ex_r.cudaConfig();
  const bool same_extractor = cl.detector_threshold == cr.detector_threshold &&
                              cl.enable_non_maximum_suppression == cr.enable_non_maximum_suppression &&
                              cl.target_number_of_keypoints == cr.target_number_of_keypoints &&
                              cl.number_of_detectors_horizontal == cr.number_of_detectors_horizontal &&
                              cl.number_of_detectors_vertical == cr.number_of_detectors_vertical;
  _meas->dim = 4;
  const bool binned = dynamic_cast<IntensityFeatureExtractorBinnedCUDA*>(&ex_l) && dynamic_cast<IntensityFeatureExtractorBinnedCUDA*>(&ex_r);
  if (epipolar && binned && same_extractor && _left.rows == _right.rows && _left.cols == _right.cols && _left.stride == _right.stride) {
    // the shipped wiring (kitti.conf / euroc.conf): ONE device call, features never leave HBM between the stages
    ex_l.prepare(_left.rows, _left.cols);  // the extractors' own init(): PARAM validation, same error texts
    ex_r.prepare(_right.rows, _right.cols);
    pslam_ctx* ctx = PslamDevice::context(_left.rows, _left.cols);
    const pslam_match_cfg mcfg = epipolar->cudaConfig();
    _meas->resize(kMaxFeatures);
    const int n = pslam_stereo_adaptor(ctx, _left.data, _right.data, _left.rows, _left.cols, _left.stride, &cl, &mcfg, kMaxFeatures,
                                       _meas->coordinates.data(), _meas->intensity.data(), _meas->descriptor.data());
    if (n < 0) _meas->resize(0);
    PslamDevice::check(n, "RawDataPreprocessorStereoProjective::compute");
    _meas->resize(n);
  } else {
    // general wiring (e.g. a brute-force finder, different extractors): compose the module calls (:77-132)
    PointIntensityDescriptorCloud features_left(3), features_right(3);
    ex_l.setFeatures(&features_left);
    ex_l.setProjections(_projections, _projection_radius);
    ex_l.compute(_left);
    // right frame: fine-grained detection for triangulation, using the left points as priors (:84-88)
    ex_r.setFeatures(&features_right);
    ex_r.setProjections(&features_left, _projection_radius);
    ex_r.compute(_right);
    _projections = nullptr;
    CorrespondenceVector stereo_matches;
    auto& finder = *param_correspondence_finder.value();
    finder.setFixed(&features_left);
    finder.setMoving(&features_right);
    finder.setCorrespondences(&stereo_matches);
    finder.compute();
    _meas->resize(stereo_matches.size());
    size_t k = 0;
    for (const Correspondence& c : stereo_matches) {
      const float* l = features_left.point(c.fixed_idx);
      const float* r = features_right.point(c.moving_idx);
      if (l[0] - r[0] < 0 || l[1] - r[1] < 0) continue;  // :120-128
      float* p = _meas->point(k);
      p[0] = l[0];
      p[1] = l[1];
      p[2] = r[0];
      p[3] = r[1];
      _meas->intensity[k] = features_left.intensity[c.fixed_idx];
      std::copy_n(features_left.descriptor.data() + 32 * (size_t) c.fixed_idx, 32, _meas->descriptor.data() + 32 * k);
      ++k;
    }
    _meas->resize(k);
  }
  _raw_data_changed_flag = false;
  _status = Ready;
}

// ---- monocular + depth adaptor ------------------------------------------------------------------------------------
bool RawDataPreprocessorMonocularDepthCUDA::setRawData(const ImageView& intensity, const DepthView& depth) {
  _raw_set = false;
  _status = Error;
  if (!intensity.data)
    throw std::runtime_error("RawDataPreprocessorMonocularDepth::setMeasurement|ERROR, ImageMessage not found - topic [ " +
                             param_topic_rgb.value() + " ]");
  if (!depth.data)
    throw std::runtime_error("RawDataPreprocessorMonocularDepth::setMeasurement|ERROR, ImageMessage not found - topic [ " +
                             param_topic_depth.value() + " ]");
  _intensity = intensity;
  _depth = depth;
  _raw_set = true;
  _raw_data_changed_flag = true;
  _status = Initializing;
  return true;
}

void RawDataPreprocessorMonocularDepthCUDA::compute() {
  _status = Error;
  if (!_raw_set) throw std::runtime_error("RawDataPreprocessorMonocularDepth::compute|ERROR: measurement not set");
  if (!_meas) throw std::runtime_error("RawDataPreprocessorMonocularDepth::compute|ERROR: destination buffer not set");
  if (!_raw_data_changed_flag) {
    _status = Ready;
    return;
  }
  _meas->clear();
  if (_intensity.rows == 0) throw std::runtime_error("RawDataPreprocessorMonocularDepth::compute|ERROR: intensity image has zero rows");
  if (_intensity.cols == 0) throw std::runtime_error("RawDataPreprocessorMonocularDepth::compute|ERROR: intensity image has zero columns");
  if (_depth.rows == 0) throw std::runtime_error("RawDataPreprocessorMonocularDepth::compute|ERROR: depth image has zero rows");
  if (_depth.cols == 0) throw std::runtime_error("RawDataPreprocessorMonocularDepth::compute|ERROR: depth image has zero columns");
  if (_depth.type != 0 && _depth.type != 1) throw std::runtime_error("RawDataPreprocessorMonocularDepth::compute|ERROR: unknown depth image type");
  if (!param_feature_extractor.value()) throw std::runtime_error("RawDataPreprocessorMonocularDepth::compute|ERROR: feature extractor not set");
  pslam_ctx* ctx = PslamDevice::context(_intensity.rows, _intensity.cols);
  if (!dynamic_cast<IntensityFeatureExtractorBinnedCUDA*>(param_feature_extractor.value().get()))
    throw std::runtime_error("RawDataPreprocessorMonocularDepth::compute|ERROR: the CUDA adaptor is wired for the binned feature extractor");
  const pslam_extract_cfg cfg = param_feature_extractor->cudaConfig();
  _meas->dim = 3;
  _meas->resize(kMaxFeatures);
  int n_in_image = 0;
  const int n = pslam_mono_depth_adaptor(ctx, _intensity.data, _intensity.rows, _intensity.cols, _intensity.stride, _depth.data,
                                         _depth.type, _depth.rows, _depth.cols, _depth.stride_elements,
                                         param_depth_scaling_factor_to_meters.value(), &cfg, kMaxFeatures,
                                         _meas->coordinates.data(), _meas->intensity.data(), _meas->descriptor.data(), &n_in_image);
  if (n < 0) _meas->resize(0);
  PslamDevice::check(n, "RawDataPreprocessorMonocularDepth::compute");
  _meas->resize(n);
  _raw_data_changed_flag = false;
  if (n_in_image == 0) std::cerr << "RawDataPreprocessorMonocularDepth::compute|WARNING: no features found" << std::endl;
  if (_meas->empty()) {
    std::cerr << "RawDataPreprocessorMonocularDepth::compute|WARNING: no adapted measurements generated" << std::endl;
    _status = Error;
    return;
  }
  const size_t without = (size_t) n_in_image - _meas->size();
  if (static_cast<float>(without) / n_in_image > 0.25)
    std::cerr << "RawDataPreprocessorMonocularDepth::compute|WARNING: high number of points without depth: " << without << "/"
              << n_in_image << std::endl;
  _status = Ready;
}

// ---- scene clipper (mapping/scene_clipper_projective_3d.cpp:9-67) -----------------------------------------------
void SceneClipperProjective3DCUDA::compute() {
  _status = Error;
  if (!param_projector.value()) throw std::runtime_error("SceneClipperProjective3D::compute|ERROR: missing projector");
  if (!_clipped_scene_in_robot) throw std::runtime_error("SceneClipperProjective3D::compute|ERROR: missing clipped scene");
  if (!_full_scene) throw std::runtime_error("SceneClipperProjective3D::compute|ERROR: missing global scene");
  if (_full_scene->empty()) {
    std::cerr << "SceneClipperProjective3D::compute|WARNING: global scene is empty, no clipping will be performed" << std::endl;
    _status = Ready;
    return;
  }
  if (_full_scene->dim != 3) throw std::runtime_error("SceneClipperProjective3D::compute|ERROR: the scene must hold 3D points");
  _clipped_scene_in_robot->clear();
  _global_indices.clear();
  _projections.clear();
  const ProjectorPinhole& projector = *param_projector.value();
  pslam_clip_cfg cfg;
  for (int i = 0; i < 9; ++i) cfg.K[i] = projector.cameraMatrix()[i];
  cfg.canvas_rows = (int) projector.param_canvas_rows.value();
  cfg.canvas_cols = (int) projector.param_canvas_cols.value();
  cfg.range_min = projector.param_range_min.value();
  cfg.range_max = projector.param_range_max.value();
  const Isometry3f camera = _robot_in_local_map * _sensor_in_robot;  // projector->setCameraPose(...), :46
  bool sensor_is_identity = true;
  const Isometry3f I;
  for (int i = 0; i < 12; ++i) {
    cfg.camera_in_map[i] = camera.m[i];
    cfg.sensor_in_robot[i] = _sensor_in_robot.m[i];
    if (_sensor_in_robot.m[i] != I.m[i]) sensor_is_identity = false;
  }
  cfg.apply_sensor_in_robot = sensor_is_identity ? 0 : 1;  // :60
  const int n = (int) _full_scene->size();
  pslam_ctx* ctx = PslamDevice::context(cfg.canvas_rows, cfg.canvas_cols);
  PointIntensityDescriptorCloud& out = *_clipped_scene_in_robot;
  out.dim = 3;
  out.number_of_optimizations.clear();
  out.resize(n);
  _global_indices.resize(n);
  _projections.resize((size_t) 3 * n);
  const int kept = pslam_scene_clip(ctx, n, _full_scene->coordinates.data(), _full_scene->descriptor.data(), &cfg, n,
                                    out.coordinates.data(), _projections.data(), _global_indices.data(), out.descriptor.data());
  if (kept < 0) {
    out.resize(0);
    _global_indices.clear();
    _projections.clear();
  }
  PslamDevice::check(kept, "SceneClipperProjective3D::compute");
  // the remaining fields of the survivors are copied like the projector copies the whole point (:52)
  for (int k = 0; k < kept; ++k) out.intensity[k] = _full_scene->intensity[_global_indices[k]];
  const bool has_stats = !_full_scene->number_of_optimizations.empty();
  out.resize(kept);
  if (has_stats) {
    out.number_of_optimizations.resize(kept);
    for (int k = 0; k < kept; ++k) out.number_of_optimizations[k] = _full_scene->number_of_optimizations[_global_indices[k]];
  }
  _global_indices.resize(kept);
  _projections.resize((size_t) 3 * kept);
  if (out.empty()) std::cerr << "SceneClipperProjective3D::compute|WARNING: clipped empty scene" << std::endl;
  _status = Successful;
}

// ---- landmark estimator (mapping/landmarks/landmark_estimator_ekf_impl.cpp) --------------------------------------
int LandmarkEstimatorEKFCUDA::computeBatch(int n, float* state_world, float* covariance, const float* measurements,
                                           float* coords_in_local_map, uint8_t* inlier) {
  if (!param_filter.value()) throw std::runtime_error("LandmarkEstimatorEKF::compute|ERROR: filter not set");
  const PointEKFCUDA& filter = *param_filter.value();
  if (filter.kind() != _kind) throw std::runtime_error("LandmarkEstimatorEKF::compute|ERROR: filter type does not match the measurement dimension");
  pslam_ekf_cfg cfg;
  cfg.kind = _kind;
  for (int i = 0; i < 9; ++i) cfg.K[i] = filter.cameraMatrix()[i];
  cfg.baseline_pixels[0] = filter.baseline()[0];
  cfg.baseline_pixels[1] = filter.baseline()[1];
  cfg.minimum_state_element_covariance = param_minimum_state_element_covariance.value();
  cfg.maximum_covariance_norm_squared = param_maximum_covariance_norm_squared.value();
  cfg.maximum_distance_geometry_meters_squared = param_maximum_distance_geometry_meters_squared.value();
  for (int i = 0; i < 12; ++i) {
    cfg.sensor_in_world[i] = _sensor_in_world.m[i];
    cfg.sensor_in_local_map[i] = _sensor_in_local_map.m[i];
  }
  pslam_ctx* ctx = PslamDevice::context();
  const int k = pslam_landmarks_ekf_update(ctx, n, state_world, covariance, measurements, &cfg, coords_in_local_map, inlier);
  PslamDevice::check(k, "LandmarkEstimatorEKF::compute");
  return k;
}

int LandmarkEstimatorWeightedMeanCUDA::computeBatch(int n, float* state_world, const int* number_of_optimizations,
                                                    const float* landmark_in_sensor, float* coords_in_local_map, uint8_t* inlier) {
  pslam_ctx* ctx = PslamDevice::context();
  const int k = pslam_landmarks_weighted_mean_update(ctx, n, state_world, number_of_optimizations, landmark_in_sensor, _sensor_in_world.m,
                                                     _sensor_in_local_map.m, param_maximum_distance_geometry_meters_squared.value(),
                                                     coords_in_local_map, inlier);
  PslamDevice::check(k, "LandmarkEstimatorWeightedMean::compute");
  return k;
}

pslam_merger_cfg MergerProjectiveCUDA::_cfg() const {
  if (!param_projector.value()) throw std::runtime_error("MergerProjective::compute|ERROR: projector not set");  // :21-24
  pslam_merger_cfg cfg;
  cfg.canvas_rows = (int) param_projector->param_canvas_rows.value();
  cfg.canvas_cols = (int) param_projector->param_canvas_cols.value();
  cfg.number_of_row_bins = (int) param_number_of_row_bins.value();
  cfg.number_of_col_bins = (int) param_number_of_col_bins.value();
  cfg.maximum_distance_appearance = param_maximum_distance_appearance.value();
  cfg.enable_binning = param_enable_binning.value() ? 1 : 0;
  cfg.kind = _kind;
  if (cfg.number_of_row_bins < 1 || (float) cfg.canvas_rows / (float) cfg.number_of_row_bins < 1)  // :36-41
    throw std::runtime_error("MergerProjective::compute|ERROR: row bin width must be at least 1 pixel - reduce param_number_of_row_bins");
  if (cfg.number_of_col_bins < 1 || (float) cfg.canvas_cols / (float) cfg.number_of_col_bins < 1)  // :42-47
    throw std::runtime_error("MergerProjective::compute|ERROR: col bin width must be at least 1 pixel - reduce param_number_of_col_bins");
  return cfg;
}

int MergerProjectiveCUDA::selectUpdates(const float* measurements, int dim, int n_meas, const int* corr_moving,
                                        const float* corr_response, int n_corr, uint8_t* selected) {
  const pslam_merger_cfg cfg = _cfg();
  _occupied.assign((size_t) pslam_merger_occupancy_words(&cfg), 0u);
  const int k = pslam_merger_select_updates(PslamDevice::context(), measurements, dim, n_meas, corr_moving, corr_response, n_corr, &cfg,
                                            selected, _occupied.data());
  PslamDevice::check(k, "MergerProjective::compute");
  return k;
}

int MergerProjectiveCUDA::plan(const float* measurements, int dim, int n_meas, const int* corr_moving, const float* corr_response,
                               int n_corr, uint8_t* selected, int* winners, int* n_winners) {
  if (param_enable_conservative_addition.value()) throw std::runtime_error("conservative addition is currently disabled");  // :262-264
  const pslam_merger_cfg cfg = _cfg();
  _occupied.assign((size_t) pslam_merger_occupancy_words(&cfg), 0u);
  const int k = pslam_merger_plan(PslamDevice::context(), measurements, dim, n_meas, corr_moving, corr_response, n_corr, &cfg, selected,
                                  _occupied.data(), winners, n_winners);
  PslamDevice::check(k, "MergerProjective::compute");
  return k;
}

bool MergerProjectiveCUDA::wantsAdditions(int number_of_merged_points, int n_meas, int n_corr) const {
  if (n_corr == 0) return true;
  return number_of_merged_points < (int) param_target_number_of_merges.value() && number_of_merged_points < n_meas;
}

int MergerProjectiveCUDA::selectAdditions(const float* measurements, int dim, int n_meas, int* winners) {
  if (param_enable_conservative_addition.value()) throw std::runtime_error("conservative addition is currently disabled");  // :262-264
  const pslam_merger_cfg cfg = _cfg();
  const bool have = _occupied.size() == (size_t) pslam_merger_occupancy_words(&cfg);
  const int k = pslam_merger_select_additions(PslamDevice::context(), measurements, dim, n_meas, &cfg, have ? _occupied.data() : nullptr,
                                              winners);
  PslamDevice::check(k, "MergerProjective::_addPoints");
  return k;
}

int LandmarkEstimatorPoseBasedSmootherCUDA::computeBatch(int n, float* state_world, int* number_of_optimizations, int n_frames,
                                                         const float* frames_sensor_in_world, const int* offsets, const int* hist_frame,
                                                         const float* hist_uv, const float* hist_point_in_camera,
                                                         float* coords_in_local_map, uint8_t* inlier) {
  pslam_smoother_cfg cfg;
  for (int i = 0; i < 9; ++i) cfg.K[i] = _K[i];
  cfg.maximum_number_of_iterations = param_maximum_number_of_iterations.value();
  cfg.convergence_criterion_minimum_chi2_delta = param_convergence_criterion_minimum_chi2_delta.value();
  cfg.maximum_reprojection_error_pixels_squared = param_maximum_reprojection_error_pixels_squared.value();
  cfg.minimum_number_of_measurements_for_optimization = param_minimum_number_of_measurements_for_optimization.value();
  cfg.maximum_distance_geometry_meters_squared = param_maximum_distance_geometry_meters_squared.value();
  for (int i = 0; i < 12; ++i) {
    cfg.sensor_in_world[i] = _sensor_in_world.m[i];
    cfg.sensor_in_local_map[i] = _sensor_in_local_map.m[i];
  }
  pslam_ctx* ctx = PslamDevice::context();
  const int k = pslam_landmarks_smoother_update(ctx, n, state_world, number_of_optimizations, n_frames, frames_sensor_in_world, offsets,
                                                hist_frame, hist_uv, hist_point_in_camera, &cfg, coords_in_local_map, inlier);
  PslamDevice::check(k, "LandmarkEstimatorPoseBasedSmoother::compute");
  return k;
}

// ---- aligner slice ----------------------------------------------------------------------------------------------
AlignerSliceProcessorProjectiveCUDA::AlignerSliceProcessorProjectiveCUDA(int kind) : _kind(kind) {
  // aligner_slice_processor_projective.cpp:7-20: saturated robustifier with chi threshold 100^2 by default
  auto r = std::make_shared<RobustifierSaturated>();
  r->setName("aligner_robustifier");
  r->param_chi_threshold.setValue(100 * 100);
  param_robustifier.setValue(r);
  setName("aligner_slice_projective");
  param_diagonal_info_matrix.setValue(std::vector<float>(kind == 2 ? 2 : 3, 0.f));  // DiagInfoVector::Zero()
}

void AlignerSliceProcessorProjectiveCUDA::bindFixed() {
  if (_kind != 0) return;
  // stereo: mean disparity over the fixed slice, fp32 accumulation in cloud order (.cpp:78-88)
  if (_fixed_slice && !_fixed_slice->empty()) {
    float accumulated_disparity = 0;
    for (size_t i = 0; i < _fixed_slice->size(); ++i) accumulated_disparity += _fixed_slice->point(i)[0] - _fixed_slice->point(i)[2];
    _mean_disparity = accumulated_disparity / _fixed_slice->size();
  } else {
    _mean_disparity = 0;
  }
}

void AlignerSliceProcessorProjectiveCUDA::diagonalInfo(float d[3]) const {
  const std::vector<float>& diag = param_diagonal_info_matrix.value();
  if ((int) diag.size() != (_kind == 2 ? 2 : 3)) throw std::runtime_error("AlignerSliceProjective_|diagonal_info_matrix has the wrong dimension");
  d[0] = diag[0];
  d[1] = diag[1];
  d[2] = diag.size() > 2 ? diag[2] : 0.f;
}

void AlignerSliceProcessorProjectiveCUDA::setupFactorConfig() {
  if (!_fixed_slice || !_moving_slice) throw std::runtime_error("AlignerSliceProjective_|fixed / moving slice not set");
  if (!param_projector.value()) throw std::runtime_error("AlignerSliceProjective_|projector not set");
  const ProjectorPinhole& projector = *param_projector.value();
  _factor.kind = _kind;
  for (int i = 0; i < 9; ++i) _factor.K[i] = projector.cameraMatrix()[i];
  _factor.image_cols = projector.param_canvas_cols.value();
  _factor.image_rows = projector.param_canvas_rows.value();
  RobustifierBase* rob = param_robustifier.value().get();
  _factor.robustifier = rob ? rob->kind() : 0;
  _factor.chi_threshold = rob ? rob->param_chi_threshold.value() : 0.0;
  _factor.baseline[0] = _factor.baseline[1] = _factor.baseline[2] = 0;
  _factor.mean_disparity = 0;
  if (_kind == 0) {
    if (!_baseline_set) {  // baseline [pixel * m] = K * t_left_in_right, fp32 (.cpp:93-103)
      const std::array<float, 9>& K = projector.cameraMatrix();
      for (int i = 0; i < 3; ++i)
        _baseline_left_in_right_pixelsmeters[i] =
          K[3 * i] * _t_left_in_right[0] + K[3 * i + 1] * _t_left_in_right[1] + K[3 * i + 2] * _t_left_in_right[2];
      _baseline_set = true;
    }
    for (int i = 0; i < 3; ++i) _factor.baseline[i] = _baseline_left_in_right_pixelsmeters[i];
    if (param_enable_inverse_depth_weighting.value()) _factor.mean_disparity = _mean_disparity;  // (.cpp:107-112)
  }
}

void AlignerSliceProcessorProjectiveCUDA::setupFactor() {
  setupFactorConfig();
  float diag[3];
  diagonalInfo(diag);
  // per-correspondence information, indexed by the FIXED index (.cpp:41-57)
  _fixed_information_diagonals.resize(3 * _fixed_slice->size());
  constexpr size_t minimum_number_of_updates = 2;
  for (const Correspondence& c : _correspondences) {
    float d[3] = {diag[0], diag[1], diag[2]};
    const int n_opt = _moving_slice->number_of_optimizations.empty() ? 0 : _moving_slice->number_of_optimizations[c.moving_idx];
    if ((size_t) n_opt > minimum_number_of_updates) {
      const float s = (float) (1 + std::log((double) n_opt));  // Vector3f *= (1 + std::log(size_t)): double, then Scalar
      for (float& x : d) x *= s;
    }
    for (int k = 0; k < 3; ++k) _fixed_information_diagonals[3 * (size_t) c.fixed_idx + k] = d[k];
  }
}

// ---- aligner ----------------------------------------------------------------------------------------------------
void MultiAligner3DQRCUDA::setMovingInFixed(const Isometry3f& T) {
  for (int i = 0; i < 12; ++i) _estimate[i] = T.m[i];
}

AlignerSliceProcessorProjectiveCUDA* MultiAligner3DQRCUDA::projectiveSlice() const {
  for (const auto& s : param_slice_processors.value())
    if (auto p = std::dynamic_pointer_cast<AlignerSliceProcessorProjectiveCUDA>(s)) return p.get();
  return nullptr;
}

AlignerSliceMotionModel3DCUDA* MultiAligner3DQRCUDA::motionModelSlice() const {
  for (const auto& s : param_slice_processors.value())
    if (auto p = std::dynamic_pointer_cast<AlignerSliceMotionModel3DCUDA>(s)) return p.get();
  return nullptr;
}

Isometry3f MotionModelConstantVelocity3D::predict(const std::vector<Isometry3f>& chunk) const {
  if (chunk.empty()) return Isometry3f::Identity();  // no history: no motion
  const Isometry3f& last = chunk.back();
  if (chunk.size() == 1) return last.inverse();
  const Isometry3f motion = chunk[chunk.size() - 2].inverse() * last;  // last relative motion, applied once more
  return (last * motion).inverse();
}

AlignerSliceMotionModel3DCUDA::AlignerSliceMotionModel3DCUDA() {
  for (int i = 0; i < 36; ++i) _information[i] = (i % 7 == 0) ? 1.0 : 0.0;
}

pslam_pose_prior AlignerSliceMotionModel3DCUDA::prior() const {
  pslam_pose_prior p;
  const MotionModelConstantVelocity3D fallback;
  const MotionModelConstantVelocity3D* mm = param_motion_model.value() ? param_motion_model.value().get() : &fallback;
  const Isometry3f Z = mm->predict(_chunk);
  for (int i = 0; i < 12; ++i) p.prediction[i] = Z.m[i];
  for (int i = 0; i < 36; ++i) p.information[i] = _information[i];
  return p;
}

void MultiAligner3DQRCUDA::compute() {
  _status = Fail;
  _stats.clear();
  _inlier_run_stats.clear();
  AlignerSliceProcessorProjectiveCUDA* slice = projectiveSlice();
  if (!slice) throw std::runtime_error("MultiAligner::compute|ERROR: no projective slice processor configured");
  if (!_fixed || !_moving) throw std::runtime_error("MultiAligner::compute|ERROR: fixed / moving not set");
  if (!slice->param_finder.value()) throw std::runtime_error("MultiAligner::compute|ERROR: slice has no correspondence finder");
  if (!param_solver.value()) throw std::runtime_error("MultiAligner::compute|ERROR: solver not set");
  if (_fixed->dim != slice->fixedDim()) throw std::runtime_error("MultiAligner::compute|ERROR: fixed cloud dimension does not match the slice's factor");
  pslam_ctx* ctx = PslamDevice::context();
  CorrespondenceFinderBase& finder = *slice->param_finder.value();
  slice->setFixed(_fixed);
  slice->setMoving(_moving);
  slice->bindFixed();
  finder.setFixed(_fixed);
  finder.setMoving(_moving);
  finder.setCorrespondences(&slice->correspondences());
  const double damping = param_solver->damping();
  // second slice: the motion model seeds the estimate and adds its pose-prior factor to every iteration
  pslam_pose_prior prior_storage;
  const pslam_pose_prior* prior = nullptr;
  if (AlignerSliceMotionModel3DCUDA* mm = motionModelSlice()) {
    prior_storage = mm->prior();
    prior = &prior_storage;
    for (int i = 0; i < 12; ++i) _estimate[i] = prior_storage.prediction[i];
  }

  // the factor reads the clouds as they are (fp32, the reference's scalar): no widening on the host, 40 B / correspondence
  const float* moving_xyz = _moving->coordinates.data();
  const float* fixed_meas = _fixed->coordinates.data();
  std::vector<int> cf, cm;
  std::vector<double> poses, stats;
  std::vector<uint8_t> factor_status;
  AlignerIterationStats last;
  bool enough = true;
  const int max_iterations = param_max_iterations.value();
  auto push_stats = [&](std::vector<AlignerIterationStats>& dst, int it, int done, int n_corr) {
    for (int j = 0; j < done; ++j) {
      AlignerIterationStats st;
      st.iteration = it + j;
      st.num_correspondences = n_corr;
      st.chi = stats[4 * (size_t) j];
      st.num_inliers = (int) stats[4 * (size_t) j + 1];
      st.num_outliers = (int) stats[4 * (size_t) j + 2];
      st.num_suppressed = (int) stats[4 * (size_t) j + 3];
      dst.push_back(st);
      last = st;
    }
  };
  auto split = [&](const CorrespondenceVector& corr) {
    cf.resize(corr.size());
    cm.resize(corr.size());
    for (size_t k = 0; k < corr.size(); ++k) {
      cf[k] = corr[k].fixed_idx;
      cm[k] = corr[k].moving_idx;
    }
  };
  // When the finder is one of the projective ones, the solver iterations that follow a search run in the SAME device round
  // trip as the search (pslam_projective_match_gn): the finder's cached fp32 clouds are the factor's clouds, the information
  // of a correspondence is diagonal x the per-point scale below (setupFactor's 1 + log(n_opt) weighting).
  FusedSolveRequest fused;
  std::vector<float> moving_scale;
  const bool can_fuse = dynamic_cast<CorrespondenceFinderProjectiveCUDA*>(&finder) != nullptr;
  if (can_fuse) {
    slice->setupFactorConfig();
    fused.factor = &slice->factorConfig();
    slice->diagonalInfo(fused.diagonal_info);
    fused.damping = damping;
    fused.prior = prior;
    moving_scale.assign(_moving->size(), 1.0f);
    if (!_moving->number_of_optimizations.empty())
      for (size_t i = 0; i < _moving->size(); ++i) {
        const int n_opt = _moving->number_of_optimizations[i];
        if (n_opt > 2) moving_scale[i] = (float) (1 + std::log((double) n_opt));
      }
    fused.moving_scale = moving_scale.data();
  }
  int it = 0;
  bool solve_failed = false;
  // The whole loop below, device resident (pslam_projective_align): the finder's state machine runs between the searches on
  // the device, one download per batch of search phases.  It hands back with `it` iterations done wherever a decision needs
  // this loop (repeat with a wider search, too few correspondences); PSLAM_ALIGN_DEVICE=0 keeps the call-by-call path.
  const char* align_env = std::getenv("PSLAM_ALIGN_DEVICE");
  if (can_fuse && !(align_env && align_env[0] == '0')) {
    fused.executed = false;
    fused.max_fused = max_iterations;
    for (int i = 0; i < 12; ++i) fused.estimate[i] = _estimate[i];
    if (finder.alignFused(fused, std::max(slice->param_min_num_correspondences.value(), 1)) && fused.executed) {
      for (size_t p = 0; 3 * p + 2 < fused.phase_log.size(); ++p) {
        const int start = fused.phase_log[3 * p], done = fused.phase_log[3 * p + 1];
        stats.assign(fused.stats.begin() + 4 * (size_t) start, fused.stats.begin() + 4 * (size_t) (start + done));
        push_stats(_stats, start, done, fused.phase_log[3 * p + 2]);
      }
      if (fused.done > 0)
        for (int i = 0; i < 12; ++i) _estimate[i] = fused.poses[12 * (size_t) (fused.done - 1) + i];
      factor_status = fused.status;
      it = fused.done;
      solve_failed = fused.stop_reason == 4;
    }
  }
  for (; it < max_iterations && !solve_failed;) {
    Isometry3f X;
    for (int i = 0; i < 12; ++i) X.m[i] = (float) _estimate[i];
    finder.setLocalMapInSensor(X);
    if (can_fuse) {
      fused.executed = false;
      fused.max_fused = max_iterations - it;
      for (int i = 0; i < 12; ++i) fused.estimate[i] = _estimate[i];
      finder.setFusedSolve(&fused);
    }
    finder.compute();
    finder.setFusedSolve(nullptr);
    const CorrespondenceVector& corr = slice->correspondences();
    if ((int) corr.size() < std::max(slice->param_min_num_correspondences.value(), 1)) {
      AlignerIterationStats st;
      st.iteration = it;
      st.num_correspondences = (int) corr.size();
      enough = false;
      _stats.push_back(st);
      break;
    }
    int done = 0, rc = PSLAM_OK;
    if (can_fuse && fused.executed) {
      done = fused.done;
      poses = fused.poses;
      stats = fused.stats;
      factor_status = fused.status;
      if (done > 0)
        for (int i = 0; i < 12; ++i) _estimate[i] = poses[12 * (size_t) (done - 1) + i];
      if (!fused.spd) rc = PSLAM_E_NOT_SPD;
    } else {
      slice->setupFactor();
      split(corr);
      // The finder keeps these correspondences for its next `quiet` calls (it only re-projects every N-th solver
      // iteration, correspondence_finder_projective_base_impl.cpp:162-178): run this iteration and those in ONE launch.
      const int quiet = finder.callsWithoutNewCorrespondences();
      const int n_fused = std::min(max_iterations - it, quiet >= max_iterations ? max_iterations : quiet + 1);
      poses.resize(12 * (size_t) n_fused);
      stats.resize(4 * (size_t) n_fused);
      factor_status.resize(corr.size());
      rc = pslam_gn_iterate_f32(ctx, &slice->factorConfig(), n_fused, damping, _estimate.data(), (int) _moving->size(),
                                moving_xyz, (int) _fixed->size(), fixed_meas, _fixed->dim, (int) corr.size(), cf.data(),
                                cm.data(), slice->informationDiagonals().data(), prior, poses.data(), stats.data(),
                                factor_status.data(), &done);
    }
    push_stats(_stats, it, done, (int) corr.size());
    if (rc == PSLAM_E_NOT_SPD) break;  // degenerate system: keep the last estimate
    PslamDevice::check(rc, "MultiAligner::compute");
    // the finder's own bookkeeping for the fused iterations (iteration counter, previous estimate): host only
    for (int j = 1; j < done; ++j) {
      for (int i = 0; i < 12; ++i) X.m[i] = (float) poses[12 * (size_t) (j - 1) + i];
      finder.setLocalMapInSensor(X);
      finder.compute();
    }
    it += done;
  }
  if (!enough) {
    _status = NotEnoughCorrespondences;
    return;
  }
  _status = last.num_inliers >= param_min_num_inliers.value() ? Success : NotEnoughInliers;
  if (_status != Success) return;
  // configurations/icl.conf:55-58 [upstream, restated in oracle_lib.align]: the correspondences whose factor was an inlier
  // in the last linearisation are frozen; with enough of them max_iterations further GN iterations run on them alone (ONE
  // launch, the finder is not consulted); keep_only_inlier_correspondences leaves exactly those in the slice
  const bool inlier_runs = param_enable_inlier_only_runs.value(), keep_only = param_keep_only_inlier_correspondences.value();
  if (!inlier_runs && !keep_only) return;
  CorrespondenceVector& corr = slice->correspondences();
  auto keep_inliers = [&]() {
    CorrespondenceVector kept;
    for (size_t k = 0; k < corr.size(); ++k)
      if (factor_status[k] == PSLAM_FACTOR_INLIER) kept.push_back(corr[k]);
    return kept;
  };
  CorrespondenceVector inliers = keep_inliers();
  if (inlier_runs && (int) inliers.size() >= param_min_num_inliers.value()) {
    const CorrespondenceVector all = corr;
    corr = inliers;
    slice->setupFactor();
    split(corr);
    poses.resize(12 * (size_t) max_iterations);
    stats.resize(4 * (size_t) max_iterations);
    factor_status.resize(corr.size());
    int done = 0;
    const int rc = pslam_gn_iterate_f32(ctx, &slice->factorConfig(), max_iterations, damping, _estimate.data(), (int) _moving->size(),
                                        moving_xyz, (int) _fixed->size(), fixed_meas, _fixed->dim, (int) corr.size(), cf.data(),
                                        cm.data(), slice->informationDiagonals().data(), prior, poses.data(), stats.data(),
                                        factor_status.data(), &done);
    push_stats(_inlier_run_stats, 0, done, (int) corr.size());
    if (rc != PSLAM_E_NOT_SPD) PslamDevice::check(rc, "MultiAligner::compute (inlier-only run)");
    if (keep_only) corr = keep_inliers();
    else corr = all;
  } else if (keep_only) {
    corr = inliers;
  }
}

// ---- registration -----------------------------------------------------------------------------------------------
namespace {
template <int Dim>
struct ExtractorD : IntensityFeatureExtractorBinnedCUDA {
  ExtractorD() : IntensityFeatureExtractorBinnedCUDA(Dim) {}
};
template <int Dim>
struct SelectiveD : IntensityFeatureExtractorSelectiveCUDA {
  SelectiveD() : IntensityFeatureExtractorSelectiveCUDA(Dim) {}
};
template <int Shape>
struct ProjectiveS : CorrespondenceFinderProjectiveCUDA {
  ProjectiveS() : CorrespondenceFinderProjectiveCUDA(Shape) {}
};
// correspondence_finder_projective_kdtree.h:24-28: the cluster size only shapes the reference's (approximate) tree; the
// device search is the exact radius query (pslam_cuda.h, shape 3)
struct ProjectiveKD : CorrespondenceFinderProjectiveCUDA {
  ProjectiveKD() : CorrespondenceFinderProjectiveCUDA(3) {}
  PARAM(PropertyUnsignedInt, minimum_number_of_points_per_cluster, "minimum number of points in the clusters", 10, nullptr);
};
template <int Kind>
struct SliceK : AlignerSliceProcessorProjectiveCUDA {
  SliceK() : AlignerSliceProcessorProjectiveCUDA(Kind) {}
};
template <int Kind>
struct FilterK : PointEKFCUDA {
  FilterK() : PointEKFCUDA(Kind) {}
};
template <int Kind>
struct EstimatorK : LandmarkEstimatorEKFCUDA {
  EstimatorK() : LandmarkEstimatorEKFCUDA(Kind) {}
};
template <int Kind>
struct MergerK : MergerProjectiveCUDA {
  MergerK() : MergerProjectiveCUDA(Kind) {}
};
template <typename T>
void reg(const std::string& reference_name) {
  PSLAM_REGISTER_CLASS_AS(T, reference_name);           // the unchanged .conf selects the CUDA-backed class
  PSLAM_REGISTER_CLASS_AS(T, reference_name + "CUDA");  // explicit name, SURVEY.md 8b
}
}  // namespace

void registerTypes() {
  static bool done = false;
  if (done) return;
  done = true;
  // sensor_processing/instances.cpp:11-17
  reg<ExtractorD<2>>("IntensityFeatureExtractorBinned2D");
  reg<ExtractorD<3>>("IntensityFeatureExtractorBinned3D");
  reg<SelectiveD<2>>("IntensityFeatureExtractorSelective2D");
  reg<SelectiveD<3>>("IntensityFeatureExtractorSelective3D");
  reg<RawDataPreprocessorStereoProjectiveCUDA>("RawDataPreprocessorStereoProjective");
  reg<RawDataPreprocessorMonocularDepthCUDA>("RawDataPreprocessorMonocularDepth");
  // registration/instances.cpp:51-76
  for (const char* dims : {"2D2D", "2D3D", "3D3D", "4D3D"}) {
    reg<CorrespondenceFinderDescriptorBasedBruteforceCUDA>(std::string("CorrespondenceFinderDescriptorBasedBruteforce") + dims);
  }
  for (const char* dims : {"2D2D", "3D3D"}) reg<CorrespondenceFinderDescriptorBasedEpipolarCUDA>(std::string("CorrespondenceFinderDescriptorBasedEpipolar") + dims);
  for (const char* dims : {"2D3D", "3D3D", "4D3D"}) {
    reg<ProjectiveS<0>>(std::string("CorrespondenceFinderProjectiveSquare") + dims);
    reg<ProjectiveS<1>>(std::string("CorrespondenceFinderProjectiveCircle") + dims);
    reg<ProjectiveS<2>>(std::string("CorrespondenceFinderProjectiveRhombus") + dims);
    reg<ProjectiveKD>(std::string("CorrespondenceFinderProjectiveKDTree") + dims);
  }
  reg<SliceK<2>>("AlignerSliceProcessorProjective");
  reg<SliceK<1>>("AlignerSliceProcessorProjectiveDepth");
  reg<SliceK<0>>("AlignerSliceProcessorProjectiveStereo");
  reg<SliceK<0>>("AlignerSliceProcessorProjectiveStereoWithSensor");  // kitti_in_baselink.conf
  reg<MultiAligner3DQRCUDA>("MultiAligner3DQR");
  reg<AlignerSliceMotionModel3DCUDA>("AlignerSliceMotionModel3D");  // srrg2_slam_interfaces; second slice of the shipped aligners
  PSLAM_REGISTER_CLASS_AS(MotionModelConstantVelocity3D, "MotionModelConstantVelocity3D");
  reg<SceneClipperProjective3DCUDA>("SceneClipperProjective3D");  // mapping/instances.cpp
  reg<FilterK<0>>("ProjectivePointEKF3D");
  reg<FilterK<1>>("ProjectiveDepthPointEKF3D");
  reg<FilterK<2>>("StereoProjectivePointEKF3D");
  for (const char* dims : {"2D3D", "3D3D", "4D3D"}) reg<LandmarkEstimatorWeightedMeanCUDA>(std::string("LandmarkEstimatorWeightedMean") + dims);
  for (const char* dims : {"2D3D", "3D3D", "4D3D"}) reg<LandmarkEstimatorPoseBasedSmootherCUDA>(std::string("LandmarkEstimatorPoseBasedSmoother") + dims);
  reg<MergerK<PSLAM_MERGER_STEREO>>("MergerRigidStereoTriangulation");  // mapping/instances.cpp:45-48
  reg<MergerK<PSLAM_MERGER_STEREO>>("MergerRigidStereoProjectiveEKF");
  reg<MergerK<PSLAM_MERGER_DEPTH>>("MergerProjectiveDepthEKF");
  reg<EstimatorK<0>>("LandmarkEstimatorProjectiveEKF3D");
  reg<EstimatorK<1>>("LandmarkEstimatorProjectiveDepthEKF3D");
  reg<EstimatorK<2>>("LandmarkEstimatorStereoProjectiveEKF3D");
  // srrg2_core / srrg2_solver modules the hot-path classes link to
  PSLAM_REGISTER_CLASS_AS(ProjectorPinhole, "PointIntensityDescriptor3fProjectorPinhole");
  PSLAM_REGISTER_CLASS_AS(RobustifierSaturated, "RobustifierSaturated");
  PSLAM_REGISTER_CLASS_AS(RobustifierClamp, "RobustifierClamp");
  PSLAM_REGISTER_CLASS_AS(IterationAlgorithmGN, "IterationAlgorithmGN");
  PSLAM_REGISTER_CLASS_AS(Solver, "Solver");
}

}  // namespace pslam_host
