// pslam_plugin.hpp -- host-side C++ mirror of the srrg2_proslam plugin classes on the frontend hot path.
//
// Same class names (so the shipped .conf files select them), same PARAM names / defaults, same call contract
// (setFixed / setMoving / setCorrespondences / compute, setFeatures / compute(image), setRawData-like setters,
// _status), same std::runtime_error texts.  Every compute() forwards the data-parallel work to the sm_100a
// kernels through the C ABI of include/pslam_cuda.h -- there is no CPU implementation of the arithmetic here;
// without a CUDA device PslamDevice::context() throws.  What stays on the host is what the reference keeps in
// its control flow: change flags, the projective finder's adaptive state machine, the aligner iteration loop.
//
// Point clouds are SoA mirrors of srrg2_core::PointIntensityDescriptor{2,3,4}fVectorCloud.
#pragma once
#include <array>
#include <cstdint>
#include <iostream>
#include <memory>
#include <string>
#include <vector>

#include "../../include/pslam_cuda.h"
#include "pslam_boss.hpp"

namespace pslam_host {

// srrg2_core::Correspondence(fixed_idx, moving_idx, response)
struct Correspondence {
  int fixed_idx = -1, moving_idx = -1;
  float response = 0;
};
using CorrespondenceVector = std::vector<Correspondence>;

// PointIntensityDescriptor_<Dim, float> vector cloud: field<0> coordinates, <1> intensity, <2> descriptor
struct PointIntensityDescriptorCloud {
  int dim = 3;
  std::vector<float> coordinates;       // [n][dim]
  std::vector<float> intensity;         // [n]
  std::vector<uint8_t> descriptor;      // [n][32]  (row of the reference's cv::Mat)
  std::vector<int> number_of_optimizations;  // statistics().numberOfOptimizations(), empty = all 0
  explicit PointIntensityDescriptorCloud(int dim_ = 3) : dim(dim_) {}
  size_t size() const { return intensity.size(); }
  bool empty() const { return intensity.empty(); }
  void clear() { resize(0); }
  void resize(size_t n) {
    coordinates.resize(n * dim);
    intensity.resize(n);
    descriptor.resize(n * 32);
    if (!number_of_optimizations.empty()) number_of_optimizations.resize(n);
  }
  float* point(size_t i) { return coordinates.data() + i * dim; }
  const float* point(size_t i) const { return coordinates.data() + i * dim; }
};

// row-major 3x4 [R|t] isometry, float like the reference's Isometry3f
struct Isometry3f {
  float m[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
  static Isometry3f Identity() { return Isometry3f(); }
  void setIdentity() { *this = Isometry3f(); }
  Isometry3f inverse() const;
  Isometry3f operator*(const Isometry3f& o) const;
};
// geometry3d::t2tnq: translation + vector part of the normalised quaternion (w >= 0)
void t2tnq(const Isometry3f& T, float v6[6]);

// an 8-bit single channel image view (cv::Mat CV_8UC1 / srrg2_core::ImageUInt8)
struct ImageView {
  const uint8_t* data = nullptr;
  int rows = 0, cols = 0, stride = 0;
};
// depth image view: TYPE_16UC1 or TYPE_32FC1
struct DepthView {
  const void* data = nullptr;
  int rows = 0, cols = 0, stride_elements = 0;
  int type = 0;  // 0 = u16, 1 = f32, anything else = unknown
};

// process-wide device context shared by all modules (the reference is single threaded)
class PslamDevice {
public:
  static pslam_ctx* context(int rows = 0, int cols = 0);  // grows on demand; throws std::runtime_error without a GPU
  static void release();
  static void setDevice(int device);
  static void check(int rc, const char* where);  // negative rc -> std::runtime_error(pslam_last_error)
};

// ---- PointIntensityDescriptor3fProjectorPinhole (srrg2_core; configurations/kitti.conf:164-179) -------------
class ProjectorPinhole : public Configurable {
public:
  PARAM(PropertyUnsignedInt, canvas_cols, "canvas width [pixels]", 0, nullptr);
  PARAM(PropertyUnsignedInt, canvas_rows, "canvas height [pixels]", 0, nullptr);
  PARAM(PropertyFloat, range_max, "maximum range [m]", 1000.f, nullptr);
  PARAM(PropertyFloat, range_min, "minimum range [m]", 0.3f, nullptr);
  void setCameraMatrix(const std::array<float, 9>& K) { _K = K; }
  const std::array<float, 9>& cameraMatrix() const { return _K; }

private:
  std::array<float, 9> _K{{1, 0, 0, 0, 1, 0, 0, 0, 1}};
};

// ---- srrg2_solver pieces the .conf wires into the aligner ----------------------------------------------------
class RobustifierBase : public Configurable {
public:
  PARAM(PropertyFloat, chi_threshold, "threshold of chi after which the kernel is active", 1.f, nullptr);
  virtual int kind() const = 0;  // pslam_linearize_cfg.robustifier
};
class RobustifierSaturated : public RobustifierBase {
public:
  int kind() const override { return 1; }
};
class RobustifierClamp : public RobustifierBase {
public:
  int kind() const override { return 2; }
};
class IterationAlgorithmGN : public Configurable {
public:
  PARAM(PropertyFloat, damping, "damping factor, the higher the closer to gradient descend. Default:0", 0.f, nullptr);
};
class Solver : public Configurable {
public:
  PARAM(PropertyConfigurable_<Configurable>, algorithm, "pointer to the optimization algorithm (GN/LM or others)",
        std::static_pointer_cast<Configurable>(std::make_shared<IterationAlgorithmGN>()), nullptr);
  PARAM(PropertyVector_<int>, max_iterations, "maximum iterations if no stopping criteria is set", std::vector<int>(), nullptr);
  float damping() const;
};

// ---- IntensityFeatureExtractor_ (sensor_processing/feature_extractors/intensity_feature_extractor_base.{h,cpp}) and its
//      two implementations ..._binned.{h,cpp}, ..._selective.{h,cpp} ------------------------------------------------
class IntensityFeatureExtractorBaseCUDA : public Configurable {
public:
  explicit IntensityFeatureExtractorBaseCUDA(int point_dim = 3) : _point_dim(point_dim) {}
  PARAM(PropertyString, descriptor_type, "OpenCV descriptor type (BRIEF-256, BRIEF-512, ORB-256, FREAK-512, ..)", "ORB-256", &_config_changed);
  PARAM(PropertyString, detector_type, "OpenCV detector type for point tracking (FAST, MSER, GFTT, BRISK-512, ORB-256, ..)", "FAST", &_config_changed);
  PARAM(PropertyFloat, detector_threshold, "scalar that maps to the first parameter of the chosen detector (if applicable)", 10, &_config_changed);
  PARAM(PropertyFloat, target_bin_width_pixels, "minimum required distance between detected features (if applicable)", 10, &_config_changed);
  PARAM(PropertyBool, enable_non_maximum_suppression, "enables non maximum suppression (filtering) of detected features (if applicable)", true, &_config_changed);
  PARAM(PropertyInt, target_number_of_keypoints, "target number of keypoints to detect (accumulative over all detectors)", 500, &_config_changed);

  virtual void init() = 0;
  void clear() { _config_changed = true; }
  void setFeatures(PointIntensityDescriptorCloud* features) { _features = features; }
  // base.h:100-106: projections enable masked detection (selective extractor; ignored by the binned one)
  void setProjections(const PointIntensityDescriptorCloud* projections, size_t projection_detection_radius) {
    _projections = projections;
    _projection_detection_radius = projection_detection_radius;
  }
  void setKeypointDetectionMask(const ImageView& mask) {  // base.h:128-132
    _mask = mask;
    _mask_set = mask.data != nullptr;
  }
  void compute(const ImageView& image);  // base.cpp:55-85
  void prepare(int rows, int cols);      // latch the image size and (re-)run init()
  size_t imageRows() const { return _image_rows; }
  size_t imageCols() const { return _image_cols; }
  virtual pslam_extract_cfg cudaConfig() const;
  int pointDim() const { return _point_dim; }

protected:
  void checkDetectorAndDescriptor() const;  // base.cpp:105-176
  // device pass: fills xy / response / intensity / descriptor, returns the feature count
  virtual int extract(pslam_ctx* ctx, const ImageView& image, int capacity, float* xy, float* response, float* intensity,
                      uint8_t* desc) = 0;
  int _point_dim;
  bool _config_changed = true;
  size_t _image_rows = 0, _image_cols = 0;
  PointIntensityDescriptorCloud* _features = nullptr;
  const PointIntensityDescriptorCloud* _projections = nullptr;
  size_t _projection_detection_radius = 0;
  ImageView _mask;
  bool _mask_set = false;
};

class IntensityFeatureExtractorBinnedCUDA : public IntensityFeatureExtractorBaseCUDA {
public:
  explicit IntensityFeatureExtractorBinnedCUDA(int point_dim = 3) : IntensityFeatureExtractorBaseCUDA(point_dim) {}
  PARAM(PropertyInt, number_of_detectors_horizontal, "number of detectors on the horizontal image axis (cols)", 3, &_config_changed);
  PARAM(PropertyInt, number_of_detectors_vertical, "number of detectors on the vertical image axis (rows)", 3, &_config_changed);
  void init() override;  // binned.cpp:7-106: validates the configuration (same error texts)
  pslam_extract_cfg cudaConfig() const override;

protected:
  int extract(pslam_ctx* ctx, const ImageView& image, int capacity, float* xy, float* response, float* intensity,
              uint8_t* desc) override;
};

class IntensityFeatureExtractorSelectiveCUDA : public IntensityFeatureExtractorBaseCUDA {
public:
  explicit IntensityFeatureExtractorSelectiveCUDA(int point_dim = 3) : IntensityFeatureExtractorBaseCUDA(point_dim) {}
  PARAM(PropertyBool, enable_full_distance_to_left, "enables complete detection from the projection to the left image border", false, &_config_changed);
  PARAM(PropertyBool, enable_full_distance_to_right, "enables complete detection from the projection to the right image border", false, &_config_changed);
  PARAM(PropertyBool, enable_seeding_when_tracking, "enables new point seeding when in tracking mode", true, &_config_changed);
  void init() override;  // selective.cpp:6-46
  size_t numberOfTrackingKeypoints() const { return _n_tracking; }
  // selective.cpp:63-144 (host-side, no device needed)
  void paintTrackingMask(int rows, int cols, const PointIntensityDescriptorCloud& projections, int projection_detection_radius,
                         std::vector<uint8_t>& mask) const;

protected:
  int extract(pslam_ctx* ctx, const ImageView& image, int capacity, float* xy, float* response, float* intensity,
              uint8_t* desc) override;  // selective.cpp:49-205
  std::vector<uint8_t> _tracking_mask;
  size_t _n_tracking = 0;
};

// ---- descriptor-based finders (registration/correspondence_finders/correspondence_finder_descriptor_based_*.h) --
// Request of the aligner loop to run its next solver iterations inside the finder's device round trip
// (pslam_projective_match_gn): filled by MultiAligner3DQRCUDA before finder->compute(), honoured by the projective finders
// when that call searches; every other finder (and every call that keeps its correspondences) leaves `executed` false and
// the aligner runs the iterations with its own launch.
struct FusedSolveRequest {
  const pslam_linearize_cfg* factor = nullptr;
  float diagonal_info[3] = {0, 0, 0};
  double damping = 0;
  double estimate[12] = {0};
  const pslam_pose_prior* prior = nullptr;
  int max_fused = 0;                  // iterations left in the aligner loop
  const float* moving_scale = nullptr;  // per moving point: 1 + log(n_opt) weighting of setupFactor (or 1)
  // results
  bool executed = false;
  bool spd = true;
  int done = 0;
  std::vector<double> poses, stats;
  std::vector<uint8_t> status;
  // alignFused only: why the device-resident registration handed back (pslam_align.stop_reason) and, per search phase,
  // (first solver iteration, iterations done, correspondences)
  int stop_reason = 0;
  std::vector<int> phase_log;
};

class CorrespondenceFinderBase : public Configurable {
public:
  virtual void setFusedSolve(FusedSolveRequest* r) { (void) r; }
  // The aligner's whole loop { compute(); solver iterations } from the current estimate, as far as this finder can take it on
  // the device (pslam_projective_align).  false: nothing done, the caller runs its loop.  true: r.done iterations were
  // executed (r.poses / r.stats / r.phase_log), the finder is in the state the call-by-call loop would have left it in, and
  // r.stop_reason tells whether the caller's loop has anything left to do.
  virtual bool alignFused(FusedSolveRequest& r, int min_num_correspondences) {
    (void) r;
    (void) min_num_correspondences;
    return false;
  }
  void setFixed(const PointIntensityDescriptorCloud* fixed) {
    _fixed = fixed;
    _fixed_changed_flag = true;
  }
  void setMoving(const PointIntensityDescriptorCloud* moving) {
    _moving = moving;
    _moving_changed_flag = true;
  }
  void setCorrespondences(CorrespondenceVector* c) { _correspondences = c; }
  void setLocalMapInSensor(const Isometry3f& T) { _local_map_in_sensor = T; }
  const Isometry3f& localMapInSensor() const { return _local_map_in_sensor; }
  virtual void compute() = 0;
  // number of upcoming compute() calls (estimate changes only) that are guaranteed NOT to change the correspondences;
  // the aligner fuses that many + 1 solver iterations into one device launch
  virtual int callsWithoutNewCorrespondences() const { return 0; }

protected:
  const PointIntensityDescriptorCloud* _fixed = nullptr;
  const PointIntensityDescriptorCloud* _moving = nullptr;
  CorrespondenceVector* _correspondences = nullptr;
  Isometry3f _local_map_in_sensor;
  bool _fixed_changed_flag = false, _moving_changed_flag = false;
};

class CorrespondenceFinderDescriptorBasedBruteforceCUDA : public CorrespondenceFinderBase {
public:
  PARAM(PropertyFloat, maximum_descriptor_distance, "maximum permitted descriptor distance for a match", 50.0f, nullptr);
  PARAM(PropertyFloat, maximum_distance_ratio_to_second_best, "Lowe's distance to drop ambiguous match candidates", 0.9f, nullptr);
  PARAM(PropertyFloat, minimum_matching_ratio, "desired minimum matching ratio with current configuration (signals transgressions)", 0.25f, nullptr);
  void compute() override;  // bruteforce_impl.cpp:6-155
  int callsWithoutNewCorrespondences() const override { return 1 << 30; }  // change flags only (:13-15)

protected:
  void _preCompute();   // :201-228
  void _postCompute();  // :230-243
};

class CorrespondenceFinderDescriptorBasedEpipolarCUDA : public CorrespondenceFinderDescriptorBasedBruteforceCUDA {
public:
  PARAM(PropertyUnsignedInt, maximum_disparity_pixels, "maximum disparity search range in pixels", 100, nullptr);
  PARAM(PropertyUnsignedInt, epipolar_line_thickness_pixels, "epipolar line search thickness in pixels (0 for perfect horizontal calibration)", 0, nullptr);
  void compute() override;  // epipolar_impl.cpp:44-219
  pslam_match_cfg cudaConfig() const;
};

// ---- projective finders (correspondence_finder_projective_base.h + square / circle / rhombus) -----------------
class CorrespondenceFinderProjectiveCUDA : public CorrespondenceFinderDescriptorBasedBruteforceCUDA {
public:
  explicit CorrespondenceFinderProjectiveCUDA(int shape = 1) : _shape(shape) {}
  PARAM(PropertyFloat, minimum_descriptor_distance, "minimum permitted descriptor distance for a match (initial)", 25.0f, &_config_changed);
  PARAM(PropertyFloat, descriptor_distance_step_size_pixels, "descriptor distance step size (increase/decrease)", 5, &_config_changed);
  PARAM(PropertyUnsignedInt, maximum_search_radius_pixels, "maximum projective region search radius in pixels", 100, &_config_changed);
  PARAM(PropertyUnsignedInt, minimum_search_radius_pixels, "minimum projective region search radius in pixels", 10, &_config_changed);
  PARAM(PropertyUnsignedInt, search_radius_step_size_pixels, "search radius step size (increase/decrease) in pixels", 5, &_config_changed);
  PARAM(PropertyConfigurable_<ProjectorPinhole>, projector, "pinhole projector used for projective descriptor matching",
        std::make_shared<ProjectorPinhole>(), &_config_changed);
  PARAM(PropertyUnsignedInt, minimum_number_of_iterations, "minimum number of iterations to perform recomputes (forced)", 10, &_config_changed);
  PARAM(PropertyFloat, maximum_estimate_change_norm_for_convergence, "maximum allowed transform estimate change threshold for assuming convergence", 1e-5f, &_config_changed);
  PARAM(PropertyUnsignedInt, number_of_solver_iterations_per_projection, "minimum number of solver iterations (on estimate) to await before reprojecting points", 25, &_config_changed);

  void compute() override;  // projective_base_impl.cpp:104-293 (state machine); search + filter on the device
  void setSearchradiusPixels(size_t r) {  // projective_base.h:82-85
    _search_radius_pixels = r;
    _config_changed = false;
  }
  size_t searchRadiusPixels() const { return _search_radius_pixels; }
  void setDescriptorDistance(float d) {  // :94-97
    _descriptor_distance = d;
    _config_changed = false;
  }
  float descriptorDistance() const { return _descriptor_distance; }
  bool hasConverged() const { return _has_converged; }
  size_t currentIteration() const { return _current_iteration; }
  int numberOfSearches() const { return _number_of_searches; }
  int callsWithoutNewCorrespondences() const override;
  void setFusedSolve(FusedSolveRequest* r) override { _fused = r; }
  bool alignFused(FusedSolveRequest& r, int min_num_correspondences) override;
  int shape() const { return _shape; }

private:
  void _prepare();                                            // :104-160: checks, device uploads, state reset on new clouds
  pslam_projective_cfg _deviceConfig() const;                 // projector + current thresholds
  void _uploadWeights(pslam_ctx* ctx, const float* moving_scale);
  void _adaptAfterConvergence();                              // :272-283
  int _shape;
  bool _config_changed = true;
  size_t _search_radius_pixels = 0;
  float _descriptor_distance = 0;
  Isometry3f _local_map_in_sensor_previous;
  bool _has_converged = false;
  size_t _current_iteration = 0;
  int _number_of_searches = 0;
  unsigned long long _device_fixed_epoch = 0, _device_moving_epoch = 0;  // stamps of OUR uploads into the shared device cache
  FusedSolveRequest* _fused = nullptr;
  const float* _device_weights_of = nullptr;        // scale table uploaded for ...
  unsigned long long _device_weights_epoch = 0;     // ... this moving-cloud epoch
};

// ---- measurement adaptors (sensor_processing/raw_data_preprocessor_{stereo_projective,monocular_depth}.{h,cpp}) --
class RawDataPreprocessorBase : public Configurable {
public:
  enum Status { Error = 0, Initializing = 1, Ready = 2 };
  Status status() const { return _status; }
  void setMeas(PointIntensityDescriptorCloud* meas) { _meas = meas; }
  // raw_data_preprocessor_descriptor_based.hpp:27-30: projections of the tracked points, handed to the extractor
  void setProjections(const PointIntensityDescriptorCloud* projections, size_t projection_radius) {
    _projections = projections;
    _projection_radius = projection_radius;
  }

protected:
  const PointIntensityDescriptorCloud* _projections = nullptr;
  size_t _projection_radius = 0;
  Status _status = Error;
  PointIntensityDescriptorCloud* _meas = nullptr;
  bool _raw_data_changed_flag = false;
};

class RawDataPreprocessorStereoProjectiveCUDA : public RawDataPreprocessorBase {
public:
  RawDataPreprocessorStereoProjectiveCUDA();
  PARAM(PropertyConfigurable_<IntensityFeatureExtractorBaseCUDA>, feature_extractor,
        "feature extractor used to detect keypoints and compute descriptors", std::static_pointer_cast<IntensityFeatureExtractorBaseCUDA>(std::make_shared<IntensityFeatureExtractorBinnedCUDA>(3)), nullptr);
  PARAM(PropertyConfigurable_<IntensityFeatureExtractorBaseCUDA>, feature_extractor_right,
        "feature extractor used to detect keypoints and compute descriptors in the right frame", std::static_pointer_cast<IntensityFeatureExtractorBaseCUDA>(std::make_shared<IntensityFeatureExtractorBinnedCUDA>(3)), nullptr);
  PARAM(PropertyConfigurable_<CorrespondenceFinderDescriptorBasedBruteforceCUDA>, correspondence_finder,
        "descriptor-based correspondence finder used to compute stereo matches",
        std::static_pointer_cast<CorrespondenceFinderDescriptorBasedBruteforceCUDA>(std::make_shared<CorrespondenceFinderDescriptorBasedEpipolarCUDA>()), nullptr);
  PARAM(PropertyString, topic_camera_left, "left rgb image topic [/camera_left/image_raw]", "/camera_left/image_raw", nullptr);
  PARAM(PropertyString, topic_camera_right, "right rgb image topic [/camera_right/image_raw]", "/camera_right/image_raw", nullptr);
  // setRawData (stereo_projective.cpp:6-44): a message pack of exactly two images
  bool setRawData(const ImageView& left, const ImageView& right);
  void compute();  // :46-134

private:
  ImageView _left, _right;
  bool _raw_set = false;
};

class RawDataPreprocessorMonocularDepthCUDA : public RawDataPreprocessorBase {
public:
  PARAM(PropertyConfigurable_<IntensityFeatureExtractorBaseCUDA>, feature_extractor,
        "feature extractor used to detect keypoints and compute descriptors", std::static_pointer_cast<IntensityFeatureExtractorBaseCUDA>(std::make_shared<IntensityFeatureExtractorBinnedCUDA>(3)), nullptr);
  PARAM(PropertyString, topic_rgb, "rgb image topic [/camera/rgb/image]", "/camera/rgb/image", nullptr);
  PARAM(PropertyString, topic_depth, "topic depth image [/camera/depth/image]", "/camera/depth/image", nullptr);
  PARAM(PropertyFloat, depth_scaling_factor_to_meters, "scaling factor used to obtain depth in meters from pixel values", 1.0f, nullptr);
  bool setRawData(const ImageView& intensity, const DepthView& depth);
  void compute();  // monocular_depth.cpp:50-180

private:
  ImageView _intensity;
  DepthView _depth;
  bool _raw_set = false;
};

// ---- aligner slice processors (registration/aligner_slice_processor_projective.{h,cpp}) -----------------------
class AlignerSliceProcessorProjectiveCUDA : public Configurable {
public:
  // factor kind: 0 rectified stereo (fixed 4-D), 1 projective depth (fixed 3-D), 2 projective (fixed 2-D)
  explicit AlignerSliceProcessorProjectiveCUDA(int kind = 2);
  PARAM(PropertyVector_<float>, diagonal_info_matrix, "value of the information matrix's diagonal", std::vector<float>(), nullptr);
  PARAM(PropertyConfigurable_<ProjectorPinhole>, projector, "link to a projector where to take the infor for the factor", nullptr, nullptr);
  PARAM(PropertyConfigurable_<CorrespondenceFinderBase>, finder, "correspondence finder used in this cue", nullptr, nullptr);
  PARAM(PropertyConfigurable_<RobustifierBase>, robustifier, "robustifier used on this slice", nullptr, nullptr);
  PARAM(PropertyString, fixed_slice_name, "name of the slice in the fixed scene", "points", nullptr);
  PARAM(PropertyString, moving_slice_name, "name of the slice in the moving scene", "points", nullptr);
  PARAM(PropertyInt, min_num_correspondences, "minimum number of correspondences in this slice", 0, nullptr);
  // stereo only (aligner_slice_processor_projective.h:129-151)
  PARAM(PropertyString, frame_camera_left, "topic for the camera left info", "camera_left", nullptr);
  PARAM(PropertyString, frame_camera_right, "topic for the camera right info", "camera_right", nullptr);
  PARAM(PropertyBool, enable_inverse_depth_weighting, "toggles point weighting by inverse depth", false, nullptr);
  PARAM(PropertyBool, enable_point_covariance_integration, "toggles individual point covariance integration", false, nullptr);

  int kind() const { return _kind; }
  int fixedDim() const { return _kind == 0 ? 4 : (_kind == 1 ? 3 : 2); }
  // platform->getTransform(left_camera_in_right, frame_camera_left, frame_camera_right) (.cpp:95-103)
  void setLeftCameraInRight(const float t_left_in_right[3]) {
    for (int i = 0; i < 3; ++i) _t_left_in_right[i] = t_left_in_right[i];
    _baseline_set = false;
  }
  void setFixed(const PointIntensityDescriptorCloud* f) { _fixed_slice = f; }
  void setMoving(const PointIntensityDescriptorCloud* m) { _moving_slice = m; }
  void bindFixed();    // stereo: mean disparity (.cpp:75-89)
  void setupFactor();  // K, image dim, per-correspondence information (.cpp:27-73), stereo extras (:91-112)
  void setupFactorConfig();  // the part of setupFactor that does not depend on the correspondences
  void diagonalInfo(float d[3]) const;
  CorrespondenceVector& correspondences() { return _correspondences; }
  const pslam_linearize_cfg& factorConfig() const { return _factor; }
  const std::vector<float>& informationDiagonals() const { return _fixed_information_diagonals; }
  float meanDisparity() const { return _mean_disparity; }
  const PointIntensityDescriptorCloud* fixedSlice() const { return _fixed_slice; }
  const PointIntensityDescriptorCloud* movingSlice() const { return _moving_slice; }

private:
  int _kind;
  const PointIntensityDescriptorCloud* _fixed_slice = nullptr;
  const PointIntensityDescriptorCloud* _moving_slice = nullptr;
  CorrespondenceVector _correspondences;
  std::vector<float> _fixed_information_diagonals;  // 3 per fixed point (fp32 like the reference's InformationMatrixVector)
  pslam_linearize_cfg _factor{};
  float _mean_disparity = 0;
  float _t_left_in_right[3] = {0, 0, 0};
  float _baseline_left_in_right_pixelsmeters[3] = {0, 0, 0};
  bool _baseline_set = false;
};

// ---- the shipped aligners' second slice (srrg2_slam_interfaces, external): AlignerSliceMotionModel3D over the
//      "trajectory_chunk" slice with a MotionModelConstantVelocity3D (configurations/kitti.conf:747-772,257-260;
//      icl.conf:268-293,660-663; euroc.conf:94-119,456-459).  [upstream, restated]: the motion model predicts the
//      estimate from the last relative motion of the chunk; the slice seeds the aligner's estimate with the prediction
//      (what tests/test_aligners.cpp:1083,1096-1102 requires) and contributes an SE3 pose-prior factor with a constant
//      information matrix to every iteration's H, b -- evaluated on the device inside the fused launch.
class MotionModelConstantVelocity3D : public Configurable {
public:
  // robot poses in the local map, oldest first -> predicted moving_in_fixed (= predicted robot_in_local_map^-1)
  Isometry3f predict(const std::vector<Isometry3f>& trajectory_chunk) const;
};
class AlignerSliceMotionModel3DCUDA : public Configurable {
public:
  AlignerSliceMotionModel3DCUDA();
  PARAM(PropertyString, base_frame_id, "name of the base frame in the tf tree", "", nullptr);
  PARAM(PropertyString, frame_id, "name of the sensor's frame in the tf tree", "", nullptr);
  PARAM(PropertyString, fixed_slice_name, "name of the slice in the fixed scene", "trajectory_chunk", nullptr);
  PARAM(PropertyString, moving_slice_name, "name of the slice in the moving scene", "trajectory_chunk", nullptr);
  PARAM(PropertyConfigurable_<MotionModelConstantVelocity3D>, motion_model, "model used to estimate inter-frame motion",
        std::make_shared<MotionModelConstantVelocity3D>(), nullptr);
  PARAM(PropertyConfigurable_<RobustifierBase>, robustifier, "robustifier used on this slice", nullptr, nullptr);
  void setTrajectoryChunk(const std::vector<Isometry3f>& chunk) { _chunk = chunk; }
  void setInformationMatrix(const double information36[36]) {
    for (int i = 0; i < 36; ++i) _information[i] = information36[i];
  }
  pslam_pose_prior prior() const;  // prediction of the motion model + the information matrix

private:
  std::vector<Isometry3f> _chunk;
  double _information[36];
};

// ---- MultiAligner3DQR (srrg2_slam_interfaces, external; iteration structure: SURVEY App. E.6) ------------------
struct AlignerIterationStats {
  int iteration = 0, num_correspondences = 0, num_inliers = 0, num_outliers = 0, num_suppressed = 0;
  double chi = 0;
};
class MultiAligner3DQRCUDA : public Configurable {
public:
  enum Status { Fail = 0, Success = 1, NotEnoughCorrespondences = 2, NotEnoughInliers = 3 };
  PARAM(PropertyBool, enable_inlier_only_runs, "toggles additional inlier only runs if sufficient inliers are available", false, nullptr);
  PARAM(PropertyBool, keep_only_inlier_correspondences, "toggles removal of correspondences which factors are not inliers in the last iteration", false, nullptr);
  PARAM(PropertyInt, max_iterations, "maximum number of iterations", 10, nullptr);
  PARAM(PropertyInt, min_num_inliers, "minimum number ofinliers", 10, nullptr);
  PARAM_VECTOR(PropertyConfigurableVector_<Configurable>, slice_processors, "slices", nullptr);
  PARAM(PropertyConfigurable_<Solver>, solver, "this solver", std::make_shared<Solver>(), nullptr);
  PARAM(PropertyConfigurable_<Configurable>, termination_criteria, "termination criteria, not set=max iterations", nullptr, nullptr);

  void setFixed(const PointIntensityDescriptorCloud* f) { _fixed = f; }
  void setMoving(const PointIntensityDescriptorCloud* m) { _moving = m; }
  void setMovingInFixed(const Isometry3f& T);
  const std::array<double, 12>& movingInFixed() const { return _estimate; }
  Status status() const { return _status; }
  void compute();
  const std::vector<AlignerIterationStats>& iterationStats() const { return _stats; }
  // iterations of the additional inlier-only run (enable_inlier_only_runs, configurations/icl.conf:55)
  const std::vector<AlignerIterationStats>& inlierRunStats() const { return _inlier_run_stats; }
  AlignerSliceProcessorProjectiveCUDA* projectiveSlice() const;
  AlignerSliceMotionModel3DCUDA* motionModelSlice() const;

private:
  const PointIntensityDescriptorCloud* _fixed = nullptr;
  const PointIntensityDescriptorCloud* _moving = nullptr;
  std::array<double, 12> _estimate{{1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0}};
  Status _status = Fail;
  std::vector<AlignerIterationStats> _stats, _inlier_run_stats;
};

// ---- SceneClipperProjective3D (mapping/scene_clipper_projective_3d.{h,cpp}) -----------------------
// srrg2_slam_interfaces::SceneClipper_ contract: setFullScene / setClippedSceneInRobot / setRobotInLocalMap /
// setSensorInRobot, compute(), status(), globalIndices().  The projection + ordered compaction of the whole local
// map runs in scene_clip_kernel (pslam_scene_clip).
class SceneClipperProjective3DCUDA : public Configurable {
public:
  enum Status { Error = 0, Ready = 1, Successful = 2 };
  PARAM(PropertyConfigurable_<ProjectorPinhole>, projector,
        "pinhole projector used to determine whether points lie in the current view", std::make_shared<ProjectorPinhole>(), nullptr);
  void setFullScene(const PointIntensityDescriptorCloud* scene) { _full_scene = scene; }
  void setClippedSceneInRobot(PointIntensityDescriptorCloud* clipped) { _clipped_scene_in_robot = clipped; }
  void setRobotInLocalMap(const Isometry3f& T) { _robot_in_local_map = T; }
  void setSensorInRobot(const Isometry3f& T) { _sensor_in_robot = T; }
  void compute();  // scene_clipper_projective_3d.cpp:9-67
  Status status() const { return _status; }
  const std::vector<int>& globalIndices() const { return _global_indices; }
  // (u, v, depth) of the survivors: the projector's third output, "can be reused for e.g. alignment" (:48-49)
  const std::vector<float>& projections() const { return _projections; }

private:
  const PointIntensityDescriptorCloud* _full_scene = nullptr;
  PointIntensityDescriptorCloud* _clipped_scene_in_robot = nullptr;
  Isometry3f _robot_in_local_map, _sensor_in_robot;
  std::vector<int> _global_indices;
  std::vector<float> _projections;
  Status _status = Error;
};

// ---- point EKFs + LandmarkEstimatorEKF_ (mapping/landmarks/filters/*.h, landmark_estimator_ekf.{h,cpp}) --
// The filter object only carries the camera model (setCameraMatrix / setBaseline); the arithmetic of
// PointEKFBase::compute runs in landmarks_ekf_kernel.
class PointEKFCUDA : public Configurable {
public:
  explicit PointEKFCUDA(int kind) : _kind(kind) {}
  int kind() const { return _kind; }  // 0 ProjectivePointEKF, 1 ProjectiveDepthPointEKF, 2 StereoProjectivePointEKF
  void setCameraMatrix(const std::array<float, 9>& K) { _K = K; }
  void setBaseline(double b_x, double b_y) {  // StereoProjectivePointEKF::setBaseline, pixels
    _b[0] = b_x;
    _b[1] = b_y;
  }
  const std::array<float, 9>& cameraMatrix() const { return _K; }
  const double* baseline() const { return _b; }

private:
  int _kind;
  std::array<float, 9> _K{{1, 0, 0, 0, 1, 0, 0, 0, 1}};
  double _b[2] = {0, 0};
};

class LandmarkEstimatorEKFCUDA : public Configurable {
public:
  explicit LandmarkEstimatorEKFCUDA(int kind) : _kind(kind) {}
  PARAM(PropertyConfigurable_<PointEKFCUDA>, filter, "filter instance used to refine the landmark position estimate", nullptr, nullptr);
  PARAM(PropertyDouble, minimum_state_element_covariance, "minimum per element covariance value (prohibits ill-shaped uncertainties)", 0.01, nullptr);
  PARAM(PropertyDouble, maximum_covariance_norm_squared, "maximum permitted covariance matrix Frobenius norm value for merging", 1, nullptr);
  PARAM(PropertyFloat, maximum_distance_geometry_meters_squared,
        "maximum distance in geometry (i.e. 3D point distance L2 norm) in meters squared", 1, nullptr);
  // LandmarkEstimatorBase_::setTransforms (landmark_estimator_base.hpp:49-58)
  void setTransforms(const Isometry3f& measurement_in_world, const Isometry3f& measurement_in_scene) {
    _sensor_in_world = measurement_in_world;
    _sensor_in_local_map = measurement_in_scene;
  }
  // setMeasurement / setLandmark / compute (landmark_estimator_ekf_impl.cpp:17-82) for ALL correspondences of one merger
  // pass: state_world [n][3] / covariance [n][9] = landmark statistics (updated in place where isInlier), measurements
  // [n][MeasurementDim], coords_in_local_map [n][3], inlier [n].  Returns the number of inliers.
  int computeBatch(int n, float* state_world, float* covariance, const float* measurements, float* coords_in_local_map,
                   uint8_t* inlier);
  int measurementDim() const { return _kind == 0 ? 2 : (_kind == 1 ? 3 : 4); }

private:
  int _kind;
  Isometry3f _sensor_in_world, _sensor_in_local_map;
};

// ---- LandmarkEstimatorWeightedMean_ (mapping/landmarks/landmark_estimator_weighted_mean.{h,cpp}) -------------------
class LandmarkEstimatorWeightedMeanCUDA : public Configurable {
public:
  PARAM(PropertyFloat, maximum_distance_geometry_meters_squared,
        "maximum distance in geometry (i.e. 3D point distance L2 norm) in meters squared", 1, nullptr);
  void setTransforms(const Isometry3f& measurement_in_world, const Isometry3f& measurement_in_scene) {
    _sensor_in_world = measurement_in_world;
    _sensor_in_local_map = measurement_in_scene;
  }
  // setLandmarkInSensor / setLandmark / compute (landmark_estimator_weighted_mean_impl.cpp:7-41) for all correspondences
  // of one merger pass; number_of_optimizations = statistics().numberOfOptimizations() of every landmark
  int computeBatch(int n, float* state_world, const int* number_of_optimizations, const float* landmark_in_sensor,
                   float* coords_in_local_map, uint8_t* inlier);

private:
  Isometry3f _sensor_in_world, _sensor_in_local_map;
};

// ---- MergerProjective_ family (mapping/mergers/merger_projective.h:14-110, merger_projective_rigid_stereo*.h,
// merger_projective_depth_ekf.h).  The module carries the .conf parameters and runs the BINNING of
// compute() / _addPoints() on the device: which correspondences reach _updatePoint, which measurements become addition
// candidates.  The landmark update itself is the linked landmark_estimator's computeBatch, the triangulation of the
// additions pslam_triangulate -- the scene container and the statistics field stay with the caller (control plane).
class MergerProjectiveCUDA : public Configurable {
public:
  explicit MergerProjectiveCUDA(int kind) : _kind(kind) {}
  // MergerCorrespondence_ (srrg2_slam_interfaces)
  PARAM(PropertyBool, enable_binning, "toggles point binning (distribution homogenization)", true, nullptr);
  PARAM(PropertyUnsignedInt, target_number_of_merges,
        "target number of points to merge, if hit no further points without correspondences will be added to moving", 100, nullptr);
  // merger_projective.h:31-66
  PARAM(PropertyConfigurable_<Configurable>, landmark_estimator,
        "landmark estimator used to refine landmark positions in the map (structure-only)", nullptr, nullptr);
  PARAM(PropertyConfigurable_<ProjectorPinhole>, projector, "pinhole projector used for projective merging",
        std::make_shared<ProjectorPinhole>(), nullptr);
  PARAM(PropertyFloat, maximum_distance_appearance, "maximum permitted correspondence response for merging a point", 50, nullptr);
  PARAM(PropertyUnsignedInt, number_of_row_bins, "number of bins in row direction (feature density regulation)", 10, nullptr);
  PARAM(PropertyUnsignedInt, number_of_col_bins, "number of bins in column direction (feature density regulation)", 30, nullptr);
  PARAM(PropertyFloat, target_merge_ratio, "target merge ratio (#merges/#correspondences)", 0.5f, nullptr);
  PARAM(PropertyBool, enable_conservative_addition,
        "enables minimal addition of new points (instead all) if merge ratio is not reached", false, nullptr);
  // merger_projective_rigid_stereo.h:24-28 / merger_projective_depth_ekf.h:27-31 (srrg2 modules outside the path: kept as links)
  PARAM(PropertyConfigurable_<Configurable>, triangulator, "rigid stereo triangulation unit", nullptr, nullptr);
  PARAM(PropertyConfigurable_<Configurable>, unprojector, "un-projector used to compute the points from the depth image", nullptr, nullptr);

  int kind() const { return _kind; }  // PSLAM_MERGER_STEREO / PSLAM_MERGER_DEPTH
  // update pass of compute() (merger_projective_impl.cpp:61-135): selected[c] = _updatePoint is reached; the blocked bins
  // are kept for the addition pass.  measurements [n][dim], dim 4 (stereo) or 3 (u, v, depth).  Returns #selected.
  int selectUpdates(const float* measurements, int dim, int n_meas, const int* corr_moving, const float* corr_response, int n_corr,
                    uint8_t* selected);
  // whether compute() goes on to _addPoints with `number_of_merged_points` successful updates (:158-165; always when
  // there were no correspondences, :56-58)
  bool wantsAdditions(int number_of_merged_points, int n_meas, int n_corr) const;
  // binning of _addPoints (:205-253) against the bins blocked by the last selectUpdates; winners has room for n_meas
  int selectAdditions(const float* measurements, int dim, int n_meas, int* winners);
  // both passes in one device round trip (the addition candidates do not depend on the estimator's verdicts, only on the
  // bins the update pass blocks); returns #selected, *n_winners = #addition candidates
  int plan(const float* measurements, int dim, int n_meas, const int* corr_moving, const float* corr_response, int n_corr,
           uint8_t* selected, int* winners, int* n_winners);
  void resetOccupancy() { _occupied.clear(); }

private:
  pslam_merger_cfg _cfg() const;
  int _kind;
  std::vector<uint32_t> _occupied;
};

// ---- LandmarkEstimatorPoseBasedSmoother_ (mapping/landmarks/landmark_estimator_pose_based_smoother.{h,cpp}) ----------
class LandmarkEstimatorPoseBasedSmootherCUDA : public Configurable {
public:
  PARAM(PropertyUnsignedInt, maximum_number_of_iterations, "maximum number of LS iterations", 100, nullptr);
  PARAM(PropertyFloat, convergence_criterion_minimum_chi2_delta, "convergence delta", 1e-5f, nullptr);
  PARAM(PropertyFloat, maximum_reprojection_error_pixels_squared, "maximum kernel reprojection error (pixels squared)", 100, nullptr);
  PARAM(PropertyUnsignedInt, minimum_number_of_measurements_for_optimization,
        "minimum number of required measurements before optimizing (otherwise averaging)", 3, nullptr);
  PARAM(PropertyFloat, maximum_distance_geometry_meters_squared,
        "maximum distance in geometry (i.e. 3D point distance L2 norm) in meters squared", 1, nullptr);
  void setCameraMatrix(const std::array<float, 9>& K) { _K = K; }
  void setTransforms(const Isometry3f& measurement_in_world, const Isometry3f& measurement_in_scene) {
    _sensor_in_world = measurement_in_world;
    _sensor_in_local_map = measurement_in_scene;
  }
  // compute() for all correspondences of one merger pass; the measurement histories (current measurement included) are
  // handed over in CSR form, see pslam_landmarks_smoother_update
  int computeBatch(int n, float* state_world, int* number_of_optimizations, int n_frames, const float* frames_sensor_in_world,
                   const int* offsets, const int* hist_frame, const float* hist_uv, const float* hist_point_in_camera,
                   float* coords_in_local_map, uint8_t* inlier);

private:
  std::array<float, 9> _K{{0, 0, 0, 0, 0, 0, 0, 0, 0}};
  Isometry3f _sensor_in_world, _sensor_in_local_map;
};

// registers every class above under the reference's names and under the ...CUDA names (idempotent)
void registerTypes();

}  // namespace pslam_host
