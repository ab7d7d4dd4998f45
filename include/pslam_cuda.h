/* =============================================================================
 * pslam_cuda.h -- C ABI of the B200-native (sm_100a) srrg2_proslam frontend.
 *
 * This is the drop-in boundary: plain pointers and sizes, no
 * torch / Eigen / OpenCV types.  Each entry point names the reference interface it
 * replaces (paths relative to the reference root; `.../` abbreviates
 * `srrg2_proslam/src/srrg2_proslam/`).  The host-side mirror of the srrg2 plugin
 * classes (srrg2_proslam_b200/host/) and the Python test harness both call ONLY these
 * functions.  There is no CPU fallback: every compute call needs a CUDA device and
 * returns PSLAM_E_CUDA when none is usable.
 *
 * Conventions
 *   - return value: 0 (PSLAM_OK) or a negative PSLAM_E_* code; functions that produce
 *     a variable-length result return the element count (>= 0) instead.
 *   - `*_dev` variants take DEVICE pointers (inputs already resident in HBM) and leave
 *     their results on the device inside the context; the plain variants take HOST
 *     pointers and do the host<->device copies themselves (these are the calls the
 *     reference-side host code makes).
 *   - descriptors are 256 bit, stored as 32 bytes (bit k of byte i = ORB pair 8i+k),
 *     i.e. the row layout of the cv::Mat the reference keeps per point.
 *   - poses are row-major 3x4 [R|t].  K is row-major 3x3.
 * ========================================================================== */
#ifndef PSLAM_CUDA_H
#define PSLAM_CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PSLAM_OK 0
#define PSLAM_E_INVALID (-1)  /* bad argument / configuration (reference: throw std::runtime_error) */
#define PSLAM_E_CUDA (-2)     /* CUDA runtime error, see pslam_last_error() */
#define PSLAM_E_CAPACITY (-3) /* a context limit (pslam_limits) was exceeded; nothing was truncated silently */
#define PSLAM_E_NOT_SPD (-4)  /* linear system not positive definite */

typedef struct pslam_ctx pslam_ctx;

/* Context limits: size the device-resident buffers once (no allocation on the hot path). */
typedef struct {
  int max_images;        /* images per batch (a stereo pair is 2 images): capacity of the result stores */
  int max_rows;          /* image height limit */
  int max_cols;          /* image width limit */
  int max_features;      /* selected features per image, <= 8192 */
  int max_raw_per_bin;   /* FAST corners (after NMS) per detection region before selection */
  int max_bins;          /* detection regions per image (nh * nv) */
  int work_images;       /* images processed per pipeline chunk (intermediate maps are sized by this,
                            the feature / stereo stores by max_images); 0 = min(max_images, 512) */
} pslam_limits;

/* PARAMs of IntensityFeatureExtractorBinned_ (.../sensor_processing/feature_extractors/
 * intensity_feature_extractor_base.h:24-58, intensity_feature_extractor_binned.h:17-27).
 * detector_type is FAST and descriptor_type ORB-256 (the only pair the shipped configs use). */
typedef struct {
  float detector_threshold;
  int enable_non_maximum_suppression;
  int target_number_of_keypoints;
  int number_of_detectors_horizontal;
  int number_of_detectors_vertical;
} pslam_extract_cfg;

/* PARAMs of CorrespondenceFinderDescriptorBased{Bruteforce,Epipolar}
 * (.../registration/correspondence_finders/correspondence_finder_descriptor_based_bruteforce.h:23-37,
 *  correspondence_finder_descriptor_based_epipolar.h:24-34). */
typedef struct {
  float maximum_descriptor_distance;
  float maximum_distance_ratio_to_second_best;
  int maximum_disparity_pixels;
  int epipolar_line_thickness_pixels;
} pslam_match_cfg;

/* ---- lifecycle ----------------------------------------------------------------- */
int pslam_create(int device, const pslam_limits* limits, pslam_ctx** out);
void pslam_destroy(pslam_ctx* ctx);
const char* pslam_last_error(const pslam_ctx* ctx);
const char* pslam_version(void);
/* number of kernels this context has launched since creation (bench.py's gpu_launches) */
long long pslam_launch_count(const pslam_ctx* ctx);
/* the context's main CUDA stream (cudaStream_t as void*): every entry point is ordered on it -- batched stage-1 calls
 * fan their chunks out over two internal streams ("lanes", see pslam_set_lanes) and join them back before they return */
void* pslam_stream(const pslam_ctx* ctx);
int pslam_synchronize(pslam_ctx* ctx);
/* Chunk lanes of the batched stage-1 entry points (pslam_extract_binned_batch_dev, pslam_stereo_frontend_batch[_dev]):
 * with 2 lanes (default) chunk i + 1's detection kernel overlaps chunk i's selection / description / matching kernels on a
 * second stream with its own chunk-level intermediates (allocated by the first batch that spans more than one chunk);
 * with 1 lane the kernels of a batch run strictly one after the other (per-kernel timing, debugging).  Results are
 * identical.  The environment variable PSLAM_LANES sets the initial value. */
int pslam_set_lanes(pslam_ctx* ctx, int n_lanes);

/* Per-kernel device timing for bench.py's roofline line (the reference's counterpart is srrg2_core's
 * Profiler / PROFILE_TIME scopes, e.g. .../feature_extractors/intensity_feature_extractor_binned.cpp:119,140,165).
 * While enabled, one CUDA event is recorded after every kernel launch; pslam_profile_mark() drops a
 * marker (call it after the host has queued copies, so that they are not billed to a kernel);
 * pslam_profile_read() synchronises and returns, per kernel name, the summed device time and launch
 * count since the last read.  names: capacity x name_len chars.  Returns the number of distinct kernels. */
int pslam_profile_enable(pslam_ctx* ctx, int enable);
int pslam_profile_mark(pslam_ctx* ctx);
int pslam_profile_read(pslam_ctx* ctx, int capacity, char* names, int name_len, double* total_ms,
                       long long* launches);

/* ---- stage 1: detect + select + describe ---------------------------------------
 * Replaces IntensityFeatureExtractorBinned_::compute(cv::Mat)
 *   (.../feature_extractors/intensity_feature_extractor_base.cpp:55-85 calling
 *    intensity_feature_extractor_binned.cpp:115-208 and base.cpp:45-53).
 * Output order, coordinates, response, intensity and descriptor bytes are those of the
 * reference; `mask` (rows x cols, non-zero = region of interest) mirrors
 * setKeypointDetectionMask (base.h:128-132) and, as in the reference, disables binning.
 * xy: 2 floats per feature; desc: 32 bytes per feature.  Returns the feature count. */
int pslam_extract_binned(pslam_ctx* ctx, const uint8_t* image, int rows, int cols, int stride,
                         const pslam_extract_cfg* cfg, const uint8_t* mask_or_null, int capacity,
                         float* xy, float* response, float* intensity, uint8_t* desc);

/* Replaces IntensityFeatureExtractorSelective_::compute in tracking mode
 *   (.../feature_extractors/intensity_feature_extractor_selective.cpp:49-205 + base.cpp:45-85):
 * FAST over the whole image once; keypoints inside `tracking_mask` (rows x cols, non-zero = the rectangles the
 * reference paints around the projections, :80-144) come first in row-major order, then -- if enable_seeding
 * (PARAM enable_seeding_when_tracking, :166-174) -- the keypoints of the complement, also row-major.  No binning,
 * no quota (the selective extractor keeps every detection).  *n_tracking receives the size of the first group.
 * Seeding mode without projections (:179-198) is pslam_extract_binned with mask_or_null. */
int pslam_extract_selective(pslam_ctx* ctx, const uint8_t* image, int rows, int cols, int stride,
                            const pslam_extract_cfg* cfg, const uint8_t* tracking_mask, int enable_seeding,
                            int capacity, float* xy, float* response, float* intensity, uint8_t* desc,
                            int* n_tracking);

/* Batched, device-resident variant: `n_images` images of identical size, image i at
 * d_images + i * image_pitch_bytes.  Results stay in the context's feature store
 * (slot i <- image i) for the matchers; read them back with pslam_download_features. */
int pslam_extract_binned_batch_dev(pslam_ctx* ctx, const uint8_t* d_images, int n_images, int rows,
                                   int cols, int stride, long long image_pitch_bytes,
                                   const pslam_extract_cfg* cfg);
int pslam_download_features(pslam_ctx* ctx, int slot, int capacity, float* xy, float* response,
                            float* intensity, uint8_t* desc);
/* counts of all slots of the last batch (n ints) */
int pslam_download_feature_counts(pslam_ctx* ctx, int n_images, int* counts);

/* Intermediate products, exposed for stage-level parity tests:
 * FAST+NMS keypoints in row-major order (cv::FastFeatureDetector::detect, called at
 * intensity_feature_extractor_binned.cpp:141-146) and the ORB blur (cv::ORB::compute's
 * 7x7 Gaussian, base.cpp:52). */
int pslam_fast_detect(pslam_ctx* ctx, const uint8_t* image, int rows, int cols, int stride,
                      int threshold, int nms, int capacity, float* xy, float* response);
int pslam_blur7(pslam_ctx* ctx, const uint8_t* image, int rows, int cols, int stride, uint8_t* out);

/* ---- stage 2a: stereo epipolar matching -----------------------------------------
 * Replaces CorrespondenceFinderDescriptorBasedEpipolar::compute
 *   (.../correspondence_finders/correspondence_finder_descriptor_based_epipolar_impl.cpp:44-219).
 * fixed = left, moving = right.  Output order = the reference's scan order. */
int pslam_match_epipolar(pslam_ctx* ctx, int n_fixed, const float* xy_fixed, const uint8_t* desc_fixed,
                         int n_moving, const float* xy_moving, const uint8_t* desc_moving,
                         const pslam_match_cfg* cfg, int capacity, int* fixed_idx, int* moving_idx,
                         float* distance);

/* ---- stage 1+2a fused per stereo pair: the measurement adaptor ------------------
 * Replaces RawDataPreprocessorStereoProjective::compute
 *   (.../sensor_processing/raw_data_preprocessor_stereo_projective.cpp:46-134):
 * extract left and right, epipolar match, emit (uL,vL,uR,vR) + left descriptor/intensity,
 * dropping negative disparities (:120-128).  Host buffers in and out. */
int pslam_stereo_adaptor(pslam_ctx* ctx, const uint8_t* left, const uint8_t* right, int rows, int cols,
                         int stride, const pslam_extract_cfg* ecfg, const pslam_match_cfg* mcfg,
                         int capacity, float* uvuv, float* intensity, uint8_t* desc);

/* ---- stage 1 + depth lookup: the RGB-D measurement adaptor ------------------------
 * Replaces RawDataPreprocessorMonocularDepth::compute / _readDepth
 *   (.../sensor_processing/raw_data_preprocessor_monocular_depth.cpp:50-180):
 * extract features, read d = depth(rint(v), rint(u)), keep the point iff d > 0 with
 * z = depth_scaling_factor_to_meters * d, preserving the cloud order (:166-179).
 * depth_type: 0 = uint16 (TYPE_16UC1), 1 = float (TYPE_32FC1); stride in ELEMENTS.
 * uvz: 3 floats per point.  *n_features_in_image receives the feature count before the depth
 * filter (the reference warns when > 25 % are lost, :147-152).  Returns the point count. */
int pslam_mono_depth_adaptor(pslam_ctx* ctx, const uint8_t* image, int rows, int cols, int stride,
                             const void* depth, int depth_type, int depth_rows, int depth_cols,
                             int depth_stride_elements, float depth_scaling_factor_to_meters,
                             const pslam_extract_cfg* ecfg, int capacity, float* uvz, float* intensity,
                             uint8_t* desc, int* n_features_in_image);

/* Batched, device-resident: images 2p (left) and 2p+1 (right) of pair p.  Runs the whole
 * frame-independent part of the frontend for n_pairs pairs; results stay on the device
 * (stereo slots p) and can be fetched per pair.  Returns PSLAM_OK. */
int pslam_stereo_frontend_batch_dev(pslam_ctx* ctx, const uint8_t* d_images, int n_pairs, int rows,
                                    int cols, int stride, long long image_pitch_bytes,
                                    const pslam_extract_cfg* ecfg, const pslam_match_cfg* mcfg);
/* same, HOST images (pinned or pageable): copies inside the call */
int pslam_stereo_frontend_batch(pslam_ctx* ctx, const uint8_t* h_images, int n_pairs, int rows,
                                int cols, int stride, long long image_pitch_bytes,
                                const pslam_extract_cfg* ecfg, const pslam_match_cfg* mcfg);
int pslam_download_stereo_counts(pslam_ctx* ctx, int n_pairs, int* counts);
/* per pair: 4 floats (uL,vL,uR,vR), left feature index, right feature index, distance */
int pslam_download_stereo_points(pslam_ctx* ctx, int pair, int capacity, float* uvuv, int* left_idx,
                                 int* right_idx, float* distance);

/* whole batch at once, packed: offsets[n_pairs + 1] (CSR), then per stereo point (uL,vL,uR,vR), the LEFT
 * feature's intensity and 32-byte descriptor (the PointIntensityDescriptor4f cloud the adaptor emits,
 * raw_data_preprocessor_stereo_projective.cpp:112-118), left / right feature index and Hamming distance.
 * Any output pointer except offsets may be NULL.  Returns the total number of points. */
int pslam_download_stereo_batch(pslam_ctx* ctx, int n_pairs, long long capacity_points, long long* offsets,
                                float* uvuv, float* intensity, uint8_t* desc, int* left_idx, int* right_idx,
                                float* distance);

/* ---- N1 (next after the path): rigid-stereo triangulation ----------------------------
 * Replaces TriangulatorRigidStereo::compute / triangulateRectifiedMidpoint
 *   (.../mapping/triangulator_rigid_stereo.cpp:7-85), the direct consumer of the stereo adaptor's (uL,vL,uR,vR)
 * cloud.  baseline_pixels_x = (K * t_right_in_left).x (:104-105).  xyz: 3 floats per input point, index-aligned
 * with the input; valid[i] = 0 marks the INVALID placeholders of points with xL - xR < minimum_disparity_pixels
 * (:38-45).  Returns the number of valid points. */
int pslam_triangulate(pslam_ctx* ctx, int n, const float* uvuv, const float* K9, float baseline_pixels_x,
                      float minimum_disparity_pixels, float infinity_depth_meters, float* xyz, uint8_t* valid);

/* ---- N2 (next after the path): projective scene clipping --------------------------------
 * Replaces SceneClipperProjective3D::compute (.../mapping/scene_clipper_projective_3d.cpp:9-67): the pinhole
 * projector (srrg2_core PointProjectorPinhole_::compute) over the WHOLE local map with the camera at
 * robot_in_local_map * sensor_in_robot (:46); survivors keep the map order (:52) and are moved into the robot frame
 * when sensor_in_robot is not the identity (:60-62).  Outputs per survivor: point in the sensor / robot frame (xyz),
 * projection (u, v, depth), index into the full scene (SceneClipper::globalIndices) and the copied descriptor.  Any
 * output pointer may be NULL.  Poses are row-major 3x4 [R|t].  Returns the number of survivors (or PSLAM_E_*;
 * PSLAM_E_CAPACITY when more than `capacity` points survive). */
typedef struct pslam_clip_cfg {
  float K[9];
  int canvas_rows, canvas_cols;
  float range_min, range_max;     /* projector range [m] (kitti.conf:172-179) */
  float camera_in_map[12];        /* robot_in_local_map * sensor_in_robot */
  float sensor_in_robot[12];
  int apply_sensor_in_robot;      /* 0: sensor_in_robot is the identity, nothing is applied (like the reference) */
} pslam_clip_cfg;
int pslam_scene_clip(pslam_ctx* ctx, int n, const float* xyz, const uint8_t* desc_or_null, const pslam_clip_cfg* cfg,
                     int capacity, float* out_xyz, float* out_uvz, int* out_index, uint8_t* out_desc);
/* device-resident map and outputs (all d_* are device pointers sized for n points; d_desc / outputs may be NULL);
 * the launch is repeated `reps` times and *ms_per_call (if not NULL) receives the mean CUDA-event time of one pass */
int pslam_scene_clip_dev(pslam_ctx* ctx, long long n, const float* d_xyz, const uint32_t* d_desc, const pslam_clip_cfg* cfg,
                         float* d_out_xyz, float* d_out_uvz, int* d_out_index, uint32_t* d_out_desc, long long* n_out,
                         int reps, double* ms_per_call);

/* ---- N3 (next after the path): per-landmark EKF update of the merger ----------------------
 * Replaces LandmarkEstimatorEKF_::compute (.../mapping/landmarks/landmark_estimator_ekf_impl.cpp:17-82) with its
 * filters PointEKFBase::_predict/_correct (.../landmarks/filters/point_ekf_base.hpp:62-131) and the three measurement
 * models (.../filters/{projective,projective_depth,stereo_projective}_point_ekf_impl.cpp), called by
 * MergerProjective_::_updatePoint once per correspondence (.../mergers/merger_projective_impl.cpp:129,176-193).
 * n landmarks that share the frame's transforms (LandmarkEstimatorBase_::setTransforms,
 * landmark_estimator_base.hpp:49-58) are filtered in one launch, double precision inside like the reference.
 * state_world [n][3] and covariance [n][9] are the landmark statistics (state(), covariance()): updated IN PLACE for
 * the landmarks the estimator accepts (what addOptimizationResult stores, :74-75); coords_in_local_map [n][3] receives
 * their new local coordinates (:79-80), inlier[n] = statistics().isInlier().  measurements: [n][E], E = 2 / 3 / 4.
 * Returns the number of inliers.
 * CALLER'S PART of addOptimizationResult: the reference also counts one optimisation per accepted landmark
 * (statistics().numberOfOptimizations(), read by the aligner's 1 + log(n) weighting, aligner_slice_processor_projective.cpp:
 * 41-57, and by the weighted mean).  The EKF and weighted-mean entry points take that counter read-only or not at all:
 * the owner of the statistics increments it for every landmark with inlier[i] != 0 (the smoother entry point updates it
 * itself, like its reference code does).  tests/test_gpu_tracker_sequence.py chains five frames this way. */
typedef struct pslam_ekf_cfg {
  int kind;                                        /* 0 projective (u,v) | 1 projective depth (u,v,z) | 2 rectified stereo (uL,vL,uR,vR) */
  float K[9];                                      /* filter->setCameraMatrix */
  double baseline_pixels[2];                       /* StereoProjectivePointEKF::setBaseline (b_x, b_y) */
  double minimum_state_element_covariance;         /* PARAM, default 0.01 (landmark_estimator_ekf.h:33) */
  double maximum_covariance_norm_squared;          /* PARAM, default 1 (:38) */
  float maximum_distance_geometry_meters_squared;  /* PARAM, default 1 (landmark_estimator_base.hpp:22) */
  float sensor_in_world[12];                       /* setTransforms(measurement_in_world, measurement_in_scene) */
  float sensor_in_local_map[12];
} pslam_ekf_cfg;
int pslam_landmarks_ekf_update(pslam_ctx* ctx, int n, float* state_world, float* covariance, const float* measurements,
                               const pslam_ekf_cfg* cfg, float* coords_in_local_map, uint8_t* inlier);
/* device-resident variant (all d_* device pointers); `reps` launches, *ms_per_call = mean CUDA-event time of one */
int pslam_landmarks_ekf_update_dev(pslam_ctx* ctx, long long n, float* d_state_world, float* d_covariance,
                                   const float* d_measurements, const pslam_ekf_cfg* cfg, float* d_coords_in_local_map,
                                   uint8_t* d_inlier, int* n_inliers, int reps, double* ms_per_call);

/* LandmarkEstimatorWeightedMean_::compute (.../mapping/landmarks/landmark_estimator_weighted_mean_impl.cpp:7-41) over the
 * n landmarks of one merger pass: running mean of the landmark's world position with the point re-observed in the sensor
 * frame (landmark_in_sensor [n][3], setLandmarkInSensor), weight = numberOfOptimizations + 1, geometric-distance gate.
 * fp32, bit exact.  state_world is updated in place for the inliers; returns their number. */
int pslam_landmarks_weighted_mean_update(pslam_ctx* ctx, int n, float* state_world, const int* number_of_optimizations,
                                         const float* landmark_in_sensor, const float* sensor_in_world12,
                                         const float* sensor_in_local_map12, float maximum_distance_geometry_meters_squared,
                                         float* coords_in_local_map, uint8_t* inlier);

/* LandmarkEstimatorPoseBasedSmoother_::compute (.../mapping/landmarks/landmark_estimator_pose_based_smoother_impl.cpp:6-148,
 * kitti.conf "landmark_estimator_smoother") over the n landmarks of one merger pass, fp32, bit exact against the CPU
 * restatement.  Every landmark brings its WHOLE measurement history, the current measurement included (the reference
 * appends it first, :15-19), in CSR form: offsets[n + 1]; per measurement the frame it was taken in (index into
 * frames_sensor_in_world [n_frames][12], row-major 3x4), its image point (u, v) and the point in that camera frame
 * (setLandmarkInSensor).  state_world / number_of_optimizations are the landmark statistics, updated in place exactly
 * where the reference updates them (mean below minimum_number_of_measurements_for_optimization, Gauss-Newton result or
 * mean reset above); coords_in_local_map[n][3], inlier[n] = isInlier.  Returns the number of inliers. */
typedef struct pslam_smoother_cfg {
  float K[9];                                                  /* setCameraMatrix */
  unsigned maximum_number_of_iterations;                       /* PARAM, default 100 */
  float convergence_criterion_minimum_chi2_delta;              /* PARAM, default 1e-5 */
  float maximum_reprojection_error_pixels_squared;             /* PARAM, default 100 */
  unsigned minimum_number_of_measurements_for_optimization;    /* PARAM, default 3 */
  float maximum_distance_geometry_meters_squared;              /* PARAM, default 1 */
  float sensor_in_world[12], sensor_in_local_map[12];          /* setTransforms of the current frame */
} pslam_smoother_cfg;
int pslam_landmarks_smoother_update(pslam_ctx* ctx, int n, float* state_world, int* number_of_optimizations, int n_frames,
                                    const float* frames_sensor_in_world, const int* offsets, const int* hist_frame,
                                    const float* hist_uv, const float* hist_point_in_camera, const pslam_smoother_cfg* cfg,
                                    float* coords_in_local_map, uint8_t* inlier);

/* MergerProjective_::compute binning (.../mapping/mergers/merger_projective_impl.cpp:8-171 update pass, :194-309
 * _addPoints), the step that decides WHICH correspondences update their landmark and WHICH measurements become new
 * landmarks; the per-landmark update itself is pslam_landmarks_*_update above, the triangulation of the additions
 * pslam_triangulate.  measurements [n_meas][dim]: dim 4 = stereo (uL, vL, uR, vR), dim 3 = (u, v, depth); a measurement's
 * bin is (round(v / (canvas_rows / number_of_row_bins)), round(u / (canvas_cols / number_of_col_bins))) (:82-83), bins
 * form an (number_of_row_bins + 1) x (number_of_col_bins + 1) grid; a bin outside it (the reference asserts it cannot
 * happen, :84-85) returns PSLAM_E_INVALID.  Decisions are identical to the reference's sequential walk, bit for bit. */
enum { PSLAM_MERGER_BASE = 0,     /* _isBetterForAddition = false (merger_projective.h:89-92): first arrival keeps the bin */
       PSLAM_MERGER_STEREO = 1,   /* larger disparity uL - uR wins (merger_projective_rigid_stereo_impl.cpp:42-52) */
       PSLAM_MERGER_DEPTH = 2 };  /* smaller depth wins (merger_projective_depth_ekf_impl.cpp:44-52) */
typedef struct pslam_merger_cfg {
  int canvas_rows, canvas_cols;                 /* param_projector: canvas_rows / canvas_cols */
  int number_of_row_bins, number_of_col_bins;   /* PARAM, defaults 10 / 30 */
  float maximum_distance_appearance;            /* PARAM, default 50 */
  int enable_binning;                           /* MergerCorrespondence_ PARAM enable_binning */
  int kind;                                     /* PSLAM_MERGER_* */
} pslam_merger_cfg;
/* number of 32-bit words of the blocked-bin bitmap (bit = bin_row * (number_of_col_bins + 1) + bin_col) */
int pslam_merger_occupancy_words(const pslam_merger_cfg* cfg);
/* update pass (:61-135): selected[c] = 1 where the reference calls _updatePoint for correspondence c (appearance gate
 * passed, bin not yet blocked by an earlier correspondence); occupied_bins receives the blocked bins for the addition
 * pass.  corr_moving[c] = measurement index, corr_response[c] = matching distance.  Returns the number selected. */
int pslam_merger_select_updates(pslam_ctx* ctx, const float* measurements, int dim, int n_meas, const int* corr_moving,
                                const float* corr_response, int n_corr, const pslam_merger_cfg* cfg, uint8_t* selected,
                                uint32_t* occupied_bins);
/* addition pass (:205-253): winners[k] = measurement index of the k-th entry of the reference's points_in_image_to_add
 * (one per free bin, in order of the bin's first arrival, the occupant chosen by _isBetterForAddition); occupied_bins
 * may be NULL (no correspondences: nothing is blocked).  winners has room for n_meas entries.  Returns their number. */
int pslam_merger_select_additions(pslam_ctx* ctx, const float* measurements, int dim, int n_meas, const pslam_merger_cfg* cfg,
                                  const uint32_t* occupied_bins, int* winners);

/* both passes in one call (one upload, two launches, one download): the addition candidates only depend on the bins the
 * update pass blocks, not on the estimator's verdicts, so they can be computed before the caller knows whether
 * compute() will reach _addPoints (:158-165) -- it ignores them if not.  Returns the number selected, *n_winners the
 * number of addition candidates. */
int pslam_merger_plan(pslam_ctx* ctx, const float* measurements, int dim, int n_meas, const int* corr_moving,
                      const float* corr_response, int n_corr, const pslam_merger_cfg* cfg, uint8_t* selected, uint32_t* occupied_bins,
                      int* winners, int* n_winners);

/* ---- stage 2b: exhaustive Hamming matching --------------------------------------
 * Replaces CorrespondenceFinderDescriptorBasedBruteforce::compute
 *   (.../correspondence_finders/correspondence_finder_descriptor_based_bruteforce_impl.cpp:6-294).
 * pslam_match_bruteforce returns the bijective, Lowe-checked correspondence set in the
 * reference's output order.  pslam_bf_best2* is the raw sweep (per fixed row: smallest and
 * second smallest distance and the argmin, first index wins ties); `second` is INT32_MAX
 * and `best_idx` -1 where absent. */
int pslam_match_bruteforce(pslam_ctx* ctx, int n_fixed, const uint8_t* desc_fixed, int n_moving,
                           const uint8_t* desc_moving, const pslam_match_cfg* cfg, int capacity,
                           int* fixed_idx, int* moving_idx, float* distance);
int pslam_bf_best2(pslam_ctx* ctx, int n_fixed, const uint8_t* desc_fixed, int n_moving,
                   const uint8_t* desc_moving, int32_t* best, int32_t* second, int32_t* best_idx);
int pslam_bf_best2_dev(pslam_ctx* ctx, int n_fixed, const uint32_t* d_desc_fixed, int n_moving,
                       const uint32_t* d_desc_moving, int32_t* d_best, int32_t* d_second,
                       int32_t* d_best_idx);

/* ---- multi-GPU: one process per GPU ---------------------------------
 * Frames of the stage-1 / stereo pipeline are independent: every rank runs the batched entry points on its own frame
 * range (pslam_shard_frames), no collective.  The exhaustive Hamming sweep shards its QUERY rows with the train set
 * replicated -- the pair loop of CorrespondenceFinderDescriptorBasedBruteforce::compute
 * (.../correspondence_finder_descriptor_based_bruteforce_impl.cpp:32-74) is a per-row reduction -- and ends with ONE
 * ncclAllGather of (best, second, argmin) on the context's stream.  libnccl is resolved at run time (the copy the host
 * process already loaded, else the system's); nccl_comm is an ncclComm_t.  A host that has no communicator yet creates
 * one with the two helpers below: rank 0 draws the id, the application broadcasts its 128 bytes, every rank calls create. */
typedef struct { char internal[128]; } PslamNcclId;  /* = ncclUniqueId */
int pslam_shard_frames(int n_frames, int rank, int world, int* begin, int* end);
/* rows [begin, end) of `rank`; returns the shard size ceil(n / world) rounded up to `align` (256 = one sweep CTA) */
int pslam_shard_rows(int n_rows, int rank, int world, int align, int* begin, int* end);
int pslam_nccl_unique_id(pslam_ctx* ctx, PslamNcclId* id);
int pslam_nccl_comm_create(pslam_ctx* ctx, const PslamNcclId* id, int rank, int world, void** nccl_comm);
int pslam_nccl_comm_destroy(pslam_ctx* ctx, void* nccl_comm);
/* The same sweep with the exchange step FUSED into the merge kernel, no collective: every rank exports one result table of
 * its HBM through CUDA IPC (pslam_p2p_table_export), the application distributes the 64-byte handles like it distributes the
 * NCCL id, every rank maps its peers' tables (pslam_p2p_table_import); the merge kernel of pslam_bf_best2_sharded_p2p_dev
 * then stores every row's (best, second, argmin) straight into ALL ranks' tables -- coalesced stores over NVLink / NVSwitch --
 * followed by one flag per source rank; a rank hands its table out as soon as all flags of the call's epoch are up.  All
 * ranks must issue the same sequence of p2p calls. */
typedef struct { char internal[64]; } PslamIpcHandle;  /* = cudaIpcMemHandle_t */
int pslam_p2p_table_export(pslam_ctx* ctx, int max_rows, PslamIpcHandle* mine);
int pslam_p2p_table_import(pslam_ctx* ctx, int rank, int world, const PslamIpcHandle* handles_of_all_ranks);
int pslam_p2p_table_release(pslam_ctx* ctx);
int pslam_bf_best2_sharded_p2p_dev(pslam_ctx* ctx, int n_fixed, const uint32_t* d_desc_fixed, int n_moving,
                                   const uint32_t* d_desc_moving, int32_t* d_best, int32_t* d_second, int32_t* d_best_idx);
/* every rank passes the SAME full descriptor sets (device pointers) and receives the full tables [n_fixed] */
int pslam_bf_best2_sharded_dev(pslam_ctx* ctx, void* nccl_comm, int rank, int world, int n_fixed,
                               const uint32_t* d_desc_fixed, int n_moving, const uint32_t* d_desc_moving,
                               int32_t* d_best, int32_t* d_second, int32_t* d_best_idx);

/* ---- stage 2c: projective window matching ---------------------------------------
 * Replaces the device-worthy part of CorrespondenceFinderProjective{Square,Circle,Rhombus}:
 * PointProjectorPinhole_::compute + _initializeDatabase + _findNearestNeighbors +
 * _filterCorrespondences
 *   (.../correspondence_finders/correspondence_finder_projective_base_impl.cpp:39-102,165-208,
 *    ..._square_impl.cpp:7-118, ..._circle_impl.cpp:7-94, ..._rhombus_impl.cpp:7-93).
 * The adaptive state machine (:104-293) stays in the host wrapper.  Output is in the order the
 * reference's std::unordered_map iteration yields.  shape: 0 square, 1 circle, 2 rhombus,
 * 3 = CorrespondenceFinderProjectiveKDTree::_findNearestNeighbors (..._kdtree_impl.cpp:28-79) as the EXACT radius query:
 * every fixed point whose squared fp32 distance to the projection is < radius^2, best initialised to
 * maximum_descriptor_distance, only the best candidate recorded.  The reference takes its candidates from srrg2_core's
 * approximate KDTree (the query's leaf cluster only): the exact query returns a superset, parity for shape 3 is unpinned. */
typedef struct {
  float K[9];
  int canvas_rows, canvas_cols;
  float range_min, range_max;
  int shape;
  int search_radius_pixels;
  float descriptor_distance;                   /* current (adaptive) threshold */
  float maximum_distance_ratio_to_second_best;
  float maximum_descriptor_distance;           /* shape 3 only: PARAM maximum_descriptor_distance (bruteforce.h:23-27) */
} pslam_projective_cfg;

/* setFixed (+ _initializeDatabase, cached until the next call) / setMoving / one search.  The
 * reference rebuilds its lattice only when the fixed cloud changes (base_impl.cpp:109-134). */
int pslam_projective_set_fixed(pslam_ctx* ctx, int n_fixed, const float* fixed_coords, int fixed_dim,
                               const uint8_t* desc_fixed);
int pslam_projective_set_moving(pslam_ctx* ctx, int n_moving, const float* moving_xyz,
                                const uint8_t* desc_moving);
/* The cached clouds live in an allocation of their own (no other entry point touches it).  Every set_fixed / set_moving
 * stamps the cache with a process-unique, never-zero epoch: a caller that shares the context with other finder instances
 * (or whose context was re-created) compares the epochs with the ones it saw after its own uploads and uploads again
 * when they differ -- the host finder mirror does exactly that. */
int pslam_projective_cache_epochs(const pslam_ctx* ctx, unsigned long long* fixed_epoch, unsigned long long* moving_epoch);
int pslam_projective_match(pslam_ctx* ctx, int n_fixed, int n_moving, const float* local_map_in_sensor12,
                           const pslam_projective_cfg* cfg, int capacity, int* fixed_idx,
                           int* moving_idx, float* distance, int* n_projected);
/* convenience: the three calls above in one */
int pslam_match_projective(pslam_ctx* ctx, int n_fixed, const float* fixed_coords, int fixed_dim,
                           const uint8_t* desc_fixed, int n_moving, const float* moving_xyz,
                           const uint8_t* desc_moving, const float* local_map_in_sensor12,
                           const pslam_projective_cfg* cfg, int capacity, int* fixed_idx,
                           int* moving_idx, float* distance, int* n_projected);

/* ---- stages 3+4: SE3 factor linearisation and H,b reduction ---------------------
 * Replaces srrg2_solver FactorCorrespondenceDriven_::compute over
 * SE3RectifiedStereoProjectiveErrorFactor / SE3ProjectiveDepthErrorFactor /
 * SE3ProjectiveErrorFactor with RobustifierSaturated / Clamp, as wired by
 * AlignerSliceProcessorProjective_::setupFactor
 *   (.../registration/aligner_slice_processor_projective.cpp:27-112; factor use in
 *    tests/test_aligners.cpp:586-638).  Arithmetic is fp64.
 * kind: 0 stereo (fixed_dim 4), 1 depth (3), 2 mono (2).  robustifier: 0 none, 1 saturated, 2 clamp.
 * info_diag: 3 doubles per FIXED point.  stats4 = {chi, inliers, outliers, suppressed}. */
typedef struct {
  int kind;
  double K[9];
  double image_cols, image_rows;
  double baseline[3];
  double mean_disparity; /* > 0 enables inverse-depth weighting */
  int robustifier;
  double chi_threshold;
} pslam_linearize_cfg;

/* SE3 pose-prior factor summed into the same 6x6 system before the solve: the second slice of the shipped aligners,
 * AlignerSliceMotionModel3D (configurations/kitti.conf:747-772, icl.conf:268-293, euroc.conf:94-119; srrg2_slam_interfaces).
 * e = t2tnq(prediction^-1 * X), constant 6x6 information (row-major); evaluated on the device in every fused iteration. */
typedef struct {
  double prediction[12];  /* predicted moving_in_fixed, row-major 3x4 */
  double information[36];
} pslam_pose_prior;

/* per-correspondence outcome of a linearisation (srrg2_solver FactorStats status): what MultiAligner's
 * enable_inlier_only_runs / keep_only_inlier_correspondences read (configurations/icl.conf:55-58) */
#define PSLAM_FACTOR_INLIER 0
#define PSLAM_FACTOR_KERNELIZED 1
#define PSLAM_FACTOR_SUPPRESSED 2

int pslam_linearize_se3(pslam_ctx* ctx, const pslam_linearize_cfg* cfg, const double* pose12,
                        int n_moving, const double* moving_xyz, int n_fixed, const double* fixed_meas,
                        int fixed_dim, int n_corr, const int* corr_fixed, const int* corr_moving,
                        const double* info_diag, double* H36, double* b6, double* stats4);
/* measurement aid for bench.py: the same linearisation on a batched synthetic, inputs resident in HBM, `reps` passes
 * timed with CUDA events; *ms_per_call receives the device time of one linearise + reduce pass */
int pslam_linearize_se3_timed(pslam_ctx* ctx, const pslam_linearize_cfg* cfg, const double* pose12, int n_moving,
                              const double* moving_xyz, int n_fixed, const double* fixed_meas, int fixed_dim, int n_corr,
                              const int* corr_fixed, const int* corr_moving, const double* info_diag, int reps,
                              double* ms_per_call);
/* n_iterations x { linearise -> H, b -> (H + damping I) dx = -b -> pose <- pose * v2t(dx) } in ONE kernel launch: what the
 * aligner runs between two re-projections of the correspondence finder (correspondences and information matrices are
 * constant in between).  pose12 in/out; poses12 (n_iterations x 12, pose after each update) and
 * stats4 (n_iterations x {chi, inliers, outliers, suppressed}) may be NULL.  *iterations_done receives the number of
 * linearised iterations (their stats are valid); PSLAM_E_NOT_SPD when the last of them could not be solved (the pose
 * keeps its last valid value), PSLAM_OK otherwise. */
int pslam_gn_iterate(pslam_ctx* ctx, const pslam_linearize_cfg* cfg, int n_iterations, double damping, double* pose12,
                     int n_moving, const double* moving_xyz, int n_fixed, const double* fixed_meas, int fixed_dim,
                     int n_corr, const int* corr_fixed, const int* corr_moving, const double* info_diag,
                     double* poses12, double* stats4, int* iterations_done);
/* The same two calls on fp32 clouds -- the reference's own cloud scalar (PointIntensityDescriptor_<D, float>): 40 B per
 * correspondence travel and stay in HBM (3 + 4 + 3 floats), the widening to fp64 happens in registers (exact), nothing is
 * converted on the host.  prior (may be NULL): pose-prior factor added to H, b (stats5[4] / not part of stats4's chi).
 * factor_status (may be NULL): n_corr bytes, PSLAM_FACTOR_* of the (last) linearisation. */
int pslam_linearize_se3_f32(pslam_ctx* ctx, const pslam_linearize_cfg* cfg, const double* pose12, int n_moving,
                            const float* moving_xyz, int n_fixed, const float* fixed_meas, int fixed_dim, int n_corr,
                            const int* corr_fixed, const int* corr_moving, const float* info_diag,
                            const pslam_pose_prior* prior, uint8_t* factor_status, double* H36, double* b6, double* stats5);
int pslam_linearize_se3_timed_f32(pslam_ctx* ctx, const pslam_linearize_cfg* cfg, const double* pose12, int n_moving,
                                  const float* moving_xyz, int n_fixed, const float* fixed_meas, int fixed_dim, int n_corr,
                                  const int* corr_fixed, const int* corr_moving, const float* info_diag, int reps,
                                  double* ms_per_call);
int pslam_gn_iterate_f32(pslam_ctx* ctx, const pslam_linearize_cfg* cfg, int n_iterations, double damping, double* pose12,
                         int n_moving, const float* moving_xyz, int n_fixed, const float* fixed_meas, int fixed_dim,
                         int n_corr, const int* corr_fixed, const int* corr_moving, const float* info_diag,
                         const pslam_pose_prior* prior, double* poses12, double* stats4, uint8_t* factor_status,
                         int* iterations_done);
/* (H + damping I) dx = -b, pose <- pose * v2t(dx) on the device (one thread, fp64 Cholesky) */
int pslam_gn_step(pslam_ctx* ctx, const double* H36, const double* b6, double damping, double* pose12,
                  double* dx6);

/* One finder phase of an aligner loop in ONE device round trip: the search + filter of pslam_projective_match and, on the
 * correspondences it finds, `n_iterations` fused solver iterations (pslam_gn_iterate_f32) -- the reference's loop calls
 * finder->compute() and then linearises / solves until the finder re-projects (MultiAligner3DQR, SURVEY App. E.6;
 * correspondence_finder_projective_base_impl.cpp:162-178).  Both clouds are the finder's cached fp32 clouds; the
 * information of a correspondence is diagonal_info x information_scale[moving index] (setupFactor's 1 + log(n_opt) weighting,
 * aligner_slice_processor_projective.cpp:41-57; pslam_projective_set_moving_weights uploads the table once per moving cloud,
 * without it every scale is 1).  The solver visits the correspondences in ascending fixed index; factor_status comes back in
 * the order of the returned correspondences.  A caller that then rejects the correspondences (too few, low matching ratio)
 * simply ignores the solver outputs. */
typedef struct pslam_fused_gn {
  const pslam_linearize_cfg* factor;
  float diagonal_info[3];
  int n_iterations;
  double damping;
  double pose12[12];             /* in: estimate (moving in fixed); out: after the last completed iteration */
  const pslam_pose_prior* prior; /* or NULL */
  double* poses12;               /* out [n_iterations][12], may be NULL */
  double* stats4;                /* out [n_iterations][4]: chi, inliers, outliers, suppressed; may be NULL */
  unsigned char* factor_status;  /* out [capacity], may be NULL */
  int iterations_done;           /* out */
  int spd;                       /* out: 0 when H + damping I was not positive definite in the last iteration */
} pslam_fused_gn;
int pslam_projective_set_moving_weights(pslam_ctx* ctx, int n_moving, const float* information_scale);
int pslam_projective_match_gn(pslam_ctx* ctx, int n_fixed, int n_moving, const float* local_map_in_sensor12,
                              const pslam_projective_cfg* cfg, int capacity, int* fixed_idx, int* moving_idx,
                              float* distance, int* n_projected, pslam_fused_gn* gn);

/* ---- the whole registration of one frame, device resident ---------------------------------------------------------
 * MultiAligner_::compute drives { finder.compute(); solver iterations } until max_iterations are spent; between two
 * searches the finder only updates a handful of scalars (CorrespondenceFinderProjective_::compute,
 * correspondence_finder_projective_base_impl.cpp:104-293: re-project every N-th call, convergence = norm of the estimate
 * change below a threshold after a minimum of calls, then the correspondences are kept).  pslam_projective_align runs that
 * state machine on the device: search -> filter -> decisions -> solver iterations, phase after phase, with ONE download at
 * the end of every batch of phases instead of one host round trip per search.  Same per-iteration poses / stats and the
 * same final correspondences as the call-by-call path.
 * It stops and hands back to the caller's loop (the state below describes the point reached) when a decision needs the
 * caller: stop_reason 2 = low matching ratio with room to widen the search (the finder repeats the call with its maximum
 * radius), 3 = fewer than min_num_correspondences, 4 = H + damping I not positive definite.  The phase that stopped has
 * changed nothing: the caller re-runs it through pslam_projective_match_gn.
 * in / out finder state: current_iteration, has_converged, previous12 (_local_map_in_sensor_previous).
 * gn: as pslam_projective_match_gn, n_iterations = max_iterations = rows of poses12 / stats4. */
#define PSLAM_ALIGN_MAX_PHASES 64
typedef struct pslam_align {
  int max_iterations;                       /* solver iterations left in the aligner's budget */
  int solver_iterations_per_projection;     /* finder PARAM number_of_solver_iterations_per_projection */
  int minimum_number_of_iterations;         /* finder PARAM */
  float maximum_estimate_change_norm_for_convergence;
  float minimum_matching_ratio;
  int can_widen_search;                     /* search radius below its maximum or descriptor distance above its minimum */
  int min_num_correspondences;              /* slice PARAM (at least 1) */
  int current_iteration;                    /* in / out */
  int has_converged;                        /* in (must be 0) / out */
  float previous12[12];                     /* in / out */
  int stop_reason;                          /* out: 1 budget spent, 2 / 3 / 4 see above */
  int iterations_done;                      /* out: solver iterations completed (rows of poses12 / stats4 that are valid) */
  int converged_with_good_ratio;            /* out: has_converged was set in a phase whose matching ratio exceeded the minimum */
  int n_projected;                          /* out: of the last phase */
  int n_phases;                             /* out: searches executed */
  int phase_log[3 * PSLAM_ALIGN_MAX_PHASES]; /* out per phase: first solver iteration, iterations done, correspondences */
} pslam_align;
int pslam_projective_align(pslam_ctx* ctx, int n_fixed, int n_moving, const pslam_projective_cfg* cfg, pslam_align* align,
                           int capacity, int* fixed_idx, int* moving_idx, float* distance, pslam_fused_gn* gn);

#ifdef __cplusplus
}
#endif
#endif /* PSLAM_CUDA_H */
