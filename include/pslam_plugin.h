/* =============================================================================
 * pslam_plugin.h -- C view of the host-side plugin mirror (libpslam_plugin.so).
 *
 * The reference wires its modules by class name in a BOSS `.conf`
 * (configurations/kitti.conf, icl.conf, euroc.conf) and drives them through the
 * srrg2 Configurable interface.  libpslam_plugin.so holds the
 * C++ mirror of that interface for the frontend hot path (srrg2_proslam_b200/host/);
 * these entry points expose it to non-C++ callers (the Python parity tests read like
 * the reference's gtest files: load the .conf, look a module up by name, set PARAMs,
 * hand it clouds / images, call compute()).  All arithmetic happens in libpslam_cuda.so.
 *
 * Every function returns 0 / a count on success and a negative value on failure;
 * psp_last_error() then holds the text of the std::runtime_error the C++ module threw
 * (the same texts the reference throws).
 * ========================================================================== */
#ifndef PSLAM_PLUGIN_H
#define PSLAM_PLUGIN_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct psp_manager psp_manager; /* ConfigurableManager */
typedef struct psp_module psp_module;   /* a Configurable owned by its manager */

const char* psp_last_error(void);
int psp_set_device(int device);
/* pslam_profile_enable / pslam_profile_read (pslam_cuda.h) on the process-wide device context the modules share */
int psp_profile_enable(int enable);
int psp_profile_read(int capacity, char* names, int name_len, double* total_ms, long long* launches);

/* ---- ConfigurableManager: srrg2_core::ConfigurableManager::read / getByName / create ---------- */
psp_manager* psp_manager_create(void);
void psp_manager_destroy(psp_manager* m);
int psp_manager_read(psp_manager* m, const char* conf_path);
int psp_manager_read_string(psp_manager* m, const char* conf_text);
/* writes the named modules and everything they link to (all modules when n_names == 0) */
int psp_manager_write(psp_manager* m, const char* conf_path, int n_names, const char* const* names);
int psp_manager_count(psp_manager* m);
psp_module* psp_manager_at(psp_manager* m, int index);
psp_module* psp_manager_get_by_name(psp_manager* m, const char* name);
psp_module* psp_manager_create_module(psp_manager* m, const char* class_name, const char* name);
/* 1 when `class_name` is backed by a CUDA class of this library, 0 when it would load as a generic module */
int psp_class_is_registered(const char* class_name);

/* ---- Configurable: class name, PARAM access, links ------------------------------------------ */
const char* psp_module_class_name(psp_module* c);
const char* psp_module_name(psp_module* c);
int psp_module_is_generic(psp_module* c);
int psp_module_has_param(psp_module* c, const char* param);
int psp_module_set_number(psp_module* c, const char* param, double value);
int psp_module_get_number(psp_module* c, const char* param, double* value);
int psp_module_set_string(psp_module* c, const char* param, const char* value);
const char* psp_module_get_string(psp_module* c, const char* param);
int psp_module_set_numbers(psp_module* c, const char* param, int n, const double* values);
int psp_module_get_numbers(psp_module* c, const char* param, int capacity, double* values);
int psp_module_set_link(psp_module* c, const char* param, psp_module* target_or_null);
psp_module* psp_module_get_link(psp_module* c, const char* param);

/* ---- IntensityFeatureExtractorBinned{2D,3D}: setFeatures + compute(image) -------------------- */
/* IntensityFeatureExtractorSelective_: the tracking detection mask (rows x cols bytes, 1 inside the rectangles) painted
 * around `n` projections for a detection radius (intensity_feature_extractor_selective.cpp:63-144).  Host-side. */
int psp_extractor_paint_tracking_mask(psp_module* extractor, int rows, int cols, int n, int dim, const float* coords, int radius,
                                      uint8_t* mask);
int psp_extractor_compute(psp_module* extractor, const uint8_t* image, int rows, int cols, int stride,
                          const uint8_t* mask_or_null, int capacity, float* xy, float* intensity, uint8_t* desc);

/* setProjections(cloud, radius) (base.h:100-106): the next compute() of a selective extractor runs in tracking mode;
 * coords: n x dim floats (x, y first).  psp_extractor_number_of_tracking_keypoints: size of the tracking group of
 * the last compute (IntensityFeatureExtractorSelective{2D,3D} only). */
int psp_extractor_set_projections(psp_module* extractor, int n, int dim, const float* coords, int radius);
int psp_extractor_number_of_tracking_keypoints(psp_module* extractor);

/* ---- RawDataPreprocessorStereoProjective: setRawData + setMeas + compute ---------------------
 * uvuv: 4 floats per point.  *status receives the adaptor's _status (0 Error, 1 Initializing, 2 Ready). */
int psp_stereo_adaptor_compute(psp_module* adaptor, const uint8_t* left, const uint8_t* right, int rows, int cols,
                               int stride, int capacity, float* uvuv, float* intensity, uint8_t* desc, int* status);
/* ---- RawDataPreprocessorMonocularDepth (depth_type 0 = uint16, 1 = float) -------------------- */
int psp_mono_depth_adaptor_compute(psp_module* adaptor, const uint8_t* image, int rows, int cols, int stride,
                                   const void* depth, int depth_type, int depth_rows, int depth_cols,
                                   int depth_stride_elements, int capacity, float* uvz, float* intensity,
                                   uint8_t* desc, int* status);

/* ---- CorrespondenceFinder*: setFixed / setMoving / setLocalMapInSensor / compute --------------
 * Clouds are copied into the module (the reference holds non-owning pointers; the copy keeps the C
 * interface free of lifetime rules).  coords: n x dim floats, desc: n x 32 bytes. */
int psp_finder_set_fixed(psp_module* finder, int n, int dim, const float* coords, const uint8_t* desc);
int psp_finder_set_moving(psp_module* finder, int n, int dim, const float* coords, const uint8_t* desc);
int psp_finder_set_local_map_in_sensor(psp_module* finder, const float* pose12);
int psp_finder_compute(psp_module* finder, int capacity, int* fixed_idx, int* moving_idx, float* response);
/* projective finders: dynamic state (projective_base.h:82-97,130-154): radius, descriptor distance,
 * iteration, converged flag, number of device searches so far */
int psp_projective_finder_state(psp_module* finder, int* search_radius_pixels, float* descriptor_distance,
                                int* current_iteration, int* has_converged, int* number_of_searches);
int psp_projective_finder_set_state(psp_module* finder, int search_radius_pixels, float descriptor_distance);
/* camera matrix of a PointIntensityDescriptor3fProjectorPinhole (set at runtime from camera info) */
int psp_projector_set_camera_matrix(psp_module* projector, const float* K9);

/* ---- MultiAligner3DQR + AlignerSliceProcessorProjective{,Depth,Stereo} ------------------------
 * fixed: the adapted measurements (dim 4 stereo / 3 depth / 2 mono); moving: 3-D scene points.
 * n_opt_or_null: statistics().numberOfOptimizations() per moving point. */
int psp_aligner_set_fixed(psp_module* aligner, int n, int dim, const float* coords, const uint8_t* desc);
int psp_aligner_set_moving(psp_module* aligner, int n, const float* xyz, const uint8_t* desc, const int* n_opt_or_null);
int psp_aligner_set_moving_in_fixed(psp_module* aligner, const float* pose12);
/* platform transform used by the stereo slice: translation of the left camera in the right camera [m] */
int psp_aligner_set_left_camera_in_right(psp_module* aligner, const float* t3);
/* returns the aligner status (0 Fail, 1 Success, 2 NotEnoughCorrespondences, 3 NotEnoughInliers) */
int psp_aligner_compute(psp_module* aligner, double* moving_in_fixed12, int* iterations, int* num_correspondences,
                        int* num_inliers, double* chi);
/* per-iteration stats of the last compute: rows of (num_correspondences, num_inliers, num_outliers, chi) */
int psp_aligner_iteration_stats(psp_module* aligner, int capacity, double* rows4);
/* aligner->param_slice_processors.setValue(index, slice) (tests/test_aligners.cpp:901) / .size() */
int psp_aligner_set_slice_processor(psp_module* aligner, int index, psp_module* slice);
int psp_aligner_num_slice_processors(psp_module* aligner);
/* the "trajectory_chunk" slice of the fixed / moving scene (robot poses in the local map, oldest first; n may be 0) read by
 * the aligner's AlignerSliceMotionModel3D, and the constant information matrix of its pose-prior factor (default identity) */
int psp_aligner_set_trajectory_chunk(psp_module* aligner, int n, const float* poses12);
int psp_aligner_set_prior_information(psp_module* aligner, const double* information36);
/* rows of the additional inlier-only run (enable_inlier_only_runs, configurations/icl.conf:55); returns their number */
int psp_aligner_inlier_run_stats(psp_module* aligner, int capacity, double* rows4);
/* correspondences of the projective slice after the last compute */
int psp_aligner_correspondences(psp_module* aligner, int capacity, int* fixed_idx, int* moving_idx, float* response);

/* ---- SceneClipperProjective3D (mapping/scene_clipper_projective_3d.h:10-40): setFullScene / setRobotInLocalMap /
 * setSensorInRobot / compute / globalIndices.  intensity may be NULL.  psp_clipper_compute returns the number of
 * clipped points (in the robot frame, map order), their projections (u, v, depth), indices into the full scene and
 * descriptors; *status receives SceneClipper::Status (0 Error, 1 Ready, 2 Successful). */
int psp_clipper_set_full_scene(psp_module* clipper, int n, const float* xyz, const float* intensity, const uint8_t* desc);
int psp_clipper_set_robot_in_local_map(psp_module* clipper, const float* pose12);
int psp_clipper_set_sensor_in_robot(psp_module* clipper, const float* pose12);
int psp_clipper_compute(psp_module* clipper, int capacity, float* xyz, float* uvz, int* global_index, uint8_t* desc,
                        int* status);

/* ---- point EKFs + LandmarkEstimatorEKF (mapping/landmarks/filters/ headers, landmark_estimator_ekf.h:12-63):
 * filter->setCameraMatrix / setBaseline, estimator->setTransforms, then the batched form of setMeasurement /
 * setLandmark / compute over the correspondences of one merger pass (see pslam_landmarks_ekf_update). */
int psp_point_filter_set_camera(psp_module* filter, const float* K9, double baseline_x_pixels, double baseline_y_pixels);
int psp_landmark_estimator_set_transforms(psp_module* estimator, const float* measurement_in_world12, const float* measurement_in_scene12);
int psp_landmark_estimator_compute_batch(psp_module* estimator, int n, float* state_world, float* covariance,
                                         const float* measurements, float* coords_in_local_map, uint8_t* inlier);

/* LandmarkEstimatorWeightedMean{2D3D,3D3D,4D3D}: psp_landmark_estimator_set_transforms, then the batched compute */
int psp_landmark_estimator_weighted_mean_batch(psp_module* estimator, int n, float* state_world, const int* number_of_optimizations,
                                               const float* landmark_in_sensor, float* coords_in_local_map, uint8_t* inlier);

/* MergerRigidStereoTriangulation / MergerRigidStereoProjectiveEKF / MergerProjectiveDepthEKF
 * (mapping/mergers/merger_projective_impl.cpp:8-309): the binning of compute() on the device.
 *   psp_merger_select_updates    update pass (:61-135): selected[c] = 1 where _updatePoint is reached; the module keeps
 *                                the blocked bins.  measurements [n_meas][dim], dim 4 (uL,vL,uR,vR) or 3 (u,v,depth).
 *   psp_merger_wants_additions   1 when compute() would call _addPoints after `merged` successful updates (:56-58,:158-165)
 *   psp_merger_select_additions  binning of _addPoints (:205-253): winners[k] = measurement behind the k-th addition
 *                                candidate (room for n_meas); returns their number. */
int psp_merger_select_updates(psp_module* merger, const float* measurements, int dim, int n_meas, const int* corr_moving,
                              const float* corr_response, int n_corr, uint8_t* selected);
int psp_merger_wants_additions(psp_module* merger, int merged, int n_meas, int n_corr);
/* both passes in one device round trip (pslam_merger_plan); returns #selected, *n_winners = #addition candidates */
int psp_merger_plan(psp_module* merger, const float* measurements, int dim, int n_meas, const int* corr_moving, const float* corr_response,
                    int n_corr, uint8_t* selected, int* winners, int* n_winners);
int psp_merger_select_additions(psp_module* merger, const float* measurements, int dim, int n_meas, int* winners);

/* LandmarkEstimatorPoseBasedSmoother{2D3D,3D3D,4D3D}: setCameraMatrix, psp_landmark_estimator_set_transforms, then the
 * batched compute over CSR measurement histories (see pslam_landmarks_smoother_update) */
int psp_landmark_smoother_set_camera_matrix(psp_module* estimator, const float* K9);
int psp_landmark_smoother_compute_batch(psp_module* estimator, int n, float* state_world, int* number_of_optimizations, int n_frames,
                                        const float* frames_sensor_in_world, const int* offsets, const int* hist_frame,
                                        const float* hist_uv, const float* hist_point_in_camera, float* coords_in_local_map,
                                        uint8_t* inlier);

#ifdef __cplusplus
}
#endif
#endif /* PSLAM_PLUGIN_H */
