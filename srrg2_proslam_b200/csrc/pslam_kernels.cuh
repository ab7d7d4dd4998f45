// pslam_kernels.cuh -- host-callable launchers of the sm_100a kernels (one per pipeline step).
#pragma once
#include "pslam_internal.cuh"

// k_fast.cu / k_detect.cu
int pslam_k_upload_pattern(pslam_ctx* ctx);
int pslam_k_blur_border(pslam_ctx* ctx, const uint8_t* d_image, int rows, int cols, int stride);
int pslam_k_fast_blur(pslam_ctx* ctx, const uint8_t* d_images, long long image_pitch, int n_images,
                      int rows, int cols, int stride, int thr, int nms);
int pslam_k_bin_select(pslam_ctx* ctx, int n_images, int rows, int cols, int nh, int nv,
                       unsigned long long quota, const uint8_t* d_mask, int mask_invert);
// slot_base: feature-store slot of the chunk's first image (maps / raw lists are chunk-local)
int pslam_k_assemble(pslam_ctx* ctx, const uint8_t* d_images, long long image_pitch, int stride,
                     int n_images, int rows, int cols, int nbins, int border, int slot_base);
int pslam_k_describe(pslam_ctx* ctx, int n_images, int rows, int cols, int slot_base);
int pslam_k_orb_tile_cap(int max_rows, int max_cols);
int pslam_k_strips_cap(int max_cols);  // K1 strips of the widest image  // description tiles of the largest image
int pslam_k_make_blur_tmap(pslam_ctx* ctx, int work_images);
int pslam_k_mono_depth(pslam_ctx* ctx, const void* d_depth, int depth_type, int depth_rows, int depth_cols,
                       int depth_stride, float scale, int slot, float* d_uvz, float* d_inten, uint32_t* d_desc,
                       int* d_count);

// k_epipolar.cu
// general != 0: arbitrary host coordinates (bitonic sort); 0: coordinates produced by the extractor (row < max_rows)
int pslam_k_epipolar(pslam_ctx* ctx, int n_pairs, const pslam_match_cfg* cfg, int general, int pair_base = 0);

// packed (CSR) stereo result of a batch, carved from the context scratch: offsets[n_pairs + 1] then SoA
struct pslam_packed_stereo {
  long long* d_offsets;
  float4* d_uvuv;
  float* d_intensity;
  uint32_t* d_desc;
  int *d_left, *d_right;
  float* d_dist;
};
int pslam_k_pack_stereo(pslam_ctx* ctx, int n_pairs, pslam_packed_stereo* out);
int pslam_k_triangulate(pslam_ctx* ctx, const float4* d_uvuv, long long n, const float* K9, float b_x, float min_disparity,
                        float infinity_depth, float* d_xyz, unsigned char* d_valid);

// k_bruteforce.cu
// sweep of this rank's rows + merge that stores (best, second, argmin) into table[3][cap_rows] of every peer at row_offset
int pslam_k_bf_best2_p2p(pslam_ctx* ctx, int n_q, const uint32_t* d_q, int n_t, const uint32_t* d_t, int row_offset, int cap_rows,
                         int parity, int world, int* const* d_peers);
int pslam_k_bf_best2(pslam_ctx* ctx, int n_fixed, const uint32_t* d_desc_fixed, int n_moving,
                     const uint32_t* d_desc_moving, int32_t* d_best, int32_t* d_second,
                     int32_t* d_best_idx);
int pslam_k_bf_match(pslam_ctx* ctx, int n_fixed, const uint32_t* d_desc_fixed, int n_moving,
                     const uint32_t* d_desc_moving, float max_dist, float max_ratio, int capacity,
                     int* h_fixed, int* h_moving, float* h_dist);

// k_projective.cu
int pslam_k_projective_set_fixed(pslam_ctx* ctx, int n_fixed, const float* h_coords, int dim,
                                 const uint8_t* h_desc);
int pslam_k_projective_set_moving(pslam_ctx* ctx, int n_moving, const float* h_xyz, const uint8_t* h_desc);
int pslam_k_projective_set_moving_weights(pslam_ctx* ctx, int n_moving, const float* h_scale);
int pslam_k_projective_match(pslam_ctx* ctx, int n_fixed, int n_moving, const float* pose12,
                             const pslam_projective_cfg* cfg, int capacity, int* h_fixed,
                             int* h_moving, float* h_dist, int* n_projected, pslam_fused_gn* gn = nullptr);
// device-resident state of one fused alignment (pslam_projective_align): the finder's state machine + the solver's estimate
struct PslamAlignState {
  double estimate[12];  // solver's estimate (moving in fixed), 3x4 row-major
  float prev[12];       // finder's _local_map_in_sensor_previous
  int current_iteration, has_converged, converged_ratio_ok;
  int it;               // solver iterations done
  int stop;             // 0 running, 1 budget spent, 2 low matching ratio, 3 too few correspondences, 4 not positive definite
  int phases, n_fused, n_corr, n_projected, pad;
  int phase_log[3 * PSLAM_ALIGN_MAX_PHASES];
};
struct PslamAlignCfg {
  int max_iterations, per_projection, min_iterations, can_widen, min_corr;
  float max_change_norm, min_matching_ratio;
};
// state != nullptr: iterations, start estimate and output row come from the state (left by the search phase), which is advanced
int pslam_k_gn_iterate_dev(pslam_ctx* ctx, const pslam_linearize_cfg* cfg, int n_iters, double damping, const double* pose12,
                           const float* d_moving_xyz, const float* d_fixed_meas, int fixed_dim, const int* d_n_corr,
                           const int* d_corr_fixed, const int* d_corr_moving, const float* d_info_diag,
                           const pslam_pose_prior* prior, double* d_out, int* d_done, uint8_t* d_status,
                           PslamAlignState* d_state = nullptr, int max_iterations = 0);
int pslam_k_projective_align(pslam_ctx* ctx, int n_fixed, int n_moving, const pslam_projective_cfg* cfg, pslam_align* align,
                             int capacity, int* h_fixed, int* h_moving, float* h_dist, pslam_fused_gn* gn);

// k_linearize.cu  (T = float | double: scalar type of the clouds)
template <typename T>
int pslam_k_linearize_t(pslam_ctx* ctx, const pslam_linearize_cfg* cfg, const double* pose12, const T* d_moving_xyz,
                        const T* d_fixed_meas, int fixed_dim, int n_corr, const int* d_corr_fixed,
                        const int* d_corr_moving, const T* d_info_diag, const pslam_pose_prior* prior, uint8_t* d_status,
                        double* h_H36, double* h_b6, double* h_stats5);
template <typename T>
int pslam_k_gn_iterate_t(pslam_ctx* ctx, const pslam_linearize_cfg* cfg, int n_iters, double damping, double* pose12,
                         int n_moving, const T* h_moving_xyz, int n_fixed, const T* h_fixed_meas, int fixed_dim,
                         int n_corr, const int* h_corr_fixed, const int* h_corr_moving, const T* h_info_diag,
                         const pslam_pose_prior* prior, double* h_out16, uint8_t* h_status, int* h_iters_done, int* h_spd);
int pslam_k_gn_step(pslam_ctx* ctx, const double* H36, const double* b6, double damping,
                    double* pose12, double* dx6);

// k_scene.cu (N2: projective scene clipping over the whole local map)
int pslam_k_scene_clip(pslam_ctx* ctx, long long n, const float* d_xyz, const uint32_t* d_desc, const float* map_in_camera12,
                       const float* sensor_in_robot12, const float* K9, int rows, int cols, float range_min, float range_max,
                       unsigned long long* d_state, float* d_out_xyz, float* d_out_uvz, int* d_out_index,
                       uint32_t* d_out_desc, long long** d_n_out);
size_t pslam_k_scene_clip_state_bytes(long long n);

// k_mapping.cu (N3: per-landmark EKF update after the alignment)
int pslam_k_landmarks_ekf(pslam_ctx* ctx, const pslam_ekf_cfg* cfg, long long n, float* d_state_world, float* d_covariance,
                          const float* d_meas, float* d_local, uint8_t* d_inlier, int* d_n_inliers);
int pslam_k_landmarks_weighted_mean(pslam_ctx* ctx, const float* sensor_in_world12, const float* sensor_in_local_map12, float max_dist2,
                                    long long n, float* d_state_world, const int* d_n_opt, const float* d_landmark_in_sensor,
                                    float* d_local, uint8_t* d_inlier, int* d_n_inliers);

// k_merger.cu (N3: MergerProjective_::compute binning)
int pslam_k_merger_select_updates(pslam_ctx* ctx, const pslam_merger_cfg* cfg, const float* d_meas, int dim, int n_meas,
                                  const int* d_corr_moving, const float* d_corr_response, int n_corr, unsigned char* d_selected,
                                  unsigned* d_occupied, int n_words, int* d_result);
int pslam_k_merger_select_additions(pslam_ctx* ctx, const pslam_merger_cfg* cfg, const float* d_meas, int dim, int n_meas,
                                    const unsigned* d_occupied, int* d_winners, int* d_result);

// k_smoother.cu (N3: LandmarkEstimatorPoseBasedSmoother)
int pslam_k_landmarks_smoother(pslam_ctx* ctx, const pslam_smoother_cfg* cfg, const float* world_in_local_map12, int n,
                               const float* d_frames_siw, const float* d_frames_wis, const int* d_offsets, const int* d_hist_frame,
                               const float* d_hist_uv, const float* d_hist_pic, float* d_state_world, int* d_n_opt, float* d_local,
                               uint8_t* d_inlier, int* d_n_inliers);
