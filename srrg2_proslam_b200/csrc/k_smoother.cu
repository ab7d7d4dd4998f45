// k_smoother.cu -- N3: LandmarkEstimatorPoseBasedSmoother over the landmarks of one merger pass (sm_100a).
//
// Reference path replaced:
//   LandmarkEstimatorPoseBasedSmoother_::compute / _setMeanCoordinatesInWorld
//     .../mapping/landmarks/landmark_estimator_pose_based_smoother_impl.cpp:6-148
//   (kitti.conf's "landmark_estimator_smoother", the estimator of merger_triangulation)
// Per landmark: structure-only Gauss-Newton over the landmark's whole measurement history (3x3 normal equations,
// saturated kernel, full-pivot LU solve, chi2-delta convergence) or, below three measurements, the mean of the
// re-observed positions.  Landmarks are independent: one thread per landmark, fp32 like the reference.  This file is
// compiled with --fmad=false so that every product and sum rounds like the CPU restatement (bit-exact parity).
// Histories are CSR (offsets[n + 1]); a measurement = (frame index, (u, v), point in the camera frame); the frame tables
// hold one sensor pose per frame (what PointStatisticsField3D::CameraMeasurement keeps per measurement).
#include "pslam_internal.cuh"
#include "pslam_kernels.cuh"

namespace {

struct SmootherParams {
  float K[9];
  unsigned max_it, min_meas;
  float delta, max_reproj2, max_dist2;
  float wl[12];  // world_in_local_map
};

__device__ __forceinline__ void apply12(const float* __restrict__ T, const float* p, float* o) {
#pragma unroll
  for (int i = 0; i < 3; ++i) o[i] = ((T[4 * i] * p[0] + T[4 * i + 1] * p[1]) + T[4 * i + 2] * p[2]) + T[4 * i + 3];
}

// Eigen::FullPivLU<Matrix3f>::solve restated (column-major scan for the pivot, rank threshold eps * 3)
__device__ __forceinline__ void full_piv_lu_solve3(const float* A_in, const float* rhs, float* x) {
  float A[9], c[3] = {rhs[0], rhs[1], rhs[2]};
#pragma unroll
  for (int i = 0; i < 9; ++i) A[i] = A_in[i];
  int cp[3] = {0, 1, 2};
  float maxpivot = 0.f;
  int nonzero = 3;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    if (nonzero != 3) break;
    int br = k, bc = k;
    float best = -1.f;
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        if (i < k || j < k) continue;
        const float v = fabsf(A[3 * i + j]);
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
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const float t = A[3 * k + j];
        A[3 * k + j] = A[3 * br + j];
        A[3 * br + j] = t;
      }
      const float t = c[k];
      c[k] = c[br];
      c[br] = t;
    }
    if (bc != k) {
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const float t = A[3 * i + k];
        A[3 * i + k] = A[3 * i + bc];
        A[3 * i + bc] = t;
      }
      const int t = cp[k];
      cp[k] = cp[bc];
      cp[bc] = t;
    }
#pragma unroll
    for (int i = 0; i < 3; ++i)
      if (i > k) A[3 * i + k] /= A[3 * k + k];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j)
        if (i > k && j > k) A[3 * i + j] -= A[3 * i + k] * A[3 * k + j];
  }
  const float thr = maxpivot * (1.1920929e-7f * 3.0f);
  int rank = 0;
#pragma unroll
  for (int k = 0; k < 3; ++k)
    if (k < nonzero) rank += fabsf(A[3 * k + k]) > thr;
  c[1] -= A[3] * c[0];
  c[2] -= A[6] * c[0];
  c[2] -= A[7] * c[1];
  float y[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 2; i >= 0; --i) {
    if (i >= rank) continue;
    float acc = c[i];
#pragma unroll
    for (int j = 0; j < 3; ++j)
      if (j > i && j < rank) acc -= A[3 * i + j] * y[j];
    y[i] = acc / A[3 * i + i];
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) {  // x[cp[k]] = y[k] without dynamic register indexing
#pragma unroll
    for (int d = 0; d < 3; ++d)
      if (cp[k] == d) x[d] = y[k];
  }
}

__global__ void __launch_bounds__(128)
landmarks_smoother_kernel(const SmootherParams p, int n, const float* __restrict__ frames_siw, const float* __restrict__ frames_wis,
                          const int* __restrict__ offsets, const int* __restrict__ hist_frame, const float* __restrict__ hist_uv,
                          const float* __restrict__ hist_pic, float* __restrict__ state_world, int* __restrict__ n_opt,
                          float* __restrict__ coords_in_local_map, uint8_t* __restrict__ inlier, int* __restrict__ n_inliers) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  bool ok = false;
  if (i < n) {
    const int o = offsets[i], M = offsets[i + 1] - o;
    const float initial[3] = {state_world[3 * i], state_world[3 * i + 1], state_world[3 * i + 2]};
    float w[3] = {initial[0], initial[1], initial[2]};
    auto mean_in_world = [&](float* out) {  // _setMeanCoordinatesInWorld (:137-146)
      float acc[3] = {0.f, 0.f, 0.f};
      for (int k = 0; k < M; ++k) {
        const float pc[3] = {hist_pic[3 * (size_t) (o + k)], hist_pic[3 * (size_t) (o + k) + 1], hist_pic[3 * (size_t) (o + k) + 2]};
        float q[3];
        apply12(frames_siw + 12 * hist_frame[o + k], pc, q);
        acc[0] += q[0];
        acc[1] += q[1];
        acc[2] += q[2];
      }
      out[0] = acc[0] / (float) M;
      out[1] = acc[1] / (float) M;
      out[2] = acc[2] / (float) M;
    };
    float local[3];
    bool write_local = false, write_state = false;
    int new_nopt = n_opt[i];
    if ((unsigned) M < p.min_meas) {  // :29-43
      mean_in_world(w);
      const float d0 = w[0] - initial[0], d1 = w[1] - initial[1], d2 = w[2] - initial[2];
      if ((d0 * d0 + d1 * d1) + d2 * d2 < p.max_dist2) {
        write_local = write_state = ok = true;
        new_nopt = M;
      }
    } else {
      float prev = 0.f;
      int n_in = 0;
      for (unsigned it = 0; it < p.max_it; ++it) {  // :49-123
        float H[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, b[3] = {0.f, 0.f, 0.f}, total = 0.f;
        int n_out = 0;
        for (int k = 0; k < M; ++k) {
          const float* T = frames_wis + 12 * hist_frame[o + k];
          float cam[3];
          apply12(T, w, cam);
          if (cam[2] <= 0.f) {
            ++n_out;
            continue;
          }
          float ph[3];
#pragma unroll
          for (int a = 0; a < 3; ++a) ph[a] = (p.K[3 * a] * cam[0] + p.K[3 * a + 1] * cam[1]) + p.K[3 * a + 2] * cam[2];
          const float c = ph[2], inv_c = 1.0f / c, inv_c2 = inv_c * inv_c;
          const float e[3] = {ph[0] / c - hist_uv[2 * (size_t) (o + k)], ph[1] / c - hist_uv[2 * (size_t) (o + k) + 1],
                              c - hist_pic[3 * (size_t) (o + k) + 2]};
          float om[3] = {1.0f, 1.0f, 10.0f};  // :56-57
          const float e2 = (e[0] * om[0] * e[0] + e[1] * om[1] * e[1]) + e[2] * om[2] * e[2];
          total += e2;
          if (e2 > p.max_reproj2) {  // saturated kernel (:80-84)
            const float s = p.max_reproj2 / e2;
            om[0] *= s;
            om[1] *= s;
            om[2] *= s;
            ++n_out;
          }
          float Jl[9], J[9];
#pragma unroll
          for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int j = 0; j < 3; ++j) Jl[3 * a + j] = (p.K[3 * a] * T[j] + p.K[3 * a + 1] * T[4 + j]) + p.K[3 * a + 2] * T[8 + j];
          const float aa = -ph[0] * inv_c2, bb = -ph[1] * inv_c2;
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            J[j] = inv_c * Jl[j] + aa * Jl[6 + j];
            J[3 + j] = inv_c * Jl[3 + j] + bb * Jl[6 + j];
            J[6 + j] = Jl[6 + j];
          }
#pragma unroll
          for (int a = 0; a < 3; ++a) {
            const float t0 = J[a] * om[0], t1 = J[3 + a] * om[1], t2 = J[6 + a] * om[2];
#pragma unroll
            for (int j = 0; j < 3; ++j) H[3 * a + j] += (t0 * J[j] + t1 * J[3 + j]) + t2 * J[6 + j];
            b[a] += (t0 * e[0] + t1 * e[1]) + t2 * e[2];
          }
        }
        const float nb[3] = {-b[0], -b[1], -b[2]};
        float dx[3] = {0.f, 0.f, 0.f};
        full_piv_lu_solve3(H, nb, dx);
        w[0] += dx[0];
        w[1] += dx[1];
        w[2] += dx[2];
        n_in = M - n_out;
        if (fabsf(total - prev) < p.delta) break;
        prev = total;
      }
      if (n_in > new_nopt) {  // :126-131
        new_nopt += 1;
        ok = true;
      } else {  // :134-139
        mean_in_world(w);
      }
      write_local = write_state = true;
    }
    if (write_state) {
      state_world[3 * i] = w[0];
      state_world[3 * i + 1] = w[1];
      state_world[3 * i + 2] = w[2];
      n_opt[i] = new_nopt;
    }
    if (write_local) {
      apply12(p.wl, w, local);
      coords_in_local_map[3 * i] = local[0];
      coords_in_local_map[3 * i + 1] = local[1];
      coords_in_local_map[3 * i + 2] = local[2];
    }
    inlier[i] = ok ? 1 : 0;
  }
  const unsigned bal = __ballot_sync(0xffffffffu, ok);
  if ((threadIdx.x & 31) == 0 && bal) atomicAdd(n_inliers, __popc(bal));
}

}  // namespace

// all d_* are device pointers; frames_world_in_sensor / world_in_local_map are computed by the caller (host, fp32)
int pslam_k_landmarks_smoother(pslam_ctx* ctx, const pslam_smoother_cfg* cfg, const float* world_in_local_map12, int n,
                               const float* d_frames_siw, const float* d_frames_wis, const int* d_offsets, const int* d_hist_frame,
                               const float* d_hist_uv, const float* d_hist_pic, float* d_state_world, int* d_n_opt, float* d_local,
                               uint8_t* d_inlier, int* d_n_inliers) {
  SmootherParams p;
  for (int i = 0; i < 9; ++i) p.K[i] = cfg->K[i];
  p.max_it = cfg->maximum_number_of_iterations;
  p.min_meas = cfg->minimum_number_of_measurements_for_optimization;
  p.delta = cfg->convergence_criterion_minimum_chi2_delta;
  p.max_reproj2 = cfg->maximum_reprojection_error_pixels_squared;
  p.max_dist2 = cfg->maximum_distance_geometry_meters_squared;
  for (int i = 0; i < 12; ++i) p.wl[i] = world_in_local_map12[i];
  PSLAM_CUDA_TRY(ctx, cudaMemsetAsync(d_n_inliers, 0, sizeof(int), ctx->stream));
  if (n == 0) return PSLAM_OK;
  landmarks_smoother_kernel<<<(n + 127) / 128, 128, 0, ctx->stream>>>(p, n, d_frames_siw, d_frames_wis, d_offsets, d_hist_frame, d_hist_uv,
                                                                      d_hist_pic, d_state_world, d_n_opt, d_local, d_inlier, d_n_inliers);
  PSLAM_LAUNCH_CHECK(ctx, "landmarks_smoother_kernel");
  return PSLAM_OK;
}
