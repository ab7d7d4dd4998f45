// k_linearize.cu -- stages 3+4 (sm_100a): SE3 projective factor error + Jacobian per
// correspondence and the fp64 reduction of the 6x6 H / 6x1 b normal equations.
//
// Replaces srrg2_solver's FactorCorrespondenceDriven_::compute over
// SE3RectifiedStereoProjectiveErrorFactor / SE3ProjectiveDepthErrorFactor / SE3ProjectiveErrorFactor
// with RobustifierSaturated / RobustifierClamp (external code; wiring at
// .../registration/aligner_slice_processor_projective.cpp:27-112, use at tests/test_aligners.cpp:586-638,
// in-tree analogue of the accumulate loop: .../mapping/landmarks/landmark_estimator_pose_based_smoother_impl.cpp:48-106).
// The arithmetic is that of oracle/pslam_oracle_solver.hpp evaluated with fused multiply-adds (see error_and_jacobian); the
// parity contract with the fp64 oracle is 1e-9 relative on H, b (tests/test_gpu_solver.py).
//
// One thread per correspondence (grid-stride), 21 + 6 + 1 fp64 partial sums and 3 counters per
// thread, warp-shuffle tree, one partial row per CTA, a last single-CTA pass adds the rows in
// index order (deterministic -- no fp64 atomics).
#include "pslam_internal.cuh"
#include "pslam_kernels.cuh"

namespace {

constexpr int LZ_THREADS = 256;
constexpr int LZ_NACC = 32;  // 21 H (upper) + 6 b + chi + inliers + outliers + suppressed + pad

struct LinParams {
  int kind, robustifier;
  double K[9];
  double cols, rows;
  double baseline[3];
  double mean_disparity, chi_threshold;
  double R[9], t[3];
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// 1 / a for a normal, finite a: the hardware's 2^-23 seed + two Newton steps (four dependent fused multiply-adds, ~1 ulp) instead
// of the correctly rounded __ddiv_rn sequence, which sits ~260 cycles deep on every dependent chain of the solver iterations
__device__ __forceinline__ double rcp_fast(double a) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
  double e = fma(-a, r, 1.0);
  r = fma(r, e, r);
  e = fma(-a, r, 1.0);
  return fma(r, e, r);
}

// fp64 fused multiply-adds throughout: the fp64 pipe issues one warp instruction every two cycles, so the instruction count IS
// the cost (batched: issue-bound; per frame: the head of every solver iteration).  Against the oracle's separately rounded
// operations this differs in the last bits (relative 1e-15 on H, b; the parity contract is 1e-9).
__device__ __forceinline__ bool error_and_jacobian(const LinParams& c, const double* pm,
                                                   const double* z, double* e, double* J) {
  const double* R = c.R;
  double pc[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) pc[i] = fma(R[3 * i], pm[0], fma(R[3 * i + 1], pm[1], fma(R[3 * i + 2], pm[2], c.t[i])));
  if (pc[2] <= 0) return false;
  const double* K = c.K;
  const double hx = fma(K[0], pc[0], fma(K[1], pc[1], K[2] * pc[2]));
  const double hy = fma(K[3], pc[0], fma(K[4], pc[1], K[5] * pc[2]));
  const double hz = fma(K[6], pc[0], fma(K[7], pc[1], K[8] * pc[2]));
  const double iz = rcp_fast(hz);
  const double u = hx * iz, v = hy * iz;
  if (u < 0 || u > c.cols || v < 0 || v > c.rows) return false;
  double Jx[18];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    Jx[6 * i + 0] = R[3 * i + 0];
    Jx[6 * i + 1] = R[3 * i + 1];
    Jx[6 * i + 2] = R[3 * i + 2];
    Jx[6 * i + 3] = -2.0 * fma(R[3 * i + 1], pm[2], -(R[3 * i + 2] * pm[1]));
    Jx[6 * i + 4] = -2.0 * fma(R[3 * i + 2], pm[0], -(R[3 * i + 0] * pm[2]));
    Jx[6 * i + 5] = -2.0 * fma(R[3 * i + 0], pm[1], -(R[3 * i + 1] * pm[0]));
  }
  double KJ[18];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 6; ++j) KJ[6 * i + j] = fma(K[3 * i], Jx[j], fma(K[3 * i + 1], Jx[6 + j], K[3 * i + 2] * Jx[12 + j]));
  const double iz2 = iz * iz;
  const double hx_iz2 = hx * iz2, hy_iz2 = hy * iz2;
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    J[j] = fma(KJ[j], iz, -(hx_iz2 * KJ[12 + j]));
    J[6 + j] = fma(KJ[6 + j], iz, -(hy_iz2 * KJ[12 + j]));
  }
  e[0] = u - z[0];
  e[1] = v - z[1];
  if (c.kind == 0) {
    const double hxr = hx + c.baseline[0];
    e[2] = fma(hxr, iz, -z[2]);
    const double hxr_iz2 = hxr * iz2;
#pragma unroll
    for (int j = 0; j < 6; ++j) J[12 + j] = fma(KJ[j], iz, -(hxr_iz2 * KJ[12 + j]));
    if (c.mean_disparity > 0) {
      double w = fma(z[0] - z[2], rcp_fast(c.mean_disparity), 0.01);
      if (w > 1) w = 1;
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) J[6 * i + j] *= w;
    }
  } else if (c.kind == 1) {
    e[2] = pc[2] - z[2];
#pragma unroll
    for (int j = 0; j < 6; ++j) J[12 + j] = Jx[12 + j];
  } else {
    e[2] = 0;
#pragma unroll
    for (int j = 0; j < 6; ++j) J[12 + j] = 0;
  }
  return true;
}

// one correspondence as the factor reads it: moving point, fixed measurement, information diagonal -- widened to fp64 (exact)
struct CorrData {
  double pm[3], z[3], om[3];
};
template <typename T>
__device__ __forceinline__ CorrData load_correspondence(const T* __restrict__ moving_xyz, const T* __restrict__ fixed_meas,
                                                        int fixed_dim, int fi, int mi, const T* __restrict__ info_diag) {
  CorrData d;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    d.pm[i] = (double) moving_xyz[3 * (size_t) mi + i];
    d.z[i] = i < fixed_dim ? (double) fixed_meas[(size_t) fixed_dim * fi + i] : 0.0;
    d.om[i] = (double) info_diag[3 * (size_t) fi + i];
  }
  return d;
}

// accumulate one correspondence into the thread's partial sums (FactorCorrespondenceDriven_::compute analogue).
// status (may be nullptr): 0 inlier, 1 kernelized, 2 suppressed.
__device__ __forceinline__ void accumulate_loaded(const LinParams& c, int edim, const CorrData& d, double* acc, uint8_t* status) {
  double e[3], J[18];
  if (!error_and_jacobian(c, d.pm, d.z, e, J)) {
    acc[30] += 1;
    if (status) *status = 2;
    return;
  }
  const double* om = d.om;
  double chi = 0;
#pragma unroll
  for (int i = 0; i < 3; ++i)  // edim is 2 or 3: unrolled under a predicate so that e, J, om stay in registers
    if (i < edim) chi = fma(e[i] * om[i], e[i], chi);
  double scale = 1;
  if (c.robustifier != 0 && chi > c.chi_threshold) {
    acc[29] += 1;
    scale = (c.robustifier == 1) ? c.chi_threshold * rcp_fast(chi) : 0.0;
    if (status) *status = 1;
  } else {
    acc[28] += 1;
    if (status) *status = 0;
  }
  acc[27] = fma(chi, scale, acc[27]);
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    if (i >= edim) break;
    const double w = om[i] * scale;
    int h = 0;
#pragma unroll
    for (int a = 0; a < 6; ++a) {
      const double Jw = J[6 * i + a] * w;
      acc[21 + a] = fma(Jw, e[i], acc[21 + a]);
#pragma unroll
      for (int b = a; b < 6; ++b, ++h) acc[h] = fma(Jw, J[6 * i + b], acc[h]);
    }
  }
}

// T = scalar type of the clouds in HBM: float (the reference's own cloud type: 40 B per correspondence, widened to
// fp64 in registers -- exact) or double.
template <typename T>
__device__ __forceinline__ void accumulate_correspondence(const LinParams& c, int edim, const T* __restrict__ moving_xyz,
                                                          const T* __restrict__ fixed_meas, int fixed_dim, int fi, int mi,
                                                          const T* __restrict__ info_diag, double* acc, uint8_t* status) {
  const CorrData d = load_correspondence(moving_xyz, fixed_meas, fixed_dim, fi, mi, info_diag);
  accumulate_loaded(c, edim, d, acc, status);
}

template <typename T>
__global__ void __launch_bounds__(LZ_THREADS, 2)
linearize_kernel(LinParams c, const T* __restrict__ moving_xyz,
                 const T* __restrict__ fixed_meas, int fixed_dim, int n_corr,
                 const int* __restrict__ corr_fixed, const int* __restrict__ corr_moving,
                 const T* __restrict__ info_diag, double* __restrict__ partials, uint8_t* __restrict__ status) {
  __shared__ double s_part[LZ_THREADS / 32][LZ_NACC];
  double acc[LZ_NACC];
#pragma unroll
  for (int i = 0; i < LZ_NACC; ++i) acc[i] = 0;
  const int edim = (c.kind == 2) ? 2 : 3;
  for (int k = blockIdx.x * LZ_THREADS + threadIdx.x; k < n_corr; k += gridDim.x * LZ_THREADS)
    accumulate_correspondence(c, edim, moving_xyz, fixed_meas, fixed_dim, corr_fixed[k], corr_moving[k], info_diag, acc,
                              status ? status + k : nullptr);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < LZ_NACC - 1; ++i) {
    const double s = warp_sum(acc[i]);
    if (lane == 0) s_part[wid][i] = s;
  }
  __syncthreads();
  if (threadIdx.x < LZ_NACC - 1) {
    double s = 0;
    for (int w = 0; w < LZ_THREADS / 32; ++w) s += s_part[w][threadIdx.x];
    partials[(size_t) blockIdx.x * LZ_NACC + threadIdx.x] = s;
  }
}

// out[0..35] H (full symmetric), [36..41] b, [42] chi, [43] inliers, [44] outliers, [45] suppressed
__global__ void linearize_finish_kernel(const double* __restrict__ partials, int n_blocks,
                                        double* __restrict__ out) {
  __shared__ double s[LZ_NACC];
  const int i = threadIdx.x;
  if (i < LZ_NACC - 1) {
    double v = 0;
    for (int b = 0; b < n_blocks; ++b) v += partials[(size_t) b * LZ_NACC + i];
    s[i] = v;
  }
  __syncthreads();
  if (i == 0) {
    int h = 0;
    for (int a = 0; a < 6; ++a)
      for (int b = a; b < 6; ++b) {
        out[6 * a + b] = s[h];
        out[6 * b + a] = s[h];
        ++h;
      }
    for (int a = 0; a < 6; ++a) out[36 + a] = s[21 + a];
    out[42] = s[27];
    out[43] = s[28];
    out[44] = s[29];
    out[45] = s[30];
  }
}

// (H + damping I) dx = -b by Cholesky; pose <- pose * v2t(dx)   (IterationAlgorithmGN, one 6x6 block).
// H: full symmetric 6x6.  R (row-major 3x3), t: updated in place.  Returns false when H + damping I is not SPD.
__device__ __forceinline__ bool gn_solve_update(const double* H, const double* bvec, double damping, double* R, double* t,
                                                double* dx) {
  double A[36], L[36];
  for (int i = 0; i < 36; ++i) {
    A[i] = H[i];
    L[i] = 0;
  }
  for (int i = 0; i < 6; ++i) A[7 * i] += damping;
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j <= i; ++j) {
      double s = A[6 * i + j];
      for (int k = 0; k < j; ++k) s = __dsub_rn(s, __dmul_rn(L[6 * i + k], L[6 * j + k]));
      if (i == j) {
        if (!(s > 0)) return false;
        L[6 * i + i] = sqrt(s);
      } else {
        L[6 * i + j] = __ddiv_rn(s, L[6 * j + j]);
      }
    }
  double y[6];
  for (int i = 0; i < 6; ++i) {
    double s = -bvec[i];
    for (int k = 0; k < i; ++k) s = __dsub_rn(s, __dmul_rn(L[6 * i + k], y[k]));
    y[i] = __ddiv_rn(s, L[6 * i + i]);
  }
  for (int i = 5; i >= 0; --i) {
    double s = y[i];
    for (int k = i + 1; k < 6; ++k) s = __dsub_rn(s, __dmul_rn(L[6 * k + i], dx[k]));
    dx[i] = __ddiv_rn(s, L[6 * i + i]);
  }
  // v2t(dx)
  double x = dx[3], yq = dx[4], z = dx[5];
  const double n2 = __dadd_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(yq, yq)), __dmul_rn(z, z));
  double w;
  if (n2 < 1.0) {
    w = sqrt(__dsub_rn(1.0, n2));
  } else {
    const double n = sqrt(n2);
    x = __ddiv_rn(x, n);
    yq = __ddiv_rn(yq, n);
    z = __ddiv_rn(z, n);
    w = 0;
  }
  double D[9];
  D[0] = __dsub_rn(1.0, __dmul_rn(2.0, __dadd_rn(__dmul_rn(yq, yq), __dmul_rn(z, z))));
  D[1] = __dmul_rn(2.0, __dsub_rn(__dmul_rn(x, yq), __dmul_rn(z, w)));
  D[2] = __dmul_rn(2.0, __dadd_rn(__dmul_rn(x, z), __dmul_rn(yq, w)));
  D[3] = __dmul_rn(2.0, __dadd_rn(__dmul_rn(x, yq), __dmul_rn(z, w)));
  D[4] = __dsub_rn(1.0, __dmul_rn(2.0, __dadd_rn(__dmul_rn(x, x), __dmul_rn(z, z))));
  D[5] = __dmul_rn(2.0, __dsub_rn(__dmul_rn(yq, z), __dmul_rn(x, w)));
  D[6] = __dmul_rn(2.0, __dsub_rn(__dmul_rn(x, z), __dmul_rn(yq, w)));
  D[7] = __dmul_rn(2.0, __dadd_rn(__dmul_rn(yq, z), __dmul_rn(x, w)));
  D[8] = __dsub_rn(1.0, __dmul_rn(2.0, __dadd_rn(__dmul_rn(x, x), __dmul_rn(yq, yq))));
  double Rn[9], tn[3];
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j)
      Rn[3 * i + j] = __dadd_rn(__dadd_rn(__dmul_rn(R[3 * i], D[j]), __dmul_rn(R[3 * i + 1], D[3 + j])),
                                __dmul_rn(R[3 * i + 2], D[6 + j]));
    tn[i] = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(R[3 * i], dx[0]), __dmul_rn(R[3 * i + 1], dx[1])),
                                __dmul_rn(R[3 * i + 2], dx[2])), t[i]);
  }
  for (int i = 0; i < 9; ++i) R[i] = Rn[i];
  for (int i = 0; i < 3; ++i) t[i] = tn[i];
  return true;
}

// io: [0..35] H, [36..41] b, [42] damping, [43..54] pose12 in/out, [55..60] dx out, [61] status
__global__ void gn_step_kernel(double* __restrict__ io) {
  if (threadIdx.x != 0) return;
  double R[9], t[3], dx[6];
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) R[3 * i + j] = io[43 + 4 * i + j];
    t[i] = io[43 + 4 * i + 3];
  }
  if (!gn_solve_update(io, io + 36, io[42], R, t, dx)) {
    io[61] = -1;
    return;
  }
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) io[43 + 4 * i + j] = R[3 * i + j];
    io[43 + 4 * i + 3] = t[i];
  }
  for (int i = 0; i < 6; ++i) io[55 + i] = dx[i];
  io[61] = 0;
}

// ---- SE3 pose-prior factor (the shipped aligners' second slice, AlignerSliceMotionModel3D: configurations/kitti.conf:747-772,
// icl.conf:268-293; srrg2_slam_interfaces / srrg2_solver, restated in oracle/pslam_oracle_solver.hpp::pose_prior_accumulate):
// e = t2tnq(Z^-1 X), d e_t / d dt = R_E, d e_q / d dq = w I + [v]x, constant information; summed into H, b before the solve.
struct PriorParams {
  int enabled;
  double Zi_R[9], Zi_t[3];  // Z^-1
  double Omega[36];
};

__device__ __forceinline__ void t2tnq_dev(const double* R, const double* t, double* v6) {
  v6[0] = t[0];
  v6[1] = t[1];
  v6[2] = t[2];
  double w, x, y, z;
  const double tr = R[0] + R[4] + R[8];
  if (tr > 0) {
    const double s = sqrt(tr + 1.0) * 2;
    w = 0.25 * s;
    x = (R[7] - R[5]) / s;
    y = (R[2] - R[6]) / s;
    z = (R[3] - R[1]) / s;
  } else if (R[0] > R[4] && R[0] > R[8]) {
    const double s = sqrt(1.0 + R[0] - R[4] - R[8]) * 2;
    w = (R[7] - R[5]) / s;
    x = 0.25 * s;
    y = (R[1] + R[3]) / s;
    z = (R[2] + R[6]) / s;
  } else if (R[4] > R[8]) {
    const double s = sqrt(1.0 + R[4] - R[0] - R[8]) * 2;
    w = (R[2] - R[6]) / s;
    x = (R[1] + R[3]) / s;
    y = 0.25 * s;
    z = (R[5] + R[7]) / s;
  } else {
    const double s = sqrt(1.0 + R[8] - R[0] - R[4]) * 2;
    w = (R[3] - R[1]) / s;
    x = (R[2] + R[6]) / s;
    y = (R[5] + R[7]) / s;
    z = 0.25 * s;
  }
  const double n = sqrt(w * w + x * x + y * y + z * z);
  const double sgn = (w < 0) ? -1.0 : 1.0;
  v6[3] = sgn * x / n;
  v6[4] = sgn * y / n;
  v6[5] = sgn * z / n;
}

// H (full 6x6), b += prior; returns chi.  One thread.
__device__ __noinline__ double pose_prior_accumulate_dev(const PriorParams& pr, const double* R, const double* t, double* H,
                                                         double* bvec) {
  double ER[9], Et[3];
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j)
      ER[3 * i + j] = (pr.Zi_R[3 * i] * R[j] + pr.Zi_R[3 * i + 1] * R[3 + j]) + pr.Zi_R[3 * i + 2] * R[6 + j];
    Et[i] = ((pr.Zi_R[3 * i] * t[0] + pr.Zi_R[3 * i + 1] * t[1]) + pr.Zi_R[3 * i + 2] * t[2]) + pr.Zi_t[i];
  }
  double e[6];
  t2tnq_dev(ER, Et, e);
  const double n2 = e[3] * e[3] + e[4] * e[4] + e[5] * e[5];
  const double w = sqrt(n2 < 1.0 ? 1.0 - n2 : 0.0);
  double J[36];
  for (int i = 0; i < 36; ++i) J[i] = 0;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) J[6 * i + j] = ER[3 * i + j];
  const double vx = e[3], vy = e[4], vz = e[5];
  const double Q[9] = {w, -vz, vy, vz, w, -vx, -vy, vx, w};
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) J[6 * (3 + i) + 3 + j] = Q[3 * i + j];
  double OJ[36], Oe[6];
  for (int i = 0; i < 6; ++i) {
    double sum = 0;
    for (int k = 0; k < 6; ++k) sum += pr.Omega[6 * i + k] * e[k];
    Oe[i] = sum;
    for (int j = 0; j < 6; ++j) {
      double a = 0;
      for (int k = 0; k < 6; ++k) a += pr.Omega[6 * i + k] * J[6 * k + j];
      OJ[6 * i + j] = a;
    }
  }
  double chi = 0;
  for (int i = 0; i < 6; ++i) chi += e[i] * Oe[i];
  for (int a = 0; a < 6; ++a) {
    double bs = 0;
    for (int k = 0; k < 6; ++k) bs += J[6 * k + a] * Oe[k];
    bvec[a] += bs;
    for (int c = 0; c < 6; ++c) {
      double hs = 0;
      for (int k = 0; k < 6; ++k) hs += J[6 * k + a] * OJ[6 * k + c];
      H[6 * a + c] += hs;
    }
  }
  return chi;
}

// ---- the same two steps spread over the lanes of ONE warp ------------------------------------------------------------------
// Inside the fused loop the pose-prior factor and the 6x6 solve were ~6 of the ~9 us of an iteration when one thread ran them
// (dependent fp64 chains, arrays in local memory).  The prior runs in a warp of its own next to the linearisation, one
// matrix element per lane, intermediate matrices in shared memory; the solve runs in registers (gn_solve_update_warp).
struct GnWork {
  double R[9], t[3];
  double Om[36];                               // the prior's information matrix (kernel parameters indexed per lane would serialise)
  double ER[9], Et[3], e6[6], Q[9], OJ[36], Oe[6];  // pose-prior scratch (its own warp)
  double Hp[36], bp[6];                        // the prior's contribution to H, b for the current pose
  double Rp[9], tp[3];                         // the pose before the last update (= after the iteration before the last)
};

// t2tnq with reciprocal square roots in place of the sqrt + divide pairs (the prior sits on the iteration's critical path
// next to the linearisation: two fp64 divide sequences less).  Last-bit differences to t2tnq_dev; parity contract 1e-9.
__device__ __forceinline__ void t2tnq_fast(const double* R, const double* t, double* v6) {
  v6[0] = t[0];
  v6[1] = t[1];
  v6[2] = t[2];
  double w, x, y, z;
  const double tr = R[0] + R[4] + R[8];
  if (tr > 0) {
    const double a = tr + 1.0, r = rsqrt(a), h = 0.5 * r;  // s = 2 sqrt(a): 0.25 s = 0.5 a r, 1 / s = 0.5 r
    w = 0.5 * a * r;
    x = (R[7] - R[5]) * h;
    y = (R[2] - R[6]) * h;
    z = (R[3] - R[1]) * h;
  } else if (R[0] > R[4] && R[0] > R[8]) {
    const double a = 1.0 + R[0] - R[4] - R[8], r = rsqrt(a), h = 0.5 * r;
    w = (R[7] - R[5]) * h;
    x = 0.5 * a * r;
    y = (R[1] + R[3]) * h;
    z = (R[2] + R[6]) * h;
  } else if (R[4] > R[8]) {
    const double a = 1.0 + R[4] - R[0] - R[8], r = rsqrt(a), h = 0.5 * r;
    w = (R[2] - R[6]) * h;
    x = (R[1] + R[3]) * h;
    y = 0.5 * a * r;
    z = (R[5] + R[7]) * h;
  } else {
    const double a = 1.0 + R[8] - R[0] - R[4], r = rsqrt(a), h = 0.5 * r;
    w = (R[3] - R[1]) * h;
    x = (R[2] + R[6]) * h;
    y = (R[5] + R[7]) * h;
    z = 0.5 * a * r;
  }
  const double rn = rsqrt(w * w + x * x + y * y + z * z);
  const double sgn = (w < 0) ? -rn : rn;
  v6[3] = sgn * x;
  v6[4] = sgn * y;
  v6[5] = sgn * z;
}

// writes the prior's terms for the pose S.R, S.t into S.Hp, S.bp (the caller adds them to H, b: the same two additions
// pose_prior_accumulate_dev makes)
__device__ __forceinline__ void pose_prior_accumulate_warp(const PriorParams& pr, GnWork& S, int lane) {
  if (lane < 9) {
    const int i = lane / 3, j = lane % 3;
    S.ER[lane] = (pr.Zi_R[3 * i] * S.R[j] + pr.Zi_R[3 * i + 1] * S.R[3 + j]) + pr.Zi_R[3 * i + 2] * S.R[6 + j];
  } else if (lane < 12) {
    const int i = lane - 9;
    S.Et[i] = ((pr.Zi_R[3 * i] * S.t[0] + pr.Zi_R[3 * i + 1] * S.t[1]) + pr.Zi_R[3 * i + 2] * S.t[2]) + pr.Zi_t[i];
  }
  __syncwarp();
  {  // e = t2tnq(E), Q = w I + [v]x: every lane evaluates it (no exchange), lane 0 publishes
    double e6[6];
    t2tnq_fast(S.ER, S.Et, e6);
    const double vx = e6[3], vy = e6[4], vz = e6[5];
    const double a = 1.0 - (vx * vx + vy * vy + vz * vz);
    const double w = a > 0 ? a * rsqrt(a) : 0.0;
    if (lane == 0) {
#pragma unroll
      for (int i = 0; i < 6; ++i) S.e6[i] = e6[i];
      S.Q[0] = w, S.Q[1] = -vz, S.Q[2] = vy;
      S.Q[3] = vz, S.Q[4] = w, S.Q[5] = -vx;
      S.Q[6] = -vy, S.Q[7] = vx, S.Q[8] = w;
    }
  }
  __syncwarp();
  // J = blockdiag(R_E, Q): Omega J, J^T (Omega J) and J^T (Omega e) only touch the non-zero 3x3 blocks (three terms per element;
  // the dropped terms are exact zeros, the sums keep their order)
  for (int el = lane; el < 36; el += 32) {
    const int i = el / 6, j = el % 6;
    const double* B = j < 3 ? S.ER + j : S.Q + (j - 3);
    const double* O = S.Om + 6 * i + (j < 3 ? 0 : 3);
    S.OJ[el] = (O[0] * B[0] + O[1] * B[3]) + O[2] * B[6];
  }
  if (lane < 6) {
    double sum = 0;
#pragma unroll
    for (int k = 0; k < 6; ++k) sum += S.Om[6 * lane + k] * S.e6[k];
    S.Oe[lane] = sum;
  }
  __syncwarp();
  if (lane < 6) {
    const double* B = lane < 3 ? S.ER + lane : S.Q + (lane - 3);
    const double* v = S.Oe + (lane < 3 ? 0 : 3);
    S.bp[lane] = (B[0] * v[0] + B[3] * v[1]) + B[6] * v[2];
  }
  for (int el = lane; el < 36; el += 32) {
    const int a = el / 6, c = el % 6;
    const double* B = a < 3 ? S.ER + a : S.Q + (a - 3);
    const double* M = S.OJ + c + (a < 3 ? 0 : 18);
    S.Hp[el] = (B[0] * M[0] + B[3] * M[6]) + B[6] * M[12];
  }
  __syncwarp();
}

// (H + damping I) dx = -b by 3x3 block elimination with closed-form (adjugate) inverses of the translation block A and of its
// Schur complement S = C - B^T A^-1 B.  Same solution as the Cholesky of gn_solve_update (H is symmetric positive definite
// and well conditioned under the damping the shipped solvers use; differences ~1e-12 relative on dx), but the dependent chain
// holds TWO reciprocals instead of six reciprocal square roots + a backward substitution: every product is independent of its
// neighbours, one thread pipelines them.  Positive definiteness = Sylvester's criterion on A and S.  Every lane computes the
// same thing (no exchange); returns false when H + damping I is not positive definite.
__device__ __forceinline__ bool gn_solve6_block(const double* U, const double* bvec, double lam, double* dx) {
  // U = upper triangle, row by row: (0,0..5) 0..5, (1,1..5) 6..10, (2,2..5) 11..14, (3,3..5) 15..17, (4,4..5) 18..19, (5,5) 20
  const double a00 = U[0] + lam, a01 = U[1], a02 = U[2], a11 = U[6] + lam, a12 = U[7], a22 = U[11] + lam;
  const double b00 = U[3], b01 = U[4], b02 = U[5], b10 = U[8], b11 = U[9], b12 = U[10], b20 = U[12], b21 = U[13], b22 = U[14];
  const double c00 = U[15] + lam, c01 = U[16], c02 = U[17], c11 = U[18] + lam, c12 = U[19], c22 = U[20] + lam;
  const double g0 = -bvec[0], g1 = -bvec[1], g2 = -bvec[2], g3 = -bvec[3], g4 = -bvec[4], g5 = -bvec[5];
  // adj(A) (symmetric) and det(A)
  const double A00 = fma(a11, a22, -(a12 * a12)), A01 = fma(a02, a12, -(a01 * a22)), A02 = fma(a01, a12, -(a02 * a11));
  const double A11 = fma(a00, a22, -(a02 * a02)), A12 = fma(a01, a02, -(a00 * a12)), A22 = fma(a00, a11, -(a01 * a01));
  const double detA = fma(a00, A00, fma(a01, A01, a02 * A02));
  if (!(a00 > 0) || !(A22 > 0) || !(detA > 0)) return false;
  const double rd = rcp_fast(detA);
  // Y' = adj(A) B, y' = adj(A) g1 (the 1 / det(A) is applied where they are used)
  const double Y00 = fma(A00, b00, fma(A01, b10, A02 * b20)), Y01 = fma(A00, b01, fma(A01, b11, A02 * b21)),
               Y02 = fma(A00, b02, fma(A01, b12, A02 * b22));
  const double Y10 = fma(A01, b00, fma(A11, b10, A12 * b20)), Y11 = fma(A01, b01, fma(A11, b11, A12 * b21)),
               Y12 = fma(A01, b02, fma(A11, b12, A12 * b22));
  const double Y20 = fma(A02, b00, fma(A12, b10, A22 * b20)), Y21 = fma(A02, b01, fma(A12, b11, A22 * b21)),
               Y22 = fma(A02, b02, fma(A12, b12, A22 * b22));
  const double y0 = fma(A00, g0, fma(A01, g1, A02 * g2)), y1 = fma(A01, g0, fma(A11, g1, A12 * g2)),
               y2 = fma(A02, g0, fma(A12, g1, A22 * g2));
  // S = C - B^T Y' / det(A) (upper triangle), h = g2 - B^T y' / det(A)
  const double s00 = fma(-rd, fma(b00, Y00, fma(b10, Y10, b20 * Y20)), c00), s01 = fma(-rd, fma(b00, Y01, fma(b10, Y11, b20 * Y21)), c01),
               s02 = fma(-rd, fma(b00, Y02, fma(b10, Y12, b20 * Y22)), c02), s11 = fma(-rd, fma(b01, Y01, fma(b11, Y11, b21 * Y21)), c11),
               s12 = fma(-rd, fma(b01, Y02, fma(b11, Y12, b21 * Y22)), c12), s22 = fma(-rd, fma(b02, Y02, fma(b12, Y12, b22 * Y22)), c22);
  const double h0 = fma(-rd, fma(b00, y0, fma(b10, y1, b20 * y2)), g3), h1 = fma(-rd, fma(b01, y0, fma(b11, y1, b21 * y2)), g4),
               h2 = fma(-rd, fma(b02, y0, fma(b12, y1, b22 * y2)), g5);
  const double S00 = fma(s11, s22, -(s12 * s12)), S01 = fma(s02, s12, -(s01 * s22)), S02 = fma(s01, s12, -(s02 * s11));
  const double S11 = fma(s00, s22, -(s02 * s02)), S12 = fma(s01, s02, -(s00 * s12)), S22 = fma(s00, s11, -(s01 * s01));
  const double detS = fma(s00, S00, fma(s01, S01, s02 * S02));
  if (!(s00 > 0) || !(S22 > 0) || !(detS > 0)) return false;
  const double rs = rcp_fast(detS);
  dx[3] = rs * fma(S00, h0, fma(S01, h1, S02 * h2));
  dx[4] = rs * fma(S01, h0, fma(S11, h1, S12 * h2));
  dx[5] = rs * fma(S02, h0, fma(S12, h1, S22 * h2));
  dx[0] = rd * (y0 - fma(Y00, dx[3], fma(Y01, dx[4], Y02 * dx[5])));
  dx[1] = rd * (y1 - fma(Y10, dx[3], fma(Y11, dx[4], Y12 * dx[5])));
  dx[2] = rd * (y2 - fma(Y20, dx[3], fma(Y21, dx[4], Y22 * dx[5])));
  return true;
}

// (H + damping I) dx = -b, pose <- pose * v2t(dx).  sums = the 21 upper-triangle sums of H followed by the 6 of b (the
// reduction's output, read in place: no 6x6 staging copy); the prior's S.Hp, S.bp are added when with_prior.  S.R, S.t in / out.
// All 32 lanes call it: every lane solves the (tiny) system and forms the update rotation in its own registers -- nothing
// is exchanged until lanes 0..11 each write one element of the new pose.
// row (shared memory, 12 doubles): the pose after this iteration in 3x4 row-major order (the unchanged pose when the solve
// fails); S.Rp, S.tp <- the pose before the update.  Both straight from the lanes' registers.
__device__ __forceinline__ bool gn_solve_update_warp(GnWork& S, const double* sums, bool with_prior, double damping, int lane,
                                                     double* row) {
  const int pos = lane < 9 ? 4 * (lane / 3) + lane % 3 : 4 * (lane - 9) + 3;  // lane's element of the pose in the row
  double old = 0;
  if (lane < 9) old = S.R[lane];
  else if (lane < 12) old = S.t[lane - 9];
  double U[21], g[6], dx[6];
  {
    int h = 0;
#pragma unroll
    for (int a = 0; a < 6; ++a)
#pragma unroll
      for (int b = a; b < 6; ++b, ++h) U[h] = with_prior ? sums[h] + S.Hp[6 * a + b] : sums[h];
#pragma unroll
    for (int a = 0; a < 6; ++a) g[a] = with_prior ? sums[21 + a] + S.bp[a] : sums[21 + a];
  }
  if (!gn_solve6_block(U, g, damping, dx)) {  // uniform: every lane evaluates the same values
    if (lane < 12) row[pos] = old;
    return false;
  }
  // v2t(dx): unit quaternion (x, y, z, w) from its imaginary part, normalised when longer than 1
  double x = dx[3], y = dx[4], z = dx[5];
  const double n2 = fma(x, x, fma(y, y, z * z));
  double w = 0;
  if (n2 < 1.0) {
    const double a = 1.0 - n2;
    w = a * rsqrt(a);
  } else {
    const double rn = rsqrt(n2);
    x *= rn;
    y *= rn;
    z *= rn;
  }
  const double xx = x * x, yy = y * y, zz = z * z, xy = x * y, xz = x * z, yz = y * z, xw = x * w, yw = y * w, zw = z * w;
  const double D0 = 1.0 - 2.0 * (yy + zz), D1 = 2.0 * (xy - zw), D2 = 2.0 * (xz + yw);
  const double D3 = 2.0 * (xy + zw), D4 = 1.0 - 2.0 * (xx + zz), D5 = 2.0 * (yz - xw);
  const double D6 = 2.0 * (xz - yw), D7 = 2.0 * (yz + xw), D8 = 1.0 - 2.0 * (xx + yy);
  double nv = 0;
  if (lane < 9) {  // R <- R D
    const int i = lane / 3, j = lane - 3 * i;
    const double d0 = j == 0 ? D0 : (j == 1 ? D1 : D2), d1 = j == 0 ? D3 : (j == 1 ? D4 : D5), d2 = j == 0 ? D6 : (j == 1 ? D7 : D8);
    nv = fma(S.R[3 * i], d0, fma(S.R[3 * i + 1], d1, S.R[3 * i + 2] * d2));
  } else if (lane < 12) {  // t <- R dt + t
    const int i = lane - 9;
    nv = fma(S.R[3 * i], dx[0], fma(S.R[3 * i + 1], dx[1], fma(S.R[3 * i + 2], dx[2], S.t[i])));
  }
  __syncwarp();
  if (lane < 9) S.R[lane] = nv, S.Rp[lane] = old;
  else if (lane < 12) S.t[lane - 9] = nv, S.tp[lane - 9] = old;
  if (lane < 12) row[pos] = nv;
  __syncwarp();
  return true;
}

// ---- fused solver iterations: n_iters x { linearise all correspondences -> H, b (+ pose prior) -> Cholesky -> pose update }
// in ONE launch, one CTA per problem.  This is what the aligner runs between two re-projections of the correspondence
// finder (the correspondences, hence the information matrices, do not change in between: SURVEY App. E.6), instead
// of 2 launches + 2 host round trips per iteration.  out (per iteration): 12 pose (after the update) + chi, inliers,
// outliers, suppressed.  iters_done: iterations completed (stops early when H + damping I is not SPD).  status (may be
// nullptr): per-correspondence factor status of the LAST linearised iteration (inlier-only runs, icl.conf:55-58).
// Latency is what counts here (a frame runs ~100 dependent iterations on a few dozen correspondences): a thread keeps its
// correspondence in registers across the iterations when there is at most one per thread, warps without correspondences
// skip the shuffle reduction, and warp 0 runs the prior + solve lane-parallel (gn_solve_update_warp).
constexpr int GN_OUT = 16;
template <typename T>
__global__ void __launch_bounds__(LZ_THREADS)
gn_iterate_kernel(LinParams c0, PriorParams prior, double damping, int n_iters, const T* __restrict__ moving_xyz,
                  const T* __restrict__ fixed_meas, int fixed_dim, int n_corr, const int* __restrict__ corr_fixed,
                  const int* __restrict__ corr_moving, const T* __restrict__ info_diag, double* __restrict__ out,
                  int* __restrict__ iters_done, uint8_t* __restrict__ status, const int* __restrict__ n_corr_dev,
                  PslamAlignState* __restrict__ state, int max_iterations) {
  pslam_pdl_enter();
  if (n_corr_dev) n_corr = *n_corr_dev;  // correspondences produced on the device by the launch before (pslam_projective_match_gn)
  if (state) {  // one phase of pslam_projective_align: budget, start estimate and output rows are the state's (uniform reads;
                // thread 0 advances the state after the last barrier of the kernel)
    if (state->stop) return;
    n_iters = state->n_fused;
    out += (size_t) state->it * GN_OUT;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
      for (int j = 0; j < 3; ++j) c0.R[3 * i + j] = state->estimate[4 * i + j];
      c0.t[i] = state->estimate[4 * i + 3];
    }
  }
  __shared__ double s_part[LZ_THREADS / 32][LZ_NACC];
  __shared__ double s_sum[LZ_NACC];
  __shared__ GnWork S;
  __shared__ int s_ok;
  // small problems (the per-frame case): the 31 partial sums of every thread go through shared memory, eight threads per PAIR
  // of sums add them up (instead of a 5-level fp64 shuffle butterfly over 31 values plus the cross-warp pass: ~1.9 k cycles).
  // 128-bit stores: the barrier that follows drains the stores in flight at tens of cycles apiece, so their NUMBER is the cost.
  constexpr int RED_CAP = 96, RED_PITCH = RED_CAP + 1;
  __shared__ double2 s_acc[(LZ_NACC / 2) * RED_PITCH];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  // kernel parameters are copied with compile-time indices, one thread per array (a per-thread index into the parameter
  // space serialises the constant loads and makes the compiler spill the struct to local memory first)
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < 9; ++i) S.R[i] = c0.R[i];
#pragma unroll
    for (int i = 0; i < 3; ++i) S.t[i] = c0.t[i];
    s_ok = 1;
  }
  if (prior.enabled && threadIdx.x == 64) {
#pragma unroll
    for (int i = 0; i < 36; ++i) S.Om[i] = prior.Omega[i];
  }
  __syncthreads();
  const int edim = (c0.kind == 2) ? 2 : 3;
  LinParams c = c0;
  const bool resident = n_corr <= LZ_THREADS;  // at most one correspondence per thread: it stays in registers
  const bool warp_has_work = wid * 32 < n_corr;
  const bool small = n_corr <= RED_CAP;
  CorrData mine;
  uint8_t my_status = 255;  // resident correspondences: the status stays in a register, ONE global store after the loop
  // The barriers below order the block's global stores as well: every store in flight at a barrier is waited for (an HBM
  // round trip per iteration when the loop stored its outputs as it went).  The per-iteration rows therefore collect in
  // shared memory and leave in bursts of OUT_RING iterations.
  constexpr int OUT_RING = 32;
  __shared__ double s_out[OUT_RING * GN_OUT];
  if (resident && (int) threadIdx.x < n_corr)
    mine = load_correspondence(moving_xyz, fixed_meas, fixed_dim, corr_fixed[threadIdx.x], corr_moving[threadIdx.x], info_diag);
  int done = 0;
  for (int it = 0; it < n_iters; ++it) {
#pragma unroll
    for (int i = 0; i < 9; ++i) c.R[i] = S.R[i];
#pragma unroll
    for (int i = 0; i < 3; ++i) c.t[i] = S.t[i];
    // the prior only depends on the pose: the last warp evaluates it while the others linearise the correspondences
    if (prior.enabled && wid == LZ_THREADS / 32 - 1) pose_prior_accumulate_warp(prior, S, lane);
    if (small) {
      if ((int) threadIdx.x < n_corr) {
        double acc[LZ_NACC];
#pragma unroll
        for (int i = 0; i < LZ_NACC; ++i) acc[i] = 0;
        accumulate_loaded(c, edim, mine, acc, &my_status);
#pragma unroll
        for (int i = 0; i < LZ_NACC / 2; ++i) s_acc[i * RED_PITCH + threadIdx.x] = make_double2(acc[2 * i], acc[2 * i + 1]);
      }
      if (wid < 4) {  // correspondences (<= 96) and the 31 x 4 adders live in warps 0..3: a barrier of their own, the prior
                      // warp is only waited for at the block barrier below
        asm volatile("bar.sync 1, 128;" ::: "memory");
        const int pr = threadIdx.x >> 3, part = threadIdx.x & 7;  // 16 pairs x 8 adders = warps 0..3
        const double2* a = s_acc + pr * RED_PITCH;
        // at most RED_CAP / 8 = 12 entries per adder: all loads issued at once (predicated), then an add tree
        double2 v[RED_CAP / 8];
#pragma unroll
        for (int k = 0; k < RED_CAP / 8; ++k) {
          const int t = part + 8 * k;
          v[k] = t < n_corr ? a[t] : make_double2(0.0, 0.0);
        }
        static_assert(RED_CAP == 96, "the add tree below is written for 12 entries per adder");
#pragma unroll
        for (int k = 0; k < 6; ++k) v[k].x += v[k + 6].x, v[k].y += v[k + 6].y;
#pragma unroll
        for (int k = 0; k < 3; ++k) v[k].x += v[k + 3].x, v[k].y += v[k + 3].y;
        v[0].x += v[1].x, v[0].y += v[1].y;
        v[0].x += v[2].x, v[0].y += v[2].y;
        double sx = v[0].x, sy = v[0].y;
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
          sx += __shfl_xor_sync(0xffffffffu, sx, o);
          sy += __shfl_xor_sync(0xffffffffu, sy, o);
        }
        if (part == 0) {
          s_sum[2 * pr] = sx;
          s_sum[2 * pr + 1] = sy;  // [31] is padding
        }
      }
    } else if (warp_has_work || !resident) {
      double acc[LZ_NACC];
#pragma unroll
      for (int i = 0; i < LZ_NACC; ++i) acc[i] = 0;
      if (resident) {
        if ((int) threadIdx.x < n_corr) accumulate_loaded(c, edim, mine, acc, &my_status);
      } else {
        for (int k = threadIdx.x; k < n_corr; k += LZ_THREADS)
          accumulate_correspondence(c, edim, moving_xyz, fixed_meas, fixed_dim, corr_fixed[k], corr_moving[k], info_diag, acc,
                                    status ? status + k : nullptr);
      }
#pragma unroll
      for (int i = 0; i < LZ_NACC - 1; ++i) {
        const double s = warp_sum(acc[i]);
        if (lane == 0) s_part[wid][i] = s;
      }
    } else if (lane < LZ_NACC - 1) {
      s_part[wid][lane] = 0.0;
    }
    __syncthreads();
    if (!small && threadIdx.x < LZ_NACC - 1) {
      double s = 0;
      for (int w = 0; w < LZ_THREADS / 32; ++w) s += s_part[w][threadIdx.x];
      s_sum[threadIdx.x] = s;
    }
    if (!small) __syncthreads();
    if (wid == 0) {
      const int slot = it % OUT_RING;
      if (lane >= 12 && lane < 16) s_out[slot * GN_OUT + lane] = s_sum[27 + (lane - 12)];
      const bool ok = gn_solve_update_warp(S, s_sum, prior.enabled != 0, damping, lane, s_out + slot * GN_OUT);  // pose untouched when not SPD
      if (lane == 0) s_ok = ok ? 1 : 0;
      if (slot == OUT_RING - 1 || it == n_iters - 1 || !ok) {  // burst: rows it - slot .. it
        __syncwarp();
        double* o = out + (size_t) (it - slot) * GN_OUT;
        for (int i = lane; i < (slot + 1) * GN_OUT; i += 32) o[i] = s_out[i];
      }
    }
    __syncthreads();
    ++done;  // the iteration was linearised (its stats are valid) even when the solve failed
    if (!s_ok) break;
  }
  if (status && resident && (int) threadIdx.x < n_corr && done > 0) status[threadIdx.x] = my_status;
  if (threadIdx.x == 0) {
    iters_done[0] = done;
    iters_done[1] = s_ok;
    if (state) {
      // MultiAligner_::compute after the solver block: estimate <- last pose; the finder's bookkeeping for the `done - 1` calls
      // that kept these correspondences (each: previous <- the pose it was given, ++iteration; none once converged; skipped
      // when the solve failed, as the caller leaves its loop before them)
      const int ph = state->phases - 1;
      if (ph >= 0 && ph < PSLAM_ALIGN_MAX_PHASES) {
        state->phase_log[3 * ph] = state->it;
        state->phase_log[3 * ph + 1] = done;
        state->phase_log[3 * ph + 2] = state->n_corr;
      }
      if (done > 0)
        for (int i = 0; i < 12; ++i) state->estimate[i] = (i & 3) == 3 ? S.t[i >> 2] : S.R[3 * (i >> 2) + (i & 3)];
      if (s_ok && !state->has_converged) {
        if (done >= 2)
          for (int i = 0; i < 12; ++i) state->prev[i] = (float) ((i & 3) == 3 ? S.tp[i >> 2] : S.Rp[3 * (i >> 2) + (i & 3)]);
        state->current_iteration += done - 1;
      }
      state->it += done;
      if (!s_ok) state->stop = 4;
      else if (state->it >= max_iterations) state->stop = 1;
    }
  }
}

// single linearisation with the optional prior folded in on the device: out46 as linearize_finish_kernel, [46] prior chi
__global__ void prior_add_kernel(PriorParams prior, LinParams c, double* __restrict__ out) {
  if (threadIdx.x != 0 || !prior.enabled) return;
  double H[36], b[6];
  for (int i = 0; i < 36; ++i) H[i] = out[i];
  for (int i = 0; i < 6; ++i) b[i] = out[36 + i];
  out[46] = pose_prior_accumulate_dev(prior, c.R, c.t, H, b);
  for (int i = 0; i < 36; ++i) out[i] = H[i];
  for (int i = 0; i < 6; ++i) out[36 + i] = b[i];
}

LinParams make_params(const pslam_linearize_cfg* cfg, const double* pose12) {
  LinParams c;
  c.kind = cfg->kind;
  c.robustifier = cfg->robustifier;
  for (int i = 0; i < 9; ++i) c.K[i] = cfg->K[i];
  c.cols = cfg->image_cols;
  c.rows = cfg->image_rows;
  for (int i = 0; i < 3; ++i) c.baseline[i] = cfg->baseline[i];
  c.mean_disparity = cfg->mean_disparity;
  c.chi_threshold = cfg->chi_threshold;
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) c.R[3 * i + j] = pose12[4 * i + j];
    c.t[i] = pose12[4 * i + 3];
  }
  return c;
}

PriorParams make_prior(const pslam_pose_prior* prior) {
  PriorParams p;
  memset(&p, 0, sizeof(p));
  if (!prior) return p;
  p.enabled = 1;
  const double* Z = prior->prediction;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) p.Zi_R[3 * i + j] = Z[4 * j + i];
  for (int i = 0; i < 3; ++i)
    p.Zi_t[i] = -((p.Zi_R[3 * i] * Z[3] + p.Zi_R[3 * i + 1] * Z[7]) + p.Zi_R[3 * i + 2] * Z[11]);
  for (int i = 0; i < 36; ++i) p.Omega[i] = prior->information[i];
  return p;
}

}  // namespace

template <typename T>
int pslam_k_linearize_t(pslam_ctx* ctx, const pslam_linearize_cfg* cfg, const double* pose12, const T* d_moving_xyz,
                        const T* d_fixed_meas, int fixed_dim, int n_corr, const int* d_corr_fixed,
                        const int* d_corr_moving, const T* d_info_diag, const pslam_pose_prior* prior, uint8_t* d_status,
                        double* h_H36, double* h_b6, double* h_stats5) {
  const LinParams c = make_params(cfg, pose12);
  int blocks = (n_corr + LZ_THREADS - 1) / LZ_THREADS;
  if (blocks < 1) blocks = 1;
  if (blocks > 592) blocks = 592;  // 4 CTAs per SM on 148 SMs
  // the tail of the scratch buffer holds the partial rows and the result
  const size_t need = sizeof(double) * ((size_t) blocks * LZ_NACC + 64);
  if (need > ctx->scratch_bytes) return pslam_set_error(ctx, PSLAM_E_CAPACITY, "linearize: scratch too small", cudaSuccess);
  double* partials = reinterpret_cast<double*>(ctx->d_scratch + ctx->scratch_bytes - need);
  double* out = partials + (size_t) blocks * LZ_NACC;
  linearize_kernel<T><<<blocks, LZ_THREADS, 0, ctx->stream>>>(c, d_moving_xyz, d_fixed_meas, fixed_dim, n_corr, d_corr_fixed,
                                                              d_corr_moving, d_info_diag, partials, d_status);
  PSLAM_LAUNCH_CHECK(ctx, "linearize_kernel");
  linearize_finish_kernel<<<1, 32, 0, ctx->stream>>>(partials, blocks, out);
  PSLAM_LAUNCH_CHECK(ctx, "linearize_finish_kernel");
  if (prior) {
    prior_add_kernel<<<1, 32, 0, ctx->stream>>>(make_prior(prior), c, out);
    PSLAM_LAUNCH_CHECK(ctx, "prior_add_kernel");
  }
  double* h = reinterpret_cast<double*>(ctx->h_pinned);
  PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(h, out, sizeof(double) * 47, cudaMemcpyDeviceToHost, ctx->stream));
  PSLAM_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  memcpy(h_H36, h, sizeof(double) * 36);
  memcpy(h_b6, h + 36, sizeof(double) * 6);
  memcpy(h_stats5, h + 42, sizeof(double) * 4);
  h_stats5[4] = prior ? h[46] : 0.0;
  return PSLAM_OK;
}
template int pslam_k_linearize_t<double>(pslam_ctx*, const pslam_linearize_cfg*, const double*, const double*, const double*, int, int,
                                         const int*, const int*, const double*, const pslam_pose_prior*, uint8_t*, double*, double*,
                                         double*);
template int pslam_k_linearize_t<float>(pslam_ctx*, const pslam_linearize_cfg*, const double*, const float*, const float*, int, int,
                                        const int*, const int*, const float*, const pslam_pose_prior*, uint8_t*, double*, double*,
                                        double*);

int pslam_k_gn_step(pslam_ctx* ctx, const double* H36, const double* b6, double damping,
                    double* pose12, double* dx6) {
  double* h = reinterpret_cast<double*>(ctx->h_pinned);
  memcpy(h, H36, sizeof(double) * 36);
  memcpy(h + 36, b6, sizeof(double) * 6);
  h[42] = damping;
  memcpy(h + 43, pose12, sizeof(double) * 12);
  double* d = reinterpret_cast<double*>(ctx->d_scratch + ctx->scratch_bytes - 64 * sizeof(double));
  PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(d, h, sizeof(double) * 62, cudaMemcpyHostToDevice, ctx->stream));
  gn_step_kernel<<<1, 32, 0, ctx->stream>>>(d);
  PSLAM_LAUNCH_CHECK(ctx, "gn_step_kernel");
  PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(h, d, sizeof(double) * 62, cudaMemcpyDeviceToHost, ctx->stream));
  PSLAM_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  if (h[61] != 0) return pslam_set_error(ctx, PSLAM_E_NOT_SPD, "gn_step: H + damping*I is not positive definite", cudaSuccess);
  memcpy(pose12, h + 43, sizeof(double) * 12);
  if (dx6) memcpy(dx6, h + 55, sizeof(double) * 6);
  return PSLAM_OK;
}

// host clouds of scalar type T (float = the reference's cloud type, no widening anywhere; double = legacy entry point)
template <typename T>
int pslam_k_gn_iterate_t(pslam_ctx* ctx, const pslam_linearize_cfg* cfg, int n_iters, double damping, double* pose12,
                         int n_moving, const T* h_moving_xyz, int n_fixed, const T* h_fixed_meas, int fixed_dim,
                         int n_corr, const int* h_corr_fixed, const int* h_corr_moving, const T* h_info_diag,
                         const pslam_pose_prior* prior, double* h_out16, uint8_t* h_status, int* h_iters_done, int* h_spd) {
  const LinParams c = make_params(cfg, pose12);
  auto al = [](size_t x) { return (x + 255) & ~(size_t) 255; };
  const size_t b_mov = al(3 * sizeof(T) * (size_t) n_moving), b_fix = al(sizeof(T) * (size_t) fixed_dim * n_fixed),
               b_cf = al(4 * (size_t) n_corr), b_info = al(3 * sizeof(T) * (size_t) n_fixed), b_out = al(8 * (size_t) GN_OUT * n_iters),
               b_status = h_status ? al((size_t) n_corr) : 0;
  if (PSLAM_SOLVER_SCRATCH_OFFSET + b_mov + b_fix + 2 * b_cf + b_info + b_out + b_status + 256 > ctx->scratch_bytes)
    return pslam_set_error(ctx, PSLAM_E_CAPACITY, "gn_iterate: problem exceeds the scratch buffer", cudaSuccess);
  uint8_t* p = ctx->d_scratch + PSLAM_SOLVER_SCRATCH_OFFSET;
  uint8_t* const d_in = p;
  T* d_mov = (T*) p; p += b_mov;
  T* d_fix = (T*) p; p += b_fix;
  int* d_cf = (int*) p; p += b_cf;
  int* d_cm = (int*) p; p += b_cf;
  T* d_info = (T*) p; p += b_info;
  const size_t in_bytes = (size_t) (p - d_in);
  int* d_done = (int*) p; p += 256;  // [iterations done, spd flag], directly in front of the per-iteration rows: one download
  double* d_out = (double*) p; p += b_out;
  uint8_t* d_status = h_status ? p : nullptr;  // directly behind the rows: still one download
  const size_t out_bytes = 256 + b_out + b_status;
  // per frame this is called ~20 times on a few hundred points: the call is bound by copy / launch latencies.  The five
  // inputs are packed into the pinned staging block and travel as ONE copy, the results come back as ONE copy.
  uint8_t* h_stage = reinterpret_cast<uint8_t*>(ctx->h_pinned);
  const bool packed = in_bytes + out_bytes <= ctx->pinned_bytes;
  if (packed) {
    memcpy(h_stage + ((uint8_t*) d_mov - d_in), h_moving_xyz, 3 * sizeof(T) * (size_t) n_moving);
    memcpy(h_stage + ((uint8_t*) d_fix - d_in), h_fixed_meas, sizeof(T) * (size_t) fixed_dim * n_fixed);
    memcpy(h_stage + ((uint8_t*) d_cf - d_in), h_corr_fixed, 4 * (size_t) n_corr);
    memcpy(h_stage + ((uint8_t*) d_cm - d_in), h_corr_moving, 4 * (size_t) n_corr);
    memcpy(h_stage + ((uint8_t*) d_info - d_in), h_info_diag, 3 * sizeof(T) * (size_t) n_fixed);
    PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(d_in, h_stage, in_bytes, cudaMemcpyHostToDevice, ctx->stream));
  } else {
    PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(d_mov, h_moving_xyz, 3 * sizeof(T) * (size_t) n_moving, cudaMemcpyHostToDevice, ctx->stream));
    PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(d_fix, h_fixed_meas, sizeof(T) * (size_t) fixed_dim * n_fixed, cudaMemcpyHostToDevice, ctx->stream));
    PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(d_cf, h_corr_fixed, 4 * (size_t) n_corr, cudaMemcpyHostToDevice, ctx->stream));
    PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(d_cm, h_corr_moving, 4 * (size_t) n_corr, cudaMemcpyHostToDevice, ctx->stream));
    PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(d_info, h_info_diag, 3 * sizeof(T) * (size_t) n_fixed, cudaMemcpyHostToDevice, ctx->stream));
  }
  gn_iterate_kernel<T><<<1, LZ_THREADS, 0, ctx->stream>>>(c, make_prior(prior), damping, n_iters, d_mov, d_fix, fixed_dim, n_corr,
                                                         d_cf, d_cm, d_info, d_out, d_done, d_status, nullptr, nullptr, 0);
  PSLAM_LAUNCH_CHECK(ctx, "gn_iterate_kernel");
  int h_small[2];
  const int* h = h_small;
  if (packed) {
    uint8_t* h_res = h_stage + in_bytes;
    PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(h_res, d_done, out_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    PSLAM_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    h = reinterpret_cast<const int*>(h_res);
    memcpy(h_out16, h_res + 256, 8 * (size_t) GN_OUT * n_iters);
    if (h_status) memcpy(h_status, h_res + 256 + b_out, (size_t) n_corr);
  } else {
    PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(h_small, d_done, 2 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(h_out16, d_out, 8 * (size_t) GN_OUT * n_iters, cudaMemcpyDeviceToHost, ctx->stream));
    if (h_status) PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(h_status, d_status, (size_t) n_corr, cudaMemcpyDeviceToHost, ctx->stream));
    PSLAM_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  }
  const int done = h[0];
  *h_iters_done = done;
  *h_spd = h[1];
  if (done > 0) memcpy(pose12, h_out16 + (size_t) (done - 1) * GN_OUT, sizeof(double) * 12);
  return PSLAM_OK;
}
// everything already on the device (fp32 clouds of the projective finder's cache, correspondences + count written by the
// launch before): launch only.  d_out: n_iters x 16 doubles, d_done: 2 ints, d_status: one byte per correspondence.
int pslam_k_gn_iterate_dev(pslam_ctx* ctx, const pslam_linearize_cfg* cfg, int n_iters, double damping, const double* pose12,
                           const float* d_moving_xyz, const float* d_fixed_meas, int fixed_dim, const int* d_n_corr,
                           const int* d_corr_fixed, const int* d_corr_moving, const float* d_info_diag,
                           const pslam_pose_prior* prior, double* d_out, int* d_done, uint8_t* d_status,
                           PslamAlignState* d_state, int max_iterations) {
  const LinParams c = make_params(cfg, pose12);
  pslam_launch_pdl(gn_iterate_kernel<float>, dim3(1), dim3(LZ_THREADS), 0, ctx->stream, c, make_prior(prior), damping, n_iters,
                   d_moving_xyz, d_fixed_meas, fixed_dim, 0, d_corr_fixed, d_corr_moving, d_info_diag, d_out, d_done, d_status, d_n_corr,
                   d_state, max_iterations);
  PSLAM_LAUNCH_CHECK(ctx, "gn_iterate_kernel");
  return PSLAM_OK;
}

template int pslam_k_gn_iterate_t<double>(pslam_ctx*, const pslam_linearize_cfg*, int, double, double*, int, const double*, int,
                                          const double*, int, int, const int*, const int*, const double*, const pslam_pose_prior*,
                                          double*, uint8_t*, int*, int*);
template int pslam_k_gn_iterate_t<float>(pslam_ctx*, const pslam_linearize_cfg*, int, double, double*, int, const float*, int,
                                         const float*, int, int, const int*, const int*, const float*, const pslam_pose_prior*,
                                         double*, uint8_t*, int*, int*);
