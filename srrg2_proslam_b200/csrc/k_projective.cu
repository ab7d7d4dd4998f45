// k_projective.cu -- stage 2c (sm_100a): projective landmark-to-keypoint window matching.
//
// Replaces, for CorrespondenceFinderProjective{Square,Circle,Rhombus}:
//   _initializeDatabase      .../correspondence_finders/correspondence_finder_projective_square_impl.cpp:7-31
//   projector->compute       srrg2_core PointProjectorPinhole_ (call site ..._projective_base_impl.cpp:165-166)
//   _findNearestNeighbors    ..._square_impl.cpp:33-118, ..._circle_impl.cpp:7-94, ..._rhombus_impl.cpp:7-93
//   _filterCorrespondences   ..._projective_base_impl.cpp:39-102
//
// The lattice is sorted by row ONLY with an unstable std::sort in the reference; the scan order
// inside equal rows breaks Hamming ties, so one lane replays libstdc++'s introsort once per fixed
// cloud (cached until the next set_fixed, like the reference's _initializeDatabase).  The window
// search runs one warp per projected point: lanes stride over the row band, each keeps its two
// lexicographically smallest (distance, lattice position) keys, a shuffle tree merges them --
// which equals the reference's sequential best / second-best update with "first wins" ties.
#include <float.h>

#include <atomic>
#include <unordered_map>
#include <vector>

#include "libstdcxx_sort.h"
#include "pslam_internal.cuh"
#include "pslam_kernels.cuh"

namespace {

struct RowLess {  // [](a, b) { return a.row < b.row; }   (square_impl.cpp:27-29)
  __device__ __forceinline__ bool operator()(const unsigned long long& a,
                                             const unsigned long long& b) const {
    return (int) (a >> 32) < (int) (b >> 32);
  }
};

// lattice element = (row << 32) | (col << 16) | index   (Element{int16 row, col, index}, square.h:37-47)
// The unstable std::sort by row is replayed move for move (libstdcxx_sort.h): warp 0 runs the introsort partitioning with
// ballots, then the whole CTA places every element at its final-insertion-sort position (113 -> ~10 us per fixed cloud:
// the lattice is rebuilt for every frame).  `tmp` / `g_rpos`: global scratch for clouds that do not fit shared memory.
constexpr int LB_THREADS = 1024;
__global__ void __launch_bounds__(LB_THREADS)
lattice_build_kernel(const float* __restrict__ coords, int dim, int n, unsigned long long* __restrict__ lattice, int smem_cap,
                     unsigned long long* __restrict__ tmp, unsigned* __restrict__ g_rpos) {
  extern __shared__ __align__(16) unsigned long long s_l[];
  __shared__ int s_end;
  const int tid = threadIdx.x;
  const bool in_smem = n <= smem_cap;
  unsigned long long* a = in_smem ? s_l : tmp;
  for (int i = tid; i < n; i += LB_THREADS) {
    const short row = (short) coords[(size_t) dim * i + 1];  // int16(coordinates(1))
    const short col = (short) coords[(size_t) dim * i + 0];
    a[i] = ((unsigned long long) (unsigned) (int) row << 32) |
           ((unsigned long long) (unsigned short) col << 16) | (unsigned long long) (unsigned short) i;
  }
  __threadfence_block();
  __syncthreads();
  constexpr int LB_SEG_CAP = 256;  // >= 4096 / 17 live segments
  __shared__ int s_seg[2 * 3 * LB_SEG_CAP], s_cnt[3];
  if (in_smem && n <= 4096) {
    // frame-sized clouds: the partitions of a recursion level run on all the warps of the block
    pslam_sort::block_std_sort_full<LB_THREADS, LB_SEG_CAP>(a, reinterpret_cast<unsigned short*>(s_l + smem_cap), n, RowLess(), s_seg,
                                                            s_cnt);
    if (tid == 0) s_end = n;
  } else if (tid < 32) {
    int se;
    if (in_smem) se = pslam_sort::warp_std_sort_prefix(a, reinterpret_cast<unsigned short*>(s_l + smem_cap), n, n, RowLess());
    else se = pslam_sort::warp_std_sort_prefix(a, g_rpos, n, n, RowLess());
    if (tid == 0) s_end = se;
  }
  __threadfence_block();
  __syncthreads();
  pslam_sort::block_final_positions<LB_THREADS>(a, s_end, n, lattice, RowLess());
}

__device__ __forceinline__ void top2_insert(unsigned& k1, unsigned& k2, unsigned k) {
  if (k < k1) {
    k2 = k1;
    k1 = k;
  } else if (k < k2) {
    k2 = k;
  }
}

struct ProjParams {
  float R[9], t[3], K[9];
  float canvas_cols, canvas_rows, range_min, range_max;
  int shape, radius;
  float max_desc_dist;  // shape 3
};

// cand[4*m + {0,1,2,3}] = {fixed_best, dist_best, fixed_second, dist_second}, -1 where absent.
// One warp per moving point.
constexpr int PJ_WARPS = 8;
__global__ void __launch_bounds__(PJ_WARPS * 32)
projective_search_kernel(ProjParams pp, const float* __restrict__ moving_xyz, int n_moving,
                         const uint32_t* __restrict__ desc_moving,
                         const unsigned long long* lattice, int n_fixed,
                         const uint32_t* __restrict__ desc_fixed, int* __restrict__ cand,
                         int* __restrict__ n_projected, const float* __restrict__ fixed_coords, int fixed_dim,
                         const PslamAlignState* __restrict__ state, int stage_lattice) {
  extern __shared__ __align__(8) int s_width[];  // circle: width per |height| (0..radius); then the staged lattice
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  pslam_pdl_enter();
  if (state) {  // a phase of pslam_projective_align: the pose is the solver's current estimate, rounded to fp32 as the caller's
                // finder.setLocalMapInSensor(X) does (uniform reads; the state is only written by later launches)
    if (state->stop) return;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
      for (int j = 0; j < 3; ++j) pp.R[3 * i + j] = (float) state->estimate[4 * i + j];
      pp.t[i] = (float) state->estimate[4 * i + 3];
    }
  }
  if (pp.shape == 1) {
    const int r2 = pp.radius * pp.radius;
    for (int h = threadIdx.x; h <= pp.radius; h += PJ_WARPS * 32)
      s_width[h] = (int) sqrt((double) (r2 - h * h)) + 1;  // circle_impl.cpp:51-54
  }
  // frame-sized fixed clouds: the lattice (8 B per point) is staged in shared memory once per CTA -- the binary search and the
  // window walk below are chains of dependent loads, an L2 round trip apiece otherwise
  if (stage_lattice) {
    unsigned long long* s_lat = reinterpret_cast<unsigned long long*>(s_width + ((pp.radius + 2) & ~1));
    for (int i = threadIdx.x; i < n_fixed; i += PJ_WARPS * 32) s_lat[i] = lattice[i];
    lattice = s_lat;
  }
  __syncthreads();
  const int m = blockIdx.x * PJ_WARPS + wid;
  if (m >= n_moving) return;
  // ---- pinhole projection, fp32, operation order of oracle/pslam_oracle_solver.hpp::project_point
  const float px = moving_xyz[3 * m], py = moving_xyz[3 * m + 1], pz = moving_xyz[3 * m + 2];
  float c[3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
    c[i] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(pp.R[3 * i], px), __fmul_rn(pp.R[3 * i + 1], py)),
                               __fmul_rn(pp.R[3 * i + 2], pz)), pp.t[i]);
  bool valid = !(c[2] < pp.range_min || c[2] > pp.range_max);
  float u = 0, v = 0;
  if (valid) {
    const float hx = __fadd_rn(__fadd_rn(__fmul_rn(pp.K[0], c[0]), __fmul_rn(pp.K[1], c[1])), __fmul_rn(pp.K[2], c[2]));
    const float hy = __fadd_rn(__fadd_rn(__fmul_rn(pp.K[3], c[0]), __fmul_rn(pp.K[4], c[1])), __fmul_rn(pp.K[5], c[2]));
    const float hz = __fadd_rn(__fadd_rn(__fmul_rn(pp.K[6], c[0]), __fmul_rn(pp.K[7], c[1])), __fmul_rn(pp.K[8], c[2]));
    u = __fdiv_rn(hx, hz);
    v = __fdiv_rn(hy, hz);
    valid = !(u < 0.0f || u > pp.canvas_cols || v < 0.0f || v > pp.canvas_rows);
  }
  int* out = cand + 4 * (size_t) m;
  if (!valid) {
    if (lane == 0) {
      out[0] = -2;  // not projected
      out[1] = 0;
      out[2] = -1;
      out[3] = 0;
    }
    return;
  }
  if (lane == 0 && n_projected) atomicAdd(n_projected, 1);
  const int row = (short) roundf(v), col = (short) roundf(u);  // std::round -> int16
  const int r = pp.radius;
  int row_min = (short) (row - r), row_max = (short) (row + r + 1);
  const int col_min = (short) (col - r - 1), col_max = (short) (col + r + 1);
  const float r2f = (float) ((size_t) r * (size_t) r);  // kdtree_impl.cpp:41-42
  if (pp.shape == 3) {
    // exact radius query: the lattice rows are int16(y) of the fixed points; |y - v| < r  =>  row in [floor(v - r), floor(v + r)]
    row_min = (int) floorf(v - (float) r) - 1;
    row_max = (int) floorf(v + (float) r) + 2;
  }
  // first lattice position with row >= row_min
  int lo = 0, hi = n_fixed;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if ((int) (lattice[mid] >> 32) < row_min) lo = mid + 1; else hi = mid;
  }
  const uint4 q0 = __ldg(reinterpret_cast<const uint4*>(desc_moving) + 2 * (size_t) m);
  const uint4 q1 = __ldg(reinterpret_cast<const uint4*>(desc_moving) + 2 * (size_t) m + 1);
  unsigned k1 = 0xffffffffu, k2 = 0xffffffffu;  // (dist << 16) | lattice position
  for (int base = lo; base < n_fixed; base += 32) {
    const int pos = base + lane;
    bool more = false;
    if (pos < n_fixed) {
      const unsigned long long e = lattice[pos];
      const int erow = (int) (e >> 32), ecol = (int) (short) ((e >> 16) & 0xffffu);
      if (erow < row_max) {
        more = true;
        bool in_window;
        if (pp.shape == 3) {
          const int fi = (int) (e & 0xffffu);
          const float dx = __fsub_rn(fixed_coords[(size_t) fixed_dim * fi], u), dy = __fsub_rn(fixed_coords[(size_t) fixed_dim * fi + 1], v);
          in_window = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)) < r2f;
        } else if (pp.shape == 0) {
          in_window = ecol > col_min && ecol < col_max;
        } else if (pp.shape == 1) {
          const int height = erow - row;
          const int width = s_width[height < 0 ? -height : height];
          in_window = ecol > col - width && ecol < col + width;
        } else {
          int width = (short) (erow - row_min + 1);
          if (width > (short) r) width = (short) (row_max - erow);
          in_window = ecol > col - width && ecol < col + width;
        }
        if (in_window) {
          const int fi = (int) (e & 0xffffu);
          const uint4 f0 = __ldg(reinterpret_cast<const uint4*>(desc_fixed) + 2 * (size_t) fi);
          const uint4 f1 = __ldg(reinterpret_cast<const uint4*>(desc_fixed) + 2 * (size_t) fi + 1);
          const unsigned d = (unsigned) hamming256(q0, q1, f0, f1);
          top2_insert(k1, k2, (d << 16) | (unsigned) pos);
        }
      }
    }
    if (!__any_sync(0xffffffffu, more)) break;
  }
  // merge the lanes' sorted pairs
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned a1 = __shfl_xor_sync(0xffffffffu, k1, o);
    const unsigned a2 = __shfl_xor_sync(0xffffffffu, k2, o);
    top2_insert(k1, k2, a1);
    top2_insert(k1, k2, a2);
  }
  if (lane == 0) {
    if (pp.shape == 3) {  // best must beat maximum_descriptor_distance (strict); no second candidate (kdtree_impl.cpp:55-78)
      const bool ok = k1 != 0xffffffffu && (float) (k1 >> 16) < pp.max_desc_dist;
      out[0] = ok ? (int) (lattice[k1 & 0xffffu] & 0xffffu) : -1;
      out[1] = ok ? (int) (k1 >> 16) : 0;
      out[2] = -1;
      out[3] = 0;
    } else {
      out[0] = (k1 != 0xffffffffu) ? (int) (lattice[k1 & 0xffffu] & 0xffffu) : -1;
      out[1] = (int) (k1 >> 16);
      out[2] = (k2 != 0xffffffffu) ? (int) (lattice[k2 & 0xffffu] & 0xffffu) : -1;
      out[3] = (int) (k2 >> 16);
    }
  }
}

constexpr int FF_THREADS = 1024;  // filter + compaction: one CTA
// ---- the finder's state machine between two searches (pslam_projective_align), thread 0 after the compaction -------------
// fp32 arithmetic of the host mirror (pslam_plugin.cpp: Isometry3f::inverse, operator*, t2tnq; separately rounded operations,
// IEEE sqrt / divide): the decisions are bit-identical to the call-by-call path's.
__device__ float align_estimate_change_norm(const float* X, const float* P) {
  float I[12], E[12];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) I[4 * i + j] = X[4 * j + i];
  for (int i = 0; i < 3; ++i)
    I[4 * i + 3] = -__fadd_rn(__fadd_rn(__fmul_rn(I[4 * i], X[3]), __fmul_rn(I[4 * i + 1], X[7])), __fmul_rn(I[4 * i + 2], X[11]));
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j)
      E[4 * i + j] = __fadd_rn(__fadd_rn(__fmul_rn(I[4 * i], P[j]), __fmul_rn(I[4 * i + 1], P[4 + j])), __fmul_rn(I[4 * i + 2], P[8 + j]));
    E[4 * i + 3] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(I[4 * i], P[3]), __fmul_rn(I[4 * i + 1], P[7])), __fmul_rn(I[4 * i + 2], P[11])),
                             I[4 * i + 3]);
  }
  float v6[6] = {E[3], E[7], E[11], 0, 0, 0};
  float w, q[3];
  float t = __fadd_rn(__fadd_rn(E[0], E[5]), E[10]);
  if (t > 0.f) {
    t = __fsqrt_rn(__fadd_rn(t, 1.f));
    w = __fmul_rn(0.5f, t);
    t = __fdiv_rn(0.5f, t);
    q[0] = __fmul_rn(__fsub_rn(E[9], E[6]), t);
    q[1] = __fmul_rn(__fsub_rn(E[2], E[8]), t);
    q[2] = __fmul_rn(__fsub_rn(E[4], E[1]), t);
  } else {
    int i = 0;
    if (E[5] > E[0]) i = 1;
    if (E[10] > (i == 0 ? E[0] : E[5])) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    t = __fsqrt_rn(__fadd_rn(__fsub_rn(__fsub_rn(E[5 * i], E[5 * j]), E[5 * k]), 1.f));
    q[i] = __fmul_rn(0.5f, t);
    t = __fdiv_rn(0.5f, t);
    w = __fmul_rn(__fsub_rn(E[4 * k + j], E[4 * j + k]), t);
    q[j] = __fmul_rn(__fadd_rn(E[4 * j + i], E[4 * i + j]), t);
    q[k] = __fmul_rn(__fadd_rn(E[4 * k + i], E[4 * i + k]), t);
  }
  const float n = __fsqrt_rn(__fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(w, w), __fmul_rn(q[0], q[0])), __fmul_rn(q[1], q[1])), __fmul_rn(q[2], q[2])));
  const float sg = __fdiv_rn(w < 0.f ? -1.f : 1.f, n);
  v6[3] = __fmul_rn(q[0], sg);
  v6[4] = __fmul_rn(q[1], sg);
  v6[5] = __fmul_rn(q[2], sg);
  float n2 = 0;
  for (int i = 0; i < 6; ++i) n2 = __fadd_rn(n2, __fmul_rn(v6[i], v6[i]));
  return __fsqrt_rn(n2);
}

// CorrespondenceFinderProjective_::compute after the search (base_impl.cpp:181-183, 228-288) + the caller's choice of how
// many solver iterations follow (the calls that keep these correspondences, :162-178)
__device__ void align_control(PslamAlignState* s, const PslamAlignCfg& a, int n, int n_fixed, int n_projected) {
  s->n_corr = n;
  s->n_projected = n_projected;
  const float ratio = __fdiv_rn((float) n, (float) n_fixed);
  if (ratio < a.min_matching_ratio && a.can_widen) {  // the finder repeats the call with its widest search: the caller's turn
    s->stop = 2;
    return;
  }
  if (n < (a.min_corr > 1 ? a.min_corr : 1)) {
    s->stop = 3;
    return;
  }
  float X[12];
  for (int i = 0; i < 12; ++i) X[i] = (float) s->estimate[i];
  const float norm = align_estimate_change_norm(X, s->prev);
  for (int i = 0; i < 12; ++i) s->prev[i] = X[i];
  const bool converged = norm < a.max_change_norm && s->current_iteration > a.min_iterations;
  int n_fused = a.max_iterations - s->it;
  if (converged) {
    s->has_converged = 1;
    s->converged_ratio_ok = ratio > a.min_matching_ratio ? 1 : 0;
  } else {
    int quiet = 0;
    for (int it = s->current_iteration + 1; !(it % a.per_projection == 0 || it == 1) && quiet < n_fused; ++it) ++quiet;
    if (quiet + 1 < n_fused) n_fused = quiet + 1;
  }
  s->n_fused = n_fused;
  s->current_iteration += 1;
  s->phases += 1;
}

// correspondences of the filter's result in ascending fixed index + their information diagonals, for the fused solver:
// acc_moving[f] >= 0  ->  (f, acc_moving[f]); info[3 f + k] = diag[k] * scale[moving]  (setupFactor, fp32 like the host)
// start of a device-resident alignment: handed to the first phase's filter kernel by value (one upload less per frame)
struct AlignInit {
  double estimate[12];
  float prev[12];
  int current_iteration, valid;
};
struct CompactArgs {
  int enabled;
  const float* scale;
  float d0, d1, d2;
  int *cf, *cm;
  float* info;
  int* n_corr;
  PslamAlignState* state;
  PslamAlignCfg acfg;
  AlignInit init;
};
__device__ __forceinline__ void corr_compact(const int* acc_moving, int n_fixed, const float* __restrict__ scale, float d0,
                                             float d1, float d2, int* __restrict__ cf, int* __restrict__ cm, float* __restrict__ info,
                                             int* __restrict__ n_corr, PslamAlignState* __restrict__ state, const PslamAlignCfg& acfg,
                                             int n_projected) {
  __shared__ int s_warp[33];
  int running = 0;
  for (int base = 0; base < n_fixed; base += FF_THREADS) {
    const int f = base + threadIdx.x;
    const int m = f < n_fixed ? acc_moving[f] : -1;
    int total;
    const int off = block_exclusive_scan<FF_THREADS>(m >= 0 ? 1 : 0, s_warp, &total);
    if (m >= 0) {
      cf[running + off] = f;
      cm[running + off] = m;
      const float s = scale ? scale[m] : 1.0f;
      info[3 * (size_t) f] = __fmul_rn(d0, s);
      info[3 * (size_t) f + 1] = __fmul_rn(d1, s);
      info[3 * (size_t) f + 2] = __fmul_rn(d2, s);
    }
    running += total;
  }
  if (threadIdx.x == 0) {
    *n_corr = running;
    if (state) align_control(state, acfg, running, n_fixed, n_projected);
  }
}

// Filtering (_filterCorrespondences, projective_base_impl.cpp:39-102): per fixed point the lowest and second lowest
// response over its candidates in insertion order (order key = 2 * moving_idx + {0 best, 1 second}; key = (dist << 32) |
// order), Lowe's ratio + distance threshold, then bijectivity (the moving point's own best candidate is its first entry).
// The three filter passes (+ the initialisation of their keys and the count of projected points) in ONE launch of one
// CTA: per frame the finder is called ~20 times on a few hundred points, so the search is bound by launch and copy
// latencies, not by work -- 3 memsets + 3 kernels + 4 copies became 1 kernel + 1 copy.  `out` is the contiguous block the
// host downloads: [n_projected][acc_moving x n_fixed][acc_dist x n_fixed (float bits)][cand x 4 n_moving].
// SMEM: frame-sized fixed clouds (<= FF_SMEM_FIXED points) keep the two key arrays and the accepted moving index in shared
// memory -- the passes below stop being L2 round trips.
constexpr int FF_SMEM_FIXED = 2048;
template <bool SMEM>
__global__ void __launch_bounds__(FF_THREADS)
filter_fused_kernel(const int* __restrict__ cand, int n_moving, int n_fixed, unsigned long long* __restrict__ g_key1,
                    unsigned long long* __restrict__ g_key2, float max_dist, float max_ratio, int* __restrict__ out, CompactArgs ca) {
  __shared__ int s_proj;
  __shared__ unsigned long long s_key1[SMEM ? FF_SMEM_FIXED : 1], s_key2[SMEM ? FF_SMEM_FIXED : 1];
  __shared__ int s_res[SMEM ? FF_SMEM_FIXED : 1];
  unsigned long long* key1 = SMEM ? s_key1 : g_key1;
  unsigned long long* key2 = SMEM ? s_key2 : g_key2;
  const int tid = threadIdx.x;
  pslam_pdl_enter();
  if (ca.state && ca.init.valid) {  // first phase: the state starts here (the search of this phase took its pose by value too)
    if (tid == 0) {
      PslamAlignState* s = ca.state;
      for (int i = 0; i < 12; ++i) {
        s->estimate[i] = ca.init.estimate[i];
        s->prev[i] = ca.init.prev[i];
      }
      s->current_iteration = ca.init.current_iteration;
      s->has_converged = s->converged_ratio_ok = s->it = s->stop = s->phases = s->n_fused = s->n_corr = s->n_projected = s->pad = 0;
    }
    __syncthreads();
  }
  if (ca.state && ca.state->stop) return;
  if (tid == 0) s_proj = 0;
  for (int f = tid; f < n_fixed; f += FF_THREADS) {
    key1[f] = ~0ULL;
    key2[f] = ~0ULL;
  }
  __syncthreads();
  int proj = 0;
  for (int i = tid; i < 2 * n_moving; i += FF_THREADS) {  // pass 1: lowest response per fixed point
    const int m = i >> 1, s = i & 1;
    const int f = cand[4 * m + 2 * s];
    if (s == 0 && f != -2) ++proj;  // -2 marks "not projected" (projective_search_kernel)
    if (f < 0) continue;
    const unsigned long long k = ((unsigned long long) (unsigned) cand[4 * m + 2 * s + 1] << 32) | (unsigned) i;
    atomicMin(key1 + f, k);
  }
  if (proj) atomicAdd(&s_proj, proj);
  __syncthreads();
  for (int i = tid; i < 2 * n_moving; i += FF_THREADS) {  // pass 2: second lowest
    const int m = i >> 1, s = i & 1;
    const int f = cand[4 * m + 2 * s];
    if (f < 0) continue;
    const unsigned long long k = ((unsigned long long) (unsigned) cand[4 * m + 2 * s + 1] << 32) | (unsigned) i;
    if (k != key1[f]) atomicMin(key2 + f, k);
  }
  __syncthreads();
  if (tid == 0) out[0] = s_proj;
  int* acc_moving = out + 1;
  int* acc_dist = out + 1 + n_fixed;
  for (int f = tid; f < n_fixed; f += FF_THREADS) {  // pass 3: thresholds + bijectivity
    const unsigned long long a = key1[f], b = key2[f];
    int res = -1;
    float dist = 0;
    if (a != ~0ULL) {
      const float lowest = (float) (unsigned) (a >> 32);
      const float second = (b != ~0ULL) ? (float) (unsigned) (b >> 32) : FLT_MAX;
      if (lowest < max_dist && __fdiv_rn(lowest, second) < max_ratio) {  // base_impl.cpp:67-69
        const unsigned order = (unsigned) (a & 0xffffffffu);
        if ((order & 1u) == 0) {  // bijectivity (:82-99)
          res = (int) (order >> 1);
          dist = lowest;
        }
      }
    }
    acc_moving[f] = res;
    acc_dist[f] = __float_as_int(dist);
    if (SMEM) s_res[f] = res;
  }
  int* cand_out = out + 1 + 2 * n_fixed;
  for (int i = tid; i < 4 * n_moving; i += FF_THREADS) cand_out[i] = cand[i];
  if (ca.enabled) {  // the fused solver's input (+ the finder's decisions) in the same launch: one kernel boundary less per search
    __syncthreads();
    corr_compact(SMEM ? s_res : acc_moving, n_fixed, ca.scale, ca.d0, ca.d1, ca.d2, ca.cf, ca.cm, ca.info, ca.n_corr, ca.state, ca.acfg,
                 s_proj);
  }
}

}  // namespace

static inline size_t al256(size_t x) { return (x + 255) & ~(size_t) 255; }

// Scratch layout of the projective matcher (persistent between set_fixed / set_moving / match):
//   [fixed coords f32 n*dim][fixed desc][lattice u64][moving xyz][moving desc][cand][key1][key2][acc]
struct ProjState {
  int n_fixed, fixed_dim, n_moving;
  float* d_fixed;
  uint32_t* d_desc_fixed;
  unsigned long long* d_lattice;
  float* d_moving;
  uint32_t* d_desc_moving;
  int* d_cand;
  unsigned long long *d_key1, *d_key2;
  int* d_acc_m;
  float* d_acc_d;
  int* d_nproj;
  float* d_mscale;            // information scale per moving point (pslam_projective_set_moving_weights)
  int *d_gn_cf, *d_gn_cm;     // correspondences of the last search in ascending fixed index (fused solver)
  float* d_gn_info;           // information diagonal per FIXED point
  int* d_gn_ncorr;
};

static std::atomic<unsigned long long> g_proj_epoch{0};  // process-wide: epochs of different contexts never coincide

// one packed upload: `a` then `b` at the next 256-byte boundary, through a buffer the thread keeps (a pageable source is staged by
// the driver before cudaMemcpyAsync returns)
static int proj_upload2(pslam_ctx* ctx, void* dst, const void* a, size_t a_bytes, const void* b, size_t b_bytes) {
  static thread_local std::vector<unsigned char> pack;
  const size_t off = al256(a_bytes), total = off + b_bytes;
  if (pack.size() < total) pack.resize(total);
  memcpy(pack.data(), a, a_bytes);
  memcpy(pack.data() + off, b, b_bytes);
  PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(dst, pack.data(), total, cudaMemcpyHostToDevice, ctx->stream));
  return PSLAM_OK;
}

// n_fixed / n_moving < 0: "as uploaded" (the set_* calls pass their own sizes after recording them in the context)
static int proj_layout(pslam_ctx* ctx, ProjState& st, int n_fixed, int dim, int n_moving) {
  const int M = PSLAM_MAX_FEATURES_HARD * 2;  // generous fixed-size regions so that set_* calls are independent
  if (n_fixed > M || n_moving > 65536)
    return pslam_set_error(ctx, PSLAM_E_CAPACITY, "projective: too many points", cudaSuccess);
  if (!ctx->d_proj) {  // the cache lives in its own allocation: nothing else carves from it
    ctx->proj_bytes = PSLAM_SOLVER_SCRATCH_OFFSET;
    PSLAM_CUDA_TRY(ctx, cudaMalloc(&ctx->d_proj, ctx->proj_bytes));
  }
  // a cloud's descriptors sit directly behind its coordinates (at the 256-byte boundary after them) inside the cloud's
  // region: coordinates + descriptors arrive in ONE copy (an upload costs the host ~4.5 us of API time whatever its size)
  uint8_t* p = ctx->d_proj;
  st.d_fixed = (float*) p;
  st.d_desc_fixed = (uint32_t*) (p + al256(sizeof(float) * (size_t) ctx->proj_fixed_dim * (size_t) ctx->proj_n_fixed));
  p += al256(sizeof(float) * 4 * (size_t) M) + al256(32 * (size_t) M);
  st.d_lattice = (unsigned long long*) p; p += al256(8 * (size_t) M);
  st.d_key1 = (unsigned long long*) p; p += al256(8 * (size_t) M);
  st.d_key2 = (unsigned long long*) p; p += al256(8 * (size_t) M);
  st.d_acc_m = (int*) p; p += al256(4 * (size_t) M);
  st.d_acc_d = (float*) p; p += al256(4 * (size_t) M);
  st.d_nproj = (int*) p; p += 256;
  st.d_moving = (float*) p;
  st.d_desc_moving = (uint32_t*) (p + al256(sizeof(float) * 3 * (size_t) ctx->proj_n_moving));
  p += al256(sizeof(float) * 3 * (size_t) 65536) + al256(32 * (size_t) 65536);
  st.d_cand = (int*) p; p += al256(16 * (size_t) 65536);
  st.d_mscale = (float*) p; p += al256(4 * (size_t) 65536);
  st.d_gn_cf = (int*) p; p += al256(4 * (size_t) M);
  st.d_gn_cm = (int*) p; p += al256(4 * (size_t) M);
  st.d_gn_info = (float*) p; p += al256(12 * (size_t) M);
  st.d_gn_ncorr = (int*) p; p += 256;
  if ((size_t) (p - ctx->d_proj) > ctx->proj_bytes)
    return pslam_set_error(ctx, PSLAM_E_CAPACITY, "projective: cache allocation too small", cudaSuccess);
  st.n_fixed = n_fixed;
  st.fixed_dim = dim;
  st.n_moving = n_moving;
  return PSLAM_OK;
}

int pslam_k_projective_set_fixed(pslam_ctx* ctx, int n_fixed, const float* h_coords, int dim,
                                 const uint8_t* h_desc) {
  if (n_fixed >= 32767)
    return pslam_set_error(ctx, PSLAM_E_INVALID, "projective: int16 lattice needs < 32767 fixed points", cudaSuccess);
  if (n_fixed > PSLAM_MAX_FEATURES_HARD * 2) return pslam_set_error(ctx, PSLAM_E_CAPACITY, "projective: too many points", cudaSuccess);
  ctx->proj_fixed_epoch = ++g_proj_epoch;
  ctx->proj_fixed_dim = dim;
  ctx->proj_n_fixed = n_fixed;
  ProjState st;
  int rc = proj_layout(ctx, st, n_fixed, dim, 0);
  if (rc) return rc;
  if (n_fixed == 0) return PSLAM_OK;
  if ((rc = proj_upload2(ctx, st.d_fixed, h_coords, sizeof(float) * dim * (size_t) n_fixed, h_desc, 32 * (size_t) n_fixed))) return rc;
  // 10 B per element in shared memory (key + stopper position); clouds above the cap sort in global memory (key1 / key2 of
  // the finder cache are free until the next match)
  const int smem_cap = n_fixed <= 4096 ? 4096 : (n_fixed <= 16384 ? 16384 : 0);
  const size_t smem = 10 * (size_t) smem_cap;
  if (smem > 48 * 1024)
    PSLAM_CUDA_TRY(ctx, cudaFuncSetAttribute(lattice_build_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
  lattice_build_kernel<<<1, LB_THREADS, smem, ctx->stream>>>(st.d_fixed, dim, n_fixed, st.d_lattice, smem_cap, st.d_key1,
                                                             reinterpret_cast<unsigned*>(st.d_key2));
  PSLAM_LAUNCH_CHECK(ctx, "lattice_build_kernel");
  return PSLAM_OK;
}

int pslam_k_projective_set_moving(pslam_ctx* ctx, int n_moving, const float* h_xyz, const uint8_t* h_desc) {
  if (n_moving > 65536) return pslam_set_error(ctx, PSLAM_E_CAPACITY, "projective: too many points", cudaSuccess);
  ctx->proj_moving_epoch = ++g_proj_epoch;
  ctx->proj_n_moving = n_moving;
  ProjState st;
  int rc = proj_layout(ctx, st, 0, 2, n_moving);
  if (rc) return rc;
  if (n_moving == 0) return PSLAM_OK;
  return proj_upload2(ctx, st.d_moving, h_xyz, sizeof(float) * 3 * (size_t) n_moving, h_desc, 32 * (size_t) n_moving);
}

int pslam_k_projective_set_moving_weights(pslam_ctx* ctx, int n_moving, const float* h_scale) {
  if (n_moving != ctx->proj_n_moving)
    return pslam_set_error(ctx, PSLAM_E_INVALID, "projective: one information scale per point of the uploaded moving cloud", cudaSuccess);
  ProjState st;
  int rc = proj_layout(ctx, st, 0, 2, n_moving);
  if (rc) return rc;
  if (n_moving > 0)
    PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(st.d_mscale, h_scale, sizeof(float) * (size_t) n_moving, cudaMemcpyHostToDevice, ctx->stream));
  ctx->proj_weights_epoch = ctx->proj_moving_epoch;  // (a pageable source is staged before cudaMemcpyAsync returns)
  return PSLAM_OK;
}

// host side of a search: the downloaded block [n_projected][acc_moving x n_fixed][acc_dist x n_fixed][cand x 4 n_moving]
// (+ the fused solver's block at gn_off: [done, spd][rows x 16 doubles][status]) -> the caller's arrays
static int proj_collect(const int* h_out, int n_fixed, int n_moving, int capacity, int* h_fixed, int* h_moving, float* h_dist,
                        int* n_projected, pslam_fused_gn* gn, size_t gn_off, int gn_rows, int done, int spd) {
  const int nproj = h_out[0];
  const int* acc_m = h_out + 1;
  const float* acc_d = reinterpret_cast<const float*>(h_out + 1 + n_fixed);
  const int* cand = h_out + 1 + 2 * (size_t) n_fixed;
  if (n_projected) *n_projected = nproj;
  // Output order of the reference = iteration order of
  //   std::unordered_map<size_t, CorrespondenceVector> reserved with fixed->size() and filled in
  //   candidate order (base_impl.cpp:185-200, 50).  Replayed with the real container: keys only.
  std::unordered_map<size_t, int> by_fixed;
  by_fixed.reserve((size_t) n_fixed);
  for (int m = 0; m < n_moving; ++m) {
    const int fb = cand[4 * (size_t) m], fs = cand[4 * (size_t) m + 2];
    if (fb >= 0) {
      by_fixed.emplace((size_t) fb, 0);
      if (fs >= 0) by_fixed.emplace((size_t) fs, 0);
    }
  }
  // fused solver block: the device list is in ascending fixed index
  const int* h_gn = gn ? h_out + gn_off : nullptr;
  const double* h_gn_out = gn ? reinterpret_cast<const double*>(h_gn + 2) : nullptr;
  const uint8_t* h_status = gn ? reinterpret_cast<const uint8_t*>(h_gn_out + 16 * (size_t) gn_rows) : nullptr;
  std::vector<int> rank_of_fixed;
  if (gn && gn->factor_status) {
    rank_of_fixed.assign((size_t) n_fixed, -1);
    int r = 0;
    for (int f = 0; f < n_fixed; ++f)
      if (acc_m[f] >= 0) rank_of_fixed[f] = r++;
  }
  int n_out = 0;
  for (const auto& kv : by_fixed) {
    const int f = (int) kv.first;
    if (acc_m[f] < 0) continue;
    if (n_out < capacity) {
      h_fixed[n_out] = f;
      h_moving[n_out] = acc_m[f];
      h_dist[n_out] = acc_d[f];
      if (gn && gn->factor_status) gn->factor_status[n_out] = h_status[rank_of_fixed[f]];
    }
    ++n_out;
  }
  if (gn) {
    if (done < 0) {
      done = h_gn[0];
      spd = h_gn[1];
    }
    gn->iterations_done = done;
    gn->spd = spd;
    for (int i = 0; i < done; ++i) {
      if (gn->poses12) memcpy(gn->poses12 + 12 * (size_t) i, h_gn_out + 16 * (size_t) i, sizeof(double) * 12);
      if (gn->stats4) memcpy(gn->stats4 + 4 * (size_t) i, h_gn_out + 16 * (size_t) i + 12, sizeof(double) * 4);
    }
    if (done > 0) memcpy(gn->pose12, h_gn_out + 16 * (size_t) (done - 1), sizeof(double) * 12);
  }
  return n_out;
}

// dynamic shared memory of the search: the circle's width table (+ the lattice when the fixed cloud is frame sized)
constexpr int PJ_STAGE_MAX = 4096;
static size_t search_smem(int radius, int n_fixed) {
  return sizeof(int) * (size_t) ((radius + 2) & ~1) + (n_fixed <= PJ_STAGE_MAX ? 8 * (size_t) n_fixed : 0);
}
static int proj_params(pslam_ctx* ctx, const pslam_projective_cfg* cfg, const float* pose12, ProjParams& pp, int n_fixed) {
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) pp.R[3 * i + j] = pose12 ? pose12[4 * i + j] : (i == j ? 1.f : 0.f);
    pp.t[i] = pose12 ? pose12[4 * i + 3] : 0.f;
  }
  for (int i = 0; i < 9; ++i) pp.K[i] = cfg->K[i];
  pp.canvas_cols = (float) cfg->canvas_cols;
  pp.canvas_rows = (float) cfg->canvas_rows;
  pp.range_min = cfg->range_min;
  pp.range_max = cfg->range_max;
  pp.shape = cfg->shape;
  pp.radius = cfg->search_radius_pixels;
  pp.max_desc_dist = cfg->maximum_descriptor_distance;
  if (pp.shape < 0 || pp.shape > 3) return pslam_set_error(ctx, PSLAM_E_INVALID, "projective: unknown window shape", cudaSuccess);
  if (pp.radius < 0 || pp.radius > 16000)
    return pslam_set_error(ctx, PSLAM_E_INVALID, "projective: bad search radius", cudaSuccess);
  const size_t smem = search_smem(pp.radius, n_fixed);
  if (smem > 48 * 1024)
    PSLAM_CUDA_TRY(ctx, cudaFuncSetAttribute(projective_search_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
  return PSLAM_OK;
}

int pslam_k_projective_match(pslam_ctx* ctx, int n_fixed, int n_moving, const float* pose12,
                             const pslam_projective_cfg* cfg, int capacity, int* h_fixed, int* h_moving,
                             float* h_dist, int* n_projected, pslam_fused_gn* gn) {
  if (n_fixed != ctx->proj_n_fixed || n_moving != ctx->proj_n_moving)
    return pslam_set_error(ctx, PSLAM_E_INVALID, "projective: cloud sizes differ from the uploaded clouds", cudaSuccess);
  ProjState st;
  int rc = proj_layout(ctx, st, n_fixed, 2, n_moving);
  if (rc) return rc;
  if (n_projected) *n_projected = 0;
  if (n_fixed == 0 || n_moving == 0) return 0;
  ProjParams pp;
  if ((rc = proj_params(ctx, cfg, pose12, pp, n_fixed))) return rc;
  const size_t smem = search_smem(pp.radius, n_fixed);
  pslam_launch_pdl(projective_search_kernel, dim3((n_moving + PJ_WARPS - 1) / PJ_WARPS), dim3(PJ_WARPS * 32), smem, ctx->stream, pp,
                   st.d_moving, n_moving, st.d_desc_moving, st.d_lattice, n_fixed, st.d_desc_fixed, st.d_cand, nullptr, st.d_fixed,
                   ctx->proj_fixed_dim, nullptr, n_fixed <= PJ_STAGE_MAX ? 1 : 0);
  PSLAM_LAUNCH_CHECK(ctx, "projective_search_kernel");
  // one filter launch, one download (see filter_fused_kernel); the result block is transient: generic scratch
  const size_t n_words = 1 + 2 * (size_t) n_fixed + 4 * (size_t) n_moving;
  if (PSLAM_SOLVER_SCRATCH_OFFSET + 4 * n_words > ctx->scratch_bytes)
    return pslam_set_error(ctx, PSLAM_E_CAPACITY, "projective: scratch too small for the result block", cudaSuccess);
  int* d_out = reinterpret_cast<int*>(ctx->d_scratch + PSLAM_SOLVER_SCRATCH_OFFSET);
  // fused solver iterations on the correspondences just found: their block sits directly behind the filter's, ONE download
  size_t gn_words = 0;  // 4-byte words: [done, spd | pad to 8 B][n_iters x 16 doubles][status bytes, padded]
  const size_t gn_off = (n_words + 1) & ~(size_t) 1;  // 8-byte aligned
  int n_gn_iters = 0;
  CompactArgs ca{};
  if (gn) {
    if (!gn->factor || gn->n_iterations < 0) return pslam_set_error(ctx, PSLAM_E_INVALID, "match_gn: factor configuration missing", cudaSuccess);
    n_gn_iters = gn->n_iterations;
    gn_words = 2 + 32 * (size_t) n_gn_iters + ((size_t) n_fixed + 3) / 4;
    if (PSLAM_SOLVER_SCRATCH_OFFSET + 4 * (gn_off + gn_words) > ctx->scratch_bytes)
      return pslam_set_error(ctx, PSLAM_E_CAPACITY, "match_gn: scratch too small", cudaSuccess);
    ca = CompactArgs{1, ctx->proj_weights_epoch == ctx->proj_moving_epoch ? st.d_mscale : nullptr, gn->diagonal_info[0],
                     gn->diagonal_info[1], gn->diagonal_info[2], st.d_gn_cf, st.d_gn_cm, st.d_gn_info, st.d_gn_ncorr, nullptr,
                     PslamAlignCfg{}};
  }
  pslam_launch_pdl(n_fixed <= FF_SMEM_FIXED ? filter_fused_kernel<true> : filter_fused_kernel<false>, dim3(1), dim3(FF_THREADS), 0,
                   ctx->stream, st.d_cand, n_moving, n_fixed, st.d_key1, st.d_key2, cfg->descriptor_distance,
                   cfg->maximum_distance_ratio_to_second_best, d_out, ca);
  PSLAM_LAUNCH_CHECK(ctx, "filter_fused_kernel");
  if (gn) {
    int* d_done = d_out + gn_off;
    double* d_gn_out = reinterpret_cast<double*>(d_done + 2);
    uint8_t* d_status = reinterpret_cast<uint8_t*>(d_gn_out + 16 * (size_t) n_gn_iters);
    if ((rc = pslam_k_gn_iterate_dev(ctx, gn->factor, n_gn_iters, gn->damping, gn->pose12, st.d_moving, st.d_fixed, ctx->proj_fixed_dim,
                                     st.d_gn_ncorr, st.d_gn_cf, st.d_gn_cm, st.d_gn_info, gn->prior, d_gn_out, d_done, d_status)))
      return rc;
  }
  const size_t all_words = gn ? gn_off + gn_words : n_words;
  std::vector<int> pageable;
  int* h_out = reinterpret_cast<int*>(ctx->h_pinned);
  if (4 * all_words > ctx->pinned_bytes) {
    pageable.resize(all_words);
    h_out = pageable.data();
  }
  PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(h_out, d_out, 4 * all_words, cudaMemcpyDeviceToHost, ctx->stream));
  PSLAM_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return proj_collect(h_out, n_fixed, n_moving, capacity, h_fixed, h_moving, h_dist, n_projected, gn, gn_off, n_gn_iters, -1, 0);
}

// The frame's whole registration on the device (see pslam_cuda.h): phases of { search, filter, compaction + the finder's
// decisions, solver iterations } are queued a batch at a time; every kernel of a phase returns at once when the state says
// the registration has stopped, so a batch costs one download whatever happens inside it.  The first batch holds exactly the
// phases that MUST happen before the finder is allowed to converge (its iteration counter advances deterministically until
// then) plus the first one that may; later batches hold two.
int pslam_k_projective_align(pslam_ctx* ctx, int n_fixed, int n_moving, const pslam_projective_cfg* cfg, pslam_align* al,
                             int capacity, int* h_fixed, int* h_moving, float* h_dist, pslam_fused_gn* gn) {
  if (!al || !gn || !gn->factor) return pslam_set_error(ctx, PSLAM_E_INVALID, "align: configuration missing", cudaSuccess);
  if (al->max_iterations <= 0 || gn->n_iterations < al->max_iterations || al->has_converged || al->solver_iterations_per_projection <= 0)
    return pslam_set_error(ctx, PSLAM_E_INVALID, "align: bad iteration budget or finder state", cudaSuccess);
  if (n_fixed != ctx->proj_n_fixed || n_moving != ctx->proj_n_moving)
    return pslam_set_error(ctx, PSLAM_E_INVALID, "projective: cloud sizes differ from the uploaded clouds", cudaSuccess);
  ProjState st;
  int rc = proj_layout(ctx, st, n_fixed, 2, n_moving);
  if (rc) return rc;
  al->stop_reason = 0;
  al->iterations_done = 0;
  al->converged_with_good_ratio = 0;
  al->n_projected = 0;
  al->n_phases = 0;
  gn->iterations_done = 0;
  gn->spd = 1;
  if (n_fixed == 0 || n_moving == 0) {
    al->stop_reason = 3;
    return 0;
  }
  ProjParams pp;
  if ((rc = proj_params(ctx, cfg, nullptr, pp, n_fixed))) return rc;
  const size_t smem = search_smem(pp.radius, n_fixed);
  const int rows = al->max_iterations;
  const size_t state_words = (sizeof(PslamAlignState) + 255) / 256 * 64;
  const size_t n_words = 1 + 2 * (size_t) n_fixed + 4 * (size_t) n_moving;
  const size_t gn_off = (n_words + 1) & ~(size_t) 1;
  const size_t gn_words = 2 + 32 * (size_t) rows + ((size_t) n_fixed + 3) / 4;
  const size_t all_words = state_words + gn_off + gn_words;
  if (PSLAM_SOLVER_SCRATCH_OFFSET + 4 * all_words > ctx->scratch_bytes)
    return pslam_set_error(ctx, PSLAM_E_CAPACITY, "align: scratch too small", cudaSuccess);
  int* d_base = reinterpret_cast<int*>(ctx->d_scratch + PSLAM_SOLVER_SCRATCH_OFFSET);
  PslamAlignState* d_state = reinterpret_cast<PslamAlignState*>(d_base);
  int* d_out = d_base + state_words;
  int* d_done = d_out + gn_off;
  double* d_gn_out = reinterpret_cast<double*>(d_done + 2);
  uint8_t* d_status = reinterpret_cast<uint8_t*>(d_gn_out + 16 * (size_t) rows);
  std::vector<int> pageable;
  int* h_all = reinterpret_cast<int*>(ctx->h_pinned);
  if (4 * all_words > ctx->pinned_bytes) {
    pageable.resize(all_words);
    h_all = pageable.data();
  }
  PslamAlignState* h_state = reinterpret_cast<PslamAlignState*>(h_all);
  AlignInit init;
  for (int i = 0; i < 12; ++i) {
    init.estimate[i] = gn->pose12[i];
    init.prev[i] = al->previous12[i];
  }
  init.current_iteration = al->current_iteration;
  init.valid = 1;
  ProjParams pp_first = pp;  // the first search takes the start estimate by value (fp32, as finder.setLocalMapInSensor(X) rounds it)
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) pp_first.R[3 * i + j] = (float) gn->pose12[4 * i + j];
    pp_first.t[i] = (float) gn->pose12[4 * i + 3];
  }
  PslamAlignCfg a;
  a.max_iterations = al->max_iterations;
  a.per_projection = al->solver_iterations_per_projection;
  a.min_iterations = al->minimum_number_of_iterations;
  a.can_widen = al->can_widen_search;
  a.min_corr = al->min_num_correspondences;
  a.max_change_norm = al->maximum_estimate_change_norm_for_convergence;
  a.min_matching_ratio = al->minimum_matching_ratio;
  CompactArgs ca{1, ctx->proj_weights_epoch == ctx->proj_moving_epoch ? st.d_mscale : nullptr, gn->diagonal_info[0],
                 gn->diagonal_info[1], gn->diagonal_info[2], st.d_gn_cf, st.d_gn_cm, st.d_gn_info, st.d_gn_ncorr, d_state, a, init};
  int first_batch = 0;
  for (int ci = al->current_iteration, it = 0; first_batch < 16;) {
    ++first_batch;
    if (ci > a.min_iterations) break;  // this phase may converge and spend the whole budget
    int n_fused = a.max_iterations - it, quiet = 0;
    for (int k = ci + 1; !(k % a.per_projection == 0 || k == 1) && quiet < n_fused; ++k) ++quiet;
    if (quiet + 1 < n_fused) n_fused = quiet + 1;
    it += n_fused;
    ci += n_fused;
    if (it >= a.max_iterations) break;
  }
  for (int queued = 0; queued < PSLAM_ALIGN_MAX_PHASES;) {
    int batch = queued == 0 ? first_batch : 2;
    if (batch > PSLAM_ALIGN_MAX_PHASES - queued) batch = PSLAM_ALIGN_MAX_PHASES - queued;  // the phase log has that many rows
    queued += batch;
    for (int p = 0; p < batch; ++p) {
      const bool first = ca.init.valid != 0;
      pslam_launch_pdl(projective_search_kernel, dim3((n_moving + PJ_WARPS - 1) / PJ_WARPS), dim3(PJ_WARPS * 32), smem, ctx->stream,
                       first ? pp_first : pp, st.d_moving, n_moving, st.d_desc_moving, st.d_lattice, n_fixed, st.d_desc_fixed, st.d_cand,
                       nullptr, st.d_fixed, ctx->proj_fixed_dim, first ? nullptr : d_state, n_fixed <= PJ_STAGE_MAX ? 1 : 0);
      PSLAM_LAUNCH_CHECK(ctx, "projective_search_kernel");
      pslam_launch_pdl(n_fixed <= FF_SMEM_FIXED ? filter_fused_kernel<true> : filter_fused_kernel<false>, dim3(1), dim3(FF_THREADS), 0,
                       ctx->stream, st.d_cand, n_moving, n_fixed, st.d_key1, st.d_key2, cfg->descriptor_distance,
                       cfg->maximum_distance_ratio_to_second_best, d_out, ca);
      PSLAM_LAUNCH_CHECK(ctx, "filter_fused_kernel");
      if ((rc = pslam_k_gn_iterate_dev(ctx, gn->factor, 0, gn->damping, gn->pose12, st.d_moving, st.d_fixed, ctx->proj_fixed_dim,
                                       st.d_gn_ncorr, st.d_gn_cf, st.d_gn_cm, st.d_gn_info, gn->prior, d_gn_out, d_done, d_status,
                                       d_state, al->max_iterations)))
        return rc;
      ca.init.valid = 0;
    }
    PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(h_all, d_base, 4 * all_words, cudaMemcpyDeviceToHost, ctx->stream));
    PSLAM_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (h_state->stop) break;
  }
  al->stop_reason = h_state->stop;
  al->iterations_done = h_state->it;
  al->converged_with_good_ratio = h_state->converged_ratio_ok;
  al->n_projected = h_state->n_projected;
  al->n_phases = h_state->phases;
  al->current_iteration = h_state->current_iteration;
  al->has_converged = h_state->has_converged;
  for (int i = 0; i < 12; ++i) al->previous12[i] = h_state->prev[i];
  memcpy(al->phase_log, h_state->phase_log, sizeof(al->phase_log));
  for (int i = 0; i < 12; ++i) gn->pose12[i] = h_state->estimate[i];
  if (h_state->phases == 0) return 0;  // the first phase already needs the caller: nothing to report
  return proj_collect(h_all + state_words, n_fixed, n_moving, capacity, h_fixed, h_moving, h_dist, nullptr, gn, gn_off, rows, h_state->it,
                      h_state->stop == 4 ? 0 : 1);
}
