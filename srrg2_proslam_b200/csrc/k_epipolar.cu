// k_epipolar.cu -- stage 2a (sm_100a): stereo epipolar matching + stereo-point assembly.
//
// Replaces CorrespondenceFinderDescriptorBasedEpipolar::compute
//   (.../correspondence_finders/correspondence_finder_descriptor_based_epipolar_impl.cpp:44-219)
// and the assembly loop of RawDataPreprocessorStereoProjective::compute
//   (.../sensor_processing/raw_data_preprocessor_stereo_projective.cpp:105-132).
//
// One CTA per stereo pair.  Both feature sets are sorted by (row, col) in shared memory (keys are unique, so any
// sort gives the reference's std::sort order, :36-41): counting sort by row for the device pipeline, bitonic sort
// for arbitrary host coordinates.  The reference's single running `index_right` only couples left features of the
// SAME image row (it is reset to the first right feature of the next row by the two skip loops, :97-126), so the
// row runs are independent: one warp per run, descriptors of the run in registers, shuffle argmin with the
// reference's tie rules.  Results are emitted in the reference's scan order by a block-wide scan.
#include <float.h>

#include "pslam_internal.cuh"
#include "pslam_kernels.cuh"

namespace {

constexpr int EP_THREADS = 256;
constexpr int EP_ROWS_LEAN = 512;   // lean variant: image rows <= 510 (KITTI 376, ICL / EuRoC 480)
constexpr int EP_ROWS = 4096;  // counting-sort path: 0 <= row < EP_ROWS (the device pipeline: row < max_rows)

// ---- (row, col) sort, general path: shared-memory bitonic sort of 64-bit keys (any 32-bit row, 16-bit col) ----
template <int NT>
__device__ __forceinline__ void ep_sort_side(const float2* __restrict__ xy, int n, int P,
                                             unsigned long long* keys, short* row, short* col,
                                             short* idx) {
  const int tid = threadIdx.x;
  for (int i = tid; i < P; i += NT) {
    unsigned long long k = ~0ULL;
    if (i < n) {
      const float2 p = xy[i];
      // Feature(row = int32(y), col = int32(x))  (epipolar_impl.cpp:8-20)
      const unsigned r = (unsigned) (int) p.y, c = (unsigned) (int) p.x;
      k = ((unsigned long long) r << 32) | ((unsigned long long) c << 16) | (unsigned) i;
    }
    keys[i] = k;
  }
  __syncthreads();
  for (int k = 2; k <= P; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < P; i += NT) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const unsigned long long a = keys[i], b = keys[ixj];
          const bool asc = (i & k) == 0;
          if ((a > b) == asc) {
            keys[i] = b;
            keys[ixj] = a;
          }
        }
      }
      __syncthreads();
    }
  }
  for (int i = tid; i < n; i += NT) {
    const unsigned long long k = keys[i];
    row[i] = (short) (k >> 32);
    col[i] = (short) ((k >> 16) & 0xffffu);
    idx[i] = (short) (k & 0xffffu);
  }
  __syncthreads();
}

// ---- (row, col) sort, pipeline path: counting sort by row + per-row insertion sort by column.  Keys are unique,
// so the result is the order std::sort gives the reference (epipolar_impl.cpp:36-41).
// hist / cursor: ROWS + 1 unsigned shorts each; tmp: n uint32 ((col << 16) | idx)
template <int ROWS, int NT>
__device__ __forceinline__ void ep_count_sort_side(const float2* __restrict__ xy, int n, unsigned short* hist,
                                                   unsigned short* cursor, uint32_t* tmp, short* row, short* col,
                                                   short* idx, int* s_warp) {
  const int tid = threadIdx.x;
  for (int r = tid; r <= ROWS; r += NT) hist[r] = 0;
  __syncthreads();
  // 16-bit shared atomics do not exist: count into the aligned 32-bit word that holds the pair
  unsigned* hist32 = reinterpret_cast<unsigned*>(hist);
  for (int i = tid; i < n; i += NT) {
    const int r = (int) xy[i].y;
    atomicAdd(&hist32[r >> 1], (r & 1) ? 0x10000u : 1u);
  }
  __syncthreads();
  // exclusive scan over rows: EP_ROWS / NT consecutive rows per thread
  constexpr int PER = ROWS / NT;
  int local[PER], sum = 0;
#pragma unroll
  for (int k = 0; k < PER; ++k) {
    local[k] = sum;
    sum += hist[tid * PER + k];
  }
  int total;
  const int base = block_exclusive_scan<NT>(sum, s_warp, &total);
#pragma unroll
  for (int k = 0; k < PER; ++k) {
    hist[tid * PER + k] = (unsigned short) (base + local[k]);  // start of row
    cursor[tid * PER + k] = (unsigned short) (base + local[k]);
  }
  if (tid == 0) hist[ROWS] = (unsigned short) total;
  __syncthreads();
  unsigned* cur32 = reinterpret_cast<unsigned*>(cursor);
  for (int i = tid; i < n; i += NT) {
    const float2 p = xy[i];
    const int r = (int) p.y, c = (int) p.x;
    const unsigned old = atomicAdd(&cur32[r >> 1], (r & 1) ? 0x10000u : 1u);
    const int pos = (r & 1) ? (int) (old >> 16) : (int) (old & 0xffffu);
    tmp[pos] = ((unsigned) c << 16) | (unsigned) i;
    row[pos] = (short) r;  // every slot of a row's range holds that row
  }
  __syncthreads();
  // order inside a row: rank of (col, idx) among the row's entries (unique values), one thread per feature
  for (int i = tid; i < n; i += NT) {
    const int r = row[i];
    const int b = hist[r], e = hist[r + 1];
    const uint32_t v = tmp[i];
    int rank = 0;
    for (int j = b; j < e; ++j) rank += tmp[j] < v ? 1 : 0;
    col[b + rank] = (short) (v >> 16);
    idx[b + rank] = (short) (v & 0xffffu);
  }
  __syncthreads();
}

// ordered in-place removal of flagged entries from (row, col, idx); tmp >= 3*n shorts
template <int NT>
__device__ __forceinline__ int ep_compact(short* row, short* col, short* idx, int n,
                                          const unsigned char* removed, short* tmp, int* s_warp) {
  const int tid = threadIdx.x;
  int running = 0;
  for (int base = 0; base < n; base += NT) {
    const int i = base + tid;
    const int keep = (i < n && !removed[i]) ? 1 : 0;
    int total;
    const int off = block_exclusive_scan<NT>(keep, s_warp, &total);
    if (keep) {
      const int o = running + off;
      tmp[3 * o] = row[i];
      tmp[3 * o + 1] = col[i];
      tmp[3 * o + 2] = idx[i];
    }
    running += total;
  }
  __syncthreads();
  for (int i = tid; i < running; i += NT) {
    row[i] = tmp[3 * i];
    col[i] = tmp[3 * i + 1];
    idx[i] = tmp[3 * i + 2];
  }
  __syncthreads();
  return running;
}

// GENERAL = true : any coordinates (host-pointer entry point), bitonic sort
// GENERAL = false: 0 <= row < EP_ROWS (device pipeline), counting sort -- no 64-bit key array in shared memory
//
// Matching: the reference's running `index_right` only couples left features of the SAME image row (it is reset to
// the first right feature of the next row by the two skip loops, :97-126).  One WARP owns a left row run: lane t
// holds the descriptor of the run's t-th left and t-th right feature in registers (one round of loads per run);
// the left features are visited in order, each lane scores "its" right candidate and a 5-step shuffle butterfly
// yields (best, second, first position of the best) with the reference's tie rules.
// LEAN (pipeline path, thickness 0, image rows <= EP_ROWS_LEAN - 2): one pass, so nothing is ever removed -- no used
// flags, no row array of the right side after its sort, run starts bounded by the image height, and a histogram of
// EP_ROWS_LEAN instead of EP_ROWS rows: 16 instead of 26 bytes of shared memory per feature, 3 instead of 2 CTAs per SM.
template <bool GENERAL, bool LEAN, int NT>
__global__ void __launch_bounds__(NT)
epipolar_kernel(const float2* __restrict__ xy, const uint32_t* __restrict__ desc,
                const int* __restrict__ count, int M, float max_dist, float max_ratio, int max_disp,
                int thickness, int* __restrict__ ep_fixed, int* __restrict__ ep_moving,
                float* __restrict__ ep_dist, int* __restrict__ ep_count,
                float4* __restrict__ st_uvuv, int* __restrict__ st_left, int* __restrict__ st_right,
                float* __restrict__ st_dist, int* __restrict__ st_count, int pair_base) {
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ int s_warp[33];
  __shared__ int s_nruns;
  __shared__ int s_next_run;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int pair = pair_base + blockIdx.x;
  const int imgL = 2 * pair, imgR = 2 * pair + 1;
  int nL = count[imgL], nR = count[imgR];

  // shared memory: [sort scratch] | rowL colL idxL rowR colR idxR | match dist | usedR usedL | runs
  // LEAN:          [hist cursor]  | rowL colL idxL colR idxR | match (= rowR during the sorts) dist | runs[ROWS]
  constexpr int ROWS = LEAN ? EP_ROWS_LEAN : EP_ROWS;
  int Pmax = 1;
  while (Pmax < M) Pmax <<= 1;
  size_t scratch = GENERAL ? (size_t) Pmax * 8 : (size_t) (ROWS + 2) * 2 * 2;
  if (!LEAN && scratch < (size_t) M * 6) scratch = (size_t) M * 6;  // ep_compact needs 3 * M shorts (same rule as ep_smem_bytes)
  scratch = (scratch + 15) & ~(size_t) 15;
  short* rowL = reinterpret_cast<short*>(smem + scratch);
  short* colL = rowL + M;
  short* idxL = colL + M;
  short* rowR = LEAN ? nullptr : idxL + M;
  short* colR = LEAN ? idxL + M : rowR + M;
  short* idxR = colR + M;
  short* match = idxR + M;
  unsigned short* dist = reinterpret_cast<unsigned short*>(match + M);
  unsigned char* usedR = LEAN ? nullptr : reinterpret_cast<unsigned char*>(dist + M);
  unsigned char* usedL = LEAN ? nullptr : usedR + M;
  short* runs = LEAN ? reinterpret_cast<short*>(dist + M) : reinterpret_cast<short*>(usedL + M);  // starts of the left row runs
  // LEAN: the sort needs a row array per side and M words of `tmp`; tmp = (match, dist) as before, the right side's
  // rows go to the tail of the run-start array's region, which is sized for them (ep_smem_bytes)
  if (LEAN) rowR = runs;

  const float2* xyL = xy + (size_t) imgL * M;
  const float2* xyR = xy + (size_t) imgR * M;
  const uint4* descL = reinterpret_cast<const uint4*>(desc + (size_t) imgL * M * 8);
  const uint4* descR = reinterpret_cast<const uint4*>(desc + (size_t) imgR * M * 8);

  if (GENERAL) {
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(smem);
    int P = 1;
    while (P < max(nL, nR)) P <<= 1;
    ep_sort_side<NT>(xyL, nL, P, keys, rowL, colL, idxL);
    ep_sort_side<NT>(xyR, nR, P, keys, rowR, colR, idxR);
  } else {
    unsigned short* hist = reinterpret_cast<unsigned short*>(smem);
    unsigned short* cursor = hist + ROWS + 2;
    uint32_t* tmp = reinterpret_cast<uint32_t*>(match);  // match + dist = 4 M bytes, not in use yet
    ep_count_sort_side<ROWS, NT>(xyL, nL, hist, cursor, tmp, rowL, colL, idxL, s_warp);
    ep_count_sort_side<ROWS, NT>(xyR, nR, hist, cursor, tmp, rowR, colR, idxR, s_warp);
  }

  int* o_fixed = ep_fixed + (size_t) pair * M;
  int* o_moving = ep_moving + (size_t) pair * M;
  float* o_dist = ep_dist + (size_t) pair * M;
  int n_out = 0;

  const int n_offsets = 2 * thickness + 1;
  for (int oi = 0; oi < n_offsets; ++oi) {
    // row offsets 0, +1, -1, +2, -2, ...  (epipolar_impl.cpp:72-79)
    const int off = (oi == 0) ? 0 : ((oi & 1) ? (oi + 1) / 2 : -(oi / 2));
    for (int i = tid; i < nL; i += NT) {
      match[i] = -1;
      if (!LEAN) usedL[i] = 0;
    }
    if (!LEAN)
      for (int i = tid; i < nR; i += NT) usedR[i] = 0;
    // ordered list of the left row-run starts
    int n_runs = 0;
    for (int base = 0; base < nL; base += NT) {
      const int i = base + tid;
      const int is_start = (i < nL && (i == 0 || rowL[i] != rowL[i - 1])) ? 1 : 0;
      int total;
      const int o = block_exclusive_scan<NT>(is_start, s_warp, &total);
      if (is_start) runs[n_runs + o] = (short) i;
      n_runs += total;
    }
    if (tid == 0) s_next_run = 0;
    __syncthreads();
    if (nR > 0) {
      // row runs are handed out dynamically (their cost varies with the row's population: with a static round-robin
      // a quarter of the kernel's stall samples sat at the barrier below)
      while (true) {
        int run = 0;
        if (lane == 0) run = atomicAdd(&s_next_run, 1);
        run = __shfl_sync(0xffffffffu, run, 0);
        if (run >= n_runs) break;
        const int i0 = runs[run];
        const int i1 = run + 1 < n_runs ? (int) runs[run + 1] : nL;
        const int row_left = rowL[i0] + off;
        int s0, s1;
        if (!GENERAL && oi == 0) {
          // first pass of the pipeline path: the counting sort of the RIGHT side left its row starts in the scratch
          // (nothing has been removed yet), no search needed
          if (row_left < 0 || row_left >= ROWS) continue;
          const unsigned short* row_start = reinterpret_cast<const unsigned short*>(smem);
          s0 = row_start[row_left];
          s1 = row_start[row_left + 1];
          if (s0 == s1) continue;
        } else {
          int lo = 0, hi = nR;  // first right feature with row >= row_left
          while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (rowR[mid] < row_left) lo = mid + 1; else hi = mid;
          }
          s0 = lo;
          if (s0 >= nR || rowR[s0] != row_left) continue;
          hi = nR;  // first right feature with row > row_left
          while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (rowR[mid] <= row_left) lo = mid + 1; else hi = mid;
          }
          s1 = lo;
        }
        const int nr = s1 - s0;
        // right side: up to 32 features of the row live in registers (lane t <-> s0 + t); longer rows are read
        // from L1 / L2 per left feature, 32 candidates at a time
        const bool cached = nr <= 32;
        uint4 b0 = make_uint4(0, 0, 0, 0), b1 = b0;
        int cr = 0;
        if (cached && lane < nr) {
          const int f = idxR[s0 + lane];
          b0 = __ldg(descR + 2 * f);
          b1 = __ldg(descR + 2 * f + 1);
          cr = colR[s0 + lane];
        }
        int ir = s0;  // the reference's index_right
        for (int jb = i0; jb < i1 && ir < s1; jb += 32) {
          // left side: 32 features of the run at a time in registers (lane t <-> jb + t)
          uint4 a0 = make_uint4(0, 0, 0, 0), a1 = a0;
          int cl = 0;
          if (jb + lane < i1) {
            const int f = idxL[jb + lane];
            a0 = __ldg(descL + 2 * f);
            a1 = __ldg(descL + 2 * f + 1);
            cl = colL[jb + lane];
          }
          const int nj = min(32, i1 - jb);
          for (int jj = 0; jj < nj; ++jj) {
            if (ir >= s1) break;  // :128-131
            const int col_left = __shfl_sync(0xffffffffu, cl, jj);
            uint4 q0, q1;
            q0.x = __shfl_sync(0xffffffffu, a0.x, jj); q0.y = __shfl_sync(0xffffffffu, a0.y, jj);
            q0.z = __shfl_sync(0xffffffffu, a0.z, jj); q0.w = __shfl_sync(0xffffffffu, a0.w, jj);
            q1.x = __shfl_sync(0xffffffffu, a1.x, jj); q1.y = __shfl_sync(0xffffffffu, a1.y, jj);
            q1.z = __shfl_sync(0xffffffffu, a1.z, jj); q1.w = __shfl_sync(0xffffffffu, a1.w, jj);
            // candidates: s >= index_right with 0 <= col_left - colR[s] <= max_disp.  The right run is sorted by
            // column, so the reference's "disparity < 0 => break" drops exactly the candidates with disparity < 0.
            int best = INT_MAX, second = INT_MAX, pos = INT_MAX;
            if (cached) {
              const int s = s0 + lane;
              const int disparity = col_left - cr;
              if (s >= ir && lane < nr && disparity >= 0 && disparity <= max_disp) {
                best = hamming256(q0, q1, b0, b1);
                pos = s;
              }
            } else {
              int lo2 = ir, hi2 = s1;  // skip the features left of the disparity window
              while (lo2 < hi2) {
                const int mid = (lo2 + hi2) >> 1;
                if (col_left - (int) colR[mid] > max_disp) lo2 = mid + 1; else hi2 = mid;
              }
              for (int sb = lo2; sb < s1; sb += 32) {
                const int s = sb + lane;
                const int disparity = s < s1 ? col_left - (int) colR[s] : -1;
                if (disparity >= 0 && disparity <= max_disp) {
                  const int f = idxR[s];
                  const int d = hamming256(q0, q1, __ldg(descR + 2 * f), __ldg(descR + 2 * f + 1));
                  if (d < best) {  // this lane's candidates arrive in scan order
                    second = best;
                    best = d;
                    pos = s;
                  } else if (d < second) {
                    second = d;
                  }
                }
                if (__any_sync(0xffffffffu, disparity < 0)) break;
              }
            }
            // warp argmin with the reference's tie rule (equal distance: the first position wins) as ONE hardware
            // reduction over (distance << 16 | position); the second best distance is the minimum over the other lanes'
            // best and the winner lane's own second (REDUX.MIN; a 5-step shuffle butterfly cost 5x as many instructions)
            {
              const unsigned key = best == INT_MAX ? 0xffffffffu : (((unsigned) best << 16) | (unsigned) pos);
              const unsigned kmin = __reduce_min_sync(0xffffffffu, key);
              const unsigned other = (key == kmin) ? (unsigned) second : (unsigned) best;
              const unsigned smin = __reduce_min_sync(0xffffffffu, other);
              if (kmin == 0xffffffffu) {
                best = INT_MAX;
              } else {
                best = (int) (kmin >> 16);
                pos = (int) (kmin & 0xffffu);
                second = (int) smin;
              }
            }
            if (best == INT_MAX) continue;
            const float fb = (float) best;
            const float fs = (second == INT_MAX) ? FLT_MAX : (float) second;
            if (fb < max_dist && __fdiv_rn(fb, fs) < max_ratio) {  // :171-173
              if (lane == 0) {
                match[jb + jj] = (short) pos;
                dist[jb + jj] = (unsigned short) best;
                if (!LEAN) {
                  usedL[jb + jj] = 1;
                  usedR[pos] = 1;
                }
              }
              ir = pos + 1;  // ordering constraint :181
            }
          }
        }
      }
    }
    __syncthreads();
    // emit this pass's matches in scan order
    for (int base = 0; base < nL; base += NT) {
      const int i = base + tid;
      const int has = (i < nL && match[i] >= 0) ? 1 : 0;
      int total;
      const int o = block_exclusive_scan<NT>(has, s_warp, &total);
      if (has) {
        const int k = n_out + o;
        if (k < M) {
          o_fixed[k] = idxL[i];
          o_moving[k] = idxR[match[i]];
          o_dist[k] = (float) dist[i];
        }
      }
      n_out += total;
    }
    if (!LEAN && oi + 1 < n_offsets) {  // prune matched candidates, keeping the order (:189-205)
      __syncthreads();
      short* ctmp = reinterpret_cast<short*>(smem);  // the sort scratch holds >= 3 * M shorts (ep_smem_bytes)
      nL = ep_compact<NT>(rowL, colL, idxL, nL, usedL, ctmp, s_warp);
      nR = ep_compact<NT>(rowR, colR, idxR, nR, usedR, ctmp, s_warp);
    }
  }
  if (n_out > M) n_out = M;
  __syncthreads();
  if (tid == 0) ep_count[pair] = n_out;
  __threadfence_block();
  __syncthreads();

  // stereo points (uL, vL, uR, vR); negative disparities dropped (stereo_projective.cpp:120-128)
  float4* s_uvuv = st_uvuv + (size_t) pair * M;
  int* s_l = st_left + (size_t) pair * M;
  int* s_r = st_right + (size_t) pair * M;
  float* s_d = st_dist + (size_t) pair * M;
  int n_st = 0;
  for (int base = 0; base < n_out; base += NT) {
    const int k = base + tid;
    int keep = 0, f = 0, m = 0;
    float4 p = make_float4(0, 0, 0, 0);
    if (k < n_out) {
      f = o_fixed[k];
      m = o_moving[k];
      const float2 l = xyL[f], r = xyR[m];
      p = make_float4(l.x, l.y, r.x, r.y);
      keep = !((l.x - r.x) < 0.0f || (l.y - r.y) < 0.0f);
    }
    int total;
    const int o = block_exclusive_scan<NT>(keep, s_warp, &total);
    if (keep) {
      s_uvuv[n_st + o] = p;
      s_l[n_st + o] = f;
      s_r[n_st + o] = m;
      s_d[n_st + o] = o_dist[k];
    }
    n_st += total;
  }
  if (tid == 0) st_count[pair] = n_st;
}

// ---- batch result packing: per-pair padded stores -> CSR (offsets + SoA), the measurement cloud the
// adaptor hands to the tracker: (uL,vL,uR,vR) + the LEFT feature's intensity and descriptor
// (raw_data_preprocessor_stereo_projective.cpp:112-118)
__global__ void __launch_bounds__(1024)
stereo_offsets_kernel(const int* __restrict__ st_count, int n_pairs, long long* __restrict__ offsets) {
  __shared__ int s_warp[33];
  long long running = 0;
  for (int base = 0; base < n_pairs; base += 1024) {
    const int i = base + threadIdx.x;
    const int v = i < n_pairs ? st_count[i] : 0;
    int total;
    const int off = block_exclusive_scan<1024>(v, s_warp, &total);
    if (i < n_pairs) offsets[i] = running + off;
    running += total;
  }
  if (threadIdx.x == 0) offsets[n_pairs] = running;
}

__global__ void __launch_bounds__(256)
stereo_pack_kernel(const long long* __restrict__ offsets, const int* __restrict__ st_count, int M,
                   const float4* __restrict__ st_uvuv, const int* __restrict__ st_left,
                   const int* __restrict__ st_right, const float* __restrict__ st_dist,
                   const float* __restrict__ inten, const uint32_t* __restrict__ desc,
                   float4* __restrict__ o_uvuv, float* __restrict__ o_int, uint4* __restrict__ o_desc,
                   int* __restrict__ o_left, int* __restrict__ o_right, float* __restrict__ o_dist) {
  const int pair = blockIdx.x;
  const int n = st_count[pair];
  const long long o = offsets[pair];
  const size_t src = (size_t) pair * M;
  const size_t left_slot = (size_t) (2 * pair) * M;
  const uint4* d4 = reinterpret_cast<const uint4*>(desc);
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int l = st_left[src + i];
    o_uvuv[o + i] = st_uvuv[src + i];
    o_left[o + i] = l;
    o_right[o + i] = st_right[src + i];
    o_dist[o + i] = st_dist[src + i];
    o_int[o + i] = inten[left_slot + l];
    o_desc[2 * (o + i)] = d4[2 * (left_slot + l)];
    o_desc[2 * (o + i) + 1] = d4[2 * (left_slot + l) + 1];
  }
}


// ---- N1: TriangulatorRigidStereo::compute / triangulateRectifiedMidpoint ---------------------------
//   .../mapping/triangulator_rigid_stereo.cpp:7-85.  fp32, operation order of the reference (no FMA contraction):
//   depth = b_x / (xL - xR) (infinity depth at zero disparity), x = 1 / f_x * (xL - c_x) * depth,
//   y = 1 / f_y * ((yL + yR) / 2 - c_y) * depth; points with xL - xR < minimum_disparity are kept as INVALID
//   placeholders (coordinates 0) so that the output stays index-aligned with the input (:38-45).
__global__ void triangulate_kernel(const float4* __restrict__ uvuv, long long n, float fx, float fy, float cx, float cy,
                                   float b_x, float min_disparity, float infinity_depth, float* __restrict__ xyz,
                                   unsigned char* __restrict__ valid) {
  const long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 p = uvuv[i];
  float x = 0.f, y = 0.f, z = 0.f;
  unsigned char ok = 0;
  if (!(__fsub_rn(p.x, p.z) < min_disparity)) {
    float depth = infinity_depth;
    if (p.x > p.z) depth = __fdiv_rn(b_x, __fsub_rn(p.x, p.z));
    z = depth;
    x = __fmul_rn(__fmul_rn(__fdiv_rn(1.f, fx), __fsub_rn(p.x, cx)), depth);
    y = __fmul_rn(__fmul_rn(__fdiv_rn(1.f, fy), __fsub_rn(__fdiv_rn(__fadd_rn(p.y, p.w), 2.f), cy)), depth);
    ok = 1;
  }
  xyz[3 * i] = x;
  xyz[3 * i + 1] = y;
  xyz[3 * i + 2] = z;
  valid[i] = ok;
}

}  // namespace

int pslam_k_pack_stereo(pslam_ctx* ctx, int n_pairs, pslam_packed_stereo* out) {
  const size_t M = ctx->lim.max_features;
  const size_t cap = (size_t) n_pairs * M;  // worst case
  uint8_t* p = ctx->d_scratch;
  auto carve = [&](size_t bytes) {
    uint8_t* r = p;
    p += (bytes + 255) & ~(size_t) 255;
    return r;
  };
  out->d_offsets = (long long*) carve(sizeof(long long) * ((size_t) n_pairs + 1));
  out->d_uvuv = (float4*) carve(sizeof(float4) * cap);
  out->d_desc = (uint32_t*) carve(32 * cap);
  out->d_intensity = (float*) carve(sizeof(float) * cap);
  out->d_left = (int*) carve(sizeof(int) * cap);
  out->d_right = (int*) carve(sizeof(int) * cap);
  out->d_dist = (float*) carve(sizeof(float) * cap);
  if ((size_t) (p - ctx->d_scratch) > ctx->scratch_bytes)
    return pslam_set_error(ctx, PSLAM_E_CAPACITY, "pack_stereo: scratch too small for this batch", cudaSuccess);
  stereo_offsets_kernel<<<1, 1024, 0, ctx->stream>>>(ctx->d_st_count, n_pairs, out->d_offsets);
  PSLAM_LAUNCH_CHECK(ctx, "stereo_offsets_kernel");
  stereo_pack_kernel<<<n_pairs, 256, 0, ctx->stream>>>(
    out->d_offsets, ctx->d_st_count, (int) M, ctx->d_st_uvuv, ctx->d_st_left, ctx->d_st_right,
    ctx->d_st_dist, ctx->d_inten, ctx->d_desc, out->d_uvuv, out->d_intensity,
    reinterpret_cast<uint4*>(out->d_desc), out->d_left, out->d_right, out->d_dist);
  PSLAM_LAUNCH_CHECK(ctx, "stereo_pack_kernel");
  return PSLAM_OK;
}

static size_t ep_smem_bytes(int M, bool general, bool lean) {
  int P = 1;
  while (P < M) P <<= 1;
  if (lean) {  // hist + cursor | 5 coordinate arrays, match, dist (shorts) | run starts; the right rows use that last region
    const size_t scratch = (((size_t) (EP_ROWS_LEAN + 2) * 2 * 2) + 15) & ~(size_t) 15;
    return scratch + (size_t) M * (7 * 2) + (size_t) M * 2;
  }
  // sort scratch (also the 3 * M shorts of ep_compact) | 6 coordinate arrays, match, dist (shorts) | usedR, usedL | runs
  size_t scratch = general ? (size_t) P * 8 : (size_t) (EP_ROWS + 2) * 2 * 2;
  if (scratch < (size_t) M * 6) scratch = (size_t) M * 6;
  scratch = (scratch + 15) & ~(size_t) 15;
  return scratch + (size_t) M * (8 * 2 + 2 + 2);
}

int pslam_k_epipolar(pslam_ctx* ctx, int n_pairs, const pslam_match_cfg* cfg, int general, int pair_base) {
  const int M = ctx->lim.max_features;
  if (ctx->lim.max_rows > EP_ROWS) general = 1;
  // a negative thickness still runs the offset-0 pass in the reference (epipolar_impl.cpp:72-79: the offset list starts
  // with 0, the +-k entries are appended in a loop that does not execute)
  const int thickness = cfg->epipolar_line_thickness_pixels < 0 ? 0 : cfg->epipolar_line_thickness_pixels;
  const bool lean = !general && thickness == 0 && ctx->lim.max_rows <= EP_ROWS_LEAN - 2;
  const size_t smem = ep_smem_bytes(M, general != 0, lean);
  // a handful of pairs (the per-frame adaptor: ONE) cannot fill the GPU with one CTA each: twice the warps per CTA take the
  // dynamically handed-out row runs in half the time (61 -> 40 us for one KITTI pair)
  const bool wide = lean && n_pairs <= 32;
  auto kernel = general ? epipolar_kernel<true, false, EP_THREADS>
                        : (lean ? (wide ? epipolar_kernel<false, true, 2 * EP_THREADS> : epipolar_kernel<false, true, EP_THREADS>)
                                : epipolar_kernel<false, false, EP_THREADS>);
  if (smem > 48 * 1024) PSLAM_CUDA_TRY(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
  kernel<<<n_pairs, wide ? 2 * EP_THREADS : EP_THREADS, smem, ctx->stream>>>(
    ctx->d_xy, ctx->d_desc, ctx->d_count, M, cfg->maximum_descriptor_distance,
    cfg->maximum_distance_ratio_to_second_best, cfg->maximum_disparity_pixels,
    thickness, ctx->d_ep_fixed, ctx->d_ep_moving, ctx->d_ep_dist,
    ctx->d_ep_count, ctx->d_st_uvuv, ctx->d_st_left, ctx->d_st_right, ctx->d_st_dist,
    ctx->d_st_count, pair_base);
  PSLAM_LAUNCH_CHECK(ctx, "epipolar_kernel");
  return PSLAM_OK;
}

int pslam_k_triangulate(pslam_ctx* ctx, const float4* d_uvuv, long long n, const float* K9, float b_x, float min_disparity,
                        float infinity_depth, float* d_xyz, unsigned char* d_valid) {
  if (n <= 0) return PSLAM_OK;
  const int threads = 256;
  triangulate_kernel<<<(unsigned) ((n + threads - 1) / threads), threads, 0, ctx->stream>>>(
    d_uvuv, n, K9[0], K9[4], K9[2], K9[5], b_x, min_disparity, infinity_depth, d_xyz, d_valid);
  PSLAM_LAUNCH_CHECK(ctx, "triangulate_kernel");
  return PSLAM_OK;
}
