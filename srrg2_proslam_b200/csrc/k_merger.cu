// k_merger.cu -- N3: the binning of MergerProjective_::compute (sm_100a).
//
// Reference path replaced:
//   MergerProjective_::compute      .../mapping/mergers/merger_projective_impl.cpp:8-171   (update pass + bin blocking)
//   MergerProjective_::_addPoints   .../mapping/mergers/merger_projective_impl.cpp:194-309 (binned addition candidates)
//   _isBetterForAddition            merger_projective.h:89-92 (never), merger_projective_rigid_stereo_impl.cpp:42-52 (larger
//                                   disparity), merger_projective_depth_ekf_impl.cpp:44-52 (smaller depth)
// The reference walks the correspondences / measurements in order through a map of maps (bin row -> bin col -> index):
// the first arrival blocks a bin, later arrivals are skipped (update pass) or replace the occupant when strictly better
// (addition pass).  Both are order statistics per bin and parallelise without changing a single decision:
//   update pass:    a correspondence is processed  <=>  it is the LOWEST correspondence index of its bin among those that pass
//                   the appearance gate  (shared-memory atomicMin per bin);
//   addition pass:  the slot of a bin in the candidate list = rank of its first arrival (atomicMin + block scan), its final
//                   occupant = the earliest measurement that attains the bin's best score (64-bit atomicMax over
//                   (ordered score, ~index)).
// One CTA per frame (a frame has <= a few thousand measurements and 11 x 31 bins); the bin tables live in shared memory.
#include <climits>

#include "pslam_internal.cuh"
#include "pslam_kernels.cuh"

namespace {

constexpr int MG_THREADS = 1024;

struct MergerParams {
  float row_bin_width, col_bin_width;  // canvas / number of bins, fp32 (merger_projective_impl.cpp:31-34)
  int n_row_bins, n_col_bins;
  float max_distance_appearance;
  int enable_binning, kind, dim;
};

// std::round(coordinate / bin width) -> size_t (:82-83, :215-216); -1 when the bin leaves the (n + 1)-wide grid (the
// reference asserts it cannot, :84-85)
__device__ __forceinline__ int merger_bin(const MergerParams& p, const float* m) {
  const float fr = roundf(__fdiv_rn(m[1], p.row_bin_width)), fc = roundf(__fdiv_rn(m[0], p.col_bin_width));
  if (!(fr >= 0.f && fr <= (float) p.n_row_bins && fc >= 0.f && fc <= (float) p.n_col_bins)) return -1;
  return (int) fr * (p.n_col_bins + 1) + (int) fc;
}

// float -> unsigned, order preserving (-0 folded onto +0: the reference compares floats, where they are equal)
__device__ __forceinline__ unsigned ordered_bits(float f) {
  const unsigned u = __float_as_uint(__fadd_rn(f, 0.f));
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__global__ void __launch_bounds__(MG_THREADS)
merger_select_updates_kernel(MergerParams p, const float* __restrict__ meas, int n_meas, const int* __restrict__ corr_moving,
                             const float* __restrict__ corr_response, int n_corr, unsigned char* __restrict__ selected,
                             unsigned* __restrict__ occupied, int n_words, int* __restrict__ result) {
  extern __shared__ int s_first[];  // per bin: lowest correspondence index that blocks it
  __shared__ int s_count, s_bad;
  const int n_bins = (p.n_row_bins + 1) * (p.n_col_bins + 1);
  for (int b = threadIdx.x; b < n_bins; b += MG_THREADS) s_first[b] = INT_MAX;
  if (threadIdx.x == 0) s_count = s_bad = 0;
  __syncthreads();
  if (p.enable_binning) {
    for (int c = threadIdx.x; c < n_corr; c += MG_THREADS) {
      if (corr_response[c] > p.max_distance_appearance) continue;  // :72-75, before the bin is looked at
      const int m = corr_moving[c];
      const int b = (m >= 0 && m < n_meas) ? merger_bin(p, meas + (size_t) m * p.dim) : -1;
      if (b < 0) s_bad = 1;
      else atomicMin(&s_first[b], c);
    }
  }
  __syncthreads();
  int mine = 0;
  for (int c = threadIdx.x; c < n_corr; c += MG_THREADS) {
    bool take = !(corr_response[c] > p.max_distance_appearance);
    if (take && p.enable_binning) {
      const int m = corr_moving[c];
      const int b = (m >= 0 && m < n_meas) ? merger_bin(p, meas + (size_t) m * p.dim) : -1;
      take = b >= 0 && s_first[b] == c;  // later correspondences of a blocked bin are skipped (:94-110)
    }
    selected[c] = take ? 1 : 0;
    mine += take;
  }
  if (mine) atomicAdd(&s_count, mine);
  // blocked bins as a bitmap, bin = row * (n_col_bins + 1) + col
  for (int w = threadIdx.x; w < n_words; w += MG_THREADS) {
    unsigned bits = 0;
    for (int k = 0; k < 32; ++k) {
      const int b = 32 * w + k;
      if (b < n_bins && s_first[b] != INT_MAX) bits |= 1u << k;
    }
    occupied[w] = bits;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    result[0] = s_count;
    result[1] = s_bad;
  }
}

__global__ void __launch_bounds__(MG_THREADS)
merger_select_additions_kernel(MergerParams p, const float* __restrict__ meas, int n_meas, const unsigned* __restrict__ occupied,
                               int* __restrict__ winners, int* __restrict__ result) {
  extern __shared__ unsigned long long s_mem[];
  __shared__ int s_warp[33];
  __shared__ int s_bad;
  const int n_bins = (p.n_row_bins + 1) * (p.n_col_bins + 1);
  unsigned long long* s_best = s_mem;                        // per bin: (ordered score << 32) | ~measurement index
  int* s_first = reinterpret_cast<int*>(s_mem + n_bins);     // per bin: first arrival, then its slot
  if (threadIdx.x == 0) s_bad = 0;
  if (!p.enable_binning) {  // all measurements are candidates (:250-253)
    for (int i = threadIdx.x; i < n_meas; i += MG_THREADS) winners[i] = i;
    if (threadIdx.x == 0) {
      result[0] = n_meas;
      result[1] = 0;
    }
    return;
  }
  for (int b = threadIdx.x; b < n_bins; b += MG_THREADS) {
    s_best[b] = 0ull;
    s_first[b] = INT_MAX;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n_meas; i += MG_THREADS) {
    const float* m = meas + (size_t) i * p.dim;
    const int b = merger_bin(p, m);
    if (b < 0) {
      s_bad = 1;
      continue;
    }
    if (occupied && ((occupied[b >> 5] >> (b & 31)) & 1u)) continue;  // bin taken by a tracked point (:221-227)
    float score = 0.f;                                              // base class: the first arrival stays
    if (p.kind == PSLAM_MERGER_STEREO) score = __fsub_rn(m[0], m[2]);  // larger disparity wins
    if (p.kind == PSLAM_MERGER_DEPTH) score = -m[2];                   // smaller depth wins
    atomicMin(&s_first[b], i);
    atomicMax(&s_best[b], ((unsigned long long) ordered_bits(score) << 32) | (unsigned) ~(unsigned) i);
  }
  __syncthreads();
  // slot of a bin = number of bins whose first arrival came earlier
  int n_out = 0;
  for (int base = 0; base < n_meas; base += MG_THREADS) {
    const int i = base + threadIdx.x;
    int b = -1, is_first = 0;
    if (i < n_meas) {
      b = merger_bin(p, meas + (size_t) i * p.dim);
      is_first = (b >= 0 && s_first[b] == i) ? 1 : 0;
    }
    int total;
    const int o = block_exclusive_scan<MG_THREADS>(is_first, s_warp, &total);
    if (is_first) winners[n_out + o] = (int) ~(unsigned) (s_best[b] & 0xffffffffull);
    n_out += total;
  }
  if (threadIdx.x == 0) {
    result[0] = n_out;
    result[1] = s_bad;
  }
}

MergerParams merger_params(const pslam_merger_cfg* cfg, int dim) {
  MergerParams p;
  p.row_bin_width = (float) cfg->canvas_rows / (float) cfg->number_of_row_bins;
  p.col_bin_width = (float) cfg->canvas_cols / (float) cfg->number_of_col_bins;
  p.n_row_bins = cfg->number_of_row_bins;
  p.n_col_bins = cfg->number_of_col_bins;
  p.max_distance_appearance = cfg->maximum_distance_appearance;
  p.enable_binning = cfg->enable_binning;
  p.kind = cfg->kind;
  p.dim = dim;
  return p;
}

}  // namespace

int pslam_k_merger_select_updates(pslam_ctx* ctx, const pslam_merger_cfg* cfg, const float* d_meas, int dim, int n_meas,
                                  const int* d_corr_moving, const float* d_corr_response, int n_corr, unsigned char* d_selected,
                                  unsigned* d_occupied, int n_words, int* d_result) {
  const size_t smem = (size_t) (cfg->number_of_row_bins + 1) * (cfg->number_of_col_bins + 1) * sizeof(int);
  if (smem > 48 * 1024)  // the shipped configurations (<= 21 x 61 bins) stay far below the default limit
    PSLAM_CUDA_TRY(ctx, cudaFuncSetAttribute(merger_select_updates_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
  merger_select_updates_kernel<<<1, MG_THREADS, smem, ctx->stream>>>(merger_params(cfg, dim), d_meas, n_meas, d_corr_moving,
                                                                      d_corr_response, n_corr, d_selected, d_occupied, n_words, d_result);
  PSLAM_LAUNCH_CHECK(ctx, "merger_select_updates_kernel");
  return PSLAM_OK;
}

int pslam_k_merger_select_additions(pslam_ctx* ctx, const pslam_merger_cfg* cfg, const float* d_meas, int dim, int n_meas,
                                    const unsigned* d_occupied, int* d_winners, int* d_result) {
  const size_t smem = (size_t) (cfg->number_of_row_bins + 1) * (cfg->number_of_col_bins + 1) * (sizeof(unsigned long long) + sizeof(int));
  if (smem > 48 * 1024)
    PSLAM_CUDA_TRY(ctx, cudaFuncSetAttribute(merger_select_additions_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
  merger_select_additions_kernel<<<1, MG_THREADS, smem, ctx->stream>>>(merger_params(cfg, dim), d_meas, n_meas, d_occupied, d_winners,
                                                                        d_result);
  PSLAM_LAUNCH_CHECK(ctx, "merger_select_additions_kernel");
  return PSLAM_OK;
}
