// k_epipolar.cu -- stage 2a (sm_100a): stereo epipolar matching + stereo-point assembly.
//
// Replaces CorrespondenceFinderDescriptorBasedEpipolar::compute
//   (.../correspondence_finders/correspondence_finder_descriptor_based_epipolar_impl.cpp:44-219)
// and the assembly loop of RawDataPreprocessorStereoProjective::compute
//   (.../sensor_processing/raw_data_preprocessor_stereo_projective.cpp:105-132).
//
// One CTA per stereo pair.  Both feature sets are sorted by (row, col) with a shared-memory bitonic
// sort (keys are unique, so any sort gives the reference's std::sort order, :36-41).  The
// reference's single running `index_right` only couples left features of the SAME image row
// (it is reset to the first right feature of the next row by the two skip loops, :97-126), so each
// run of equal-row left features is scanned by one thread exactly as the reference does, all
// runs in parallel.  Results are emitted in the reference's scan order by a block-wide scan.
#include <float.h>

#include "pslam_internal.cuh"
#include "pslam_kernels.cuh"

namespace {

constexpr int EP_THREADS = 512;

struct EpSmem {
  unsigned long long* keys;  // [P]
  short *rowL, *colL, *idxL, *rowR, *colR, *idxR;  // [M] each
  short* match;              // [M] sorted right position matched to sorted left position, -1 none
  unsigned short* dist;      // [M]
  unsigned char* usedR;      // [M]
};

__device__ __forceinline__ void ep_sort_side(const float2* __restrict__ xy, int n, int P,
                                             unsigned long long* keys, short* row, short* col,
                                             short* idx) {
  const int tid = threadIdx.x;
  for (int i = tid; i < P; i += EP_THREADS) {
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
      for (int i = tid; i < P; i += EP_THREADS) {
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
  for (int i = tid; i < n; i += EP_THREADS) {
    const unsigned long long k = keys[i];
    row[i] = (short) (k >> 32);
    col[i] = (short) ((k >> 16) & 0xffffu);
    idx[i] = (short) (k & 0xffffu);
  }
  __syncthreads();
}

// ordered in-place removal of flagged entries from (row, col, idx); tmp >= 3*n shorts
__device__ __forceinline__ int ep_compact(short* row, short* col, short* idx, int n,
                                          const unsigned char* removed, short* tmp, int* s_warp) {
  const int tid = threadIdx.x;
  int running = 0;
  for (int base = 0; base < n; base += EP_THREADS) {
    const int i = base + tid;
    const int keep = (i < n && !removed[i]) ? 1 : 0;
    int total;
    const int off = block_exclusive_scan<EP_THREADS>(keep, s_warp, &total);
    if (keep) {
      const int o = running + off;
      tmp[3 * o] = row[i];
      tmp[3 * o + 1] = col[i];
      tmp[3 * o + 2] = idx[i];
    }
    running += total;
  }
  __syncthreads();
  for (int i = tid; i < running; i += EP_THREADS) {
    row[i] = tmp[3 * i];
    col[i] = tmp[3 * i + 1];
    idx[i] = tmp[3 * i + 2];
  }
  __syncthreads();
  return running;
}

__global__ void __launch_bounds__(EP_THREADS)
epipolar_kernel(const float2* __restrict__ xy, const uint32_t* __restrict__ desc,
                const int* __restrict__ count, int M, float max_dist, float max_ratio, int max_disp,
                int thickness, int* __restrict__ ep_fixed, int* __restrict__ ep_moving,
                float* __restrict__ ep_dist, int* __restrict__ ep_count,
                float4* __restrict__ st_uvuv, int* __restrict__ st_left, int* __restrict__ st_right,
                float* __restrict__ st_dist, int* __restrict__ st_count) {
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ int s_warp[33];
  const int tid = threadIdx.x;
  const int pair = blockIdx.x;
  const int imgL = 2 * pair, imgR = 2 * pair + 1;
  int nL = count[imgL], nR = count[imgR];
  int P = 1;
  while (P < max(nL, nR)) P <<= 1;

  unsigned long long* keys = reinterpret_cast<unsigned long long*>(smem);
  int Pmax = 1;
  while (Pmax < M) Pmax <<= 1;
  short* rowL = reinterpret_cast<short*>(smem + (size_t) Pmax * 8);
  short* colL = rowL + M;
  short* idxL = colL + M;
  short* rowR = idxL + M;
  short* colR = rowR + M;
  short* idxR = colR + M;
  short* match = idxR + M;
  unsigned short* dist = reinterpret_cast<unsigned short*>(match + M);
  unsigned char* usedR = reinterpret_cast<unsigned char*>(dist + M);
  unsigned char* usedL = usedR + M;

  const float2* xyL = xy + (size_t) imgL * M;
  const float2* xyR = xy + (size_t) imgR * M;
  const uint4* descL = reinterpret_cast<const uint4*>(desc + (size_t) imgL * M * 8);
  const uint4* descR = reinterpret_cast<const uint4*>(desc + (size_t) imgR * M * 8);

  ep_sort_side(xyL, nL, P, keys, rowL, colL, idxL);
  ep_sort_side(xyR, nR, P, keys, rowR, colR, idxR);

  int* o_fixed = ep_fixed + (size_t) pair * M;
  int* o_moving = ep_moving + (size_t) pair * M;
  float* o_dist = ep_dist + (size_t) pair * M;
  int n_out = 0;

  const int n_offsets = 2 * thickness + 1;
  for (int oi = 0; oi < n_offsets; ++oi) {
    // row offsets 0, +1, -1, +2, -2, ...  (epipolar_impl.cpp:72-79)
    const int off = (oi == 0) ? 0 : ((oi & 1) ? (oi + 1) / 2 : -(oi / 2));
    for (int i = tid; i < nL; i += EP_THREADS) {
      match[i] = -1;
      usedL[i] = 0;
    }
    for (int i = tid; i < nR; i += EP_THREADS) usedR[i] = 0;
    __syncthreads();
    if (nR > 0) {
      for (int i = tid; i < nL; i += EP_THREADS) {
        if (i != 0 && rowL[i] == rowL[i - 1]) continue;  // not the start of a row run
        const int row_left = rowL[i] + off;
        int lo = 0, hi = nR;  // first right feature with row >= row_left
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if (rowR[mid] < row_left) lo = mid + 1; else hi = mid;
        }
        int index_right = lo;
        for (int j = i; j < nL && rowL[j] == rowL[i]; ++j) {
          if (index_right >= nR || rowR[index_right] != row_left) break;
          const int col_left = colL[j];
          const uint4 a0 = __ldg(descL + 2 * (int) idxL[j]), a1 = __ldg(descL + 2 * (int) idxL[j] + 1);
          int best = INT_MAX, second = INT_MAX, best_pos = 0;
          for (int s = index_right; s < nR && rowR[s] == row_left; ++s) {
            const int disparity = col_left - colR[s];
            if (disparity < 0) break;
            if (disparity > max_disp) continue;
            const uint4 b0 = __ldg(descR + 2 * (int) idxR[s]), b1 = __ldg(descR + 2 * (int) idxR[s] + 1);
            const int d = hamming256(a0, a1, b0, b1);
            if (d < best) {
              second = best;
              best = d;
              best_pos = s;
            } else if (d < second) {
              second = d;
            }
          }
          if (best == INT_MAX) continue;
          const float fb = (float) best;
          const float fs = (second == INT_MAX) ? FLT_MAX : (float) second;
          if (fb < max_dist && __fdiv_rn(fb, fs) < max_ratio) {  // :171-173
            match[j] = (short) best_pos;
            dist[j] = (unsigned short) best;
            usedL[j] = 1;
            usedR[best_pos] = 1;
            index_right = best_pos + 1;  // ordering constraint :181
          }
        }
      }
    }
    __syncthreads();
    // emit this pass's matches in scan order
    for (int base = 0; base < nL; base += EP_THREADS) {
      const int i = base + tid;
      const int has = (i < nL && match[i] >= 0) ? 1 : 0;
      int total;
      const int o = block_exclusive_scan<EP_THREADS>(has, s_warp, &total);
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
    if (oi + 1 < n_offsets) {  // prune matched candidates, keeping the order (:189-205)
      __syncthreads();
      nL = ep_compact(rowL, colL, idxL, nL, usedL, reinterpret_cast<short*>(keys), s_warp);
      nR = ep_compact(rowR, colR, idxR, nR, usedR, reinterpret_cast<short*>(keys), s_warp);
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
  for (int base = 0; base < n_out; base += EP_THREADS) {
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
    const int o = block_exclusive_scan<EP_THREADS>(keep, s_warp, &total);
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

static size_t ep_smem_bytes(int M) {
  int P = 1;
  while (P < M) P <<= 1;
  return (size_t) P * 8 + (size_t) M * (7 * 2 + 2 + 2);
}

int pslam_k_epipolar(pslam_ctx* ctx, int n_pairs, const pslam_match_cfg* cfg) {
  const int M = ctx->lim.max_features;
  const size_t smem = ep_smem_bytes(M);
  if (smem > 48 * 1024) {
    PSLAM_CUDA_TRY(ctx, cudaFuncSetAttribute(epipolar_kernel,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
  }
  epipolar_kernel<<<n_pairs, EP_THREADS, smem, ctx->stream>>>(
    ctx->d_xy, ctx->d_desc, ctx->d_count, M, cfg->maximum_descriptor_distance,
    cfg->maximum_distance_ratio_to_second_best, cfg->maximum_disparity_pixels,
    cfg->epipolar_line_thickness_pixels, ctx->d_ep_fixed, ctx->d_ep_moving, ctx->d_ep_dist,
    ctx->d_ep_count, ctx->d_st_uvuv, ctx->d_st_left, ctx->d_st_right, ctx->d_st_dist,
    ctx->d_st_count);
  PSLAM_LAUNCH_CHECK(ctx, "epipolar_kernel");
  return PSLAM_OK;
}
