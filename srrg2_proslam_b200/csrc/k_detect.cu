// k_detect.cu -- stage 1 kernels (sm_100a): FAST-9/16 + NMS fused with the ORB 7x7 integer blur,
// ordered per-region compaction + std::sort-exact selection, feature assembly, ORB-256 description.
//
// Reference path replaced (see include/pslam_cuda.h):
//   IntensityFeatureExtractorBinned_::computeKeypoints  .../feature_extractors/intensity_feature_extractor_binned.cpp:115-208
//   IntensityFeatureExtractor_::computeDescriptors/compute  .../intensity_feature_extractor_base.cpp:45-85
#include "libstdcxx_sort.h"
#include "pslam_internal.cuh"
#include "pslam_kernels.cuh"

#include "../../include/pslam_orb_pattern.h"

namespace {

// ---------------------------------------------------------------------------------------------
// K1: one pass over the image tile in shared memory produces BOTH
//   nms_map : FAST-9/16 corner response after 3x3 non-max suppression (0 / response+1)
//   blur    : 7x7 sigma-2 Gaussian, OpenCV-3 fixed point ((sum k_i k_j p + 2^15) >> 16)
// so the image is read from HBM exactly once (algorithmic bytes: W*H read, 2*W*H written).
// Tile 64x32 output pixels, 4-pixel halo (3 for the ring / the taps + 1 for the NMS ring).
// ---------------------------------------------------------------------------------------------
constexpr int TW = 64, TH = 32, HALO = 4;
constexpr int SW = TW + 2 * HALO;  // 72
constexpr int SH = TH + 2 * HALO;  // 40
constexpr int SCW = TW + 2;        // score region 66 x 34
constexpr int SCH = TH + 2;
constexpr int SCP = 68;            // score row pitch
constexpr int K1_THREADS = 256;

__device__ __forceinline__ int fast_strength(const uint8_t* c /* centre in smem */) {
  // d_k = I(p) - I(p + off_k); s = max over 9-arcs of min(d) and of min(-d)
  const int v = c[0];
  int d[16];
  d[0] = v - c[3 * SW + 0];
  d[1] = v - c[3 * SW + 1];
  d[2] = v - c[2 * SW + 2];
  d[3] = v - c[1 * SW + 3];
  d[4] = v - c[0 * SW + 3];
  d[5] = v - c[-1 * SW + 3];
  d[6] = v - c[-2 * SW + 2];
  d[7] = v - c[-3 * SW + 1];
  d[8] = v - c[-3 * SW + 0];
  d[9] = v - c[-3 * SW - 1];
  d[10] = v - c[-2 * SW - 2];
  d[11] = v - c[-1 * SW - 3];
  d[12] = v - c[0 * SW - 3];
  d[13] = v - c[1 * SW - 3];
  d[14] = v - c[2 * SW - 2];
  d[15] = v - c[3 * SW - 1];
  int lo3[16], hi3[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    lo3[k] = min(min(d[k], d[(k + 1) & 15]), d[(k + 2) & 15]);
    hi3[k] = max(max(d[k], d[(k + 1) & 15]), d[(k + 2) & 15]);
  }
  // NOTE: keep the two polarities in separate accumulators and negate ONCE at the end.  Folding
  // `s = max(s, max(mn, -mx))` into the loop is miscompiled by nvcc 12.9 for sm_100a (the negation
  // is dropped when the expression is fused into VIMNMX3; caught by the GPU parity tests).
  int s_dark = -1024, s_bright = 1024;
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    const int mn = min(min(lo3[k], lo3[(k + 3) & 15]), lo3[(k + 6) & 15]);
    const int mx = max(max(hi3[k], hi3[(k + 3) & 15]), hi3[(k + 6) & 15]);
    s_dark = max(s_dark, mn);
    s_bright = min(s_bright, mx);
  }
  return max(s_dark, -s_bright);
}

__global__ void __launch_bounds__(K1_THREADS)
fast_blur_kernel(const uint8_t* __restrict__ images, long long image_pitch, int rows, int cols,
                 int stride, int thr, int nms, uint8_t* __restrict__ nms_map,
                 uint8_t* __restrict__ blur, int map_pitch, long long map_slot) {
  __shared__ __align__(16) uint8_t s_img[SH * SW];
  __shared__ __align__(16) uint8_t s_score[SCH * SCP];
  __shared__ __align__(16) uint16_t s_h[(TH + 6) * TW];
  __shared__ uint16_t s_cand[SCH * SCW];
  __shared__ int s_ncand;

  const int tid = threadIdx.x;
  const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
  const uint8_t* img = images + (size_t) blockIdx.z * image_pitch;
  uint8_t* out_nms = nms_map + (size_t) blockIdx.z * map_slot;
  uint8_t* out_blur = blur + (size_t) blockIdx.z * map_slot;

  if (tid == 0) s_ncand = 0;
  // ---- tile load, reflect-101 outside the image (FAST never uses those cells) ----
  for (int i = tid; i < SH * SW; i += K1_THREADS) {
    const int ty = i / SW, tx = i - ty * SW;
    const int gy = reflect101(y0 - HALO + ty, rows);
    const int gx = reflect101(x0 - HALO + tx, cols);
    s_img[i] = __ldg(img + (size_t) gy * stride + gx);
  }
  for (int i = tid; i < SCH * SCP / 4; i += K1_THREADS) reinterpret_cast<uint32_t*>(s_score)[i] = 0u;
  __syncthreads();

  // ---- blur, horizontal pass: rows [HALO-3, HALO+TH+3), cols of the tile ----
  for (int i = tid; i < (TH + 6) * TW; i += K1_THREADS) {
    const int hy = i / TW, hx = i - hy * TW;
    const uint8_t* p = s_img + (hy + HALO - 3) * SW + hx + HALO;
    const int acc = 18 * (p[-3] + p[3]) + 34 * (p[-2] + p[2]) + 49 * (p[-1] + p[1]) + 55 * p[0];
    s_h[i] = (uint16_t) acc;  // <= 257*255 = 65535
  }

  // ---- FAST phase 1: cheap compass reject over the 66x34 score region ----
  const int hi_thr = thr, lo_thr = -thr;
  for (int i = tid; i < SCH * SCW; i += K1_THREADS) {
    const int sy = i / SCW, sx = i - sy * SCW;
    const int gx = x0 - 1 + sx, gy = y0 - 1 + sy;
    if (gx < 3 || gx >= cols - 3 || gy < 3 || gy >= rows - 3) continue;
    const uint8_t* c = s_img + (sy + 3) * SW + (sx + 3);
    const int v = c[0];
    const int d0 = v - c[3 * SW], d4 = v - c[3], d8 = v - c[-3 * SW], d12 = v - c[-3];
    // ring darker than centre (d > thr) or brighter (d < -thr) on two adjacent compass points
    const bool k0 = d0 > hi_thr, k4 = d4 > hi_thr, k8 = d8 > hi_thr, k12 = d12 > hi_thr;
    const bool b0 = d0 < lo_thr, b4 = d4 < lo_thr, b8 = d8 < lo_thr, b12 = d12 < lo_thr;
    const bool cand = (k0 & k4) | (k4 & k8) | (k8 & k12) | (k12 & k0) | (b0 & b4) | (b4 & b8) |
                      (b8 & b12) | (b12 & b0);
    if (cand) {
      const int slot = atomicAdd(&s_ncand, 1);
      s_cand[slot] = (uint16_t) i;
    }
  }
  __syncthreads();

  // ---- FAST phase 2: full 16-ring strength, dense over the candidate list ----
  const int ncand = s_ncand;
  for (int k = tid; k < ncand; k += K1_THREADS) {
    const int i = s_cand[k];
    const int sy = i / SCW, sx = i - sy * SCW;
    const int s = fast_strength(s_img + (sy + 3) * SW + (sx + 3));
    if (s > thr) s_score[sy * SCP + sx] = (uint8_t) s;  // 1..255 (response = s - 1)
  }
  __syncthreads();

  // ---- outputs: 4 consecutive pixels per thread, one aligned 32-bit store each ----
  for (int i = tid; i < TH * TW / 4; i += K1_THREADS) {
    const int py = i / (TW / 4), px = (i - py * (TW / 4)) * 4;
    const int gy = y0 + py, gx = x0 + px;
    if (gy >= rows || gx >= map_pitch) continue;
    uint32_t w_nms = 0, w_blur = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      // blur, vertical pass
      const uint16_t* h = s_h + (py + 3) * TW + px + j;
      const int acc = 18 * (h[-3 * TW] + h[3 * TW]) + 34 * (h[-2 * TW] + h[2 * TW]) +
                      49 * (h[-TW] + h[TW]) + 55 * h[0];
      const int b = min(255, (acc + 32768) >> 16);
      // NMS: response strictly greater than all 8 neighbours (non-corners count as 0)
      const uint8_t* sc = s_score + (py + 1) * SCP + (px + j + 1);
      const int s = sc[0];
      int o = 0;
      if (s > 0) {
        if (nms) {
          const int r = s - 1;
          int m = max(max(sc[-SCP - 1], sc[-SCP]), max(sc[-SCP + 1], sc[-1]));
          m = max(m, max(max(sc[1], sc[SCP - 1]), max(sc[SCP], sc[SCP + 1])));
          const int nb = m > 0 ? m - 1 : 0;
          o = (r > nb) ? s : 0;  // stored value = response + 1
        } else {
          o = 1;  // response 0 without NMS
        }
      }
      if (gx + j < cols) {
        w_nms |= (uint32_t) o << (8 * j);
        w_blur |= (uint32_t) b << (8 * j);
      }
    }
    *reinterpret_cast<uint32_t*>(out_nms + (size_t) gy * map_pitch + gx) = w_nms;
    *reinterpret_cast<uint32_t*>(out_blur + (size_t) gy * map_pitch + gx) = w_blur;
  }
}

// ---------------------------------------------------------------------------------------------
// K2a: per (image, detection region): ordered (row-major) compaction of the NMS map into the
// region's raw list (the bucket loop of intensity_feature_extractor_binned.cpp:168-175; the
// row-major order is the order cv::FAST emits keypoints in).
// ---------------------------------------------------------------------------------------------
constexpr int K2_THREADS = 256;

__global__ void __launch_bounds__(K2_THREADS)
bin_compact_kernel(const uint8_t* __restrict__ nms_map, const uint8_t* __restrict__ mask,
                   int map_pitch, long long map_slot, int rows, int cols, int nh, int nv,
                   float pixel_rows_per_detector, float pixel_cols_per_detector,
                   uint32_t* __restrict__ raw, int max_raw_per_bin, int max_bins,
                   int* __restrict__ raw_count, int* __restrict__ flags) {
  extern __shared__ __align__(16) uint8_t s_colbin[];  // [map_pitch]
  __shared__ int s_warp[33];
  __shared__ int s_rbegin, s_rend;

  const int tid = threadIdx.x;
  const int bin = blockIdx.x, image = blockIdx.y;
  const int rb = bin / nh;
  const uint8_t* map = nms_map + (size_t) image * map_slot;
  uint32_t* seg = raw + ((size_t) image * max_bins + bin) * max_raw_per_bin;

  if (tid == 0) {
    s_rbegin = rows;
    s_rend = 0;
  }
  __syncthreads();
  // rows of this region row: floor(r / pixel_rows_per_detector) == rb   (binned.cpp:84-85)
  for (int r = tid; r < rows; r += K2_THREADS) {
    const float q = floorf(__fdiv_rn((float) r, pixel_rows_per_detector));
    if ((int) q == rb) {
      atomicMin(&s_rbegin, r);
      atomicMax(&s_rend, r + 1);
    }
  }
  // column -> region index for this region row: (size_t)((float)row_region + (float)c / pc)  (:86-88)
  const float row_region = __fmul_rn((float) rb, (float) nh);
  for (int c = tid; c < map_pitch; c += K2_THREADS) {
    const float f = __fadd_rn((float) (unsigned) row_region, __fdiv_rn((float) c, pixel_cols_per_detector));
    const unsigned idx = (unsigned) f;
    s_colbin[c] = (c < cols && idx == (unsigned) bin) ? 1 : 0;
  }
  __syncthreads();
  const int rbegin = s_rbegin, rend = s_rend;
  const int cpr = map_pitch / 16;  // 16-byte chunks per row
  const int total_chunks = (rend > rbegin) ? (rend - rbegin) * cpr : 0;
  int running = 0;
  for (int base = 0; base < total_chunks; base += K2_THREADS) {
    const int q = base + tid;
    uint32_t w[4] = {0, 0, 0, 0};
    int row = 0, c0 = 0;
    if (q < total_chunks) {
      row = rbegin + q / cpr;
      c0 = (q % cpr) * 16;
      const uint4 v = *reinterpret_cast<const uint4*>(map + (size_t) row * map_pitch + c0);
      w[0] = v.x;
      w[1] = v.y;
      w[2] = v.z;
      w[3] = v.w;
      if (mask && (w[0] | w[1] | w[2] | w[3])) {
        const uint4 m = *reinterpret_cast<const uint4*>(mask + (size_t) row * map_pitch + c0);
        const uint32_t mm[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (((mm[k] >> (8 * j)) & 0xffu) == 0) w[k] &= ~(0xffu << (8 * j));
      }
    }
    int cnt = 0;
    if (w[0] | w[1] | w[2] | w[3]) {
#pragma unroll
      for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int c = c0 + 4 * k + j;
          if (((w[k] >> (8 * j)) & 0xffu) && s_colbin[c]) ++cnt;
        }
    }
    int total;
    const int off = block_exclusive_scan<K2_THREADS>(cnt, s_warp, &total);
    if (cnt) {
      int o = running + off;
#pragma unroll
      for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int c = c0 + 4 * k + j;
          const uint32_t val = (w[k] >> (8 * j)) & 0xffu;
          if (val && s_colbin[c]) {
            if (o < max_raw_per_bin) seg[o] = ((uint32_t) (row * cols + c) << 8) | val;
            ++o;
          }
        }
    }
    running += total;
  }
  if (tid == 0) {
    if (running > max_raw_per_bin) {
      atomicOr(flags, PSLAM_FLAG_RAW_OVERFLOW);
      running = max_raw_per_bin;
    }
    raw_count[image * max_bins + bin] = running;
  }
}

// ---------------------------------------------------------------------------------------------
// K2b: selection.  A region with fewer than `quota` corners keeps all of them in detection
// order; otherwise the reference runs std::sort(response descending) -- unstable -- and keeps the
// first `quota` (binned.cpp:180-200).  One warp per region: the list is staged in shared memory,
// lane 0 replays libstdc++'s introsort (prefix-pruned, libstdcxx_sort.h), the warp writes back.
// ---------------------------------------------------------------------------------------------
struct RespGreater {
  __device__ __forceinline__ bool operator()(const uint32_t& a, const uint32_t& b) const {
    return (a & 0xffu) > (b & 0xffu);
  }
};

__global__ void __launch_bounds__(32)
bin_sort_kernel(uint32_t* __restrict__ raw, int max_raw_per_bin, int max_bins,
                const int* __restrict__ raw_count, int* __restrict__ sel_count,
                unsigned long long quota, int sort_cap) {
  extern __shared__ __align__(16) uint32_t s_sort[];
  const int lane = threadIdx.x;
  const int bin = blockIdx.x, image = blockIdx.y;
  uint32_t* seg = raw + ((size_t) image * max_bins + bin) * max_raw_per_bin;
  const int n = raw_count[image * max_bins + bin];
  int kept = n;
  if ((unsigned long long) n >= quota) {
    kept = (int) quota;
    if (kept > 0) {
      if (n <= sort_cap) {
        for (int i = lane; i < n; i += 32) s_sort[i] = seg[i];
        __syncwarp();
        if (lane == 0) pslam_sort::std_sort_prefix(s_sort, n, kept, RespGreater());
        __syncwarp();
        for (int i = lane; i < kept; i += 32) seg[i] = s_sort[i];
      } else if (lane == 0) {
        pslam_sort::std_sort_prefix(seg, n, kept, RespGreater());
      }
    }
  }
  if (lane == 0) sel_count[image * max_bins + bin] = kept;
}

// ---------------------------------------------------------------------------------------------
// K3a: per image: concatenate the regions' kept keypoints in region order, drop those closer than
// 31 px to the border (cv::ORB::compute -> KeyPointsFilter::runByImageBorder), fill the SoA store.
// ---------------------------------------------------------------------------------------------
constexpr int K3_THREADS = 256;

__global__ void __launch_bounds__(K3_THREADS)
assemble_features_kernel(const uint8_t* __restrict__ images, long long image_pitch, int stride,
                         int rows, int cols, int nbins, int max_bins, int max_raw_per_bin,
                         const uint32_t* __restrict__ raw, const int* __restrict__ sel_count,
                         int border, int max_features, int slot_base, float2* __restrict__ xy,
                         float* __restrict__ resp, float* __restrict__ inten,
                         int* __restrict__ count, int* __restrict__ flags) {
  __shared__ int s_warp[33];
  const int tid = threadIdx.x, image = blockIdx.x;
  const uint8_t* img = images + (size_t) image * image_pitch;
  const size_t slot = (size_t) slot_base + image;
  float2* o_xy = xy + slot * max_features;
  float* o_resp = resp + slot * max_features;
  float* o_int = inten + slot * max_features;
  int running = 0;
  for (int b = 0; b < nbins; ++b) {
    const uint32_t* seg = raw + ((size_t) image * max_bins + b) * max_raw_per_bin;
    const int n = sel_count[image * max_bins + b];
    for (int base = 0; base < n; base += K3_THREADS) {
      const int i = base + tid;
      uint32_t e = 0;
      int x = 0, y = 0, keep = 0;
      if (i < n) {
        e = seg[i];
        const int pix = (int) (e >> 8);
        y = pix / cols;
        x = pix - y * cols;
        keep = (x >= border && x < cols - border && y >= border && y < rows - border) ? 1 : 0;
      }
      int total;
      const int off = block_exclusive_scan<K3_THREADS>(keep, s_warp, &total);
      if (keep) {
        const int o = running + off;
        if (o < max_features) {
          o_xy[o] = make_float2((float) x, (float) y);
          o_resp[o] = (float) ((int) (e & 0xffu) - 1);
          o_int[o] = (float) __ldg(img + (size_t) y * stride + x);
        }
      }
      running += total;
    }
  }
  if (tid == 0) {
    if (running > max_features) {
      atomicOr(flags, PSLAM_FLAG_FEATURE_OVERFLOW);
      running = max_features;
    }
    count[slot] = running;
  }
}

// ---------------------------------------------------------------------------------------------
// K3b: ORB-256 (angle 0, bit_pattern_31_) on the blurred image: one warp per keypoint, the
// 31x31 patch staged in shared memory, lane l computes descriptor byte l (pairs 8l..8l+7).
// ---------------------------------------------------------------------------------------------
constexpr int K4_WARPS = 8;
constexpr int PATCH = 31, PATCH_PITCH = 33;

__constant__ signed char c_pattern[256 * 4];

__global__ void __launch_bounds__(K4_WARPS * 32)
orb_describe_kernel(const uint8_t* __restrict__ blur, int map_pitch, long long map_slot,
                    const float2* __restrict__ xy, const int* __restrict__ count, int max_features,
                    int slot_base, uint32_t* __restrict__ desc) {
  __shared__ uint8_t s_patch[K4_WARPS][PATCH * PATCH_PITCH + 3];
  __shared__ uint16_t s_off[16][32];  // [2*k + {a,b}][lane] byte offsets into the patch
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 512; i += K4_WARPS * 32) {
    const int l = i & 31, j = i >> 5;  // j = 2*k + ab
    const int pair = 8 * l + (j >> 1);
    const int xo = c_pattern[4 * pair + 2 * (j & 1)], yo = c_pattern[4 * pair + 2 * (j & 1) + 1];
    s_off[j][l] = (uint16_t) ((yo + 15) * PATCH_PITCH + (xo + 15));
  }
  __syncthreads();
  uint16_t off[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) off[j] = s_off[j][lane];

  const int image = blockIdx.y;
  const size_t slot = (size_t) slot_base + image;
  const int n = count[slot];
  const uint8_t* b = blur + (size_t) image * map_slot;
  uint8_t* patch = s_patch[wid];
  for (int i = blockIdx.x * K4_WARPS + wid; i < n; i += gridDim.x * K4_WARPS) {
    const float2 p = xy[slot * max_features + i];
    const int x = (int) p.x, y = (int) p.y;
    const uint8_t* src = b + (size_t) (y - 15) * map_pitch + (x - 15);
    __syncwarp();
    if (lane < PATCH) {
#pragma unroll 4
      for (int r = 0; r < PATCH; ++r) patch[r * PATCH_PITCH + lane] = __ldg(src + (size_t) r * map_pitch + lane);
    }
    __syncwarp();
    uint32_t byte = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int a = patch[off[2 * k]], c = patch[off[2 * k + 1]];
      byte |= (uint32_t) (a < c) << k;
    }
    // gather 4 bytes into one word: lanes 4w..4w+3 -> word w
    const uint32_t b1 = __shfl_down_sync(0xffffffffu, byte, 1);
    const uint32_t b2 = __shfl_down_sync(0xffffffffu, byte, 2);
    const uint32_t b3 = __shfl_down_sync(0xffffffffu, byte, 3);
    if ((lane & 3) == 0) {
      desc[(slot * max_features + i) * 8 + (lane >> 2)] = byte | (b1 << 8) | (b2 << 16) | (b3 << 24);
    }
  }
}

}  // namespace

// ---- host-side launchers -----------------------------------------------------------------------
int pslam_k_upload_pattern(pslam_ctx* ctx) {
  PSLAM_CUDA_TRY(ctx, cudaMemcpyToSymbol(c_pattern, PSLAM_ORB_PATTERN, sizeof(PSLAM_ORB_PATTERN)));
  return PSLAM_OK;
}

int pslam_k_fast_blur(pslam_ctx* ctx, const uint8_t* d_images, long long image_pitch, int n_images,
                      int rows, int cols, int stride, int thr, int nms) {
  dim3 grid((cols + TW - 1) / TW, (rows + TH - 1) / TH, n_images);
  thr = thr < 0 ? 0 : (thr > 255 ? 255 : thr);
  fast_blur_kernel<<<grid, K1_THREADS, 0, ctx->stream>>>(d_images, image_pitch, rows, cols, stride,
                                                         thr, nms, ctx->d_nms, ctx->d_blur,
                                                         ctx->map_pitch, (long long) ctx->map_slot);
  PSLAM_LAUNCH_CHECK(ctx, "fast_blur_kernel");
  return PSLAM_OK;
}

int pslam_k_bin_select(pslam_ctx* ctx, int n_images, int rows, int cols, int nh, int nv,
                       unsigned long long quota, const uint8_t* d_mask) {
  // float arithmetic of IntensityFeatureExtractorBinned_::init (binned.cpp:49-52)
  const float pr = static_cast<float>(rows) / static_cast<float>((size_t) nv);
  const float pc = static_cast<float>(cols) / static_cast<float>((size_t) nh);
  dim3 grid(nh * nv, n_images);
  bin_compact_kernel<<<grid, K2_THREADS, ctx->map_pitch, ctx->stream>>>(
    ctx->d_nms, d_mask, ctx->map_pitch, (long long) ctx->map_slot, rows, cols, nh, nv, pr, pc,
    ctx->d_raw, ctx->lim.max_raw_per_bin, ctx->lim.max_bins, ctx->d_raw_count, ctx->d_flags);
  PSLAM_LAUNCH_CHECK(ctx, "bin_compact_kernel");
  const int sort_cap = ctx->lim.max_raw_per_bin < 16384 ? ctx->lim.max_raw_per_bin : 16384;
  const size_t smem = (size_t) sort_cap * 4;
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(bin_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
  bin_sort_kernel<<<grid, 32, smem, ctx->stream>>>(ctx->d_raw, ctx->lim.max_raw_per_bin,
                                                   ctx->lim.max_bins, ctx->d_raw_count,
                                                   ctx->d_sel_count, quota, sort_cap);
  PSLAM_LAUNCH_CHECK(ctx, "bin_sort_kernel");
  return PSLAM_OK;
}

int pslam_k_assemble(pslam_ctx* ctx, const uint8_t* d_images, long long image_pitch, int stride,
                     int n_images, int rows, int cols, int nbins, int border, int slot_base) {
  assemble_features_kernel<<<n_images, K3_THREADS, 0, ctx->stream>>>(
    d_images, image_pitch, stride, rows, cols, nbins, ctx->lim.max_bins, ctx->lim.max_raw_per_bin,
    ctx->d_raw, ctx->d_sel_count, border, ctx->lim.max_features, slot_base, ctx->d_xy, ctx->d_resp,
    ctx->d_inten, ctx->d_count, ctx->d_flags);
  PSLAM_LAUNCH_CHECK(ctx, "assemble_features_kernel");
  return PSLAM_OK;
}

int pslam_k_describe(pslam_ctx* ctx, int n_images, int slot_base) {
  // enough warps for max_features per image, capped: grid-stride over the image's features
  int bx = (ctx->lim.max_features + K4_WARPS - 1) / K4_WARPS;
  if (bx > 64) bx = 64;
  dim3 grid(bx, n_images);
  orb_describe_kernel<<<grid, K4_WARPS * 32, 0, ctx->stream>>>(
    ctx->d_blur, ctx->map_pitch, (long long) ctx->map_slot, ctx->d_xy, ctx->d_count,
    ctx->lim.max_features, slot_base, ctx->d_desc);
  PSLAM_LAUNCH_CHECK(ctx, "orb_describe_kernel");
  return PSLAM_OK;
}
