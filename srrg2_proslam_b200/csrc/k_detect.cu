// k_detect.cu -- stage 1 back half (sm_100a): feature assembly and ORB-256 description
// (detection, blur and selection live in k_fast.cu).
//
// Reference path replaced (see include/pslam_cuda.h):
//   IntensityFeatureExtractorBinned_::computeKeypoints  .../feature_extractors/intensity_feature_extractor_binned.cpp:115-208
//   IntensityFeatureExtractor_::computeDescriptors/compute  .../intensity_feature_extractor_base.cpp:45-85
#include "pslam_internal.cuh"
#include "pslam_kernels.cuh"

#include "../../include/pslam_orb_pattern.h"

namespace {

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


// ---- monocular + depth adaptor: order-preserving compaction of the features with a valid depth ---------------
// Replaces RawDataPreprocessorMonocularDepth::_readDepth (.../sensor_processing/raw_data_preprocessor_monocular_depth.cpp:157-180):
// d = depth(rint(v), rint(u)) converted to float; keep the point iff d > 0; z = scale * d (one fp32 multiply).
// One CTA per image; chunks of MD_THREADS features, block scan keeps the cloud order.
constexpr int MD_THREADS = 256;
__global__ void __launch_bounds__(MD_THREADS)
mono_depth_kernel(const void* __restrict__ depth, int depth_type, int depth_rows, int depth_cols, int depth_stride,
                  float scale, const float2* __restrict__ xy, const float* __restrict__ inten,
                  const uint32_t* __restrict__ desc, const int* __restrict__ count, int max_features, int slot,
                  float* __restrict__ out_uvz, float* __restrict__ out_inten, uint32_t* __restrict__ out_desc,
                  int* __restrict__ out_count) {
  __shared__ int s_warp[33];
  const size_t base = (size_t) slot * max_features;
  const int n = count[slot];
  int running = 0;
  for (int b = 0; b < n; b += MD_THREADS) {
    const int i = b + threadIdx.x;
    float d = 0.f;
    float2 p = make_float2(0.f, 0.f);
    if (i < n) {
      p = xy[base + i];
      const int r = (int) rintf(p.y), c = (int) rintf(p.x);
      if (r >= 0 && r < depth_rows && c >= 0 && c < depth_cols) {
        const size_t o = (size_t) r * depth_stride + c;
        d = depth_type == 0 ? (float) reinterpret_cast<const unsigned short*>(depth)[o] : reinterpret_cast<const float*>(depth)[o];
      }
    }
    const int keep = (i < n && d > 0.f) ? 1 : 0;
    int total;
    const int off = block_exclusive_scan<MD_THREADS>(keep, s_warp, &total);
    if (keep) {
      const size_t o = (size_t) running + off;
      out_uvz[3 * o] = p.x;
      out_uvz[3 * o + 1] = p.y;
      out_uvz[3 * o + 2] = __fmul_rn(scale, d);
      out_inten[o] = inten[base + i];
      const uint4* src = reinterpret_cast<const uint4*>(desc + (base + i) * 8);
      uint4* dst = reinterpret_cast<uint4*>(out_desc + o * 8);
      dst[0] = src[0];
      dst[1] = src[1];
    }
    running += total;
  }
  if (threadIdx.x == 0) *out_count = running;
}

}  // namespace

// ---- host-side launchers -----------------------------------------------------------------------
int pslam_k_upload_pattern(pslam_ctx* ctx) {
  PSLAM_CUDA_TRY(ctx, cudaMemcpyToSymbol(c_pattern, PSLAM_ORB_PATTERN, sizeof(PSLAM_ORB_PATTERN)));
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

int pslam_k_mono_depth(pslam_ctx* ctx, const void* d_depth, int depth_type, int depth_rows, int depth_cols,
                       int depth_stride, float scale, int slot, float* d_uvz, float* d_inten, uint32_t* d_desc,
                       int* d_count) {
  mono_depth_kernel<<<1, MD_THREADS, 0, ctx->stream>>>(d_depth, depth_type, depth_rows, depth_cols, depth_stride, scale,
                                                      ctx->d_xy, ctx->d_inten, ctx->d_desc, ctx->d_count,
                                                      ctx->lim.max_features, slot, d_uvz, d_inten, d_desc, d_count);
  PSLAM_LAUNCH_CHECK(ctx, "mono_depth_kernel");
  return PSLAM_OK;
}
