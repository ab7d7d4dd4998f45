// k_detect.cu -- stage 1 back half (sm_100a): feature assembly and ORB-256 description
// (detection, blur and selection live in k_fast.cu).
//
// Reference path replaced (see include/pslam_cuda.h):
//   IntensityFeatureExtractorBinned_::computeKeypoints  .../feature_extractors/intensity_feature_extractor_binned.cpp:115-208
//   IntensityFeatureExtractor_::computeDescriptors/compute  .../intensity_feature_extractor_base.cpp:45-85
#include "pslam_internal.cuh"
#include "pslam_kernels.cuh"

#include "../../include/pslam_orb_pattern.h"
#include "orb_schedule.h"

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
// K3b: ORB-256 (angle 0, bit_pattern_31_) on the blurred image: one warp per keypoint, lane l computes
// descriptor byte l (pairs 8l..8l+7).
//   * staging: ONE TMA box load (cp.async.bulk.tensor.3d) per keypoint brings the patch from the blur map (L2 / HBM)
//     straight into shared memory -- no LSU wavefronts for the patch bytes.  A TMA box has to start on a 16-byte
//     boundary (an unaligned start coordinate faults: tools/debug/dbg_tma.cu), so the box is the 48 x 31 bytes at
//     ((x-15) & ~15, y-15, image) and every sample offset gets shift = (x-15) & 15 added.  Two patch buffers per
//     warp: the box of the next keypoint is in flight while the current one is sampled; one mbarrier per buffer.
//     (The LDG-based staging this replaces ran at 90 % l1tex utilisation: profiles/r01d_other_kernels_ncu.txt.)
//   * sampling: 16 byte loads per lane from the dense 48-byte rows in the bank-conflict-minimising order of
//     orb_schedule.h (which pair a lane visits when, and which end it reads first, is free: the bits are
//     reassembled in registers).
// ---------------------------------------------------------------------------------------------
constexpr int K4_WARPS = 8;
// bit_pattern_31_ only reaches +-13 pixels (include/pslam_orb_pattern.h), so the box holds the 27 rows y - 13 .. y + 13:
// rows y +- 14, 15 of the nominal 31 x 31 patch are never sampled and stay in L2 (the kernel is L2 bound: -13 % traffic).
// The schedule's byte offsets are relative to row y - 15; dropping two rows shifts every sample by the same 96 bytes
// (24 words), which rotates all banks alike and keeps the schedule conflict-free.
constexpr int PATCH = 27, PATCH_ROW0 = 13, PATCH_SKIP = (15 - PATCH_ROW0) * PSLAM_ORB_PATCH_PITCH;
constexpr int PATCH_BOX_BYTES = PATCH * PSLAM_ORB_PATCH_PITCH, PATCH_BUF = 1536;
static_assert(PSLAM_ORB_PATCH_PITCH == 48, "the TMA box rows are 48 bytes: 15 bytes of alignment slack + 31 + 2");

__constant__ signed char c_pattern[256 * 4];
// per-lane schedule tables: read with lane-varying indices, so they live in global memory (coalesced, L1 / L2
// resident) -- the constant cache would serialise the 32 different addresses of a warp
__device__ uint16_t g_sched_off[16][32];
__device__ uint8_t g_sched_bit[8][32];

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }

__device__ __forceinline__ void tma_patch_load(const CUtensorMap* tmap, uint32_t dst, uint32_t bar, int x0, int y0, int img) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(PATCH_BOX_BYTES) : "memory");
  asm volatile(
    "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
    ::"r"(dst), "l"(tmap), "r"(bar), "r"(x0), "r"(y0), "r"(img)
    : "memory");
}

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
    "{\n"
    ".reg .pred p;\n"
    "WAIT_%=:\n"
    "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
    "@p bra DONE_%=;\n"
    "bra WAIT_%=;\n"
    "DONE_%=:\n"
    "}" ::"r"(bar), "r"(parity)
    : "memory");
}

__global__ void __launch_bounds__(K4_WARPS * 32)
orb_describe_kernel(const __grid_constant__ CUtensorMap tmap, const float2* __restrict__ xy,
                    const int* __restrict__ count, int max_features, int slot_base, uint32_t* __restrict__ desc) {
  __shared__ __align__(128) uint8_t s_patch[K4_WARPS][2][PATCH_BUF];
  __shared__ __align__(8) unsigned long long s_bar[K4_WARPS][2];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_bar[wid][0])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_bar[wid][1])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncwarp();
  // this lane's schedule: byte offsets of the 16 samples, (bit position | flip << 3) of the 8 pairs
  unsigned off[8], bits = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    off[j] = (unsigned) __ldg(&g_sched_off[2 * j][lane]) | ((unsigned) __ldg(&g_sched_off[2 * j + 1][lane]) << 16);
    bits |= (unsigned) __ldg(&g_sched_bit[j][lane]) << (4 * j);
  }
  const int image = blockIdx.y;
  const size_t slot = (size_t) slot_base + image;
  const int n = count[slot];
  const float2* pxy = xy + slot * max_features;
  const int stride = gridDim.x * K4_WARPS;
  const uint32_t bar0 = smem_u32(&s_bar[wid][0]), buf0 = smem_u32(&s_patch[wid][0][0]);
  int i = blockIdx.x * K4_WARPS + wid;
  if (i < n && lane == 0) {
    const float2 p = pxy[i];
    tma_patch_load(&tmap, buf0, bar0, ((int) p.x - 15) & ~15, (int) p.y - PATCH_ROW0, image);
  }
  for (int k = 0; i < n; i += stride, ++k) {
    const int b = k & 1;
    // all lanes are done with buffer b ^ 1 (sampled in the previous iteration): refill it with the next patch
    __syncwarp();
    if (lane == 0 && i + stride < n) {
      const float2 p = pxy[i + stride];
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      tma_patch_load(&tmap, buf0 + (b ^ 1) * PATCH_BUF, bar0 + (b ^ 1) * 8, ((int) p.x - 15) & ~15, (int) p.y - PATCH_ROW0, image);
    }
    const int shift = ((int) pxy[i].x - 15) & 15;  // column of the patch inside the aligned box
    mbar_wait(bar0 + b * 8, (unsigned) (k >> 1) & 1u);
    const uint8_t* patch8 = s_patch[wid][b] + shift - PATCH_SKIP;  // offsets of the schedule start at row y - 15
    uint32_t byte = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int first = patch8[off[j] & 0xffffu], second = patch8[off[j] >> 16];
      const unsigned bj = bits >> (4 * j);
      const int m = -(int) ((bj >> 3) & 1u);          // flip: (second, first) were read as (first, second)
      const int d = ((first - second) ^ m) - m;        // a - c  with a = end A, c = end B of the pair
      byte |= ((unsigned) d >> 31) << (bj & 7u);       // bit = (a < c)
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
  PSLAM_CUDA_TRY(ctx, cudaMemcpyToSymbol(g_sched_off, PSLAM_ORB_SCHED_OFF, sizeof(PSLAM_ORB_SCHED_OFF)));
  PSLAM_CUDA_TRY(ctx, cudaMemcpyToSymbol(g_sched_bit, PSLAM_ORB_SCHED_BIT, sizeof(PSLAM_ORB_SCHED_BIT)));
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
  // few CTAs per image: every warp amortises its schedule load over ~max_features / (16 * 8) keypoints
  int bx = (ctx->lim.max_features + K4_WARPS - 1) / K4_WARPS;
  if (bx > 16) bx = 16;
  dim3 grid(bx, n_images);
  orb_describe_kernel<<<grid, K4_WARPS * 32, 0, ctx->stream>>>(ctx->blur_tmap, ctx->d_xy, ctx->d_count,
                                                              ctx->lim.max_features, slot_base, ctx->d_desc);
  PSLAM_LAUNCH_CHECK(ctx, "orb_describe_kernel");
  return PSLAM_OK;
}

// TMA descriptor of the blur maps (created once per context).  cuTensorMapEncodeTiled is fetched through the
// runtime (cudaGetDriverEntryPoint), so the library does not link against libcuda.
int pslam_k_make_blur_tmap(pslam_ctx* ctx, int work_images) {
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  PSLAM_CUDA_TRY(ctx, cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  if (!fn || qres != cudaDriverEntryPointSuccess)
    return pslam_set_error(ctx, PSLAM_E_CUDA, "cuTensorMapEncodeTiled is not available in this driver", cudaSuccess);
  const cuuint64_t dims[3] = {(cuuint64_t) ctx->map_pitch, (cuuint64_t) ctx->lim.max_rows, (cuuint64_t) work_images};
  const cuuint64_t strides[2] = {(cuuint64_t) ctx->map_pitch, (cuuint64_t) ctx->map_slot};  // bytes, dims 1 and 2
  const cuuint32_t box[3] = {PSLAM_ORB_PATCH_PITCH, PATCH, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUresult r = ((EncodeFn) fn)(&ctx->blur_tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, ctx->d_blur, dims, strides, box, estr,
                                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char msg[96];
    snprintf(msg, sizeof msg, "cuTensorMapEncodeTiled failed (CUresult %d)", (int) r);
    return pslam_set_error(ctx, PSLAM_E_CUDA, msg, cudaSuccess);
  }
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
