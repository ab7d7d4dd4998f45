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
// description tiles (orb_describe_tiles_kernel): ORB_TW x ORB_TH pixels of the blur map + the pattern's reach around them
constexpr int ORB_TW = 128, ORB_TH = 64;

__global__ void __launch_bounds__(K3_THREADS)
assemble_features_kernel(const uint8_t* __restrict__ images, long long image_pitch, int stride,
                         int rows, int cols, int nbins, int max_bins, int max_raw_per_bin,
                         const uint32_t* __restrict__ raw, const int* __restrict__ sel_count,
                         int border, int max_features, int slot_base, float2* __restrict__ xy,
                         float* __restrict__ resp, float* __restrict__ inten,
                         int* __restrict__ count, int* __restrict__ flags, int tiles_x, int n_tiles, int tile_cap,
                         int* __restrict__ tile_start, uint32_t* __restrict__ tile_order) {
  extern __shared__ int s_tile[];  // [n_tiles] histogram / cursors of the description tiles
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
  if (running > max_features) {
    if (tid == 0) atomicOr(flags, PSLAM_FLAG_FEATURE_OVERFLOW);
    running = max_features;
  }
  if (tid == 0) count[slot] = running;
  // ---- bucket the kept keypoints by description tile (ORB_TW x ORB_TH pixels) for orb_describe_tiles_kernel: the
  // descriptors are written to the keypoints' own slots, so the order inside a tile does not matter
  int* t_start = tile_start + slot * (size_t) (tile_cap + 1);
  uint32_t* t_order = tile_order + slot * (size_t) max_features;
  for (int t = tid; t < n_tiles; t += K3_THREADS) s_tile[t] = 0;
  __syncthreads();  // also: the xy written above are visible to the whole CTA
  for (int i = tid; i < running; i += K3_THREADS) {
    const float2 p = o_xy[i];
    atomicAdd(&s_tile[((int) p.y / ORB_TH) * tiles_x + (int) p.x / ORB_TW], 1);
  }
  __syncthreads();
  int carry = 0;
  for (int base = 0; base < n_tiles; base += K3_THREADS) {
    const int t = base + tid;
    const int c = t < n_tiles ? s_tile[t] : 0;
    int total;
    const int off = block_exclusive_scan<K3_THREADS>(c, s_warp, &total);
    if (t < n_tiles) {
      t_start[t] = carry + off;
      s_tile[t] = carry + off;  // cursor
    }
    carry += total;
    __syncthreads();
  }
  if (tid == 0) t_start[n_tiles] = carry;
  __syncthreads();
  for (int i = tid; i < running; i += K3_THREADS) {
    const float2 p = o_xy[i];
    const int x = (int) p.x, y = (int) p.y;
    const int pos = atomicAdd(&s_tile[(y / ORB_TH) * tiles_x + x / ORB_TW], 1);
    // keypoint index | column inside the tile << 13 | row inside the tile << 20: ONE load per keypoint in the consumer
    t_order[pos] = (uint32_t) i | ((uint32_t) (x % ORB_TW) << 13) | ((uint32_t) (y % ORB_TH) << 20);
  }
}

// ---------------------------------------------------------------------------------------------
// K3b: ORB-256 (angle 0, bit_pattern_31_) on the blurred image, TILE-MAJOR: one CTA per (tile, image).
//   * staging: ONE TMA box load (cp.async.bulk.tensor.3d) brings the tile's ORB_TW x ORB_TH pixels of the blur map plus
//     the pattern's reach (+-13 px; bit_pattern_31_ never samples rows / columns +-14, 15 of the nominal 31 x 31 patch)
//     into shared memory: a 176 x 90 byte box at (tx * 128 - 16, ty * 64 - 13) -- TMA boxes start on 16-byte boundaries,
//     coordinates outside the map are zero-filled (never sampled: keypoints keep 31 px from the border).  Every
//     keypoint of the tile is described from that copy: 60 boxes = 0.95 MB of L2 reads per KITTI image instead of one
//     48 x 27 box per keypoint (4.2 MB at 3 244 keypoints -- the keypoint-major kernel ran at 71 % of the L2 peak).
//   * sampling: one warp per keypoint, lane l computes descriptor byte l (pairs 8l .. 8l+7) with 16 byte loads in the
//     bank-conflict-minimising order of orb_schedule.h.  That schedule was generated for a 48-byte patch pitch; the tile
//     pitch of 176 bytes = 44 words is congruent to 12 words mod 32 banks, so the same schedule stays conflict-minimal.
//   * per pair 2 LDS + 2 IMAD + SHF + LOP3: the orientation of a pair ("which end is read first" is part of the
//     schedule) is folded into per-lane multipliers s = +-1: t = first * s - second * s, bit = sign(t); the bit lands
//     in its place with one (t >> 31) & mask OR.  All per-lane constants (16 offsets, 8 multipliers, 8 masks) live in
//     registers across the keypoints of the tile.
// ---------------------------------------------------------------------------------------------
constexpr int K4_WARPS = 2;  // measured us / image: 1 warp per tile CTA 0.461, 2 0.425, 3 0.458, 4 0.461, 8 0.604 (15.9 kB tile per CTA: 14 CTAs per SM)
constexpr int ORB_REACH = 13;                           // bit_pattern_31_ reaches +-13 pixels (include/pslam_orb_pattern.h)
constexpr int ORB_BOX_W = 176, ORB_BOX_H = ORB_TH + 2 * ORB_REACH, ORB_BOX_X0 = 16;  // box = [tx*128 - 16, +176) x [ty*64 - 13, +90)
constexpr int ORB_BOX_BYTES = ORB_BOX_W * ORB_BOX_H;
static_assert(PSLAM_ORB_PATCH_PITCH == 48, "orb_schedule.h is generated for a 48-byte pitch");
static_assert((ORB_BOX_W / 4) % 32 == (PSLAM_ORB_PATCH_PITCH / 4) % 32, "tile pitch must map rows to banks like the schedule's pitch");
static_assert(ORB_TW <= 128 && ORB_TH <= 64 && PSLAM_MAX_FEATURES_HARD <= 8192, "tile entry packing: 13 + 7 + 6 bits");
static_assert(ORB_BOX_X0 >= ORB_REACH && ORB_BOX_W - ORB_BOX_X0 - ORB_TW >= ORB_REACH + 1 && ORB_BOX_W % 16 == 0, "tile box geometry");

__constant__ signed char c_pattern[256 * 4];
// per-lane schedule tables: read with lane-varying indices, so they live in global memory (coalesced, L1 / L2
// resident) -- the constant cache would serialise the 32 different addresses of a warp
// [lane][40]: off_a[8], off_b[8] (byte offsets relative to the keypoint, tile pitch), mul[8], -mul[8], mask[8] -- expanded on
// the host from orb_schedule.h once per context, read as ten 16-byte loads per lane
__device__ int4 g_tile_sched[32][10];

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }

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
orb_describe_tiles_kernel(const __grid_constant__ CUtensorMap tmap, const float2* __restrict__ xy, int max_features,
                          int slot_base, int tiles_x, int tile_cap, const int* __restrict__ tile_start,
                          const uint32_t* __restrict__ tile_order, uint32_t* __restrict__ desc) {
  __shared__ __align__(128) uint8_t s_tile[ORB_BOX_BYTES];
  __shared__ __align__(8) unsigned long long s_bar;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int tile = blockIdx.x, image = blockIdx.y;
  const size_t slot = (size_t) slot_base + image;
  const int* t_start = tile_start + slot * (size_t) (tile_cap + 1);
  const int k_begin = t_start[tile], k_end = t_start[tile + 1];
  if (k_begin == k_end) return;  // nothing to describe here: the tile is never loaded
  const int tx = tile % tiles_x, ty = tile / tiles_x;
  const int bx0 = tx * ORB_TW - ORB_BOX_X0, by0 = ty * ORB_TH - ORB_REACH;
  const uint32_t bar = smem_u32(&s_bar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(ORB_BOX_BYTES) : "memory");
    asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(s_tile)), "l"(&tmap), "r"(bar), "r"(bx0), "r"(by0), "r"(image)
      : "memory");
  }
  if (k_begin + wid >= k_end) return;  // fewer keypoints than warps (an exited warp does not hold up the barrier below)
  // this lane's schedule while the box is in flight (expanded on the host: pslam_k_upload_pattern)
  int off_a[8], off_b[8], mul[8], nmul[8];
  unsigned mask[8];
  {
    int4 v[10];
#pragma unroll
    for (int q = 0; q < 10; ++q) v[q] = __ldg(&g_tile_sched[lane][q]);
    const int* f = reinterpret_cast<const int*>(v);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      off_a[j] = f[j];
      off_b[j] = f[8 + j];
      mul[j] = f[16 + j];
      nmul[j] = f[24 + j];
      mask[j] = (unsigned) f[32 + j];
    }
  }
  const uint32_t* order = tile_order + slot * (size_t) max_features;
  uint32_t e_next = __ldg(order + k_begin + wid);
  __syncthreads();  // the barrier is initialised before anybody waits on it
  mbar_wait(bar, 0u);
#pragma unroll 1
  for (int k = k_begin + wid; k < k_end; k += K4_WARPS) {
    const uint32_t e = e_next;
    if (k + K4_WARPS < k_end) e_next = __ldg(order + k + K4_WARPS);  // the next keypoint's entry is in flight while this one is sampled
    const int i = (int) (e & 0x1fffu);
    const uint8_t* c = s_tile + ((int) (e >> 20) + ORB_REACH) * ORB_BOX_W + ((int) ((e >> 13) & 0x7fu) + ORB_BOX_X0);
    uint32_t byte = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int first = c[off_a[j]], second = c[off_b[j]];
      // t = first * s - second * s = a - c (a = end A, c = end B of the pair), two IMADs on the otherwise idle FMA pipe
      // (the compiler's own form re-derives a predicate per pair and negates conditionally); then
      // byte |= (t >> 31) & mask: bit = (a < c), as SHF + LOP3 (instead of a compare, a select and an OR)
      asm("{\n.reg .s32 t, x;\nmul.lo.s32 t, %2, %4;\nmad.lo.s32 t, %1, %3, t;\nshr.s32 x, t, 31;\nlop3.b32 %0, %0, x, %5, 0xF8;\n}"
          : "+r"(byte) : "r"(first), "r"(second), "r"(mul[j]), "r"(nmul[j]), "r"(mask[j]));
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
  // sample offsets re-based from the schedule's 48-byte patch pitch (origin: row y - 15, column x - 15) to the tile pitch,
  // relative to the keypoint's own byte; multiplier +-1 (flip: the ends were scheduled as (second, first)) and bit mask per pair
  static int table[32][40];
  for (int lane = 0; lane < 32; ++lane)
    for (int j = 0; j < 8; ++j) {
      const int oa = PSLAM_ORB_SCHED_OFF[2 * j][lane], ob = PSLAM_ORB_SCHED_OFF[2 * j + 1][lane];
      const int bj = PSLAM_ORB_SCHED_BIT[j][lane];
      table[lane][j] = (oa / PSLAM_ORB_PATCH_PITCH - 15) * ORB_BOX_W + (oa % PSLAM_ORB_PATCH_PITCH - 15);
      table[lane][8 + j] = (ob / PSLAM_ORB_PATCH_PITCH - 15) * ORB_BOX_W + (ob % PSLAM_ORB_PATCH_PITCH - 15);
      table[lane][16 + j] = (bj & 8) ? -1 : 1;
      table[lane][24 + j] = (bj & 8) ? 1 : -1;
      table[lane][32 + j] = 1 << (bj & 7);
    }
  PSLAM_CUDA_TRY(ctx, cudaMemcpyToSymbol(g_tile_sched, table, sizeof(table)));
  return PSLAM_OK;
}

int pslam_k_assemble(pslam_ctx* ctx, const uint8_t* d_images, long long image_pitch, int stride,
                     int n_images, int rows, int cols, int nbins, int border, int slot_base) {
  const int tiles_x = (cols + ORB_TW - 1) / ORB_TW, n_tiles = tiles_x * ((rows + ORB_TH - 1) / ORB_TH);
  assemble_features_kernel<<<n_images, K3_THREADS, sizeof(int) * (size_t) n_tiles, ctx->stream>>>(
    d_images, image_pitch, stride, rows, cols, nbins, ctx->lim.max_bins, ctx->lim.max_raw_per_bin,
    ctx->d_raw, ctx->d_sel_count, border, ctx->lim.max_features, slot_base, ctx->d_xy, ctx->d_resp,
    ctx->d_inten, ctx->d_count, ctx->d_flags, tiles_x, n_tiles, ctx->tile_cap, ctx->d_tile_start, ctx->d_tile_order);
  PSLAM_LAUNCH_CHECK(ctx, "assemble_features_kernel");
  return PSLAM_OK;
}

int pslam_k_orb_tile_cap(int max_rows, int max_cols) {
  return ((max_cols + ORB_TW - 1) / ORB_TW) * ((max_rows + ORB_TH - 1) / ORB_TH);
}

int pslam_k_describe(pslam_ctx* ctx, int n_images, int rows, int cols, int slot_base) {
  const int tiles_x = (cols + ORB_TW - 1) / ORB_TW, n_tiles = tiles_x * ((rows + ORB_TH - 1) / ORB_TH);
  dim3 grid(n_tiles, n_images);
  orb_describe_tiles_kernel<<<grid, K4_WARPS * 32, 0, ctx->stream>>>(ctx->blur_tmap, ctx->d_xy, ctx->lim.max_features, slot_base,
                                                                    tiles_x, ctx->tile_cap, ctx->d_tile_start, ctx->d_tile_order,
                                                                    ctx->d_desc);
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
  const cuuint32_t box[3] = {ORB_BOX_W, ORB_BOX_H, 1};
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
