// pslam_internal.cuh -- context layout and device helpers shared by the sm_100a kernels.
#pragma once
#include <cuda.h>  // CUtensorMap (types only: the driver entry point is fetched through the runtime)
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/pslam_cuda.h"

#define PSLAM_MAX_FEATURES_HARD 8192
// scratch partition: [0, PSLAM_SOLVER_SCRATCH_OFFSET) holds the projective finder's cached clouds / lattice (they persist
// between set_fixed / set_moving / match calls), the stage-3/4 entry points carve their temporaries above it so that an
// aligner loop can interleave finder and solver calls
#define PSLAM_SOLVER_SCRATCH_OFFSET ((size_t) 8 << 20)

// ---- HBM-resident state of one context ------------------------------------------------------
// Layout (all sized once from pslam_limits; nothing is allocated on the hot path):
// Stage 1 runs in chunks of `work_images` images so that the intermediate maps of a chunk stay
// L2-resident between the kernels that produce and consume them; only the stores are batch-sized.
//   images   u8  [2][work_images][max_rows][img_pitch]      double-buffered staging (host-pointer entry points)
//   row_kp   u32 [work_images][max_rows][map_pitch]         per-row keypoint lists after NMS, (col << 8) | response + 1,
//   row_count int [work_images][max_rows]                   ordered by column (only the used prefix is ever touched)
//   blur     u8  [work_images][max_rows][map_pitch]         ORB 7x7 integer Gaussian
//   raw      u32 [work_images][max_bins][max_raw_per_bin]   (pixel index << 8 | response + 1), row-major per bin
//   features SoA [max_images][max_features]: xy float2, response f32, intensity f32, desc 8 x u32
//   stereo   SoA [max_images/2][max_features]: uvuv float4, left/right feature index, distance
// map_pitch and img_pitch are multiples of 128 so every row starts on a 128 B line.
// One LANE = the chunk-level intermediates of the stage-1 pipeline plus the stream they are produced / consumed on.  A batch
// alternates its chunks between two lanes: chunk i + 1's detection kernel (issue bound, long CTAs) runs while chunk i's
// selection / description / matching kernels (latency / L2 bound, short CTAs) drain, so no SM idles in a kernel's tail.
struct pslam_lane {
  cudaStream_t stream;
  uint32_t* d_row_kp;
  int* d_row_count;
  uint8_t* d_blur;
  CUtensorMap blur_tmap;
  uint32_t* d_raw;
  int* d_raw_count;
  int* d_sel_count;
  cudaEvent_t ev_done;
};

struct pslam_ctx {
  int device;
  cudaStream_t stream;       // the ACTIVE lane's stream: every launcher enqueues here (lane 0 outside batched stage 1)
  pslam_lane lane[2];        // lane 1 is allocated by the first batch that spans more than one chunk
  int n_lanes, cur_lane, lanes_wanted;
  cudaStream_t copy_stream;  // host->device image uploads of the batched host entry point
  cudaEvent_t ev_ready[2], ev_free[2];  // double-buffered staging hand-shake
  int work_images;           // chunk size (images) of the stage-1 pipeline
  pslam_limits lim;
  char err[512];
  long long launches;
  int img_pitch, map_pitch;
  size_t img_slot, map_slot;  // bytes per image slot
  uint8_t* d_images;
  uint32_t* d_row_kp;      // K1 -> K2: [work_images][max_rows][strips_cap][256] keypoints after NMS, (col << 8) | response + 1
  int* d_row_count;        // [work_images][max_rows][strips_cap]
  int strips_cap;          // 252-pixel strips of a max_cols wide image (k_fast.cu)
  size_t k1_smem_set;
  int* d_sel_bounds;      // K2: per region (row begin, row end, col begin, col end), valid for sel_bounds_key
  int sel_bounds_key[4];  // rows, cols, nh, nv the table was built for
  uint8_t* d_blur;
  CUtensorMap blur_tmap;  // TMA view of the blur maps: u8 [work_images][max_rows][map_pitch], box = one description tile (k_detect.cu)
  // keypoints bucketed by description tile (written by assemble_features_kernel, read by orb_describe_tiles_kernel)
  int tile_cap;                   // tiles of a max_rows x max_cols image
  int* d_tile_start;              // [max_images][tile_cap + 1]
  uint32_t* d_tile_order;         // [max_images][max_features] (keypoint index | x in tile << 13 | y in tile << 20), grouped by tile
  uint8_t* d_mask;  // [max_rows][map_pitch], single image (host entry point only)
  uint32_t* d_raw;
  int* d_raw_count;
  int* d_sel_count;
  float2* d_xy;
  float* d_resp;
  float* d_inten;
  uint32_t* d_desc;
  int* d_count;
  // stereo results
  float4* d_st_uvuv;
  int* d_st_left;
  int* d_st_right;
  float* d_st_dist;
  int* d_st_count;
  // raw epipolar correspondences (before the disparity filter), per pair
  int* d_ep_fixed;
  int* d_ep_moving;
  float* d_ep_dist;
  int* d_ep_count;
  int* d_flags;  // [0] capacity overflow bits
  // generic scratch for the host-pointer matchers / solver
  uint8_t* d_scratch;
  size_t scratch_bytes;
  void* h_pinned;  // small pinned staging (counts, flags, H/b)
  size_t pinned_bytes;
  uint8_t* h_frame_stage;  // per-frame adaptor: pinned staging of the two images + the packed result (allocated on first use)
  size_t frame_stage_bytes;
  // projective finder cache (fixed cloud + row-sorted lattice, moving cloud): its OWN allocation (made on first use), so
  // that no other entry point's scratch use can clobber it; the epochs change with every upload and are unique per
  // process, so a caller can tell whether the cache still holds what IT uploaded (pslam_projective_cache_epochs)
  uint8_t* d_proj;
  size_t proj_bytes;
  unsigned long long proj_fixed_epoch, proj_moving_epoch;
  int proj_fixed_dim;  // floats per fixed point of the cached cloud
  int proj_n_fixed, proj_n_moving;  // sizes of the cached clouds: a cloud's descriptors sit directly behind its coordinates (ONE upload)
  unsigned long long proj_weights_epoch;  // == proj_moving_epoch while the information-scale table belongs to the cached moving cloud
  // peer-to-peer result tables of the sharded Hamming sweep (k_sharded.cu): every rank owns one table in its HBM, exported
  // through CUDA IPC; the ranks' merge kernels store their rows straight into every peer's table over NVLink
  int* d_p2p_table;        // [2 parities][3][p2p_cap_rows] ints + 64 flags (epoch of the last complete write, per source rank)
  int p2p_cap_rows, p2p_world, p2p_rank;
  int* p2p_peer[64];       // table base of every rank as seen from this device (own pointer for p2p_rank)
  int** d_p2p_peers;       // the same array on the device
  int p2p_epoch;
  // geometry of the last batch
  int rows, cols, n_images;
  // optional per-kernel device timing (pslam_profile_*): one CUDA event after every launch; the interval
  // between consecutive events on the in-order stream is attributed to the kernel that ends it
  int prof_enabled;
  cudaEvent_t* prof_ev;
  const char** prof_name;
  int* prof_lane;
  int prof_n, prof_cap;
};

// make lane `i` the active one: its stream and chunk-level buffers become the ones the launchers see
static inline void pslam_use_lane(pslam_ctx* ctx, int i) {
  pslam_lane& L = ctx->lane[i];
  ctx->cur_lane = i;
  ctx->stream = L.stream;
  ctx->d_row_kp = L.d_row_kp;
  ctx->d_row_count = L.d_row_count;
  ctx->d_blur = L.d_blur;
  ctx->blur_tmap = L.blur_tmap;
  ctx->d_raw = L.d_raw;
  ctx->d_raw_count = L.d_raw_count;
  ctx->d_sel_count = L.d_sel_count;
}

void pslam_prof_mark(pslam_ctx* ctx, const char* name);  // pslam_capi.cu

#define PSLAM_FLAG_RAW_OVERFLOW 1
#define PSLAM_FLAG_FEATURE_OVERFLOW 2
#define PSLAM_FLAG_CANDIDATE_OVERFLOW 4
#define PSLAM_FLAG_P2P_TIMEOUT 8

static inline int pslam_set_error(pslam_ctx* ctx, int code, const char* what, cudaError_t e) {
  if (ctx) {
    if (e != cudaSuccess)
      snprintf(ctx->err, sizeof(ctx->err), "%s: %s", what, cudaGetErrorString(e));
    else
      snprintf(ctx->err, sizeof(ctx->err), "%s", what);
  }
  return code;
}

#define PSLAM_CUDA_TRY(ctx, call)                                              \
  do {                                                                         \
    cudaError_t e__ = (call);                                                  \
    if (e__ != cudaSuccess) return pslam_set_error(ctx, PSLAM_E_CUDA, #call, e__); \
  } while (0)

#define PSLAM_LAUNCH_CHECK(ctx, name)                                          \
  do {                                                                         \
    (ctx)->launches++;                                                         \
    if ((ctx)->prof_enabled) pslam_prof_mark(ctx, name);                       \
    cudaError_t e__ = cudaGetLastError();                                      \
    if (e__ != cudaSuccess) return pslam_set_error(ctx, PSLAM_E_CUDA, name, e__); \
  } while (0)

#ifdef __CUDACC__
// ---- programmatic dependent launch --------------------------------------------------------------
// The per-frame paths are chains of short dependent kernels on one stream; between two of them the GPU idles for the launch
// latency of the second (~2 us).  Launched with pslam_launch_pdl, a kernel is scheduled while its predecessor still runs
// and parks in pslam_pdl_enter() -- its first statement -- until the predecessor has completed and its writes are visible;
// only the launch overlaps, never the work.  (A kernel launched the ordinary way passes through pslam_pdl_enter() at once.)
__device__ __forceinline__ void pslam_pdl_enter() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}
template <typename... KArgs, typename... Args>
static inline cudaError_t pslam_launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                           Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ---- device helpers ---------------------------------------------------------------------------
__device__ __forceinline__ int reflect101(int i, int n) {
  // single reflection is enough: |overhang| <= 4 << n
  if (i < 0) i = -i;
  if (i >= n) i = 2 * (n - 1) - i;
  return i;
}

// exclusive scan of one int per thread over a block of NT threads (NT multiple of 32, <= 1024);
// returns the exclusive prefix, *total receives the block sum.  s_warp: >= 33 ints of smem.
template <int NT>
__device__ __forceinline__ int block_exclusive_scan(int v, int* s_warp, int* total) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  __syncthreads();  // protect s_warp reuse across calls
  if (lane == 31) s_warp[wid] = incl;
  __syncthreads();
  if (wid == 0) {
    int w = (lane < NT / 32) ? s_warp[lane] : 0;
    int wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o) wi += t;
    }
    if (lane < NT / 32) s_warp[lane] = wi - w;  // exclusive warp offsets
    if (lane == 31) s_warp[32] = wi;
  }
  __syncthreads();
  *total = s_warp[32];
  return s_warp[wid] + incl - v;
}

__device__ __forceinline__ int hamming256(const uint4 a0, const uint4 a1, const uint4 b0, const uint4 b1) {
  return __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
         __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
}
#endif
