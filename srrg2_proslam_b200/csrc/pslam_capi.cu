// pslam_capi.cu -- the C ABI (include/pslam_cuda.h): context management and the entry points that
// sequence the sm_100a kernels on the context's stream.  No CPU fallback exists: every compute entry
// point launches kernels; without a usable CUDA device pslam_create fails with PSLAM_E_CUDA.
#include <math.h>
#include <stdlib.h>

#include <new>
#include <vector>

#include "pslam_internal.cuh"
#include "pslam_kernels.cuh"

void pslam_prof_mark(pslam_ctx* ctx, const char* name) {
  if (ctx->prof_n == ctx->prof_cap) {
    const int cap = ctx->prof_cap ? 2 * ctx->prof_cap : 1024;
    cudaEvent_t* ev = (cudaEvent_t*) realloc(ctx->prof_ev, sizeof(cudaEvent_t) * cap);
    const char** nm = (const char**) realloc(ctx->prof_name, sizeof(char*) * cap);
    int* ln = (int*) realloc(ctx->prof_lane, sizeof(int) * cap);
    if (!ev || !nm || !ln) {
      if (ev) ctx->prof_ev = ev;
      if (nm) ctx->prof_name = nm;
      if (ln) ctx->prof_lane = ln;
      return;
    }
    ctx->prof_ev = ev;
    ctx->prof_name = nm;
    ctx->prof_lane = ln;
    for (int i = ctx->prof_cap; i < cap; ++i) cudaEventCreate(&ctx->prof_ev[i]);
    ctx->prof_cap = cap;
  }
  cudaEventRecord(ctx->prof_ev[ctx->prof_n], ctx->stream);
  ctx->prof_name[ctx->prof_n] = name;
  ctx->prof_lane[ctx->prof_n] = ctx->cur_lane;
  ctx->prof_n++;
}

namespace {

int round_up(int x, int m) { return (x + m - 1) / m * m; }

template <typename T>
cudaError_t dmalloc_t(T** p, size_t n) { return cudaMalloc((void**) p, n * sizeof(T)); }

// chunk-level intermediates + stream of lane `i` (lane 0: pslam_create; lane 1: the first multi-chunk batch)
int alloc_lane(pslam_ctx* ctx, int i) {
  pslam_lane& L = ctx->lane[i];
  const size_t NW = ctx->work_images;
  const pslam_limits& lim = ctx->lim;
  PSLAM_CUDA_TRY(ctx, cudaStreamCreateWithFlags(&L.stream, cudaStreamNonBlocking));
  PSLAM_CUDA_TRY(ctx, cudaEventCreateWithFlags(&L.ev_done, cudaEventDisableTiming));
  PSLAM_CUDA_TRY(ctx, dmalloc_t(&L.d_row_kp, NW * lim.max_rows * (size_t) ctx->strips_cap * 256));
  PSLAM_CUDA_TRY(ctx, dmalloc_t(&L.d_row_count, NW * lim.max_rows * (size_t) ctx->strips_cap));
  PSLAM_CUDA_TRY(ctx, dmalloc_t(&L.d_blur, NW * ctx->map_slot));
  PSLAM_CUDA_TRY(ctx, dmalloc_t(&L.d_raw, NW * lim.max_bins * (size_t) lim.max_raw_per_bin));
  PSLAM_CUDA_TRY(ctx, dmalloc_t(&L.d_raw_count, NW * lim.max_bins));
  PSLAM_CUDA_TRY(ctx, dmalloc_t(&L.d_sel_count, NW * lim.max_bins));
  const int back = ctx->cur_lane;
  if (i + 1 > ctx->n_lanes) ctx->n_lanes = i + 1;
  pslam_use_lane(ctx, i);
  const int rc = pslam_k_make_blur_tmap(ctx, (int) NW);  // encodes ctx->d_blur into ctx->blur_tmap
  L.blur_tmap = ctx->blur_tmap;
  pslam_use_lane(ctx, back);
  return rc;
}

// A batch that spans several chunks alternates them between the two lanes; the guard joins lane 1 back into lane 0 (the
// context's main stream) and makes lane 0 the active one again on every exit path.
struct LaneScope {
  pslam_ctx* ctx;
  bool multi;
  LaneScope(pslam_ctx* c, int n_images) : ctx(c), multi(false) {
    if (n_images <= c->work_images || c->lanes_wanted < 2 || (c->work_images & 1)) return;
    if (c->n_lanes < 2 && alloc_lane(c, 1) != PSLAM_OK) return;  // out of memory: stay on one lane
    // lane 1 starts after everything already queued on the main stream
    if (cudaEventRecord(c->lane[0].ev_done, c->lane[0].stream) != cudaSuccess ||
        cudaStreamWaitEvent(c->lane[1].stream, c->lane[0].ev_done, 0) != cudaSuccess)
      return;
    multi = true;
  }
  void use(int chunk) {
    if (multi) pslam_use_lane(ctx, chunk & 1);
  }
  ~LaneScope() {
    if (!multi) return;
    cudaEventRecord(ctx->lane[1].ev_done, ctx->lane[1].stream);
    pslam_use_lane(ctx, 0);
    cudaStreamWaitEvent(ctx->lane[0].stream, ctx->lane[1].ev_done, 0);
  }
};

template <typename T>
cudaError_t dmalloc(T** p, size_t n) {
  return cudaMalloc(reinterpret_cast<void**>(p), n * sizeof(T));
}

int check_flags(pslam_ctx* ctx) {
  int* h = reinterpret_cast<int*>(ctx->h_pinned);
  PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(h, ctx->d_flags, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  PSLAM_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  if (h[0]) {
    char msg[160];
    snprintf(msg, sizeof(msg), "capacity exceeded (flags=%d: 1 max_raw_per_bin, 2 max_features, 4 candidates, 8 a peer never signalled its rows of a sharded sweep)", h[0]);
    PSLAM_CUDA_TRY(ctx, cudaMemsetAsync(ctx->d_flags, 0, sizeof(int), ctx->stream));
    return pslam_set_error(ctx, PSLAM_E_CAPACITY, msg, cudaSuccess);
  }
  return PSLAM_OK;
}

int validate_extract(pslam_ctx* ctx, int n_images, int rows, int cols, const pslam_extract_cfg* cfg) {
  if (!ctx || !cfg) return PSLAM_E_INVALID;
  if (n_images < 0 || n_images > ctx->lim.max_images || rows <= 0 || cols <= 0 ||
      rows > ctx->lim.max_rows || cols > ctx->lim.max_cols)
    return pslam_set_error(ctx, PSLAM_E_CAPACITY, "image batch exceeds the context limits", cudaSuccess);
  // reference: IntensityFeatureExtractorBinned_::init throws on these (binned.cpp:13-28)
  if (cfg->number_of_detectors_vertical <= 0)
    return pslam_set_error(ctx, PSLAM_E_INVALID, "IntensityFeatureExtractor::init|ERROR: invalid number of vertical detectors (check configuration!)", cudaSuccess);
  if (cfg->number_of_detectors_horizontal <= 0)
    return pslam_set_error(ctx, PSLAM_E_INVALID, "IntensityFeatureExtractor::init|ERROR: invalid number of horizontal detectors (check configuration!)", cudaSuccess);
  if (cfg->number_of_detectors_horizontal * cfg->number_of_detectors_vertical > ctx->lim.max_bins)
    return pslam_set_error(ctx, PSLAM_E_CAPACITY, "more detection regions than pslam_limits.max_bins", cudaSuccess);
  if (rows < 8 || cols < 8)
    return pslam_set_error(ctx, PSLAM_E_INVALID, "image smaller than 8x8", cudaSuccess);
  if ((long long) rows * cols >= (1LL << 24) || cols > 4032)
    return pslam_set_error(ctx, PSLAM_E_CAPACITY, "image larger than 2^24 pixels or wider than 4032", cudaSuccess);
  return PSLAM_OK;
}

// quota + region grid of IntensityFeatureExtractorBinned_::init (binned.cpp:72-75), mask => no binning (:167)
struct SelectPlan {
  int nh, nv;
  unsigned long long quota;
};
SelectPlan make_plan(const pslam_extract_cfg* cfg, bool masked) {
  SelectPlan p;
  p.nh = cfg->number_of_detectors_horizontal;
  p.nv = cfg->number_of_detectors_vertical;
  const float qf = static_cast<float>(cfg->target_number_of_keypoints) / static_cast<float>((size_t) (p.nh * p.nv));
  p.quota = qf <= 0.0f ? 0ULL : (unsigned long long) qf;
  if (masked) {
    p.nh = p.nv = 1;
    p.quota = ~0ULL;
  }
  return p;
}

// detect -> select -> assemble -> describe for ONE chunk (<= work_images images) already on the device;
// features land in store slots slot_base .. slot_base + n_images - 1
int run_extract_chunk(pslam_ctx* ctx, const uint8_t* d_images, long long image_pitch, int n_images, int rows,
                      int cols, int stride, const pslam_extract_cfg* cfg, const uint8_t* d_mask, int slot_base) {
  int rc;
  if ((rc = pslam_k_fast_blur(ctx, d_images, image_pitch, n_images, rows, cols, stride,
                              (int) cfg->detector_threshold, cfg->enable_non_maximum_suppression))) return rc;
  const SelectPlan plan = make_plan(cfg, d_mask != nullptr);
  if ((rc = pslam_k_bin_select(ctx, n_images, rows, cols, plan.nh, plan.nv, plan.quota, d_mask, 0))) return rc;
  if ((rc = pslam_k_assemble(ctx, d_images, image_pitch, stride, n_images, rows, cols, plan.nh * plan.nv, 31, slot_base))) return rc;
  if ((rc = pslam_k_describe(ctx, n_images, rows, cols, slot_base))) return rc;
  return PSLAM_OK;
}

// the whole batch, device-resident images, in chunks of work_images
// `mcfg` != nullptr (stereo batches, even work_images): the stereo pairs of a chunk are matched right after the chunk's
// descriptors were written, while they are still L2 resident (and, for host images, while the next chunk uploads)
int run_extract(pslam_ctx* ctx, const uint8_t* d_images, long long image_pitch, int n_images, int rows,
                int cols, int stride, const pslam_extract_cfg* cfg, const uint8_t* d_mask, const pslam_match_cfg* mcfg = nullptr) {
  ctx->rows = rows;
  ctx->cols = cols;
  ctx->n_images = n_images;
  LaneScope lanes(ctx, n_images);
  int chunk = 0;
  for (int base = 0; base < n_images; base += ctx->work_images, ++chunk) {
    const int n = n_images - base < ctx->work_images ? n_images - base : ctx->work_images;
    lanes.use(chunk);
    int rc = run_extract_chunk(ctx, d_images + (size_t) base * image_pitch, image_pitch, n, rows, cols, stride, cfg, d_mask, base);
    if (rc) return rc;
    if (mcfg && n / 2 > 0 && (rc = pslam_k_epipolar(ctx, n / 2, mcfg, 0, base / 2))) return rc;
  }
  return PSLAM_OK;
}

// the whole batch, HOST images: uploads on copy_stream into the double-buffered staging area overlap the
// kernels of the previous chunk on the compute stream
int run_extract_host(pslam_ctx* ctx, const uint8_t* h_images, long long image_pitch, int n_images, int rows,
                     int cols, int stride, const pslam_extract_cfg* cfg, const pslam_match_cfg* mcfg = nullptr) {
  ctx->rows = rows;
  ctx->cols = cols;
  ctx->n_images = n_images;
  const size_t half = (size_t) ctx->work_images * ctx->img_slot;
  // the staging buffers may still be read by kernels of an earlier call: order the uploads after them
  PSLAM_CUDA_TRY(ctx, cudaEventRecord(ctx->ev_free[0], ctx->stream));
  PSLAM_CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_free[0], 0));
  LaneScope lanes(ctx, n_images);
  int chunk = 0;
  for (int base = 0; base < n_images; base += ctx->work_images, ++chunk) {
    const int n = n_images - base < ctx->work_images ? n_images - base : ctx->work_images;
    const int buf = chunk & 1;
    lanes.use(chunk);
    uint8_t* stage = ctx->d_images + buf * half;
    if (chunk >= 2) PSLAM_CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_free[buf], 0));
    const uint8_t* src = h_images + (size_t) base * image_pitch;
    // A chunk whose images are evenly spaced in host memory and fit the staging slots travels as ONE
    // linear copy and is processed in its host layout (stride / pitch as given): a 2-D re-pitching copy
    // of 1241-byte rows runs at a fifth of the PCIe rate.
    const bool linear = image_pitch >= (long long) (rows - 1) * stride + cols && (size_t) image_pitch <= ctx->img_slot;
    const uint8_t* d_src = stage;
    long long d_pitch = (long long) ctx->img_slot;
    int d_stride = ctx->img_pitch;
    if (linear) {
      const size_t bytes = (size_t) (n - 1) * image_pitch + (size_t) (rows - 1) * stride + cols;
      PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(stage, src, bytes, cudaMemcpyHostToDevice, ctx->copy_stream));
      d_pitch = image_pitch;
      d_stride = stride;
    } else {
      for (int i = 0; i < n; ++i)
        PSLAM_CUDA_TRY(ctx, cudaMemcpy2DAsync(stage + (size_t) i * ctx->img_slot, ctx->img_pitch, src + (size_t) i * image_pitch, stride, cols, rows, cudaMemcpyHostToDevice, ctx->copy_stream));
    }
    PSLAM_CUDA_TRY(ctx, cudaEventRecord(ctx->ev_ready[buf], ctx->copy_stream));
    PSLAM_CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_ready[buf], 0));
    if (ctx->prof_enabled) pslam_prof_mark(ctx, nullptr);  // the wait for the upload is not kernel time
    int rc = run_extract_chunk(ctx, d_src, d_pitch, n, rows, cols, d_stride, cfg, nullptr, base);
    if (rc) return rc;
    PSLAM_CUDA_TRY(ctx, cudaEventRecord(ctx->ev_free[buf], ctx->stream));  // the matcher below does not read the staging buffer
    if (mcfg && n / 2 > 0 && (rc = pslam_k_epipolar(ctx, n / 2, mcfg, 0, base / 2))) return rc;
  }
  return PSLAM_OK;
}

// single-image / single-pair entry points: upload into staging buffer 0 on the compute stream
// one host image into staging slot `slot` (per-frame entry points).  An image that fits the slot in its host layout
// travels as ONE linear copy and is processed with its host stride: the 2-D re-pitching copy of 1241-byte rows cost
// ~230 us per KITTI image from pageable memory, three quarters of the per-frame adaptor latency.
static int upload_one(pslam_ctx* ctx, int slot, const uint8_t* h, int rows, int cols, int stride, int* d_stride) {
  uint8_t* dst = ctx->d_images + (size_t) slot * ctx->img_slot;
  const size_t bytes = (size_t) (rows - 1) * stride + cols;
  if (stride >= cols && bytes <= ctx->img_slot) {
    PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(dst, h, bytes, cudaMemcpyHostToDevice, ctx->stream));
    *d_stride = stride;
  } else {
    PSLAM_CUDA_TRY(ctx, cudaMemcpy2DAsync(dst, ctx->img_pitch, h, stride, cols, rows, cudaMemcpyHostToDevice, ctx->stream));
    *d_stride = ctx->img_pitch;
  }
  return PSLAM_OK;
}

int upload_images(pslam_ctx* ctx, const uint8_t* h, int n_images, int rows, int cols, int stride,
                  long long image_pitch_bytes) {
  if (n_images > ctx->work_images)
    return pslam_set_error(ctx, PSLAM_E_CAPACITY, "more images than pslam_limits.work_images", cudaSuccess);
  for (int i = 0; i < n_images; ++i) {
    PSLAM_CUDA_TRY(ctx, cudaMemcpy2DAsync(ctx->d_images + (size_t) i * ctx->img_slot, ctx->img_pitch,
                                          h + (size_t) i * image_pitch_bytes, stride, cols, rows,
                                          cudaMemcpyHostToDevice, ctx->stream));
  }
  return PSLAM_OK;
}

}  // namespace

extern "C" {

const char* pslam_version(void) { return "srrg2_proslam_b200 0.1 (sm_100a)"; }

int pslam_create(int device, const pslam_limits* lim, pslam_ctx** out) {
  if (!lim || !out) return PSLAM_E_INVALID;
  *out = nullptr;
  if (lim->max_images <= 0 || lim->max_rows <= 0 || lim->max_cols <= 0 || lim->max_features <= 0 ||
      lim->max_features > PSLAM_MAX_FEATURES_HARD || lim->max_raw_per_bin <= 0 || lim->max_bins <= 0)
    return PSLAM_E_INVALID;
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev <= 0 || device < 0 || device >= n_dev) {
    cudaGetLastError();
    return PSLAM_E_CUDA;  // no CPU fallback
  }
  pslam_ctx* ctx = new (std::nothrow) pslam_ctx();
  if (!ctx) return PSLAM_E_INVALID;
  memset(ctx, 0, sizeof(*ctx));
  ctx->device = device;
  ctx->lim = *lim;
  // per-image capacities are strides of 16-bit / 32-bit shared-memory arrays in the matchers (k_epipolar.cu): an odd
  // capacity would misalign them.  The capacity is an upper bound, so it is rounded up (8 <= PSLAM_MAX_FEATURES_HARD's grain)
  ctx->lim.max_features = (lim->max_features + 7) & ~7;
  *out = ctx;  // returned even on failure so that pslam_last_error is readable; caller destroys
  PSLAM_CUDA_TRY(ctx, cudaSetDevice(device));
  PSLAM_CUDA_TRY(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
  {
    const char* e = getenv("PSLAM_LANES");  // 1 = chunks strictly one after the other on one stream (tuning / debugging)
    ctx->lanes_wanted = e ? atoi(e) : 2;
  }
  for (int i = 0; i < 2; ++i) {
    PSLAM_CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->ev_ready[i], cudaEventDisableTiming));
    PSLAM_CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->ev_free[i], cudaEventDisableTiming));
  }
  ctx->work_images = lim->work_images > 0 ? lim->work_images : (lim->max_images < 512 ? lim->max_images : 512);
  if (ctx->work_images > lim->max_images) ctx->work_images = lim->max_images;
  if (ctx->work_images > 65535) ctx->work_images = 65535;  // gridDim.z / gridDim.y
  ctx->img_pitch = round_up(lim->max_cols, 128);
  ctx->map_pitch = round_up(lim->max_cols, 128);
  ctx->img_slot = (size_t) ctx->img_pitch * lim->max_rows;
  ctx->map_slot = (size_t) ctx->map_pitch * lim->max_rows;
  const size_t NI = lim->max_images, MF = ctx->lim.max_features, NP = (lim->max_images + 1) / 2;
  const size_t NW = ctx->work_images;
  PSLAM_CUDA_TRY(ctx, dmalloc(&ctx->d_images, 2 * NW * ctx->img_slot));
  ctx->strips_cap = pslam_k_strips_cap(lim->max_cols);
  PSLAM_CUDA_TRY(ctx, dmalloc(&ctx->d_mask, ctx->map_slot));
  PSLAM_CUDA_TRY(ctx, dmalloc(&ctx->d_xy, NI * MF));
  PSLAM_CUDA_TRY(ctx, dmalloc(&ctx->d_resp, NI * MF));
  PSLAM_CUDA_TRY(ctx, dmalloc(&ctx->d_inten, NI * MF));
  PSLAM_CUDA_TRY(ctx, dmalloc(&ctx->d_desc, NI * MF * 8));
  PSLAM_CUDA_TRY(ctx, dmalloc(&ctx->d_count, NI));
  ctx->tile_cap = pslam_k_orb_tile_cap(lim->max_rows, lim->max_cols);
  PSLAM_CUDA_TRY(ctx, dmalloc(&ctx->d_tile_start, NI * ((size_t) ctx->tile_cap + 1)));
  PSLAM_CUDA_TRY(ctx, dmalloc(&ctx->d_tile_order, NI * MF));
  PSLAM_CUDA_TRY(ctx, dmalloc(&ctx->d_st_uvuv, NP * MF));
  PSLAM_CUDA_TRY(ctx, dmalloc(&ctx->d_st_left, NP * MF));
  PSLAM_CUDA_TRY(ctx, dmalloc(&ctx->d_st_right, NP * MF));
  PSLAM_CUDA_TRY(ctx, dmalloc(&ctx->d_st_dist, NP * MF));
  PSLAM_CUDA_TRY(ctx, dmalloc(&ctx->d_st_count, NP));
  PSLAM_CUDA_TRY(ctx, dmalloc(&ctx->d_ep_fixed, NP * MF));
  PSLAM_CUDA_TRY(ctx, dmalloc(&ctx->d_ep_moving, NP * MF));
  PSLAM_CUDA_TRY(ctx, dmalloc(&ctx->d_ep_dist, NP * MF));
  PSLAM_CUDA_TRY(ctx, dmalloc(&ctx->d_ep_count, NP));
  PSLAM_CUDA_TRY(ctx, dmalloc(&ctx->d_flags, 4));
  PSLAM_CUDA_TRY(ctx, dmalloc(&ctx->d_sel_bounds, 4 * (size_t) lim->max_bins));
  // scratch: matcher / solver temporaries, and the packed (CSR) stereo result of a whole batch
  ctx->scratch_bytes = (size_t) 64 << 20;
  const size_t pack_bytes = NP * MF * 64 + NP * 8 + (1 << 20);
  if (NI > 2 && pack_bytes > ctx->scratch_bytes) ctx->scratch_bytes = pack_bytes;
  PSLAM_CUDA_TRY(ctx, dmalloc(&ctx->d_scratch, ctx->scratch_bytes));
  ctx->pinned_bytes = 1 << 20;
  PSLAM_CUDA_TRY(ctx, cudaMallocHost(&ctx->h_pinned, ctx->pinned_bytes));
  int rc = alloc_lane(ctx, 0);  // the main stream and the chunk-level intermediates
  if (rc) return rc;
  pslam_use_lane(ctx, 0);
  PSLAM_CUDA_TRY(ctx, cudaMemsetAsync(ctx->d_flags, 0, 4 * sizeof(int), ctx->stream));
  PSLAM_CUDA_TRY(ctx, cudaMemsetAsync(ctx->d_count, 0, NI * sizeof(int), ctx->stream));
  if ((rc = pslam_k_upload_pattern(ctx))) return rc;
  PSLAM_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return PSLAM_OK;
}

void pslam_destroy(pslam_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  for (int l = 0; l < 2; ++l) {
    pslam_lane& L = ctx->lane[l];
    if (L.stream) cudaStreamSynchronize(L.stream);
    void* lb[] = {L.d_row_kp, L.d_row_count, L.d_blur, L.d_raw, L.d_raw_count, L.d_sel_count};
    for (void* b : lb)
      if (b) cudaFree(b);
    if (L.ev_done) cudaEventDestroy(L.ev_done);
    if (L.stream) cudaStreamDestroy(L.stream);
  }
  void* bufs[] = {ctx->d_images, ctx->d_mask, ctx->d_xy, ctx->d_resp, ctx->d_inten, ctx->d_desc, ctx->d_count,
                  ctx->d_st_uvuv, ctx->d_st_left, ctx->d_st_right, ctx->d_st_dist, ctx->d_st_count,
                  ctx->d_ep_fixed, ctx->d_ep_moving, ctx->d_ep_dist, ctx->d_ep_count, ctx->d_flags,
                  ctx->d_sel_bounds, ctx->d_scratch, ctx->d_proj, ctx->d_tile_start, ctx->d_tile_order};
  for (void* b : bufs)
    if (b) cudaFree(b);
  pslam_p2p_table_release(ctx);
  if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
  if (ctx->h_frame_stage) cudaFreeHost(ctx->h_frame_stage);
  for (int i = 0; i < ctx->prof_cap; ++i) cudaEventDestroy(ctx->prof_ev[i]);
  free(ctx->prof_ev);
  free(ctx->prof_name);
  free(ctx->prof_lane);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  for (int i = 0; i < 2; ++i) {
    if (ctx->ev_ready[i]) cudaEventDestroy(ctx->ev_ready[i]);
    if (ctx->ev_free[i]) cudaEventDestroy(ctx->ev_free[i]);
  }
  delete ctx;
}

const char* pslam_last_error(const pslam_ctx* ctx) { return ctx ? ctx->err : "null context"; }
long long pslam_launch_count(const pslam_ctx* ctx) { return ctx ? ctx->launches : 0; }
void* pslam_stream(const pslam_ctx* ctx) { return ctx ? (void*) ctx->lane[0].stream : nullptr; }
int pslam_set_lanes(pslam_ctx* ctx, int n_lanes) {
  if (!ctx || n_lanes < 1 || n_lanes > 2) return PSLAM_E_INVALID;
  ctx->lanes_wanted = n_lanes;
  return PSLAM_OK;
}
int pslam_synchronize(pslam_ctx* ctx) {
  if (!ctx) return PSLAM_E_INVALID;
  PSLAM_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return PSLAM_OK;
}

// ---- per-kernel device timing ---------------------------------------------------------------------
int pslam_profile_enable(pslam_ctx* ctx, int enable) {
  if (!ctx) return PSLAM_E_INVALID;
  PSLAM_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->prof_enabled = enable ? 1 : 0;
  ctx->prof_n = 0;
  return PSLAM_OK;
}

int pslam_profile_mark(pslam_ctx* ctx) {
  if (!ctx) return PSLAM_E_INVALID;
  if (ctx->prof_enabled) pslam_prof_mark(ctx, nullptr);
  return PSLAM_OK;
}

int pslam_profile_read(pslam_ctx* ctx, int capacity, char* names, int name_len, double* total_ms,
                       long long* launches) {
  if (!ctx || capacity <= 0 || !names || name_len <= 1 || !total_ms || !launches) return PSLAM_E_INVALID;
  for (int l = 0; l < ctx->n_lanes; ++l) PSLAM_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->lane[l].stream));
  int n_names = 0;
  std::vector<const char*> seen;
  int prev_of_lane[2] = {-1, -1};
  for (int i = 0; i < ctx->prof_n; ++i) {
    // the interval between consecutive events of the SAME lane (in-order stream) belongs to the kernel that ends it; when two
    // lanes overlap, a kernel's interval includes the time it shared the GPU with the other lane's kernels
    const int prev = prev_of_lane[ctx->prof_lane[i]];
    prev_of_lane[ctx->prof_lane[i]] = i;
    const char* nm = ctx->prof_name[i];
    if (!nm || prev < 0) continue;  // interval ending at a marker: host gap / copies, not a kernel
    float ms = 0.0f;
    if (cudaEventElapsedTime(&ms, ctx->prof_ev[prev], ctx->prof_ev[i]) != cudaSuccess) {
      cudaGetLastError();
      continue;
    }
    int k = -1;
    for (int j = 0; j < n_names; ++j)
      if (seen[j] == nm || strcmp(seen[j], nm) == 0) { k = j; break; }
    if (k < 0) {
      if (n_names == capacity) continue;
      k = n_names++;
      seen.push_back(nm);
      snprintf(names + (size_t) k * name_len, name_len, "%s", nm);
      total_ms[k] = 0.0;
      launches[k] = 0;
    }
    total_ms[k] += ms;
    launches[k] += 1;
  }
  ctx->prof_n = 0;
  return n_names;
}

// ---- stage 1 ------------------------------------------------------------------------------------
int pslam_extract_binned_batch_dev(pslam_ctx* ctx, const uint8_t* d_images, int n_images, int rows,
                                   int cols, int stride, long long image_pitch_bytes,
                                   const pslam_extract_cfg* cfg) {
  int rc = validate_extract(ctx, n_images, rows, cols, cfg);
  if (rc) return rc;
  PSLAM_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  return run_extract(ctx, d_images, image_pitch_bytes, n_images, rows, cols, stride, cfg, nullptr);
}

int pslam_download_feature_counts(pslam_ctx* ctx, int n_images, int* counts) {
  if (!ctx || n_images > ctx->lim.max_images) return PSLAM_E_INVALID;
  int rc = check_flags(ctx);
  if (rc) return rc;
  PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(counts, ctx->d_count, sizeof(int) * n_images, cudaMemcpyDeviceToHost, ctx->stream));
  PSLAM_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return PSLAM_OK;
}

int pslam_download_features(pslam_ctx* ctx, int slot, int capacity, float* xy, float* response,
                            float* intensity, uint8_t* desc) {
  if (!ctx || slot < 0 || slot >= ctx->lim.max_images) return PSLAM_E_INVALID;
  int rc = check_flags(ctx);
  if (rc) return rc;
  int* h = reinterpret_cast<int*>(ctx->h_pinned);
  PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(h, ctx->d_count + slot, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  PSLAM_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  const int n = h[0];
  const int m = n < capacity ? n : capacity;
  const size_t o = (size_t) slot * ctx->lim.max_features;
  if (m > 0) {
    if (xy) PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(xy, ctx->d_xy + o, sizeof(float2) * m, cudaMemcpyDeviceToHost, ctx->stream));
    if (response) PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(response, ctx->d_resp + o, sizeof(float) * m, cudaMemcpyDeviceToHost, ctx->stream));
    if (intensity) PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(intensity, ctx->d_inten + o, sizeof(float) * m, cudaMemcpyDeviceToHost, ctx->stream));
    if (desc) PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(desc, ctx->d_desc + o * 8, 32 * (size_t) m, cudaMemcpyDeviceToHost, ctx->stream));
    PSLAM_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return n;
}

int pslam_extract_binned(pslam_ctx* ctx, const uint8_t* image, int rows, int cols, int stride,
                         const pslam_extract_cfg* cfg, const uint8_t* mask, int capacity, float* xy,
                         float* response, float* intensity, uint8_t* desc) {
  int rc = validate_extract(ctx, 1, rows, cols, cfg);
  if (rc) return rc;
  if (!image) return PSLAM_E_INVALID;
  PSLAM_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  int d_stride = 0;
  if ((rc = upload_one(ctx, 0, image, rows, cols, stride, &d_stride))) return rc;
  const uint8_t* d_mask = nullptr;
  if (mask) {
    PSLAM_CUDA_TRY(ctx, cudaMemsetAsync(ctx->d_mask, 0, ctx->map_slot, ctx->stream));
    PSLAM_CUDA_TRY(ctx, cudaMemcpy2DAsync(ctx->d_mask, ctx->map_pitch, mask, cols, cols, rows, cudaMemcpyHostToDevice, ctx->stream));
    d_mask = ctx->d_mask;
  }
  if ((rc = run_extract(ctx, ctx->d_images, (long long) ctx->img_slot, 1, rows, cols, d_stride, cfg, d_mask))) return rc;
  return pslam_download_features(ctx, 0, capacity, xy, response, intensity, desc);
}

int pslam_extract_selective(pslam_ctx* ctx, const uint8_t* image, int rows, int cols, int stride,
                            const pslam_extract_cfg* cfg, const uint8_t* tracking_mask, int enable_seeding,
                            int capacity, float* xy, float* response, float* intensity, uint8_t* desc,
                            int* n_tracking) {
  int rc = validate_extract(ctx, 2, rows, cols, cfg);
  if (rc) return rc;
  if (!image || !tracking_mask) return PSLAM_E_INVALID;
  PSLAM_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  if ((rc = upload_images(ctx, image, 1, rows, cols, stride, 0))) return rc;
  PSLAM_CUDA_TRY(ctx, cudaMemsetAsync(ctx->d_mask, 0, ctx->map_slot, ctx->stream));
  PSLAM_CUDA_TRY(ctx, cudaMemcpy2DAsync(ctx->d_mask, ctx->map_pitch, tracking_mask, cols, cols, rows, cudaMemcpyHostToDevice, ctx->stream));
  ctx->rows = rows;
  ctx->cols = cols;
  ctx->n_images = 1;
  // ONE FAST + blur pass; the tracking keypoints (mask != 0) go to feature slot 0, the seeding keypoints
  // (complement of the mask, selective.cpp:78-79) to slot 1 -- each in row-major order, no binning, no quota
  if ((rc = pslam_k_fast_blur(ctx, ctx->d_images, (long long) ctx->img_slot, 1, rows, cols, ctx->img_pitch,
                              (int) cfg->detector_threshold, cfg->enable_non_maximum_suppression))) return rc;
  const int passes = enable_seeding ? 2 : 1;
  for (int pass = 0; pass < passes; ++pass) {
    if ((rc = pslam_k_bin_select(ctx, 1, rows, cols, 1, 1, ~0ULL, ctx->d_mask, pass))) return rc;
    if ((rc = pslam_k_assemble(ctx, ctx->d_images, (long long) ctx->img_slot, ctx->img_pitch, 1, rows, cols, 1, 31, pass))) return rc;
    if ((rc = pslam_k_describe(ctx, 1, rows, cols, pass))) return rc;
  }
  const int n0 = pslam_download_features(ctx, 0, capacity, xy, response, intensity, desc);
  if (n0 < 0) return n0;
  if (n_tracking) *n_tracking = n0;
  int n1 = 0;
  if (enable_seeding) {
    const int used = n0 < capacity ? n0 : capacity;
    n1 = pslam_download_features(ctx, 1, capacity - used, xy ? xy + 2 * (size_t) used : nullptr,
                                 response ? response + used : nullptr, intensity ? intensity + used : nullptr,
                                 desc ? desc + 32 * (size_t) used : nullptr);
    if (n1 < 0) return n1;
  }
  return n0 + n1;
}

int pslam_fast_detect(pslam_ctx* ctx, const uint8_t* image, int rows, int cols, int stride,
                      int threshold, int nms, int capacity, float* xy, float* response) {
  pslam_extract_cfg cfg{(float) threshold, nms, 0, 1, 1};
  int rc = validate_extract(ctx, 1, rows, cols, &cfg);
  if (rc) return rc;
  PSLAM_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  if ((rc = upload_images(ctx, image, 1, rows, cols, stride, 0))) return rc;
  if ((rc = pslam_k_fast_blur(ctx, ctx->d_images, (long long) ctx->img_slot, 1, rows, cols, ctx->img_pitch, threshold, nms))) return rc;
  if ((rc = pslam_k_bin_select(ctx, 1, rows, cols, 1, 1, ~0ULL, nullptr, 0))) return rc;
  if ((rc = check_flags(ctx))) return rc;
  // one region, no quota: the region's raw list IS the row-major keypoint list (pixel << 8 | response + 1).
  // Read it back directly so that this stage-level entry point is bounded by max_raw_per_bin only.
  int* h = reinterpret_cast<int*>(ctx->h_pinned);
  PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(h, ctx->d_raw_count, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  PSLAM_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  const int n = h[0];
  const int m = n < capacity ? n : capacity;
  if (m > 0) {
    std::vector<uint32_t> raw((size_t) m);
    PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(raw.data(), ctx->d_raw, sizeof(uint32_t) * (size_t) m, cudaMemcpyDeviceToHost, ctx->stream));
    PSLAM_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < m; ++i) {
      const int pix = (int) (raw[i] >> 8);
      if (xy) {
        xy[2 * i] = (float) (pix % cols);
        xy[2 * i + 1] = (float) (pix / cols);
      }
      if (response) response[i] = (float) ((int) (raw[i] & 0xffu) - 1);
    }
  }
  return n;
}

int pslam_blur7(pslam_ctx* ctx, const uint8_t* image, int rows, int cols, int stride, uint8_t* out) {
  pslam_extract_cfg cfg{255.0f, 1, 0, 1, 1};
  int rc = validate_extract(ctx, 1, rows, cols, &cfg);
  if (rc) return rc;
  PSLAM_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  if ((rc = upload_images(ctx, image, 1, rows, cols, stride, 0))) return rc;
  if ((rc = pslam_k_fast_blur(ctx, ctx->d_images, (long long) ctx->img_slot, 1, rows, cols, ctx->img_pitch, 255, 1))) return rc;
  // the pipeline's blur map is exact except on the 3-pixel frame nothing reads; patch it for this entry point
  if ((rc = pslam_k_blur_border(ctx, ctx->d_images, rows, cols, ctx->img_pitch))) return rc;
  PSLAM_CUDA_TRY(ctx, cudaMemcpy2DAsync(out, cols, ctx->d_blur, ctx->map_pitch, cols, rows, cudaMemcpyDeviceToHost, ctx->stream));
  PSLAM_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return PSLAM_OK;
}

// ---- stage 2a -----------------------------------------------------------------------------------
static int upload_cloud(pslam_ctx* ctx, int slot, int n, const float* xy, const uint8_t* desc) {
  if (n > ctx->lim.max_features) return pslam_set_error(ctx, PSLAM_E_CAPACITY, "cloud larger than pslam_limits.max_features", cudaSuccess);
  const size_t o = (size_t) slot * ctx->lim.max_features;
  if (n > 0) {
    PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_xy + o, xy, sizeof(float2) * n, cudaMemcpyHostToDevice, ctx->stream));
    PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_desc + o * 8, desc, 32 * (size_t) n, cudaMemcpyHostToDevice, ctx->stream));
  }
  PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_count + slot, &n, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  PSLAM_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));  // &n is a stack variable
  return PSLAM_OK;
}

int pslam_match_epipolar(pslam_ctx* ctx, int n_fixed, const float* xy_fixed, const uint8_t* desc_fixed,
                         int n_moving, const float* xy_moving, const uint8_t* desc_moving,
                         const pslam_match_cfg* cfg, int capacity, int* fixed_idx, int* moving_idx,
                         float* distance) {
  if (!ctx || !cfg || n_fixed < 0 || n_moving < 0 || ctx->lim.max_images < 2) return PSLAM_E_INVALID;
  // Feature{int32 row, col} (epipolar_impl.cpp:8-20) travels as a packed sort key with 16-bit columns and 15-bit rows:
  // coordinates outside [0, 65536) x [0, 32768) would order differently from the reference's signed std::sort -- refuse
  auto in_range = [](int n, const float* xy) {
    for (int i = 0; i < n; ++i) {
      const float x = xy[2 * (size_t) i], y = xy[2 * (size_t) i + 1];
      if (!(x >= 0.f && x < 65536.f && y >= 0.f && y < 32768.f)) return false;
    }
    return true;
  };
  if (!in_range(n_fixed, xy_fixed) || !in_range(n_moving, xy_moving))
    return pslam_set_error(ctx, PSLAM_E_INVALID, "match_epipolar: coordinates outside [0, 65536) x [0, 32768)", cudaSuccess);
  PSLAM_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  int rc;
  if ((rc = upload_cloud(ctx, 0, n_fixed, xy_fixed, desc_fixed))) return rc;
  if ((rc = upload_cloud(ctx, 1, n_moving, xy_moving, desc_moving))) return rc;
  if ((rc = pslam_k_epipolar(ctx, 1, cfg, 1))) return rc;
  int* h = reinterpret_cast<int*>(ctx->h_pinned);
  PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(h, ctx->d_ep_count, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  PSLAM_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  const int n = h[0];
  const int m = n < capacity ? n : capacity;
  if (m > 0) {
    PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(fixed_idx, ctx->d_ep_fixed, sizeof(int) * m, cudaMemcpyDeviceToHost, ctx->stream));
    PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(moving_idx, ctx->d_ep_moving, sizeof(int) * m, cudaMemcpyDeviceToHost, ctx->stream));
    PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(distance, ctx->d_ep_dist, sizeof(float) * m, cudaMemcpyDeviceToHost, ctx->stream));
    PSLAM_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return n;
}

// ---- stage 1 + 2a -------------------------------------------------------------------------------
int pslam_stereo_frontend_batch_dev(pslam_ctx* ctx, const uint8_t* d_images, int n_pairs, int rows,
                                    int cols, int stride, long long image_pitch_bytes,
                                    const pslam_extract_cfg* ecfg, const pslam_match_cfg* mcfg) {
  if (!mcfg) return PSLAM_E_INVALID;
  int rc = validate_extract(ctx, 2 * n_pairs, rows, cols, ecfg);
  if (rc) return rc;
  PSLAM_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  if (ctx->prof_enabled) pslam_prof_mark(ctx, nullptr);
  // device-resident images: ONE matcher launch over all pairs after the last chunk (one CTA per pair: a launch per
  // 256-pair chunk leaves a third of the SMs idle and measured 0.94 instead of 0.60 us / image)
  if ((rc = run_extract(ctx, d_images, image_pitch_bytes, 2 * n_pairs, rows, cols, stride, ecfg, nullptr))) return rc;
  if (n_pairs > 0 && (rc = pslam_k_epipolar(ctx, n_pairs, mcfg, 0))) return rc;
  return PSLAM_OK;
}

int pslam_stereo_frontend_batch(pslam_ctx* ctx, const uint8_t* h_images, int n_pairs, int rows,
                                int cols, int stride, long long image_pitch_bytes,
                                const pslam_extract_cfg* ecfg, const pslam_match_cfg* mcfg) {
  if (!mcfg || !h_images) return PSLAM_E_INVALID;
  int rc = validate_extract(ctx, 2 * n_pairs, rows, cols, ecfg);
  if (rc) return rc;
  PSLAM_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  // host images: the pipeline is bound by the PCIe uploads, so every chunk's pairs are matched as soon as its
  // descriptors exist -- hidden behind the next chunk's upload instead of running after the last one (+6.5 % end to end)
  const bool per_chunk = (ctx->work_images & 1) == 0;  // pairs never straddle a chunk
  if ((rc = run_extract_host(ctx, h_images, image_pitch_bytes, 2 * n_pairs, rows, cols, stride, ecfg, per_chunk ? mcfg : nullptr))) return rc;
  if (!per_chunk && n_pairs > 0 && (rc = pslam_k_epipolar(ctx, n_pairs, mcfg, 0))) return rc;
  return PSLAM_OK;
}

int pslam_download_stereo_counts(pslam_ctx* ctx, int n_pairs, int* counts) {
  if (!ctx || 2 * n_pairs > ctx->lim.max_images + 1) return PSLAM_E_INVALID;
  int rc = check_flags(ctx);
  if (rc) return rc;
  PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(counts, ctx->d_st_count, sizeof(int) * n_pairs, cudaMemcpyDeviceToHost, ctx->stream));
  PSLAM_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return PSLAM_OK;
}

int pslam_download_stereo_points(pslam_ctx* ctx, int pair, int capacity, float* uvuv, int* left_idx,
                                 int* right_idx, float* distance) {
  if (!ctx || pair < 0 || 2 * pair >= ctx->lim.max_images) return PSLAM_E_INVALID;
  int rc = check_flags(ctx);
  if (rc) return rc;
  int* h = reinterpret_cast<int*>(ctx->h_pinned);
  PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(h, ctx->d_st_count + pair, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  PSLAM_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  const int n = h[0];
  const int m = n < capacity ? n : capacity;
  const size_t o = (size_t) pair * ctx->lim.max_features;
  if (m > 0) {
    if (uvuv) PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(uvuv, ctx->d_st_uvuv + o, sizeof(float4) * m, cudaMemcpyDeviceToHost, ctx->stream));
    if (left_idx) PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(left_idx, ctx->d_st_left + o, sizeof(int) * m, cudaMemcpyDeviceToHost, ctx->stream));
    if (right_idx) PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(right_idx, ctx->d_st_right + o, sizeof(int) * m, cudaMemcpyDeviceToHost, ctx->stream));
    if (distance) PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(distance, ctx->d_st_dist + o, sizeof(float) * m, cudaMemcpyDeviceToHost, ctx->stream));
    PSLAM_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return n;
}

int pslam_download_stereo_batch(pslam_ctx* ctx, int n_pairs, long long capacity_points, long long* offsets,
                                float* uvuv, float* intensity, uint8_t* desc, int* left_idx, int* right_idx,
                                float* distance) {
  if (!ctx || n_pairs < 0 || 2 * n_pairs > ctx->lim.max_images + 1 || !offsets) return PSLAM_E_INVALID;
  PSLAM_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  int rc = check_flags(ctx);
  if (rc) return rc;
  offsets[0] = 0;
  if (n_pairs == 0) return 0;
  pslam_packed_stereo pk;
  if ((rc = pslam_k_pack_stereo(ctx, n_pairs, &pk))) return rc;
  PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(offsets, pk.d_offsets, sizeof(long long) * ((size_t) n_pairs + 1), cudaMemcpyDeviceToHost, ctx->stream));
  PSLAM_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  const long long total = offsets[n_pairs];
  if (total > capacity_points)
    return pslam_set_error(ctx, PSLAM_E_CAPACITY, "download_stereo_batch: capacity_points too small", cudaSuccess);
  if (total > 0) {
    const size_t n = (size_t) total;
    if (uvuv) PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(uvuv, pk.d_uvuv, sizeof(float4) * n, cudaMemcpyDeviceToHost, ctx->stream));
    if (intensity) PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(intensity, pk.d_intensity, sizeof(float) * n, cudaMemcpyDeviceToHost, ctx->stream));
    if (desc) PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(desc, pk.d_desc, 32 * n, cudaMemcpyDeviceToHost, ctx->stream));
    if (left_idx) PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(left_idx, pk.d_left, sizeof(int) * n, cudaMemcpyDeviceToHost, ctx->stream));
    if (right_idx) PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(right_idx, pk.d_right, sizeof(int) * n, cudaMemcpyDeviceToHost, ctx->stream));
    if (distance) PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(distance, pk.d_dist, sizeof(float) * n, cudaMemcpyDeviceToHost, ctx->stream));
    PSLAM_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return total > 0x7fffffffLL ? 0x7fffffff : (int) total;
}

int pslam_stereo_adaptor(pslam_ctx* ctx, const uint8_t* left, const uint8_t* right, int rows, int cols,
                         int stride, const pslam_extract_cfg* ecfg, const pslam_match_cfg* mcfg,
                         int capacity, float* uvuv, float* intensity, uint8_t* desc) {
  if (!ctx || !left || !right || !mcfg) return PSLAM_E_INVALID;
  int rc = validate_extract(ctx, 2, rows, cols, ecfg);
  if (rc) return rc;
  PSLAM_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  // Per-frame latency path (one call per stereo frame of a sequence).  The images go through a pinned staging buffer so
  // that the copy of the left image is on the wire while the host stages the right one (a pageable cudaMemcpyAsync blocks
  // the host for its whole duration); the result comes back as the packed block of the batch path -- measurement
  // coordinates, left-feature intensity and descriptor gathered on the device -- with the capacity flags in ONE
  // synchronisation instead of six.
  const size_t M = ctx->lim.max_features;
  const size_t img_bytes = (size_t) (rows - 1) * stride + cols;
  const bool linear = stride >= cols && img_bytes <= ctx->img_slot;
  const size_t res_bytes = 512 + (sizeof(float4) + 32 + sizeof(float)) * M + 3 * 256;
  const size_t need = 2 * ctx->img_slot + res_bytes;
  if (!ctx->h_frame_stage || ctx->frame_stage_bytes < need) {
    if (ctx->h_frame_stage) cudaFreeHost(ctx->h_frame_stage);
    ctx->h_frame_stage = nullptr;
    PSLAM_CUDA_TRY(ctx, cudaMallocHost((void**) &ctx->h_frame_stage, need));
    ctx->frame_stage_bytes = need;
  }
  int d_stride = ctx->img_pitch;
  if (linear) {
    uint8_t* hs = ctx->h_frame_stage;
    memcpy(hs, left, img_bytes);
    PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_images, hs, img_bytes, cudaMemcpyHostToDevice, ctx->stream));
    memcpy(hs + ctx->img_slot, right, img_bytes);
    PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_images + ctx->img_slot, hs + ctx->img_slot, img_bytes, cudaMemcpyHostToDevice, ctx->stream));
    d_stride = stride;
  } else {
    int d_stride_r = 0;
    if ((rc = upload_one(ctx, 0, left, rows, cols, stride, &d_stride))) return rc;
    if ((rc = upload_one(ctx, 1, right, rows, cols, stride, &d_stride_r))) return rc;  // same stride, same decision
  }
  if ((rc = run_extract(ctx, ctx->d_images, (long long) ctx->img_slot, 2, rows, cols, d_stride, ecfg, nullptr))) return rc;
  if ((rc = pslam_k_epipolar(ctx, 1, mcfg, 0))) return rc;
  pslam_packed_stereo pk;
  if ((rc = pslam_k_pack_stereo(ctx, 1, &pk))) return rc;
  // [flags | offsets] [uvuv] [desc] [intensity]: the three regions are carved back to back by pslam_k_pack_stereo
  uint8_t* hr = ctx->h_frame_stage + 2 * ctx->img_slot;
  const size_t blk = (size_t) ((uint8_t*) pk.d_left - (uint8_t*) pk.d_uvuv);  // uvuv + desc + intensity regions (256-aligned)
  if (256 + 256 + blk > res_bytes) return pslam_set_error(ctx, PSLAM_E_CAPACITY, "stereo_adaptor: result staging too small", cudaSuccess);
  PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(hr, ctx->d_flags, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(hr + 256, pk.d_offsets, 2 * sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
  PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(hr + 512, pk.d_uvuv, blk, cudaMemcpyDeviceToHost, ctx->stream));
  PSLAM_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  const int flags = *reinterpret_cast<const int*>(hr);
  if (flags) {
    char msg[160];
    snprintf(msg, sizeof(msg), "capacity exceeded (flags=%d: 1 max_raw_per_bin, 2 max_features, 4 candidates)", flags);
    PSLAM_CUDA_TRY(ctx, cudaMemsetAsync(ctx->d_flags, 0, sizeof(int), ctx->stream));
    return pslam_set_error(ctx, PSLAM_E_CAPACITY, msg, cudaSuccess);
  }
  const int n = (int) reinterpret_cast<const long long*>(hr + 256)[1];
  const int m = n < capacity ? n : capacity;
  if (m > 0) {
    const uint8_t* base = hr + 512;
    if (uvuv) memcpy(uvuv, base, sizeof(float4) * (size_t) m);
    if (desc) memcpy(desc, base + ((uint8_t*) pk.d_desc - (uint8_t*) pk.d_uvuv), 32 * (size_t) m);
    if (intensity) memcpy(intensity, base + ((uint8_t*) pk.d_intensity - (uint8_t*) pk.d_uvuv), sizeof(float) * (size_t) m);
  }
  return n;
}


int pslam_mono_depth_adaptor(pslam_ctx* ctx, const uint8_t* image, int rows, int cols, int stride, const void* depth,
                             int depth_type, int depth_rows, int depth_cols, int depth_stride_elements,
                             float depth_scaling_factor_to_meters, const pslam_extract_cfg* ecfg, int capacity,
                             float* uvz, float* intensity, uint8_t* desc, int* n_features_in_image) {
  if (!ctx || !image || !depth) return PSLAM_E_INVALID;
  int rc = validate_extract(ctx, 1, rows, cols, ecfg);
  if (rc) return rc;
  if (depth_type != 0 && depth_type != 1)
    return pslam_set_error(ctx, PSLAM_E_INVALID, "RawDataPreprocessorMonocularDepth::compute|ERROR: unknown depth image type", cudaSuccess);
  if (depth_rows <= 0)
    return pslam_set_error(ctx, PSLAM_E_INVALID, "RawDataPreprocessorMonocularDepth::compute|ERROR: depth image has zero rows", cudaSuccess);
  if (depth_cols <= 0 || depth_stride_elements < depth_cols)
    return pslam_set_error(ctx, PSLAM_E_INVALID, "RawDataPreprocessorMonocularDepth::compute|ERROR: depth image has zero columns", cudaSuccess);
  PSLAM_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  const size_t esz = depth_type == 0 ? 2 : 4;
  const size_t depth_bytes = ((size_t) depth_rows * depth_stride_elements * esz + 255) & ~(size_t) 255;
  const size_t mf = (size_t) ctx->lim.max_features;
  const size_t need = depth_bytes + mf * (12 + 4 + 32) + 256;
  if (need > ctx->scratch_bytes) return pslam_set_error(ctx, PSLAM_E_CAPACITY, "mono depth adaptor: depth image exceeds the scratch buffer", cudaSuccess);
  uint8_t* s = ctx->d_scratch;
  void* d_depth = s;
  float* d_uvz = reinterpret_cast<float*>(s + depth_bytes);
  float* d_in = d_uvz + 3 * mf;
  uint32_t* d_de = reinterpret_cast<uint32_t*>(d_in + mf);
  int* d_n = reinterpret_cast<int*>(d_de + 8 * mf);
  int d_stride = 0;
  if ((rc = upload_one(ctx, 0, image, rows, cols, stride, &d_stride))) return rc;
  PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(d_depth, depth, (size_t) depth_rows * depth_stride_elements * esz, cudaMemcpyHostToDevice, ctx->stream));
  if ((rc = run_extract(ctx, ctx->d_images, (long long) ctx->img_slot, 1, rows, cols, d_stride, ecfg, nullptr))) return rc;
  if ((rc = pslam_k_mono_depth(ctx, d_depth, depth_type, depth_rows, depth_cols, depth_stride_elements,
                               depth_scaling_factor_to_meters, 0, d_uvz, d_in, d_de, d_n))) return rc;
  if ((rc = check_flags(ctx))) return rc;
  int* h = reinterpret_cast<int*>(ctx->h_pinned);
  PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(h, d_n, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(h + 1, ctx->d_count, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  PSLAM_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  const int n = h[0];
  if (n_features_in_image) *n_features_in_image = h[1];
  const int m = n < capacity ? n : capacity;
  if (m > 0) {
    if (uvz) PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(uvz, d_uvz, sizeof(float) * 3 * m, cudaMemcpyDeviceToHost, ctx->stream));
    if (intensity) PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(intensity, d_in, sizeof(float) * m, cudaMemcpyDeviceToHost, ctx->stream));
    if (desc) PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(desc, d_de, 32 * (size_t) m, cudaMemcpyDeviceToHost, ctx->stream));
    PSLAM_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return n;
}

// ---- N1: rigid-stereo triangulation ---------------------------------------------------------------
int pslam_triangulate(pslam_ctx* ctx, int n, const float* uvuv, const float* K9, float baseline_pixels_x,
                      float minimum_disparity_pixels, float infinity_depth_meters, float* xyz, uint8_t* valid) {
  if (!ctx || n < 0 || !K9 || (n > 0 && (!uvuv || !xyz))) return PSLAM_E_INVALID;
  if (n == 0) return 0;
  PSLAM_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  const size_t need = (size_t) n * (16 + 12 + 1) + 1024;
  if (need > ctx->scratch_bytes) return pslam_set_error(ctx, PSLAM_E_CAPACITY, "triangulate: too many points for the scratch buffer", cudaSuccess);
  float4* d_in = reinterpret_cast<float4*>(ctx->d_scratch);
  float* d_xyz = reinterpret_cast<float*>(ctx->d_scratch + (((size_t) n * 16 + 255) & ~(size_t) 255));
  unsigned char* d_valid = reinterpret_cast<unsigned char*>(d_xyz + 3 * (size_t) n);
  PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(d_in, uvuv, (size_t) n * 16, cudaMemcpyHostToDevice, ctx->stream));
  int rc = pslam_k_triangulate(ctx, d_in, n, K9, baseline_pixels_x, minimum_disparity_pixels, infinity_depth_meters, d_xyz, d_valid);
  if (rc) return rc;
  PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(xyz, d_xyz, (size_t) n * 12, cudaMemcpyDeviceToHost, ctx->stream));
  std::vector<unsigned char> v((size_t) n);
  PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(v.data(), d_valid, (size_t) n, cudaMemcpyDeviceToHost, ctx->stream));
  PSLAM_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  int n_valid = 0;
  for (int i = 0; i < n; ++i) {
    if (valid) valid[i] = v[i];
    n_valid += v[i];
  }
  return n_valid;
}

// ---- N2: projective scene clipping ----------------------------------------------------------------
// inverse of a row-major 3x4 isometry, fp32, operation order of the oracle's Pose::inverse
static void clip_invert(const float* T, float* inv) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) inv[4 * i + j] = T[4 * j + i];
  for (int i = 0; i < 3; ++i) inv[4 * i + 3] = -((inv[4 * i] * T[3] + inv[4 * i + 1] * T[7]) + inv[4 * i + 2] * T[11]);
}

int pslam_scene_clip_dev(pslam_ctx* ctx, long long n, const float* d_xyz, const uint32_t* d_desc, const pslam_clip_cfg* cfg,
                         float* d_out_xyz, float* d_out_uvz, int* d_out_index, uint32_t* d_out_desc, long long* n_out,
                         int reps, double* ms_per_call) {
  if (!ctx || !cfg || n < 0 || (n > 0 && !d_xyz) || (d_out_desc && !d_desc) || reps < 1) return PSLAM_E_INVALID;
  PSLAM_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  const size_t state_bytes = pslam_k_scene_clip_state_bytes(n);
  if (PSLAM_SOLVER_SCRATCH_OFFSET + state_bytes > ctx->scratch_bytes)
    return pslam_set_error(ctx, PSLAM_E_CAPACITY, "scene_clip: too many points for the scratch buffer", cudaSuccess);
  unsigned long long* d_state = reinterpret_cast<unsigned long long*>(ctx->d_scratch + PSLAM_SOLVER_SCRATCH_OFFSET);
  float inv[12];
  clip_invert(cfg->camera_in_map, inv);
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (ms_per_call) {
    PSLAM_CUDA_TRY(ctx, cudaEventCreate(&e0));
    PSLAM_CUDA_TRY(ctx, cudaEventCreate(&e1));
    cudaEventRecord(e0, ctx->stream);
  }
  long long* d_n = nullptr;
  int rc = PSLAM_OK;
  for (int r = 0; r < reps && rc == PSLAM_OK; ++r)
    rc = pslam_k_scene_clip(ctx, n, d_xyz, d_desc, inv, cfg->apply_sensor_in_robot ? cfg->sensor_in_robot : nullptr, cfg->K,
                            cfg->canvas_rows, cfg->canvas_cols, cfg->range_min, cfg->range_max, d_state, d_out_xyz, d_out_uvz,
                            d_out_index, d_out_desc, &d_n);
  if (ms_per_call) {
    if (rc == PSLAM_OK) {
      cudaEventRecord(e1, ctx->stream);
      cudaEventSynchronize(e1);
      float ms = 0;
      cudaEventElapsedTime(&ms, e0, e1);
      *ms_per_call = (double) ms / reps;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
  }
  if (rc) return rc;
  long long* h_n = reinterpret_cast<long long*>(ctx->h_pinned);
  PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(h_n, d_n, sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
  PSLAM_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  if (n_out) *n_out = *h_n;
  return PSLAM_OK;
}

int pslam_scene_clip(pslam_ctx* ctx, int n, const float* xyz, const uint8_t* desc, const pslam_clip_cfg* cfg, int capacity,
                     float* out_xyz, float* out_uvz, int* out_index, uint8_t* out_desc) {
  if (!ctx || !cfg || n < 0 || capacity < 0 || (n > 0 && !xyz) || (out_desc && !desc)) return PSLAM_E_INVALID;
  if (n == 0) return 0;
  PSLAM_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  // carve: [state][xyz in][desc in][xyz out][uvz out][index out][desc out], above the finder's cache partition
  auto al = [](size_t b) { return (b + 255) & ~(size_t) 255; };
  const size_t b_state = al(pslam_k_scene_clip_state_bytes(n)), b_xyz = al((size_t) n * 12), b_desc = desc ? al((size_t) n * 32) : 0,
               b_idx = al((size_t) n * 4);
  const size_t need = b_state + 3 * b_xyz + 2 * b_desc + b_idx;
  // maps that do not fit the context scratch (64 MB) get a temporary device block for this call; a caller that clips
  // every frame keeps the map in HBM and uses pslam_scene_clip_dev
  uint8_t* tmp = nullptr;
  uint8_t* p = ctx->d_scratch + PSLAM_SOLVER_SCRATCH_OFFSET + b_state;
  if (PSLAM_SOLVER_SCRATCH_OFFSET + need > ctx->scratch_bytes) {
    PSLAM_CUDA_TRY(ctx, cudaMalloc(&tmp, need));
    p = tmp;
  }
  float* d_xyz = reinterpret_cast<float*>(p);
  p += b_xyz;
  uint32_t* d_desc = desc ? reinterpret_cast<uint32_t*>(p) : nullptr;
  p += b_desc;
  float* d_oxyz = reinterpret_cast<float*>(p);
  p += b_xyz;
  float* d_ouvz = reinterpret_cast<float*>(p);
  p += b_xyz;
  int* d_oidx = reinterpret_cast<int*>(p);
  p += b_idx;
  uint32_t* d_odesc = desc ? reinterpret_cast<uint32_t*>(p) : nullptr;
  long long kept = 0;
  auto run = [&]() -> int {
    PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(d_xyz, xyz, (size_t) n * 12, cudaMemcpyHostToDevice, ctx->stream));
    if (desc) PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(d_desc, desc, (size_t) n * 32, cudaMemcpyHostToDevice, ctx->stream));
    const int rc = pslam_scene_clip_dev(ctx, n, d_xyz, d_desc, cfg, d_oxyz, d_ouvz, d_oidx, out_desc ? d_odesc : nullptr, &kept, 1, nullptr);
    if (rc) return rc;
    if (kept > capacity) return pslam_set_error(ctx, PSLAM_E_CAPACITY, "scene_clip: more survivors than the output capacity", cudaSuccess);
    if (kept > 0) {
      if (out_xyz) PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(out_xyz, d_oxyz, (size_t) kept * 12, cudaMemcpyDeviceToHost, ctx->stream));
      if (out_uvz) PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(out_uvz, d_ouvz, (size_t) kept * 12, cudaMemcpyDeviceToHost, ctx->stream));
      if (out_index) PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(out_index, d_oidx, (size_t) kept * 4, cudaMemcpyDeviceToHost, ctx->stream));
      if (out_desc) PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(out_desc, d_odesc, (size_t) kept * 32, cudaMemcpyDeviceToHost, ctx->stream));
    }
    PSLAM_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return PSLAM_OK;
  };
  const int rc = run();
  if (tmp) cudaFree(tmp);
  if (rc) return rc;
  return (int) kept;
}

// ---- N3: per-landmark EKF update -----------------------------------------------------------------
int pslam_landmarks_ekf_update_dev(pslam_ctx* ctx, long long n, float* d_state_world, float* d_covariance,
                                   const float* d_measurements, const pslam_ekf_cfg* cfg, float* d_coords_in_local_map,
                                   uint8_t* d_inlier, int* n_inliers, int reps, double* ms_per_call) {
  if (!ctx || !cfg || n < 0 || reps < 1 || cfg->kind < 0 || cfg->kind > 2 ||
      (n > 0 && (!d_state_world || !d_covariance || !d_measurements || !d_coords_in_local_map || !d_inlier)))
    return PSLAM_E_INVALID;
  PSLAM_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  int* d_cnt = reinterpret_cast<int*>(ctx->d_scratch + PSLAM_SOLVER_SCRATCH_OFFSET);
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (ms_per_call) {
    PSLAM_CUDA_TRY(ctx, cudaEventCreate(&e0));
    PSLAM_CUDA_TRY(ctx, cudaEventCreate(&e1));
    cudaEventRecord(e0, ctx->stream);
  }
  int rc = PSLAM_OK;
  for (int r = 0; r < reps && rc == PSLAM_OK; ++r)
    rc = pslam_k_landmarks_ekf(ctx, cfg, n, d_state_world, d_covariance, d_measurements, d_coords_in_local_map, d_inlier, d_cnt);
  if (ms_per_call) {
    if (rc == PSLAM_OK) {
      cudaEventRecord(e1, ctx->stream);
      cudaEventSynchronize(e1);
      float ms = 0;
      cudaEventElapsedTime(&ms, e0, e1);
      *ms_per_call = (double) ms / reps;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
  }
  if (rc) return rc;
  int* h = reinterpret_cast<int*>(ctx->h_pinned);
  PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(h, d_cnt, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  PSLAM_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  if (n_inliers) *n_inliers = *h;
  return PSLAM_OK;
}

int pslam_landmarks_ekf_update(pslam_ctx* ctx, int n, float* state_world, float* covariance, const float* measurements,
                               const pslam_ekf_cfg* cfg, float* coords_in_local_map, uint8_t* inlier) {
  if (!ctx || !cfg || n < 0 || cfg->kind < 0 || cfg->kind > 2 ||
      (n > 0 && (!state_world || !covariance || !measurements || !coords_in_local_map || !inlier)))
    return PSLAM_E_INVALID;
  if (n == 0) return 0;
  PSLAM_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  const int E = cfg->kind == 0 ? 2 : (cfg->kind == 1 ? 3 : 4);
  auto al = [](size_t b) { return (b + 255) & ~(size_t) 255; };
  const size_t b_st = al((size_t) n * 12), b_cov = al((size_t) n * 36), b_ms = al((size_t) n * 4 * E), b_in = al((size_t) n);
  if (PSLAM_SOLVER_SCRATCH_OFFSET + 256 + 2 * b_st + b_cov + b_ms + b_in > ctx->scratch_bytes)
    return pslam_set_error(ctx, PSLAM_E_CAPACITY, "landmarks_ekf: too many landmarks for the scratch buffer (use the _dev variant)", cudaSuccess);
  uint8_t* p = ctx->d_scratch + PSLAM_SOLVER_SCRATCH_OFFSET + 256;  // [0, 256): the inlier counter of the _dev call
  float* d_st = reinterpret_cast<float*>(p);
  p += b_st;
  float* d_cov = reinterpret_cast<float*>(p);
  p += b_cov;
  float* d_ms = reinterpret_cast<float*>(p);
  p += b_ms;
  float* d_loc = reinterpret_cast<float*>(p);
  p += b_st;
  uint8_t* d_in = p;
  PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(d_st, state_world, (size_t) n * 12, cudaMemcpyHostToDevice, ctx->stream));
  PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(d_cov, covariance, (size_t) n * 36, cudaMemcpyHostToDevice, ctx->stream));
  PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(d_ms, measurements, (size_t) n * 4 * E, cudaMemcpyHostToDevice, ctx->stream));
  PSLAM_CUDA_TRY(ctx, cudaMemsetAsync(d_loc, 0, (size_t) n * 12, ctx->stream));
  int n_inliers = 0;
  const int rc = pslam_landmarks_ekf_update_dev(ctx, n, d_st, d_cov, d_ms, cfg, d_loc, d_in, &n_inliers, 1, nullptr);
  if (rc) return rc;
  PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(state_world, d_st, (size_t) n * 12, cudaMemcpyDeviceToHost, ctx->stream));
  PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(covariance, d_cov, (size_t) n * 36, cudaMemcpyDeviceToHost, ctx->stream));
  PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(coords_in_local_map, d_loc, (size_t) n * 12, cudaMemcpyDeviceToHost, ctx->stream));
  PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(inlier, d_in, (size_t) n, cudaMemcpyDeviceToHost, ctx->stream));
  PSLAM_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return n_inliers;
}

int pslam_landmarks_weighted_mean_update(pslam_ctx* ctx, int n, float* state_world, const int* number_of_optimizations,
                                         const float* landmark_in_sensor, const float* sensor_in_world12,
                                         const float* sensor_in_local_map12, float maximum_distance_geometry_meters_squared,
                                         float* coords_in_local_map, uint8_t* inlier) {
  if (!ctx || n < 0 || !sensor_in_world12 || !sensor_in_local_map12 ||
      (n > 0 && (!state_world || !number_of_optimizations || !landmark_in_sensor || !coords_in_local_map || !inlier)))
    return PSLAM_E_INVALID;
  if (n == 0) return 0;
  PSLAM_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  auto al = [](size_t b) { return (b + 255) & ~(size_t) 255; };
  const size_t b3 = al((size_t) n * 12), b1 = al((size_t) n * 4), bi = al((size_t) n);
  if (PSLAM_SOLVER_SCRATCH_OFFSET + 256 + 3 * b3 + b1 + bi > ctx->scratch_bytes)
    return pslam_set_error(ctx, PSLAM_E_CAPACITY, "landmarks_weighted_mean: too many landmarks for the scratch buffer", cudaSuccess);
  int* d_cnt = reinterpret_cast<int*>(ctx->d_scratch + PSLAM_SOLVER_SCRATCH_OFFSET);
  uint8_t* p = ctx->d_scratch + PSLAM_SOLVER_SCRATCH_OFFSET + 256;
  float* d_st = reinterpret_cast<float*>(p);
  p += b3;
  float* d_ls = reinterpret_cast<float*>(p);
  p += b3;
  float* d_loc = reinterpret_cast<float*>(p);
  p += b3;
  int* d_no = reinterpret_cast<int*>(p);
  p += b1;
  uint8_t* d_in = p;
  PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(d_st, state_world, (size_t) n * 12, cudaMemcpyHostToDevice, ctx->stream));
  PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(d_ls, landmark_in_sensor, (size_t) n * 12, cudaMemcpyHostToDevice, ctx->stream));
  PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(d_no, number_of_optimizations, (size_t) n * 4, cudaMemcpyHostToDevice, ctx->stream));
  PSLAM_CUDA_TRY(ctx, cudaMemsetAsync(d_loc, 0, (size_t) n * 12, ctx->stream));
  const int rc = pslam_k_landmarks_weighted_mean(ctx, sensor_in_world12, sensor_in_local_map12, maximum_distance_geometry_meters_squared, n,
                                                 d_st, d_no, d_ls, d_loc, d_in, d_cnt);
  if (rc) return rc;
  int* h = reinterpret_cast<int*>(ctx->h_pinned);
  PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(h, d_cnt, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(state_world, d_st, (size_t) n * 12, cudaMemcpyDeviceToHost, ctx->stream));
  PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(coords_in_local_map, d_loc, (size_t) n * 12, cudaMemcpyDeviceToHost, ctx->stream));
  PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(inlier, d_in, (size_t) n, cudaMemcpyDeviceToHost, ctx->stream));
  PSLAM_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return *h;
}

// ---- N3: MergerProjective_::compute binning ------------------------------------------------------------------
static int merger_cfg_ok(const pslam_merger_cfg* cfg, int dim) {
  if (!cfg || (dim != 3 && dim != 4) || cfg->number_of_row_bins < 1 || cfg->number_of_col_bins < 1 || cfg->canvas_rows < 1 ||
      cfg->canvas_cols < 1 || cfg->kind < PSLAM_MERGER_BASE || cfg->kind > PSLAM_MERGER_DEPTH ||
      (cfg->kind == PSLAM_MERGER_STEREO && dim != 4))
    return 0;
  // "row / col bin width must be at least 1 pixel" (merger_projective_impl.cpp:36-47)
  if ((float) cfg->canvas_rows / (float) cfg->number_of_row_bins < 1 || (float) cfg->canvas_cols / (float) cfg->number_of_col_bins < 1)
    return 0;
  return 1;
}

int pslam_merger_occupancy_words(const pslam_merger_cfg* cfg) {
  if (!cfg || cfg->number_of_row_bins < 1 || cfg->number_of_col_bins < 1) return PSLAM_E_INVALID;
  const long long bins = (long long) (cfg->number_of_row_bins + 1) * (cfg->number_of_col_bins + 1);
  return (int) ((bins + 31) / 32);
}

int pslam_merger_select_updates(pslam_ctx* ctx, const float* measurements, int dim, int n_meas, const int* corr_moving,
                                const float* corr_response, int n_corr, const pslam_merger_cfg* cfg, uint8_t* selected,
                                uint32_t* occupied_bins) {
  if (!ctx || !merger_cfg_ok(cfg, dim) || n_meas < 0 || n_corr < 0 || !occupied_bins || (n_meas > 0 && !measurements) ||
      (n_corr > 0 && (!corr_moving || !corr_response || !selected)))
    return PSLAM_E_INVALID;
  const int n_words = pslam_merger_occupancy_words(cfg);
  if ((size_t) n_words * 32 * 12 > 200 * 1024)
    return pslam_set_error(ctx, PSLAM_E_CAPACITY, "merger: bin grid does not fit shared memory", cudaSuccess);
  if (n_corr == 0) {
    for (int w = 0; w < n_words; ++w) occupied_bins[w] = 0;
    return 0;
  }
  PSLAM_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  auto al = [](size_t b) { return (b + 255) & ~(size_t) 255; };
  const size_t bm = al((size_t) n_meas * dim * 4), bc = al((size_t) n_corr * 4), bs = al((size_t) n_corr), bw = al((size_t) n_words * 4);
  if (PSLAM_SOLVER_SCRATCH_OFFSET + 256 + bm + 2 * bc + bs + bw > ctx->scratch_bytes)
    return pslam_set_error(ctx, PSLAM_E_CAPACITY, "merger: too many measurements for the scratch buffer", cudaSuccess);
  int* d_res = reinterpret_cast<int*>(ctx->d_scratch + PSLAM_SOLVER_SCRATCH_OFFSET);
  uint8_t* p = ctx->d_scratch + PSLAM_SOLVER_SCRATCH_OFFSET + 256;
  float* d_meas = reinterpret_cast<float*>(p);
  p += bm;
  int* d_mv = reinterpret_cast<int*>(p);
  p += bc;
  float* d_rs = reinterpret_cast<float*>(p);
  p += bc;
  uint8_t* d_sel = p;
  p += bs;
  unsigned* d_occ = reinterpret_cast<unsigned*>(p);
  if (n_meas > 0)
    PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(d_meas, measurements, (size_t) n_meas * dim * 4, cudaMemcpyHostToDevice, ctx->stream));
  PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(d_mv, corr_moving, (size_t) n_corr * 4, cudaMemcpyHostToDevice, ctx->stream));
  PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(d_rs, corr_response, (size_t) n_corr * 4, cudaMemcpyHostToDevice, ctx->stream));
  const int rc = pslam_k_merger_select_updates(ctx, cfg, d_meas, dim, n_meas, d_mv, d_rs, n_corr, d_sel, d_occ, n_words, d_res);
  if (rc) return rc;
  int* h = reinterpret_cast<int*>(ctx->h_pinned);
  PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(h, d_res, 2 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(selected, d_sel, (size_t) n_corr, cudaMemcpyDeviceToHost, ctx->stream));
  PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(occupied_bins, d_occ, (size_t) n_words * 4, cudaMemcpyDeviceToHost, ctx->stream));
  PSLAM_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  if (h[1]) return pslam_set_error(ctx, PSLAM_E_INVALID, "merger: measurement outside the bin grid / correspondence index out of range", cudaSuccess);
  return h[0];
}

int pslam_merger_select_additions(pslam_ctx* ctx, const float* measurements, int dim, int n_meas, const pslam_merger_cfg* cfg,
                                  const uint32_t* occupied_bins, int* winners) {
  if (!ctx || !merger_cfg_ok(cfg, dim) || n_meas < 0 || (n_meas > 0 && (!measurements || !winners))) return PSLAM_E_INVALID;
  if (n_meas == 0) return 0;
  const int n_words = pslam_merger_occupancy_words(cfg);
  if ((size_t) n_words * 32 * 12 > 200 * 1024)
    return pslam_set_error(ctx, PSLAM_E_CAPACITY, "merger: bin grid does not fit shared memory", cudaSuccess);
  PSLAM_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  auto al = [](size_t b) { return (b + 255) & ~(size_t) 255; };
  const size_t bm = al((size_t) n_meas * dim * 4), bi = al((size_t) n_meas * 4), bw = al((size_t) n_words * 4);
  if (PSLAM_SOLVER_SCRATCH_OFFSET + 256 + bm + bi + bw > ctx->scratch_bytes)
    return pslam_set_error(ctx, PSLAM_E_CAPACITY, "merger: too many measurements for the scratch buffer", cudaSuccess);
  int* d_res = reinterpret_cast<int*>(ctx->d_scratch + PSLAM_SOLVER_SCRATCH_OFFSET);
  uint8_t* p = ctx->d_scratch + PSLAM_SOLVER_SCRATCH_OFFSET + 256;
  float* d_meas = reinterpret_cast<float*>(p);
  p += bm;
  int* d_win = reinterpret_cast<int*>(p);
  p += bi;
  unsigned* d_occ = reinterpret_cast<unsigned*>(p);
  PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(d_meas, measurements, (size_t) n_meas * dim * 4, cudaMemcpyHostToDevice, ctx->stream));
  if (occupied_bins)
    PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(d_occ, occupied_bins, (size_t) n_words * 4, cudaMemcpyHostToDevice, ctx->stream));
  const int rc = pslam_k_merger_select_additions(ctx, cfg, d_meas, dim, n_meas, occupied_bins ? d_occ : nullptr, d_win, d_res);
  if (rc) return rc;
  int* h = reinterpret_cast<int*>(ctx->h_pinned);
  PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(h, d_res, 2 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  PSLAM_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  if (h[1]) return pslam_set_error(ctx, PSLAM_E_INVALID, "merger: measurement outside the bin grid", cudaSuccess);
  if (h[0] > 0) {
    PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(winners, d_win, (size_t) h[0] * 4, cudaMemcpyDeviceToHost, ctx->stream));
    PSLAM_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return h[0];
}

int pslam_merger_plan(pslam_ctx* ctx, const float* measurements, int dim, int n_meas, const int* corr_moving,
                      const float* corr_response, int n_corr, const pslam_merger_cfg* cfg, uint8_t* selected, uint32_t* occupied_bins,
                      int* winners, int* n_winners) {
  if (!ctx || !merger_cfg_ok(cfg, dim) || n_meas < 0 || n_corr < 0 || !occupied_bins || !n_winners ||
      (n_meas > 0 && (!measurements || !winners)) || (n_corr > 0 && (!corr_moving || !corr_response || !selected)))
    return PSLAM_E_INVALID;
  const int n_words = pslam_merger_occupancy_words(cfg);
  auto al = [](size_t b) { return (b + 255) & ~(size_t) 255; };
  // device block: [meas | moving | response] in, [results (4 ints) | selected | blocked bins | winners] out
  const size_t bm = al((size_t) n_meas * dim * 4), bc = al((size_t) n_corr * 4);
  const size_t in_bytes = bm + 2 * bc;
  const size_t bs = al((size_t) n_corr), bw = al((size_t) n_words * 4), bi = al((size_t) n_meas * 4);
  const size_t out_bytes = 256 + bs + bw + bi;
  if (in_bytes + out_bytes > ctx->pinned_bytes || (size_t) n_words * 32 * 12 > 200 * 1024 ||
      PSLAM_SOLVER_SCRATCH_OFFSET + in_bytes + out_bytes > ctx->scratch_bytes) {
    // too large for the one-copy staging block: the two passes one after the other
    const int k = pslam_merger_select_updates(ctx, measurements, dim, n_meas, corr_moving, corr_response, n_corr, cfg, selected, occupied_bins);
    if (k < 0) return k;
    const int a = pslam_merger_select_additions(ctx, measurements, dim, n_meas, cfg, occupied_bins, winners);
    if (a < 0) return a;
    *n_winners = a;
    return k;
  }
  PSLAM_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  uint8_t* d_in = ctx->d_scratch + PSLAM_SOLVER_SCRATCH_OFFSET;
  float* d_meas = reinterpret_cast<float*>(d_in);
  int* d_mv = reinterpret_cast<int*>(d_in + bm);
  float* d_rs = reinterpret_cast<float*>(d_in + bm + bc);
  uint8_t* d_out = d_in + in_bytes;
  int* d_res = reinterpret_cast<int*>(d_out);
  uint8_t* d_sel = d_out + 256;
  unsigned* d_occ = reinterpret_cast<unsigned*>(d_out + 256 + bs);
  int* d_win = reinterpret_cast<int*>(d_out + 256 + bs + bw);
  uint8_t* h_stage = reinterpret_cast<uint8_t*>(ctx->h_pinned);
  if (n_meas > 0) memcpy(h_stage, measurements, (size_t) n_meas * dim * 4);
  if (n_corr > 0) {
    memcpy(h_stage + bm, corr_moving, (size_t) n_corr * 4);
    memcpy(h_stage + bm + bc, corr_response, (size_t) n_corr * 4);
  }
  if (in_bytes > 0) PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(d_in, h_stage, in_bytes, cudaMemcpyHostToDevice, ctx->stream));
  int rc = pslam_k_merger_select_updates(ctx, cfg, d_meas, dim, n_meas, d_mv, d_rs, n_corr, d_sel, d_occ, n_words, d_res);
  if (rc) return rc;
  if ((rc = pslam_k_merger_select_additions(ctx, cfg, d_meas, dim, n_meas, d_occ, d_win, d_res + 2))) return rc;
  uint8_t* h_res = h_stage + in_bytes;
  PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(h_res, d_out, out_bytes, cudaMemcpyDeviceToHost, ctx->stream));
  PSLAM_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  const int* h = reinterpret_cast<const int*>(h_res);
  if (h[1] || h[3]) return pslam_set_error(ctx, PSLAM_E_INVALID, "merger: measurement outside the bin grid / correspondence index out of range", cudaSuccess);
  if (n_corr > 0) memcpy(selected, h_res + 256, (size_t) n_corr);
  memcpy(occupied_bins, h_res + 256 + bs, (size_t) n_words * 4);
  if (h[2] > 0) memcpy(winners, h_res + 256 + bs + bw, (size_t) h[2] * 4);
  *n_winners = h[2];
  return h[0];
}

int pslam_landmarks_smoother_update(pslam_ctx* ctx, int n, float* state_world, int* number_of_optimizations, int n_frames,
                                    const float* frames_sensor_in_world, const int* offsets, const int* hist_frame,
                                    const float* hist_uv, const float* hist_point_in_camera, const pslam_smoother_cfg* cfg,
                                    float* coords_in_local_map, uint8_t* inlier) {
  if (!ctx || !cfg || n < 0 || n_frames < 0 ||
      (n > 0 && (!state_world || !number_of_optimizations || !frames_sensor_in_world || !offsets || !hist_frame || !hist_uv ||
                 !hist_point_in_camera || !coords_in_local_map || !inlier)))
    return PSLAM_E_INVALID;
  if (n == 0) return 0;
  const int total = offsets[n];
  if (offsets[0] != 0 || total < 0) return pslam_set_error(ctx, PSLAM_E_INVALID, "landmarks_smoother: offsets must start at 0", cudaSuccess);
  for (int i = 0; i < n; ++i)
    if (offsets[i + 1] < offsets[i]) return pslam_set_error(ctx, PSLAM_E_INVALID, "landmarks_smoother: offsets must not decrease", cudaSuccess);
  for (int k = 0; k < total; ++k)
    if (hist_frame[k] < 0 || hist_frame[k] >= n_frames) return pslam_set_error(ctx, PSLAM_E_INVALID, "landmarks_smoother: frame index out of range", cudaSuccess);
  PSLAM_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  // frames: world_in_sensor = sensor_in_world^-1 (what the measurements store, :16-19); current frame: world_in_local_map
  std::vector<float> wis((size_t) 12 * n_frames);
  auto invert = [](const float* T, float* inv) {
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) inv[4 * i + j] = T[4 * j + i];
    for (int i = 0; i < 3; ++i) inv[4 * i + 3] = -((inv[4 * i] * T[3] + inv[4 * i + 1] * T[7]) + inv[4 * i + 2] * T[11]);
  };
  for (int f = 0; f < n_frames; ++f) invert(frames_sensor_in_world + 12 * f, wis.data() + 12 * f);
  float cur_inv[12], wl[12];
  invert(cfg->sensor_in_world, cur_inv);
  const float* L = cfg->sensor_in_local_map;
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) wl[4 * i + j] = (L[4 * i] * cur_inv[j] + L[4 * i + 1] * cur_inv[4 + j]) + L[4 * i + 2] * cur_inv[8 + j];
    wl[4 * i + 3] = ((L[4 * i] * cur_inv[3] + L[4 * i + 1] * cur_inv[7]) + L[4 * i + 2] * cur_inv[11]) + L[4 * i + 3];
  }
  auto al = [](size_t b) { return (b + 255) & ~(size_t) 255; };
  const size_t b_fr = al((size_t) n_frames * 48), b_off = al((size_t) (n + 1) * 4), b_hf = al((size_t) total * 4), b_uv = al((size_t) total * 8),
               b_pc = al((size_t) total * 12), b3 = al((size_t) n * 12), b1 = al((size_t) n * 4), bi = al((size_t) n);
  if (PSLAM_SOLVER_SCRATCH_OFFSET + 256 + 2 * b_fr + b_off + b_hf + b_uv + b_pc + 2 * b3 + b1 + bi > ctx->scratch_bytes)
    return pslam_set_error(ctx, PSLAM_E_CAPACITY, "landmarks_smoother: histories exceed the scratch buffer", cudaSuccess);
  int* d_cnt = reinterpret_cast<int*>(ctx->d_scratch + PSLAM_SOLVER_SCRATCH_OFFSET);
  uint8_t* p = ctx->d_scratch + PSLAM_SOLVER_SCRATCH_OFFSET + 256;
  auto take = [&](size_t bytes) {
    uint8_t* r = p;
    p += bytes;
    return r;
  };
  float* d_siw = reinterpret_cast<float*>(take(b_fr));
  float* d_wis = reinterpret_cast<float*>(take(b_fr));
  int* d_off = reinterpret_cast<int*>(take(b_off));
  int* d_hf = reinterpret_cast<int*>(take(b_hf));
  float* d_uv = reinterpret_cast<float*>(take(b_uv));
  float* d_pc = reinterpret_cast<float*>(take(b_pc));
  float* d_st = reinterpret_cast<float*>(take(b3));
  float* d_loc = reinterpret_cast<float*>(take(b3));
  int* d_no = reinterpret_cast<int*>(take(b1));
  uint8_t* d_in = take(bi);
  cudaStream_t s = ctx->stream;
  PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(d_siw, frames_sensor_in_world, (size_t) n_frames * 48, cudaMemcpyHostToDevice, s));
  PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(d_wis, wis.data(), (size_t) n_frames * 48, cudaMemcpyHostToDevice, s));
  PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(d_off, offsets, (size_t) (n + 1) * 4, cudaMemcpyHostToDevice, s));
  if (total > 0) {
    PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(d_hf, hist_frame, (size_t) total * 4, cudaMemcpyHostToDevice, s));
    PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(d_uv, hist_uv, (size_t) total * 8, cudaMemcpyHostToDevice, s));
    PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(d_pc, hist_point_in_camera, (size_t) total * 12, cudaMemcpyHostToDevice, s));
  }
  PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(d_st, state_world, (size_t) n * 12, cudaMemcpyHostToDevice, s));
  PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(d_no, number_of_optimizations, (size_t) n * 4, cudaMemcpyHostToDevice, s));
  PSLAM_CUDA_TRY(ctx, cudaMemsetAsync(d_loc, 0, (size_t) n * 12, s));
  PSLAM_CUDA_TRY(ctx, cudaStreamSynchronize(s));  // `wis` is pageable host memory owned by this call
  const int rc = pslam_k_landmarks_smoother(ctx, cfg, wl, n, d_siw, d_wis, d_off, d_hf, d_uv, d_pc, d_st, d_no, d_loc, d_in, d_cnt);
  if (rc) return rc;
  int* h = reinterpret_cast<int*>(ctx->h_pinned);
  PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(h, d_cnt, sizeof(int), cudaMemcpyDeviceToHost, s));
  PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(state_world, d_st, (size_t) n * 12, cudaMemcpyDeviceToHost, s));
  PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(number_of_optimizations, d_no, (size_t) n * 4, cudaMemcpyDeviceToHost, s));
  PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(coords_in_local_map, d_loc, (size_t) n * 12, cudaMemcpyDeviceToHost, s));
  PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(inlier, d_in, (size_t) n, cudaMemcpyDeviceToHost, s));
  PSLAM_CUDA_TRY(ctx, cudaStreamSynchronize(s));
  return *h;
}

// ---- stage 2b -----------------------------------------------------------------------------------
static int bf_upload(pslam_ctx* ctx, int nf, const uint8_t* df, int nm, const uint8_t* dm,
                     uint32_t** d_f, uint32_t** d_m, size_t* used) {
  const size_t bf = ((size_t) 32 * nf + 255) & ~(size_t) 255, bm = ((size_t) 32 * nm + 255) & ~(size_t) 255;
  if (bf + bm + (16 << 20) > ctx->scratch_bytes)
    return pslam_set_error(ctx, PSLAM_E_CAPACITY, "bruteforce: descriptor sets exceed the scratch buffer", cudaSuccess);
  // descriptors live at the END of the scratch buffer, the kernels carve from the front
  *d_f = reinterpret_cast<uint32_t*>(ctx->d_scratch + ctx->scratch_bytes - bf - bm);
  *d_m = reinterpret_cast<uint32_t*>(ctx->d_scratch + ctx->scratch_bytes - bm);
  if (nf > 0) PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(*d_f, df, (size_t) 32 * nf, cudaMemcpyHostToDevice, ctx->stream));
  if (nm > 0) PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(*d_m, dm, (size_t) 32 * nm, cudaMemcpyHostToDevice, ctx->stream));
  *used = bf + bm;
  return PSLAM_OK;
}

int pslam_bf_best2_dev(pslam_ctx* ctx, int n_fixed, const uint32_t* d_desc_fixed, int n_moving,
                       const uint32_t* d_desc_moving, int32_t* d_best, int32_t* d_second,
                       int32_t* d_best_idx) {
  if (!ctx || n_fixed < 0 || n_moving < 0) return PSLAM_E_INVALID;
  PSLAM_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  return pslam_k_bf_best2(ctx, n_fixed, d_desc_fixed, n_moving, d_desc_moving, d_best, d_second, d_best_idx);
}

int pslam_bf_best2(pslam_ctx* ctx, int n_fixed, const uint8_t* desc_fixed, int n_moving,
                   const uint8_t* desc_moving, int32_t* best, int32_t* second, int32_t* best_idx) {
  if (!ctx || n_fixed < 0 || n_moving < 0) return PSLAM_E_INVALID;
  PSLAM_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  uint32_t *d_f, *d_m;
  size_t used;
  int rc = bf_upload(ctx, n_fixed, desc_fixed, n_moving, desc_moving, &d_f, &d_m, &used);
  if (rc) return rc;
  if (n_fixed == 0) return PSLAM_OK;
  const size_t ob = ((size_t) 4 * n_fixed + 255) & ~(size_t) 255;
  int32_t* d_out = reinterpret_cast<int32_t*>(ctx->d_scratch + ctx->scratch_bytes - used - 3 * ob);
  const size_t saved = ctx->scratch_bytes;
  ctx->scratch_bytes = saved - used - 3 * ob;  // kernels may only carve below the outputs
  rc = pslam_k_bf_best2(ctx, n_fixed, d_f, n_moving, d_m, d_out, d_out + ob / 4, d_out + 2 * ob / 4);
  ctx->scratch_bytes = saved;
  if (rc) return rc;
  PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(best, d_out, 4 * (size_t) n_fixed, cudaMemcpyDeviceToHost, ctx->stream));
  PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(second, d_out + ob / 4, 4 * (size_t) n_fixed, cudaMemcpyDeviceToHost, ctx->stream));
  PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(best_idx, d_out + 2 * ob / 4, 4 * (size_t) n_fixed, cudaMemcpyDeviceToHost, ctx->stream));
  PSLAM_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return PSLAM_OK;
}

int pslam_match_bruteforce(pslam_ctx* ctx, int n_fixed, const uint8_t* desc_fixed, int n_moving,
                           const uint8_t* desc_moving, const pslam_match_cfg* cfg, int capacity,
                           int* fixed_idx, int* moving_idx, float* distance) {
  if (!ctx || !cfg || n_fixed < 0 || n_moving < 0) return PSLAM_E_INVALID;
  PSLAM_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  uint32_t *d_f, *d_m;
  size_t used;
  int rc = bf_upload(ctx, n_fixed, desc_fixed, n_moving, desc_moving, &d_f, &d_m, &used);
  if (rc) return rc;
  const size_t saved = ctx->scratch_bytes;
  ctx->scratch_bytes = saved - used;
  rc = pslam_k_bf_match(ctx, n_fixed, d_f, n_moving, d_m, cfg->maximum_descriptor_distance,
                        cfg->maximum_distance_ratio_to_second_best, capacity, fixed_idx, moving_idx, distance);
  ctx->scratch_bytes = saved;
  return rc;
}

// ---- stage 2c -----------------------------------------------------------------------------------
int pslam_projective_set_fixed(pslam_ctx* ctx, int n_fixed, const float* coords, int dim, const uint8_t* desc) {
  if (!ctx || n_fixed < 0 || dim < 2 || dim > 4) return PSLAM_E_INVALID;
  PSLAM_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  return pslam_k_projective_set_fixed(ctx, n_fixed, coords, dim, desc);
}
int pslam_projective_set_moving(pslam_ctx* ctx, int n_moving, const float* xyz, const uint8_t* desc) {
  if (!ctx || n_moving < 0) return PSLAM_E_INVALID;
  PSLAM_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  return pslam_k_projective_set_moving(ctx, n_moving, xyz, desc);
}
int pslam_projective_cache_epochs(const pslam_ctx* ctx, unsigned long long* fixed_epoch, unsigned long long* moving_epoch) {
  if (!ctx) return PSLAM_E_INVALID;
  if (fixed_epoch) *fixed_epoch = ctx->proj_fixed_epoch;
  if (moving_epoch) *moving_epoch = ctx->proj_moving_epoch;
  return PSLAM_OK;
}
int pslam_projective_match(pslam_ctx* ctx, int n_fixed, int n_moving, const float* pose12,
                           const pslam_projective_cfg* cfg, int capacity, int* fixed_idx,
                           int* moving_idx, float* distance, int* n_projected) {
  if (!ctx || !cfg || !pose12) return PSLAM_E_INVALID;
  if (ctx->proj_fixed_epoch == 0 || ctx->proj_moving_epoch == 0)
    return pslam_set_error(ctx, PSLAM_E_INVALID, "projective_match: set_fixed / set_moving have not been called on this context", cudaSuccess);
  PSLAM_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  return pslam_k_projective_match(ctx, n_fixed, n_moving, pose12, cfg, capacity, fixed_idx, moving_idx, distance, n_projected);
}
int pslam_projective_set_moving_weights(pslam_ctx* ctx, int n_moving, const float* information_scale) {
  if (!ctx || n_moving < 0 || (n_moving > 0 && !information_scale)) return PSLAM_E_INVALID;
  PSLAM_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  return pslam_k_projective_set_moving_weights(ctx, n_moving, information_scale);
}
int pslam_projective_match_gn(pslam_ctx* ctx, int n_fixed, int n_moving, const float* pose12, const pslam_projective_cfg* cfg,
                              int capacity, int* fixed_idx, int* moving_idx, float* distance, int* n_projected, pslam_fused_gn* gn) {
  if (!ctx || !pose12 || !cfg || !gn || capacity < 0) return PSLAM_E_INVALID;
  PSLAM_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  gn->iterations_done = 0;
  gn->spd = 1;
  return pslam_k_projective_match(ctx, n_fixed, n_moving, pose12, cfg, capacity, fixed_idx, moving_idx, distance, n_projected, gn);
}
int pslam_projective_align(pslam_ctx* ctx, int n_fixed, int n_moving, const pslam_projective_cfg* cfg, pslam_align* align,
                           int capacity, int* fixed_idx, int* moving_idx, float* distance, pslam_fused_gn* gn) {
  if (!ctx || !cfg || !align || !gn || capacity < 0) return PSLAM_E_INVALID;
  if (ctx->proj_fixed_epoch == 0 || ctx->proj_moving_epoch == 0)
    return pslam_set_error(ctx, PSLAM_E_INVALID, "projective_align: set_fixed / set_moving have not been called on this context", cudaSuccess);
  PSLAM_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  return pslam_k_projective_align(ctx, n_fixed, n_moving, cfg, align, capacity, fixed_idx, moving_idx, distance, gn);
}
int pslam_match_projective(pslam_ctx* ctx, int n_fixed, const float* fixed_coords, int fixed_dim,
                           const uint8_t* desc_fixed, int n_moving, const float* moving_xyz,
                           const uint8_t* desc_moving, const float* pose12,
                           const pslam_projective_cfg* cfg, int capacity, int* fixed_idx,
                           int* moving_idx, float* distance, int* n_projected) {
  int rc;
  if ((rc = pslam_projective_set_fixed(ctx, n_fixed, fixed_coords, fixed_dim, desc_fixed))) return rc;
  if ((rc = pslam_projective_set_moving(ctx, n_moving, moving_xyz, desc_moving))) return rc;
  return pslam_projective_match(ctx, n_fixed, n_moving, pose12, cfg, capacity, fixed_idx, moving_idx, distance, n_projected);
}

// ---- stages 3 + 4 -------------------------------------------------------------------------------
}  // extern "C"
namespace {
template <typename T>
int linearize_host(pslam_ctx* ctx, const pslam_linearize_cfg* cfg, const double* pose12, int n_moving, const T* moving_xyz,
                   int n_fixed, const T* fixed_meas, int fixed_dim, int n_corr, const int* corr_fixed, const int* corr_moving,
                   const T* info_diag, const pslam_pose_prior* prior, uint8_t* factor_status, double* H36, double* b6,
                   double* stats5) {
  if (!ctx || !cfg || !pose12 || n_corr < 0 || fixed_dim < 2 || fixed_dim > 4) return PSLAM_E_INVALID;
  if (cfg->kind < 0 || cfg->kind > 2 || cfg->robustifier < 0 || cfg->robustifier > 2) return PSLAM_E_INVALID;
  PSLAM_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  // validate indices on the host (the reference asserts; we refuse)
  for (int k = 0; k < n_corr; ++k)
    if (corr_fixed[k] < 0 || corr_fixed[k] >= n_fixed || corr_moving[k] < 0 || corr_moving[k] >= n_moving)
      return pslam_set_error(ctx, PSLAM_E_INVALID, "linearize: correspondence index out of range", cudaSuccess);
  uint8_t* p = ctx->d_scratch + PSLAM_SOLVER_SCRATCH_OFFSET;
  auto carve = [&](size_t bytes) {
    uint8_t* r = p;
    p += (bytes + 255) & ~(size_t) 255;
    return r;
  };
  T* d_mv = (T*) carve(sizeof(T) * 3 * (size_t) n_moving);
  T* d_fx = (T*) carve(sizeof(T) * fixed_dim * (size_t) n_fixed);
  T* d_info = (T*) carve(sizeof(T) * 3 * (size_t) n_fixed);
  int* d_cf = (int*) carve(sizeof(int) * (size_t) n_corr);
  int* d_cm = (int*) carve(sizeof(int) * (size_t) n_corr);
  uint8_t* d_status = factor_status ? carve((size_t) n_corr) : nullptr;
  if ((size_t) (p - ctx->d_scratch) + (1 << 20) > ctx->scratch_bytes)
    return pslam_set_error(ctx, PSLAM_E_CAPACITY, "linearize: scratch too small", cudaSuccess);
  if (n_moving) PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(d_mv, moving_xyz, sizeof(T) * 3 * (size_t) n_moving, cudaMemcpyHostToDevice, ctx->stream));
  if (n_fixed) {
    PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(d_fx, fixed_meas, sizeof(T) * fixed_dim * (size_t) n_fixed, cudaMemcpyHostToDevice, ctx->stream));
    PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(d_info, info_diag, sizeof(T) * 3 * (size_t) n_fixed, cudaMemcpyHostToDevice, ctx->stream));
  }
  if (n_corr) {
    PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(d_cf, corr_fixed, sizeof(int) * (size_t) n_corr, cudaMemcpyHostToDevice, ctx->stream));
    PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(d_cm, corr_moving, sizeof(int) * (size_t) n_corr, cudaMemcpyHostToDevice, ctx->stream));
  }
  int rc = pslam_k_linearize_t<T>(ctx, cfg, pose12, d_mv, d_fx, fixed_dim, n_corr, d_cf, d_cm, d_info, prior, d_status, H36, b6, stats5);
  if (rc == PSLAM_OK && factor_status && n_corr) {
    PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(factor_status, d_status, (size_t) n_corr, cudaMemcpyDeviceToHost, ctx->stream));
    PSLAM_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return rc;
}

// bench.py's H,b throughput line: inputs uploaded once into temporary device buffers (a batched synthetic does not fit
// the context scratch), `reps` linearise + reduce passes timed with CUDA events on the context's stream
template <typename T>
int linearize_timed_host(pslam_ctx* ctx, const pslam_linearize_cfg* cfg, const double* pose12, int n_moving, const T* moving_xyz,
                         int n_fixed, const T* fixed_meas, int fixed_dim, int n_corr, const int* corr_fixed, const int* corr_moving,
                         const T* info_diag, int reps, double* ms_per_call) {
  if (!ctx || !cfg || !pose12 || !ms_per_call || n_corr <= 0 || reps <= 0 || fixed_dim < 2 || fixed_dim > 4) return PSLAM_E_INVALID;
  PSLAM_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  T *d_mv = nullptr, *d_fx = nullptr, *d_info = nullptr;
  int *d_cf = nullptr, *d_cm = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  int rc = PSLAM_OK;
  auto cleanup = [&]() {
    cudaFree(d_mv);
    cudaFree(d_fx);
    cudaFree(d_info);
    cudaFree(d_cf);
    cudaFree(d_cm);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
  };
#define TIMED_TRY(call)                                                        \
  do {                                                                         \
    cudaError_t e__ = (call);                                                  \
    if (e__ != cudaSuccess) {                                                  \
      cleanup();                                                               \
      return pslam_set_error(ctx, PSLAM_E_CUDA, #call, e__);                   \
    }                                                                          \
  } while (0)
  TIMED_TRY(cudaMalloc(&d_mv, sizeof(T) * 3 * (size_t) n_moving));
  TIMED_TRY(cudaMalloc(&d_fx, sizeof(T) * fixed_dim * (size_t) n_fixed));
  TIMED_TRY(cudaMalloc(&d_info, sizeof(T) * 3 * (size_t) n_fixed));
  TIMED_TRY(cudaMalloc(&d_cf, sizeof(int) * (size_t) n_corr));
  TIMED_TRY(cudaMalloc(&d_cm, sizeof(int) * (size_t) n_corr));
  TIMED_TRY(cudaMemcpyAsync(d_mv, moving_xyz, sizeof(T) * 3 * (size_t) n_moving, cudaMemcpyHostToDevice, ctx->stream));
  TIMED_TRY(cudaMemcpyAsync(d_fx, fixed_meas, sizeof(T) * fixed_dim * (size_t) n_fixed, cudaMemcpyHostToDevice, ctx->stream));
  TIMED_TRY(cudaMemcpyAsync(d_info, info_diag, sizeof(T) * 3 * (size_t) n_fixed, cudaMemcpyHostToDevice, ctx->stream));
  TIMED_TRY(cudaMemcpyAsync(d_cf, corr_fixed, sizeof(int) * (size_t) n_corr, cudaMemcpyHostToDevice, ctx->stream));
  TIMED_TRY(cudaMemcpyAsync(d_cm, corr_moving, sizeof(int) * (size_t) n_corr, cudaMemcpyHostToDevice, ctx->stream));
  TIMED_TRY(cudaEventCreate(&e0));
  TIMED_TRY(cudaEventCreate(&e1));
  double H[36], b[6], st[5];
  for (int w = 0; w < 2 && rc == PSLAM_OK; ++w)
    rc = pslam_k_linearize_t<T>(ctx, cfg, pose12, d_mv, d_fx, fixed_dim, n_corr, d_cf, d_cm, d_info, nullptr, nullptr, H, b, st);
  if (rc == PSLAM_OK) {
    TIMED_TRY(cudaEventRecord(e0, ctx->stream));
    for (int r = 0; r < reps && rc == PSLAM_OK; ++r)
      rc = pslam_k_linearize_t<T>(ctx, cfg, pose12, d_mv, d_fx, fixed_dim, n_corr, d_cf, d_cm, d_info, nullptr, nullptr, H, b, st);
    TIMED_TRY(cudaEventRecord(e1, ctx->stream));
    TIMED_TRY(cudaEventSynchronize(e1));
    float ms = 0;
    TIMED_TRY(cudaEventElapsedTime(&ms, e0, e1));
    *ms_per_call = (double) ms / reps;
  }
#undef TIMED_TRY
  cleanup();
  return rc;
}

template <typename T>
int gn_iterate_host(pslam_ctx* ctx, const pslam_linearize_cfg* cfg, int n_iterations, double damping, double* pose12, int n_moving,
                    const T* moving_xyz, int n_fixed, const T* fixed_meas, int fixed_dim, int n_corr, const int* corr_fixed,
                    const int* corr_moving, const T* info_diag, const pslam_pose_prior* prior, double* poses12, double* stats4,
                    uint8_t* factor_status, int* iterations_done) {
  if (iterations_done) *iterations_done = 0;
  if (!ctx || !cfg || !pose12 || n_corr < 0 || n_iterations < 0 || fixed_dim < 2 || fixed_dim > 4) return PSLAM_E_INVALID;
  if (cfg->kind < 0 || cfg->kind > 2 || cfg->robustifier < 0 || cfg->robustifier > 2) return PSLAM_E_INVALID;
  if (n_iterations == 0) return PSLAM_OK;
  if (n_iterations > 4096) return pslam_set_error(ctx, PSLAM_E_INVALID, "gn_iterate: more than 4096 iterations per call", cudaSuccess);
  PSLAM_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  for (int k = 0; k < n_corr; ++k)
    if (corr_fixed[k] < 0 || corr_fixed[k] >= n_fixed || corr_moving[k] < 0 || corr_moving[k] >= n_moving)
      return pslam_set_error(ctx, PSLAM_E_INVALID, "gn_iterate: correspondence index out of range", cudaSuccess);
  std::vector<double> out(16 * (size_t) n_iterations);
  int done = 0, spd = 1;
  int rc = pslam_k_gn_iterate_t<T>(ctx, cfg, n_iterations, damping, pose12, n_moving, moving_xyz, n_fixed, fixed_meas, fixed_dim,
                                   n_corr, corr_fixed, corr_moving, info_diag, prior, out.data(), factor_status, &done, &spd);
  if (rc) return rc;
  for (int i = 0; i < done; ++i) {
    if (poses12) memcpy(poses12 + 12 * (size_t) i, out.data() + 16 * (size_t) i, sizeof(double) * 12);
    if (stats4) memcpy(stats4 + 4 * (size_t) i, out.data() + 16 * (size_t) i + 12, sizeof(double) * 4);
  }
  if (iterations_done) *iterations_done = done;
  if (!spd) return pslam_set_error(ctx, PSLAM_E_NOT_SPD, "gn_iterate: H + damping*I is not positive definite", cudaSuccess);
  return PSLAM_OK;
}
}  // namespace
extern "C" {

int pslam_linearize_se3(pslam_ctx* ctx, const pslam_linearize_cfg* cfg, const double* pose12,
                        int n_moving, const double* moving_xyz, int n_fixed, const double* fixed_meas,
                        int fixed_dim, int n_corr, const int* corr_fixed, const int* corr_moving,
                        const double* info_diag, double* H36, double* b6, double* stats4) {
  double st[5];
  const int rc = linearize_host<double>(ctx, cfg, pose12, n_moving, moving_xyz, n_fixed, fixed_meas, fixed_dim, n_corr, corr_fixed,
                                        corr_moving, info_diag, nullptr, nullptr, H36, b6, st);
  if (rc == PSLAM_OK) memcpy(stats4, st, sizeof(double) * 4);
  return rc;
}
int pslam_linearize_se3_f32(pslam_ctx* ctx, const pslam_linearize_cfg* cfg, const double* pose12, int n_moving,
                            const float* moving_xyz, int n_fixed, const float* fixed_meas, int fixed_dim, int n_corr,
                            const int* corr_fixed, const int* corr_moving, const float* info_diag,
                            const pslam_pose_prior* prior, uint8_t* factor_status, double* H36, double* b6, double* stats5) {
  return linearize_host<float>(ctx, cfg, pose12, n_moving, moving_xyz, n_fixed, fixed_meas, fixed_dim, n_corr, corr_fixed,
                               corr_moving, info_diag, prior, factor_status, H36, b6, stats5);
}
int pslam_linearize_se3_timed(pslam_ctx* ctx, const pslam_linearize_cfg* cfg, const double* pose12, int n_moving,
                              const double* moving_xyz, int n_fixed, const double* fixed_meas, int fixed_dim, int n_corr,
                              const int* corr_fixed, const int* corr_moving, const double* info_diag, int reps,
                              double* ms_per_call) {
  return linearize_timed_host<double>(ctx, cfg, pose12, n_moving, moving_xyz, n_fixed, fixed_meas, fixed_dim, n_corr, corr_fixed,
                                      corr_moving, info_diag, reps, ms_per_call);
}
int pslam_linearize_se3_timed_f32(pslam_ctx* ctx, const pslam_linearize_cfg* cfg, const double* pose12, int n_moving,
                                  const float* moving_xyz, int n_fixed, const float* fixed_meas, int fixed_dim, int n_corr,
                                  const int* corr_fixed, const int* corr_moving, const float* info_diag, int reps,
                                  double* ms_per_call) {
  return linearize_timed_host<float>(ctx, cfg, pose12, n_moving, moving_xyz, n_fixed, fixed_meas, fixed_dim, n_corr, corr_fixed,
                                     corr_moving, info_diag, reps, ms_per_call);
}
int pslam_gn_iterate(pslam_ctx* ctx, const pslam_linearize_cfg* cfg, int n_iterations, double damping, double* pose12,
                     int n_moving, const double* moving_xyz, int n_fixed, const double* fixed_meas, int fixed_dim,
                     int n_corr, const int* corr_fixed, const int* corr_moving, const double* info_diag,
                     double* poses12, double* stats4, int* iterations_done) {
  return gn_iterate_host<double>(ctx, cfg, n_iterations, damping, pose12, n_moving, moving_xyz, n_fixed, fixed_meas, fixed_dim,
                                 n_corr, corr_fixed, corr_moving, info_diag, nullptr, poses12, stats4, nullptr, iterations_done);
}
int pslam_gn_iterate_f32(pslam_ctx* ctx, const pslam_linearize_cfg* cfg, int n_iterations, double damping, double* pose12,
                         int n_moving, const float* moving_xyz, int n_fixed, const float* fixed_meas, int fixed_dim,
                         int n_corr, const int* corr_fixed, const int* corr_moving, const float* info_diag,
                         const pslam_pose_prior* prior, double* poses12, double* stats4, uint8_t* factor_status,
                         int* iterations_done) {
  return gn_iterate_host<float>(ctx, cfg, n_iterations, damping, pose12, n_moving, moving_xyz, n_fixed, fixed_meas, fixed_dim,
                                n_corr, corr_fixed, corr_moving, info_diag, prior, poses12, stats4, factor_status, iterations_done);
}

int pslam_gn_step(pslam_ctx* ctx, const double* H36, const double* b6, double damping, double* pose12,
                  double* dx6) {
  if (!ctx || !H36 || !b6 || !pose12) return PSLAM_E_INVALID;
  PSLAM_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  return pslam_k_gn_step(ctx, H36, b6, damping, pose12, dx6);
}

}  // extern "C"
