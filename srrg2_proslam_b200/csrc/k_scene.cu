// k_scene.cu -- N2: projective scene clipping over the WHOLE local map (sm_100a).
//
// Reference path replaced:
//   SceneClipperProjective3D::compute  .../mapping/scene_clipper_projective_3d.cpp:9-67
//     = srrg2_core PointProjectorPinhole_::compute(full_scene, points_in_camera, projections, indices) (:52)
//       followed by transformInPlace(sensor_in_robot) when that is not the identity (:60-62)
// The projector arithmetic (fp32, operation order of oracle/pslam_oracle_solver.hpp::project_point) is the one the
// projective finders use (k_projective.cu); here it runs over 10^4 .. 10^6 map points and the survivors are
// compacted IN ORDER (the reference appends them in map order, and `indices` maps local -> global).
//
// Three launches: (A) scene_flags_kernel streams the map once (TMA bulk staging per tile), projects every point and
// leaves one validity bit per point + one count per tile; (S) scene_scan_kernel scans the few thousand tile counts;
// (B) scene_emit_kernel touches the survivors only (re-projection, robot-frame transform, ordered writes, descriptor
// gather).  Algorithmic traffic: 12 B read per map point; per survivor 12 B re-read + 12 + 12 + 4 B written and the
// 32 B descriptor copied -- pass A is HBM bound and carries the roofline.
#include "pslam_internal.cuh"
#include "pslam_kernels.cuh"

namespace {

constexpr int CLIP_THREADS = 256, CLIP_ITEMS = 8, CLIP_TILE = CLIP_THREADS * CLIP_ITEMS, CLIP_WARPS = CLIP_THREADS / 32;

struct ClipParams {
  float R[9], t[3];    // map_in_camera = (robot_in_local_map * sensor_in_robot)^-1
  float K[9];
  float canvas_cols, canvas_rows, range_min, range_max;
  float Rs[9], ts[3];  // sensor_in_robot, applied to the survivors when apply_sensor != 0
  int apply_sensor;
  float P[12];         // K * [R|t], rounded from double: the fused pre-test (clip_classify)
  float t_max, band_h, band_z;  // max |t_i| and the relative widths of the rounding bands of the pre-test
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }

// pinhole projection of one map point, fp32, operation order of oracle/pslam_oracle_solver.hpp::project_point
__device__ __forceinline__ bool clip_project(const ClipParams& pp, float px, float py, float pz, float c[3], float& u, float& v) {
#pragma unroll
  for (int k = 0; k < 3; ++k)
    c[k] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(pp.R[3 * k], px), __fmul_rn(pp.R[3 * k + 1], py)),
                               __fmul_rn(pp.R[3 * k + 2], pz)), pp.t[k]);
  if (c[2] < pp.range_min || c[2] > pp.range_max) return false;
  const float hx = __fadd_rn(__fadd_rn(__fmul_rn(pp.K[0], c[0]), __fmul_rn(pp.K[1], c[1])), __fmul_rn(pp.K[2], c[2]));
  const float hy = __fadd_rn(__fadd_rn(__fmul_rn(pp.K[3], c[0]), __fmul_rn(pp.K[4], c[1])), __fmul_rn(pp.K[5], c[2]));
  const float hz = __fadd_rn(__fadd_rn(__fmul_rn(pp.K[6], c[0]), __fmul_rn(pp.K[7], c[1])), __fmul_rn(pp.K[8], c[2]));
  u = __fdiv_rn(hx, hz);
  v = __fdiv_rn(hy, hz);
  return !(u < 0.0f || u > pp.canvas_cols || v < 0.0f || v > pp.canvas_rows);
}

// Validity of a point WITHOUT the two IEEE divisions: h = (K [R|t]) p with fused multiply-adds and two tests
//   m_h = min(hx, cols hz - hx, hy, rows hz - hy)  against  +- band_h * (|p|_1 + |t|_max)   (image borders)
//   m_z = min(z - range_min, range_max - z, hz)    against  +- band_z * (|p|_1 + |t|_max)   (range, forward-looking)
// Returns 1 (certainly valid), 0 (certainly invalid) or -1 (inside the rounding band of a border: clip_project decides).
// The fused evaluation and the reference's mul / add sequence both lie within a few ulp * (operand magnitudes) of the
// exact value; band_h = 128 * 2^-24 * (max row sum of |K| rows 0-1 + max(cols, rows) * row sum of |K| row 2 + 1) and
// band_z = 128 * 2^-24 * (row sum of |K| row 2 + 1) over-estimate their distance (and the rounding of the quotient) at
// least 8-fold, so outside the bands both agree.  About 1 point in 10^4 falls into a band.
struct ClipFast {  // copied into registers once per thread (the kernel parameter lives in the constant bank)
  float P[12], Rz[3], tz, cols, rows, rmin, rmax, t_max, band_h, band_z;
};
__device__ __forceinline__ int clip_classify(const ClipFast& f, float px, float py, float pz) {
  const float mag = fabsf(px) + fabsf(py) + fabsf(pz) + f.t_max;
  const float hx = fmaf(f.P[0], px, fmaf(f.P[1], py, fmaf(f.P[2], pz, f.P[3])));
  const float hy = fmaf(f.P[4], px, fmaf(f.P[5], py, fmaf(f.P[6], pz, f.P[7])));
  const float hz = fmaf(f.P[8], px, fmaf(f.P[9], py, fmaf(f.P[10], pz, f.P[11])));
  const float z = fmaf(f.Rz[0], px, fmaf(f.Rz[1], py, fmaf(f.Rz[2], pz, f.tz)));
  const float m_h = fminf(fminf(hx, fmaf(f.cols, hz, -hx)), fminf(hy, fmaf(f.rows, hz, -hy)));
  const float m_z = fminf(fminf(z - f.rmin, f.rmax - z), hz);
  const float bh = mag * f.band_h, bz = mag * f.band_z;
  if (m_h > bh && m_z > bz) return 1;
  if (m_z < -bz && fminf(z - f.rmin, f.rmax - z) < -bz) return 0;  // out of range for certain
  if (m_h < -bh && hz > bz) return 0;                                // outside the image for certain
  return -1;
}

// ---- pass A: project every point, one validity bit per point + one survivor count per tile -------------------
// Persistent CTAs (4 per SM) walk the tiles of CLIP_TILE points with a two-stage shared-memory ring: a tile's 24 KB
// of coordinates arrive by ONE cp.async.bulk (TMA; a single thread issues it, the copy engine completes the
// mbarrier) while the previous tile is projected -- a pure stream over the map with no inter-CTA dependency.
// (A single-pass variant with a decoupled look-back was measured first: with ~600 tiles in flight the look-back
// walks ~20 rounds of L2 latency per tile and the kernel stalls at 1.0-1.6 TB/s; see profiles/.)
__global__ void __launch_bounds__(CLIP_THREADS)
scene_flags_kernel(const ClipParams pp, const float* __restrict__ xyz, long long n, unsigned* __restrict__ flags,
                   int* __restrict__ tile_count, int n_tiles, int bulk_ok) {
  extern __shared__ __align__(128) float s_ring[];  // [2][3 * CLIP_TILE]: two 24 KB stages
  __shared__ __align__(8) unsigned long long s_bar[2];
  __shared__ int s_total[2];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;

  auto issue = [&](int tile, int buf) {  // thread 0: bulk copy of a FULL tile into stage `buf`
    const long long base = (long long) tile * CLIP_TILE;
    if (!bulk_ok || tile >= n_tiles || n - base < CLIP_TILE) return;
    const uint32_t bar = smem_u32(&s_bar[buf]), bytes = 3 * CLIP_TILE * 4;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(s_ring + buf * 3 * CLIP_TILE)), "l"(xyz + 3 * base), "r"(bytes), "r"(bar) : "memory");
  };
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_bar[0])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_bar[1])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    s_total[0] = s_total[1] = 0;
    issue(blockIdx.x, 0);
  }
  __syncthreads();
  ClipFast cf;
#pragma unroll
  for (int i = 0; i < 12; ++i) cf.P[i] = pp.P[i];
  cf.Rz[0] = pp.R[6];
  cf.Rz[1] = pp.R[7];
  cf.Rz[2] = pp.R[8];
  cf.tz = pp.t[2];
  cf.cols = pp.canvas_cols;
  cf.rows = pp.canvas_rows;
  cf.rmin = pp.range_min;
  cf.rmax = pp.range_max;
  cf.t_max = pp.t_max;
  cf.band_h = pp.band_h;
  cf.band_z = pp.band_z;
  unsigned phase = 0;  // bit b = parity the next wait on stage b expects
  int buf = 0;
  // persistent CTAs: while tile t is projected, the bulk copy of tile t + gridDim.x is already in flight
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, buf ^= 1) {
    if (threadIdx.x == 0) issue(tile + gridDim.x, buf ^ 1);  // stage buf ^ 1 was released by the barrier that ended the previous round
    const long long base = (long long) tile * CLIP_TILE;
    const int in_tile = (int) (n - base < CLIP_TILE ? n - base : CLIP_TILE);
    float* s_xyz = s_ring + buf * 3 * CLIP_TILE;
    if (bulk_ok && in_tile == CLIP_TILE) {
      asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}" ::"r"(smem_u32(&s_bar[buf])), "r"((phase >> buf) & 1u) : "memory");
      phase ^= 1u << buf;
    } else {  // ragged last tile / map not 16-byte aligned: plain coalesced loads
      for (int k = threadIdx.x; k < 3 * in_tile; k += CLIP_THREADS) s_xyz[k] = __ldg(xyz + 3 * base + k);
      __syncthreads();
    }
    // fast pass: no divisions, no branches; `unsure` marks the few points inside a rounding band
    int mine = 0;
    unsigned vbits = 0, unsure = 0;
    unsigned* fw = flags + (size_t) tile * (CLIP_TILE / 32) + wid;
#pragma unroll
    for (int j = 0; j < CLIP_ITEMS; ++j) {
      const int q = j * CLIP_THREADS + threadIdx.x;  // word stride 3 between lanes: no bank conflicts
      const int verdict = clip_classify(cf, s_xyz[3 * q], s_xyz[3 * q + 1], s_xyz[3 * q + 2]);  // (stale words past a ragged tile: masked below)
      const bool in = q < in_tile;
      if (in && verdict > 0) vbits |= 1u << j;
      if (in && verdict < 0) unsure |= 1u << j;
    }
    if (__any_sync(0xffffffffu, unsure != 0u)) {  // rounding band: the reference's exact operation sequence decides
      for (int j = 0; j < CLIP_ITEMS; ++j) {
        if (!((unsure >> j) & 1u)) continue;
        const int q = j * CLIP_THREADS + threadIdx.x;
        float c[3], u, v;
        if (clip_project(pp, s_xyz[3 * q], s_xyz[3 * q + 1], s_xyz[3 * q + 2], c, u, v)) vbits |= 1u << j;
      }
    }
#pragma unroll
    for (int j = 0; j < CLIP_ITEMS; ++j) {
      const unsigned bal = __ballot_sync(0xffffffffu, (vbits >> j) & 1u);
      if (lane == 0) {
        fw[j * CLIP_WARPS] = bal;  // bit = point base + 32 * word + lane
        mine += __popc(bal);
      }
    }
    if (lane == 0 && mine) atomicAdd(&s_total[buf], mine);
    __syncthreads();  // all reads of stage `buf` done: it may be refilled in the next round
    if (threadIdx.x == 0) {
      tile_count[tile] = s_total[buf];
      s_total[buf] = 0;
    }
  }
}

// ---- pass S: exclusive scan of the tile counts (one CTA; every thread owns a run of consecutive tiles) --------
__global__ void __launch_bounds__(1024)
scene_scan_kernel(const int* __restrict__ tile_count, int n_tiles, long long* __restrict__ tile_offset, long long* __restrict__ n_out) {
  __shared__ long long s_warp[32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int per = (n_tiles + 1023) / 1024, begin = threadIdx.x * per, end = min(begin + per, n_tiles);
  long long v = 0;
  for (int i = begin; i < end; ++i) v += tile_count[i];
  long long incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const long long t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) s_warp[wid] = incl;
  __syncthreads();
  if (wid == 0) {
    long long w = s_warp[lane], wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const long long t = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o) wi += t;
    }
    s_warp[lane] = wi - w;
    if (lane == 31) *n_out = wi;
  }
  __syncthreads();
  long long run = s_warp[wid] + incl - v;
  for (int i = begin; i < end; ++i) {
    tile_offset[i] = run;
    run += tile_count[i];
  }
}

// ---- pass B: survivors only -- re-project (the reference's exact arithmetic), move into the robot frame, write in
//      map order.  64 threads per tile (4 tiles per CTA): the tile's 64 flag words are expanded into an ordered list of
//      survivor positions in shared memory, then one thread per survivor does the work -- at a few percent survivors
//      a loop over all points would run the projection with one or two active lanes per warp.
constexpr int EMIT_TILES = 4, EMIT_GROUP = 64;
__global__ void __launch_bounds__(EMIT_TILES * EMIT_GROUP)
scene_emit_kernel(const ClipParams pp, const float* __restrict__ xyz, const uint4* __restrict__ desc,
                  const unsigned* __restrict__ flags, const int* __restrict__ tile_count, const long long* __restrict__ tile_offset,
                  int n_tiles, float* __restrict__ out_xyz, float* __restrict__ out_uvz, int* __restrict__ out_index,
                  uint4* __restrict__ out_desc) {
  __shared__ unsigned short s_list[EMIT_TILES][CLIP_TILE];
  __shared__ int s_half[EMIT_TILES];
  const int g = threadIdx.x / EMIT_GROUP, t = threadIdx.x % EMIT_GROUP, lane = threadIdx.x & 31;
  const int tile = blockIdx.x * EMIT_TILES + g;
  const int count = tile < n_tiles ? tile_count[tile] : 0;
  // word t of the tile (points 32 t .. 32 t + 31): exclusive offset by a two-warp scan, then expand the set bits
  unsigned w = count ? flags[(size_t) tile * (CLIP_TILE / 32) + t] : 0u;
  const int c = __popc(w);
  int incl = c;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int tt = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += tt;
  }
  if (t == 31) s_half[g] = incl;
  __syncthreads();
  int pos = incl - c + (t >= 32 ? s_half[g] : 0);
  while (w) {
    const int b = __ffs(w) - 1;
    w &= w - 1;
    s_list[g][pos++] = (unsigned short) (32 * t + b);
  }
  __syncthreads();
  const long long prefix = count ? tile_offset[tile] : 0, base = (long long) tile * CLIP_TILE;
  for (int k = t; k < count; k += EMIT_GROUP) {
    const long long o = prefix + k, i = base + s_list[g][k];
    float cc[3], u, v;
    clip_project(pp, __ldg(xyz + 3 * i), __ldg(xyz + 3 * i + 1), __ldg(xyz + 3 * i + 2), cc, u, v);
    float qx = cc[0], qy = cc[1], qz = cc[2];
    if (pp.apply_sensor) {  // transformInPlace(sensor_in_robot), scene_clipper_projective_3d.cpp:60-62
      const float a0 = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(pp.Rs[0], qx), __fmul_rn(pp.Rs[1], qy)), __fmul_rn(pp.Rs[2], qz)), pp.ts[0]);
      const float a1 = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(pp.Rs[3], qx), __fmul_rn(pp.Rs[4], qy)), __fmul_rn(pp.Rs[5], qz)), pp.ts[1]);
      const float a2 = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(pp.Rs[6], qx), __fmul_rn(pp.Rs[7], qy)), __fmul_rn(pp.Rs[8], qz)), pp.ts[2]);
      qx = a0;
      qy = a1;
      qz = a2;
    }
    if (out_xyz) {
      out_xyz[3 * o] = qx;
      out_xyz[3 * o + 1] = qy;
      out_xyz[3 * o + 2] = qz;
    }
    if (out_uvz) {
      out_uvz[3 * o] = u;
      out_uvz[3 * o + 1] = v;
      out_uvz[3 * o + 2] = cc[2];
    }
    if (out_index) out_index[o] = (int) i;
    if (out_desc) {
      out_desc[2 * o] = __ldg(desc + 2 * i);
      out_desc[2 * o + 1] = __ldg(desc + 2 * i + 1);
    }
  }
}

}  // namespace

// d_state: pslam_k_scene_clip_state_bytes(n) bytes of scratch: [n_out][tile_offset x n_tiles][tile_count x n_tiles]
// [flags: 1 bit per point]; all pointers are device pointers.
// map_in_camera12 / sensor_in_robot12: row-major 3x4 [R|t]; sensor_in_robot12 == nullptr => not applied.
int pslam_k_scene_clip(pslam_ctx* ctx, long long n, const float* d_xyz, const uint32_t* d_desc, const float* map_in_camera12,
                       const float* sensor_in_robot12, const float* K9, int rows, int cols, float range_min, float range_max,
                       unsigned long long* d_state, float* d_out_xyz, float* d_out_uvz, int* d_out_index,
                       uint32_t* d_out_desc, long long** d_n_out) {
  const int n_tiles = (int) ((n + CLIP_TILE - 1) / CLIP_TILE);
  ClipParams pp;
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) {
      pp.R[3 * i + j] = map_in_camera12[4 * i + j];
      pp.Rs[3 * i + j] = sensor_in_robot12 ? sensor_in_robot12[4 * i + j] : (i == j ? 1.f : 0.f);
    }
    pp.t[i] = map_in_camera12[4 * i + 3];
    pp.ts[i] = sensor_in_robot12 ? sensor_in_robot12[4 * i + 3] : 0.f;
  }
  for (int i = 0; i < 9; ++i) pp.K[i] = K9[i];
  pp.canvas_cols = (float) cols;
  pp.canvas_rows = (float) rows;
  pp.range_min = range_min;
  pp.range_max = range_max;
  pp.apply_sensor = sensor_in_robot12 != nullptr;
  pp.t_max = fmaxf(fabsf(pp.t[0]), fmaxf(fabsf(pp.t[1]), fabsf(pp.t[2])));
  double k01 = 0, k2 = 0;
  for (int i = 0; i < 3; ++i) {
    const double rs = fabs((double) pp.K[3 * i]) + fabs((double) pp.K[3 * i + 1]) + fabs((double) pp.K[3 * i + 2]);
    if (i < 2) k01 = rs > k01 ? rs : k01; else k2 = rs;
    for (int j = 0; j < 4; ++j) {
      double acc = 0;
      for (int k = 0; k < 3; ++k) acc += (double) pp.K[3 * i + k] * (double) (j < 3 ? pp.R[3 * k + j] : pp.t[k]);
      pp.P[4 * i + j] = (float) acc;
    }
  }
  pp.band_h = (float) (128.0 * 5.9604645e-8 * (k01 + (cols > rows ? cols : rows) * k2 + 1.0));
  pp.band_z = (float) (128.0 * 5.9604645e-8 * (k2 + 1.0));
  long long* d_n = reinterpret_cast<long long*>(d_state);
  long long* d_offset = d_n + 1;
  int* d_count = reinterpret_cast<int*>(d_offset + n_tiles);
  unsigned* d_flags = reinterpret_cast<unsigned*>(d_count + ((n_tiles + 1) & ~1));
  *d_n_out = d_n;
  if (n_tiles == 0) {
    PSLAM_CUDA_TRY(ctx, cudaMemsetAsync(d_n, 0, 8, ctx->stream));
    return PSLAM_OK;
  }
  const size_t ring = (size_t) 2 * 3 * CLIP_TILE * 4;
  static bool attr_set = false;
  if (!attr_set) {
    PSLAM_CUDA_TRY(ctx, cudaFuncSetAttribute(scene_flags_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) ring));
    attr_set = true;
  }
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
  const int resident = 4 * sms;  // 4 CTAs x 48 KB per SM
  scene_flags_kernel<<<n_tiles < resident ? n_tiles : resident, CLIP_THREADS, ring, ctx->stream>>>(
    pp, d_xyz, n, d_flags, d_count, n_tiles, ((uintptr_t) d_xyz & 15u) == 0 ? 1 : 0);
  PSLAM_LAUNCH_CHECK(ctx, "scene_flags_kernel");
  scene_scan_kernel<<<1, 1024, 0, ctx->stream>>>(d_count, n_tiles, d_offset, d_n);
  PSLAM_LAUNCH_CHECK(ctx, "scene_scan_kernel");
  scene_emit_kernel<<<(n_tiles + EMIT_TILES - 1) / EMIT_TILES, EMIT_TILES * EMIT_GROUP, 0, ctx->stream>>>(
    pp, d_xyz, reinterpret_cast<const uint4*>(d_desc), d_flags, d_count, d_offset, n_tiles, d_out_xyz, d_out_uvz, d_out_index,
    reinterpret_cast<uint4*>(d_out_desc));
  PSLAM_LAUNCH_CHECK(ctx, "scene_emit_kernel");
  return PSLAM_OK;
}

size_t pslam_k_scene_clip_state_bytes(long long n) {
  const size_t n_tiles = (size_t) ((n + CLIP_TILE - 1) / CLIP_TILE);
  return 8 + n_tiles * 8 + ((n_tiles + 1) & ~(size_t) 1) * 4 + n_tiles * (CLIP_TILE / 8);
}
