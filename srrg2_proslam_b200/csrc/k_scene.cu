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
// One pass over HBM: a CTA takes the next tile of CLIP_TILE points (atomic ticket => tiles start in order), stages its
// 24 KB of coordinates in shared memory with ONE cp.async.bulk (TMA, mbarrier completion), projects CLIP_ITEMS points
// per thread (striped: consecutive lanes take consecutive points), counts survivors with warp ballots, and obtains the
// tile's output offset with a warp-wide decoupled look-back over the tile-state words (status | inclusive / aggregate
// count) instead of a second pass.  Per point: 12 B read,
// survivors write 12 B (point in the robot frame) + 12 B (u, v, depth) + 4 B (global index) and copy the 32 B
// descriptor -- the kernel is HBM bound.
#include "pslam_internal.cuh"
#include "pslam_kernels.cuh"

namespace {

constexpr int CLIP_THREADS = 256, CLIP_ITEMS = 8, CLIP_TILE = CLIP_THREADS * CLIP_ITEMS, CLIP_WARPS = CLIP_THREADS / 32;
constexpr unsigned long long ST_AGGREGATE = 1ull << 62, ST_INCLUSIVE = 2ull << 62, ST_MASK = 3ull << 62;

struct ClipParams {
  float R[9], t[3];    // map_in_camera = (robot_in_local_map * sensor_in_robot)^-1
  float K[9];
  float canvas_cols, canvas_rows, range_min, range_max;
  float Rs[9], ts[3];  // sensor_in_robot, applied to the survivors when apply_sensor != 0
  int apply_sensor;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(CLIP_THREADS)
scene_clip_kernel(const ClipParams pp, const float* __restrict__ xyz, const uint4* __restrict__ desc, long long n,
                  unsigned long long* __restrict__ tile_state, unsigned* __restrict__ ticket,
                  float* __restrict__ out_xyz, float* __restrict__ out_uvz, int* __restrict__ out_index,
                  uint4* __restrict__ out_desc, long long* __restrict__ n_out, int n_tiles, int bulk_ok) {
  __shared__ __align__(128) float s_xyz[3 * CLIP_TILE];  // 24 KB: the tile's points, staged by ONE bulk copy (TMA)
  __shared__ __align__(8) unsigned long long s_bar;
  __shared__ unsigned s_tile;
  __shared__ int s_cnt[CLIP_ITEMS * CLIP_WARPS];
  __shared__ long long s_prefix;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    s_tile = atomicAdd(ticket, 1u);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  const unsigned tile = s_tile;
  const long long base = (long long) tile * CLIP_TILE;
  const int in_tile = (int) (n - base < CLIP_TILE ? n - base : CLIP_TILE);
  // ---- stage the tile: full tiles of a 16-byte aligned map with cp.async.bulk (one thread issues 24 KB, the copy
  //      engine completes the mbarrier), the ragged last tile / unaligned maps with plain coalesced loads
  if (bulk_ok && in_tile == CLIP_TILE) {
    if (threadIdx.x == 0) {
      const uint32_t bar = smem_u32(&s_bar), bytes = 3 * CLIP_TILE * 4;
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(smem_u32(s_xyz)), "l"(xyz + 3 * base), "r"(bytes), "r"(bar) : "memory");
    }
    asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}" ::"r"(smem_u32(&s_bar)), "r"(0u) : "memory");
  } else {
    for (int k = threadIdx.x; k < 3 * in_tile; k += CLIP_THREADS) s_xyz[k] = __ldg(xyz + 3 * base + k);
    __syncthreads();
  }

  float cx[CLIP_ITEMS], cy[CLIP_ITEMS], cz[CLIP_ITEMS], u[CLIP_ITEMS], v[CLIP_ITEMS];
  unsigned flags = 0;
#pragma unroll
  for (int j = 0; j < CLIP_ITEMS; ++j) {
    const int q = j * CLIP_THREADS + threadIdx.x;  // word stride 3 between lanes: no bank conflicts
    bool valid = q < in_tile;
    cx[j] = cy[j] = cz[j] = u[j] = v[j] = 0.f;
    if (valid) {
      const float px = s_xyz[3 * q], py = s_xyz[3 * q + 1], pz = s_xyz[3 * q + 2];
      float c[3];
#pragma unroll
      for (int k = 0; k < 3; ++k)
        c[k] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(pp.R[3 * k], px), __fmul_rn(pp.R[3 * k + 1], py)),
                                   __fmul_rn(pp.R[3 * k + 2], pz)), pp.t[k]);
      valid = !(c[2] < pp.range_min || c[2] > pp.range_max);
      cx[j] = c[0];
      cy[j] = c[1];
      cz[j] = c[2];
      if (valid) {
        const float hx = __fadd_rn(__fadd_rn(__fmul_rn(pp.K[0], c[0]), __fmul_rn(pp.K[1], c[1])), __fmul_rn(pp.K[2], c[2]));
        const float hy = __fadd_rn(__fadd_rn(__fmul_rn(pp.K[3], c[0]), __fmul_rn(pp.K[4], c[1])), __fmul_rn(pp.K[5], c[2]));
        const float hz = __fadd_rn(__fadd_rn(__fmul_rn(pp.K[6], c[0]), __fmul_rn(pp.K[7], c[1])), __fmul_rn(pp.K[8], c[2]));
        u[j] = __fdiv_rn(hx, hz);
        v[j] = __fdiv_rn(hy, hz);
        valid = !(u[j] < 0.0f || u[j] > pp.canvas_cols || v[j] < 0.0f || v[j] > pp.canvas_rows);
      }
    }
    const unsigned bal = __ballot_sync(0xffffffffu, valid);
    if (valid) flags |= 1u << j;
    if (lane == 0) s_cnt[j * CLIP_WARPS + wid] = __popc(bal);
  }
  __syncthreads();
  // exclusive scan of the CLIP_ITEMS * CLIP_WARPS group counts (order: item-major, warp-minor = point order)
  if (wid == 0) {
    int a = s_cnt[lane], b = s_cnt[32 + lane];
    int ia = a, ib = b;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int ta = __shfl_up_sync(0xffffffffu, ia, o), tb = __shfl_up_sync(0xffffffffu, ib, o);
      if (lane >= o) {
        ia += ta;
        ib += tb;
      }
    }
    const int sum_a = __shfl_sync(0xffffffffu, ia, 31);
    const int total = sum_a + __shfl_sync(0xffffffffu, ib, 31);
    s_cnt[lane] = ia - a;
    s_cnt[32 + lane] = sum_a + ib - b;
    // ---- decoupled look-back, one warp: publish this tile's aggregate, then walk back 32 predecessors at a time
    //      until one of them carries an inclusive prefix
    long long prefix = 0;
    if (tile > 0) {
      if (lane == 0) atomicExch(tile_state + tile, ST_AGGREGATE | (unsigned long long) total);
      long long look = (long long) tile - 1;
      while (true) {
        const long long idx = look - lane;
        unsigned long long st = ST_INCLUSIVE;  // before tile 0: an inclusive prefix of zero
        if (idx >= 0) {
          do {
            st = *reinterpret_cast<volatile unsigned long long*>(tile_state + idx);
          } while ((st & ST_MASK) == 0);
        }
        const unsigned inc = __ballot_sync(0xffffffffu, (st & ST_MASK) == ST_INCLUSIVE);
        const int first = inc ? __ffs(inc) - 1 : 31;  // nearest predecessor with an inclusive prefix
        long long val = lane <= first ? (long long) (st & ~ST_MASK) : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) val += __shfl_xor_sync(0xffffffffu, val, o);
        prefix += val;
        if (inc) break;
        look -= 32;
      }
    }
    if (lane == 0) {
      __threadfence();
      atomicExch(tile_state + tile, ST_INCLUSIVE | (unsigned long long) (prefix + total));
      s_prefix = prefix;
      if ((int) tile == n_tiles - 1) *n_out = prefix + total;
    }
  }
  __syncthreads();
  const long long prefix = s_prefix;
#pragma unroll
  for (int j = 0; j < CLIP_ITEMS; ++j) {
    const bool valid = (flags >> j) & 1u;
    const unsigned bal = __ballot_sync(0xffffffffu, valid);
    if (!valid) continue;
    const long long o = prefix + s_cnt[j * CLIP_WARPS + wid] + __popc(bal & ((1u << lane) - 1u));
    const long long i = base + j * CLIP_THREADS + threadIdx.x;
    float qx = cx[j], qy = cy[j], qz = cz[j];
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
      out_uvz[3 * o] = u[j];
      out_uvz[3 * o + 1] = v[j];
      out_uvz[3 * o + 2] = cz[j];
    }
    if (out_index) out_index[o] = (int) i;
    if (out_desc) {
      out_desc[2 * o] = __ldg(desc + 2 * i);
      out_desc[2 * o + 1] = __ldg(desc + 2 * i + 1);
    }
  }
}

}  // namespace

// d_state: n_tiles + 2 words of scratch (tile states, ticket, result); all pointers are device pointers.
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
  PSLAM_CUDA_TRY(ctx, cudaMemsetAsync(d_state, 0, ((size_t) n_tiles + 2) * 8, ctx->stream));
  unsigned* d_ticket = reinterpret_cast<unsigned*>(d_state + n_tiles);
  long long* d_n = reinterpret_cast<long long*>(d_state + n_tiles + 1);
  *d_n_out = d_n;
  if (n_tiles == 0) return PSLAM_OK;
  scene_clip_kernel<<<n_tiles, CLIP_THREADS, 0, ctx->stream>>>(pp, d_xyz, reinterpret_cast<const uint4*>(d_desc), n, d_state, d_ticket,
                                                               d_out_xyz, d_out_uvz, d_out_index,
                                                               reinterpret_cast<uint4*>(d_out_desc), d_n, n_tiles,
                                                               ((uintptr_t) d_xyz & 15u) == 0 ? 1 : 0);
  PSLAM_LAUNCH_CHECK(ctx, "scene_clip_kernel");
  return PSLAM_OK;
}

size_t pslam_k_scene_clip_state_bytes(long long n) { return (size_t) ((n + CLIP_TILE - 1) / CLIP_TILE + 2) * 8; }
