// k_bruteforce.cu -- stage 2b (sm_100a): exhaustive Hamming sweep over 256-bit descriptors.
//
// Replaces CorrespondenceFinderDescriptorBasedBruteforce::compute (+ checkLowesRatio,
// _processCorrespondencePool)
//   (.../correspondence_finders/correspondence_finder_descriptor_based_bruteforce_impl.cpp:6-294).
//
// Sweep: one thread owns QPT query (fixed) descriptors in registers and walks a slice of the
// train (moving) set staged in shared memory (all lanes read the same train descriptor -> smem
// broadcast).  Per pair: 8 LOP3(xor) + 8 POPC + adds + best/second update; the POPC pipe
// (16 results/clk/SM) is the roofline, HBM traffic is ~0 (32*(Nq+Nt) bytes for Nq*Nt pairs).
// Query rows x train slices form the grid so that the grid is a multiple of the SM count; a tiny
// merge kernel combines the slices (first index wins ties, like the reference's sequential scan).
//
// Bijective resolve: candidates (d < max_dist) are written in the reference's row-major
// generation order, lane 0 replays std::sort on them (the reference's unstable sort decides the
// output order inside an equal-distance pool), then one CTA walks the distance levels.
#include <float.h>

#include "libstdcxx_sort.h"
#include "pslam_internal.cuh"
#include "pslam_kernels.cuh"

namespace {

constexpr int BF_THREADS = 128;
constexpr int BF_QPT_MAX = 4;                   // queries per thread (template parameter: 2 or 4)
constexpr int BF_TTILE = 256;                   // train descriptors per smem stage (8 KB)
constexpr int BF_MAX_SLICE = 65536;             // the packed best-2 key carries the train index within the slice in 16 bits

struct Best2 {
  int best, second, idx;
};

__device__ __forceinline__ unsigned lop3_xor3(unsigned a, unsigned b, unsigned c) {
  unsigned r;
  asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
  return r;
}
__device__ __forceinline__ unsigned lop3_maj(unsigned a, unsigned b, unsigned c) {
  unsigned r;
  asm("lop3.b32 %0, %1, %2, %3, 0xe8;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
  return r;
}

// 256-bit Hamming distance with a two-level carry-save reduction of the eight XOR words:
//   (x0,x1,x2) -> (s0,c0)   (x3,x4,x5) -> (s1,c1)   (s0,s1,x6) -> (s2,c2)
//   d = popc(s2) + popc(x7) + 2 (popc(c0) + popc(c1) + popc(c2))
// 5 POPC + 14 LOP3 instead of 8 POPC + 8 LOP3.  Measured pipe rates on B200 (profiles/int_peaks.json): POPC 16 / clk / SM
// (XU pipe), LOP3 / IADD3 64 / clk / SM (ALU pipe).  The plain form is XU-bound at 8 / 16 = 0.5 clk per pair (ncu: XU 94 %,
// ALU 49 %, profiles/r04b_bf_sweep_ncu.txt); this form balances the two pipes at ~0.31 clk per pair.  The full Harley-Seal
// tree (4 POPC, 22 LOP3) would be ALU-bound at 0.39 -- worse.
__device__ __forceinline__ int hamming256_csa(const uint4 a0, const uint4 a1, const uint4 b0, const uint4 b1) {
  const unsigned x0 = a0.x ^ b0.x, x1 = a0.y ^ b0.y, x2 = a0.z ^ b0.z, x3 = a0.w ^ b0.w;
  const unsigned x4 = a1.x ^ b1.x, x5 = a1.y ^ b1.y, x6 = a1.z ^ b1.z, x7 = a1.w ^ b1.w;
  const unsigned s0 = lop3_xor3(x0, x1, x2), c0 = lop3_maj(x0, x1, x2);
  const unsigned s1 = lop3_xor3(x3, x4, x5), c1 = lop3_maj(x3, x4, x5);
  const unsigned s2 = lop3_xor3(s0, s1, x6), c2 = lop3_maj(s0, s1, x6);
  return __popc(s2) + __popc(x7) + 2 * (__popc(c0) + __popc(c1) + __popc(c2));
}

// MODE 0: best / second / argmin per (query, slice)
// MODE 2: MODE 0 + the number of candidates (d < max_dist_i) per (query, slice)
// MODE 1: write candidates (d < max_dist_i) at cand_offset[query * n_slices + slice] in m order
// Best-2 bookkeeping (MODE 0 / 2) on packed keys (distance << 16 | train index within the slice): k1 = smallest,
// k2 = second smallest key -- min / max only, and "first index wins ties" (the reference's sequential scan,
// bruteforce_impl.cpp:36-66) is the key order itself.
template <int MODE, int BF_QPT>
__global__ void __launch_bounds__(BF_THREADS)
bf_sweep_kernel(const uint4* __restrict__ q, int nq, const uint4* __restrict__ t, int nt,
                int slice_len, int n_slices, int max_dist_i, int* __restrict__ part_best,
                int* __restrict__ part_second, int* __restrict__ part_idx,
                int* __restrict__ cand_count, const int* __restrict__ cand_offset,
                unsigned long long* __restrict__ cand, int cand_capacity) {
  constexpr int BF_QTILE = BF_THREADS * BF_QPT;
  __shared__ uint4 s_t[2][BF_TTILE * 2];
  const int tid = threadIdx.x;
  const int qtile = blockIdx.x, slice = blockIdx.y;
  const int t_begin = slice * slice_len;
  const int t_end = min(nt, t_begin + slice_len);

  uint4 qa[BF_QPT][2];
  int qi[BF_QPT];
  unsigned k1[BF_QPT], k2[BF_QPT];
  int cnt[BF_QPT];
  int wpos[BF_QPT];
#pragma unroll
  for (int k = 0; k < BF_QPT; ++k) {
    qi[k] = qtile * BF_QTILE + k * BF_THREADS + tid;
    const int src = min(qi[k], nq - 1);
    qa[k][0] = __ldg(q + 2 * (size_t) src);
    qa[k][1] = __ldg(q + 2 * (size_t) src + 1);
    k1[k] = 0xffffffffu;
    k2[k] = 0xffffffffu;
    cnt[k] = 0;
    wpos[k] = 0;
    if (MODE == 1 && qi[k] < nq) wpos[k] = cand_offset[(size_t) qi[k] * n_slices + slice];
  }

  const int n_stages = (t_end - t_begin + BF_TTILE - 1) / BF_TTILE;
  auto load_stage = [&](int stage, int buf) {
    const int base = t_begin + stage * BF_TTILE;
    for (int i = tid; i < BF_TTILE * 2; i += BF_THREADS) {
      const int g = base + (i >> 1);
      uint4 v = make_uint4(0, 0, 0, 0);
      if (g < t_end) v = __ldg(t + 2 * (size_t) g + (i & 1));
      s_t[buf][i] = v;
    }
  };
  if (n_stages > 0) load_stage(0, 0);
  __syncthreads();
  for (int stage = 0; stage < n_stages; ++stage) {
    const int buf = stage & 1;
    if (stage + 1 < n_stages) load_stage(stage + 1, buf ^ 1);
    const int base = t_begin + stage * BF_TTILE;
    const int lim = min(BF_TTILE, t_end - base);
    unsigned jj = (unsigned) (stage * BF_TTILE);  // index within the slice
#pragma unroll 4
    for (int j = 0; j < lim; ++j, ++jj) {
      const uint4 t0 = s_t[buf][2 * j], t1 = s_t[buf][2 * j + 1];
#pragma unroll
      for (int k = 0; k < BF_QPT; ++k) {
        const int d = hamming256_csa(qa[k][0], qa[k][1], t0, t1);
        if (MODE != 1) {
          const unsigned key = ((unsigned) d << 16) | jj;
          k2[k] = min(k2[k], max(k1[k], key));
          k1[k] = min(k1[k], key);
          if (MODE == 2) cnt[k] += (d < max_dist_i) ? 1 : 0;
        } else {
          if (d < max_dist_i && qi[k] < nq) {
            if (wpos[k] < cand_capacity)
              cand[wpos[k]] = ((unsigned long long) d << 48) | ((unsigned long long) qi[k] << 24) |
                              (unsigned long long) (base + j);
            ++wpos[k];
          }
        }
      }
    }
    __syncthreads();
  }
  if (MODE != 1) {
#pragma unroll
    for (int k = 0; k < BF_QPT; ++k) {
      if (qi[k] < nq) {
        const size_t o = (size_t) slice * nq + qi[k];
        part_best[o] = (k1[k] != 0xffffffffu) ? (int) (k1[k] >> 16) : INT_MAX;
        part_second[o] = (k2[k] != 0xffffffffu) ? (int) (k2[k] >> 16) : INT_MAX;
        part_idx[o] = (k1[k] != 0xffffffffu) ? t_begin + (int) (k1[k] & 0xffffu) : -1;
        if (MODE == 2) cand_count[(size_t) qi[k] * n_slices + slice] = cnt[k];
      }
    }
  }
}

__global__ void bf_merge_kernel(int nq, int n_slices, const int* __restrict__ part_best,
                                const int* __restrict__ part_second, const int* __restrict__ part_idx,
                                int* __restrict__ best, int* __restrict__ second, int* __restrict__ idx) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nq) return;
  Best2 b{INT_MAX, INT_MAX, -1};
  for (int s = 0; s < n_slices; ++s) {  // slice order == train index order: first index wins ties
    const size_t o = (size_t) s * nq + i;
    const int pb = part_best[o], ps = part_second[o];
    if (pb < b.best) {
      b.second = min(b.best, ps);
      b.best = pb;
      b.idx = part_idx[o];
    } else {
      b.second = min(b.second, pb);
    }
  }
  best[i] = b.best;
  second[i] = b.second;
  idx[i] = b.idx;
}

// The same merge, fused with the exchange step of the sharded sweep: every row's result is stored straight into the result
// table of EVERY rank (peer HBM mapped through CUDA IPC: plain coalesced stores that travel over NVLink / NVSwitch), so no
// collective follows -- only a flag per source rank (k_sharded.cu).  peers[r] = table base of rank r: [2][3][cap_rows] ints.
__global__ void bf_merge_p2p_kernel(int nq, int n_slices, const int* __restrict__ part_best, const int* __restrict__ part_second,
                                    const int* __restrict__ part_idx, int row_offset, int cap_rows, int parity, int world,
                                    int* const* __restrict__ peers) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nq) return;
  Best2 b{INT_MAX, INT_MAX, -1};
  for (int s = 0; s < n_slices; ++s) {
    const size_t o = (size_t) s * nq + i;
    const int pb = part_best[o], ps = part_second[o];
    if (pb < b.best) {
      b.second = min(b.best, ps);
      b.best = pb;
      b.idx = part_idx[o];
    } else {
      b.second = min(b.second, pb);
    }
  }
  const size_t base = (size_t) parity * 3 * cap_rows + row_offset + i;
  for (int r = 0; r < world; ++r) {
    int* t = peers[r];
    t[base] = b.best;
    t[base + cap_rows] = b.second;
    t[base + 2 * (size_t) cap_rows] = b.idx;
  }
}

// exclusive scan of `n` ints by a single CTA (n up to a few million); writes total to *total
__global__ void __launch_bounds__(1024)
bf_scan_kernel(const int* __restrict__ in, int* __restrict__ out, int n, int* __restrict__ total) {
  __shared__ int s_warp[33];
  int running = 0;
  for (int base = 0; base < n; base += 1024) {
    const int i = base + threadIdx.x;
    const int v = (i < n) ? in[i] : 0;
    int tot;
    const int off = block_exclusive_scan<1024>(v, s_warp, &tot);
    if (i < n) out[i] = running + off;
    running += tot;
  }
  if (threadIdx.x == 0) *total = running;
}

struct CandLess {  // [](a, b) { return a.response < b.response; }   (bruteforce_impl.cpp:89-92)
  __device__ __forceinline__ bool operator()(const unsigned long long& a,
                                             const unsigned long long& b) const {
    return (a >> 48) < (b >> 48);
  }
};

// presence bitmaps of distances per fixed / per moving index (the reference's sorted distance
// lists, :36-66 and :95-97, reduced to what checkLowesRatio reads: the list size and the first
// entry strictly greater than a given distance)
constexpr int BM_WORDS = 9;  // distances 0..256
__global__ void bf_bitmap_kernel(const unsigned long long* __restrict__ cand, int n,
                                 unsigned* __restrict__ bm_f, int* __restrict__ cnt_f,
                                 unsigned* __restrict__ bm_m, int* __restrict__ cnt_m) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned long long c = cand[i];
  const int d = (int) (c >> 48), f = (int) ((c >> 24) & 0xffffffu), m = (int) (c & 0xffffffu);
  atomicOr(bm_f + (size_t) f * BM_WORDS + (d >> 5), 1u << (d & 31));
  atomicOr(bm_m + (size_t) m * BM_WORDS + (d >> 5), 1u << (d & 31));
  atomicAdd(cnt_f + f, 1);
  atomicAdd(cnt_m + m, 1);
}

__device__ __forceinline__ bool lowes_list(const unsigned* bm, int list_size, int d, float ratio) {
  if (list_size == 1) return true;  // :185-188
  // first entry strictly greater than d (:190-196); none -> second == best -> false (:165-166)
  int w = (d + 1) >> 5;
  unsigned bits = bm[w] & (0xffffffffu << ((d + 1) & 31));
  while (bits == 0 && ++w < BM_WORDS) bits = bm[w];
  if (bits == 0) return false;
  const int second = (w << 5) + __ffs(bits) - 1;
  return __fdiv_rn((float) d, (float) second) < ratio;
}

// Candidate sort = the reference's unstable std::sort by response (bruteforce_impl.cpp:89-92), replayed move for move
// (libstdcxx_sort.h): warp 0 runs the introsort partitioning with ballots, then all 1024 threads place every element at its
// final-insertion-sort position.  Candidates that fit shared memory are sorted there, larger sets in place in global memory.
constexpr int BS_THREADS = 1024;
__global__ void __launch_bounds__(BS_THREADS)
bf_sort_kernel(unsigned long long* __restrict__ cand, unsigned long long* __restrict__ sorted, unsigned* __restrict__ g_rpos, int n,
               int smem_cap) {
  extern __shared__ __align__(16) unsigned long long s_c[];
  __shared__ int s_end;
  const int tid = threadIdx.x;
  if (n <= smem_cap) {
    unsigned short* s_rpos = reinterpret_cast<unsigned short*>(s_c + smem_cap);
    for (int i = tid; i < n; i += BS_THREADS) s_c[i] = cand[i];
    __syncthreads();
    if (tid < 32) {
      const int se = pslam_sort::warp_std_sort_prefix(s_c, s_rpos, n, n, CandLess());
      if (tid == 0) s_end = se;
    }
    __syncthreads();
    pslam_sort::block_final_positions<BS_THREADS>(s_c, s_end, n, sorted, CandLess());
  } else {
    if (tid < 32) {
      const int se = pslam_sort::warp_std_sort_prefix(cand, g_rpos, n, n, CandLess());
      if (tid == 0) s_end = se;
    }
    __threadfence_block();
    __syncthreads();
    pslam_sort::block_final_positions<BS_THREADS>(cand, s_end, n, sorted, CandLess());
  }
}

// One CTA walks the sorted candidates level by level (a level = one distance value = one
// "pool" of the reference, :100-152): pool members are the level's candidates whose endpoints
// were not registered by a lower level; a member is accepted iff no other member shares its
// fixed or moving index (:257-268) and Lowe's check holds on both sides (:280-287).
__global__ void __launch_bounds__(1024)
bf_resolve_kernel(const unsigned long long* __restrict__ cand, int n, const unsigned* __restrict__ bm_f,
                  const int* __restrict__ cnt_f, const unsigned* __restrict__ bm_m,
                  const int* __restrict__ cnt_m, unsigned char* __restrict__ reg_f,
                  unsigned char* __restrict__ reg_m, int* __restrict__ pool_f, int* __restrict__ pool_m,
                  float ratio, int* __restrict__ out_f, int* __restrict__ out_m,
                  float* __restrict__ out_d, int* __restrict__ out_n) {
  __shared__ int s_warp[33];
  __shared__ int s_level_end;
  const int tid = threadIdx.x;
  int n_out = 0;
  int start = 0;
  if (n == 1) {  // trivial case of the reference (:80-85): the single candidate is taken unchecked
    if (tid == 0) {
      const unsigned long long c = cand[0];
      out_f[0] = (int) ((c >> 24) & 0xffffffu);
      out_m[0] = (int) (c & 0xffffffu);
      out_d[0] = (float) (int) (c >> 48);
      *out_n = 1;
    }
    return;
  }
  while (start < n) {
    const int d = (int) (cand[start] >> 48);
    // end of this level: first index whose distance differs (sorted ascending)
    if (tid == 0) {
      int lo = start, hi = n;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if ((int) (cand[mid] >> 48) <= d) lo = mid + 1; else hi = mid;
      }
      s_level_end = lo;
    }
    __syncthreads();
    const int end = s_level_end;
    const int n_level_begin = n_out;
    // pass A: pool membership + per-index multiplicity inside the pool
    for (int i = start + tid; i < end; i += 1024) {
      const unsigned long long c = cand[i];
      const int f = (int) ((c >> 24) & 0xffffffu), m = (int) (c & 0xffffffu);
      if (!reg_f[f] && !reg_m[m]) {
        atomicAdd(pool_f + f, 1);
        atomicAdd(pool_m + m, 1);
      }
    }
    __syncthreads();
    // pass B: acceptance, ordered emission
    for (int base = start; base < end; base += 1024) {
      const int i = base + tid;
      int acc = 0, f = 0, m = 0;
      if (i < end) {
        const unsigned long long c = cand[i];
        f = (int) ((c >> 24) & 0xffffffu);
        m = (int) (c & 0xffffffu);
        if (!reg_f[f] && !reg_m[m] && pool_f[f] == 1 && pool_m[m] == 1) {
          acc = lowes_list(bm_f + (size_t) f * BM_WORDS, cnt_f[f], d, ratio) &&
                lowes_list(bm_m + (size_t) m * BM_WORDS, cnt_m[m], d, ratio);
        }
      }
      int total;
      const int off = block_exclusive_scan<1024>(acc, s_warp, &total);
      if (acc) {
        out_f[n_out + off] = f;
        out_m[n_out + off] = m;
        out_d[n_out + off] = (float) d;
      }
      n_out += total;
    }
    __syncthreads();
    // pass C: register accepted, clear pool counters touched by this level
    for (int i = start + tid; i < end; i += 1024) {
      const unsigned long long c = cand[i];
      const int f = (int) ((c >> 24) & 0xffffffu), m = (int) (c & 0xffffffu);
      pool_f[f] = 0;
      pool_m[m] = 0;
    }
    __syncthreads();
    for (int i = n_level_begin + tid; i < n_out; i += 1024) {
      reg_f[out_f[i]] = 1;
      reg_m[out_m[i]] = 1;
    }
    __syncthreads();
    start = end;
  }
  if (tid == 0) *out_n = n_out;
}

int bf_qpt() {  // queries per thread: 4 measured best on B200 (PSLAM_BF_QPT=2 selects the narrower variant for tuning runs)
  static const int v = [] {
    const char* e = getenv("PSLAM_BF_QPT");
    return (e && atoi(e) == 2) ? 2 : 4;
  }();
  return v;
}

int choose_slices(int nq, int nt, int sm_count) {
  const int BF_QTILE = BF_THREADS * bf_qpt();
  const int qtiles = (nq + BF_QTILE - 1) / BF_QTILE;
  int s = (sm_count * 8 + qtiles - 1) / qtiles;
  const int max_s = (nt + 1023) / 1024;
  if (s > max_s) s = max_s;
  if (s < 1) s = 1;
  if (s > 64) s = 64;
  const int min_s = (nt + BF_MAX_SLICE - 1) / BF_MAX_SLICE;  // 16-bit train index within a slice (packed best-2 keys)
  if (s < min_s) s = min_s;
  return s;
}

}  // namespace

// scratch layout helper
static inline size_t align256(size_t x) { return (x + 255) & ~(size_t) 255; }

int pslam_k_bf_best2(pslam_ctx* ctx, int nq, const uint32_t* d_q, int nt, const uint32_t* d_t,
                     int32_t* d_best, int32_t* d_second, int32_t* d_idx) {
  if (nq <= 0) return PSLAM_OK;
  if (nt <= 0) {
    // nothing to compare against: best/second absent
    PSLAM_CUDA_TRY(ctx, cudaMemsetAsync(d_idx, 0xff, sizeof(int) * (size_t) nq, ctx->stream));
    PSLAM_CUDA_TRY(ctx, cudaMemsetAsync(d_best, 0x7f, sizeof(int) * (size_t) nq, ctx->stream));
    PSLAM_CUDA_TRY(ctx, cudaMemsetAsync(d_second, 0x7f, sizeof(int) * (size_t) nq, ctx->stream));
    return PSLAM_OK;
  }
  int sm_count = 148;
  cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, ctx->device);
  const int S = choose_slices(nq, nt, sm_count);
  int slice_len = (nt + S - 1) / S;
  slice_len = (slice_len + BF_TTILE - 1) / BF_TTILE * BF_TTILE;
  const int n_slices = (nt + slice_len - 1) / slice_len;
  const size_t part = align256(sizeof(int) * (size_t) n_slices * nq);
  if (3 * part > ctx->scratch_bytes) return pslam_set_error(ctx, PSLAM_E_CAPACITY, "bf_best2: scratch too small", cudaSuccess);
  int* pb = reinterpret_cast<int*>(ctx->d_scratch);
  int* ps = reinterpret_cast<int*>(ctx->d_scratch + part);
  int* pi = reinterpret_cast<int*>(ctx->d_scratch + 2 * part);
  const int BF_QTILE = BF_THREADS * bf_qpt();
  dim3 grid((nq + BF_QTILE - 1) / BF_QTILE, n_slices);
  auto sweep = bf_qpt() == 4 ? bf_sweep_kernel<0, 4> : bf_sweep_kernel<0, 2>;
  sweep<<<grid, BF_THREADS, 0, ctx->stream>>>(
    reinterpret_cast<const uint4*>(d_q), nq, reinterpret_cast<const uint4*>(d_t), nt, slice_len,
    n_slices, 0, pb, ps, pi, nullptr, nullptr, nullptr, 0);
  PSLAM_LAUNCH_CHECK(ctx, "bf_sweep_kernel<0>");
  bf_merge_kernel<<<(nq + 255) / 256, 256, 0, ctx->stream>>>(nq, n_slices, pb, ps, pi, d_best,
                                                             d_second, d_idx);
  PSLAM_LAUNCH_CHECK(ctx, "bf_merge_kernel");
  return PSLAM_OK;
}

int pslam_k_bf_best2_p2p(pslam_ctx* ctx, int nq, const uint32_t* d_q, int nt, const uint32_t* d_t, int row_offset, int cap_rows,
                         int parity, int world, int* const* d_peers) {
  if (nq <= 0) return PSLAM_OK;
  if (nt <= 0) return pslam_set_error(ctx, PSLAM_E_INVALID, "bf_best2_p2p: empty train set", cudaSuccess);
  int sm_count = 148;
  cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, ctx->device);
  const int S = choose_slices(nq, nt, sm_count);
  int slice_len = (nt + S - 1) / S;
  slice_len = (slice_len + BF_TTILE - 1) / BF_TTILE * BF_TTILE;
  const int n_slices = (nt + slice_len - 1) / slice_len;
  const size_t part = align256(sizeof(int) * (size_t) n_slices * nq);
  if (3 * part > ctx->scratch_bytes) return pslam_set_error(ctx, PSLAM_E_CAPACITY, "bf_best2: scratch too small", cudaSuccess);
  int* pb = reinterpret_cast<int*>(ctx->d_scratch);
  int* ps = reinterpret_cast<int*>(ctx->d_scratch + part);
  int* pi = reinterpret_cast<int*>(ctx->d_scratch + 2 * part);
  const int BF_QTILE = BF_THREADS * bf_qpt();
  dim3 grid((nq + BF_QTILE - 1) / BF_QTILE, n_slices);
  auto sweep = bf_qpt() == 4 ? bf_sweep_kernel<0, 4> : bf_sweep_kernel<0, 2>;
  sweep<<<grid, BF_THREADS, 0, ctx->stream>>>(
    reinterpret_cast<const uint4*>(d_q), nq, reinterpret_cast<const uint4*>(d_t), nt, slice_len,
    n_slices, 0, pb, ps, pi, nullptr, nullptr, nullptr, 0);
  PSLAM_LAUNCH_CHECK(ctx, "bf_sweep_kernel<0>");
  bf_merge_p2p_kernel<<<(nq + 255) / 256, 256, 0, ctx->stream>>>(nq, n_slices, pb, ps, pi, row_offset, cap_rows, parity, world, d_peers);
  PSLAM_LAUNCH_CHECK(ctx, "bf_merge_p2p_kernel");
  return PSLAM_OK;
}

int pslam_k_bf_match(pslam_ctx* ctx, int nf, const uint32_t* d_f, int nm, const uint32_t* d_m,
                     float max_dist, float max_ratio, int capacity, int* h_fixed, int* h_moving,
                     float* h_dist) {
  if (nf <= 0 || nm <= 0) return 0;
  if (nf >= (1 << 24) || nm >= (1 << 24))
    return pslam_set_error(ctx, PSLAM_E_INVALID, "bf_match: more than 2^24 descriptors", cudaSuccess);
  // integer threshold equivalent to (float)d < max_dist
  int max_dist_i = (int) ceilf(max_dist);
  if (max_dist_i > 257) max_dist_i = 257;
  if (max_dist_i < 0) max_dist_i = 0;
  int sm_count = 148;
  cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, ctx->device);
  const int S = choose_slices(nf, nm, sm_count);
  int slice_len = (nm + S - 1) / S;
  slice_len = (slice_len + BF_TTILE - 1) / BF_TTILE * BF_TTILE;
  const int n_slices = (nm + slice_len - 1) / slice_len;

  // scratch carve-up
  uint8_t* p = ctx->d_scratch;
  const size_t part = align256(sizeof(int) * (size_t) n_slices * nf);
  int* pb = (int*) p; p += part;
  int* ps = (int*) p; p += part;
  int* pi = (int*) p; p += part;
  int* cnt = (int*) p; p += part;
  int* offs = (int*) p; p += part;
  int* total = (int*) p; p += 256;
  unsigned* bm_f = (unsigned*) p; p += align256(sizeof(unsigned) * BM_WORDS * (size_t) nf);
  unsigned* bm_m = (unsigned*) p; p += align256(sizeof(unsigned) * BM_WORDS * (size_t) nm);
  int* cnt_f = (int*) p; p += align256(sizeof(int) * (size_t) nf);
  int* cnt_m = (int*) p; p += align256(sizeof(int) * (size_t) nm);
  int* pool_f = (int*) p; p += align256(sizeof(int) * (size_t) nf);
  int* pool_m = (int*) p; p += align256(sizeof(int) * (size_t) nm);
  unsigned char* reg_f = p; p += align256((size_t) nf);
  unsigned char* reg_m = p; p += align256((size_t) nm);
  const int out_cap = nf < nm ? nf : nm;
  int* out_f = (int*) p; p += align256(sizeof(int) * (size_t) out_cap);
  int* out_m = (int*) p; p += align256(sizeof(int) * (size_t) out_cap);
  float* out_d = (float*) p; p += align256(sizeof(float) * (size_t) out_cap);
  uint8_t* zero_end = p;
  if ((size_t) (p - ctx->d_scratch) + 4096 > ctx->scratch_bytes)
    return pslam_set_error(ctx, PSLAM_E_CAPACITY, "bf_match: scratch too small", cudaSuccess);
  // candidates | sorted candidates | right-stopper positions of the partition replay
  const size_t cand_cap_sz = (ctx->scratch_bytes - (size_t) (p - ctx->d_scratch) - 1024) / 20;
  const int cand_cap = cand_cap_sz > 0x7fffffff ? 0x7fffffff : (int) cand_cap_sz;
  unsigned long long* cand = (unsigned long long*) p; p += align256(8 * (size_t) cand_cap);
  unsigned long long* sorted = (unsigned long long*) p; p += align256(8 * (size_t) cand_cap);
  unsigned* g_rpos = (unsigned*) p;

  PSLAM_CUDA_TRY(ctx, cudaMemsetAsync(bm_f, 0, (size_t) (zero_end - (uint8_t*) bm_f), ctx->stream));
  const int BF_QTILE = BF_THREADS * bf_qpt();
  dim3 grid((nf + BF_QTILE - 1) / BF_QTILE, n_slices);
  auto sweep2 = bf_qpt() == 4 ? bf_sweep_kernel<2, 4> : bf_sweep_kernel<2, 2>;
  auto sweep1 = bf_qpt() == 4 ? bf_sweep_kernel<1, 4> : bf_sweep_kernel<1, 2>;
  sweep2<<<grid, BF_THREADS, 0, ctx->stream>>>(
    reinterpret_cast<const uint4*>(d_f), nf, reinterpret_cast<const uint4*>(d_m), nm, slice_len,
    n_slices, max_dist_i, pb, ps, pi, cnt, nullptr, nullptr, 0);
  PSLAM_LAUNCH_CHECK(ctx, "bf_sweep_kernel<2>");
  bf_scan_kernel<<<1, 1024, 0, ctx->stream>>>(cnt, offs, nf * n_slices, total);
  PSLAM_LAUNCH_CHECK(ctx, "bf_scan_kernel");
  int* h = reinterpret_cast<int*>(ctx->h_pinned);
  PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(h, total, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  PSLAM_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  const int n_cand = h[0];
  if (n_cand == 0) return 0;
  if (n_cand > cand_cap)
    return pslam_set_error(ctx, PSLAM_E_CAPACITY, "bf_match: candidate buffer too small", cudaSuccess);
  sweep1<<<grid, BF_THREADS, 0, ctx->stream>>>(
    reinterpret_cast<const uint4*>(d_f), nf, reinterpret_cast<const uint4*>(d_m), nm, slice_len,
    n_slices, max_dist_i, nullptr, nullptr, nullptr, nullptr, offs, cand, cand_cap);
  PSLAM_LAUNCH_CHECK(ctx, "bf_sweep_kernel<1>");
  bf_bitmap_kernel<<<(n_cand + 255) / 256, 256, 0, ctx->stream>>>(cand, n_cand, bm_f, cnt_f, bm_m, cnt_m);
  PSLAM_LAUNCH_CHECK(ctx, "bf_bitmap_kernel");
  const unsigned long long* resolve_in = cand;
  if (n_cand > 1) {
    // 10 B per candidate in shared memory (8 B key + 2 B stopper position): up to 20 k candidates; larger sets in place
    int smem_cap = n_cand <= 4096 ? 4096 : 20 * 1024;
    if (n_cand > smem_cap) smem_cap = 0;
    const size_t smem = (size_t) smem_cap * 10;
    if (smem > 48 * 1024)
      PSLAM_CUDA_TRY(ctx, cudaFuncSetAttribute(bf_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    bf_sort_kernel<<<1, BS_THREADS, smem, ctx->stream>>>(cand, sorted, g_rpos, n_cand, smem_cap);
    PSLAM_LAUNCH_CHECK(ctx, "bf_sort_kernel");
    resolve_in = sorted;
  }
  bf_resolve_kernel<<<1, 1024, 0, ctx->stream>>>(resolve_in, n_cand, bm_f, cnt_f, bm_m, cnt_m, reg_f, reg_m,
                                                 pool_f, pool_m, max_ratio, out_f, out_m, out_d, total);
  PSLAM_LAUNCH_CHECK(ctx, "bf_resolve_kernel");
  PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(h, total, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  PSLAM_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  const int n_out = h[0];
  const int n_copy = n_out < capacity ? n_out : capacity;
  if (n_copy > 0) {
    PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(h_fixed, out_f, sizeof(int) * n_copy, cudaMemcpyDeviceToHost, ctx->stream));
    PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(h_moving, out_m, sizeof(int) * n_copy, cudaMemcpyDeviceToHost, ctx->stream));
    PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(h_dist, out_d, sizeof(float) * n_copy, cudaMemcpyDeviceToHost, ctx->stream));
    PSLAM_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return n_out;
}
