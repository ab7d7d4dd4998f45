// libstdcxx_sort.h -- sequential re-statement of libstdc++'s std::sort (GCC 4.7 .. 14:
// introsort with median-of-3 to first, unguarded Hoare partition, heap-sort fallback at depth
// 2*floor(log2 n), threshold 16, final insertion sort).
//
// Why: the reference selects and orders keypoints / lattice elements / match candidates with
// an UNSTABLE std::sort whose comparator looks at one field only
//   (.../feature_extractors/intensity_feature_extractor_binned.cpp:188-192,
//    .../correspondence_finders/correspondence_finder_projective_square_impl.cpp:27-29,
//    .../correspondence_finder_descriptor_based_bruteforce_impl.cpp:89-92),
// so WHICH equal-key elements survive a cut and in which order they come out is defined by
// this algorithm's exact sequence of moves.  To be bit-exact with the reference the device has
// to replay it; one thread does so over data staged in shared memory.
// Compiles for host (g++) and device (nvcc); tests/test_sort_emulation.py checks the host build
// against the real std::sort, tests/test_gpu_*.py check the device build against the oracle.
#pragma once

#if defined(__CUDACC__)
#define PSLAM_HD __host__ __device__ __forceinline__
#else
#define PSLAM_HD inline
#endif

namespace pslam_sort {

template <typename T>
PSLAM_HD void swap_(T& a, T& b) {
  T t = a;
  a = b;
  b = t;
}

PSLAM_HD int lg_(int n) {  // std::__lg: floor(log2(n)), n > 0
  int k = 0;
  while (n > 1) {
    n >>= 1;
    ++k;
  }
  return k;
}

template <typename T, typename Comp>
PSLAM_HD void push_heap_(T* first, int hole, int top, T value, Comp comp) {
  int parent = (hole - 1) / 2;
  while (hole > top && comp(first[parent], value)) {
    first[hole] = first[parent];
    hole = parent;
    parent = (hole - 1) / 2;
  }
  first[hole] = value;
}

template <typename T, typename Comp>
PSLAM_HD void adjust_heap_(T* first, int hole, int len, T value, Comp comp) {
  const int top = hole;
  int child = hole;
  while (child < (len - 1) / 2) {
    child = 2 * (child + 1);
    if (comp(first[child], first[child - 1])) --child;
    first[hole] = first[child];
    hole = child;
  }
  if ((len & 1) == 0 && child == (len - 2) / 2) {
    child = 2 * (child + 1);
    first[hole] = first[child - 1];
    hole = child - 1;
  }
  push_heap_(first, hole, top, value, comp);
}

template <typename T, typename Comp>
PSLAM_HD void heap_sort_(T* first, int len, Comp comp) {  // __partial_sort(first, last, last)
  if (len >= 2) {                                          // __make_heap
    int parent = (len - 2) / 2;
    while (true) {
      T value = first[parent];
      adjust_heap_(first, parent, len, value, comp);
      if (parent == 0) break;
      --parent;
    }
  }
  int last = len;  // __sort_heap
  while (last > 1) {
    --last;
    T value = first[last];  // __pop_heap(first, last, last)
    first[last] = first[0];
    adjust_heap_(first, 0, last, value, comp);
  }
}

template <typename T, typename Comp>
PSLAM_HD void move_median_to_first_(T* a, int result, int ia, int ib, int ic, Comp comp) {
  if (comp(a[ia], a[ib])) {
    if (comp(a[ib], a[ic]))
      swap_(a[result], a[ib]);
    else if (comp(a[ia], a[ic]))
      swap_(a[result], a[ic]);
    else
      swap_(a[result], a[ia]);
  } else if (comp(a[ia], a[ic]))
    swap_(a[result], a[ia]);
  else if (comp(a[ib], a[ic]))
    swap_(a[result], a[ic]);
  else
    swap_(a[result], a[ib]);
}

template <typename T, typename Comp>
PSLAM_HD int unguarded_partition_(T* a, int first, int last, int pivot, Comp comp) {
  while (true) {
    while (comp(a[first], a[pivot])) ++first;
    --last;
    while (comp(a[pivot], a[last])) --last;
    if (!(first < last)) return first;
    swap_(a[first], a[last]);
    ++first;
  }
}

template <typename T, typename Comp>
PSLAM_HD void unguarded_linear_insert_(T* a, int last, Comp comp) {
  T val = a[last];
  int next = last - 1;
  while (comp(val, a[next])) {
    a[last] = a[next];
    last = next;
    --next;
  }
  a[last] = val;
}

template <typename T, typename Comp>
PSLAM_HD void insertion_sort_(T* a, int first, int last, Comp comp) {
  if (first == last) return;
  for (int i = first + 1; i != last; ++i) {
    if (comp(a[i], a[first])) {
      T val = a[i];
      for (int j = i; j > first; --j) a[j] = a[j - 1];  // move_backward
      a[first] = val;
    } else {
      unguarded_linear_insert_(a, i, comp);
    }
  }
}

// std::sort(a, a + n, comp), restricted to what decides the first `need` output positions.
// With need >= n this is the complete algorithm.  With need < n, sub-ranges that start at or
// beyond `need` are left unpartitioned: a Hoare partition only permutes inside its own range
// and every element right of a cut is "not less" than every element left of it, so neither the
// introsort recursion nor the final insertion pass can move such an element into [0, need).
// Positions [0, need) therefore hold exactly what the full std::sort would put there
// (this is what the reference keeps: binned.cpp:195-199 takes the first `quota` elements).
template <typename T, typename Comp>
PSLAM_HD void std_sort_prefix(T* a, int n, int need, Comp comp) {
  if (n <= 0) return;
  const int threshold = 16;
  // __introsort_loop with the recursion on the right part made explicit
  int stack_first[64], stack_last[64], stack_depth[64];
  int sp = 0;
  int sorted_end = n;  // elements at >= sorted_end are not guaranteed to be block-sorted
  stack_first[0] = 0;
  stack_last[0] = n;
  stack_depth[0] = 2 * lg_(n);
  sp = 1;
  while (sp > 0) {
    --sp;
    int first = stack_first[sp], last = stack_last[sp], depth = stack_depth[sp];
    while (last - first > threshold) {
      if (depth == 0) {
        heap_sort_(a + first, last - first, comp);
        break;
      }
      --depth;
      const int mid = first + (last - first) / 2;
      move_median_to_first_(a, first, first + 1, mid, last - 1, comp);
      const int cut = unguarded_partition_(a, first + 1, last, first, comp);
      // recursive call __introsort_loop(cut, last, depth) runs BEFORE the loop continues on
      // [first, cut); the two ranges are disjoint so deferring it does not change the result.
      if (cut < need) {
        stack_first[sp] = cut;
        stack_last[sp] = last;
        stack_depth[sp] = depth;
        ++sp;
      } else if (cut < sorted_end) {
        sorted_end = cut;
      }
      last = cut;
    }
  }
  // __final_insertion_sort (stops where the untouched tail begins)
  if (n > threshold) {
    insertion_sort_(a, 0, threshold, comp);
    for (int i = threshold; i < sorted_end; ++i) unguarded_linear_insert_(a, i, comp);
  } else {
    insertion_sort_(a, 0, n, comp);
  }
}

template <typename T, typename Comp>
PSLAM_HD void std_sort(T* a, int n, Comp comp) {
  std_sort_prefix(a, n, n, comp);
}


#if defined(__CUDACC__)
// ---------------------------------------------------------------------------------------------------------
// Warp-parallel replay of the same algorithm (device only).  Bit-exact with std_sort_prefix:
//   * the unguarded Hoare partition is a pure function of the range's ORIGINAL contents: the k-th swap pairs the
//     k-th "left stopper" (ascending index i with !comp(a[i], pivot)) with the k-th "right stopper" (descending
//     index j with !comp(pivot, a[j])) for as long as L_k < R_k; the returned cut is L_0 when nothing is swapped,
//     else min(L_K, R_{K-1}) with K the number of swaps.  Stoppers are found with ballots, 32 positions per step.
//   * the final insertion sort is a STABLE sort of a sequence whose inversions are confined to quicksort leaves of
//     <= 16 elements, so the final position of element i is i - #{j in [i-16, i): comp(a[i], a[j])}
//     + #{j in (i, i+16]: comp(a[j], a[i])} -- evaluated by all threads of the block at once.
// Control flow (introsort loop, pruning to the first `need` outputs, depth limit) is executed uniformly by the warp;
// median-of-3 and the heap-sort fallback stay on lane 0.  tools/../tests: tests/test_gpu_stage1.py compares the
// selected keypoints with the oracle (which calls the real std::sort) on every golden image.
// a: shared or global memory.  rpos: scratch of n indices (uint16 for n <= 65536, else uint32), same address space rules.
// Must be called by all 32 lanes of one warp.
// Returns sorted_end (see std_sort_prefix); the caller finishes with block_final_positions().
template <typename T, typename RP, typename Comp>
__device__ __forceinline__ int warp_partition_(T* a, RP* rpos, int first, int last, int pivot, Comp comp) {
  const unsigned FULLM = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const unsigned lt = (1u << lane) - 1u;
  const T p = a[pivot];
  int nR = 0;
  for (int base = last; base > first; base -= 32) {  // right stoppers, descending
    const int i = base - 1 - lane;
    const bool f = i >= first && !comp(p, a[i]);
    const unsigned bal = __ballot_sync(FULLM, f);
    if (f) rpos[nR + __popc(bal & lt)] = (RP) i;
    nR += __popc(bal);
  }
  __syncwarp();
  int nL = 0, K = 0, lnext = 0x7fffffff;
  for (int base = first; base < last; base += 32) {  // left stoppers, ascending; swap while L_k < R_k
    const int i = base + lane;
    const bool f = i < last && !comp(a[i], p);
    const unsigned bal = __ballot_sync(FULLM, f);
    const int k = nL + __popc(bal & lt);
    int y = -1;
    if (f && k < nR) y = rpos[k];
    const bool do_swap = f && y > i;
    const unsigned sw = __ballot_sync(FULLM, do_swap);
    __syncwarp();
    if (do_swap) {
      const T t = a[y];
      a[y] = a[i];
      a[i] = t;
    }
    __syncwarp();  // the next round may read a swapped position (compute-sanitizer racecheck)
    K += __popc(sw);
    nL += __popc(bal);
    const unsigned stop = bal & ~sw;  // first stopper that is not swapped = L_K
    if (stop) {
      lnext = base + (__ffs(stop) - 1);
      break;
    }
  }
  __syncwarp();
  if (K > 0) {
    const int r = rpos[K - 1];
    return lnext < r ? lnext : r;
  }
  return lnext;
}

template <typename T, typename RP, typename Comp>
__device__ __forceinline__ int warp_std_sort_prefix(T* a, RP* rpos, int n, int need, Comp comp) {
  const int threshold = 16;
  const int lane = threadIdx.x & 31;
  if (n <= threshold) return n;
  // explicit stack of deferred right parts: entry s lives in lane s (depth <= 2 * floor(log2 n) <= 30 entries)
  int st_first = 0, st_last = 0, st_depth = 0;
  int sp = 0;
  int sorted_end = n;
  if (lane == 0) {
    st_first = 0;
    st_last = n;
    st_depth = 2 * lg_(n);
  }
  sp = 1;
  while (sp > 0) {
    --sp;
    int first = __shfl_sync(0xffffffffu, st_first, sp), last = __shfl_sync(0xffffffffu, st_last, sp),
        depth = __shfl_sync(0xffffffffu, st_depth, sp);
    while (last - first > threshold) {
      if (depth == 0) {
        if (lane == 0) heap_sort_(a + first, last - first, comp);
        __syncwarp();
        break;
      }
      --depth;
      const int mid = first + (last - first) / 2;
      if (lane == 0) move_median_to_first_(a, first, first + 1, mid, last - 1, comp);
      __syncwarp();
      const int cut = warp_partition_(a, rpos, first + 1, last, first, comp);
      if (cut < need) {
        if (lane == sp) {
          st_first = cut;
          st_last = last;
          st_depth = depth;
        }
        ++sp;
      } else if (cut < sorted_end) {
        sorted_end = cut;
      }
      last = cut;
    }
  }
  __syncwarp();
  return sorted_end;
}

// The same replay with the partitions of one recursion level spread over the warps of the block (full sort, need == n).
// The introsort recursion is a binary tree whose nodes touch disjoint ranges: [first, cut) and [cut, last) both continue with
// the decremented depth limit, so the order in which the nodes of a level are processed does not matter.  Level by level:
// every warp takes segments of the current list, partitions them exactly as warp_std_sort_prefix does and appends the
// children longer than the insertion-sort threshold to the next list (at most n / 17 segments are alive at a time).
// s_seg: 2 x SEG_CAP x 3 ints, s_cnt: 3 ints (shared memory).  rpos: n entries; a segment uses the slice of its own range.
// Call with all NT threads; follow with __syncthreads() + block_final_positions(a, n, ...).
// need < n: only the first `need` outputs are wanted (std_sort_prefix's pruning: a right part that starts at or beyond `need`
// is dropped and the smallest such cut is the returned sorted_end, kept in s_cnt[2]); need == n: full sort, returns n.
template <int NT, int SEG_CAP, typename T, typename RP, typename Comp>
__device__ __forceinline__ int block_std_sort_prefix(T* a, RP* rpos, int n, int need, Comp comp, int* s_seg, int* s_cnt) {
  const int threshold = 16;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (n <= threshold) return n;
  if (threadIdx.x == 0) {
    s_seg[0] = 0;
    s_seg[1] = n;
    s_seg[2] = 2 * lg_(n);
    s_cnt[0] = 1;
    s_cnt[1] = 0;
    s_cnt[2] = n;
  }
  __syncthreads();
  int cur = 0;
  for (;;) {
    const int n_cur = s_cnt[cur];
    if (n_cur == 0) break;
    const int* seg = s_seg + cur * (3 * SEG_CAP);
    int* nxt = s_seg + (cur ^ 1) * (3 * SEG_CAP);
    for (int sidx = wid; sidx < n_cur; sidx += NT / 32) {
      const int first = seg[3 * sidx], last = seg[3 * sidx + 1];
      int depth = seg[3 * sidx + 2];
      if (depth == 0) {
        if (lane == 0) heap_sort_(a + first, last - first, comp);
        __syncwarp();
        continue;
      }
      --depth;
      const int mid = first + (last - first) / 2;
      if (lane == 0) move_median_to_first_(a, first, first + 1, mid, last - 1, comp);
      __syncwarp();
      const int cut = warp_partition_(a, rpos + first, first + 1, last, first, comp);
      if (lane == 0) {
        if (cut >= need) {
          atomicMin(&s_cnt[2], cut);  // pruned right part: everything from `cut` on is beyond the wanted prefix
        } else if (last - cut > threshold) {
          const int k = atomicAdd(&s_cnt[cur ^ 1], 1);  // k < SEG_CAP: live segments are disjoint and longer than 16
          nxt[3 * k] = cut;
          nxt[3 * k + 1] = last;
          nxt[3 * k + 2] = depth;
        }
        if (cut - first > threshold) {
          const int k = atomicAdd(&s_cnt[cur ^ 1], 1);
          nxt[3 * k] = first;
          nxt[3 * k + 1] = cut;
          nxt[3 * k + 2] = depth;
        }
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) s_cnt[cur] = 0;
    cur ^= 1;
    __syncthreads();
  }
  return s_cnt[2];
}
template <int NT, int SEG_CAP, typename T, typename RP, typename Comp>
__device__ __forceinline__ void block_std_sort_full(T* a, RP* rpos, int n, Comp comp, int* s_seg, int* s_cnt) {
  block_std_sort_prefix<NT, SEG_CAP>(a, rpos, n, n, comp, s_seg, s_cnt);
}

// final insertion sort as a windowed stable rank; thread-parallel over the block.  out[pos] receives a[i] for the
// positions pos < keep (the caller only needs the kept prefix).  Call after a __syncthreads().
template <int NT, typename T, typename Comp>
__device__ __forceinline__ void block_final_positions(const T* a, int sorted_end, int keep, T* out, Comp comp) {
  for (int i = threadIdx.x; i < sorted_end; i += NT) {
    const T v = a[i];
    int pos = i;
    const int lo = i - 16 > 0 ? i - 16 : 0, hi = i + 16 < sorted_end - 1 ? i + 16 : sorted_end - 1;
    for (int j = lo; j < i; ++j) pos -= comp(v, a[j]) ? 1 : 0;
    for (int j = i + 1; j <= hi; ++j) pos += comp(a[j], v) ? 1 : 0;
    if (pos < keep) out[pos] = v;
  }
}
#endif  // __CUDACC__

}  // namespace pslam_sort
