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

}  // namespace pslam_sort
