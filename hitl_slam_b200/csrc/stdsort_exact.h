// stdsort_exact.h — libstdc++'s std::sort, restated so that it runs on the device.
//
// Why: the reference builds every KD-tree node with `std::sort` on ONE coordinate and a strict `<`
// comparator (perception_tools/kdtree.cpp:106-139).  std::sort is not stable, so where two points share
// a coordinate the tree shape depends on the exact sequence of swaps libstdc++ performs.  The GPU
// builder (kdtree_gpu.cu) sorts segments with a radix sort — any correct sort gives the same answer when
// all keys differ — and re-sorts the rare segments that contain equal keys with this function, which
// performs libstdc++'s introsort step for step (GCC 13, bits/stl_algo.h + bits/stl_heap.h: __sort,
// __introsort_loop, __unguarded_partition_pivot, __move_median_to_first, __unguarded_partition,
// __final_insertion_sort, __insertion_sort, __unguarded_linear_insert, and the heap-sort fallback
// __partial_sort / __heap_select / __make_heap / __adjust_heap / __push_heap / __pop_heap / __sort_heap).
// The algorithm is a third-party dependency of the reference, absent from /root/reference; it is pinned
// by tests/test_cpu_oracle.py::test_stdsort_restatement_matches_std_sort against this image's libstdc++
// (the same one the host builder and the compiled reference kdtree.cpp use).
//
// Elements are 32-bit ids; `Key` maps an id to the float it is ordered by.  comp(a, b) = key(a) < key(b).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define STDSORT_HD __host__ __device__ __forceinline__
#else
#define STDSORT_HD inline
#endif

namespace hitl {
namespace stdsort {

template <typename Key> struct Sorter {
  uint32_t* v;     // the range being sorted (ids)
  Key key;
  uint32_t heap_sorts = 0;   // times the depth limit was hit (diagnostics)
  STDSORT_HD bool lt(uint32_t a, uint32_t b) const { return key(a) < key(b); }
  STDSORT_HD void swap_at(int64_t a, int64_t b) { const uint32_t t = v[a]; v[a] = v[b]; v[b] = t; }

  // ---- bits/stl_heap.h ----
  STDSORT_HD void push_heap(int64_t first, int64_t hole, int64_t top, uint32_t value) {
    int64_t parent = (hole - 1) / 2;
    while (hole > top && lt(v[first + parent], value)) {
      v[first + hole] = v[first + parent];
      hole = parent;
      parent = (hole - 1) / 2;
    }
    v[first + hole] = value;
  }
  STDSORT_HD void adjust_heap(int64_t first, int64_t hole, int64_t len, uint32_t value) {
    const int64_t top = hole;
    int64_t second = hole;
    while (second < (len - 1) / 2) {
      second = 2 * (second + 1);
      if (lt(v[first + second], v[first + (second - 1)])) second--;
      v[first + hole] = v[first + second];
      hole = second;
    }
    if ((len & 1) == 0 && second == (len - 2) / 2) {
      second = 2 * (second + 1);
      v[first + hole] = v[first + (second - 1)];
      hole = second - 1;
    }
    push_heap(first, hole, top, value);
  }
  STDSORT_HD void make_heap(int64_t first, int64_t last) {
    if (last - first < 2) return;
    const int64_t len = last - first;
    int64_t parent = (len - 2) / 2;
    for (;;) {
      const uint32_t value = v[first + parent];
      adjust_heap(first, parent, len, value);
      if (parent == 0) return;
      parent--;
    }
  }
  STDSORT_HD void pop_heap(int64_t first, int64_t last, int64_t result) {
    const uint32_t value = v[result];
    v[result] = v[first];
    adjust_heap(first, 0, last - first, value);
  }
  // __partial_sort(first, middle = last, last): __heap_select degenerates to make_heap, then __sort_heap
  STDSORT_HD void heap_sort(int64_t first, int64_t last) {
    ++heap_sorts;
    make_heap(first, last);
    while (last - first > 1) { --last; pop_heap(first, last, last); }
  }

  // ---- bits/stl_algo.h ----
  STDSORT_HD void move_median_to_first(int64_t result, int64_t a, int64_t b, int64_t c) {
    if (lt(v[a], v[b])) {
      if (lt(v[b], v[c])) swap_at(result, b);
      else if (lt(v[a], v[c])) swap_at(result, c);
      else swap_at(result, a);
    } else if (lt(v[a], v[c])) swap_at(result, a);
    else if (lt(v[b], v[c])) swap_at(result, c);
    else swap_at(result, b);
  }
  STDSORT_HD int64_t unguarded_partition(int64_t first, int64_t last, int64_t pivot) {
    for (;;) {
      while (lt(v[first], v[pivot])) ++first;
      --last;
      while (lt(v[pivot], v[last])) --last;
      if (!(first < last)) return first;
      swap_at(first, last);
      ++first;
    }
  }
  STDSORT_HD void unguarded_linear_insert(int64_t last) {
    const uint32_t val = v[last];
    int64_t next = last - 1;
    while (lt(val, v[next])) { v[last] = v[next]; last = next; --next; }
    v[last] = val;
  }
  STDSORT_HD void insertion_sort(int64_t first, int64_t last) {
    if (first == last) return;
    for (int64_t i = first + 1; i != last; ++i) {
      if (lt(v[i], v[first])) {
        const uint32_t val = v[i];
        for (int64_t k = i; k > first; --k) v[k] = v[k - 1];     // move_backward(first, i, i + 1)
        v[first] = val;
      } else {
        unguarded_linear_insert(i);
      }
    }
  }
  STDSORT_HD void final_insertion_sort(int64_t first, int64_t last) {
    if (last - first > 16) {
      insertion_sort(first, first + 16);
      for (int64_t i = first + 16; i != last; ++i) unguarded_linear_insert(i);
    } else {
      insertion_sort(first, last);
    }
  }
  // __introsort_loop recurses on the right part and loops on the left; the recursion is unrolled on an explicit
  // stack of (first, last, depth_limit) — at most one pending right part per level, depth <= 2 * lg(n) <= 64.
  STDSORT_HD void introsort_loop(int64_t first0, int64_t last0, int depth0) {
    int64_t st_first[66], st_last[66];
    int st_depth[66];
    int sp = 0;
    int64_t first = first0, last = last0;
    int depth = depth0;
    for (;;) {
      // the while loop of __introsort_loop(first, last, depth)
      bool returned = false;
      while (last - first > 16) {
        if (depth == 0) { heap_sort(first, last); returned = true; break; }
        --depth;
        const int64_t mid = first + (last - first) / 2;
        move_median_to_first(first, first + 1, mid, last - 1);
        const int64_t cut = unguarded_partition(first + 1, last, first);
        // the reference recurses into [cut, last) NOW and continues with [first, cut) afterwards: process the right
        // part first and remember the left one
        st_first[sp] = first; st_last[sp] = cut; st_depth[sp] = depth; ++sp;
        first = cut;
      }
      (void)returned;
      if (sp == 0) return;
      --sp;
      first = st_first[sp]; last = st_last[sp]; depth = st_depth[sp];
    }
  }
  STDSORT_HD void sort(int64_t n) {
    if (n <= 0) return;
    int lg = 0;
    for (uint64_t m = (uint64_t)n; m > 1; m >>= 1) ++lg;        // std::__lg
    introsort_loop(0, n, 2 * lg);
    final_insertion_sort(0, n);
  }
};

template <typename Key> STDSORT_HD uint32_t sort_ids(uint32_t* ids, int64_t n, Key key) {
  Sorter<Key> s{ids, key};
  s.sort(n);
  return s.heap_sorts;
}

}  // namespace stdsort
}  // namespace hitl
