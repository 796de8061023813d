// walt_stdsort.cuh -- the order libstdc++'s std::sort leaves EQUIVALENT elements in.
//
// The reference sorts every hash bucket with std::sort (reference.cpp:290-300), which is
// unstable: where several suffixes compare equal under SortHashTableBucketCMP
// (reference.cpp:258-288) -- repeats longer than the 178 compared bases, or suffixes cut by the
// same chromosome end -- their order in the .dbindex file is whatever the introsort's sequence
// of swaps produced from the bucket's initial arrangement (ascending position, HashToBucket,
// reference.cpp:231-256).  That order is observable: SingleEndMapping keeps the LAST of
// several equal-best candidates (mapping.cpp:306-313).  To write byte-identical index files the
// device builder first radix-sorts (which yields, per element, the rank of its equivalence
// class), then replays the exact libstdc++ algorithm per bucket on those integer ranks, so every
// comparison has the outcome the reference's comparator had and every swap happens in the same
// place, without touching the genome again.
//
// Transcribed from GCC 13 libstdc++ (the toolchain the reference is built with here):
//   bits/stl_algo.h:84-101 (__move_median_to_first), 1792-1866 (insertion sorts, threshold 16),
//   1871-1900 (__unguarded_partition[_pivot]), 1905-1913 (__partial_sort), 1918-1936
//   (__introsort_loop), 1941-1952 (__sort);  bits/stl_heap.h:135-147 (__push_heap), 224-249
//   (__adjust_heap), 254-267 (__pop_heap), 340-362 (__make_heap), 419-427 (__sort_heap).
// Compiles for the device and, for the CPU-only tests, for the host (tests/emu).
#pragma once

#include <stdint.h>

#if defined(__CUDACC__)
#define WALT_SORT_HD __host__ __device__ __forceinline__
#else
#define WALT_SORT_HD inline
#endif

namespace waltsort {

// An element is (position, class rank): two parallel u32 arrays addressed by one index.
struct Elem { uint32_t pos, cls; };

struct PairSeq {
  uint32_t* pos;
  uint32_t* cls;
  WALT_SORT_HD Elem get(int64_t i) const { return Elem{pos[i], cls[i]}; }
  WALT_SORT_HD void set(int64_t i, Elem v) const { pos[i] = v.pos; cls[i] = v.cls; }
  WALT_SORT_HD void move(int64_t dst, int64_t src) const { pos[dst] = pos[src]; cls[dst] = cls[src]; }
  WALT_SORT_HD void swap(int64_t a, int64_t b) const { const Elem t = get(a); move(a, b); set(b, t); }
};

// the two orders used: by class rank (= the reference comparator's outcome) and by position
struct ByClass {
  WALT_SORT_HD bool operator()(const Elem& a, const Elem& b) const { return a.cls < b.cls; }
};
struct ByPos {
  WALT_SORT_HD bool operator()(const Elem& a, const Elem& b) const { return a.pos < b.pos; }
};

// ---- stl_heap.h ----------------------------------------------------------------------------
template <class Cmp>
WALT_SORT_HD void push_heap_(const PairSeq& s, int64_t first, int64_t hole, int64_t top, Elem value, Cmp comp) {
  int64_t parent = (hole - 1) / 2;
  while (hole > top && comp(s.get(first + parent), value)) {
    s.move(first + hole, first + parent);
    hole = parent;
    parent = (hole - 1) / 2;
  }
  s.set(first + hole, value);
}

template <class Cmp>
WALT_SORT_HD void adjust_heap_(const PairSeq& s, int64_t first, int64_t hole, int64_t len, Elem value, Cmp comp) {
  const int64_t top = hole;
  int64_t child = hole;
  while (child < (len - 1) / 2) {
    child = 2 * (child + 1);
    if (comp(s.get(first + child), s.get(first + (child - 1)))) child--;
    s.move(first + hole, first + child);
    hole = child;
  }
  if ((len & 1) == 0 && child == (len - 2) / 2) {
    child = 2 * (child + 1);
    s.move(first + hole, first + (child - 1));
    hole = child - 1;
  }
  push_heap_(s, first, hole, top, value, comp);
}

template <class Cmp>
WALT_SORT_HD void pop_heap_(const PairSeq& s, int64_t first, int64_t last, int64_t result, Cmp comp) {
  const Elem value = s.get(result);
  s.move(result, first);
  adjust_heap_(s, first, 0, last - first, value, comp);
}

template <class Cmp>
WALT_SORT_HD void make_heap_(const PairSeq& s, int64_t first, int64_t last, Cmp comp) {
  if (last - first < 2) return;
  const int64_t len = last - first;
  int64_t parent = (len - 2) / 2;
  for (;;) {
    const Elem value = s.get(first + parent);
    adjust_heap_(s, first, parent, len, value, comp);
    if (parent == 0) return;
    parent--;
  }
}

template <class Cmp>
WALT_SORT_HD void sort_heap_(const PairSeq& s, int64_t first, int64_t last, Cmp comp) {
  while (last - first > 1) {
    --last;
    pop_heap_(s, first, last, last, comp);
  }
}

// __partial_sort(first, last, last): __heap_select's loop over [middle, last) is empty
template <class Cmp>
WALT_SORT_HD void heap_sort_(const PairSeq& s, int64_t first, int64_t last, Cmp comp) {
  make_heap_(s, first, last, comp);
  sort_heap_(s, first, last, comp);
}

// ---- stl_algo.h ----------------------------------------------------------------------------
template <class Cmp>
WALT_SORT_HD void move_median_to_first_(const PairSeq& s, int64_t result, int64_t a, int64_t b, int64_t c, Cmp comp) {
  const Elem ea = s.get(a), eb = s.get(b), ec = s.get(c);
  if (comp(ea, eb)) {
    if (comp(eb, ec)) s.swap(result, b);
    else if (comp(ea, ec)) s.swap(result, c);
    else s.swap(result, a);
  } else if (comp(ea, ec)) s.swap(result, a);
  else if (comp(eb, ec)) s.swap(result, c);
  else s.swap(result, b);
}

template <class Cmp>
WALT_SORT_HD int64_t unguarded_partition_(const PairSeq& s, int64_t first, int64_t last, int64_t pivot, Cmp comp) {
  const Elem pv = s.get(pivot);   // the pivot slot (range start) is never written by the loop
  for (;;) {
    while (comp(s.get(first), pv)) ++first;
    --last;
    while (comp(pv, s.get(last))) --last;
    if (!(first < last)) return first;
    s.swap(first, last);
    ++first;
  }
}

template <class Cmp>
WALT_SORT_HD int64_t unguarded_partition_pivot_(const PairSeq& s, int64_t first, int64_t last, Cmp comp) {
  const int64_t mid = first + (last - first) / 2;
  move_median_to_first_(s, first, first + 1, mid, last - 1, comp);
  return unguarded_partition_(s, first + 1, last, first, comp);
}

template <class Cmp>
WALT_SORT_HD void unguarded_linear_insert_(const PairSeq& s, int64_t last, Cmp comp) {
  const Elem val = s.get(last);
  int64_t next = last - 1;
  while (comp(val, s.get(next))) {
    s.move(last, next);
    last = next;
    --next;
  }
  s.set(last, val);
}

template <class Cmp>
WALT_SORT_HD void insertion_sort_(const PairSeq& s, int64_t first, int64_t last, Cmp comp) {
  if (first == last) return;
  for (int64_t i = first + 1; i != last; ++i) {
    if (comp(s.get(i), s.get(first))) {
      const Elem val = s.get(i);
      for (int64_t j = i; j > first; --j) s.move(j, j - 1);   // move_backward(first, i, i + 1)
      s.set(first, val);
    } else {
      unguarded_linear_insert_(s, i, comp);
    }
  }
}

template <class Cmp>
WALT_SORT_HD void final_insertion_sort_(const PairSeq& s, int64_t first, int64_t last, Cmp comp) {
  constexpr int64_t THRESHOLD = 16;
  if (last - first > THRESHOLD) {
    insertion_sort_(s, first, first + THRESHOLD, comp);
    for (int64_t i = first + THRESHOLD; i != last; ++i) unguarded_linear_insert_(s, i, comp);
  } else {
    insertion_sort_(s, first, last, comp);
  }
}

// std::sort(first, last, comp) over s[first, last).  The recursion of __introsort_loop
// (recurse on the right part, loop on the left) is unrolled onto an explicit stack; its depth is
// bounded by depth_limit = 2 * floor(log2(n)) <= 62 for n < 2^32.  Returns how many ranges fell
// back to the heap sort (depth limit reached) -- of interest to the tests only.
template <class Cmp>
WALT_SORT_HD int std_sort(const PairSeq& s, int64_t first, int64_t last, Cmp comp) {
  if (first == last) return 0;
  int fallbacks = 0;
  constexpr int64_t THRESHOLD = 16;
  int lg = 0;
  for (uint64_t n = (uint64_t)(last - first); n > 1; n >>= 1) ++lg;   // std::__lg
  struct Frame { int64_t first, last; int depth; };
  Frame stack[72];
  int sp = 0;
  stack[sp++] = Frame{first, last, 2 * lg};
  while (sp > 0) {
    Frame f = stack[--sp];
    // one activation of __introsort_loop(f.first, f.last, f.depth): each iteration calls itself
    // on [cut, last) BEFORE continuing with [first, cut), so the right part must be fully
    // processed first -- push the continuation (left), then the right part on top of it.
    if (f.last - f.first > THRESHOLD) {
      if (f.depth == 0) {
        heap_sort_(s, f.first, f.last, comp);
        ++fallbacks;
        continue;
      }
      const int depth = f.depth - 1;
      const int64_t cut = unguarded_partition_pivot_(s, f.first, f.last, comp);
      stack[sp++] = Frame{f.first, cut, depth};   // "last = cut" and loop: same depth counter
      stack[sp++] = Frame{cut, f.last, depth};    // the recursive call, runs first
    }
  }
  final_insertion_sort_(s, first, last, comp);
  return fallbacks;
}

// ---- the same sort as a tree of independent tasks ----------------------------------------------
// After a partition the two sides never interact again: __introsort_loop sorts [cut, last) by
// recursion and [first, cut) by iteration, and the closing __final_insertion_sort never moves an
// element across a partition boundary (everything left of a cut is <= everything right of it,
// and the insertion loops stop at the first element that is not greater).  Inside one leaf
// (a range of <= 16 elements, or a heap-sorted range) the final pass is a plain stable insertion
// sort.  So std::sort == process(first, last, 2 * lg(n)) with
//     process(f, l, d):  l - f <= 16 -> stable insertion sort
//                        d == 0      -> heap sort
//                        else        -> cut = partition(f, l); process(f, cut, d-1); process(cut, l, d-1)
// and the two recursive calls may run concurrently.  The device builder runs this level by level
// (one thread per task, children appended to the next level's list), which turns the n log n
// serial steps of a 500 000-entry bucket into ~2n.
struct SortTask { uint32_t first, last, depth; };   // slots [first, last) of the whole index

WALT_SORT_HD uint32_t depth_limit_for(uint64_t n) {
  uint32_t lg = 0;
  for (; n > 1; n >>= 1) ++lg;
  return 2u * lg;
}

// One task.  Children that still need partitioning are written to out[0..return value).
template <class Cmp>
WALT_SORT_HD int sort_task_step(const PairSeq& s, SortTask t, Cmp comp, SortTask out[2]) {
  constexpr int64_t THRESHOLD = 16;
  const int64_t first = t.first, last = t.last;
  if (last - first <= THRESHOLD) { insertion_sort_(s, first, last, comp); return 0; }
  if (t.depth == 0) { heap_sort_(s, first, last, comp); return 0; }
  const int64_t cut = unguarded_partition_pivot_(s, first, last, comp);
  int k = 0;
  if (cut - first <= THRESHOLD) insertion_sort_(s, first, cut, comp);
  else out[k++] = SortTask{(uint32_t)first, (uint32_t)cut, t.depth - 1u};
  if (last - cut <= THRESHOLD) insertion_sort_(s, cut, last, comp);
  else out[k++] = SortTask{(uint32_t)cut, (uint32_t)last, t.depth - 1u};
  return k;
}

}  // namespace waltsort
