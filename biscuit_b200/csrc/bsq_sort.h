// In-place introsort whose sequence of comparisons and swaps matches klib's ks_introsort
// (lib/aln/ksort.h:184-233 with its comb-sort fallback :163-183 and final insertion sort
// :150-157).  The reference's sorts are NOT stable and their tie order is part of the output
// contract, so the partition scheme -- median of (first, middle+1, last),
// pivot parked at the right end, sub-ranges of <= 16 elements left for one final insertion
// sort, depth limit 2*ceil(log2 n) -- is reproduced step for step; the code is written against
// integer indices so it runs unchanged in a GPU thread.
#pragma once
#include "bsq_common.h"

template <typename T, typename LT>
BSQ_HD void bsq_insertion_sort(T *a, int64_t n, LT lt) {
  for (int64_t i = 1; i < n; ++i)
    for (int64_t j = i; j > 0 && lt(a[j], a[j - 1]); --j) {
      T t = a[j]; a[j] = a[j - 1]; a[j - 1] = t;
    }
}

template <typename T, typename LT>
BSQ_HD void bsq_combsort(T *a, int64_t n, LT lt) {
  const double shrink = 1.2473309501039786540366528676643;
  uint64_t gap = (uint64_t)n;
  bool swapped;
  do {
    if (gap > 2) {
      gap = (uint64_t)((double)gap / shrink);
      if (gap == 9 || gap == 10) gap = 11;
    }
    swapped = false;
    for (int64_t i = 0; i + (int64_t)gap < n; ++i) {
      int64_t j = i + (int64_t)gap;
      if (lt(a[j], a[i])) { T t = a[i]; a[i] = a[j]; a[j] = t; swapped = true; }
    }
  } while (swapped || gap > 2);
  if (gap != 1) bsq_insertion_sort(a, n, lt);
}

// FINAL = false stops before the final insertion sort: what remains is a *stable* sort of
// the array as the partition phase left it, which the caller may do in any way it likes (bsq_chain_warp: in parallel).
template <bool FINAL = true, typename T, typename LT>
BSQ_HD void bsq_introsort(T *a, int64_t n, LT lt) {
  if (n < 1) return;
  if (n == 2) {
    if (lt(a[1], a[0])) { T t = a[0]; a[0] = a[1]; a[1] = t; }
    return;
  }
  int d = 2;
  while ((1ull << d) < (uint64_t)n) ++d;
  // explicit stack of pending (left, right, depth) ranges; 64-bit n needs at most 8*d+2 slots,
  // here n is a per-read quantity (< 2^31): 2*31+2 frames are plenty (one push per level).
  int64_t st_l[72], st_r[72];
  int st_d[72], top = 0;
  int64_t s = 0, t = n - 1;
  d <<= 1;
  for (;;) {
    if (s < t) {
      if (--d == 0) {
        bsq_combsort(a + s, t - s + 1, lt);
        t = s;
        continue;
      }
      int64_t i = s, j = t, k = i + ((j - i) >> 1) + 1;
      if (lt(a[k], a[i])) {
        if (lt(a[k], a[j])) k = j;
      } else k = lt(a[j], a[i]) ? i : j;
      T rp = a[k];
      if (k != t) { T x = a[k]; a[k] = a[t]; a[t] = x; }
      for (;;) {
        do ++i; while (lt(a[i], rp));
        do --j; while (i <= j && lt(rp, a[j]));
        if (j <= i) break;
        T x = a[i]; a[i] = a[j]; a[j] = x;
      }
      { T x = a[i]; a[i] = a[t]; a[t] = x; }
      if (i - s > t - i) {
        if (i - s > 16) { st_l[top] = s; st_r[top] = i - 1; st_d[top] = d; ++top; }
        s = t - i > 16 ? i + 1 : t;
      } else {
        if (t - i > 16) { st_l[top] = i + 1; st_r[top] = t; st_d[top] = d; ++top; }
        t = i - s > 16 ? i - 1 : s;
      }
    } else {
      if (top == 0) {
        if (FINAL) bsq_insertion_sort(a, n, lt);
        return;
      }
      --top;
      s = st_l[top]; t = st_r[top]; d = st_d[top];
    }
  }
}
