/* bq_sort.h -- typed instances of the introsort of bq_core.c (bq_introsort): the same algorithm, statement for statement
 * (klib's ks_introsort with its exact comparison / swap sequence, ksort.h:150-233 -- the order of equal keys in the
 * output depends on it and shows in the SAM text), generated per element type so that the comparator is inlined and an
 * element is moved by assignment instead of three memcpy calls through a byte buffer.  LT(x, y) takes two pointers. */
#ifndef BQ_SORT_H
#define BQ_SORT_H
#include <stddef.h>

#define BQ_INTROSORT_DEFINE(NAME, T, LT)                                                                         \
  static void NAME##_ins(T *s, T *t) {                                                                           \
    for (T *i = s + 1; i < t; ++i)                                                                               \
      for (T *j = i; j > s && LT(j, j - 1); --j) { T tmp_ = *j; *j = *(j - 1); *(j - 1) = tmp_; }                   \
  }                                                                                                              \
  static void NAME##_comb(T *a, size_t n) {                                                                      \
    const double shrink = 1.2473309501039786540366528676643;                                                     \
    size_t gap = n;                                                                                              \
    int swapped;                                                                                                 \
    do {                                                                                                         \
      if (gap > 2) { gap = (size_t)(gap / shrink); if (gap == 9 || gap == 10) gap = 11; }                        \
      swapped = 0;                                                                                               \
      for (T *i = a; i < a + (n - gap); ++i) {                                                                   \
        T *j = i + gap;                                                                                          \
        if (LT(j, i)) { T tmp_ = *i; *i = *j; *j = tmp_; swapped = 1; }                                           \
      }                                                                                                          \
    } while (swapped || gap > 2);                                                                                \
    if (gap != 1) NAME##_ins(a, a + n);                                                                          \
  }                                                                                                              \
  static void NAME(T *a, size_t n) {                                                                             \
    struct { T *l, *r; int d; } st[80], *top = st;                                                               \
    T rp;                                                                                                        \
    if (n < 1) return;                                                                                           \
    if (n == 2) { if (LT(a + 1, a)) { T tmp_ = a[0]; a[0] = a[1]; a[1] = tmp_; } return; }                       \
    int d = 2;                                                                                                   \
    while ((1ul << d) < n) ++d;                                                                                  \
    T *s = a, *t = a + (n - 1);                                                                                  \
    d <<= 1;                                                                                                     \
    for (;;) {                                                                                                   \
      if (s < t) {                                                                                               \
        if (--d == 0) { NAME##_comb(s, (size_t)(t - s) + 1); t = s; continue; }                                  \
        T *i = s, *j = t, *k = i + ((size_t)(j - i) >> 1) + 1;                                                   \
        if (LT(k, i)) { if (LT(k, j)) k = j; }                                                                   \
        else k = LT(j, i) ? i : j;                                                                               \
        rp = *k;                                                                                                 \
        if (k != t) { T tmp_ = *k; *k = *t; *t = tmp_; }                                                         \
        for (;;) {                                                                                               \
          do ++i; while (LT(i, &rp));                                                                            \
          do --j; while (i <= j && LT(&rp, j));                                                                  \
          if (j <= i) break;                                                                                     \
          { T tmp_ = *i; *i = *j; *j = tmp_; }                                                                   \
        }                                                                                                        \
        { T tmp_ = *i; *i = *t; *t = tmp_; }                                                                     \
        if (i - s > t - i) {                                                                                     \
          if (i - s > 16) { top->l = s; top->r = i - 1; top->d = d; ++top; }                                     \
          s = t - i > 16 ? i + 1 : t;                                                                            \
        } else {                                                                                                 \
          if (t - i > 16) { top->l = i + 1; top->r = t; top->d = d; ++top; }                                     \
          t = i - s > 16 ? i - 1 : s;                                                                            \
        }                                                                                                        \
      } else {                                                                                                   \
        if (top == st) { NAME##_ins(a, a + n); return; }                                                         \
        --top; s = top->l; t = top->r; d = top->d;                                                               \
      }                                                                                                          \
    }                                                                                                            \
  }
#endif
