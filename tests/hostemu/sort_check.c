/* TEST ONLY: the typed introsort instances of biscuit_b200/host/bq_sort.h against the generic bq_introsort (bq_core.c) on
 * arrays with many equal keys: the two must perform the same comparisons and swaps, i.e. leave equal keys in the same
 * order (the payload tells them apart). */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../../biscuit_b200/host/bq.h"
#include "../../biscuit_b200/host/bq_sort.h"

typedef struct { int key; int payload; char pad[120]; } big_t;   /* the size of bq_reg_t */
typedef struct { unsigned long x, y; } small_t;
static int lt_big(const void *a, const void *b) { return ((const big_t *)a)->key < ((const big_t *)b)->key; }
static int lt_small(const void *a_, const void *b_) { const small_t *a = a_, *b = b_; return a->x < b->x || (a->x == b->x && (a->y >> 40) < (b->y >> 40)); }
BQ_INTROSORT_DEFINE(sort_big, big_t, lt_big)
BQ_INTROSORT_DEFINE(sort_small, small_t, lt_small)

static unsigned long rng_state = 88172645463325252ul;
static unsigned long rnd(void) { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return rng_state; }

int main(void) {
  long n_cases = 0;
  for (int trial = 0; trial < 6000; ++trial) {
    const int n = trial < 3000 ? (int)(rnd() % 40) : (int)(rnd() % 3000);
    const int n_keys = 1 + (int)(rnd() % (trial % 3 == 0 ? 3 : (trial % 3 == 1 ? 17 : 100000)));
    const int shape = trial % 5; /* random, ascending, descending, organ pipe, few runs */
    big_t *a = malloc(sizeof(big_t) * (size_t)(n + 1)), *b = malloc(sizeof(big_t) * (size_t)(n + 1));
    small_t *c = malloc(sizeof(small_t) * (size_t)(n + 1)), *d = malloc(sizeof(small_t) * (size_t)(n + 1));
    for (int i = 0; i < n; ++i) {
      int k = (int)(rnd() % (unsigned long)n_keys);
      if (shape == 1) k = i * n_keys / (n + 1);
      else if (shape == 2) k = (n - i) * n_keys / (n + 1);
      else if (shape == 3) k = (i < n / 2 ? i : n - i) * n_keys / (n + 1);
      else if (shape == 4) k = (i / 7) % n_keys;
      memset(&a[i], 0, sizeof a[i]);
      a[i].key = k; a[i].payload = i;
      c[i].x = (unsigned long)k; c[i].y = rnd() << 40 | (unsigned long)i;
    }
    memcpy(b, a, sizeof(big_t) * (size_t)n); memcpy(d, c, sizeof(small_t) * (size_t)n);
    bq_introsort(a, (size_t)n, sizeof(big_t), lt_big); sort_big(b, (size_t)n);
    bq_introsort(c, (size_t)n, sizeof(small_t), lt_small); sort_small(d, (size_t)n);
    if (memcmp(a, b, sizeof(big_t) * (size_t)n) || memcmp(c, d, sizeof(small_t) * (size_t)n)) { printf("MISMATCH trial %d n %d keys %d shape %d\n", trial, n, n_keys, shape); return 1; }
    for (int i = 1; i < n; ++i) if (a[i].key < a[i - 1].key) { printf("NOT SORTED trial %d\n", trial); return 1; }
    ++n_cases;
    free(a); free(b); free(c); free(d);
  }
  printf("ok %ld cases\n", n_cases);
  return 0;
}
