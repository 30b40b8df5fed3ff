/* TEST INFRASTRUCTURE ONLY -- stand-in for huishenlab/utils wqueue.h (see README.md): bounded blocking FIFO between
 * producer and consumer threads (pileup.c:1140-1212). */
#ifndef BSQ_SHIM_WQUEUE_H
#define BSQ_SHIM_WQUEUE_H
#include <pthread.h>
#include <stdlib.h>
#define wqueue_t(name) wqueue_##name##_t
#define DEFINE_WQUEUE(name, type)                                                                   \
  typedef struct {                                                                                  \
    type *buf; size_t cap, head, n;                                                                 \
    pthread_mutex_t mu; pthread_cond_t not_full, not_empty;                                         \
  } wqueue_##name##_t;                                                                              \
  static inline wqueue_##name##_t *wqueue_init_##name(size_t cap) {                                 \
    wqueue_##name##_t *q = (wqueue_##name##_t *)calloc(1, sizeof *q);                               \
    q->cap = cap; q->buf = (type *)malloc(cap * sizeof(type));                                      \
    pthread_mutex_init(&q->mu, 0); pthread_cond_init(&q->not_full, 0); pthread_cond_init(&q->not_empty, 0); \
    return q;                                                                                       \
  }                                                                                                 \
  static inline void wqueue_destroy_##name(wqueue_##name##_t *q) {                                  \
    pthread_mutex_destroy(&q->mu); pthread_cond_destroy(&q->not_full); pthread_cond_destroy(&q->not_empty); \
    free(q->buf); free(q);                                                                          \
  }                                                                                                 \
  static inline void wqueue_put_##name(wqueue_##name##_t *q, const type *e) {                       \
    pthread_mutex_lock(&q->mu);                                                                     \
    while (q->n == q->cap) pthread_cond_wait(&q->not_full, &q->mu);                                 \
    q->buf[(q->head + q->n++) % q->cap] = *e;                                                       \
    pthread_cond_signal(&q->not_empty);                                                             \
    pthread_mutex_unlock(&q->mu);                                                                   \
  }                                                                                                 \
  static inline void wqueue_put2_##name(wqueue_##name##_t *q, type e) { wqueue_put_##name(q, &e); } \
  static inline void wqueue_get_##name(wqueue_##name##_t *q, type *e) {                             \
    pthread_mutex_lock(&q->mu);                                                                     \
    while (q->n == 0) pthread_cond_wait(&q->not_empty, &q->mu);                                     \
    *e = q->buf[q->head]; q->head = (q->head + 1) % q->cap; q->n--;                                 \
    pthread_cond_signal(&q->not_full);                                                              \
    pthread_mutex_unlock(&q->mu);                                                                   \
  }
#define wqueue_init(name, cap) wqueue_init_##name(cap)
#define wqueue_destroy(name, q) wqueue_destroy_##name(q)
#define wqueue_put(name, q, e) wqueue_put_##name(q, e)
#define wqueue_put2(name, q, e) wqueue_put2_##name(q, e)
#define wqueue_get(name, q, e) wqueue_get_##name(q, e)
#endif
