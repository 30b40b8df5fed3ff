/* TEST INFRASTRUCTURE ONLY -- stand-in for huishenlab/utils wvec.h (see README.md): a typed growable array with the
 * accessor names the reference's call sites use (init_/free_/ref_/get_/next_ref_/push_). */
#ifndef BSQ_SHIM_WVEC_H
#define BSQ_SHIM_WVEC_H
#include <stdlib.h>
#include <string.h>
#define DEFINE_VECTOR(name, element_type)                                                                     \
  typedef struct { element_type *buffer; size_t size; size_t capacity; } name;                                 \
  static inline name *init_##name(size_t init_capacity) {                                                      \
    name *v = (name *)calloc(1, sizeof(name));                                                                 \
    v->capacity = init_capacity ? init_capacity : 2;                                                           \
    v->buffer = (element_type *)malloc(v->capacity * sizeof(element_type));                                    \
    return v;                                                                                                  \
  }                                                                                                            \
  static inline void free_##name(name *v) { if (v) { free(v->buffer); free(v); } }                             \
  static inline element_type *ref_##name(name *v, size_t i) { return v->buffer + i; }                          \
  static inline element_type get_##name(name *v, size_t i) { return v->buffer[i]; }                            \
  static inline element_type *next_ref_##name(name *v) {                                                       \
    if (v->size + 1 > v->capacity) {                                                                           \
      v->capacity <<= 1;                                                                                       \
      v->buffer = (element_type *)realloc(v->buffer, v->capacity * sizeof(element_type));                      \
    }                                                                                                          \
    return v->buffer + v->size++;                                                                              \
  }                                                                                                            \
  static inline void push_##name(name *v, element_type e) { *next_ref_##name(v) = e; }                         \
  static inline void clear_##name(name *v) { v->size = 0; }
#endif
