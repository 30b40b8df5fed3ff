/* TEST INFRASTRUCTURE ONLY -- stand-in for htslib kstring.h (see README.md). */
#ifndef BSQ_SHIM_HTS_KSTRING_H
#define BSQ_SHIM_HTS_KSTRING_H
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifndef kroundup32
#define kroundup32(x) (--(x), (x)|=(x)>>1, (x)|=(x)>>2, (x)|=(x)>>4, (x)|=(x)>>8, (x)|=(x)>>16, ++(x))
#endif
typedef struct kstring_t { size_t l, m; char *s; } kstring_t;
static inline int ks_resize(kstring_t *s, size_t size) {
  if (s->m < size) {
    size_t m = size < 16 ? 16 : size;
    m += m >> 1;
    char *t = (char *)realloc(s->s, m);
    if (!t) return -1;
    s->s = t; s->m = m;
  }
  return 0;
}
static inline int kputsn(const char *p, size_t l, kstring_t *s) {
  if (ks_resize(s, s->l + l + 2) < 0) return EOF;
  memcpy(s->s + s->l, p, l); s->l += l; s->s[s->l] = 0;
  return (int)l;
}
static inline int kputs(const char *p, kstring_t *s) { return kputsn(p, strlen(p), s); }
static inline int kputc(int c, kstring_t *s) {
  if (ks_resize(s, s->l + 2) < 0) return EOF;
  s->s[s->l++] = (char)c; s->s[s->l] = 0;
  return c;
}
static inline int kputuw(unsigned x, kstring_t *s) {
  char buf[16]; int n = snprintf(buf, sizeof buf, "%u", x);
  return kputsn(buf, (size_t)n, s);
}
static inline int kvsprintf(kstring_t *s, const char *fmt, va_list ap) {
  va_list args;
  va_copy(args, ap);
  int l = vsnprintf(s->s ? s->s + s->l : NULL, s->s ? s->m - s->l : 0, fmt, args);
  va_end(args);
  if (l < 0) return -1;
  if ((size_t)l + 1 > s->m - s->l || !s->s) {
    if (ks_resize(s, s->l + (size_t)l + 2) < 0) return -1;
    va_copy(args, ap);
    l = vsnprintf(s->s + s->l, s->m - s->l, fmt, args);
    va_end(args);
  }
  s->l += (size_t)l;
  return l;
}
static inline int ksprintf(kstring_t *s, const char *fmt, ...) {
  va_list ap; va_start(ap, fmt);
  int l = kvsprintf(s, fmt, ap);
  va_end(ap);
  return l;
}
#endif
