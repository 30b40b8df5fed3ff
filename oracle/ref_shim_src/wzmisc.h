/* TEST INFRASTRUCTURE ONLY -- stand-in for huishenlab/utils wzmisc.h as used by /root/reference/src (see README.md). */
#ifndef BSQ_SHIM_SRC_WZMISC_H
#define BSQ_SHIM_SRC_WZMISC_H
#include <ctype.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifndef min
#define min(a, b) ({ __typeof__(a) _a = (a); __typeof__(b) _b = (b); _a > _b ? _b : _a; })
#endif
#ifndef max
#define max(a, b) ({ __typeof__(a) _a = (a); __typeof__(b) _b = (b); _a > _b ? _a : _b; })
#endif
static inline void wzfatal(const char *fmt, ...) {
  va_list ap; va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap);
  fflush(stderr); exit(1);
}
static inline void wzstrupr(char *s) { for (; *s; ++s) *s = (char)toupper((unsigned char)*s); }
/* plain decimal number: optional sign, digits, optional fraction / exponent */
static inline int is_number(const char *s) {
  if (!s || !*s) return 0;
  char *e; (void)strtod(s, &e);
  return e != s && *e == 0;
}
static inline char *strcpy_realloc(char *dest, const char *src) {
  dest = (char *)realloc(dest, strlen(src) + 1);
  strcpy(dest, src);
  return dest;
}
static inline void free_char_array(char **a, int n) {
  if (!a) return;
  for (int i = 0; i < n; ++i) free(a[i]);
  free(a);
}
/* split on any character of sep; empty fields are kept */
static inline void line_get_fields(const char *line, const char *sep, char ***fields, int *nfields) {
  int n = 0, cap = 16;
  char **f = (char **)malloc((size_t)cap * sizeof(char *));
  const char *p = line;
  for (;;) {
    size_t l = strcspn(p, sep);
    if (n == cap) { cap *= 2; f = (char **)realloc(f, (size_t)cap * sizeof(char *)); }
    f[n] = (char *)malloc(l + 1);
    memcpy(f[n], p, l); f[n][l] = 0; ++n;
    if (!p[l]) break;
    p += l + 1;
  }
  *fields = f; *nfields = n;
}
#endif
