/* Shim: ksprintf/kvsprintf are declared in lib/aln/kstring.h:72-73 and normally come from
 * htslib's kstring.c (absent). TEST INFRASTRUCTURE ONLY. */
#include <stdarg.h>
#include <stdio.h>
#include "kstring.h"

int kvsprintf(kstring_t *s, const char *fmt, va_list ap) {
  va_list args;
  int l;
  va_copy(args, ap);
  l = vsnprintf(s->s + s->l, s->m - s->l, fmt, args);
  va_end(args);
  if (l + 1 > (int)(s->m - s->l)) {
    ks_resize(s, s->l + l + 2);
    va_copy(args, ap);
    l = vsnprintf(s->s + s->l, s->m - s->l, fmt, args);
    va_end(args);
  }
  s->l += l;
  return l;
}

int ksprintf(kstring_t *s, const char *fmt, ...) {
  va_list ap; int l;
  va_start(ap, fmt);
  l = kvsprintf(s, fmt, ap);
  va_end(ap);
  return l;
}
