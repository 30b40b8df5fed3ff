/* Shim for huishenlab/utils wzmisc.h (absent dependency, CMakeLists.txt:45-54 of the
 * reference). TEST INFRASTRUCTURE ONLY: lets /root/reference/lib/aln compile untouched
 * into oracle/_ref. Provides only what lib/aln uses: wzfatal, min, max. */
#ifndef BSQ_SHIM_WZMISC_H
#define BSQ_SHIM_WZMISC_H
#include <stdio.h>
#include <stdlib.h>
#include <stdarg.h>
#ifndef min
#define min(a, b) ({ __typeof__(a) _a = (a); __typeof__(b) _b = (b); _a < _b ? _a : _b; })
#endif
#ifndef max
#define max(a, b) ({ __typeof__(a) _a = (a); __typeof__(b) _b = (b); _a > _b ? _a : _b; })
#endif
static inline void wzfatal(const char *fmt, ...) {
  va_list ap; va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap);
  fputc('\n', stderr); exit(1);
}
#endif
