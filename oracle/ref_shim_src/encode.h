/* TEST INFRASTRUCTURE ONLY -- stand-in for huishenlab/utils encode.h as used by /root/reference/src (see README.md). */
#ifndef BSQ_SHIM_SRC_ENCODE_H
#define BSQ_SHIM_SRC_ENCODE_H
#include <stdint.h>
/* A/a 0, C/c 1, G/g 2, T/t 3, everything else 4 (call sites: pileup.c:423,800,808) */
static const uint8_t nt256char_to_nt256int8_table[256] = {
    4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4, 4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,
    4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4, 4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,
    4,0,4,1,4,4,4,2,4,4,4,4,4,4,4,4, 4,4,4,4,3,4,4,4,4,4,4,4,4,4,4,4,
    4,0,4,1,4,4,4,2,4,4,4,4,4,4,4,4, 4,4,4,4,3,4,4,4,4,4,4,4,4,4,4,4,
    4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4, 4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,
    4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4, 4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,
    4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4, 4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,
    4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4, 4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4};
static inline char nt256char_complement(char c) {
  switch (c) {
    case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A';
    case 'a': return 't'; case 'c': return 'g'; case 'g': return 'c'; case 't': return 'a';
    default: return c;
  }
}
/* reverse complement in place (call site: bisc_utils.c:52) */
static inline void nt256char_rev_ip(char *s, int len) {
  for (int i = 0, j = len - 1; i <= j; ++i, --j) {
    char a = nt256char_complement(s[i]), b = nt256char_complement(s[j]);
    s[i] = b; s[j] = a;
  }
}
#endif
