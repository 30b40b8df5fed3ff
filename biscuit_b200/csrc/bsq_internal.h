// Internals shared by the translation units of libbsq.so (not part of the ABI).
#pragma once
#include <stdarg.h>
#include <stdint.h>
#include "bsq_common.h"

struct bsq_index {
  bsq_devidx_t d;  // device pointers
  int device;
  void *allocs[32];
  int n_allocs;
  uint64_t bwt_words[2], n_sa[2];
  int64_t build_stats[4];  // GPU index build: chunks, refinement passes, largest chunk
};

void bsq_set_error(const char *fmt, ...);
bsq_index *bsq_index_alloc(int device);
void bsq_index_adopt(bsq_index *ix, void *dev_ptr);  // freed by bsq_index_free
// full suffix array in HBM?  BSQ_FULL_SA=0/1 forces it; default: when `halves_left` arrays of (n+1) x 8 B fit with 40 GB to spare
bool bsq_want_full_sa(uint64_t n, int halves_left);
// 32-byte rank blocks of both halves (bsq_fm_t::b32), derived on the device from the reference-layout blocks
int bsq_index_derive_b32(bsq_index *ix);
