// Result and scratch types of the banded affine-gap extension (ksw_extend2, lib/aln/ksw.c:380-479).  The device
// implementation is bsq_ksw_warp.cuh (one DP row striped over a warp); the scalar form used by the test-only host
// emulation lives in tests/hostemu/bsq_ksw_scalar.h.
#pragma once
#include "bsq_common.h"

struct bsq_ext_result_t {
  int32_t score, qle, tle, gtle, gscore, max_off;
};

struct bsq_eh_t {
  int32_t h, e;
};

// scratch row: H(i-1,j-1) / E(i,j) per query column (the reference's eh[] array)
struct bsq_ksw_scratch_t {
  bsq_eh_t eh[BSQ_MAX_READ_LEN + 2];
};

BSQ_HD int bsq_iabs(int v) { return v < 0 ? -v : v; }
