// Per-task entry points shared by the CUDA kernels (bsq_kernels.cu) and the test-only host
// emulation (tests/hostemu).  A task is one read against one bisulfite conversion
// (mem_align1_core, lib/aln/bwamem.c:183-208).
#pragma once
#include "bsq_chain.h"
#include "bsq_region.h"
#include "bsq_seed.h"

// bseq_bsconvert (lib/aln/bwamem.c:161-178): parent -> C>T, daughter -> G>A
BSQ_HD void bsq_bsconvert(const uint8_t *seq, int len, int parent, uint8_t *out) {
  for (int i = 0; i < len; ++i) {
    uint8_t c = seq[i];
    out[i] = parent ? (c == 1 ? 3 : c) : (c == 2 ? 0 : c);
  }
}

// Seeding of one task.  Returns the interval count (or -1 on overflow) and, in *n_sa, how many
// suffix-array lookups the chaining stage will need up front: min(x[2], max_occ) per interval
// (the occurrences mem_chain always visits, memchain.c:325-326).
BSQ_HD int bsq_task_seed(const bsq_devopt_t &opt, const bsq_devidx_t &ix, const uint8_t *seq, int len, int parent,
                         bool pipeline, bsq_seed_scratch_t &scr, bsq_intv_t *out, int32_t *n_sa) {
  uint8_t q[BSQ_MAX_READ_LEN];
  *n_sa = 0;
  if (pipeline && len < opt.min_seed_len) return 0;  // mem_chain returns before seeding (memchain.c:280)
  bsq_bsconvert(seq, len, parent, q);
  int n = bsq_collect_intv(opt, ix.fm[parent], ix.fm[!parent], len, q, scr, out, BSQ_MAX_INTV);
  if (n < 0) return -1;
  int64_t tot = 0;
  for (int i = 0; i < n; ++i) tot += (int64_t)(out[i].x[2] < (uint64_t)(uint32_t)opt.max_occ ? out[i].x[2] : (uint64_t)(uint32_t)opt.max_occ);
  *n_sa = (int32_t)tot;
  return n;
}

// BWT ranks whose text positions the chaining stage needs, in visiting order.
BSQ_HD void bsq_task_expand(const bsq_devopt_t &opt, const bsq_intv_t *intv, int n, uint64_t *ranks) {
  int64_t o = 0;
  for (int i = 0; i < n; ++i) {
    uint64_t m = intv[i].x[2] < (uint64_t)(uint32_t)opt.max_occ ? intv[i].x[2] : (uint64_t)(uint32_t)opt.max_occ;
    for (uint64_t k = 0; k < m; ++k) ranks[o++] = intv[i].x[0] + k;
  }
}
