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

// Seeding of one task, driven sequentially (host emulation; the CUDA kernel k_seed drives the same
// machine with all lanes of a warp meeting at the extend).  Returns the interval count (or -1 on
// overflow); out[] is packed and sorted; *n_sa = SA lookups chaining will need up front.
BSQ_HD int bsq_task_seed(const bsq_devopt_t &opt, const bsq_devidx_t &ix, const uint8_t *seq, int len, int parent,
                         bool pipeline, bsq_seed_scratch_t &scr, bsq_pk_t *out, int32_t *n_sa) {
  *n_sa = 0;
  if (pipeline && len < opt.min_seed_len) return 0;  // mem_chain returns before seeding (memchain.c:280)
  scr.bind(seq, parent);
  bsq_seed_machine_t m;
  bsq_sm_init(m, opt, len, BSQ_MAX_INTV);
  bsq_ext_req_t req;
  const bsq_fm_t &fm = ix.fm[parent], &fmc = ix.fm[!parent];
  while (bsq_sm_next(m, fm, fmc, scr, out, req)) {
    uint64_t o0, o1, o2;
    bsq_extend1(fm, fmc, req, o0, o1, o2);
    bsq_sm_consume(m, scr, out, req, o0, o1, o2);
  }
  if (m.overflow) return -1;
  uint32_t keys[BSQ_MAX_INTV];
  *n_sa = bsq_seed_sort(opt, out, m.n_out, keys);
  return m.n_out;
}

// BWT ranks whose text positions the chaining stage needs, in visiting order.
BSQ_HD void bsq_task_expand(const bsq_devopt_t &opt, const bsq_pk_t *intv, int n, uint64_t tag, uint64_t *ranks) {
  int64_t o = 0;
  for (int i = 0; i < n; ++i) {
    const uint64_t x2 = bsq_pk_x2(intv[i]), x0 = bsq_pk_x0(intv[i]);
    const uint64_t m = x2 < (uint64_t)(uint32_t)opt.max_occ ? x2 : (uint64_t)(uint32_t)opt.max_occ;
    for (uint64_t k = 0; k < m; ++k) ranks[o++] = (x0 + k) | tag;
  }
}
