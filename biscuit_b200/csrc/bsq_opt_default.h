// bsq_opt_default(): the defaults of mem_opt_init (lib/aln/bwamem.c:77-128) and the three scoring
// matrices of bwa_fill_scmat_ct / _ga (lib/aln/bwa.c:158-182) for the fields that reach the GPU.
#pragma once
#include <string.h>

// mat[ref*5 + read]; N scores -1.  C>T matrix: a read T against a reference C is a match;
// G>A matrix: a read A against a reference G is a match.
static void bsq_fill_bsmat(int a, int b, int ct, int8_t mat[25]) {
  int k = 0;
  for (int i = 0; i < 4; ++i) {
    for (int j = 0; j < 4; ++j) mat[k++] = i == j ? a : -b;
    mat[k++] = -1;
  }
  for (int j = 0; j < 5; ++j) mat[k++] = -1;
  if (ct) mat[1 * 5 + 3] = a; else mat[2 * 5 + 0] = a;
}

extern "C" void bsq_opt_default(bsq_opt *o) {
  memset(o, 0, sizeof *o);
  o->a = 1; o->b = 2; o->o_del = o->o_ins = 6; o->e_del = o->e_ins = 1;
  o->pen_clip5 = o->pen_clip3 = 10; o->w = 100; o->zdrop = 100;
  o->min_seed_len = 19; o->split_width = 10; o->max_occ = 500; o->max_chain_gap = 10000;
  o->min_chain_weight = 0; o->max_chain_extend = 1 << 30; o->max_mem_intv = 20;
  o->split_len = (int)(19 * 1.5f + .499);  // (int)(min_seed_len * split_factor + .499), memchain.c:55
  o->self_ovlp = 0; o->bsstrand = 0;
  o->mask_level = 0.50f; o->drop_ratio = 0.50f;
  bsq_fill_bsmat(o->a, o->b, 1, o->ctmat);
  bsq_fill_bsmat(o->a, o->b, 0, o->gamat);
}
