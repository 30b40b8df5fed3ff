/* TEST INFRASTRUCTURE ONLY -- never linked into the product.
 *
 * Flat, ctypes-friendly entry points into the UNMODIFIED reference (lib/aln) so that tests
 * and tools/make_golden.py can dump kernel-level golden vectors.  memchain.c is #included
 * where it lies under /root/reference (it is left out of the object list of
 * libbiscuit_ref.so) so that its `static` functions (mem_collect_intv, memchain.c:50) are
 * reachable.  Nothing here restates reference logic: every function only marshals. */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "memchain.c" /* resolved through -I/root/reference/lib/aln */
#include "bwa.h"
#include "mem_alnreg.h"

typedef struct {
  bwaidx_t *idx;
  mem_opt_t *opt;
} refp_t;

refp_t *refp_open(const char *prefix) {
  refp_t *h = calloc(1, sizeof(refp_t));
  bwa_verbose = 1;
  h->idx = bwa_idx_load(prefix, BWA_IDX_ALL);
  if (!h->idx) { free(h); return 0; }
  h->opt = mem_opt_init();
  h->opt->flag |= MEM_F_PE;
  return h;
}

void refp_close(refp_t *h) {
  if (!h) return;
  bwa_idx_destroy(h->idx);
  free(h->opt);
  free(h);
}

mem_opt_t *refp_opt(refp_t *h) { return h->opt; }

/* index facts: out[0]=l_pac out[1]=n_seqs out[2]=primary[0] out[3]=primary[1]
 * out[4]=seq_len out[5..9]=L2 of bwt[0]  out[10..14]=L2 of bwt[1] */
void refp_index_info(refp_t *h, int64_t *out) {
  int i;
  out[0] = h->idx->bns->l_pac; out[1] = h->idx->bns->n_seqs;
  out[2] = h->idx->bwt[0].primary; out[3] = h->idx->bwt[1].primary;
  out[4] = h->idx->bwt[0].seq_len;
  for (i = 0; i < 5; ++i) out[5 + i] = h->idx->bwt[0].L2[i], out[10 + i] = h->idx->bwt[1].L2[i];
}

/* bwt_occ4 (bwt.c:173) on bwt[which] */
void refp_occ4(refp_t *h, int which, int n, const int64_t *k, uint64_t *cnt) {
  int i;
  for (i = 0; i < n; ++i) bwt_occ4(&h->idx->bwt[which], (bwtint_t)k[i], cnt + 4 * i);
}

/* bwt_sa (bwt.c:87) on bwt[which] */
void refp_sa(refp_t *h, int which, int n, const uint64_t *k, uint64_t *pos) {
  int i;
  for (i = 0; i < n; ++i) pos[i] = bwt_sa(&h->idx->bwt[which], k[i]);
}

/* bwt_smem1a (bwt.c:307).  q is the already converted read.  out: 4 u64 per interval. */
int refp_smem1a(refp_t *h, int parent, int len, const uint8_t *q, int x, int min_intv,
                uint64_t max_intv, uint64_t *out, int cap, int *n_out) {
  bwtintv_v mem = {0, 0, 0};
  int ret = bwt_smem1a(&h->idx->bwt[parent], &h->idx->bwt[!parent], len, q, x, min_intv, max_intv, &mem, 0);
  size_t i;
  *n_out = (int)mem.n;
  for (i = 0; i < mem.n && (int)i < cap; ++i) memcpy(out + 4 * i, &mem.a[i], 32);
  free(mem.a);
  return ret;
}

/* bwt_seed_strategy1 (bwt.c:376) */
int refp_seed_strategy1(refp_t *h, int parent, int len, const uint8_t *q, int x, int min_len,
                        int max_intv, uint64_t *out) {
  bwtintv_t m;
  int ret = bwt_seed_strategy1(&h->idx->bwt[parent], &h->idx->bwt[!parent], len, q, x, min_len, max_intv, &m);
  memcpy(out, &m, 32);
  return ret;
}

static void bsconv(int len, const uint8_t *seq, int parent, uint8_t *out) {
  int i; /* same mapping as bseq_bsconvert, bwamem.c:161-178 */
  for (i = 0; i < len; ++i) out[i] = parent ? (seq[i] == 1 ? 3 : seq[i]) : (seq[i] == 2 ? 0 : seq[i]);
}

/* mem_collect_intv (memchain.c:50-106) on the converted read; seq is the UNCONVERTED nt4 read */
int refp_collect_intv(refp_t *h, int parent, int len, const uint8_t *seq, uint64_t *out, int cap) {
  bwtintv_cache_t *c = bwtintv_cache_init();
  uint8_t *bis = malloc(len ? len : 1);
  size_t i; int n;
  bsconv(len, seq, parent, bis);
  mem_collect_intv(h->opt, &h->idx->bwt[parent], &h->idx->bwt[!parent], len, bis, c);
  n = (int)c->mem.n;
  for (i = 0; i < c->mem.n && (int)i < cap; ++i) memcpy(out + 4 * i, &c->mem.a[i], 32);
  bwtintv_cache_destroy(c);
  free(bis);
  return n;
}

/* ksw_extend2 (ksw.c:380); out = {score,qle,tle,gtle,gscore,max_off} */
void refp_extend2(int qlen, const uint8_t *query, int tlen, const uint8_t *target, const int8_t *mat,
                  int o_del, int e_del, int o_ins, int e_ins, int w, int end_bonus, int zdrop, int h0, int *out) {
  out[0] = ksw_extend2(qlen, query, tlen, target, 5, mat, o_del, e_del, o_ins, e_ins, w, end_bonus, zdrop, h0,
                       &out[1], &out[2], &out[3], &out[4], &out[5]);
}

/* ksw_global2 (ksw.c:504); returns score, cigar copied to out (cap words) */
int refp_global2(int qlen, const uint8_t *query, int tlen, const uint8_t *target, const int8_t *mat,
                 int o_del, int e_del, int o_ins, int e_ins, int w, int *n_cigar, uint32_t *cig, int cap) {
  uint32_t *c = 0; int i;
  int sc = ksw_global2(qlen, query, tlen, target, 5, mat, o_del, e_del, o_ins, e_ins, w, n_cigar, &c);
  for (i = 0; i < *n_cigar && i < cap; ++i) cig[i] = c[i];
  free(c);
  return sc;
}

/* bns_fetch_seq (bntseq.c:428): returns length, writes rid and clipped [beg,end) */
int refp_fetch_seq(refp_t *h, int64_t *beg, int64_t mid, int64_t *end, int *rid, uint8_t *out, int cap) {
  uint8_t *s = bns_fetch_seq(h->idx->bns, h->idx->pac, beg, mid, end, rid);
  int n = (int)(*end - *beg);
  memcpy(out, s, n < cap ? n : cap);
  free(s);
  return n;
}

static bseq1_t mk_bseq(int len, const uint8_t *seq, uint8_t *buf) {
  bseq1_t b; memset(&b, 0, sizeof b);
  memcpy(buf, seq, len);
  b.l_seq = len; b.seq = buf; b.name = (char *)"probe";
  return b;
}

/* Flattened chain dump.  Per chain: {pos, rid, w, kept, first, is_alt, n_seeds, n_extra} (8 x i64),
 * then n_seeds + n_extra seeds of {rbeg, qbeg, len, score} (4 x i64).
 * stage 0 = after mem_chain (memchain.c:268), 1 = after mem_chain_flt (memchain.c:406). */
static int64_t dump_chains(const mem_chain_v *ch, int64_t *out, int64_t cap) {
  int64_t o = 0; size_t i, j;
  for (i = 0; i < ch->n; ++i) {
    const mem_chain_t *c = &ch->a[i];
    if (o + 8 + 4 * (int64_t)(c->seeds.n + c->seeds_extra.n) > cap) return -1;
    out[o++] = c->pos; out[o++] = c->rid; out[o++] = c->w; out[o++] = c->kept; out[o++] = c->first;
    out[o++] = c->is_alt; out[o++] = c->seeds.n; out[o++] = c->seeds_extra.n;
    for (j = 0; j < c->seeds.n; ++j) {
      out[o++] = c->seeds.a[j].rbeg; out[o++] = c->seeds.a[j].qbeg; out[o++] = c->seeds.a[j].len; out[o++] = c->seeds.a[j].score;
    }
    for (j = 0; j < c->seeds_extra.n; ++j) {
      out[o++] = c->seeds_extra.a[j].rbeg; out[o++] = c->seeds_extra.a[j].qbeg; out[o++] = c->seeds_extra.a[j].len; out[o++] = c->seeds_extra.a[j].score;
    }
  }
  return o;
}

int64_t refp_chain(refp_t *h, int parent, int len, const uint8_t *seq, int stage, int *n_chains,
                   float *frac_rep, int64_t *out, int64_t cap) {
  uint8_t *buf = malloc(len + 1), *bis = malloc(len + 1);
  bseq1_t b = mk_bseq(len, seq, buf);
  mem_chain_v ch; int64_t o;
  bsconv(len, seq, parent, bis);
  b.bisseq[parent] = bis;
  ch = mem_chain(h->opt, h->idx->bwt, h->idx->bns, &b, 0, (uint8_t)parent);
  if (stage >= 1) mem_chain_flt(h->opt, &ch);
  *n_chains = (int)ch.n;
  *frac_rep = ch.n ? ch.a[0].frac_rep : 0.f;
  o = dump_chains(&ch, out, cap);
  free_mem_chain_v(ch);
  free(buf); free(bis);
  return o;
}

/* Region dump: 16 x i64 per region:
 * rb re qb qe rid score truesc sub csub sub_n w seedcov seedlen0 secondary bss|parent<<1 frac_rep(bits) */
static void dump_reg(const mem_alnreg_t *r, int64_t *o) {
  uint32_t fb; memcpy(&fb, &r->frac_rep, 4);
  o[0] = r->rb; o[1] = r->re; o[2] = r->qb; o[3] = r->qe; o[4] = r->rid; o[5] = r->score; o[6] = r->truesc;
  o[7] = r->sub; o[8] = r->csub; o[9] = r->sub_n; o[10] = r->w; o[11] = r->seedcov; o[12] = r->seedlen0;
  o[13] = r->secondary; o[14] = r->bss | r->parent << 1; o[15] = fb;
}

/* One mem_align1_core-equivalent (bwamem.c:183-208) for one read and one conversion:
 * mem_chain -> mem_chain_flt -> mem_flt_chained_seeds -> mem_chain2region. Regions are
 * returned BEFORE mem_merge_regions. */
int refp_align1(refp_t *h, int parent, int len, const uint8_t *seq, int64_t *out, int cap_regs) {
  uint8_t *buf = malloc(len + 1), *bis = malloc(len + 1);
  bseq1_t b = mk_bseq(len, seq, buf);
  mem_alnreg_v regs; mem_chain_v ch; size_t i; int n;
  kv_init(regs); regs.n_pri = 0;
  bsconv(len, seq, parent, bis);
  b.bisseq[parent] = bis;
  ch = mem_chain(h->opt, h->idx->bwt, h->idx->bns, &b, 0, (uint8_t)parent);
  mem_chain_flt(h->opt, &ch);
  mem_flt_chained_seeds(h->opt, h->idx->bns, h->idx->pac, &b, &ch, (uint8_t)parent);
  mem_chain2region(h->opt, h->idx->bns, h->idx->pac, &b, (uint8_t)parent, &ch, &regs);
  free_mem_chain_v(ch);
  n = (int)regs.n;
  for (i = 0; i < regs.n && (int)i < cap_regs; ++i) dump_reg(&regs.a[i], out + 16 * i);
  free(regs.a); free(buf); free(bis);
  return n;
}

/* bis_worker1-equivalent for one read (bwamem.c:311-375): conversions in the reference's
 * order (read 1: parent then daughter; read 2: daughter then parent), then
 * mem_merge_regions (mem_alnreg.c).  which_read = 0/1. */
int refp_worker1(refp_t *h, int which_read, int len, const uint8_t *seq, int64_t *out, int cap_regs) {
  uint8_t *buf = malloc(len + 1);
  bseq1_t b = mk_bseq(len, seq, buf);
  mem_alnreg_v regs; size_t i; int n, t;
  kv_init(regs); regs.n_pri = 0;
  for (t = 0; t < 2; ++t) {
    int parent = which_read == 0 ? (t == 0) : (t == 1);
    mem_chain_v ch;
    bseq_bsconvert(&b, (uint8_t)parent);
    ch = mem_chain(h->opt, h->idx->bwt, h->idx->bns, &b, 0, (uint8_t)parent);
    mem_chain_flt(h->opt, &ch);
    mem_flt_chained_seeds(h->opt, h->idx->bns, h->idx->pac, &b, &ch, (uint8_t)parent);
    mem_chain2region(h->opt, h->idx->bns, h->idx->pac, &b, (uint8_t)parent, &ch, &regs);
    free_mem_chain_v(ch);
  }
  mem_merge_regions(h->opt, h->idx->bns, h->idx->pac, &b, &regs);
  n = (int)regs.n;
  for (i = 0; i < regs.n && (int)i < cap_regs; ++i) dump_reg(&regs.a[i], out + 16 * i);
  free(regs.a); free(buf); free(b.bisseq[0]); free(b.bisseq[1]);
  return n;
}

/* mem_process_seqs (bwamem.c:432-476), the reference's batch API, on n
 * interleaved reads given as nt4 rows.  Returns total SAM bytes; if sam_out != NULL the SAM text
 * of all reads is concatenated into it (cap bytes).  Timed by the caller. */
int64_t refp_process_seqs(refp_t *h, int n_threads, int64_t n_processed, int n, const uint8_t *seqs, int stride,
                          const int32_t *lens, const uint8_t *quals, char *sam_out, int64_t cap) {
  bseq1_t *bs = calloc(n, sizeof(bseq1_t));
  int i; int64_t tot = 0;
  char nm[64];
  for (i = 0; i < n; ++i) {
    bs[i].l_seq = lens[i];
    bs[i].seq = malloc(lens[i] + 1); memcpy(bs[i].seq, seqs + (int64_t)i * stride, lens[i]);
    bs[i].qual = malloc(lens[i] + 1);
    if (quals) memcpy(bs[i].qual, quals + (int64_t)i * stride, lens[i]); else memset(bs[i].qual, 'I', lens[i]);
    bs[i].qual[lens[i]] = 0;
    snprintf(nm, sizeof nm, "r%lld", (long long)((n_processed + i) >> 1));
    bs[i].name = strdup(nm);
    bs[i].id = i;
  }
  h->opt->n_threads = n_threads;
  mem_process_seqs(h->opt, h->idx->bwt, h->idx->bns, h->idx->pac, n_processed, n, bs, 0);
  for (i = 0; i < n; ++i) {
    int64_t l = bs[i].sam ? (int64_t)strlen(bs[i].sam) : 0;
    if (sam_out && tot + l < cap) memcpy(sam_out + tot, bs[i].sam, l);
    tot += l;
    free(bs[i].sam); free(bs[i].seq0 ? bs[i].seq0 : bs[i].seq); free(bs[i].qual); free(bs[i].name);
    free(bs[i].bisseq[0]); free(bs[i].bisseq[1]);
  }
  if (sam_out && tot < cap) sam_out[tot] = 0;
  free(bs);
  return tot;
}

/* bis_bwa_gen_cigar2 (bwa.c:290) on the loaded index; out = {score, n_cigar, NM, ZC, ZR, bss_u}; the CIGAR words
 * go to cig (cap words), the MD text (stored behind the CIGAR by the reference) to md (cap_md bytes).
 * Returns 1 when a CIGAR came back, 0 otherwise.  The query is modified and restored by the callee. */
int refp_gen_cigar2(refp_t *h, const int8_t *mat, int o_del, int e_del, int o_ins, int e_ins, int w, int l_query, uint8_t *query,
                    int64_t rb, int64_t re, uint8_t parent, int *out, uint32_t *cig, int cap, char *md, int cap_md) {
  int score = 0, n_cigar = 0, NM = -1, bss_u = 0, i;
  uint32_t ZC = 0, ZR = 0;
  uint32_t *c = bis_bwa_gen_cigar2(mat, o_del, e_del, o_ins, e_ins, w, h->idx->bns->l_pac, h->idx->pac, l_query, query, rb, re, &score,
                                   &n_cigar, &NM, &ZC, &ZR, &bss_u, parent);
  out[0] = score; out[1] = n_cigar; out[2] = NM; out[3] = (int)ZC; out[4] = (int)ZR; out[5] = bss_u;
  if (!c) return 0;
  for (i = 0; i < n_cigar && i < cap; ++i) cig[i] = c[i];
  strncpy(md, (char *)(c + n_cigar), cap_md - 1);
  md[cap_md - 1] = 0;
  free(c);
  return 1;
}

/* ksw_align2 (ksw.c:343); out = {score, te, qe, score2, te2, tb, qb}.  query / target are modified and restored by the callee. */
void refp_align2(int qlen, uint8_t *query, int tlen, uint8_t *target, const int8_t *mat, int o_del, int e_del, int o_ins, int e_ins,
                 int xtra, int *out) {
  kswr_t r = ksw_align2(qlen, query, tlen, target, 5, mat, o_del, e_del, o_ins, e_ins, xtra, 0);
  out[0] = r.score; out[1] = r.te; out[2] = r.qe; out[3] = r.score2; out[4] = r.te2; out[5] = r.tb; out[6] = r.qb;
}

/* bns_get_seq (bntseq.c:402): bases [beg,end) in forward-reverse coordinates */
int refp_get_seq(refp_t *h, int64_t beg, int64_t end, uint8_t *out, int cap) {
  int64_t len = 0;
  uint8_t *s = bns_get_seq(h->idx->bns->l_pac, h->idx->pac, beg, end, &len);
  memcpy(out, s, len < cap ? len : cap);
  free(s);
  return (int)len;
}
