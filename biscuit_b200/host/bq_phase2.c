/* bq_phase2.c -- what happens to the alignment regions after the GPU: region merging, insert-size
 * statistics, mate rescue, primary marking, pairing, mapQ, SAM records.  Restated from
 * lib/aln/mem_alnreg.c, mem_pair.c, mem_alnreg_format.c and bwamem.c (line references at each function);
 * all libm use of the aligner lives here. */
#include <limits.h>
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include "bq.h"
#include "bq_sort.h"
#include <pthread.h>
#include <time.h>
/* BQ_PROF=1: thread-seconds per part of phase 2, printed per batch (diagnostics) */
static int g_prof = -1;
static double g_t_mate, g_t_mark, g_t_sam, g_t_pair, g_t_setsam, g_t_fmt;
static pthread_mutex_t g_prof_mu = PTHREAD_MUTEX_INITIALIZER;
static double bq_now(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + ts.tv_nsec * 1e-9; }

_Static_assert(sizeof(bq_reg_t) == 128, "bq_reg_t is laid out to fill two cache lines");
#define MINV(a, b) ((a) < (b) ? (a) : (b))
#define MAXV(a, b) ((a) > (b) ? (a) : (b))

static void regv_push(bq_regv_t *v, const bq_reg_t *r) {
  if (v->n == v->m) {
    const size_t m = v->m ? v->m << 1 : 4;
    if (v->pooled) { /* leave the pool slice / stack array behind */
      bq_reg_t *na = malloc(m * sizeof(bq_reg_t));
      memcpy(na, v->a, v->n * sizeof(bq_reg_t));
      v->a = na; v->pooled = 0;
    } else v->a = realloc(v->a, m * sizeof(bq_reg_t));
    v->m = m;
  }
  v->a[v->n++] = *r;
}

/* mem_alnreg_isize / mem_infer_isize (mem_alnreg.h:74-93): "insert" measured between the two rb of the
 * forward-reverse coordinates -- includes the aligned length of the reverse read */
static int infer_isize(int64_t pos1, int64_t pos2, int isrev1, int isrev2, int64_t len1, int64_t len2, int64_t *isize) {
  if (isrev1 && !isrev2) { *isize = pos1 - pos2 + len1; return 1; }
  if (isrev2 && !isrev1) { *isize = pos2 - pos1 + len2; return 1; }
  return 0;
}
static int reg_isize(const bq_ref_t *ref, const bq_reg_t *r1, const bq_reg_t *r2, int64_t *isize) {
  if (r1->rid != r2->rid) return 0;
  const int isrev1 = r1->rb > ref->l_pac, isrev2 = r2->rb > ref->l_pac;
  const int64_t pos1 = isrev1 ? (ref->l_pac << 1) - 1 - r1->rb : r1->rb, pos2 = isrev2 ? (ref->l_pac << 1) - 1 - r2->rb : r2->rb;
  return infer_isize(pos1, pos2, isrev1, isrev2, r1->qe - r1->qb, r2->qe - r2->qb, isize);
}
static int is_proper_pair(const bq_ref_t *ref, const bq_reg_t *r1, const bq_reg_t *r2, bq_pestat_t pes) {
  int64_t is;
  if (!reg_isize(ref, r1, r2, &is)) return 0;
  return is >= pes.low && is <= pes.high;
}
static int region_depos(const bq_ref_t *ref, const bq_reg_t *reg, int *is_rev) {
  int tmp;
  int64_t rpos = bq_depos(ref, reg->rb < ref->l_pac ? reg->rb : reg->re - 1, is_rev ? is_rev : &tmp);
  return (int)(rpos - ref->anns[reg->rid].offset);
}

/* ---------------- region merging: mem_alnreg.c:63-227 ---------------- */

static int lt_re(const void *a, const void *b) { return ((const bq_reg_t *)a)->re < ((const bq_reg_t *)b)->re; }
static int lt_score_rb_qb(const void *a_, const void *b_) {
  const bq_reg_t *a = a_, *b = b_;
  return a->score > b->score || (a->score == b->score && (a->rb < b->rb || (a->rb == b->rb && a->qb < b->qb)));
}
BQ_INTROSORT_DEFINE(sort_regs_re, bq_reg_t, lt_re)
BQ_INTROSORT_DEFINE(sort_regs_score, bq_reg_t, lt_score_rb_qb)

/* score of joining two colinear regions through a global alignment, 0 when they should stay apart */
static int test_concatenation(const bq_opt_t *opt, const bq_ref_t *ref, uint8_t *query, const bq_reg_t *a, const bq_reg_t *b, int *w_out) {
  if (ref == 0 || query == 0) return 0;
  if (a->rb < ref->l_pac && b->rb >= ref->l_pac) return 0;
  if (a->qb >= b->qb || a->qe >= b->qe || a->re >= b->re) return 0;
  int w = (int)((a->re - b->rb) - (a->qe - b->qb));
  w = w > 0 ? w : -w;
  double r = (double)(a->re - b->rb) / (b->re - a->rb) - (double)(a->qe - b->qb) / (b->qe - a->qb);
  r = r > 0. ? r : -r;
  if (a->re < b->rb || a->qe < b->qb) { if (w > opt->w << 1 || r >= 0.05f) return 0; }
  else if (w > opt->w << 2 || r >= 0.05f * 2) return 0;
  w += a->w + b->w;
  w = MINV(w, opt->w << 2);
  int score;
  bq_gen_cigar(a->parent ? opt->ctmat : opt->gamat, opt->o_del, opt->e_del, opt->o_ins, opt->e_ins, w, ref->l_pac, ref->pac, b->qe - a->qb,
               query + a->qb, a->rb, b->re, &score, 0, 0, 0, 0, 0, a->parent);
  int q_s = (int)((double)(b->qe - a->qb) / ((b->qe - b->qb) + (a->qe - a->qb)) * (b->score + a->score) + .499);
  int r_s = (int)((double)(b->re - a->rb) / ((b->re - b->rb) + (a->re - a->rb)) * (b->score + a->score) + .499);
  if ((double)score / MAXV(q_s, r_s) < 0.90f) return 0;
  *w_out = w;
  return score;
}

static void sort_dedup(const bq_opt_t *opt, const bq_ref_t *ref, uint8_t *query, bq_regv_t *regs) {
  if (regs->n <= 1) return;
  sort_regs_re(regs->a, regs->n);
  int i, m;
  for (i = 1; (size_t)i < regs->n; ++i) {
    bq_reg_t *p = regs->a + i;
    for (int j = i - 1; j >= 0 && p->rid == regs->a[j].rid && p->rb < regs->a[j].re + opt->max_chain_gap; --j) {
      bq_reg_t *q = regs->a + j;
      if (q->qe == q->qb) continue;
      int64_t orr = q->re - p->rb, oq = q->qb < p->qb ? q->qe - p->qb : p->qe - q->qb;
      int64_t mr = MINV(q->re - q->rb, p->re - p->rb), mq = MINV(q->qe - q->qb, p->qe - p->qb);
      int score, w;
      if (orr > opt->mask_level_redun * mr && oq > opt->mask_level_redun * mq) {
        if (p->score < q->score) { p->qe = p->qb; break; }
        else q->qe = q->qb;
      } else if (q->rb < p->rb && (score = test_concatenation(opt, ref, query, q, p, &w)) > 0) {
        p->seedcov = p->seedcov > q->seedcov ? p->seedcov : q->seedcov;
        p->sub = MAXV(p->sub, q->sub);
        p->csub = MAXV(p->csub, q->csub);
        p->truesc = p->score = score;
        p->qb = q->qb; p->rb = q->rb; p->w = w;
        q->qb = q->qe;
      }
    }
  }
  for (i = 0, m = 0; (size_t)i < regs->n; ++i)
    if (regs->a[i].qe > regs->a[i].qb) { if (m != i) regs->a[m++] = regs->a[i]; else ++m; }
  regs->n = (size_t)m;
  sort_regs_score(regs->a, regs->n);
  for (i = 1; (size_t)i < regs->n; ++i)
    if (regs->a[i].score == regs->a[i - 1].score && regs->a[i].rb == regs->a[i - 1].rb && regs->a[i].qb == regs->a[i - 1].qb)
      regs->a[i].qe = regs->a[i].qb;
  for (i = 1, m = 1; (size_t)i < regs->n; ++i)
    if (regs->a[i].qe > regs->a[i].qb) { if (m != i) regs->a[m++] = regs->a[i]; else ++m; }
  regs->n = (size_t)m;
}

void bq_merge_regions(const bq_opt_t *opt, const bq_ref_t *ref, const uint8_t *query, int l_query, bq_regv_t *regs) {
  sort_dedup(opt, ref, (uint8_t *)query, regs);
  if ((opt->flag & BQ_F_SELF_OVLP) && regs->n && regs->a[0].truesc == l_query * opt->a) { /* mem_test_and_remove_exact */
    memmove(regs->a, regs->a + 1, (regs->n - 1) * sizeof(bq_reg_t));
    regs->n--;
  }
  for (size_t i = 0; i < regs->n; ++i)
    if (regs->a[i].rid >= 0 && ref->anns[regs->a[i].rid].is_alt) regs->a[i].is_alt = 1;
}

/* ---------------- insert size statistics: mem_pair.c:42-144 ---------------- */

static int cal_sub(const bq_opt_t *opt, const bq_regv_t *regs) {
  const bq_reg_t *best = &regs->a[0], *p = 0;
  size_t j;
  for (j = 1; j < regs->n; ++j) {
    p = &regs->a[j];
    int b_max = MAXV(p->qb, best->qb), e_min = MINV(p->qe, best->qe);
    if (e_min > b_max) {
      int min_l = MINV(p->qe - p->qb, best->qe - best->qb);
      if (e_min - b_max >= min_l * opt->mask_level) break;
    }
  }
  return j < regs->n ? p->score : opt->min_seed_len * opt->a;
}

static int lt_i64(const void *a, const void *b) { return *(const int64_t *)a < *(const int64_t *)b; }

/* the insert size a pair contributes to the statistics, if any (mem_pair.c:84-100) */
static int pestat_candidate(const bq_opt_t *opt, const bq_ref_t *ref, const bq_regv_t *r0, const bq_regv_t *r1, int64_t *is_out) {
  int64_t is;
  if (r0->n == 0 || r1->n == 0) return 0;
  if (cal_sub(opt, r0) > 0.8 * r0->a[0].score) return 0; /* MIN_RATIO, mem_pair.c:35 */
  if (cal_sub(opt, r1) > 0.8 * r1->a[0].score) return 0;
  if (r0->a[0].rid != r1->a[0].rid) return 0;
  if (r0->a[0].bss != r1->a[0].bss) return 0;
  if (!reg_isize(ref, &r0->a[0], &r1->a[0], &is)) return 0;
  if (!(is <= opt->max_ins && is >= -opt->max_ins)) return 0;
  *is_out = is;
  return 1;
}

static bq_pestat_t pestat_finish(const bq_opt_t *opt, int64_t *isz, size_t n_is);

bq_pestat_t bq_pestat(const bq_opt_t *opt, const bq_ref_t *ref, int n, const bq_regv_t *regs) {
  int64_t *isz = malloc(sizeof(int64_t) * (size_t)(n / 2 + 1));
  size_t n_is = 0;
  for (int i = 0; i < n >> 1; ++i) {
    int64_t is;
    if (pestat_candidate(opt, ref, &regs[i << 1], &regs[i << 1 | 1], &is)) isz[n_is++] = is;
  }
  return pestat_finish(opt, isz, n_is);
}

/* isz: the candidate insert sizes in any order (they are sorted here); freed */
static bq_pestat_t pestat_finish(const bq_opt_t *opt, int64_t *isz, size_t n_is) {
  bq_pestat_t pes;
  memset(&pes, 0, sizeof pes);
  if (bq_verbose >= 3) fprintf(stderr, "[M::mem_pestat] # candidate unique pairs: %ld\n", (long)n_is);
  if (n_is < 10) {
    fprintf(stderr, "[M:mem_pestat] There are not enough pairs for insert size inference\n");
    free(isz);
    pes.failed = 1;
    return pes;
  }
  /* ascending order (ks_introsort_64, mem_pair.c:104).  The keys are plain integers in [-max_ins, max_ins], so any
   * correct sort gives the same array; a counting sort replaces ~1.7 M comparator calls per batch on this serial path */
  if (opt->max_ins > 0 && opt->max_ins <= (1 << 22)) {
    const int64_t lo = -(int64_t)opt->max_ins, range = 2 * (int64_t)opt->max_ins + 1;
    uint32_t *cnt = calloc((size_t)range, sizeof(uint32_t));
    for (size_t k = 0; k < n_is; ++k) ++cnt[isz[k] - lo];
    size_t o = 0;
    for (int64_t v = 0; v < range; ++v)
      for (uint32_t c = cnt[v]; c > 0; --c) isz[o++] = v + lo;
    free(cnt);
  } else bq_introsort(isz, n_is, sizeof(int64_t), lt_i64);
  int p25 = (int)isz[(int)(.25 * n_is + .499)], p50 = (int)isz[(int)(.50 * n_is + .499)], p75 = (int)isz[(int)(.75 * n_is + .499)];
  pes.low = (int)(p25 - 2.0 * (p75 - p25) + .499);
  pes.high = (int)(p75 + 2.0 * (p75 - p25) + .499);
  fprintf(stderr, "[M::mem_pestat] (25, 50, 75) percentile: (%d, %d, %d)\n", p25, p50, p75);
  fprintf(stderr, "[M::mem_pestat] low and high boundaries for computing mean and std.dev: (%d, %d)\n", pes.low, pes.high);
  int x = 0;
  size_t i;
  for (i = 0, pes.avg = 0; i < n_is; ++i)
    if (isz[i] >= pes.low && isz[i] <= pes.high) { pes.avg += isz[i]; ++x; }
  pes.avg /= x;
  for (i = 0, pes.std = 0; i < n_is; ++i)
    if (isz[i] >= pes.low && isz[i] <= pes.high) pes.std += (isz[i] - pes.avg) * (isz[i] - pes.avg);
  pes.std = sqrt(pes.std / x);
  fprintf(stderr, "[M::mem_pestat] mean and std.dev: (%.2f, %.2f)\n", pes.avg, pes.std);
  pes.low = (int)(p25 - 3.0 * (p75 - p25) + .499);
  pes.high = (int)(p75 + 3.0 * (p75 - p25) + .499);
  if (pes.low > pes.avg - 4.0 * pes.std) pes.low = (int)(pes.avg - 4.0 * pes.std + .499);
  if (pes.high < pes.avg + 4.0 * pes.std) pes.high = (int)(pes.avg + 4.0 * pes.std + .499);
  fprintf(stderr, "[M::mem_pestat] low and high boundaries for proper pairs: (%d, %d)\n", pes.low, pes.high);
  free(isz);
  return pes;
}

/* ---------------- mate rescue: mem_alnreg.c:395-513 ---------------- */

/* counters of the batched DP (bq_dp_stats) */
static int64_t g_dp_stat[6];
void bq_dp_stats(int64_t out[6]) { for (int i = 0; i < 6; ++i) out[i] = __atomic_load_n(&g_dp_stat[i], __ATOMIC_RELAXED); }

/* does the mate already have a hit at a proper distance from reg?  (mem_alnreg.c:401-407) */
static int mate_in_range(const bq_ref_t *ref, bq_pestat_t pes, const bq_reg_t *reg, const bq_regv_t *mregs) {
  for (size_t i = 0; i < mregs->n; ++i) {
    int64_t is;
    if (reg_isize(ref, reg, &mregs->a[i], &is) && is >= pes.low && is <= pes.high) return 1;
  }
  return 0;
}

/* the reference window searched for the mate of reg (mem_alnreg.c:418-427): bns_fetch_seq's clipping to the contig of the
 * window's middle (bntseq.c:428-452) without fetching the bases; 0 when there is nothing to search */
static int matesw_window(const bq_opt_t *opt, const bq_ref_t *ref, bq_pestat_t pes, const bq_reg_t *reg, int l_ms, int64_t *rb_, int64_t *re_) {
  const int64_t l_pac = ref->l_pac;
  int64_t rb = MAXV(0, reg->rb + pes.low - l_ms), re = MINV(l_pac << 1, reg->rb + pes.high);
  if (rb >= re) return 0;
  int is_rev;
  const int rid = bq_pos2rid(ref, bq_depos(ref, (rb + re) >> 1, &is_rev));
  int64_t far_beg = ref->anns[rid].offset, far_end = far_beg + ref->anns[rid].len;
  if (is_rev) { int64_t t = far_beg; far_beg = (l_pac << 1) - far_end; far_end = (l_pac << 1) - t; }
  if (rb < far_beg) rb = far_beg;
  if (re > far_end) re = far_end;
  if (reg->rid != rid || re - rb < opt->min_seed_len) return 0;
  *rb_ = rb; *re_ = re;
  return 1;
}
static inline int matesw_xtra(const bq_opt_t *opt, int l_ms) { return BQ_XSUBO | BQ_XSTART | (l_ms * opt->a < 250 ? BQ_XBYTE : 0) | (opt->min_seed_len * opt->a); }

/* pre: the local alignment of this call computed on the GPU (bsq_dp_matesw), or NULL: computed here */
static void matesw_core(const bq_opt_t *opt, const bq_ref_t *ref, bq_pestat_t pes, const bq_reg_t *reg, int l_ms, const uint8_t *ms,
                        bq_regv_t *mregs, const bsq_matesw_res *pre) {
  const int64_t l_pac = ref->l_pac;
  int i;
  if (mate_in_range(ref, pes, reg, mregs)) return;
  int64_t rb, re;
  if (!matesw_window(opt, ref, pes, reg, l_ms, &rb, &re)) return;
  const uint8_t parent = reg->bss ^ (reg->rb < l_pac);
  bq_swr_t aln;
  if (pre && !pre->pad_) {
    aln.score = pre->score; aln.te = pre->te; aln.qe = pre->qe; aln.score2 = pre->score2; aln.te2 = pre->te2; aln.tb = pre->tb; aln.qb = pre->qb;
    __atomic_fetch_add(&g_dp_stat[4], 1, __ATOMIC_RELAXED);
  } else {
    uint8_t revbuf[512], *rev = l_ms < (int)sizeof revbuf ? revbuf : malloc((size_t)l_ms + 1);
    for (i = 0; i < l_ms; ++i) rev[l_ms - 1 - i] = ms[i] < 4 ? 3 - ms[i] : 4;
    int rid = -1;
    uint8_t *rseq = bq_fetch_seq(ref, &rb, (rb + re) >> 1, &re, &rid);
    aln = bq_local_align(l_ms, rev, (int)(re - rb), rseq, parent ? opt->gamat : opt->ctmat, opt->o_del, opt->e_del, opt->o_ins, opt->e_ins,
                         matesw_xtra(opt, l_ms));
    if (rev != revbuf) free(rev);
    free(rseq);
    __atomic_fetch_add(&g_dp_stat[5], 1, __ATOMIC_RELAXED);
  }
  if (aln.score >= opt->min_seed_len && aln.qb >= 0) {
    bq_reg_t b;
    memset(&b, 0, sizeof b);
    b.rid = reg->rid; b.is_alt = reg->is_alt;
    b.qb = l_ms - (aln.qe + 1); b.qe = l_ms - aln.qb;
    b.rb = (l_pac << 1) - (rb + aln.te + 1); b.re = (l_pac << 1) - (rb + aln.tb);
    b.score = aln.score; b.csub = aln.score2; b.secondary = -1;
    b.seedcov = (int)(MINV(b.re - b.rb, (int64_t)(b.qe - b.qb)) >> 1);
    b.bss = reg->bss; b.parent = 1 - parent;
    regv_push(mregs, &b);
    for (i = 0; (size_t)i < mregs->n - 1; ++i)
      if (mregs->a[i].score < b.score) break;
    int at = i;
    for (i = (int)mregs->n - 1; i > at; --i) mregs->a[i] = mregs->a[i - 1];
    mregs->a[i] = b;
    sort_dedup(opt, 0, 0, mregs);
  }
}

/* The rescue attempts of one pair in the reference's order (mem_alnreg.c:495-513).  Three uses:
 *   jobs == NULL, pre == NULL  the reference's loop, every local alignment computed on the host (bq_matesw);
 *   jobs != NULL               planning: nothing is changed; every attempt that the state of the pair BEFORE any rescue
 *                              does not rule out is written as one bsq_dp_matesw job (row[] = device rows of the two reads)
 *                              with its (end, rank) key; returns the number of jobs (jobs may be a counting dummy, n_max 0);
 *   pre != NULL                replay with the results of those jobs.  An attempt is ruled out dynamically when the mate
 *                              gained a hit in range from an earlier rescue, so the planned set is a superset of what is
 *                              used -- except when region de-duplication removed the hit that ruled an attempt out at
 *                              planning time; such an attempt finds no job and is computed on the host. */
typedef struct { const bsq_matesw_res *res; const uint32_t *key; int64_t cur, end; } ms_pre_t;
static int matesw_pair(const bq_opt_t *opt, const bq_ref_t *ref, bq_pestat_t pes, bq_read_t s[2], bq_regv_t regs[2], ms_pre_t *pre,
                       bsq_matesw_job *jobs, uint32_t *keys, const int64_t row[2], int fill) {
  int i, n_jobs = 0;
  size_t j;
  const int plan = row != 0;
  if (plan) { /* nothing changes while planning: the candidates are read in place (rank j = position among the good ones) */
    for (i = 0; i < 2; ++i) {
      size_t rank = 0;
      for (j = 0; j < regs[i].n && (int)rank < opt->max_matesw; ++j) {
        const bq_reg_t *reg = &regs[i].a[j];
        if (reg->score < regs[i].a[0].score - opt->pen_unpaired) continue;
        const size_t jr = rank++;
        int64_t rb, re;
        if (mate_in_range(ref, pes, reg, &regs[!i]) || !matesw_window(opt, ref, pes, reg, s[!i].l_seq, &rb, &re)) continue;
        if (fill) {
          bsq_matesw_job *jb = &jobs[n_jobs];
          memset(jb, 0, sizeof *jb);
          jb->rb = rb; jb->re = re; jb->row = (int32_t)row[!i]; jb->xtra = matesw_xtra(opt, s[!i].l_seq);
          jb->use_ga = (uint8_t)(reg->bss ^ (reg->rb < ref->l_pac));
          keys[n_jobs] = (uint32_t)i << 16 | (uint32_t)jr;
        }
        ++n_jobs;
      }
    }
    return n_jobs;
  }
  bq_reg_t gbuf[2][8];
  bq_regv_t good[2] = {{0, 8, 0, gbuf[0], 1}, {0, 8, 0, gbuf[1], 1}};
  for (i = 0; i < 2; ++i)
    for (j = 0; j < regs[i].n; ++j)
      if (regs[i].a[j].score >= regs[i].a[0].score - opt->pen_unpaired) regv_push(&good[i], &regs[i].a[j]);
  for (i = 0; i < 2; ++i)
    for (j = 0; j < good[i].n && (int)j < opt->max_matesw; ++j) {
      const bq_reg_t *reg = &good[i].a[j];
      const bsq_matesw_res *r = 0;
      if (pre) {
        const uint32_t key = (uint32_t)i << 16 | (uint32_t)j;
        while (pre->cur < pre->end && pre->key[pre->cur] < key) ++pre->cur; /* jobs of attempts that were skipped */
        if (pre->cur < pre->end && pre->key[pre->cur] == key) r = &pre->res[pre->cur++];
      }
      matesw_core(opt, ref, pes, reg, s[!i].l_seq, s[!i].seq, &regs[!i], r);
    }
  for (i = 0; i < 2; ++i) if (!good[i].pooled) free(good[i].a);
  return n_jobs;
}

void bq_matesw(const bq_opt_t *opt, const bq_ref_t *ref, bq_pestat_t pes, bq_read_t s[2], bq_regv_t regs[2]) {
  matesw_pair(opt, ref, pes, s, regs, 0, 0, 0, 0, 0);
}

/* ---------------- primary marking: mem_alnreg.c:242-380 ---------------- */

static int lt_hash(const void *a_, const void *b_) {
  const bq_reg_t *a = a_, *b = b_;
  return a->score > b->score || (a->score == b->score && (a->is_alt < b->is_alt || (a->is_alt == b->is_alt && a->hash < b->hash)));
}
static int lt_hash2(const void *a_, const void *b_) {
  const bq_reg_t *a = a_, *b = b_;
  return a->is_alt < b->is_alt || (a->is_alt == b->is_alt && (a->score > b->score || (a->score == b->score && a->hash < b->hash)));
}
BQ_INTROSORT_DEFINE(sort_regs_hash, bq_reg_t, lt_hash)
BQ_INTROSORT_DEFINE(sort_regs_hash2, bq_reg_t, lt_hash2)

static void mark_core(const bq_opt_t *opt, int n_mark, bq_regv_t *regs, int *z, int *nz) {
  int tmp = opt->a + opt->b;
  tmp = MAXV(opt->o_del + opt->e_del, tmp);
  tmp = MAXV(opt->o_ins + opt->e_ins, tmp);
  *nz = 0;
  z[(*nz)++] = 0;
  for (int i = 1; i < n_mark; ++i) {
    bq_reg_t *a = regs->a + i;
    int k;
    for (k = 0; k < *nz; ++k) {
      bq_reg_t *b = regs->a + z[k];
      int b_max = MAXV(a->qb, b->qb), e_min = MINV(a->qe, b->qe);
      if (e_min > b_max) {
        int min_l = MINV(a->qe - a->qb, b->qe - b->qb);
        if (e_min - b_max >= min_l * opt->mask_level) {
          if (b->sub == 0) b->sub = a->score;
          if (b->score - a->score <= tmp && (b->is_alt || !a->is_alt)) ++b->sub_n;
          break;
        }
      }
    }
    if (k == *nz) z[(*nz)++] = i;
    else a->secondary = z[k];
  }
}

void bq_mark_primary(const bq_opt_t *opt, bq_regv_t *regs, int64_t id) {
  regs->n_pri = 0;
  if (regs->n == 0) return;
  int i, nz;
  for (i = 0; (size_t)i < regs->n; ++i) {
    bq_reg_t *p = regs->a + i;
    p->sub = p->alt_sc = 0;
    p->secondary = p->secondary_all = -1;
    p->hash = bq_hash64((uint64_t)(id + i));
    if (!p->is_alt) ++regs->n_pri;
  }
  sort_regs_hash(regs->a, regs->n);
  int zbuf[64], *z = regs->n < 64 ? zbuf : malloc(sizeof(int) * (regs->n + 1));
  mark_core(opt, (int)regs->n, regs, z, &nz);
  for (i = 0; (size_t)i < regs->n; ++i) {
    bq_reg_t *p = regs->a + i;
    p->secondary_all = i;
    if (!p->is_alt && p->secondary >= 0 && regs->a[p->secondary].is_alt) p->alt_sc = regs->a[p->secondary].score;
  }
  if (regs->n_pri > 0 && regs->n_pri < regs->n) {
    sort_regs_hash2(regs->a, regs->n);
    for (i = 0; (size_t)i < regs->n; ++i) z[regs->a[i].secondary_all] = i;
    for (i = 0; (size_t)i < regs->n; ++i) {
      if (regs->a[i].secondary >= 0) {
        regs->a[i].secondary_all = z[regs->a[i].secondary];
        if (regs->a[i].is_alt) regs->a[i].secondary = INT_MAX;
      } else regs->a[i].secondary_all = -1;
    }
    for (i = 0; (size_t)i < regs->n_pri; ++i) { regs->a[i].sub = 0; regs->a[i].secondary = -1; }
    mark_core(opt, (int)regs->n_pri, regs, z, &nz);
  } else
    for (i = 0; (size_t)i < regs->n; ++i) regs->a[i].secondary_all = regs->a[i].secondary;
  if (z != zbuf) free(z);
}

/* ---------------- mapQ: bwamem.c:134-157 ---------------- */

int bq_approx_mapq_se(const bq_opt_t *opt, const bq_reg_t *a) {
  int mapq, l, sub = a->sub ? a->sub : opt->min_seed_len * opt->a;
  double identity;
  sub = a->csub > sub ? a->csub : sub;
  if (sub >= a->score) return 0;
  l = a->qe - a->qb > a->re - a->rb ? a->qe - a->qb : (int)(a->re - a->rb);
  identity = 1. - (double)(l * opt->a - a->score) / (opt->a + opt->b) / l;
  if (a->score == 0) mapq = 0;
  else if (opt->mapQ_coef_len > 0) {
    double tmp = l < opt->mapQ_coef_len ? 1. : opt->mapQ_coef_fac / log(l);
    tmp *= identity * identity;
    mapq = (int)(6.02 * (a->score - sub) / opt->a * tmp * tmp + .499);
  } else {
    mapq = (int)(30.0 * (1. - (double)sub / a->score) * log(a->seedcov) + .499);
    mapq = identity < 0.95 ? (int)(mapq * identity * identity + .499) : mapq;
  }
  if (a->sub_n > 0) mapq -= (int)(4.343 * log(a->sub_n + 1) + .499);
  if (mapq > 60) mapq = 60;
  if (mapq < 0) mapq = 0;
  mapq = (int)(mapq * (1. - a->frac_rep) + .499);
  return mapq;
}

/* ---------------- pairing: mem_pair.c:147-270 ---------------- */

typedef struct { uint64_t x, y, z; } trio_t;
typedef struct { uint64_t x, y; } pair_t;
static int lt_xy3(const void *a_, const void *b_) { const trio_t *a = a_, *b = b_; return a->x < b->x || (a->x == b->x && a->y < b->y); }
static int lt_xy2(const void *a_, const void *b_) { const pair_t *a = a_, *b = b_; return a->x < b->x || (a->x == b->x && a->y < b->y); }
BQ_INTROSORT_DEFINE(sort_trio, trio_t, lt_xy3)
BQ_INTROSORT_DEFINE(sort_pair, pair_t, lt_xy2)

/* The insert-size term of the pair score (mem_pair.c:174-176): .721 * log(2 * erfc(|z| / sqrt 2)) * a, a function of the
 * insert size alone once the batch's statistics are known.  Phase 2 tabulates it once per batch over [low, high] (the only
 * values that reach it) instead of two libm calls per candidate pair; the entries are the doubles the expression gives. */
typedef struct { int low, high; const double *term; } pair_tab_t;
static __thread const pair_tab_t *tl_pair_tab;
static inline double pair_term(const bq_opt_t *opt, const bq_pestat_t *pes, int64_t is) {
  const double zscore = (is - pes->avg) / pes->std;
  return .721 * log(2. * erfc(fabs(zscore) * M_SQRT1_2)) * opt->a;
}

static void pair_regs(const bq_opt_t *opt, const bq_ref_t *ref, bq_pestat_t pes, bq_regv_t regs[2], int id, int *score, int *sub, int *n_sub,
                      int z[2]) {
  const int64_t l_pac = ref->l_pac;
  size_t nv = regs[0].n_pri + regs[1].n_pri, n = 0, np = 0, mp = 16;
  trio_t vbuf[64], *v = nv < 64 ? vbuf : malloc(sizeof(trio_t) * (nv + 1));
  pair_t pbuf[16], *pp = pbuf;
  int i, k, r;
  for (r = 0; r < 2; ++r)
    for (i = 0; (size_t)i < regs[r].n_pri; ++i) {
      const bq_reg_t *p = &regs[r].a[i];
      v[n].x = (uint64_t)p->bss << 63 | (uint64_t)p->rid << 32 | (uint64_t)(int64_t)region_depos(ref, p, 0);
      v[n].y = (uint64_t)p->score << 32 | (uint64_t)(int64_t)(i << 2 | (p->rb >= l_pac) << 1 | r);
      v[n].z = (uint64_t)(p->qe - p->qb);
      ++n;
    }
  sort_trio(v, n);
  for (i = 0; (size_t)i < n; ++i)
    for (k = i - 1; k >= 0; --k) {
      if (v[i].x >> 32 != v[k].x >> 32) break;
      if (v[i].x >> 63 != v[k].x >> 63) break;
      if ((int64_t)(v[i].x & 0xffffffffU) - (int64_t)(v[k].x & 0xffffffffU) > MAXV(pes.low, pes.high)) break;
      if ((v[i].y & 1) == (v[k].y & 1)) break;
      int64_t is = 0;
      if (infer_isize((int64_t)v[k].x, (int64_t)v[i].x, (v[k].y >> 1) & 1, (v[i].y >> 1) & 1, (int64_t)v[k].z, (int64_t)v[i].z, &is) &&
          is >= pes.low && is <= pes.high) {
        const pair_tab_t *pt = tl_pair_tab;
        const double term = pt && is >= pt->low && is <= pt->high ? pt->term[is - pt->low] : pair_term(opt, &pes, is);
        int sc = (int)((v[i].y >> 32) + (v[k].y >> 32) + term + .499);
        sc = MAXV(0, sc);
        if (np == mp) {
          mp <<= 1;
          if (pp == pbuf) { pp = malloc(mp * sizeof(pair_t)); memcpy(pp, pbuf, np * sizeof(pair_t)); }
          else pp = realloc(pp, mp * sizeof(pair_t));
        }
        pp[np].y = (uint64_t)k << 32 | (uint64_t)i;
        pp[np].x = (uint64_t)sc << 32 | (bq_hash64(pp[np].y ^ (uint64_t)(int64_t)(id << 8)) & 0xffffffffU);
        ++np;
      }
    }
  if (np) {
    sort_pair(pp, np);
    i = (int)(pp[np - 1].y >> 32);
    k = (int)(pp[np - 1].y << 32 >> 32);
    z[v[i].y & 1] = (int)(v[i].y << 32 >> 34);
    z[v[k].y & 1] = (int)(v[k].y << 32 >> 34);
    *score = (int)(pp[np - 1].x >> 32);
    *sub = np > 1 ? (int)(pp[np - 2].x >> 32) : 0;
    int tmp = opt->a + opt->b;
    tmp = MAXV(tmp, opt->o_del + opt->e_del);
    tmp = MAXV(tmp, opt->o_ins + opt->e_ins);
    *n_sub = 0;
    for (long u = (long)np - 2; u >= 0; --u)
      if (*sub - (int)(pp[u].x >> 32) <= tmp) ++*n_sub;
  } else { *score = 0; *sub = 0; *n_sub = 0; z[0] = z[1] = -1; }
  if (pp != pbuf) free(pp);
  if (v != vbuf) free(v);
}

/* ---------------- SAM: mem_alnreg_format.c ---------------- */

static int infer_bw(int l1, int l2, int score, int a, int q, int r) { /* bwamem.h:192 */
  int w;
  if (l1 == l2 && l1 * a - score < (q + r - a) << 1) return 0;
  w = (int)((double)((l1 < l2 ? l1 : l2) * a - score - q) / r + 2.);
  if (w < abs(l1 - l2)) w = abs(l1 - l2);
  return w;
}
static int get_rlen(int n_cigar, const uint32_t *cigar) {
  int l = 0;
  for (int k = 0; k < n_cigar; ++k) { int op = cigar[k] & 0xf; if (op == 0 || op == 2) l += (int)(cigar[k] >> 4); }
  return l;
}

/* the first band of mem_alnreg_setSAM (:46-51) */
static int set_sam_band(const bq_opt_t *opt, const bq_reg_t *reg) {
  int w1 = infer_bw(reg->qe - reg->qb, (int)(reg->re - reg->rb), reg->truesc, opt->a, opt->o_del, opt->e_del);
  int w2 = infer_bw(reg->qe - reg->qb, (int)(reg->re - reg->rb), reg->truesc, opt->a, opt->o_ins, opt->e_ins);
  int w = MAXV(w1, w2);
  if (w > opt->w) w = MINV(w, reg->w);
  return w;
}

/* CIGARs of the batch computed on the GPU (bsq_dp_cigar): set by the phase-2 workers around reg2sam */
typedef struct { const bsq_cigar_res *res; const uint32_t *blob; int64_t n; const int64_t *thr_base; int n_thr; } dp_cig_t;
/* result of the job a region carries, or NULL */
static inline const bsq_cigar_res *dp_cig_find(const dp_cig_t *d, uint32_t dp_job) {
  if (!d || !dp_job) return 0;
  const uint32_t x = dp_job - 1, tid = x >> 24, local = x & 0xffffffu;
  if ((int)tid >= d->n_thr) return 0;
  const int64_t k = d->thr_base[tid] + local;
  return k < d->thr_base[tid + 1] && k < d->n ? &d->res[k] : 0;
}
static __thread const dp_cig_t *tl_dp_cig;

/* mem_alnreg_setSAM (:40-123): final CIGAR with band doubling, position, clipping */
static void set_sam(const bq_opt_t *opt, const bq_ref_t *ref, bq_read_t *s, bq_reg_t *reg) {
  if (reg->n_cigar > 0) return;
  const bsq_cigar_res *r = dp_cig_find(tl_dp_cig, reg->dp_job);
  if (r && r->n_cigar > 0) {
    /* the band loop, the global alignment, MD / NM / ZC / ZR and the clean-up of the CIGAR were done by k_cigar on the
     * job built from this region (cigar_job); what is left is the position */
    int is_rev;
    int64_t rpos = bq_depos(ref, reg->rb < ref->l_pac ? reg->rb : reg->re - 1, &is_rev);
    reg->is_rev = is_rev;
    reg->flag |= reg->is_rev ? 0x10 : 0;
    reg->NM = r->NM; reg->ZC = (uint32_t)r->ZC; reg->ZR = (uint32_t)r->ZR; reg->bss_u = (uint8_t)r->bss_u;
    reg->n_cigar = r->n_cigar;
    reg->cigar = (uint32_t *)(tl_dp_cig->blob + r->off); reg->cigar_ext = 1; /* read-only from here on */
    reg->pos = (int)(rpos + r->lead_del - ref->anns[reg->rid].offset);
    __atomic_fetch_add(&g_dp_stat[1], 1, __ATOMIC_RELAXED);
    return;
  }
  if (tl_dp_cig) __atomic_fetch_add(&g_dp_stat[2], 1, __ATOMIC_RELAXED);
  uint8_t qbuf[512], *query = s->l_seq < (int)sizeof qbuf ? qbuf : malloc((size_t)s->l_seq + 1);
  int i;
  for (i = 0; i < s->l_seq; ++i) query[i] = s->seq[i] < 5 ? s->seq[i] : 4;
  int w = set_sam_band(opt, reg);
  uint32_t *cigar = 0;
  int n_cigar = 0, score = 0, last_sc = -(1 << 30), bss_u_ = 0;
  for (i = 0; i < 3; ++i, w <<= 1, last_sc = score) {
    free(cigar);
    w = MINV(w, opt->w << 2);
    cigar = bq_gen_cigar(reg->parent ? opt->ctmat : opt->gamat, opt->o_del, opt->e_del, opt->o_ins, opt->e_ins, w, ref->l_pac, ref->pac,
                         reg->qe - reg->qb, &query[reg->qb], reg->rb, reg->re, &score, &n_cigar, &reg->NM, &reg->ZC, &reg->ZR, &bss_u_,
                         reg->parent);
    if (score == last_sc) break;
    if (w == opt->w << 2) break;
    if (score >= reg->truesc - opt->a) break;
  }
  reg->bss_u = (uint8_t)bss_u_;
  int l_MD = cigar ? (int)strlen((char *)(cigar + n_cigar)) + 1 : 0;
  int is_rev;
  int64_t rpos = bq_depos(ref, reg->rb < ref->l_pac ? reg->rb : reg->re - 1, &is_rev);
  reg->is_rev = is_rev;
  reg->flag |= reg->is_rev ? 0x10 : 0;
  if (n_cigar > 0) {
    if ((cigar[0] & 0xf) == 2) {
      rpos += cigar[0] >> 4;
      --n_cigar;
      memmove(cigar, cigar + 1, (size_t)n_cigar * 4 + l_MD);
    } else if ((cigar[n_cigar - 1] & 0xf) == 2) {
      --n_cigar;
      memmove(cigar + n_cigar, cigar + n_cigar + 1, (size_t)l_MD);
    }
  }
  if (reg->qb != 0 || reg->qe != s->l_seq || s->clip5 || s->clip3) {
    int clip5 = reg->is_rev ? s->l_seq - reg->qe + s->clip3 : reg->qb + s->clip5;
    int clip3 = reg->is_rev ? reg->qb + s->clip5 : s->l_seq - reg->qe + s->clip3;
    cigar = realloc(cigar, 4 * (size_t)(n_cigar + 2) + l_MD);
    if (clip5) { memmove(cigar + 1, cigar, (size_t)n_cigar * 4 + l_MD); cigar[0] = (uint32_t)clip5 << 4 | 3; ++n_cigar; }
    if (clip3) { memmove(cigar + n_cigar + 1, cigar + n_cigar, (size_t)l_MD); cigar[n_cigar++] = (uint32_t)clip3 << 4 | 3; }
  }
  if (query != qbuf) free(query);
  reg->n_cigar = n_cigar;
  if (reg->n_cigar > 0) reg->cigar = cigar; else free(cigar);
  reg->pos = (int)(rpos - ref->anns[reg->rid].offset);
}

static int get_pri_idx(double XA_drop_ratio, const bq_reg_t *a, int i) { /* mem_alnreg.h:127-131 */
  int k = a[i].secondary_all;
  if (k >= 0 && a[i].score >= a[k].score * XA_drop_ratio) return k;
  return -1;
}

/* One bsq_dp_cigar job = the set_sam call on this region (row = device row of the read) */
static void cigar_job(const bq_opt_t *opt, const bq_ref_t *ref, const bq_read_t *s, const bq_reg_t *reg, int64_t row, bsq_cigar_job *jb) {
  memset(jb, 0, sizeof *jb);
  jb->rb = reg->rb; jb->re = reg->re; jb->row = (int32_t)row; jb->qb = reg->qb; jb->qe = reg->qe;
  jb->w = set_sam_band(opt, reg); jb->truesc = reg->truesc; jb->parent = reg->parent;
  const int is_rev = reg->rb >= ref->l_pac; /* what bns_depos says for rb (forward) / re - 1 (reverse), mem_alnreg.c:77 */
  if (reg->qb != 0 || reg->qe != s->l_seq || s->clip5 || s->clip3) { /* mem_alnreg.c:96-108 */
    jb->clip5 = is_rev ? s->l_seq - reg->qe + s->clip3 : reg->qb + s->clip5;
    jb->clip3 = is_rev ? reg->qb + s->clip5 : s->l_seq - reg->qe + s->clip3;
  }
}

/* Which regions of a read will mem_alnreg_setSAM be called on?  Decided after primary marking, before pairing: every
 * region that can be printed as a record of its own (not a secondary, score >= T; with -a the secondaries that pass
 * the drop ratio too: mem_alnreg_select_format :445-488, mem_reg2sam_pe :680-693) and the secondaries listed in an XA
 * tag (mem_alnreg_format.c:91-134).  Pairing may still pick a secondary that is in neither group (mem_pair); set_sam then
 * finds no job and does that one on the host.  The jobs go to the calling worker's list. */
typedef struct { bsq_cigar_job *a; int64_t n, m; } cj_vec_t; /* the CIGAR jobs one worker thread found */
static int predict_cigar_jobs(const bq_opt_t *opt, const bq_ref_t *ref, const bq_read_t *s, bq_regv_t *regs, int64_t row, cj_vec_t *jobs, int tid) {
  int n = 0;
  /* number of XA candidates per primary (tag_xaxb prints the tag only up to max_XA_hits) */
  uint16_t cbuf[2][128], *cnt_pri = cbuf[0], *cnt_alt = cbuf[1];
  const int xa = !(opt->flag & BQ_F_ALL) && regs->n > 1;
  if (xa) {
    if (regs->n > 128) { cnt_pri = malloc(sizeof(uint16_t) * 2 * regs->n); cnt_alt = cnt_pri + regs->n; }
    memset(cnt_pri, 0, sizeof(uint16_t) * regs->n); memset(cnt_alt, 0, sizeof(uint16_t) * regs->n);
    for (size_t i = 0; i < regs->n; ++i) {
      const int r = get_pri_idx(opt->XA_drop_ratio, regs->a, (int)i);
      if (r >= 0 && (size_t)r < regs->n) { uint16_t *c = regs->a[i].is_alt ? &cnt_alt[r] : &cnt_pri[r]; if (*c < 0xffff) ++*c; }
    }
  }
  for (size_t k = 0; k < regs->n; ++k) {
    bq_reg_t *p = &regs->a[k];
    int want = 0;
    if (p->rb < 0 || p->re < 0 || p->rid < 0) continue;
    if (p->secondary < 0) want = p->score >= opt->T;
    else if ((opt->flag & BQ_F_ALL) && !p->is_alt) want = p->score >= opt->T && (p->secondary >= INT_MAX || p->score >= regs->a[p->secondary].score * opt->drop_ratio);
    if (!want && xa) { /* XA of its primary, if that one prints the tag at all */
      const int r = get_pri_idx(opt->XA_drop_ratio, regs->a, (int)k);
      if (r >= 0 && (size_t)r < regs->n && regs->a[r].score >= opt->T) want = cnt_pri[r] <= opt->max_XA_hits && cnt_alt[r] <= opt->max_XA_hits_alt;
    }
    if (!want) continue;
    if (jobs->n == jobs->m) { jobs->m = jobs->m ? jobs->m << 1 : 4096; jobs->a = realloc(jobs->a, (size_t)jobs->m * sizeof(bsq_cigar_job)); }
    if (jobs->n >= (1 << 24)) continue; /* beyond what the region's tag can address: this one is done on the host */
    cigar_job(opt, ref, s, p, row, &jobs->a[jobs->n]);
    p->dp_job = ((uint32_t)tid << 24 | (uint32_t)jobs->n) + 1;
    ++jobs->n;
    ++n;
  }
  if (cnt_pri != cbuf[0]) free(cnt_pri);
  return n;
}

static void tag_xaxb(const bq_opt_t *opt, const bq_ref_t *ref, bq_read_t *s, const bq_reg_t *p0, const bq_regv_t *regs0, bq_str_t *out) {
  if (!regs0 || (opt->flag & BQ_F_ALL)) return;
  int cnt_pri = 0, cnt_alt = 0;
  size_t i;
  for (i = 0; i < regs0->n; ++i) {
    int r = get_pri_idx(opt->XA_drop_ratio, regs0->a, (int)i);
    if (r >= 0 && regs0->a + r == p0) { if (regs0->a[i].is_alt) ++cnt_alt; else ++cnt_pri; }
  }
  if (cnt_pri <= opt->max_XA_hits && cnt_alt <= opt->max_XA_hits_alt) {
    bq_str_t str = {0, 0, 0};
    int n = 0;
    for (i = 0; i < regs0->n; ++i) {
      bq_reg_t *q = regs0->a + i;
      int r = get_pri_idx(opt->XA_drop_ratio, regs0->a, (int)i);
      if (r < 0 || regs0->a + r != p0) continue;
      if (q->n_cigar == 0) { set_sam(opt, ref, s, q); if (q->n_cigar == 0) continue; }
      if (n) bq_kputc(&str, ';');
      bq_kputs(&str, ref->anns[q->rid].name); bq_kputc(&str, ','); bq_kputc(&str, "+-"[q->is_rev]); bq_kputl(&str, q->pos + 1);
      bq_kputc(&str, ',');
      for (int k = 0; k < q->n_cigar; ++k) { bq_kputw(&str, (int)(q->cigar[k] >> 4)); bq_kputc(&str, "MIDSHN"[q->cigar[k] & 0xf]); }
      bq_kputc(&str, ','); bq_kputw(&str, q->NM);
      ++n;
    }
    if (str.l) { bq_kputsn(out, "\tXA:Z:", 6); bq_kputs(out, str.s); }
    free(str.s);
  }
  if (cnt_pri > 0 || cnt_alt > 0) { bq_kputsn(out, "\tXB:Z:", 6); bq_kputw(out, cnt_pri); bq_kputc(out, ','); bq_kputw(out, cnt_alt); }
}

static void tag_sa(const bq_ref_t *ref, const bq_reg_t *p0, const bq_regv_t *regs0, bq_str_t *out) {
  if (!regs0 || (p0->flag & 0x100)) return;
  bq_str_t str = {0, 0, 0};
  for (size_t i = 0; i < regs0->n; ++i) {
    const bq_reg_t *q = regs0->a + i;
    if (q == p0 || q->n_cigar == 0 || (q->flag & 0x100)) continue;
    bq_kputs(&str, ref->anns[q->rid].name); bq_kputc(&str, ','); bq_kputl(&str, q->pos + 1); bq_kputc(&str, ',');
    bq_kputc(&str, "+-"[q->is_rev]); bq_kputc(&str, ',');
    for (int k = 0; k < q->n_cigar; ++k) { bq_kputw(&str, (int)(q->cigar[k] >> 4)); bq_kputc(&str, "MIDSH"[q->cigar[k] & 0xf]); }
    bq_kputc(&str, ','); bq_kputw(&str, (int)q->mapq); bq_kputc(&str, ','); bq_kputw(&str, q->NM); bq_kputc(&str, ';');
  }
  if (str.l) { bq_kputsn(out, "\tSA:Z:", 6); bq_kputs(out, str.s); }
  free(str.s);
}

/* Fixed-layout part of a SAM record written through a bare pointer: the caller reserves an upper bound once, so the ~100
 * pieces of a record cost no capacity check and no terminator each (they were a third of format_sam's time). */
static const char fw_dig2[201] =
  "0001020304050607080910111213141516171819202122232425262728293031323334353637383940414243444546474849"
  "5051525354555657585960616263646566676869707172737475767778798081828384858687888990919293949596979899";
static inline char *fw_num(char *w, long v) { /* decimal, two digits per division */
  char b[24];
  int n = 0;
  unsigned long u = v < 0 ? 0ul - (unsigned long)v : (unsigned long)v;
  if (v < 0) *w++ = '-';
  while (u >> 32) { const unsigned r = (unsigned)(u % 100); u /= 100; b[n++] = fw_dig2[2 * r + 1]; b[n++] = fw_dig2[2 * r]; }
  unsigned x = (unsigned)u;
  while (x >= 100) { const unsigned r = x % 100; x /= 100; b[n++] = fw_dig2[2 * r + 1]; b[n++] = fw_dig2[2 * r]; }
  if (x >= 10) { b[n++] = fw_dig2[2 * x + 1]; b[n++] = fw_dig2[2 * x]; }
  else b[n++] = (char)('0' + x);
  while (n) *w++ = b[--n];
  return w;
}
static inline char *fw_str(char *w, const char *p) { const size_t n = strlen(p); memcpy(w, p, n); return w + n; }
static inline char *fw_mem(char *w, const char *p, size_t n) { memcpy(w, p, n); return w + n; }
static char *fw_cigar(char *w, const bq_opt_t *opt, const bq_reg_t *r, int is_primary) {
  for (int i = 0; i < r->n_cigar; ++i) {
    int c = r->cigar[i] & 0xf;
    if (!(opt->flag & BQ_F_SOFTCLIP) && !r->is_alt && (c == 3 || c == 4)) c = is_primary ? 3 : 4;
    w = fw_num(w, (long)(r->cigar[i] >> 4)); *w++ = "MIDSH"[c];
  }
  return w;
}

/* SEQ / QUAL of a record: nt4 codes to letters (reverse-complemented for a reverse-strand hit), qualities copied or
 * reversed.  16 bases per step with a byte shuffle where the CPU has one (a third of format_sam's time went into the
 * byte loops); the plain loops otherwise and for the tails. */
static char *seq_out_plain(char *w, const uint8_t *s, int qb, int qe, int rev) {
  if (rev) for (int i = qe - 1; i >= qb; --i) *w++ = "TGCAN"[s[i]];
  else for (int i = qb; i < qe; ++i) *w++ = "ACGTN"[s[i]];
  return w;
}
static char *qual_rev_plain(char *w, const char *q, int qb, int qe) {
  for (int i = qe - 1; i >= qb; --i) *w++ = q[i];
  return w;
}
#if defined(__x86_64__) && defined(__GNUC__)
#include <tmmintrin.h>
__attribute__((target("ssse3"))) static char *seq_out_ssse3(char *w, const uint8_t *s, int qb, int qe, int rev) {
  const __m128i fw = _mm_setr_epi8('A', 'C', 'G', 'T', 'N', 'N', 'N', 'N', 'N', 'N', 'N', 'N', 'N', 'N', 'N', 'N');
  const __m128i rc = _mm_setr_epi8('T', 'G', 'C', 'A', 'N', 'N', 'N', 'N', 'N', 'N', 'N', 'N', 'N', 'N', 'N', 'N');
  const __m128i flip = _mm_setr_epi8(15, 14, 13, 12, 11, 10, 9, 8, 7, 6, 5, 4, 3, 2, 1, 0);
  if (!rev) {
    int i = qb;
    for (; i + 16 <= qe; i += 16, w += 16) _mm_storeu_si128((__m128i *)w, _mm_shuffle_epi8(fw, _mm_loadu_si128((const __m128i *)(s + i))));
    return seq_out_plain(w, s, i, qe, 0);
  }
  int e = qe;
  for (; e - 16 >= qb; e -= 16, w += 16)
    _mm_storeu_si128((__m128i *)w, _mm_shuffle_epi8(_mm_shuffle_epi8(rc, _mm_loadu_si128((const __m128i *)(s + e - 16))), flip));
  return seq_out_plain(w, s, qb, e, 1);
}
__attribute__((target("ssse3"))) static char *qual_rev_ssse3(char *w, const char *q, int qb, int qe) {
  const __m128i flip = _mm_setr_epi8(15, 14, 13, 12, 11, 10, 9, 8, 7, 6, 5, 4, 3, 2, 1, 0);
  int e = qe;
  for (; e - 16 >= qb; e -= 16, w += 16) _mm_storeu_si128((__m128i *)w, _mm_shuffle_epi8(_mm_loadu_si128((const __m128i *)(q + e - 16)), flip));
  return qual_rev_plain(w, q, qb, e);
}
static int have_ssse3(void) { static int v = -1; if (v < 0) v = __builtin_cpu_supports("ssse3") ? 1 : 0; return v; }
static inline char *seq_out(char *w, const uint8_t *s, int qb, int qe, int rev) { return have_ssse3() ? seq_out_ssse3(w, s, qb, qe, rev) : seq_out_plain(w, s, qb, qe, rev); }
static inline char *qual_rev(char *w, const char *q, int qb, int qe) { return have_ssse3() ? qual_rev_ssse3(w, q, qb, qe) : qual_rev_plain(w, q, qb, qe); }
#else
#define seq_out seq_out_plain
#define qual_rev qual_rev_plain
#endif

/* mem_alnreg_formatSAM (:237-436) */
static void format_sam(const bq_opt_t *opt, const bq_ref_t *ref, bq_str_t *str, bq_read_t *s, const bq_reg_t *p0, const bq_reg_t *m0,
                       const bq_regv_t *regs0, int is_primary, const bq_pestat_t *pes, const char *rg_id) {
  bq_reg_t p = *p0, m;
  memset(&m, 0, sizeof m);
  if (m0) m = *m0;
  p.flag |= m0 ? 0x1 : 0;
  p.flag |= m0 && m.rid < 0 ? 0x8 : 0;
  if (m0 && m0->bss_u == 0) p.bss_u = 0;
  if (p.rid >= 0 && m0 && m.rid >= 0 && pes && is_proper_pair(ref, &p, &m, *pes)) { p.flag |= 2; m.flag |= 2; }
  if (p.rid < 0 && m0 && m.rid >= 0) { p.rid = m.rid; p.pos = m.pos; p.is_rev = m.is_rev; p.n_cigar = 0; }
  if (m0 && m.rid < 0 && p.rid >= 0) { m.rid = p.rid; m.pos = p.pos; m.is_rev = p.is_rev; m.n_cigar = 0; }
  p.flag |= m0 && m.is_rev ? 0x20 : 0;
  /* upper bound of everything up to and including the RG tag: names, eleven numbers of at most 20 digits, both CIGARs,
   * sequence and qualities, MD text, tag names */
  const char *md = p.n_cigar ? (const char *)(p.cigar + p.n_cigar) : "";
  const size_t l_name = strlen(s->name), l_cmt = s->comment ? strlen(s->comment) : 0, l_md = strlen(md);
  const size_t l_rn = p.rid >= 0 ? strlen(ref->anns[p.rid].name) : 1, l_mn = m0 && m.rid >= 0 ? strlen(ref->anns[m.rid].name) : 1;
  const size_t l_rg = rg_id ? strlen(rg_id) : 0;
  bq_str_reserve(str, l_name + l_cmt + l_rn + l_mn + l_md + l_rg + 2 * (size_t)s->l_seq0 + 12 * ((size_t)p.n_cigar + 1) + 400);
  char *w = str->s + str->l;
  w = fw_mem(w, s->name, l_name);
  if (s->comment) { *w++ = '_'; w = fw_mem(w, s->comment, l_cmt); }
  *w++ = '\t';
  w = fw_num(w, (p.flag & 0xffff) | (p.flag & 0x10000 ? 0x100 : 0)); *w++ = '\t';
  if (p.rid >= 0) {
    w = fw_mem(w, ref->anns[p.rid].name, l_rn); *w++ = '\t';
    w = fw_num(w, p.pos + 1); *w++ = '\t';
    w = fw_num(w, (int)p.mapq); *w++ = '\t';
    if (p.n_cigar) w = fw_cigar(w, opt, &p, is_primary);
    else *w++ = '*';
  } else w = fw_mem(w, "*\t0\t0\t*", 7);
  *w++ = '\t';
  if (m0 && m.rid >= 0) {
    if (p.rid == m.rid) *w++ = '='; else w = fw_mem(w, ref->anns[m.rid].name, l_mn);
    *w++ = '\t'; w = fw_num(w, m.pos + 1); *w++ = '\t';
    if (p.rid == m.rid) { /* biscuit-specific TLEN (:304-311) */
      int64_t q0 = -1, q1 = -1;
      if (p.is_rev) q1 = p.pos + get_rlen(p.n_cigar, p.cigar) - 1; else q0 = p.pos;
      if (m.is_rev) q1 = m.pos + get_rlen(m.n_cigar, m.cigar) - 1; else q0 = m.pos;
      if (p.n_cigar > 0 && m.n_cigar > 0 && q0 >= 0 && q1 >= 0) w = fw_num(w, (long)(q1 - q0 + 1));
      else *w++ = '0';
    } else *w++ = '0';
  } else w = fw_mem(w, "*\t0\t0", 5);
  *w++ = '\t';
  if (p.flag & 0x100) w = fw_mem(w, "*\t*", 3);
  else {
    int qb = 0, qe = s->l_seq0;
    const int hard = p.n_cigar && !is_primary && !(opt->flag & BQ_F_SOFTCLIP) && !p.is_alt;
    if (p.is_rev) {
      if (hard) {
        if ((p.cigar[0] & 0xf) == 4 || (p.cigar[0] & 0xf) == 3) qe -= (int)(p.cigar[0] >> 4);
        if ((p.cigar[p.n_cigar - 1] & 0xf) == 4 || (p.cigar[p.n_cigar - 1] & 0xf) == 3) qb += (int)(p.cigar[p.n_cigar - 1] >> 4);
      }
      w = seq_out(w, s->seq0, qb, qe, 1);
      *w++ = '\t';
      if (s->qual) w = qual_rev(w, s->qual, qb, qe);
      else *w++ = '*';
    } else {
      if (hard) {
        if ((p.cigar[0] & 0xf) == 4 || (p.cigar[0] & 0xf) == 3) qb += (int)(p.cigar[0] >> 4);
        if ((p.cigar[p.n_cigar - 1] & 0xf) == 4 || (p.cigar[p.n_cigar - 1] & 0xf) == 3) qe -= (int)(p.cigar[p.n_cigar - 1] >> 4);
      }
      w = seq_out(w, s->seq0, qb, qe, 0);
      *w++ = '\t';
      if (s->qual) { if (qe > qb) w = fw_mem(w, s->qual + qb, (size_t)(qe - qb)); }
      else *w++ = '*';
    }
  }
  if (p.n_cigar) {
    w = fw_mem(w, "\tNM:i:", 6); w = fw_num(w, p.NM);
    w = fw_mem(w, "\tMD:Z:", 6); w = fw_mem(w, md, l_md);
    w = fw_mem(w, "\tZC:i:", 6); w = fw_num(w, (int)p.ZC);
    w = fw_mem(w, "\tZR:i:", 6); w = fw_num(w, (int)p.ZR);
  }
  if (p.score >= 0) { w = fw_mem(w, "\tAS:i:", 6); w = fw_num(w, p.score); }
  if (p.sub >= 0) { w = fw_mem(w, "\tXS:i:", 6); w = fw_num(w, MAXV(p.sub, p.csub)); }
  if (rg_id && rg_id[0]) { w = fw_mem(w, "\tRG:Z:", 6); w = fw_mem(w, rg_id, l_rg); }
  *w = 0;
  str->l = (size_t)(w - str->s);
  /* variable-length tails through the checked writers */
  if (regs0) tag_sa(ref, p0, regs0, str);
  if (is_primary && p.alt_sc > 0) { char b[64]; snprintf(b, sizeof b, "\tPA:f:%.3f", (double)p.score / p.alt_sc); bq_kputs(str, b); }
  bq_kputsn(str, "\tXL:i:", 6); bq_kputw(str, s->l_seq);
  if (regs0) tag_xaxb(opt, ref, s, p0, regs0, str);
  if ((opt->flag & BQ_F_REF_HDR) && p.rid >= 0 && ref->anns[p.rid].anno != 0 && ref->anns[p.rid].anno[0] != 0) {
    bq_kputsn(str, "\tXR:Z:", 6);
    size_t t0 = str->l;
    bq_kputs(str, ref->anns[p.rid].anno);
    for (size_t i = t0; i < str->l; ++i) if (str->s[i] == '\t') str->s[i] = ' ';
  }
  if (s->barcode) { bq_kputsn(str, "\tCB:Z:", 6); bq_kputs(str, s->barcode); }
  if (s->umi) { bq_kputsn(str, "\tRX:Z:", 6); bq_kputs(str, s->umi); }
  bq_str_reserve(str, 12 * ((size_t)m.n_cigar + 1) + 64);
  w = str->s + str->l;
  w = fw_mem(w, "\tMC:Z:", 6);
  if (m.n_cigar) w = fw_cigar(w, opt, &m, is_primary); else *w++ = '*';
  w = fw_mem(w, "\tMQ:i:", 6); w = fw_num(w, (int)m.mapq);
  w = fw_mem(w, "\tYD:A:", 6);
  *w++ = p.bss_u ? 'u' : "fr"[p.bss];
  *w++ = '\n';
  *w = 0;
  str->l = (size_t)(w - str->s);
}

/* Where the SAM text of a read goes: phase-2 workers point tl_sam_slab at their own slab and the text is formatted
 * in place (offset kept in sam_off, pointers fixed up when the batch is done); without it (direct callers of
 * bq_reg2sam_*) every read gets its own string as in the reference. */
static __thread bq_str_t *tl_sam_slab;
typedef struct { bq_str_t own, *str; size_t off0; } sam_out_t;
static bq_str_t *sam_out_begin(sam_out_t *o, const bq_read_t *s) {
  o->own.l = o->own.m = 0; o->own.s = 0;
  o->str = tl_sam_slab ? tl_sam_slab : &o->own;
  o->off0 = o->str->l;
  bq_str_reserve(o->str, 2 * (size_t)s->l_seq0 + 400);
  return o->str;
}
static void sam_out_end(sam_out_t *o, bq_read_t *s) {
  if (o->str == &o->own) { s->sam = o->own.s; s->sam_len = o->own.l; return; }
  bq_str_reserve(o->str, 1);
  s->sam_len = o->str->l - o->off0;
  o->str->s[o->str->l++] = 0; /* keep the terminating NUL inside the slab */
  s->sam = 0; s->sam_in_slab = 1; s->sam_off = o->off0;
}

/* mem_alnreg_select_format (:445-488) */
static int *select_format(const bq_opt_t *opt, const bq_ref_t *ref, bq_read_t *s, bq_regv_t *regs, int *n_out, int *buf, int n_buf) {
  int *out = regs->n < (size_t)n_buf ? buf : malloc(sizeof(int) * (regs->n + 1)), l = 0;
  for (size_t k = 0; k < regs->n; ++k) {
    bq_reg_t *p = regs->a + k;
    if (p->rb < 0 || p->re < 0) continue;
    if (p->score < opt->T) continue;
    if (p->secondary >= 0 && (p->is_alt || !(opt->flag & BQ_F_ALL))) continue;
    if (p->secondary >= 0 && p->secondary < INT_MAX && p->score < regs->a[p->secondary].score * opt->drop_ratio) continue;
    if (l && p->secondary < 0) p->flag |= (opt->flag & BQ_F_NO_MULTI) ? 0x10000 : 0x800;
    if (p->secondary >= 0) p->flag |= 0x100;
    p->mapq = p->secondary < 0 ? (unsigned)bq_approx_mapq_se(opt, p) : 0;
    if (!(opt->flag & BQ_F_KEEP_SUPP_MAPQ) && l && !p->is_alt) p->mapq = MINV(p->mapq, regs->a[0].mapq);
    set_sam(opt, ref, s, p);
    out[l++] = (int)k;
  }
  *n_out = l;
  return out;
}

void bq_reg2sam_se(const bq_opt_t *opt, const bq_ref_t *ref, bq_read_t *s, bq_regv_t *regs, const char *rg_id) {
  sam_out_t so;
  bq_str_t *str = sam_out_begin(&so, s);
  int tobuf[64];
  int n, *to = select_format(opt, ref, s, regs, &n, tobuf, 64);
  if (n > 0) for (int i = 0; i < n; ++i) format_sam(opt, ref, str, s, &regs->a[to[i]], 0, regs, !i, 0, rg_id);
  else {
    bq_reg_t reg;
    memset(&reg, 0, sizeof reg);
    reg.rid = -1; reg.flag = 0x4;
    format_sam(opt, ref, str, s, &reg, 0, regs, 1, 0, rg_id);
  }
  sam_out_end(&so, s);
  if (to != tobuf) free(to);
}

static void reg2sam_pe_nopairing(const bq_opt_t *opt, const bq_ref_t *ref, bq_read_t s[2], bq_regv_t regs[2], bq_pestat_t pes,
                                 const char *rg_id) {
  bq_reg_t *best[2] = {0, 0}, unmapped[2];
  int tobuf[2][64];
  int *to[2], n_to[2], i;
  for (i = 0; i < 2; ++i) {
    to[i] = select_format(opt, ref, &s[i], &regs[i], &n_to[i], tobuf[i], 64);
    if (n_to[i] > 0) best[i] = &regs[i].a[to[i][0]];
    else {
      memset(&unmapped[i], 0, sizeof(bq_reg_t));
      unmapped[i].rid = -1; unmapped[i].flag = 0x40 << i | 0x1 | 0x4;
      best[i] = &unmapped[i];
    }
  }
  for (i = 0; i < 2; ++i) {
    sam_out_t so;
    bq_str_t *str = sam_out_begin(&so, &s[i]);
    if (n_to[i]) {
      for (int j = 0; j < n_to[i]; ++j) format_sam(opt, ref, str, &s[i], &regs[i].a[to[i][j]], best[!i], &regs[i], !j, &pes, rg_id);
    } else format_sam(opt, ref, str, &s[i], best[i], best[!i], 0, 1, &pes, rg_id);
    sam_out_end(&so, &s[i]);
  }
  for (i = 0; i < 2; ++i) if (to[i] != tobuf[i]) free(to[i]);
}

#define RAW_MAPQ(diff, a) ((int)(6.02 * (diff) / (a) + .499))

/* mem_reg2sam_pe (:562-696) */
void bq_reg2sam_pe(const bq_opt_t *opt, const bq_ref_t *ref, uint64_t id, bq_read_t s[2], bq_regv_t regs[2], bq_pestat_t pes,
                   const char *rg_id) {
  int i;
  size_t k, j;
  for (i = 0; i < 2; ++i)
    for (k = 0; k < regs[i].n; ++k) regs[i].a[k].flag |= (0x40 << i) | 1;
  if ((opt->flag & BQ_F_NOPAIRING) || regs[0].n_pri == 0 || regs[1].n_pri == 0) { reg2sam_pe_nopairing(opt, ref, s, regs, pes, rg_id); return; }
  int is_multi[2];
  for (i = 0; i < 2; ++i) {
    for (j = 1; j < regs[i].n_pri; ++j)
      if (regs[i].a[j].secondary < 0 && regs[i].a[j].score >= opt->T) break;
    is_multi[i] = j < regs[i].n_pri ? 1 : 0;
  }
  if (is_multi[0] || is_multi[1]) { reg2sam_pe_nopairing(opt, ref, s, regs, pes, rg_id); return; }
  int pscore, sub_pscore, n_sub, z[2] = {0, 0};
  const double q0 = g_prof > 0 ? bq_now() : 0;
  pair_regs(opt, ref, pes, regs, (int)id, &pscore, &sub_pscore, &n_sub, z);
  if (g_prof > 0) { const double q1 = bq_now(); pthread_mutex_lock(&g_prof_mu); g_t_pair += q1 - q0; pthread_mutex_unlock(&g_prof_mu); }
  if (pscore <= 0) { reg2sam_pe_nopairing(opt, ref, s, regs, pes, rg_id); return; }
  int score_unpaired = regs[0].a[0].score + regs[1].a[0].score - opt->pen_unpaired;
  if (pscore > score_unpaired) {
    sub_pscore = MAXV(sub_pscore, score_unpaired);
    int q_pe = RAW_MAPQ(pscore - sub_pscore, opt->a);
    if (n_sub > 0) q_pe -= (int)(4.343 * log(n_sub + 1) + .499);
    q_pe = MAXV(0, MINV(60, q_pe));
    q_pe = (int)(q_pe * (1. - .5 * (regs[0].a[0].frac_rep + regs[1].a[0].frac_rep)) + .499);
    int q_se[2];
    bq_reg_t *c[2] = {&regs[0].a[z[0]], &regs[1].a[z[1]]};
    for (i = 0; i < 2; ++i) {
      if (c[i]->secondary >= 0) { c[i]->sub = regs[i].a[c[i]->secondary].score; c[i]->secondary = -2; }
      q_se[i] = bq_approx_mapq_se(opt, c[i]);
    }
    q_se[0] = MAXV(q_se[0], MINV(q_pe, q_se[0] + 40));
    q_se[1] = MAXV(q_se[1], MINV(q_pe, q_se[1] + 40));
    c[0]->mapq = (unsigned)MINV(q_se[0], RAW_MAPQ(c[0]->score - c[0]->csub, opt->a));
    c[1]->mapq = (unsigned)MINV(q_se[1], RAW_MAPQ(c[1]->score - c[1]->csub, opt->a));
  } else {
    z[0] = z[1] = 0;
    regs[0].a[0].mapq = (unsigned)bq_approx_mapq_se(opt, &regs[0].a[0]);
    regs[1].a[0].mapq = (unsigned)bq_approx_mapq_se(opt, &regs[1].a[0]);
  }
  for (i = 0; i < 2; ++i) { /* a chosen secondary swaps roles with its primary */
    bq_regv_t *r = &regs[i];
    int kk = r->a[z[i]].secondary_all;
    if (kk >= 0 && (size_t)kk < r->n_pri) {
      for (j = 0; j < r->n; ++j)
        if (r->a[j].secondary_all == kk || j == (size_t)kk) r->a[j].secondary_all = z[i];
      r->a[z[i]].secondary_all = -1;
    }
  }
  const double q2 = g_prof > 0 ? bq_now() : 0;
  for (i = 0; i < 2; ++i) set_sam(opt, ref, &s[i], &regs[i].a[z[i]]);
  if (g_prof > 0) { const double q3 = bq_now(); pthread_mutex_lock(&g_prof_mu); g_t_setsam += q3 - q2; pthread_mutex_unlock(&g_prof_mu); }
  for (i = 0; i < 2; ++i) {
    sam_out_t so;
    bq_regv_t *r = &regs[i];
    bq_str_t *str = sam_out_begin(&so, &s[i]);
    format_sam(opt, ref, str, &s[i], r->a + z[i], regs[!i].a + z[!i], r, 1, &pes, rg_id);
    if (r->n_pri < r->n) {
      bq_reg_t *p = &r->a[r->n_pri];
      if (p->score >= opt->T && p->secondary < 0) {
        p->flag |= 0x800;
        set_sam(opt, ref, &s[i], p);
        format_sam(opt, ref, str, &s[i], p, 0, r, 0, &pes, rg_id);
      }
    }
    sam_out_end(&so, &s[i]);
  }
}

/* ---------------- read clipping: bwamem.c:238-303 ---------------- */

static const uint8_t *find_sub(const uint8_t *hay, size_t hlen, const uint8_t *needle, size_t nlen) {
  if (!nlen) return 0;
  for (size_t i = 0; i + nlen <= hlen; ++i)
    if (hay[i] == needle[0] && memcmp(hay + i, needle, nlen) == 0) return hay + i;
  return 0;
}

void bq_read_clipping(bq_read_t *s, const uint8_t *adaptor, int l_adaptor, const bq_opt_t *opt) {
  if (adaptor == 0) s->l_adaptor = 0;
  else {
    const uint8_t *hit = find_sub(s->seq, (size_t)s->l_seq, adaptor, (size_t)l_adaptor);
    if (hit) s->l_adaptor = s->l_seq - (int)(hit - s->seq);
    else {
      int i;
      for (i = l_adaptor - 1; i; --i)
        if (memcmp(s->seq + s->l_seq - i, adaptor, (size_t)i) == 0) break;
      s->l_adaptor = i;
    }
  }
  s->clip5 = opt->clip5;
  s->clip3 = opt->clip3 + s->l_adaptor;
  if (s->qual) {
    for (; s->clip5 < s->l_seq - s->clip3; s->clip5++)
      if (s->qual[s->clip5] >= opt->min_base_qual + 33) break;
    for (; s->l_seq - s->clip3 >= s->clip5; s->clip3++)
      if (s->qual[s->l_seq - s->clip3 - 1] >= opt->min_base_qual + 33) break;
  }
  s->seq0 = s->seq; s->l_seq0 = s->l_seq;
  s->seq += s->clip5;
  s->l_seq = s->l_seq - s->clip3 - s->clip5;
  if (s->l_seq < 0) s->l_seq = 0;
}

/* ---------------- the batch driver: mem_process_seqs (bwamem.c:432-476) ---------------- */


/* stages of the host phase 2 (one parallel section each, items = reads or pairs) */
enum { ST_GENERIC = 0, ST_MERGE, ST_PESTAT, ST_PLAN, ST_MS_FILL, ST_MARK, ST_SAM };
#define PES_NONE INT64_MIN

typedef struct {
  const bq_opt_t *opt; const bq_ref_t *ref; bq_read_t *seqs; bq_regv_t *regs; bq_pestat_t pes; int64_t n_processed;
  const char *rg_id; int n_items, n_threads, pe, stage;
  long next_item; /* work distribution (thr_main) */
  const bsq_reg *dev_regs; const int64_t *reg_off; const int64_t *task_of_read; /* first task of each read */
  const uint8_t *n_task_of_read;
  bq_reg_t *reg_pool; const int64_t *pool_off; /* regions of read i: reg_pool[pool_off[i] .. pool_off[i + 1]) */
  bq_str_t *sam_slab;   /* per worker thread: SAM text of the reads it formatted */
  size_t *sam_off;      /* per read: offset of its text in its thread's slab */
  uint8_t *sam_thr;     /* per read: which thread's slab */
  int64_t *pes_is;      /* per pair: the insert size it contributes to mem_pestat, or PES_NONE */
  void (*gen_fn)(void *ctx, long i); void *gen_ctx; int no_spawn; /* ST_GENERIC: a plain parallel loop (batch preparation) */
  /* batched DP on the GPU (bsq_dp_*), when the batch has a context */
  int use_dp;
  int64_t *ms_first;            /* per pair (+1): its mate-rescue jobs [ms_first[p], ms_first[p+1]) */
  bsq_matesw_job *mjobs; bsq_matesw_res *mres; uint32_t *mkeys;
  cj_vec_t tjobs[256];          /* per worker thread: the CIGAR jobs it found (regions carry thread << 24 | index) */
  int64_t thr_base[257];        /* where each thread's jobs start in cjobs / cres */
  bsq_cigar_job *cjobs; bsq_cigar_res *cres;
  dp_cig_t cig;                 /* results as set_sam sees them */
  pair_tab_t pair_tab; double *pair_term; /* insert-size term of the pair score, tabulated per batch */
} work_t;

static void reg_from_dev(const bsq_reg *d, bq_reg_t *r) {
  memset(r, 0, sizeof *r);
  r->rb = d->rb; r->re = d->re; r->qb = d->qb; r->qe = d->qe; r->rid = d->rid; r->score = d->score; r->truesc = d->truesc; r->w = d->w;
  r->seedcov = d->seedcov; r->seedlen0 = (int16_t)d->seedlen0; r->frac_rep = d->frac_rep; r->bss = d->bss; r->parent = d->parent;
}

/* primary marking of an item (read or pair) and, with a DP context, its CIGAR jobs into the calling worker's list */
static void mark_and_predict(work_t *w, long i, int tid) {
  const double p1 = g_prof > 0 ? bq_now() : 0;
  if (!w->pe) {
    bq_mark_primary(w->opt, &w->regs[i], w->n_processed + i);
    for (size_t k = 0; k < w->regs[i].n; ++k) w->regs[i].a[k].flag = 0;
    if (w->use_dp && w->n_task_of_read[i]) predict_cigar_jobs(w->opt, w->ref, &w->seqs[i], &w->regs[i], w->task_of_read[i], &w->tjobs[tid], tid);
  } else {
    bq_mark_primary(w->opt, &w->regs[i << 1 | 0], i << 1 | 0); /* PE ids lack n_processed (bwamem.c:408,413) */
    bq_mark_primary(w->opt, &w->regs[i << 1 | 1], i << 1 | 1);
    for (int e = 0; e < 2; ++e)
      for (size_t k = 0; k < w->regs[i << 1 | e].n; ++k) w->regs[i << 1 | e].a[k].flag = 0;
    if (w->use_dp)
      for (int e = 0; e < 2; ++e)
        if (w->n_task_of_read[i << 1 | e])
          predict_cigar_jobs(w->opt, w->ref, &w->seqs[i << 1 | e], &w->regs[i << 1 | e], w->task_of_read[i << 1 | e], &w->tjobs[tid], tid);
  }
  if (g_prof > 0) { const double p2 = bq_now(); pthread_mutex_lock(&g_prof_mu); g_t_mark += p2 - p1; pthread_mutex_unlock(&g_prof_mu); }
}

static __thread bq_reg_t *tl_scr; /* per worker: where the regions of one read are gathered and merged */
static __thread size_t tl_scr_cap;

static void work_item(work_t *w, long i, int tid) {
  switch (w->stage) {
  case ST_GENERIC: w->gen_fn(w->gen_ctx, i); return;
  case ST_MERGE: { /* gather the regions of read i in the reference's order and merge them */
    bq_regv_t *rv = &w->regs[i];
    rv->n = rv->n_pri = 0;
    /* The regions of read i live in a slice of the batch's pool: room for every device region of its tasks + 2 (mate
     * rescue may add hits; a push beyond that moves the vector to the heap, regv_push).  Gathering and merging run in a
     * scratch array of the worker that stays in cache -- two thirds of the device regions do not survive the merge --
     * and only the survivors are written to the slice. */
    bq_reg_t *slice = w->reg_pool + w->pool_off[i];
    const size_t room = (size_t)(w->pool_off[i + 1] - w->pool_off[i]);
    if (tl_scr_cap < room) {
      free(tl_scr);
      tl_scr_cap = room * 2 + 64;
      tl_scr = aligned_alloc(64, tl_scr_cap * sizeof(bq_reg_t));
    }
    rv->a = tl_scr; rv->m = room; rv->pooled = 1;
    for (int t = 0; t < w->n_task_of_read[i]; ++t) {
      const int64_t task = w->task_of_read[i] + t;
      for (int64_t k = w->reg_off[task]; k < w->reg_off[task + 1]; ++k) reg_from_dev(&w->dev_regs[k], &rv->a[rv->n++]); /* there is room for all of them */
    }
    bq_merge_regions(w->opt, w->ref, w->seqs[i].seq, w->seqs[i].l_seq, rv);
    if (rv->a == tl_scr) { memcpy(slice, tl_scr, rv->n * sizeof(bq_reg_t)); rv->a = slice; } /* else: moved to the heap by a push (not expected here) */
    return;
  }
  case ST_PESTAT: { /* pair i: its candidate for the insert-size statistics (the serial part only sorts and sums) */
    int64_t is;
    w->pes_is[i] = pestat_candidate(w->opt, w->ref, &w->regs[i << 1], &w->regs[i << 1 | 1], &is) ? is : PES_NONE;
    return;
  }
  case ST_PLAN: case ST_MS_FILL: { /* pair i: its mate-rescue alignments as jobs of one bsq_dp_matesw batch */
    const int64_t row[2] = {w->task_of_read[i << 1], w->task_of_read[(i << 1) + 1]};
    const int both = w->n_task_of_read[i << 1] != 0 && w->n_task_of_read[(i << 1) + 1] != 0;
    if (w->stage == ST_MS_FILL) {
      if (both && w->ms_first[i + 1] > w->ms_first[i])
        matesw_pair(w->opt, w->ref, w->pes, &w->seqs[i << 1], &w->regs[i << 1], 0, w->mjobs + w->ms_first[i], w->mkeys + w->ms_first[i], row, 1);
      return;
    }
    const int n_ms = both ? matesw_pair(w->opt, w->ref, w->pes, &w->seqs[i << 1], &w->regs[i << 1], 0, 0, 0, row, 0) : 0;
    w->ms_first[i + 1] = n_ms;
    /* nothing to try for this pair (every attempt is ruled out by the state the pair is in, which no rescue will change):
     * primary marking and the CIGAR jobs follow at once, while the pair's regions are in cache */
    if (n_ms == 0) mark_and_predict(w, i, tid);
    return;
  }
  case ST_MARK: { /* mate rescue (replaying the GPU's alignments) for the pairs that had attempts planned, then as above;
                   * without a DP context / without rescue: every item */
    if (w->pe && w->ms_first) {
      if (w->ms_first[i] == w->ms_first[i + 1]) return; /* done in ST_PLAN */
      ms_pre_t pre = {w->mres, w->mkeys, w->ms_first[i], w->ms_first[i + 1]};
      matesw_pair(w->opt, w->ref, w->pes, &w->seqs[i << 1], &w->regs[i << 1], &pre, 0, 0, 0, 0);
    } else if (w->pe && !(w->opt->flag & BQ_F_NO_RESCUE)) {
      const double p0 = g_prof > 0 ? bq_now() : 0;
      matesw_pair(w->opt, w->ref, w->pes, &w->seqs[i << 1], &w->regs[i << 1], 0, 0, 0, 0, 0);
      if (g_prof > 0) { const double p1 = bq_now(); pthread_mutex_lock(&g_prof_mu); g_t_mate += p1 - p0; pthread_mutex_unlock(&g_prof_mu); }
    }
    mark_and_predict(w, i, tid);
    return;
  }
  default: break;
  }
  /* ST_SAM: pairing, mapQ, SAM text */
  tl_sam_slab = &w->sam_slab[tid];
  tl_dp_cig = w->use_dp ? &w->cig : 0;
  tl_pair_tab = w->pair_term ? &w->pair_tab : 0;
  if (!w->pe) {
    if (tl_sam_slab->m == 0) { tl_sam_slab->s = bq_big_alloc((size_t)(w->n_items / w->n_threads + 1) * (2 * (size_t)w->seqs[i].l_seq0 + 256), &tl_sam_slab->m); tl_sam_slab->s[0] = 0; }
    bq_reg2sam_se(w->opt, w->ref, &w->seqs[i], &w->regs[i], w->rg_id);
  } else {
    if (tl_sam_slab->m == 0) { tl_sam_slab->s = bq_big_alloc((size_t)(w->n_items / w->n_threads + 1) * 2 * (2 * (size_t)w->seqs[i << 1].l_seq0 + 256), &tl_sam_slab->m); tl_sam_slab->s[0] = 0; }
    const double p2 = g_prof > 0 ? bq_now() : 0;
    bq_reg2sam_pe(w->opt, w->ref, (uint64_t)((w->n_processed >> 1) + i), &w->seqs[i << 1], &w->regs[i << 1], w->pes, w->rg_id);
    if (g_prof > 0) {
      const double p3 = bq_now();
      pthread_mutex_lock(&g_prof_mu);
      g_t_sam += p3 - p2;
      pthread_mutex_unlock(&g_prof_mu);
    }
  }
  { /* the regions of this item are done with: release them here, on the worker */
    const long lo = w->pe ? i << 1 : i, hi = w->pe ? (i << 1) + 2 : i + 1;
    tl_sam_slab = 0;
    tl_dp_cig = 0;
    tl_pair_tab = 0;
    for (long r = lo; r < hi; ++r) { /* the text is already in this thread's slab (sam_out_end); a stray own string is moved there */
      bq_read_t *rd = &w->seqs[r];
      if (rd->sam_in_slab) { w->sam_off[r] = rd->sam_off; w->sam_thr[r] = (uint8_t)tid; continue; }
      if (!rd->sam) continue;
      bq_str_t *sl = &w->sam_slab[tid];
      const size_t l = strlen(rd->sam);
      w->sam_off[r] = sl->l; w->sam_thr[r] = (uint8_t)tid;
      bq_kputsn(sl, rd->sam, l + 1); /* with the terminating NUL */
      free(rd->sam);
      rd->sam = 0; rd->sam_in_slab = 1;
    }
    for (long r = lo; r < hi; ++r) {
      for (size_t k = 0; k < w->regs[r].n; ++k) if (w->regs[r].a[k].n_cigar > 0 && !w->regs[r].a[k].cigar_ext) free(w->regs[r].a[k].cigar);
      if (!w->regs[r].pooled) free(w->regs[r].a);
      w->regs[r].a = 0; w->regs[r].n = 0;
    }
  }
}

typedef struct { work_t *w; int tid; } thr_t;
/* Items are handed out in chunks of neighbouring reads: neighbours share cache lines in the per-read arrays (regs,
 * seqs, sam_off), so a strided split makes every thread write into every other thread's lines.  Which thread
 * formats a read does not show in the output (the text is found through sam_thr / sam_off). */
#define BQ_WORK_CHUNK 64
static void *thr_main(void *a) {
  thr_t *t = a;
  work_t *w = t->w;
  for (;;) {
    const long lo = __atomic_fetch_add(&w->next_item, BQ_WORK_CHUNK, __ATOMIC_RELAXED);
    if (lo >= w->n_items) break;
    const long hi = lo + BQ_WORK_CHUNK < w->n_items ? lo + BQ_WORK_CHUNK : w->n_items;
    for (long i = lo; i < hi; ++i) work_item(w, i, t->tid);
  }
  return 0;
}
/* Worker threads are created once and parked between calls: a batch runs two parallel sections (region merge, then
 * pairing + SAM), and creating the threads for each of them costs milliseconds per thread where mmap / clone are slow
 * (measured 4 ms per pthread_create in a micro-VM), i.e. more than the work itself for small batches. */
static struct {
  pthread_mutex_t user;            /* one parallel section at a time */
  pthread_mutex_t mu; pthread_cond_t go, done;
  pthread_t th[256]; thr_t arg[256];
  int n, active, running; unsigned long gen; work_t *w; /* active: pool threads taking part in the current section */
} g_tp = {PTHREAD_MUTEX_INITIALIZER, PTHREAD_MUTEX_INITIALIZER, PTHREAD_COND_INITIALIZER, PTHREAD_COND_INITIALIZER, {0}, {{0, 0}}, 0, 0, 0, 0, 0};

static void *pool_main(void *a) {
  thr_t *t = a;
  unsigned long seen = 0;
  for (;;) {
    pthread_mutex_lock(&g_tp.mu);
    while (g_tp.gen == seen) pthread_cond_wait(&g_tp.go, &g_tp.mu);
    seen = g_tp.gen;
    t->w = g_tp.w;
    const int take_part = t->tid <= g_tp.active; /* a section may use fewer threads than the pool holds */
    pthread_mutex_unlock(&g_tp.mu);
    if (!take_part) continue;
    thr_main(t);
    pthread_mutex_lock(&g_tp.mu);
    if (--g_tp.running == 0) pthread_cond_signal(&g_tp.done);
    pthread_mutex_unlock(&g_tp.mu);
  }
  return 0;
}

static void run_threads(work_t *w, int n_items) {
  w->n_items = n_items;
  w->next_item = 0;
  int nt = w->n_threads < 1 ? 1 : w->n_threads;
  if (nt > 255) nt = 255;
  w->n_threads = nt;
  if (nt == 1) { for (long i = 0; i < n_items; ++i) work_item(w, i, 0); return; }
  /* the batch preparation (no_spawn) only borrows the pool when it is free; phase 2 waits for it (a preparation loop
   * takes a few milliseconds) rather than paying for threads of its own */
  if (w->no_spawn ? pthread_mutex_trylock(&g_tp.user) == 0 : pthread_mutex_lock(&g_tp.user) == 0) {
    /* workers 1 .. nt-1 come from the pool (grown on demand), the caller is worker 0 */
    pthread_mutex_lock(&g_tp.mu);
    while (g_tp.n < nt - 1) {
      g_tp.arg[g_tp.n].tid = g_tp.n + 1; g_tp.arg[g_tp.n].w = 0;
      if (pthread_create(&g_tp.th[g_tp.n], 0, pool_main, &g_tp.arg[g_tp.n]) != 0) break;
      pthread_detach(g_tp.th[g_tp.n]);
      ++g_tp.n;
    }
    const int have = g_tp.n >= nt - 1;
    if (have) { g_tp.w = w; g_tp.active = nt - 1; g_tp.running = nt - 1; ++g_tp.gen; pthread_cond_broadcast(&g_tp.go); }
    pthread_mutex_unlock(&g_tp.mu);
    if (have) {
      thr_t me = {w, 0};
      thr_main(&me);
      pthread_mutex_lock(&g_tp.mu);
      while (g_tp.running > 0) pthread_cond_wait(&g_tp.done, &g_tp.mu);
      pthread_mutex_unlock(&g_tp.mu);
      pthread_mutex_unlock(&g_tp.user);
      return;
    }
    pthread_mutex_unlock(&g_tp.user);
  }
  if (w->no_spawn) { for (long i = 0; i < n_items; ++i) work_item(w, i, 0); return; } /* pool busy: not worth threads of its own */
  /* pool busy (another caller inside a parallel section) or it could not be built: threads for this call only */
  pthread_t *th = malloc(sizeof(pthread_t) * (size_t)nt);
  thr_t *ta = malloc(sizeof(thr_t) * (size_t)nt);
  for (int t = 0; t < nt; ++t) { ta[t].w = w; ta[t].tid = t; pthread_create(&th[t], 0, thr_main, &ta[t]); }
  for (int t = 0; t < nt; ++t) pthread_join(th[t], 0);
  free(th); free(ta);
}

/* The batch in three parts so that callers can pipeline them (bq_pipeline_run): host preparation (clipping, task
 * list), the GPU part (stage / run / fetch through page-locked buffers) and the host phase 2.  (The reference
 * overlaps I/O and compute the same way with kt_pipeline, align.c:577.) */
typedef struct bq_slot {
  void *tseq; size_t tseq_cap;   /* page-locked task rows */
  void *regs; size_t regs_cap;   /* page-locked regions */
  void *mjobs, *mres, *cjobs, *cres; size_t mjobs_cap, mres_cap, cjobs_cap, cres_cap; /* page-locked DP jobs / results */
  struct bq_slot *next;
} bq_slot_t;
static pthread_mutex_t slot_mu = PTHREAD_MUTEX_INITIALIZER;
static bq_slot_t *slot_free_list;

static bq_slot_t *slot_get(void) {
  pthread_mutex_lock(&slot_mu);
  bq_slot_t *s = slot_free_list;
  if (s) slot_free_list = s->next;
  pthread_mutex_unlock(&slot_mu);
  if (!s) s = calloc(1, sizeof *s);
  return s;
}
static void slot_put(bq_slot_t *s) {
  if (!s) return;
  pthread_mutex_lock(&slot_mu);
  s->next = slot_free_list; slot_free_list = s;
  pthread_mutex_unlock(&slot_mu);
}
static int slot_reserve(void **p, size_t *cap, size_t need) {
  if (need <= *cap) return 0;
  if (*p) bsq_host_free(*p);
  *p = 0; *cap = 0;
  const size_t want = need + need / 4 + 4096;
  const int rc = bsq_host_alloc(p, want);
  if (rc) return rc;
  *cap = want;
  return 0;
}

struct bq_batch {
  int n;
  int64_t n_processed;
  bq_read_t *seqs;
  bsq_reg *dregs;
  int64_t *reg_off, *task_of_read;
  uint8_t *n_task;
  /* prepared for the GPU */
  int64_t nt;
  int stride;
  int32_t *tlen;
  uint8_t *par;
  bq_slot_t *slot;
  /* host phase 2 between its two halves (bq_batch_finish_a / _b) */
  bsq_aligner *al; int out_slot; /* where the regions wait on the device (bq_batch_fetch) */
  bsq_dp *dp;
  struct bq_fin *fin;
};

/* the two per-read loops of the preparation run on the phase-2 workers when those are idle (the start of a run, a
 * GPU-bound pipeline), otherwise on the calling thread */
typedef struct { const bq_opt_t *opt; bq_read_t *seqs; int pe; uint8_t *tseq; int stride; const int64_t *task_of_read; const uint8_t *n_task; int32_t *tlen; } prep_t;
static void prep_clip(void *ctx, long i) {
  prep_t *p = ctx;
  const int second = p->pe && (i & 1);
  bq_read_clipping(&p->seqs[i], second ? p->opt->adaptor2 : p->opt->adaptor1, second ? p->opt->l_adaptor2 : p->opt->l_adaptor1, p->opt);
}
static void prep_rows(void *ctx, long i) {
  prep_t *p = ctx;
  for (int t = 0; t < p->n_task[i]; ++t) {
    uint8_t *row = p->tseq + (size_t)(p->task_of_read[i] + t) * p->stride;
    memcpy(row, p->seqs[i].seq, (size_t)p->seqs[i].l_seq);
    memset(row + p->seqs[i].l_seq, 0, (size_t)(p->stride - p->seqs[i].l_seq));
    p->tlen[p->task_of_read[i] + t] = p->seqs[i].l_seq;
  }
}
void bq_parallel_for(int n_threads, long n, void (*fn)(void *ctx, long i), void *ctx) {
  work_t w;
  memset(&w, 0, sizeof w);
  w.stage = ST_GENERIC; w.gen_fn = fn; w.gen_ctx = ctx; w.no_spawn = 1;
  w.n_threads = n_threads > 8 ? 8 : n_threads; /* memory-bound loops: a few threads are enough */
  run_threads(&w, (int)n);
}
static void prep_loop(const bq_opt_t *opt, prep_t *p, int n, void (*fn)(void *, long)) { bq_parallel_for(opt->n_threads, n, fn, p); }

bq_batch_t *bq_batch_prep(const bq_opt_t *opt, int64_t n_processed, int n, bq_read_t *seqs, int *rc_out) {
  const int pe = (opt->flag & BQ_F_PE) != 0;
  int i, max_len = 1;
  if (rc_out) *rc_out = 0;
  /* paired reads must carry the same name, or names that differ in a trailing 1 / 2 (check_paired_read_names,
   * bwamem.c:210-216, called first thing in bis_worker1): desynchronised FASTQ files must not be paired silently */
  if (pe)
    for (i = 0; i + 1 < n; i += 2) {
      const char *n1 = seqs[i].name, *n2 = seqs[i + 1].name;
      if (strcmp(n1, n2) == 0) continue;
      const size_t l = strlen(n1);
      if (l > 0 && n1[l - 1] == '1' && strlen(n2) >= l && n2[l - 1] == '2' && strncmp(n1, n2, l - 1) == 0) continue;
      bq_fatal("[check_paired_read_names] paired reads have different names: \"%s\", \"%s\"\n", n1, n2);
    }
  prep_t pp;
  memset(&pp, 0, sizeof pp);
  pp.opt = opt; pp.seqs = seqs; pp.pe = pe;
  /* clipping (bwamem.c:322,343-344) */
  prep_loop(opt, &pp, n, prep_clip);
  int n_long = 0;
  const char *first_long = 0;
  for (i = 0; i < n; ++i) {
    if (seqs[i].l_seq > BSQ_MAX_READ_LEN) { if (!n_long++) first_long = seqs[i].name; continue; }
    if (seqs[i].l_seq > max_len) max_len = seqs[i].l_seq;
  }
  /* reads beyond the device kernels' length limit get no tasks: they come out as unaligned records (their mates are
   * aligned as usual) instead of failing the batch -- and with it a run that may already have written output */
  if (n_long)
    fprintf(stderr, "[W::%s] %d read(s) longer than %d bases after clipping (first: %s) are reported unaligned: the GPU kernels take reads up to "
            "that length\n", __func__, n_long, BSQ_MAX_READ_LEN, first_long);
  /* (read, conversion) tasks in the order bis_worker1 runs them (bwamem.c:325-372) */
  bq_batch_t *b = calloc(1, sizeof *b);
  b->n = n; b->n_processed = n_processed; b->seqs = seqs;
  b->task_of_read = malloc(sizeof(int64_t) * (size_t)(n + 1));
  b->n_task = malloc((size_t)n + 1);
  uint8_t *par = malloc((size_t)n * 2 + 2);
  int64_t nt = 0;
  for (i = 0; i < n; ++i) {
    b->task_of_read[i] = nt;
    int k0 = (int)nt;
    if (seqs[i].l_seq > BSQ_MAX_READ_LEN) { b->n_task[i] = 0; continue; }
    if (!pe) {
      if (!(opt->parent & 1) || opt->parent >> 1) par[nt++] = 0;
      if (!(opt->parent & 1) || !(opt->parent >> 1)) par[nt++] = 1;
    } else if (!(i & 1)) { par[nt++] = 1; if (!opt->parent) par[nt++] = 0; }
    else { par[nt++] = 0; if (!opt->parent) par[nt++] = 1; }
    b->n_task[i] = (uint8_t)(nt - k0);
  }
  const int stride = (max_len + 15) & ~15;
  b->slot = slot_get();
  const int rc = slot_reserve(&b->slot->tseq, &b->slot->tseq_cap, (size_t)nt * stride + 16);
  if (rc) { if (rc_out) *rc_out = rc; slot_put(b->slot); free(par); free(b->task_of_read); free(b->n_task); free(b); return 0; }
  int32_t *tlen = malloc(sizeof(int32_t) * (size_t)(nt + 1));
  pp.tseq = b->slot->tseq; pp.stride = stride; pp.task_of_read = b->task_of_read; pp.n_task = b->n_task; pp.tlen = tlen;
  prep_loop(opt, &pp, n, prep_rows);
  b->reg_off = malloc(sizeof(int64_t) * (size_t)(nt + 1));
  b->nt = nt; b->stride = stride; b->tlen = tlen; b->par = par;
  return b;
}

/* GPU part: H2D of the task rows and the phase-1 kernels.  The regions stay in one of the aligner's two result slots;
 * they are copied to the host (page-locked memory) by bq_batch_fetch at the start of the host phase 2, on another
 * thread and another stream, while this lane already stages and runs the next batch. */
int bq_batch_run(bsq_aligner *al, bsq_dp *dp, bq_batch_t *b) {
  int rc;
  int64_t n_regs = 0;
  b->dp = dp; b->al = al; b->out_slot = -1;
  if (b->nt == 0) { b->reg_off[0] = 0; return 0; }
  const double t0 = getenv("BQ_TIMING") ? bq_now() : 0;
  if ((rc = bsq_aligner_stage(al, b->nt, b->slot->tseq, b->stride, b->tlen, b->par))) return rc;
  const double t1 = t0 > 0 ? bq_now() : 0;
  if ((rc = bsq_aligner_run(al, &n_regs))) return rc;
  const double t2 = t0 > 0 ? bq_now() : 0;
  if ((rc = bsq_aligner_result_slot(al, &b->out_slot, 0, 0))) return rc;
  if ((rc = slot_reserve(&b->slot->regs, &b->slot->regs_cap, (size_t)(n_regs + 1) * sizeof(bsq_reg)))) return rc;
  {
    int64_t c[16];
    if (bsq_aligner_counters(al, c, 16) == 0 && c[3])
      fprintf(stderr, "[W::%s] reads %lld..%lld: a read with more than %d SMEM intervals was left unseeded for that conversion\n", __func__,
              (long long)b->n_processed, (long long)(b->n_processed + b->n - 1), BSQ_MAX_INTV);
  }
  if (t0 > 0) {
    int64_t c[16];
    bsq_aligner_counters(al, c, 16);
    fprintf(stderr, "[bq_batch_run] stage %.4f run %.4f (kernels %.4f) s\n", t1 - t0, t2 - t1, c[10] * 1e-6);
  }
  return 0;
}

/* D2H of the regions of the batch from its result slot (a no-op when there is nothing to fetch or it is done) */
static int bq_batch_fetch(bq_batch_t *b) {
  if (b->dregs || b->nt == 0 || b->out_slot < 0) return 0;
  const int rc = bsq_aligner_fetch_slot(b->al, b->out_slot, b->slot->regs, b->reg_off);
  if (rc) return rc;
  b->dregs = b->slot->regs;
  return 0;
}

static void batch_free(bq_batch_t *b) {
  if (b->al && b->out_slot >= 0 && !b->dregs && b->nt > 0) bsq_aligner_release_slot(b->al, b->out_slot); /* never fetched (a failed batch) */
  slot_put(b->slot);
  free(b->task_of_read); free(b->n_task); free(b->reg_off); free(b->tlen); free(b->par);
  free(b);
}

void bq_batch_discard(bq_batch_t *b) { if (b) batch_free(b); }

bq_batch_t *bq_batch_gpu(const bq_opt_t *opt, bsq_aligner *al, bsq_dp *dp, int64_t n_processed, int n, bq_read_t *seqs, int *rc_out) {
  int rc = 0;
  bq_batch_t *b = bq_batch_prep(opt, n_processed, n, seqs, &rc);
  if (b && (rc = bq_batch_run(al, dp, b))) { batch_free(b); b = 0; }
  if (rc_out) *rc_out = rc;
  return b;
}

/* the region pool of the last batch is kept for the next one (its pages are already mapped) */
static pthread_mutex_t g_pool_mu = PTHREAD_MUTEX_INITIALIZER;
static bq_reg_t *g_pool;
static size_t g_pool_cap;
static bq_reg_t *pool_take(size_t n, size_t *cap) {
  bq_reg_t *p = 0;
  pthread_mutex_lock(&g_pool_mu);
  if (g_pool && g_pool_cap >= n) { p = g_pool; *cap = g_pool_cap; g_pool = 0; g_pool_cap = 0; }
  pthread_mutex_unlock(&g_pool_mu);
  if (!p) { *cap = n + (n >> 3); p = aligned_alloc(64, *cap * sizeof(bq_reg_t)); } /* regions on cache-line boundaries (128 bytes each) */
  return p;
}
static void pool_give(bq_reg_t *p, size_t cap) {
  pthread_mutex_lock(&g_pool_mu);
  if (!g_pool || g_pool_cap < cap) { bq_reg_t *old = g_pool; g_pool = p; g_pool_cap = cap; p = old; }
  pthread_mutex_unlock(&g_pool_mu);
  free(p);
}

/* what the first half of phase 2 leaves for the second */
typedef struct bq_fin {
  work_t w;
  size_t pool_cap;
  int64_t *pool_off;
  int64_t n_cjobs;
  int cig_submitted;
  double t0, t_a;
} bq_fin_t;

static int64_t prefix_counts(int64_t *first, int n_items) { /* first[i + 1] holds the count of item i on entry */
  first[0] = 0;
  for (int i = 0; i < n_items; ++i) first[i + 1] += first[i];
  return first[n_items];
}

int bq_batch_finish_a(const bq_opt_t *opt, const bq_ref_t *ref, bq_batch_t *b, const bq_pestat_t *pes0) {
  const int pe = (opt->flag & BQ_F_PE) != 0, n = b->n;
  int rc = 0;
  const double t_in_ = getenv("BQ_TIMING") ? bq_now() : 0;
  if ((rc = bq_batch_fetch(b))) return rc;
  bq_fin_t *f = calloc(1, sizeof *f);
  work_t *w = &f->w;
  b->fin = f;
  w->opt = opt; w->ref = ref; w->seqs = b->seqs; w->n_processed = b->n_processed; w->n_threads = opt->n_threads; w->pe = pe;
  w->use_dp = b->dp != 0 && b->nt > 0;
  size_t regs_cap_ = 0;
  w->regs = bq_big_alloc(((size_t)n + 1) * sizeof(bq_regv_t), &regs_cap_);
  memset(w->regs, 0, ((size_t)n + 1) * sizeof(bq_regv_t));
  w->dev_regs = b->dregs; w->reg_off = b->reg_off; w->task_of_read = b->task_of_read; w->n_task_of_read = b->n_task;
  /* one pool for the host regions of the whole batch instead of one allocation per read (the per-read vectors were
   * allocated by one worker and freed by another, which glibc's per-thread caches cannot serve) */
  int64_t *pool_off = malloc(sizeof(int64_t) * (size_t)(n + 1));
  pool_off[0] = 0;
  for (int i = 0; i < n; ++i) {
    int64_t tot = 2;
    for (int t = 0; t < b->n_task[i]; ++t) tot += b->reg_off[b->task_of_read[i] + t + 1] - b->reg_off[b->task_of_read[i] + t];
    pool_off[i + 1] = pool_off[i] + tot;
  }
  w->reg_pool = pool_take((size_t)pool_off[n] + 1, &f->pool_cap);
  w->pool_off = f->pool_off = pool_off;
  f->t0 = getenv("BQ_TIMING") ? bq_now() : 0;
  /* the read rows go to the DP context while the regions are merged (asynchronous copy from the page-locked rows) */
  if (w->use_dp && (rc = bsq_dp_set_reads(b->dp, b->nt, b->slot->tseq, b->stride, b->tlen))) return rc;
  w->stage = ST_MERGE;
  run_threads(w, n);
  const double tm_ = f->t0 > 0 ? bq_now() : 0;
  if (pe) {
    if (pes0) w->pes = *pes0;
    else { /* mem_pestat (mem_pair.c:74-138): candidates collected on the workers, then the serial sort + moments */
      w->pes_is = malloc(sizeof(int64_t) * (size_t)(n / 2 + 1));
      w->stage = ST_PESTAT;
      run_threads(w, n >> 1);
      size_t n_is = 0;
      for (int i = 0; i < n >> 1; ++i) if (w->pes_is[i] != PES_NONE) w->pes_is[n_is++] = w->pes_is[i];
      w->pes = pestat_finish(opt, w->pes_is, n_is); /* frees the array */
      w->pes_is = 0;
    }
  }
  const double tp_ = f->t0 > 0 ? bq_now() : 0;
  if (g_prof < 0) g_prof = getenv("BQ_PROF") != 0;
  g_t_mate = g_t_mark = g_t_sam = g_t_pair = g_t_setsam = g_t_fmt = 0;
  const int n_items = pe ? n >> 1 : n;
  int64_t n_mjobs = 0;
  const int plan = w->use_dp && pe && !(opt->flag & BQ_F_NO_RESCUE);
  if (plan) { /* mate rescue: plan (pairs with nothing to try are finished in the same pass), one kernel, replay (ST_MARK) */
    w->ms_first = calloc((size_t)n_items + 2, sizeof(int64_t));
    w->stage = ST_PLAN;
    run_threads(w, n_items);
    n_mjobs = prefix_counts(w->ms_first, n_items);
    if (n_mjobs > 0) {
      bq_slot_t *sl = b->slot;
      if ((rc = slot_reserve(&sl->mjobs, &sl->mjobs_cap, (size_t)n_mjobs * sizeof(bsq_matesw_job))) ||
          (rc = slot_reserve(&sl->mres, &sl->mres_cap, (size_t)n_mjobs * sizeof(bsq_matesw_res))))
        return rc;
      w->mjobs = sl->mjobs; w->mres = sl->mres;
      w->mkeys = malloc(sizeof(uint32_t) * (size_t)n_mjobs);
      w->stage = ST_MS_FILL;
      run_threads(w, n_items);
      if ((rc = bsq_dp_matesw_submit(b->dp, n_mjobs, w->mjobs, w->mres)) || (rc = bsq_dp_matesw_wait(b->dp))) return rc;
      __atomic_fetch_add(&g_dp_stat[3], n_mjobs, __ATOMIC_RELAXED);
    }
  }
  const double tr_ = f->t0 > 0 ? bq_now() : 0;
  double ts_ = 0;
  if (!plan || n_mjobs > 0) { /* what ST_PLAN could not finish: the pairs with rescue attempts -- or, without planning, every item */
    w->stage = ST_MARK;
    run_threads(w, n_items);
  }
  if (w->use_dp) { /* the final CIGARs of the batch as one asynchronous kernel; _b picks them up */
    w->thr_base[0] = 0;
    for (int t = 0; t < 256; ++t) w->thr_base[t + 1] = w->thr_base[t] + w->tjobs[t].n;
    f->n_cjobs = w->thr_base[256];
    bq_slot_t *sl = b->slot;
    if ((rc = slot_reserve(&sl->cjobs, &sl->cjobs_cap, (size_t)(f->n_cjobs + 1) * sizeof(bsq_cigar_job))) ||
        (rc = slot_reserve(&sl->cres, &sl->cres_cap, (size_t)(f->n_cjobs + 1) * sizeof(bsq_cigar_res))))
      return rc;
    w->cjobs = sl->cjobs; w->cres = sl->cres;
    for (int t = 0; t < 256; ++t) { /* the threads' lists, one behind the other, into the page-locked job array */
      if (w->tjobs[t].n) memcpy(w->cjobs + w->thr_base[t], w->tjobs[t].a, (size_t)w->tjobs[t].n * sizeof(bsq_cigar_job));
      free(w->tjobs[t].a); w->tjobs[t].a = 0; w->tjobs[t].m = 0;
    }
    ts_ = f->t0 > 0 ? bq_now() : 0;
    if ((rc = bsq_dp_cigar_submit(b->dp, f->n_cjobs, w->cjobs, w->cres))) return rc;
    ts_ = f->t0 > 0 ? bq_now() - ts_ : 0;
    f->cig_submitted = 1;
    __atomic_fetch_add(&g_dp_stat[0], f->n_cjobs, __ATOMIC_RELAXED);
  }
  if (f->t0 > 0) {
    f->t_a = bq_now() - f->t0;
    fprintf(stderr, "[bq_finish_a] merge %.3f pestat %.3f rescue plan + primary marking %.3f s (%lld rescue jobs) replay + job list %.3f s, cigar jobs %lld (submit %.3f s); fetch + set-up %.3f, total %.3f s\n", tm_ - f->t0, tp_ - tm_,
            tr_ - tp_, (long long)n_mjobs, bq_now() - tr_ - ts_, (long long)f->n_cjobs, ts_, f->t0 - t_in_, bq_now() - t_in_);
  }
  return 0;
}

int bq_batch_finish_wait(bq_batch_t *b) {
  bq_fin_t *f = b->fin;
  if (!f || !f->cig_submitted) return 0;
  f->cig_submitted = 0;
  const uint32_t *blob = 0;
  int64_t words = 0;
  const int rc = bsq_dp_cigar_wait(b->dp, &blob, &words);
  if (rc) return rc;
  f->w.cig.res = f->w.cres; f->w.cig.blob = blob; f->w.cig.n = f->n_cjobs; f->w.cig.thr_base = f->w.thr_base; f->w.cig.n_thr = 256;
  return 0;
}

void bq_batch_finish_b(const bq_opt_t *opt, const bq_ref_t *ref, bq_batch_t *b, const char *rg_id) {
  bq_fin_t *f = b->fin;
  work_t *w = &f->w;
  const int pe = w->pe, n = b->n;
  const double t1_ = f->t0 > 0 ? bq_now() : 0;
  w->rg_id = rg_id;
  {
    const int nt = opt->n_threads < 1 ? 1 : (opt->n_threads > 255 ? 255 : opt->n_threads);
    w->sam_slab = calloc((size_t)nt, sizeof(bq_str_t));
    w->sam_off = malloc(sizeof(size_t) * (size_t)(n + 1));
    w->sam_thr = malloc((size_t)n + 1);
  }
  if (pe && !w->pes.failed && w->pes.high >= w->pes.low && w->pes.high - w->pes.low < (1 << 16) && w->pes.std > 0) {
    const int nt_ = w->pes.high - w->pes.low + 1;
    w->pair_term = malloc(sizeof(double) * (size_t)nt_);
    for (int k = 0; k < nt_; ++k) w->pair_term[k] = pair_term(opt, &w->pes, (int64_t)w->pes.low + k);
    w->pair_tab.low = w->pes.low; w->pair_tab.high = w->pes.high; w->pair_tab.term = w->pair_term;
  }
  w->stage = ST_SAM;
  run_threads(w, pe ? n >> 1 : n);
  free(w->pair_term); w->pair_term = 0;
  if (n > 0) { /* .sam pointers into the (now final) slabs; the slabs belong to the first read of the batch */
    const int nt = w->n_threads;
    for (int i = 0; i < n; ++i)
      if (b->seqs[i].sam_in_slab) b->seqs[i].sam = w->sam_slab[w->sam_thr[i]].s + w->sam_off[i];
    b->seqs[0].sam_slabs = malloc(sizeof(char *) * (size_t)nt);
    b->seqs[0].n_sam_slabs = nt;
    for (int k = 0; k < nt; ++k) b->seqs[0].sam_slabs[k] = w->sam_slab[k].s;
  }
  free(w->sam_slab); free(w->sam_off); free(w->sam_thr);
  const double t2_ = f->t0 > 0 ? bq_now() : 0;
  if (g_prof > 0) fprintf(stderr, "[bq_prof] thread-seconds: matesw %.3f mark_primary %.3f reg2sam %.3f (pair %.3f set_sam %.3f format %.3f)\n",
                          g_t_mate, g_t_mark, g_t_sam, g_t_pair, g_t_setsam, g_t_fmt);
  bq_big_free(w->regs);
  pool_give(w->reg_pool, f->pool_cap);
  free(f->pool_off); free(w->ms_first); free(w->mkeys);
  if (f->t0 > 0) fprintf(stderr, "[bq_finish_b] pairing + SAM %.3f s, total %.3f s\n", t2_ - t1_, bq_now() - t1_);
  free(f);
  b->fin = 0;
  batch_free(b);
}

static void fin_abandon(bq_batch_t *b) { /* a DP call failed: release what _a built (the reads stay with the caller) */
  bq_fin_t *f = b->fin;
  if (f) {
    if (f->cig_submitted) { const uint32_t *bl; int64_t wd; bsq_dp_cigar_wait(b->dp, &bl, &wd); }
    work_t *w = &f->w;
    for (int r = 0; r < b->n; ++r) if (w->regs[r].a && !w->regs[r].pooled) free(w->regs[r].a);
    bq_big_free(w->regs);
    pool_give(w->reg_pool, f->pool_cap);
    free(f->pool_off); free(w->ms_first); free(w->mkeys);
    for (int t = 0; t < 256; ++t) free(w->tjobs[t].a);
    free(f);
    b->fin = 0;
  }
  batch_free(b);
}
void bq_batch_abandon(bq_batch_t *b) { if (b) fin_abandon(b); }

int bq_batch_finish(const bq_opt_t *opt, const bq_ref_t *ref, bq_batch_t *b, const bq_pestat_t *pes0, const char *rg_id) {
  int rc = bq_batch_finish_a(opt, ref, b, pes0);
  if (!rc) rc = bq_batch_finish_wait(b);
  if (rc) { fin_abandon(b); return rc; }
  bq_batch_finish_b(opt, ref, b, rg_id);
  return 0;
}

int bq_process_seqs(const bq_opt_t *opt, bsq_aligner *al, bsq_dp *dp, const bq_ref_t *ref, int64_t n_processed, int n, bq_read_t *seqs,
                    const bq_pestat_t *pes0, const char *rg_id) {
  int rc = 0;
  bq_batch_t *b = bq_batch_gpu(opt, al, dp, n_processed, n, seqs, &rc);
  if (!b) return rc ? rc : BSQ_ENOMEM;
  return bq_batch_finish(opt, ref, b, pes0, rg_id);
}
