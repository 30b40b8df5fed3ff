/* bq_pileup.c -- `biscuit pileup` (reference: src/pileup.c).
 *
 * main_pileup (src/pileup.c:1014-1225) dispatches 100 kb windows to threads that walk the BAM through
 * htslib iterators.  Here the GPU does the per-read / per-locus integer work for a chunk of windows at
 * a time (bsq_plp_*), and the host
 *   - streams each coordinate-sorted BAM front to back per contig (bq_bam.c), contigs in name order,
 *   - computes the genotype likelihoods and formats the VCF text (plp_format, src/pileup.c:415-640),
 *   - keeps the per-window methylation sums and adds them in window order, as write_func does
 *     (src/pileup.c:145-234), so <out>_meth_average.tsv is reproduced digit for digit.
 *
 * Genotype likelihoods: the reference takes genotype_lnlik / ln_sum3 / pval2qual from huishenlab/utils
 * stats.h @5f4aeab, which is not vendored under /root/reference and cannot be fetched.  gt_lnlik(),
 * ln_sum3() and pval2qual() below restate them as plain binomial likelihoods -- PARITY UNPINNED for the VCF
 * fields QUAL, FILTER, GT, GL1, GQ.  Every other field is a function of code present in the
 * reference tree and of the integer counts.
 */
#include <errno.h>
#include <getopt.h>
#include <libgen.h>
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>
#include "bq_plp.h"

enum { B_A = 0, B_C, B_G, B_T, B_N, B_Y, B_R };
enum { M_RET = 0, M_CONV = 1 };
static const char basecode[7] = "ACGTNYR";                                                        /* bisc_utils.c:28 */
static const char *ctx_name[7] = {"CG", "CHG", "CHH", "CG", "CHG", "CHH", "CN"};                  /* bisc_utils.c:29 */
static const char *ctx_name_nome[7] = {"HCG", "HCHG", "HCHH", "GCG", "GCH", "GCH", "CN"};         /* bisc_utils.c:30 */

/* ---- genotype math (see the header: restated, parity unpinned) ---- */
static double gt_lnlik(int gt, int n_ref, int n_alt, double error, double contam) {
  double p_alt; /* probability that a read shows the alternative allele */
  if (gt == 0) p_alt = error + contam;
  else if (gt == 1) p_alt = 0.5;
  else p_alt = 1.0 - error - contam;
  if (p_alt < 1e-300) p_alt = 1e-300;
  if (p_alt > 1.0 - 1e-16) p_alt = 1.0 - 1e-16;
  return n_alt * log(p_alt) + n_ref * log(1.0 - p_alt);
}

static double ln_sum3(double a, double b, double c) {
  double m = a > b ? a : b;
  if (c > m) m = c;
  return m + log(exp(a - m) + exp(b - m) + exp(c - m));
}

static double pval2qual(double pval) {
  if (pval <= 0) return 1000;
  double q = -10.0 * log10(pval);
  return q > 1000 ? 1000 : q;
}

typedef struct { char gt[4]; double gl0, gl1, gl2, gq; } gt_call_t;

/* pileup_genotype, src/pileup.c:389-413 */
static void genotype(const bq_plp_fmt_t *cf, int cref, int altsupp, gt_call_t *g) {
  g->gl0 = log(cf->prior0) + gt_lnlik(0, cref, altsupp, cf->error, cf->contam);
  g->gl1 = log(cf->prior1) + gt_lnlik(1, cref, altsupp, cf->error, cf->contam);
  g->gl2 = log(cf->prior2) + gt_lnlik(2, cref, altsupp, cf->error, cf->contam);
  const double s = ln_sum3(g->gl0, g->gl1, g->gl2);
  if (g->gl0 > g->gl1) {
    if (g->gl0 > g->gl2) { g->gq = pval2qual(1 - exp(g->gl0 - s)); strcpy(g->gt, "0/0"); }
    else { g->gq = pval2qual(1 - exp(g->gl2 - s)); strcpy(g->gt, "1/1"); }
  } else if (g->gl1 > g->gl2) { g->gq = pval2qual(1 - exp(g->gl1 - s)); strcpy(g->gt, "0/1"); }
  else { g->gq = pval2qual(1 - exp(g->gl2 - s)); strcpy(g->gt, "1/1"); }
}

/* ---- text tables: the same few small-integer arguments recur billions of times ---- */
#define GT_TAB 96   /* (nref, nalt) < GT_TAB: cached "\tGT:gl0,gl1,gl2:gq" text + gq */
#define FR_TAB 256  /* k/n with n < FR_TAB: cached "%1.3f" and "%1.2f" text */
typedef struct {
  bq_plp_fmt_t cf;
  struct { char s[40]; uint8_t l, set; double gq; } *gt;
  char (*f3)[6]; /* [n*FR_TAB+k] = "%1.3f" of k/n */
  char (*f2)[5];
  uint8_t *f_set;
} fmt_tab_t;

static fmt_tab_t *tab_new(const bq_plp_fmt_t *cf) {
  fmt_tab_t *t = calloc(1, sizeof *t);
  t->cf = *cf;
  t->gt = calloc((size_t)GT_TAB * GT_TAB, sizeof *t->gt);
  t->f3 = calloc((size_t)FR_TAB * FR_TAB, 6);
  t->f2 = calloc((size_t)FR_TAB * FR_TAB, 5);
  t->f_set = calloc((size_t)FR_TAB * FR_TAB, 1);
  return t;
}

static void tab_free(fmt_tab_t *t) { free(t->gt); free(t->f3); free(t->f2); free(t->f_set); free(t); }

static inline void put_uint(bq_str_t *s, uint32_t v) {
  char b[12];
  int n = 0;
  do { b[n++] = (char)('0' + v % 10); v /= 10; } while (v);
  char *o = s->s + s->l;
  while (n) *o++ = b[--n];
  s->l = (size_t)(o - s->s);
}

static inline void put_mem(bq_str_t *s, const char *p, size_t n) { memcpy(s->s + s->l, p, n); s->l += n; }
#define PUTS(s, lit) put_mem((s), (lit), sizeof(lit) - 1)

static inline void put_frac(fmt_tab_t *t, bq_str_t *s, int k, int n, int digits3) {
  if (n < FR_TAB) {
    const size_t i = (size_t)n * FR_TAB + (size_t)k;
    if (!t->f_set[i]) {
      char b[32];
      snprintf(b, sizeof b, "%1.3f", k / (double)n); memcpy(t->f3[i], b, 5); t->f3[i][5] = 0;
      snprintf(b, sizeof b, "%1.2f", k / (double)n); memcpy(t->f2[i], b, 4); t->f2[i][4] = 0;
      t->f_set[i] = 1;
    }
    if (digits3) put_mem(s, t->f3[i], 5); else put_mem(s, t->f2[i], 4);
  } else s->l += (size_t)sprintf(s->s + s->l, digits3 ? "%1.3f" : "%1.2f", k / (double)n);
}

/* one locus (n_bams records) -> one VCF line; returns gq-independent pieces through the tables */
static void format_locus(fmt_tab_t *t, const char *chrm, size_t l_chrm, const bsq_plp_rec *r, bq_str_t *s, double *wbeta, int64_t *wcnt) {
  const bq_plp_fmt_t *cf = &t->cf;
  const int nb = cf->n_bams, rb_code = r[0].rb_code, cm1 = r[0].cm1, ctt = r[0].ctx;
  const char rb = basecode[rb_code];
  bq_str_reserve(s, l_chrm + 160 + (size_t)nb * 200);
  /* genotype each sample (src/pileup.c:466-502) */
  gt_call_t gts[8];
  double gq[8], lowest_gq = 0;
  const char *gtxt[8];
  int gtl[8];
  char gbuf[8][64];
  for (int sid = 0; sid < nb; ++sid) {
    const int nref = r[sid].base_redist[rb_code], nalt = cm1 >= 0 ? r[sid].base_redist[cm1] : 0;
    gq[sid] = 0; gtxt[sid] = 0; gtl[sid] = 0;
    if (nref + nalt > 0) {
      if (nref < GT_TAB && nalt < GT_TAB) {
        const size_t i = (size_t)nref * GT_TAB + (size_t)nalt;
        if (!t->gt[i].set) {
          genotype(cf, nref, nalt, &gts[sid]);
          t->gt[i].gq = gts[sid].gq;
          t->gt[i].l = (uint8_t)snprintf(t->gt[i].s, sizeof t->gt[i].s, "\t%s:%1.0f,%1.0f,%1.0f:%1.0f", gts[sid].gt,
                                         gts[sid].gl0 > -1000 ? gts[sid].gl0 : -1000.0, gts[sid].gl1 > -1000 ? gts[sid].gl1 : -1000.0,
                                         gts[sid].gl2 > -1000 ? gts[sid].gl2 : -1000.0, gts[sid].gq);
          t->gt[i].set = 1;
        }
        gq[sid] = t->gt[i].gq; gtxt[sid] = t->gt[i].s; gtl[sid] = t->gt[i].l;
      } else {
        genotype(cf, nref, nalt, &gts[sid]);
        gq[sid] = gts[sid].gq;
        gtl[sid] = snprintf(gbuf[sid], sizeof gbuf[sid], "\t%s:%1.0f,%1.0f,%1.0f:%1.0f", gts[sid].gt,
                            gts[sid].gl0 > -1000 ? gts[sid].gl0 : -1000.0, gts[sid].gl1 > -1000 ? gts[sid].gl1 : -1000.0,
                            gts[sid].gl2 > -1000 ? gts[sid].gl2 : -1000.0, gts[sid].gq);
        gtxt[sid] = gbuf[sid];
      }
    }
    if (gq[sid] < lowest_gq || !sid) lowest_gq = gq[sid];
  }
  /* CHROM POS ID REF ALT */
  put_mem(s, chrm, l_chrm);
  s->s[s->l++] = '\t';
  put_uint(s, (uint32_t)r[0].pos);
  PUTS(s, "\t.\t");
  s->s[s->l++] = rb; s->s[s->l++] = '\t';
  if (cm1 >= 0) s->s[s->l++] = (cm1 == B_Y || cm1 == B_R) ? 'N' : basecode[cm1];
  else s->s[s->l++] = '.';
  /* QUAL FILTER */
  s->s[s->l++] = '\t';
  {
    int q = (int)lowest_gq;
    if (q < 0) { s->s[s->l++] = '-'; q = -q; }
    put_uint(s, (uint32_t)q);
  }
  if (lowest_gq > 5) PUTS(s, "\tPASS\t"); else PUTS(s, "\tLowQual\t");
  /* INFO */
  PUTS(s, "NS=");
  put_uint(s, (uint32_t)nb);
  if (rb == 'C' || rb == 'G') {
    PUTS(s, ";CX=");
    const char *cx = cf->is_nome ? ctx_name_nome[ctt] : ctx_name[ctt];
    put_mem(s, cx, strlen(cx));
    PUTS(s, ";N5=");
    put_mem(s, r[0].n5, 5);
  }
  if (cm1 == B_Y || cm1 == B_R) { PUTS(s, ";AB="); s->s[s->l++] = basecode[cm1]; }
  /* FORMAT */
  PUTS(s, "\tGT:GL1:GQ:DP:SP");
  if (cm1 >= 0) PUTS(s, ":AC:AF1");
  const int any_callable = r[0].any_callable;
  if (any_callable) PUTS(s, ":CV:BT");
  for (int sid = 0; sid < nb; ++sid) {
    const bsq_plp_rec *q = r + sid;
    if (gq[sid] > 0 && q->dp) put_mem(s, gtxt[sid], (size_t)gtl[sid]);
    else PUTS(s, "\t./.:.,.,.:0");
    s->s[s->l++] = ':';
    put_uint(s, (uint32_t)q->dp);
    /* SP */
    s->s[s->l++] = ':';
    int added = 0;
    if (q->base[rb_code]) { s->s[s->l++] = rb; put_uint(s, (uint32_t)q->base[rb_code]); added = 1; }
    for (int i = 0; i < 7; ++i) {
      if (i == B_N || i == rb_code || q->base[i] <= 0) continue;
      s->s[s->l++] = basecode[i]; put_uint(s, (uint32_t)q->base[i]); added = 1;
    }
    if (!added) s->s[s->l++] = '.';
    /* AC AF1 */
    if (cm1 >= 0) {
      const int nref = q->base_redist[rb_code], nalt = q->base_redist[cm1];
      s->s[s->l++] = ':';
      put_uint(s, (uint32_t)(nref + nalt));
      s->s[s->l++] = ':';
      if (nref + nalt) put_frac(t, s, nalt, nref + nalt, 0);
      else s->s[s->l++] = '.';
    }
    /* CV BT */
    if (any_callable) {
      if (q->methcallable) {
        const int cv = q->meth[M_RET] + q->meth[M_CONV];
        if (ctt != 6) {
          wbeta[sid * BQ_NCTX + ctt] += (double)q->meth[M_RET] / (double)cv;
          wcnt[sid * BQ_NCTX + ctt]++;
        }
        s->s[s->l++] = ':';
        put_uint(s, (uint32_t)cv);
        s->s[s->l++] = ':';
        put_frac(t, s, q->meth[M_RET], cv, 1);
      } else PUTS(s, ":0:.");
    }
  }
  s->s[s->l++] = '\n';
  s->s[s->l] = 0;
}

typedef struct {
  fmt_tab_t *tab;
  const char *chrm;
  const bsq_plp_rec *recs;
  int64_t lo, hi, w0, step;
  bq_str_t out;
  double *wbeta;
  int64_t *wcnt;
} fmt_job_t;

static void *fmt_worker(void *arg) {
  fmt_job_t *j = arg;
  const int nb = j->tab->cf.n_bams;
  const size_t l_chrm = strlen(j->chrm);
  for (int64_t i = j->lo; i < j->hi; ++i) {
    const bsq_plp_rec *r = j->recs + i * nb;
    const int64_t w = (r->pos - j->w0) / j->step;
    format_locus(j->tab, j->chrm, l_chrm, r, &j->out, j->wbeta + w * nb * BQ_NCTX, j->wcnt + w * nb * BQ_NCTX);
  }
  return 0;
}

void bq_plp_format(const bq_plp_fmt_t *cf, const char *chrm, const bsq_plp_rec *recs, int64_t n_loci, int64_t w0, int64_t step, int n_win,
                   bq_str_t *out, double *wbeta, int64_t *wcnt) {
  if (n_loci <= 0) return;
  int nt = cf->n_threads < 1 ? 1 : cf->n_threads > 64 ? 64 : cf->n_threads;
  const int nb = cf->n_bams;
  /* thread ranges are cut at window boundaries: a window's methylation sum is accumulated by one thread, in
   * locus order, exactly as plp_format does within process_func */
  int64_t cut[65];
  cut[0] = 0;
  int n_rng = 0;
  for (int t = 1; t <= nt; ++t) {
    int64_t target = n_loci * t / nt;
    if (t < nt) {
      if (target >= n_loci) target = n_loci;
      else { /* advance to the first locus of the next window */
        const int64_t w = (recs[target * nb].pos - w0) / step;
        int64_t lo = target, hi = n_loci;
        while (lo < hi) { int64_t m = (lo + hi) >> 1; if ((recs[m * nb].pos - w0) / step <= w) lo = m + 1; else hi = m; }
        target = lo;
      }
    } else target = n_loci;
    if (target > cut[n_rng]) cut[++n_rng] = target;
  }
  fmt_job_t jobs[64];
  pthread_t th[64];
  fmt_tab_t *tabs[64];
  for (int t = 0; t < n_rng; ++t) {
    tabs[t] = tab_new(cf);
    memset(&jobs[t], 0, sizeof jobs[t]);
    jobs[t].tab = tabs[t]; jobs[t].chrm = chrm; jobs[t].recs = recs; jobs[t].lo = cut[t]; jobs[t].hi = cut[t + 1]; jobs[t].w0 = w0;
    jobs[t].step = step; jobs[t].wbeta = wbeta; jobs[t].wcnt = wcnt;
  }
  (void)n_win;
  for (int t = 1; t < n_rng; ++t) pthread_create(&th[t], 0, fmt_worker, &jobs[t]);
  fmt_worker(&jobs[0]);
  for (int t = 1; t < n_rng; ++t) pthread_join(th[t], 0);
  for (int t = 0; t < n_rng; ++t) {
    if (jobs[t].out.l) bq_kputsn(out, jobs[t].out.s, jobs[t].out.l);
    free(jobs[t].out.s);
    tab_free(tabs[t]);
  }
}

/* ------------------------------------------------------------------ command line ---- */

typedef struct {
  int n_threads, step, is_nome, somatic, verbose;
  bsq_plp_conf filt;
  double error, mu, mu_somatic, contam, prior0, prior1, prior2;
} plp_conf_t;

static void conf_init(plp_conf_t *c) { /* pileup_conf_init, src/pileup.c:944-963; bisc_utils.h:40-113 */
  memset(c, 0, sizeof *c);
  c->n_threads = 3; c->step = 100000;
  bsq_plp_conf_default(&c->filt);
  c->error = 0.001; c->mu = 0.001; c->mu_somatic = 0.001; c->contam = 0.01; c->prior1 = 0.33333; c->prior2 = 0.33333;
  c->prior0 = 1.0 - c->prior1 - c->prior2;
}

static int usage(const plp_conf_t *conf) { /* src/pileup.c:965-1012 */
  fprintf(stderr, "\n");
  fprintf(stderr, "Usage: biscuit pileup [options] <ref.fa> <in1.bam> [in2.bam in3.bam ...]\n");
  fprintf(stderr, "Som. Mode Usage: biscuit pileup [options] <-S -T tum.bam -I norm.bam> <ref.fa>\n");
  fprintf(stderr, "\n");
  fprintf(stderr, "Options:\n");
  fprintf(stderr, "    -g STR      Region (optional, will process the whole bam if not specified)\n");
  fprintf(stderr, "    -@ INT      Number of threads [%d]\n", conf->n_threads);
  fprintf(stderr, "    -s INT      Step of window dispatching [%d]\n", conf->step);
  fprintf(stderr, "    -N          NOMe-seq mode [off]\n");
  fprintf(stderr, "    -S          Somatic mode, must provide -T and -I arguments [off]\n");
  fprintf(stderr, "    -T STR      Somatic mode, tumor BAM\n");
  fprintf(stderr, "    -I STR      Somatic mode, normal BAM\n");
  fprintf(stderr, "\n");
  fprintf(stderr, "Output options:\n");
  fprintf(stderr, "    -o STR      Output file [stdout]\n");
  fprintf(stderr, "    -w STR      Pileup statistics output prefix [same as output]\n");
  fprintf(stderr, "    -v INT      Verbosity level (0: no added info printed, 0<INT<=5: print\n");
  fprintf(stderr, "                    diagnostic info, INT>5: print diagnostic and debug info) [0]\n");
  fprintf(stderr, "\n");
  fprintf(stderr, "Filter options:\n");
  fprintf(stderr, "    -b INT      Minimum base quality [%u]\n", conf->filt.min_base_qual);
  fprintf(stderr, "    -m INT      Minimum mapping quality [%u]\n", conf->filt.min_mapq);
  fprintf(stderr, "    -a INT      Minimum alignment score (from AS-tag) [%u]\n", conf->filt.min_score);
  fprintf(stderr, "    -t INT      Maximum cytosine retention in a read [%u]\n", conf->filt.max_retention);
  fprintf(stderr, "    -l INT      Minimum read length [%u]\n", conf->filt.min_read_len);
  fprintf(stderr, "    -5 INT      Minimum distance to 5' end of a read [%u]\n", conf->filt.min_dist_end_5p);
  fprintf(stderr, "    -3 INT      Minimum distance to 3' end of a read [%u]\n", conf->filt.min_dist_end_3p);
  fprintf(stderr, "    -r          NO redistribution of ambiguous (Y/R) calls in SNP genotyping\n");
  fprintf(stderr, "    -c          NO filtering secondary mapping\n");
  fprintf(stderr, "    -d          Double count cytosines in overlapping mate reads (avoided\n");
  fprintf(stderr, "                    by default)\n");
  fprintf(stderr, "    -u          NO filtering of duplicate flagged reads\n");
  fprintf(stderr, "    -p          NO filtering of improper pair flagged reads\n");
  fprintf(stderr, "    -n INT      Maximum NM tag [%d]\n", conf->filt.max_nm);
  fprintf(stderr, "\n");
  fprintf(stderr, "Genotyping options:\n");
  fprintf(stderr, "    -E FLOAT    Error rate [%1.3f]\n", conf->error);
  fprintf(stderr, "    -M FLOAT    Mutation rate [%1.3f]\n", conf->mu);
  fprintf(stderr, "    -x FLOAT    Somatic mutation rate [%1.3f]\n", conf->mu_somatic);
  fprintf(stderr, "    -C FLOAT    Contamination rate [%1.3f]\n", conf->contam);
  fprintf(stderr, "    -P FLOAT    Prior probability for heterozygous variant [%1.3f]\n", conf->prior1);
  fprintf(stderr, "    -Q FLOAT    Prior probability for homozygous variant [%1.3f]\n", conf->prior2);
  fprintf(stderr, "    -h          This help\n");
  fprintf(stderr, "\n");
  return 1;
}

typedef struct { int tid; char *name; int32_t len; } target_t;
static int cmp_target(const void *a, const void *b) { return strcmp(((const target_t *)a)->name, ((const target_t *)b)->name); }

/* print_vcf_header, src/pileup.c:874-942 (the verbose-only FORMAT lines are not produced: -v > 0 is rejected) */
static void vcf_header(bq_str_t *h, const char *reffn, const target_t *targets, int n_targets, char **argv, int argc, const plp_conf_t *conf,
                       char **in_fns, int n_fns) {
  char tmp[64];
  bq_kputs(h, "##fileformat=VCFv4.1\n");
  bq_kputs(h, "##reference="); bq_kputs(h, reffn); bq_kputc(h, '\n');
  bq_kputs(h, "##source=biscuitV" BQ_VERSION "\n");
  for (int j = 0; j < n_targets; ++j) {
    bq_kputs(h, "##contig=<ID="); bq_kputs(h, targets[j].name);
    snprintf(tmp, sizeof tmp, ",length=%d>\n", targets[j].len); bq_kputs(h, tmp);
  }
  bq_kputs(h, "##program=<cmd=biscuit");
  for (int i = 0; i < argc; ++i) { bq_kputc(h, ' '); bq_kputs(h, argv[i]); }
  bq_kputs(h, ">\n");
  bq_kputs(h, "##FILTER=<ID=PASS,Description=\"All filters passed\">\n");
  bq_kputs(h, "##FILTER=<ID=LowQual,Description=\"Genotype quality smaller than 5\">\n");
  bq_kputs(h, "##INFO=<ID=NS,Number=1,Type=Integer,Description=\"Number of samples with data\">\n");
  if (conf->is_nome) bq_kputs(h, "##INFO=<ID=CX,Number=1,Type=String,Description=\"Cytosine context (HCG, HCHG, HCHH, GCG, GCH)\">\n");
  else bq_kputs(h, "##INFO=<ID=CX,Number=1,Type=String,Description=\"Cytosine context (CG, CHH or CHG)\">\n");
  bq_kputs(h, "##INFO=<ID=N5,Number=1,Type=String,Description=\"5-nucleotide context, centered around target cytosine\">\n");
  bq_kputs(h, "##INFO=<ID=AB,Number=A,Type=String,Description=\"When true alt-allele is ambiguous, ALT field will be N and true alt-allele is "
              "stored here, following IUPAC code convention. This option does not appear when ALT != N.\">\n");
  bq_kputs(h, "##FORMAT=<ID=DP,Number=1,Type=Integer,Description=\"Raw read depth\">\n");
  bq_kputs(h, "##FORMAT=<ID=SP,Number=.,Type=String,Description=\"Allele support (considering bisulfite conversion, with filtering)\">\n");
  bq_kputs(h, "##FORMAT=<ID=AC,Number=.,Type=Integer,Description=\"Depth in calculating alternative allele frequency (after inference, with "
              "filtering)\">\n");
  bq_kputs(h, "##FORMAT=<ID=AF1,Number=.,Type=Float,Description=\"Alternative allele frequency (after inference, with filtering)\">\n");
  bq_kputs(h, "##FORMAT=<ID=CV,Number=1,Type=Integer,Description=\"Effective (strand-specific) coverage on cytosine\">\n");
  bq_kputs(h, "##FORMAT=<ID=BT,Number=1,Type=Float,Description=\"Cytosine methylation fraction (aka beta value, with filtering)\">\n");
  bq_kputs(h, "##FORMAT=<ID=GT,Number=1,Type=String,Description=\"Genotype from normal\">\n");
  bq_kputs(h, "##FORMAT=<ID=GL1,Number=3,Type=Float,Description=\"Genotype likelihoods for the first alternative allele\">\n");
  bq_kputs(h, "##FORMAT=<ID=GQ,Number=1,Type=Integer,Description=\"Genotype quality (phred-scaled)\">\n");
  bq_kputs(h, "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT");
  for (int sid = 0; sid < n_fns; ++sid) { /* sample name = BAM file name without directory and .bam */
    bq_kputc(h, '\t');
    char *path = strdup(in_fns[sid]);
    char *bname = basename(path);
    const size_t l = strlen(bname);
    if (l >= 4 && strcmp(bname + l - 4, ".bam") == 0) bname[l - 4] = 0;
    bq_kputs(h, bname);
    free(path);
  }
  bq_kputc(h, '\n');
}

/* print_meth_average_1chrom / print_meth_average1, src/pileup.c:74-143 */
static void meth_average_1chrom(FILE *out, const char *sample, const char *chrom, const double *betasum, const int64_t *cnt, int is_nome) {
  enum { HCG = 0, HCHG, HCHH, GCG, GCHG, GCHH };
  if (is_nome) {
    const int64_t k_hcg = cnt[HCG], k_hchg = cnt[HCHG], k_hchh = cnt[HCHH], k_hch = k_hchg + k_hchh;
    const double b_hcg = betasum[HCG], b_hchg = betasum[HCHG], b_hchh = betasum[HCHH], b_hch = b_hchg + b_hchh;
    const int64_t k_gc = cnt[GCG] + cnt[GCHG] + cnt[GCHH];
    const double b_gc = betasum[GCG] + betasum[GCHG] + betasum[GCHH];
    if (k_hcg > 0) {
      fprintf(out, "%s\t%s", sample, chrom);
      fprintf(out, "\t%ld\t%1.3f%%", (long)k_hcg, b_hcg / (double)k_hcg * 100);
      fprintf(out, "\t%ld\t%1.3f%%", (long)k_hchg, b_hchg / (double)k_hchg * 100);
      fprintf(out, "\t%ld\t%1.3f%%", (long)k_hchh, b_hchh / (double)k_hchh * 100);
      fprintf(out, "\t%ld\t%1.3f%%", (long)k_hch, b_hch / (double)k_hch * 100);
      fprintf(out, "\t%ld\t%1.3f%%", (long)k_gc, b_gc / (double)k_gc * 100);
      fputc('\n', out);
    }
  } else {
    const int64_t k_cg = cnt[GCG] + cnt[HCG], k_chg = cnt[GCHG] + cnt[HCHG], k_chh = cnt[GCHH] + cnt[HCHH], k_ch = k_chg + k_chh;
    const double b_cg = betasum[GCG] + betasum[HCG], b_chg = betasum[GCHG] + betasum[HCHG], b_chh = betasum[GCHH] + betasum[HCHH];
    const double b_ch = b_chg + b_chh;
    if (k_cg > 0) {
      fprintf(out, "%s\t%s", sample, chrom);
      fprintf(out, "\t%ld\t%1.3f%%", (long)k_cg, b_cg / (double)k_cg * 100);
      fprintf(out, "\t%ld\t%1.3f%%", (long)k_chg, b_chg / (double)k_chg * 100);
      fprintf(out, "\t%ld\t%1.3f%%", (long)k_chh, b_chh / (double)k_chh * 100);
      fprintf(out, "\t%ld\t%1.3f%%", (long)k_ch, b_ch / (double)k_ch * 100);
      fputc('\n', out);
    }
  }
}

/* hts_parse_reg-style "chr", "chr:beg", "chr:beg-end" (1-based inclusive in the string, commas allowed);
 * returns beg0 (0-based) and end (exclusive) like biscuit_parse_region (src/bisc_utils.h:165-180) */
static int parse_region(const char *reg, const bq_bam_hdr_t *hdr, int *tid, int64_t *beg, int64_t *end) {
  *beg = 0; *end = INT32_MAX;
  for (int i = 0; i < hdr->n_targets; ++i)
    if (strcmp(hdr->name[i], reg) == 0) { *tid = i; return 0; }
  const char *colon = strrchr(reg, ':');
  if (!colon) return -1;
  char *name = strndup(reg, (size_t)(colon - reg));
  *tid = -1;
  for (int i = 0; i < hdr->n_targets; ++i)
    if (strcmp(hdr->name[i], name) == 0) { *tid = i; break; }
  free(name);
  if (*tid < 0) return -1;
  char num[64];
  int n = 0;
  const char *p = colon + 1;
  for (; *p && *p != '-' && n < 62; ++p) if (*p != ',') num[n++] = *p;
  num[n] = 0;
  *beg = atoll(num) - 1;
  if (*beg < 0) *beg = 0;
  if (*p == '-') {
    n = 0;
    for (++p; *p && n < 62; ++p) if (*p != ',') num[n++] = *p;
    num[n] = 0;
    *end = atoll(num);
  }
  return *beg < *end ? 0 : -1;
}

typedef struct {
  bq_bgzf_t *fp;
  bq_bai_t bai;
  int cur_tid_done; /* reader has run past the current contig */
} bam_in_t;

/* ---- small blocking queue of pointers (pipeline plumbing) ---- */
#define PQ_CAP 8
typedef struct { pthread_mutex_t mu; pthread_cond_t cv; void *a[PQ_CAP]; int head, n; } pq_t;
static void pq_init(pq_t *q) { memset(q, 0, sizeof *q); pthread_mutex_init(&q->mu, 0); pthread_cond_init(&q->cv, 0); }
static void pq_put(pq_t *q, void *p) {
  pthread_mutex_lock(&q->mu);
  while (q->n == PQ_CAP) pthread_cond_wait(&q->cv, &q->mu);
  q->a[(q->head + q->n++) % PQ_CAP] = p;
  pthread_cond_broadcast(&q->cv);
  pthread_mutex_unlock(&q->mu);
}
static void *pq_get(pq_t *q) {
  pthread_mutex_lock(&q->mu);
  while (q->n == 0) pthread_cond_wait(&q->cv, &q->mu);
  void *p = q->a[q->head];
  q->head = (q->head + 1) % PQ_CAP; --q->n;
  pthread_cond_broadcast(&q->cv);
  pthread_mutex_unlock(&q->mu);
  return p;
}

/* one chunk of windows on its way from the BAM decoder to the GPU stage */
typedef struct {
  int end, tid;
  int64_t cb, ce;
  bq_plp_batch_t *bt;
  uint8_t *ref; /* nt4 codes of the contig, set on its first chunk (ownership passes to the consumer) */
  int64_t ref_len;
} chunk_msg_t;

typedef struct { bq_str_t text; int end; } text_msg_t;

typedef struct {
  /* decoder */
  int nb, n_work, n_threads;
  bam_in_t *in;
  const bq_bam_hdr_t *hdr;
  const bq_fasta_t *fa;
  const char *reffn;
  struct work_item { int tid; int64_t beg, end; } *work;
  int64_t chunk;
  pq_t q_chunks, q_free_batches;
  double t_dec;
  /* writer */
  FILE *out;
  pq_t q_text, q_free_text;
  double t_wr;
} plp_pipe_t;

static double now_s(void);

/* stage 1: stream the BAMs contig by contig and cut them into chunks of decoded records */
static void *decoder_main(void *arg) {
  plp_pipe_t *P = arg;
  const int nb = P->nb;
  for (int wi = 0; wi < P->n_work; ++wi) {
    const int tid = P->work[wi].tid;
    const int64_t beg = P->work[wi].beg, end = P->work[wi].end;
    if (beg >= end) continue;
    int any = 0;
    for (int s = 0; s < nb; ++s) {
      const uint64_t v = bq_bai_start(&P->in[s].bai, tid, beg - 1);
      P->in[s].cur_tid_done = v == UINT64_MAX;
      if (!P->in[s].cur_tid_done) { bq_bgzf_seek(P->in[s].fp, v); any = 1; }
    }
    if (!any) continue; /* no reads on this contig: the reference emits nothing for it */
    uint8_t *ref = 0;
    const int64_t ref_len = bq_fasta_fetch_nt4(P->fa, P->hdr->name[tid], &ref);
    if (ref_len < 0) bq_fatal("[pileup] contig %s is not in %s\n", P->hdr->name[tid], P->reffn);
    if (ref_len < end) bq_fatal("[pileup] contig %s: reference has %ld bases, BAM header says %d\n", P->hdr->name[tid], (long)ref_len, P->hdr->len[tid]);
    bq_plp_batch_t carry; /* reads reaching into the next chunk (a few dozen): private, plain memory */
    memset(&carry, 0, sizeof carry);
    carry.plain = 1;
    for (int64_t cb = beg; cb < end; cb += P->chunk) {
      const int64_t ce = cb + P->chunk < end ? cb + P->chunk : end;
      bq_plp_batch_t *bt = pq_get(&P->q_free_batches); /* not timed: waiting for the consumer is not decoding */
      const double t0 = now_s();
      bq_plp_batch_reset(bt);
      for (int64_t i = 0; i < carry.n; ++i) bq_plp_batch_copy1(bt, &carry, i);
      /* reads of this contig starting before the chunk end (0-based pos < ce - 1), decoded on n_threads threads */
      for (int s = 0; s < nb; ++s)
        if (!P->in[s].cur_tid_done) {
          bq_plp_batch_fill(bt, P->in[s].fp, s, tid, ce - 1, P->n_threads);
          uint32_t len;
          const uint8_t *r = bq_bam_peek(P->in[s].fp, &len);
          if (!r || (int32_t)(r[0] | r[1] << 8 | r[2] << 16 | (uint32_t)r[3] << 24) != tid) P->in[s].cur_tid_done = 1;
        }
      /* the reads that reach into the next chunk are carried over */
      bq_plp_batch_reset(&carry);
      for (int64_t i = 0; i < bt->n; ++i)
        if (bt->end[i] >= ce) bq_plp_batch_copy1(&carry, bt, i);
      P->t_dec += now_s() - t0;
      chunk_msg_t *m = calloc(1, sizeof *m);
      m->tid = tid; m->cb = cb; m->ce = ce; m->bt = bt; m->ref = ref; m->ref_len = ref_len;
      ref = 0;
      pq_put(&P->q_chunks, m);
    }
    bq_plp_batch_free(&carry);
    free(ref);
  }
  chunk_msg_t *m = calloc(1, sizeof *m);
  m->end = 1;
  pq_put(&P->q_chunks, m);
  return 0;
}

/* stage 3: ordered output */
static void *writer_main(void *arg) {
  plp_pipe_t *P = arg;
  for (;;) {
    text_msg_t *t = pq_get(&P->q_text);
    if (t->end) { free(t); return 0; }
    const double t0 = now_s();
    if (t->text.l && fwrite(t->text.s, 1, t->text.l, P->out) != t->text.l && errno == EPIPE) exit(1);
    P->t_wr += now_s() - t0;
    pq_put(&P->q_free_text, t);
  }
}

static double now_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + ts.tv_nsec * 1e-9;
}

int bq_main_pileup(int argc, char **argv) {
  int c;
  char *reg = 0, *tum = 0, *nor = 0, *outfn = 0, *statsfn = 0;
  plp_conf_t conf;
  conf_init(&conf);
  /* Multi-GPU: one process per GPU, launched like any one-rank-per-GPU job (RANK / WORLD_SIZE / LOCAL_RANK in the
   * environment, e.g. under torchrun; BSQ_RANK / BSQ_WORLD override).  Contigs (in name order) are dealt to the ranks
   * round-robin, each rank piles up its own on its own device, and rank 0 merges: VCF bodies concatenated in contig-name
   * order, per-contig methylation statistics added in rank order (every entry has a single contributor, so the sums
   * equal a one-process run bit for bit).  The merge goes through files next to the output: KBs per rank. */
  int device = 0, rank = 0, world = 1;
  if (getenv("WORLD_SIZE")) world = atoi(getenv("WORLD_SIZE"));
  if (getenv("RANK")) rank = atoi(getenv("RANK"));
  if (getenv("BSQ_WORLD")) world = atoi(getenv("BSQ_WORLD"));
  if (getenv("BSQ_RANK")) rank = atoi(getenv("BSQ_RANK"));
  if (world < 1 || rank < 0 || rank >= world) bq_fatal("[pileup] bad rank %d of %d\n", rank, world);
  if (world > 1 && getenv("LOCAL_RANK")) device = atoi(getenv("LOCAL_RANK"));
  if (getenv("BSQ_DEVICE")) device = atoi(getenv("BSQ_DEVICE"));
  if (argc < 2) return usage(&conf);
  while ((c = getopt(argc, argv, ":o:w:g:@:5:3:b:s:E:M:x:C:P:Q:t:n:m:a:l:T:I:SNrcdupv:h")) >= 0) {
    switch (c) {
      case 'g': reg = optarg; break;
      case '@': conf.n_threads = atoi(optarg); break;
      case 's': conf.step = atoi(optarg); break;
      case 'N': conf.is_nome = 1; break;
      case 'S': conf.somatic = 1; break;
      case 'T': tum = optarg; break;
      case 'I': nor = optarg; break;
      case 'o': outfn = optarg; break;
      case 'w': statsfn = strdup(optarg); break;
      case 'v': conf.verbose = atoi(optarg); break;
      case 'b': conf.filt.min_base_qual = atoi(optarg); break;
      case 'm': conf.filt.min_mapq = atoi(optarg); break;
      case 'a': conf.filt.min_score = atoi(optarg); break;
      case 't': conf.filt.max_retention = atoi(optarg); break;
      case 'l': conf.filt.min_read_len = atoi(optarg); break;
      case '5': conf.filt.min_dist_end_5p = atoi(optarg); break;
      case '3': conf.filt.min_dist_end_3p = atoi(optarg); break;
      case 'r': conf.filt.ambi_redist = 0; break;
      case 'c': conf.filt.filter_secondary = 0; break;
      case 'd': conf.filt.filter_doublecnt = 0; break;
      case 'u': conf.filt.filter_duplicate = 0; break;
      case 'p': conf.filt.filter_ppair = 0; break;
      case 'n': conf.filt.max_nm = atoi(optarg); break;
      case 'E': conf.error = atof(optarg); break;
      case 'M': conf.mu = atof(optarg); break;
      case 'x': conf.mu_somatic = atof(optarg); break;
      case 'C': conf.contam = atof(optarg); break;
      case 'P': conf.prior1 = atof(optarg); break;
      case 'Q': conf.prior2 = atof(optarg); break;
      case 'h': return usage(&conf);
      case ':': usage(&conf); bq_fatal("Option needs an argument: -%c\n", optopt); break;
      case '?': usage(&conf); bq_fatal("Unrecognized option: -%c\n", optopt); break;
      default: return usage(&conf);
    }
  }
  conf.filt.is_nome = conf.is_nome;
  if (conf.somatic || tum || nor)
    bq_fatal("[pileup] somatic mode (-S/-T/-I) is not available: somatic_posterior() lives in huishenlab/utils, which is not part of the "
             "reference tree\n");
  if (conf.verbose > 0 && conf.verbose <= 5)
    bq_fatal("[pileup] -v 1..5 (per-read DIAGNOSE columns) is not available in the GPU pileup; use -v 0 or -v >5 for progress messages\n");
  const int progress = conf.verbose > 5;
  conf.filt.verbose = 0;
  if (conf.step < 1) bq_fatal("[pileup] -s must be positive\n");
  if (optind + 2 > argc) { usage(&conf); bq_fatal("Reference or bam input is missing\n"); }
  const char *reffn = argv[optind++];
  /* the genotype fields rest on a restatement of a header that is absent from the reference tree (see the top of this
   * file): say so on every run instead of printing them as if they were verified against upstream */
  if (rank == 0 && !getenv("BSQ_PLP_QUIET"))
    fprintf(stderr, "[W::pileup] QUAL, FILTER, GT, GL1 and GQ are computed from a restatement of huishenlab/utils stats.h (not part of the "
                    "reference tree): these VCF fields are not verified against an upstream build; all other fields are\n");
  int n_fns = argc - optind;
  char **in_fns = argv + optind;
  if (n_fns > 8) bq_fatal("[pileup] at most 8 BAM files\n");
  const int nb = n_fns;

  const double t_start = now_s();
  bam_in_t in[8];
  bq_bam_hdr_t hdr, h2;
  memset(&hdr, 0, sizeof hdr);
  for (int s = 0; s < nb; ++s) {
    in[s].fp = bq_bgzf_open(in_fns[s], conf.n_threads);
    if (!in[s].fp) { fprintf(stderr, "[%s:%d] Cannot open %s\nAbort.\n", __func__, __LINE__, in_fns[s]); exit(1); }
    if (bq_bam_read_header(in[s].fp, s ? &h2 : &hdr) != 0) bq_fatal("[pileup] %s is not a BAM file\n", in_fns[s]);
    if (s) bq_bam_hdr_free(&h2); /* all BAMs are assumed to share the header of the first (src/pileup.c:1119) */
    if (bq_bai_load(in_fns[s], &in[s].bai) != 0) bq_fatal("[pileup] Cannot load index of %s (expected %s.bai)\n", in_fns[s], in_fns[s]);
  }
  target_t *targets = calloc((size_t)hdr.n_targets + 1, sizeof *targets);
  for (int i = 0; i < hdr.n_targets; ++i) { targets[i].tid = i; targets[i].name = hdr.name[i]; targets[i].len = hdr.len[i]; }
  qsort(targets, (size_t)hdr.n_targets, sizeof *targets, cmp_target); /* src/pileup.c:1135 */

  const double t_bam = now_s();
  bq_fasta_t fa;
  if (bq_fasta_load(reffn, &fa) != 0) bq_fatal("[pileup] Cannot open reference %s\n", reffn);
  const double t_fa = now_s();

  FILE *out = stdout;
  char *partfn = 0;
  const char *run_id = getenv("TORCHELASTIC_RUN_ID") ? getenv("TORCHELASTIC_RUN_ID") : (getenv("MASTER_PORT") ? getenv("MASTER_PORT") : "0");
  if (world > 1) {
    if (!outfn) bq_fatal("[pileup] with several ranks the output must be a file (-o)\n");
    partfn = calloc(strlen(outfn) + strlen(run_id) + 64, 1);
    sprintf(partfn, "%s.%s.done%d", outfn, run_id, rank);
    remove(partfn); /* a marker left by an earlier, interrupted run */
    sprintf(partfn, "%s.%s.part%d", outfn, run_id, rank);
  }
  if (outfn) {
    out = fopen(partfn ? partfn : outfn, "w");
    if (!out) { fprintf(stderr, "[%s:%d] Cannot open output file: %s\nAbort.\n", __func__, __LINE__, partfn ? partfn : outfn); exit(1); }
  }
  setvbuf(out, 0, _IOFBF, 1 << 22);
  bq_str_t vcf_hdr = {0, 0, 0};
  vcf_header(&vcf_hdr, reffn, targets, hdr.n_targets, argv, argc, &conf, in_fns, n_fns);
  if (world == 1) fputs(vcf_hdr.s, out); /* sharded: rank 0 writes it when it merges */

  bsq_plp *plp = 0;
  int rc = bsq_plp_create(device, nb, &plp);
  if (rc) bq_fatal("[pileup] bsq_plp_create: %s (%s)\n", bsq_strerror(rc), bsq_last_error());
  const double t_cuda = now_s();

  /* work list: (target index, beg, end) with 1-based beg, exclusive end (src/pileup.c:1171-1200) */
  int n_work = 0;
  struct work_item *work = calloc((size_t)hdr.n_targets + 1, sizeof *work);
  if (reg) {
    int tid = -1; int64_t beg, end;
    if (parse_region(reg, &hdr, &tid, &beg, &end) != 0) bq_fatal("[pileup] cannot parse region %s\n", reg);
    beg++;
    if (beg <= 0) beg = 1;
    if (end > hdr.len[tid]) end = hdr.len[tid];
    if (rank == 0) { work[n_work].tid = tid; work[n_work].beg = beg; work[n_work].end = end; n_work++; } /* a region is not sharded */
  } else {
    for (int j = 0; j < hdr.n_targets; ++j)
      if (j % world == rank) { work[n_work].tid = targets[j].tid; work[n_work].beg = 1; work[n_work].end = targets[j].len; n_work++; }
  }
  int64_t *seg_bytes = calloc((size_t)hdr.n_targets + 1, sizeof(int64_t)); /* VCF text bytes per BAM contig, this rank */

  /* statistics: [sid][tid][ctx] (write_func, src/pileup.c:160-185) */
  const int smpl_block = hdr.n_targets * BQ_NCTX;
  double *betasum = calloc((size_t)nb * smpl_block + 1, sizeof(double));
  int64_t *cnt = calloc((size_t)nb * smpl_block + 1, sizeof(int64_t));

  bq_plp_fmt_t fcf;
  fcf.n_bams = nb; fcf.is_nome = conf.is_nome; fcf.n_threads = conf.n_threads; fcf.error = conf.error; fcf.contam = conf.contam;
  /* prior0 is the value set with the defaults (1 - 0.33333 - 0.33333): the reference derives it in pileup_conf_init
   * (src/pileup.c:955), before -P / -Q are parsed, and never again -- kept, so that GL1/GQ agree when they are given */
  fcf.prior0 = conf.prior0; fcf.prior1 = conf.prior1; fcf.prior2 = conf.prior2;

  /* chunk = a whole number of windows */
  /* about a million loci per chunk: small enough that decode, GPU and text overlap well and that the page-locked
   * staging stays small, large enough that the per-chunk launches do not matter */
  int64_t win_per_chunk = 1000000 / conf.step;
  if (win_per_chunk < 1) win_per_chunk = 1;
  const int64_t chunk = win_per_chunk * conf.step;
  /* pipeline: BAM decode (thread) | GPU + VCF text (this thread, text on conf.n_threads threads) | output (thread) */
  plp_pipe_t P;
  memset(&P, 0, sizeof P);
  P.nb = nb; P.n_work = n_work; P.n_threads = conf.n_threads; P.in = in; P.hdr = &hdr; P.fa = &fa; P.reffn = reffn; P.work = work; P.chunk = chunk; P.out = out;
  pq_init(&P.q_chunks); pq_init(&P.q_free_batches); pq_init(&P.q_text); pq_init(&P.q_free_text);
  bq_plp_batch_t B[3];  /* one being filled, one carrying over, one on the GPU */
  memset(B, 0, sizeof B);
  for (int i = 0; i < 3; ++i) pq_put(&P.q_free_batches, &B[i]);
  text_msg_t *texts[3];
  for (int i = 0; i < 3; ++i) { texts[i] = calloc(1, sizeof(text_msg_t)); pq_put(&P.q_free_text, texts[i]); }
  pthread_t th_dec, th_wr;
  pthread_create(&th_dec, 0, decoder_main, &P);
  pthread_create(&th_wr, 0, writer_main, &P);
  bsq_plp_rec *recs = 0; /* page-locked */
  int64_t recs_cap = 0;
  double *wbeta = calloc((size_t)win_per_chunk * nb * BQ_NCTX, sizeof(double));
  int64_t *wcnt = calloc((size_t)win_per_chunk * nb * BQ_NCTX, sizeof(int64_t));
  double t_gpu = 0, t_fmt = 0, t_wait = 0;
  int64_t tot_reads = 0, tot_loci = 0, tot_emit = 0;
  for (;;) {
    double t0 = now_s();
    chunk_msg_t *m = pq_get(&P.q_chunks);
    t_wait += now_s() - t0; t0 = now_s();
    if (m->end) { free(m); break; }
    const int tid = m->tid;
    const int64_t cb = m->cb, ce = m->ce;
    bq_plp_batch_t *bt = m->bt;
    if (m->ref) {
      if ((rc = bsq_plp_set_contig(plp, m->ref, (int32_t)m->ref_len))) bq_fatal("[pileup] bsq_plp_set_contig: %s (%s)\n", bsq_strerror(rc), bsq_last_error());
      free(m->ref);
    }
    int64_t n_loci = 0;
    const int64_t n_reads_chunk = bt->n;
    if (bt->n > 0) {
      bsq_plp_reads view;
      bq_plp_batch_view(bt, &view);
      if ((rc = bsq_plp_stage(plp, &view))) bq_fatal("[pileup] bsq_plp_stage: %s (%s)\n", bsq_strerror(rc), bsq_last_error());
      if ((rc = bsq_plp_run(plp, &conf.filt, (int32_t)cb, (int32_t)ce, &n_loci))) bq_fatal("[pileup] bsq_plp_run: %s (%s)\n", bsq_strerror(rc), bsq_last_error());
    }
    pq_put(&P.q_free_batches, bt); /* bsq_plp_run has synchronised: the staged copies are complete */
    if (n_loci > recs_cap) {
      if (recs) bsq_host_free(recs);
      recs_cap = n_loci * 5 / 4 + 1024;
      void *pp_ = 0;
      if (bsq_host_alloc(&pp_, (size_t)recs_cap * nb * sizeof *recs)) bq_fatal("[pileup] out of page-locked memory\n");
      recs = pp_;
    }
    if (n_loci > 0 && (rc = bsq_plp_fetch(plp, recs))) bq_fatal("[pileup] bsq_plp_fetch: %s (%s)\n", bsq_strerror(rc), bsq_last_error());
    t_gpu += now_s() - t0; t0 = now_s();
    tot_reads += n_reads_chunk; tot_loci += ce - cb; tot_emit += n_loci;
    const int n_win = (int)((ce - cb + conf.step - 1) / conf.step);
    if (n_loci > 0) {
      memset(wbeta, 0, (size_t)n_win * nb * BQ_NCTX * sizeof(double));
      memset(wcnt, 0, (size_t)n_win * nb * BQ_NCTX * sizeof(int64_t));
      text_msg_t *tm = pq_get(&P.q_free_text);
      tm->text.l = 0;
      bq_plp_format(&fcf, hdr.name[tid], recs, n_loci, cb, conf.step, n_win, &tm->text, wbeta, wcnt);
      seg_bytes[tid] += (int64_t)tm->text.l;
      pq_put(&P.q_text, tm);
      for (int w = 0; w < n_win; ++w) /* one record per window, added in block order (write_func) */
        for (int s = 0; s < nb; ++s)
          for (int i = 0; i < BQ_NCTX; ++i) {
            betasum[s * smpl_block + tid * BQ_NCTX + i] += wbeta[((size_t)w * nb + s) * BQ_NCTX + i];
            cnt[s * smpl_block + tid * BQ_NCTX + i] += wcnt[((size_t)w * nb + s) * BQ_NCTX + i];
          }
      t_fmt += now_s() - t0;
    }
    if (progress) fprintf(stderr, "[pileup] %s:%ld-%ld reads %ld emitted %ld\n", hdr.name[tid], (long)cb, (long)ce, (long)n_reads_chunk, (long)n_loci);
    free(m);
  }
  {
    text_msg_t *e = calloc(1, sizeof *e);
    e->end = 1;
    pq_put(&P.q_text, e);
  }
  pthread_join(th_dec, 0); pthread_join(th_wr, 0);
  const double t_dec = P.t_dec, t_wr = P.t_wr;

  if (world > 1) { /* hand this rank's part over; rank 0 merges */
    fclose(out); out = 0;
    const size_t n_stat = (size_t)nb * smpl_block;
    char *fn = calloc(strlen(outfn) + strlen(run_id) + 64, 1), *fn2 = calloc(strlen(outfn) + strlen(run_id) + 64, 1);
    sprintf(fn, "%s.%s.stats%d", outfn, run_id, rank);
    FILE *sf = fopen(fn, "wb");
    if (!sf || fwrite(seg_bytes, sizeof(int64_t), (size_t)hdr.n_targets, sf) != (size_t)hdr.n_targets || fwrite(cnt, sizeof(int64_t), n_stat, sf) != n_stat ||
        fwrite(betasum, sizeof(double), n_stat, sf) != n_stat || fclose(sf) != 0)
      bq_fatal("[pileup] cannot write %s\n", fn);
    sprintf(fn2, "%s.%s.done%d", outfn, run_id, rank);
    if (rename(fn, fn2) != 0) bq_fatal("[pileup] cannot write %s\n", fn2);
    if (rank != 0) { fprintf(stderr, "[main] Real time: %.3f sec (rank %d of %d)\n", now_s() - t_start, rank, world); fflush(stderr); _exit(0); }
    /* rank 0: wait for every rank, add the statistics in rank order, concatenate the bodies in contig-name order */
    double wait_max = getenv("BSQ_PLP_MERGE_TIMEOUT") ? atof(getenv("BSQ_PLP_MERGE_TIMEOUT")) : 86400.;
    int64_t *segs = calloc((size_t)world * hdr.n_targets + 1, sizeof(int64_t)), *c2 = malloc(sizeof(int64_t) * (n_stat + 1));
    double *b2 = malloc(sizeof(double) * (n_stat + 1));
    memset(cnt, 0, n_stat * sizeof(int64_t)); memset(betasum, 0, n_stat * sizeof(double));
    for (int r = 0; r < world; ++r) {
      sprintf(fn2, "%s.%s.done%d", outfn, run_id, r);
      const double tw = now_s();
      FILE *df;
      while (!(df = fopen(fn2, "rb"))) {
        if (now_s() - tw > wait_max) bq_fatal("[pileup] rank %d did not finish within %.0f s (no %s)\n", r, wait_max, fn2);
        struct timespec nap = {0, 20000000};
        nanosleep(&nap, 0);
      }
      if (fread(segs + (size_t)r * hdr.n_targets, sizeof(int64_t), (size_t)hdr.n_targets, df) != (size_t)hdr.n_targets ||
          fread(c2, sizeof(int64_t), n_stat, df) != n_stat || fread(b2, sizeof(double), n_stat, df) != n_stat)
        bq_fatal("[pileup] truncated %s\n", fn2);
      fclose(df);
      for (size_t i = 0; i < n_stat; ++i) { cnt[i] += c2[i]; betasum[i] += b2[i]; }
    }
    out = fopen(outfn, "w");
    if (!out) { fprintf(stderr, "[%s:%d] Cannot open output file: %s\nAbort.\n", __func__, __LINE__, outfn); exit(1); }
    setvbuf(out, 0, _IOFBF, 1 << 22);
    fputs(vcf_hdr.s, out);
    FILE **pf = calloc((size_t)world, sizeof(FILE *));
    for (int r = 0; r < world; ++r) {
      sprintf(fn, "%s.%s.part%d", outfn, run_id, r);
      if (!(pf[r] = fopen(fn, "rb"))) bq_fatal("[pileup] cannot read %s\n", fn);
    }
    char *buf = malloc(1 << 22);
    for (int j = 0; j < hdr.n_targets; ++j) { /* each rank wrote its contigs in this same order */
      const int r = reg ? 0 : j % world;
      int64_t left = segs[(size_t)r * hdr.n_targets + targets[j].tid];
      while (left > 0) {
        const size_t want = left > (1 << 22) ? (size_t)(1 << 22) : (size_t)left;
        if (fread(buf, 1, want, pf[r]) != want) bq_fatal("[pileup] truncated part of rank %d\n", r);
        fwrite(buf, 1, want, out);
        left -= (int64_t)want;
      }
    }
    free(buf);
    for (int r = 0; r < world; ++r) {
      fclose(pf[r]);
      sprintf(fn, "%s.%s.part%d", outfn, run_id, r); remove(fn);
      sprintf(fn2, "%s.%s.done%d", outfn, run_id, r); remove(fn2);
    }
    free(pf); free(segs); free(c2); free(b2); free(fn); free(fn2);
  }

  if (!statsfn && outfn) statsfn = strdup(outfn);
  if (statsfn) { /* src/pileup.c:201-222 */
    char *fn = calloc(strlen(statsfn) + 20, 1);
    strcpy(fn, statsfn); strcat(fn, "_meth_average.tsv");
    FILE *so = fopen(fn, "w");
    if (!so) bq_fatal("[pileup] cannot write %s\n", fn);
    if (conf.is_nome) fprintf(so, "sample\tchrm\tHCGn\tHCGb\tHCHGn\tHCHGb\tHCHHn\tHCHHb\tHCHn\tHCHb\tGCn\tGCb\n");
    else fprintf(so, "sample\tchrm\tCGn\tCGb\tCHGn\tCHGb\tCHHn\tCHHb\tCHn\tCHb\n");
    for (int s = 0; s < nb; ++s) {
      double b0[BQ_NCTX] = {0};
      int64_t c0[BQ_NCTX] = {0};
      /* print_meth_average1 (src/pileup.c:121-143): row k holds the sums of BAM contig k and is labelled
       * targets[targets[k].tid].name -- kept as is */
      for (int k = 0; k < hdr.n_targets; ++k) {
        const int t = targets[k].tid;
        meth_average_1chrom(so, in_fns[s], targets[t].name, betasum + s * smpl_block + k * BQ_NCTX, cnt + s * smpl_block + k * BQ_NCTX, conf.is_nome);
        for (int i = 0; i < BQ_NCTX; ++i) { c0[i] += cnt[s * smpl_block + k * BQ_NCTX + i]; b0[i] += betasum[s * smpl_block + k * BQ_NCTX + i]; }
      }
      meth_average_1chrom(so, in_fns[s], "WholeGenome", b0, c0, conf.is_nome);
    }
    fclose(so);
    free(fn);
  }
  if (progress || getenv("BSQ_PLP_TIMING"))
    { extern double bq_bgzf_t_read, bq_bgzf_t_inflate, bq_plp_t_fill;
      fprintf(stderr, "[pileup] decode thread: file read %.2fs, inflate %.2fs, record decode %.2fs\n", bq_bgzf_t_read, bq_bgzf_t_inflate, bq_plp_t_fill); }
    fprintf(stderr, "[pileup] start-up: BAM/BAI open %.2fs, FASTA load %.2fs, CUDA context %.2fs; main loop %.2fs\n", t_bam - t_start, t_fa - t_bam,
            t_cuda - t_fa, now_s() - t_cuda);
  if (progress || getenv("BSQ_PLP_TIMING"))
    fprintf(stderr, "[pileup] reads %ld loci %ld emitted %ld | decode %.2fs (thread) | wait %.2fs gpu %.2fs format %.2fs | write %.2fs (thread)\n",
            (long)tot_reads, (long)tot_loci, (long)tot_emit, t_dec, t_wait, t_gpu, t_fmt, t_wr);
  if (outfn) fclose(out); else fflush(out);
  if (!getenv("BSQ_PLP_FULL_TEARDOWN")) {
    /* All output is written.  Unpinning several hundred MB of page-locked staging memory and tearing the CUDA context
     * down takes 0.5-1 s and frees nothing the exiting process will not free anyway. */
    fprintf(stderr, "[main] Real time: %.3f sec\n", now_s() - t_start);
    fflush(stderr);
    _exit(0);
  }
  bsq_plp_destroy(plp);
  for (int s = 0; s < nb; ++s) { bq_bgzf_close(in[s].fp); bq_bai_free(&in[s].bai); }
  for (int i = 0; i < 3; ++i) bq_plp_batch_free(&B[i]);
  for (int i = 0; i < 3; ++i) { free(texts[i]->text.s); free(texts[i]); }
  if (recs) bsq_host_free(recs);
  free(wbeta); free(wcnt); free(betasum); free(cnt); free(work); free(targets); free(statsfn);
  bq_fasta_free(&fa);
  bq_bam_hdr_free(&hdr);
  return 0;
}
