/* CPU restatement of the text half of BISCUIT's pileup -- TEST INFRASTRUCTURE ONLY (see bsq_oracle.h).
 *
 * Follows plp_format (src/pileup.c:415-640) statement by statement from the counts of one emitted locus
 * (bsqo_plp_rec, produced by bsqo_plp_region) to the VCF line, pileup_genotype (src/pileup.c:389-413), and
 * the per-record methylation sums that write_func adds up (src/pileup.c:178-185).
 *
 * PARITY UNPINNED for QUAL, FILTER, GT, GL1, GQ: genotype_lnlik(), ln_sum3() and pval2qual() come from
 * huishenlab/utils stats.h @5f4aeab (CMakeLists.txt:45-54), which is neither vendored under /root/reference
 * nor fetchable here, and the reference holds no test vector for them.  They are restated below as the
 * binomial genotype model the call sites imply (error+contamination as the alt-read probability of a
 * homozygous-reference site, 1/2 for a heterozygous one), log-sum-exp, and -10 log10(p) capped at 1000.
 * All other fields (CHROM POS REF ALT NS CX N5 AB DP SP AC AF1 CV BT) are pinned by code present in the tree.
 */
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "bsq_oracle.h"

enum { B_A = 0, B_C, B_G, B_T, B_N, B_Y, B_R };
static const char basecode[8] = "ACGTNYR";
static const char *cytosine_context[] = {"CG", "CHG", "CHH", "CG", "CHG", "CHH", "CN"};
static const char *cytosine_context_nome[] = {"HCG", "HCHG", "HCHH", "GCG", "GCH", "GCH", "CN"};

static double o_genotype_lnlik(int gt, int ref_cnt, int alt_cnt, double error, double contam) {
  double p = gt == 0 ? error + contam : gt == 1 ? 0.5 : 1.0 - error - contam;
  if (p < 1e-300) p = 1e-300;
  if (p > 1.0 - 1e-16) p = 1.0 - 1e-16;
  return alt_cnt * log(p) + ref_cnt * log(1.0 - p);
}

static double o_ln_sum3(double a, double b, double c) {
  double m = a;
  if (b > m) m = b;
  if (c > m) m = c;
  return m + log(exp(a - m) + exp(b - m) + exp(c - m));
}

static double o_pval2qual(double pval) {
  if (pval <= 0) return 1000;
  double q = -10.0 * log10(pval);
  return q > 1000 ? 1000 : q;
}

static void o_pileup_genotype(int cref, int altsupp, const bsqo_vcf_conf *conf, char gt[4], double *_gl0, double *_gl1, double *_gl2, double *_gq) {
  double gl0 = -1, gl1 = -1, gl2 = -1, gq = -1;
  if (cref >= 0 || altsupp >= 0) {
    gl0 = log(conf->prior0) + o_genotype_lnlik(0, cref, altsupp, conf->error, conf->contam);
    gl1 = log(conf->prior1) + o_genotype_lnlik(1, cref, altsupp, conf->error, conf->contam);
    gl2 = log(conf->prior2) + o_genotype_lnlik(2, cref, altsupp, conf->error, conf->contam);
    if (gl0 > gl1) {
      if (gl0 > gl2) { gq = o_pval2qual(1 - exp(gl0 - o_ln_sum3(gl0, gl1, gl2))); strcpy(gt, "0/0"); }
      else { gq = o_pval2qual(1 - exp(gl2 - o_ln_sum3(gl0, gl1, gl2))); strcpy(gt, "1/1"); }
    } else if (gl1 > gl2) { gq = o_pval2qual(1 - exp(gl1 - o_ln_sum3(gl0, gl1, gl2))); strcpy(gt, "0/1"); }
    else { gq = o_pval2qual(1 - exp(gl2 - o_ln_sum3(gl0, gl1, gl2))); strcpy(gt, "1/1"); }
  }
  *_gl0 = gl0; *_gl1 = gl1; *_gl2 = gl2; *_gq = gq;
}

typedef struct { char *s; size_t l, m; } ostr;
static void oprintf(ostr *s, const char *fmt, ...) {
  va_list ap;
  for (;;) {
    va_start(ap, fmt);
    int n = vsnprintf(s->s + s->l, s->m - s->l, fmt, ap);
    va_end(ap);
    if ((size_t)n < s->m - s->l) { s->l += (size_t)n; return; }
    s->m = (s->m + (size_t)n + 64) * 2;
    s->s = realloc(s->s, s->m);
  }
}

#define omax(a, b) ((a) > (b) ? (a) : (b))

/* VCF text of n_loci emitted loci; betasum/cnt[sid*6+ctx] accumulate over these loci in order.
 * Returns a malloc()ed NUL-terminated string (free with bsqo_free). */
char *bsqo_plp_vcf(const bsqo_vcf_conf *conf, const char *chrm, const bsqo_plp_rec *recs, int64_t n_loci, int n_bams, double *betasum,
                   int64_t *cnt) {
  ostr S = {malloc(1024), 0, 1024};
  S.s[0] = 0;
  for (int64_t l = 0; l < n_loci; ++l) {
    const bsqo_plp_rec *r = recs + l * n_bams;
    ostr *s = &S;
    const int rb_code = r[0].rb_code, cm1 = r[0].cm1;
    const char rb = basecode[rb_code];
    char gt[8][4];
    double gl0[8], gl1[8], gl2[8], gq[8];
    double lowest_gq = 0;
    int any_methcallable = 0, sid;
    for (sid = 0; sid < n_bams; ++sid) {
      strcpy(gt[sid], "./.");
      gl0[sid] = -1; gl1[sid] = -1; gl2[sid] = -1; gq[sid] = 0;
      const int nref = r[sid].base_redist[rb_code];
      const int nalt = cm1 >= 0 ? r[sid].base_redist[cm1] : 0;
      if (nref + nalt > 0) o_pileup_genotype(nref, nalt, conf, gt[sid], gl0 + sid, gl1 + sid, gl2 + sid, gq + sid);
      if (gq[sid] < lowest_gq || !sid) lowest_gq = gq[sid];
      if (r[sid].methcallable) any_methcallable = 1;
    }
    oprintf(s, "%s\t%u\t.\t%c\t", chrm, (unsigned)r[0].pos, rb);
    if (cm1 >= 0) oprintf(s, "%c", (cm1 == B_Y || cm1 == B_R) ? 'N' : basecode[cm1]);
    else oprintf(s, ".");
    oprintf(s, "\t%d", (int)lowest_gq);
    if (lowest_gq > 5) oprintf(s, "\tPASS\t"); else oprintf(s, "\tLowQual\t");
    const int ctt = r[0].ctx;
    oprintf(s, "NS=%d", n_bams);
    if (rb == 'C' || rb == 'G') {
      oprintf(s, ";CX=%s", conf->is_nome ? cytosine_context_nome[ctt] : cytosine_context[ctt]);
      oprintf(s, ";N5=%.5s", r[0].n5);
    }
    if (cm1 >= 0 && (cm1 == B_Y || cm1 == B_R)) oprintf(s, ";AB=%c", basecode[cm1]);
    oprintf(s, "\tGT:GL1:GQ:DP");
    oprintf(s, ":SP");
    if (cm1 >= 0) oprintf(s, ":AC:AF1");
    if (any_methcallable) oprintf(s, ":CV:BT");
    for (sid = 0; sid < n_bams; ++sid) {
      const bsqo_plp_rec *q = r + sid;
      const int dp = q->dp;
      if (gq[sid] > 0 && dp)
        oprintf(s, "\t%s:%1.0f,%1.0f,%1.0f:%1.0f", gt[sid], omax(-1000, gl0[sid]), omax(-1000, gl1[sid]), omax(-1000, gl2[sid]), gq[sid]);
      else oprintf(s, "\t./.:.,.,.:0");
      if (dp) oprintf(s, ":%d", dp); else oprintf(s, ":0");
      oprintf(s, ":");
      int added = 0, i;
      if (q->base[rb_code]) { oprintf(s, "%c%d", rb, q->base[rb_code]); added = 1; }
      for (i = 0; i < 7; ++i) {
        if (i == B_N) continue;
        if (i == rb_code) continue;
        if (q->base[i] <= 0) continue;
        oprintf(s, "%c%d", basecode[i], q->base[i]);
        added = 1;
      }
      if (!added) oprintf(s, ".");
      if (cm1 >= 0) {
        const int nref = q->base_redist[rb_code], nalt = q->base_redist[cm1];
        oprintf(s, ":%d:", nref + nalt);
        if (nref + nalt) oprintf(s, "%1.2f", nalt / (double)(nref + nalt));
        else oprintf(s, ".");
      }
      if (any_methcallable) {
        if (q->methcallable) {
          const double beta = (double)q->meth[0] / (double)(q->meth[0] + q->meth[1]);
          if (ctt != 6) { betasum[sid * 6 + ctt] += beta; cnt[sid * 6 + ctt]++; }
          oprintf(s, ":%d:%1.3f", q->meth[0] + q->meth[1], beta);
        } else oprintf(s, ":0:.");
      }
    }
    oprintf(s, "\n");
  }
  return S.s;
}

void bsqo_free(void *p) { free(p); }
