/* TEST INFRASTRUCTURE ONLY -- stand-in for huishenlab/utils stats.h @5f4aeab (see README.md).
 *
 * PARITY UNPINNED.  genotype_lnlik, ln_sum3, pval2qual and somatic_posterior are defined in a dependency that is
 * neither vendored under /root/reference nor fetchable here, and the reference holds no test vector for them.  They are
 * restated as the model their call sites imply (pileup.c:389-413, :509): binomial genotype likelihood with
 * error + contamination as the alt-read probability of a homozygous-reference site and 1/2 for a heterozygous one,
 * log-sum-exp, and -10 log10(p) capped at 1000.  The VCF columns that depend on them (QUAL, FILTER, GT, GL1, GQ, SS, SC)
 * are excluded from the pinned comparison in tests/. */
#ifndef BSQ_SHIM_SRC_STATS_H
#define BSQ_SHIM_SRC_STATS_H
#include <math.h>
typedef enum { HOMOREF, HET, HOMOVAR } genotype_t;
static inline double genotype_lnlik(genotype_t gt, int ref_cnt, int alt_cnt, double error, double contam) {
  double p = gt == HOMOREF ? error + contam : gt == HET ? 0.5 : 1.0 - error - contam;
  if (p < 1e-300) p = 1e-300;
  if (p > 1.0 - 1e-16) p = 1.0 - 1e-16;
  return alt_cnt * log(p) + ref_cnt * log(1.0 - p);
}
static inline double ln_sum3(double a, double b, double c) {
  double m = a;
  if (b > m) m = b;
  if (c > m) m = c;
  return m + log(exp(a - m) + exp(b - m) + exp(c - m));
}
static inline double pval2qual(double pval) {
  if (pval <= 0) return 1000;
  double q = -10.0 * log10(pval);
  return q > 1000 ? 1000 : q;
}
/* somatic mode (-S) only; posterior that the tumour carries a variant absent from the normal */
static inline double somatic_posterior(int cref_t, int calt_t, int cref_n, int calt_n, double error, double mu, double mu_somatic, double contam) {
  double t0 = genotype_lnlik(HOMOREF, cref_t, calt_t, error, contam), t1 = genotype_lnlik(HET, cref_t, calt_t, error, contam);
  double n0 = genotype_lnlik(HOMOREF, cref_n, calt_n, error, contam), n1 = genotype_lnlik(HET, cref_n, calt_n, error, contam);
  double som = log(mu_somatic) + t1 + log(1 - mu) + n0;
  double germ = log(mu) + t1 + n1;
  double wild = log(1 - mu) + log(1 - mu_somatic) + t0 + n0;
  return 1.0 - exp(som - ln_sum3(som, germ, wild));
}
#endif
