/* bsq_oracle.h -- CPU restatement of the reference algorithms.  TEST INFRASTRUCTURE ONLY: used as the
 * checker by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg; never linked into or
 * called from the product (biscuit_b200/, libbsq.so).
 *
 * Pileup (src/pileup.c cannot be compiled offline: htslib + huishenlab/utils are absent):
 *   "parity unpinned" for everything that depends on utils/stats.h (QUAL, FILTER, GT, GL1, GQ) -- the
 *   reference holds no test or golden vector for the path; the integer path (events, filters, counts,
 *   redistribution, top mutant, emit rule, methcallable, context, CV/BT operands) is restated here line
 *   by line from code that IS present under /root/reference/src.
 * Align: the unmodified reference itself is compiled into oracle/_ref (see Makefile); the restatement in
 *   bsq_oracle_align.c covers the kernel-level functions and is pinned against oracle/_ref in tests.
 */
#ifndef BSQ_ORACLE_H
#define BSQ_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* same layout as bsq_plp_reads / bsq_plp_conf / bsq_plp_rec in include/bsq.h (kept textually separate
 * on purpose: the oracle must not include product headers) */
typedef struct {
  int64_t n_reads;
  const int32_t *pos;        /* 0-based leftmost reference position (bam1_core_t.pos) */
  const int32_t *mpos;       /* mate position, 0-based */
  const int32_t *mate_rlen;  /* reference length of the mate from the MC tag, -1 if the tag is absent */
  const int32_t *l_qseq;
  const int32_t *nm;         /* NM tag, INT32_MIN if absent */
  const int32_t *as;         /* AS tag, INT32_MIN if absent */
  const uint16_t *flag;
  const uint8_t *mapq;
  const int8_t *bss_tag;     /* 0: YD:f / ZS:+ / XG:CT, 1: YD:r / ZS:- / XG:GA, -1: none of the tags (infer) */
  const uint8_t *sid;        /* sample index */
  const int32_t *n_cigar;
  const int64_t *cigar_off;  /* into cigar[] */
  const uint32_t *cigar;     /* BAM encoding: len<<4 | op */
  const int64_t *seq_off;    /* byte offset into seq[]; 4-bit packed as in BAM, two bases per byte */
  const uint8_t *seq;
  const int64_t *qual_off;
  const uint8_t *qual;
} bsqo_plp_reads;

typedef struct {
  int32_t min_base_qual, min_read_len, min_dist_end_5p, min_dist_end_3p, min_mapq, min_score, max_nm, max_retention;
  int32_t filter_ppair, filter_secondary, filter_duplicate, filter_qcfail, filter_doublecnt;
  int32_t ambi_redist, verbose, is_nome;
} bsqo_plp_conf;

typedef struct {
  int32_t pos, dp;
  int32_t meth[3];
  int32_t base[7];
  int32_t base_redist[7];
  uint8_t rb_code;
  int8_t cm1;
  uint8_t ctx, methcallable;
  char n5[5];
  uint8_t any_callable, pad_[2];
} bsqo_plp_rec;

void bsqo_plp_conf_default(bsqo_plp_conf *c);

/* Pile up reads of one contig over loci [beg, end) (1-based, end exclusive).  ref: nt4 codes
 * (A0 C1 G2 T3 N4) of the whole contig, ref_len bases.  Emitted loci are appended to out (n_bams records
 * per locus); returns the number of emitted loci or -1 if cap_loci is too small. */
int64_t bsqo_plp_region(const bsqo_plp_conf *conf, const uint8_t *ref, int32_t ref_len, int32_t beg, int32_t end,
                        const bsqo_plp_reads *reads, int n_bams, bsqo_plp_rec *out, int64_t cap_loci);

/* genotyping / text options of plp_format (pileup_conf_t, src/pileup.h:49-63; defaults src/pileup.c:944-963) */
typedef struct {
  double error, contam, prior0, prior1, prior2;
  int32_t is_nome, pad_;
} bsqo_vcf_conf;

/* VCF lines (src/pileup.c:521-636) of n_loci emitted loci of contig chrm; betasum/cnt[sid*6+ctx] accumulate the
 * methylation statistics of these loci in order (src/pileup.c:610-616).  malloc()ed text, free with bsqo_free.
 * QUAL/FILTER/GT/GL1/GQ: parity unpinned (see the header of bsq_oracle_vcf.c). */
char *bsqo_plp_vcf(const bsqo_vcf_conf *conf, const char *chrm, const bsqo_plp_rec *recs, int64_t n_loci, int n_bams, double *betasum,
                   int64_t *cnt);
void bsqo_free(void *p);

#ifdef __cplusplus
}
#endif
#endif
