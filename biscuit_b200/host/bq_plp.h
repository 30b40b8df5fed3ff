/* bq_plp.h -- host side (C) of `biscuit pileup`, `biscuit vcf2bed`, `biscuit mergecg`.
 *
 * The GPU does the per-read event generation and the per-locus integer decisions (libbsq.so, bsq_plp_*);
 * the host reads BGZF/BAM/FASTA, forms the chunks, computes genotype likelihoods (double + libm),
 * formats VCF text and the <out>_meth_average.tsv statistics (src/pileup.c:145-234, :389-640, :874-1225).
 */
#ifndef BQ_PLP_H
#define BQ_PLP_H
#include "bq.h"

/* ---- bq_bam.c ---- */
typedef struct bq_bgzf bq_bgzf_t;
typedef struct {
  int32_t n_targets;
  char **name;
  int32_t *len;
  char *text;
} bq_bam_hdr_t;

bq_bgzf_t *bq_bgzf_open(const char *fn, int n_threads);
void bq_bgzf_close(bq_bgzf_t *b);
void bq_bgzf_seek(bq_bgzf_t *b, uint64_t voffset);
int bq_bam_read_header(bq_bgzf_t *b, bq_bam_hdr_t *h);
void bq_bam_hdr_free(bq_bam_hdr_t *h);
const uint8_t *bq_bam_next(bq_bgzf_t *b, uint32_t *len);
const uint8_t *bq_bam_peek(bq_bgzf_t *b, uint32_t *len);
void bq_bam_skip(bq_bgzf_t *b, uint32_t len);

/* BAI index (SAM/BAM spec 5.2): only what a front-to-back reader needs -- where a contig's records start and,
 * for -g regions, the linear index */
typedef struct {
  int32_t n_ref;
  uint64_t *first;    /* virtual offset of the first record of each reference, UINT64_MAX if none */
  int32_t *n_intv;
  uint64_t **ioffset; /* linear index per reference (16 kb windows) */
} bq_bai_t;
int bq_bai_load(const char *bam_fn, bq_bai_t *bai);
void bq_bai_free(bq_bai_t *bai);
uint64_t bq_bai_start(const bq_bai_t *bai, int tid, int64_t beg0);

/* structure-of-arrays batch of decoded records = the storage behind a bsq_plp_reads view */
typedef struct {
  int64_t n, cap, n_cig, cap_cig, n_seq, cap_seq, n_qual, cap_qual;
  int32_t *pos, *mpos, *mate_rlen, *l_qseq, *nm, *as;
  uint16_t *flag;
  uint8_t *mapq;
  int8_t *bss_tag;
  uint8_t *sid;
  int32_t *n_cigar;
  int64_t *cigar_off;
  uint32_t *cigar;
  int64_t *seq_off;
  uint8_t *seq;
  int64_t *qual_off;
  uint8_t *qual;
  int64_t *end; /* 0-based exclusive reference end (pos + 1 for reads without reference span) */
  int plain;    /* 1: ordinary memory (small private batches); 0: page-locked, staged to the GPU directly */
} bq_plp_batch_t;
void bq_plp_batch_reset(bq_plp_batch_t *B);
void bq_plp_batch_free(bq_plp_batch_t *B);
int bq_plp_batch_push(bq_plp_batch_t *B, const uint8_t *rec, uint32_t len, int sid);
int bq_plp_batch_fill(bq_plp_batch_t *B, bq_bgzf_t *b, int sid, int tid, int64_t pos_lt, int n_threads);
void bq_plp_batch_copy1(bq_plp_batch_t *dst, const bq_plp_batch_t *src, int64_t i);
void bq_plp_batch_view(const bq_plp_batch_t *B, bsq_plp_reads *v);

typedef struct {
  char *buf;
  size_t n_buf;
  int n;
  char **name;
  size_t *beg, *endp;
} bq_fasta_t;
int bq_fasta_load(const char *fn, bq_fasta_t *fa);
void bq_fasta_free(bq_fasta_t *fa);
int64_t bq_fasta_fetch_nt4(const bq_fasta_t *fa, const char *name, uint8_t **out);
int bq_main_bamdump(int argc, char **argv);

/* ---- bq_pileup.c ---- */
#define BQ_NCTX 6
typedef struct {
  int n_bams, is_nome, n_threads;
  double error, contam, prior0, prior1, prior2;
} bq_plp_fmt_t;

/* VCF lines of n_loci emitted loci (n_bams records each, ascending position) of contig `chrm` appended to
 * out; per-window methylation sums (windows [w0 + k*step, w0 + (k+1)*step), k < n_win) added into
 * wbeta/wcnt[k][sid*6+ctx] in locus order (plp_format, src/pileup.c:415-640). */
void bq_plp_format(const bq_plp_fmt_t *cf, const char *chrm, const bsq_plp_rec *recs, int64_t n_loci, int64_t w0, int64_t step, int n_win,
                   bq_str_t *out, double *wbeta, int64_t *wcnt);
int bq_main_pileup(int argc, char **argv);

/* ---- bq_vcf2bed.c ---- */
int bq_main_vcf2bed(int argc, char **argv);
int bq_main_mergecg(int argc, char **argv);
/* ---- bq_sortbam.c ---- */
int bq_main_sortbam(int argc, char **argv);
#endif
