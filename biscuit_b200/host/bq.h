/* bq.h -- host side (C) of the B200 BISCUIT hot paths.
 *
 * The GPU does phase 1 of `biscuit align` (seeding, chaining, extension: libbsq.so, include/bsq.h).
 * Everything here is the host shell around it, restated from the reference's lib/aln in this project's
 * own code: reference sequence access, region merging, insert-size statistics, mate rescue, primary
 * marking, pairing, mapQ, CIGAR/MD generation and SAM text -- the parts of mem_process_seqs
 * (lib/aln/bwamem.c:432-476) that involve libm or text formatting -- plus the
 * `biscuit index|align` command lines and the on-disk index format.
 */
#ifndef BQ_H
#define BQ_H
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/bsq.h"

#define BQ_VERSION "1.6.1-dev-b200"

/* ---- options: mem_opt_t (lib/aln/bwamem.h:54-124), defaults mem_opt_init (bwamem.c:77-128) ---- */
#define BQ_F_PE 0x2
#define BQ_F_NOPAIRING 0x4
#define BQ_F_ALL 0x8
#define BQ_F_NO_MULTI 0x10
#define BQ_F_NO_RESCUE 0x20
#define BQ_F_SELF_OVLP 0x40
#define BQ_F_REF_HDR 0x100
#define BQ_F_SOFTCLIP 0x200
#define BQ_F_KEEP_SUPP_MAPQ 0x1000

typedef struct {
  int a, b, o_del, e_del, o_ins, e_ins, pen_unpaired, pen_clip5, pen_clip3, w, zdrop;
  uint64_t max_mem_intv;
  int T, flag, min_seed_len, min_chain_weight;
  uint32_t max_chain_extend;
  float split_factor;
  int split_width;
  uint32_t max_occ;
  int max_chain_gap, n_threads, chunk_size;
  float mask_level, drop_ratio, XA_drop_ratio, mask_level_redun, mapQ_coef_len;
  int mapQ_coef_fac, max_ins, max_matesw, max_XA_hits, max_XA_hits_alt;
  int8_t mat[25], ctmat[25], gamat[25];
  uint8_t parent, bsstrand;
  uint8_t *adaptor1, *adaptor2;
  int l_adaptor1, l_adaptor2, clip5, clip3, min_base_qual;
  uint8_t has_bc;
} bq_opt_t;

typedef struct {
  int low, high, set, failed;
  double avg, std;
} bq_pestat_t;

/* ---- reference meta data: bntseq_t / bntann1_t (lib/aln/bntseq.h:41-64) ---- */
typedef struct {
  int64_t offset;
  int32_t len, n_ambs;
  uint32_t gi;
  int32_t is_alt;
  char *name, *anno;
} bq_ann_t;

typedef struct {
  int64_t offset;
  int32_t len;
  char amb;
} bq_amb_t;

typedef struct {
  int64_t l_pac;
  int32_t n_seqs;
  uint32_t seed;
  bq_ann_t *anns;
  int32_t n_holes;
  bq_amb_t *ambs;
  uint8_t *pac; /* forward-only 2-bit packed reference */
} bq_ref_t;

/* host copy of one FM-index half as stored on disk */
typedef struct {
  uint64_t primary, L2[5], seq_len, bwt_words, n_sa;
  int sa_intv;
  uint32_t *bwt;
  uint64_t *sa;
} bq_fm_t;

typedef struct {
  bq_fm_t fm[2]; /* [0] daughter, [1] parent */
  bq_ref_t ref;
} bq_index_t;

/* ---- reads: bseq1_t (lib/aln/bwa.h:52-61) ---- */
typedef struct {
  int l_seq, id;
  char *name, *comment, *barcode, *umi, *qual, *sam;
  uint8_t *seq;  /* nt4, after clipping */
  uint8_t *seq0; /* nt4, as read */
  int l_seq0, l_adaptor, clip5, clip3;
  /* batch readers place name / seq0 / qual of all reads of a batch in one slab (owned by the first read of the
   * batch) instead of three allocations per read; bq_reads_free knows both conventions */
  char *slab;
  uint8_t in_slab;
  /* phase 2 writes the SAM text of a batch into one buffer per worker thread; .sam then points into it
   * (sam_in_slab) and the buffers hang off the first read of the batch */
  uint8_t sam_in_slab;
  size_t sam_off; /* phase-2 workers format straight into their slab: offset of this read's text until the batch is done */
  size_t sam_len; /* length of .sam when phase 2 knows it (0 = unknown: use strlen), so that the sinks need not scan the text */
  int n_sam_slabs;
  char **sam_slabs;
} bq_read_t;

/* ---- alignment region: mem_alnreg_t (lib/aln/mem_alnreg.h:34-66) ----
 * Same fields and meanings; laid out in 128 bytes (two cache lines): a batch holds about 1.7 M of them and every stage of
 * phase 2 streams through them, so the flags are bytes, seedlen0 is 16 bits, and the reference's bookkeeping fields that
 * nothing reads (n_comp, sam_set, read_in_pair) are left out. */
typedef struct {
  int64_t rb, re;
  uint64_t hash;
  uint32_t *cigar; /* n_cigar words followed by the NUL-terminated MD string */
  int qb, qe, rid, score, truesc, sub, alt_sc, csub, sub_n, w, seedcov, secondary, secondary_all;
  int pos, flag, NM, n_cigar;
  float frac_rep;
  unsigned mapq;
  uint32_t ZC, ZR;
  /* batched phase-2 DP on the GPU (bsq_dp_*): the CIGAR job predicted for this region, 0 = none, else
   * 1 + (worker thread << 24 | index in that thread's job list) */
  uint32_t dp_job;
  int16_t seedlen0;
  uint8_t bss, parent, is_alt, is_rev, bss_u;
  uint8_t cigar_ext; /* .cigar points into the batch's result blob (not owned by the region) */
} bq_reg_t;

typedef struct {
  size_t n, m, n_pri;
  bq_reg_t *a;
  int pooled; /* a[] is a slice of a batch pool / a stack array: never freed, copied out by the first push beyond m */
} bq_regv_t;

typedef struct {
  size_t l, m;
  char *s;
} bq_str_t;

/* bq_core.c */
void bq_opt_init(bq_opt_t *o);
void bq_fill_scmat(int a, int b, int8_t mat[25]);
void bq_fill_scmat_bis(int a, int b, int ct, int8_t mat[25]);
void bq_opt_to_dev(const bq_opt_t *o, bsq_opt *d);
uint64_t bq_hash64(uint64_t key);
void bq_introsort(void *base, size_t n, size_t sz, int (*lt)(const void *, const void *));
int bq_pos2rid(const bq_ref_t *r, int64_t pos_f);
int64_t bq_depos(const bq_ref_t *r, int64_t pos, int *is_rev);
uint8_t *bq_get_seq(int64_t l_pac, const uint8_t *pac, int64_t beg, int64_t end, int64_t *len);
uint8_t *bq_fetch_seq(const bq_ref_t *r, int64_t *beg, int64_t mid, int64_t *end, int *rid);
int bq_global_align(int qlen, const uint8_t *query, int tlen, const uint8_t *target, const int8_t *mat, int o_del, int e_del, int o_ins,
                    int e_ins, int w, int *n_cigar, uint32_t **cigar);
typedef struct { int score, te, qe, score2, te2, tb, qb; } bq_swr_t;
#define BQ_XBYTE 0x10000
#define BQ_XSTOP 0x20000
#define BQ_XSUBO 0x40000
#define BQ_XSTART 0x80000
bq_swr_t bq_local_align(int qlen, uint8_t *query, int tlen, uint8_t *target, const int8_t *mat, int o_del, int e_del, int o_ins, int e_ins,
                        int xtra);
uint32_t *bq_gen_cigar(const int8_t mat[25], int o_del, int e_del, int o_ins, int e_ins, int w_, int64_t l_pac, const uint8_t *pac,
                       int l_query, uint8_t *query, int64_t rb, int64_t re, int *score, int *n_cigar, int *NM, uint32_t *ZC, uint32_t *ZR,
                       int *bss_u, uint8_t parent);
/* string buffer (kstring-like); inline: SAM formatting calls these ~100 times per record */
#include <stdlib.h>
#include <string.h>
static inline void bq_str_reserve(bq_str_t *s, size_t extra) {
  if (s->l + extra + 1 > s->m) {
    size_t m = s->m ? s->m : 64;
    while (m < s->l + extra + 1) m <<= 1;
    s->s = (char *)realloc(s->s, m);
    s->m = m;
  }
}
static inline void bq_kputsn(bq_str_t *s, const char *p, size_t n) { bq_str_reserve(s, n); memcpy(s->s + s->l, p, n); s->l += n; s->s[s->l] = 0; }
static inline void bq_kputs(bq_str_t *s, const char *p) { bq_kputsn(s, p, strlen(p)); }
static inline void bq_kputc(bq_str_t *s, int c) { bq_str_reserve(s, 1); s->s[s->l++] = (char)c; s->s[s->l] = 0; }
static inline void bq_kputl(bq_str_t *s, long v) { /* decimal without printf: numbers are a large share of the SAM text */
  char b[24];
  int n = 0;
  unsigned long u = v < 0 ? 0ul - (unsigned long)v : (unsigned long)v;
  do { b[n++] = (char)('0' + u % 10); u /= 10; } while (u);
  if (v < 0) b[n++] = '-';
  bq_str_reserve(s, (size_t)n);
  while (n) s->s[s->l++] = b[--n];
  s->s[s->l] = 0;
}
static inline void bq_kputw(bq_str_t *s, int v) { bq_kputl(s, v); }

/* bq_phase2.c */
void bq_merge_regions(const bq_opt_t *opt, const bq_ref_t *ref, const uint8_t *query, int l_query, bq_regv_t *regs);
bq_pestat_t bq_pestat(const bq_opt_t *opt, const bq_ref_t *ref, int n, const bq_regv_t *regs);
void bq_matesw(const bq_opt_t *opt, const bq_ref_t *ref, bq_pestat_t pes, bq_read_t s[2], bq_regv_t regs[2]);
void bq_mark_primary(const bq_opt_t *opt, bq_regv_t *regs, int64_t id);
int bq_approx_mapq_se(const bq_opt_t *opt, const bq_reg_t *a);
void bq_reg2sam_se(const bq_opt_t *opt, const bq_ref_t *ref, bq_read_t *s, bq_regv_t *regs, const char *rg_id);
void bq_reg2sam_pe(const bq_opt_t *opt, const bq_ref_t *ref, uint64_t id, bq_read_t s[2], bq_regv_t regs[2], bq_pestat_t pes,
                   const char *rg_id);
void bq_read_clipping(bq_read_t *s, const uint8_t *adaptor, int l_adaptor, const bq_opt_t *opt);
/* mem_process_seqs (lib/aln/bwamem.c:432-476) on top of the GPU aligner: fills seqs[i].sam */
int bq_process_seqs(const bq_opt_t *opt, bsq_aligner *al, bsq_dp *dp, const bq_ref_t *ref, int64_t n_processed, int n, bq_read_t *seqs,
                    const bq_pestat_t *pes0, const char *rg_id);

typedef struct bq_batch bq_batch_t;
/* GPU half (clipping, task list, bsq_align_phase1) and host half (merge, pestat, phase 2, SAM) of one batch */
bq_batch_t *bq_batch_prep(const bq_opt_t *opt, int64_t n_processed, int n, bq_read_t *seqs, int *rc);
int bq_batch_run(bsq_aligner *al, bsq_dp *dp, bq_batch_t *b);
void bq_batch_discard(bq_batch_t *b);
bq_batch_t *bq_batch_gpu(const bq_opt_t *opt, bsq_aligner *al, bsq_dp *dp, int64_t n_processed, int n, bq_read_t *seqs, int *rc);
/* Host phase 2 in two halves around the batched DP of the batch's bsq_dp context (NULL: everything on the host):
 *   _a  region merge, insert-size statistics, mate rescue (its local alignments as one bsq_dp_matesw batch), primary
 *       marking; predicts which regions will need a final CIGAR and submits them as one bsq_dp_cigar batch (asynchronous)
 *   _wait  blocks until the CIGARs are back
 *   _b  pairing, mapQ, SAM text (mem_alnreg_setSAM looks its CIGAR up; anything not predicted is done on the host)
 * bq_batch_finish = the three in a row.  A pipeline runs _a of batch k+1 between _wait and _b of batch k, so that the
 * CIGAR kernel of one batch overlaps the SAM formatting of the one before (bq_pipe.c). */
int bq_batch_finish_a(const bq_opt_t *opt, const bq_ref_t *ref, bq_batch_t *b, const bq_pestat_t *pes0);
int bq_batch_finish_wait(bq_batch_t *b);
void bq_batch_finish_b(const bq_opt_t *opt, const bq_ref_t *ref, bq_batch_t *b, const char *rg_id);
int bq_batch_finish(const bq_opt_t *opt, const bq_ref_t *ref, bq_batch_t *b, const bq_pestat_t *pes0, const char *rg_id);
void bq_batch_abandon(bq_batch_t *b); /* after a failed _a / _wait: releases the batch (not the reads) */
/* fn(ctx, i) for i in [0, n) on the phase-2 worker pool when it is idle (at most 8 threads), else on the calling thread */
void bq_parallel_for(int n_threads, long n, void (*fn)(void *ctx, long i), void *ctx);
/* counters of the batched DP since the start: [0] CIGAR jobs sent to the GPU, [1] setSAM calls answered from them, [2] setSAM
 * calls done on the host (not predicted / outside the kernel's limits), [3] mate-rescue alignments sent to the GPU,
 * [4] used from there, [5] done on the host */
void bq_dp_stats(int64_t out[6]);

/* bq_pipe.c: source -> prep | GPU | phase 2 -> sink, batches in order.  src returns the next batch (NULL / n <= 0 at the end);
 * sink receives the reads with .sam filled and frees them (n < 0: the batch failed, free only). */
typedef bq_read_t *(*bq_source_fn)(void *ctx, int *n);
typedef void (*bq_sink_fn)(void *ctx, bq_read_t *seqs, int n);
#define BQ_MAX_LANES 16 /* aligner contexts (GPUs) one pipeline can drive */
int bq_pipeline_run(const bq_opt_t *opt, const bq_ref_t *ref, bsq_aligner *const *als, bsq_dp *const *dps /* one per lane, or NULL */,
                    int n_al /* one lane per aligner context: batches round-robin */,
                    bq_source_fn src, void *src_ctx, bq_sink_fn sink, void *sink_ctx, const bq_pestat_t *pes0, const char *rg_id);

/* bq_io.c */
int bq_index_load(const char *prefix, bq_index_t *idx);
void bq_index_free(bq_index_t *idx);
int bq_index_to_device(const bq_index_t *idx, int device, bsq_index **out);
int bq_main_index(int argc, char **argv);
typedef struct bq_fastq bq_fastq_t;
bq_fastq_t *bq_fastq_open(const char *fn);
void bq_fastq_close(bq_fastq_t *f);
bq_read_t *bq_read_batch(int chunk_size, int has_bc, int keep_comment, int *n, bq_fastq_t *f1, bq_fastq_t *f2);
/* large per-batch buffers (read slabs, SAM slabs, region pools) are recycled between batches: fresh memory costs a
 * page fault per 4 KB, and under a hypervisor those faults are slow and serialise the worker threads */
void *bq_big_alloc(size_t n, size_t *cap);
void bq_big_free(void *p);
void bq_reads_free(bq_read_t *seqs, int n); /* reads of one batch incl. their .sam strings and the array itself */
void bq_print_sam_hdr(const bq_ref_t *ref, const char *hdr_line, const char *pg_line);
void bq_fatal(const char *fmt, ...);

/* bq_main.c */
int bq_main_align(int argc, char **argv);

extern int bq_verbose;
#endif
