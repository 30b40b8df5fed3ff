/* bsq.h -- C ABI of libbsq.so, the B200 (sm_100a) implementation of BISCUIT's two data-parallel
 * hot paths (bisulfite seed-and-extend alignment, methylation pileup).
 *
 * Plain C: pointers and sizes only.  All buffers named `h_*` or documented as "host" are HOST
 * memory owned by the caller; the library stages them to the GPU, runs the kernels and copies the
 * results back.  Every function returns 0 on success or a negative BSQ_E* code; nothing in the
 * library calls exit().  There is NO CPU implementation behind these entry points: if no CUDA
 * device is usable they return BSQ_ENODEV.
 *
 * The entry points replace these interfaces of the reference (zhou-lab/biscuit @ b682768):
 *   bsq_index_upload        <- bwa_idx_load_from_disk()            lib/aln/bwa.c:525-554 (device copy of bwaidx_t)
 *   bsq_align_phase1        <- bis_worker1() body up to mem_merge_regions   lib/aln/bwamem.c:311-375
 *                              (= mem_align1_core x2: mem_chain, mem_chain_flt, mem_chain2region)
 *   bsq_collect_intv        <- mem_collect_intv()                  lib/aln/memchain.c:50-106
 *                              (bwt_smem1a lib/aln/bwt.c:307, bwt_seed_strategy1 bwt.c:376)
 *   bsq_occ4                <- bwt_occ4()                          lib/aln/bwt.c:173-200
 *   bsq_sa_lookup           <- bwt_sa()                            lib/aln/bwt.c:87-97
 *   bsq_extend_batch        <- ksw_extend2()                       lib/aln/ksw.c:380-479
 *   bsq_dp_cigar_*          <- mem_alnreg_setSAM() band loop       lib/aln/mem_alnreg.c:40-123 around
 *                              bis_bwa_gen_cigar2()                 lib/aln/bwa.c:290-428 (ksw_global2, ksw.c:504-606)
 *   bsq_dp_matesw_*         <- ksw_align2()                        lib/aln/ksw.c:343-365, called by mem_matesw()
 *                                                                  lib/aln/mem_alnreg.c:395-493
 *   bsq_plp_*               <- process_func() hot loop + plp_getcnts  src/pileup.c:707-831, :372-387
 */
#ifndef BSQ_H
#define BSQ_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BSQ_OK 0
#define BSQ_ENODEV (-1)    /* no usable CUDA device / CUDA runtime error */
#define BSQ_EINVAL (-2)    /* bad argument (e.g. read longer than BSQ_MAX_READ_LEN) */
#define BSQ_EOVERFLOW (-3) /* a per-read capacity was exceeded; nothing was silently dropped */
#define BSQ_ENOMEM (-4)

#define BSQ_MAX_READ_LEN 256
#define BSQ_MAX_INTV 384

/* same layout as the reference's bwtintv_t (lib/aln/bwt.h:80-82) */
typedef struct { uint64_t x[3], info; } bsq_intv;

/* phase-1 alignment region: the fields of mem_alnreg_t (lib/aln/mem_alnreg.h:34-66) that are
 * set by mem_chain2region1 (lib/aln/memchain.c:742-871) */
typedef struct {
  int64_t rb, re;
  int32_t qb, qe;
  int32_t rid, score, truesc, w;
  int32_t seedcov, seedlen0;
  float frac_rep;
  uint8_t bss, parent, pad_[2];
} bsq_reg;

/* alignment options that reach the device: subset of mem_opt_t (lib/aln/bwamem.h:54-124),
 * defaults from mem_opt_init (lib/aln/bwamem.c:77-128) via bsq_opt_default() */
typedef struct {
  int32_t a, b, o_del, e_del, o_ins, e_ins, pen_clip5, pen_clip3, w, zdrop;
  int32_t min_seed_len, split_width, max_occ, max_chain_gap, min_chain_weight, max_chain_extend;
  int32_t max_mem_intv, split_len, self_ovlp, bsstrand;
  float mask_level, drop_ratio;
  int8_t ctmat[25], gamat[25];
  int8_t pad_[2];
} bsq_opt;

/* host view of a loaded index (bwaidx_t, lib/aln/bwa.h:42-50).  which: 0 = daughter (G>A,
 * <prefix>.dau.*), 1 = parent (C>T, <prefix>.par.*) as in bwa.c:535-536 */
typedef struct {
  const uint32_t *bwt[2];  /* body of the .bwt file after the 40-byte header */
  uint64_t bwt_words[2];   /* number of u32 words in bwt[] */
  uint64_t primary[2];
  uint64_t L2[2][5];       /* L2[.][0] = 0 */
  uint64_t seq_len;        /* 2 * l_pac */
  const uint64_t *sa[2];   /* n_sa entries, sa[0] = (uint64_t)-1 */
  uint64_t n_sa[2];
  int32_t sa_intv[2];
  const uint8_t *pac;      /* .bis.pac body, (l_pac+3)/4 bytes */
  int64_t l_pac;
  int32_t n_seqs;
  const int64_t *ann_offset;
  const int32_t *ann_len;
  const int32_t *ann_is_alt;
} bsq_index_desc;

typedef struct bsq_index bsq_index; /* device-resident index, opaque */

void bsq_opt_default(bsq_opt *opt);
const char *bsq_strerror(int code);
const char *bsq_last_error(void); /* text of the last CUDA error seen by this thread */

int bsq_index_upload(const bsq_index_desc *desc, int device, bsq_index **out);
void bsq_index_free(bsq_index *idx);

/* `biscuit index` on the GPU (replaces bwt_bwtgen/bwt_pac2bwt + bwt_bwtupdate_core + bwt_cal_sa,
 * lib/aln/bwtindex.c:258-344): builds both converted FM-indices from the forward 2-bit pac (host
 * buffer, N already replaced, i.e. the content of <prefix>.bis.pac) and leaves them resident on
 * `device`.  bsq_index_sizes/_download return the arrays exactly as the reference stores them in
 * <prefix>.{par,dau}.{bwt,sa} (which: 0 = dau, 1 = par). */
int bsq_index_build(const uint8_t *pac, int64_t l_pac, int32_t n_seqs, const int64_t *ann_offset, const int32_t *ann_len,
                    const int32_t *ann_is_alt, int device, bsq_index **out);
int bsq_index_sizes(const bsq_index *idx, uint64_t bwt_words[2], uint64_t n_sa[2], uint64_t primary[2], uint64_t L2[10],
                    int64_t stats[3]);
int bsq_index_download(const bsq_index *idx, int which, uint32_t *bwt, uint64_t *sa);

/* ---- kernel-level entry points (SoA, host buffers) ---- */

/* cnt[4*i..4*i+3] = ranks of A,C,G,T up to BWT position k[i] in index `which` */
int bsq_occ4(const bsq_index *idx, int which, int64_t n, const uint64_t *k, uint64_t *cnt);

/* pos[i] = text position of BWT rank k[i] */
int bsq_sa_lookup(const bsq_index *idx, int which, int64_t n, const uint64_t *k, uint64_t *pos);

/* SMEM seeding of n_tasks (read, conversion) tasks.  seqs: UNCONVERTED nt4 reads, task t at
 * seqs + t*stride, lens[t] bases; parent[t] selects the conversion (1: C>T vs parent index,
 * 0: G>A vs daughter index).  out: n_tasks * BSQ_MAX_INTV intervals; n_out[t] = count. */
int bsq_collect_intv(const bsq_index *idx, const bsq_opt *opt, int64_t n_tasks, const uint8_t *seqs, int32_t stride,
                     const int32_t *lens, const uint8_t *parent, bsq_intv *out, int32_t *n_out);

/* Batched ksw_extend2.  Job j: query = qbuf + qoff[j] (qlen[j] nt4 codes), target = tbuf + toff[j]
 * (tlen[j] codes), matrix = is_parent[j] ? opt->ctmat : opt->gamat, band w[j], initial score h0[j],
 * end bonus = opt->pen_clip5.  out[6*j..] = {score, qle, tle, gtle, gscore, max_off}. */
int bsq_extend_batch(const bsq_opt *opt, int64_t n_jobs, const uint8_t *qbuf, const int64_t *qoff, const int32_t *qlen,
                     const uint8_t *tbuf, const int64_t *toff, const int32_t *tlen, const uint8_t *is_parent,
                     const int32_t *w, const int32_t *h0, int32_t *out);

/* ---- phase 1 of the aligner for a batch of reads ---- */

typedef struct bsq_aligner bsq_aligner; /* per-GPU context: streams + reusable device buffers */

int bsq_aligner_create(const bsq_index *idx, const bsq_opt *opt, bsq_aligner **out);
void bsq_aligner_destroy(bsq_aligner *al);

/* Seed -> chain -> filter -> extend for n_tasks (read, conversion) tasks (same task layout as
 * bsq_collect_intv).  Regions of task t are written to regs[reg_off[t] .. reg_off[t+1]) in the
 * order mem_chain2region (lib/aln/memchain.c:873-904) produces them.  `regs` is malloc()ed by the
 * library (caller frees with bsq_free); reg_off has n_tasks+1 entries (caller-allocated). */
int bsq_align_phase1(bsq_aligner *al, int64_t n_tasks, const uint8_t *seqs, int32_t stride, const int32_t *lens,
                     const uint8_t *parent, bsq_reg **regs, int64_t *reg_off);
void bsq_free(void *p);

/* The same call split in three so that a caller can pipeline batches (and so that bench.py can time the
 * device part with inputs already resident in HBM): stage = host->device copy of the reads,
 * run = the five kernels (returns the number of regions), fetch = device->host copy into caller
 * buffers (regs: *n_regs entries, reg_off: n_tasks+1).  bsq_host_alloc returns page-locked host
 * memory for these buffers. */
int bsq_aligner_stage(bsq_aligner *al, int64_t n_tasks, const uint8_t *seqs, int32_t stride, const int32_t *lens,
                      const uint8_t *parent);
int bsq_aligner_run(bsq_aligner *al, int64_t *n_regs);
int bsq_aligner_fetch(bsq_aligner *al, bsq_reg *regs, int64_t *reg_off);
/* Deferred fetch for pipelined callers.  bsq_aligner_run leaves its regions in one of two device-side result slots,
 * alternating from run to run.  bsq_aligner_result_slot names the slot of the last run (with its task and region
 * counts) and CLAIMS it: the run that would overwrite it blocks until the slot has been fetched or released.
 * bsq_aligner_fetch_slot copies the slot to the host on a stream of its own and releases it; it may be called from
 * another thread while later batches are staged and run on the same context.  A claimed slot that will not be fetched
 * (a failed batch) must be given back with bsq_aligner_release_slot. */
int bsq_aligner_result_slot(bsq_aligner *al, int *slot, int64_t *n_tasks, int64_t *n_regs);
int bsq_aligner_fetch_slot(bsq_aligner *al, int slot, bsq_reg *regs, int64_t *reg_off);
int bsq_aligner_release_slot(bsq_aligner *al, int slot);
/* How this library's calls wait for the GPU: 0 = spinning (cudaStreamSynchronize; lowest latency, holds a core), 1 = sleeping
 * on an event (gives the core to the caller's other threads).  Process-wide; BSQ_SPIN_WAIT=0/1 in the environment overrides. */
void bsq_set_wait_mode(int blocking);
int bsq_host_alloc(void **p, size_t bytes);
/* work counters for the roofline arithmetic: only the instrumented build (libbsq_count.so) has them,
 * libbsq.so returns BSQ_EINVAL.  out[0]=64-B index blocks fetched, [1]=bwt_extend calls,
 * [2]=ksw_extend2 calls, [3]=DP cells, [4]=reference bases decoded, [5]=32-byte derived rank blocks fetched by the
 * seeding kernels ([0] then counts what the same extensions touch in the reference's 64-byte layout, bwt.c:204-236) */
int bsq_work_counters(uint64_t *out, int n, int reset);
/* Measurement aid for the roofline: rate of dependent random 32-byte gathers (one DRAM sector each, the access pattern
 * of bwt_occ over the derived rank blocks) from a scratch buffer of `bytes` bytes on `device`, all SMs busy.
 * *gathers_per_s = sectors per second.  Not used by any alignment path. */
int bsq_measure_gather(int device, uint64_t bytes, double *gathers_per_s);
void bsq_host_free(void *p);

/* counters of the last bsq_align_phase1 call (for the roofline arithmetic in bench.py):
 * c[0]=tasks c[1]=index halves with a resident full suffix array (0..2: SA lookups are then one 8-byte read,
 * no LF walk) c[2]=seeds(SA lookups) c[3]=unused c[4]=regions
 * c[5..8] = device time of the seed / expand+sa / chain / extend kernels, c[9] = scans+compaction,
 * c[10] = first kernel to last kernel, all in integer microseconds (CUDA events on the aligner's stream);
 * instrumented build only: c[11..13] = running count of 64-B index blocks fetched before k_seed / after
 * k_seed / after k_sa; c[14] = tasks whose chaining went through the exact fallback kernel,
 * c[15] = microseconds of k_chain_warp alone (c[7] = both chaining kernels) */
int bsq_aligner_counters(const bsq_aligner *al, int64_t *c, int n);

/* ---- the two dynamic-programming steps of phase 2 (after pairing), batched on the GPU ----
 *   bsq_dp_cigar   <- mem_alnreg_setSAM's band loop  lib/aln/mem_alnreg.c:40-75  around
 *                     bis_bwa_gen_cigar2             lib/aln/bwa.c:290-428 (ksw_global2, lib/aln/ksw.c:504-606; MD/NM/ZC/ZR)
 *   bsq_dp_matesw  <- ksw_align2                     lib/aln/ksw.c:343-365 (ksw_u8 :111-220, ksw_i16 :222-334) as called
 *                     by mem_matesw                  lib/aln/mem_alnreg.c:395-493
 * A bsq_dp context belongs to one device; it has its own stream and buffers and may be driven by another host thread
 * than the bsq_aligner of the same device.  Reads are the nt4 rows of a batch (the rows handed to bsq_aligner_stage
 * serve: row r at seqs + r*stride, lens[r] bases); jobs refer to them by row. */
typedef struct bsq_dp bsq_dp;

/* one final CIGAR (+MD) = one mem_alnreg_setSAM call: query = bases [qb,qe) of `row` (codes > 4 read as 4), reference
 * [rb,re) in forward-reverse coordinates, first band w, doubled up to three times while the global score stays below
 * truesc - a (mem_alnreg.c:60-70); clip5 / clip3 = soft-clip lengths to put in front / behind (0 = none) */
typedef struct {
  int64_t rb, re;
  int32_t row, qb, qe, w, truesc, clip5, clip3;
  uint8_t parent, pad_[3];
} bsq_cigar_job;

/* n_cigar < 0: the job is outside the kernel's limits (reference span > 1024) and must be done by the caller;
 * n_cigar == 0: bis_bwa_gen_cigar2 returned no alignment.  Otherwise blob[off .. off+n_cigar) holds the final CIGAR
 * words (BAM encoding; a leading / trailing deletion already removed, lead_del = length of the removed leading one;
 * clips added) directly followed by the NUL-terminated MD text. */
typedef struct {
  int32_t n_cigar, NM, ZC, ZR, score, lead_del, bss_u;
  uint32_t off; /* in 4-byte words */
} bsq_cigar_res;

/* one ksw_align2 call of mate rescue: query = reverse complement of `row` (l_ms = its length), target = reference
 * [rb,re) in forward-reverse coordinates, matrix = use_ga ? gamat : ctmat, xtra as in ksw.h:34-37 */
typedef struct {
  int64_t rb, re;
  int32_t row, xtra;
  uint8_t use_ga, pad_[7];
} bsq_matesw_job;

/* kswr_t (lib/aln/ksw.h:39-43) */
typedef struct { int32_t score, te, qe, score2, te2, tb, qb, pad_; } bsq_matesw_res;

int bsq_dp_create(const bsq_index *idx, const bsq_opt *opt, bsq_dp **out);
void bsq_dp_destroy(bsq_dp *dp);
/* host->device copy of the rows (asynchronous on the context's stream; the host buffer must stay valid until the next
 * bsq_dp_*_wait or bsq_dp_sync) */
int bsq_dp_set_reads(bsq_dp *dp, int64_t n_rows, const uint8_t *seqs, int32_t stride, const int32_t *lens);
int bsq_dp_sync(bsq_dp *dp);
/* submit = H2D of the jobs + kernel + D2H of the results, all asynchronous; wait = block until they are back.
 * res (n_jobs entries) is caller memory; the CIGAR/MD blob belongs to the context and stays valid until the next
 * bsq_dp_cigar_wait on it (a new submission does not disturb it).  One submission of each kind may be in flight
 * per context. */
int bsq_dp_cigar_submit(bsq_dp *dp, int64_t n_jobs, const bsq_cigar_job *jobs, bsq_cigar_res *res);
int bsq_dp_cigar_wait(bsq_dp *dp, const uint32_t **blob, int64_t *blob_words);
int bsq_dp_matesw_submit(bsq_dp *dp, int64_t n_jobs, const bsq_matesw_job *jobs, bsq_matesw_res *res);
int bsq_dp_matesw_wait(bsq_dp *dp);
/* c[0] = CIGAR jobs, c[1] = of those ungapped, c[2] = DP cells of ksw_global2, c[3] = microseconds of the last k_cigar,
 * c[4] = mate-rescue jobs, c[5] = microseconds of the last k_matesw, c[6] = blob re-runs (capacity grown) */
int bsq_dp_counters(const bsq_dp *dp, int64_t *c, int n);

/* ================= pileup (methylation caller) =================
 * Replaces the per-window hot loop of process_func (src/pileup.c:707-831: read filters, bisulfite-strand
 * inference, mate-overlap skip, per-base retention/conversion events), plp_getcnts (src/pileup.c:372-387)
 * and the integer part of plp_format (src/pileup.c:415-485: ambiguity redistribution, top mutant, emit
 * rule, methcallable, 5-mer cytosine context of src/bisc_utils.c:33-74).  Genotype likelihoods (double,
 * utils/stats.h) and VCF/BED text stay on the host (biscuit_b200/host).
 *
 * Input = the BAM records of one contig as structure-of-arrays, i.e. exactly the fields of bam1_core_t
 * plus CIGAR / 4-bit SEQ / QUAL / the tags the reference reads (NM AS MC YD ZS XG), coordinate sorted. */
typedef struct {
  int64_t n_reads;
  const int32_t *pos;        /* 0-based leftmost reference position */
  const int32_t *mpos;       /* mate position, 0-based */
  const int32_t *mate_rlen;  /* reference length of the mate from the MC tag, -1 if absent */
  const int32_t *l_qseq;
  const int32_t *nm;         /* NM tag, INT32_MIN if absent */
  const int32_t *as;         /* AS tag, INT32_MIN if absent */
  const uint16_t *flag;
  const uint8_t *mapq;
  const int8_t *bss_tag;     /* 0: YD:f / ZS:+ / XG:CT, 1: YD:r / ZS:- / XG:GA, -1: none (infer from the read) */
  const uint8_t *sid;        /* sample (BAM file) index, < n_bams */
  const int32_t *n_cigar;
  const int64_t *cigar_off;
  const uint32_t *cigar;     /* BAM encoding: len<<4 | op */
  const int64_t *seq_off;    /* byte offset into seq[]; 4-bit packed as in BAM */
  const uint8_t *seq;
  const int64_t *qual_off;
  const uint8_t *qual;
} bsq_plp_reads;

/* meth_filter_t + the pileup_conf_t switches that affect counting (src/bisc_utils.h:95-113, pileup.h:49-63) */
typedef struct {
  int32_t min_base_qual, min_read_len, min_dist_end_5p, min_dist_end_3p, min_mapq, min_score, max_nm, max_retention;
  int32_t filter_ppair, filter_secondary, filter_duplicate, filter_qcfail, filter_doublecnt;
  int32_t ambi_redist, verbose, is_nome;
} bsq_plp_conf;

/* one emitted locus x one sample (n_bams consecutive records per locus) */
typedef struct {
  int32_t pos, dp;           /* 1-based position; DP = all events of the sample (pileup.c:572) */
  int32_t meth[3];           /* retention, conversion, NA (after base filters) */
  int32_t base[7];           /* A C G T N Y R (after base filters) */
  int32_t base_redist[7];    /* after redistribute_cnts */
  uint8_t rb_code;           /* reference base 0..3 */
  int8_t cm1;                /* top mutant base code or -1 */
  uint8_t ctx, methcallable; /* cytosine_context_t (6 = NA); methcallable of this sample */
  char n5[5];
  uint8_t any_callable, pad_[2];
} bsq_plp_rec;

typedef struct bsq_plp bsq_plp;

void bsq_plp_conf_default(bsq_plp_conf *c);
int bsq_plp_create(int device, int n_bams, bsq_plp **out);
void bsq_plp_destroy(bsq_plp *p);
/* upload one contig: nt4 codes (A0 C1 G2 T3 N4), ref_len bases */
int bsq_plp_set_contig(bsq_plp *p, const uint8_t *ref_nt4, int32_t ref_len);
/* host->device copy of the reads; run = kernels over loci [beg,end) (1-based, end exclusive, clipped so that
 * the last base of the contig is never piled, pileup.c:1191-1196), returns the number of emitted loci;
 * fetch = device->host copy of n_loci*n_bams records */
int bsq_plp_stage(bsq_plp *p, const bsq_plp_reads *reads);
int bsq_plp_run(bsq_plp *p, const bsq_plp_conf *conf, int32_t beg, int32_t end, int64_t *n_loci);
int bsq_plp_fetch(bsq_plp *p, bsq_plp_rec *out);
/* c[0] = reads staged, c[1] = loci scanned, c[2] = loci emitted, c[3] = base events (after read filters),
 * c[4..5] = device microseconds of the event kernel / the per-locus kernel of the last run */
int bsq_plp_counters(const bsq_plp *p, int64_t *c, int n);

#ifdef __cplusplus
}
#endif
#endif
