/* bq_main.c -- `biscuit index | align | version` command lines (reference: src/main.c:105-159,
 * lib/aln/align.c:226-598).  `align` keeps the reference's options, batch boundaries (chunk_size x threads
 * bases, even read count: bwa.c:842, align.c:576) and output order; batches go to the GPU through
 * bq_process_seqs. */
#include <ctype.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>
#include <unistd.h>
#include "bq_plp.h"

#include <pthread.h>

/* source and sink of the batch pipeline (bq_pipe.c): FASTQ batches in, SAM text out (cf. process() steps 0 and 2,
 * lib/aln/align.c:70-167) */
typedef struct { bq_opt_t *opt; bq_fastq_t *f1, *f2; int chunk, copy_comment; } src_ctx_t;

static bq_read_t *fastq_source(void *ctx, int *n) {
  src_ctx_t *p = ctx;
  bq_read_t *seqs = bq_read_batch(p->chunk, p->opt->has_bc, p->copy_comment, n, p->f1, p->f2);
  if (seqs && *n > 0 && bq_verbose >= 3) {
    int64_t size = 0;
    for (int i = 0; i < *n; ++i) size += seqs[i].l_seq;
    fprintf(stderr, "[M::process] read %d sequences (%ld bp)...\n", *n, (long)size);
  }
  return seqs;
}

/* The FASTQ reader on a thread of its own, two batches ahead of the pipeline's preparation stage (which used to parse
 * and prepare on one thread: ~0.7 us per read, below one GPU's rate).  Batches come out in file order. */
typedef struct {
  src_ctx_t *sc;
  pthread_t th;
  pthread_mutex_t mu;
  pthread_cond_t cv;
  bq_read_t *seqs[2];
  int n[2], count, head, stop;
} ahead_src_t;
static void *ahead_main(void *arg) {
  ahead_src_t *a = arg;
  for (;;) {
    int n = 0;
    bq_read_t *seqs = fastq_source(a->sc, &n);
    pthread_mutex_lock(&a->mu);
    while (a->count == 2 && !a->stop) pthread_cond_wait(&a->cv, &a->mu);
    if (a->stop) { pthread_mutex_unlock(&a->mu); if (seqs && n > 0) bq_reads_free(seqs, n); else free(seqs); return 0; }
    const int at = (a->head + a->count) & 1;
    a->seqs[at] = seqs; a->n[at] = n; a->count++;
    pthread_cond_broadcast(&a->cv);
    pthread_mutex_unlock(&a->mu);
    if (!seqs || n <= 0) return 0; /* the end marker has been queued */
  }
}
static bq_read_t *ahead_source(void *ctx, int *n) {
  ahead_src_t *a = ctx;
  pthread_mutex_lock(&a->mu);
  while (a->count == 0) pthread_cond_wait(&a->cv, &a->mu);
  bq_read_t *seqs = a->seqs[a->head];
  *n = a->n[a->head];
  if (seqs && *n > 0) { a->head ^= 1; a->count--; } /* the end marker stays queued: every further call sees it */
  pthread_cond_broadcast(&a->cv);
  pthread_mutex_unlock(&a->mu);
  if (!seqs || *n <= 0) { *n = 0; return 0; }
  return seqs;
}
static void ahead_stop(ahead_src_t *a) {
  pthread_mutex_lock(&a->mu);
  a->stop = 1;
  pthread_cond_broadcast(&a->cv);
  pthread_mutex_unlock(&a->mu);
  pthread_join(a->th, 0);
  for (; a->count > 0; a->head ^= 1, a->count--) { /* batches read but never taken (the run failed) */
    if (a->seqs[a->head] && a->n[a->head] > 0) bq_reads_free(a->seqs[a->head], a->n[a->head]); else free(a->seqs[a->head]);
  }
}

static void sam_sink(void *ctx, bq_read_t *seqs, int n) {
  const int ok = n >= 0;
  if (n < 0) n = -n;
  if (ok)
    for (int i = 0; i < n; ++i)
      if (seqs[i].sam) { if (seqs[i].sam_len) fwrite(seqs[i].sam, 1, seqs[i].sam_len, stdout); else fputs(seqs[i].sam, stdout); }
  bq_reads_free(seqs, n);
  if (ok && bq_verbose >= 3) fprintf(stderr, "[M::mem_process_seqs] Processed %d reads\n", n);
}

/* -p: interleaved input whose neighbours with equal names are pairs, everything else single-end (bseq_classify,
 * bwa.c:119-138, and the MEM_F_SMARTPE branch of process(), align.c:109-143).  Each batch becomes up to two
 * mem_process_seqs-style calls: the single-end reads first (no insert-size prior), then the pairs with
 * n_processed advanced by the number of single-end reads; SAM lines leave in input order.  The sub-batches are
 * shallow copies of the batch's reads (names / sequences stay owned by the batch, the SAM text by the copies). */
static void free_sub_sam(bq_read_t *sub, int n) {
  if (n <= 0) return;
  for (int i = 0; i < n; ++i) if (!sub[i].sam_in_slab) free(sub[i].sam);
  for (int k = 0; k < sub[0].n_sam_slabs; ++k) bq_big_free(sub[0].sam_slabs[k]);
  free(sub[0].sam_slabs);
}

static int align_smart_pairing(const bq_opt_t *opt, const bq_ref_t *ref, bsq_aligner *al, bsq_dp *dp, src_ctx_t *sc, const bq_pestat_t *pes0,
                               const char *rg_id) {
  int64_t n_processed = 0;
  for (;;) {
    int n = 0, i, has_last, m[2] = {0, 0};
    bq_read_t *seqs = fastq_source(sc, &n);
    if (!seqs || n <= 0) { free(seqs); break; }
    bq_read_t *sep[2];
    sep[0] = calloc((size_t)n + 1, sizeof(bq_read_t)); sep[1] = calloc((size_t)n + 1, sizeof(bq_read_t));
    for (i = 1, has_last = 1; i < n; ++i) {
      if (has_last) {
        if (strcmp(seqs[i].name, seqs[i - 1].name) == 0) { sep[1][m[1]++] = seqs[i - 1]; sep[1][m[1]++] = seqs[i]; has_last = 0; }
        else sep[0][m[0]++] = seqs[i - 1];
      } else has_last = 1;
    }
    if (has_last) sep[0][m[0]++] = seqs[i - 1];
    if (bq_verbose >= 3) fprintf(stderr, "[bseq_classify] %d SE sequences; %d PE sequences\n", m[0], m[1]);
    const char **sam = calloc((size_t)n, sizeof(char *));
    for (int k = 0; k < 2; ++k) {
      if (!m[k]) continue;
      bq_opt_t tmp = *opt;
      if (k) tmp.flag |= BQ_F_PE; else tmp.flag &= ~BQ_F_PE;
      for (i = 0; i < m[k]; ++i) { sep[k][i].slab = 0; sep[k][i].sam = 0; sep[k][i].sam_slabs = 0; sep[k][i].n_sam_slabs = 0; sep[k][i].sam_in_slab = 0; }
      const int rc = bq_process_seqs(&tmp, al, dp, ref, n_processed + (k ? m[0] : 0), m[k], sep[k], k ? pes0 : 0, rg_id);
      if (rc) return rc;
      for (i = 0; i < m[k]; ++i) sam[sep[k][i].id] = sep[k][i].sam;
    }
    for (i = 0; i < n; ++i) if (sam[i]) fputs(sam[i], stdout);
    if (bq_verbose >= 3) fprintf(stderr, "[M::mem_process_seqs] Processed %d reads\n", n);
    free(sam);
    free_sub_sam(sep[0], m[0]); free_sub_sam(sep[1], m[1]);
    free(sep[0]); free(sep[1]);
    n_processed += n;
    bq_reads_free(seqs, n);
  }
  return 0;
}

static double now(void) { struct timeval tv; gettimeofday(&tv, 0); return tv.tv_sec + tv.tv_usec * 1e-6; }

static int align_usage(void) {
  fprintf(stderr, "\nUsage: biscuit align [options] <fai-index base> <in1.fq> [in2.fq]\n\n"
                  "Options follow `biscuit align` of BISCUIT %s: -@ -b -f -k -w -d -r -y -c -D -W -m -S -P -e -9 -A -B -O -E -L -U\n"
                  "    -1 -2 -i -R -H -j -q -T -g -a -C -V -Y -M -I -v -J -K -z -5 -3 -p, plus\n    -G LIST  CUDA device(s), e.g. 0, 0-7 or 0,2,5: FASTQ batches are dealt to them round-robin [0]\n\n", BQ_VERSION);
  return 1;
}

static char *escape_hdr(char *s) { /* bwa_escape, bwa.c:686-700 */
  char *p, *q;
  for (p = q = s; *p; ++p) {
    if (*p == '\\') {
      ++p;
      if (*p == 't') *q++ = '\t'; else if (*p == 'n') *q++ = '\n'; else if (*p == 'r') *q++ = '\r'; else if (*p == '\\') *q++ = '\\';
    } else *q++ = *p;
  }
  *q = 0;
  return s;
}

static char *insert_header(const char *s, char *hdr) { /* bwa_insert_header, bwa.c:731-743 */
  size_t len = 0;
  if (s == 0 || s[0] != '@') return hdr;
  if (hdr) { len = strlen(hdr); hdr = realloc(hdr, len + strlen(s) + 2); hdr[len++] = '\n'; strcpy(hdr + len, s); }
  else hdr = strdup(s);
  escape_hdr(hdr + len);
  return hdr;
}

static void infer_alt(bq_ref_t *r) { /* infer_alt_chromosomes, align.c:184-224 */
  int i, n, found[25];
  for (i = 0; i < r->n_seqs; ++i) if (r->anns[i].is_alt) return;
  memset(found, 0, sizeof found);
  for (i = 0; i < r->n_seqs; ++i) {
    const char *nm = r->anns[i].name;
    if (strncmp(nm, "chr", 3) != 0) continue;
    if (strlen(nm) == 4) {
      int c = toupper((unsigned char)nm[3]);
      if (c == 'X') found[22] = 1; else if (c == 'Y') found[23] = 1; else if (c == 'M') found[24] = 1;
      else if (isdigit(c)) { n = nm[3] - '0'; if (n > 0 && n <= 22) found[n - 1] = 1; }
    } else if (strlen(nm) == 5 && isdigit((unsigned char)nm[3]) && isdigit((unsigned char)nm[4])) {
      n = atoi(nm + 3);
      if (n > 0 && n <= 22) found[n - 1] = 1;
    }
  }
  for (i = n = 0; i < 25; ++i) if (found[i]) ++n;
  if (n < 20) return;
  for (i = 0; i < r->n_seqs; ++i) {
    const char *nm = r->anns[i].name;
    if (strncmp(nm, "chrUn", 5) == 0 || strstr(nm, "_random") || strstr(nm, "_hap") || strstr(nm, "_alt")) r->anns[i].is_alt = 1;
  }
}

int bq_main_align(int argc, char **argv) {
  bq_opt_t opt, set;
  bq_opt_init(&opt);
  opt.flag |= BQ_F_NO_MULTI; /* align.c:334 */
  memset(&set, 0, sizeof set);
  int c, i, ignore_alt = 0, auto_alt = 1, copy_comment = 0, no_hdr = 0, smart_pe = 0;
  int devices[BQ_MAX_LANES] = {0}, n_dev = 1;
  char *p, *rg_line = 0, *hdr_line = 0, *seq1 = 0, *seq2 = 0, rg_id[256] = "";
  bq_pestat_t *pes0 = 0;
  const uint8_t *t4 = 0;
  { static uint8_t t[256]; memset(t, 4, 256); t['A'] = t['a'] = 0; t['C'] = t['c'] = 1; t['G'] = t['g'] = 2; t['T'] = t['t'] = 3; t['-'] = 5; t4 = t; }
  if (argc < 2) return align_usage();
  while ((c = getopt(argc, argv, ":@:1:2:3:5:9ab:c:d:ef:g:hijk:m:pqr:s:v:w:y:z:A:B:CD:E:FG:H:I:J:K:L:MN:O:PQ:R:ST:U:VW:X:Y")) >= 0) {
    if (c == 'k') { opt.min_seed_len = atoi(optarg); set.min_seed_len = 1; }
    else if (c == '1') seq1 = strdup(optarg);
    else if (c == '2') seq2 = strdup(optarg);
    else if (c == 'b') opt.parent = (uint8_t)atoi(optarg);
    else if (c == 'f') opt.bsstrand = (uint8_t)atoi(optarg);
    else if (c == 'i') auto_alt = 0;
    else if (c == 'w') { opt.w = atoi(optarg); set.w = 1; }
    else if (c == 'A') { opt.a = atoi(optarg); set.a = 1; }
    else if (c == 'B') { opt.b = atoi(optarg); set.b = 1; }
    else if (c == 'T') { opt.T = atoi(optarg); set.T = 1; }
    else if (c == 'U') { opt.pen_unpaired = atoi(optarg); set.pen_unpaired = 1; }
    else if (c == '@') { opt.n_threads = atoi(optarg); if (opt.n_threads < 1) opt.n_threads = 1; }
    else if (c == 'P') opt.flag |= BQ_F_NOPAIRING;
    else if (c == 'a') opt.flag |= BQ_F_ALL;
    else if (c == 'q') opt.flag |= BQ_F_KEEP_SUPP_MAPQ;
    else if (c == 'M') opt.flag |= BQ_F_NO_MULTI;
    else if (c == 'S') opt.flag |= BQ_F_NO_RESCUE;
    else if (c == 'e') opt.flag |= BQ_F_SELF_OVLP;
    else if (c == 'F') no_hdr = 1;
    else if (c == 'Y') opt.flag |= BQ_F_SOFTCLIP;
    else if (c == 'V') opt.flag |= BQ_F_REF_HDR;
    else if (c == 'c') opt.max_occ = (uint32_t)atoi(optarg);
    else if (c == 'd') { opt.zdrop = atoi(optarg); set.zdrop = 1; }
    else if (c == 'v') bq_verbose = atoi(optarg);
    else if (c == 'j') ignore_alt = 1;
    else if (c == 'r') opt.split_factor = (float)atof(optarg);
    else if (c == 'D') opt.drop_ratio = (float)atof(optarg);
    else if (c == 'm') opt.max_matesw = atoi(optarg);
    else if (c == 's') opt.split_width = atoi(optarg);
    else if (c == 'N') opt.max_chain_extend = (uint32_t)atoi(optarg);
    else if (c == 'W') opt.min_chain_weight = atoi(optarg);
    else if (c == 'y') opt.max_mem_intv = (uint64_t)atol(optarg);
    else if (c == 'C') copy_comment = 1;
    else if (c == 'G') { /* the reference's hidden -G (max_chain_gap) is not exposed; here: CUDA device(s): "3", "0-7", "0,2,5" */
      n_dev = 0;
      for (p = optarg; *p && n_dev < BQ_MAX_LANES;) {
        const int a = (int)strtol(p, &p, 10);
        int b = a;
        if (*p == '-') b = (int)strtol(p + 1, &p, 10);
        for (int d = a; d <= b && n_dev < BQ_MAX_LANES; ++d) devices[n_dev++] = d;
        if (*p == ',') ++p; else break;
      }
      if (n_dev < 1) { devices[0] = 0; n_dev = 1; }
    }
    else if (c == 'J' || c == 'K') {
      int l = (int)strlen(optarg);
      uint8_t *a = calloc((size_t)l + 1, 1);
      for (i = 0; i < l; ++i) a[i] = t4[(unsigned char)optarg[i]];
      if (c == 'J') { opt.adaptor1 = a; opt.l_adaptor1 = l; } else { opt.adaptor2 = a; opt.l_adaptor2 = l; }
    } else if (c == 'z') opt.min_base_qual = atoi(optarg);
    else if (c == '5') opt.clip5 = atoi(optarg);
    else if (c == '3') opt.clip3 = atoi(optarg);
    else if (c == '9') opt.has_bc = 1;
    else if (c == 'X') opt.mask_level = (float)atof(optarg);
    else if (c == 'g') {
      opt.max_XA_hits = opt.max_XA_hits_alt = (int)strtol(optarg, &p, 10);
      if (*p != 0 && ispunct((unsigned char)*p) && isdigit((unsigned char)p[1])) opt.max_XA_hits_alt = (int)strtol(p + 1, &p, 10);
    } else if (c == 'Q') {
      opt.mapQ_coef_len = (float)atoi(optarg);
      opt.mapQ_coef_fac = opt.mapQ_coef_len > 0 ? (int)log(opt.mapQ_coef_len) : 0;
    } else if (c == 'O') {
      set.o_del = set.o_ins = 1;
      opt.o_del = opt.o_ins = (int)strtol(optarg, &p, 10);
      if (*p != 0 && ispunct((unsigned char)*p) && isdigit((unsigned char)p[1])) opt.o_ins = (int)strtol(p + 1, &p, 10);
    } else if (c == 'E') {
      set.e_del = set.e_ins = 1;
      opt.e_del = opt.e_ins = (int)strtol(optarg, &p, 10);
      if (*p != 0 && ispunct((unsigned char)*p) && isdigit((unsigned char)p[1])) opt.e_ins = (int)strtol(p + 1, &p, 10);
    } else if (c == 'L') {
      set.pen_clip5 = set.pen_clip3 = 1;
      opt.pen_clip5 = opt.pen_clip3 = (int)strtol(optarg, &p, 10);
      if (*p != 0 && ispunct((unsigned char)*p) && isdigit((unsigned char)p[1])) opt.pen_clip3 = (int)strtol(p + 1, &p, 10);
    } else if (c == 'R') { /* bwa_set_rg, bwa.c:702-729 */
      if (strstr(optarg, "@RG") != optarg) { fprintf(stderr, "[E::bwa_set_rg] the read group line is not started with @RG\n"); return 1; }
      rg_line = escape_hdr(strdup(optarg));
      char *id = strstr(rg_line, "\tID:");
      if (!id) { fprintf(stderr, "[E::bwa_set_rg] no ID at the read group line\n"); return 1; }
      id += 4;
      for (i = 0; id[i] && id[i] != '\t' && id[i] != '\n' && i < 255; ++i) rg_id[i] = id[i];
      rg_id[i] = 0;
    } else if (c == 'H') {
      if (optarg[0] != '@') {
        FILE *fp = fopen(optarg, "r");
        if (fp) {
          char *buf = calloc(1, 0x10000);
          while (fgets(buf, 0xffff, fp)) { size_t l = strlen(buf); if (l && buf[l - 1] == '\n') buf[l - 1] = 0; hdr_line = insert_header(buf, hdr_line); }
          free(buf);
          fclose(fp);
        }
      } else hdr_line = insert_header(optarg, hdr_line);
    } else if (c == 'I') {
      pes0 = calloc(1, sizeof *pes0);
      pes0->avg = strtod(optarg, &p);
      pes0->std = pes0->avg * .1;
      if (*p != 0 && ispunct((unsigned char)*p) && isdigit((unsigned char)p[1])) pes0->std = strtod(p + 1, &p);
      pes0->high = (int)(pes0->avg + 4. * pes0->std + .499);
      pes0->low = (int)(pes0->avg - 4. * pes0->std + .499);
      if (*p != 0 && ispunct((unsigned char)*p) && isdigit((unsigned char)p[1])) pes0->high = (int)(strtod(p + 1, &p) + .499);
      if (*p != 0 && ispunct((unsigned char)*p) && isdigit((unsigned char)p[1])) pes0->low = (int)(strtod(p + 1, &p) + .499);
      if (bq_verbose >= 3)
        fprintf(stderr, "[M::main_align] mean insert size: %.3f, stddev: %.3f, max: %d, min: %d\n", pes0->avg, pes0->std, pes0->high, pes0->low);
    } else if (c == 'p') { opt.flag |= BQ_F_PE; smart_pe = 1; }
    else if (c == ':') { align_usage(); bq_fatal("Option needs an argument: -%c", optopt); }
    else if (c == '?') { align_usage(); bq_fatal("Unrecognized option: -%c", optopt); }
    else return align_usage();
  }
  if (rg_line) { hdr_line = insert_header(rg_line, hdr_line); free(rg_line); }
  if ((optind + 1 >= argc || optind + 3 < argc) && !seq1) { align_usage(); bq_fatal("Missing fai-index base or FASTQ file"); }
  if (set.a) { /* update_a, align.c:169-182 */
    if (!set.b) opt.b *= opt.a;
    if (!set.T) opt.T *= opt.a;
    if (!set.o_del) opt.o_del *= opt.a;
    if (!set.e_del) opt.e_del *= opt.a;
    if (!set.o_ins) opt.o_ins *= opt.a;
    if (!set.e_ins) opt.e_ins *= opt.a;
    if (!set.zdrop) opt.zdrop *= opt.a;
    if (!set.pen_clip5) opt.pen_clip5 *= opt.a;
    if (!set.pen_clip3) opt.pen_clip3 *= opt.a;
    if (!set.pen_unpaired) opt.pen_unpaired *= opt.a;
  }
  bq_fill_scmat(opt.a, opt.b, opt.mat);
  bq_fill_scmat_bis(opt.a, opt.b, 1, opt.ctmat);
  bq_fill_scmat_bis(opt.a, opt.b, 0, opt.gamat);

  bq_index_t idx;
  double t0 = now();
  if (bq_index_load(argv[optind], &idx)) { fprintf(stderr, "[E::main_align] fail to locate the index files\n"); return 1; }
  if (auto_alt) infer_alt(&idx.ref);
  if (ignore_alt) for (i = 0; i < idx.ref.n_seqs; ++i) idx.ref.anns[i].is_alt = 0;
  /* one index replica and one aligner context per device: FASTQ batches are dealt to the devices round-robin by the
   * pipeline (bq_pipe.c), no data moves between GPUs */
  bsq_index *dxs[BQ_MAX_LANES] = {0};
  bsq_aligner *als[BQ_MAX_LANES] = {0};
  bsq_dp *dps[BQ_MAX_LANES] = {0}; /* batched phase-2 DP (final CIGARs, mate-rescue alignments), one context per lane */
  bsq_opt dopt;
  bq_opt_to_dev(&opt, &dopt);
  int rc = 0;
  for (i = 0; i < n_dev; ++i) {
    if ((rc = bq_index_to_device(&idx, devices[i], &dxs[i])))
      bq_fatal("cannot stage the index on CUDA device %d: %s (%s)", devices[i], bsq_strerror(rc), bsq_last_error());
    if ((rc = bsq_aligner_create(dxs[i], &dopt, &als[i]))) bq_fatal("bsq_aligner_create (device %d): %s", devices[i], bsq_strerror(rc));
    if ((rc = bsq_dp_create(dxs[i], &dopt, &dps[i]))) bq_fatal("bsq_dp_create (device %d): %s", devices[i], bsq_strerror(rc));
  }
  bsq_aligner *al = als[0];
  int n_al = n_dev;
  if (n_dev == 1 && getenv("BQ_TWO_CONTEXTS") && !bsq_aligner_create(dxs[0], &dopt, &als[1]) && !bsq_dp_create(dxs[0], &dopt, &dps[1])) n_al = 2; /* off by default: measured no gain */
  if (bq_verbose >= 3) fprintf(stderr, "[M::main_align] index loaded and staged on %d GPU(s) in %.3f sec\n", n_dev, now() - t0);

  bq_fastq_t *f1 = 0, *f2 = 0;
  if (!seq1) {
    if (!(f1 = bq_fastq_open(argv[optind + 1]))) { fprintf(stderr, "[E::main_align] fail to open file `%s'.\n", argv[optind + 1]); return 1; }
    if (optind + 2 < argc) {
      if (smart_pe) {
        if (bq_verbose >= 2) fprintf(stderr, "[W::main_align] when '-p' is in use, the second query file is ignored.\n");
      } else if (!(f2 = bq_fastq_open(argv[optind + 2]))) { fprintf(stderr, "[E::main_align] fail to open file `%s'.\n", argv[optind + 2]); return 1; }
      opt.flag |= BQ_F_PE;
    }
  }
  if (!no_hdr) bq_print_sam_hdr(&idx.ref, hdr_line, getenv("BISCUIT_PG_LINE"));
  if (getenv("BQ_CHUNK_SIZE")) opt.chunk_size = atoi(getenv("BQ_CHUNK_SIZE")); /* test hook: bases per thread and batch */
  const int chunk = opt.chunk_size * opt.n_threads;
  if (seq1) { /* -1/-2: literal reads (align.c:77-81) */
    int n = seq2 ? 2 : 1;
    bq_read_t *seqs = calloc((size_t)n, sizeof(bq_read_t));
    char *lit[2] = {seq1, seq2};
    for (i = 0; i < n; ++i) {
      seqs[i].name = strdup("inputread");
      seqs[i].l_seq = seqs[i].l_seq0 = (int)strlen(lit[i]);
      seqs[i].seq = seqs[i].seq0 = malloc((size_t)seqs[i].l_seq + 1);
      for (int k = 0; k < seqs[i].l_seq; ++k) seqs[i].seq[k] = t4[(unsigned char)lit[i][k]];
    }
    if (seq2) opt.flag |= BQ_F_PE;
    if ((rc = bq_process_seqs(&opt, al, dps[0], &idx.ref, 0, n, seqs, pes0, rg_id))) bq_fatal("alignment failed: %s (%s)", bsq_strerror(rc), bsq_last_error());
    for (i = 0; i < n; ++i) if (seqs[i].sam) fputs(seqs[i].sam, stdout);
    bq_reads_free(seqs, n);
  } else {
    src_ctx_t sc = {&opt, f1, f2, chunk, copy_comment};
    if (smart_pe) {
      if ((rc = align_smart_pairing(&opt, &idx.ref, al, dps[0], &sc, pes0, rg_id))) bq_fatal("alignment batch failed: %s (%s)", bsq_strerror(rc), bsq_last_error());
    } else {
      ahead_src_t ah;
      memset(&ah, 0, sizeof ah);
      ah.sc = &sc;
      pthread_mutex_init(&ah.mu, 0); pthread_cond_init(&ah.cv, 0);
      const int threaded = !getenv("BQ_FQ_NO_AHEAD") && pthread_create(&ah.th, 0, ahead_main, &ah) == 0;
      rc = threaded ? bq_pipeline_run(&opt, &idx.ref, als, dps, n_al, ahead_source, &ah, sam_sink, 0, pes0, rg_id)
                    : bq_pipeline_run(&opt, &idx.ref, als, dps, n_al, fastq_source, &sc, sam_sink, 0, pes0, rg_id);
      if (threaded) ahead_stop(&ah);
      if (rc) bq_fatal("alignment batch failed: %s (%s)", bsq_strerror(rc), bsq_last_error());
    }
  }
  if (getenv("BQ_TIMING") || bq_verbose >= 4) {
    int64_t st[6];
    bq_dp_stats(st);
    fprintf(stderr, "[M::main_align] phase-2 DP on the GPU: %lld CIGAR jobs, %lld used, %lld setSAM calls on the host; mate rescue: %lld jobs, %lld used, %lld on the host\n",
            (long long)st[0], (long long)st[1], (long long)st[2], (long long)st[3], (long long)st[4], (long long)st[5]);
  }
  for (i = 0; i < BQ_MAX_LANES; ++i) if (dps[i]) bsq_dp_destroy(dps[i]);
  for (i = 0; i < BQ_MAX_LANES; ++i) if (als[i]) bsq_aligner_destroy(als[i]);
  for (i = 0; i < BQ_MAX_LANES; ++i) if (dxs[i]) bsq_index_free(dxs[i]);
  bq_index_free(&idx);
  bq_fastq_close(f1); bq_fastq_close(f2);
  free(hdr_line); free(pes0); free(opt.adaptor1); free(opt.adaptor2);
  return 0;
}

int bq_main_pileup(int argc, char **argv);

int main(int argc, char **argv) {
  if (argc < 2) {
    fprintf(stderr, "\nProgram: biscuit (B200 build of the index/align/pileup hot paths)\nVersion: %s\n\nUsage: biscuit <index|align|sortbam|pileup|vcf2bed|mergecg|version> [options]\n\n", BQ_VERSION);
    return 1;
  }
  int ret;
  double t0 = now();
  if (strcmp(argv[1], "index") == 0) ret = bq_main_index(argc - 1, argv + 1);
  else if (strcmp(argv[1], "align") == 0) ret = bq_main_align(argc - 1, argv + 1);
  else if (strcmp(argv[1], "pileup") == 0) ret = bq_main_pileup(argc - 1, argv + 1);
  else if (strcmp(argv[1], "sortbam") == 0) ret = bq_main_sortbam(argc - 1, argv + 1);
  else if (strcmp(argv[1], "bamdump") == 0) ret = bq_main_bamdump(argc - 1, argv + 1);
  else if (strcmp(argv[1], "vcf2bed") == 0) ret = bq_main_vcf2bed(argc - 1, argv + 1);
  else if (strcmp(argv[1], "mergecg") == 0) ret = bq_main_mergecg(argc - 1, argv + 1);
  else if (strcmp(argv[1], "version") == 0) { fprintf(stderr, "BISCUIT Version: %s\n", BQ_VERSION); return 0; }
  else { fprintf(stderr, "Unrecognized subcommand: %s\n", argv[1]); return 1; }
  fflush(stdout);
  if (ret == 0) fprintf(stderr, "[main] Real time: %.3f sec\n", now() - t0);
  return ret;
}
