/* bq_io.c -- on-disk index (reference's layout), `biscuit index`, FASTA/FASTQ input, SAM header.
 *
 * Index files (reference lib/aln/bwt.c:402-454, bntseq.c:514-539,588-633, bwtindex.c:206-347):
 *   <p>.{par,dau}.bwt  u64 primary, u64 L2[1..4], u32 bwt[]     (occ checkpoints interleaved every 128 symbols)
 *   <p>.{par,dau}.sa   u64 primary, u64 L2[1..4], u64 32, u64 seq_len, u64 sa[1..]
 *   <p>.bis.pac        2-bit forward reference + trailer;  <p>.bis.ann / .bis.amb  text
 * `biscuit index` here packs the FASTA on the host exactly like the reference (N -> lrand48()&3 after srand48(11))
 * and hands the packed reference to bsq_index_build (GPU suffix sorting) instead of bwt_gen/is.c. */
#include <ctype.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include <zlib.h>
#include "bq.h"

/* ---------------- buffered reader + FASTA/FASTQ records (kseq.h semantics) ---------------- */

struct bq_fastq {
  gzFile fp;
  unsigned char *buf;
  int beg, end, eof, last_char;
  bq_str_t name, comment, seq, qual;
  int comment_ever; /* the comment buffer has been allocated at least once (see bq_main_index) */
  struct fqa *ahead; /* paired input: this (second) file is parsed by a helper thread that runs ahead of the batcher */
};
static void fqa_stop(struct fqa *a);

#define FQ_BUF 0x40000

static int fq_getc(bq_fastq_t *f) {
  if (f->eof && f->beg >= f->end) return -1;
  if (f->beg >= f->end) {
    f->beg = 0;
    f->end = gzread(f->fp, f->buf, FQ_BUF);
    if (f->end <= 0) { f->eof = 1; f->end = 0; return -1; }
  }
  return f->buf[f->beg++];
}

/* append bytes up to a delimiter; mode 0: isspace, 1: newline.  Returns the delimiter or -1 at EOF (with nothing read) */
static int fq_until(bq_fastq_t *f, int line, bq_str_t *s, int append, int *dret) {
  int got = 0;
  if (dret) *dret = 0;
  if (!append) s->l = 0;
  for (;;) {
    if (f->beg >= f->end) {
      if (f->eof) break;
      f->beg = 0;
      f->end = gzread(f->fp, f->buf, FQ_BUF);
      if (f->end <= 0) { f->eof = 1; f->end = 0; break; }
    }
    int i;
    if (line) { unsigned char *p = memchr(f->buf + f->beg, '\n', (size_t)(f->end - f->beg)); i = p ? (int)(p - f->buf) : f->end; }
    else for (i = f->beg; i < f->end; ++i) if (isspace(f->buf[i])) break;
    bq_kputsn(s, (char *)f->buf + f->beg, (size_t)(i - f->beg));
    got = 1;
    f->beg = i + 1;
    if (i < f->end) { if (dret) *dret = f->buf[i]; goto done; }
  }
  if (!got && f->eof) return -1;
done:
  if (s->s == 0) { bq_str_reserve(s, 1); s->s[0] = 0; }
  if (line && s->l > 1 && s->s[s->l - 1] == '\r') s->s[--s->l] = 0;
  return 0;
}

bq_fastq_t *bq_fastq_open(const char *fn) {
  gzFile fp = strcmp(fn, "-") == 0 ? gzdopen(STDIN_FILENO, "r") : gzopen(fn, "r");
  if (!fp) return 0;
  bq_fastq_t *f = calloc(1, sizeof *f);
  f->fp = fp;
  f->buf = malloc(FQ_BUF);
  return f;
}

void bq_fastq_close(bq_fastq_t *f) {
  if (!f) return;
  if (f->ahead) fqa_stop(f->ahead);
  gzclose(f->fp);
  free(f->buf); free(f->name.s); free(f->comment.s); free(f->seq.s); free(f->qual.s);
  free(f);
}

/* >= 0 sequence length, -1 EOF, -2 truncated quality */
static int fq_read(bq_fastq_t *f) {
  int c, d;
  if (f->last_char == 0) {
    while ((c = fq_getc(f)) >= 0 && c != '>' && c != '@') {}
    if (c < 0) return -1;
    f->last_char = c;
  }
  f->comment.l = f->seq.l = f->qual.l = 0;
  if (fq_until(f, 0, &f->name, 0, &d) < 0) return -1;
  if (d != '\n') { fq_until(f, 1, &f->comment, 0, 0); f->comment_ever = 1; }
  while ((c = fq_getc(f)) >= 0 && c != '>' && c != '+' && c != '@') {
    if (c == '\n') continue;
    bq_kputc(&f->seq, c);
    fq_until(f, 1, &f->seq, 1, 0);
  }
  if (c == '>' || c == '@') f->last_char = c;
  bq_str_reserve(&f->seq, 1);
  f->seq.s[f->seq.l] = 0;
  if (c != '+') return (int)f->seq.l;
  while ((c = fq_getc(f)) >= 0 && c != '\n') {}
  if (c == -1) return -2;
  while (fq_until(f, 1, &f->qual, 1, 0) >= 0 && f->qual.l < f->seq.l) {}
  f->last_char = 0;
  if (f->seq.l != f->qual.l) return -2;
  return (int)f->seq.l;
}

static const uint8_t *nt4_table(void) { /* nst_nt4_table, bntseq.c:49-66 */
  static uint8_t t[256];
  static int init = 0;
  if (!init) {
    memset(t, 4, 256);
    t['A'] = t['a'] = 0; t['C'] = t['c'] = 1; t['G'] = t['g'] = 2; t['T'] = t['t'] = 3; t['-'] = 5;
    init = 1;
  }
  return t;
}

static void trim_readno(bq_str_t *s) { /* bwa.c:58-63 */
  if (s->l > 2 && s->s[s->l - 2] == '/' && isdigit((unsigned char)s->s[s->l - 1])) { s->l -= 2; s->s[s->l] = 0; }
}

static void to_read(bq_fastq_t *f, bq_read_t *s, int has_bc, int keep_comment) { /* bis_kseq2bseq1, bwa.c:766-815 */
  const uint8_t *t = nt4_table();
  memset(s, 0, sizeof *s);
  s->name = strdup(f->name.s);
  s->comment = (keep_comment && f->comment.l) ? strdup(f->comment.s) : 0;
  if (has_bc) {
    char *tmp = strdup(f->name.s), *tok, *bc = 0, *umi = 0;
    tok = strtok(tmp, "_");
    if (!tok) fprintf(stderr, "[W::bis_kseq2bseq1] barcode and UMI extraction requested but could not include be extracted\n");
    bc = strtok(0, "_"); umi = strtok(0, "_");
    while ((tok = strtok(0, "_")) != 0) { bc = umi; umi = tok; }
    s->barcode = bc ? strdup(bc) : 0; s->umi = umi ? strdup(umi) : 0;
    free(tmp);
  }
  s->l_seq = s->l_seq0 = (int)f->seq.l;
  s->seq = s->seq0 = malloc(f->seq.l + 1);
  for (size_t i = 0; i < f->seq.l; ++i) s->seq[i] = t[(unsigned char)f->seq.s[i]];
  s->qual = f->qual.l ? strdup(f->qual.s) : 0;
}

/* slab variant of to_read for the common case (no barcode extraction, no comments): offsets first, pointers once
 * the slab has stopped growing */
static void to_read_slab(bq_fastq_t *f, bq_read_t *s, bq_str_t *slab, size_t off[3]) {
  const uint8_t *t = nt4_table();
  memset(s, 0, sizeof *s);
  bq_str_reserve(slab, f->name.l + 1 + 2 * (f->seq.l + 1) + 16);
  off[0] = slab->l; memcpy(slab->s + slab->l, f->name.s, f->name.l + 1); slab->l += f->name.l + 1;
  off[1] = slab->l;
  for (size_t i = 0; i < f->seq.l; ++i) slab->s[slab->l + i] = (char)t[(unsigned char)f->seq.s[i]];
  slab->l += f->seq.l + 1;
  off[2] = (size_t)-1;
  if (f->qual.l) { off[2] = slab->l; memcpy(slab->s + slab->l, f->qual.s, f->qual.l + 1); slab->l += f->qual.l + 1; }
  s->l_seq = s->l_seq0 = (int)f->seq.l;
  s->in_slab = 1;
}

/* The common record -- four lines, '@name[ comment]', bases, '+...', as many quality characters, all inside the read
 * buffer -- goes from the buffer into the slab in one pass.  Anything else (multi-line records, FASTA, CR/LF, a
 * record cut by the end of the buffer, the last record of a file without a newline) returns 0 with nothing
 * consumed and takes the general kseq path (fq_read + to_read_slab); both produce the same bytes. */
static int fq_fast_slab(bq_fastq_t *f, bq_read_t *s, bq_str_t *slab, size_t off[3]) {
  static int off_switch = -1;
  if (off_switch < 0) off_switch = getenv("BQ_FQ_SLOW") != 0; /* test hook: general path only */
  if (off_switch || f->last_char != 0 || f->beg >= f->end) return 0;
  const unsigned char *p = f->buf + f->beg, *e = f->buf + f->end;
  if (p[0] != '@') return 0;
  const unsigned char *nl1 = memchr(p, '\n', (size_t)(e - p));
  if (!nl1) return 0;
  const unsigned char *ne = p + 1;
  while (ne < nl1 && !isspace(*ne)) ++ne;
  const unsigned char *q = nl1 + 1;
  if (q >= e || *q == '+' || *q == '>' || *q == '@' || *q == '\n') return 0;
  const unsigned char *nl2 = memchr(q, '\n', (size_t)(e - q));
  if (!nl2 || nl2[-1] == '\r') return 0;
  const unsigned char *r = nl2 + 1;
  if (r >= e || *r != '+') return 0;
  const unsigned char *nl3 = memchr(r, '\n', (size_t)(e - r));
  if (!nl3) return 0;
  const unsigned char *u = nl3 + 1;
  const size_t l_seq = (size_t)(nl2 - q);
  if ((size_t)(e - u) <= l_seq || u[l_seq] != '\n' || memchr(u, '\n', l_seq) || u[l_seq - 1] == '\r') return 0;
  size_t l_name = (size_t)(ne - (p + 1));
  if (l_name > 2 && p[1 + l_name - 2] == '/' && isdigit(p[1 + l_name - 1])) l_name -= 2; /* trim_readno */
  const uint8_t *t = nt4_table();
  memset(s, 0, sizeof *s);
  bq_str_reserve(slab, l_name + 1 + 2 * (l_seq + 1) + 16);
  char *d = slab->s + slab->l;
  off[0] = slab->l; memcpy(d, p + 1, l_name); d[l_name] = 0; d += l_name + 1;
  off[1] = (size_t)(d - slab->s);
  for (size_t i = 0; i < l_seq; ++i) d[i] = (char)t[q[i]];
  d += l_seq + 1;
  off[2] = (size_t)(d - slab->s); memcpy(d, u, l_seq); d[l_seq] = 0; d += l_seq + 1;
  slab->l = (size_t)(d - slab->s);
  s->l_seq = s->l_seq0 = (int)l_seq;
  s->in_slab = 1;
  f->beg = (int)(u + l_seq + 1 - f->buf);
  return 1;
}

/* ---------------- recycled large buffers ---------------- */
#include <malloc.h>
#include <pthread.h>
#define BQ_BIG_SLOTS 48
#define BQ_BIG_MIN (256u << 10) /* smaller blocks go back to malloc */
static pthread_mutex_t g_big_mu = PTHREAD_MUTEX_INITIALIZER;
static struct { void *p; size_t cap; } g_big[BQ_BIG_SLOTS];

void *bq_big_alloc(size_t n, size_t *cap) {
  void *p = 0;
  pthread_mutex_lock(&g_big_mu);
  int best = -1;
  for (int i = 0; i < BQ_BIG_SLOTS; ++i)
    if (g_big[i].p && g_big[i].cap >= n && (best < 0 || g_big[i].cap < g_big[best].cap)) best = i;
  if (best >= 0 && g_big[best].cap <= 4 * n + (1u << 20)) { p = g_big[best].p; *cap = g_big[best].cap; g_big[best].p = 0; }
  pthread_mutex_unlock(&g_big_mu);
  if (!p) {
    size_t want = n + (n >> 3) + 64;
    p = malloc(want);
    if (!p) bq_fatal("out of memory (%zu bytes)", want);
    *cap = malloc_usable_size(p);
  }
  return p;
}

void bq_big_free(void *p) {
  if (!p) return;
  const size_t cap = malloc_usable_size(p);
  if (cap >= BQ_BIG_MIN) {
    pthread_mutex_lock(&g_big_mu);
    int slot = -1, smallest = -1;
    for (int i = 0; i < BQ_BIG_SLOTS; ++i) {
      if (!g_big[i].p) { slot = i; break; }
      if (smallest < 0 || g_big[i].cap < g_big[smallest].cap) smallest = i;
    }
    if (slot < 0 && g_big[smallest].cap < cap) { void *q = g_big[smallest].p; g_big[smallest].p = p; g_big[smallest].cap = cap; p = q; }
    else if (slot >= 0) { g_big[slot].p = p; g_big[slot].cap = cap; p = 0; }
    pthread_mutex_unlock(&g_big_mu);
  }
  free(p);
}

/* ---------------- the second file of a pair, parsed ahead ----------------
 * bis_bseq_read (bwa.c:817-850) takes one record of file 1, then one of file 2.  Parsing is the cost of the reader
 * (~0.45 us per record on one core); with two files a helper thread parses file 2 into blocks of records while the
 * batcher parses file 1 and copies the matching record out of the helper's block -- the same records in the same
 * order with the same batch boundaries, at about twice the rate. */
#define FQA_RECS 4096
#define FQA_DEPTH 4
typedef struct fqa_block { bq_str_t slab; size_t offs[3 * FQA_RECS]; int lens[FQA_RECS]; int n, eof; struct fqa_block *next; } fqa_block_t;
struct fqa {
  bq_fastq_t *f;
  pthread_t th;
  pthread_mutex_t mu;
  pthread_cond_t cv;
  fqa_block_t *head, *tail, *freel; /* filled blocks in order; recycled blocks */
  int n_q, stop;
  fqa_block_t *cur; int cur_i;      /* consumer side */
};
static void *fqa_main(void *arg) {
  struct fqa *a = arg;
  bq_fastq_t *f = a->f;
  for (;;) {
    pthread_mutex_lock(&a->mu);
    fqa_block_t *b = a->freel;
    if (b) a->freel = b->next;
    pthread_mutex_unlock(&a->mu);
    if (!b) b = calloc(1, sizeof *b);
    b->n = 0; b->eof = 0; b->slab.l = 0; b->next = 0;
    while (b->n < FQA_RECS) {
      bq_read_t tmp;
      size_t *off = b->offs + 3 * (size_t)b->n;
      if (!fq_fast_slab(f, &tmp, &b->slab, off)) {
        if (fq_read(f) < 0) { b->eof = 1; break; }
        trim_readno(&f->name);
        to_read_slab(f, &tmp, &b->slab, off);
      }
      b->lens[b->n++] = tmp.l_seq;
    }
    pthread_mutex_lock(&a->mu);
    while (a->n_q >= FQA_DEPTH && !a->stop) pthread_cond_wait(&a->cv, &a->mu);
    if (a->tail) a->tail->next = b; else a->head = b;
    a->tail = b; a->n_q++;
    const int end = b->eof || a->stop;
    pthread_cond_broadcast(&a->cv);
    pthread_mutex_unlock(&a->mu);
    if (end) return 0;
  }
}
static struct fqa *fqa_start(bq_fastq_t *f) {
  struct fqa *a = calloc(1, sizeof *a);
  a->f = f;
  pthread_mutex_init(&a->mu, 0); pthread_cond_init(&a->cv, 0);
  if (pthread_create(&a->th, 0, fqa_main, a) != 0) { free(a); return 0; }
  return a;
}
/* next record of the file: pointers into the helper's block (valid until the next call); 0 at the end of the file */
static int fqa_next(struct fqa *a, const char **name, const uint8_t **seq, const char **qual, int *l_seq) {
  while (!a->cur || a->cur_i >= a->cur->n) {
    if (a->cur && a->cur->eof) return 0;
    pthread_mutex_lock(&a->mu);
    if (a->cur) { a->cur->next = a->freel; a->freel = a->cur; a->cur = 0; }
    while (!a->head) pthread_cond_wait(&a->cv, &a->mu);
    a->cur = a->head; a->head = a->cur->next; if (!a->head) a->tail = 0;
    a->n_q--; a->cur_i = 0;
    pthread_cond_broadcast(&a->cv);
    pthread_mutex_unlock(&a->mu);
  }
  const fqa_block_t *b = a->cur;
  const size_t *off = b->offs + 3 * (size_t)a->cur_i;
  *name = b->slab.s + off[0]; *seq = (const uint8_t *)b->slab.s + off[1];
  *qual = off[2] == (size_t)-1 ? 0 : b->slab.s + off[2];
  *l_seq = b->lens[a->cur_i++];
  return 1;
}
static void fqa_stop(struct fqa *a) {
  pthread_mutex_lock(&a->mu);
  a->stop = 1;
  pthread_cond_broadcast(&a->cv);
  pthread_mutex_unlock(&a->mu);
  /* the helper leaves after its current block; blocks it may be waiting to queue are accepted because of `stop` */
  pthread_join(a->th, 0);
  fqa_block_t *lists[3] = {a->head, a->freel, a->cur};
  for (int k = 0; k < 3; ++k)
    for (fqa_block_t *b = lists[k]; b;) { fqa_block_t *nx = k == 2 ? 0 : b->next; free(b->slab.s); free(b); b = nx; }
  pthread_mutex_destroy(&a->mu); pthread_cond_destroy(&a->cv);
  free(a);
}

bq_read_t *bq_read_batch(int chunk_size, int has_bc, int keep_comment, int *n_, bq_fastq_t *f1, bq_fastq_t *f2) { /* bis_bseq_read, bwa.c:817-850 */
  int size = 0, m = 0, n = 0;
  bq_read_t *seqs = 0;
  const int use_slab = !has_bc && !keep_comment;
  bq_str_t slab = {0, 0, 0};
  if (use_slab) { slab.s = bq_big_alloc((size_t)chunk_size * 2 + ((size_t)chunk_size >> 2) + 4096, &slab.m); slab.s[0] = 0; }
  size_t *offs = 0;
  if (f2 && use_slab && !f2->ahead && !getenv("BQ_FQ_NO_AHEAD")) f2->ahead = fqa_start(f2); /* from here on file 2 belongs to the helper */
  struct fqa *ahead = f2 && use_slab ? f2->ahead : 0;
  for (;;) {
    if (n + 2 > m) {
      m = m ? m << 1 : 256;
      seqs = realloc(seqs, (size_t)m * sizeof(bq_read_t));
      if (use_slab) offs = realloc(offs, (size_t)m * 3 * sizeof(size_t));
    }
    /* the read of file 1, then (paired input) the read of file 2: fast path first, general path otherwise */
    const int fast1 = use_slab && fq_fast_slab(f1, &seqs[n], &slab, offs + 3 * (size_t)n);
    if (!fast1 && fq_read(f1) < 0) break;
    const char *a_name = 0, *a_qual = 0; const uint8_t *a_seq = 0; int a_len = 0;
    if (ahead && !fqa_next(ahead, &a_name, &a_seq, &a_qual, &a_len)) { fprintf(stderr, "[W::bis_bseq_read] the 2nd file has fewer sequences.\n"); break; }
    const int fast2 = !ahead && f2 && use_slab && fq_fast_slab(f2, &seqs[n + 1], &slab, offs + 3 * (size_t)(n + 1));
    if (!ahead && f2 && !fast2 && fq_read(f2) < 0) { fprintf(stderr, "[W::bis_bseq_read] the 2nd file has fewer sequences.\n"); break; }
    if (!fast1) {
      trim_readno(&f1->name);
      if (use_slab) to_read_slab(f1, &seqs[n], &slab, offs + 3 * (size_t)n); else to_read(f1, &seqs[n], has_bc, keep_comment);
    }
    seqs[n].id = n;
    size += seqs[n++].l_seq;
    if (ahead) { /* the record the helper parsed: copied into this batch's slab */
      const size_t ln = strlen(a_name);
      size_t *off = offs + 3 * (size_t)n;
      memset(&seqs[n], 0, sizeof seqs[n]);
      bq_str_reserve(&slab, ln + 1 + 2 * ((size_t)a_len + 1) + 16);
      char *d = slab.s + slab.l;
      off[0] = slab.l; memcpy(d, a_name, ln + 1); d += ln + 1;
      off[1] = (size_t)(d - slab.s); memcpy(d, a_seq, (size_t)a_len); d += a_len + 1;
      off[2] = (size_t)-1;
      if (a_qual) { off[2] = (size_t)(d - slab.s); memcpy(d, a_qual, (size_t)a_len + 1); d += a_len + 1; }
      slab.l = (size_t)(d - slab.s);
      seqs[n].l_seq = seqs[n].l_seq0 = a_len;
      seqs[n].in_slab = 1;
      seqs[n].id = n;
      size += seqs[n++].l_seq;
    } else if (f2) {
      if (!fast2) {
        trim_readno(&f2->name);
        if (use_slab) to_read_slab(f2, &seqs[n], &slab, offs + 3 * (size_t)n); else to_read(f2, &seqs[n], has_bc, keep_comment);
      }
      seqs[n].id = n;
      size += seqs[n++].l_seq;
    }
    if (size >= chunk_size && (n & 1) == 0) break;
  }
  if (size == 0 && f2) {
    const char *a_name, *a_qual; const uint8_t *a_seq; int a_len;
    if (ahead ? fqa_next(ahead, &a_name, &a_seq, &a_qual, &a_len) : fq_read(f2) >= 0) fprintf(stderr, "[W::bis_bseq_read] the 1st file has fewer sequences.\n");
  }
  if (use_slab && n > 0) {
    for (int i = 0; i < n; ++i) {
      seqs[i].name = slab.s + offs[3 * (size_t)i];
      seqs[i].seq = seqs[i].seq0 = (uint8_t *)slab.s + offs[3 * (size_t)i + 1];
      seqs[i].qual = offs[3 * (size_t)i + 2] == (size_t)-1 ? 0 : slab.s + offs[3 * (size_t)i + 2];
    }
    seqs[0].slab = slab.s;
  } else bq_big_free(slab.s);
  free(offs);
  *n_ = n;
  return seqs;
}

void bq_reads_free(bq_read_t *seqs, int n) {
  if (!seqs) return;
  char *slab = n > 0 ? seqs[0].slab : 0;
  char **sam_slabs = n > 0 ? seqs[0].sam_slabs : 0;
  const int n_sam_slabs = n > 0 ? seqs[0].n_sam_slabs : 0;
  for (int i = 0; i < n; ++i) {
    if (!seqs[i].in_slab) { free(seqs[i].name); free(seqs[i].seq0); free(seqs[i].qual); }
    free(seqs[i].comment); free(seqs[i].barcode); free(seqs[i].umi);
    if (!seqs[i].sam_in_slab) free(seqs[i].sam);
  }
  for (int k = 0; k < n_sam_slabs; ++k) bq_big_free(sam_slabs[k]);
  free(sam_slabs);
  bq_big_free(slab);
  free(seqs);
}

/* ---------------- index files ---------------- */

static void *read_file_part(FILE *fp, size_t bytes, const char *fn) {
  void *p = malloc(bytes + 64);
  if (fread(p, 1, bytes, fp) != bytes) bq_fatal("unexpected end of file in %s", fn);
  return p;
}

static int load_half(const char *prefix, const char *tag, bq_fm_t *fm) {
  char fn[4096];
  snprintf(fn, sizeof fn, "%s.%s.bwt", prefix, tag);
  FILE *fp = fopen(fn, "rb");
  if (!fp) { fprintf(stderr, "[E::bq_index_load] cannot open %s\n", fn); return -1; }
  fseek(fp, 0, SEEK_END);
  const long sz = ftell(fp);
  fseek(fp, 0, SEEK_SET);
  uint64_t hdr[5];
  if (fread(hdr, 8, 5, fp) != 5) bq_fatal("truncated %s", fn);
  fm->primary = hdr[0]; fm->L2[0] = 0;
  for (int i = 0; i < 4; ++i) fm->L2[i + 1] = hdr[i + 1];
  fm->seq_len = fm->L2[4];
  fm->bwt_words = (uint64_t)(sz - 40) >> 2;
  fm->bwt = 0;
  if (posix_memalign((void **)&fm->bwt, 64, fm->bwt_words * 4 + 64)) bq_fatal("out of memory");
  if (fread(fm->bwt, 4, fm->bwt_words, fp) != fm->bwt_words) bq_fatal("truncated %s", fn);
  fclose(fp);
  snprintf(fn, sizeof fn, "%s.%s.sa", prefix, tag);
  fp = fopen(fn, "rb");
  if (!fp) { fprintf(stderr, "[E::bq_index_load] cannot open %s\n", fn); return -1; }
  uint64_t h2[7];
  if (fread(h2, 8, 7, fp) != 7) bq_fatal("truncated %s", fn);
  if (h2[0] != fm->primary) bq_fatal("SA-BWT inconsistency: primary is not the same.");
  if (h2[6] != fm->seq_len) bq_fatal("SA-BWT inconsistency: seq_len is not the same.");
  fm->sa_intv = (int)h2[5];
  fm->n_sa = (fm->seq_len + (uint64_t)fm->sa_intv) / (uint64_t)fm->sa_intv;
  fm->sa = malloc(fm->n_sa * 8 + 8);
  fm->sa[0] = (uint64_t)-1;
  if (fread(fm->sa + 1, 8, fm->n_sa - 1, fp) != fm->n_sa - 1) bq_fatal("truncated %s", fn);
  fclose(fp);
  return 0;
}

int bq_index_load(const char *prefix, bq_index_t *idx) { /* bwa_idx_load_from_disk + bns_restore, bwa.c:525-554, bntseq.c:96-216 */
  char fn[4096], str[8192];
  memset(idx, 0, sizeof *idx);
  if (load_half(prefix, "dau", &idx->fm[0]) || load_half(prefix, "par", &idx->fm[1])) return -1;
  bq_ref_t *r = &idx->ref;
  snprintf(fn, sizeof fn, "%s.bis.ann", prefix);
  FILE *fp = fopen(fn, "r");
  if (!fp) { fprintf(stderr, "[E::bq_index_load] cannot open %s\n", fn); return -1; }
  long long xx;
  if (fscanf(fp, "%lld%d%u", &xx, &r->n_seqs, &r->seed) != 3) bq_fatal("Parse error reading %s", fn);
  r->l_pac = xx;
  r->anns = calloc((size_t)r->n_seqs, sizeof(bq_ann_t));
  for (int i = 0; i < r->n_seqs; ++i) {
    bq_ann_t *p = r->anns + i;
    char *q = str;
    int c;
    if (fscanf(fp, "%u%8191s", &p->gi, str) != 2) bq_fatal("Parse error reading %s", fn);
    p->name = strdup(str);
    while ((size_t)(q - str) < sizeof(str) - 1 && (c = fgetc(fp)) != '\n' && c != EOF) *q++ = (char)c;
    *q = 0;
    p->anno = (q - str > 1 && strcmp(str, " (null)") != 0) ? strdup(str + 1) : strdup("");
    if (fscanf(fp, "%lld%d%d", &xx, &p->len, &p->n_ambs) != 3) bq_fatal("Parse error reading %s", fn);
    p->offset = xx;
  }
  fclose(fp);
  snprintf(fn, sizeof fn, "%s.bis.amb", prefix);
  if ((fp = fopen(fn, "r")) != 0) {
    int ns;
    if (fscanf(fp, "%lld%d%d", &xx, &ns, &r->n_holes) != 3) bq_fatal("Parse error reading %s", fn);
    r->ambs = r->n_holes ? calloc((size_t)r->n_holes, sizeof(bq_amb_t)) : 0;
    for (int i = 0; i < r->n_holes; ++i) {
      if (fscanf(fp, "%lld%d%8191s", &xx, &r->ambs[i].len, str) != 3) bq_fatal("Parse error reading %s", fn);
      r->ambs[i].offset = xx; r->ambs[i].amb = str[0];
    }
    fclose(fp);
  }
  snprintf(fn, sizeof fn, "%s.bis.pac", prefix);
  fp = fopen(fn, "rb");
  if (!fp) { fprintf(stderr, "[E::bq_index_load] cannot open %s\n", fn); return -1; }
  r->pac = read_file_part(fp, (size_t)(r->l_pac / 4 + 1), fn);
  fclose(fp);
  snprintf(fn, sizeof fn, "%s.alt", prefix);
  if ((fp = fopen(fn, "r")) != 0) { /* ALT contigs, bntseq.c:189-214 */
    while (fgets(str, sizeof str, fp)) {
      if (str[0] == '@') continue;
      str[strcspn(str, "\t\r\n")] = 0;
      for (int i = 0; i < r->n_seqs; ++i) if (strcmp(r->anns[i].name, str) == 0) r->anns[i].is_alt = 1;
    }
    fclose(fp);
  }
  return 0;
}

void bq_index_free(bq_index_t *idx) {
  for (int w = 0; w < 2; ++w) { free(idx->fm[w].bwt); free(idx->fm[w].sa); }
  for (int i = 0; i < idx->ref.n_seqs; ++i) { free(idx->ref.anns[i].name); free(idx->ref.anns[i].anno); }
  free(idx->ref.anns); free(idx->ref.ambs); free(idx->ref.pac);
  memset(idx, 0, sizeof *idx);
}

int bq_index_to_device(const bq_index_t *idx, int device, bsq_index **out) {
  bsq_index_desc d;
  memset(&d, 0, sizeof d);
  const int n = idx->ref.n_seqs;
  int64_t *off = malloc(sizeof(int64_t) * (size_t)n);
  int32_t *len = malloc(sizeof(int32_t) * (size_t)n), *alt = malloc(sizeof(int32_t) * (size_t)n);
  for (int i = 0; i < n; ++i) { off[i] = idx->ref.anns[i].offset; len[i] = idx->ref.anns[i].len; alt[i] = idx->ref.anns[i].is_alt; }
  for (int w = 0; w < 2; ++w) {
    d.bwt[w] = idx->fm[w].bwt; d.bwt_words[w] = idx->fm[w].bwt_words; d.primary[w] = idx->fm[w].primary;
    for (int i = 0; i < 5; ++i) d.L2[w][i] = idx->fm[w].L2[i];
    d.sa[w] = idx->fm[w].sa; d.n_sa[w] = idx->fm[w].n_sa; d.sa_intv[w] = idx->fm[w].sa_intv;
  }
  d.seq_len = idx->fm[0].seq_len; d.pac = idx->ref.pac; d.l_pac = idx->ref.l_pac; d.n_seqs = n;
  d.ann_offset = off; d.ann_len = len; d.ann_is_alt = alt;
  int rc = bsq_index_upload(&d, device, out);
  free(off); free(len); free(alt);
  return rc;
}

/* ---------------- biscuit index ---------------- */

static void write_or_die(const void *p, size_t sz, size_t n, FILE *fp, const char *fn) {
  if (fwrite(p, sz, n, fp) != n) bq_fatal("failed to write %s", fn);
}

int bq_main_index(int argc, char **argv) {
  char *prefix = 0;
  int c, device = 0;
  while ((c = getopt(argc, argv, ":6a:p:hg:")) >= 0) {
    if (c == 'p') prefix = strdup(optarg);
    else if (c == 'a') { /* -a is|bwtsw|div: accepted for compatibility; the GPU builder replaces all three */ }
    else if (c == '6') {}
    else if (c == 'g') device = atoi(optarg);
    else {
      fprintf(stderr, "\nUsage: biscuit index [options] <in.fasta>\n\nOptions:\n    -a STR    BWT construction algorithm (ignored: built on the GPU)\n"
                      "    -p STR    Prefix of the index [same as fasta name]\n    -g INT    CUDA device [0]\n    -h        This help\n\n");
      return 1;
    }
  }
  if (optind + 1 > argc) bq_fatal("Missing FASTA reference");
  const char *fa = argv[optind];
  if (!prefix) prefix = strdup(fa);
  bq_fastq_t *f = bq_fastq_open(fa);
  if (!f) bq_fatal("fail to open %s", fa);
  const uint8_t *t4 = nt4_table();
  /* pack: bis_add1 / add1 (bntseq.c:236-282,459-505) */
  bq_ref_t r;
  memset(&r, 0, sizeof r);
  r.seed = 11;
  srand48(r.seed);
  int m_seqs = 8, m_holes = 8;
  int64_t m_pac = 0x10000;
  r.anns = calloc((size_t)m_seqs, sizeof(bq_ann_t));
  r.ambs = calloc((size_t)m_holes, sizeof(bq_amb_t));
  uint8_t *pac = calloc((size_t)m_pac / 4, 1);
  bq_amb_t *q = r.ambs;
  fprintf(stderr, "[main_biscuit_index] Pack bisulfite FASTA...\n");
  while (fq_read(f) >= 0) {
    if (r.n_seqs == m_seqs) { m_seqs <<= 1; r.anns = realloc(r.anns, (size_t)m_seqs * sizeof(bq_ann_t)); }
    bq_ann_t *p = r.anns + r.n_seqs;
    p->name = strdup(f->name.s);
    /* quirk kept from bis_add1 (bntseq.c:469): the comment buffer is tested for existence, not for length, so a
     * header without a comment inherits the text of the previous header's comment */
    p->anno = f->comment_ever ? strdup(f->comment.s ? f->comment.s : "") : strdup("(null)");
    p->gi = 0; p->len = (int32_t)f->seq.l; p->is_alt = 0;
    p->offset = r.n_seqs == 0 ? 0 : (p - 1)->offset + (p - 1)->len;
    p->n_ambs = 0;
    int lasts = 0;
    for (size_t i = 0; i < f->seq.l; ++i) {
      int ch = (unsigned char)f->seq.s[i], code = t4[ch];
      if (code >= 4) {
        if (lasts == ch) ++q->len;
        else {
          if (r.n_holes == m_holes) { m_holes <<= 1; r.ambs = realloc(r.ambs, (size_t)m_holes * sizeof(bq_amb_t)); }
          q = r.ambs + r.n_holes;
          q->len = 1; q->offset = p->offset + (int64_t)i; q->amb = (char)ch;
          ++p->n_ambs; ++r.n_holes;
        }
      }
      lasts = ch;
      if (code >= 4) code = (int)(lrand48() & 3);
      if (r.l_pac == m_pac) {
        m_pac <<= 1;
        pac = realloc(pac, (size_t)m_pac / 4);
        memset(pac + r.l_pac / 4, 0, (size_t)(m_pac - r.l_pac) / 4);
      }
      pac[r.l_pac >> 2] |= (uint8_t)(code << ((~r.l_pac & 3) << 1));
      ++r.l_pac;
    }
    ++r.n_seqs;
  }
  bq_fastq_close(f);
  if (r.l_pac == 0) bq_fatal("no sequence in %s", fa);
  char fn[4096];
  FILE *fp;
  snprintf(fn, sizeof fn, "%s.bis.ann", prefix);
  if (!(fp = fopen(fn, "w"))) bq_fatal("cannot write %s", fn);
  fprintf(fp, "%lld %d %u\n", (long long)r.l_pac, r.n_seqs, r.seed);
  for (int i = 0; i < r.n_seqs; ++i) {
    bq_ann_t *p = r.anns + i;
    fprintf(fp, "%d %s", p->gi, p->name);
    if (p->anno[0]) fprintf(fp, " %s\n", p->anno); else fprintf(fp, "\n");
    fprintf(fp, "%lld %d %d\n", (long long)p->offset, p->len, p->n_ambs);
  }
  fclose(fp);
  snprintf(fn, sizeof fn, "%s.bis.amb", prefix);
  if (!(fp = fopen(fn, "w"))) bq_fatal("cannot write %s", fn);
  fprintf(fp, "%lld %d %u\n", (long long)r.l_pac, r.n_seqs, (unsigned)r.n_holes);
  for (int i = 0; i < r.n_holes; ++i) fprintf(fp, "%lld %d %c\n", (long long)r.ambs[i].offset, r.ambs[i].len, r.ambs[i].amb);
  fclose(fp);
  snprintf(fn, sizeof fn, "%s.bis.pac", prefix);
  if (!(fp = fopen(fn, "wb"))) bq_fatal("cannot write %s", fn);
  write_or_die(pac, 1, (size_t)((r.l_pac >> 2) + ((r.l_pac & 3) == 0 ? 0 : 1)), fp, fn);
  { uint8_t ct = 0; if (r.l_pac % 4 == 0) write_or_die(&ct, 1, 1, fp, fn); ct = (uint8_t)(r.l_pac % 4); write_or_die(&ct, 1, 1, fp, fn); }
  fclose(fp);
  /* both FM-indices on the GPU */
  fprintf(stderr, "[main_biscuit_index] Construct BWT, Occ and SA for the parent and daughter strands on the GPU...\n");
  int64_t *off = malloc(sizeof(int64_t) * (size_t)r.n_seqs);
  int32_t *len = malloc(sizeof(int32_t) * (size_t)r.n_seqs);
  for (int i = 0; i < r.n_seqs; ++i) { off[i] = r.anns[i].offset; len[i] = r.anns[i].len; }
  bsq_index *dx = 0;
  int rc = bsq_index_build(pac, r.l_pac, r.n_seqs, off, len, 0, device, &dx);
  if (rc) bq_fatal("bsq_index_build: %s (%s)", bsq_strerror(rc), bsq_last_error());
  uint64_t words[2], n_sa[2], primary[2], L2[10];
  int64_t stats[3];
  bsq_index_sizes(dx, words, n_sa, primary, L2, stats);
  const char *tags[2] = {"dau", "par"};
  for (int w = 0; w < 2; ++w) {
    uint32_t *bwt = malloc(words[w] * 4 + 64);
    uint64_t *sa = malloc(n_sa[w] * 8 + 64);
    if ((rc = bsq_index_download(dx, w, bwt, sa))) bq_fatal("bsq_index_download: %s", bsq_strerror(rc));
    uint64_t hdr[5] = {primary[w], L2[5 * w + 1], L2[5 * w + 2], L2[5 * w + 3], L2[5 * w + 4]};
    snprintf(fn, sizeof fn, "%s.%s.bwt", prefix, tags[w]);
    if (!(fp = fopen(fn, "wb"))) bq_fatal("cannot write %s", fn);
    write_or_die(hdr, 8, 5, fp, fn); write_or_die(bwt, 4, words[w], fp, fn);
    fclose(fp);
    snprintf(fn, sizeof fn, "%s.%s.sa", prefix, tags[w]);
    if (!(fp = fopen(fn, "wb"))) bq_fatal("cannot write %s", fn);
    uint64_t h2[2] = {32, (uint64_t)r.l_pac * 2};
    write_or_die(hdr, 8, 5, fp, fn); write_or_die(h2, 8, 2, fp, fn); write_or_die(sa + 1, 8, n_sa[w] - 1, fp, fn);
    fclose(fp);
    free(bwt); free(sa);
  }
  bsq_index_free(dx);
  fprintf(stderr, "[main_biscuit_index] done: %lld bp, %d sequences, %lld sort passes\n", (long long)r.l_pac, r.n_seqs, (long long)stats[0]);
  free(off); free(len); free(pac); free(prefix);
  return 0;
}

/* ---------------- SAM header (bwa_print_sam_hdr, bwa.c:654-684) ---------------- */

static int lt_name(const void *a, const void *b) { return strcmp((*(bq_ann_t *const *)a)->name, (*(bq_ann_t *const *)b)->name) < 0; }

void bq_print_sam_hdr(const bq_ref_t *ref, const char *hdr_line, const char *pg_line) {
  int n_SQ = 0;
  if (hdr_line) {
    const char *p = hdr_line;
    while ((p = strstr(p, "@SQ\t")) != 0) { if (p == hdr_line || *(p - 1) == '\n') ++n_SQ; p += 4; }
  }
  if (n_SQ == 0) { /* @SQ lines sorted by name (bwa.c:668-674) */
    bq_ann_t **ap = malloc(sizeof(bq_ann_t *) * (size_t)ref->n_seqs);
    for (int i = 0; i < ref->n_seqs; ++i) ap[i] = ref->anns + i;
    bq_introsort(ap, (size_t)ref->n_seqs, sizeof(bq_ann_t *), lt_name);
    for (int i = 0; i < ref->n_seqs; ++i) printf("@SQ\tSN:%s\tLN:%d\n", ap[i]->name, ap[i]->len);
    free(ap);
  } else if (n_SQ != ref->n_seqs && bq_verbose >= 2)
    fprintf(stderr, "[W::bwa_print_sam_hdr] %d @SQ lines provided with -H; %d sequences in the index. Continue anyway.\n", n_SQ, ref->n_seqs);
  if (hdr_line) printf("%s\n", hdr_line);
  if (pg_line) printf("%s\n", pg_line);
}
