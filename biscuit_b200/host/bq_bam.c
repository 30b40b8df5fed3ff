/* bq_bam.c -- BGZF / BAM / FASTA input for `biscuit pileup` (the reference uses htslib 1.18 for this:
 * hts_open/sam_hdr_read/sam_itr_next/bam_aux_get/faidx, src/pileup.c:651-704, src/refcache.h:82-113; htslib
 * is not vendored, so the formats are read from their specifications: SAM/BAM v1 section 4, BGZF section 4.1).
 *
 * The reader streams a coordinate-sorted BAM front to back; BGZF blocks are inflated a batch at a time on
 * several threads.  Records are handed out as raw pointers into the inflated stream (no per-record copy);
 * bq_plp_reads_push appends the fields pileup needs to a structure-of-arrays batch for bsq_plp_stage.
 */
#include <ctype.h>
#include <limits.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <zlib.h>
#include "bq_plp.h"

/* ------------------------------------------------------------------ BGZF ---- */

#define BGZF_BATCH 512 /* blocks inflated per refill (<= 32 MiB of output) */

struct bq_bgzf {
  FILE *fp;
  char *fn;
  int n_threads, eof;
  /* raw compressed batch */
  uint8_t *craw;
  size_t craw_cap;
  /* per-block descriptors of the current batch */
  struct { size_t coff; uint32_t clen, ulen; size_t uoff; } blk[BGZF_BATCH];
  int n_blk;
  /* inflated stream window: [ubeg, uend) valid, data may be moved to the front on refill */
  uint8_t *u;
  size_t ucap, ubeg, uend;
};

static uint32_t le32(const uint8_t *p) { return (uint32_t)p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16 | (uint32_t)p[3] << 24; }
static uint16_t le16(const uint8_t *p) { return (uint16_t)(p[0] | p[1] << 8); }

bq_bgzf_t *bq_bgzf_open(const char *fn, int n_threads) {
  FILE *fp = fopen(fn, "rb");
  if (!fp) return 0;
  bq_bgzf_t *b = calloc(1, sizeof *b);
  b->fp = fp; b->fn = strdup(fn); b->n_threads = n_threads < 1 ? 1 : n_threads;
  b->craw_cap = (size_t)BGZF_BATCH * 65536; b->craw = malloc(b->craw_cap);
  b->ucap = (size_t)BGZF_BATCH * 65536 * 2; b->u = malloc(b->ucap);
  setvbuf(fp, 0, _IOFBF, 1 << 22);
  return b;
}

void bq_bgzf_close(bq_bgzf_t *b) {
  if (!b) return;
  fclose(b->fp); free(b->fn); free(b->craw); free(b->u); free(b);
}

typedef struct { bq_bgzf_t *b; int t; int err; } inflate_job_t;

static void *inflate_worker(void *arg) {
  inflate_job_t *j = arg;
  bq_bgzf_t *b = j->b;
  z_stream zs;
  memset(&zs, 0, sizeof zs);
  if (inflateInit2(&zs, -15) != Z_OK) { j->err = 1; return 0; }
  for (int i = j->t; i < b->n_blk; i += b->n_threads) {
    const uint8_t *c = b->craw + b->blk[i].coff;
    const int xlen = le16(c + 10);
    zs.next_in = (Bytef *)(c + 12 + xlen); zs.avail_in = b->blk[i].clen - 12 - xlen - 8;
    zs.next_out = b->u + b->blk[i].uoff; zs.avail_out = b->blk[i].ulen;
    if (b->blk[i].ulen) {
      int rc = inflate(&zs, Z_FINISH);
      if (rc != Z_STREAM_END || zs.avail_out != 0) { j->err = 1; break; }
      if ((uint32_t)crc32(crc32(0, 0, 0), b->u + b->blk[i].uoff, b->blk[i].ulen) != le32(c + b->blk[i].clen - 8)) { j->err = 1; break; }
    }
    inflateReset(&zs);
  }
  inflateEnd(&zs);
  return 0;
}

double bq_bgzf_t_read, bq_bgzf_t_inflate, bq_plp_t_fill; /* diagnostics (BSQ_PLP_TIMING) */
static double bam_now(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + ts.tv_nsec * 1e-9; }

/* read and inflate the next batch of blocks behind the unread bytes; returns bytes added (0 at EOF) */
static size_t bgzf_refill(bq_bgzf_t *b) {
  if (b->eof) return 0;
  const double t0_ = bam_now();
  if (b->ubeg > 0) { /* keep the unread tail at the front */
    memmove(b->u, b->u + b->ubeg, b->uend - b->ubeg);
    b->uend -= b->ubeg; b->ubeg = 0;
  }
  size_t coff = 0, uoff = b->uend;
  b->n_blk = 0;
  while (b->n_blk < BGZF_BATCH) {
    uint8_t hdr[18];
    size_t got = fread(hdr, 1, 18, b->fp);
    if (got == 0) { b->eof = 1; break; }
    if (got < 18 || hdr[0] != 0x1f || hdr[1] != 0x8b || hdr[2] != 8 || !(hdr[3] & 4)) bq_fatal("[bgzf] %s: not a BGZF block\n", b->fn);
    const int xlen = le16(hdr + 10);
    /* find the BC subfield (normally first) */
    uint8_t extra[65536];
    memcpy(extra, hdr + 12, 6);
    if (xlen > 6 && fread(extra + 6, 1, (size_t)xlen - 6, b->fp) != (size_t)xlen - 6) bq_fatal("[bgzf] %s: truncated block\n", b->fn);
    int bsize = -1;
    for (int o = 0; o + 4 <= xlen;) {
      const int slen = le16(extra + o + 2);
      if (extra[o] == 'B' && extra[o + 1] == 'C' && slen == 2) { bsize = le16(extra + o + 4); break; }
      o += 4 + slen;
    }
    if (bsize < 0) bq_fatal("[bgzf] %s: block without BC field\n", b->fn);
    const uint32_t clen = (uint32_t)bsize + 1;
    if (coff + clen > b->craw_cap) bq_fatal("[bgzf] internal: batch overflow\n");
    uint8_t *c = b->craw + coff;
    memcpy(c, hdr, 12); memcpy(c + 12, extra, (size_t)xlen);
    const size_t rest = clen - 12 - (size_t)xlen;
    if (fread(c + 12 + xlen, 1, rest, b->fp) != rest) bq_fatal("[bgzf] %s: truncated block\n", b->fn);
    const uint32_t ulen = le32(c + clen - 4);
    if (ulen > 65536) bq_fatal("[bgzf] %s: bad ISIZE\n", b->fn);
    if (uoff + ulen > b->ucap) { b->ucap = (uoff + ulen) * 2; b->u = realloc(b->u, b->ucap); }
    b->blk[b->n_blk].coff = coff; b->blk[b->n_blk].clen = clen; b->blk[b->n_blk].ulen = ulen; b->blk[b->n_blk].uoff = uoff;
    b->n_blk++;
    coff += clen; uoff += ulen;
  }
  if (b->n_blk == 0) return 0;
  const double t1_ = bam_now();
  bq_bgzf_t_read += t1_ - t0_;
  int nt = b->n_threads < b->n_blk ? b->n_threads : b->n_blk;
  inflate_job_t jobs[64];
  pthread_t th[64];
  if (nt > 64) nt = 64;
  const int save = b->n_threads;
  b->n_threads = nt;
  for (int t = 0; t < nt; ++t) { jobs[t].b = b; jobs[t].t = t; jobs[t].err = 0; }
  for (int t = 1; t < nt; ++t) pthread_create(&th[t], 0, inflate_worker, &jobs[t]);
  inflate_worker(&jobs[0]);
  for (int t = 1; t < nt; ++t) pthread_join(th[t], 0);
  b->n_threads = save;
  for (int t = 0; t < nt; ++t) if (jobs[t].err) bq_fatal("[bgzf] %s: inflate/CRC error\n", b->fn);
  const size_t added = uoff - b->uend;
  b->uend = uoff;
  bq_bgzf_t_inflate += bam_now() - t1_;
  return added;
}

/* make at least n unread bytes available; returns pointer or NULL at EOF (fatal on a partial item) */
static const uint8_t *bgzf_need(bq_bgzf_t *b, size_t n) {
  while (b->uend - b->ubeg < n) {
    if (bgzf_refill(b) == 0 && b->eof) {
      if (b->uend - b->ubeg == 0) return 0;
      if (b->uend - b->ubeg < n) bq_fatal("[bam] %s: truncated file\n", b->fn);
    }
  }
  return b->u + b->ubeg;
}

/* ------------------------------------------------------------------ BAM ---- */

int bq_bam_read_header(bq_bgzf_t *b, bq_bam_hdr_t *h) {
  memset(h, 0, sizeof *h);
  const uint8_t *p = bgzf_need(b, 8);
  if (!p || memcmp(p, "BAM\1", 4) != 0) return -1;
  const uint32_t l_text = le32(p + 4);
  p = bgzf_need(b, 8 + (size_t)l_text + 4);
  if (!p) return -1;
  h->text = malloc((size_t)l_text + 1);
  memcpy(h->text, p + 8, l_text); h->text[l_text] = 0;
  h->n_targets = (int32_t)le32(p + 8 + l_text);
  b->ubeg += 8 + (size_t)l_text + 4;
  h->name = calloc((size_t)h->n_targets + 1, sizeof(char *));
  h->len = calloc((size_t)h->n_targets + 1, sizeof(int32_t));
  for (int i = 0; i < h->n_targets; ++i) {
    p = bgzf_need(b, 4);
    if (!p) return -1;
    const uint32_t l_name = le32(p);
    p = bgzf_need(b, 4 + (size_t)l_name + 4);
    if (!p) return -1;
    h->name[i] = malloc((size_t)l_name + 1);
    memcpy(h->name[i], p + 4, l_name); h->name[i][l_name] = 0;
    h->len[i] = (int32_t)le32(p + 4 + l_name);
    b->ubeg += 4 + (size_t)l_name + 4;
  }
  return 0;
}

void bq_bam_hdr_free(bq_bam_hdr_t *h) {
  for (int i = 0; i < h->n_targets; ++i) free(h->name[i]);
  free(h->name); free(h->len); free(h->text);
  memset(h, 0, sizeof *h);
}

/* next record: pointer to the block after block_size (valid until the next call), length in *len */
const uint8_t *bq_bam_next(bq_bgzf_t *b, uint32_t *len) {
  const uint8_t *p = bgzf_need(b, 4);
  if (!p) return 0;
  const uint32_t bs = le32(p);
  if (bs < 32) bq_fatal("[bam] %s: bad record size %u\n", b->fn, bs);
  p = bgzf_need(b, 4 + (size_t)bs);
  if (!p) return 0;
  b->ubeg += 4 + (size_t)bs;
  *len = bs;
  return p + 4;
}

/* leave the record in the stream */
const uint8_t *bq_bam_peek(bq_bgzf_t *b, uint32_t *len) {
  const uint8_t *p = bgzf_need(b, 4);
  if (!p) return 0;
  const uint32_t bs = le32(p);
  if (bs < 32) bq_fatal("[bam] %s: bad record size %u\n", b->fn, bs);
  p = bgzf_need(b, 4 + (size_t)bs);
  if (!p) return 0;
  *len = bs;
  return p + 4;
}

void bq_bam_skip(bq_bgzf_t *b, uint32_t len) { b->ubeg += 4 + (size_t)len; }

/* auxiliary field lookup: returns pointer to the type byte of tag, or NULL (bam_aux_get) */
static const uint8_t *aux_skip(const uint8_t *s, const uint8_t *end) {
  if (s >= end) return 0;
  const int t = *s++;
  switch (t) {
    case 'A': case 'c': case 'C': return s + 1 <= end ? s + 1 : 0;
    case 's': case 'S': return s + 2 <= end ? s + 2 : 0;
    case 'i': case 'I': case 'f': return s + 4 <= end ? s + 4 : 0;
    case 'd': return s + 8 <= end ? s + 8 : 0;
    case 'Z': case 'H':
      while (s < end && *s) ++s;
      return s < end ? s + 1 : 0;
    case 'B': {
      if (s + 5 > end) return 0;
      const int st = *s;
      const uint32_t n = le32(s + 1);
      const int sz = (st == 'c' || st == 'C') ? 1 : (st == 's' || st == 'S') ? 2 : (st == 'i' || st == 'I' || st == 'f') ? 4 : 0;
      if (!sz) return 0;
      s += 5 + (size_t)n * sz;
      return s <= end ? s : 0;
    }
  }
  return 0;
}

static const uint8_t *aux_get(const uint8_t *aux, const uint8_t *end, const char tag[2]) {
  const uint8_t *s = aux;
  while (s && s + 3 <= end) {
    if (s[0] == (uint8_t)tag[0] && s[1] == (uint8_t)tag[1]) return s + 2;
    s = aux_skip(s + 2, end);
  }
  return 0;
}

/* bam_aux2i: integer value of an integer-typed field, 0 for any other type */
static int64_t aux2i(const uint8_t *s) {
  switch (*s) {
    case 'c': return (int8_t)s[1];
    case 'C': return s[1];
    case 's': return (int16_t)le16(s + 1);
    case 'S': return le16(s + 1);
    case 'i': return (int32_t)le32(s + 1);
    case 'I': return le32(s + 1);
  }
  return 0;
}

/* get_mate_length (src/bisc_utils.c:124-161): reference length of the mate from its MC CIGAR string */
static int32_t mc_rlen(const char *mc) {
  if (*mc == '*') return 0;
  int64_t len = 0;
  int n_op = 0;
  const char *q = mc;
  while (*q) {
    char *e;
    const long v = strtol(q, &e, 10);
    const int c = (unsigned char)*e;
    if (c == 0) bq_fatal("No CIGAR operations found in MC tag\n");
    if (!strchr("MIDNSHP=XB", c)) bq_fatal("Unrecognized CIGAR operator\n");
    if (c == 'M' || c == 'D' || c == 'N' || c == '=' || c == 'X') len += v;
    ++n_op;
    q = e + 1;
  }
  if (n_op == 0) bq_fatal("No CIGAR operations found in MC tag\n");
  if (n_op >= 65536) bq_fatal("Too many CIGAR operations found in MC tag\n");
  return (int32_t)len;
}

/* ------------------------------------------------------------------ SoA batch ---- */

/* The batch arrays are page-locked (bsq_host_alloc) so that bsq_plp_stage copies them at full PCIe speed; they only
 * ever grow and batches are reused, so the (slow) pinned allocations amortise. */
static int g_pinned = 1; /* bamdump (no device involved) switches to plain memory */
/* Page-locking is a heavy call whatever the size (the three batches of a run hold 18 arrays each), so the arrays are
 * carved out of a few large page-locked slabs: a bump allocator whose blocks are never returned one by one (the arrays
 * only grow, geometrically, so what a regrown array leaves behind is bounded by its final size); the slabs live until
 * the process ends. */
#define PIN_SLAB_MIN ((size_t)96 << 20)
static struct { char *base; size_t cap, used; } g_pin_slab[64];
static int g_n_pin_slab;
static pthread_mutex_t g_pin_mu = PTHREAD_MUTEX_INITIALIZER;
static void *pinned_carve(size_t bytes) {
  bytes = (bytes + 255) & ~(size_t)255;
  pthread_mutex_lock(&g_pin_mu);
  int k = g_n_pin_slab - 1;
  if (k < 0 || g_pin_slab[k].used + bytes > g_pin_slab[k].cap) {
    if (g_n_pin_slab == 64) { pthread_mutex_unlock(&g_pin_mu); return 0; }
    size_t cap = bytes > PIN_SLAB_MIN ? bytes : PIN_SLAB_MIN;
    void *p = 0;
    if (bsq_host_alloc(&p, cap) != 0) { pthread_mutex_unlock(&g_pin_mu); return 0; }
    k = g_n_pin_slab++;
    g_pin_slab[k].base = p; g_pin_slab[k].cap = cap; g_pin_slab[k].used = 0;
  }
  void *r = g_pin_slab[k].base + g_pin_slab[k].used;
  g_pin_slab[k].used += bytes;
  pthread_mutex_unlock(&g_pin_mu);
  return r;
}
static void batch_mem_free(void *p, int pinned) { if (!p) return; if (!pinned) free(p); /* page-locked blocks stay in their slab */ }
#define BATCH_PINNED(B) (g_pinned && !(B)->plain)
static void *pinned_grow(void *old, size_t old_bytes, size_t new_bytes, int pinned) {
  void *p = 0;
  if (!pinned) { p = malloc(new_bytes); if (!p) bq_fatal("[pileup] out of memory\n"); }
  else if (!(p = pinned_carve(new_bytes))) bq_fatal("[pileup] cannot allocate %zu bytes of page-locked memory: %s\n", new_bytes, bsq_last_error());
  if (old) { memcpy(p, old, old_bytes); batch_mem_free(old, pinned); }
  return p;
}
#define GROW(ptr, n, cap, extra)                                                                              \
  do {                                                                                                        \
    if ((n) + (extra) > (cap)) {                                                                              \
      int64_t ncap_ = ((n) + (extra)) * 5 / 4 + (1 << 16);                                                     \
      if (pin_ && ncap_ < (int64_t)(1 << 20)) ncap_ = (int64_t)(1 << 20); /* page-locking is slow: few, large steps */ \
      if (pin_ && ncap_ < 2 * (int64_t)(cap)) ncap_ = 2 * (int64_t)(cap);                                          \
      (ptr) = pinned_grow((ptr), (size_t)(n) * sizeof *(ptr), (size_t)ncap_ * sizeof *(ptr), pin_);                 \
      (cap) = ncap_;                                                                                          \
    }                                                                                                         \
  } while (0)

void bq_plp_batch_reset(bq_plp_batch_t *B) { B->n = 0; B->n_cig = 0; B->n_seq = 0; B->n_qual = 0; }

void bq_plp_batch_free(bq_plp_batch_t *B) {
  void *ps[] = {B->pos, B->mpos, B->mate_rlen, B->l_qseq, B->nm, B->as, B->flag, B->mapq, B->bss_tag, B->sid, B->n_cigar, B->cigar_off, B->cigar,
                B->seq_off, B->seq, B->qual_off, B->qual, B->end};
  for (size_t i = 0; i < sizeof ps / sizeof ps[0]; ++i) batch_mem_free(ps[i], BATCH_PINNED(B));
  memset(B, 0, sizeof *B);
}

static void batch_room(bq_plp_batch_t *B, int64_t n_cig, int64_t n_seq, int64_t n_qual) {
  const int pin_ = BATCH_PINNED(B);
  if (B->n + 1 > B->cap) {
    int64_t ncap = (B->n + 1) * 2 + (1 << 14);
    if (pin_ && ncap < 262144) ncap = 262144;
#define R(f) B->f = pinned_grow(B->f, (size_t)B->n * sizeof *B->f, (size_t)ncap * sizeof *B->f, pin_)
    R(pos); R(mpos); R(mate_rlen); R(l_qseq); R(nm); R(as); R(flag); R(mapq); R(bss_tag); R(sid); R(n_cigar); R(cigar_off); R(seq_off);
    R(qual_off); R(end);
#undef R
    B->cap = ncap;
  }
  GROW(B->cigar, B->n_cig, B->cap_cig, n_cig);
  GROW(B->seq, B->n_seq, B->cap_seq, n_seq);
  GROW(B->qual, B->n_qual, B->cap_qual, n_qual);
}

/* Fill slot i of the batch from BAM record `r` (after block_size); the variable-length parts go to the given offsets
 * (the caller has made room). */
static void batch_fill1(bq_plp_batch_t *B, int64_t i, int64_t cig_off, int64_t seq_off, int64_t qual_off, const uint8_t *r, uint32_t len, int sid) {
  const int32_t pos = (int32_t)le32(r + 4);
  const uint32_t l_read_name = r[8], mapq = r[9], n_cig = le16(r + 12), flag = le16(r + 14);
  const int32_t l_seq = (int32_t)le32(r + 16), mpos = (int32_t)le32(r + 24);
  const uint8_t *cig = r + 32 + l_read_name;
  const uint8_t *seq = cig + 4 * (size_t)n_cig;
  const uint8_t *qual = seq + ((size_t)l_seq + 1) / 2;
  const uint8_t *aux = qual + l_seq, *end = r + len;
  if (aux > end) bq_fatal("[bam] corrupt record\n");
  B->pos[i] = pos; B->mpos[i] = mpos; B->l_qseq[i] = l_seq; B->flag[i] = (uint16_t)flag; B->mapq[i] = (uint8_t)mapq; B->sid[i] = (uint8_t)sid;
  B->n_cigar[i] = (int32_t)n_cig; B->cigar_off[i] = cig_off; B->seq_off[i] = seq_off; B->qual_off[i] = qual_off;
  int64_t rlen = 0;
  for (uint32_t k = 0; k < n_cig; ++k) {
    const uint32_t c = le32(cig + 4 * k), op = c & 0xf;
    B->cigar[cig_off + k] = c;
    if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) rlen += c >> 4;
  }
  B->end[i] = (int64_t)pos + (rlen > 0 ? rlen : 1);
  memcpy(B->seq + seq_off, seq, ((size_t)l_seq + 1) / 2);
  memcpy(B->qual + qual_off, qual, (size_t)l_seq);
  /* tags read by process_func (src/pileup.c:709-746) */
  const uint8_t *t;
  t = aux_get(aux, end, "NM"); B->nm[i] = t ? (int32_t)aux2i(t) : INT32_MIN;
  t = aux_get(aux, end, "AS"); B->as[i] = t ? (int32_t)aux2i(t) : INT32_MIN;
  t = aux_get(aux, end, "MC"); B->mate_rlen[i] = (t && (*t == 'Z' || *t == 'H')) ? mc_rlen((const char *)t + 1) : (t ? 0 : -1);
  /* get_bsstrand (src/bisc_utils.c:208-238): YD, then ZS, then XG, else infer from the read */
  int bss = -1;
  t = aux_get(aux, end, "YD");
  if (t) { if (t[1] == 'f') bss = 0; else if (t[1] == 'r') bss = 1; }
  if (bss < 0) { t = aux_get(aux, end, "ZS"); if (t) { if (t[1] == '+') bss = 0; else if (t[1] == '-') bss = 1; } }
  if (bss < 0) {
    t = aux_get(aux, end, "XG");
    if (t && *t == 'Z') { if (strcmp((const char *)t + 1, "CT") == 0) bss = 0; else if (strcmp((const char *)t + 1, "GA") == 0) bss = 1; }
  }
  B->bss_tag[i] = (int8_t)bss;
}

/* Append BAM record `r` (after block_size) of sample `sid`.  Returns 0, or -1 if the record was dropped
 * because it can never produce an event (unmapped / no CIGAR). */
int bq_plp_batch_push(bq_plp_batch_t *B, const uint8_t *r, uint32_t len, int sid) {
  const uint32_t n_cig = le16(r + 12);
  const int32_t l_seq = (int32_t)le32(r + 16);
  if (n_cig == 0) return -1;
  batch_room(B, n_cig, ((int64_t)l_seq + 1) / 2, l_seq);
  batch_fill1(B, B->n, B->n_cig, B->n_seq, B->n_qual, r, len, sid);
  B->n_cig += n_cig; B->n_seq += ((int64_t)l_seq + 1) / 2; B->n_qual += l_seq;
  B->n++;
  return 0;
}

/* ---- bulk append: every following record of reference `tid` that starts before pos_lt ----
 * The record headers of the inflated window are walked serially (a few nanoseconds each: sizes and offsets), the
 * records are then decoded into the batch by n_threads threads.  Returns 1 when the stream has moved past the range
 * (next record belongs to another reference / starts at or after pos_lt / end of file), 0 never. */
typedef struct { const uint8_t *r; uint32_t len; int64_t cig_off, seq_off, qual_off; } rec_ref_t;
typedef struct { bq_plp_batch_t *B; const rec_ref_t *recs; int64_t base, lo, hi; int sid; } fill_job_t;
static void *fill_worker(void *arg) {
  fill_job_t *j = arg;
  for (int64_t k = j->lo; k < j->hi; ++k)
    batch_fill1(j->B, j->base + k, j->recs[k].cig_off, j->recs[k].seq_off, j->recs[k].qual_off, j->recs[k].r, j->recs[k].len, j->sid);
  return 0;
}

int bq_plp_batch_fill(bq_plp_batch_t *B, bq_bgzf_t *b, int sid, int tid, int64_t pos_lt, int n_threads) {
  const int pin_ = BATCH_PINNED(B);
  rec_ref_t *recs = 0;
  int64_t cap = 0;
  for (;;) {
    uint32_t len0;
    if (!bq_bam_peek(b, &len0)) return 1; /* end of file */
    /* walk the complete records that are in the inflated window now */
    int64_t n = 0, n_cig = 0, n_seq = 0, n_qual = 0;
    size_t off = b->ubeg;
    int past = 0;
    while (off + 4 <= b->uend) {
      const uint8_t *p = b->u + off;
      const uint32_t bs = le32(p);
      if (bs < 32) bq_fatal("[bam] %s: bad record size %u\n", b->fn, bs);
      if (off + 4 + (size_t)bs > b->uend) break; /* continues in the next window */
      const uint8_t *r = p + 4;
      if ((int32_t)le32(r) != tid || (int32_t)le32(r + 4) >= pos_lt) { past = 1; break; }
      const uint32_t nc = le16(r + 12);
      const int32_t l_seq = (int32_t)le32(r + 16);
      if (nc > 0) {
        if (n == cap) { cap = cap ? cap * 2 : 1 << 16; recs = realloc(recs, (size_t)cap * sizeof *recs); }
        recs[n].r = r; recs[n].len = bs; recs[n].cig_off = B->n_cig + n_cig; recs[n].seq_off = B->n_seq + n_seq; recs[n].qual_off = B->n_qual + n_qual;
        ++n; n_cig += nc; n_seq += ((int64_t)l_seq + 1) / 2; n_qual += l_seq;
      }
      off += 4 + (size_t)bs;
    }
    if (n > 0) {
      /* room for all of them, then decode in parallel */
      if (B->n + n > B->cap) {
        int64_t ncap = (B->n + n) * 5 / 4 + (1 << 14);
        if (pin_ && ncap < 262144) ncap = 262144;
#define R(f) B->f = pinned_grow(B->f, (size_t)B->n * sizeof *B->f, (size_t)ncap * sizeof *B->f, pin_)
        R(pos); R(mpos); R(mate_rlen); R(l_qseq); R(nm); R(as); R(flag); R(mapq); R(bss_tag); R(sid); R(n_cigar); R(cigar_off); R(seq_off);
        R(qual_off); R(end);
#undef R
        B->cap = ncap;
        if (pin_) { /* page-locked payload arrays in one step each, sized from what the records of this window carry per read */
          const int64_t room = ncap - B->n;
          GROW(B->cigar, B->n_cig, B->cap_cig, n_cig * room / n + 1);
          GROW(B->seq, B->n_seq, B->cap_seq, n_seq * room / n + 1);
          GROW(B->qual, B->n_qual, B->cap_qual, n_qual * room / n + 1);
        }
      }
      GROW(B->cigar, B->n_cig, B->cap_cig, n_cig);
      GROW(B->seq, B->n_seq, B->cap_seq, n_seq);
      GROW(B->qual, B->n_qual, B->cap_qual, n_qual);
      int nt = n_threads < 1 ? 1 : (n_threads > 32 ? 32 : n_threads);
      if (n < 4096) nt = 1;
      const double tf_ = bam_now();
      fill_job_t jobs[32];
      pthread_t th[32];
      for (int t = 0; t < nt; ++t) { jobs[t].B = B; jobs[t].recs = recs; jobs[t].base = B->n; jobs[t].lo = n * t / nt; jobs[t].hi = n * (t + 1) / nt; jobs[t].sid = sid; }
      for (int t = 1; t < nt; ++t) pthread_create(&th[t], 0, fill_worker, &jobs[t]);
      fill_worker(&jobs[0]);
      for (int t = 1; t < nt; ++t) pthread_join(th[t], 0);
      bq_plp_t_fill += bam_now() - tf_;
      B->n += n; B->n_cig += n_cig; B->n_seq += n_seq; B->n_qual += n_qual;
    }
    b->ubeg = off; /* consumed */
    if (past) { free(recs); return 1; }
    /* window exhausted (possibly in the middle of a record): the next peek refills */
  }
}

void bq_plp_batch_view(const bq_plp_batch_t *B, bsq_plp_reads *v) {
  v->n_reads = B->n; v->pos = B->pos; v->mpos = B->mpos; v->mate_rlen = B->mate_rlen; v->l_qseq = B->l_qseq; v->nm = B->nm; v->as = B->as;
  v->flag = B->flag; v->mapq = B->mapq; v->bss_tag = B->bss_tag; v->sid = B->sid; v->n_cigar = B->n_cigar; v->cigar_off = B->cigar_off;
  v->cigar = B->cigar; v->seq_off = B->seq_off; v->seq = B->seq; v->qual_off = B->qual_off; v->qual = B->qual;
}

/* ------------------------------------------------------------------ FASTA ---- */

/* Index a FASTA file held in memory: names (up to the first white space) and the byte range of each
 * record's sequence lines.  (The reference goes through htslib faidx, which needs/creates <fa>.fai; reading
 * the file itself gives the same bases without the side file.) */
int bq_fasta_load(const char *fn, bq_fasta_t *fa) {
  memset(fa, 0, sizeof *fa);
  gzFile fp = gzopen(fn, "rb");
  if (!fp) return -1;
  gzbuffer(fp, 1 << 20);
  size_t cap = 1 << 26, n = 0;
  char *buf = malloc(cap);
  for (;;) {
    if (n + (1 << 24) > cap) { cap *= 2; buf = realloc(buf, cap); }
    const int got = gzread(fp, buf + n, 1 << 24);
    if (got < 0) { free(buf); gzclose(fp); return -1; }
    if (got == 0) break;
    n += (size_t)got;
  }
  gzclose(fp);
  fa->buf = buf; fa->n_buf = n;
  size_t i = 0;
  while (i < n) {
    if (buf[i] != '>') { /* skip to the next header */
      const char *nl = memchr(buf + i, '\n', n - i);
      if (!nl) break;
      i = (size_t)(nl - buf) + 1;
      continue;
    }
    const size_t h0 = i + 1;
    const char *nl = memchr(buf + i, '\n', n - i);
    const size_t hend = nl ? (size_t)(nl - buf) : n;
    size_t ne = h0;
    while (ne < hend && !isspace((unsigned char)buf[ne])) ++ne;
    if (fa->n % 64 == 0) {
      fa->name = realloc(fa->name, (size_t)(fa->n + 64) * sizeof(char *));
      fa->beg = realloc(fa->beg, (size_t)(fa->n + 64) * sizeof(size_t));
      fa->endp = realloc(fa->endp, (size_t)(fa->n + 64) * sizeof(size_t));
    }
    fa->name[fa->n] = strndup(buf + h0, ne - h0);
    size_t s0 = hend < n ? hend + 1 : n, s1 = s0;
    /* the record runs until the next line that starts with '>' */
    while (s1 < n) {
      if (buf[s1] == '>') break;
      const char *q = memchr(buf + s1, '\n', n - s1);
      s1 = q ? (size_t)(q - buf) + 1 : n;
    }
    fa->beg[fa->n] = s0; fa->endp[fa->n] = s1;
    fa->n++;
    i = s1;
  }
  return 0;
}

void bq_fasta_free(bq_fasta_t *fa) {
  for (int i = 0; i < fa->n; ++i) free(fa->name[i]);
  free(fa->name); free(fa->beg); free(fa->endp); free(fa->buf);
  memset(fa, 0, sizeof *fa);
}

/* nt4 codes (A0 C1 G2 T3, anything else 4) of contig `name`; returns length or -1 */
int64_t bq_fasta_fetch_nt4(const bq_fasta_t *fa, const char *name, uint8_t **out) {
  static uint8_t tab[256];
  static int init = 0;
  if (!init) {
    memset(tab, 4, 256);
    tab['A'] = tab['a'] = 0; tab['C'] = tab['c'] = 1; tab['G'] = tab['g'] = 2; tab['T'] = tab['t'] = 3;
    tab['\n'] = tab['\r'] = tab[' '] = tab['\t'] = 255;
    init = 1;
  }
  for (int i = 0; i < fa->n; ++i) {
    if (strcmp(fa->name[i], name) != 0) continue;
    const size_t b = fa->beg[i], e = fa->endp[i];
    uint8_t *o = malloc(e - b + 16);
    int64_t n = 0;
    for (size_t k = b; k < e; ++k) {
      const uint8_t c = tab[(uint8_t)fa->buf[k]];
      if (c != 255) o[n++] = c;
    }
    *out = o;
    return n;
  }
  return -1;
}

/* ------------------------------------------------------------------ seek / BAI ---- */

void bq_bgzf_seek(bq_bgzf_t *b, uint64_t voffset) {
  if (fseeko(b->fp, (off_t)(voffset >> 16), SEEK_SET) != 0) bq_fatal("[bgzf] %s: seek failed\n", b->fn);
  b->eof = 0; b->ubeg = b->uend = 0; b->n_blk = 0;
  const size_t within = (size_t)(voffset & 0xffff);
  if (within) {
    if (bgzf_refill(b) == 0 || b->uend < within) bq_fatal("[bgzf] %s: bad virtual offset\n", b->fn);
    b->ubeg = within;
  }
}

void bq_plp_batch_copy1(bq_plp_batch_t *D, const bq_plp_batch_t *S, int64_t i) {
  const int64_t nc = S->n_cigar[i], ls = S->l_qseq[i] > 0 ? S->l_qseq[i] : 0;
  batch_room(D, nc, (ls + 1) / 2, ls);
  const int64_t j = D->n;
  D->pos[j] = S->pos[i]; D->mpos[j] = S->mpos[i]; D->mate_rlen[j] = S->mate_rlen[i]; D->l_qseq[j] = S->l_qseq[i]; D->nm[j] = S->nm[i];
  D->as[j] = S->as[i]; D->flag[j] = S->flag[i]; D->mapq[j] = S->mapq[i]; D->bss_tag[j] = S->bss_tag[i]; D->sid[j] = S->sid[i];
  D->n_cigar[j] = S->n_cigar[i]; D->end[j] = S->end[i];
  D->cigar_off[j] = D->n_cig; D->seq_off[j] = D->n_seq; D->qual_off[j] = D->n_qual;
  memcpy(D->cigar + D->n_cig, S->cigar + S->cigar_off[i], (size_t)nc * 4);
  memcpy(D->seq + D->n_seq, S->seq + S->seq_off[i], (size_t)(ls + 1) / 2);
  memcpy(D->qual + D->n_qual, S->qual + S->qual_off[i], (size_t)ls);
  D->n_cig += nc; D->n_seq += (ls + 1) / 2; D->n_qual += ls;
  D->n++;
}

int bq_bai_load(const char *bam_fn, bq_bai_t *bai) {
  memset(bai, 0, sizeof *bai);
  char *fn = malloc(strlen(bam_fn) + 8);
  sprintf(fn, "%s.bai", bam_fn);
  FILE *fp = fopen(fn, "rb");
  if (!fp) { /* foo.bam -> foo.bai */
    const size_t l = strlen(bam_fn);
    if (l > 4 && strcmp(bam_fn + l - 4, ".bam") == 0) { strcpy(fn, bam_fn); strcpy(fn + l - 4, ".bai"); fp = fopen(fn, "rb"); }
  }
  free(fn);
  if (!fp) return -1;
  uint8_t h[8];
  if (fread(h, 1, 8, fp) != 8 || memcmp(h, "BAI\1", 4) != 0) { fclose(fp); return -1; }
  bai->n_ref = (int32_t)le32(h + 4);
  bai->first = malloc((size_t)(bai->n_ref + 1) * 8);
  bai->n_intv = calloc((size_t)bai->n_ref + 1, 4);
  bai->ioffset = calloc((size_t)bai->n_ref + 1, sizeof(uint64_t *));
  for (int r = 0; r < bai->n_ref; ++r) {
    uint8_t w[8];
    uint64_t first = UINT64_MAX;
    if (fread(w, 1, 4, fp) != 4) goto bad;
    const int32_t n_bin = (int32_t)le32(w);
    for (int32_t k = 0; k < n_bin; ++k) {
      if (fread(w, 1, 8, fp) != 8) goto bad;
      const uint32_t bin = le32(w);
      const int32_t n_chunk = (int32_t)le32(w + 4);
      for (int32_t c = 0; c < n_chunk; ++c) {
        uint8_t ch[16];
        if (fread(ch, 1, 16, fp) != 16) goto bad;
        if (bin == 37450) continue; /* pseudo-bin: offsets / counts, not record chunks */
        const uint64_t beg = (uint64_t)le32(ch) | (uint64_t)le32(ch + 4) << 32;
        if (beg < first) first = beg;
      }
    }
    if (fread(w, 1, 4, fp) != 4) goto bad;
    const int32_t n_intv = (int32_t)le32(w);
    bai->n_intv[r] = n_intv;
    bai->ioffset[r] = malloc((size_t)(n_intv + 1) * 8);
    for (int32_t k = 0; k < n_intv; ++k) {
      if (fread(w, 1, 8, fp) != 8) goto bad;
      bai->ioffset[r][k] = (uint64_t)le32(w) | (uint64_t)le32(w + 4) << 32;
    }
    bai->first[r] = first;
  }
  fclose(fp);
  return 0;
bad:
  fclose(fp);
  bq_bai_free(bai);
  return -1;
}

void bq_bai_free(bq_bai_t *bai) {
  if (bai->ioffset) for (int r = 0; r < bai->n_ref; ++r) free(bai->ioffset[r]);
  free(bai->ioffset); free(bai->n_intv); free(bai->first);
  memset(bai, 0, sizeof *bai);
}

/* virtual offset from which a front-to-back scan sees every record of `tid` overlapping position >= beg0 */
uint64_t bq_bai_start(const bq_bai_t *bai, int tid, int64_t beg0) {
  if (tid < 0 || tid >= bai->n_ref || bai->first[tid] == UINT64_MAX) return UINT64_MAX;
  uint64_t v = bai->first[tid];
  if (beg0 > 0 && bai->n_intv[tid] > 0) {
    int64_t w = beg0 >> 14;
    if (w >= bai->n_intv[tid]) w = bai->n_intv[tid] - 1;
    /* the linear index gives the smallest offset of a record overlapping window w; empty windows hold 0 */
    while (w >= 0 && bai->ioffset[tid][w] == 0) --w;
    if (w >= 0 && bai->ioffset[tid][w] > v) v = bai->ioffset[tid][w];
  }
  return v;
}

/* `biscuit bamdump in.bam [tid [beg0]]` -- decoded view of the records as pileup sees them (diagnostics / tests):
 * one line per record: tid pos mpos flag mapq l_qseq NM AS mate_rlen bss n_cigar cigar... */
int bq_main_bamdump(int argc, char **argv) {
  if (argc < 2) { fprintf(stderr, "Usage: biscuit bamdump <in.bam> [tid [beg0]]\n"); return 1; }
  g_pinned = 0;
  bq_bgzf_t *fp = bq_bgzf_open(argv[1], 2);
  if (!fp) bq_fatal("Cannot open %s\n", argv[1]);
  bq_bam_hdr_t h;
  if (bq_bam_read_header(fp, &h) != 0) bq_fatal("%s is not a BAM file\n", argv[1]);
  int only_tid = -1;
  if (argc > 2) {
    only_tid = atoi(argv[2]);
    bq_bai_t bai;
    if (bq_bai_load(argv[1], &bai) != 0) bq_fatal("Cannot load index of %s\n", argv[1]);
    const uint64_t v = bq_bai_start(&bai, only_tid, argc > 3 ? atoll(argv[3]) : 0);
    bq_bai_free(&bai);
    if (v == UINT64_MAX) { bq_bam_hdr_free(&h); bq_bgzf_close(fp); return 0; }
    bq_bgzf_seek(fp, v);
  }
  for (int i = 0; i < h.n_targets; ++i) printf("@\t%s\t%d\n", h.name[i], h.len[i]);
  bq_plp_batch_t B;
  memset(&B, 0, sizeof B);
  uint32_t len;
  const uint8_t *r;
  while ((r = bq_bam_next(fp, &len))) {
    const int32_t tid = (int32_t)le32(r);
    if (only_tid >= 0 && tid != only_tid) break;
    bq_plp_batch_reset(&B);
    if (bq_plp_batch_push(&B, r, len, 0) != 0) { printf("%d\t%d\tno-cigar\n", tid, (int32_t)le32(r + 4)); continue; }
    printf("%d\t%d\t%d\t%u\t%u\t%d\t%d\t%d\t%d\t%d\t%d", tid, B.pos[0], B.mpos[0], B.flag[0], B.mapq[0], B.l_qseq[0], B.nm[0], B.as[0],
           B.mate_rlen[0], B.bss_tag[0], B.n_cigar[0]);
    for (int k = 0; k < B.n_cigar[0]; ++k) printf("\t%u", B.cigar[k]);
    uint32_t cs = 0;
    for (int64_t k = 0; k < B.n_seq; ++k) cs = cs * 31 + B.seq[k];
    for (int64_t k = 0; k < B.n_qual; ++k) cs = cs * 31 + B.qual[k];
    printf("\t%u\n", cs);
  }
  bq_plp_batch_free(&B);
  bq_bam_hdr_free(&h);
  bq_bgzf_close(fp);
  return 0;
}
