/* bq_sortbam.c -- `biscuit sortbam [-@ T] [-l LEVEL] -o out.bam in.sam`: SAM text -> coordinate-sorted BAM + BAI.
 *
 * Not part of the reference: its workflow pipes `biscuit align` into samtools for this step (README.md:33-38), and
 * samtools/htslib are not available here.  It closes the gap between `biscuit align` and `biscuit pileup` so that the
 * whole index -> align -> pileup -> vcf2bed chain runs from this tree.  Formats from the
 * SAM/BAM v1 specification (sections 4.2, 4.1 BGZF, 5.2 BAI).  The whole file is held in memory: a utility for
 * moderate inputs, not a replacement for an external-memory sorter.
 */
#include <ctype.h>
#include <getopt.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>
#include "bq_plp.h"

#define BGZF_BLOCK 0xff00

typedef struct { uint8_t *s; size_t l, m; } buf_t;
static void buf_room(buf_t *b, size_t n) {
  if (b->l + n > b->m) { b->m = (b->l + n) * 2 + 65536; b->s = realloc(b->s, b->m); }
}
static void put8(buf_t *b, unsigned v) { buf_room(b, 1); b->s[b->l++] = (uint8_t)v; }
static void put16(buf_t *b, unsigned v) { buf_room(b, 2); b->s[b->l++] = v & 0xff; b->s[b->l++] = (v >> 8) & 0xff; }
static void put32(buf_t *b, uint32_t v) { buf_room(b, 4); for (int i = 0; i < 4; ++i) b->s[b->l++] = (v >> (8 * i)) & 0xff; }
static void putn(buf_t *b, const void *p, size_t n) { buf_room(b, n); memcpy(b->s + b->l, p, n); b->l += n; }

static int reg2bin(int64_t beg, int64_t end) {
  --end;
  if (beg >> 14 == end >> 14) return (int)(((1 << 15) - 1) / 7 + (beg >> 14));
  if (beg >> 17 == end >> 17) return (int)(((1 << 12) - 1) / 7 + (beg >> 17));
  if (beg >> 20 == end >> 20) return (int)(((1 << 9) - 1) / 7 + (beg >> 20));
  if (beg >> 23 == end >> 23) return (int)(((1 << 6) - 1) / 7 + (beg >> 23));
  if (beg >> 26 == end >> 26) return (int)(((1 << 3) - 1) / 7 + (beg >> 26));
  return 0;
}

typedef struct { int32_t tid, pos; int64_t end; size_t off; uint32_t len; int64_t idx; } rec_t;
static int cmp_rec(const void *a, const void *b) {
  const rec_t *x = a, *y = b;
  const uint32_t tx = (uint32_t)x->tid, ty = (uint32_t)y->tid; /* unmapped (tid -1) last */
  if (tx != ty) return tx < ty ? -1 : 1;
  if (x->pos != y->pos) return x->pos < y->pos ? -1 : 1;
  return x->idx < y->idx ? -1 : x->idx > y->idx;
}

static int name2tid(char **names, int n, const char *s) {
  for (int i = 0; i < n; ++i) if (strcmp(names[i], s) == 0) return i;
  return -1;
}

static void put_int_tag(buf_t *b, long long v) { /* smallest type that holds the value */
  if (v >= 0) {
    if (v <= 255) { put8(b, 'C'); put8(b, (unsigned)v); }
    else if (v <= 65535) { put8(b, 'S'); put16(b, (unsigned)v); }
    else { put8(b, 'I'); put32(b, (uint32_t)v); }
  } else {
    if (v >= -128) { put8(b, 'c'); put8(b, (unsigned)(v & 0xff)); }
    else if (v >= -32768) { put8(b, 's'); put16(b, (unsigned)(v & 0xffff)); }
    else { put8(b, 'i'); put32(b, (uint32_t)(int32_t)v); }
  }
}

typedef struct { const uint8_t *src; size_t n_blocks, total; uint8_t **out; uint32_t *clen; int level, t, nt; } dj_t;
static void *deflate_worker(void *arg) {
  dj_t *j = arg;
  for (size_t k = (size_t)j->t; k < j->n_blocks; k += (size_t)j->nt) {
    const size_t off = k * BGZF_BLOCK, ulen = off + BGZF_BLOCK <= j->total ? BGZF_BLOCK : j->total - off;
    uint8_t *o = malloc(BGZF_BLOCK + 1024);
    z_stream zs;
    memset(&zs, 0, sizeof zs);
    deflateInit2(&zs, j->level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY);
    zs.next_in = (Bytef *)(j->src + off); zs.avail_in = (uInt)ulen;
    zs.next_out = o + 18; zs.avail_out = BGZF_BLOCK + 1024 - 26;
    deflate(&zs, Z_FINISH);
    const uint32_t cl = (uint32_t)zs.total_out;
    deflateEnd(&zs);
    static const uint8_t hdr[16] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0};
    memcpy(o, hdr, 16);
    const uint32_t bsize = cl + 25;
    o[16] = bsize & 0xff; o[17] = (bsize >> 8) & 0xff;
    const uint32_t crc = (uint32_t)crc32(crc32(0, 0, 0), j->src + off, (uInt)ulen);
    uint8_t *t = o + 18 + cl;
    for (int i = 0; i < 4; ++i) t[i] = (crc >> (8 * i)) & 0xff;
    for (int i = 0; i < 4; ++i) t[4 + i] = ((uint32_t)ulen >> (8 * i)) & 0xff;
    j->out[k] = o; j->clen[k] = cl + 26;
  }
  return 0;
}

int bq_main_sortbam(int argc, char **argv) {
  int c, n_threads = 4, level = 6;
  const char *outfn = 0;
  while ((c = getopt(argc, argv, "@:l:o:h")) >= 0) {
    if (c == '@') n_threads = atoi(optarg);
    else if (c == 'l') level = atoi(optarg);
    else if (c == 'o') outfn = optarg;
    else { fprintf(stderr, "Usage: biscuit sortbam [-@ threads] [-l level] -o out.bam in.sam\n"); return 1; }
  }
  if (!outfn || optind >= argc) { fprintf(stderr, "Usage: biscuit sortbam [-@ threads] [-l level] -o out.bam in.sam\n"); return 1; }
  if (n_threads < 1) n_threads = 1;
  if (n_threads > 64) n_threads = 64;
  gzFile fp = strcmp(argv[optind], "-") == 0 ? gzdopen(0, "r") : gzopen(argv[optind], "r");
  if (!fp) bq_fatal("[sortbam] cannot open %s\n", argv[optind]);
  gzbuffer(fp, 1 << 20);
  buf_t text = {0, 0, 0}, recs_b = {0, 0, 0};
  char **names = 0;
  int32_t *lens = 0;
  int n_ref = 0;
  rec_t *recs = 0;
  size_t n_rec = 0, m_rec = 0;
  size_t cap = 1 << 16;
  char *line = malloc(cap);
  static const char *ops = "MIDNSHP=XB";
  uint8_t nt16[256];
  memset(nt16, 15, 256);
  { const char *t = "=ACMGRSVTWYHKDBN"; for (int i = 0; i < 16; ++i) { nt16[(uint8_t)t[i]] = (uint8_t)i; nt16[(uint8_t)tolower(t[i])] = (uint8_t)i; } }
  for (;;) {
    size_t n = 0;
    for (;;) { /* one line of any length */
      if (!gzgets(fp, line + n, (int)(cap - n))) break;
      n += strlen(line + n);
      if (n && line[n - 1] == '\n') break;
      if (cap - n < 4096) { cap *= 2; line = realloc(line, cap); }
    }
    if (n == 0) break;
    while (n && (line[n - 1] == '\n' || line[n - 1] == '\r')) line[--n] = 0;
    if (line[0] == '@') {
      putn(&text, line, n); put8(&text, '\n');
      if (strncmp(line, "@SQ", 3) == 0) {
        char *sn = strstr(line, "\tSN:"), *ln = strstr(line, "\tLN:");
        if (!sn || !ln) bq_fatal("[sortbam] malformed @SQ line\n");
        sn += 4;
        char *e = strchr(sn, '\t');
        names = realloc(names, sizeof(char *) * (size_t)(n_ref + 1));
        lens = realloc(lens, sizeof(int32_t) * (size_t)(n_ref + 1));
        names[n_ref] = e ? strndup(sn, (size_t)(e - sn)) : strdup(sn);
        lens[n_ref] = atoi(ln + 4);
        ++n_ref;
      }
      continue;
    }
    /* alignment line */
    char *f[12] = {0};
    int nf = 0;
    char *p = line;
    while (nf < 11) { f[nf++] = p; p = strchr(p, '\t'); if (!p) break; *p++ = 0; }
    if (nf < 11) bq_fatal("[sortbam] alignment line with %d fields\n", nf);
    char *tags = p; /* NULL when there are none */
    const int flag = atoi(f[1]), tid = strcmp(f[2], "*") ? name2tid(names, n_ref, f[2]) : -1;
    const int32_t pos = atoi(f[3]) - 1;
    const int mtid = strcmp(f[6], "=") == 0 ? tid : (strcmp(f[6], "*") ? name2tid(names, n_ref, f[6]) : -1);
    const size_t start = recs_b.l;
    put32(&recs_b, 0); /* block_size, patched below */
    put32(&recs_b, (uint32_t)tid); put32(&recs_b, (uint32_t)pos);
    const size_t l_name = strlen(f[0]) + 1;
    put8(&recs_b, (unsigned)l_name); put8(&recs_b, (unsigned)atoi(f[4]));
    const size_t bin_at = recs_b.l;
    put16(&recs_b, 0);
    /* CIGAR */
    uint32_t cig[65536];
    int n_cig = 0;
    int64_t rlen = 0;
    if (strcmp(f[5], "*")) {
      for (char *q = f[5]; *q;) {
        char *e;
        const long v = strtol(q, &e, 10);
        const char *o = strchr(ops, *e);
        if (!o || !*e) bq_fatal("[sortbam] bad CIGAR %s\n", f[5]);
        const int op = (int)(o - ops);
        if (n_cig == 65535) bq_fatal("[sortbam] CIGAR with more than 65535 operations\n");
        cig[n_cig++] = (uint32_t)v << 4 | (uint32_t)op;
        if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) rlen += v;
        q = e + 1;
      }
    }
    put16(&recs_b, (unsigned)n_cig); put16(&recs_b, (unsigned)flag);
    const size_t l_seq = strcmp(f[9], "*") ? strlen(f[9]) : 0;
    put32(&recs_b, (uint32_t)l_seq); put32(&recs_b, (uint32_t)mtid); put32(&recs_b, (uint32_t)(atoi(f[7]) - 1)); put32(&recs_b, (uint32_t)atoi(f[8]));
    putn(&recs_b, f[0], l_name);
    for (int k = 0; k < n_cig; ++k) put32(&recs_b, cig[k]);
    for (size_t k = 0; k < l_seq; k += 2) put8(&recs_b, (unsigned)(nt16[(uint8_t)f[9][k]] << 4 | (k + 1 < l_seq ? nt16[(uint8_t)f[9][k + 1]] : 0)));
    if (strcmp(f[10], "*") == 0) for (size_t k = 0; k < l_seq; ++k) put8(&recs_b, 0xff);
    else { if (strlen(f[10]) != l_seq) bq_fatal("[sortbam] SEQ and QUAL differ in length\n"); for (size_t k = 0; k < l_seq; ++k) put8(&recs_b, (unsigned)(f[10][k] - 33)); }
    for (char *t = tags; t && *t;) { /* TAG:TYPE:VALUE */
      char *e = strchr(t, '\t');
      if (e) *e = 0;
      if (strlen(t) < 5 || t[2] != ':' || t[4] != ':') bq_fatal("[sortbam] bad tag %s\n", t);
      putn(&recs_b, t, 2);
      const char ty = t[3], *v = t + 5;
      if (ty == 'i') put_int_tag(&recs_b, atoll(v));
      else if (ty == 'A') { put8(&recs_b, 'A'); put8(&recs_b, (unsigned)v[0]); }
      else if (ty == 'f') { float x = (float)atof(v); put8(&recs_b, 'f'); putn(&recs_b, &x, 4); }
      else if (ty == 'Z' || ty == 'H') { put8(&recs_b, (unsigned)ty); putn(&recs_b, v, strlen(v) + 1); }
      else bq_fatal("[sortbam] tag type %c is not supported\n", ty);
      t = e ? e + 1 : 0;
    }
    const int64_t end = pos + (rlen > 0 ? rlen : 1);
    const int bin = reg2bin(pos < 0 ? 0 : pos, end < 1 ? 1 : end);
    recs_b.s[bin_at] = bin & 0xff; recs_b.s[bin_at + 1] = (bin >> 8) & 0xff;
    const uint32_t bs = (uint32_t)(recs_b.l - start - 4);
    for (int i = 0; i < 4; ++i) recs_b.s[start + i] = (bs >> (8 * i)) & 0xff;
    if (n_rec == m_rec) { m_rec = m_rec ? m_rec * 2 : 1 << 16; recs = realloc(recs, m_rec * sizeof *recs); }
    recs[n_rec].tid = tid; recs[n_rec].pos = pos; recs[n_rec].end = end; recs[n_rec].off = start; recs[n_rec].len = bs + 4; recs[n_rec].idx = (int64_t)n_rec;
    ++n_rec;
  }
  gzclose(fp);
  qsort(recs, n_rec, sizeof *recs, cmp_rec);
  /* the uncompressed BAM stream: header, then the records in order */
  buf_t bam = {0, 0, 0};
  putn(&bam, "BAM\1", 4);
  put32(&bam, (uint32_t)text.l); putn(&bam, text.s, text.l);
  put32(&bam, (uint32_t)n_ref);
  for (int i = 0; i < n_ref; ++i) { const size_t l = strlen(names[i]) + 1; put32(&bam, (uint32_t)l); putn(&bam, names[i], l); put32(&bam, (uint32_t)lens[i]); }
  uint64_t *ustart = malloc(sizeof(uint64_t) * (n_rec + 1));
  for (size_t k = 0; k < n_rec; ++k) { ustart[k] = bam.l; putn(&bam, recs_b.s + recs[k].off, recs[k].len); }
  ustart[n_rec] = bam.l;
  /* BGZF: fixed-size blocks of the stream, deflated in parallel */
  const size_t n_blocks = (bam.l + BGZF_BLOCK - 1) / BGZF_BLOCK;
  uint8_t **cb = calloc(n_blocks + 1, sizeof *cb);
  uint32_t *clen = calloc(n_blocks + 1, sizeof *clen);
  dj_t jobs[64];
  pthread_t th[64];
  for (int t = 0; t < n_threads; ++t) { jobs[t].src = bam.s; jobs[t].n_blocks = n_blocks; jobs[t].total = bam.l; jobs[t].out = cb; jobs[t].clen = clen; jobs[t].level = level; jobs[t].t = t; jobs[t].nt = n_threads; }
  for (int t = 1; t < n_threads; ++t) pthread_create(&th[t], 0, deflate_worker, &jobs[t]);
  deflate_worker(&jobs[0]);
  for (int t = 1; t < n_threads; ++t) pthread_join(th[t], 0);
  uint64_t *coff = malloc(sizeof(uint64_t) * (n_blocks + 1));
  FILE *out = fopen(outfn, "wb");
  if (!out) bq_fatal("[sortbam] cannot write %s\n", outfn);
  uint64_t off = 0;
  for (size_t k = 0; k < n_blocks; ++k) { coff[k] = off; fwrite(cb[k], 1, clen[k], out); off += clen[k]; free(cb[k]); }
  coff[n_blocks] = off;
  static const uint8_t eof_block[28] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0, 0x1b, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  fwrite(eof_block, 1, 28, out);
  fclose(out);
#define VOFF(u) (((u) / BGZF_BLOCK < n_blocks ? coff[(u) / BGZF_BLOCK] : coff[n_blocks]) << 16 | ((u) / BGZF_BLOCK < n_blocks ? (u) % BGZF_BLOCK : 0))
  /* BAI: per reference the bins (chunks of consecutive records merged) and the 16 kb linear index */
  char *bai_fn = malloc(strlen(outfn) + 8);
  sprintf(bai_fn, "%s.bai", outfn);
  FILE *bf = fopen(bai_fn, "wb");
  if (!bf) bq_fatal("[sortbam] cannot write %s\n", bai_fn);
  buf_t bai = {0, 0, 0};
  putn(&bai, "BAI\1", 4); put32(&bai, (uint32_t)n_ref);
  size_t k = 0;
  for (int r = 0; r < n_ref; ++r) {
    const size_t k0 = k;
    while (k < n_rec && recs[k].tid == r) ++k;
    /* bins: sort record indices by (bin, position in file) */
    typedef struct { uint32_t bin; size_t i; } bk_t;
    const size_t nr = k - k0;
    bk_t *bk = malloc(sizeof(bk_t) * (nr + 1));
    for (size_t i = 0; i < nr; ++i) { bk[i].bin = (uint32_t)reg2bin(recs[k0 + i].pos < 0 ? 0 : recs[k0 + i].pos, recs[k0 + i].end < 1 ? 1 : recs[k0 + i].end); bk[i].i = k0 + i; }
    /* stable counting by bin */
    size_t n_bins = 0;
    uint32_t *cnt = calloc(37451, sizeof(uint32_t));
    for (size_t i = 0; i < nr; ++i) if (cnt[bk[i].bin]++ == 0) ++n_bins;
    put32(&bai, (uint32_t)n_bins);
    size_t *start = calloc(37452, sizeof(size_t));
    for (int b = 0; b < 37450; ++b) start[b + 1] = start[b] + cnt[b];
    size_t *ord = malloc(sizeof(size_t) * (nr + 1)), *fill = calloc(37451, sizeof(size_t));
    for (size_t i = 0; i < nr; ++i) ord[start[bk[i].bin] + fill[bk[i].bin]++] = bk[i].i;
    for (int b = 0; b < 37450; ++b) {
      if (!cnt[b]) continue;
      /* chunks: runs of records that are adjacent in the file */
      buf_t ch = {0, 0, 0};
      uint32_t n_chunk = 0;
      size_t i = start[b];
      while (i < start[b] + cnt[b]) {
        size_t j = i;
        while (j + 1 < start[b] + cnt[b] && ord[j + 1] == ord[j] + 1) ++j;
        const uint64_t v0 = VOFF(ustart[ord[i]]), v1 = VOFF(ustart[ord[j] + 1]);
        put32(&ch, (uint32_t)v0); put32(&ch, (uint32_t)(v0 >> 32)); put32(&ch, (uint32_t)v1); put32(&ch, (uint32_t)(v1 >> 32));
        ++n_chunk;
        i = j + 1;
      }
      put32(&bai, (uint32_t)b); put32(&bai, n_chunk); putn(&bai, ch.s, ch.l);
      free(ch.s);
    }
    /* linear index */
    int64_t max_end = 0;
    for (size_t i = k0; i < k; ++i) if (recs[i].end > max_end) max_end = recs[i].end;
    const int32_t n_intv = nr ? (int32_t)(((max_end - 1) >> 14) + 1) : 0;
    uint64_t *lin = calloc((size_t)n_intv + 1, sizeof(uint64_t));
    for (size_t i = k0; i < k; ++i) {
      const uint64_t v = VOFF(ustart[i]);
      for (int64_t w = (recs[i].pos < 0 ? 0 : recs[i].pos) >> 14; w <= (recs[i].end - 1) >> 14 && w < n_intv; ++w)
        if (lin[w] == 0 || v < lin[w]) lin[w] = v;
    }
    put32(&bai, (uint32_t)n_intv);
    uint64_t last = 0;
    for (int32_t w = 0; w < n_intv; ++w) { if (lin[w]) last = lin[w]; put32(&bai, (uint32_t)last); put32(&bai, (uint32_t)(last >> 32)); }
    free(lin); free(bk); free(cnt); free(start); free(ord); free(fill);
  }
  { uint64_t n_no_coor = n_rec - k; put32(&bai, (uint32_t)n_no_coor); put32(&bai, (uint32_t)(n_no_coor >> 32)); }
  fwrite(bai.s, 1, bai.l, bf);
  fclose(bf);
  fprintf(stderr, "[sortbam] %zu records, %d references, %zu BGZF blocks\n", n_rec, n_ref, n_blocks);
  free(bai.s); free(bai_fn); free(coff); free(cb); free(clen); free(ustart); free(bam.s); free(recs); free(recs_b.s); free(text.s); free(line);
  for (int i = 0; i < n_ref; ++i) free(names[i]);
  free(names); free(lens);
  return 0;
}
