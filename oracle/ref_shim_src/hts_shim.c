/* TEST INFRASTRUCTURE ONLY -- the part of htslib 1.18 that /root/reference/src/{pileup.c,bisc_utils.c,refcache.h,
 * mergecg.c} call, written from the SAM/BAM specification (see README.md).  Whole files are inflated / read into memory
 * and shared between handles, which is all the test sizes need; the region iterator applies htslib's own overlap rule
 * (hts_itr_next: same tid, pos < end, bam_endpos > beg) to the coordinate-sorted records. */
#include <ctype.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>
#include "faidx.h"
#include "sam.h"

const char seq_nt16_str[] = "=ACMGRSVTWYHKDBN";
const int8_t bam_cigar_table[256] = {
#define X -1
    X,X,X,X,X,X,X,X,X,X,X,X,X,X,X,X, X,X,X,X,X,X,X,X,X,X,X,X,X,X,X,X,
    X,X,X,X,X,X,X,X,X,X,X,X,X,X,X,X, X,X,X,X,X,X,X,X,X,X,X,X,X,BAM_CEQUAL,X,X,
    X,X,BAM_CBACK,X,BAM_CDEL,X,X,X,BAM_CHARD_CLIP,BAM_CINS,X,X,X,BAM_CMATCH,BAM_CREF_SKIP,X,
    BAM_CPAD,X,X,BAM_CSOFT_CLIP,X,X,X,X,BAM_CDIFF,X,X,X,X,X,X,X,
    X,X,X,X,X,X,X,X,X,X,X,X,X,X,X,X, X,X,X,X,X,X,X,X,X,X,X,X,X,X,X,X,
    X,X,X,X,X,X,X,X,X,X,X,X,X,X,X,X, X,X,X,X,X,X,X,X,X,X,X,X,X,X,X,X,
    X,X,X,X,X,X,X,X,X,X,X,X,X,X,X,X, X,X,X,X,X,X,X,X,X,X,X,X,X,X,X,X,
    X,X,X,X,X,X,X,X,X,X,X,X,X,X,X,X, X,X,X,X,X,X,X,X,X,X,X,X,X,X,X,X
#undef X
};

/* ---- one BAM file in memory ---- */
typedef struct {
  char *path;
  uint8_t *raw;
  size_t n_raw;
  bam_hdr_t hdr;
  size_t n_rec;
  size_t *off;        /* offset of each record's block_size field */
  int64_t *maxend;    /* running max of bam_endpos within the record's tid */
  size_t *tid_beg, *tid_end; /* record ranges per tid (sorted input) */
  int refcnt;
} bamfile_t;
struct htsFile { bamfile_t *f; };
struct hts_idx_t { bamfile_t *f; };
struct hts_itr_t { bamfile_t *f; size_t cur, stop; int tid; hts_pos_t beg, end; };

static pthread_mutex_t g_lock = PTHREAD_MUTEX_INITIALIZER;
static bamfile_t *g_files[64];
static int g_nfiles;

static int32_t rd_i32(const uint8_t *p) { int32_t v; memcpy(&v, p, 4); return v; }
static uint16_t rd_u16(const uint8_t *p) { uint16_t v; memcpy(&v, p, 2); return v; }

static void rec_view(const bamfile_t *f, size_t i, bam1_t *b) {
  const uint8_t *p = f->raw + f->off[i];
  int32_t bs = rd_i32(p);
  p += 4;
  bam1_core_t *c = &b->core;
  c->tid = rd_i32(p);
  c->pos = rd_i32(p + 4);
  c->l_qname = p[8];
  c->qual = p[9];
  c->bin = rd_u16(p + 10);
  c->n_cigar = rd_u16(p + 12);
  c->flag = rd_u16(p + 14);
  c->l_qseq = rd_i32(p + 16);
  c->mtid = rd_i32(p + 20);
  c->mpos = rd_i32(p + 24);
  c->isize = rd_i32(p + 28);
  c->l_extranul = 0;
  b->l_data = bs - 32;
  if ((uint32_t)b->l_data > b->m_data) {
    b->m_data = (uint32_t)b->l_data + 64;
    b->data = (uint8_t *)realloc(b->data, b->m_data);
  }
  memcpy(b->data, p + 32, (size_t)b->l_data);
}

static bamfile_t *bam_load(const char *fn) {
  gzFile g = gzopen(fn, "rb");
  if (!g) return NULL;
  gzbuffer(g, 1 << 20);
  bamfile_t *f = (bamfile_t *)calloc(1, sizeof *f);
  size_t cap = 1 << 24;
  f->raw = (uint8_t *)malloc(cap);
  for (;;) {
    if (f->n_raw + (1 << 22) > cap) { cap *= 2; f->raw = (uint8_t *)realloc(f->raw, cap); }
    int n = gzread(g, f->raw + f->n_raw, 1 << 22);
    if (n <= 0) break;
    f->n_raw += (size_t)n;
  }
  gzclose(g);
  if (f->n_raw < 12 || memcmp(f->raw, "BAM\1", 4) != 0) { free(f->raw); free(f); return NULL; }
  const uint8_t *p = f->raw + 4;
  int32_t l_text = rd_i32(p);
  p += 4;
  f->hdr.text = (char *)malloc((size_t)l_text + 1);
  memcpy(f->hdr.text, p, (size_t)l_text);
  f->hdr.text[l_text] = 0;
  f->hdr.l_text = (size_t)l_text;
  p += l_text;
  int32_t n_ref = rd_i32(p);
  p += 4;
  f->hdr.n_targets = n_ref;
  f->hdr.target_len = (uint32_t *)calloc((size_t)n_ref + 1, sizeof(uint32_t));
  f->hdr.target_name = (char **)calloc((size_t)n_ref + 1, sizeof(char *));
  for (int i = 0; i < n_ref; ++i) {
    int32_t l_name = rd_i32(p);
    p += 4;
    f->hdr.target_name[i] = strdup((const char *)p);
    p += l_name;
    f->hdr.target_len[i] = (uint32_t)rd_i32(p);
    p += 4;
  }
  size_t o = (size_t)(p - f->raw), mrec = 1 << 16;
  f->off = (size_t *)malloc(mrec * sizeof(size_t));
  while (o + 4 <= f->n_raw) {
    int32_t bs = rd_i32(f->raw + o);
    if (bs < 32 || o + 4 + (size_t)bs > f->n_raw) break;
    if (f->n_rec == mrec) { mrec *= 2; f->off = (size_t *)realloc(f->off, mrec * sizeof(size_t)); }
    f->off[f->n_rec++] = o;
    o += 4 + (size_t)bs;
  }
  f->maxend = (int64_t *)malloc((f->n_rec + 1) * sizeof(int64_t));
  f->tid_beg = (size_t *)calloc((size_t)n_ref + 1, sizeof(size_t));
  f->tid_end = (size_t *)calloc((size_t)n_ref + 1, sizeof(size_t));
  bam1_t *b = bam_init1();
  int cur_tid = -2;
  int64_t mx = 0;
  for (size_t i = 0; i < f->n_rec; ++i) {
    rec_view(f, i, b);
    if (b->core.tid != cur_tid) {
      cur_tid = b->core.tid;
      mx = 0;
      if (cur_tid >= 0 && cur_tid < n_ref) f->tid_beg[cur_tid] = i;
    }
    if (cur_tid >= 0 && cur_tid < n_ref) f->tid_end[cur_tid] = i + 1;
    int64_t e = bam_endpos(b);
    if (e > mx) mx = e;
    f->maxend[i] = mx;
  }
  bam_destroy1(b);
  f->path = strdup(fn);
  return f;
}

htsFile *hts_open(const char *fn, const char *mode) {
  (void)mode;
  pthread_mutex_lock(&g_lock);
  bamfile_t *f = NULL;
  for (int i = 0; i < g_nfiles; ++i)
    if (strcmp(g_files[i]->path, fn) == 0) f = g_files[i];
  if (!f) {
    f = bam_load(fn);
    if (f && g_nfiles < 64) g_files[g_nfiles++] = f;
  }
  if (f) f->refcnt++;
  pthread_mutex_unlock(&g_lock);
  if (!f) return NULL;
  htsFile *h = (htsFile *)calloc(1, sizeof *h);
  h->f = f;
  return h;
}
int hts_close(htsFile *fp) { free(fp); return 0; } /* files stay cached for the life of the process */

bam_hdr_t *sam_hdr_read(htsFile *fp) {
  bam_hdr_t *h = (bam_hdr_t *)calloc(1, sizeof *h), *s = &fp->f->hdr;
  h->n_targets = s->n_targets;
  h->target_len = (uint32_t *)calloc((size_t)s->n_targets + 1, sizeof(uint32_t));
  h->target_name = (char **)calloc((size_t)s->n_targets + 1, sizeof(char *));
  for (int i = 0; i < s->n_targets; ++i) {
    h->target_len[i] = s->target_len[i];
    h->target_name[i] = strdup(s->target_name[i]);
  }
  h->text = strdup(s->text);
  h->l_text = s->l_text;
  return h;
}
void bam_hdr_destroy(bam_hdr_t *h) {
  if (!h) return;
  for (int i = 0; i < h->n_targets; ++i) free(h->target_name[i]);
  free(h->target_name); free(h->target_len); free(h->text); free(h);
}
int bam_name2id(bam_hdr_t *h, const char *ref) {
  for (int i = 0; i < h->n_targets; ++i)
    if (strcmp(h->target_name[i], ref) == 0) return i;
  return -1;
}

hts_idx_t *sam_index_load(htsFile *fp, const char *fn) {
  /* like htslib: the .bai must exist next to the BAM; its content is not needed here (records are in memory) */
  char *p = (char *)malloc(strlen(fn) + 8);
  sprintf(p, "%s.bai", fn);
  FILE *t = fopen(p, "rb");
  if (!t) {
    size_t l = strlen(fn);
    if (l > 4 && strcmp(fn + l - 4, ".bam") == 0) { sprintf(p, "%.*s.bai", (int)(l - 4), fn); t = fopen(p, "rb"); }
  }
  free(p);
  if (!t) return NULL;
  fclose(t);
  hts_idx_t *x = (hts_idx_t *)calloc(1, sizeof *x);
  x->f = fp->f;
  return x;
}
void hts_idx_destroy(hts_idx_t *idx) { free(idx); }

hts_itr_t *sam_itr_queryi(const hts_idx_t *idx, int tid, hts_pos_t beg, hts_pos_t end) {
  hts_itr_t *it = (hts_itr_t *)calloc(1, sizeof *it);
  bamfile_t *f = idx->f;
  it->f = f; it->tid = tid; it->beg = beg < 0 ? 0 : beg; it->end = end;
  if (tid < 0 || tid >= f->hdr.n_targets) { it->cur = it->stop = 0; return it; }
  size_t lo = f->tid_beg[tid], hi = f->tid_end[tid];
  it->stop = hi;
  /* first record whose running max end exceeds beg: nothing before it can overlap */
  while (lo < hi) {
    size_t mid = (lo + hi) / 2;
    if (f->maxend[mid] > it->beg) hi = mid; else lo = mid + 1;
  }
  it->cur = lo;
  return it;
}
int sam_itr_next(htsFile *fp, hts_itr_t *it, bam1_t *b) {
  (void)fp;
  while (it->cur < it->stop) {
    rec_view(it->f, it->cur++, b);
    if (b->core.tid != it->tid || b->core.pos >= it->end) { it->cur = it->stop; return -1; }
    if (bam_endpos(b) > it->beg) return b->l_data + 32;
  }
  return -1;
}
void hts_itr_destroy(hts_itr_t *iter) { free(iter); }

bam1_t *bam_init1(void) { return (bam1_t *)calloc(1, sizeof(bam1_t)); }
void bam_destroy1(bam1_t *b) { if (b) { free(b->data); free(b); } }

hts_pos_t bam_cigar2rlen(int n_cigar, const uint32_t *cigar) {
  hts_pos_t l = 0;
  for (int k = 0; k < n_cigar; ++k)
    if (bam_cigar_type(bam_cigar_op(cigar[k])) & 2) l += bam_cigar_oplen(cigar[k]);
  return l;
}
hts_pos_t bam_endpos(const bam1_t *b) {
  hts_pos_t rlen = (!(b->core.flag & BAM_FUNMAP) && b->core.n_cigar > 0) ? bam_cigar2rlen((int)b->core.n_cigar, bam_get_cigar(b)) : 1;
  if (rlen == 0) rlen = 1;
  return b->core.pos + rlen;
}

static const uint8_t *aux_skip(const uint8_t *s, const uint8_t *end) {
  int t = *s++;
  switch (t) {
    case 'A': case 'c': case 'C': return s + 1;
    case 's': case 'S': return s + 2;
    case 'i': case 'I': case 'f': return s + 4;
    case 'd': return s + 8;
    case 'Z': case 'H':
      while (s < end && *s) ++s;
      return s + 1;
    case 'B': {
      int st = *s++;
      int32_t n = rd_i32(s);
      s += 4;
      int sz = (st == 'c' || st == 'C') ? 1 : (st == 's' || st == 'S') ? 2 : 4;
      return s + (size_t)n * (size_t)sz;
    }
    default: return end;
  }
}
uint8_t *bam_aux_get(const bam1_t *b, const char tag[2]) {
  const uint8_t *s = bam_get_aux(b), *end = b->data + b->l_data;
  while (s + 3 <= end) {
    if (s[0] == (uint8_t)tag[0] && s[1] == (uint8_t)tag[1]) return (uint8_t *)(s + 2);
    s = aux_skip(s + 2, end);
  }
  return NULL;
}
int64_t bam_aux2i(const uint8_t *s) {
  int t = *s++;
  switch (t) {
    case 'c': return (int8_t)*s;
    case 'C': return *s;
    case 's': { int16_t v; memcpy(&v, s, 2); return v; }
    case 'S': { uint16_t v; memcpy(&v, s, 2); return v; }
    case 'i': { int32_t v; memcpy(&v, s, 4); return v; }
    case 'I': { uint32_t v; memcpy(&v, s, 4); return v; }
    default: return 0;
  }
}

const char *hts_parse_reg(const char *s, int *beg, int *end) {
  const char *colon = strrchr(s, ':');
  *beg = 0; *end = INT_MAX;
  if (!colon) return s + strlen(s);
  /* numbers with thousands separators */
  char buf[64]; size_t k = 0;
  for (const char *p = colon + 1; *p && k + 1 < sizeof buf; ++p)
    if (*p != ',') buf[k++] = *p;
  buf[k] = 0;
  char *e;
  long long b = strtoll(buf, &e, 10);
  if (e == buf) return NULL;
  *beg = (int)(b - 1);
  if (*beg < 0) *beg = 0;
  if (*e == '-') {
    char *e2;
    long long x = strtoll(e + 1, &e2, 10);
    if (e2 != e + 1) *end = (int)x;
    e = e2;
  }
  if (*e) return NULL;
  if (*beg > *end) return NULL;
  return colon;
}

/* ---- FASTA ---- */
typedef struct { char *name; char *seq; int len; } faseq_t;
struct faidx_t { faseq_t *s; int n; char *path; int refcnt; };
static faidx_t *g_fa[16];
static int g_nfa;

faidx_t *fai_load(const char *fn) {
  pthread_mutex_lock(&g_lock);
  for (int i = 0; i < g_nfa; ++i)
    if (strcmp(g_fa[i]->path, fn) == 0) { g_fa[i]->refcnt++; pthread_mutex_unlock(&g_lock); return g_fa[i]; }
  gzFile g = gzopen(fn, "rb");
  if (!g) { pthread_mutex_unlock(&g_lock); return NULL; }
  gzbuffer(g, 1 << 20);
  faidx_t *fa = (faidx_t *)calloc(1, sizeof *fa);
  int cap = 0, scap = 0;
  char *line = (char *)malloc(1 << 16);
  faseq_t *cur = NULL;
  while (gzgets(g, line, 1 << 16)) {
    size_t l = strlen(line);
    while (l && (line[l - 1] == '\n' || line[l - 1] == '\r')) line[--l] = 0;
    if (line[0] == '>') {
      if (fa->n == cap) { cap = cap ? cap * 2 : 32; fa->s = (faseq_t *)realloc(fa->s, (size_t)cap * sizeof(faseq_t)); }
      cur = &fa->s[fa->n++];
      size_t k = 1;
      while (line[k] && !isspace((unsigned char)line[k])) ++k;
      cur->name = strndup(line + 1, k - 1);
      cur->seq = NULL; cur->len = 0; scap = 0;
    } else if (cur) {
      if (cur->len + (int)l + 1 > scap) { scap = (cur->len + (int)l + 1) * 2; cur->seq = (char *)realloc(cur->seq, (size_t)scap); }
      for (size_t k = 0; k < l; ++k)
        if (isgraph((unsigned char)line[k])) cur->seq[cur->len++] = line[k];
    }
  }
  free(line);
  gzclose(g);
  fa->path = strdup(fn);
  fa->refcnt = 1;
  if (g_nfa < 16) g_fa[g_nfa++] = fa;
  pthread_mutex_unlock(&g_lock);
  return fa;
}
void fai_destroy(faidx_t *fai) { (void)fai; } /* cached for the life of the process */
int faidx_seq_len(const faidx_t *fai, const char *seq) {
  for (int i = 0; i < fai->n; ++i)
    if (strcmp(fai->s[i].name, seq) == 0) return fai->s[i].len;
  return -1;
}
char *faidx_fetch_seq(const faidx_t *fai, const char *c_name, int p_beg_i, int p_end_i, int *len) {
  for (int i = 0; i < fai->n; ++i)
    if (strcmp(fai->s[i].name, c_name) == 0) {
      const faseq_t *s = &fai->s[i];
      if (p_end_i < p_beg_i) p_beg_i = p_end_i;
      if (p_beg_i < 0) p_beg_i = 0; else if (s->len <= p_beg_i) p_beg_i = s->len - 1;
      if (p_end_i < 0) p_end_i = 0; else if (s->len <= p_end_i) p_end_i = s->len - 1;
      int l = p_end_i - p_beg_i + 1;
      char *r = (char *)malloc((size_t)l + 1);
      memcpy(r, s->seq + p_beg_i, (size_t)l);
      r[l] = 0;
      *len = l;
      return r;
    }
  *len = -2;
  return NULL;
}
