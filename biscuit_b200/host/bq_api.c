/* bq_api.c -- the batch boundary B1 (mem_process_seqs, lib/aln/bwamem.c:432-476) as a flat C entry point of
 * libbiscuit_host.so, for callers that already hold reads in memory (bench.py, tests): nt4 read rows in, SAM
 * text out.  GPU phase 1 through libbsq.so, phase 2 on host threads. */
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include "bq.h"
#include <stdio.h>

typedef struct {
  bq_opt_t opt;
  bq_ref_t ref;
  bsq_aligner *al, *al2;
  bsq_dp *dp, *dp2; /* batched phase-2 DP of the same device */
} bq_session;

bq_session *bq_session_create(bsq_index *dx, const uint8_t *pac, int64_t l_pac, int n_seqs, const char *const *names, const int64_t *offs,
                              const int32_t *lens, int n_threads, int paired) {
  bq_session *s = calloc(1, sizeof *s);
  bq_opt_init(&s->opt);
  s->opt.flag |= BQ_F_NO_MULTI;
  if (paired) s->opt.flag |= BQ_F_PE;
  s->opt.n_threads = n_threads > 0 ? n_threads : 1;
  s->ref.l_pac = l_pac; s->ref.n_seqs = n_seqs; s->ref.seed = 11;
  s->ref.anns = calloc((size_t)n_seqs, sizeof(bq_ann_t));
  for (int i = 0; i < n_seqs; ++i) {
    s->ref.anns[i].name = strdup(names[i]); s->ref.anns[i].anno = strdup("");
    s->ref.anns[i].offset = offs[i]; s->ref.anns[i].len = lens[i];
  }
  s->ref.pac = (uint8_t *)pac; /* borrowed */
  bsq_opt d;
  bq_opt_to_dev(&s->opt, &d);
  if (bsq_aligner_create(dx, &d, &s->al)) { free(s->ref.anns); free(s); return 0; }
  if (bsq_dp_create(dx, &d, &s->dp)) { bsq_aligner_destroy(s->al); free(s->ref.anns); free(s); return 0; }
  if (!getenv("BQ_TWO_CONTEXTS") || bsq_aligner_create(dx, &d, &s->al2) || bsq_dp_create(dx, &d, &s->dp2)) s->al2 = 0; /* measured: no gain from a second context */
  return s;
}

void bq_session_destroy(bq_session *s) {
  if (!s) return;
  bsq_dp_destroy(s->dp);
  if (s->dp2) bsq_dp_destroy(s->dp2);
  bsq_aligner_destroy(s->al);
  if (s->al2) bsq_aligner_destroy(s->al2);
  for (int i = 0; i < s->ref.n_seqs; ++i) { free(s->ref.anns[i].name); free(s->ref.anns[i].anno); }
  free(s->ref.anns);
  free(s);
}

/* Align n reads (interleaved pairs when the session is paired).  Returns the total number of SAM bytes, or a
 * negative BSQ_E* code; the SAM text is copied to sam_out when it is non-NULL and fits in cap. */
int64_t bq_session_align(bq_session *s, int64_t n_processed, int n, const uint8_t *seqs, int stride, const int32_t *lens, const uint8_t *quals,
                         char *sam_out, int64_t cap) {
  bq_read_t *rd = calloc((size_t)n + 1, sizeof(bq_read_t));
  char nm[64];
  for (int i = 0; i < n; ++i) {
    rd[i].l_seq = rd[i].l_seq0 = lens[i];
    rd[i].seq = rd[i].seq0 = malloc((size_t)lens[i] + 1);
    memcpy(rd[i].seq, seqs + (size_t)i * stride, (size_t)lens[i]);
    rd[i].qual = malloc((size_t)lens[i] + 1);
    if (quals) memcpy(rd[i].qual, quals + (size_t)i * stride, (size_t)lens[i]); else memset(rd[i].qual, 'I', (size_t)lens[i]);
    rd[i].qual[lens[i]] = 0;
    snprintf(nm, sizeof nm, "r%lld", (long long)((n_processed + i) >> 1));
    rd[i].name = strdup(nm);
    rd[i].id = i;
  }
  int rc = bq_process_seqs(&s->opt, s->al, s->dp, &s->ref, n_processed, n, rd, 0, "");
  int64_t tot = 0;
  for (int i = 0; i < n; ++i) {
    int64_t l = rd[i].sam ? (int64_t)strlen(rd[i].sam) : 0;
    if (sam_out && tot + l < cap) memcpy(sam_out + tot, rd[i].sam, (size_t)l);
    tot += l;
  }
  if (sam_out && tot < cap) sam_out[tot] = 0;
  bq_reads_free(rd, n);
  return rc ? rc : tot;
}

/* n_batches batches of the same n reads through the three-stage pipeline of the CLI (bq_pipe.c).
 * Returns the SAM bytes of the last batch (negative BSQ_E* on error). */
typedef struct { int n_batches, b, n, stride, n_threads; const uint8_t *seqs, *quals; const int32_t *lens; int64_t sam_bytes; } stream_t;

/* what bq_read_batch produces for a FASTQ batch: one slab with name, nt4 sequence and quality of every read */
typedef struct { bq_read_t *rd; char *slab; int64_t n_processed; const uint8_t *seqs, *quals; const int32_t *lens; int stride; size_t rec; } mk_t;
static void make_read(void *ctx, long i) {
  mk_t *m = ctx;
  bq_read_t *r = &m->rd[i];
  char *p = m->slab + (size_t)i * m->rec;
  r->name = p; sprintf(p, "r%lld", (long long)((m->n_processed + i) >> 1)); p += 24;
  r->l_seq = r->l_seq0 = m->lens[i];
  r->seq = r->seq0 = (uint8_t *)p; memcpy(p, m->seqs + (size_t)i * m->stride, (size_t)m->lens[i]); p += m->lens[i] + 1;
  r->qual = p;
  if (m->quals) memcpy(p, m->quals + (size_t)i * m->stride, (size_t)m->lens[i]); else memset(p, 'I', (size_t)m->lens[i]);
  p[m->lens[i]] = 0;
  r->id = (int)i; r->in_slab = 1;
}
static bq_read_t *make_reads(int n_threads, int64_t n_processed, int n, const uint8_t *seqs, int stride, const int32_t *lens, const uint8_t *quals) {
  bq_read_t *rd = calloc((size_t)n + 1, sizeof(bq_read_t));
  int max_len = 0;
  for (int i = 0; i < n; ++i) if (lens[i] > max_len) max_len = lens[i];
  mk_t m = {rd, 0, n_processed, seqs, quals, lens, stride, 24 + 2 * ((size_t)max_len + 1)}; /* fixed record size: the reads fill in parallel */
  size_t cap_ = 0;
  m.slab = bq_big_alloc((size_t)n * m.rec + 16, &cap_);
  bq_parallel_for(n_threads, n, make_read, &m);
  if (n > 0) rd[0].slab = m.slab; else bq_big_free(m.slab);
  return rd;
}

static bq_read_t *stream_source(void *ctx, int *n) {
  stream_t *st = ctx;
  if (st->b >= st->n_batches) { *n = 0; return 0; }
  bq_read_t *rd = make_reads(st->n_threads, (int64_t)st->b * st->n, st->n, st->seqs, st->stride, st->lens, st->quals);
  st->b++;
  *n = st->n;
  return rd;
}

static void stream_sink(void *ctx, bq_read_t *rd, int n) {
  stream_t *st = ctx;
  const int ok = n >= 0;
  if (n < 0) n = -n;
  int64_t tot = 0;
  for (int i = 0; i < n; ++i) tot += rd[i].sam ? (int64_t)(rd[i].sam_len ? rd[i].sam_len : strlen(rd[i].sam)) : 0;
  bq_reads_free(rd, n);
  if (ok) st->sam_bytes = tot;
}

int64_t bq_session_align_stream(bq_session *s, int n_batches, int n, const uint8_t *seqs, int stride, const int32_t *lens, const uint8_t *quals) {
  stream_t st;
  memset(&st, 0, sizeof st);
  st.n_batches = n_batches; st.n = n; st.stride = stride; st.seqs = seqs; st.quals = quals; st.lens = lens; st.n_threads = s->opt.n_threads;
  bsq_aligner *als[2] = {s->al, s->al2};
  bsq_dp *dps[2] = {s->dp, s->dp2};
  const int rc = bq_pipeline_run(&s->opt, &s->ref, als, dps, s->al2 ? 2 : 1, stream_source, &st, stream_sink, &st, 0, "");
  return rc ? rc : st.sam_bytes;
}

/* counters of the batched phase-2 DP since the library was loaded (bq_dp_stats) */
void bq_session_dp_stats(int64_t out[6]) { bq_dp_stats(out); }
