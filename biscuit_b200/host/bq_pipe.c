/* bq_pipe.c -- three-stage batch pipeline of `biscuit align` on one GPU:
 *
 *   stage A (thread)  next batch of reads from the source + host preparation (clipping, task rows in page-locked memory)
 *   stage B (thread)  GPU: H2D, phase-1 kernels, D2H                                              (bq_batch_run)
 *   stage C (caller)  host phase 2 on opt->n_threads threads
 *   stage D (thread)  the sink (SAM output, freeing the reads), batches in order
 *
 * The reference runs the same shape with kt_pipeline (read / mem_process_seqs / write, lib/aln/align.c:70-167,577);
 * here the compute step is split at the device boundary so that the GPU never waits for host work of its own batch.
 * Batches retire strictly in order; at most one batch sits in each queue.
 */
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "bq.h"

typedef struct {
  pthread_mutex_t mu;
  pthread_cond_t cv;
  bq_batch_t *b;
  bq_read_t *seqs;
  int n, full, rc, end;
} q1_t;

static void q_init(q1_t *q) { memset(q, 0, sizeof *q); pthread_mutex_init(&q->mu, 0); pthread_cond_init(&q->cv, 0); }
static void q_put(q1_t *q, bq_batch_t *b, bq_read_t *seqs, int n, int rc, int end) {
  pthread_mutex_lock(&q->mu);
  while (q->full) pthread_cond_wait(&q->cv, &q->mu);
  q->b = b; q->seqs = seqs; q->n = n; q->rc = rc; q->end = end; q->full = 1;
  pthread_cond_broadcast(&q->cv);
  pthread_mutex_unlock(&q->mu);
}
static void q_get(q1_t *q, bq_batch_t **b, bq_read_t **seqs, int *n, int *rc, int *end) {
  pthread_mutex_lock(&q->mu);
  while (!q->full) pthread_cond_wait(&q->cv, &q->mu);
  *b = q->b; *seqs = q->seqs; *n = q->n; *rc = q->rc; *end = q->end;
  q->full = 0;
  pthread_cond_broadcast(&q->cv);
  pthread_mutex_unlock(&q->mu);
}

typedef struct {
  const bq_opt_t *opt;
  bsq_aligner *al;
  bq_source_fn src;
  void *src_ctx;
  q1_t qa, qb, qc;
  int64_t n_processed;
  double t_src, t_prep, t_gpu, t_sink;
  bq_sink_fn sink;
  void *sink_ctx;
} pipe_ctx_t;

static double pnow(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + ts.tv_nsec * 1e-9; }

static void *stage_a(void *arg) {
  pipe_ctx_t *p = arg;
  for (;;) {
    int n = 0, rc = 0;
    double t0 = pnow();
    bq_read_t *seqs = p->src(p->src_ctx, &n);
    p->t_src += pnow() - t0; t0 = pnow();
    if (!seqs || n <= 0) { free(seqs); q_put(&p->qa, 0, 0, 0, 0, 1); return 0; }
    bq_batch_t *b = bq_batch_prep(p->opt, p->n_processed, n, seqs, &rc);
    p->t_prep += pnow() - t0;
    p->n_processed += n;
    q_put(&p->qa, b, seqs, n, rc, b == 0);
    if (!b) return 0;
  }
}

static void *stage_b(void *arg) {
  pipe_ctx_t *p = arg;
  int failed = 0;
  for (;;) {
    bq_batch_t *b; bq_read_t *seqs; int n, rc, end;
    q_get(&p->qa, &b, &seqs, &n, &rc, &end);
    if (b && !failed) {
      const double t0 = pnow();
      rc = bq_batch_run(p->al, b);
      p->t_gpu += pnow() - t0;
      if (rc) failed = rc;
    } else if (b) rc = failed;  /* after a failure the remaining batches are only drained */
    q_put(&p->qb, b, seqs, n, rc, end);
    if (end) return 0;
  }
}

/* stage D: the sink (SAM output, freeing the reads) runs beside phase 2 of the next batch */
static void *stage_d(void *arg) {
  pipe_ctx_t *p = arg;
  for (;;) {
    bq_batch_t *b; bq_read_t *seqs; int n, rc, end;
    q_get(&p->qc, &b, &seqs, &n, &rc, &end);
    if (seqs) {
      const double t0 = pnow();
      p->sink(p->sink_ctx, seqs, rc ? -n : n);
      p->t_sink += pnow() - t0;
    }
    if (end) return 0;
  }
}

int bq_pipeline_run(const bq_opt_t *opt, const bq_ref_t *ref, bsq_aligner *al, bq_source_fn src, void *src_ctx, bq_sink_fn sink, void *sink_ctx,
                    const bq_pestat_t *pes0, const char *rg_id) {
  pipe_ctx_t p;
  memset(&p, 0, sizeof p);
  p.opt = opt; p.al = al; p.src = src; p.src_ctx = src_ctx;
  p.sink = sink; p.sink_ctx = sink_ctx;
  q_init(&p.qa); q_init(&p.qb); q_init(&p.qc);
  pthread_t ta, tb, td;
  pthread_create(&ta, 0, stage_a, &p);
  pthread_create(&tb, 0, stage_b, &p);
  pthread_create(&td, 0, stage_d, &p);
  int ret = 0;
  double t_wait = 0, t_fin = 0;
  for (;;) {
    bq_batch_t *b; bq_read_t *seqs; int n, rc, end;
    double t0 = pnow();
    q_get(&p.qb, &b, &seqs, &n, &rc, &end);
    t_wait += pnow() - t0; t0 = pnow();
    if (b && rc == 0) {
      bq_batch_finish(opt, ref, b, pes0, rg_id);
      t_fin += pnow() - t0;
      q_put(&p.qc, 0, seqs, n, 0, end);
    } else {
      if (rc && !ret) ret = rc;
      if (b) bq_batch_discard(b);
      q_put(&p.qc, 0, seqs, n, rc ? rc : -1, end); /* failed batch: the sink only frees */
    }
    if (end) break;
  }
  pthread_join(ta, 0); pthread_join(tb, 0); pthread_join(td, 0);
  if (getenv("BQ_TIMING"))
    fprintf(stderr, "[bq_pipeline] source %.3f prep %.3f | gpu %.3f | wait %.3f phase2 %.3f sink %.3f s\n", p.t_src, p.t_prep, p.t_gpu, t_wait, t_fin,
            p.t_sink);
  return ret;
}
