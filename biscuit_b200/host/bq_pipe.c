/* bq_pipe.c -- batch pipeline of `biscuit align` over one or several GPUs:
 *
 *   stage A (thread)  next batch of reads from the source + host preparation (clipping, task rows in page-locked memory)
 *   stage B (one thread per lane; a lane = one aligner context = one GPU with its index replica; batches go to the lanes
 *                     round-robin) GPU: H2D, phase-1 kernels, D2H  (bq_batch_run).  `biscuit align -G 0-7` runs eight lanes:
 *                     reads shard by batch with no exchange between GPUs; the batch is the unit
 *                     because mem_pestat is per batch (bwamem.c:464-467), so the output equals a one-GPU run
 *   stage C (caller)  host phase 2 on opt->n_threads threads, batches in sequence order
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
static int q_ready(q1_t *q) { /* is there something to get right now? */
  pthread_mutex_lock(&q->mu);
  const int r = q->full;
  pthread_mutex_unlock(&q->mu);
  return r;
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
  bsq_aligner *al[BQ_MAX_LANES];
  bsq_dp *dp[BQ_MAX_LANES]; /* batched phase-2 DP of the lane's device (or NULL) */
  int n_al;
  volatile int abort_; /* set on the first failure: the source stops reading, the lanes stop computing */
  bq_source_fn src;
  void *src_ctx;
  q1_t qa[BQ_MAX_LANES], qb[BQ_MAX_LANES], qc;
  int64_t n_processed;
  double t_src, t_prep, t_gpu, t_sink;
  bq_sink_fn sink;
  void *sink_ctx;
} pipe_ctx_t;

static double pnow(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + ts.tv_nsec * 1e-9; }

static void *stage_a(void *arg) {
  pipe_ctx_t *p = arg;
  for (int seq = 0;; ++seq) {
    q1_t *qa = &p->qa[seq % p->n_al];
    int n = 0, rc = 0;
    double t0 = pnow();
    bq_read_t *seqs = p->abort_ ? 0 : p->src(p->src_ctx, &n); /* after a failure nothing more is read */
    p->t_src += pnow() - t0; t0 = pnow();
    if (!seqs || n <= 0) { /* end marker to every GPU lane, in sequence order */
      free(seqs);
      for (int k = 0; k < p->n_al; ++k) q_put(&p->qa[(seq + k) % p->n_al], 0, 0, 0, 0, 1);
      return 0;
    }
    bq_batch_t *b = bq_batch_prep(p->opt, p->n_processed, n, seqs, &rc);
    p->t_prep += pnow() - t0;
    p->n_processed += n;
    if (!b) { /* preparation failed: report it on this lane, end the others */
      p->abort_ = 1;
      q_put(qa, 0, seqs, n, rc, 1);
      for (int k = 1; k < p->n_al; ++k) q_put(&p->qa[(seq + k) % p->n_al], 0, 0, 0, 0, 1);
      return 0;
    }
    q_put(qa, b, seqs, n, rc, 0);
  }
}

typedef struct { pipe_ctx_t *p; int lane; double t_gpu; } lane_t;

static void *stage_b(void *arg) {
  lane_t *L = arg;
  pipe_ctx_t *p = L->p;
  int failed = 0;
  for (;;) {
    bq_batch_t *b; bq_read_t *seqs; int n, rc, end;
    q_get(&p->qa[L->lane], &b, &seqs, &n, &rc, &end);
    if (b && !failed && !p->abort_) {
      const double t0 = pnow();
      rc = bq_batch_run(p->al[L->lane], p->dp[L->lane], b);
      L->t_gpu += pnow() - t0;
      if (rc) { failed = rc; p->abort_ = 1; }
    } else if (b) rc = failed ? failed : BSQ_EINVAL;  /* after a failure (here or on another lane) batches are only drained */
    q_put(&p->qb[L->lane], b, seqs, n, rc, end);
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

int bq_pipeline_run(const bq_opt_t *opt, const bq_ref_t *ref, bsq_aligner *const *als, bsq_dp *const *dps, int n_al, bq_source_fn src, void *src_ctx,
                    bq_sink_fn sink, void *sink_ctx, const bq_pestat_t *pes0, const char *rg_id) {
  pipe_ctx_t p;
  memset(&p, 0, sizeof p);
  if (n_al < 1 || n_al > BQ_MAX_LANES) return BSQ_EINVAL;
  p.opt = opt; p.n_al = n_al; p.src = src; p.src_ctx = src_ctx;
  for (int k = 0; k < n_al; ++k) { p.al[k] = als[k]; p.dp[k] = dps ? dps[k] : 0; }
  p.sink = sink; p.sink_ctx = sink_ctx;
  for (int k = 0; k < n_al; ++k) { q_init(&p.qa[k]); q_init(&p.qb[k]); }
  q_init(&p.qc);
  pthread_t ta, tb[BQ_MAX_LANES], td;
  lane_t lanes[BQ_MAX_LANES];
  for (int k = 0; k < n_al; ++k) { lanes[k].p = &p; lanes[k].lane = k; lanes[k].t_gpu = 0; }
  pthread_create(&ta, 0, stage_a, &p);
  for (int k = 0; k < p.n_al; ++k) pthread_create(&tb[k], 0, stage_b, &lanes[k]);
  pthread_create(&td, 0, stage_d, &p);
  int ret = 0;
  double t_wait = 0, t_fin = 0;
  const int adaptive_wait = getenv("BQ_ADAPTIVE_WAIT") && atoi(getenv("BQ_ADAPTIVE_WAIT")) != 0;
  /* Stage C around the asynchronous CIGAR kernel of the batch's DP context.  When the next batch is already waiting
   * (the host is the slower side), the first half of phase 2 of batch k+1 (merge .. primary marking, CIGAR jobs
   * submitted) runs before the second half of batch k (pairing, SAM text), so the kernel of k+1 overlaps the formatting
   * of k; `held` = batch k between its halves.  When nothing is waiting (the GPU is the slower side) batch k is
   * finished at once: this thread would idle anyway, and the output leaves one batch earlier. */
  struct { bq_batch_t *b; bq_read_t *seqs; int n; } held = {0, 0, 0};
#define FLUSH_HELD(END) do { if (held.b) { bq_batch_finish_b(opt, ref, held.b, rg_id); q_put(&p.qc, 0, held.seqs, held.n, 0, END); held.b = 0; } } while (0)
  for (int seq = 0;; ++seq) { /* batches come back in sequence order: lane seq % n_al */
    bq_batch_t *b; bq_read_t *seqs; int n, rc, end;
    double t0 = pnow();
    q_get(&p.qb[seq % p.n_al], &b, &seqs, &n, &rc, &end);
    /* which side is the slower one?  If this thread had to wait for the GPU lane, the lane's waits should be as short
     * as they can (spinning); if the batch was already there, phase 2 is the slower side and its workers need the core
     * a spinning lane thread would hold (the lane then sleeps on an event).  Two batches of hysteresis. */
    if (adaptive_wait) { /* BQ_ADAPTIVE_WAIT=1: measured once, on different boxes, without a clear gain (profiles/README.md): off by default */
      static int host_bound = 0;
      const double waited = pnow() - t0;
      if (waited < 0.002) { if (host_bound < 2 && ++host_bound == 2) bsq_set_wait_mode(1); }
      else if (host_bound > 0 && --host_bound == 0) bsq_set_wait_mode(0);
    }
    t_wait += pnow() - t0; t0 = pnow();
    int rcw;
    if (held.b && (rcw = bq_batch_finish_wait(held.b))) { /* its CIGARs did not come back: the batch fails */
      if (!ret) ret = rcw;
      p.abort_ = 1;
      bq_batch_abandon(held.b);
      q_put(&p.qc, 0, held.seqs, held.n, rcw, 0);
      held.b = 0;
    }
    if (b && rc == 0 && (rc = bq_batch_finish_a(opt, ref, b, pes0)) == 0) {
      FLUSH_HELD(0);
      held.b = b; held.seqs = seqs; held.n = n;
      if (!q_ready(&p.qb[(seq + 1) % p.n_al])) { /* nothing to overlap with: finish this batch now */
        if ((rcw = bq_batch_finish_wait(held.b))) {
          if (!ret) ret = rcw;
          p.abort_ = 1;
          bq_batch_abandon(held.b);
          q_put(&p.qc, 0, held.seqs, held.n, rcw, 0);
          held.b = 0;
        } else FLUSH_HELD(0);
      }
      t_fin += pnow() - t0;
    } else {
      FLUSH_HELD(0);
      t_fin += pnow() - t0;
      if (rc && !ret) ret = rc;
      if (rc) p.abort_ = 1;
      if (b) { if (rc) bq_batch_abandon(b); else bq_batch_discard(b); }
      q_put(&p.qc, 0, seqs, n, rc ? rc : -1, end); /* failed batch / end marker: the sink only frees */
    }
    if (end) { /* drain the end markers of the other lanes */
      for (int k = 1; k < p.n_al; ++k) {
        bq_batch_t *b2; bq_read_t *s2; int n2, rc2, e2;
        q_get(&p.qb[(seq + k) % p.n_al], &b2, &s2, &n2, &rc2, &e2);
      }
      break;
    }
  }
#undef FLUSH_HELD
  pthread_join(ta, 0);
  for (int k = 0; k < p.n_al; ++k) pthread_join(tb[k], 0);
  pthread_join(td, 0);
  for (int k = 0; k < p.n_al; ++k) p.t_gpu += lanes[k].t_gpu;
  if (getenv("BQ_TIMING"))
    fprintf(stderr, "[bq_pipeline] source %.3f prep %.3f | gpu %.3f | wait %.3f phase2 %.3f sink %.3f s\n", p.t_src, p.t_prep, p.t_gpu, t_wait, t_fin,
            p.t_sink);
  return ret;
}
