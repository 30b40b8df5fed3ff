/* TEST-ONLY host emulation of the bsq_dp_* entry points of include/bsq.h (the batched phase-2 dynamic programming).
 * The CUDA kernels (biscuit_b200/csrc/bsq_dp.cu) have no per-task host form; on a machine without a GPU this stand-in
 * answers the same jobs with the host's scalar routines (biscuit_b200/host/bq_core.c: bq_gen_cigar, bq_local_align), so
 * that the job construction, result plumbing and SAM text of the host side can be checked in the CPU suite.  The
 * kernels themselves are compared with the reference on the GPU box (tests/test_dp.py -m gpu, tests/test_align_sam.py). */
#include <stdlib.h>
#include <string.h>
#include "../../biscuit_b200/host/bq.h"

const uint8_t *hostemu_index_pac(const bsq_index *ix, int64_t *l_pac);

struct bsq_dp {
  const bsq_index *idx;
  bsq_opt opt;
  uint8_t *seqs; int32_t *lens; int64_t n_rows; int32_t stride;
  uint32_t *blob; size_t blob_words, blob_cap;       /* written by submit */
  uint32_t *ready; size_t ready_words;               /* what the last wait handed out (valid until the next wait) */
  int64_t c_n, m_n, counters[8];
  int c_pending, m_pending;
};

int bsq_dp_create(const bsq_index *idx, const bsq_opt *opt, bsq_dp **out) {
  if (!idx || !opt || !out) return BSQ_EINVAL;
  bsq_dp *dp = calloc(1, sizeof *dp);
  dp->idx = idx; dp->opt = *opt;
  *out = dp;
  return 0;
}
void bsq_dp_destroy(bsq_dp *dp) { if (dp) { free(dp->seqs); free(dp->lens); free(dp->blob); free(dp->ready); free(dp); } }
int bsq_dp_set_reads(bsq_dp *dp, int64_t n_rows, const uint8_t *seqs, int32_t stride, const int32_t *lens) {
  if (!dp || n_rows < 0 || stride <= 0) return BSQ_EINVAL;
  free(dp->seqs); free(dp->lens);
  dp->seqs = malloc((size_t)n_rows * stride + 1); dp->lens = malloc((size_t)n_rows * 4 + 4);
  memcpy(dp->seqs, seqs, (size_t)n_rows * stride); memcpy(dp->lens, lens, (size_t)n_rows * 4);
  dp->n_rows = n_rows; dp->stride = stride;
  return 0;
}
int bsq_dp_sync(bsq_dp *dp) { return dp ? 0 : BSQ_EINVAL; }

static uint32_t *blob_room(bsq_dp *dp, size_t words) {
  if (dp->blob_words + words > dp->blob_cap) {
    dp->blob_cap = (dp->blob_words + words) * 2 + 1024;
    dp->blob = realloc(dp->blob, dp->blob_cap * 4);
  }
  return dp->blob + dp->blob_words;
}

int bsq_dp_cigar_submit(bsq_dp *dp, int64_t n_jobs, const bsq_cigar_job *jobs, bsq_cigar_res *res) {
  if (!dp || n_jobs < 0 || dp->c_pending) return BSQ_EINVAL;
  int64_t l_pac = 0;
  const uint8_t *pac = hostemu_index_pac(dp->idx, &l_pac);
  const bsq_opt *o = &dp->opt;
  dp->blob_words = 0; dp->c_n = n_jobs; dp->counters[0] = n_jobs; dp->counters[1] = 0;
  for (int64_t j = 0; j < n_jobs; ++j) {
    const bsq_cigar_job *jb = &jobs[j];
    bsq_cigar_res *r = &res[j];
    memset(r, 0, sizeof *r);
    r->NM = -1;
    if (jb->row < 0 || jb->row >= dp->n_rows) return BSQ_EINVAL;
    const int lq = jb->qe - jb->qb;
    if (lq <= 0 || lq > BSQ_MAX_READ_LEN || jb->re - jb->rb > 1024) { r->n_cigar = lq <= 0 ? 0 : -1; continue; }
    uint8_t q[BSQ_MAX_READ_LEN + 1];
    const uint8_t *row = dp->seqs + (size_t)jb->row * dp->stride;
    for (int i = 0; i < lq; ++i) q[i] = row[jb->qb + i] < 5 ? row[jb->qb + i] : 4;
    uint32_t *cigar = 0;
    int n_cigar = 0, score = 0, last_sc = -(1 << 30), w = jb->w, NM = -1, bss_u = 0;
    uint32_t ZC = 0, ZR = 0;
    for (int i = 0; i < 3; ++i, w <<= 1, last_sc = score) { /* mem_alnreg.c:60-70 */
      free(cigar);
      w = w < o->w << 2 ? w : o->w << 2;
      cigar = bq_gen_cigar(jb->parent ? o->ctmat : o->gamat, o->o_del, o->e_del, o->o_ins, o->e_ins, w, l_pac, pac, lq, q, jb->rb, jb->re, &score,
                           &n_cigar, &NM, &ZC, &ZR, &bss_u, jb->parent);
      if (score == last_sc) break;
      if (w == o->w << 2) break;
      if (score >= jb->truesc - o->a) break;
    }
    if (!cigar || n_cigar <= 0) { free(cigar); continue; }
    if (lq == jb->re - jb->rb && jb->w == 0) dp->counters[1]++;
    const char *md = (const char *)(cigar + n_cigar);
    const size_t l_md = strlen(md);
    int first = 0, last = n_cigar;
    if ((cigar[0] & 0xf) == 2) { r->lead_del = (int32_t)(cigar[0] >> 4); first = 1; }
    else if ((cigar[n_cigar - 1] & 0xf) == 2) last = n_cigar - 1;
    const int n_final = (jb->clip5 ? 1 : 0) + (last - first) + (jb->clip3 ? 1 : 0);
    const size_t words = (size_t)n_final + ((l_md + 1 + 3) >> 2);
    uint32_t *dst = blob_room(dp, words);
    r->off = (uint32_t)dp->blob_words;
    dp->blob_words += words;
    int k = 0;
    if (jb->clip5) dst[k++] = (uint32_t)jb->clip5 << 4 | 3u;
    for (int c = first; c < last; ++c) dst[k++] = cigar[c];
    if (jb->clip3) dst[k++] = (uint32_t)jb->clip3 << 4 | 3u;
    memcpy(dst + k, md, l_md + 1);
    r->n_cigar = n_final; r->NM = NM; r->ZC = (int32_t)ZC; r->ZR = (int32_t)ZR; r->score = score; r->bss_u = bss_u;
    free(cigar);
  }
  dp->c_pending = 1;
  return 0;
}
int bsq_dp_cigar_wait(bsq_dp *dp, const uint32_t **blob, int64_t *blob_words) {
  if (!dp || !dp->c_pending) return BSQ_EINVAL;
  dp->c_pending = 0;
  free(dp->ready);
  dp->ready = dp->blob; dp->ready_words = dp->blob_words;
  dp->blob = 0; dp->blob_words = dp->blob_cap = 0;
  if (blob) *blob = dp->ready_words ? dp->ready : 0;
  if (blob_words) *blob_words = (int64_t)dp->ready_words;
  return 0;
}

int bsq_dp_matesw_submit(bsq_dp *dp, int64_t n_jobs, const bsq_matesw_job *jobs, bsq_matesw_res *res) {
  if (!dp || n_jobs < 0 || dp->m_pending) return BSQ_EINVAL;
  int64_t l_pac = 0;
  const uint8_t *pac = hostemu_index_pac(dp->idx, &l_pac);
  const bsq_opt *o = &dp->opt;
  dp->m_n = n_jobs; dp->counters[4] = n_jobs;
  for (int64_t j = 0; j < n_jobs; ++j) {
    const bsq_matesw_job *jb = &jobs[j];
    if (jb->row < 0 || jb->row >= dp->n_rows) return BSQ_EINVAL;
    const int l_ms = dp->lens[jb->row];
    const uint8_t *ms = dp->seqs + (size_t)jb->row * dp->stride;
    uint8_t rev[BSQ_MAX_READ_LEN + 1];
    for (int i = 0; i < l_ms; ++i) rev[l_ms - 1 - i] = ms[i] < 4 ? 3 - ms[i] : 4;
    int64_t len = 0;
    uint8_t *rseq = bq_get_seq(l_pac, pac, jb->rb, jb->re, &len);
    bq_swr_t a = bq_local_align(l_ms, rev, (int)len, rseq, jb->use_ga ? o->gamat : o->ctmat, o->o_del, o->e_del, o->o_ins, o->e_ins, jb->xtra);
    free(rseq);
    res[j].score = a.score; res[j].te = a.te; res[j].qe = a.qe; res[j].score2 = a.score2; res[j].te2 = a.te2; res[j].tb = a.tb; res[j].qb = a.qb;
    res[j].pad_ = 0;
  }
  dp->m_pending = 1;
  return 0;
}
int bsq_dp_matesw_wait(bsq_dp *dp) {
  if (!dp || !dp->m_pending) return BSQ_EINVAL;
  dp->m_pending = 0;
  return 0;
}
int bsq_dp_counters(const bsq_dp *dp, int64_t *c, int n) {
  if (!dp || !c) return BSQ_EINVAL;
  for (int i = 0; i < n && i < 8; ++i) c[i] = dp->counters[i];
  return 0;
}
