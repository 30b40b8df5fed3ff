/* bq_core.c -- options, sorting, packed-reference access, host DP (global + local), CIGAR/MD generation. */
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include "bq.h"

int bq_verbose = 3;

void bq_fatal(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  fputs("[biscuit] ", stderr);
  vfprintf(stderr, fmt, ap);
  va_end(ap);
  fputc('\n', stderr);
  exit(1);
}

/* ---------------- options ---------------- */

void bq_fill_scmat(int a, int b, int8_t mat[25]) { /* plain matrix, bwa.c:146-155 */
  for (int r = 0; r < 5; ++r)
    for (int c = 0; c < 5; ++c) mat[r * 5 + c] = (r == 4 || c == 4) ? -1 : (r == c ? a : -b);
}

/* asymmetric matrices, bwa.c:158-182: ct -> read T on reference C scores a; else read A on reference G */
void bq_fill_scmat_bis(int a, int b, int ct, int8_t mat[25]) {
  bq_fill_scmat(a, b, mat);
  if (ct) mat[1 * 5 + 3] = a; else mat[2 * 5 + 0] = a;
}

void bq_opt_init(bq_opt_t *o) { /* mem_opt_init, bwamem.c:77-128 */
  memset(o, 0, sizeof *o);
  o->a = 1; o->b = 2; o->o_del = o->o_ins = 6; o->e_del = o->e_ins = 1;
  o->w = 100; o->T = 30; o->zdrop = 100; o->pen_unpaired = 17; o->pen_clip5 = o->pen_clip3 = 10;
  o->max_mem_intv = 20; o->min_seed_len = 19; o->split_width = 10; o->max_occ = 500; o->max_chain_gap = 10000;
  o->max_ins = 5000; o->mask_level = 0.50f; o->drop_ratio = 0.50f; o->XA_drop_ratio = 0.80f; o->split_factor = 1.5f;
  o->chunk_size = 10000000; o->n_threads = 1; o->max_XA_hits = 5; o->max_XA_hits_alt = 5; o->max_matesw = 50;
  o->mask_level_redun = 0.95f; o->min_chain_weight = 0; o->max_chain_extend = 1u << 30;
  o->mapQ_coef_len = 50; o->mapQ_coef_fac = (int)log(o->mapQ_coef_len); /* int field: log(50) truncates to 3 (bwamem.h:81) */
  bq_fill_scmat(o->a, o->b, o->mat);
  bq_fill_scmat_bis(o->a, o->b, 1, o->ctmat);
  bq_fill_scmat_bis(o->a, o->b, 0, o->gamat);
}

void bq_opt_to_dev(const bq_opt_t *o, bsq_opt *d) {
  memset(d, 0, sizeof *d);
  d->a = o->a; d->b = o->b; d->o_del = o->o_del; d->e_del = o->e_del; d->o_ins = o->o_ins; d->e_ins = o->e_ins;
  d->pen_clip5 = o->pen_clip5; d->pen_clip3 = o->pen_clip3; d->w = o->w; d->zdrop = o->zdrop;
  d->min_seed_len = o->min_seed_len; d->split_width = o->split_width; d->max_occ = (int32_t)o->max_occ;
  d->max_chain_gap = o->max_chain_gap; d->min_chain_weight = o->min_chain_weight; d->max_chain_extend = (int32_t)o->max_chain_extend;
  d->max_mem_intv = (int32_t)o->max_mem_intv;
  d->split_len = (int)(o->min_seed_len * o->split_factor + .499); /* memchain.c:55 */
  d->self_ovlp = (o->flag & BQ_F_SELF_OVLP) != 0; d->bsstrand = o->bsstrand;
  d->mask_level = o->mask_level; d->drop_ratio = o->drop_ratio;
  memcpy(d->ctmat, o->ctmat, 25); memcpy(d->gamat, o->gamat, 25);
}

uint64_t bq_hash64(uint64_t key) { /* utils.h:107 */
  key += ~(key << 32); key ^= (key >> 22); key += ~(key << 13); key ^= (key >> 8);
  key += (key << 3); key ^= (key >> 15); key += ~(key << 27); key ^= (key >> 31);
  return key;
}

/* ---------------- string buffer ---------------- */

/* bq_str_reserve / bq_kput*: static inline in bq.h */

/* ---------------- introsort with klib's exact comparison/swap sequence (ksort.h:150-233) ---------------- */

#define SWP(x, y) do { memcpy(tmp, (x), sz); memcpy((x), (y), sz); memcpy((y), tmp, sz); } while (0)

static void ins_sort(char *s, char *t, size_t sz, int (*lt)(const void *, const void *), char *tmp) {
  for (char *i = s + sz; i < t; i += sz)
    for (char *j = i; j > s && lt(j, j - sz); j -= sz) SWP(j, j - sz);
}

static void comb_sort(char *a, size_t n, size_t sz, int (*lt)(const void *, const void *), char *tmp) {
  const double shrink = 1.2473309501039786540366528676643;
  size_t gap = n;
  int swapped;
  do {
    if (gap > 2) { gap = (size_t)(gap / shrink); if (gap == 9 || gap == 10) gap = 11; }
    swapped = 0;
    for (char *i = a; i < a + (n - gap) * sz; i += sz) {
      char *j = i + gap * sz;
      if (lt(j, i)) { SWP(i, j); swapped = 1; }
    }
  } while (swapped || gap > 2);
  if (gap != 1) ins_sort(a, a + n * sz, sz, lt, tmp);
}

void bq_introsort(void *base, size_t n, size_t sz, int (*lt)(const void *, const void *)) {
  char *a = base, tmp[256], rp[256];
  struct { char *l, *r; int d; } st[80], *top = st;
  if (n < 1 || sz > sizeof tmp) { if (sz > sizeof tmp) bq_fatal("bq_introsort: element too large"); return; }
  if (n == 2) { if (lt(a + sz, a)) SWP(a, a + sz); return; }
  int d = 2;
  while ((1ul << d) < n) ++d;
  char *s = a, *t = a + (n - 1) * sz;
  d <<= 1;
  for (;;) {
    if (s < t) {
      if (--d == 0) { comb_sort(s, (size_t)(t - s) / sz + 1, sz, lt, tmp); t = s; continue; }
      char *i = s, *j = t, *k = i + (((size_t)(j - i) / sz) >> 1) * sz + sz;
      if (lt(k, i)) { if (lt(k, j)) k = j; }
      else k = lt(j, i) ? i : j;
      memcpy(rp, k, sz);
      if (k != t) SWP(k, t);
      for (;;) {
        do i += sz; while (lt(i, rp));
        do j -= sz; while (i <= j && lt(rp, j));
        if (j <= i) break;
        SWP(i, j);
      }
      SWP(i, t);
      if (i - s > t - i) {
        if ((size_t)(i - s) > 16 * sz) { top->l = s; top->r = i - sz; top->d = d; ++top; }
        s = (size_t)(t - i) > 16 * sz ? i + sz : t;
      } else {
        if ((size_t)(t - i) > 16 * sz) { top->l = i + sz; top->r = t; top->d = d; ++top; }
        t = (size_t)(i - s) > 16 * sz ? i - sz : s;
      }
    } else {
      if (top == st) { ins_sort(a, a + n * sz, sz, lt, tmp); return; }
      --top; s = top->l; t = top->r; d = top->d;
    }
  }
}

/* ---------------- packed reference ---------------- */

#define PAC(pac, l) ((pac)[(l) >> 2] >> ((~(l) & 3) << 1) & 3)

int bq_pos2rid(const bq_ref_t *r, int64_t pos_f) { /* bntseq.c:356-369 */
  if (pos_f >= r->l_pac) return -1;
  int left = 0, mid = 0, right = r->n_seqs;
  while (left < right) {
    mid = (left + right) >> 1;
    if (pos_f >= r->anns[mid].offset) {
      if (mid == r->n_seqs - 1) break;
      if (pos_f < r->anns[mid + 1].offset) break;
      left = mid + 1;
    } else right = mid;
  }
  return mid;
}

int64_t bq_depos(const bq_ref_t *r, int64_t pos, int *is_rev) { return (*is_rev = (pos >= r->l_pac)) ? (r->l_pac << 1) - 1 - pos : pos; }

uint8_t *bq_get_seq(int64_t l_pac, const uint8_t *pac, int64_t beg, int64_t end, int64_t *len) { /* bntseq.c:402-422 */
  if (end < beg) { int64_t t = beg; beg = end; end = t; }
  if (end > l_pac << 1) end = l_pac << 1;
  if (beg < 0) beg = 0;
  if (!(beg >= l_pac || end <= l_pac)) { *len = 0; return 0; } /* bridging the forward-reverse boundary */
  uint8_t *seq = malloc((size_t)(end - beg) + 1);
  int64_t l = 0;
  *len = end - beg;
  if (beg >= l_pac) {
    const int64_t beg_f = (l_pac << 1) - 1 - end, end_f = (l_pac << 1) - 1 - beg;
    for (int64_t k = end_f; k > beg_f; --k) seq[l++] = 3 - PAC(pac, k);
  } else
    for (int64_t k = beg; k < end; ++k) seq[l++] = PAC(pac, k);
  return seq;
}

uint8_t *bq_fetch_seq(const bq_ref_t *r, int64_t *beg, int64_t mid, int64_t *end, int *rid) { /* bntseq.c:428-452 */
  int is_rev;
  int64_t len;
  if (*end < *beg) { int64_t t = *beg; *beg = *end; *end = t; }
  *rid = bq_pos2rid(r, bq_depos(r, mid, &is_rev));
  int64_t far_beg = r->anns[*rid].offset, far_end = far_beg + r->anns[*rid].len;
  if (is_rev) { int64_t t = far_beg; far_beg = (r->l_pac << 1) - far_end; far_end = (r->l_pac << 1) - t; }
  if (*beg < far_beg) *beg = far_beg;
  if (*end > far_end) *end = far_end;
  uint8_t *seq = bq_get_seq(r->l_pac, r->pac, *beg, *end, &len);
  if (seq == 0 || *end - *beg != len) bq_fatal("bq_fetch_seq: begin=%ld mid=%ld end=%ld len=%ld rid=%d", (long)*beg, (long)mid, (long)*end, (long)len, *rid);
  return seq;
}

/* ---------------- banded global alignment with traceback (ksw_global2, ksw.c:504-606) ---------------- */

#define NEG_INF (-0x40000000)

static uint32_t *cig_push(int *n, int *m, uint32_t *cig, int op, int len) {
  if (*n == 0 || (uint32_t)op != (cig[*n - 1] & 0xf)) {
    if (*n == *m) { *m = *m ? *m << 1 : 4; cig = realloc(cig, (size_t)*m << 2); }
    cig[(*n)++] = (uint32_t)len << 4 | (uint32_t)op;
  } else cig[*n - 1] += (uint32_t)len << 4;
  return cig;
}

int bq_global_align(int qlen, const uint8_t *query, int tlen, const uint8_t *target, const int8_t *mat, int o_del, int e_del, int o_ins,
                    int e_ins, int w, int *n_cigar_, uint32_t **cigar_) {
  const int oe_del = o_del + e_del, oe_ins = o_ins + e_ins;
  const int want = n_cigar_ && cigar_;
  const int n_col = qlen < 2 * w + 1 ? qlen : 2 * w + 1;
  uint8_t *z = want ? malloc((size_t)n_col * tlen + 1) : 0;
  int32_t *H = malloc((size_t)(qlen + 1) * 4), *E = malloc((size_t)(qlen + 1) * 4);
  int i, j;
  if (n_cigar_) *n_cigar_ = 0;
  H[0] = 0; E[0] = NEG_INF;
  for (j = 1; j <= qlen && j <= w; ++j) { H[j] = -(o_ins + e_ins * j); E[j] = NEG_INF; }
  for (; j <= qlen; ++j) H[j] = E[j] = NEG_INF;
  for (i = 0; i < tlen; ++i) {
    int32_t f = NEG_INF, h1;
    const int8_t *row = mat + 5 * target[i];
    const int beg = i > w ? i - w : 0, end = i + w + 1 < qlen ? i + w + 1 : qlen;
    uint8_t *zi = z ? z + (size_t)i * n_col : 0;
    h1 = beg == 0 ? -(o_del + e_del * (i + 1)) : NEG_INF;
    for (j = beg; j < end; ++j) {
      /* H[j] holds H(i-1,j-1), E[j] holds E(i,j); direction bits: h (2), e-extension (1), f-extension (1) */
      int32_t m = H[j] + row[query[j]], e = E[j], h, t;
      uint8_t d;
      H[j] = h1;
      d = m >= e ? 0 : 1; h = m >= e ? m : e;
      d = h >= f ? d : 2; h = h >= f ? h : f;
      h1 = h;
      t = m - oe_del; e -= e_del;
      d |= e > t ? 1 << 2 : 0; e = e > t ? e : t;
      E[j] = e;
      t = m - oe_ins; f -= e_ins;
      d |= f > t ? 2 << 4 : 0; f = f > t ? f : t;
      if (zi) zi[j - beg] = d;
    }
    H[end] = h1; E[end] = NEG_INF;
  }
  const int score = H[qlen];
  if (want) {
    int n = 0, m = 0, which = 0, k;
    uint32_t *cig = 0;
    i = tlen - 1; k = (i + w + 1 < qlen ? i + w + 1 : qlen) - 1;
    while (i >= 0 && k >= 0) {
      which = z[(size_t)i * n_col + (k - (i > w ? i - w : 0))] >> (which << 1) & 3;
      if (which == 0) { cig = cig_push(&n, &m, cig, 0, 1); --i; --k; }
      else if (which == 1) { cig = cig_push(&n, &m, cig, 2, 1); --i; }
      else { cig = cig_push(&n, &m, cig, 1, 1); --k; }
    }
    if (i >= 0) cig = cig_push(&n, &m, cig, 2, i + 1);
    if (k >= 0) cig = cig_push(&n, &m, cig, 1, k + 1);
    for (i = 0; i < n >> 1; ++i) { uint32_t t = cig[i]; cig[i] = cig[n - 1 - i]; cig[n - 1 - i] = t; }
    *n_cigar_ = n; *cigar_ = cig;
  }
  free(H); free(E); free(z);
  return score;
}

/* ---------------- local alignment used by mate rescue ----------------
 * ksw_align2 -> ksw_u8 / ksw_i16 (ksw.c:111-365) are 16 x u8 / 8 x i16 striped SSE2 kernels whose results
 * depend on the striping (E is taken before the lazy-F correction, the row maximum before it too) and on
 * unsigned saturation.  They are re-enacted lane by lane in scalar code: P lanes, segment length slen,
 * query position of (segment j, lane l) = j + l*slen.   */

typedef struct { int P, slen, qlen, shift, mdiff, max, is8; int *qp; } swq_t;

static swq_t *swq_init(int size, int qlen, const uint8_t *query, const int8_t *mat) {
  swq_t *q = calloc(1, sizeof *q);
  q->is8 = size == 1; q->P = q->is8 ? 16 : 8; q->slen = (qlen + q->P - 1) / q->P; q->qlen = qlen;
  int mn = 127, mx = 0;
  for (int a = 0; a < 25; ++a) { if (mat[a] < mn) mn = mat[a]; if (mat[a] > mx) mx = mat[a]; }
  q->max = mx; q->shift = (256 - mn) & 0xff; q->mdiff = (mx + q->shift) & 0xff;
  q->qp = malloc(sizeof(int) * 5 * (size_t)q->slen * q->P + 16);
  for (int a = 0; a < 5; ++a)
    for (int j = 0; j < q->slen; ++j)
      for (int l = 0; l < q->P; ++l) {
        int k = j + l * q->slen, v = k >= qlen ? 0 : mat[a * 5 + query[k]];
        q->qp[((size_t)a * q->slen + j) * q->P + l] = q->is8 ? ((v + q->shift) & 0xff) : v;
      }
  return q;
}

static inline int sat_sub_u(int a, int b) { return a > b ? a - b : 0; }

static bq_swr_t sw_striped(const swq_t *q, int tlen, const uint8_t *target, int o_del, int e_del, int o_ins, int e_ins, int xtra) {
  const int P = q->P, slen = q->slen, oe_del = o_del + e_del, oe_ins = o_ins + e_ins;
  const int cap = q->is8 ? 255 : 32767;
  bq_swr_t r = {0, -1, -1, -1, -1, -1, -1};
  const int minsc = (xtra & BQ_XSUBO) ? xtra & 0xffff : 0x10000, endsc = (xtra & BQ_XSTOP) ? xtra & 0xffff : 0x10000;
  size_t nv = (size_t)slen * P;
  int *H0 = calloc(nv, 4), *H1 = calloc(nv, 4), *E = calloc(nv, 4), *Hmax = calloc(nv, 4);
  int *h = malloc(4 * P), *f = malloc(4 * P), *mxv = malloc(4 * P);
  uint64_t *b = 0; int n_b = 0, m_b = 0, gmax = 0, te = -1, i, j, l;
  for (i = 0; i < tlen; ++i) {
    const int *S = q->qp + (size_t)target[i] * slen * P;
    for (l = 0; l < P; ++l) { f[l] = 0; mxv[l] = 0; }
    /* h = H0[slen-1] shifted by one lane */
    h[0] = 0;
    for (l = 1; l < P; ++l) h[l] = H0[(size_t)(slen - 1) * P + l - 1];
    for (j = 0; j < slen; ++j) {
      int *e = E + (size_t)j * P, *h1 = H1 + (size_t)j * P;
      const int *s = S + (size_t)j * P, *h0 = H0 + (size_t)j * P;
      for (l = 0; l < P; ++l) {
        int v;
        if (q->is8) { v = h[l] + s[l]; if (v > 255) v = 255; v = sat_sub_u(v, q->shift); }
        else { v = h[l] + s[l]; if (v > 32767) v = 32767; if (v < -32768) v = -32768; }
        if (v < e[l]) v = e[l];
        if (v < f[l]) v = f[l];
        if (mxv[l] < v) mxv[l] = v;
        h1[l] = v;
        int ee = sat_sub_u(e[l], e_del), t = sat_sub_u(v, oe_del);
        e[l] = ee > t ? ee : t;
        int ff = sat_sub_u(f[l], e_ins);
        t = sat_sub_u(v, oe_ins);
        f[l] = ff > t ? ff : t;
        h[l] = h0[l];
      }
    }
    /* lazy-F loop */
    int done = 0;
    for (int k = 0; k < 16 && !done; ++k) {
      for (l = P - 1; l > 0; --l) f[l] = f[l - 1];
      f[0] = 0;
      for (j = 0; j < slen; ++j) {
        int *h1 = H1 + (size_t)j * P, all = 1;
        for (l = 0; l < P; ++l) {
          int v = h1[l] > f[l] ? h1[l] : f[l];
          h1[l] = v;
          v = sat_sub_u(v, oe_ins);
          f[l] = sat_sub_u(f[l], e_ins);
          if (q->is8) { if (sat_sub_u(f[l], v) != 0) all = 0; }
          else { if (f[l] > v) all = 0; }
        }
        if (all) { done = 1; break; }
      }
    }
    int imax = 0;
    for (l = 0; l < P; ++l) if (mxv[l] > imax) imax = mxv[l];
    if (imax >= minsc) {
      if (n_b == 0 || (int32_t)b[n_b - 1] + 1 != i) {
        if (n_b == m_b) { m_b = m_b ? m_b << 1 : 8; b = realloc(b, 8 * (size_t)m_b); }
        b[n_b++] = (uint64_t)imax << 32 | (uint32_t)i;
      } else if ((int)(b[n_b - 1] >> 32) < imax) b[n_b - 1] = (uint64_t)imax << 32 | (uint32_t)i;
    }
    if (imax > gmax) {
      gmax = imax; te = i;
      memcpy(Hmax, H1, nv * 4);
      if (q->is8) { if (gmax + q->shift >= 255 || gmax >= endsc) break; }
      else if (gmax >= endsc) break;
    }
    { int *t = H1; H1 = H0; H0 = t; }
  }
  (void)cap;
  r.score = q->is8 ? (gmax + q->shift < 255 ? gmax : 255) : gmax;
  r.te = te;
  if (!q->is8 || r.score != 255) {
    int mx = -1, tmp, low, high;
    const int n = slen * P;
    if (!q->is8) r.qe = -1;
    for (i = 0; i < n; ++i) { /* memory order of the vector array: segment i / P, lane i % P */
      int v = Hmax[i];
      if (v > mx) { mx = v; r.qe = i / P + i % P * slen; }
      else if (v == mx && (tmp = i / P + i % P * slen) < r.qe) r.qe = tmp;
    }
    if (b) {
      i = (r.score + q->max - 1) / q->max;
      low = te - i; high = te + i;
      for (i = 0; i < n_b; ++i) {
        int e = (int32_t)b[i];
        if ((e < low || e > high) && (int)(b[i] >> 32) > r.score2) { r.score2 = (int)(b[i] >> 32); r.te2 = e; }
      }
    }
  }
  free(b); free(H0); free(H1); free(E); free(Hmax); free(h); free(f); free(mxv);
  return r;
}

static void rev_bytes(int l, uint8_t *s) { for (int i = 0; i < l >> 1; ++i) { uint8_t t = s[i]; s[i] = s[l - 1 - i]; s[l - 1 - i] = t; } }

bq_swr_t bq_local_align(int qlen, uint8_t *query, int tlen, uint8_t *target, const int8_t *mat, int o_del, int e_del, int o_ins, int e_ins,
                        int xtra) { /* ksw_align2, ksw.c:343-365 */
  const int size = (xtra & BQ_XBYTE) ? 1 : 2;
  swq_t *q = swq_init(size, qlen, query, mat);
  bq_swr_t r = sw_striped(q, tlen, target, o_del, e_del, o_ins, e_ins, xtra), rr;
  free(q->qp); free(q);
  if ((xtra & BQ_XSTART) == 0 || ((xtra & BQ_XSUBO) && r.score < (xtra & 0xffff))) return r;
  rev_bytes(r.qe + 1, query); rev_bytes(r.te + 1, target);
  q = swq_init(size, r.qe + 1, query, mat);
  rr = sw_striped(q, tlen, target, o_del, e_del, o_ins, e_ins, BQ_XSTOP | r.score);
  rev_bytes(r.qe + 1, query); rev_bytes(r.te + 1, target);
  free(q->qp); free(q);
  if (r.score == rr.score) { r.tb = r.te - rr.te; r.qb = r.qe - rr.qe; }
  return r;
}

/* ---------------- CIGAR + MD + NM/ZC/ZR (bis_bwa_gen_cigar2, bwa.c:290-428) ---------------- */

/* forward bases [beg, end) of the 2-bit packed reference, four per table lookup (same values as PAC()) */
static void pac_decode(const uint8_t *pac, int64_t beg, int64_t end, uint8_t *out) {
  static uint32_t lut[256];
  static int lut_ready;
  if (!__atomic_load_n(&lut_ready, __ATOMIC_ACQUIRE)) { /* idempotent: racing threads write the same values */
    for (int v = 0; v < 256; ++v) lut[v] = (uint32_t)(v >> 6 & 3) | (uint32_t)(v >> 4 & 3) << 8 | (uint32_t)(v >> 2 & 3) << 16 | (uint32_t)(v & 3) << 24;
    __atomic_store_n(&lut_ready, 1, __ATOMIC_RELEASE);
  }
  int64_t k = beg;
  for (; k < end && (k & 3); ++k) *out++ = PAC(pac, k);
  for (; k + 4 <= end; k += 4, out += 4) { const uint32_t w = lut[pac[k >> 2]]; memcpy(out, &w, 4); } /* little endian: first base in the low byte */
  for (; k < end; ++k) *out++ = PAC(pac, k);
}

uint32_t *bq_gen_cigar(const int8_t mat[25], int o_del, int e_del, int o_ins, int e_ins, int w_, int64_t l_pac, const uint8_t *pac,
                       int l_query, uint8_t *query, int64_t rb, int64_t re, int *score, int *n_cigar, int *NM, uint32_t *ZC, uint32_t *ZR,
                       int *bss_u, uint8_t parent) {
  uint32_t *cigar = 0;
  int64_t rlen;
  int i;
  if (n_cigar) *n_cigar = 0;
  if (NM) *NM = -1;
  if (l_query <= 0 || rb >= re || (rb < l_pac && re > l_pac)) return 0;
  /* this runs once or twice per read: the reference window, the single-M CIGAR and the MD text live on the stack
   * whenever they fit, and the returned block is allocated once */
  uint8_t rbuf[1024], *rseq;
  uint32_t one_m = 0;
  int cigar_local = 0;
  if (re - rb <= (int64_t)sizeof rbuf && rb >= 0 && re <= l_pac << 1) { /* bq_get_seq without the allocation (bntseq.c:402-422) */
    rseq = rbuf; rlen = re - rb;
    if (rb >= l_pac) { /* reverse strand: forward bases (beg_f, end_f] read backwards and complemented */
      const int64_t beg_f = (l_pac << 1) - 1 - re, end_f = (l_pac << 1) - 1 - rb;
      pac_decode(pac, beg_f + 1, end_f + 1, rseq);
      for (int64_t a = 0, b = rlen - 1; a < b; ++a, --b) { const uint8_t t = rseq[a]; rseq[a] = 3 - rseq[b]; rseq[b] = 3 - t; }
      if (rlen & 1) rseq[rlen >> 1] = 3 - rseq[rlen >> 1];
    } else pac_decode(pac, rb, re, rseq);
  } else rseq = bq_get_seq(l_pac, pac, rb, re, &rlen);
  if (re - rb != rlen) { if (rseq != rbuf) free(rseq); return 0; }
  if (rb >= l_pac) { rev_bytes(l_query, query); rev_bytes((int)rlen, rseq); } /* left-align indels on the forward strand */
  if (l_query == re - rb && w_ == 0) { /* ungapped: one M, no DP (bwa.c:314-322) */
    if (n_cigar) { one_m = (uint32_t)l_query << 4; cigar = &one_m; cigar_local = 1; *n_cigar = 1; }
    for (i = 0, *score = 0; i < l_query; ++i) *score += mat[rseq[i] * 5 + query[i]];
  } else {
    int max_ins = (int)((double)(((l_query + 1) >> 1) * mat[0] - o_ins) / e_ins + 1.);
    int max_del = (int)((double)(((l_query + 1) >> 1) * mat[0] - o_del) / e_del + 1.);
    int max_gap = max_ins > max_del ? max_ins : max_del;
    max_gap = max_gap > 1 ? max_gap : 1;
    int w = (int)((max_gap + llabs(rlen - l_query) + 1) >> 1);
    w = w < w_ ? w : w_;
    int min_w = (int)llabs(rlen - l_query) + 3;
    w = w > min_w ? w : min_w;
    *score = bq_global_align(l_query, query, (int)rlen, rseq, mat, o_del, e_del, o_ins, e_ins, w, n_cigar, &cigar);
  }
  if (NM && n_cigar) {
    int k, x = 0, y = 0, u = 0, n_mm = 0, n_gap = 0, n_conv_ct = 0, n_ret_c = 0, n_conv_ga = 0, n_ret_g = 0;
    char mdbuf[2048];
    bq_str_t md = {0, 0, 0};
    /* a mismatch costs the digits of the run before it (at most run + 1 characters) and a letter, a deletion '^' and
     * its bases: fewer than 3 characters per reference base and CIGAR operation */
    const int md_on_stack = 3 * (size_t)(rlen + *n_cigar) + 32 < sizeof mdbuf;
    if (md_on_stack) { md.s = mdbuf; md.m = sizeof mdbuf; mdbuf[0] = 0; }
    const char *int2base = rb < l_pac ? "ACGTN" : "TGCAN";
    for (k = 0; k < *n_cigar; ++k) {
      const int op = cigar[k] & 0xf, len = (int)(cigar[k] >> 4);
      if (op == 0) {
        for (i = 0; i < len; ++i) {
          const unsigned q_ = query[x + i], r_ = rseq[y + i];
          if (q_ == r_) { if (q_ == 1) ++n_ret_c; if (q_ == 2) ++n_ret_g; ++u; }
          else {
            bq_kputw(&md, u); bq_kputc(&md, int2base[r_]); u = 0;
            if (parent && q_ == 3 && r_ == 1) ++n_conv_ct;
            else if (!parent && q_ == 0 && r_ == 2) ++n_conv_ga;
            else ++n_mm;
          }
        }
        x += len; y += len;
      } else if (op == 2) {
        if (k > 0 && k < *n_cigar - 1) {
          bq_kputw(&md, u); bq_kputc(&md, '^');
          for (i = 0; i < len; ++i) bq_kputc(&md, int2base[rseq[y + i]]);
          u = 0; n_gap += len;
        }
        y += len;
      } else if (op == 1) { x += len; n_gap += len; }
    }
    bq_kputw(&md, u);
    /* MD string stored right behind the CIGAR words; 8 spare bytes for the two clipping operations set_sam may add */
    if (cigar_local) { cigar = malloc((size_t)*n_cigar * 4 + md.l + 1 + 8); cigar[0] = one_m; cigar_local = 0; }
    else cigar = realloc(cigar, (size_t)*n_cigar * 4 + md.l + 1 + 8);
    memcpy((char *)(cigar + *n_cigar), md.s, md.l + 1);
    if (!md_on_stack) free(md.s);
    *NM = n_mm + n_gap;
    *ZC = parent ? (uint32_t)n_conv_ct : (uint32_t)n_conv_ga;
    *ZR = parent ? (uint32_t)n_ret_c : (uint32_t)n_ret_g;
    *bss_u = (n_conv_ct == 0 && n_conv_ga == 0) ? 1 : 0;
  }
  if (rb >= l_pac) rev_bytes(l_query, query);
  if (rseq != rbuf) free(rseq);
  if (cigar_local) { cigar = malloc(4); cigar[0] = one_m; } /* caller did not ask for NM / MD */
  return cigar;
}
