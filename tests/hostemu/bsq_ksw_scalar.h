// TEST-ONLY: the scalar form of the banded extension and the scalar execution policy of the region builder, used by
// the host emulation (hostemu.cpp) to run the per-task device code of biscuit_b200/csrc in one thread.  Not part of
// the product: libbsq.so runs the warp-cooperative kernel (biscuit_b200/csrc/bsq_ksw_warp.cuh).
//
// Banded affine-gap extension with the reference's exact tie rules (ksw_extend2, lib/aln/ksw.c:380-479).  Row i walks
// the target, column j the query.  Query and target are read through accessors so callers can extend leftwards
// (reversed prefixes) or straight from the 2-bit packed reference without materialising windows.
#pragma once
#include "../../biscuit_b200/csrc/bsq_region.h"

template <typename QGet, typename TGet>
BSQ_HD bsq_ext_result_t bsq_ksw_extend(int qlen, QGet qget, int tlen, TGet tget, const int8_t *mat, int o_del, int e_del,
                                       int o_ins, int e_ins, int w, int end_bonus, int zdrop, int h0,
                                       bsq_ksw_scratch_t &scr) {
  bsq_eh_t *eh = scr.eh;
  BSQ_CTR(BSQ_CTR_KSW, 1);
  const int oe_del = o_del + e_del, oe_ins = o_ins + e_ins;
  int i, j;
  for (j = 0; j <= qlen; ++j) eh[j].h = eh[j].e = 0;
  // first row: gap-extension tail of h0 (ksw.c:395-397)
  eh[0].h = h0;
  if (qlen >= 1) eh[1].h = h0 > oe_ins ? h0 - oe_ins : 0;
  for (j = 2; j <= qlen && eh[j - 1].h > e_ins; ++j) eh[j].h = eh[j - 1].h - e_ins;
  // cap the band by the longest gap that can still score (ksw.c:399-407)
  int mx = 0;
  for (i = 0; i < 25; ++i) mx = mx > mat[i] ? mx : mat[i];
  int max_ins = (int)((double)(qlen * mx + end_bonus - o_ins) / e_ins + 1.);
  max_ins = max_ins > 1 ? max_ins : 1;
  w = w < max_ins ? w : max_ins;
  int max_del = (int)((double)(qlen * mx + end_bonus - o_del) / e_del + 1.);
  max_del = max_del > 1 ? max_del : 1;
  w = w < max_del ? w : max_del;
  int max = h0, max_i = -1, max_j = -1, max_ie = -1, gscore = -1, max_off = 0;
  int beg = 0, end = qlen;
  for (i = 0; i < tlen; ++i) {
    int f = 0, h1, m = 0, mj = -1;
    const int8_t *row = mat + 5 * tget(i);
    if (beg < i - w) beg = i - w;
    if (end > i + w + 1) end = i + w + 1;
    if (end > qlen) end = qlen;
    BSQ_CTR(BSQ_CTR_CELLS, end > beg ? end - beg : 0);
    if (beg == 0) {
      h1 = h0 - (o_del + e_del * (i + 1));
      if (h1 < 0) h1 = 0;
    } else h1 = 0;
    for (j = beg; j < end; ++j) {
      // eh[j] = { H(i-1,j-1), E(i,j) }, f = F(i,j), h1 = H(i,j-1)
      int M = eh[j].h, e = eh[j].e, h, t;
      eh[j].h = h1;
      M = M ? M + row[qget(j)] : 0;  // a zero cell cannot be restarted (ksw.c:433)
      h = M > e ? M : e;
      h = h > f ? h : f;
      h1 = h;
      mj = m > h ? mj : j;  // last column wins ties within a row
      m = m > h ? m : h;
      t = M - oe_del; t = t > 0 ? t : 0;
      e -= e_del; e = e > t ? e : t;
      eh[j].e = e;
      t = M - oe_ins; t = t > 0 ? t : 0;
      f -= e_ins; f = f > t ? f : t;
    }
    eh[end].h = h1; eh[end].e = 0;
    if (j == qlen) {  // reached the end of the query: later rows win ties
      max_ie = gscore > h1 ? max_ie : i;
      gscore = gscore > h1 ? gscore : h1;
    }
    if (m == 0) break;
    if (m > max) {  // first row wins ties
      max = m; max_i = i; max_j = mj;
      max_off = max_off > bsq_iabs(mj - i) ? max_off : bsq_iabs(mj - i);
    } else if (zdrop > 0) {
      if (i - max_i > mj - max_j) {
        if (max - m - ((i - max_i) - (mj - max_j)) * e_del > zdrop) break;
      } else {
        if (max - m - ((mj - max_j) - (i - max_i)) * e_ins > zdrop) break;
      }
    }
    // shrink the band to the non-zero extent of this row (ksw.c:466-469)
    for (j = beg; j < end && eh[j].h == 0 && eh[j].e == 0; ++j) {}
    beg = j;
    for (j = end; j >= beg && eh[j].h == 0 && eh[j].e == 0; --j) {}
    end = j + 2 < qlen ? j + 2 : qlen;
  }
  bsq_ext_result_t r;
  r.score = max; r.qle = max_j + 1; r.tle = max_i + 1; r.gtle = max_ie + 1; r.gscore = gscore; r.max_off = max_off;
  return r;
}

struct bsq_scalar_policy {
  BSQ_HD static bool leader() { return true; }
  BSQ_HD static int max_gap(const bsq_devopt_t &opt, int qlen) { return bsq_cal_max_gap(opt, qlen); }
  BSQ_HD static void sync() {}
  // asymmetric_flt_seed (memchain.c:138-149): ref T under read C, or ref A under read G, inside the seed
  BSQ_HD static bool asym_conflict(const bsq_devidx_t &ix, const bsq_seed_t &s, const uint8_t *query) {
    for (int i = 0; i < s.len; ++i) {
      const int r = bsq_ref_base(ix, s.rbeg + i), qv = query[s.qbeg + i];
      if ((r == 3 && qv == 1) || (r == 0 && qv == 2)) return true;
    }
    return false;
  }
  BSQ_HD static bsq_ext_result_t extend(int qlen, bsq_qacc_t qget, int tlen, bsq_tacc_t tget, const int8_t *mat, int o_del, int e_del, int o_ins,
                                        int e_ins, int w, int end_bonus, int zdrop, int h0, bsq_ksw_scratch_t *scr) {
    return bsq_ksw_extend(qlen, qget, tlen, tget, mat, o_del, e_del, o_ins, e_ins, w, end_bonus, zdrop, h0, *scr);
  }
};
