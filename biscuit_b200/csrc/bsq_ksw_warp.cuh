// Warp-cooperative banded extension: ksw_extend2 (lib/aln/ksw.c:380-479) with one DP row spread over
// the 32 lanes of a warp.  Device only.
//
// The reference's band is adaptive: after every row it is trimmed to the non-zero extent of that row
// (ksw.c:466-469), and the trimming is not value-neutral, so rows must be finished one at a time
//  -- an anti-diagonal wavefront inside one extension would read cells the
// reference never computes.  Rows are therefore processed synchronously:
//   * lane L owns the C = ceil((qlen+1)/32) consecutive query columns [L*C, L*C+C); their eh[] state
//     (H of the previous row shifted by one, E of this row) lives in registers, including stale cells
//     outside the current band exactly as the reference's eh[] array keeps them;
//   * the horizontal gap state F(i,j+1) = max(F(i,j) - e_ins, max(M(i,j) - oe_ins, 0)) is a max-plus
//     prefix scan over the columns: A_k = t_k + k*e_ins, F_j = max_{k<j} A_k - (j-1)*e_ins, done with a
//     per-lane serial pass and a 5-step shuffle scan across lanes;
//   * row maximum (last column wins ties), band trimming (first / last non-zero cell) and the
//     to-end score are warp reductions.
// All control decisions are warp-uniform, all arithmetic is the reference's int32 arithmetic.
#pragma once
#include "bsq_region.h"

#define BSQ_NEG_INF (-0x40000000)

// single-instruction warp reductions (REDUX)
__device__ __forceinline__ int bsq_warp_max(int v) { return __reduce_max_sync(0xffffffffu, v); }
__device__ __forceinline__ int bsq_warp_min(int v) { return __reduce_min_sync(0xffffffffu, v); }

// One instantiation per column count, never inlined: the DP body is large and is reached from four
// call sites (left/right extension of seeds and of backup seeds); keeping one copy keeps the
// instruction cache warm (the first version stalled mostly on instruction fetch, profiles/README.md).
template <int CMAX>
__device__ __noinline__ bsq_ext_result_t bsq_ksw_extend_warp_c(int qlen, bsq_qacc_t qget, int tlen, bsq_tacc_t tget, const int8_t *mat, int o_del,
                                                               int e_del, int o_ins, int e_ins, int w, int end_bonus, int zdrop, int h0) {
  const int lane = threadIdx.x & 31;
  const int C = (qlen + 1 + 31) >> 5;  // columns 0..qlen
  const int j0 = lane * C;
  const int oe_del = o_del + e_del, oe_ins = o_ins + e_ins;
  int H[CMAX], E[CMAX], Q[CMAX];
  BSQ_CTR(BSQ_CTR_KSW, lane == 0);
  // first row (ksw.c:395-397)
  const int h_1 = h0 > oe_ins ? h0 - oe_ins : 0;
#pragma unroll
  for (int k = 0; k < CMAX; ++k) {
    const int j = j0 + k;
    int v = 0;
    if (k < C && j <= qlen) {
      if (j == 0) v = h0;
      else if (j == 1) v = h_1;
      else v = (h_1 - (j - 2) * e_ins > e_ins) ? h_1 - (j - 1) * e_ins : 0;
    }
    H[k] = v; E[k] = 0;
    Q[k] = (k < C && j < qlen) ? qget(j) : 4;
  }
  // band cap (ksw.c:399-407)
  int mx = 0;
  for (int i = 0; i < 25; ++i) mx = mx > mat[i] ? mx : mat[i];
  int max_ins = (int)((double)(qlen * mx + end_bonus - o_ins) / e_ins + 1.);
  max_ins = max_ins > 1 ? max_ins : 1;
  w = w < max_ins ? w : max_ins;
  int max_del = (int)((double)(qlen * mx + end_bonus - o_del) / e_del + 1.);
  max_del = max_del > 1 ? max_del : 1;
  w = w < max_del ? w : max_del;
  int max = h0, max_i = -1, max_j = -1, max_ie = -1, gscore = -1, max_off = 0;
  int beg = 0, end = qlen;
  for (int i = 0; i < tlen; ++i) {
    const int8_t *row = mat + 5 * tget(i);
    const int r0 = row[0], r1 = row[1], r2 = row[2], r3 = row[3], r4 = row[4];  // uniform loads; picked per column below
    if (beg < i - w) beg = i - w;
    if (end > i + w + 1) end = i + w + 1;
    if (end > qlen) end = qlen;
    BSQ_CTR(BSQ_CTR_CELLS, lane == 0 ? (end > beg ? end - beg : 0) : 0);
    int h1_first;
    if (beg == 0) { h1_first = h0 - (o_del + e_del * (i + 1)); if (h1_first < 0) h1_first = 0; }
    else h1_first = 0;
    // ---- M, E', and the scan operand A ----
    int M[CMAX], A[CMAX];
    int run = BSQ_NEG_INF;  // serial prefix max of A inside the lane (exclusive)
    int Pk[CMAX];
#pragma unroll
    for (int k = 0; k < CMAX; ++k) {
      const int j = j0 + k;
      const bool act = k < C && j >= beg && j < end;
      int m_ = 0;
      if (act) {
        const int q = Q[k];
        const int sc = q == 0 ? r0 : q == 1 ? r1 : q == 2 ? r2 : q == 3 ? r3 : r4;
        m_ = H[k]; m_ = m_ ? m_ + sc : 0;
      }
      M[k] = m_;
      int t = m_ - oe_ins; t = t > 0 ? t : 0;
      A[k] = act ? t + j * e_ins : BSQ_NEG_INF;
      Pk[k] = run;
      run = run > A[k] ? run : A[k];
    }
    // exclusive prefix max of the lane totals across lanes
    int incl = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl = incl > v ? incl : v;
    }
    int excl = __shfl_up_sync(0xffffffffu, incl, 1);
    if (lane == 0) excl = BSQ_NEG_INF;
    // ---- H(i,j), row max, new E ----
    int hrow[CMAX];
    int m = 0, mj = -1;
#pragma unroll
    for (int k = 0; k < CMAX; ++k) {
      const int j = j0 + k;
      const bool act = k < C && j >= beg && j < end;
      int h = 0;
      if (act) {
        int p = Pk[k] > excl ? Pk[k] : excl;
        int f = p - (j - 1) * e_ins; f = f > 0 ? f : 0;  // F(i,beg) = 0; t >= 0 keeps F >= 0
        int e = E[k];
        h = M[k] > e ? M[k] : e;
        h = h > f ? h : f;
        if (h >= m) { m = h; mj = j; }
        int t = M[k] - oe_del; t = t > 0 ? t : 0;
        e -= e_del; e = e > t ? e : t;
        E[k] = e;
      }
      hrow[k] = h;
    }
    // warp row max; ties: the last column wins (ksw.c:437).  h < 2^15 and j < 2^9 pack into one key.
    {
      const int key = bsq_warp_max((m << 10) | (mj + 1));
      m = key >> 10;
      mj = (key & 1023) - 1;
    }
    // ---- shift: eh[j].h = H(i,j-1) for j in (beg,end], eh[beg].h = h1_first, eh[end].e = 0 ----
    int lastv = 0;
#pragma unroll
    for (int k = 0; k < CMAX; ++k) if (k == C - 1) lastv = hrow[k];
    const int prev_last = __shfl_up_sync(0xffffffffu, lastv, 1);  // H(i, j0-1) from the lane to the left
    const bool nonempty = beg < end;
    int h_end1 = 0;  // H(i,end-1), needed by the to-end score
#pragma unroll
    for (int k = 0; k < CMAX; ++k) {
      const int j = j0 + k;
      if (k < C) {
        const int left = k == 0 ? prev_last : hrow[k - 1];
        if (nonempty) {
          if (j == beg) H[k] = h1_first;
          else if (j > beg && j <= end) H[k] = left;
          if (j == end - 1) h_end1 = hrow[k];
        } else if (j == end) H[k] = h1_first;  // empty band: only eh[end] is touched (ksw.c:449)
        if (j == end) E[k] = 0;
      }
    }
    if (nonempty) h_end1 = bsq_warp_max(h_end1);  // scores are >= 0 and exactly one lane holds the value
    else h_end1 = h1_first;
    if ((nonempty ? end : beg) == qlen) {  // the column loop stopped at the query end (ksw.c:450-453)
      max_ie = gscore > h_end1 ? max_ie : i;
      gscore = gscore > h_end1 ? gscore : h_end1;
    }
    if (m == 0) break;
    if (m > max) {
      max = m; max_i = i; max_j = mj;
      max_off = max_off > bsq_iabs(mj - i) ? max_off : bsq_iabs(mj - i);
    } else if (zdrop > 0) {
      if (i - max_i > mj - max_j) {
        if (max - m - ((i - max_i) - (mj - max_j)) * e_del > zdrop) break;
      } else {
        if (max - m - ((mj - max_j) - (i - max_i)) * e_ins > zdrop) break;
      }
    }
    // ---- trim the band to the non-zero extent (ksw.c:466-469) ----
    int first_nz = end, last_nz = -1;
#pragma unroll
    for (int k = 0; k < CMAX; ++k) {
      const int j = j0 + k;
      if (k < C && j >= beg && j <= end && (H[k] != 0 || E[k] != 0)) {
        if (j < end && j < first_nz) first_nz = j;
        if (j > last_nz) last_nz = j;
      }
    }
    first_nz = bsq_warp_min(first_nz);
    last_nz = bsq_warp_max(last_nz);
    beg = first_nz;                                   // first non-zero cell in [beg,end), else end
    const int jj = last_nz >= beg ? last_nz : beg - 1;  // last non-zero cell in [beg,end], else beg-1
    end = jj + 2 < qlen ? jj + 2 : qlen;
  }
  bsq_ext_result_t r;
  r.score = max; r.qle = max_j + 1; r.tle = max_i + 1; r.gtle = max_ie + 1; r.gscore = gscore; r.max_off = max_off;
  return r;
}

// dispatch on the number of columns per lane
__device__ __forceinline__ bsq_ext_result_t bsq_ksw_extend_warp(int qlen, bsq_qacc_t qget, int tlen, bsq_tacc_t tget, const int8_t *mat, int o_del,
                                                                int e_del, int o_ins, int e_ins, int w, int end_bonus, int zdrop, int h0) {
  const int C = (qlen + 1 + 31) >> 5;
  if (C <= 2) return bsq_ksw_extend_warp_c<2>(qlen, qget, tlen, tget, mat, o_del, e_del, o_ins, e_ins, w, end_bonus, zdrop, h0);
  if (C <= 5) return bsq_ksw_extend_warp_c<5>(qlen, qget, tlen, tget, mat, o_del, e_del, o_ins, e_ins, w, end_bonus, zdrop, h0);
  return bsq_ksw_extend_warp_c<9>(qlen, qget, tlen, tget, mat, o_del, e_del, o_ins, e_ins, w, end_bonus, zdrop, h0);
}

struct bsq_warp_policy {
  __device__ static bool leader() { return (threadIdx.x & 31) == 0; }
  __device__ static void sync() { __syncwarp(); }
  // asymmetric_flt_seed (memchain.c:138-149), 32 seed positions per step
  __device__ static bool asym_conflict(const bsq_devidx_t &ix, const bsq_seed_t &s, const uint8_t *query) {
    const int lane = threadIdx.x & 31;
    for (int i0 = 0; i0 < s.len; i0 += 32) {
      const int i = i0 + lane;
      bool bad = false;
      if (i < s.len) {
        const int r = bsq_ref_base(ix, s.rbeg + i), qv = query[s.qbeg + i];
        bad = (r == 3 && qv == 1) || (r == 0 && qv == 2);
      }
      if (__any_sync(0xffffffffu, bad)) return true;
    }
    return false;
  }
  __device__ static bsq_ext_result_t extend(int qlen, bsq_qacc_t qget, int tlen, bsq_tacc_t tget, const int8_t *mat, int o_del, int e_del, int o_ins,
                                            int e_ins, int w, int end_bonus, int zdrop, int h0, bsq_ksw_scratch_t *) {
    return bsq_ksw_extend_warp(qlen, qget, tlen, tget, mat, o_del, e_del, o_ins, e_ins, w, end_bonus, zdrop, h0);
  }
};
