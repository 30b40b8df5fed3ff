// Warp-cooperative banded extension: ksw_extend2 (lib/aln/ksw.c:380-479) with one DP row spread over
// the 32 lanes of a warp.  Device only.
//
// The reference's band is adaptive: after every row it is trimmed to the non-zero extent of that row
// (ksw.c:466-469), and the trimming is not value-neutral, so rows must be finished one at a time
//  -- an anti-diagonal wavefront inside one extension would read cells the
// reference never computes.  Rows are therefore processed synchronously, and the lanes stripe the
// CURRENT band (typically 20-40 columns), not the whole query:
//   * the reference's eh[] array (H of the previous row shifted by one, E of this row, stale cells
//     outside the band included) lives in shared memory, one slice per warp; in row i lane L owns
//     columns beg + L, beg + L + 32, ...  Each lane reads and writes only its own index, so the row is
//     updated in place like the reference does;
//   * the horizontal gap state F(i,j+1) = max(F(i,j) - e_ins, max(M(i,j) - oe_ins, 0)) is a max-plus
//     prefix scan over the columns: A_k = t_k + k*e_ins, F_j = max_{k<j} A_k - (j-1)*e_ins, a 5-step
//     shuffle scan per 32 columns with a carry between chunks;
//   * row maximum (last column wins ties) is a REDUX over packed keys, band trimming (first / last
//     non-zero cell) comes from ballots;
//   * the 32 target bases of a block of rows are decoded by 32 lanes at once and broadcast row by row.
// All control decisions are warp-uniform, all arithmetic is the reference's int32 arithmetic.
#pragma once
#include "bsq_region.h"

#define BSQ_NEG_INF (-0x40000000)
#define BSQ_KSW_WARPS 4  // warps per CTA of every kernel that calls bsq_ksw_extend_warp

// single-instruction warp reductions (REDUX)
__device__ __forceinline__ int bsq_warp_max(int v) { return __reduce_max_sync(0xffffffffu, v); }
__device__ __forceinline__ int bsq_warp_min(int v) { return __reduce_min_sync(0xffffffffu, v); }

// Never inlined: the DP body is reached from four call sites (left/right extension of seeds and of
// backup seeds); one copy keeps the instruction cache warm.
__device__ __noinline__ bsq_ext_result_t bsq_ksw_extend_warp(int qlen, bsq_qacc_t qget, int tlen, bsq_tacc_t tget, const int8_t *mat, int o_del,
                                                             int e_del, int o_ins, int e_ins, int w, int end_bonus, int zdrop, int h0) {
  __shared__ int2 s_eh[BSQ_KSW_WARPS][BSQ_MAX_READ_LEN + 2];
  __shared__ uint8_t s_q[BSQ_KSW_WARPS][BSQ_MAX_READ_LEN + 8];
  __shared__ int8_t s_mat[BSQ_KSW_WARPS][32];
  const int lane = threadIdx.x & 31, wid = (threadIdx.x >> 5) & (BSQ_KSW_WARPS - 1);
  int2 *eh = s_eh[wid];
  uint8_t *qs = s_q[wid];
  int8_t *sm = s_mat[wid];
  const int oe_del = o_del + e_del, oe_ins = o_ins + e_ins;
  BSQ_CTR(BSQ_CTR_KSW, lane == 0);
  __syncwarp();  // the previous extension of this warp is done with the slices
  // first row (ksw.c:395-397), query codes, scoring matrix
  const int h_1 = h0 > oe_ins ? h0 - oe_ins : 0;
  for (int j = lane; j <= qlen; j += 32) {
    int v;
    if (j == 0) v = h0;
    else if (j == 1) v = h_1;
    else v = (h_1 - (j - 2) * e_ins > e_ins) ? h_1 - (j - 1) * e_ins : 0;
    eh[j] = make_int2(v, 0);
    if (j < qlen) qs[j] = (uint8_t)qget(j);
  }
  int mx = 0;
  {
    const int v = lane < 25 ? mat[lane] : 0;
    if (lane < 25) sm[lane] = (int8_t)v;
    mx = bsq_warp_max(v);
    mx = mx > 0 ? mx : 0;
  }
  __syncwarp();
  // band cap (ksw.c:399-407)
  int max_ins = (int)((double)(qlen * mx + end_bonus - o_ins) / e_ins + 1.);
  max_ins = max_ins > 1 ? max_ins : 1;
  w = w < max_ins ? w : max_ins;
  int max_del = (int)((double)(qlen * mx + end_bonus - o_del) / e_del + 1.);
  max_del = max_del > 1 ? max_del : 1;
  w = w < max_del ? w : max_del;
  int max = h0, max_i = -1, max_j = -1, max_ie = -1, gscore = -1, max_off = 0;
  int beg = 0, end = qlen;
  int tblk = 0;  // target bases of rows [i & ~31, +32), one per lane
  for (int i = 0; i < tlen; ++i) {
    if ((i & 31) == 0) tblk = i + lane < tlen ? tget(i + lane) : 4;
    const int8_t *row = sm + 5 * __shfl_sync(0xffffffffu, tblk, i & 31);
    if (beg < i - w) beg = i - w;
    if (end > i + w + 1) end = i + w + 1;
    if (end > qlen) end = qlen;
    BSQ_CTR(BSQ_CTR_CELLS, lane == 0 ? (end > beg ? end - beg : 0) : 0);
    int h1_first;
    if (beg == 0) { h1_first = h0 - (o_del + e_del * (i + 1)); if (h1_first < 0) h1_first = 0; }
    else h1_first = 0;
    int key = 0;               // (row max << 10) | (column + 1): ties go to the last column (ksw.c:437)
    int carry = BSQ_NEG_INF;   // max of A over the chunks already done
    int prevh = h1_first;      // H(i, j-1) for the first column of the chunk; eh[beg].h = h1_first
    int nz_first = end, nz_last = -1;
    for (int jb = beg; jb < end; jb += 32) {
      const int j = jb + lane;
      const bool act = j < end;
      int2 cur = make_int2(0, 0);
      int sc = 0;
      if (act) { cur = eh[j]; sc = row[qs[j]]; }
      const int M = cur.x ? cur.x + sc : 0;  // no restart from 0 (ksw.c:433)
      int t = M - oe_ins; t = t > 0 ? t : 0;
      int incl = act ? t + j * e_ins : BSQ_NEG_INF;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);  // lanes below o get their own value back: max is a no-op
        incl = incl > v ? incl : v;
      }
      int p = __shfl_up_sync(0xffffffffu, incl, 1);
      if (lane == 0) p = BSQ_NEG_INF;
      p = p > carry ? p : carry;
      int f = p - (j - 1) * e_ins; f = f > 0 ? f : 0;  // F(i,beg) = 0; t >= 0 keeps F >= 0
      int h = M > cur.y ? M : cur.y;
      h = h > f ? h : f;
      h = act ? h : 0;
      const int k2 = act ? (h << 10) | (j + 1) : 0;
      key = key > k2 ? key : k2;
      int e = M - oe_del; e = e > 0 ? e : 0;
      { const int e2 = cur.y - e_del; e = e > e2 ? e : e2; }
      int left = __shfl_up_sync(0xffffffffu, h, 1);
      if (lane == 0) left = prevh;
      if (act) eh[j] = make_int2(left, e);  // eh[j].h = H(i,j-1), eh[j].e = E(i+1,j)
      const unsigned nzb = __ballot_sync(0xffffffffu, act && (left != 0 || e != 0));
      if (nzb) {
        if (nz_last < 0) nz_first = jb + __ffs(nzb) - 1;
        nz_last = jb + 31 - __clz(nzb);
      }
      const int nact = end - jb < 32 ? end - jb : 32;
      prevh = __shfl_sync(0xffffffffu, h, nact - 1);
      const int ctot = __shfl_sync(0xffffffffu, incl, 31);
      carry = carry > ctot ? carry : ctot;
    }
    // eh[end].h = H(i,end-1) (or h1_first for an empty band), eh[end].e = 0 (ksw.c:449)
    const int h_end1 = prevh;
    if (lane == 0) eh[end] = make_int2(h_end1, 0);
    if (h_end1 != 0) nz_last = end;  // index end counts for the last, never for the first non-zero cell
    __syncwarp();
    key = bsq_warp_max(key);
    const int m = key >> 10, mj = (key & 1023) - 1;
    if ((beg < end ? end : beg) == qlen) {  // the column loop stopped at the query end (ksw.c:450-453)
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
    // first non-zero cell in [beg,end) else end; last non-zero cell in [beg,end] else beg-1
    const int nbeg = nz_first < end ? nz_first : end;
    const int jj = nz_last >= nbeg ? nz_last : nbeg - 1;
    beg = nbeg;
    end = jj + 2 < qlen ? jj + 2 : qlen;
  }
  bsq_ext_result_t r;
  r.score = max; r.qle = max_j + 1; r.tle = max_i + 1; r.gtle = max_ie + 1; r.gscore = gscore; r.max_off = max_off;
  return r;
}

// cal_max_gap (memchain.c:576-582) depends on the query length only: one table per CTA instead of two integer
// divisions per call (it is evaluated for every seed of every chain)
__device__ __forceinline__ int32_t *bsq_gap_tab() {
  __shared__ int32_t tab[BSQ_MAX_READ_LEN + 2];
  return tab;
}
__device__ __forceinline__ void bsq_gap_tab_init(const bsq_devopt_t &opt) {  // all threads of the CTA
  int32_t *t = bsq_gap_tab();
  for (int q = threadIdx.x; q <= BSQ_MAX_READ_LEN; q += blockDim.x) t[q] = bsq_cal_max_gap(opt, q);
  __syncthreads();
}

struct bsq_warp_policy {
  __device__ static bool leader() { return (threadIdx.x & 31) == 0; }
  __device__ static int max_gap(const bsq_devopt_t &opt, int qlen) {
    return (unsigned)qlen <= (unsigned)BSQ_MAX_READ_LEN ? bsq_gap_tab()[qlen] : bsq_cal_max_gap(opt, qlen);
  }
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
