// libbsq.so -- pileup half: methylation counting over coordinate-sorted reads (sm_100a).
//
//   k_plp_win     coordinate-sorted reads (the normal case): one CTA per window of 1024 loci whose counters
//                 [locus][sample][12] live in SHARED memory.  The warps of the CTA walk the reads that can touch the
//                 window (k_plp_ranges: binary search on the sorted positions); per read one WARP does
//                 bisulfite-strand inference, read filters and cnt_retention as warp reductions over the aligned
//                 bases (lanes stride the bases, coalesced SEQ/QUAL/REF loads), then one retention/conversion/base
//                 event per aligned base as shared-memory atomics; finally one thread per locus takes the
//                 per-locus decisions (ambiguity redistribution, top mutant, emit rule, methcallable, 5-mer
//                 cytosine context) straight from shared memory -> flags + dense records.  The counters never
//                 travel through HBM.
//   k_plp_pile, k_plp_locus   the same two steps over a counter tile in global memory, for reads that are not
//                 coordinate-sorted (several BAMs concatenated).
//   k_plp_compact emitted loci -> contiguous output (order = position), via a device prefix sum.
// Reference: src/pileup.c:707-831 (events), :372-387 (plp_getcnts), :312-370 and :415-485 (per locus),
// src/bisc_utils.c:33-122,163-238.  Integer work only; bit-exact against oracle/bsq_oracle_pileup.c.
#include <cuda_runtime.h>
#include <limits.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <thread>
#include <vector>
#include <cub/device/device_scan.cuh>

#include "../../include/bsq.h"
#include "bsq_internal.h"

#define PLP_NCNT 12            // meth[3] base[7] dp pad
#define PLP_TILE (8 << 20)     // loci per internal tile

#define CKP(call)                                                                                  \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) {                                                                       \
      bsq_set_error("%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_));            \
      return e_ == cudaErrorMemoryAllocation ? BSQ_ENOMEM : BSQ_ENODEV;                            \
    }                                                                                              \
  } while (0)

enum { M_RET = 0, M_CONV = 1, M_NA = 2 };
enum { B_A = 0, B_C, B_G, B_T, B_N, B_Y, B_R };
enum { CT_HCG = 0, CT_HCHG, CT_HCHH, CT_GCG, CT_GCHG, CT_GCHH, CT_NA };

struct DBuf {
  void *p = nullptr; size_t cap = 0;
  int need(size_t b) {
    if (b <= cap) return 0;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    size_t want = b + b / 4 + 256;
    if (cudaMalloc(&p, want) != cudaSuccess) { bsq_set_error("cudaMalloc(%zu) failed", want); return BSQ_ENOMEM; }
    cap = want;
    return 0;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
  template <typename T> T *as() const { return (T *)p; }
};

// device copy of bsq_plp_reads
struct DevReads {
  int64_t n_reads;
  const int32_t *pos, *mpos, *mate_rlen, *l_qseq, *nm, *as;
  const uint16_t *flag;
  const uint8_t *mapq;
  const int8_t *bss_tag;
  const uint8_t *sid;
  const int32_t *n_cigar;
  const int64_t *cigar_off;
  const uint32_t *cigar;
  const int64_t *seq_off;
  const uint8_t *seq;
  const int64_t *qual_off;
  const uint8_t *qual;
};

struct bsq_plp {
  int device, n_bams;
  cudaStream_t stream;
  cudaEvent_t ev[4];
  DBuf ref, b_pos, b_mpos, b_mrl, b_lq, b_nm, b_as, b_flag, b_mapq, b_bss, b_sid, b_nc, b_coff, b_cig, b_soff, b_seq, b_qoff, b_qual;
  DBuf cnt, flags, dense, offs, out, cub_tmp, scal, wr0, wr1;
  int32_t ref_len;
  DevReads dr;
  bool sorted;     // reads are coordinate-sorted: windows find their reads by binary search on the device (k_plp_ranges)
  int64_t n_reads;
  int32_t max_span;
  int64_t n_out;
  int64_t counters[8];
};

__constant__ uint8_t c_nt16_to_nt4[16] = {4, 0, 1, 4, 2, 4, 4, 4, 3, 4, 4, 4, 4, 4, 4, 4};

__device__ __forceinline__ int rd_base(const uint8_t *seq, int q) {
  uint8_t b = seq[q >> 1];
  return c_nt16_to_nt4[(q & 1) ? (b & 0xf) : (b >> 4)];
}

// One warp, one read: bisulfite-strand inference, read filters, then one event per aligned base inside [beg, end),
// accumulated with integer atomics into cnt[((p - beg) * n_bams + sid) * PLP_NCNT + k] (global tile or a shared-memory
// window).  Returns the number of events of this lane.
__device__ __forceinline__ unsigned long long plp_read_events(const DevReads &rd, int64_t i, const bsq_plp_conf &cf, const uint8_t *ref, int32_t ref_len,
                                                              int32_t beg, int32_t end, int n_bams, int *cnt) {
  const int lane = threadIdx.x & 31;
  const uint32_t *cig = rd.cigar + rd.cigar_off[i];
  const uint8_t *seq = rd.seq + rd.seq_off[i];
  const uint8_t *qual = rd.qual + rd.qual_off[i];
  const int nc = rd.n_cigar[i], flag = rd.flag[i], sid = rd.sid[i];
  const uint32_t pos1 = (uint32_t)rd.pos[i] + 1;
  // quick reject: the read cannot touch [beg,end) -- only an optimisation, the per-base test below decides
  int bsstrand = rd.bss_tag[i];
  const int lq = rd.l_qseq[i];
  // ---- pass A: strand inference + retention count (warp reductions over aligned bases) ----
  int nC2T = 0, nG2A = 0, nCC = 0, nGG = 0;
  uint32_t read_length = 0;
  {
    uint32_t rpos = pos1, qpos = 0;
    for (int k = 0; k < nc; ++k) {
      const uint32_t op = cig[k] & 0xf, ol = cig[k] >> 4;
      if (op == 0 || op == 7 || op == 8) {
        for (uint32_t j = lane; j < ol; j += 32) {
          const uint32_t p = rpos + j;
          if (p < 1 || p > (uint32_t)ref_len || qpos + j >= (uint32_t)lq) continue;
          const int rb = ref[p - 1], qb = rd_base(seq, qpos + j);
          if (rb == 1 && qb == 1) nCC++;
          if (rb == 2 && qb == 2) nGG++;
          if (qual[qpos + j] < (uint32_t)cf.min_base_qual) continue;
          if (rb == 1 && qb == 3) nC2T++;
          if (rb == 2 && qb == 0) nG2A++;
        }
        rpos += ol; qpos += ol; read_length += ol;
      } else if (op == 1 || op == 4 || op == 5) qpos += ol;  // H advances qpos like the reference (pileup.c:822); bases past SEQ are guarded below
      else if (op == 2) { rpos += ol; read_length += ol; }
      else if (op == 3) read_length += ol;
    }
  }
  if (bsstrand < 0) {
    nC2T = __reduce_add_sync(0xffffffffu, nC2T);
    nG2A = __reduce_add_sync(0xffffffffu, nG2A);
    bsstrand = nC2T >= nG2A ? 0 : 1;
  }
  // ---- read-level filters (pileup.c:713-729) ----
  if (rd.mapq[i] < cf.min_mapq) return 0;
  if (lq < 0 || lq < cf.min_read_len) return 0;
  if (flag > 0) {
    if (cf.filter_secondary && (flag & 0x100)) return 0;
    if (cf.filter_duplicate && (flag & 0x400)) return 0;
    if (cf.filter_ppair && (flag & 0x1) && !(flag & 0x2)) return 0;
    if (cf.filter_qcfail && (flag & 0x200)) return 0;
  }
  if (rd.nm[i] != INT_MIN && rd.nm[i] > cf.max_nm) return 0;
  if (rd.as[i] != INT_MIN && rd.as[i] < cf.min_score) return 0;
  {
    const uint32_t c = (uint32_t)__reduce_add_sync(0xffffffffu, bsstrand ? nCC : nGG);  // cnt_retention quirk: C/C on BSC
    if (c > (uint32_t)cf.max_retention) return 0;
  }
  // ---- pass B: events ----
  const uint32_t rmpos = (uint32_t)rd.mpos[i] + 1;
  const uint32_t mate_length = rd.mate_rlen[i] >= 0 ? (uint32_t)rd.mate_rlen[i] : read_length;
  const uint32_t rend = pos1 + read_length - 1, rmend = rmpos + mate_length - 1;
  const uint32_t ov_hi = rend < rmend ? rend : rmend;
  const bool dbl = cf.filter_doublecnt && (flag & 0x80);
  uint32_t rpos = pos1, qpos = 0;
  unsigned long long ev = 0;
  for (int k = 0; k < nc; ++k) {
    const uint32_t op = cig[k] & 0xf, ol = cig[k] >> 4;
    if (op == 0 || op == 7 || op == 8) {
      const uint32_t ov_lo = rpos > rmpos ? rpos : rmpos;
      for (uint32_t j = lane; j < ol; j += 32) {
        const uint32_t p = rpos + j;
        if (p < (uint32_t)beg || p >= (uint32_t)end) continue;
        if (dbl && p >= ov_lo && p <= ov_hi) continue;
        int *lc = cnt + ((int64_t)(p - beg) * n_bams + sid) * PLP_NCNT;
        atomicAdd(lc + 10, 1);  // DP: every event
        ++ev;
        if (qpos + j >= (uint32_t)lq) continue;  // past SEQ after a leading H: the reference's 3'-distance rule drops it from the counts
        const int rb = ref[p - 1], qb = rd_base(seq, qpos + j);
        int meth, base;
        if (bsstrand) { meth = rb == 2 ? (qb == 0 ? M_CONV : qb == 2 ? M_RET : M_NA) : M_NA; base = qb == 0 ? B_R : qb; }
        else { meth = rb == 1 ? (qb == 3 ? M_CONV : qb == 1 ? M_RET : M_NA) : M_NA; base = qb == 3 ? B_Y : qb; }
        const uint32_t q7 = qual[qpos + j] & 0x7f, qp = (qpos + j + 1) & 0xffff, rl = (uint32_t)lq & 0xffff;
        if (q7 < (uint32_t)cf.min_base_qual) continue;
        if (qp <= (uint32_t)cf.min_dist_end_5p || rl < qp + (uint32_t)cf.min_dist_end_3p) continue;
        atomicAdd(lc + meth, 1);
        atomicAdd(lc + 3 + base, 1);
      }
      rpos += ol; qpos += ol;
    } else if (op == 1 || op == 4 || op == 5) qpos += ol;  // H advances qpos like the reference (pileup.c:822); bases past SEQ are guarded below
    else if (op == 2) rpos += ol;
  }
  return ev;
}

// one warp per read, counters in a global tile (used when the reads are not coordinate-sorted)
__global__ void __launch_bounds__(256) k_plp_pile(DevReads rd, int64_t r0, int64_t r1, bsq_plp_conf cf, const uint8_t *ref, int32_t ref_len,
                                                  int32_t beg, int32_t end, int n_bams, int *cnt, unsigned long long *n_events) {
  const int64_t i = r0 + (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  if (i >= r1) return;
  unsigned long long ev = plp_read_events(rd, i, cf, ref, ref_len, beg, end, n_bams, cnt);
  ev = __reduce_add_sync(0xffffffffu, (unsigned)ev);
  if ((threadIdx.x & 31) == 0 && ev) atomicAdd(n_events, ev);
}

__device__ __forceinline__ char nt4_char(int c) { return c == 0 ? 'A' : c == 1 ? 'C' : c == 2 ? 'G' : c == 3 ? 'T' : 'N'; }

// One locus: ambiguity redistribution, top mutant, emit rule, methcallable, cytosine context (plp_format, pileup.c:415-485)
// from its counters lc0[n_bams][PLP_NCNT]; writes flags[l] and, when the locus is emitted, dense[l * n_bams + s].
__device__ __forceinline__ void plp_locus(const int *lc0, const bsq_plp_conf &cf, const uint8_t *ref, int32_t ref_len, int32_t rpos, int64_t l, int n_bams,
                                          int32_t *flags, bsq_plp_rec *dense) {
  int touched = 0;
  for (int s = 0; s < n_bams; ++s) touched |= lc0[s * PLP_NCNT + 10];
  flags[l] = 0;
  if (!touched) return;
  const int rb = ref[rpos - 1];
  if (rb > 3) return;
  int raw_all[7], all_base[7], all_meth[3] = {0, 0, 0};
  for (int b = 0; b < 7; ++b) { raw_all[b] = 0; all_base[b] = 0; }
  for (int s = 0; s < n_bams; ++s)
    for (int b = 0; b < 7; ++b) raw_all[b] += lc0[s * PLP_NCNT + 3 + b];
  const bool yT = cf.ambi_redist && (rb == B_T || raw_all[B_T]) && raw_all[B_C] == 0 && rb != B_C;
  const bool yC = cf.ambi_redist && (rb == B_C || raw_all[B_C]) && raw_all[B_T] == 0 && rb != B_T;
  const bool rA = cf.ambi_redist && (rb == B_A || raw_all[B_A]) && raw_all[B_G] == 0 && rb != B_G;
  const bool rG = cf.ambi_redist && (rb == B_G || raw_all[B_G]) && raw_all[B_A] == 0 && rb != B_A;
  for (int s = 0; s < n_bams; ++s) {
    const int *lc = lc0 + s * PLP_NCNT;
    int c1[7];
    for (int b = 0; b < 7; ++b) c1[b] = lc[3 + b];
    if (yT) { c1[B_T] += c1[B_Y]; c1[B_Y] = 0; }
    if (yC) { c1[B_C] += c1[B_Y]; c1[B_Y] = 0; }
    if (rA) { c1[B_A] += c1[B_R]; c1[B_R] = 0; }
    if (rG) { c1[B_G] += c1[B_R]; c1[B_R] = 0; }
    for (int b = 0; b < 7; ++b) all_base[b] += c1[b];
    for (int b = 0; b < 3; ++b) all_meth[b] += lc[b];
  }
  // top_mutant: highest count first, ties in base-code order (glibc qsort is stable for 7 items)
  int cm1 = -1;
  {
    int best = 0;
    for (int b = 0; b < 7; ++b) {
      if (b == B_N || b == rb) continue;
      if (b == B_R && (rb == B_A || rb == B_G)) continue;
      if (b == B_Y && (rb == B_C || rb == B_T)) continue;
      if (all_base[b] > best) { best = all_base[b]; cm1 = b; }
    }
  }
  if (cm1 < 0 && !cf.verbose && all_meth[M_RET] == 0 && all_meth[M_CONV] == 0) return;
  char n5[5] = {'N', 'N', 'N', 'N', 'N'};
  int ctx = CT_NA;
  if (rb == B_C || rb == B_G) {
    for (int q = 0; q < 5; ++q) {
      const int32_t p = rpos - 2 + q;
      n5[q] = (p >= 1 && p <= ref_len) ? nt4_char(ref[p - 1]) : 'N';
    }
    if (rb == B_G) {
      char t[5];
      for (int q = 0; q < 5; ++q) { char c = n5[4 - q]; t[q] = c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : c == 'T' ? 'A' : 'N'; }
      for (int q = 0; q < 5; ++q) n5[q] = t[q];
    }
    bool has_n = false;
    for (int q = 0; q < 5; ++q) has_n |= n5[q] == 'N';
    if (!has_n) {
      if (n5[3] == 'G') ctx = n5[1] == 'G' ? CT_GCG : CT_HCG;
      else if (n5[4] == 'G') ctx = n5[1] == 'G' ? CT_GCHG : CT_HCHG;
      else ctx = n5[1] == 'G' ? CT_GCHH : CT_HCHH;
    }
  }
  int any_callable = 0;
  for (int s = 0; s < n_bams; ++s) {
    const int *lc = lc0 + s * PLP_NCNT;
    bsq_plp_rec r;
    memset(&r, 0, sizeof r);
    int c1[7];
    for (int b = 0; b < 7; ++b) { c1[b] = lc[3 + b]; r.base[b] = c1[b]; }
    if (yT) { c1[B_T] += c1[B_Y]; c1[B_Y] = 0; }
    if (yC) { c1[B_C] += c1[B_Y]; c1[B_Y] = 0; }
    if (rA) { c1[B_A] += c1[B_R]; c1[B_R] = 0; }
    if (rG) { c1[B_G] += c1[B_R]; c1[B_R] = 0; }
    int callable = 0;
    if (lc[M_RET] + lc[M_CONV] > 0) {
      if (rb == B_C) {
        if (c1[B_T] == 0) callable = 1;
        else if (c1[B_C] > 0 && __ddiv_rn((double)c1[B_T], (double)c1[B_C]) < 0.05) callable = 1;
      }
      if (rb == B_G) {
        if (c1[B_A] == 0) callable = 1;
        else if (c1[B_G] > 0 && __ddiv_rn((double)c1[B_A], (double)c1[B_G]) < 0.05) callable = 1;
      }
    }
    any_callable |= callable;
    r.pos = rpos; r.dp = lc[10];
    for (int b = 0; b < 3; ++b) r.meth[b] = lc[b];
    for (int b = 0; b < 7; ++b) r.base_redist[b] = c1[b];
    r.rb_code = (uint8_t)rb; r.cm1 = (int8_t)cm1; r.ctx = (uint8_t)ctx; r.methcallable = (uint8_t)callable;
    for (int q = 0; q < 5; ++q) r.n5[q] = n5[q];
    dense[l * n_bams + s] = r;
  }
  for (int s = 0; s < n_bams; ++s) dense[l * n_bams + s].any_callable = (uint8_t)any_callable;
  flags[l] = 1;
}

// one thread per locus of the tile, counters in the global tile
__global__ void k_plp_locus(const int *cnt, bsq_plp_conf cf, const uint8_t *ref, int32_t ref_len, int32_t beg, int64_t nl, int n_bams,
                            int32_t *flags, bsq_plp_rec *dense) {
  const int64_t l = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= nl) return;
  plp_locus(cnt + l * n_bams * PLP_NCNT, cf, ref, ref_len, beg + (int32_t)l, l, n_bams, flags, dense);
}

// Coordinate-sorted reads: one CTA per window of W loci.  The window's counters live in shared memory, the warps of the
// CTA walk the reads that can touch the window (range from k_plp_ranges), events are shared-memory atomics, and the
// per-locus decisions are taken straight from shared memory -- the counters never travel through HBM.
__global__ void k_plp_ranges(const int32_t *pos, int64_t n_reads, int32_t tb, int32_t te, int W, int32_t max_span, int64_t *r0, int64_t *r1) {
  const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t wb = (int64_t)tb + w * W;
  if (wb >= te) return;
  const int64_t we = wb + W < te ? wb + W : te;
  // reads with pos in [wb - 1 - max_span, we - 1) can touch [wb, we)
  const int64_t lo_v = wb - 1 - (int64_t)max_span, hi_v = we - 1;
  int64_t lo = 0, hi = n_reads;
  while (lo < hi) { const int64_t m = (lo + hi) >> 1; if (pos[m] < lo_v) lo = m + 1; else hi = m; }
  r0[w] = lo;
  hi = n_reads;
  while (lo < hi) { const int64_t m = (lo + hi) >> 1; if (pos[m] < hi_v) lo = m + 1; else hi = m; }
  r1[w] = lo;
}

__global__ void __launch_bounds__(256) k_plp_win(DevReads rd, const int64_t *r0s, const int64_t *r1s, bsq_plp_conf cf, const uint8_t *ref, int32_t ref_len,
                                                 int32_t tb, int32_t te, int W, int n_bams, int32_t *flags, bsq_plp_rec *dense,
                                                 unsigned long long *n_events) {
  extern __shared__ int s_cnt[];  // [W][n_bams][PLP_NCNT]
  const int32_t wb = tb + (int32_t)blockIdx.x * W;
  const int32_t we = wb + W < te ? wb + W : te;
  const int n_cnt = (we - wb) * n_bams * PLP_NCNT;
  for (int k = threadIdx.x; k < n_cnt; k += blockDim.x) s_cnt[k] = 0;
  __syncthreads();
  const int64_t r0 = r0s[blockIdx.x], r1 = r1s[blockIdx.x];
  unsigned long long ev = 0;
  for (int64_t i = r0 + (threadIdx.x >> 5); i < r1; i += blockDim.x >> 5) ev += plp_read_events(rd, i, cf, ref, ref_len, wb, we, n_bams, s_cnt);
  ev = __reduce_add_sync(0xffffffffu, (unsigned)ev);
  if ((threadIdx.x & 31) == 0 && ev) atomicAdd(n_events, ev);
  __syncthreads();
  for (int l = threadIdx.x; l < we - wb; l += blockDim.x)
    plp_locus(s_cnt + l * n_bams * PLP_NCNT, cf, ref, ref_len, wb + l, (int64_t)(wb - tb) + l, n_bams, flags, dense);
}

__global__ void k_plp_compact(const int32_t *flags, const int64_t *offs, const bsq_plp_rec *dense, int64_t nl, int n_bams, int64_t out_base,
                              bsq_plp_rec *out) {
  const int64_t l = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= nl || !flags[l]) return;
  for (int s = 0; s < n_bams; ++s) out[(out_base + offs[l]) * n_bams + s] = dense[l * n_bams + s];
}

static inline unsigned nbk(int64_t n, int bs) { return (unsigned)((n + bs - 1) / bs); }

extern "C" {

void bsq_plp_conf_default(bsq_plp_conf *c) {
  memset(c, 0, sizeof *c);
  c->min_base_qual = 20; c->min_read_len = 10; c->min_dist_end_5p = 3; c->min_dist_end_3p = 3; c->min_mapq = 40;
  c->min_score = 40; c->max_nm = 999999; c->max_retention = 999999;
  c->filter_ppair = c->filter_secondary = c->filter_duplicate = c->filter_qcfail = c->filter_doublecnt = 1;
  c->ambi_redist = 1;
}

int bsq_plp_create(int device, int n_bams, bsq_plp **out) {
  if (!out || n_bams < 1 || n_bams > 8) return BSQ_EINVAL;
  int ndev = 0;
  CKP(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) { bsq_set_error("device %d of %d", device, ndev); return BSQ_ENODEV; }
  CKP(cudaSetDevice(device));
  bsq_plp *p = new bsq_plp();
  p->device = device; p->n_bams = n_bams; p->ref_len = 0; p->sorted = false; p->n_reads = 0; p->max_span = 0; p->n_out = 0;
  memset(p->counters, 0, sizeof p->counters);
  CKP(cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking));
  for (int i = 0; i < 4; ++i) CKP(cudaEventCreate(&p->ev[i]));
  *out = p;
  return 0;
}

void bsq_plp_destroy(bsq_plp *p) {
  if (!p) return;
  cudaSetDevice(p->device);
  DBuf *bufs[] = {&p->ref, &p->b_pos, &p->b_mpos, &p->b_mrl, &p->b_lq, &p->b_nm, &p->b_as, &p->b_flag, &p->b_mapq, &p->b_bss, &p->b_sid,
                  &p->b_nc, &p->b_coff, &p->b_cig, &p->b_soff, &p->b_seq, &p->b_qoff, &p->b_qual, &p->cnt, &p->flags, &p->dense, &p->offs,
                  &p->out, &p->cub_tmp, &p->scal, &p->wr0, &p->wr1};
  for (DBuf *b : bufs) b->release();
  for (int i = 0; i < 4; ++i) cudaEventDestroy(p->ev[i]);
  cudaStreamDestroy(p->stream);
  delete p;
}

int bsq_plp_set_contig(bsq_plp *p, const uint8_t *ref, int32_t ref_len) {
  if (!p || !ref || ref_len <= 0) return BSQ_EINVAL;
  CKP(cudaSetDevice(p->device));
  int rc = p->ref.need((size_t)ref_len + 16);
  if (rc) return rc;
  CKP(cudaMemcpyAsync(p->ref.p, ref, (size_t)ref_len, cudaMemcpyHostToDevice, p->stream));
  CKP(cudaStreamSynchronize(p->stream));
  p->ref_len = ref_len;
  return 0;
}

#define UP(buf, src, bytes)                                                                          \
  do {                                                                                               \
    int rc_ = p->buf.need((bytes) + 16);                                                             \
    if (rc_) return rc_;                                                                             \
    if ((bytes) > 0) CKP(cudaMemcpyAsync(p->buf.p, src, (bytes), cudaMemcpyHostToDevice, p->stream)); \
  } while (0)

int bsq_plp_stage(bsq_plp *p, const bsq_plp_reads *r) {
  if (!p || !r || r->n_reads < 0) return BSQ_EINVAL;
  CKP(cudaSetDevice(p->device));
  const int64_t n = r->n_reads;
  // pool sizes, reference span, sortedness and validation of ops / sample ids on the host: one linear pass over the
  // records, cut into slices for a few threads (50 M reads of a chr1-sized contig at 30x take 0.7 s on one)
  int64_t cig_tot = 0, seq_tot = 0, qual_tot = 0;
  int32_t max_span = 0;
  bool sorted = true;
  {
    struct slice_t { int64_t cig = 0, seq = 0, qual = 0, bad = -1; int32_t span = 0; uint32_t bad_op = 0; bool sorted = true; };
    const int nt = n > (1 << 20) ? 8 : 1;
    std::vector<slice_t> sl(nt);
    auto work = [&](int t) {
      slice_t &o = sl[t];
      const int64_t i0 = n * t / nt, i1 = n * (t + 1) / nt;
      for (int64_t i = i0; i < i1; ++i) {
        if (r->sid[i] >= p->n_bams || r->n_cigar[i] < 0) { if (o.bad < 0) { o.bad = i; o.bad_op = ~0u; } continue; }
        const uint32_t *c = r->cigar + r->cigar_off[i];
        int64_t span = 0;
        for (int k = 0; k < r->n_cigar[i]; ++k) {
          const uint32_t op = c[k] & 0xf;
          if ((op == 3 || op == 6 || op > 8) && o.bad < 0) { o.bad = i; o.bad_op = op; }  // the reference abort()s on N / P (pileup.c:826-828)
          if (op == 0 || op == 2 || op == 7 || op == 8) span += c[k] >> 4;
        }
        if (span > o.span) o.span = (int32_t)(span > INT_MAX ? INT_MAX : span);
        if (r->cigar_off[i] + r->n_cigar[i] > o.cig) o.cig = r->cigar_off[i] + r->n_cigar[i];
        const int64_t lq = r->l_qseq[i] > 0 ? r->l_qseq[i] : 0;
        if (r->seq_off[i] + (lq + 1) / 2 > o.seq) o.seq = r->seq_off[i] + (lq + 1) / 2;
        if (r->qual_off[i] + lq > o.qual) o.qual = r->qual_off[i] + lq;
        if (i && r->pos[i] < r->pos[i - 1]) o.sorted = false;
      }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < nt; ++t) th.emplace_back(work, t);
    work(0);
    for (auto &x : th) x.join();
    for (int t = 0; t < nt; ++t) {
      const slice_t &o = sl[t];
      if (o.bad >= 0) {
        if (o.bad_op == ~0u) bsq_set_error("read %lld: bad sample id / n_cigar", (long long)o.bad);
        else bsq_set_error("read %lld: CIGAR operator %u is not supported by pileup", (long long)o.bad, o.bad_op);
        return BSQ_EINVAL;
      }
      if (o.cig > cig_tot) cig_tot = o.cig;
      if (o.seq > seq_tot) seq_tot = o.seq;
      if (o.qual > qual_tot) qual_tot = o.qual;
      if (o.span > max_span) max_span = o.span;
      sorted = sorted && o.sorted;
    }
  }
  UP(b_pos, r->pos, n * 4); UP(b_mpos, r->mpos, n * 4); UP(b_mrl, r->mate_rlen, n * 4); UP(b_lq, r->l_qseq, n * 4);
  UP(b_nm, r->nm, n * 4); UP(b_as, r->as, n * 4); UP(b_flag, r->flag, n * 2); UP(b_mapq, r->mapq, n); UP(b_bss, r->bss_tag, n);
  UP(b_sid, r->sid, n); UP(b_nc, r->n_cigar, n * 4); UP(b_coff, r->cigar_off, n * 8); UP(b_cig, r->cigar, cig_tot * 4);
  UP(b_soff, r->seq_off, n * 8); UP(b_seq, r->seq, seq_tot); UP(b_qoff, r->qual_off, n * 8); UP(b_qual, r->qual, qual_tot);
  DevReads &d = p->dr;
  d.n_reads = n;
  d.pos = p->b_pos.as<int32_t>(); d.mpos = p->b_mpos.as<int32_t>(); d.mate_rlen = p->b_mrl.as<int32_t>(); d.l_qseq = p->b_lq.as<int32_t>();
  d.nm = p->b_nm.as<int32_t>(); d.as = p->b_as.as<int32_t>(); d.flag = p->b_flag.as<uint16_t>(); d.mapq = p->b_mapq.as<uint8_t>();
  d.bss_tag = p->b_bss.as<int8_t>(); d.sid = p->b_sid.as<uint8_t>(); d.n_cigar = p->b_nc.as<int32_t>(); d.cigar_off = p->b_coff.as<int64_t>();
  d.cigar = p->b_cig.as<uint32_t>(); d.seq_off = p->b_soff.as<int64_t>(); d.seq = p->b_seq.as<uint8_t>(); d.qual_off = p->b_qoff.as<int64_t>();
  d.qual = p->b_qual.as<uint8_t>();
  p->sorted = sorted;  // unsorted input: every tile scans every read
  p->n_reads = n; p->max_span = max_span;
  p->counters[0] = n;
  return 0;
}

int bsq_plp_run(bsq_plp *p, const bsq_plp_conf *cf, int32_t beg, int32_t end, int64_t *n_loci) {
  if (!p || !cf || !n_loci || p->ref_len <= 0) return BSQ_EINVAL;
  CKP(cudaSetDevice(p->device));
  if (end > p->ref_len) end = p->ref_len;  // the last base of a contig is never piled
  if (beg < 1) beg = 1;
  *n_loci = 0; p->n_out = 0;
  p->counters[1] = p->counters[2] = p->counters[3] = p->counters[4] = p->counters[5] = 0;
  if (end <= beg) return 0;
  cudaStream_t s = p->stream;
  const int nb = p->n_bams;
  int rc;
  if ((rc = p->scal.need(64))) return rc;
  CKP(cudaMemsetAsync(p->scal.p, 0, 64, s));
  // worst case output: every locus emitted; grow lazily per tile instead
  int64_t out_cap = (int64_t)(p->out.cap / ((size_t)nb * sizeof(bsq_plp_rec))), n_out = 0;  // the output buffer is kept between runs
  for (int64_t tb = beg; tb < end; tb += PLP_TILE) {
    const int64_t te = tb + PLP_TILE < end ? tb + PLP_TILE : end;
    const int64_t nl = te - tb;
    if ((rc = p->flags.need((size_t)nl * 4))) return rc;
    if ((rc = p->offs.need((size_t)(nl + 1) * 8))) return rc;
    if ((rc = p->dense.need((size_t)nl * nb * sizeof(bsq_plp_rec)))) return rc;
    CKP(cudaEventRecord(p->ev[0], s));
    if (p->sorted) {
      // coordinate-sorted reads: windows of W loci with their counters in shared memory (k_plp_win)
      int W = 1024 / nb;
      if (W < 128) W = 128;
      const int64_t n_win = (nl + W - 1) / W;
      if ((rc = p->wr0.need((size_t)n_win * 8))) return rc;
      if ((rc = p->wr1.need((size_t)n_win * 8))) return rc;
      k_plp_ranges<<<nbk(n_win, 256), 256, 0, s>>>(p->dr.pos, p->n_reads, (int32_t)tb, (int32_t)te, W, p->max_span, p->wr0.as<int64_t>(),
                                                    p->wr1.as<int64_t>());
      CKP(cudaGetLastError());
      const size_t smem = (size_t)W * nb * PLP_NCNT * sizeof(int);
      static bool attr_set_[64];  // per device
      int dev_ = 0; cudaGetDevice(&dev_);
      bool &attr_set = attr_set_[dev_ < 0 || dev_ >= 64 ? 0 : dev_];
      if (!attr_set) { CKP(cudaFuncSetAttribute(k_plp_win, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024)); attr_set = true; }
      k_plp_win<<<(unsigned)n_win, 256, smem, s>>>(p->dr, p->wr0.as<int64_t>(), p->wr1.as<int64_t>(), *cf, p->ref.as<uint8_t>(), p->ref_len, (int32_t)tb,
                                                    (int32_t)te, W, nb, p->flags.as<int32_t>(), p->dense.as<bsq_plp_rec>(),
                                                    p->scal.as<unsigned long long>());
      CKP(cudaGetLastError());
      CKP(cudaEventRecord(p->ev[1], s));
    } else {
      // unsorted reads: every read is tried against the tile, counters in a global tile
      if ((rc = p->cnt.need((size_t)nl * nb * PLP_NCNT * 4))) return rc;
      CKP(cudaMemsetAsync(p->cnt.p, 0, (size_t)nl * nb * PLP_NCNT * 4, s));
      const int64_t r0 = 0, r1 = p->n_reads;
      if (r1 > r0) {
        k_plp_pile<<<nbk((r1 - r0) * 32, 256), 256, 0, s>>>(p->dr, r0, r1, *cf, p->ref.as<uint8_t>(), p->ref_len, (int32_t)tb, (int32_t)te, nb,
                                                             p->cnt.as<int>(), p->scal.as<unsigned long long>());
        CKP(cudaGetLastError());
      }
      CKP(cudaEventRecord(p->ev[1], s));
      k_plp_locus<<<nbk(nl, 256), 256, 0, s>>>(p->cnt.as<int>(), *cf, p->ref.as<uint8_t>(), p->ref_len, (int32_t)tb, nl, nb, p->flags.as<int32_t>(),
                                                p->dense.as<bsq_plp_rec>());
      CKP(cudaGetLastError());
    }
    size_t tmpb = 0;
    CKP(cub::DeviceScan::ExclusiveSum(nullptr, tmpb, p->flags.as<int32_t>(), p->offs.as<int64_t>(), (int)nl, s));
    if ((rc = p->cub_tmp.need(tmpb))) return rc;
    CKP(cub::DeviceScan::ExclusiveSum(p->cub_tmp.p, tmpb, p->flags.as<int32_t>(), p->offs.as<int64_t>(), (int)nl, s));
    int64_t last_off = 0; int32_t last_flag = 0;
    CKP(cudaMemcpyAsync(&last_off, p->offs.as<int64_t>() + (nl - 1), 8, cudaMemcpyDeviceToHost, s));
    CKP(cudaMemcpyAsync(&last_flag, p->flags.as<int32_t>() + (nl - 1), 4, cudaMemcpyDeviceToHost, s));
    CKP(cudaStreamSynchronize(s));
    const int64_t n_tile = last_off + last_flag;
    if (n_out + n_tile > out_cap) {  // grow the output, keeping what earlier tiles produced
      int64_t want = (n_out + n_tile) * 2 + 1024;
      void *np_ = nullptr;
      CKP(cudaMalloc(&np_, (size_t)want * nb * sizeof(bsq_plp_rec)));
      if (n_out) CKP(cudaMemcpyAsync(np_, p->out.p, (size_t)n_out * nb * sizeof(bsq_plp_rec), cudaMemcpyDeviceToDevice, s));
      CKP(cudaStreamSynchronize(s));
      p->out.release();
      p->out.p = np_; p->out.cap = (size_t)want * nb * sizeof(bsq_plp_rec);
      out_cap = want;
    }
    k_plp_compact<<<nbk(nl, 256), 256, 0, s>>>(p->flags.as<int32_t>(), p->offs.as<int64_t>(), p->dense.as<bsq_plp_rec>(), nl, nb, n_out,
                                                p->out.as<bsq_plp_rec>());
    CKP(cudaGetLastError());
    CKP(cudaEventRecord(p->ev[2], s));
    CKP(cudaStreamSynchronize(s));
    float a = 0, b = 0;
    cudaEventElapsedTime(&a, p->ev[0], p->ev[1]);
    cudaEventElapsedTime(&b, p->ev[1], p->ev[2]);
    p->counters[4] += (int64_t)(a * 1000); p->counters[5] += (int64_t)(b * 1000);
    n_out += n_tile;
    p->counters[1] += nl;
  }
  unsigned long long evs = 0;
  CKP(cudaMemcpy(&evs, p->scal.p, 8, cudaMemcpyDeviceToHost));
  p->counters[2] = n_out; p->counters[3] = (int64_t)evs;
  p->n_out = n_out;
  *n_loci = n_out;
  return 0;
}

int bsq_plp_fetch(bsq_plp *p, bsq_plp_rec *out) {
  if (!p) return BSQ_EINVAL;
  if (p->n_out == 0) return 0;
  if (!out) return BSQ_EINVAL;
  CKP(cudaSetDevice(p->device));
  CKP(cudaMemcpyAsync(out, p->out.p, (size_t)p->n_out * p->n_bams * sizeof(bsq_plp_rec), cudaMemcpyDeviceToHost, p->stream));
  CKP(cudaStreamSynchronize(p->stream));
  return 0;
}

int bsq_plp_counters(const bsq_plp *p, int64_t *c, int n) {
  if (!p || !c) return BSQ_EINVAL;
  for (int i = 0; i < n && i < 8; ++i) c[i] = p->counters[i];
  return 0;
}

}  // extern "C"
