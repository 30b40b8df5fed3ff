// SMEM seeding on the device, third organisation: the passes of mem_collect_intv (lib/aln/memchain.c:50-106) as
// separate kernels with one uniform inner loop each, over a 32-byte FM-index block layout derived in HBM.
//
// Why.  k_seed2 (bsq_seed_dev.cuh) keeps the whole of mem_collect_intv -- forward sweeps, backward sweeps, re-seeding,
// greedy seeds -- in one per-lane state machine: ~630 warp instructions per extension step at 95 registers and 20
// warps/SM, bound by the length of the dependent chain per lane (profiles/README.md, r01 v7 / r02).  The reference's
// loops decompose into pieces that are independent of each other:
//   * the forward sweeps of pass 1 form one chain per task: the sweep of the next bwt_smem1a call starts where the
//     previous forward sweep stopped (ret = end of the longest match, bwt.c:343), whatever the backward sweep finds;
//   * every backward sweep (bwt.c:346-364) only needs the candidate list of its own forward sweep;
//   * pass 2 (memchain.c:76-85) is one more forward + backward sweep per long, rare SMEM of pass 1;
//   * pass 3 (bwt_seed_strategy1, bwt.c:376-396) only needs the read.
// and the final ks_introsort by (start, end) (memchain.c:105) makes the order of emission irrelevant (records with
// equal keys are the same substring, hence identical).  So:
//   k_s3_fwd<1>   one lane per task      forward sweeps of pass 1; candidate lists -> HBM, one work item per call
//   k_s3_bwd      one lane per call      backward sweep; SMEMs -> the task's interval list; long rare SMEMs of
//                                        pass 1 -> pass-2 items
//   k_s3_fwd<2>   one lane per item      forward sweep of a pass-2 call
//   k_s3_bwd      (again, pass-2 calls)
//   k_s3_greedy   one lane per task      pass 3
// Every kernel is a persistent loop "refill the lanes that ran out of work (divergent, rare) -- one bwt_extend for
// all lanes (convergent)"; the per-lane state is an interval and a few counters (<= 64 registers, 32 warps/SM).
//
// FM-index gathers.  The reference's block (bwt.h:93-101) is 64 bytes: u64 occ[4] + 128 symbols.  For the device a
// 32-byte block is derived from it at upload (k_derive_b32): three 40-bit cumulative counts + 64 symbols, i.e. exactly
// one DRAM sector, fetched with one 256-bit load (LDG.E.256); bwt_occ for the one symbol that is extended
// (bwt.c:278-293 needs the rank of c and the number of symbols > c in the interval) costs four 32-bit words per
// position instead of eight.  The ranks are the same integers by construction (checked against bwt_occ4 of the
// reference in tests/test_phase1.py).
#pragma once
#include "bsq_seed.h"

#if !defined(__CUDACC__)
// Test-only: tests/hostemu compiles these kernels with g++ and runs each of them as ONE sequential lane (a lane keeps
// pulling work until the queue is empty), so the decomposition can be checked against the oracle without a GPU.
#ifndef BSQ_SEED3_HOSTEMU
#error "bsq_seed3.cuh is device code; only tests/hostemu may compile it for the host"
#endif
struct uint4 { uint32_t x, y, z, w; };
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { uint4 v = {x, y, z, w}; return v; }
struct s3_dim_t { unsigned x; };
static s3_dim_t blockIdx, blockDim, threadIdx;
#define __global__
#define __device__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __grid_constant__
static inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { const unsigned long long o = *p; *p += v; return o; }
static inline int atomicAdd(int *p, int v) { const int o = *p; *p += v; return o; }
static inline unsigned long long atomicOr(unsigned long long *p, unsigned long long v) { const unsigned long long o = *p; *p |= v; return o; }
static inline int __popc(uint32_t v) { return __builtin_popcount(v); }
template <typename T> static inline T __ldg(const T *p) { return *p; }
static inline bool __all_sync(unsigned, bool p) { return p; }
#endif

#ifndef BSQ_S3_CTAS
#define BSQ_S3_CTAS 4  // resident CTAs of 128 threads per SM the seeding kernels run with.  Measured 8: 33.2, 7: 30.5, 6: 27.9, 5: 25.7, 4: 23.7, 3: 25.4 ms per 400 k tasks
                       // (profiles/kab_r02_l.jsonl, _m): the kernels sit at the random-sector rate of the DRAM (tools/micro/gather_peak.cu), and with fewer lanes in flight
                       // the candidate lists are re-read from L2 before the index traffic evicts them
#endif
#define BSQ_CTR_BLOCKS32 5  // 32-byte derived blocks fetched (distinct per extension)

// ---- derived block: w[0..2] = low words of S1,S2,S3 (S_c = number of symbols >= c before the block), w[3] = their
//      bits 32..39 (one byte each), w[4..7] = 64 symbols, first symbol in the top bits (as in the reference) ----
__global__ void k_derive_b32(const uint32_t *blocks, uint64_t n_sym, uint64_t n_half, uint32_t *b32) {
  const uint64_t h = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (h >= n_half) return;
  const uint32_t *src = blocks + (h >> 1) * 16;
  uint64_t occ[4];
  uint32_t sym[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) sym[j] = ((h >> 1) * 128 + 16 * (uint64_t)j < n_sym) ? src[8 + j] : 0u;  // the last block may be short
  if ((h >> 1) * 128 < n_sym) {
#pragma unroll
    for (int c = 0; c < 4; ++c) occ[c] = (uint64_t)src[2 * c] | (uint64_t)src[2 * c + 1] << 32;
  } else {  // padding block behind the text: never addressed
    occ[0] = occ[1] = occ[2] = occ[3] = 0;
  }
  if (h & 1) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint32_t w = sym[j], a = w & 0x55555555u, b = (w >> 1) & 0x55555555u;
      occ[1] += __popc(a & ~b); occ[2] += __popc(b & ~a); occ[3] += __popc(a & b);
    }
  }
  const uint64_t s3 = occ[3], s2 = s3 + occ[2], s1 = s2 + occ[1];
  uint32_t *o = b32 + h * 8;
  o[0] = (uint32_t)s1; o[1] = (uint32_t)s2; o[2] = (uint32_t)s3;
  o[3] = (uint32_t)(s1 >> 32) | (uint32_t)(s2 >> 32) << 8 | (uint32_t)(s3 >> 32) << 16;
  const int q = (h & 1) ? 4 : 0;
  o[4] = sym[q]; o[5] = sym[q + 1]; o[6] = sym[q + 2]; o[7] = sym[q + 3];
}

__device__ __forceinline__ void s3_ld256(const uint32_t *p, uint32_t (&w)[8]) {
#if defined(__CUDACC__)
  asm("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
      : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]) : "l"(p));
#else
  memcpy(w, p, 32);
#endif
}

// per-symbol constants of one extension: A/B select the symbol in the match mask, M switches the "> c" mask between
// "high bit" (c = 1) and "both bits" (c = 2); c = 0 and c = 3 are completed arithmetically
struct s3_sym_t {
  uint32_t A, B, M;
  int c;
};
__device__ __forceinline__ s3_sym_t s3_sym(int c) {
  s3_sym_t s;
  s.c = c;
  s.A = (c & 1) ? 0u : 0x55555555u;
  s.B = (c & 2) ? 0u : 0x55555555u;
  s.M = c == 1 ? 0x55555555u : 0u;
  return s;
}

// E = number of symbols == c, G = number of symbols > c among BWT[0..k2] ('$' already skipped in k2)
__device__ __forceinline__ void s3_rank(const uint32_t (&w)[8], uint64_t k2, const s3_sym_t &s, uint64_t &E, uint64_t &G) {
  const int n = (int)(k2 & 63) + 1;
  uint32_t e = 0, g = 0;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    int sh = 32 + 32 * j - 2 * n;  // bits of word j behind the prefix
    sh = sh < 0 ? 0 : sh;
    uint32_t keep;
#if defined(__CUDACC__)
    asm("shl.b32 %0, %1, %2;" : "=r"(keep) : "r"(0xffffffffu), "r"(sh));  // shift counts >= 32 give 0
#else
    keep = sh >= 32 ? 0u : 0xffffffffu << sh;
#endif
    const uint32_t a = w[4 + j] & keep & 0x55555555u, b = (w[4 + j] & keep & 0xaaaaaaaau) >> 1;
    e += __popc((a ^ s.A) & (b ^ s.B));
    g += __popc(b & (a | s.M));
  }
  const int c = s.c;
  if (c == 0) e -= (uint32_t)(64 - n);  // the cleared tail reads as symbol 0
  g = c == 0 ? (uint32_t)n - e : (c == 3 ? 0u : g);
  const uint64_t s1 = (uint64_t)w[0] | (uint64_t)(w[3] & 0xffu) << 32;
  const uint64_t s2 = (uint64_t)w[1] | (uint64_t)((w[3] >> 8) & 0xffu) << 32;
  const uint64_t s3 = (uint64_t)w[2] | (uint64_t)((w[3] >> 16) & 0xffu) << 32;
  const uint64_t tot = k2 & ~(uint64_t)63;  // symbols before the block
  const uint64_t hi = c == 0 ? tot : (c == 1 ? s1 : (c == 2 ? s2 : s3));   // symbols >= c
  const uint64_t lo = c == 0 ? s1 : (c == 1 ? s2 : (c == 2 ? s3 : 0ull));  // symbols > c
  E = hi - lo + e;
  G = lo + g;
}

// the index that is stepped by one kernel, in registers
struct s3_fm_t {
  const uint32_t *b32;
  uint64_t primary;
};

// bwt_extend (bwt.c:278-293) restricted to the child that is used: one backward step by symbol c in the stepped index;
// xa = interval start in the stepped index, xb = start in the other index.
__device__ __forceinline__ void s3_extend(const s3_fm_t &f, uint64_t L2c1 /* L2[c] + 1 */, uint64_t xa, uint64_t xb, uint64_t x2, int c,
                                          uint64_t &na, uint64_t &nb, uint64_t &o2) {
  const uint64_t k = xa - 1, l = k + x2;
  const uint64_t k2 = k - (k >= f.primary), l2 = l - (l >= f.primary);
  uint32_t wk[8], wl[8];
  s3_ld256(f.b32 + (k2 >> 6) * 8, wk);
  s3_ld256(f.b32 + (l2 >> 6) * 8, wl);
  BSQ_CTR(BSQ_CTR_EXTENDS, 1);
  BSQ_CTR(BSQ_CTR_BLOCKS32, 1 + ((k2 >> 6) != (l2 >> 6)));
  BSQ_CTR(BSQ_CTR_BLOCKS, 1 + ((k2 >> 7) != (l2 >> 7)));  // what bwt_2occ4 touches in the reference's 64-byte layout (N_occblk)
  const s3_sym_t s = s3_sym(c);
  uint64_t ek, gk, el, gl;
  s3_rank(wk, k2, s, ek, gk);
  s3_rank(wl, l2, s, el, gl);
  na = L2c1 + ek;
  o2 = el - ek;
  nb = xb + (xa <= f.primary && xa + x2 - 1 >= f.primary) + (gl - gk);
}

// ---- candidate record of a sweep (16 bytes): three 34-bit coordinates + the read end of the match ----
__device__ __forceinline__ uint4 s3_pack(uint64_t x0, uint64_t x1, uint64_t x2, int end) {
  return make_uint4((uint32_t)x0, (uint32_t)x1, (uint32_t)x2,
                    (uint32_t)(x0 >> 32) | (uint32_t)(x1 >> 32) << 4 | (uint32_t)(x2 >> 32) << 8 | (uint32_t)end << 12);
}
__device__ __forceinline__ uint64_t s3_x0(const uint4 &v) { return (uint64_t)v.x | (uint64_t)(v.w & 15u) << 32; }
__device__ __forceinline__ uint64_t s3_x1(const uint4 &v) { return (uint64_t)v.y | (uint64_t)((v.w >> 4) & 15u) << 32; }
__device__ __forceinline__ uint64_t s3_x2(const uint4 &v) { return (uint64_t)v.z | (uint64_t)((v.w >> 8) & 15u) << 32; }
__device__ __forceinline__ int s3_end(const uint4 &v) { return (int)(v.w >> 12); }

// one backward sweep to run: candidates cand[0 .. top) in push order (the last push is the longest match)
struct s3_call_t {
  uint32_t task;
  uint32_t xt;    // x | top << 9 | pass2 << 31
  uint64_t cand;  // first candidate, in records (40 bits) | min_intv << 40
};
// one pass-2 call to run (memchain.c:76-85)
struct s3_item_t {
  uint32_t task;
  uint32_t xm;  // x | min_intv << 9
};
// queue heads / fills, zeroed before every batch
struct s3_q_t {
  unsigned long long next_task, n_calls1, next_call1, n_items, next_item, cand2_used, next_task3, overflow, n_calls2, next_call2;
};

// base i of the read as it is searched: in-silico conversion of bseq_bsconvert (bwamem.c:161-178); 4 outside the read
__device__ __forceinline__ int s3_q(const uint8_t *row, int i, int par) {
  const int c = __ldg(row + i);
  return par ? (c == 1 ? 3 : c) : (c == 2 ? 0 : c);
}
__device__ __forceinline__ int s3_qs(const uint8_t *row, int i, int len, int par) { return i >= 0 && i < len ? s3_q(row, i, par) : 4; }

// Atomics whose result is not needed at once.  A lane that reserves a slot (interval list of a task, work queue) does
// so from divergent code; if the reservation were used on the spot the whole warp would sit out the round trip to L2
// in every iteration (measured: a third of the stall samples of the first version, profiles/README.md r02).  The
// reservation is issued, the record kept in registers, and the store done at the top of the next iteration, by which
// time the FM-index gather of that iteration has covered the latency.  (Inline PTX: the compiler would turn
// atomicAdd on a common address into a warp-aggregated one whose shuffle needs the result immediately.)
__device__ __forceinline__ int s3_reserve32(int32_t *p) {
#if defined(__CUDACC__)
  int r;
  asm volatile("atom.global.add.s32 %0, [%1], 1;" : "=r"(r) : "l"(p) : "memory");
  return r;
#else
  return (*p)++;
#endif
}
__device__ __forceinline__ unsigned long long s3_reserve64(unsigned long long *p) {
#if defined(__CUDACC__)
  unsigned long long r;
  asm volatile("atom.global.add.u64 %0, [%1], 1;" : "=l"(r) : "l"(p) : "memory");
  return r;
#else
  return (*p)++;
#endif
}

// an SMEM waiting for its slot in the task's interval list
struct s3_pend_t {
  bsq_pk_t rec;
  uint32_t task;
  int slot;
  bool on;
};
__device__ __forceinline__ void s3_pend_flush(s3_pend_t &pe, bsq_pk_t *intv) {
  if (pe.on) {
    if (pe.slot < BSQ_MAX_INTV) intv[(size_t)pe.task * BSQ_MAX_INTV + pe.slot] = pe.rec;
    pe.on = false;
  }
}

// an SMEM of the task [beg, end): reserve its slot (stored by s3_pend_flush); pass 1 also queues the re-seeding of long,
// rare SMEMs (memchain.c:79-81)
__device__ __forceinline__ void s3_emit(const bsq_devopt_t &opt, s3_pend_t &pe, bsq_pk_t *intv, int32_t *n_intv, uint32_t t, uint64_t x0, uint64_t x1,
                                        uint64_t x2, int beg, int end, bool pass1, s3_item_t *items, unsigned long long items_cap, s3_q_t *q) {
  if (end - beg < opt.min_seed_len) return;  // memchain.c:69-71
  s3_pend_flush(pe, intv);
  pe.slot = s3_reserve32(n_intv + t);
  pe.rec = bsq_pk_make(x0, x1, x2, beg, end);
  pe.task = t;
  pe.on = true;
  if (pass1 && end - beg >= opt.split_len && x2 <= (uint64_t)opt.split_width) {
    const unsigned long long k = atomicAdd(&q->n_items, 1ull);
    if (k < items_cap) { s3_item_t it; it.task = t; it.xm = (uint32_t)((beg + end) >> 1) | (uint32_t)(x2 + 1) << 9; items[k] = it; }
    else atomicOr(&q->overflow, 1ull);
  }
}

// Forward sweep(s) (bwt.c:324-343).  PASS 1: one lane per task runs the forward sweeps of all its bwt_smem1a calls
// back to back (memchain.c:65-73).  PASS 2: one lane per re-seeding item.  A call whose backward sweep is trivial
// (x == 0 or an ambiguous base at x - 1: every candidate stops at once and only the longest is kept,
// bwt.c:350-356) emits its SMEM here.  The base of the next step is loaded one step ahead.
template <int PASS>
__global__ void __launch_bounds__(128, BSQ_S3_CTAS) k_s3_fwd(const __grid_constant__ bsq_devopt_t opt, const __grid_constant__ bsq_devidx_t ix, int64_t n_tasks,
                                                   const uint8_t *seqs, int stride, const int32_t *lens, const uint8_t *parent, int pipeline,
                                                   uint4 *cand, unsigned long long cand_cap, const s3_item_t *items_in, s3_call_t *calls,
                                                   unsigned long long calls_cap, s3_item_t *items, unsigned long long items_cap, s3_q_t *q,
                                                   bsq_pk_t *intv, int32_t *n_intv) {
  unsigned long long *const n_calls = PASS == 1 ? &q->n_calls1 : &q->n_calls2;
  const int start_width = opt.self_ovlp ? 2 : 1;
  bool need = true, exhausted = false, triv = false;
  int64_t t = -1;
  int len = 0, par = 0, x = 0, i = 0, ik_end = 0, npush = 0, min_intv = 1, c_cur = 4, c_nx = 4;
  uint64_t ik0 = 0, ik1 = 0, ik2 = 0, cbase = 0;
  const uint8_t *row = nullptr;
  s3_fm_t f; f.b32 = nullptr; f.primary = 0;
  s3_pend_t pe; pe.on = false; pe.slot = 0; pe.task = 0; pe.rec.w0 = pe.rec.w1 = 0;
  // a call waiting for its slot in the queue
  bool cl_on = false;
  unsigned long long cl_k = 0;
  s3_call_t cl; cl.task = 0; cl.xt = 0; cl.cand = 0;
  unsigned long long n_in = 0;
  if (PASS == 2) n_in = q->n_items < items_cap ? q->n_items : items_cap;
  for (;;) {
    s3_pend_flush(pe, intv);
    if (cl_on) {
      if (cl_k < calls_cap) calls[cl_k] = cl;
      else atomicOr(&q->overflow, PASS == 1 ? 4ull : 8ull);
      cl_on = false;
    }
    if (need && !exhausted) {  // ---- refill (divergent): next call of the task, or the next task / item
      for (;;) {
        if (t < 0) {
          if (PASS == 1) {
            t = (int64_t)atomicAdd(&q->next_task, 1ull);
            if (t >= n_tasks) { exhausted = true; break; }
            len = lens[t]; x = 0; min_intv = start_width;
            if (pipeline && len < opt.min_seed_len) { t = -1; continue; }  // mem_chain returns before seeding (memchain.c:280)
            cbase = (uint64_t)t * (uint64_t)stride;
          } else {
            const unsigned long long k = atomicAdd(&q->next_item, 1ull);
            if (k >= n_in) { exhausted = true; break; }
            const s3_item_t it = items_in[k];
            t = it.task; len = lens[t]; x = (int)(it.xm & 511u); min_intv = (int)(it.xm >> 9);
            const unsigned long long need_rec = (unsigned long long)(len - x);
            cbase = atomicAdd(&q->cand2_used, need_rec);
            if (cbase + need_rec > cand_cap) { atomicOr(&q->overflow, 2ull); t = -1; continue; }
            cbase -= (uint64_t)x;  // pushes go to cbase + x + k like in pass 1
          }
          par = parent[t] != 0;
          row = seqs + (size_t)t * stride;
          f.b32 = ix.fm[!par].b32; f.primary = ix.fm[!par].primary;
        }
        int c = 4;
        if (PASS == 1) { while (x < len && (c = s3_q(row, x, par)) > 3) ++x; }
        else c = s3_q(row, x, par);  // inside an SMEM: never ambiguous
        if (x >= len || c > 3) { t = -1; continue; }
        ik0 = ix.fm[par].L2[c] + 1; ik2 = ix.fm[par].L2[c + 1] - ix.fm[par].L2[c]; ik1 = ix.fm[!par].L2[3 - c] + 1;  // bwt_set_intv
        ik_end = x + 1; i = x + 1; npush = 0;
        triv = x == 0 || s3_q(row, x - 1, par) > 3;
        c_cur = s3_qs(row, i, len, par); c_nx = s3_qs(row, i + 1, len, par);
        need = false;
        break;
      }
    }
    if (__all_sync(0xffffffffu, exhausted)) break;
    // ---- one forward step (convergent)
    const bool act = !need;
    const int c = act ? c_cur : 4;
    const bool ext = c <= 3;
    uint64_t o0 = 0, o1 = 0, o2 = 0;
    if (ext) s3_extend(f, ix.fm[!par].L2[3 - c] + 1, ik1, ik0, ik2, 3 - c, o1, o0, o2);
    if (act) {
      bool fin = false;
      if (!ext || o2 != ik2) {  // size change, read end or ambiguous base: the current match is a candidate
        cand[cbase + (uint64_t)(x + npush)] = s3_pack(ik0, ik1, ik2, ik_end);
        ++npush;
        fin = !ext || o2 < (uint64_t)min_intv;
      }
      if (!fin) {
        ik0 = o0; ik1 = o1; ik2 = o2; ik_end = i + 1; ++i;
        c_cur = c_nx; c_nx = s3_qs(row, i + 1, len, par);
      } else {
        if (triv) s3_emit(opt, pe, intv, n_intv, (uint32_t)t, ik0, ik1, ik2, x, ik_end, PASS == 1, items, items_cap, q);
        else {
          cl_k = s3_reserve64(n_calls);
          cl.task = (uint32_t)t; cl.cand = (cbase + (uint64_t)x) | (uint64_t)min_intv << 40;
          cl.xt = (uint32_t)x | (uint32_t)npush << 9 | (PASS == 2 ? 1u << 31 : 0u);
          cl_on = true;
        }
        need = true;
        if (PASS == 1) x = ik_end; else t = -1;  // bwt.c:343: the next call starts where this sweep stopped
      }
    }
  }
}

// Backward sweep of one call (bwt.c:346-364): candidates, longest match first, are extended to the left column by
// column; the list is compacted in place (entry n_curr <= j is written after entry j was read).  The next candidate
// of a column is loaded one step ahead, the first candidate of the next column is the first record kept in this one
// (held in registers), and the base of the next column is loaded when the current column starts.
__global__ void __launch_bounds__(128, BSQ_S3_CTAS) k_s3_bwd(const __grid_constant__ bsq_devopt_t opt, const __grid_constant__ bsq_devidx_t ix,
                                                   const uint8_t *seqs, int stride, const uint8_t *parent, uint4 *cand, const s3_call_t *calls,
                                                   unsigned long long calls_cap, int pass, s3_item_t *items,
                                                   unsigned long long items_cap, s3_q_t *q, bsq_pk_t *intv, int32_t *n_intv) {
  unsigned long long *const next_call = pass == 1 ? &q->next_call1 : &q->next_call2;
  const unsigned long long n_made = pass == 1 ? q->n_calls1 : q->n_calls2;
  bool need = true, exhausted = false, pass1 = true, any = false;
  uint32_t t = 0;
  int par = 0, i = 0, j = 0, n_prev = 0, n_curr = 0, top = 0, last_beg = 0, min_intv = 1, c = 4, c_nx = 4;
  uint64_t last_x2 = 0;
  uint4 *lst = nullptr;
  const uint8_t *row = nullptr;
  s3_fm_t f; f.b32 = nullptr; f.primary = 0;
  uint4 nxt = make_uint4(0, 0, 0, 0), first = make_uint4(0, 0, 0, 0);
  bool nxt_ok = false;
  s3_pend_t pe; pe.on = false; pe.slot = 0; pe.task = 0; pe.rec.w0 = pe.rec.w1 = 0;
  const unsigned long long n_calls = n_made < calls_cap ? n_made : calls_cap;
  for (;;) {
    s3_pend_flush(pe, intv);
    if (need && !exhausted) {
      const unsigned long long k = atomicAdd(next_call, 1ull);
      if (k >= n_calls) exhausted = true;
      else {
        const s3_call_t cl = calls[k];
        t = cl.task;
        const int x = (int)(cl.xt & 511u);
        top = (int)((cl.xt >> 9) & 511u); min_intv = (int)(cl.cand >> 40); pass1 = (cl.xt >> 31) == 0;
        lst = cand + (cl.cand & ((1ull << 40) - 1));
        par = parent[t] != 0;
        row = seqs + (size_t)t * stride;
        f.b32 = ix.fm[par].b32; f.primary = ix.fm[par].primary;
        i = x - 1; j = 0; n_prev = top; n_curr = 0; any = false; nxt_ok = false;
        c = s3_q(row, i, par);  // x >= 1: trivial calls never get here
        c_nx = i > 0 ? s3_q(row, i - 1, par) : 4;
        need = false;
      }
    }
    if (__all_sync(0xffffffffu, exhausted)) break;
    const bool act = !need;
    uint4 p = nxt;
    if (act && !nxt_ok) p = lst[top - 1 - j];
    if (act) {  // entry j + 1 is untouched by this step
      nxt_ok = j + 1 < n_prev;
      if (nxt_ok) nxt = lst[top - 2 - j];
    }
    const uint64_t x0 = s3_x0(p), x1 = s3_x1(p), x2 = s3_x2(p);
    const bool ext = act && c <= 3;
    uint64_t o0 = 0, o1 = 0, o2 = 0;
    if (ext) s3_extend(f, ix.fm[par].L2[c] + 1, x0, x1, x2, c, o0, o1, o2);
    if (act) {
      if (!ext || o2 < (uint64_t)min_intv) {  // cannot be extended: an SMEM unless contained in a longer one (bwt.c:350-356)
        if (n_curr == 0 && (!any || i + 1 < last_beg)) {
          s3_emit(opt, pe, intv, n_intv, t, x0, x1, x2, i + 1, s3_end(p), pass1, items, items_cap, q);
          last_beg = i + 1; any = true;
        }
      } else if (n_curr == 0 || o2 != last_x2) {
        const uint4 rec = s3_pack(o0, o1, o2, s3_end(p));
        lst[top - 1 - n_curr] = rec;
        if (n_curr == 0) first = rec;
        ++n_curr; last_x2 = o2;
      }
      ++j;
      if (j == n_prev) {  // next column (bwt.c:362-363)
        if (n_curr == 0) need = true;
        else {
          n_prev = n_curr; n_curr = 0; --i; j = 0;
          nxt = first; nxt_ok = true;
          c = c_nx; c_nx = i > 0 ? s3_q(row, i - 1, par) : 4;
        }
      }
    }
  }
}

// Pass 3: greedy forward seeds (memchain.c:88-103 over bwt_seed_strategy1, bwt.c:376-396).
__global__ void __launch_bounds__(128, BSQ_S3_CTAS) k_s3_greedy(const __grid_constant__ bsq_devopt_t opt, const __grid_constant__ bsq_devidx_t ix, int64_t n_tasks,
                                                      const uint8_t *seqs, int stride, const int32_t *lens, const uint8_t *parent, int pipeline,
                                                      s3_q_t *q, bsq_pk_t *intv, int32_t *n_intv) {
  bool need = true, exhausted = false;
  int64_t t = -1;
  int len = 0, par = 0, x = 0, i = 0, c_cur = 4, c_nx = 4;
  uint64_t ik0 = 0, ik1 = 0, ik2 = 0;
  const uint8_t *row = nullptr;
  s3_fm_t f; f.b32 = nullptr; f.primary = 0;
  s3_pend_t pe; pe.on = false; pe.slot = 0; pe.task = 0; pe.rec.w0 = pe.rec.w1 = 0;
  const int min_seed_len = opt.min_seed_len, max_mem_intv = opt.max_mem_intv;
  if (max_mem_intv <= 0) return;
  for (;;) {
    s3_pend_flush(pe, intv);
    if (need && !exhausted) {
      for (;;) {
        if (t < 0) {
          t = (int64_t)atomicAdd(&q->next_task3, 1ull);
          if (t >= n_tasks) { exhausted = true; break; }
          len = lens[t]; x = 0;
          if (pipeline && len < min_seed_len) { t = -1; continue; }
          par = parent[t] != 0;
          row = seqs + (size_t)t * stride;
          f.b32 = ix.fm[!par].b32; f.primary = ix.fm[!par].primary;
        }
        int c = 4;
        while (x < len && (c = s3_q(row, x, par)) > 3) ++x;
        if (x >= len) { t = -1; continue; }
        ik0 = ix.fm[par].L2[c] + 1; ik2 = ix.fm[par].L2[c + 1] - ix.fm[par].L2[c]; ik1 = ix.fm[!par].L2[3 - c] + 1;
        i = x + 1;
        c_cur = s3_qs(row, i, len, par); c_nx = s3_qs(row, i + 1, len, par);
        need = false;
        break;
      }
    }
    if (__all_sync(0xffffffffu, exhausted)) break;
    const bool act = !need;
    const int c = act ? c_cur : 4;
    const bool ext = c <= 3;
    uint64_t o0 = 0, o1 = 0, o2 = 0;
    if (ext) s3_extend(f, ix.fm[!par].L2[3 - c] + 1, ik1, ik0, ik2, 3 - c, o1, o0, o2);
    if (act) {
      if (!ext) { x = i == len ? len : i + 1; need = true; }  // read end / ambiguous base: no seed from x (bwt.c:393-395)
      else if (o2 < (uint64_t)max_mem_intv && i - x >= min_seed_len) {
        if (o2 > 0) {  // memchain.c:95
          pe.slot = s3_reserve32(n_intv + t);
          pe.rec = bsq_pk_make(o0, o1, o2, x, i + 1);
          pe.task = (uint32_t)t;
          pe.on = true;
        }
        x = i + 1; need = true;
      } else {
        ik0 = o0; ik1 = o1; ik2 = o2; ++i;
        c_cur = c_nx; c_nx = s3_qs(row, i + 1, len, par);
      }
    }
  }
}
