// k_seed: SMEM seeding on the device, second organisation of the state machine of bsq_seed.h.
//
// Same algorithm and the same results as bsq_sm_next / bsq_sm_consume (= mem_collect_intv,
// lib/aln/memchain.c:50-106, over bwt_smem1a lib/aln/bwt.c:307-370 and bwt_seed_strategy1 bwt.c:376-396); the host
// emulation keeps using bsq_seed.h, this file is what runs on the GPU.  Profiling the first organisation showed the
// kernel bound by instruction issue, with 55 % of the warp instructions in per-state control code executed by ~7 of 32
// lanes (profiles/README.md): lanes of a warp are in different states (forward sweep, backward sweep, greedy seeds),
// and every state's code path was issued separately each round.  Here
//   * the three stepping states share one request / extension / consume path, written with selects rather than
//     branches wherever the states differ only in data (which interval, which direction, which symbol);
//   * only the rare transitions (start of a sweep, end of a sweep, next task) are separate, divergent code;
//   * the forward sweep's candidate list is never reversed: the backward sweep reads it through an index transform,
//     and the most recent CAP pushes (the longest matches, which the backward sweep works on) stay in shared memory
//     as a ring, older ones move to local memory.
#pragma once
#include "bsq_seed.h"

enum { S2_IDLE = 0, S2_FWD = 1, S2_BWD = 2, S2_S1 = 3, S2_NEXT = 4, S2_FINFWD = 5, S2_FINBWD = 6 };

template <int CAP>
struct seed2_list_t {
  uint4 *sm;  // this lane's ring; slot s at sm[s * 32]
  bsq_pk_t spill[BSQ_MAX_READ_LEN + 1];
  __device__ __forceinline__ static bsq_pk_t from4(const uint4 v) {
    bsq_pk_t p;
    p.w0 = (uint64_t)v.x | (uint64_t)v.y << 32; p.w1 = (uint64_t)v.z | (uint64_t)v.w << 32;
    return p;
  }
  __device__ __forceinline__ static uint4 to4(const bsq_pk_t &p) {
    return make_uint4((uint32_t)p.w0, (uint32_t)(p.w0 >> 32), (uint32_t)p.w1, (uint32_t)(p.w1 >> 32));
  }
  // forward sweep: push number k (0, 1, 2, ...)
  __device__ __forceinline__ void push(int k, const bsq_pk_t &p) {
    uint4 *slot = sm + (k & (CAP - 1)) * 32;
    if (k >= CAP) spill[k - CAP] = from4(*slot);  // push k - CAP leaves the ring
    *slot = to4(p);
  }
  // backward sweep: logical entry L of a list whose forward sweep made `top` pushes (entry 0 = last push)
  __device__ __forceinline__ bsq_pk_t get(int top, int L) const {
    const int k = top - 1 - L;
    if (L < CAP) return from4(sm[(k & (CAP - 1)) * 32]);
    return spill[k];
  }
  __device__ __forceinline__ void set(int top, int L, const bsq_pk_t &p) {
    const int k = top - 1 - L;
    if (L < CAP) sm[(k & (CAP - 1)) * 32] = to4(p);
    else spill[k] = p;
  }
};

// V = 1 (default) adds four things the source-level profile of V = 0 asked for (profiles/README.md, r01 v7: only a
// fifth of the stall samples sat on the FM-index gathers, a third on candidate-list / interval-array accesses in
// code executed by one to three lanes):
//   * the SMEMs of one bwt_smem1a call are left in emission order -- k_seed_sort orders the whole list by (start, end)
//     afterwards and records with equal keys are identical (same substring, same bi-interval), so the reversal of
//     bwt.c:365 (a single-lane loop over global memory) changes nothing;
//   * pass 2 (memchain.c:76-85) takes its re-seeding positions from a four-entry queue filled when pass 1 stores a
//     long, rare SMEM, instead of re-reading every pass-1 interval from global memory (falls back to the scan when
//     more than four qualify);
//   * the next candidate of a backward column is loaded one step ahead (entries beyond the shared-memory ring
//     live in local memory, and its address feeds the next gather);
//   * reads whose row is not 8-byte aligned are converted from aligned 8-byte words (funnel shift), not bytewise.
// (L2 prefetches of the next candidate's blocks were tried -- prefetch.global.L2 and LDGSTS into a scratch slot --
// and measured slower: 67.5 / 56 ms against 47.5 ms.)
template <int CAP, int V>
__global__ void __launch_bounds__(128, BSQ_SEED_CTAS) k_seed2(const __grid_constant__ bsq_devopt_t opt, const __grid_constant__ bsq_devidx_t ix, int64_t n_tasks,
                                                              const uint8_t *seqs, int stride, const int32_t *lens, const uint8_t *parent, int pipeline,
                                                              bsq_pk_t *intv, int32_t *n_intv, int32_t *status, unsigned long long *next_task) {
  extern __shared__ uint4 seed_smem[];
  seed2_list_t<CAP> lst;
  lst.sm = seed_smem + (threadIdx.x >> 5) * (CAP * 32) + (threadIdx.x & 31);
  uint32_t *rd = reinterpret_cast<uint32_t *>(seed_smem + 128 * CAP) + threadIdx.x;  // converted read, [word][thread]
  // options
  const int min_seed_len = opt.min_seed_len, split_len = opt.split_len, split_width = opt.split_width;
  const int start_width = opt.self_ovlp ? 2 : 1, max_mem_intv = opt.max_mem_intv;
  // lane state
  int st = S2_IDLE, pass = 1, len = 0, par = 0;
  int x = 0, i = 0, j = 0, k2 = 0, old_n = 0;
  int n_curr = 0, n_prev = 0, top = 0, n_tmp = 0, n_keep = 0, n_out = 0, min_intv = 1, ret = 0, overflow = 0;
  uint64_t ik0 = 0, ik1 = 0, ik2 = 0, last_x2 = 0;
  int ik_end = 0, last_beg = 0;
  int64_t t = -1;
  bool exhausted = false;
  bsq_pk_t *out = nullptr;
  // V = 1: pass-2 queue (16 bits per entry: position << 7 | min_intv), preloaded candidate j + 1
  uint64_t q2 = 0;
  int q2n = 0;      // entries queued; -1: more than four qualified (or do not fit), pass 2 scans the list
  bsq_pk_t nxt; nxt.w0 = nxt.w1 = 0;
  bool nxt_ok = false;  // nxt holds entry j of the current column (loaded while entry j - 1 was extended)

#define S2_Q(pos) ((int)(rd[((pos) >> 3) * 128] >> (((pos) & 7) * 4)) & 0xf)
  // a candidate cannot be extended further to the left (bwt.c:350-356).  n_tmp counts every SMEM of this call like the
  // reference's mem vector does; only those of at least min_seed_len (memchain.c:69-71) are stored (n_keep of them)
#define S2_BWD_STOP(X0, X1, X2, END)                                                        \
  do {                                                                                      \
    if (n_curr == 0 && (n_tmp == 0 || i + 1 < last_beg)) {                                  \
      if (n_out + n_tmp >= BSQ_MAX_INTV) overflow = 1;                                      \
      else {                                                                                \
        if ((END) - (i + 1) >= min_seed_len) out[n_out + n_keep++] = bsq_pk_make((X0), (X1), (X2), i + 1, (END)); \
        if (V && pass == 1 && (END) - (i + 1) >= min_seed_len && (END) - (i + 1) >= split_len && (X2) <= (uint64_t)split_width) { \
          if (q2n >= 0 && q2n < 4 && (X2) < 127) { q2 |= (uint64_t)(((i + 1 + (END)) >> 1) << 7 | (int)((X2) + 1)) << (16 * q2n); ++q2n; } \
          else q2n = -1;                                                                    \
        }                                                                                   \
        last_beg = i + 1;                                                                   \
        ++n_tmp;                                                                            \
      }                                                                                     \
    }                                                                                       \
  } while (0)
  // one candidate of the current column done (bwt.c:362-363)
#define S2_COL_STEP()                                                                       \
  do {                                                                                      \
    ++j;                                                                                    \
    if (j == n_prev) {                                                                      \
      if (n_curr == 0) st = S2_FINBWD;                                                      \
      else { n_prev = n_curr; n_curr = 0; --i; j = 0; if (i < -1) st = S2_FINBWD; }         \
    }                                                                                       \
  } while (0)

  for (;;) {
    // ---------------- rare transitions (divergent) ----------------
    if (st == S2_IDLE || st >= S2_NEXT) {
      if (st == S2_FINFWD) {  // end of the forward sweep (bwt.c:340-345); the list is read backwards from here on
        ret = ik_end;  // read end of the last candidate pushed
        top = n_curr; n_prev = n_curr; n_curr = 0;
        i = x - 1; j = 0; n_tmp = 0; n_keep = 0;
        nxt_ok = false;
        st = S2_BWD;
      } else if (st == S2_FINBWD) {  // end of one bwt_smem1a call: the kept SMEMs ordered by start (memchain.c:69-71)
        if (!V) {
          bsq_pk_t *tt = out + n_out;
          for (int a = 0, b = n_keep - 1; a < b; ++a, --b) { const bsq_pk_t s_ = tt[a]; tt[a] = tt[b]; tt[b] = s_; }
        }
        n_out += n_keep; n_tmp = 0; n_keep = 0;
        if (pass == 1) x = ret;
        st = S2_NEXT;
      }
      if (st == S2_IDLE && !exhausted) {  // next task
        t = (int64_t)atomicAdd(next_task, 1ull);
        if (t >= n_tasks) exhausted = true;
        else {
          len = lens[t];
          par = parent[t] != 0;
          out = intv + t * BSQ_MAX_INTV;
          if (pipeline && len < min_seed_len) n_intv[t] = 0;  // mem_chain returns before seeding
          else {
            // the read, converted for the index it is searched in (bseq_bsconvert, bwamem.c:161-178), 4 bits per base
            const uint8_t *s = seqs + t * stride;
            const int nw = (len + 7) >> 3;
            const bool al8 = ((uintptr_t)s & 7) == 0;
            const uint64_t *s8 = reinterpret_cast<const uint64_t *>((uintptr_t)s & ~(uintptr_t)7);
            const int sh = (int)((uintptr_t)s & 7);  // bytes of the first aligned word that precede the row
            q2 = 0; q2n = 0; nxt_ok = false;
            for (int w = 0; w < nw; ++w) {
              uint64_t v;
              if (al8) v = __ldg(reinterpret_cast<const uint64_t *>(s) + w);
              else if (V) {  // bases 8w .. 8w+7 from two aligned words; the second only if it holds a base of this row
                v = __ldg(s8 + w) >> (8 * sh);
                if (8 * w + 8 - sh < len) v |= __ldg(s8 + w + 1) << (64 - 8 * sh);
              } else { v = 0; for (int k = 0; k < 8; ++k) if (8 * w + k < len) v |= (uint64_t)s[8 * w + k] << (8 * k); }
              uint32_t pk = 0;
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                int c = (int)(v >> (8 * k)) & 0xf;
                c = par ? (c == 1 ? 3 : c) : (c == 2 ? 0 : c);
                pk |= (uint32_t)c << (4 * k);
              }
              rd[w * 128] = pk;
            }
            pass = 1; x = 0; i = j = k2 = old_n = 0;
            n_curr = n_prev = top = n_tmp = n_keep = n_out = 0; min_intv = 1; ret = 0; overflow = 0;
            st = S2_NEXT;
          }
        }
      }
      if (st == S2_NEXT) {
        const bsq_fm_t &fm = ix.fm[par], &fmc = ix.fm[!par];
        bool done = overflow != 0;
        int c = 4, sx = -1, smin = 1;  // sx >= 0: start bwt_smem1a at sx
        if (!done && pass == 1) {  // every SMEM (memchain.c:65-73)
          while (x < len && (c = S2_Q(x)) > 3) ++x;
          if (x >= len) { pass = 2; old_n = n_out; k2 = 0; }
          else { sx = x; smin = start_width; }
        }
        if (V && !done && pass == 2 && sx < 0 && q2n >= 0) {  // re-seeding positions queued by pass 1
          if (q2n > 0) {
            const int e = (int)(q2 & 0xffff);
            q2 >>= 16; --q2n;
            sx = e >> 7; smin = e & 127; c = S2_Q(sx);
          } else { pass = 3; x = 0; }
        } else if (!done && pass == 2 && sx < 0) {  // re-seed from the middle of long, rare SMEMs (memchain.c:76-85)
          while (k2 < old_n) {
            const bsq_pk_t p = out[k2++];
            const int start = bsq_pk_beg(p), end = bsq_pk_end(p);
            if (end - start < split_len || bsq_pk_x2(p) > (uint64_t)split_width) continue;
            sx = (start + end) >> 1; c = S2_Q(sx); smin = (int)(bsq_pk_x2(p) + 1);
            break;
          }
          if (sx < 0) { pass = 3; x = 0; }
        }
        if (!done && sx >= 0) {  // begin bwt_smem1a at sx (bwt.c:313-322)
          n_tmp = 0; n_keep = 0; n_curr = 0; x = sx;
          if (c > 3) ret = x + 1;  // not reachable for positions inside a read / an SMEM; kept for symmetry with bsq_seed.h
          else {
            min_intv = smin < 1 ? 1 : smin;
            ik0 = fm.L2[c] + 1; ik2 = fm.L2[c + 1] - fm.L2[c]; ik1 = fmc.L2[3 - c] + 1;
            ik_end = x + 1; i = x + 1;
            st = S2_FWD;
          }
        } else if (!done && pass == 3) {  // greedy forward seeds (memchain.c:88-103)
          if (max_mem_intv <= 0) done = true;
          else {
            while (x < len && (c = S2_Q(x)) > 3) ++x;
            if (x >= len) done = true;
            else {
              ik0 = fm.L2[c] + 1; ik2 = fm.L2[c + 1] - fm.L2[c]; ik1 = fmc.L2[3 - c] + 1;
              i = x + 1;
              st = S2_S1;
            }
          }
        }
        if (done) {
          if (overflow) { atomicOr(status, 1); n_intv[t] = 0; }
          else n_intv[t] = n_out;
          st = S2_IDLE;
        }
      }
    }
    if (__all_sync(0xffffffffu, exhausted && st == S2_IDLE)) break;

    // ---------------- one extension step, shared by the three stepping states ----------------
    const bool act = st >= S2_FWD && st <= S2_S1;
    const bool isb = st == S2_BWD;
    int c = 4;
    if (act && i >= 0 && i < len) c = S2_Q(i);
    uint64_t x0 = ik0, x1 = ik1, x2 = ik2;
    int p_end = 0;
    if (isb) {
      bsq_pk_t p;
      if (V && nxt_ok) p = nxt; else p = lst.get(top, j);
      x0 = bsq_pk_x0(p); x1 = bsq_pk_x1(p); x2 = bsq_pk_x2(p); p_end = bsq_pk_end(p);
      if (V) {  // entry j + 1 is untouched by this step: the in-place compaction writes entries <= j
        nxt_ok = j + 1 < n_prev;  // false at the last candidate of a column, so a new column / sweep starts clean
        if (nxt_ok) nxt = lst.get(top, j + 1);
      }
    }
    const bool issue = act && c <= 3;
    if (act && !issue) {  // read end, read start or an ambiguous base: no extension
      if (st == S2_FWD) {  // closes the sweep (bwt.c:335-340)
        lst.push(n_curr++, bsq_pk_make(ik0, ik1, ik2, 0, ik_end));
        st = S2_FINFWD;
      } else if (st == S2_S1) { x = i == len ? len : i + 1; st = S2_NEXT; }
      else { S2_BWD_STOP(x0, x1, x2, p_end); S2_COL_STEP(); }
    }
    if (issue) {
      bsq_ext_req_t req;
      req.x0 = x0; req.x1 = x1; req.x2 = x2; req.back = isb; req.c = isb ? c : 3 - c;
      uint64_t o0, o1, o2;
      bsq_extend1(ix.fm[par], ix.fm[!par], req, o0, o1, o2);
      if (isb) {  // bwt.c:349-360
        if (o2 < (uint64_t)min_intv) S2_BWD_STOP(x0, x1, x2, p_end);
        else if (n_curr == 0 || o2 != last_x2) {
          lst.set(top, n_curr++, bsq_pk_make(o0, o1, o2, 0, p_end));  // entry n_curr <= j: in-place compaction
          last_x2 = o2;
        }
        S2_COL_STEP();
      } else {
        bool stop;
        if (st == S2_FWD) {  // bwt.c:326-334
          stop = false;
          if (o2 != ik2) {
            lst.push(n_curr++, bsq_pk_make(ik0, ik1, ik2, 0, ik_end));
            stop = o2 < (uint64_t)min_intv;
          }
          if (stop) st = S2_FINFWD; else ik_end = i + 1;
        } else {  // S2_S1, bwt.c:387-392
          stop = o2 < (uint64_t)max_mem_intv && i - x >= min_seed_len;
          if (stop) {
            if (o2 > 0) {  // memchain.c:95
              if (n_out >= BSQ_MAX_INTV) overflow = 1;
              else out[n_out++] = bsq_pk_make(o0, o1, o2, x, i + 1);
            }
            x = i + 1;
            st = S2_NEXT;
          }
        }
        if (!stop) { ik0 = o0; ik1 = o1; ik2 = o2; ++i; }
      }
    }
  }
#undef S2_Q
#undef S2_BWD_STOP
#undef S2_COL_STEP
}
