// SMEM seeding for one (read, conversion) task: the three passes of mem_collect_intv
// (lib/aln/memchain.c:50-106) over bwt_smem1a (lib/aln/bwt.c:307-370) and bwt_seed_strategy1
// (lib/aln/bwt.c:376-396), re-organised for SIMT execution.
//
// The reference's loops are turned inside out into a resumable state machine whose only
// suspension point is "I need one bwt_extend": next() runs the cheap control logic up to the
// next extension request, the caller performs the extension (the random 64-byte FM-index
// gathers -- the one long-latency operation) and hands the result to consume().  In the CUDA
// kernel all 32 lanes of a warp therefore meet at a single extend site every iteration and
// their DRAM round trips overlap, instead of each lane serialising its own (v1 measured 4 of
// 32 lanes active, profiles/README.md).  Candidate intervals are kept as 16-byte packed
// records (3 x 34-bit positions + 2 x 9-bit read coordinates) to keep the thread-local scratch
// that spills through L2 small.
#pragma once
#include "bsq_fm.h"
#include "bsq_sort.h"

// ---- packed bi-interval: x0,x1,x2 < 2^34, beg,end < 2^9 ----
struct bsq_pk_t {
  uint64_t w0, w1;
};
BSQ_HD bsq_pk_t bsq_pk_make(uint64_t x0, uint64_t x1, uint64_t x2, int beg, int end) {
  bsq_pk_t p;
  p.w0 = x0 | (x1 << 34);
  p.w1 = (x1 >> 30) | (x2 << 4) | ((uint64_t)(uint32_t)beg << 38) | ((uint64_t)(uint32_t)end << 47);
  return p;
}
BSQ_HD uint64_t bsq_pk_x0(const bsq_pk_t &p) { return p.w0 & 0x3FFFFFFFFull; }
BSQ_HD uint64_t bsq_pk_x1(const bsq_pk_t &p) { return (p.w0 >> 34) | ((p.w1 & 0xF) << 30); }
BSQ_HD uint64_t bsq_pk_x2(const bsq_pk_t &p) { return (p.w1 >> 4) & 0x3FFFFFFFFull; }
BSQ_HD int bsq_pk_beg(const bsq_pk_t &p) { return (int)((p.w1 >> 38) & 0x1FF); }
BSQ_HD int bsq_pk_end(const bsq_pk_t &p) { return (int)((p.w1 >> 47) & 0x1FF); }
BSQ_HD uint64_t bsq_pk_info(const bsq_pk_t &p) { return (uint64_t)bsq_pk_beg(p) << 32 | (uint32_t)bsq_pk_end(p); }
BSQ_HD bsq_intv_t bsq_pk_unpack(const bsq_pk_t &p) {
  bsq_intv_t v;
  v.x[0] = bsq_pk_x0(p); v.x[1] = bsq_pk_x1(p); v.x[2] = bsq_pk_x2(p); v.info = bsq_pk_info(p);
  return v;
}

// Per-task scratch of the state machine.  A scratch type provides
//   get(i) / set(i, v)  the candidate list (the reference's tmpvec[0..1]; ONE list here: the backward
//                       phase compacts it in place, entry n_curr <= j is only written after entry j was read)
//   q(i)                base i of the read after in-silico bisulfite conversion (C>T parent, G>A daughter)
// The plain version below is used by the host emulation; k_seed keeps the first entries of each lane's
// list in shared memory ([entry][lane], conflict-free 16-byte accesses) and spills the rest to local memory.
struct bsq_seed_scratch_t {
  bsq_pk_t a[BSQ_MAX_READ_LEN + 1];
  const uint8_t *seq;
  int parent;
  BSQ_HD void bind(const uint8_t *s, int par) { seq = s; parent = par; }
  BSQ_HD bsq_pk_t get(int i) const { return a[i]; }
  BSQ_HD void set(int i, const bsq_pk_t &v) { a[i] = v; }
  BSQ_HD int q(int i) const {
    const int c = seq[i];
    return parent ? (c == 1 ? 3 : c) : (c == 2 ? 0 : c);
  }
};

// one pending bwt_extend: interval (x0,x1,x2), direction, and the symbol whose child is wanted
struct bsq_ext_req_t {
  uint64_t x0, x1, x2;
  int back, c;
};

// bwt_extend (bwt.c:278-293) restricted to the one child that is used.  back=1 extends to the
// left in `fm`; back=0 is the forward extension = a backward step in the complementary index.
BSQ_HD uint64_t bsq_sel4(int c, uint64_t a0, uint64_t a1, uint64_t a2, uint64_t a3) {  // a[c] without indexing (keeps a[] in registers)
  const uint64_t lo = (c & 1) ? a1 : a0, hi = (c & 1) ? a3 : a2;
  return (c & 2) ? hi : lo;
}

BSQ_HD void bsq_extend1(const bsq_fm_t &fm, const bsq_fm_t &fmc, const bsq_ext_req_t &r, uint64_t &o0, uint64_t &o1, uint64_t &o2) {
  const bsq_fm_t &f = r.back ? fm : fmc;
  const uint64_t xa = r.back ? r.x0 : r.x1;  // coordinate in the index being stepped
  const uint64_t xb = r.back ? r.x1 : r.x0;  // coordinate in the other index
  uint64_t tk[4], tl[4];
  BSQ_CTR(BSQ_CTR_EXTENDS, 1);
  bsq_2occ4_flat(f, xa - 1, xa - 1 + r.x2, tk, tl);
  const int c = r.c;
  const uint64_t tkc = bsq_sel4(c, tk[0], tk[1], tk[2], tk[3]), tlc = bsq_sel4(c, tl[0], tl[1], tl[2], tl[3]);
  const uint64_t na = bsq_sel4(c, f.L2[0], f.L2[1], f.L2[2], f.L2[3]) + 1 + tkc;
  uint64_t nb = xb + (xa <= f.primary && xa + r.x2 - 1 >= f.primary);
#pragma unroll
  for (int s = 3; s > 0; --s)
    if (s > c) nb += tl[s] - tk[s];
  o2 = tlc - tkc;
  if (r.back) { o0 = na; o1 = nb; } else { o1 = na; o0 = nb; }
}

enum { BSQ_ST_NEXT = 0, BSQ_ST_FWD, BSQ_ST_BWD, BSQ_ST_S1, BSQ_ST_DONE };

struct bsq_seed_machine_t {
  // immutable per task
  int len, cap;
  int min_seed_len, split_len, split_width, start_width, max_mem_intv;
  // state
  int st, pass, x, i, j, k2, old_n;
  int n_curr, n_prev, n_tmp, n_out, min_intv, ret, overflow;
  uint64_t ik0, ik1, ik2, last_x2;  // last_x2: x[2] of the candidate most recently kept in the backward phase
  int ik_end, p_end, last_beg;      // p_end: read end of the candidate whose extension is pending; last_beg: start of the SMEM emitted last
};

BSQ_HD void bsq_sm_init(bsq_seed_machine_t &m, const bsq_devopt_t &opt, int len, int cap) {
  m.len = len; m.cap = cap;
  m.min_seed_len = opt.min_seed_len; m.split_len = opt.split_len; m.split_width = opt.split_width;
  m.start_width = opt.self_ovlp ? 2 : 1; m.max_mem_intv = opt.max_mem_intv;
  m.st = BSQ_ST_NEXT; m.pass = 1; m.x = 0; m.i = m.j = m.k2 = m.old_n = 0;
  m.n_curr = m.n_prev = m.n_tmp = m.n_out = 0; m.min_intv = 1; m.ret = 0; m.overflow = 0;
  m.ik0 = m.ik1 = m.ik2 = m.last_x2 = 0; m.ik_end = m.p_end = m.last_beg = 0;
}

// begin bwt_smem1a at query position x (bwt.c:313-322)
BSQ_HD void bsq_sm_start_smem(bsq_seed_machine_t &m, const bsq_fm_t &fm, const bsq_fm_t &fmc, int c, int x, int min_intv) {
  m.n_tmp = 0; m.n_curr = 0;
  m.x = x;
  if (c > 3) { m.ret = x + 1; m.st = BSQ_ST_NEXT; return; }
  m.min_intv = min_intv < 1 ? 1 : min_intv;
  m.ik0 = fm.L2[c] + 1; m.ik2 = fm.L2[c + 1] - fm.L2[c]; m.ik1 = fmc.L2[3 - c] + 1;
  m.ik_end = x + 1; m.i = x + 1;
  m.st = BSQ_ST_FWD;
}

// end of the forward sweep: longest matches first, then walk left (bwt.c:340-345)
template <class S>
BSQ_HD void bsq_sm_finish_fwd(bsq_seed_machine_t &m, S &scr) {
  for (int a = 0, b = m.n_curr - 1; a < b; ++a, --b) { const bsq_pk_t t = scr.get(a), u = scr.get(b); scr.set(a, u); scr.set(b, t); }
  m.ret = m.ik_end;  // = read end of the last candidate pushed, now entry 0
  m.n_prev = m.n_curr; m.n_curr = 0;  // the sweep result becomes `prev`
  m.i = m.x - 1; m.j = 0; m.n_tmp = 0;
  m.st = BSQ_ST_BWD;
}

// a candidate cannot be extended further to the left (bwt.c:350-356)
BSQ_HD void bsq_sm_bwd_stop(bsq_seed_machine_t &m, uint64_t x0, uint64_t x1, uint64_t x2, int end, bsq_pk_t *out) {
  if (m.n_curr != 0) return;  // contained in a longer match kept in this round
  if (m.n_tmp == 0 || m.i + 1 < m.last_beg) {
    if (m.n_out + m.n_tmp >= m.cap) { m.overflow = 1; return; }
    out[m.n_out + m.n_tmp] = bsq_pk_make(x0, x1, x2, m.i + 1, end);
    m.last_beg = m.i + 1;
    ++m.n_tmp;
  }
}

// end of one bwt_smem1a call: order by start, keep seeds of at least min_seed_len (memchain.c:69-71)
BSQ_HD void bsq_sm_finish_bwd(bsq_seed_machine_t &m, bsq_pk_t *out) {
  bsq_pk_t *t = out + m.n_out;
  for (int a = 0, b = m.n_tmp - 1; a < b; ++a, --b) { bsq_pk_t s = t[a]; t[a] = t[b]; t[b] = s; }
  int kept = 0;
  for (int a = 0; a < m.n_tmp; ++a)
    if (bsq_pk_end(t[a]) - bsq_pk_beg(t[a]) >= m.min_seed_len) t[kept++] = t[a];
  m.n_out += kept;
  m.n_tmp = 0;
  if (m.pass == 1) m.x = m.ret;
  m.st = BSQ_ST_NEXT;
}

// Run the control logic up to the next extension.  Returns false when the task is finished
// (out[0..n_out) then holds the unsorted interval list).
template <class S>
BSQ_HD bool bsq_sm_next(bsq_seed_machine_t &m, const bsq_fm_t &fm, const bsq_fm_t &fmc, S &scr, bsq_pk_t *out, bsq_ext_req_t &req) {
  for (;;) {
    if (m.overflow) { m.st = BSQ_ST_DONE; return false; }
    switch (m.st) {
      case BSQ_ST_NEXT: {
        if (m.pass == 1) {  // every SMEM (memchain.c:65-73)
          int c = 4;
          while (m.x < m.len && (c = scr.q(m.x)) > 3) ++m.x;
          if (m.x >= m.len) { m.pass = 2; m.old_n = m.n_out; m.k2 = 0; break; }
          bsq_sm_start_smem(m, fm, fmc, c, m.x, m.start_width);
        } else if (m.pass == 2) {  // re-seed from the middle of long, rare SMEMs (memchain.c:76-85)
          int found = 0;
          while (m.k2 < m.old_n) {
            const bsq_pk_t p = out[m.k2++];
            const int start = bsq_pk_beg(p), end = bsq_pk_end(p);
            if (end - start < m.split_len || bsq_pk_x2(p) > (uint64_t)m.split_width) continue;
            const int mid = (start + end) >> 1;
            bsq_sm_start_smem(m, fm, fmc, scr.q(mid), mid, (int)(bsq_pk_x2(p) + 1));
            found = 1;
            break;
          }
          if (!found) { m.pass = 3; m.x = 0; }
        } else {  // greedy forward seeds (memchain.c:88-103)
          if (m.max_mem_intv <= 0) { m.st = BSQ_ST_DONE; return false; }
          int c = 4;
          while (m.x < m.len && (c = scr.q(m.x)) > 3) ++m.x;
          if (m.x >= m.len) { m.st = BSQ_ST_DONE; return false; }
          m.ik0 = fm.L2[c] + 1; m.ik2 = fm.L2[c + 1] - fm.L2[c]; m.ik1 = fmc.L2[3 - c] + 1;
          m.i = m.x + 1;
          m.st = BSQ_ST_S1;
        }
        break;
      }
      case BSQ_ST_FWD: {
        const int c = m.i == m.len ? 4 : scr.q(m.i);
        if (c > 3) {  // read end or ambiguous base closes the sweep (bwt.c:335-340)
          scr.set(m.n_curr++, bsq_pk_make(m.ik0, m.ik1, m.ik2, 0, m.ik_end));
          bsq_sm_finish_fwd(m, scr);
          break;
        }
        req.x0 = m.ik0; req.x1 = m.ik1; req.x2 = m.ik2; req.back = 0; req.c = 3 - c;
        return true;
      }
      case BSQ_ST_BWD: {
        if (m.j == m.n_prev) {  // one column to the left done (bwt.c:362-363)
          if (m.n_curr == 0) { bsq_sm_finish_bwd(m, out); break; }
          m.n_prev = m.n_curr; m.n_curr = 0; --m.i; m.j = 0;
          if (m.i < -1) bsq_sm_finish_bwd(m, out);
          break;
        }
        const int c = m.i < 0 ? 4 : scr.q(m.i);
        const bsq_pk_t p = scr.get(m.j);
        req.x0 = bsq_pk_x0(p); req.x1 = bsq_pk_x1(p); req.x2 = bsq_pk_x2(p); req.back = 1; req.c = c;
        m.p_end = bsq_pk_end(p);
        if (c > 3) { bsq_sm_bwd_stop(m, req.x0, req.x1, req.x2, m.p_end, out); ++m.j; break; }
        return true;
      }
      case BSQ_ST_S1: {
        if (m.i == m.len) { m.x = m.len; m.st = BSQ_ST_NEXT; break; }
        const int c = scr.q(m.i);
        if (c > 3) { m.x = m.i + 1; m.st = BSQ_ST_NEXT; break; }
        req.x0 = m.ik0; req.x1 = m.ik1; req.x2 = m.ik2; req.back = 0; req.c = 3 - c;
        return true;
      }
      default:
        return false;
    }
  }
}

// Hand the result of the requested extension (still described by req) back to the machine.
template <class S>
BSQ_HD void bsq_sm_consume(bsq_seed_machine_t &m, S &scr, bsq_pk_t *out, const bsq_ext_req_t &req, uint64_t o0, uint64_t o1, uint64_t o2) {
  if (m.st == BSQ_ST_FWD) {  // bwt.c:326-334
    if (o2 != m.ik2) {
      scr.set(m.n_curr++, bsq_pk_make(m.ik0, m.ik1, m.ik2, 0, m.ik_end));
      if (o2 < (uint64_t)m.min_intv) { bsq_sm_finish_fwd(m, scr); return; }
    }
    m.ik0 = o0; m.ik1 = o1; m.ik2 = o2; m.ik_end = m.i + 1; ++m.i;
  } else if (m.st == BSQ_ST_BWD) {  // bwt.c:349-360
    if (o2 < (uint64_t)m.min_intv) bsq_sm_bwd_stop(m, req.x0, req.x1, req.x2, m.p_end, out);
    else if (m.n_curr == 0 || o2 != m.last_x2) {
      scr.set(m.n_curr++, bsq_pk_make(o0, o1, o2, 0, m.p_end));  // entry n_curr <= j: in-place compaction
      m.last_x2 = o2;
    }
    ++m.j;
  } else {  // BSQ_ST_S1, bwt.c:387-392
    if (o2 < (uint64_t)m.max_mem_intv && m.i - m.x >= m.min_seed_len) {
      if (o2 > 0) {  // memchain.c:95
        if (m.n_out >= m.cap) { m.overflow = 1; return; }
        out[m.n_out++] = bsq_pk_make(o0, o1, o2, m.x, m.i + 1);
      }
      m.x = m.i + 1;
      m.st = BSQ_ST_NEXT;
    } else {
      m.ik0 = o0; m.ik1 = o1; m.ik2 = o2; ++m.i;
    }
  }
}

struct bsq_key_less {
  BSQ_HD bool operator()(uint32_t a, uint32_t b) const { return (a >> 9) < (b >> 9); }
};

// Sort a finished list the way ks_introsort(mem_intv) does (memchain.c:105; not stable, so the exact
// comparison sequence matters: it only depends on the keys) and count the SA lookups chaining will need:
// min(x[2], max_occ) per interval (memchain.c:325-326).  The sort runs on 27-bit keys
// (beg:9 | end:9 | index:9) in `keys`; the 16-byte records are then permuted in place, cycle by cycle.
BSQ_HD int32_t bsq_seed_sort(const bsq_devopt_t &opt, bsq_pk_t *out, int n, uint32_t *keys) {
  int64_t tot = 0;
  for (int i = 0; i < n; ++i) {
    const bsq_pk_t p = out[i];
    const uint64_t x2 = bsq_pk_x2(p);
    tot += (int64_t)(x2 < (uint64_t)(uint32_t)opt.max_occ ? x2 : (uint64_t)(uint32_t)opt.max_occ);
    keys[i] = (uint32_t)bsq_pk_beg(p) << 18 | (uint32_t)bsq_pk_end(p) << 9 | (uint32_t)i;
  }
  bsq_introsort(keys, (int64_t)n, bsq_key_less());
  for (int i = 0; i < n; ++i) {
    if ((int)(keys[i] & 511) == i) continue;
    const bsq_pk_t first = out[i];
    int j = i;
    for (;;) {
      const int src = (int)(keys[j] & 511);
      keys[j] = (keys[j] & ~511u) | (uint32_t)j;
      if (src == i) { out[j] = first; break; }
      out[j] = out[src];
      j = src;
    }
  }
  return (int32_t)tot;
}
