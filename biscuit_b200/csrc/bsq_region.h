// Chains -> alignment regions for one (read, conversion) task:
//   mem_chain2region   lib/aln/memchain.c:873-904
//   mem_chain2region1  lib/aln/memchain.c:742-871
//   left/right_extend_seed_set_align_*  memchain.c:613-730, mem_chain_reference_span :585-605,
//   cal_max_gap :576-582, asymmetric_flt_seed :138-149, bns_fetch_seq bntseq.c:428-452.
// The reference window of a chain is never materialised: target bases are decoded straight
// from the 2-bit packed forward reference (reverse strand = complement read backwards,
// bntseq.c:411-416), which stays L2-resident for the few hundred bases a chain touches.
#pragma once
#include "bsq_chain.h"
#include "bsq_ksw.h"

BSQ_HD int bsq_pac_base(const uint8_t *pac, int64_t l) { return pac[l >> 2] >> ((~l & 3) << 1) & 3; }

// base at forward-reverse coordinate pos in [0, 2*l_pac)
BSQ_HD int bsq_ref_base(const bsq_devidx_t &ix, int64_t pos) {
  BSQ_CTR(BSQ_CTR_REFB, 1);
  return pos < ix.l_pac ? bsq_pac_base(ix.pac, pos) : 3 - bsq_pac_base(ix.pac, (ix.l_pac << 1) - 1 - pos);
}

BSQ_HD int bsq_cal_max_gap(const bsq_devopt_t &opt, int qlen) {
  // (int)((double)x / e + 1.) of memchain.c:577-578 in integer arithmetic: x and e are small integers, e > 0, so
  // the double quotient is never within rounding distance of an integer it does not equal, and truncating
  // x/e + 1 toward zero is the C division (x + e) / e
  int l_del = opt.e_del > 0 ? (qlen * opt.a - opt.o_del + opt.e_del) / opt.e_del : (int)((double)(qlen * opt.a - opt.o_del) / opt.e_del + 1.);
  int l_ins = opt.e_ins > 0 ? (qlen * opt.a - opt.o_ins + opt.e_ins) / opt.e_ins : (int)((double)(qlen * opt.a - opt.o_ins) / opt.e_ins + 1.);
  int l = l_del > l_ins ? l_del : l_ins;
  l = l > 1 ? l : 1;
  return l < opt.w << 1 ? l : opt.w << 1;
}

// Query / target accessors shared by left (step = -1) and right (step = +1) extensions, so that the
// DP code is instantiated once.  The target comes either from the packed reference (ix != 0) or from
// a plain nt4 buffer (kernel-level test entry point).
struct bsq_qacc_t {
  const uint8_t *q;
  int step;
  BSQ_HD int operator()(int j) const { return q[j * step]; }
};
struct bsq_tacc_t {
  const bsq_devidx_t *ix;
  const uint8_t *buf;
  int64_t p0;
  int step;
  BSQ_HD int operator()(int i) const { return ix ? bsq_ref_base(*ix, p0 + (int64_t)i * step) : buf[i]; }
};

// Execution policy of the region builder.  The CUDA warp policy (bsq_ksw_warp.cuh) runs the control flow
// redundantly and uniformly in all 32 lanes, lets lane 0 do the stores, and spreads each DP row
// over the lanes.
// (A scalar policy that runs everything in the calling thread exists for the test-only host emulation:
// tests/hostemu/bsq_ksw_scalar.h.)

// What k_region_prep (bsq_align.cu) computes ahead of k_region, one chain / one seed per lane instead of redundantly in all
// 32 lanes of the task's warp: the seed order of mem_chain2region1 (keys of one chain in srt[seed_off ..], backup seeds
// behind the main ones), the verdict of asymmetric_flt_seed for every seed (aflag[], indexed like the seeds), and the
// chain's reference span.

// asymmetric_flt_seed (memchain.c:138-149) for one seed, one thread: ref T under read C, or ref A under read G
BSQ_HD bool bsq_asym_seed(const bsq_devidx_t &ix, const bsq_seed_t &s, const uint8_t *query) {
  const bool rev = s.rbeg >= ix.l_pac;  // a seed never spans the strand boundary (memchain.c:339)
  int64_t f = rev ? (ix.l_pac << 1) - 1 - s.rbeg : s.rbeg;  // forward coordinate of the first base; the reverse strand reads backwards
  const int step = rev ? -1 : 1, comp = rev ? 3 : 0;
  const uint8_t *q = query + s.qbeg;
  for (int b = 0; b < s.len; ++b, f += step) {
    const int r = (ix.pac[f >> 2] >> ((~f & 3) << 1) & 3) ^ comp, qv = q[b];
    if ((r == 3 && qv == 1) || (r == 0 && qv == 2)) return true;
  }
  return false;
}

// mem_chain2region1 for one seed list.  regs[reg0..*n_regs) are the regions of this task so far.
// PREP: srt[0..n_seeds) already holds the sorted keys and aflag[i] the asymmetric-filter verdict of seeds[i].
template <typename X, bool PREP = false>
BSQ_HD void bsq_chain2region1(const bsq_devopt_t &opt, const bsq_devidx_t &ix, int64_t rmax0, int64_t rmax1, int rid,
                              int l_query, const uint8_t *query, const bsq_seed_t *seeds, int n_seeds, int parent,
                              float frac_rep, uint64_t *srt, bsq_ksw_scratch_t *ksw, bsq_reg_t *regs, int *n_regs, const uint8_t *aflag = nullptr) {
  const int8_t *mat = parent ? opt.ctmat : opt.gamat;
  struct u64_less { BSQ_HD bool operator()(uint64_t a, uint64_t b) const { return a < b; } };
  if (!PREP) {
    if (X::leader()) {
      for (int i = 0; i < n_seeds; ++i) srt[i] = (uint64_t)(uint32_t)seeds[i].len << 32 | (uint32_t)i;  // score == len
      bsq_introsort(srt, (int64_t)n_seeds, u64_less());
    }
    X::sync();
  }
  for (int k = n_seeds - 1; k >= 0; --k) {
    const uint64_t key = srt[k];
    const bsq_seed_t &s = seeds[(uint32_t)key];
    // asymmetric_flt_seed: reject ref T/read C and ref A/read G inside the seed
    if (PREP ? aflag[(uint32_t)key] != 0 : X::asym_conflict(ix, s, query)) continue;
    // was this seed already covered by an earlier extension?
    int u;
    for (u = 0; u < *n_regs; ++u) {
      const bsq_reg_t &reg = regs[u];
      if (s.rbeg < reg.rb || s.rbeg + s.len > reg.re || s.qbeg < reg.qb || s.qbeg + s.len > reg.qe) continue;
      if ((double)(s.len - reg.seedlen0) > .1 * l_query) continue;
      int qd = s.qbeg - reg.qb;
      int64_t rd = s.rbeg - reg.rb;
      int max_gap = X::max_gap(opt, (int)(qd < rd ? qd : rd));
      int w = max_gap < reg.w ? max_gap : reg.w;
      if (qd - rd < w && rd - qd < w) break;
      qd = reg.qe - (s.qbeg + s.len);
      rd = reg.re - (s.rbeg + s.len);
      max_gap = X::max_gap(opt, (int)(qd < rd ? qd : rd));
      w = max_gap < reg.w ? max_gap : reg.w;
      if (qd - rd < w && rd - qd < w) break;
    }
    if (u < *n_regs) {
      // almost contained: extend anyway only if an overlapping seed sits on another diagonal
      int i;
      for (i = k + 1; i < n_seeds; ++i) {
        if (srt[i] == 0) continue;
        const bsq_seed_t &t = seeds[(uint32_t)srt[i]];
        if ((double)t.len < s.len * .95) continue;
        if (s.qbeg <= t.qbeg && s.qbeg + s.len - t.qbeg >= s.len >> 2 && t.qbeg - s.qbeg != t.rbeg - s.rbeg) break;
        if (t.qbeg <= s.qbeg && t.qbeg + t.len - s.qbeg >= s.len >> 2 && s.qbeg - t.qbeg != s.rbeg - t.rbeg) break;
      }
      if (i == n_seeds) {
        X::sync();  // every lane has finished reading srt[]
        if (X::leader()) srt[k] = 0;
        X::sync();
        continue;
      }
    }
    // ---- extension ----
    bsq_reg_t reg;
    int aw0 = opt.w, aw1 = opt.w;
    reg.rb = reg.re = 0; reg.qb = reg.qe = 0; reg.w = opt.w; reg.score = reg.truesc = -1; reg.rid = rid;
    reg.seedcov = 0; reg.seedlen0 = 0; reg.frac_rep = 0.f; reg.bss = reg.parent = 0; reg.pad_[0] = reg.pad_[1] = 0;
    if (s.qbeg == 0) {
      reg.score = reg.truesc = s.len * opt.a; reg.qb = 0; reg.rb = s.rbeg;
    } else {
      bsq_qacc_t qa; qa.q = query + s.qbeg - 1; qa.step = -1;
      bsq_tacc_t ta; ta.ix = &ix; ta.buf = nullptr; ta.p0 = s.rbeg - 1; ta.step = -1;
      const int tlen = (int)(s.rbeg - rmax0);
      bsq_ext_result_t r;
      r.score = 0; r.qle = r.tle = r.gtle = 0; r.gscore = -1; r.max_off = 0;
      for (int i = 0; i < 2; ++i) {
        const int prev = reg.score;
        aw0 = opt.w << i;
        r = X::extend(s.qbeg, qa, tlen, ta, mat, opt.o_del, opt.e_del, opt.o_ins, opt.e_ins, aw0, opt.pen_clip5,
                      opt.zdrop, s.len * opt.a, ksw);
        reg.score = r.score;
        if (reg.score == prev || r.max_off < (aw0 >> 1) + (aw0 >> 2)) break;
      }
      if (r.gscore <= 0 || r.gscore <= reg.score - opt.pen_clip5) {
        reg.qb = s.qbeg - r.qle; reg.rb = s.rbeg - r.tle; reg.truesc = reg.score;
      } else {
        reg.qb = 0; reg.rb = s.rbeg - r.gtle; reg.truesc = r.gscore;
      }
    }
    if (s.qbeg + s.len == l_query) {
      reg.qe = l_query; reg.re = s.rbeg + s.len;
    } else {
      const int sc0 = reg.score, qe = s.qbeg + s.len;
      bsq_qacc_t qa; qa.q = query + qe; qa.step = 1;
      bsq_tacc_t ta; ta.ix = &ix; ta.buf = nullptr; ta.p0 = s.rbeg + s.len; ta.step = 1;
      const int tlen = (int)(rmax1 - (s.rbeg + s.len));
      bsq_ext_result_t r;
      r.score = 0; r.qle = r.tle = r.gtle = 0; r.gscore = -1; r.max_off = 0;
      for (int i = 0; i < 2; ++i) {
        const int prev = reg.score;
        aw1 = opt.w << i;
        r = X::extend(l_query - qe, qa, tlen, ta, mat, opt.o_del, opt.e_del, opt.o_ins, opt.e_ins, aw1,
                      opt.pen_clip3, opt.zdrop, sc0, ksw);
        reg.score = r.score;
        if (reg.score == prev || r.max_off < (aw1 >> 1) + (aw1 >> 2)) break;
      }
      if (r.gscore <= 0 || r.gscore <= reg.score - opt.pen_clip3) {
        reg.qe = qe + r.qle; reg.re = s.rbeg + s.len + r.tle; reg.truesc += reg.score - sc0;
      } else {
        reg.qe = l_query; reg.re = s.rbeg + s.len + r.gtle; reg.truesc += r.gscore - sc0;
      }
    }
    reg.bss = (uint8_t)bsq_getbss(ix, parent, reg.rb);
    reg.parent = (uint8_t)parent;
    if (bsq_getbss(ix, parent, reg.re) != reg.bss) continue;  // crosses the strand boundary: dropped
    int cov = 0;
    for (int i = 0; i < n_seeds; ++i) {
      const bsq_seed_t &t = seeds[i];
      if (t.qbeg >= reg.qb && t.qbeg + t.len <= reg.qe && t.rbeg >= reg.rb && t.rbeg + t.len <= reg.re) cov += t.len;
    }
    reg.seedcov = cov;
    reg.w = aw0 > aw1 ? aw0 : aw1;
    reg.seedlen0 = s.len;
    reg.frac_rep = frac_rep;
    if (X::leader()) regs[*n_regs] = reg;
    ++(*n_regs);
    X::sync();
  }
}

// mem_chain_reference_span (memchain.c:585-605) + the clipping of bns_fetch_seq (bntseq.c:428-452) for one chain
template <typename X>
BSQ_HD void bsq_chain_span(const bsq_devopt_t &opt, const bsq_devidx_t &ix, int l_query, const bsq_chain_t &c, const bsq_seed_t *cs, int64_t &rmax0_out,
                           int64_t &rmax1_out) {
  const int64_t l_pac = ix.l_pac;
  {
    int64_t rmax0 = l_pac << 1, rmax1 = 0;
    for (int i = 0; i < c.n_seeds; ++i) {
      const bsq_seed_t &s = cs[i];
      int64_t b = s.rbeg - (s.qbeg + X::max_gap(opt, s.qbeg));
      int64_t e = s.rbeg + s.len + ((l_query - s.qbeg - s.len) + X::max_gap(opt, l_query - s.qbeg - s.len));
      rmax0 = rmax0 < b ? rmax0 : b;
      rmax1 = rmax1 > e ? rmax1 : e;
    }
    rmax0 = rmax0 > 0 ? rmax0 : 0;
    rmax1 = rmax1 < l_pac << 1 ? rmax1 : l_pac << 1;
    if (rmax0 < l_pac && l_pac < rmax1) {
      if (cs[0].rbeg < l_pac) rmax1 = l_pac; else rmax0 = l_pac;
    }
    // bns_fetch_seq: clip to the contig (and strand) that holds the first seed
    // contig of the first seed: the chain record carries it (set from the same seed by the chaining kernels)
    const int is_rev = cs[0].rbeg >= l_pac;
    const int rid = c.rid;
    int64_t far_beg = ix.ann_offset[rid], far_end = far_beg + ix.ann_len[rid];
    if (is_rev) {
      int64_t t = far_beg;
      far_beg = (l_pac << 1) - far_end;
      far_end = (l_pac << 1) - t;
    }
    rmax0 = rmax0 > far_beg ? rmax0 : far_beg;
    rmax1 = rmax1 < far_end ? rmax1 : far_end;
    rmax0_out = rmax0; rmax1_out = rmax1;
  }
}

// mem_chain2region for one task.  Returns the number of regions written to regs[].
// PREP: spans[2 ci], spans[2 ci + 1] and the srt slices of every chain were filled by k_region_prep.
template <typename X, bool PREP = false>
BSQ_HD int bsq_chain2region(const bsq_devopt_t &opt, const bsq_devidx_t &ix, int parent, int l_query, const uint8_t *query,
                            const bsq_chain_t *chains, int n_chains, const bsq_seed_t *seeds, float frac_rep,
                            uint64_t *srt, bsq_ksw_scratch_t *ksw, bsq_reg_t *regs, const int64_t *spans = nullptr, const uint8_t *aflag = nullptr) {
  int n_regs = 0;
  for (int ci = 0; ci < n_chains; ++ci) {
    const bsq_chain_t &c = chains[ci];
    if (c.n_seeds == 0) continue;
    const bsq_seed_t *cs = seeds + c.seed_off;
    int64_t rmax0, rmax1;
    if (PREP) { rmax0 = spans[2 * ci]; rmax1 = spans[2 * ci + 1]; }
    else bsq_chain_span<X>(opt, ix, l_query, c, cs, rmax0, rmax1);
    const int rid = c.rid;
    const int n0 = n_regs;
    bsq_chain2region1<X, PREP>(opt, ix, rmax0, rmax1, rid, l_query, query, cs, c.n_seeds, parent, frac_rep, PREP ? srt + c.seed_off : srt, ksw, regs,
                               &n_regs, PREP ? aflag + c.seed_off : nullptr);
    if (n_regs == n0 && c.n_extra > 0)
      bsq_chain2region1<X, PREP>(opt, ix, rmax0, rmax1, rid, l_query, query, cs + c.n_seeds, c.n_extra, parent, frac_rep,
                                 PREP ? srt + c.seed_off + c.n_seeds : srt, ksw, regs, &n_regs, PREP ? aflag + c.seed_off + c.n_seeds : nullptr);
  }
  return n_regs;
}
