// Warp-per-task chaining + chain filter with all per-task state in shared memory (k_chain v2).
//
// Same results as bsq_chain_task (= mem_chain + mem_chain_flt, lib/aln/memchain.c:268-488); different
// organisation.  v1 ran one task per thread with its B-tree, seed lists and chain records in global memory:
// every step was a dependent DRAM access, serialised across the diverged lanes of a warp (40 % of the
// GRCh38-sized step, profiles/README.md).  Here one warp owns a task and keeps everything in shared memory:
//
//  1. all lanes load the occurrences of the task (coalesced) and resolve contig ids;
//  2. the seeds are sorted by reference position (bitonic, in shared memory) and cut into clusters wherever two
//     neighbours are at least G = BSQ_MAX_READ_LEN + w + 1 apart.  A seed can only ever join the chain that
//     precedes it in reference order (merge_seed_to_chain needs |qdist - rdist| <= w and containment needs
//     overlap), so clusters cannot interact and their chains simply concatenate in position order -- the
//     order the reference reads out of its B-tree;
//  3. inside a cluster the reference's sequential rule is replayed in arrival order against a small sorted
//     list of the cluster's chains (lane 0; the list is a handful of entries);
//  4. weights, the introsort by weight (partitions: exact comparison sequence on lane 0; final insertion sort: a
//     parallel stable sort) and the greedy overlap filter, turned inside out: instead of walking every candidate over
//     the kept chains (memchain.c:427-457), every newly kept chain sweeps all later candidates that are still alive,
//     one candidate per lane.  A candidate meets the kept chains in the same order as in the reference and stops at the
//     same one, so `first`, `kept` and the kept list come out identical, in (kept chains) x (candidates / 32) steps
//     instead of (candidates) steps with mostly idle lanes.
//
// Exactness guard: the decomposition is only equivalent when (F1) no interval has more than max_occ occurrences
// (otherwise the reference's `count` cap couples clusters, memchain.c:325-326), (F2) the task fits the
// shared-memory capacity and (F3) no two chains get the same position (the B-tree's equal-key behaviour is
// shape dependent).  Any violation returns BSQ_CW_FALLBACK and the task is redone by bsq_chain_task.
#pragma once
#include "bsq_chain.h"

#define BSQ_CW_CAP 256   // default capacity (host emulation); the CUDA kernels use 64 / 128 / 256 / 512
#define BSQ_CW_OK 0
#define BSQ_CW_FALLBACK 1
#define BSQ_CW_NONE 0xFFFFu

// CAP = capacity in seeds (= SA lookups of the task); the kernel is instantiated for several capacities so that
// small tasks (the majority) run at high occupancy and only the few large ones pay for a large slice
template <int CAP_>
struct bsq_cw_smem_tt {
  static const int CAP = CAP_;
  int64_t rbeg[CAP_];       // seed reference position, arrival order
  static const int KEYS = CAP_ <= 64 ? 64 : CAP_ <= 128 ? 128 : CAP_ <= 256 ? 256 : CAP_ <= 512 ? 512 : 1024;  // bitonic sort pads to a power of two
  uint64_t key[KEYS];       // sort keys: rbeg << 10 | arrival index
  int64_t c_last_rbeg[CAP_];  // chain state, indexed by the arrival index of the chain's first seed
  int32_t c_w[CAP_];
  uint16_t qbeg[CAP_], slen[CAP_];
  int16_t rid[CAP_];        // < 0: seed dropped (bridges contigs / strands, memchain.c:339-346)
  uint16_t next[CAP_];      // seed lists
  uint16_t c_last_q[CAP_], c_last_len[CAP_], c_tail[CAP_], c_n[CAP_];
  uint16_t c_xhead[CAP_], c_xtail[CAP_], c_xn[CAP_];
  int16_t c_first[CAP_];
  uint16_t clist[CAP_];     // chains in position order
  uint16_t ord[CAP_];       // chains in filter order
  uint16_t keep[CAP_];
  uint8_t c_kept[CAP_];
  uint8_t c_alt[CAP_];       // is_alt of the seed's contig
  uint16_t iv_off[CAP_ + 2];  // first seed of every interval (prefix sums of the occurrence counts; n_intv <= n_sa)
  int32_t pub[4];                 // lane 0 -> all lanes: number of chains / fallback request
};
typedef bsq_cw_smem_tt<BSQ_CW_CAP> bsq_cw_smem_t;

// scalar policy (host emulation): one "lane"
struct bsq_cw_scalar {
  BSQ_HD static int lane() { return 0; }
  BSQ_HD static int nl() { return 1; }
  BSQ_HD static void sync() {}
  BSQ_HD static int first_true(bool p) { return p ? 0 : -1; }
  BSQ_HD static bool any(bool p) { return p; }
  BSQ_HD static int sum(int v) { return v; }
  BSQ_HD static int min(int v) { return v; }
  BSQ_HD static int scan_excl(int v, int &total) { total = v; return 0; }  // exclusive prefix sum over the lanes
  BSQ_HD static void sort_keys(uint64_t *k, int n) {
    for (int i = 1; i < n; ++i) { uint64_t v = k[i]; int j = i; while (j > 0 && k[j - 1] > v) { k[j] = k[j - 1]; --j; } k[j] = v; }
  }
  // the partition phase of ks_introsort over (weight << 16 | chain) keys, compared by weight only (defined below)
  BSQ_HD static void weight_partitions(uint32_t *k32, int n, uint16_t *scratch);
};

// filter order: sort (weight << 16 | chain) by weight only, descending -- the chain id rides along but takes no part
// in the comparison, so the comparison/swap sequence is the reference's (flt_lt, memchain.c:402)
struct bsq_cw_by_weight {  // 32-bit keys: weight (at most the read length) << 16 | chain
  BSQ_HD bool operator()(uint32_t a, uint32_t b) const { return (a >> 16) > (b >> 16); }
};

BSQ_HD void bsq_cw_scalar::weight_partitions(uint32_t *k32, int n, uint16_t *) { bsq_introsort<false>(k32, (int64_t)n, bsq_cw_by_weight()); }

// merge_seed_to_chain (memchain.c:227-256) against chain L (created by seed L)
template <typename S>
BSQ_HD int bsq_cw_merge(const bsq_devopt_t &opt, int64_t l_pac, S &s, int L, int a) {
  if (s.rid[a] != s.rid[L]) return 0;
  const int64_t f_rbeg = s.rbeg[L], l_rbeg = s.c_last_rbeg[L], rb = s.rbeg[a];
  const int f_q = s.qbeg[L], l_q = s.c_last_q[L], l_len = s.c_last_len[L], qb = s.qbeg[a], ln = s.slen[a];
  if (qb >= f_q && qb + ln <= l_q + l_len && rb >= f_rbeg && rb + ln <= l_rbeg + l_len) {
    if (s.c_xn[L] == 0) s.c_xhead[L] = (uint16_t)a; else s.next[s.c_xtail[L]] = (uint16_t)a;
    s.c_xtail[L] = (uint16_t)a; ++s.c_xn[L];
    return 1;
  }
  if ((l_rbeg < l_pac || f_rbeg < l_pac) && rb >= l_pac) return 0;
  const int64_t qdist = qb - l_q, rdist = rb - l_rbeg;
  if (rdist >= 0 && qdist - rdist <= opt.w && rdist - qdist <= opt.w && qdist - l_len < opt.max_chain_gap && rdist - l_len < opt.max_chain_gap) {
    s.next[s.c_tail[L]] = (uint16_t)a;
    s.c_tail[L] = (uint16_t)a; ++s.c_n[L];
    s.c_last_rbeg[L] = rb; s.c_last_q[L] = (uint16_t)qb; s.c_last_len[L] = (uint16_t)ln;
    return 1;
  }
  return 0;
}

template <typename S>
BSQ_HD int bsq_cw_weight(const S &s, int c) {  // mem_chain_weight (memchain.c:158-180)
  int64_t end = 0;
  int w = 0, tmp, j;
  for (j = c; j != (int)BSQ_CW_NONE; j = s.next[j]) {
    const int q = s.qbeg[j], l = s.slen[j];
    if (q >= end) w += l; else if (q + l > end) w += (int)(q + l - end);
    end = end > q + l ? end : q + l;
  }
  tmp = w; w = 0; end = 0;
  for (j = c; j != (int)BSQ_CW_NONE; j = s.next[j]) {
    const int64_t r = s.rbeg[j]; const int l = s.slen[j];
    if (r >= end) w += l; else if (r + l > end) w += (int)(r + l - end);
    end = end > r + l ? end : r + l;
  }
  w = w < tmp ? w : tmp;
  return w < 1 << 30 ? w : (1 << 30) - 1;
}

// Returns BSQ_CW_OK (res filled, outputs written) or BSQ_CW_FALLBACK (nothing written that matters).
template <typename W, typename S>
BSQ_HD int bsq_chain_warp(const bsq_devopt_t &opt, const bsq_devidx_t &ix, int parent, int l_seq, const bsq_pk_t *intv, int n_intv,
                          const uint64_t *sa_pos, int n_sa, S &s, bsq_chain_t *out_chains, bsq_seed_t *out_seeds,
                          bsq_chain_result_t &res) {
  const int lane = W::lane(), NL = W::nl();
  res.n_chains = 0; res.n_seeds = 0; res.status = 0; res.frac_rep = 0.f;
  if (l_seq < opt.min_seed_len) return BSQ_CW_OK;
  if (n_sa > S::CAP || n_intv > S::CAP || ix.n_seqs > 32767) return BSQ_CW_FALLBACK;  // (F2)
  const uint64_t max_occ = (uint64_t)(uint32_t)opt.max_occ;
  // ---- 1. seeds of the task (arrival order = interval order, then occurrence order) ----
  // 1a. occurrence counts of all intervals, lanes in parallel; prefix sums by lane 0
  {
    bool big = false;
    for (int i = lane; i < n_intv; i += NL) {
      const uint64_t x2 = bsq_pk_x2(intv[i]);
      if (x2 > max_occ) big = true;
      s.iv_off[i + 1] = (uint16_t)(x2 > max_occ ? 0 : x2);
    }
    if (W::any(big)) return BSQ_CW_FALLBACK;  // (F1)
    W::sync();
    if (lane == 0) {
      int o = 0;
      s.iv_off[0] = 0;
      for (int i = 0; i < n_intv; ++i) { o += s.iv_off[i + 1]; s.iv_off[i + 1] = (uint16_t)o; }
    }
    W::sync();
  }
  // 1b. one seed per lane step: interval by binary search in the prefix sums, position from the SA-lookup kernel
  for (int a = lane; a < n_sa; a += NL) {
    int lo = 0, hi = n_intv;  // last interval with iv_off[i] <= a
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (s.iv_off[mid] <= a) lo = mid; else hi = mid; }
    const bsq_pk_t pk = intv[lo];
    const int qb = bsq_pk_beg(pk), ln = bsq_pk_end(pk) - qb;
    const int64_t rb = (int64_t)sa_pos[a];
    int rid = bsq_intv2rid(ix, rb, rb + ln);
    if (rid >= 0 && (opt.bsstrand & 1) && bsq_getbss(ix, parent, rb) != (opt.bsstrand >> 1)) rid = -1;
    s.rbeg[a] = rb; s.qbeg[a] = (uint16_t)qb; s.slen[a] = (uint16_t)ln; s.rid[a] = (int16_t)(rid < 0 ? -1 : rid);
    s.c_alt[a] = (uint8_t)(rid >= 0 && ix.ann_is_alt[rid] != 0);
    s.next[a] = (uint16_t)BSQ_CW_NONE;
    s.key[a] = rid < 0 ? ~0ull : ((uint64_t)rb << 10 | (uint64_t)a);
  }
  W::sync();
  // ---- 2. order by reference position ----
  W::sort_keys(s.key, n_sa);
  W::sync();
  // ---- 3. clusters, replay of the merge rule: one cluster per lane at a time ----
  // Cluster [u, j) of the sorted keys keeps its chains (in position order) in clist[u, u + nc) and its arrival-order
  // scratch in ord[u, j): private ranges, and all chain state is indexed by seed, so clusters never touch the same
  // entry.  keep[u] = nc at cluster starts, 0 elsewhere; lane 0 then compacts clist left to right.
  const int64_t G = (int64_t)BSQ_MAX_READ_LEN + opt.w + 1;
  int n_ch = 0;
  int n_valid;
  {
    int nv = 0;
    for (int u = lane; u < n_sa; u += NL) { nv += s.key[u] != ~0ull; s.keep[u] = 0; }
    n_valid = W::sum(nv);  // dropped seeds sort to the end
  }
  W::sync();
  bool dup = false;
  for (int u = lane; u < n_valid; u += NL) {
    if (u > 0 && (int64_t)(s.key[u] >> 10) - (int64_t)(s.key[u - 1] >> 10) < G) continue;  // not a cluster start
    int j = u + 1;
    while (j < n_valid && (int64_t)(s.key[j] >> 10) - (int64_t)(s.key[j - 1] >> 10) < G) ++j;
    const int cl0 = u;
    int nc = cl0;  // end of this cluster's chain list
    // members in arrival order: insertion sort of the arrival indices of key[u..j)
    for (int v = u; v < j; ++v) s.ord[v] = (uint16_t)(s.key[v] & 1023);
    for (int v = u + 1; v < j; ++v) { uint16_t x = s.ord[v]; int t = v; while (t > u && s.ord[t - 1] > x) { s.ord[t] = s.ord[t - 1]; --t; } s.ord[t] = x; }
    for (int v = u; v < j && !dup; ++v) {
      const int a = s.ord[v];
      int p = cl0 - 1;  // predecessor inside the cluster (cl0 - 1: none)
      for (int c = cl0; c < nc; ++c) { if (s.rbeg[s.clist[c]] <= s.rbeg[a]) p = c; else break; }
      if (p >= cl0 && bsq_cw_merge(opt, ix.l_pac, s, s.clist[p], a)) continue;
      if (p >= cl0 && s.rbeg[s.clist[p]] == s.rbeg[a]) { dup = true; break; }  // (F3)
      for (int c = nc; c > p + 1; --c) s.clist[c] = s.clist[c - 1];
      s.clist[p + 1] = (uint16_t)a;
      ++nc;
      s.c_last_rbeg[a] = s.rbeg[a]; s.c_last_q[a] = s.qbeg[a]; s.c_last_len[a] = s.slen[a];
      s.c_tail[a] = (uint16_t)a; s.c_n[a] = 1; s.c_xhead[a] = s.c_xtail[a] = (uint16_t)BSQ_CW_NONE; s.c_xn[a] = 0;
    }
    s.keep[u] = (uint16_t)(nc - cl0);
  }
  if (W::any(dup)) return BSQ_CW_FALLBACK;
  W::sync();
  {  // compaction of the per-cluster chain lists into position order: prefix sums of keep[], copy through ord[]
    int base = 0;
    for (int u0 = 0; u0 < n_valid; u0 += NL) {
      const int u = u0 + lane;
      const int nc = u < n_valid ? s.keep[u] : 0;
      int tot;
      const int off = base + W::scan_excl(nc, tot);
      for (int c = 0; c < nc; ++c) s.ord[off + c] = s.clist[u + c];
      base += tot;
    }
    n_ch = base;
    W::sync();
    for (int c = lane; c < n_ch; c += NL) s.clist[c] = s.ord[c];
    W::sync();
  }
  // ---- 4a. weights (one chain per lane), then the order for the filter ----
  for (int c = lane; c < n_ch; c += NL) {
    const int a = s.clist[c];
    s.c_first[a] = -1; s.c_kept[a] = 0;
    s.c_w[a] = bsq_cw_weight(s, a);
  }
  W::sync();
  {  // a weight that does not fit the 16-bit key field cannot occur for reads of at most BSQ_MAX_READ_LEN bases
    bool big = false;
    for (int c = lane; c < n_ch; c += NL) big |= s.c_w[s.clist[c]] >= 65536;
    if (W::any(big)) return BSQ_CW_FALLBACK;
  }
  // ks_introsort = quicksort partitions that leave ranges of <= 16 elements, then ONE insertion sort over the whole
  // array (ksort.h:150-157,184-233).  An insertion sort is stable, so the second part is "the stable sort by weight of
  // whatever the partition phase left": lane 0 replays only the partitions (exact comparison/swap sequence), all
  // lanes then sort (weight descending, slot after partitioning) keys, which are unique.
  uint32_t *k32 = reinterpret_cast<uint32_t *>(s.c_last_rbeg);  // free since step 3; s.key takes the 64-bit keys
  {
    int base = 0;
    for (int c0 = 0; c0 < n_ch; c0 += NL) {
      const int c = c0 + lane;
      const int a = c < n_ch ? s.clist[c] : 0;
      const int ok = c < n_ch && s.c_w[a] >= opt.min_chain_weight;
      int tot;
      const int off = base + W::scan_excl(ok, tot);
      if (ok) k32[off] = (uint32_t)s.c_w[a] << 16 | (uint32_t)a;
      base += tot;
    }
    n_ch = base;
    W::sync();
    // (the whole range is partitioned once whatever its size, ksort.h:196-221: only sub-ranges of <= 16 are left alone)
    W::weight_partitions(k32, n_ch, reinterpret_cast<uint16_t *>(s.key));  // s.key is free between steps 3 and 4b
    W::sync();
  }
  for (int c = lane; c < n_ch; c += NL) {
    const uint32_t v = k32[c];
    s.key[c] = (uint64_t)(0xffffu - (v >> 16)) << 32 | (uint64_t)c << 16 | (uint64_t)(v & 0xffffu);
  }
  W::sync();
  W::sort_keys(s.key, n_ch);
  W::sync();
  for (int c = lane; c < n_ch; c += NL) s.ord[c] = (uint16_t)(s.key[c] & 0xffffu);
  if (lane == 0 && n_ch > 0) { s.c_kept[(uint16_t)(s.key[0] & 0xffffu)] = 3; s.keep[0] = 0; }
  W::sync();
  if (n_ch == 0) return BSQ_CW_OK;
  // ---- 4b. greedy overlap filter (memchain.c:427-457), one sweep per kept chain over the later candidates ----
  // key[i] (free after the sort) = query begin | query end << 16 | weight << 32 | is_alt << 48 of candidate i;
  // clist[i] (free after 4a) = bit 0: dropped, bit 1: has a significant overlap with a kept chain
  for (int i = lane; i < n_ch; i += NL) {
    const int c = s.ord[i];
    const uint64_t beg = s.qbeg[c], end = (uint64_t)s.c_last_q[c] + s.c_last_len[c];
    s.key[i] = beg | end << 16 | (uint64_t)(uint32_t)s.c_w[c] << 32 | (uint64_t)(s.c_alt[c] != 0) << 48;
    s.clist[i] = 0;
  }
  W::sync();
  int n_keep = 1;
  for (int cur = 0;;) {
    const uint64_t kk = s.key[cur];
    const int ck = s.ord[cur];
    const int ck_beg = (int)(kk & 0xffff), ck_end = (int)(kk >> 16 & 0xffff), ck_w = (int)(kk >> 32 & 0xffff);
    const bool ck_alt = (kk >> 48 & 1) != 0;
    const float wk = (float)ck_w * opt.drop_ratio;
    int first_i = 0x7fffffff, next = 0x7fffffff;
    for (int i = cur + 1 + lane; i < n_ch; i += NL) {
      int fl = s.clist[i];
      if (fl & 1) continue;
      const uint64_t ki = s.key[i];
      const int ci_beg = (int)(ki & 0xffff), ci_end = (int)(ki >> 16 & 0xffff), ci_w = (int)(ki >> 32 & 0xffff);
      const bool ci_alt = (ki >> 48 & 1) != 0;
      const int b_max = ck_beg > ci_beg ? ck_beg : ci_beg, e_min = ck_end < ci_end ? ck_end : ci_end;
      if (e_min > b_max && (!ck_alt || ci_alt)) {
        const int li = ci_end - ci_beg, lj = ck_end - ck_beg, min_l = li < lj ? li : lj;
        const float thr = (float)min_l * opt.mask_level;
        if ((float)(e_min - b_max) >= thr && min_l < opt.max_chain_gap) {
          fl |= 2;
          first_i = first_i < i ? first_i : i;
          if ((float)ci_w < wk && ck_w - ci_w >= (opt.min_seed_len << 1)) fl |= 1;
          s.clist[i] = (uint16_t)fl;
        }
      }
      if (!(fl & 1)) next = next < i ? next : i;  // i ascends per lane: the first survivor of this lane
    }
    first_i = W::min(first_i);
    next = W::min(next);
    W::sync();  // the flags written by the other lanes are visible to lane 0
    if (lane == 0 && first_i != 0x7fffffff && s.c_first[ck] < 0) s.c_first[ck] = (int16_t)first_i;
    if (next == 0x7fffffff) break;
    // the smallest survivor has now met every chain kept before it: it is kept itself
    if (lane == 0) { s.keep[n_keep] = (uint16_t)next; s.c_kept[s.ord[next]] = (s.clist[next] & 2) ? 2 : 3; }
    ++n_keep;
    cur = next;
    W::sync();
  }
  W::sync();
  // ---- 4c. kept = 1 for shadowed firsts, max_chain_extend (memchain.c:459-473), output slots ----
  for (int i = lane; i < n_keep; i += NL) {
    const int c = s.ord[s.keep[i]];
    if (s.c_first[c] >= 0) s.c_kept[s.ord[s.c_first[c]]] = 1;  // several kept chains may shadow the same one: same value
  }
  W::sync();
  {
    int cnt = 0;
    for (int i = lane; i < n_ch; i += NL) { const int kept = s.c_kept[s.ord[i]]; cnt += kept == 1 || kept == 2; }
    if ((uint32_t)W::sum(cnt) >= (uint32_t)opt.max_chain_extend) {  // only with -X given: the reference's sequential rule
      W::sync();
      if (lane == 0) {
        int i; uint32_t kk = 0;
        for (i = 0; i < n_ch; ++i) {
          const int kept = s.c_kept[s.ord[i]];
          if (kept == 0 || kept == 3) continue;
          if (++kk >= (uint32_t)opt.max_chain_extend) break;
        }
        for (; i < n_ch; ++i) if (s.c_kept[s.ord[i]] < 3) s.c_kept[s.ord[i]] = 0;
      }
      W::sync();
    }
  }
  // output slots: keep[o] = chain, iv_off[o] = first seed slot of output chain o (prefix sums in filter order)
  {
    int n_out_ = 0, s_out_ = 0;
    for (int i0 = 0; i0 < n_ch; i0 += NL) {
      const int i = i0 + lane;
      const int c = i < n_ch ? s.ord[i] : 0;
      const int live = i < n_ch && s.c_kept[c] != 0;
      const int ns = live ? s.c_n[c] + s.c_xn[c] : 0;
      int t1, t2;
      const int o1 = n_out_ + W::scan_excl(live, t1), o2 = s_out_ + W::scan_excl(ns, t2);
      if (live) { s.keep[o1] = (uint16_t)c; s.iv_off[o1] = (uint16_t)o2; }
      n_out_ += t1; s_out_ += t2;
    }
    if (lane == 0) { s.pub[1] = n_out_; s.pub[2] = s_out_; }
  }
  W::sync();
  const int n_out = s.pub[1];
  for (int o_ = lane; o_ < n_out; o_ += NL) {  // emit, one chain per lane
    const int c = s.keep[o_];
    int s_out = s.iv_off[o_];
    bsq_chain_t &o = out_chains[o_];
    o.pos = s.rbeg[c]; o.rid = s.rid[c]; o.w = s.c_w[c]; o.first = s.c_first[c]; o.kept = s.c_kept[c];
    o.is_alt = s.c_alt[c];
    o.seed_off = s_out; o.n_seeds = s.c_n[c]; o.n_extra = s.c_xn[c];
    for (int q = 0; q < 6; ++q) o.pad_[q] = 0;
    for (int j = c; j != (int)BSQ_CW_NONE; j = s.next[j]) { bsq_seed_t &d = out_seeds[s_out++]; d.rbeg = s.rbeg[j]; d.qbeg = s.qbeg[j]; d.len = s.slen[j]; }
    for (int j = s.c_xn[c] ? s.c_xhead[c] : (int)BSQ_CW_NONE; j != (int)BSQ_CW_NONE; j = s.next[j]) {
      bsq_seed_t &d = out_seeds[s_out++]; d.rbeg = s.rbeg[j]; d.qbeg = s.qbeg[j]; d.len = s.slen[j];
    }
  }
  res.n_chains = n_out; res.n_seeds = s.pub[2];
  return BSQ_CW_OK;
}
