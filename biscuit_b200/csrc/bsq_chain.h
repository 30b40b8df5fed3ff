// Seed chaining and chain filtering for one (read, conversion) task:
//   mem_chain      lib/aln/memchain.c:268-393  (merge_seed_to_chain :227-256)
//   mem_chain_flt  lib/aln/memchain.c:406-488  (mem_chain_weight :158-180)
//
// The reference keeps the growing chains in a klib B-tree (kbtree.h, t = 3 for the 72-byte
// mem_chain_t: 5 keys per node) and asks it for the predecessor of every new seed.  With equal
// keys the answer depends on the tree shape, so the same B-tree is built
// here -- over 32-bit chain ids instead of by-value structs, from a per-task pool, without
// recursion -- rather than a sorted array.
#pragma once
#include "bsq_common.h"
#include "bsq_fm.h"
#include "bsq_seed.h"
#include "bsq_sort.h"

#define BSQ_BT_T 3
#define BSQ_BT_MAXKEYS (2 * BSQ_BT_T - 1)

struct bsq_bnode_t {
  int32_t n, internal;
  int32_t key[BSQ_BT_MAXKEYS];
  int32_t ptr[BSQ_BT_MAXKEYS + 1];
};

// seed list node (a chain's `seeds` and `seeds_extra` vectors are singly linked lists)
struct bsq_snode_t {
  int64_t rbeg;
  int32_t qbeg, len;
  int32_t next, pad_;
};

struct bsq_wchain_t {
  int64_t pos;
  int64_t first_rbeg, last_rbeg;
  int32_t first_qbeg, first_len, last_qbeg, last_len;
  int32_t rid, is_alt;
  int32_t head, tail, n;     // seeds
  int32_t xhead, xtail, xn;  // seeds_extra
  int32_t w, first, kept, pad_;
};

// Per-task workspace carved out of device pools by the host (all sized by `cap` seeds).
struct bsq_chain_ws_t {
  int32_t cap;
  bsq_snode_t *snodes;   // cap
  bsq_wchain_t *chains;  // cap
  bsq_bnode_t *bnodes;   // cap + 2
  int32_t *order;        // cap
};

struct bsq_btree_t {
  bsq_bnode_t *nodes;
  const bsq_wchain_t *chains;
  int32_t root, n_nodes, n_keys;
};

BSQ_HD int bsq_bt_cmp(int64_t a, int64_t b) { return (int)(b < a) - (int)(a < b); }

BSQ_HD int bsq_bt_new_node(bsq_btree_t &t, int internal) {
  int id = t.n_nodes++;
  bsq_bnode_t &x = t.nodes[id];
  x.n = 0; x.internal = internal;
  for (int i = 0; i <= BSQ_BT_MAXKEYS; ++i) x.ptr[i] = -1;
  return id;
}

BSQ_HD void bsq_bt_init(bsq_btree_t &t, bsq_bnode_t *nodes, const bsq_wchain_t *chains) {
  t.nodes = nodes; t.chains = chains; t.n_nodes = 0; t.n_keys = 0;
  t.root = bsq_bt_new_node(t, 0);
}

// Position of `pos` inside one node (kbtree.h __kb_getp_aux): index of the first key equal to
// pos (r = 0), else of the last key below it (possibly -1, r != 0).
BSQ_HD int bsq_bt_locate(const bsq_btree_t &t, const bsq_bnode_t &x, int64_t pos, int *r) {
  int begin = 0, end = x.n;
  if (x.n == 0) return -1;
  while (begin < end) {
    int mid = (begin + end) >> 1;
    if (bsq_bt_cmp(t.chains[x.key[mid]].pos, pos) < 0) begin = mid + 1;
    else end = mid;
  }
  if (begin == x.n) { *r = 1; return x.n - 1; }
  if ((*r = bsq_bt_cmp(pos, t.chains[x.key[begin]].pos)) < 0) --begin;
  return begin;
}

// kb_intervalp restricted to the lower bound: id of the chain the descent reports as the
// closest one at or below pos, -1 if none.
BSQ_HD int bsq_bt_lower(const bsq_btree_t &t, int64_t pos) {
  int lower = -1, r = 0, x = t.root;
  while (x >= 0) {
    const bsq_bnode_t &nd = t.nodes[x];
    int i = bsq_bt_locate(t, nd, pos, &r);
    if (i >= 0 && r == 0) return nd.key[i];
    if (i >= 0) lower = nd.key[i];
    if (!nd.internal) return lower;
    x = nd.ptr[i + 1];
  }
  return lower;
}

// split the full child y = x.ptr[i] (kbtree.h __kb_split)
BSQ_HD void bsq_bt_split(bsq_btree_t &t, int xi, int i, int yi) {
  int zi = bsq_bt_new_node(t, t.nodes[yi].internal);
  bsq_bnode_t &x = t.nodes[xi], &y = t.nodes[yi], &z = t.nodes[zi];
  z.n = BSQ_BT_T - 1;
  for (int a = 0; a < BSQ_BT_T - 1; ++a) z.key[a] = y.key[BSQ_BT_T + a];
  if (y.internal)
    for (int a = 0; a < BSQ_BT_T; ++a) z.ptr[a] = y.ptr[BSQ_BT_T + a];
  y.n = BSQ_BT_T - 1;
  for (int a = x.n; a > i; --a) x.ptr[a + 1] = x.ptr[a];
  x.ptr[i + 1] = zi;
  for (int a = x.n - 1; a >= i; --a) x.key[a + 1] = x.key[a];
  x.key[i] = y.key[BSQ_BT_T - 1];
  ++x.n;
}

// kb_putp: insert chain id `cid` (key = chains[cid].pos)
BSQ_HD void bsq_bt_put(bsq_btree_t &t, int cid) {
  const int64_t pos = t.chains[cid].pos;
  int r = 0;
  ++t.n_keys;
  if (t.nodes[t.root].n == BSQ_BT_MAXKEYS) {
    int s = bsq_bt_new_node(t, 1);
    t.nodes[s].ptr[0] = t.root;
    bsq_bt_split(t, s, 0, t.root);
    t.root = s;
  }
  int xi = t.root;
  for (;;) {
    bsq_bnode_t &x = t.nodes[xi];
    if (!x.internal) {
      int i = bsq_bt_locate(t, x, pos, &r);
      for (int a = x.n - 1; a > i; --a) x.key[a + 1] = x.key[a];
      x.key[i + 1] = cid;
      ++x.n;
      return;
    }
    int i = bsq_bt_locate(t, x, pos, &r) + 1;
    if (t.nodes[x.ptr[i]].n == BSQ_BT_MAXKEYS) {
      bsq_bt_split(t, xi, i, x.ptr[i]);
      if (bsq_bt_cmp(pos, t.chains[t.nodes[xi].key[i]].pos) > 0) ++i;
    }
    xi = t.nodes[xi].ptr[i];
  }
}

// in-order walk; writes chain ids to out[], returns their number
BSQ_HD int bsq_bt_inorder(const bsq_btree_t &t, int32_t *out) {
  int st_node[40], st_i[40], top = 0, n = 0;
  if (t.n_keys == 0) return 0;
  st_node[0] = t.root; st_i[0] = 0; top = 1;
  while (top > 0) {
    const bsq_bnode_t &x = t.nodes[st_node[top - 1]];
    int i = st_i[top - 1];
    if (x.internal) {
      // state i even: descend into child i/2 ; odd: emit key (i-1)/2
      if ((i & 1) == 0) {
        int c = i >> 1;
        st_i[top - 1] = i + 1;
        if (c <= x.n) { st_node[top] = x.ptr[c]; st_i[top] = 0; ++top; }
      } else {
        int k = i >> 1;
        if (k < x.n) { out[n++] = x.key[k]; st_i[top - 1] = i + 1; }
        else --top;
      }
    } else {
      for (int k = 0; k < x.n; ++k) out[n++] = x.key[k];
      --top;
    }
  }
  return n;
}

// bns_pos2rid (lib/aln/bntseq.c:356-369)
BSQ_HD int bsq_pos2rid(const bsq_devidx_t &ix, int64_t pos_f) {
  if (pos_f >= ix.l_pac) return -1;
  int left = 0, mid = 0, right = ix.n_seqs;
  while (left < right) {
    mid = (left + right) >> 1;
    if (pos_f >= ix.ann_offset[mid]) {
      if (mid == ix.n_seqs - 1) break;
      if (pos_f < ix.ann_offset[mid + 1]) break;
      left = mid + 1;
    } else right = mid;
  }
  return mid;
}

BSQ_HD int64_t bsq_depos(const bsq_devidx_t &ix, int64_t pos, int *is_rev) {
  return (*is_rev = (pos >= ix.l_pac)) ? (ix.l_pac << 1) - 1 - pos : pos;
}

// bns_intv2rid (lib/aln/bntseq.c:371-378)
BSQ_HD int bsq_intv2rid(const bsq_devidx_t &ix, int64_t rb, int64_t re) {
  int is_rev;
  if (rb < ix.l_pac && re > ix.l_pac) return -2;
  int rid_b = bsq_pos2rid(ix, bsq_depos(ix, rb, &is_rev));
  int rid_e = rb < re ? bsq_pos2rid(ix, bsq_depos(ix, re - 1, &is_rev)) : rid_b;
  return rid_b == rid_e ? rid_b : -1;
}

// mem_getbss (lib/aln/memchain.c:265)
BSQ_HD int bsq_getbss(const bsq_devidx_t &ix, int parent, int64_t rb) { return ((rb > ix.l_pac) == (parent != 0)) ? 1 : 0; }

// merge_seed_to_chain (memchain.c:227-256).  snode `si` already holds the seed.
BSQ_HD int bsq_merge_seed(const bsq_devopt_t &opt, int64_t l_pac, bsq_wchain_t &c, bsq_snode_t *sn, int si, int seed_rid) {
  const bsq_snode_t &s = sn[si];
  if (seed_rid != c.rid) return 0;
  if (s.qbeg >= c.first_qbeg && s.qbeg + s.len <= c.last_qbeg + c.last_len && s.rbeg >= c.first_rbeg &&
      s.rbeg + s.len <= c.last_rbeg + c.last_len) {
    if (c.xn == 0) c.xhead = si; else sn[c.xtail].next = si;
    c.xtail = si; ++c.xn;
    return 1;  // contained: parked in the backup list
  }
  if ((c.last_rbeg < l_pac || c.first_rbeg < l_pac) && s.rbeg >= l_pac) return 0;  // other strand
  int64_t qdist = s.qbeg - c.last_qbeg, rdist = s.rbeg - c.last_rbeg;
  if (rdist >= 0 && qdist - rdist <= opt.w && rdist - qdist <= opt.w && qdist - c.last_len < opt.max_chain_gap &&
      rdist - c.last_len < opt.max_chain_gap) {
    sn[c.tail].next = si;
    c.tail = si; ++c.n;
    c.last_rbeg = s.rbeg; c.last_qbeg = s.qbeg; c.last_len = s.len;
    return 1;
  }
  return 0;
}

// mem_chain_weight (memchain.c:158-180)
BSQ_HD int bsq_chain_weight(const bsq_wchain_t &c, const bsq_snode_t *sn) {
  int64_t end = 0;
  int w = 0, tmp, j;
  for (j = c.head; j >= 0; j = sn[j].next) {
    const bsq_snode_t &s = sn[j];
    if (s.qbeg >= end) w += s.len;
    else if (s.qbeg + s.len > end) w += (int)(s.qbeg + s.len - end);
    end = end > s.qbeg + s.len ? end : s.qbeg + s.len;
  }
  tmp = w; w = 0; end = 0;
  for (j = c.head; j >= 0; j = sn[j].next) {
    const bsq_snode_t &s = sn[j];
    if (s.rbeg >= end) w += s.len;
    else if (s.rbeg + s.len > end) w += (int)(s.rbeg + s.len - end);
    end = end > s.rbeg + s.len ? end : s.rbeg + s.len;
  }
  w = w < tmp ? w : tmp;
  return w < 1 << 30 ? w : (1 << 30) - 1;
}

struct bsq_chain_w_desc {
  const bsq_wchain_t *ch;
  BSQ_HD bool operator()(int32_t a, int32_t b) const { return ch[a].w > ch[b].w; }
};

struct bsq_chain_result_t {
  int32_t n_chains;  // kept chains written to out_chains
  int32_t n_seeds;   // seed slots used in out_seeds
  int32_t status;    // 0 ok, 1 workspace overflow
  float frac_rep;
};

// mem_chain + mem_chain_flt for one task.  intv[0..n_intv) is the sorted interval list,
// sa_pos[] the text positions of its first min(x[2], max_occ) occurrences each, interval after
// interval (computed by the SA-lookup kernel).  Results go to out_chains / out_seeds (both with
// room for ws.cap entries).
BSQ_HD bsq_chain_result_t bsq_chain_task(const bsq_devopt_t &opt, const bsq_devidx_t &ix, int parent, int l_seq,
                                         const bsq_pk_t *intv, int n_intv, const uint64_t *sa_pos, bsq_chain_ws_t &ws,
                                         bsq_chain_t *out_chains, bsq_seed_t *out_seeds) {
  bsq_chain_result_t res;
  res.n_chains = 0; res.n_seeds = 0; res.status = 0; res.frac_rep = 0.f;
  if (l_seq < opt.min_seed_len) return res;
  const int64_t l_pac = ix.l_pac;
  const bsq_fm_t &fm = ix.fm[parent];
  // length of the read covered by repetitive seeds (memchain.c:294-301)
  int b = 0, e = 0, l_rep = 0;
  for (int i = 0; i < n_intv; ++i) {
    if (bsq_pk_x2(intv[i]) <= (uint64_t)(uint32_t)opt.max_occ) continue;
    int sb = bsq_pk_beg(intv[i]), se = bsq_pk_end(intv[i]);
    if (sb > e) { l_rep += e - b; b = sb; e = se; }
    else e = e > se ? e : se;
  }
  l_rep += e - b;
  res.frac_rep = (float)l_rep / (float)l_seq;

  bsq_btree_t tree;
  bsq_bt_init(tree, ws.bnodes, ws.chains);
  int n_sn = 0, n_ch = 0;
  int64_t sa_i = 0;
  for (int i = 0; i < n_intv; ++i) {
    const bsq_pk_t pk = intv[i];
    const uint64_t v_x0 = bsq_pk_x0(pk), v_x2 = bsq_pk_x2(pk);
    const int v_beg = bsq_pk_beg(pk);
    const int slen = bsq_pk_end(pk) - v_beg;
    const uint64_t n_pre = v_x2 < (uint64_t)(uint32_t)opt.max_occ ? v_x2 : (uint64_t)(uint32_t)opt.max_occ;
    uint32_t count = 0;
    uint64_t k;
    for (k = 0; k < v_x2 && count < (uint32_t)opt.max_occ && ((count > 5 && k < (uint64_t)(uint32_t)opt.max_occ) || count <= 5); ++k) {
      const int64_t rbeg = (int64_t)(k < n_pre ? sa_pos[sa_i + (int64_t)k] : bsq_sa(fm, v_x0 + k));
      const int rid = bsq_intv2rid(ix, rbeg, rbeg + slen);
      if (rid < 0) continue;
      if ((opt.bsstrand & 1) && bsq_getbss(ix, parent, rbeg) != (opt.bsstrand >> 1)) continue;
      if (n_sn >= ws.cap) { res.status = 1; return res; }
      const int si = n_sn++;
      bsq_snode_t &s = ws.snodes[si];
      s.rbeg = rbeg; s.qbeg = v_beg; s.len = slen; s.next = -1; s.pad_ = 0;
      bool to_add = true;
      if (tree.n_keys > 0) {
        int lower = bsq_bt_lower(tree, rbeg);
        if (lower >= 0 && bsq_merge_seed(opt, l_pac, ws.chains[lower], ws.snodes, si, rid)) to_add = false;
      }
      if (to_add) {
        ++count;
        const int ci = n_ch++;
        bsq_wchain_t &c = ws.chains[ci];
        c.pos = rbeg; c.rid = rid; c.is_alt = ix.ann_is_alt[rid] != 0;
        c.first_rbeg = c.last_rbeg = rbeg; c.first_qbeg = c.last_qbeg = s.qbeg; c.first_len = c.last_len = slen;
        c.head = c.tail = si; c.n = 1; c.xhead = c.xtail = -1; c.xn = 0;
        c.w = 0; c.first = -1; c.kept = 0; c.pad_ = 0;
        bsq_bt_put(tree, ci);
      }
    }
    sa_i += (int64_t)n_pre;
  }

  // ---- mem_chain_flt ----
  int32_t *ord = ws.order;
  int n = bsq_bt_inorder(tree, ord);
  if (n == 0) return res;
  {
    int kk = 0;
    for (int i = 0; i < n; ++i) {
      bsq_wchain_t &c = ws.chains[ord[i]];
      c.first = -1; c.kept = 0;
      c.w = bsq_chain_weight(c, ws.snodes);
      if (c.w >= opt.min_chain_weight) ord[kk++] = ord[i];
    }
    n = kk;
  }
  if (n > 0) {
    bsq_chain_w_desc lt; lt.ch = ws.chains;
    bsq_introsort(ord, (int64_t)n, lt);
    // greedy overlap filter; the list of kept chains reuses the front of out_chains' seed_off
    // field is not available yet, so keep indices in the snode pad of... a plain scan instead:
    ws.chains[ord[0]].kept = 3;
    // `keep[]`: positions (in ord) of chains accepted so far; stored in bnodes memory, which is
    // no longer needed once the in-order walk is done.
    int32_t *keep = reinterpret_cast<int32_t *>(ws.bnodes);
    int n_keep = 0;
    keep[n_keep++] = 0;
    for (int i = 1; i < n; ++i) {
      bsq_wchain_t &ci = ws.chains[ord[i]];
      const int ci_beg = ci.first_qbeg, ci_end = ci.last_qbeg + ci.last_len;
      int large_overlap = 0, k;
      for (k = 0; k < n_keep; ++k) {
        bsq_wchain_t &ck = ws.chains[ord[keep[k]]];
        const int ck_beg = ck.first_qbeg, ck_end = ck.last_qbeg + ck.last_len;
        const int b_max = ck_beg > ci_beg ? ck_beg : ci_beg;
        const int e_min = ck_end < ci_end ? ck_end : ci_end;
        if (e_min > b_max && (!ck.is_alt || ci.is_alt)) {
          const int li = ci_end - ci_beg, lj = ck_end - ck_beg;
          const int min_l = li < lj ? li : lj;
          const float thr = (float)min_l * opt.mask_level;
          if ((float)(e_min - b_max) >= thr && min_l < opt.max_chain_gap) {
            large_overlap = 1;
            if (ck.first < 0) ck.first = i;
            const float wk = (float)ck.w * opt.drop_ratio;
            if ((float)ci.w < wk && ck.w - ci.w >= (opt.min_seed_len << 1)) break;
          }
        }
      }
      if (k == n_keep) {
        keep[n_keep++] = i;
        ci.kept = large_overlap ? 2 : 3;
      }
    }
    for (int i = 0; i < n_keep; ++i) {
      const bsq_wchain_t &c = ws.chains[ord[keep[i]]];
      if (c.first >= 0) ws.chains[ord[c.first]].kept = 1;
    }
    // cap the number of kept=1/2 chains that get extended
    {
      int i; uint32_t kk = 0;
      for (i = 0; i < n; ++i) {
        const int kept = ws.chains[ord[i]].kept;
        if (kept == 0 || kept == 3) continue;
        if (++kk >= (uint32_t)opt.max_chain_extend) break;
      }
      for (; i < n; ++i)
        if (ws.chains[ord[i]].kept < 3) ws.chains[ord[i]].kept = 0;
    }
  }
  // ---- emit kept chains with contiguous seed arrays ----
  int n_out = 0, s_out = 0;
  for (int i = 0; i < n; ++i) {
    const bsq_wchain_t &c = ws.chains[ord[i]];
    if (c.kept == 0) continue;
    bsq_chain_t &o = out_chains[n_out++];
    o.pos = c.pos; o.rid = c.rid; o.w = c.w; o.first = c.first; o.kept = (uint8_t)c.kept; o.is_alt = (uint8_t)c.is_alt;
    o.seed_off = s_out; o.n_seeds = c.n; o.n_extra = c.xn;
    for (int q = 0; q < 6; ++q) o.pad_[q] = 0;
    for (int j = c.head; j >= 0; j = ws.snodes[j].next) {
      bsq_seed_t &d = out_seeds[s_out++];
      d.rbeg = ws.snodes[j].rbeg; d.qbeg = ws.snodes[j].qbeg; d.len = ws.snodes[j].len;
    }
    for (int j = c.xhead; j >= 0; j = ws.snodes[j].next) {
      bsq_seed_t &d = out_seeds[s_out++];
      d.rbeg = ws.snodes[j].rbeg; d.qbeg = ws.snodes[j].qbeg; d.len = ws.snodes[j].len;
    }
  }
  res.n_chains = n_out; res.n_seeds = s_out;
  return res;
}
