// TEST-ONLY host emulation of the bsq.h ABI.
//
// The per-task device code in biscuit_b200/csrc/bsq_*.h is plain C++ (BSQ_HD); this file compiles
// it with g++ and drives it with sequential loops so the algorithmic logic can be checked against
// the oracle on a machine without a GPU (`pytest -m "not gpu"`).  It is NOT part of the product:
// nothing under biscuit_b200/ loads it, and libbsq.so has no host execution path.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <condition_variable>
#include <mutex>
#include <vector>
#include "../../include/bsq.h"
#include "../../biscuit_b200/csrc/bsq_task.h"
#include "bsq_ksw_scalar.h"
#include "../../biscuit_b200/csrc/bsq_chain_warp.h"
#define BSQ_SEED3_HOSTEMU 1
#include "../../biscuit_b200/csrc/bsq_seed3.cuh"

static_assert(sizeof(bsq_intv) == sizeof(bsq_intv_t), "abi");
static_assert(sizeof(bsq_reg) == sizeof(bsq_reg_t), "abi");
static_assert(sizeof(bsq_opt) == sizeof(bsq_devopt_t), "abi");

struct bsq_index { bsq_devidx_t d; };
struct bsq_aligner {
  const bsq_index *idx; bsq_devopt_t opt; int64_t counters[16];
  // staged execution (bsq_aligner_stage / _run / _fetch): copies of the inputs and the results of the last run
  std::vector<uint8_t> st_seqs, st_par; std::vector<int32_t> st_lens; int32_t st_stride = 0; int64_t st_n = 0;
  bsq_reg *res_regs = nullptr; std::vector<int64_t> res_off; int64_t res_n = -1;
  // the two result slots of the deferred fetch (bsq_aligner_result_slot / _fetch_slot)
  std::mutex slot_mu; std::condition_variable slot_cv; bool claimed[2] = {false, false};
  int out_slot = 0; std::vector<bsq_reg> slot_regs[2]; std::vector<int64_t> slot_off[2]; int64_t slot_tasks[2] = {0, 0}, slot_n[2] = {-1, -1};
};

#include "../../biscuit_b200/csrc/bsq_opt_default.h"

extern "C" {
const char *bsq_strerror(int code) {
  switch (code) { case 0: return "ok"; case BSQ_ENODEV: return "no device"; case BSQ_EINVAL: return "invalid argument";
    case BSQ_EOVERFLOW: return "capacity overflow"; case BSQ_ENOMEM: return "out of memory"; }
  return "unknown";
}
const char *bsq_last_error(void) { return "hostemu"; }

int bsq_index_upload(const bsq_index_desc *h, int device, bsq_index **out) {
  (void)device;
  bsq_index *ix = new bsq_index();
  memset(&ix->d, 0, sizeof ix->d);
  for (int w = 0; w < 2; ++w) {
    bsq_fm_t &f = ix->d.fm[w];
    f.blocks = h->bwt[w]; f.sa = h->sa[w]; f.full_sa = nullptr; f.primary = h->primary[w]; f.seq_len = h->seq_len; f.sa_intv = h->sa_intv[w];
    for (int i = 0; i < 5; ++i) f.L2[i] = h->L2[w][i];
  }
  ix->d.pac = h->pac; ix->d.l_pac = h->l_pac; ix->d.n_seqs = h->n_seqs;
  // the caller may free its contig tables after the call (the CUDA library copies them to the device)
  int64_t *off = (int64_t *)malloc(sizeof(int64_t) * h->n_seqs);
  int32_t *len = (int32_t *)malloc(sizeof(int32_t) * h->n_seqs), *alt = (int32_t *)malloc(sizeof(int32_t) * h->n_seqs);
  memcpy(off, h->ann_offset, sizeof(int64_t) * h->n_seqs);
  memcpy(len, h->ann_len, sizeof(int32_t) * h->n_seqs);
  memcpy(alt, h->ann_is_alt, sizeof(int32_t) * h->n_seqs);
  ix->d.ann_offset = off; ix->d.ann_len = len; ix->d.ann_is_alt = alt;
  *out = ix;
  return 0;
}
void bsq_index_free(bsq_index *ix) {
  if (!ix) return;
  free((void *)ix->d.ann_offset); free((void *)ix->d.ann_len); free((void *)ix->d.ann_is_alt);
  delete ix;
}

int bsq_occ4(const bsq_index *ix, int which, int64_t n, const uint64_t *k, uint64_t *cnt) {
  for (int64_t i = 0; i < n; ++i) bsq_occ4(ix->d.fm[which], k[i], cnt + 4 * i);
  return 0;
}
int bsq_sa_lookup(const bsq_index *ix, int which, int64_t n, const uint64_t *k, uint64_t *pos) {
  for (int64_t i = 0; i < n; ++i) pos[i] = bsq_sa(ix->d.fm[which], k[i]);
  return 0;
}
// the per-pass seeding kernels of bsq_seed3.cuh, each run as one sequential lane: unsorted interval lists
static int seed3_emulate(const bsq_devidx_t &ix0, const bsq_devopt_t &opt, int64_t n, const uint8_t *seqs, int32_t stride, const int32_t *lens,
                         const uint8_t *parent, int pipeline, std::vector<bsq_pk_t> &intv, std::vector<int32_t> &n_intv) {
  bsq_devidx_t ix = ix0;
  std::vector<uint32_t> b32[2];
  for (int w = 0; w < 2; ++w) {
    const uint64_t n_half = 2 * ((ix.fm[w].seq_len + 127) / 128);
    b32[w].assign((n_half + 1) * 8, 0);
    blockDim.x = 1; threadIdx.x = 0;
    for (uint64_t h = 0; h < n_half; ++h) { blockIdx.x = (unsigned)h; k_derive_b32(ix.fm[w].blocks, ix.fm[w].seq_len, n_half, b32[w].data()); }
    ix.fm[w].b32 = b32[w].data();
  }
  long long qcap = 8 * n + 64;
  if (const char *e = getenv("BSQ_SEED_QCAP")) if (atoll(e) > 0) qcap = atoll(e);
  intv.assign((size_t)n * BSQ_MAX_INTV, bsq_pk_t());
  for (int attempt = 0; attempt < 30; ++attempt) {
    std::vector<uint4> cand1((size_t)n * stride + 1), cand2((size_t)qcap * 8 + 1);
    std::vector<s3_call_t> calls1(qcap), calls2(qcap);
    std::vector<s3_item_t> items(qcap);
    s3_q_t q; memset(&q, 0, sizeof q);
    n_intv.assign(n, 0);
    k_s3_fwd<1>(opt, ix, n, seqs, stride, lens, parent, pipeline, cand1.data(), 0ull, nullptr, calls1.data(), (unsigned long long)qcap, items.data(),
                (unsigned long long)qcap, &q, intv.data(), n_intv.data());
    k_s3_greedy(opt, ix, n, seqs, stride, lens, parent, pipeline, &q, intv.data(), n_intv.data());
    k_s3_bwd(opt, ix, seqs, stride, parent, cand1.data(), calls1.data(), (unsigned long long)qcap, 1, items.data(), (unsigned long long)qcap, &q,
             intv.data(), n_intv.data());
    k_s3_fwd<2>(opt, ix, n, seqs, stride, lens, parent, pipeline, cand2.data(), (unsigned long long)qcap * 8, items.data(), calls2.data(),
                (unsigned long long)qcap, items.data(), (unsigned long long)qcap, &q, intv.data(), n_intv.data());
    k_s3_bwd(opt, ix, seqs, stride, parent, cand2.data(), calls2.data(), (unsigned long long)qcap, 2, items.data(), (unsigned long long)qcap, &q,
             intv.data(), n_intv.data());
    if (!q.overflow) return 0;
    qcap *= 2;  // same policy as the CUDA library: grow the queues and seed the batch again
  }
  return BSQ_EOVERFLOW;
}

static bool pk_less(const bsq_pk_t &a, const bsq_pk_t &b) { return a.w1 != b.w1 ? a.w1 < b.w1 : a.w0 < b.w0; }

// Seeds every task twice -- with the sequential state machine (bsq_seed.h) and with the per-pass kernels the GPU runs
// (bsq_seed3.cuh) -- and refuses to answer when the two interval multisets differ.
int bsq_collect_intv(const bsq_index *ix, const bsq_opt *opt_, int64_t n_tasks, const uint8_t *seqs, int32_t stride,
                     const int32_t *lens, const uint8_t *parent, bsq_intv *out, int32_t *n_out) {
  bsq_devopt_t opt; memcpy(&opt, opt_, sizeof opt);
  bsq_seed_scratch_t *scr = new bsq_seed_scratch_t();
  int rc = 0;
  for (int64_t t = 0; t < n_tasks && !rc; ++t) if (lens[t] > BSQ_MAX_READ_LEN) rc = BSQ_EINVAL;
  std::vector<bsq_pk_t> intv3; std::vector<int32_t> n3;
  if (!rc) rc = seed3_emulate(ix->d, opt, n_tasks, seqs, stride, lens, parent, 0, intv3, n3);
  for (int64_t t = 0; t < n_tasks && !rc; ++t) {
    int32_t n_sa;
    bsq_pk_t pk[BSQ_MAX_INTV];
    int n = bsq_task_seed(opt, ix->d, seqs + t * stride, lens[t], parent[t], false, *scr, pk, &n_sa);
    if (n < 0) { rc = BSQ_EOVERFLOW; break; }
    for (int i = 0; i < n; ++i) ((bsq_intv_t *)out)[t * BSQ_MAX_INTV + i] = bsq_pk_unpack(pk[i]);
    n_out[t] = n;
    std::vector<bsq_pk_t> a(pk, pk + n), b;
    if (n3[t] <= BSQ_MAX_INTV) b.assign(intv3.begin() + t * BSQ_MAX_INTV, intv3.begin() + t * BSQ_MAX_INTV + n3[t]);
    std::sort(a.begin(), a.end(), pk_less); std::sort(b.begin(), b.end(), pk_less);
    bool same = n3[t] == n;
    for (int i = 0; same && i < n; ++i) same = a[i].w0 == b[i].w0 && a[i].w1 == b[i].w1;
    if (!same) { fprintf(stderr, "hostemu: task %lld: per-pass seeding kernels give %d intervals, the state machine %d (or different records)\n", (long long)t, n3[t], n); rc = BSQ_EINVAL; }
  }
  delete scr;
  return rc;
}
int bsq_extend_batch(const bsq_opt *opt_, int64_t n_jobs, const uint8_t *qbuf, const int64_t *qoff, const int32_t *qlen,
                     const uint8_t *tbuf, const int64_t *toff, const int32_t *tlen, const uint8_t *is_parent,
                     const int32_t *w, const int32_t *h0, int32_t *out) {
  bsq_devopt_t opt; memcpy(&opt, opt_, sizeof opt);
  bsq_ksw_scratch_t scr;
  for (int64_t j = 0; j < n_jobs; ++j) {
    if (qlen[j] > BSQ_MAX_READ_LEN) return BSQ_EINVAL;
    bsq_qacc_t qa; qa.q = qbuf + qoff[j]; qa.step = 1;
    bsq_tacc_t ta; ta.ix = nullptr; ta.buf = tbuf + toff[j]; ta.p0 = 0; ta.step = 1;
    bsq_ext_result_t r = bsq_ksw_extend(qlen[j], qa, tlen[j], ta, is_parent[j] ? opt.ctmat : opt.gamat, opt.o_del, opt.e_del,
                                        opt.o_ins, opt.e_ins, w[j], opt.pen_clip5, opt.zdrop, h0[j], scr);
    memcpy(out + 6 * j, &r, 24);
  }
  return 0;
}

int bsq_aligner_create(const bsq_index *ix, const bsq_opt *opt, bsq_aligner **out) {
  bsq_aligner *a = new bsq_aligner();
  a->idx = ix; memcpy(&a->opt, opt, sizeof a->opt); memset(a->counters, 0, sizeof a->counters);
  *out = a;
  return 0;
}
void bsq_aligner_destroy(bsq_aligner *a) { if (a) free(a->res_regs); delete a; }
void bsq_free(void *p) { free(p); }
int bsq_aligner_counters(const bsq_aligner *a, int64_t *c, int n) { for (int i = 0; i < n && i < 16; ++i) c[i] = a->counters[i]; return 0; }

int bsq_align_phase1(bsq_aligner *al, int64_t n_tasks, const uint8_t *seqs, int32_t stride, const int32_t *lens,
                     const uint8_t *parent, bsq_reg **regs_out, int64_t *reg_off) {
  const bsq_devopt_t &opt = al->opt;
  const bsq_devidx_t &ix = al->idx->d;
  std::vector<bsq_reg_t> all;
  bsq_seed_scratch_t *scr = new bsq_seed_scratch_t();
  bsq_ksw_scratch_t *ksw = new bsq_ksw_scratch_t();
  std::vector<bsq_pk_t> intv(BSQ_MAX_INTV);
  int rc = 0;
  memset(al->counters, 0, sizeof al->counters);
  for (int64_t t = 0; t < n_tasks && rc == 0; ++t) {
    reg_off[t] = (int64_t)all.size();
    if (lens[t] > BSQ_MAX_READ_LEN) { rc = BSQ_EINVAL; break; }
    const uint8_t *seq = seqs + t * stride;
    int32_t n_sa;
    int n = bsq_task_seed(opt, ix, seq, lens[t], parent[t], true, *scr, intv.data(), &n_sa);
    if (n < 0) { rc = BSQ_EOVERFLOW; break; }
    std::vector<uint64_t> ranks(n_sa + 1), pos(n_sa + 1);
    bsq_task_expand(opt, intv.data(), n, 0, ranks.data());
    for (int i = 0; i < n_sa; ++i) pos[i] = bsq_sa(ix.fm[parent[t]], ranks[i]);
    const int cap = n_sa + 64;
    std::vector<bsq_snode_t> sn(cap); std::vector<bsq_wchain_t> wc(cap); std::vector<bsq_bnode_t> bn(cap + 2);
    std::vector<int32_t> ord(cap); std::vector<bsq_chain_t> och(cap); std::vector<bsq_seed_t> osd(cap);
    bsq_chain_ws_t ws; ws.cap = cap; ws.snodes = sn.data(); ws.chains = wc.data(); ws.bnodes = bn.data(); ws.order = ord.data();
    // same order of attempts as the CUDA path: shared-memory decomposition first, exact B-tree replay as fallback
    bsq_chain_result_t cr;
    static thread_local bsq_cw_smem_t cw;
    if (getenv("BSQ_EMU_NO_CW") || bsq_chain_warp<bsq_cw_scalar>(opt, ix, parent[t], lens[t], intv.data(), n, pos.data(), n_sa, cw, och.data(), osd.data(), cr) != BSQ_CW_OK) {
      cr = bsq_chain_task(opt, ix, parent[t], lens[t], intv.data(), n, pos.data(), ws, och.data(), osd.data());
      al->counters[11]++;
    }
    if (cr.status) { rc = BSQ_EOVERFLOW; break; }
    std::vector<uint64_t> srt(cap); std::vector<bsq_reg_t> rg(cap);
    int nr = bsq_chain2region<bsq_scalar_policy>(opt, ix, parent[t], lens[t], seq, och.data(), cr.n_chains, osd.data(), cr.frac_rep, srt.data(), ksw, rg.data());
    all.insert(all.end(), rg.begin(), rg.begin() + nr);
    al->counters[0]++; al->counters[1] += n; al->counters[2] += n_sa; al->counters[3] += cr.n_chains; al->counters[4] += nr;
  }
  reg_off[n_tasks] = (int64_t)all.size();
  delete scr; delete ksw;
  if (rc) return rc;
  *regs_out = (bsq_reg *)malloc(all.size() * sizeof(bsq_reg) + 8);
  memcpy(*regs_out, all.data(), all.size() * sizeof(bsq_reg));
  return 0;
}

// test hook: chains after mem_chain + mem_chain_flt for one task, flattened like oracle's refp_chain
int64_t hostemu_chain(const bsq_index *ixp, const bsq_opt *opt_, int parent, int len, const uint8_t *seq, int *n_chains,
                      float *frac_rep, int64_t *out, int64_t cap_out) {
  bsq_devopt_t opt; memcpy(&opt, opt_, sizeof opt);
  const bsq_devidx_t &ix = ixp->d;
  bsq_seed_scratch_t *scr = new bsq_seed_scratch_t();
  std::vector<bsq_pk_t> intv(BSQ_MAX_INTV);
  int32_t n_sa;
  int n = bsq_task_seed(opt, ix, seq, len, parent, true, *scr, intv.data(), &n_sa);
  delete scr;
  if (n < 0) return -1;
  std::vector<uint64_t> ranks(n_sa + 1), pos(n_sa + 1);
  bsq_task_expand(opt, intv.data(), n, 0, ranks.data());
  for (int i = 0; i < n_sa; ++i) pos[i] = bsq_sa(ix.fm[parent], ranks[i]);
  const int cap = n_sa + 64;
  std::vector<bsq_snode_t> sn(cap); std::vector<bsq_wchain_t> wc(cap); std::vector<bsq_bnode_t> bn(cap + 2);
  std::vector<int32_t> ord(cap); std::vector<bsq_chain_t> och(cap); std::vector<bsq_seed_t> osd(cap);
  bsq_chain_ws_t ws; ws.cap = cap; ws.snodes = sn.data(); ws.chains = wc.data(); ws.bnodes = bn.data(); ws.order = ord.data();
  bsq_chain_result_t cr;
  static thread_local bsq_cw_smem_t cw;
  if (getenv("BSQ_EMU_NO_CW") || bsq_chain_warp<bsq_cw_scalar>(opt, ix, parent, len, intv.data(), n, pos.data(), n_sa, cw, och.data(), osd.data(), cr) != BSQ_CW_OK)
    cr = bsq_chain_task(opt, ix, parent, len, intv.data(), n, pos.data(), ws, och.data(), osd.data());
  if (cr.status) return -1;
  *n_chains = cr.n_chains; *frac_rep = cr.frac_rep;
  int64_t o = 0;
  for (int i = 0; i < cr.n_chains; ++i) {
    const bsq_chain_t &c = och[i];
    if (o + 8 + 4 * (int64_t)(c.n_seeds + c.n_extra) > cap_out) return -1;
    out[o++] = c.pos; out[o++] = c.rid; out[o++] = c.w; out[o++] = c.kept; out[o++] = c.first; out[o++] = c.is_alt;
    out[o++] = c.n_seeds; out[o++] = c.n_extra;
    for (int j = 0; j < c.n_seeds + c.n_extra; ++j) {
      const bsq_seed_t &s = osd[c.seed_off + j];
      out[o++] = s.rbeg; out[o++] = s.qbeg; out[o++] = s.len; out[o++] = s.len;
    }
  }
  return o;
}
}  // extern "C"

// for the C half of the emulation (hostemu_dp.c)
extern "C" const uint8_t *hostemu_index_pac(const bsq_index *ix, int64_t *l_pac) { *l_pac = ix->d.l_pac; return ix->d.pac; }

// staged execution on top of bsq_align_phase1 (plain host memory stands in for page-locked memory)
extern "C" {
int bsq_aligner_stage(bsq_aligner *al, int64_t n, const uint8_t *seqs, int32_t stride, const int32_t *lens, const uint8_t *parent) {
  al->st_seqs.assign(seqs, seqs + n * stride); al->st_lens.assign(lens, lens + n); al->st_par.assign(parent, parent + n);
  al->st_stride = stride; al->st_n = n; al->res_n = -1;
  return 0;
}
int bsq_aligner_run(bsq_aligner *al, int64_t *n_regs) {
  free(al->res_regs); al->res_regs = nullptr;
  al->res_off.assign(al->st_n + 1, 0);
  if (al->st_n == 0) { al->res_n = 0; if (n_regs) *n_regs = 0; return 0; }
  int rc = bsq_align_phase1(al, al->st_n, al->st_seqs.data(), al->st_stride, al->st_lens.data(), al->st_par.data(), &al->res_regs, al->res_off.data());
  if (rc) return rc;
  al->res_n = al->res_off[al->st_n];
  if (n_regs) *n_regs = al->res_n;
  {
    std::unique_lock<std::mutex> lk(al->slot_mu);
    al->slot_cv.wait(lk, [&] { return !al->claimed[al->out_slot ^ 1]; });
    al->out_slot ^= 1;
  }
  const int sl = al->out_slot;
  al->slot_regs[sl].assign(al->res_regs, al->res_regs + al->res_n); al->slot_off[sl] = al->res_off; al->slot_tasks[sl] = al->st_n; al->slot_n[sl] = al->res_n;
  return 0;
}
int bsq_aligner_release_slot(bsq_aligner *al, int slot) {
  if (!al || slot < 0 || slot > 1) return BSQ_EINVAL;
  { std::lock_guard<std::mutex> lk(al->slot_mu); al->claimed[slot] = false; }
  al->slot_cv.notify_all();
  return 0;
}
int bsq_aligner_result_slot(bsq_aligner *al, int *slot, int64_t *n_tasks, int64_t *n_regs) {
  if (!al || !slot || al->res_n < 0) return BSQ_EINVAL;
  { std::lock_guard<std::mutex> lk(al->slot_mu); al->claimed[al->out_slot] = true; }
  *slot = al->out_slot;
  if (n_tasks) *n_tasks = al->st_n;
  if (n_regs) *n_regs = al->res_n;
  return 0;
}
int bsq_aligner_fetch_slot(bsq_aligner *al, int slot, bsq_reg *regs, int64_t *reg_off) {
  if (!al || slot < 0 || slot > 1 || al->slot_n[slot] < 0 || !reg_off) return BSQ_EINVAL;
  if (al->slot_tasks[slot] == 0) { reg_off[0] = 0; bsq_aligner_release_slot(al, slot); return 0; }
  if (al->slot_n[slot] > 0) memcpy(regs, al->slot_regs[slot].data(), (size_t)al->slot_n[slot] * sizeof(bsq_reg));
  memcpy(reg_off, al->slot_off[slot].data(), (size_t)(al->slot_tasks[slot] + 1) * 8);
  bsq_aligner_release_slot(al, slot);
  return 0;
}
int bsq_aligner_fetch(bsq_aligner *al, bsq_reg *regs, int64_t *reg_off) {
  if (al->res_n < 0) return BSQ_EINVAL;
  if (al->res_n > 0) memcpy(regs, al->res_regs, (size_t)al->res_n * sizeof(bsq_reg));
  memcpy(reg_off, al->res_off.data(), (size_t)(al->st_n + 1) * 8);
  return 0;
}
void bsq_set_wait_mode(int) {}
int bsq_host_alloc(void **p, size_t bytes) { *p = malloc(bytes ? bytes : 1); return *p ? 0 : BSQ_ENOMEM; }
void bsq_host_free(void *p) { free(p); }
}

// Entry points that only exist on the GPU (index construction, pinned memory, staged execution):
// the host emulation says so instead of pretending.
extern "C" {
int bsq_index_build(const uint8_t *, int64_t, int32_t, const int64_t *, const int32_t *, const int32_t *, int, bsq_index **) { return BSQ_ENODEV; }
int bsq_index_sizes(const bsq_index *, uint64_t *, uint64_t *, uint64_t *, uint64_t *, int64_t *) { return BSQ_ENODEV; }
int bsq_index_download(const bsq_index *, int, uint32_t *, uint64_t *) { return BSQ_ENODEV; }
// Pileup: the device kernels have no per-task host form, so the emulation hands the staged reads to the oracle's
// restatement (oracle/bsq_oracle_pileup.c, linked into this test-only library; same record layouts).  What that
// checks on a machine without a GPU is everything around the kernels: BGZF/BAM/BAI/FASTA readers, chunking and the
// carry-over of reads between chunks, VCF text, methylation averages (biscuit_b200/host/bq_pileup.c, bq_bam.c).
// The kernels themselves are compared with the same oracle on the GPU box (tests/test_pileup.py).
}  // extern "C"
extern "C" {
#include "../../oracle/bsq_oracle.h"
}
struct bsq_plp {
  int n_bams = 1;
  std::vector<uint8_t> ref;
  // deep copy of the staged reads (bsq_plp_stage is a host->device copy: the caller may reuse its buffers)
  int64_t n = 0;
  std::vector<int32_t> pos, mpos, mate_rlen, l_qseq, nm, as, n_cigar;
  std::vector<uint16_t> flag;
  std::vector<uint8_t> mapq, sid, seq, qual;
  std::vector<int8_t> bss_tag;
  std::vector<int64_t> cigar_off, seq_off, qual_off;
  std::vector<uint32_t> cigar;
  std::vector<bsq_plp_rec> out;
  int64_t n_out = 0, counters[8] = {0, 0, 0, 0, 0, 0, 0, 0};
};
extern "C" {
void bsq_plp_conf_default(bsq_plp_conf *c) {
  static_assert(sizeof(bsq_plp_conf) == sizeof(bsqo_plp_conf) && sizeof(bsq_plp_rec) == sizeof(bsqo_plp_rec) &&
                sizeof(bsq_plp_reads) == sizeof(bsqo_plp_reads), "oracle and ABI layouts differ");
  bsqo_plp_conf_default(reinterpret_cast<bsqo_plp_conf *>(c));
}
int bsq_plp_create(int, int n_bams, bsq_plp **out) {
  if (!out || n_bams < 1) return BSQ_EINVAL;
  *out = new bsq_plp();
  (*out)->n_bams = n_bams;
  return 0;
}
void bsq_plp_destroy(bsq_plp *p) { delete p; }
int bsq_plp_set_contig(bsq_plp *p, const uint8_t *ref_nt4, int32_t ref_len) {
  if (!p || !ref_nt4 || ref_len <= 0) return BSQ_EINVAL;
  p->ref.assign(ref_nt4, ref_nt4 + ref_len);
  return 0;
}
int bsq_plp_stage(bsq_plp *p, const bsq_plp_reads *r) {
  if (!p || !r || r->n_reads < 0) return BSQ_EINVAL;
  const int64_t n = r->n_reads;
  int64_t cig = 0, sq = 0, ql = 0;
  for (int64_t i = 0; i < n; ++i) {
    if (r->sid[i] >= p->n_bams || r->n_cigar[i] < 0) return BSQ_EINVAL;
    cig = std::max(cig, r->cigar_off[i] + r->n_cigar[i]);
    sq = std::max(sq, r->seq_off[i] + (r->l_qseq[i] + 1) / 2);
    ql = std::max(ql, r->qual_off[i] + r->l_qseq[i]);
  }
  p->n = n;
  p->pos.assign(r->pos, r->pos + n); p->mpos.assign(r->mpos, r->mpos + n); p->mate_rlen.assign(r->mate_rlen, r->mate_rlen + n);
  p->l_qseq.assign(r->l_qseq, r->l_qseq + n); p->nm.assign(r->nm, r->nm + n); p->as.assign(r->as, r->as + n);
  p->flag.assign(r->flag, r->flag + n); p->mapq.assign(r->mapq, r->mapq + n); p->bss_tag.assign(r->bss_tag, r->bss_tag + n);
  p->sid.assign(r->sid, r->sid + n); p->n_cigar.assign(r->n_cigar, r->n_cigar + n);
  p->cigar_off.assign(r->cigar_off, r->cigar_off + n); p->seq_off.assign(r->seq_off, r->seq_off + n); p->qual_off.assign(r->qual_off, r->qual_off + n);
  p->cigar.assign(r->cigar, r->cigar + cig); p->seq.assign(r->seq, r->seq + sq); p->qual.assign(r->qual, r->qual + ql);
  p->counters[0] = n;
  return 0;
}
int bsq_plp_run(bsq_plp *p, const bsq_plp_conf *cf, int32_t beg, int32_t end, int64_t *n_loci) {
  if (!p || !cf || !n_loci || p->ref.empty()) return BSQ_EINVAL;
  const int32_t ref_len = (int32_t)p->ref.size();
  if (end > ref_len) end = ref_len;  // the last base of a contig is never piled (as in the CUDA library)
  if (beg < 1) beg = 1;
  *n_loci = 0; p->n_out = 0;
  if (end <= beg) return 0;
  bsqo_plp_reads rd;
  rd.n_reads = p->n; rd.pos = p->pos.data(); rd.mpos = p->mpos.data(); rd.mate_rlen = p->mate_rlen.data(); rd.l_qseq = p->l_qseq.data();
  rd.nm = p->nm.data(); rd.as = p->as.data(); rd.flag = p->flag.data(); rd.mapq = p->mapq.data(); rd.bss_tag = p->bss_tag.data();
  rd.sid = p->sid.data(); rd.n_cigar = p->n_cigar.data(); rd.cigar_off = p->cigar_off.data(); rd.cigar = p->cigar.data();
  rd.seq_off = p->seq_off.data(); rd.seq = p->seq.data(); rd.qual_off = p->qual_off.data(); rd.qual = p->qual.data();
  const int64_t cap = (int64_t)end - beg;
  p->out.resize((size_t)cap * p->n_bams);
  const int64_t n = bsqo_plp_region(reinterpret_cast<const bsqo_plp_conf *>(cf), p->ref.data(), ref_len, beg, end, &rd, p->n_bams,
                                    reinterpret_cast<bsqo_plp_rec *>(p->out.data()), cap);
  if (n < 0) return BSQ_EOVERFLOW;
  p->n_out = n; *n_loci = n;
  p->counters[1] = cap; p->counters[2] = n;
  return 0;
}
int bsq_plp_fetch(bsq_plp *p, bsq_plp_rec *out) {
  if (!p || !out) return BSQ_EINVAL;
  memcpy(out, p->out.data(), (size_t)p->n_out * p->n_bams * sizeof(bsq_plp_rec));
  return 0;
}
int bsq_plp_counters(const bsq_plp *p, int64_t *c, int n) {
  if (!p || !c) return BSQ_EINVAL;
  for (int i = 0; i < n && i < 8; ++i) c[i] = p->counters[i];
  return 0;
}
}
