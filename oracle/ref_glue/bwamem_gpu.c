/* INTEGRATION.md section 2, compiled: the reference's batch API mem_process_seqs (lib/aln/bwamem.c:432-476) with step 1
 * (kt_for over bis_worker1 = mem_chain -> mem_chain_flt -> mem_chain2region for both conversions) replaced by ONE call
 * into libbsq.so; read clipping, mem_merge_regions, mem_pestat and bis_worker2 (pairing, mapQ, CIGAR, SAM text) are the
 * reference's own code, untouched.
 *
 * TEST INFRASTRUCTURE (oracle/Makefile target _ref/biscuit_ref_gpu): proves that the C ABI of include/bsq.h drops into the
 * reference's own source tree.  The reference file is not copied or patched: it is #included where it lies with its
 * mem_process_seqs renamed, so this translation unit sees its static helpers (read_clipping, bis_worker2, worker_t).
 * tests/test_boundary.py asserts that `biscuit_ref_gpu align` and `biscuit_ref align` print the same SAM. */
#define mem_process_seqs mem_process_seqs_cpu
#include "bwamem.c" /* /root/reference/lib/aln/bwamem.c through -I */
#undef mem_process_seqs
#include "bsq.h"

static bsq_index *g_idx;
static bsq_aligner *g_al;
static const bwt_t *g_bwt_seen;

static void gpu_fatal(const char *what, int rc) {
  err_fatal(what, "%s: %s", bsq_strerror(rc), bsq_last_error());
}

/* device copy of bwaidx_t (bwa.h:42-50) + the aligner for these options; once per process */
static void gpu_init(const mem_opt_t *opt, const bwt_t *bwt, const bntseq_t *bns, const uint8_t *pac) {
  if (g_al && g_bwt_seen == bwt) return;
  bsq_index_desc d;
  memset(&d, 0, sizeof d);
  int w, i;
  for (w = 0; w < 2; ++w) { /* bwt[0] = daughter, bwt[1] = parent (bwa.c:535-536) */
    d.bwt[w] = bwt[w].bwt; d.bwt_words[w] = bwt[w].bwt_size; d.primary[w] = bwt[w].primary;
    for (i = 0; i < 5; ++i) d.L2[w][i] = bwt[w].L2[i];
    d.sa[w] = (const uint64_t *)bwt[w].sa; d.n_sa[w] = bwt[w].n_sa; d.sa_intv[w] = bwt[w].sa_intv;
  }
  d.seq_len = bwt[0].seq_len; d.pac = pac; d.l_pac = bns->l_pac; d.n_seqs = bns->n_seqs;
  int64_t *off = malloc(sizeof(int64_t) * bns->n_seqs);
  int32_t *len = malloc(sizeof(int32_t) * bns->n_seqs), *alt = malloc(sizeof(int32_t) * bns->n_seqs);
  for (i = 0; i < bns->n_seqs; ++i) { off[i] = bns->anns[i].offset; len[i] = bns->anns[i].len; alt[i] = bns->anns[i].is_alt; }
  d.ann_offset = off; d.ann_len = len; d.ann_is_alt = alt;
  const char *dev = getenv("BSQ_DEVICE");
  int rc = bsq_index_upload(&d, dev ? atoi(dev) : 0, &g_idx);
  if (rc) gpu_fatal("bsq_index_upload", rc);
  free(off); free(len); free(alt);
  bsq_opt o;
  bsq_opt_default(&o);
  o.a = opt->a; o.b = opt->b; o.o_del = opt->o_del; o.e_del = opt->e_del; o.o_ins = opt->o_ins; o.e_ins = opt->e_ins;
  o.pen_clip5 = opt->pen_clip5; o.pen_clip3 = opt->pen_clip3; o.w = opt->w; o.zdrop = opt->zdrop;
  o.min_seed_len = opt->min_seed_len; o.split_width = opt->split_width; o.max_occ = opt->max_occ; o.max_chain_gap = opt->max_chain_gap;
  o.min_chain_weight = opt->min_chain_weight; o.max_chain_extend = opt->max_chain_extend; o.max_mem_intv = opt->max_mem_intv;
  o.split_len = (int)(opt->min_seed_len * opt->split_factor + .499); /* memchain.c:55 */
  o.self_ovlp = (opt->flag & MEM_F_SELF_OVLP) != 0; o.bsstrand = opt->bsstrand;
  o.mask_level = opt->mask_level; o.drop_ratio = opt->drop_ratio;
  memcpy(o.ctmat, opt->ctmat, 25); memcpy(o.gamat, opt->gamat, 25);
  rc = bsq_aligner_create(g_idx, &o, &g_al);
  if (rc) gpu_fatal("bsq_aligner_create", rc);
  g_bwt_seen = bwt;
}

void mem_process_seqs(const mem_opt_t *opt, const bwt_t *bwt, const bntseq_t *bns, const uint8_t *pac, int64_t n_processed, int n,
                      bseq1_t *seqs, const mem_pestat_t *pes0) {
  int i, k;
  double ctime = cputime(), rtime = realtime();
  gpu_init(opt, bwt, bns, pac);
  worker_t w;
  w.regs = malloc(n * sizeof(mem_alnreg_v));
  w.opt = opt; w.bwt = bwt; w.bns = bns; w.pac = pac; w.seqs = seqs; w.n_processed = n_processed;
  w.intv_cache = 0;

  /***** Step 1 on the GPU: what bis_worker1 does around mem_align1_core stays here (bwamem.c:311-375) *****/
  const int pe = (opt->flag & MEM_F_PE) != 0;
  int stride = 8;
  for (i = 0; i < n; ++i) {
    if (pe && !(i & 1)) check_paired_read_names(seqs[i].name, seqs[i + 1].name);
    if (!pe || !(i & 1)) read_clipping(&seqs[i], opt->adaptor1, opt->l_adaptor1, opt);
    else read_clipping(&seqs[i], opt->adaptor2, opt->l_adaptor2, opt);
    if (seqs[i].l_seq > stride) stride = seqs[i].l_seq;
  }
  stride = (stride + 7) & ~7;
  /* (read, conversion) tasks in the order the regions are appended: SE daughter, parent (:326-334);
   * PE read 1 parent, daughter (:351-357), read 2 daughter, parent (:368-373); opt->parent restricts */
  int64_t n_tasks = 0;
  int *first_task = malloc(sizeof(int) * (n + 1));
  uint8_t *tseq = calloc((size_t)2 * n + 1, (size_t)stride), *tpar = malloc((size_t)2 * n + 1);
  int32_t *tlen = malloc(sizeof(int32_t) * ((size_t)2 * n + 1));
  for (i = 0; i < n; ++i) {
    int order[2], no = 0;
    if (!pe) {
      if (!(opt->parent & 1) || opt->parent >> 1) order[no++] = 0;
      if (!(opt->parent & 1) || !(opt->parent >> 1)) order[no++] = 1;
    } else if (!(i & 1)) { order[no++] = 1; if (!opt->parent) order[no++] = 0; }
    else { order[no++] = 0; if (!opt->parent) order[no++] = 1; }
    first_task[i] = (int)n_tasks;
    for (k = 0; k < no; ++k) {
      memcpy(tseq + n_tasks * stride, seqs[i].seq, (size_t)seqs[i].l_seq);
      tlen[n_tasks] = seqs[i].l_seq; tpar[n_tasks] = (uint8_t)order[k];
      ++n_tasks;
    }
  }
  first_task[n] = (int)n_tasks;
  bsq_reg *regs = 0;
  int64_t *reg_off = malloc(sizeof(int64_t) * (n_tasks + 1));
  int rc = bsq_align_phase1(g_al, n_tasks, tseq, stride, tlen, tpar, &regs, reg_off);
  if (rc) gpu_fatal("bsq_align_phase1", rc);
  for (i = 0; i < n; ++i) {
    mem_alnreg_v *rv = &w.regs[i];
    kv_init(*rv); rv->n_pri = 0;
    int64_t t, r;
    for (t = first_task[i]; t < first_task[i + 1]; ++t)
      for (r = reg_off[t]; r < reg_off[t + 1]; ++r) {
        mem_alnreg_t *a = kv_pushp(mem_alnreg_t, *rv);
        memset(a, 0, sizeof *a); /* mem_chain2region1 resets every new region (memchain.c:762) */
        a->rb = regs[r].rb; a->re = regs[r].re; a->qb = regs[r].qb; a->qe = regs[r].qe; a->rid = regs[r].rid;
        a->score = regs[r].score; a->truesc = regs[r].truesc; a->w = regs[r].w; a->seedcov = regs[r].seedcov;
        a->seedlen0 = regs[r].seedlen0; a->frac_rep = regs[r].frac_rep; a->bss = regs[r].bss; a->parent = regs[r].parent;
      }
    mem_merge_regions(opt, bns, pac, &seqs[i], rv);
  }
  bsq_free(regs);
  free(reg_off); free(first_task); free(tseq); free(tpar); free(tlen);

  /***** Step 2 and 3: the reference's own code *****/
  if (pe) { if (pes0) w.pes = *pes0; else w.pes = mem_pestat(opt, w.bns, n, w.regs); }
  kt_for(opt->n_threads, bis_worker2, &w, pe ? n >> 1 : n);
  free(w.regs);
  if (bwa_verbose >= 3)
    fprintf(stderr, "[M::%s] Processed %d reads in %.3f CPU sec, %.3f real sec (step 1 on the GPU)\n", __func__, n, cputime() - ctime,
            realtime() - rtime);
}
