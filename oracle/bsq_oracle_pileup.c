/* CPU restatement of BISCUIT's pileup hot path -- TEST INFRASTRUCTURE ONLY (see bsq_oracle.h).
 *
 * Follows, in plain C and in the reference's own order of operations:
 *   process_func read loop      src/pileup.c:707-831   (read filters, mate-overlap skip, per-base events)
 *   get_bsstrand/infer_bsstrand src/bisc_utils.c:163-238
 *   cnt_retention               src/bisc_utils.c:76-122  (incl. its strand quirk: counts C/C when bsstrand=1)
 *   plp_getcnts                 src/pileup.c:372-387
 *   redistribute_cnts           src/pileup.c:339-370
 *   top_mutant                  src/pileup.c:312-334    (qsort by count only; glibc sorts 7 items stably)
 *   plp_format emit rule,       src/pileup.c:415-485, 572
 *     methcallable, DP
 *   fivenuc_context             src/bisc_utils.c:33-74
 * Per-window dispatch (pileup.c:1189-1199) is folded into one [beg,end) range: every read contributes
 * every one of its in-range bases exactly once either way.
 * PINNED against the reference itself: oracle/_ref/biscuit_ref_src is the unmodified src/pileup.c + src/bisc_utils.c
 * compiled over header stand-ins (oracle/ref_shim_src), and tests/test_pileup_ref.py compares VCF bodies byte for byte.
 * Hard clips: the reference advances qpos on H (pileup.c:822-824, bisc_utils.c:108), so after a leading H it pairs
 * reference positions with SEQ bases shifted by the clip length and indexes SEQ/QUAL past their ends for the last ones.
 * Those out-of-range events can never reach the counts (d->rlen < d->qpos + min_dist_end_3p holds for every qpos >
 * l_qseq, pileup.c:383) but they are counted in DP (pileup.c:572).  That behaviour is reproduced exactly, without the
 * out-of-range memory reads.  The only undefined case left is strand inference / retention counting over such bases
 * (reads without YD/ZS/XG tags, or -t given): there the out-of-range bases are skipped here, whereas the reference
 * compares whatever bytes follow SEQ in the record.
 * Nothing here touches utils/stats.h; see bsq_oracle.h. */
#include <limits.h>
#include <stdlib.h>
#include <string.h>
#include "bsq_oracle.h"

enum { M_RET = 0, M_CONV = 1, M_NA = 2 };
enum { B_A = 0, B_C, B_G, B_T, B_N, B_Y, B_R };
enum { CT_HCG = 0, CT_HCHG, CT_HCHH, CT_GCG, CT_GCHG, CT_GCHH, CT_NA };

/* seq_nt16_str "=ACMGRSVTWYHKDBN" folded with nt256char_to_nt256int8_table: A C G T, everything else N */
static const uint8_t nt16_to_nt4[16] = {4, 0, 1, 4, 2, 4, 4, 4, 3, 4, 4, 4, 4, 4, 4, 4};

void bsqo_plp_conf_default(bsqo_plp_conf *c) { /* meth_filter_init, src/bisc_utils.h:95-113; pileup_conf_init :944 */
  memset(c, 0, sizeof *c);
  c->min_base_qual = 20; c->min_read_len = 10; c->min_dist_end_5p = 3; c->min_dist_end_3p = 3; c->min_mapq = 40;
  c->min_score = 40; c->max_nm = 999999; c->max_retention = 999999;
  c->filter_ppair = c->filter_secondary = c->filter_duplicate = c->filter_qcfail = c->filter_doublecnt = 1;
  c->ambi_redist = 1;
}

static inline int read_base(const bsqo_plp_reads *r, int64_t i, int q) {
  uint8_t b = r->seq[r->seq_off[i] + (q >> 1)];
  return nt16_to_nt4[(q & 1) ? (b & 0xf) : (b >> 4)];
}

typedef struct { int32_t meth[3], base[7], dp; } locus_cnt;

int64_t bsqo_plp_region(const bsqo_plp_conf *cf, const uint8_t *ref, int32_t ref_len, int32_t beg, int32_t end,
                        const bsqo_plp_reads *rd, int n_bams, bsqo_plp_rec *out, int64_t cap_loci) {
  if (end > ref_len) end = ref_len; /* the last base of a contig is never piled (pileup.c:1191-1196) */
  if (beg < 1) beg = 1;
  if (end <= beg) return 0;
  const int64_t nl = (int64_t)end - beg;
  locus_cnt *cnt = calloc((size_t)nl * n_bams, sizeof(locus_cnt));
  uint8_t *touched = calloc((size_t)nl, 1);
  int64_t i, n_out = 0;
  for (i = 0; i < rd->n_reads; ++i) {
    const uint32_t *cig = rd->cigar + rd->cigar_off[i];
    const uint8_t *qual = rd->qual + rd->qual_off[i];
    const int nc = rd->n_cigar[i], flag = rd->flag[i], sid = rd->sid[i];
    int k, bsstrand = rd->bss_tag[i];
    uint32_t j;
    if (bsstrand < 0) { /* infer_bsstrand */
      int nC2T = 0, nG2A = 0;
      uint32_t rpos = (uint32_t)rd->pos[i] + 1, qpos = 0;
      for (k = 0; k < nc; ++k) {
        uint32_t op = cig[k] & 0xf, ol = cig[k] >> 4;
        if (op == 0 || op == 7 || op == 8) {
          for (j = 0; j < ol; ++j) {
            if (rpos + j < 1 || rpos + j > (uint32_t)ref_len) continue;
            if (qpos + j >= (uint32_t)rd->l_qseq[i]) continue;
            int rb = ref[rpos + j - 1], qb = read_base(rd, i, qpos + j);
            if (qual[qpos + j] < (uint32_t)cf->min_base_qual) continue;
            if (rb == 1 && qb == 3) nC2T++;
            if (rb == 2 && qb == 0) nG2A++;
          }
          rpos += ol; qpos += ol;
        } else if (op == 1 || op == 4 || op == 5) qpos += ol; /* H advances qpos like the reference; note at the top */
        else if (op == 2) rpos += ol;
        else { free(cnt); free(touched); return -2; } /* the reference abort()s on N/P */
      }
      bsstrand = nC2T >= nG2A ? 0 : 1;
    }
    /* read-level filters (pileup.c:713-729) */
    if (rd->mapq[i] < cf->min_mapq) continue;
    if (rd->l_qseq[i] < 0 || rd->l_qseq[i] < cf->min_read_len) continue;
    if (flag > 0) {
      if (cf->filter_secondary && (flag & 0x100)) continue;
      if (cf->filter_duplicate && (flag & 0x400)) continue;
      if (cf->filter_ppair && (flag & 0x1) && !(flag & 0x2)) continue;
      if (cf->filter_qcfail && (flag & 0x200)) continue;
    }
    if (rd->nm[i] != INT32_MIN && rd->nm[i] > cf->max_nm) continue;
    if (rd->as[i] != INT32_MIN && rd->as[i] < cf->min_score) continue;
    { /* cnt_retention */
      uint32_t c = 0, rpos = (uint32_t)rd->pos[i] + 1, qpos = 0;
      for (k = 0; k < nc; ++k) {
        uint32_t op = cig[k] & 0xf, ol = cig[k] >> 4;
        if (op == 0 || op == 7 || op == 8) {
          for (j = 0; j < ol; ++j) {
            if (rpos + j < 1 || rpos + j > (uint32_t)ref_len) continue;
            if (qpos + j >= (uint32_t)rd->l_qseq[i]) continue;
            int rb = ref[rpos + j - 1], qb = read_base(rd, i, qpos + j);
            if (bsstrand) { if (rb == 1 && qb == 1) c++; } else { if (rb == 2 && qb == 2) c++; }
          }
          rpos += ol; qpos += ol;
        } else if (op == 1 || op == 4 || op == 5) qpos += ol; /* H advances qpos like the reference; note at the top */
        else if (op == 2) rpos += ol;
        else { free(cnt); free(touched); return -2; }
      }
      if (c > (uint32_t)cf->max_retention) continue;
    }
    {
      uint32_t rpos = (uint32_t)rd->pos[i] + 1, qpos = 0, rmpos = (uint32_t)rd->mpos[i] + 1, read_length = 0;
      for (k = 0; k < nc; ++k) { uint32_t op = cig[k] & 0xf; if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) read_length += cig[k] >> 4; }
      uint32_t mate_length = rd->mate_rlen[i] >= 0 ? (uint32_t)rd->mate_rlen[i] : read_length;
      uint32_t rend = rpos + read_length - 1, rmend = rmpos + mate_length - 1;
      for (k = 0; k < nc; ++k) {
        uint32_t op = cig[k] & 0xf, ol = cig[k] >> 4;
        if (op == 0 || op == 7 || op == 8) {
          for (j = 0; j < ol; ++j) {
            uint32_t p = rpos + j;
            if (p < (uint32_t)beg || p >= (uint32_t)end) continue;
            if (cf->filter_doublecnt && (flag & 0x80) && p >= (rpos > rmpos ? rpos : rmpos) && p <= (rend < rmend ? rend : rmend)) continue;
            locus_cnt *lc = &cnt[(int64_t)(p - beg) * n_bams + sid];
            touched[p - beg] = 1;
            lc->dp++;
            if (qpos + j >= (uint32_t)rd->l_qseq[i]) continue; /* past SEQ (leading H): an event for DP, never counted (rlen < qpos + 3') */
            int rb = ref[p - 1], qb = read_base(rd, i, qpos + j);
            int meth, base;
            if (bsstrand) { /* BSC */
              meth = rb == 2 ? (qb == 0 ? M_CONV : qb == 2 ? M_RET : M_NA) : M_NA;
              base = qb == 0 ? B_R : qb;
            } else { /* BSW */
              meth = rb == 1 ? (qb == 3 ? M_CONV : qb == 1 ? M_RET : M_NA) : M_NA;
              base = qb == 3 ? B_Y : qb;
            }
            /* plp_getcnts: base quality and distance-to-end filters */
            uint32_t q7 = qual[qpos + j] & 0x7f, qp = (qpos + j + 1) & 0xffff, rl = (uint32_t)rd->l_qseq[i] & 0xffff;
            if (q7 < (uint32_t)cf->min_base_qual) continue;
            if (qp <= (uint32_t)cf->min_dist_end_5p || rl < qp + (uint32_t)cf->min_dist_end_3p) continue;
            lc->meth[meth]++; lc->base[base]++;
          }
          rpos += ol; qpos += ol;
        } else if (op == 1 || op == 4 || op == 5) qpos += ol; /* H advances qpos like the reference; note at the top */
        else if (op == 2) rpos += ol;
      }
    }
  }
  /* per-locus decisions (plp_format) */
  int64_t l;
  for (l = 0; l < nl; ++l) {
    if (!touched[l]) continue;
    const int32_t rpos = beg + (int32_t)l;
    const int rb = ref[rpos - 1];
    if (rb > 3) continue;
    int sid, b, redist[8][7], all_base[7] = {0}, all_meth[3] = {0}, raw_all[7] = {0};
    if (n_bams > 8) { free(cnt); free(touched); return -3; }
    for (sid = 0; sid < n_bams; ++sid) {
      locus_cnt *lc = &cnt[l * n_bams + sid];
      for (b = 0; b < 7; ++b) { redist[sid][b] = lc->base[b]; raw_all[b] += lc->base[b]; }
    }
    if (cf->ambi_redist) {
      for (sid = 0; sid < n_bams; ++sid) {
        int *c1 = redist[sid];
        if ((rb == B_T || raw_all[B_T]) && raw_all[B_C] == 0 && rb != B_C) { c1[B_T] += c1[B_Y]; c1[B_Y] = 0; }
        if ((rb == B_C || raw_all[B_C]) && raw_all[B_T] == 0 && rb != B_T) { c1[B_C] += c1[B_Y]; c1[B_Y] = 0; }
        if ((rb == B_A || raw_all[B_A]) && raw_all[B_G] == 0 && rb != B_G) { c1[B_A] += c1[B_R]; c1[B_R] = 0; }
        if ((rb == B_G || raw_all[B_G]) && raw_all[B_A] == 0 && rb != B_A) { c1[B_G] += c1[B_R]; c1[B_R] = 0; }
      }
    }
    for (sid = 0; sid < n_bams; ++sid) {
      locus_cnt *lc = &cnt[l * n_bams + sid];
      for (b = 0; b < 3; ++b) all_meth[b] += lc->meth[b];
      for (b = 0; b < 7; ++b) all_base[b] += redist[sid][b];
    }
    /* top_mutant: stable sort by count descending */
    int cm1 = -1, order[7], t, u;
    uint32_t supp[7];
    for (b = 0; b < 7; ++b) { supp[b] = b != B_N ? ((uint32_t)all_base[b] << 4) | (uint32_t)b : 0; order[b] = b; }
    for (t = 1; t < 7; ++t) { /* insertion sort = stable */
      int v = order[t];
      for (u = t; u > 0 && (supp[order[u - 1]] >> 4) < (supp[v] >> 4); --u) order[u] = order[u - 1];
      order[u] = v;
    }
    for (t = 0; t < 7; ++t) {
      int base = supp[order[t]] & 0xf;
      if (base == B_R && (rb == B_A || rb == B_G)) continue;
      if (base == B_Y && (rb == B_C || rb == B_T)) continue;
      if (base != B_N && base != rb && (supp[order[t]] >> 4) > 0) { cm1 = base; break; }
    }
    if (cm1 < 0 && !cf->verbose && all_meth[M_RET] == 0 && all_meth[M_CONV] == 0) continue;
    if (n_out >= cap_loci) { free(cnt); free(touched); return -1; }
    /* context */
    char n5[5] = {'N', 'N', 'N', 'N', 'N'};
    int ctx = CT_NA;
    if (rb == B_C || rb == B_G) {
      int q;
      for (q = 0; q < 5; ++q) {
        int32_t p = rpos - 2 + q; /* 1-based */
        n5[q] = (p >= 1 && p <= ref_len) ? "ACGTN"[ref[p - 1] > 3 ? 4 : ref[p - 1]] : 'N';
      }
      if (rb == B_G) { /* reverse complement */
        char tmp[5];
        for (q = 0; q < 5; ++q) { char c = n5[4 - q]; tmp[q] = c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : c == 'T' ? 'A' : 'N'; }
        memcpy(n5, tmp, 5);
      }
      int has_n = 0;
      for (q = 0; q < 5; ++q) if (n5[q] == 'N') has_n = 1;
      if (!has_n) {
        if (n5[3] == 'G') ctx = n5[1] == 'G' ? CT_GCG : CT_HCG;
        else if (n5[4] == 'G') ctx = n5[1] == 'G' ? CT_GCHG : CT_HCHG;
        else ctx = n5[1] == 'G' ? CT_GCHH : CT_HCHH;
      }
    }
    int any_callable = 0;
    uint8_t callable[8];
    for (sid = 0; sid < n_bams; ++sid) {
      locus_cnt *lc = &cnt[l * n_bams + sid];
      const int *c1 = redist[sid];
      callable[sid] = 0;
      if (lc->meth[M_RET] + lc->meth[M_CONV] > 0) {
        if (rb == B_C) {
          if (c1[B_T] == 0) callable[sid] = 1;
          else if (c1[B_C] > 0 && c1[B_T] / (double)c1[B_C] < 0.05) callable[sid] = 1;
        }
        if (rb == B_G) {
          if (c1[B_A] == 0) callable[sid] = 1;
          else if (c1[B_G] > 0 && c1[B_A] / (double)c1[B_G] < 0.05) callable[sid] = 1;
        }
      }
      if (callable[sid]) any_callable = 1;
    }
    for (sid = 0; sid < n_bams; ++sid) {
      locus_cnt *lc = &cnt[l * n_bams + sid];
      bsqo_plp_rec *o = &out[n_out * n_bams + sid];
      memset(o, 0, sizeof *o);
      o->pos = rpos; o->dp = lc->dp;
      memcpy(o->meth, lc->meth, sizeof o->meth);
      memcpy(o->base, lc->base, sizeof o->base);
      for (b = 0; b < 7; ++b) o->base_redist[b] = redist[sid][b];
      o->rb_code = (uint8_t)rb; o->cm1 = (int8_t)cm1; o->ctx = (uint8_t)ctx; o->methcallable = callable[sid];
      memcpy(o->n5, n5, 5);
      o->any_callable = (uint8_t)any_callable;
    }
    ++n_out;
  }
  free(cnt); free(touched);
  return n_out;
}
