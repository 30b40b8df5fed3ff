// SMEM seeding for one (read, conversion) task: the three passes of mem_collect_intv
// (lib/aln/memchain.c:50-106) over bwt_smem1a (lib/aln/bwt.c:307-370) and
// bwt_seed_strategy1 (lib/aln/bwt.c:376-396).  Per-task scratch lives in thread-local arrays
// (interleaved local memory on the GPU), the result list goes to the caller's buffer.
#pragma once
#include "bsq_fm.h"
#include "bsq_sort.h"

// Scratch entry for the forward/backward sweeps: interval + end coordinate on the read.
struct bsq_cand_t {
  uint64_t x0, x1, x2;
  int32_t end, pad_;
};

struct bsq_seed_scratch_t {
  bsq_cand_t a[2][BSQ_MAX_READ_LEN + 1];
  bsq_intv_t one[BSQ_MAX_READ_LEN + 1];  // SMEMs of a single bwt_smem1a call (the reference's `_mem`)
};

BSQ_HD bsq_intv_t bsq_cand2intv(const bsq_cand_t &c) {
  bsq_intv_t r;
  r.x[0] = c.x0; r.x[1] = c.x1; r.x[2] = c.x2; r.info = (uint64_t)(uint32_t)c.end;
  return r;
}
BSQ_HD bsq_cand_t bsq_intv2cand(const bsq_intv_t &v, int end) {
  bsq_cand_t c;
  c.x0 = v.x[0]; c.x1 = v.x[1]; c.x2 = v.x[2]; c.end = end; c.pad_ = 0;
  return c;
}

// All SMEMs through query position x whose interval size is >= min_intv (bwt.c:307-370 with
// max_intv == 0, the only way the reference calls it).  q is the converted read.  Returns the
// next x; the SMEMs, sorted by start, are left in scr.one[0..*n_out).
BSQ_HD int bsq_smem1(const bsq_fm_t &fm, const bsq_fm_t &fmc, int len, const uint8_t *q, int x, int min_intv,
                     bsq_seed_scratch_t &scr, int *n_out) {
  *n_out = 0;
  if (q[x] > 3) return x + 1;
  if (min_intv < 1) min_intv = 1;
  bsq_cand_t *curr = scr.a[0], *prev = scr.a[1];
  int n_curr = 0, n_prev;
  bsq_intv_t ik, ok[4];
  bsq_set_intv(fm, fmc, q[x], ik);
  int ik_end = x + 1, i;
  // forward sweep: remember the interval every time its size changes
  for (i = x + 1; i < len; ++i) {
    if (q[i] < 4) {
      int c = 3 - q[i];
      bsq_extend<0>(fmc, ik, ok);
      if (ok[c].x[2] != ik.x[2]) {
        curr[n_curr++] = bsq_intv2cand(ik, ik_end);
        if (ok[c].x[2] < (uint64_t)min_intv) break;
      }
      ik = ok[c];
      ik_end = i + 1;
    } else {
      curr[n_curr++] = bsq_intv2cand(ik, ik_end);
      break;
    }
  }
  if (i == len) curr[n_curr++] = bsq_intv2cand(ik, ik_end);
  // longest matches first
  for (int a = 0, b = n_curr - 1; a < b; ++a, --b) { bsq_cand_t t = curr[a]; curr[a] = curr[b]; curr[b] = t; }
  const int ret = curr[0].end;
  { bsq_cand_t *t = curr; curr = prev; prev = t; }
  n_prev = n_curr;
  // backward sweep
  int n_mem = 0;
  for (i = x - 1; i >= -1; --i) {
    const int c = i < 0 ? -1 : (q[i] < 4 ? q[i] : -1);
    n_curr = 0;
    for (int j = 0; j < n_prev; ++j) {
      const bsq_cand_t &p = prev[j];
      bool stop = c < 0;
      if (!stop) {
        bsq_extend<1>(fm, bsq_cand2intv(p), ok);
        stop = ok[c].x[2] < (uint64_t)min_intv;
      }
      if (stop) {
        // cannot be extended further to the left: a MEM unless a longer one was kept in this round
        if (n_curr == 0) {
          if (n_mem == 0 || (uint32_t)(i + 1) < (uint32_t)(scr.one[n_mem - 1].info >> 32)) {
            bsq_intv_t m = bsq_cand2intv(p);
            m.info |= (uint64_t)(i + 1) << 32;
            scr.one[n_mem++] = m;
          }
        }
      } else if (n_curr == 0 || ok[c].x[2] != curr[n_curr - 1].x2) {
        curr[n_curr++] = bsq_intv2cand(ok[c], p.end);
      }
    }
    if (n_curr == 0) break;
    { bsq_cand_t *t = curr; curr = prev; prev = t; }
    n_prev = n_curr;
  }
  for (int a = 0, b = n_mem - 1; a < b; ++a, --b) { bsq_intv_t t = scr.one[a]; scr.one[a] = scr.one[b]; scr.one[b] = t; }
  *n_out = n_mem;
  return ret;
}

// bwt_seed_strategy1 (bwt.c:376-396).  m.x[2] == 0 when nothing was found.
BSQ_HD int bsq_seed_strategy1(const bsq_fm_t &fm, const bsq_fm_t &fmc, int len, const uint8_t *q, int x, int min_len,
                              int max_intv, bsq_intv_t &m) {
  m.x[0] = m.x[1] = m.x[2] = m.info = 0;
  if (q[x] > 3) return x + 1;
  bsq_intv_t ik, ok[4];
  bsq_set_intv(fm, fmc, q[x], ik);
  for (int i = x + 1; i < len; ++i) {
    if (q[i] >= 4) return i + 1;
    int c = 3 - q[i];
    bsq_extend<0>(fmc, ik, ok);
    if (ok[c].x[2] < (uint64_t)max_intv && i - x >= min_len) {
      m = ok[c];
      m.info = (uint64_t)x << 32 | (uint32_t)(i + 1);
      return i + 1;
    }
    ik = ok[c];
  }
  return len;
}

struct bsq_intv_less {
  BSQ_HD bool operator()(const bsq_intv_t &a, const bsq_intv_t &b) const { return a.info < b.info; }
};

// mem_collect_intv (memchain.c:50-106).  `out` has room for `cap` intervals; returns the number
// found, or -1 when `cap` is too small.  On return out[] is sorted the way ks_introsort leaves it.
BSQ_HD int bsq_collect_intv(const bsq_devopt_t &opt, const bsq_fm_t &fm, const bsq_fm_t &fmc, int len, const uint8_t *q,
                            bsq_seed_scratch_t &scr, bsq_intv_t *out, int cap) {
  int n = 0, x = 0, n1;
  const int start_width = opt.self_ovlp ? 2 : 1;
  // pass 1: every SMEM of at least min_seed_len
  while (x < len) {
    if (q[x] < 4) {
      x = bsq_smem1(fm, fmc, len, q, x, start_width, scr, &n1);
      for (int i = 0; i < n1; ++i) {
        const bsq_intv_t &m = scr.one[i];
        if ((uint32_t)m.info - (uint32_t)(m.info >> 32) >= (uint32_t)opt.min_seed_len) {
          if (n == cap) return -1;
          out[n++] = m;
        }
      }
    } else ++x;
  }
  // pass 2: re-seed from the middle of long, rare SMEMs
  const int old_n = n;
  for (int k = 0; k < old_n; ++k) {
    const int start = (int)(out[k].info >> 32), end = (int32_t)out[k].info;
    if (end - start < opt.split_len || out[k].x[2] > (uint64_t)opt.split_width) continue;
    bsq_smem1(fm, fmc, len, q, (start + end) >> 1, (int)(out[k].x[2] + 1), scr, &n1);
    for (int i = 0; i < n1; ++i) {
      const bsq_intv_t &m = scr.one[i];
      if ((uint32_t)m.info - (uint32_t)(m.info >> 32) >= (uint32_t)opt.min_seed_len) {
        if (n == cap) return -1;
        out[n++] = m;
      }
    }
  }
  // pass 3: greedy forward seeds with fewer than max_mem_intv occurrences
  if (opt.max_mem_intv > 0) {
    x = 0;
    while (x < len) {
      if (q[x] < 4) {
        bsq_intv_t m;
        x = bsq_seed_strategy1(fm, fmc, len, q, x, opt.min_seed_len, opt.max_mem_intv, m);
        if (m.x[2] > 0) {
          if (n == cap) return -1;
          out[n++] = m;
        }
      } else ++x;
    }
  }
  bsq_introsort(out, n, bsq_intv_less());
  return n;
}
