// bsq_dp.cu -- the two dynamic-programming steps of the aligner's phase 2 as batched sm_100a kernels:
//
//   k_cigar   one warp per mem_alnreg_setSAM call (lib/aln/mem_alnreg.c:40-123): the band-doubling loop around
//             bis_bwa_gen_cigar2 (lib/aln/bwa.c:290-428) = banded global alignment with traceback (ksw_global2,
//             lib/aln/ksw.c:504-606) + MD text + NM / ZC / ZR / bss_u, and the CIGAR clean-up of setSAM (leading /
//             trailing deletion dropped, soft clips added).
//   k_matesw  one warp per ksw_align2 call of mate rescue (lib/aln/ksw.c:343-365; mem_matesw, mem_alnreg.c:395-493):
//             the 16 x u8 / 8 x i16 striped SSE2 kernels ksw_u8 / ksw_i16 (ksw.c:111-334) re-enacted with one LANE per
//             SSE element, so that the striping-dependent results (E before the lazy-F correction, unsigned
//             saturation, lazy-F early exit) come out as in the reference.
//
// Mapping of ksw_global2 onto a warp: rows are processed one after the other, the band [beg,end) of a row striped
// over the lanes 32 columns at a time.  In ksw_global2 both gap states are fed from the diagonal term only
// (t = H(i-1,j-1)+s - gapoe, ksw.c:561-569), so the horizontal state F(i,j) is a max-plus prefix scan over values that
// are all known when the row starts: with g(j) = F_in(j) + j*e_ins, g(j+1) = max(g(j), M(j) - oe_ins + (j+1)*e_ins) --
// a 5-step shuffle scan per 32 columns, carry between chunks.  All values are the reference's int32 values (the
// -2^30 "minus infinity" terms included: max-plus over integers is exact), hence so are the direction bits.
// H/E of the row live in shared memory, the direction matrix z in a per-warp HBM scratch slice (it stays in L2),
// the traceback and the MD text are written by lane 0, mismatches of an M run found 32 bases at a time by ballot.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#include "../../include/bsq.h"
#include "bsq_internal.h"

#define DP_WARPS 4                 // warps per CTA
#define DP_MAX_RLEN 1024           // longest reference span of a CIGAR job
#define DP_ZCAP (128 * 1024)       // direction bytes per warp (n_col * tlen)
#define DP_MAX_CIG (DP_MAX_RLEN + BSQ_MAX_READ_LEN + 8)
#define DP_MD_CAP 8192
#define DP_MAX_TLEN 8192           // longest mate-rescue window
#define DP_NEG_INF (-0x40000000)

#define CKD(call)                                                                                       \
  do {                                                                                                  \
    cudaError_t e_ = (call);                                                                            \
    if (e_ != cudaSuccess) {                                                                            \
      bsq_set_error("%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_));                 \
      return e_ == cudaErrorMemoryAllocation ? BSQ_ENOMEM : BSQ_ENODEV;                                 \
    }                                                                                                   \
  } while (0)

namespace {

struct DpBuf {  // grow-only device buffer
  void *p = nullptr;
  size_t cap = 0;
  int reserve(size_t bytes) {
    if (bytes <= cap) return 0;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    const size_t want = bytes + bytes / 4 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) { bsq_set_error("cudaMalloc(%zu): %s", want, cudaGetErrorString(e)); return BSQ_ENOMEM; }
    cap = want;
    return 0;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
  template <typename T> T *as() const { return reinterpret_cast<T *>(p); }
};

struct HostBuf {  // grow-only page-locked host buffer
  void *p = nullptr;
  size_t cap = 0;
  int reserve(size_t bytes) {
    if (bytes <= cap) return 0;
    if (p) cudaFreeHost(p);
    p = nullptr; cap = 0;
    const size_t want = bytes + bytes / 4 + 4096;
    if (cudaMallocHost(&p, want) != cudaSuccess) { bsq_set_error("cudaMallocHost(%zu) failed", want); return BSQ_ENOMEM; }
    cap = want;
    return 0;
  }
  void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

// per-warp HBM scratch of k_cigar
struct CigScratch {
  uint8_t z[DP_ZCAP];
  uint32_t cig[DP_MAX_CIG];
  char md[DP_MD_CAP];
};

// device-side counters of one launch
struct DpCtr {
  unsigned long long next_job, blob_used, cells, ungapped;
};

__device__ __forceinline__ int dp_pac_base(const uint8_t *pac, int64_t l) { return pac[l >> 2] >> ((~l & 3) << 1) & 3; }
// base at forward-reverse coordinate pos in [0, 2*l_pac): bns_get_seq (lib/aln/bntseq.c:402-422)
__device__ __forceinline__ int dp_ref_base(const bsq_devidx_t &ix, int64_t pos) {
  return pos < ix.l_pac ? dp_pac_base(ix.pac, pos) : 3 - dp_pac_base(ix.pac, (ix.l_pac << 1) - 1 - pos);
}
__device__ __forceinline__ int dp_iabs(int v) { return v < 0 ? -v : v; }

// ---------------------------------------------------------------------------------------------------------------
// ksw_global2 on a warp.  q: lq codes, r: rlen codes (shared memory), H/E: lq+1 ints each (shared memory).
// Returns the score in all lanes; the CIGAR (already in left-to-right order) in cig[0..*n_cig).
// ---------------------------------------------------------------------------------------------------------------
__device__ int dp_global_warp(int lq, const uint8_t *q, int rlen, const uint8_t *r, const int8_t *mat, int o_del, int e_del, int o_ins, int e_ins,
                              int w, int *H, int *E, uint8_t *z, uint32_t *cig, int *n_cig, unsigned long long *cells) {
  const int lane = threadIdx.x & 31;
  const int oe_del = o_del + e_del, oe_ins = o_ins + e_ins;
  const int n_col = lq < 2 * w + 1 ? lq : 2 * w + 1;
  // first row (ksw.c:529-533)
  for (int j = lane; j <= lq; j += 32) {
    int h;
    if (j == 0) h = 0;
    else if (j <= w) h = -(o_ins + e_ins * j);
    else h = DP_NEG_INF;
    H[j] = h; E[j] = DP_NEG_INF;
  }
  __syncwarp();
  unsigned long long ncell = 0;
  for (int i = 0; i < rlen; ++i) {
    const int beg = i > w ? i - w : 0, end = i + w + 1 < lq ? i + w + 1 : lq;
    const int8_t *row = mat + 5 * r[i];
    int prev_h = beg == 0 ? -(o_del + e_del * (i + 1)) : DP_NEG_INF;  // h1 before the first column
    int carry = DP_NEG_INF + beg * e_ins;                              // g(beg) = F_in(beg) + beg*e_ins, F_in(beg) = -inf
    uint8_t *zi = z + (size_t)i * n_col;
    if (end > beg) ncell += (unsigned)(end - beg);
    for (int jb = beg; jb < end; jb += 32) {
      const int j = jb + lane;
      const bool act = j < end;
      int hj = 0, e = 0, sc = 0;
      if (act) { hj = H[j]; e = E[j]; sc = row[q[j]]; }
      const int m = hj + sc;
      // inclusive prefix max of c(k) = M(k) - oe_ins + (k+1)*e_ins over the lanes of the chunk
      int incl = act ? m - oe_ins + (j + 1) * e_ins : (int)0x80000000;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl = incl > v ? incl : v;
      }
      int g = __shfl_up_sync(0xffffffffu, incl, 1);  // exclusive
      if (lane == 0) g = (int)0x80000000;
      g = g > carry ? g : carry;
      const int f = g - j * e_ins;  // F_in(j)
      uint8_t d = m >= e ? 0 : 1;
      int h = m >= e ? m : e;
      d = h >= f ? d : 2;
      h = h >= f ? h : f;
      int t = m - oe_del;
      int e2 = e - e_del;
      d |= e2 > t ? 1 << 2 : 0;
      e2 = e2 > t ? e2 : t;
      t = m - oe_ins;
      const int f2 = f - e_ins;
      d |= f2 > t ? 2 << 4 : 0;
      int left = __shfl_up_sync(0xffffffffu, h, 1);
      if (lane == 0) left = prev_h;
      if (act) { H[j] = left; E[j] = e2; zi[j - beg] = d; }
      const int nact = end - jb < 32 ? end - jb : 32;
      prev_h = __shfl_sync(0xffffffffu, h, nact - 1);
      const int ctot = __shfl_sync(0xffffffffu, incl, nact - 1);
      carry = carry > ctot ? carry : ctot;
    }
    if (lane == 0) { H[end] = prev_h; E[end] = DP_NEG_INF; }
    __syncwarp();
  }
  if (lane == 0 && cells) atomicAdd(cells, ncell);
  const int score = H[lq];
  __syncwarp();
  // traceback (ksw.c:584-602) by lane 0; the direction bytes were written by all lanes of this warp
  int n = 0;
  if (lane == 0) {
    int i = rlen - 1, k = (i + w + 1 < lq ? i + w + 1 : lq) - 1, which = 0;
    while (i >= 0 && k >= 0) {
      const int b = i > w ? i - w : 0;
      which = __ldcg(z + (size_t)i * n_col + (k - b)) >> (which << 1) & 3;
      int op;
      if (which == 0) { op = 0; --i; --k; }
      else if (which == 1) { op = 2; --i; }
      else { op = 1; --k; }
      if (n == 0 || (cig[n - 1] & 0xf) != (uint32_t)op) cig[n++] = 1u << 4 | (uint32_t)op;
      else cig[n - 1] += 1u << 4;
    }
    if (i >= 0) { if (n == 0 || (cig[n - 1] & 0xf) != 2u) cig[n++] = (uint32_t)(i + 1) << 4 | 2u; else cig[n - 1] += (uint32_t)(i + 1) << 4; }
    if (k >= 0) { if (n == 0 || (cig[n - 1] & 0xf) != 1u) cig[n++] = (uint32_t)(k + 1) << 4 | 1u; else cig[n - 1] += (uint32_t)(k + 1) << 4; }
    for (int a = 0; a < n >> 1; ++a) { const uint32_t t = cig[a]; cig[a] = cig[n - 1 - a]; cig[n - 1 - a] = t; }
  }
  n = __shfl_sync(0xffffffffu, n, 0);
  __syncwarp();
  *n_cig = n;
  return score;
}

// decimal digits of a non-negative number, written by lane 0; returns the new length (all lanes)
__device__ __forceinline__ int dp_put_num(char *md, int l, int v, bool writer) {
  int nd = 1;
  for (int t = v; t >= 10; t /= 10) ++nd;
  if (writer) { int t = v; for (int k = nd - 1; k >= 0; --k) { md[l + k] = (char)('0' + t % 10); t /= 10; } }
  return l + nd;
}

__global__ void __launch_bounds__(32 * DP_WARPS) k_cigar(const __grid_constant__ bsq_devopt_t opt, const __grid_constant__ bsq_devidx_t ix, int64_t n_jobs,
                                                          const bsq_cigar_job *jobs, const uint8_t *seqs, int32_t stride, const int32_t *lens,
                                                          bsq_cigar_res *res, uint32_t *blob, uint64_t blob_cap, CigScratch *scratch, DpCtr *ctr) {
  __shared__ int s_H[DP_WARPS][BSQ_MAX_READ_LEN + 2], s_E[DP_WARPS][BSQ_MAX_READ_LEN + 2];
  __shared__ uint8_t s_q[DP_WARPS][BSQ_MAX_READ_LEN + 8], s_r[DP_WARPS][DP_MAX_RLEN + 8];
  __shared__ int8_t s_mat[2][32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (threadIdx.x < 25) { s_mat[0][threadIdx.x] = opt.gamat[threadIdx.x]; s_mat[1][threadIdx.x] = opt.ctmat[threadIdx.x]; }
  __syncthreads();
  CigScratch *scr = scratch + ((size_t)blockIdx.x * DP_WARPS + wid);
  uint8_t *q = s_q[wid], *r = s_r[wid];
  for (;;) {
    long long jid = 0;
    if (lane == 0) jid = (long long)atomicAdd(&ctr->next_job, 1ull);
    jid = __shfl_sync(0xffffffffu, jid, 0);
    if (jid >= n_jobs) break;
    const bsq_cigar_job jb = jobs[jid];
    bsq_cigar_res out;
    out.n_cigar = 0; out.NM = -1; out.ZC = 0; out.ZR = 0; out.score = 0; out.lead_del = 0; out.bss_u = 0; out.off = 0;
    const int lq = jb.qe - jb.qb;
    const int64_t rlen64 = jb.re - jb.rb;
    const bool rev = jb.rb >= ix.l_pac;
    // bis_bwa_gen_cigar2 gives up on empty input and on a span bridging the two strands (bwa.c:300,305-306)
    bool none = lq <= 0 || jb.rb >= jb.re || (jb.rb < ix.l_pac && jb.re > ix.l_pac) || jb.rb < 0 || jb.re > ix.l_pac << 1;
    bool unsupported = !none && (rlen64 > DP_MAX_RLEN || lq > BSQ_MAX_READ_LEN || jb.qb < 0 || jb.qe > lens[jb.row]);
    int n_cig = 0, score = 0;
    const int rlen = (int)rlen64;
    const int8_t *mat = s_mat[jb.parent ? 1 : 0];
    if (!none && !unsupported) {
      // query and reference in the orientation of the forward strand (bwa.c:311-312: both reversed for a reverse-strand hit)
      const uint8_t *rowp = seqs + (size_t)jb.row * stride;
      for (int x = lane; x < lq; x += 32) { const uint8_t c = rev ? rowp[jb.qe - 1 - x] : rowp[jb.qb + x]; q[x] = c < 5 ? c : 4; }
      for (int y = lane; y < rlen; y += 32) r[y] = (uint8_t)dp_ref_base(ix, rev ? jb.re - 1 - y : jb.rb + y);
      __syncwarp();
      int w = jb.w, last_sc = -(1 << 30);
      for (int it = 0; it < 3; ++it, w <<= 1, last_sc = score) {  // mem_alnreg.c:60-70
        w = w < opt.w << 2 ? w : opt.w << 2;
        if (lq == rlen && w == 0) {  // ungapped (bwa.c:314-322)
          int s = 0;
          for (int x = lane; x < lq; x += 32) s += mat[r[x] * 5 + q[x]];
#pragma unroll
          for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
          score = s; n_cig = 1;
          if (lane == 0) { scr->cig[0] = (uint32_t)lq << 4; if (it == 0) atomicAdd(&ctr->ungapped, 1ull); }
          __syncwarp();
        } else {
          int max_ins = (int)((double)(((lq + 1) >> 1) * mat[0] - opt.o_ins) / opt.e_ins + 1.);
          int max_del = (int)((double)(((lq + 1) >> 1) * mat[0] - opt.o_del) / opt.e_del + 1.);
          int max_gap = max_ins > max_del ? max_ins : max_del;
          max_gap = max_gap > 1 ? max_gap : 1;
          int w2 = (max_gap + dp_iabs(rlen - lq) + 1) >> 1;
          w2 = w2 < w ? w2 : w;
          const int min_w = dp_iabs(rlen - lq) + 3;
          w2 = w2 > min_w ? w2 : min_w;
          const int n_col = lq < 2 * w2 + 1 ? lq : 2 * w2 + 1;
          if ((size_t)n_col * rlen + 1 > DP_ZCAP) { unsupported = true; break; }
          score = dp_global_warp(lq, q, rlen, r, mat, opt.o_del, opt.e_del, opt.o_ins, opt.e_ins, w2, s_H[wid], s_E[wid], scr->z, scr->cig, &n_cig,
                                 &ctr->cells);
        }
        if (score == last_sc) break;
        if (w == opt.w << 2) break;
        if (score >= jb.truesc - opt.a) break;
      }
    }
    if (unsupported) out.n_cigar = -1;
    else if (!none && n_cig > 0) {
      // ---- MD, NM, ZC, ZR, bss_u (bwa.c:343-419) ----
      const char *int2base = rev ? "TGCAN" : "ACGTN";
      char *md = scr->md;
      const bool wr = lane == 0;
      int x = 0, y = 0, u = 0, l = 0, n_mm = 0, n_gap = 0, n_conv_ct = 0, n_ret_c = 0, n_conv_ga = 0, n_ret_g = 0;
      for (int k = 0; k < n_cig; ++k) {
        const uint32_t cw = __ldcg(scr->cig + k);
        const int op = cw & 0xf, len = (int)(cw >> 4);
        if (op == 0) {
          for (int i0 = 0; i0 < len; i0 += 32) {
            const int i = i0 + lane;
            const bool act = i < len;
            const int qv = act ? q[x + i] : 0, rv = act ? r[y + i] : 0;
            const bool eq = act && qv == rv, ne = act && qv != rv;
            unsigned mm = __ballot_sync(0xffffffffu, ne);
            n_ret_c += __popc(__ballot_sync(0xffffffffu, eq && qv == 1));
            n_ret_g += __popc(__ballot_sync(0xffffffffu, eq && qv == 2));
            const int ct = __popc(__ballot_sync(0xffffffffu, ne && jb.parent && qv == 3 && rv == 1));
            const int ga = __popc(__ballot_sync(0xffffffffu, ne && !jb.parent && qv == 0 && rv == 2));
            n_conv_ct += ct; n_conv_ga += ga; n_mm += __popc(mm) - ct - ga;
            int prev = 0;
            while (mm) {
              const int b = __ffs(mm) - 1;
              u += b - prev;
              l = dp_put_num(md, l, u, wr);
              if (wr) md[l] = int2base[r[y + i0 + b]];
              ++l; u = 0; prev = b + 1;
              mm &= mm - 1;
            }
            u += (len - i0 < 32 ? len - i0 : 32) - prev;
          }
          x += len; y += len;
        } else if (op == 2) {
          if (k > 0 && k < n_cig - 1) {
            l = dp_put_num(md, l, u, wr);
            if (wr) md[l] = '^';
            ++l;
            for (int i = lane; i < len; i += 32) md[l + i] = int2base[r[y + i]];
            l += len; u = 0; n_gap += len;
          }
          y += len;
        } else if (op == 1) { x += len; n_gap += len; }
      }
      l = dp_put_num(md, l, u, wr);
      if (wr) md[l] = 0;
      // ---- setSAM clean-up (mem_alnreg.c:86-108) ----
      int first = 0, last = n_cig;
      const uint32_t c0 = __ldcg(scr->cig), cl = __ldcg(scr->cig + n_cig - 1);
      if ((c0 & 0xf) == 2) { out.lead_del = (int)(c0 >> 4); first = 1; }
      else if ((cl & 0xf) == 2) last = n_cig - 1;
      const int n_final = (jb.clip5 ? 1 : 0) + (last - first) + (jb.clip3 ? 1 : 0);
      const uint32_t words = (uint32_t)n_final + (uint32_t)((l + 1 + 3) >> 2);
      unsigned long long off = 0;
      if (lane == 0) off = atomicAdd(&ctr->blob_used, (unsigned long long)words);
      off = __shfl_sync(0xffffffffu, off, 0);
      __syncwarp();
      if (off + words <= blob_cap) {
        uint32_t *dst = blob + off;
        int o = 0;
        if (jb.clip5) { if (lane == 0) dst[0] = (uint32_t)jb.clip5 << 4 | 3u; o = 1; }
        for (int k = first + lane; k < last; k += 32) dst[o + k - first] = __ldcg(scr->cig + k);
        o += last - first;
        if (jb.clip3) { if (lane == 0) dst[o] = (uint32_t)jb.clip3 << 4 | 3u; ++o; }
        char *mdst = (char *)(dst + o);
        for (int k = lane; k <= l; k += 32) mdst[k] = __ldcg(md + k);
      }
      out.n_cigar = n_final; out.NM = n_mm + n_gap;
      out.ZC = jb.parent ? n_conv_ct : n_conv_ga;
      out.ZR = jb.parent ? n_ret_c : n_ret_g;
      out.bss_u = (n_conv_ct == 0 && n_conv_ga == 0) ? 1 : 0;
      out.score = score; out.off = (uint32_t)off;
    }
    if (lane == 0) res[jid] = out;
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------------------------
// mate rescue: ksw_u8 / ksw_i16 with one lane per SSE element.  P lanes (16 / 8) carry the striped vectors; element
// (segment j, lane l) is query position j + l*slen.  Arrays H0 / H1 / E / Hmax: [j*P + l] in shared memory.
// ---------------------------------------------------------------------------------------------------------------
struct SwRes { int score, te, qe, score2, te2; };

__device__ __forceinline__ int dp_subs(int a, int b) { return a > b ? a - b : 0; }  // unsigned saturating subtraction

// tget(i): target base of row i
template <bool IS8, typename TGet>
__device__ SwRes dp_sw_striped(int qlen, const uint8_t *q, int tlen, TGet tget, const int8_t *mat, int o_del, int e_del, int o_ins, int e_ins, int xtra,
                               int *A0, int *A1, int *E, int *Hmax, unsigned long long *bl) {
  constexpr int P = IS8 ? 16 : 8;
  constexpr unsigned PM = IS8 ? 0xffffu : 0xffu;
  const int lane = threadIdx.x & 31;
  const bool on = lane < P;
  const int l = on ? lane : 0;  // idle lanes shadow lane 0 without storing
  const int slen = (qlen + P - 1) / P;
  const int oe_del = o_del + e_del, oe_ins = o_ins + e_ins;
  int mn = 127, mxm = 0;
  for (int a = 0; a < 25; ++a) { mn = mat[a] < mn ? mat[a] : mn; mxm = mat[a] > mxm ? mat[a] : mxm; }
  const int shift = (256 - mn) & 0xff;
  const int minsc = (xtra & 0x40000) ? xtra & 0xffff : 0x10000, endsc = (xtra & 0x20000) ? xtra & 0xffff : 0x10000;
  SwRes res; res.score = 0; res.te = -1; res.qe = -1; res.score2 = -1; res.te2 = -1;
  int *H0 = A0, *H1 = A1;
  if (on) for (int j = 0; j < slen; ++j) { H0[j * P + l] = 0; H1[j * P + l] = 0; E[j * P + l] = 0; Hmax[j * P + l] = 0; }
  __syncwarp();
  int n_b = 0, gmax = 0, te = -1;
  unsigned long long b_last = 0;
  int tblk = 4;
  for (int i = 0; i < tlen; ++i) {
    if ((i & 31) == 0) tblk = i + lane < tlen ? tget(i + lane) : 4;
    const int tb = __shfl_sync(0xffffffffu, tblk, i & 31);
    const int8_t *srow = mat + tb * 5;
    int f = 0, mxv = 0;
    int h = slen > 0 ? H0[(slen - 1) * P + l] : 0;
    h = __shfl_up_sync(0xffffffffu, h, 1);
    if (lane == 0) h = 0;
    for (int j = 0; j < slen; ++j) {
      const int k = j + l * slen;
      int s = k >= qlen ? 0 : srow[q[k]];
      if (IS8) s = (s + shift) & 0xff;
      int e = E[j * P + l];
      int v = h + s;
      if (IS8) { v = v > 255 ? 255 : v; v = dp_subs(v, shift); }
      else { v = v > 32767 ? 32767 : v; v = v < -32768 ? -32768 : v; }
      v = v < e ? e : v;
      v = v < f ? f : v;
      mxv = mxv < v ? v : mxv;
      const int h0 = H0[j * P + l];
      if (on) H1[j * P + l] = v;
      const int ee = dp_subs(e, e_del);
      int t = dp_subs(v, oe_del);
      if (on) E[j * P + l] = ee > t ? ee : t;
      const int ff = dp_subs(f, e_ins);
      t = dp_subs(v, oe_ins);
      f = ff > t ? ff : t;
      h = h0;
    }
    __syncwarp();
    // lazy-F loop (ksw.c:176-190 / :289-301)
    bool done = false;
    for (int k = 0; k < 16 && !done; ++k) {
      f = __shfl_up_sync(0xffffffffu, f, 1);
      if (lane == 0) f = 0;
      for (int j = 0; j < slen; ++j) {
        int v = H1[j * P + l];
        v = v > f ? v : f;
        if (on) H1[j * P + l] = v;
        v = dp_subs(v, oe_ins);
        f = dp_subs(f, e_ins);
        const bool ok = IS8 ? dp_subs(f, v) == 0 : !(f > v);
        if ((__ballot_sync(0xffffffffu, ok) & PM) == PM) { done = true; break; }
      }
    }
    __syncwarp();
    int imax = on ? mxv : 0;
#pragma unroll
    for (int o = 8; o; o >>= 1) { const int v = __shfl_xor_sync(0xffffffffu, imax, o); imax = imax > v ? imax : v; }
    imax = __shfl_sync(0xffffffffu, imax, 0);
    if (imax >= minsc) {
      if (n_b == 0 || (int)(uint32_t)b_last + 1 != i) {
        if (n_b < DP_MAX_TLEN) { b_last = (unsigned long long)imax << 32 | (uint32_t)i; if (lane == 0) bl[n_b] = b_last; ++n_b; }
      } else if ((int)(b_last >> 32) < imax) { b_last = (unsigned long long)imax << 32 | (uint32_t)i; if (lane == 0) bl[n_b - 1] = b_last; }
    }
    if (imax > gmax) {
      gmax = imax; te = i;
      if (on) for (int j = 0; j < slen; ++j) Hmax[j * P + l] = H1[j * P + l];
      if (IS8) { if (gmax + shift >= 255 || gmax >= endsc) break; }
      else if (gmax >= endsc) break;
    }
    { int *t = H1; H1 = H0; H0 = t; }
  }
  __syncwarp();
  res.score = IS8 ? (gmax + shift < 255 ? gmax : 255) : gmax;
  res.te = te;
  if (!IS8 || res.score != 255) {
    // end of the best alignment on the query: largest Hmax, ties to the smallest position (ksw.c:206-214)
    int key = -1;
    if (on) for (int j = 0; j < slen; ++j) { const int k2 = Hmax[j * P + l] << 16 | (0xffff - (j + l * slen)); key = key > k2 ? key : k2; }
#pragma unroll
    for (int o = 8; o; o >>= 1) { const int v = __shfl_xor_sync(0xffffffffu, key, o); key = key > v ? key : v; }
    key = __shfl_sync(0xffffffffu, key, 0);
    res.qe = key < 0 ? -1 : 0xffff - (key & 0xffff);
    if (n_b > 0) {  // second best score away from the best end (ksw.c:215-226): first maximum in list order
      const int d = (res.score + mxm - 1) / mxm, low = te - d, high = te + d;
      long long best = -1;
      for (int k = lane; k < n_b; k += 32) {
        const unsigned long long be = __ldcg(bl + k);
        const int e = (int)(uint32_t)be, sc = (int)(be >> 32);
        if (e < low || e > high) {
          const long long k2 = (long long)sc << 32 | (uint32_t)(0x7fffffff - k);
          best = best > k2 ? best : k2;
        }
      }
#pragma unroll
      for (int o = 16; o; o >>= 1) { const long long v = __shfl_xor_sync(0xffffffffu, best, o); best = best > v ? best : v; }
      if (best >= 0 && (int)(best >> 32) > res.score2) {
        const int idx = 0x7fffffff - (int)(uint32_t)best;
        res.score2 = (int)(best >> 32);
        res.te2 = (int)(uint32_t)__ldcg(bl + idx);
      }
    }
  }
  __syncwarp();
  return res;
}

struct TFwd {
  const bsq_devidx_t *ix; int64_t rb;
  __device__ int operator()(int i) const { return dp_ref_base(*ix, rb + i); }
};
struct TRevPrefix {  // the first te+1 bases reversed, the rest as they are (ksw.c:359: revseq(r.te + 1, target))
  const bsq_devidx_t *ix; int64_t rb; int te;
  __device__ int operator()(int i) const { return dp_ref_base(*ix, rb + (i <= te ? te - i : i)); }
};

__global__ void __launch_bounds__(32 * DP_WARPS) k_matesw(const __grid_constant__ bsq_devopt_t opt, const __grid_constant__ bsq_devidx_t ix, int64_t n_jobs,
                                                           const bsq_matesw_job *jobs, const uint8_t *seqs, int32_t stride, const int32_t *lens,
                                                           bsq_matesw_res *res, unsigned long long *bscratch, DpCtr *ctr) {
  __shared__ int s_a[DP_WARPS][4][BSQ_MAX_READ_LEN + 32];
  __shared__ uint8_t s_q[DP_WARPS][BSQ_MAX_READ_LEN + 8], s_q2[DP_WARPS][BSQ_MAX_READ_LEN + 8];
  __shared__ int8_t s_mat[2][32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (threadIdx.x < 25) { s_mat[0][threadIdx.x] = opt.ctmat[threadIdx.x]; s_mat[1][threadIdx.x] = opt.gamat[threadIdx.x]; }
  __syncthreads();
  unsigned long long *bl = bscratch + ((size_t)blockIdx.x * DP_WARPS + wid) * DP_MAX_TLEN;
  uint8_t *q = s_q[wid], *q2 = s_q2[wid];
  for (;;) {
    long long jid = 0;
    if (lane == 0) jid = (long long)atomicAdd(&ctr->next_job, 1ull);
    jid = __shfl_sync(0xffffffffu, jid, 0);
    if (jid >= n_jobs) break;
    const bsq_matesw_job jb = jobs[jid];
    bsq_matesw_res out;
    out.score = 0; out.te = -1; out.qe = -1; out.score2 = -1; out.te2 = -1; out.tb = -1; out.qb = -1; out.pad_ = 0;
    const int l_ms = lens[jb.row];
    const int64_t tlen64 = jb.re - jb.rb;
    if (l_ms <= 0 || l_ms > BSQ_MAX_READ_LEN || tlen64 <= 0 || tlen64 > DP_MAX_TLEN) {
      out.pad_ = 1;  // outside the kernel's limits: the caller does this one itself
    } else {
      const int tlen = (int)tlen64;
      const uint8_t *rowp = seqs + (size_t)jb.row * stride;
      for (int x = lane; x < l_ms; x += 32) { const uint8_t c = rowp[x]; q[l_ms - 1 - x] = c < 4 ? 3 - c : 4; }  // reverse complement (mem_alnreg.c:414-416)
      __syncwarp();
      const int8_t *mat = s_mat[jb.use_ga ? 1 : 0];
      const bool is8 = (jb.xtra & 0x10000) != 0;
      TFwd t1{&ix, jb.rb};
      SwRes r = is8 ? dp_sw_striped<true>(l_ms, q, tlen, t1, mat, opt.o_del, opt.e_del, opt.o_ins, opt.e_ins, jb.xtra, s_a[wid][0], s_a[wid][1], s_a[wid][2], s_a[wid][3], bl)
                    : dp_sw_striped<false>(l_ms, q, tlen, t1, mat, opt.o_del, opt.e_del, opt.o_ins, opt.e_ins, jb.xtra, s_a[wid][0], s_a[wid][1], s_a[wid][2], s_a[wid][3], bl);
      out.score = r.score; out.te = r.te; out.qe = r.qe; out.score2 = r.score2; out.te2 = r.te2;
      const bool second = (jb.xtra & 0x80000) != 0 && !((jb.xtra & 0x40000) && r.score < (jb.xtra & 0xffff));  // ksw.c:355-356
      if (second) {
        const int ql2 = r.qe + 1;
        SwRes rr; rr.score = 0; rr.te = -1; rr.qe = -1;
        if (ql2 > 0) {
          for (int x = lane; x < ql2; x += 32) q2[x] = q[r.qe - x];
          __syncwarp();
          TRevPrefix t2{&ix, jb.rb, r.te};
          const int x2 = 0x20000 | r.score;
          rr = is8 ? dp_sw_striped<true>(ql2, q2, tlen, t2, mat, opt.o_del, opt.e_del, opt.o_ins, opt.e_ins, x2, s_a[wid][0], s_a[wid][1], s_a[wid][2], s_a[wid][3], bl)
                   : dp_sw_striped<false>(ql2, q2, tlen, t2, mat, opt.o_del, opt.e_del, opt.o_ins, opt.e_ins, x2, s_a[wid][0], s_a[wid][1], s_a[wid][2], s_a[wid][3], bl);
        }
        if (r.score == rr.score) { out.tb = r.te - rr.te; out.qb = r.qe - rr.qe; }
      }
    }
    if (lane == 0) res[jid] = out;
    __syncwarp();
  }
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------------
// host side of the C ABI
// ---------------------------------------------------------------------------------------------------------------
struct bsq_dp {
  const bsq_index *idx;
  bsq_devopt_t opt;
  cudaStream_t stream;
  cudaEvent_t ev[4];
  DpBuf seqs, lens, cjobs, cres, blob, cscratch, mjobs, mres, mscratch, ctr;
  HostBuf h_blob, h_ctr;
  int64_t n_rows = 0;
  int32_t stride = 0;
  int grid = 0;
  // CIGAR submission in flight
  int64_t c_n = 0;
  const bsq_cigar_job *c_jobs = nullptr;
  bsq_cigar_res *c_res = nullptr;
  bool c_pending = false, m_pending = false;
  int64_t m_n = 0;
  int64_t counters[8];
};

static int dp_launch_cigar(bsq_dp *dp) {
  cudaStream_t s = dp->stream;
  DpCtr *ctr = dp->ctr.as<DpCtr>();
  CKD(cudaMemsetAsync(ctr, 0, sizeof(DpCtr), s));
  CKD(cudaEventRecord(dp->ev[0], s));
  k_cigar<<<dp->grid, 32 * DP_WARPS, 0, s>>>(dp->opt, dp->idx->d, dp->c_n, dp->cjobs.as<bsq_cigar_job>(), dp->seqs.as<uint8_t>(), dp->stride,
                                             dp->lens.as<int32_t>(), dp->cres.as<bsq_cigar_res>(), dp->blob.as<uint32_t>(), dp->blob.cap / 4,
                                             dp->cscratch.as<CigScratch>(), ctr);
  CKD(cudaGetLastError());
  CKD(cudaEventRecord(dp->ev[1], s));
  CKD(cudaMemcpyAsync(dp->c_res, dp->cres.p, (size_t)dp->c_n * sizeof(bsq_cigar_res), cudaMemcpyDeviceToHost, s));
  CKD(cudaMemcpyAsync(dp->h_ctr.p, ctr, sizeof(DpCtr), cudaMemcpyDeviceToHost, s));
  return 0;
}

extern "C" {

int bsq_dp_create(const bsq_index *idx, const bsq_opt *opt, bsq_dp **out) {
  if (!idx || !opt || !out) return BSQ_EINVAL;
  CKD(cudaSetDevice(idx->device));
  bsq_dp *dp = new bsq_dp();
  dp->idx = idx;
  memcpy(&dp->opt, opt, sizeof dp->opt);
  memset(dp->counters, 0, sizeof dp->counters);
  int lo = 0, hi = 0;
  cudaDeviceGetStreamPriorityRange(&lo, &hi);
  CKD(cudaStreamCreateWithPriority(&dp->stream, cudaStreamNonBlocking, hi));  // ahead of the phase-1 kernels of the next batch
  for (int i = 0; i < 4; ++i) CKD(cudaEventCreate(&dp->ev[i]));
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, idx->device);
  dp->grid = sms * 4;  // 4 CTAs of 4 warps per SM: one warp per job, jobs pulled from a counter
  int rc;
  if ((rc = dp->ctr.reserve(sizeof(DpCtr))) || (rc = dp->h_ctr.reserve(sizeof(DpCtr)))) { delete dp; return rc; }
  *out = dp;
  return 0;
}

void bsq_dp_destroy(bsq_dp *dp) {
  if (!dp) return;
  cudaSetDevice(dp->idx->device);
  cudaStreamSynchronize(dp->stream);
  DpBuf *bufs[] = {&dp->seqs, &dp->lens, &dp->cjobs, &dp->cres, &dp->blob, &dp->cscratch, &dp->mjobs, &dp->mres, &dp->mscratch, &dp->ctr};
  for (DpBuf *b : bufs) b->release();
  dp->h_blob.release(); dp->h_ctr.release();
  for (int i = 0; i < 4; ++i) cudaEventDestroy(dp->ev[i]);
  cudaStreamDestroy(dp->stream);
  delete dp;
}

int bsq_dp_set_reads(bsq_dp *dp, int64_t n_rows, const uint8_t *seqs, int32_t stride, const int32_t *lens) {
  if (!dp || n_rows < 0 || stride <= 0) return BSQ_EINVAL;
  dp->n_rows = 0;
  if (n_rows == 0) return 0;
  CKD(cudaSetDevice(dp->idx->device));
  int rc;
  if ((rc = dp->seqs.reserve((size_t)n_rows * stride)) || (rc = dp->lens.reserve((size_t)n_rows * 4))) return rc;
  CKD(cudaMemcpyAsync(dp->seqs.p, seqs, (size_t)n_rows * stride, cudaMemcpyHostToDevice, dp->stream));
  CKD(cudaMemcpyAsync(dp->lens.p, lens, (size_t)n_rows * 4, cudaMemcpyHostToDevice, dp->stream));
  dp->n_rows = n_rows; dp->stride = stride;
  return 0;
}

int bsq_dp_sync(bsq_dp *dp) {
  if (!dp) return BSQ_EINVAL;
  CKD(cudaSetDevice(dp->idx->device));
  CKD(bsq_stream_wait(dp->stream));
  return 0;
}

int bsq_dp_cigar_submit(bsq_dp *dp, int64_t n_jobs, const bsq_cigar_job *jobs, bsq_cigar_res *res) {
  if (!dp || n_jobs < 0 || dp->c_pending) return BSQ_EINVAL;
  dp->c_n = n_jobs; dp->c_jobs = jobs; dp->c_res = res;
  dp->counters[0] = n_jobs;
  if (n_jobs == 0) { dp->c_pending = true; return 0; }
  if (!jobs || !res || dp->n_rows == 0) return BSQ_EINVAL;
  for (int64_t j = 0; j < n_jobs; ++j)
    if (jobs[j].row < 0 || jobs[j].row >= dp->n_rows) return BSQ_EINVAL;
  CKD(cudaSetDevice(dp->idx->device));
  int rc;
  if ((rc = dp->cjobs.reserve((size_t)n_jobs * sizeof(bsq_cigar_job))) || (rc = dp->cres.reserve((size_t)n_jobs * sizeof(bsq_cigar_res))) ||
      (rc = dp->blob.reserve((size_t)n_jobs * 48 + (1 << 20))) || (rc = dp->cscratch.reserve((size_t)dp->grid * DP_WARPS * sizeof(CigScratch))))
    return rc;
  CKD(cudaMemcpyAsync(dp->cjobs.p, jobs, (size_t)n_jobs * sizeof(bsq_cigar_job), cudaMemcpyHostToDevice, dp->stream));
  if ((rc = dp_launch_cigar(dp))) return rc;
  dp->c_pending = true;
  return 0;
}

int bsq_dp_cigar_wait(bsq_dp *dp, const uint32_t **blob, int64_t *blob_words) {
  if (!dp || !dp->c_pending) return BSQ_EINVAL;
  dp->c_pending = false;
  if (blob) *blob = nullptr;
  if (blob_words) *blob_words = 0;
  if (dp->c_n == 0) return 0;
  CKD(cudaSetDevice(dp->idx->device));
  CKD(bsq_stream_wait(dp->stream));
  const DpCtr *hc = (const DpCtr *)dp->h_ctr.p;
  if (hc->blob_used > dp->blob.cap / 4) {  // the CIGAR/MD text did not fit: grow the blob to what was asked for and run again
    int rc = dp->blob.reserve((size_t)hc->blob_used * 4 + 4096);
    if (rc) return rc;
    if ((rc = dp_launch_cigar(dp))) return rc;
    CKD(bsq_stream_wait(dp->stream));
    dp->counters[6]++;
    if (hc->blob_used > dp->blob.cap / 4) { bsq_set_error("k_cigar: blob overflow after growing"); return BSQ_EOVERFLOW; }
  }
  float ms = 0;
  cudaEventElapsedTime(&ms, dp->ev[0], dp->ev[1]);
  dp->counters[1] = (int64_t)hc->ungapped; dp->counters[2] = (int64_t)hc->cells; dp->counters[3] = (int64_t)(ms * 1000);
  const size_t bytes = (size_t)hc->blob_used * 4;
  int rc = dp->h_blob.reserve(bytes + 16);
  if (rc) return rc;
  if (bytes) {
    CKD(cudaMemcpyAsync(dp->h_blob.p, dp->blob.p, bytes, cudaMemcpyDeviceToHost, dp->stream));
    CKD(bsq_stream_wait(dp->stream));
  }
  if (blob) *blob = (const uint32_t *)dp->h_blob.p;
  if (blob_words) *blob_words = (int64_t)hc->blob_used;
  return 0;
}

int bsq_dp_matesw_submit(bsq_dp *dp, int64_t n_jobs, const bsq_matesw_job *jobs, bsq_matesw_res *res) {
  if (!dp || n_jobs < 0 || dp->m_pending) return BSQ_EINVAL;
  dp->m_n = n_jobs;
  dp->counters[4] = n_jobs;
  if (n_jobs == 0) { dp->m_pending = true; return 0; }
  if (!jobs || !res || dp->n_rows == 0) return BSQ_EINVAL;
  for (int64_t j = 0; j < n_jobs; ++j)
    if (jobs[j].row < 0 || jobs[j].row >= dp->n_rows) return BSQ_EINVAL;
  CKD(cudaSetDevice(dp->idx->device));
  int rc;
  if ((rc = dp->mjobs.reserve((size_t)n_jobs * sizeof(bsq_matesw_job))) || (rc = dp->mres.reserve((size_t)n_jobs * sizeof(bsq_matesw_res))) ||
      (rc = dp->mscratch.reserve((size_t)dp->grid * DP_WARPS * DP_MAX_TLEN * 8)))
    return rc;
  cudaStream_t s = dp->stream;
  CKD(cudaMemcpyAsync(dp->mjobs.p, jobs, (size_t)n_jobs * sizeof(bsq_matesw_job), cudaMemcpyHostToDevice, s));
  DpCtr *ctr = dp->ctr.as<DpCtr>();
  CKD(cudaMemsetAsync(ctr, 0, sizeof(DpCtr), s));
  CKD(cudaEventRecord(dp->ev[2], s));
  k_matesw<<<dp->grid, 32 * DP_WARPS, 0, s>>>(dp->opt, dp->idx->d, n_jobs, dp->mjobs.as<bsq_matesw_job>(), dp->seqs.as<uint8_t>(), dp->stride,
                                              dp->lens.as<int32_t>(), dp->mres.as<bsq_matesw_res>(), dp->mscratch.as<unsigned long long>(), ctr);
  CKD(cudaGetLastError());
  CKD(cudaEventRecord(dp->ev[3], s));
  CKD(cudaMemcpyAsync(res, dp->mres.p, (size_t)n_jobs * sizeof(bsq_matesw_res), cudaMemcpyDeviceToHost, s));
  dp->m_pending = true;
  return 0;
}

int bsq_dp_matesw_wait(bsq_dp *dp) {
  if (!dp || !dp->m_pending) return BSQ_EINVAL;
  dp->m_pending = false;
  if (dp->m_n == 0) return 0;
  CKD(cudaSetDevice(dp->idx->device));
  CKD(bsq_stream_wait(dp->stream));
  float ms = 0;
  cudaEventElapsedTime(&ms, dp->ev[2], dp->ev[3]);
  dp->counters[5] = (int64_t)(ms * 1000);
  return 0;
}

int bsq_dp_counters(const bsq_dp *dp, int64_t *c, int n) {
  if (!dp || !c) return BSQ_EINVAL;
  for (int i = 0; i < n && i < 8; ++i) c[i] = dp->counters[i];
  return 0;
}

}  // extern "C"
