// GPU construction of the two bisulfite FM-indices (`biscuit index`).
//
// Output is bit-identical to what the reference writes (lib/aln/bwtindex.c:206-347):
//   text  = conv(fwd) || conv(revcomp(fwd)), conv = C>T (parent) or G>A (daughter)   bntseq.c:588-600
//   .bwt  = BWT of text$ without the '$', interleaved with occ checkpoints every 128   bwtindex.c:130-154
//   .sa   = SA[32*i], i >= 1                                                          bwt.c:63-85
// but it is built differently: the reference grows the BWT incrementally on one CPU core (bwt_gen.c,
// hours for a 3-Gb genome); here suffixes are bucketed by their first 12 symbols, each group of buckets
// is radix-sorted on 31-symbol (62-bit) keys, the few remaining ties are refined 31 symbols at a time,
// and BWT symbols / SA samples are scattered straight from the sorted order.  cub::DeviceRadixSort /
// DeviceScan / DeviceSelect are used as library plumbing; this is index construction, not the
// alignment hot path.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>
#include <cub/iterator/counting_input_iterator.cuh>

#include "../../include/bsq.h"
#include "bsq_common.h"
#include "bsq_internal.h"

#define PFX_SYMS 12
#define PFX_BUCKETS (1u << (2 * PFX_SYMS))

#define CKB(call)                                                                                  \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) {                                                                       \
      bsq_set_error("%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_));            \
      rc = e_ == cudaErrorMemoryAllocation ? BSQ_ENOMEM : BSQ_ENODEV;                              \
      goto done;                                                                                   \
    }                                                                                              \
  } while (0)

// ---- text access: T packed 16 symbols per u32, first symbol in the top bits; 2 spare words of zeros at the end ----

__device__ __forceinline__ int txt_sym(const uint32_t *T, uint64_t i) { return (T[i >> 4] >> ((~i & 15) << 1)) & 3; }

// Sort key of the suffix starting at i, 31 symbols at a time.  Bit 63 = 1 while the suffix still has
// symbols (i < n), followed by 31 symbols (bits 62..1, zero padded past the end of the text).  A suffix
// that is exhausted (i >= n) sorts before every live one ('$' is the smallest symbol) and, among
// exhausted suffixes of one tie group, the shorter one first: key = 64 - (i - n).
__device__ __forceinline__ uint64_t txt_key31(const uint32_t *T, uint64_t n, uint64_t i) {
  if (i >= n) { uint64_t e = i - n; return 64 - (e < 63 ? e : 63); }
  const uint64_t w = i >> 4;
  const int sh = (int)(i & 15) << 1;  // bits to drop from the first word
  uint64_t hi = ((uint64_t)T[w] << 32) | T[w + 1];
  uint64_t lo = (uint64_t)T[w + 2] << 32;
  uint64_t v = sh ? (hi << sh) | (lo >> (64 - sh)) : hi;  // 32 symbols starting at i (words past the end are zero)
  uint64_t rem = n - i;                                    // real symbols available
  if (rem < 31) v &= ~((1ull << ((32 - rem) << 1)) - 1);
  return (1ull << 63) | ((v >> 2) << 1);
}

__device__ __forceinline__ uint32_t txt_prefix(const uint32_t *T, uint64_t n, uint64_t i) {
  return (uint32_t)(txt_key31(T, n, i) >> (63 - 2 * PFX_SYMS)) & (PFX_BUCKETS - 1);
}

// ---- kernels ----

// converted, doubled text from the forward 2-bit pac (one thread per output word)
__global__ void k_make_text(const uint8_t *pac, int64_t l_pac, int parent, uint32_t *T, uint64_t n_words, unsigned long long *sym_cnt) {
  uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long c[4] = {0, 0, 0, 0};
  if (w < n_words) {
    uint32_t v = 0;
    const uint64_t n = (uint64_t)l_pac * 2;
    for (int j = 0; j < 16; ++j) {
      uint64_t i = w * 16 + j;
      int s = 0;
      if (i < n) {
        if (i < (uint64_t)l_pac) s = (pac[i >> 2] >> ((~i & 3) << 1)) & 3;
        else { uint64_t k = n - 1 - i; s = 3 - ((pac[k >> 2] >> ((~k & 3) << 1)) & 3); }
        if (parent) { if (s == 1) s = 3; } else { if (s == 2) s = 0; }
        ++c[s];
      }
      v |= (uint32_t)s << ((15 - j) << 1);
    }
    T[w] = v;
  }
  // block-level reduction of the symbol counts
  __shared__ unsigned long long sh[4];
  if (threadIdx.x < 4) sh[threadIdx.x] = 0;
  __syncthreads();
  for (int s = 0; s < 4; ++s) {
    unsigned long long v = c[s];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(&sh[s], v);
  }
  __syncthreads();
  if (threadIdx.x < 4 && sh[threadIdx.x]) atomicAdd(&sym_cnt[threadIdx.x], sh[threadIdx.x]);
}

__global__ void k_prefix_hist(const uint32_t *T, uint64_t n, unsigned int *hist) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) atomicAdd(&hist[txt_prefix(T, n, i)], 1u);
}

struct InBucketRange {
  const uint32_t *T;
  uint64_t n;
  uint32_t lo, hi;
  __device__ bool operator()(uint64_t i) const {
    uint32_t p = txt_prefix(T, n, i);
    return p >= lo && p < hi;
  }
};

__global__ void k_make_keys(const uint32_t *T, uint64_t n, const uint64_t *pos, uint64_t m, uint64_t depth, uint64_t *keys) {
  uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j < m) keys[j] = txt_key31(T, n, pos[j] + depth);
}

// flag[j] = 1 when element j ties with a neighbour on (grp, key)
__global__ void k_mark_ties(const uint64_t *keys, const uint64_t *grp, uint64_t m, uint8_t *flag) {
  uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= m) return;
  bool t = false;
  if (j > 0) t |= keys[j] == keys[j - 1] && (!grp || grp[j] == grp[j - 1]);
  if (j + 1 < m) t |= keys[j] == keys[j + 1] && (!grp || grp[j] == grp[j + 1]);
  flag[j] = t;
}

// new group id = index of the first element of the (grp,key) run the element belongs to
__global__ void k_group_heads(const uint64_t *keys, const uint64_t *grp, uint64_t m, uint64_t *head) {
  uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= m) return;
  bool is_head = j == 0 || keys[j] != keys[j - 1] || (grp && grp[j] != grp[j - 1]);
  head[j] = is_head ? j : 0;
}

struct MaxOp { __device__ uint64_t operator()(uint64_t a, uint64_t b) const { return a > b ? a : b; } };

__global__ void k_gather_u64(const uint64_t *src, const uint64_t *idx, uint64_t m, uint64_t *dst) {
  uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j < m) dst[j] = src[idx[j]];
}
__global__ void k_scatter_u64(const uint64_t *src, const uint64_t *idx, uint64_t m, uint64_t *dst) {
  uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j < m) dst[idx[j]] = src[j];
}

// sorted suffix positions of one chunk -> BWT symbol bytes, SA samples, primary.
// rank of element j = rank0 + j (rank 0 is the '$' suffix).
__global__ void k_emit(const uint32_t *T, const uint64_t *pos, uint64_t m, uint64_t rank0, uint8_t *bwt_sym, uint64_t *sa, int sa_intv,
                       unsigned long long *primary, uint64_t *full_sa) {
  uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= m) return;
  const uint64_t r = rank0 + j, p = pos[j];
  if (p == 0) { *primary = r; bwt_sym[r] = 0; }
  else bwt_sym[r] = (uint8_t)txt_sym(T, p - 1);
  if (r % sa_intv == 0) sa[r / sa_intv] = p;
  if (full_sa) full_sa[r] = p;
}

// per 128-symbol block of the '$'-less BWT string: symbol counts (packed 4 x 16 bit -> u64)
__global__ void k_block_counts(const uint8_t *bwt_sym, uint64_t n, uint64_t primary, uint64_t n_blocks, uint64_t *cnt4 /*4 arrays of n_blocks+1*/) {
  uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n_blocks) return;
  uint32_t c[4] = {0, 0, 0, 0};
  for (int j = 0; j < 128; ++j) {
    uint64_t k = b * 128 + j;  // index in the '$'-less string
    if (k >= n) break;
    uint64_t r = k + (k >= primary);
    ++c[bwt_sym[r]];
  }
  for (int s = 0; s < 4; ++s) cnt4[(uint64_t)s * (n_blocks + 1) + b] = c[s];
}

// interleaved layout: block b = {u64 occ[4] (counts before the block), 8 x u32 symbols}; trailing occ[4]
__global__ void k_write_blocks(const uint8_t *bwt_sym, uint64_t n, uint64_t primary, uint64_t n_blocks, const uint64_t *occ4, uint32_t *out,
                               uint64_t out_words) {
  uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b > n_blocks) return;
  // the trailing occ[4] sits right after the last symbol word (the last block may be partial)
  uint32_t *o = b < n_blocks ? out + b * 16 : out + (out_words - 8);
  for (int s = 0; s < 4; ++s) {
    uint64_t v = occ4[(uint64_t)s * (n_blocks + 1) + b];
    o[2 * s] = (uint32_t)v; o[2 * s + 1] = (uint32_t)(v >> 32);
  }
  if (b == n_blocks) return;
  for (int wd = 0; wd < 8; ++wd) {
    uint64_t k0 = b * 128 + (uint64_t)wd * 16;
    if (k0 >= n) break;
    uint32_t v = 0;
    for (int j = 0; j < 16; ++j) {
      uint64_t k = k0 + j;
      if (k >= n) break;
      uint64_t r = k + (k >= primary);
      v |= (uint32_t)bwt_sym[r] << ((15 - j) << 1);
    }
    if (b * 16 + 8 + wd < out_words) o[8 + wd] = v;
  }
}

static inline unsigned nb(uint64_t n, int bs) { return (unsigned)((n + bs - 1) / bs); }

struct Scratch {
  void *p = nullptr; size_t cap = 0;
  cudaError_t need(size_t b) {
    if (b <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    cudaError_t e = cudaMalloc(&p, b + (b >> 3) + 256);
    if (e == cudaSuccess) cap = b + (b >> 3) + 256;
    return e;
  }
  ~Scratch() { if (p) cudaFree(p); }
};

// Refine ties inside pos[0..m) (already sorted by the depth-0 key in keys[]) until the order is total.
static int refine_ties(const uint32_t *T, uint64_t n, uint64_t *pos, uint64_t *keys, uint64_t m, Scratch &tmp, int *n_pass) {
  int rc = 0;
  uint8_t *flag = nullptr;
  uint64_t *slots = nullptr, *u_pos = nullptr, *u_key = nullptr, *u_grp = nullptr, *a1 = nullptr, *a2 = nullptr, *a3 = nullptr, *d_cnt = nullptr;
  uint64_t n_u = 0, depth = 0;
  size_t tb = 0;
  CKB(cudaMalloc(&flag, m + 1));
  CKB(cudaMalloc(&d_cnt, 8));
  k_mark_ties<<<nb(m, 256), 256>>>(keys, nullptr, m, flag);
  CKB(cudaGetLastError());
  // unresolved slots = indices j with flag[j]
  {
    cub::CountingInputIterator<uint64_t> it(0);
    CKB(cudaMalloc(&slots, 8 * (m + 1)));  // upper bound; shrinks logically
    CKB(cub::DeviceSelect::Flagged(nullptr, tb, it, flag, slots, d_cnt, (int64_t)m));
    CKB(tmp.need(tb));
    CKB(cub::DeviceSelect::Flagged(tmp.p, tb, it, flag, slots, d_cnt, (int64_t)m));
    CKB(cudaMemcpy(&n_u, d_cnt, 8, cudaMemcpyDeviceToHost));
  }
  *n_pass = 0;
  if (n_u == 0) goto done;
  CKB(cudaMalloc(&u_pos, 8 * n_u)); CKB(cudaMalloc(&u_key, 8 * n_u)); CKB(cudaMalloc(&u_grp, 8 * n_u));
  CKB(cudaMalloc(&a1, 8 * n_u)); CKB(cudaMalloc(&a2, 8 * n_u)); CKB(cudaMalloc(&a3, 8 * n_u));
  k_gather_u64<<<nb(n_u, 256), 256>>>(pos, slots, n_u, u_pos);
  k_gather_u64<<<nb(n_u, 256), 256>>>(keys, slots, n_u, u_key);
  // initial groups: runs of equal depth-0 keys
  k_group_heads<<<nb(n_u, 256), 256>>>(u_key, nullptr, n_u, a1);
  CKB(cudaGetLastError());
  CKB(cub::DeviceScan::InclusiveScan(nullptr, tb, a1, u_grp, MaxOp(), (int64_t)n_u));
  CKB(tmp.need(tb));
  CKB(cub::DeviceScan::InclusiveScan(tmp.p, tb, a1, u_grp, MaxOp(), (int64_t)n_u));
  while (n_u > 0) {
    ++*n_pass;
    depth += 31;
    // sort the unresolved set by (group, next 31 symbols): stable sort by key, then by group
    k_make_keys<<<nb(n_u, 256), 256>>>(T, n, u_pos, n_u, depth, u_key);
    CKB(cudaGetLastError());
    // (key, pos) -> (a1, a2); carry group along with a second pass keyed the same way
    CKB(cub::DeviceRadixSort::SortPairs(nullptr, tb, u_key, a1, u_pos, a2, (int64_t)n_u));
    CKB(tmp.need(tb));
    CKB(cub::DeviceRadixSort::SortPairs(tmp.p, tb, u_key, a1, u_pos, a2, (int64_t)n_u));
    CKB(cub::DeviceRadixSort::SortPairs(tmp.p, tb, u_key, a1, u_grp, a3, (int64_t)n_u));
    // now a1 = keys sorted, a2 = pos, a3 = grp (same permutation).  Stable sort by group.
    CKB(cub::DeviceRadixSort::SortPairs(nullptr, tb, a3, u_grp, a2, u_pos, (int64_t)n_u));
    CKB(tmp.need(tb));
    CKB(cub::DeviceRadixSort::SortPairs(tmp.p, tb, a3, u_grp, a2, u_pos, (int64_t)n_u));
    CKB(cub::DeviceRadixSort::SortPairs(tmp.p, tb, a3, u_grp, a1, u_key, (int64_t)n_u));
    // u_grp, u_pos, u_key are ordered by (group, key); slot order is unchanged -> write back
    k_scatter_u64<<<nb(n_u, 256), 256>>>(u_pos, slots, n_u, pos);
    CKB(cudaGetLastError());
    // still tied?
    k_mark_ties<<<nb(n_u, 256), 256>>>(u_key, u_grp, n_u, flag);
    k_group_heads<<<nb(n_u, 256), 256>>>(u_key, u_grp, n_u, a1);
    CKB(cudaGetLastError());
    CKB(cub::DeviceScan::InclusiveScan(nullptr, tb, a1, a3, MaxOp(), (int64_t)n_u));
    CKB(tmp.need(tb));
    CKB(cub::DeviceScan::InclusiveScan(tmp.p, tb, a1, a3, MaxOp(), (int64_t)n_u));  // a3 = new group ids
    // compact slots / pos / group by flag
    uint64_t n_next = 0;
    CKB(cub::DeviceSelect::Flagged(nullptr, tb, slots, flag, a1, d_cnt, (int64_t)n_u));
    CKB(tmp.need(tb));
    CKB(cub::DeviceSelect::Flagged(tmp.p, tb, slots, flag, a1, d_cnt, (int64_t)n_u));
    CKB(cudaMemcpy(&n_next, d_cnt, 8, cudaMemcpyDeviceToHost));
    if (n_next) {
      CKB(cudaMemcpy(slots, a1, 8 * n_next, cudaMemcpyDeviceToDevice));
      CKB(cub::DeviceSelect::Flagged(tmp.p, tb, u_pos, flag, a1, d_cnt, (int64_t)n_u));
      CKB(cudaMemcpy(u_pos, a1, 8 * n_next, cudaMemcpyDeviceToDevice));
      CKB(cub::DeviceSelect::Flagged(tmp.p, tb, a3, flag, a1, d_cnt, (int64_t)n_u));
      CKB(cudaMemcpy(u_grp, a1, 8 * n_next, cudaMemcpyDeviceToDevice));
    }
    n_u = n_next;
    if (*n_pass > 100000) { bsq_set_error("suffix refinement did not converge (highly repetitive text)"); rc = BSQ_EOVERFLOW; goto done; }
  }
done:
  cudaFree(flag); cudaFree(slots); cudaFree(u_pos); cudaFree(u_key); cudaFree(u_grp); cudaFree(a1); cudaFree(a2); cudaFree(a3); cudaFree(d_cnt);
  return rc;
}

// Build one FM-index half on the current device.  d_pac: forward pac on the device.
static int build_half(const uint8_t *d_pac, int64_t l_pac, int parent, int sa_intv, uint64_t chunk_max, bsq_fm_t *fm, void **alloc_bwt, void **alloc_sa,
                      uint64_t *bwt_words_out, uint64_t *n_sa_out, int64_t *stats, uint64_t *full_sa) {
  int rc = 0;
  const uint64_t n = (uint64_t)l_pac * 2;
  const uint64_t n_words = (n + 15) / 16;
  const uint64_t n_blocks = (n + 127) / 128;
  const uint64_t out_words = ((n + 15) >> 4) + (n_blocks + 1) * 8;
  const uint64_t n_sa = (n + sa_intv) / sa_intv;
  uint32_t *T = nullptr, *blocks = nullptr;
  unsigned int *hist = nullptr;
  unsigned long long *d_small = nullptr;  // [0..3] symbol counts, [4] primary
  uint8_t *bwt_sym = nullptr;
  uint64_t *sa = nullptr, *pos = nullptr, *pos2 = nullptr, *keys = nullptr, *keys2 = nullptr, *cnt4 = nullptr, *d_sel = nullptr;
  unsigned int *h_hist = nullptr;
  unsigned long long h_small[5];
  Scratch tmp;
  size_t tb = 0;
  uint64_t rank0 = 1, max_chunk = 0;
  int tot_pass = 0, n_chunks = 0;

  CKB(cudaMalloc(&T, (n_words + 4) * 4));
  CKB(cudaMemset(T, 0, (n_words + 4) * 4));
  CKB(cudaMalloc(&d_small, 5 * 8));
  CKB(cudaMemset(d_small, 0, 5 * 8));
  k_make_text<<<nb(n_words, 256), 256>>>(d_pac, l_pac, parent, T, n_words, d_small);
  CKB(cudaGetLastError());
  CKB(cudaMalloc(&hist, PFX_BUCKETS * 4));
  CKB(cudaMemset(hist, 0, PFX_BUCKETS * 4));
  k_prefix_hist<<<nb(n, 256), 256>>>(T, n, hist);
  CKB(cudaGetLastError());
  h_hist = (unsigned int *)malloc(PFX_BUCKETS * 4);
  CKB(cudaMemcpy(h_hist, hist, PFX_BUCKETS * 4, cudaMemcpyDeviceToHost));
  CKB(cudaMalloc(&bwt_sym, n + 1));
  CKB(cudaMalloc(&sa, n_sa * 8));
  CKB(cudaMalloc(&d_sel, 8));
  // rank 0 = '$' suffix: BWT symbol T[n-1]; sa[0] = -1 by convention (bwt.c:84)
  {
    // bucket groups of at most chunk_max suffixes
    for (uint32_t b = 0; b < PFX_BUCKETS;) {
      uint64_t m = 0; uint32_t e = b;
      while (e < PFX_BUCKETS && (m == 0 || m + h_hist[e] <= chunk_max)) { m += h_hist[e]; ++e; }
      if (m > max_chunk) max_chunk = m;
      b = e;
    }
    CKB(cudaMalloc(&pos, 8 * (max_chunk + 1))); CKB(cudaMalloc(&pos2, 8 * (max_chunk + 1)));
    CKB(cudaMalloc(&keys, 8 * (max_chunk + 1))); CKB(cudaMalloc(&keys2, 8 * (max_chunk + 1)));
    for (uint32_t b = 0; b < PFX_BUCKETS;) {
      uint64_t m = 0; uint32_t e = b;
      while (e < PFX_BUCKETS && (m == 0 || m + h_hist[e] <= chunk_max)) { m += h_hist[e]; ++e; }
      if (m > 0) {
        ++n_chunks;
        InBucketRange pred{T, n, b, e};
        cub::CountingInputIterator<uint64_t> it(0);
        if (b == 0 && e == PFX_BUCKETS) {
          // single chunk: every suffix; skip the select
          CKB(cub::DeviceSelect::If(nullptr, tb, it, pos, d_sel, (int64_t)n, pred));
          CKB(tmp.need(tb));
          CKB(cub::DeviceSelect::If(tmp.p, tb, it, pos, d_sel, (int64_t)n, pred));
        } else {
          CKB(cub::DeviceSelect::If(nullptr, tb, it, pos, d_sel, (int64_t)n, pred));
          CKB(tmp.need(tb));
          CKB(cub::DeviceSelect::If(tmp.p, tb, it, pos, d_sel, (int64_t)n, pred));
        }
        uint64_t got = 0;
        CKB(cudaMemcpy(&got, d_sel, 8, cudaMemcpyDeviceToHost));
        if (got != m) { bsq_set_error("index build: bucket count mismatch %llu vs %llu", (unsigned long long)got, (unsigned long long)m); rc = BSQ_ENODEV; goto done; }
        k_make_keys<<<nb(m, 256), 256>>>(T, n, pos, m, 0, keys);
        CKB(cudaGetLastError());
        CKB(cub::DeviceRadixSort::SortPairs(nullptr, tb, keys, keys2, pos, pos2, (int64_t)m));
        CKB(tmp.need(tb));
        CKB(cub::DeviceRadixSort::SortPairs(tmp.p, tb, keys, keys2, pos, pos2, (int64_t)m));
        int np = 0;
        if ((rc = refine_ties(T, n, pos2, keys2, m, tmp, &np))) goto done;
        tot_pass += np;
        k_emit<<<nb(m, 256), 256>>>(T, pos2, m, rank0, bwt_sym, sa, sa_intv, d_small + 4, full_sa);
        CKB(cudaGetLastError());
        rank0 += m;
      }
      b = e;
    }
  }
  if (rank0 != n + 1) { bsq_set_error("index build: ranked %llu of %llu suffixes", (unsigned long long)rank0, (unsigned long long)(n + 1)); rc = BSQ_ENODEV; goto done; }
  {
    // '$' suffix (rank 0): preceded by the last text symbol
    uint32_t last_word;
    CKB(cudaMemcpy(&last_word, T + ((n - 1) >> 4), 4, cudaMemcpyDeviceToHost));
    uint8_t s = (uint8_t)((last_word >> ((~(n - 1) & 15) << 1)) & 3);
    CKB(cudaMemcpy(bwt_sym, &s, 1, cudaMemcpyHostToDevice));
    uint64_t m1 = ~0ull;
    CKB(cudaMemcpy(sa, &m1, 8, cudaMemcpyHostToDevice));
    if (full_sa) CKB(cudaMemcpy(full_sa, &m1, 8, cudaMemcpyHostToDevice));
  }
  CKB(cudaMemcpy(h_small, d_small, 5 * 8, cudaMemcpyDeviceToHost));
  cudaFree(pos); pos = nullptr; cudaFree(pos2); pos2 = nullptr; cudaFree(keys); keys = nullptr; cudaFree(keys2); keys2 = nullptr;
  cudaFree(hist); hist = nullptr; cudaFree(T); T = nullptr;
  // occ checkpoints + interleaved layout
  CKB(cudaMalloc(&cnt4, 4 * (n_blocks + 1) * 8));
  CKB(cudaMemset(cnt4, 0, 4 * (n_blocks + 1) * 8));
  k_block_counts<<<nb(n_blocks, 128), 128>>>(bwt_sym, n, h_small[4], n_blocks, cnt4);
  CKB(cudaGetLastError());
  for (int s = 0; s < 4; ++s) {
    uint64_t *p = cnt4 + (uint64_t)s * (n_blocks + 1);
    CKB(cub::DeviceScan::ExclusiveSum(nullptr, tb, p, p, (int64_t)(n_blocks + 1)));
    CKB(tmp.need(tb));
    CKB(cub::DeviceScan::ExclusiveSum(tmp.p, tb, p, p, (int64_t)(n_blocks + 1)));
  }
  CKB(cudaMalloc(&blocks, (out_words + 16) * 4));
  CKB(cudaMemset(blocks, 0, (out_words + 16) * 4));
  k_write_blocks<<<nb(n_blocks + 1, 128), 128>>>(bwt_sym, n, h_small[4], n_blocks, cnt4, blocks, out_words);
  CKB(cudaGetLastError());
  CKB(cudaDeviceSynchronize());
  fm->blocks = blocks; fm->sa = sa; fm->full_sa = full_sa; fm->primary = h_small[4]; fm->seq_len = n; fm->sa_intv = sa_intv;
  fm->L2[0] = 0;
  for (int s = 0; s < 4; ++s) fm->L2[s + 1] = fm->L2[s] + h_small[s];
  *alloc_bwt = blocks; *alloc_sa = sa; *bwt_words_out = out_words; *n_sa_out = n_sa;
  blocks = nullptr; sa = nullptr;
  if (stats) { stats[0] += n_chunks; stats[1] += tot_pass; stats[2] = (int64_t)max_chunk; }
done:
  free(h_hist);
  cudaFree(T); cudaFree(hist); cudaFree(d_small); cudaFree(bwt_sym); cudaFree(sa); cudaFree(pos); cudaFree(pos2); cudaFree(keys); cudaFree(keys2);
  cudaFree(cnt4); cudaFree(d_sel); cudaFree(blocks);
  return rc;
}

extern "C" int bsq_index_build(const uint8_t *pac, int64_t l_pac, int32_t n_seqs, const int64_t *ann_offset, const int32_t *ann_len,
                               const int32_t *ann_is_alt, int device, bsq_index **out) {
  if (!pac || l_pac <= 0 || n_seqs <= 0 || !out) return BSQ_EINVAL;
  int rc = 0, ndev = 0;
  bsq_index *ix = nullptr;
  uint8_t *d_pac = nullptr;
  int64_t stats[4] = {0, 0, 0, 0};
  CKB(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) { bsq_set_error("device %d of %d", device, ndev); return BSQ_ENODEV; }
  CKB(cudaSetDevice(device));
  ix = bsq_index_alloc(device);
  CKB(cudaMalloc(&d_pac, (size_t)(l_pac / 4 + 1)));
  CKB(cudaMemcpy(d_pac, pac, (size_t)((l_pac + 3) / 4), cudaMemcpyHostToDevice));
  bsq_index_adopt(ix, d_pac);
  ix->d.pac = d_pac; d_pac = nullptr;
  ix->d.l_pac = l_pac; ix->d.n_seqs = n_seqs;
  {
    void *p;
    CKB(cudaMalloc(&p, (size_t)n_seqs * 8)); bsq_index_adopt(ix, p);
    CKB(cudaMemcpy(p, ann_offset, (size_t)n_seqs * 8, cudaMemcpyHostToDevice)); ix->d.ann_offset = (const int64_t *)p;
    CKB(cudaMalloc(&p, (size_t)n_seqs * 4)); bsq_index_adopt(ix, p);
    CKB(cudaMemcpy(p, ann_len, (size_t)n_seqs * 4, cudaMemcpyHostToDevice)); ix->d.ann_len = (const int32_t *)p;
    CKB(cudaMalloc(&p, (size_t)n_seqs * 4)); bsq_index_adopt(ix, p);
    if (ann_is_alt) CKB(cudaMemcpy(p, ann_is_alt, (size_t)n_seqs * 4, cudaMemcpyHostToDevice));
    else CKB(cudaMemset(p, 0, (size_t)n_seqs * 4));
    ix->d.ann_is_alt = (const int32_t *)p;
  }
  for (int parent = 1; parent >= 0; --parent) {
    void *a = nullptr, *b = nullptr;
    uint64_t chunk_max = 1ull << 29;  // suffixes sorted per pass (4 x 8 bytes each); BSQ_INDEX_CHUNK overrides (tests)
    if (const char *e = getenv("BSQ_INDEX_CHUNK")) { long long v = atoll(e); if (v > 0) chunk_max = (uint64_t)v; }
    uint64_t *full = nullptr;
    if (bsq_want_full_sa((uint64_t)l_pac * 2, parent == 1 ? 2 : 1)) {
      if (cudaMalloc(&full, ((uint64_t)l_pac * 2 + 1) * 8) != cudaSuccess) { full = nullptr; cudaGetLastError(); }
    }
    rc = build_half(ix->d.pac, l_pac, parent, 32, chunk_max, &ix->d.fm[parent], &a, &b, &ix->bwt_words[parent], &ix->n_sa[parent], stats, full);
    if (rc) { cudaFree(full); goto done; }
    bsq_index_adopt(ix, a); bsq_index_adopt(ix, b); bsq_index_adopt(ix, full);
  }
  ix->build_stats[0] = stats[0]; ix->build_stats[1] = stats[1]; ix->build_stats[2] = stats[2];
  if ((rc = bsq_index_derive_b32(ix))) goto done;
  *out = ix; ix = nullptr;
done:
  cudaFree(d_pac);
  if (ix) bsq_index_free(ix);
  return rc;
}

extern "C" int bsq_index_sizes(const bsq_index *ix, uint64_t *bwt_words, uint64_t *n_sa, uint64_t *primary, uint64_t *L2, int64_t *stats) {
  if (!ix) return BSQ_EINVAL;
  for (int w = 0; w < 2; ++w) {
    if (bwt_words) bwt_words[w] = ix->bwt_words[w];
    if (n_sa) n_sa[w] = ix->n_sa[w];
    if (primary) primary[w] = ix->d.fm[w].primary;
    if (L2) for (int i = 0; i < 5; ++i) L2[5 * w + i] = ix->d.fm[w].L2[i];
  }
  if (stats) for (int i = 0; i < 3; ++i) stats[i] = ix->build_stats[i];
  return 0;
}

extern "C" int bsq_index_download(const bsq_index *ix, int which, uint32_t *bwt, uint64_t *sa) {
  if (!ix || which < 0 || which > 1) return BSQ_EINVAL;
  int rc = 0;
  CKB(cudaSetDevice(ix->device));
  if (bwt) CKB(cudaMemcpy(bwt, ix->d.fm[which].blocks, ix->bwt_words[which] * 4, cudaMemcpyDeviceToHost));
  if (sa) CKB(cudaMemcpy(sa, ix->d.fm[which].sa, ix->n_sa[which] * 8, cudaMemcpyDeviceToHost));
done:
  return rc;
}
