// libbsq.so -- aligner half: CUDA kernels (sm_100a) and the C ABI of include/bsq.h.
//
// Phase 1 of `biscuit align` for a batch of (read, conversion) tasks runs as five kernels:
//   k_seed    one thread per task      SMEM seeding over the two FM-indices (random 64-B gathers)
//   k_expand  one thread per task      BWT ranks of the occurrences chaining will visit
//   k_sa      one thread per rank      sampled-SA lookup (LF walk, random 64-B gathers)
//   k_chain   one thread per task      B-tree chaining + chain filter
//   k_region  one thread per task      banded extension of the chained seeds -> regions
// Buffers between the stages are sized exactly from device-side prefix sums, so nothing is
// truncated; capacity violations raise BSQ_EOVERFLOW.
#include <cuda_runtime.h>
#include <condition_variable>
#include <mutex>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <cub/device/device_scan.cuh>

#include "../../include/bsq.h"
#include "bsq_internal.h"
#ifdef BSQ_INSTRUMENT
static __device__ unsigned long long bsq_ctr[8];
#endif
#include "bsq_task.h"
#include "bsq_seed3.cuh"
#include "bsq_ksw_warp.cuh"
#include "bsq_chain_warp.h"
#include "bsq_opt_default.h"

static_assert(sizeof(bsq_intv) == sizeof(bsq_intv_t), "abi");
static_assert(sizeof(bsq_reg) == sizeof(bsq_reg_t), "abi");
static_assert(sizeof(bsq_opt) == sizeof(bsq_devopt_t), "abi");

#define BSQ_TAIL_SLACK 64  // extra seed slots per task for the k >= max_occ tail (memchain.c:325-326)

static thread_local char g_err[512] = "";

#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) {                                                                       \
      snprintf(g_err, sizeof g_err, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
      return e_ == cudaErrorMemoryAllocation ? BSQ_ENOMEM : BSQ_ENODEV;                            \
    }                                                                                              \
  } while (0)

int g_bsq_wait_blocking = 0;
extern "C" void bsq_set_wait_mode(int blocking) { g_bsq_wait_blocking = blocking != 0; }

void bsq_set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
}

bsq_index *bsq_index_alloc(int device) {
  bsq_index *ix = (bsq_index *)calloc(1, sizeof(bsq_index));
  ix->device = device;
  return ix;
}

bool bsq_want_full_sa(uint64_t n, int halves_left) {
  const char *e = getenv("BSQ_FULL_SA");
  if (e) return atoi(e) != 0;
  size_t free_b = 0, total_b = 0;
  if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) return false;
  const double need = (double)halves_left * (double)(n + 1) * 8.0 + 40e9;  // leaves room for two aligner contexts
  return (double)free_b > need;
}

// SA of every rank by walking LF from each rank (for indices loaded from the reference's files, which only
// hold every 32nd entry): one thread per rank, same arithmetic as bwt_sa.
__global__ void k_derive_full_sa(bsq_fm_t fm, uint64_t n_ranks, uint64_t *full) {
  uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_ranks) return;
  bsq_fm_t f = fm;
  f.full_sa = nullptr;
  full[k] = bsq_sa(f, k);
}

void bsq_index_adopt(bsq_index *ix, void *p) {
  if (p && ix->n_allocs < 32) ix->allocs[ix->n_allocs++] = p;
}

int bsq_index_derive_b32(bsq_index *ix) {
  for (int w = 0; w < 2; ++w) {
    bsq_fm_t &f = ix->d.fm[w];
    const uint64_t n_half = 2 * ((f.seq_len + 127) / 128);
    uint32_t *b = nullptr;
    CK(cudaMalloc(&b, (n_half + 1) * 32));
    bsq_index_adopt(ix, b);
    CK(cudaMemsetAsync(b + n_half * 8, 0, 32));
    k_derive_b32<<<(unsigned)((n_half + 255) / 256), 256>>>(f.blocks, f.seq_len, n_half, b);
    CK(cudaGetLastError());
    f.b32 = b;
  }
  CK(cudaDeviceSynchronize());
  return 0;
}

// grow-only device buffer
struct DevBuf {
  void *p = nullptr;
  size_t cap = 0;
  int reserve(size_t bytes) {
    if (bytes <= cap) return 0;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    size_t want = bytes + bytes / 4 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {
      snprintf(g_err, sizeof g_err, "cudaMalloc(%zu): %s", want, cudaGetErrorString(e));
      return BSQ_ENOMEM;
    }
    cap = want;
    return 0;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
  template <typename T> T *as() const { return reinterpret_cast<T *>(p); }
};

// workspace of the seeding kernels (bsq_seed3.cuh): candidate lists of the forward sweeps, work queues
struct SeedWs {
  DevBuf cand1, cand2, calls1, calls2, items, q;
  int64_t calls1_cap = 0, calls2_cap = 0, items_cap = 0, cand2_cap = 0;  // grown when a batch needs more (the batch is re-seeded)
  void release() { cand1.release(); cand2.release(); calls1.release(); calls2.release(); items.release(); q.release(); }
};

struct bsq_aligner {
  const bsq_index *idx;
  SeedWs seed_ws;
  bsq_devopt_t opt;
  cudaStream_t stream, stream2;  // stream2: the large-task chaining tier, concurrent with the small tiers
  cudaEvent_t ev[8], ev_fork, ev_join;
  DevBuf seqs, lens, parent, intv, n_intv, n_sa, sa_off, ranks, pos, status;
  DevBuf snodes, wchains, bnodes, order, ochains, oseeds, n_chains, frac_rep, srt, regs_tmp, n_regs, reg_off, regs;
  DevBuf cub_tmp, scalars, fb_flag, tiers, spans, aflag;
  // results live in one of two slots (regs / reg_off and regs2 / reg_off2), alternating from run to run, so that a pipelined
  // caller can copy the regions of batch k to the host (bsq_aligner_fetch_slot, own stream) while batch k+1 is being run
  DevBuf regs2, reg_off2;
  cudaStream_t stream_out;
  int out_slot = 0;
  int64_t out_tasks[2] = {0, 0}, out_regs[2] = {-1, -1};
  // a slot named by bsq_aligner_result_slot is claimed until it is fetched (or released): the run that would overwrite it waits
  std::mutex slot_mu;
  std::condition_variable slot_cv;
  bool claimed[2] = {false, false};
  int64_t counters[16];
  int64_t n_staged = 0, n_regs_total = -1;
  int64_t fb_cap = 0;  // entries of the fallback-chaining workspace pools
  int64_t seed_fills[8] = {0, 0, 0, 0, 0, 0, 0, 0};  // seeding work-queue fills of the last batch: pass-1 calls, pass-2 items, pass-2 calls, pass-2 candidate records, retries
  int32_t stride = 0;
};

// ------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------

// Scratch of the seeding state machine on the device (interface: bsq_seed.h).  The first CAP candidates
// of each lane live in shared memory, [entry][lane] with 16-byte elements: whatever entries the 32 lanes
// address, every quarter-warp touches 8 distinct 16-byte bank groups, so accesses are conflict-free.
// Longer lists (rare: a list has one entry per distinct interval size along the sweep) continue in local
// memory.  The read is not copied: q() converts on the fly from an 8-byte register window over the batch
// buffer (the enclosing aligned word always lies inside the 256-byte-granular device allocation).
template <int CAP>
struct bsq_seed_scratch_dev {
  uint4 *sm;  // this lane's column; entry e at sm[e * 32]
  bsq_pk_t spill[BSQ_MAX_READ_LEN + 1 - CAP];
  const uint32_t *rd;  // this lane's converted read, 4 bits per base, word w at rd[w * 128] (shared memory)
  // Copy the read of the task into shared memory, converted for the index it is searched in (C>T parent, G>A daughter;
  // bseq_bsconvert, bwamem.c:161-178).  q() is on the critical path of every extension step: from here it is one
  // shared-memory word instead of a global load.
  __device__ __forceinline__ void bind(uint32_t *rd_sm, const uint8_t *s, int len, int par) {
    rd = rd_sm;
    const int nw = (len + 7) >> 3;
    const bool al8 = ((uintptr_t)s & 7) == 0;
    for (int w = 0; w < nw; ++w) {
      uint64_t v;
      if (al8) v = __ldg(reinterpret_cast<const uint64_t *>(s) + w);  // the row is padded to its stride
      else { v = 0; for (int k = 0; k < 8; ++k) if (8 * w + k < len) v |= (uint64_t)s[8 * w + k] << (8 * k); }
      uint32_t pk = 0;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        int c = (int)(v >> (8 * k)) & 0xf;
        c = par ? (c == 1 ? 3 : c) : (c == 2 ? 0 : c);
        pk |= (uint32_t)c << (4 * k);
      }
      rd_sm[w * 128] = pk;
    }
  }
  __device__ __forceinline__ bsq_pk_t get(int i) const {
    if (i >= CAP) return spill[i - CAP];
    const uint4 v = sm[i * 32];
    bsq_pk_t p;
    p.w0 = (uint64_t)v.x | (uint64_t)v.y << 32; p.w1 = (uint64_t)v.z | (uint64_t)v.w << 32;
    return p;
  }
  __device__ __forceinline__ void set(int i, const bsq_pk_t &p) {
    if (i >= CAP) { spill[i - CAP] = p; return; }
    sm[i * 32] = make_uint4((uint32_t)p.w0, (uint32_t)(p.w0 >> 32), (uint32_t)p.w1, (uint32_t)(p.w1 >> 32));
  }
  __device__ __forceinline__ int q(int i) const { return (int)(rd[(i >> 3) * 128] >> ((i & 7) * 4)) & 0xf; }
};

#ifndef BSQ_SEED_CAP
#define BSQ_SEED_CAP 8   // shared-memory candidates per lane (16 KB per 128-thread CTA); 16 measured 7 % slower: the smaller slice leaves more of the SM's 256 KB to L1, which the FM-index gathers use (profiles/README.md, r02)
#endif
#ifndef BSQ_SEED_CTAS
#define BSQ_SEED_CTAS 5   // resident CTAs per SM (registers: 96 per thread)
#endif

#include "bsq_seed_dev.cuh"

// First organisation of the seeding kernel (kept for comparison: BSQ_SEED_V1=1 in the environment selects it).
// Each lane owns one (read, conversion) task at a time and pulls the next one from a
// global counter when it finishes; all lanes of the warp meet at the single bsq_extend1 site per
// iteration so that their FM-index gathers are in flight together (see bsq_seed.h).  Starting and
// finishing a task cost a handful of instructions (no read conversion pass, no sort: k_seed_sort), so a
// lane that switches tasks does not hold up the other 31.
__global__ void __launch_bounds__(128, BSQ_SEED_CTAS) k_seed(const __grid_constant__ bsq_devopt_t opt, const __grid_constant__ bsq_devidx_t ix, int64_t n_tasks, const uint8_t *seqs, int stride,
                                              const int32_t *lens, const uint8_t *parent, int pipeline, bsq_pk_t *intv,
                                              int32_t *n_intv, int32_t *status, unsigned long long *next_task) {
  extern __shared__ uint4 seed_smem[];
  bsq_seed_scratch_dev<BSQ_SEED_CAP> scr;
  scr.sm = seed_smem + (threadIdx.x >> 5) * (BSQ_SEED_CAP * 32) + (threadIdx.x & 31);
  uint32_t *rd_sm = reinterpret_cast<uint32_t *>(seed_smem + 128 * BSQ_SEED_CAP) + threadIdx.x;  // [word][thread]
  bsq_seed_machine_t m;
  bsq_ext_req_t req;
  int64_t t = -1;
  int par = 0;
  bool have = false, exhausted = false;
  bsq_pk_t *out = nullptr;
  for (;;) {
    if (!have && !exhausted) {
      t = (int64_t)atomicAdd(next_task, 1ull);
      if (t >= n_tasks) exhausted = true;
      else {
        const int len = lens[t];
        par = parent[t] != 0;
        out = intv + t * BSQ_MAX_INTV;
        if (pipeline && len < opt.min_seed_len) n_intv[t] = 0;  // mem_chain returns before seeding
        else {
          scr.bind(rd_sm, seqs + t * stride, len, par);
          bsq_sm_init(m, opt, len, BSQ_MAX_INTV);
          have = true;
        }
      }
    }
    bool need = false;
    if (have) {
      need = bsq_sm_next(m, ix.fm[par], ix.fm[!par], scr, out, req);
      if (!need) {
        if (m.overflow) { atomicOr(status, 1); n_intv[t] = 0; }
        else n_intv[t] = m.n_out;
        have = false;
      }
    }
    if (__all_sync(0xffffffffu, exhausted && !have)) break;
    if (need) {
      uint64_t o0, o1, o2;
      bsq_extend1(ix.fm[par], ix.fm[!par], req, o0, o1, o2);
      bsq_sm_consume(m, scr, out, req, o0, o1, o2);
    }
  }
}

// Order each task's interval list (memchain.c:105) and count its SA lookups; one thread per task, all
// lanes busy (inside k_seed a finishing lane would sort while 31 lanes wait).
__global__ void __launch_bounds__(128) k_seed_sort(const __grid_constant__ bsq_devopt_t opt, int64_t n_tasks, bsq_pk_t *intv, int32_t *n_intv, int32_t *n_sa, int32_t *status) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_tasks) return;
  uint32_t keys[BSQ_MAX_INTV];
  int n = n_intv[t];
  if (n > BSQ_MAX_INTV) { n = 0; n_intv[t] = 0; atomicOr(status, 1); }  // the seeding kernels count past the capacity without storing
  n_sa[t] = n > 0 ? bsq_seed_sort(opt, intv + t * BSQ_MAX_INTV, n, keys) : 0;
}

__global__ void k_expand(const __grid_constant__ bsq_devopt_t opt, int64_t n_tasks, const bsq_pk_t *intv, const int32_t *n_intv, const uint8_t *parent,
                         const int64_t *sa_off, uint64_t *ranks) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_tasks) return;
  bsq_task_expand(opt, intv + t * BSQ_MAX_INTV, n_intv[t], (uint64_t)(parent[t] != 0) << 63, ranks + sa_off[t]);
}

__global__ void k_unpack_intv(int64_t n, const bsq_pk_t *in, bsq_intv_t *out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = bsq_pk_unpack(in[i]);
}

__global__ void k_sa(bsq_devidx_t ix, int64_t n, const uint64_t *ranks, uint64_t *pos) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint64_t r = ranks[i];
  pos[i] = bsq_sa(ix.fm[r >> 63], r & ~(1ull << 63));
}

__global__ void k_sa_plain(bsq_devidx_t ix, int which, int64_t n, const uint64_t *ranks, uint64_t *pos) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  pos[i] = bsq_sa(ix.fm[which], ranks[i]);
}

// bwt_occ4 (bwt.c:173-200) through the derived 32-byte blocks the seeding kernels use; cross-checked on the device against
// the reference-layout blocks (a mismatch returns all ones, which no rank equals)
__global__ void k_occ4(bsq_devidx_t ix, int which, int64_t n, const uint64_t *k, uint64_t *cnt) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint64_t c[4];
  bsq_occ4(ix.fm[which], k[i], c);
  const bsq_fm_t &f = ix.fm[which];
  if (f.b32 && k[i] != (uint64_t)-1) {
    const uint64_t k2 = k[i] - (k[i] >= f.primary);
    uint32_t w[8];
    s3_ld256(f.b32 + (k2 >> 6) * 8, w);
    uint64_t gt_prev = 0;
    for (int s = 3; s >= 0; --s) {
      uint64_t e, g;
      s3_rank(w, k2, s3_sym(s), e, g);
      if (e != c[s] || g != gt_prev) c[s] = ~0ull;
      gt_prev += e;
    }
  }
  cnt[4 * i] = c[0]; cnt[4 * i + 1] = c[1]; cnt[4 * i + 2] = c[2]; cnt[4 * i + 3] = c[3];
}

// per-task workspace offset: sa_off[t] + t * BSQ_TAIL_SLACK entries
__device__ __forceinline__ int64_t ws_off(const int64_t *sa_off, int64_t t) { return sa_off[t] + t * BSQ_TAIL_SLACK; }

// Exact chaining (B-tree replay), one thread per task, only for the tasks k_chain_warp flagged (fb_flag == 1).  Their
// workspaces come from a small pool by bump allocation (the flagged tasks are a handful per batch); if the pool is
// too small the task stays flagged and status bit 4 asks the host to retry with a larger pool -- nothing is dropped.
__global__ void __launch_bounds__(128) k_chain(const __grid_constant__ bsq_devopt_t opt, const __grid_constant__ bsq_devidx_t ix, int64_t n_tasks,
                                               const int32_t *lens, const uint8_t *parent, const bsq_pk_t *intv, const int32_t *n_intv,
                                               const int64_t *sa_off, const uint64_t *pos, bsq_snode_t *snodes,
                                               bsq_wchain_t *wchains, bsq_bnode_t *bnodes, int32_t *order, int64_t fb_cap,
                                               unsigned long long *fb_cursor, bsq_chain_t *ochains,
                                               bsq_seed_t *oseeds, int32_t *n_chains, float *frac_rep, int32_t *status,
                                               uint8_t *fb_flag) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_tasks) return;
  if (fb_flag[t] != 1) return;  // done by k_chain_warp (0) or by an earlier pass of this kernel (2)
  const int64_t wo = ws_off(sa_off, t);
  bsq_chain_ws_t ws;
  ws.cap = (int32_t)(sa_off[t + 1] - sa_off[t]) + BSQ_TAIL_SLACK;
  const int64_t fo = (int64_t)atomicAdd(fb_cursor, (unsigned long long)(ws.cap + 2));
  if (fo + ws.cap + 2 > fb_cap) { atomicOr(status, 4); return; }
  ws.snodes = snodes + fo; ws.chains = wchains + fo; ws.bnodes = bnodes + fo; ws.order = order + fo;
  bsq_chain_result_t r = bsq_chain_task(opt, ix, parent[t], lens[t], intv + t * BSQ_MAX_INTV, n_intv[t], pos + sa_off[t], ws,
                                        ochains + wo, oseeds + wo);
  if (r.status) { atomicOr(status, 2); r.n_chains = 0; }
  n_chains[t] = r.n_chains;
  frac_rep[t] = r.frac_rep;
  fb_flag[t] = 2;
}

// per-device one-time set-up (function attributes are per device; one host thread drives one device)
static inline int bsq_cur_device() { int dev = 0; cudaGetDevice(&dev); return dev < 0 || dev >= 64 ? 0 : dev; }
static inline int bsq_sm_count() {  // SMs of the current device (148 on a B200)
  static int n_sm = 0;
  if (!n_sm) { int dev = 0; cudaGetDevice(&dev); if (cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n_sm < 1) n_sm = 148; }
  return n_sm;
}

// warp policy of bsq_chain_warp
struct bsq_cw_warp {
  __device__ static int lane() { return threadIdx.x & 31; }
  __device__ static int nl() { return 32; }
  __device__ static void sync() { __syncwarp(); }
  __device__ static int first_true(bool p) { unsigned b = __ballot_sync(0xffffffffu, p); return b ? __ffs(b) - 1 : -1; }
  __device__ static bool any(bool p) { return __any_sync(0xffffffffu, p); }
  __device__ static int sum(int v) { return __reduce_add_sync(0xffffffffu, v); }
  __device__ static int min(int v) { return __reduce_min_sync(0xffffffffu, v); }
  __device__ static int scan_excl(int v, int &total) {  // exclusive prefix sum over the lanes (small non-negative values)
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int u = __shfl_up_sync(0xffffffffu, incl, o);
      if ((threadIdx.x & 31) >= o) incl += u;
    }
    total = __shfl_sync(0xffffffffu, incl, 31);
    return incl - v;
  }
  // bitonic sort of n keys (unique, so the result is the total order).  Up to 256 keys are sorted in registers: key i
  // sits in register i / 32 of lane i % 32, partners at distance >= 32 are registers of the same lane, partners at
  // distance < 32 come by shuffle -- no shared-memory round trip and no barrier per stage.  Larger tasks (the 1024-seed
  // tier) sort in shared memory.
  template <int R>
  __device__ static void sort_regs(uint64_t *k, int n) {
    const int lane = threadIdx.x & 31;
    uint64_t v[R];
#pragma unroll
    for (int m = 0; m < R; ++m) { const int i = m * 32 + lane; v[m] = i < n ? k[i] : ~0ull; }
#pragma unroll
    for (int kk = 2; kk <= 32 * R; kk <<= 1) {
#pragma unroll
      for (int j = kk >> 1; j > 0; j >>= 1) {
        if (j >= 32) {
#pragma unroll
          for (int m = 0; m < R; ++m) {
            const int x = m ^ (j >> 5);
            if (x > m) {
              const bool up = ((m * 32) & kk) == 0;
              const uint64_t a = v[m], b = v[x];
              if ((a > b) == up) { v[m] = b; v[x] = a; }
            }
          }
        } else {
#pragma unroll
          for (int m = 0; m < R; ++m) {
            const uint64_t o = __shfl_xor_sync(0xffffffffu, v[m], j);
            const bool up = ((m * 32 + lane) & kk) == 0, lower = (lane & j) == 0;
            v[m] = (up == lower) ? (v[m] < o ? v[m] : o) : (v[m] > o ? v[m] : o);
          }
        }
      }
    }
    __syncwarp();
#pragma unroll
    for (int m = 0; m < R; ++m) { const int i = m * 32 + lane; if (i < n) k[i] = v[m]; }
    __syncwarp();
  }
  // The partition phase of ks_introsort (ksort.h:184-221; bsq_introsort<false> is the sequential statement), with every
  // Hoare partition done by the whole warp.  Sequentially the i scan stops at the positions x > s with !(a[x] < pivot), in
  // ascending order, the j scan at the positions x < t with !(pivot < a[x]), in descending order; both scans only ever
  // look at values nobody has swapped yet, except that the i scan also stops at the position of the last swap's right
  // element.  So swap number k exchanges the k-th stop of i (I_k) with the k-th stop of j (J_k) as long as I_k < J_k, the
  // number of swaps K is the count of such k (the predicate is monotone), and the scan ends at i = min(I_{K+1}, J_K)
  // (the pivot, parked at t, is the last I).  The stops come from ballots, 32 positions at a time; the swaps are disjoint.
  // a[] and the two position lists live in shared memory; all lanes hold the same control state.
  __device__ static void weight_partitions(uint32_t *a, int n, uint16_t *scratch) {
    const int lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    bsq_cw_by_weight lt;
    if (n < 1) return;
#ifdef BSQ_CW_SEQ_PARTITION  // A/B switch: the sequential replay on lane 0
    if (lane == 0) bsq_introsort<false>(a, (int64_t)n, lt);
    __syncwarp();
    return;
#endif
    if (n == 2) {
      if (lane == 0 && lt(a[1], a[0])) { const uint32_t x = a[0]; a[0] = a[1]; a[1] = x; }
      __syncwarp();
      return;
    }
    uint16_t *ipos = scratch, *jpos = scratch + n + 1;
    int d = 2;
    while ((1u << d) < (unsigned)n) ++d;
    int st_l[40], st_r[40], st_d[40], top = 0;
    int s = 0, t = n - 1;
    d <<= 1;
    __syncwarp();
    for (;;) {
      if (s < t) {
        if (--d == 0) {  // depth limit: comb sort of the range (ksort.h:198-202), sequential
          if (lane == 0) bsq_combsort(a + s, (int64_t)(t - s + 1), lt);
          __syncwarp();
          t = s;
          continue;
        }
        int k = s + ((t - s) >> 1) + 1;
        if (lt(a[k], a[s])) { if (lt(a[k], a[t])) k = t; }
        else k = lt(a[t], a[s]) ? s : t;
        const uint32_t rp = a[k];
        __syncwarp();
        if (k != t && lane == 0) { a[k] = a[t]; a[t] = rp; }
        __syncwarp();
        int n_i = 0, n_j = 0;
        for (int x0 = s + 1; x0 <= t; x0 += 32) {
          const int x = x0 + lane;
          const bool g = x <= t && !lt(a[x], rp);  // true at x == t (the pivot)
          const unsigned m = __ballot_sync(0xffffffffu, g);
          if (g) ipos[n_i + __popc(m & lt_mask)] = (uint16_t)x;
          n_i += __popc(m);
        }
        for (int x0 = t - 1; x0 > s; x0 -= 32) {
          const int x = x0 - lane;
          const bool l = x > s && !lt(rp, a[x]);
          const unsigned m = __ballot_sync(0xffffffffu, l);
          if (l) jpos[n_j + __popc(m & lt_mask)] = (uint16_t)x;
          n_j += __popc(m);
        }
        __syncwarp();
        const int n_min = n_i < n_j ? n_i : n_j;
        int K = 0;
        for (int k0 = 0; k0 < n_min; k0 += 32) {
          const int kk = k0 + lane;
          const unsigned m = __ballot_sync(0xffffffffu, kk < n_min && ipos[kk] < jpos[kk]);
          K += __popc(m);
          if (m != 0xffffffffu) break;
        }
        int i = ipos[K];  // K < n_i: the pivot's position t is the last stop of i and no stop of j is right of it
        if (K > 0) { const int jl = jpos[K - 1]; i = i < jl ? i : jl; }
        for (int kk = lane; kk < K; kk += 32) { const int xi = ipos[kk], xj = jpos[kk]; const uint32_t v = a[xi]; a[xi] = a[xj]; a[xj] = v; }
        __syncwarp();
        if (lane == 0) { const uint32_t v = a[i]; a[i] = a[t]; a[t] = v; }
        __syncwarp();
        if (i - s > t - i) {
          if (i - s > 16) { st_l[top] = s; st_r[top] = i - 1; st_d[top] = d; ++top; }
          s = t - i > 16 ? i + 1 : t;
        } else {
          if (t - i > 16) { st_l[top] = i + 1; st_r[top] = t; st_d[top] = d; ++top; }
          t = i - s > 16 ? i - 1 : s;
        }
      } else {
        if (top == 0) return;
        --top;
        s = st_l[top]; t = st_r[top]; d = st_d[top];
      }
    }
  }
  __device__ static void sort_keys(uint64_t *k, int n) {
    const int lane = threadIdx.x & 31;
    if (n <= 32) { sort_regs<1>(k, n); return; }
    if (n <= 64) { sort_regs<2>(k, n); return; }
    if (n <= 128) { sort_regs<4>(k, n); return; }
    if (n <= 256) { sort_regs<8>(k, n); return; }
    int P = 512;
    while (P < n) P <<= 1;
    for (int i = n + lane; i < P; i += 32) k[i] = ~0ull;
    __syncwarp();
    for (int kk = 2; kk <= P; kk <<= 1)
      for (int j = kk >> 1; j > 0; j >>= 1) {
        for (int i = lane; i < P; i += 32) {
          const int x = i ^ j;
          if (x > i) {
            const uint64_t a = k[i], b = k[x];
            if ((a > b) == ((i & kk) == 0)) { k[i] = b; k[x] = a; }
          }
        }
        __syncwarp();
      }
  }
};

// Tasks binned by their number of seeds into BSQ_N_TIERS capacity classes (k_chain_warp is instantiated once per
// class).  Order inside a tier is arbitrary (every task writes to its own slots).
#define BSQ_N_TIERS 7
__constant__ int c_tier_cap[BSQ_N_TIERS] = {64, 96, 128, 160, 192, 256, 1 << 30};
__global__ void k_chain_tiers(int64_t n_tasks, const int32_t *n_sa, int32_t *tier_list, unsigned long long *tier_cnt) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_tasks) return;
  const int ns = n_sa[t];
  int tier = 0;
  while (ns > c_tier_cap[tier]) ++tier;
  // one atomic per (warp, tier)
  for (int q = 0; q < BSQ_N_TIERS; ++q) {
    const unsigned m = __ballot_sync(__activemask(), tier == q);
    if (tier == q) {
      const int leader = __ffs(m) - 1, lane = threadIdx.x & 31;
      unsigned long long base = 0;
      if (lane == leader) base = atomicAdd(&tier_cnt[q], (unsigned long long)__popc(m));
      base = __shfl_sync(m, base, leader);
      tier_list[q * n_tasks + base + __popc(m & ((1u << lane) - 1))] = (int32_t)t;
    }
  }
}

// Chaining + chain filter, one WARP per task, state in shared memory (bsq_chain_warp.h).  Instantiated for several
// slice capacities (one launch per tier of k_chain_tiers), so small tasks run at high occupancy.  Tasks the decomposition cannot take exactly are
// flagged for k_chain (thread-per-task, exact B-tree replay).
template <int CAP, int WPB>
__global__ void __launch_bounds__(32 * WPB) k_chain_warp(const __grid_constant__ bsq_devopt_t opt, const __grid_constant__ bsq_devidx_t ix, int64_t n_tasks,
                                                    const int32_t *tier_list, const unsigned long long *tier_cnt, const int32_t *lens,
                                                    const uint8_t *parent, const bsq_pk_t *intv, const int32_t *n_intv,
                                                    const int32_t *n_sa, const int64_t *sa_off, const uint64_t *pos, bsq_chain_t *ochains,
                                                    bsq_seed_t *oseeds, int32_t *n_chains, float *frac_rep, uint8_t *fb_flag,
                                                    unsigned long long *n_fallback, unsigned long long *cursor) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  typedef bsq_cw_smem_tt<CAP> S;
  S *sm = reinterpret_cast<S *>(smem_raw) + (threadIdx.x >> 5);
  const int64_t n_tier = (int64_t)*tier_cnt;
  // persistent warps: the grid is one wave of resident CTAs and every warp pulls the next task of the tier when it is
  // done (tasks differ by an order of magnitude in work; a CTA that waits for its slowest warp leaves slots idle)
  for (;;) {
    unsigned long long wi_ = 0;
    if ((threadIdx.x & 31) == 0) wi_ = atomicAdd(cursor, 1ull);
    const int64_t wi = (int64_t)__shfl_sync(0xffffffffu, wi_, 0);
    if (wi >= n_tier) return;  // whole warps leave together
    const int64_t t = tier_list[wi];
    const int ns = n_sa[t];
    const int64_t wo = ws_off(sa_off, t);
    bsq_chain_result_t r;
    const int rc = bsq_chain_warp<bsq_cw_warp>(opt, ix, parent[t], lens[t], intv + t * BSQ_MAX_INTV, n_intv[t], pos + sa_off[t], ns, *sm,
                                               ochains + wo, oseeds + wo, r);
    if ((threadIdx.x & 31) == 0) {
      if (rc == BSQ_CW_OK) { n_chains[t] = r.n_chains; frac_rep[t] = r.frac_rep; fb_flag[t] = 0; }
      else { fb_flag[t] = 1; atomicAdd(n_fallback, 1ull); }
    }
    __syncwarp();  // the next task reuses the warp's shared-memory slice
  }
}

template <int CAP, int WPB>
static int launch_chain_warp(cudaStream_t s, const bsq_devopt_t &opt, const bsq_devidx_t &ix, int64_t n, const int32_t *tier_list,
                             const unsigned long long *tier_cnt, const int32_t *lens,
                             const uint8_t *parent, const bsq_pk_t *intv, const int32_t *n_intv, const int32_t *n_sa, const int64_t *sa_off,
                             const uint64_t *pos, bsq_chain_t *ochains, bsq_seed_t *oseeds, int32_t *n_chains, float *frac_rep, uint8_t *fb_flag,
                             unsigned long long *n_fallback, unsigned long long *cursor) {
  static int resident_[64];  // CTAs of this instantiation per SM, per device
  int &resident = resident_[bsq_cur_device()];
  const size_t smem = WPB * sizeof(bsq_cw_smem_tt<CAP>);
  if (!resident) {
    CK(cudaFuncSetAttribute(k_chain_warp<CAP, WPB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int r = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&r, k_chain_warp<CAP, WPB>, 32 * WPB, smem));
    resident = r < 1 ? 1 : r;
  }
  const int64_t want = (n + WPB - 1) / WPB, wave = (int64_t)bsq_sm_count() * resident;
  k_chain_warp<CAP, WPB><<<(unsigned)(want < wave ? want : wave), 32 * WPB, smem, s>>>(opt, ix, n, tier_list, tier_cnt, lens, parent, intv, n_intv, n_sa, sa_off, pos,
                                                                                    ochains, oseeds, n_chains, frac_rep, fb_flag, n_fallback, cursor);
  CK(cudaGetLastError());
  return 0;
}

// Inputs of k_region that do not depend on the regions found so far, computed one chain / one seed per lane (k_region runs
// its control flow redundantly in the 32 lanes of the task's warp, so everything done here costs a fraction of what it cost
// there; measured shares of k_region before: chain loop 14 %, seed sort 7 %, asymmetric filter 13 %, profiles/README.md r02):
//   spans[2 c], spans[2 c + 1]   reference window of chain c (mem_chain_reference_span + bns_fetch_seq clipping)
//   srt[seed_off ..]             seed order of mem_chain2region1 (memchain.c:750-757: by score = len, then index; the keys are
//                                unique, so any sort gives the reference's order), backup seeds behind the main ones
//   aflag[seed]                  verdict of asymmetric_flt_seed (memchain.c:138-149), one seed per lane
__global__ void __launch_bounds__(128) k_region_prep(const __grid_constant__ bsq_devopt_t opt, const __grid_constant__ bsq_devidx_t ix, int64_t n_tasks,
                                                     const uint8_t *seqs, int stride, const int32_t *lens, const int64_t *sa_off, const bsq_chain_t *ochains,
                                                     const bsq_seed_t *oseeds, const int32_t *n_chains, uint64_t *srt, int64_t *spans, uint8_t *aflag) {
  bsq_gap_tab_init(opt);
  struct u64_less { __device__ bool operator()(uint64_t a, uint64_t b) const { return a < b; } };
  const int lane = threadIdx.x & 31;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t t = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; t < n_tasks; t += n_warps) {
    const int64_t wo = ws_off(sa_off, t);
    const int nc = n_chains[t], lq = lens[t];
    const uint8_t *query = seqs + t * stride;
    int n_seeds_task = 0;  // the chains' seed slices tile the task's seed pool from 0
    for (int ci = lane; ci < nc; ci += 32) {
      const bsq_chain_t c = ochains[wo + ci];
      if (c.n_seeds == 0) continue;
      const bsq_seed_t *cs = oseeds + wo + c.seed_off;
      const int e = c.seed_off + c.n_seeds + c.n_extra;
      n_seeds_task = n_seeds_task > e ? n_seeds_task : e;
      int64_t r0, r1;
      bsq_chain_span<bsq_warp_policy>(opt, ix, lq, c, cs, r0, r1);
      spans[2 * (wo + ci)] = r0; spans[2 * (wo + ci) + 1] = r1;
      for (int part = 0; part < 2; ++part) {
        const bsq_seed_t *sd = part ? cs + c.n_seeds : cs;
        const int n = part ? c.n_extra : c.n_seeds;
        uint64_t *keys = srt + wo + c.seed_off + (part ? c.n_seeds : 0);
        for (int i = 0; i < n; ++i) keys[i] = (uint64_t)(uint32_t)sd[i].len << 32 | (uint32_t)i;
        if (n > 1) bsq_introsort(keys, (int64_t)n, u64_less());
      }
    }
    n_seeds_task = __reduce_max_sync(0xffffffffu, n_seeds_task);
    for (int j = lane; j < n_seeds_task; j += 32) aflag[wo + j] = bsq_asym_seed(ix, oseeds[wo + j], query) ? 1 : 0;
  }
}

// Chains -> regions, one WARP per task: the control flow of mem_chain2region runs uniformly in all
// lanes, every banded extension is spread over the lanes (bsq_ksw_warp.cuh), lane 0 stores.
#ifndef BSQ_REGION_CTAS
#define BSQ_REGION_CTAS 8
#endif
__global__ void __launch_bounds__(128, BSQ_REGION_CTAS) k_region(const __grid_constant__ bsq_devopt_t opt, const __grid_constant__ bsq_devidx_t ix, int64_t n_tasks, const uint8_t *seqs, int stride,
                                                const int32_t *lens, const uint8_t *parent, const int64_t *sa_off,
                                                const bsq_chain_t *ochains, const bsq_seed_t *oseeds, const int32_t *n_chains,
                                                const float *frac_rep, uint64_t *srt, const int64_t *spans, const uint8_t *aflag, bsq_reg_t *regs_tmp, int32_t *n_regs,
                                                unsigned long long *cursor) {
  bsq_gap_tab_init(opt);
  for (;;) {  // persistent warps: one wave of resident CTAs, every warp pulls the next task (see k_chain_warp)
    unsigned long long t_ = 0;
    if ((threadIdx.x & 31) == 0) t_ = atomicAdd(cursor, 1ull);
    const int64_t t = (int64_t)__shfl_sync(0xffffffffu, t_, 0);
    if (t >= n_tasks) return;  // whole warps leave together
    const int64_t wo = ws_off(sa_off, t);
    const int n = bsq_chain2region<bsq_warp_policy, true>(opt, ix, parent[t], lens[t], seqs + t * stride, ochains + wo, n_chains[t], oseeds + wo,
                                                          frac_rep[t], srt + wo, nullptr, regs_tmp + wo, spans + 2 * wo, aflag + wo);
    if ((threadIdx.x & 31) == 0) n_regs[t] = n;
  }
}

__global__ void k_compact_regs(int64_t n_tasks, const int64_t *sa_off, const int32_t *n_regs, const int64_t *reg_off,
                               const bsq_reg_t *regs_tmp, bsq_reg_t *regs) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_tasks) return;
  const bsq_reg_t *src = regs_tmp + ws_off(sa_off, t);
  bsq_reg_t *dst = regs + reg_off[t];
  for (int i = 0; i < n_regs[t]; ++i) dst[i] = src[i];
}


// batched ksw_extend2 jobs through the warp-cooperative kernel that k_region uses (one warp per job)
__global__ void __launch_bounds__(128) k_extend_warp(const __grid_constant__ bsq_devopt_t opt, int64_t n_jobs, const uint8_t *qbuf, const int64_t *qoff,
                                                     const int32_t *qlen, const uint8_t *tbuf, const int64_t *toff, const int32_t *tlen,
                                                     const uint8_t *is_parent, const int32_t *w, const int32_t *h0, int32_t *out) {
  const int64_t j = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (j >= n_jobs) return;
  bsq_qacc_t qa; qa.q = qbuf + qoff[j]; qa.step = 1;
  bsq_tacc_t ta; ta.ix = nullptr; ta.buf = tbuf + toff[j]; ta.p0 = 0; ta.step = 1;
  bsq_ext_result_t r = bsq_ksw_extend_warp(qlen[j], qa, tlen[j], ta, is_parent[j] ? opt.ctmat : opt.gamat, opt.o_del, opt.e_del,
                                           opt.o_ins, opt.e_ins, w[j], opt.pen_clip5, opt.zdrop, h0[j]);
  if ((threadIdx.x & 31) != 0) return;
  int32_t *o = out + 6 * j;
  o[0] = r.score; o[1] = r.qle; o[2] = r.tle; o[3] = r.gtle; o[4] = r.gscore; o[5] = r.max_off;
}

// ------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------

static inline unsigned nblk(int64_t n, int bs) { return (unsigned)((n + bs - 1) / bs); }


// k_seed is persistent (lanes pull tasks from a counter): one wave of 148 SMs x 3 resident CTAs of 128
// shared memory of k_seed: candidate lists + the converted reads (4 bits per base, as many words as the longest row needs)
static inline size_t seed_smem_bytes(int stride) {  // reads are at most BSQ_MAX_READ_LEN long whatever the row stride
  const int len = stride < BSQ_MAX_READ_LEN ? stride : BSQ_MAX_READ_LEN;
  return (size_t)128 * BSQ_SEED_CAP * 16 + (size_t)((len + 7) >> 3) * 128 * 4;
}
// BSQ_SEED_VARIANT=0 launches k_seed2 without the round-1 v7 changes (bsq_seed_dev.cuh).  Read at every launch:
// tools/ab_seed.py toggles it inside one process.
static inline int seed_variant() { const char *e = getenv("BSQ_SEED_VARIANT"); return e ? atoi(e) != 0 : 1; }
static void launch_seed2(int variant, unsigned grid, cudaStream_t s, const bsq_devopt_t &opt, const bsq_devidx_t &ix, int64_t n, const uint8_t *seqs, int stride,
                         const int32_t *lens, const uint8_t *parent, int pipeline, bsq_pk_t *intv, int32_t *n_intv, int32_t *status,
                         unsigned long long *next_task) {
  const size_t smem = seed_smem_bytes(stride);
  if (variant) k_seed2<BSQ_SEED_CAP, 1><<<grid, 128, smem, s>>>(opt, ix, n, seqs, stride, lens, parent, pipeline, intv, n_intv, status, next_task);
  else k_seed2<BSQ_SEED_CAP, 0><<<grid, 128, smem, s>>>(opt, ix, n, seqs, stride, lens, parent, pipeline, intv, n_intv, status, next_task);
}
static inline bool seed_v1() { static int v = -1; if (v < 0) { const char *e = getenv("BSQ_SEED_V1"); v = e && atoi(e) != 0; } return v != 0; }
static inline unsigned seed_grid(int64_t n) {
  static bool attr_set_[64];
  bool &attr_set = attr_set_[bsq_cur_device()];
  if (!attr_set) {
    cudaFuncSetAttribute(k_seed, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)seed_smem_bytes(BSQ_MAX_READ_LEN + 8));
    cudaFuncSetAttribute(k_seed2<BSQ_SEED_CAP, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)seed_smem_bytes(BSQ_MAX_READ_LEN + 8));
    cudaFuncSetAttribute(k_seed2<BSQ_SEED_CAP, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)seed_smem_bytes(BSQ_MAX_READ_LEN + 8));
    attr_set = true;
  }
  int64_t want = (n + 127) / 128;
  return (unsigned)(want < 148 * BSQ_SEED_CTAS ? want : 148 * BSQ_SEED_CTAS);
}

// BSQ_SEED_IMPL=2 selects the single-kernel state machine (k_seed2) instead of the per-pass kernels of bsq_seed3.cuh
static inline int seed_impl() { const char *e = getenv("BSQ_SEED_IMPL"); return e ? atoi(e) : 3; }

// mem_collect_intv for n tasks: unsorted interval lists in intv / n_intv (n_intv zeroed here).  Synchronises the stream
// once (the queue fills come back to size the retry when a work queue was too small).
static int seed3_run(SeedWs &ws, cudaStream_t s, const bsq_devopt_t &opt, const bsq_devidx_t &ix, int64_t n, const uint8_t *seqs, int stride,
                     const int32_t *lens, const uint8_t *parent, int pipeline, bsq_pk_t *intv, int32_t *n_intv, int64_t *fills) {
  int rc;
  if (!ix.fm[0].b32 || !ix.fm[1].b32) { snprintf(g_err, sizeof g_err, "index without derived rank blocks"); return BSQ_EINVAL; }
  if (ws.calls1_cap < 8 * n + 1024) ws.calls1_cap = 8 * n + 1024;
  if (ws.items_cap < 4 * n + 1024) ws.items_cap = 4 * n + 1024;
  if (ws.calls2_cap < ws.items_cap) ws.calls2_cap = ws.items_cap;
  if (ws.cand2_cap < n * (int64_t)stride) ws.cand2_cap = n * (int64_t)stride;
  { const char *e = getenv("BSQ_SEED_QCAP"); if (e && atoll(e) > 0 && fills) { ws.calls1_cap = ws.calls2_cap = ws.items_cap = atoll(e); ws.cand2_cap = 64 * atoll(e); } }  // test hook: force the retry path
  const unsigned grid = (unsigned)((n + 127) / 128 < 148 * BSQ_S3_CTAS ? (n + 127) / 128 : 148 * BSQ_S3_CTAS);
  for (int attempt = 0; attempt < 12; ++attempt) {
    if ((rc = ws.cand1.reserve((size_t)n * stride * 16))) return rc;
    if ((rc = ws.cand2.reserve((size_t)ws.cand2_cap * 16))) return rc;
    if ((rc = ws.calls1.reserve((size_t)ws.calls1_cap * sizeof(s3_call_t)))) return rc;
    if ((rc = ws.calls2.reserve((size_t)ws.calls2_cap * sizeof(s3_call_t)))) return rc;
    if ((rc = ws.items.reserve((size_t)ws.items_cap * sizeof(s3_item_t)))) return rc;
    if ((rc = ws.q.reserve(sizeof(s3_q_t)))) return rc;
    s3_q_t *q = ws.q.as<s3_q_t>();
    CK(cudaMemsetAsync(q, 0, sizeof(s3_q_t), s));
    CK(cudaMemsetAsync(n_intv, 0, (size_t)n * 4, s));
    k_s3_fwd<1><<<grid, 128, 0, s>>>(opt, ix, n, seqs, stride, lens, parent, pipeline, ws.cand1.as<uint4>(), 0ull, nullptr, ws.calls1.as<s3_call_t>(),
                                     (unsigned long long)ws.calls1_cap, ws.items.as<s3_item_t>(), (unsigned long long)ws.items_cap, q, intv, n_intv);
    k_s3_greedy<<<grid, 128, 0, s>>>(opt, ix, n, seqs, stride, lens, parent, pipeline, q, intv, n_intv);
    k_s3_bwd<<<grid, 128, 0, s>>>(opt, ix, seqs, stride, parent, ws.cand1.as<uint4>(), ws.calls1.as<s3_call_t>(), (unsigned long long)ws.calls1_cap, 1,
                                  ws.items.as<s3_item_t>(), (unsigned long long)ws.items_cap, q, intv, n_intv);
    k_s3_fwd<2><<<grid, 128, 0, s>>>(opt, ix, n, seqs, stride, lens, parent, pipeline, ws.cand2.as<uint4>(), (unsigned long long)ws.cand2_cap,
                                     ws.items.as<s3_item_t>(), ws.calls2.as<s3_call_t>(), (unsigned long long)ws.calls2_cap, ws.items.as<s3_item_t>(),
                                     (unsigned long long)ws.items_cap, q, intv, n_intv);
    k_s3_bwd<<<grid, 128, 0, s>>>(opt, ix, seqs, stride, parent, ws.cand2.as<uint4>(), ws.calls2.as<s3_call_t>(), (unsigned long long)ws.calls2_cap, 2,
                                  ws.items.as<s3_item_t>(), (unsigned long long)ws.items_cap, q, intv, n_intv);
    CK(cudaGetLastError());
    s3_q_t h;
    CK(cudaMemcpyAsync(&h, q, sizeof h, cudaMemcpyDeviceToHost, s));
    CK(bsq_stream_wait(s));
    if (fills) { fills[0] = (int64_t)h.n_calls1; fills[1] = (int64_t)h.n_items; fills[2] = (int64_t)h.n_calls2; fills[3] = (int64_t)h.cand2_used; fills[4] = attempt; }
    if (!h.overflow) return 0;
    // a queue was too small: size it from what this attempt asked for and seed the batch again
    if ((int64_t)h.n_calls1 > ws.calls1_cap) ws.calls1_cap = (int64_t)h.n_calls1 + (int64_t)h.n_calls1 / 4;
    if ((int64_t)h.n_items > ws.items_cap) ws.items_cap = (int64_t)h.n_items + (int64_t)h.n_items / 4;
    if (ws.calls2_cap < ws.items_cap) ws.calls2_cap = ws.items_cap;
    if ((int64_t)h.n_calls2 > ws.calls2_cap) ws.calls2_cap = (int64_t)h.n_calls2 + (int64_t)h.n_calls2 / 4;
    if ((int64_t)h.cand2_used > ws.cand2_cap) ws.cand2_cap = (int64_t)h.cand2_used + (int64_t)h.cand2_used / 4;
    if (h.overflow & 1ull) { ws.cand2_cap *= 2; }  // items were dropped, so cand2_used undercounts
  }
  snprintf(g_err, sizeof g_err, "seeding work queues still too small after 12 attempts");
  return BSQ_EOVERFLOW;
}

extern "C" {

const char *bsq_strerror(int code) {
  switch (code) {
    case BSQ_OK: return "ok";
    case BSQ_ENODEV: return "CUDA device/runtime error";
    case BSQ_EINVAL: return "invalid argument";
    case BSQ_EOVERFLOW: return "per-read capacity exceeded";
    case BSQ_ENOMEM: return "out of device memory";
  }
  return "unknown error";
}

const char *bsq_last_error(void) { return g_err; }

static int upload_array(bsq_index *ix, const void *src, size_t bytes, const void **dst) {
  void *p = nullptr;
  CK(cudaMalloc(&p, bytes ? bytes : 16));
  ix->allocs[ix->n_allocs++] = p;
  if (bytes) CK(cudaMemcpy(p, src, bytes, cudaMemcpyHostToDevice));
  *dst = p;
  return 0;
}

int bsq_index_upload(const bsq_index_desc *h, int device, bsq_index **out) {
  if (!h || !out) return BSQ_EINVAL;
  int ndev = 0;
  CK(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) { snprintf(g_err, sizeof g_err, "device %d of %d", device, ndev); return BSQ_ENODEV; }
  CK(cudaSetDevice(device));
  bsq_index *ix = bsq_index_alloc(device);
  int rc = 0;
  for (int w = 0; w < 2; ++w) { ix->bwt_words[w] = h->bwt_words[w]; ix->n_sa[w] = h->n_sa[w]; }
  for (int w = 0; w < 2 && !rc; ++w) {
    bsq_fm_t &f = ix->d.fm[w];
    f.primary = h->primary[w]; f.seq_len = h->seq_len; f.sa_intv = h->sa_intv[w];
    for (int i = 0; i < 5; ++i) f.L2[i] = h->L2[w][i];
    rc = upload_array(ix, h->bwt[w], h->bwt_words[w] * 4, (const void **)&f.blocks);
    if (!rc) rc = upload_array(ix, h->sa[w], h->n_sa[w] * 8, (const void **)&f.sa);
  }
  if (!rc) rc = upload_array(ix, h->pac, (size_t)(h->l_pac / 4 + 1), (const void **)&ix->d.pac);
  if (!rc) rc = upload_array(ix, h->ann_offset, (size_t)h->n_seqs * 8, (const void **)&ix->d.ann_offset);
  if (!rc) rc = upload_array(ix, h->ann_len, (size_t)h->n_seqs * 4, (const void **)&ix->d.ann_len);
  if (!rc) rc = upload_array(ix, h->ann_is_alt, (size_t)h->n_seqs * 4, (const void **)&ix->d.ann_is_alt);
  ix->d.l_pac = h->l_pac; ix->d.n_seqs = h->n_seqs;
  if (!rc) rc = bsq_index_derive_b32(ix);
  if (rc) { bsq_index_free(ix); return rc; }
  for (int w = 1; w >= 0; --w) {  // optional full SA (see bsq_want_full_sa)
    ix->d.fm[w].full_sa = nullptr;
    if (!bsq_want_full_sa(h->seq_len, w + 1)) continue;
    uint64_t *full = nullptr;
    if (cudaMalloc(&full, (h->seq_len + 1) * 8) != cudaSuccess) { cudaGetLastError(); continue; }
    k_derive_full_sa<<<(unsigned)((h->seq_len + 1 + 255) / 256), 256>>>(ix->d.fm[w], h->seq_len + 1, full);
    if (cudaDeviceSynchronize() != cudaSuccess) { cudaGetLastError(); cudaFree(full); continue; }
    ix->allocs[ix->n_allocs++] = full;
    ix->d.fm[w].full_sa = full;
  }
  *out = ix;
  return 0;
}

void bsq_index_free(bsq_index *ix) {
  if (!ix) return;
  for (int i = 0; i < ix->n_allocs; ++i) cudaFree(ix->allocs[i]);
  free(ix);
}

int bsq_occ4(const bsq_index *ix, int which, int64_t n, const uint64_t *k, uint64_t *cnt) {
  if (!ix || which < 0 || which > 1 || n < 0) return BSQ_EINVAL;
  if (n == 0) return 0;
  CK(cudaSetDevice(ix->device));
  uint64_t *dk = nullptr, *dc = nullptr;
  CK(cudaMalloc(&dk, n * 8)); CK(cudaMalloc(&dc, n * 32));
  CK(cudaMemcpy(dk, k, n * 8, cudaMemcpyHostToDevice));
  k_occ4<<<nblk(n, 256), 256>>>(ix->d, which, n, dk, dc);
  CK(cudaGetLastError());
  CK(cudaMemcpy(cnt, dc, n * 32, cudaMemcpyDeviceToHost));
  cudaFree(dk); cudaFree(dc);
  return 0;
}

int bsq_sa_lookup(const bsq_index *ix, int which, int64_t n, const uint64_t *k, uint64_t *pos) {
  if (!ix || which < 0 || which > 1 || n < 0) return BSQ_EINVAL;
  if (n == 0) return 0;
  CK(cudaSetDevice(ix->device));
  uint64_t *dk = nullptr, *dp = nullptr;
  CK(cudaMalloc(&dk, n * 8)); CK(cudaMalloc(&dp, n * 8));
  CK(cudaMemcpy(dk, k, n * 8, cudaMemcpyHostToDevice));
  k_sa_plain<<<nblk(n, 256), 256>>>(ix->d, which, n, dk, dp);
  CK(cudaGetLastError());
  CK(cudaMemcpy(pos, dp, n * 8, cudaMemcpyDeviceToHost));
  cudaFree(dk); cudaFree(dp);
  return 0;
}

static int check_lens(int64_t n, const int32_t *lens, int32_t stride) {
  for (int64_t i = 0; i < n; ++i)
    if (lens[i] < 0 || lens[i] > BSQ_MAX_READ_LEN || lens[i] > stride) {
      snprintf(g_err, sizeof g_err, "read %lld has length %d (max %d, stride %d)", (long long)i, lens[i], BSQ_MAX_READ_LEN, stride);
      return BSQ_EINVAL;
    }
  return 0;
}

int bsq_collect_intv(const bsq_index *ix, const bsq_opt *opt_, int64_t n, const uint8_t *seqs, int32_t stride,
                     const int32_t *lens, const uint8_t *parent, bsq_intv *out, int32_t *n_out) {
  if (!ix || !opt_ || n < 0) return BSQ_EINVAL;
  if (n == 0) return 0;
  int rc = check_lens(n, lens, stride);
  if (rc) return rc;
  CK(cudaSetDevice(ix->device));
  bsq_devopt_t opt; memcpy(&opt, opt_, sizeof opt);
  uint8_t *dseq = nullptr, *dpar = nullptr; int32_t *dlen = nullptr, *dn = nullptr, *dnsa = nullptr, *dst = nullptr;
  bsq_pk_t *dpk = nullptr; bsq_intv_t *dint = nullptr; unsigned long long *dnext = nullptr;
  CK(cudaMalloc(&dseq, n * stride)); CK(cudaMalloc(&dpar, n)); CK(cudaMalloc(&dlen, n * 4)); CK(cudaMalloc(&dn, n * 4));
  CK(cudaMalloc(&dnsa, n * 4)); CK(cudaMalloc(&dst, 4)); CK(cudaMalloc(&dnext, 8));
  CK(cudaMalloc(&dpk, n * BSQ_MAX_INTV * sizeof(bsq_pk_t))); CK(cudaMalloc(&dint, n * BSQ_MAX_INTV * sizeof(bsq_intv_t)));
  CK(cudaMemcpy(dseq, seqs, n * stride, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dpar, parent, n, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dlen, lens, n * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(dst, 0, 4)); CK(cudaMemset(dnext, 0, 8));
  CK(cudaMemset(dpk, 0, n * BSQ_MAX_INTV * sizeof(bsq_pk_t)));
  if (seed_v1()) k_seed<<<seed_grid(n), 128, seed_smem_bytes(stride)>>>(opt, ix->d, n, dseq, stride, dlen, dpar, 0, dpk, dn, dst, dnext);
  else if (seed_impl() == 2) launch_seed2(seed_variant(), seed_grid(n), 0, opt, ix->d, n, dseq, stride, dlen, dpar, 0, dpk, dn, dst, dnext);
  else {
    SeedWs ws;
    const int rc3 = seed3_run(ws, 0, opt, ix->d, n, dseq, stride, dlen, dpar, 0, dpk, dn, nullptr);
    ws.release();
    if (rc3) { cudaFree(dseq); cudaFree(dpar); cudaFree(dlen); cudaFree(dn); cudaFree(dnsa); cudaFree(dst); cudaFree(dpk); cudaFree(dint); cudaFree(dnext); return rc3; }
  }
  CK(cudaGetLastError());
  k_seed_sort<<<nblk(n, 128), 128>>>(opt, n, dpk, dn, dnsa, dst);
  CK(cudaGetLastError());
  k_unpack_intv<<<nblk(n * BSQ_MAX_INTV, 256), 256>>>(n * BSQ_MAX_INTV, dpk, dint);
  CK(cudaGetLastError());
  int32_t st = 0;
  CK(cudaMemcpy(&st, dst, 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(n_out, dn, n * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(out, dint, n * BSQ_MAX_INTV * sizeof(bsq_intv_t), cudaMemcpyDeviceToHost));
  cudaFree(dseq); cudaFree(dpar); cudaFree(dlen); cudaFree(dn); cudaFree(dnsa); cudaFree(dst); cudaFree(dpk); cudaFree(dint); cudaFree(dnext);
  return st ? BSQ_EOVERFLOW : 0;
}

int bsq_extend_batch(const bsq_opt *opt_, int64_t n, const uint8_t *qbuf, const int64_t *qoff, const int32_t *qlen,
                     const uint8_t *tbuf, const int64_t *toff, const int32_t *tlen, const uint8_t *is_parent,
                     const int32_t *w, const int32_t *h0, int32_t *out) {
  if (!opt_ || n < 0) return BSQ_EINVAL;
  if (n == 0) return 0;
  int64_t qtot = 0, ttot = 0;
  for (int64_t j = 0; j < n; ++j) {
    if (qlen[j] < 0 || qlen[j] > BSQ_MAX_READ_LEN || tlen[j] < 0 || h0[j] <= 0) return BSQ_EINVAL;
    if (qoff[j] + qlen[j] > qtot) qtot = qoff[j] + qlen[j];
    if (toff[j] + tlen[j] > ttot) ttot = toff[j] + tlen[j];
  }
  bsq_devopt_t opt; memcpy(&opt, opt_, sizeof opt);
  uint8_t *dq, *dt, *dp; int64_t *dqo, *dto; int32_t *dql, *dtl, *dw, *dh, *dout;
  CK(cudaMalloc(&dq, qtot + 1)); CK(cudaMalloc(&dt, ttot + 1)); CK(cudaMalloc(&dp, n));
  CK(cudaMalloc(&dqo, n * 8)); CK(cudaMalloc(&dto, n * 8));
  CK(cudaMalloc(&dql, n * 4)); CK(cudaMalloc(&dtl, n * 4)); CK(cudaMalloc(&dw, n * 4)); CK(cudaMalloc(&dh, n * 4));
  CK(cudaMalloc(&dout, n * 24));
  CK(cudaMemcpy(dq, qbuf, qtot, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dt, tbuf, ttot, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dp, is_parent, n, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dqo, qoff, n * 8, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dto, toff, n * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dql, qlen, n * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dtl, tlen, n * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dw, w, n * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dh, h0, n * 4, cudaMemcpyHostToDevice));
  k_extend_warp<<<nblk(n * 32, 128), 128>>>(opt, n, dq, dqo, dql, dt, dto, dtl, dp, dw, dh, dout);
  CK(cudaGetLastError());
  CK(cudaMemcpy(out, dout, n * 24, cudaMemcpyDeviceToHost));
  cudaFree(dq); cudaFree(dt); cudaFree(dp); cudaFree(dqo); cudaFree(dto); cudaFree(dql); cudaFree(dtl); cudaFree(dw);
  cudaFree(dh); cudaFree(dout);
  return 0;
}

int bsq_aligner_create(const bsq_index *ix, const bsq_opt *opt, bsq_aligner **out) {
  if (!ix || !opt || !out) return BSQ_EINVAL;
  CK(cudaSetDevice(ix->device));
  bsq_aligner *al = new bsq_aligner();
  al->idx = ix;
  memcpy(&al->opt, opt, sizeof al->opt);
  memset(al->counters, 0, sizeof al->counters);
  CK(cudaStreamCreateWithFlags(&al->stream, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&al->stream2, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&al->stream_out, cudaStreamNonBlocking));
  CK(cudaEventCreateWithFlags(&al->ev_fork, cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&al->ev_join, cudaEventDisableTiming));
  for (int i = 0; i < 8; ++i) CK(cudaEventCreate(&al->ev[i]));
  *out = al;
  return 0;
}

void bsq_aligner_destroy(bsq_aligner *al) {
  if (!al) return;
  cudaSetDevice(al->idx->device);
  DevBuf *bufs[] = {&al->seqs, &al->lens, &al->parent, &al->intv, &al->n_intv, &al->n_sa, &al->sa_off, &al->ranks, &al->pos,
                    &al->status, &al->snodes, &al->wchains, &al->bnodes, &al->order, &al->ochains, &al->oseeds, &al->n_chains,
                    &al->frac_rep, &al->srt, &al->regs_tmp, &al->n_regs, &al->reg_off, &al->regs, &al->cub_tmp, &al->scalars, &al->fb_flag, &al->tiers, &al->spans, &al->aflag, &al->regs2, &al->reg_off2};
  for (DevBuf *b : bufs) b->release();
  al->seed_ws.release();
  for (int i = 0; i < 8; ++i) cudaEventDestroy(al->ev[i]);
  cudaStreamDestroy(al->stream);
  cudaStreamDestroy(al->stream_out);
  cudaStreamDestroy(al->stream2); cudaEventDestroy(al->ev_fork); cudaEventDestroy(al->ev_join);
  delete al;
}

void bsq_free(void *p) { free(p); }

int bsq_aligner_counters(const bsq_aligner *al, int64_t *c, int n) {
  if (!al || !c) return BSQ_EINVAL;
  for (int i = 0; i < n && i < 16; ++i) c[i] = al->counters[i];
  return 0;
}

// exclusive prefix sum of n int32 counts into n+1 int64 offsets (offsets[n] = total)
static int scan_counts(bsq_aligner *al, const int32_t *d_counts, int64_t *d_off, int64_t n, int64_t *total) {
  // shift-by-one trick: run an inclusive sum into d_off+1 and zero d_off[0]
  size_t tmp = 0;
  CK(cub::DeviceScan::InclusiveSum(nullptr, tmp, d_counts, d_off + 1, (int)n, al->stream));
  int rc = al->cub_tmp.reserve(tmp);
  if (rc) return rc;
  CK(cudaMemsetAsync(d_off, 0, 8, al->stream));
  CK(cub::DeviceScan::InclusiveSum(al->cub_tmp.p, tmp, d_counts, d_off + 1, (int)n, al->stream));
  CK(cudaMemcpyAsync(total, d_off + n, 8, cudaMemcpyDeviceToHost, al->stream));
  CK(bsq_stream_wait(al->stream));
  return 0;
}

#ifdef BSQ_INSTRUMENT
// blocks fetched so far (stream-synchronising; instrumented build only)
static void snap_blocks(bsq_aligner *al, int slot) {
  unsigned long long h[8];
  cudaStreamSynchronize(al->stream);
  cudaMemcpyFromSymbol(h, bsq_ctr, sizeof h);
  al->counters[slot] = (int64_t)h[BSQ_CTR_BLOCKS];
}
#define SNAP(slot) snap_blocks(al, slot)
#else
#define SNAP(slot) ((void)0)
#endif

// Device-resident core of phase 1: inputs already in al->seqs / lens / parent.
static int phase1_device(bsq_aligner *al, int64_t n, int32_t stride, int64_t *total_regs) {
  const bsq_devidx_t &ix = al->idx->d;
  const bsq_devopt_t &opt = al->opt;
  cudaStream_t s = al->stream;
  int rc;
#define RES(buf, bytes) if ((rc = al->buf.reserve(bytes))) return rc
  RES(intv, (size_t)n * BSQ_MAX_INTV * sizeof(bsq_pk_t)); RES(scalars, 256);
  RES(n_intv, n * 4); RES(n_sa, n * 4); RES(sa_off, (n + 1) * 8); RES(status, 4);
  RES(n_chains, n * 4); RES(frac_rep, n * 4); RES(n_regs, n * 4); RES(reg_off, (n + 1) * 8);
  CK(cudaMemsetAsync(al->status.p, 0, 4, s));
  CK(cudaMemsetAsync(al->scalars.p, 0, 256, s));
  SNAP(11);
  CK(cudaEventRecord(al->ev[0], s));
  if (seed_v1())
    k_seed<<<seed_grid(n), 128, seed_smem_bytes(stride), s>>>(opt, ix, n, al->seqs.as<uint8_t>(), stride, al->lens.as<int32_t>(), al->parent.as<uint8_t>(), 1,
                                                              al->intv.as<bsq_pk_t>(), al->n_intv.as<int32_t>(), al->status.as<int32_t>(),
                                                              al->scalars.as<unsigned long long>());
  else if (seed_impl() == 2)
    launch_seed2(seed_variant(), seed_grid(n), s, opt, ix, n, al->seqs.as<uint8_t>(), stride, al->lens.as<int32_t>(), al->parent.as<uint8_t>(), 1,
                 al->intv.as<bsq_pk_t>(), al->n_intv.as<int32_t>(), al->status.as<int32_t>(), al->scalars.as<unsigned long long>());
  else if ((rc = seed3_run(al->seed_ws, s, opt, ix, n, al->seqs.as<uint8_t>(), stride, al->lens.as<int32_t>(), al->parent.as<uint8_t>(), 1,
                           al->intv.as<bsq_pk_t>(), al->n_intv.as<int32_t>(), al->seed_fills)))
    return rc;
  CK(cudaGetLastError());
  k_seed_sort<<<nblk(n, 128), 128, 0, s>>>(opt, n, al->intv.as<bsq_pk_t>(), al->n_intv.as<int32_t>(), al->n_sa.as<int32_t>(), al->status.as<int32_t>());
  CK(cudaGetLastError());
  CK(cudaEventRecord(al->ev[1], s));
  SNAP(12);
  int64_t total_sa = 0;
  if ((rc = scan_counts(al, al->n_sa.as<int32_t>(), al->sa_off.as<int64_t>(), n, &total_sa))) return rc;
  const int64_t pool = total_sa + n * BSQ_TAIL_SLACK;
  RES(ranks, (total_sa + 1) * 8); RES(pos, (total_sa + 1) * 8);
  // workspace of the exact fallback chaining: small by default (bump-allocated to the few flagged tasks)
  int64_t fb_cap = pool / 16 + 65536;
  { const char *e = getenv("BSQ_FB_POOL"); if (e && atoll(e) > 0) fb_cap = atoll(e); }  // test hook: force the retry path
  if (al->fb_cap > fb_cap) fb_cap = al->fb_cap;  // keep a pool that was grown earlier
  al->fb_cap = fb_cap;
  RES(snodes, fb_cap * sizeof(bsq_snode_t)); RES(wchains, fb_cap * sizeof(bsq_wchain_t));
  RES(bnodes, fb_cap * sizeof(bsq_bnode_t)); RES(order, fb_cap * 4);
  RES(ochains, pool * sizeof(bsq_chain_t)); RES(oseeds, pool * sizeof(bsq_seed_t));
  RES(srt, pool * 8); RES(regs_tmp, pool * sizeof(bsq_reg_t)); RES(spans, pool * 16); RES(aflag, pool);
  CK(cudaEventRecord(al->ev[2], s));
  k_expand<<<nblk(n, 128), 128, 0, s>>>(opt, n, al->intv.as<bsq_pk_t>(), al->n_intv.as<int32_t>(), al->parent.as<uint8_t>(),
                                         al->sa_off.as<int64_t>(), al->ranks.as<uint64_t>());
  CK(cudaGetLastError());
  if (total_sa > 0) {
    k_sa<<<nblk(total_sa, 256), 256, 0, s>>>(ix, total_sa, al->ranks.as<uint64_t>(), al->pos.as<uint64_t>());
    CK(cudaGetLastError());
  }
  CK(cudaEventRecord(al->ev[3], s));
  SNAP(13);
  RES(fb_flag, n);
  {
    RES(tiers, (size_t)n * 4 * BSQ_N_TIERS);
    unsigned long long *tcnt = al->scalars.as<unsigned long long>() + 8;
    k_chain_tiers<<<nblk(n, 256), 256, 0, s>>>(n, al->n_sa.as<int32_t>(), al->tiers.as<int32_t>(), tcnt);
    CK(cudaGetLastError());
#define CW_LAUNCH(CAP, WPB, Q, STREAM) if ((rc = launch_chain_warp<CAP, WPB>(STREAM, opt, ix, n, al->tiers.as<int32_t>() + (size_t)(Q) * n, tcnt + (Q), al->lens.as<int32_t>(), al->parent.as<uint8_t>(),  \
                al->intv.as<bsq_pk_t>(), al->n_intv.as<int32_t>(), al->n_sa.as<int32_t>(), al->sa_off.as<int64_t>(), al->pos.as<uint64_t>(), \
                al->ochains.as<bsq_chain_t>(), al->oseeds.as<bsq_seed_t>(), al->n_chains.as<int32_t>(), al->frac_rep.as<float>(),           \
                al->fb_flag.as<uint8_t>(), al->scalars.as<unsigned long long>() + 1, al->scalars.as<unsigned long long>() + 16 + (Q)))) return rc
    CK(cudaEventRecord(al->ev_fork, s));
    CK(cudaStreamWaitEvent(al->stream2, al->ev_fork, 0));
    // two streams: the kernels are persistent (one wave each), so the CTAs of one tier fill the SMs that the tail of
    // another leaves idle; the few large tasks (1024-seed tier) start first
    CW_LAUNCH(1024, 2, 6, al->stream2);
    CW_LAUNCH(256, 4, 5, s);
    CW_LAUNCH(192, 4, 4, al->stream2);
    CW_LAUNCH(160, 4, 3, s);
    CW_LAUNCH(128, 4, 2, al->stream2);
    CW_LAUNCH(96, 4, 1, s);
    CW_LAUNCH(64, 4, 0, al->stream2);
    CK(cudaEventRecord(al->ev_join, al->stream2));
    CK(cudaStreamWaitEvent(s, al->ev_join, 0));
#undef CW_LAUNCH
    CK(cudaEventRecord(al->ev[7], s));
  }
  for (int pass = 0; pass < 2; ++pass) {
    unsigned long long *fb_cursor = al->scalars.as<unsigned long long>() + 6;
    k_chain<<<nblk(n, 128), 128, 0, s>>>(opt, ix, n, al->lens.as<int32_t>(), al->parent.as<uint8_t>(), al->intv.as<bsq_pk_t>(),
                                          al->n_intv.as<int32_t>(), al->sa_off.as<int64_t>(), al->pos.as<uint64_t>(),
                                          al->snodes.as<bsq_snode_t>(), al->wchains.as<bsq_wchain_t>(), al->bnodes.as<bsq_bnode_t>(),
                                          al->order.as<int32_t>(), fb_cap, fb_cursor, al->ochains.as<bsq_chain_t>(),
                                          al->oseeds.as<bsq_seed_t>(), al->n_chains.as<int32_t>(), al->frac_rep.as<float>(),
                                          al->status.as<int32_t>(), al->fb_flag.as<uint8_t>());
    CK(cudaGetLastError());
    int32_t st_now = 0;
    CK(cudaMemcpyAsync(&st_now, al->status.p, 4, cudaMemcpyDeviceToHost, s));
    CK(bsq_stream_wait(s));
    if (!(st_now & 4)) break;
    if (pass == 1) { snprintf(g_err, sizeof g_err, "fallback chaining workspace exhausted twice"); return BSQ_EOVERFLOW; }
    // many flagged tasks (e.g. repeats with more than max_occ occurrences): give the fallback the full-size pool and redo the rest
    fb_cap = al->fb_cap = pool + 2 * n;
    RES(snodes, fb_cap * sizeof(bsq_snode_t)); RES(wchains, fb_cap * sizeof(bsq_wchain_t));
    RES(bnodes, fb_cap * sizeof(bsq_bnode_t)); RES(order, fb_cap * 4);
    st_now &= ~4;
    CK(cudaMemcpyAsync(al->status.p, &st_now, 4, cudaMemcpyHostToDevice, s));
    CK(cudaMemsetAsync(fb_cursor, 0, 8, s));
    CK(bsq_stream_wait(s));
  }
  CK(cudaEventRecord(al->ev[4], s));
  {
    const unsigned want = nblk(n * 32, 128), wave = (unsigned)(bsq_sm_count() * BSQ_REGION_CTAS);
    k_region_prep<<<nblk(n * 32, 128), 128, 0, s>>>(opt, ix, n, al->seqs.as<uint8_t>(), stride, al->lens.as<int32_t>(), al->sa_off.as<int64_t>(),
                                                    al->ochains.as<bsq_chain_t>(), al->oseeds.as<bsq_seed_t>(), al->n_chains.as<int32_t>(),
                                                    al->srt.as<uint64_t>(), al->spans.as<int64_t>(), al->aflag.as<uint8_t>());
    k_region<<<want < wave ? want : wave, 128, 0, s>>>(opt, ix, n, al->seqs.as<uint8_t>(), stride, al->lens.as<int32_t>(), al->parent.as<uint8_t>(),
                                                       al->sa_off.as<int64_t>(), al->ochains.as<bsq_chain_t>(), al->oseeds.as<bsq_seed_t>(),
                                                       al->n_chains.as<int32_t>(), al->frac_rep.as<float>(), al->srt.as<uint64_t>(), al->spans.as<int64_t>(), al->aflag.as<uint8_t>(),
                                                       al->regs_tmp.as<bsq_reg_t>(), al->n_regs.as<int32_t>(), al->scalars.as<unsigned long long>() + 24);
  }
  CK(cudaGetLastError());
  CK(cudaEventRecord(al->ev[5], s));
  DevBuf &out_regs = al->out_slot ? al->regs2 : al->regs, &out_off = al->out_slot ? al->reg_off2 : al->reg_off;  // this run's result slot
  if ((rc = out_off.reserve((n + 1) * 8))) return rc;
  if ((rc = scan_counts(al, al->n_regs.as<int32_t>(), out_off.as<int64_t>(), n, total_regs))) return rc;
  if ((rc = out_regs.reserve((*total_regs + 1) * sizeof(bsq_reg_t)))) return rc;
  k_compact_regs<<<nblk(n, 128), 128, 0, s>>>(n, al->sa_off.as<int64_t>(), al->n_regs.as<int32_t>(), out_off.as<int64_t>(),
                                               al->regs_tmp.as<bsq_reg_t>(), out_regs.as<bsq_reg_t>());
  CK(cudaGetLastError());
  CK(cudaEventRecord(al->ev[6], s));
  int32_t st = 0;
  unsigned long long n_fb = 0;
  CK(cudaMemcpyAsync(&st, al->status.p, 4, cudaMemcpyDeviceToHost, s));
  CK(cudaMemcpyAsync(&n_fb, al->scalars.as<unsigned long long>() + 1, 8, cudaMemcpyDeviceToHost, s));
  CK(bsq_stream_wait(s));
  al->counters[14] = (int64_t)n_fb;  // tasks chained by the exact fallback kernel
  { float w_ms = 0; cudaEventElapsedTime(&w_ms, al->ev[3], al->ev[7]); al->counters[15] = (int64_t)(w_ms * 1000); }
#undef RES
  // bit 0: a (read, conversion) task produced more than BSQ_MAX_INTV SMEM intervals.  That task is left without seeds
  // (n_intv = 0, so the read may come out unaligned for that conversion) and the batch goes on; the caller reads
  // counters[3] and warns -- one pathological read must not take the whole run down.  The other bits are real errors.
  al->counters[3] = st & 1;
  if (st & ~1) { snprintf(g_err, sizeof g_err, "device status 0x%x (2: chain workspace, 4: fallback pool)", st); return BSQ_EOVERFLOW; }
  float ms[6];
  for (int i = 0; i < 6; ++i) cudaEventElapsedTime(&ms[i], al->ev[i], al->ev[i + 1]);
  { float tot; cudaEventElapsedTime(&tot, al->ev[0], al->ev[6]); al->counters[10] = (int64_t)(tot * 1000); }
  al->counters[0] = n; al->counters[2] = total_sa; al->counters[4] = *total_regs;
  al->counters[1] = (ix.fm[0].full_sa != nullptr) + (ix.fm[1].full_sa != nullptr);
  al->counters[5] = (int64_t)(ms[0] * 1000); al->counters[6] = (int64_t)(ms[2] * 1000);
  al->counters[7] = (int64_t)(ms[3] * 1000); al->counters[8] = (int64_t)(ms[4] * 1000);
  al->counters[9] = (int64_t)((ms[1] + ms[5]) * 1000);
  return 0;
}

int bsq_aligner_stage(bsq_aligner *al, int64_t n, const uint8_t *seqs, int32_t stride, const int32_t *lens, const uint8_t *parent) {
  if (!al || n < 0) return BSQ_EINVAL;
  al->n_staged = 0;
  if (n == 0) return 0;
  int rc = check_lens(n, lens, stride);
  if (rc) return rc;
  CK(cudaSetDevice(al->idx->device));
  if ((rc = al->seqs.reserve((size_t)n * stride))) return rc;
  if ((rc = al->lens.reserve(n * 4))) return rc;
  if ((rc = al->parent.reserve(n))) return rc;
  cudaStream_t s = al->stream;
  CK(cudaMemcpyAsync(al->seqs.p, seqs, (size_t)n * stride, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(al->lens.p, lens, n * 4, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(al->parent.p, parent, n, cudaMemcpyHostToDevice, s));
  al->n_staged = n; al->stride = stride; al->n_regs_total = -1;
  return 0;
}

int bsq_aligner_run(bsq_aligner *al, int64_t *n_regs) {
  if (!al) return BSQ_EINVAL;
  if (al->n_staged == 0) { if (n_regs) *n_regs = 0; al->n_regs_total = 0; return 0; }
  CK(cudaSetDevice(al->idx->device));
  int64_t total = 0;
  {
    std::unique_lock<std::mutex> lk(al->slot_mu);
    al->slot_cv.wait(lk, [&] { return !al->claimed[al->out_slot ^ 1]; });  // its previous content is still to be fetched by another thread
    al->out_slot ^= 1;
    al->out_regs[al->out_slot] = -1;
  }
  int rc = phase1_device(al, al->n_staged, al->stride, &total);
  if (rc) return rc;
  al->n_regs_total = total;
  al->out_tasks[al->out_slot] = al->n_staged; al->out_regs[al->out_slot] = total;
  if (n_regs) *n_regs = total;
  return 0;
}

int bsq_aligner_result_slot(bsq_aligner *al, int *slot, int64_t *n_tasks, int64_t *n_regs) {
  if (!al || !slot || al->n_regs_total < 0) return BSQ_EINVAL;
  {
    std::lock_guard<std::mutex> lk(al->slot_mu);
    al->claimed[al->out_slot] = true;
  }
  *slot = al->out_slot;
  if (n_tasks) *n_tasks = al->n_staged;
  if (n_regs) *n_regs = al->n_regs_total;
  return 0;
}

int bsq_aligner_release_slot(bsq_aligner *al, int slot) {
  if (!al || slot < 0 || slot > 1) return BSQ_EINVAL;
  {
    std::lock_guard<std::mutex> lk(al->slot_mu);
    al->claimed[slot] = false;
  }
  al->slot_cv.notify_all();
  return 0;
}

static int fetch_slot_copy(bsq_aligner *al, int slot, bsq_reg *regs, int64_t *reg_off);
int bsq_aligner_fetch_slot(bsq_aligner *al, int slot, bsq_reg *regs, int64_t *reg_off) {
  if (!al || slot < 0 || slot > 1 || al->out_regs[slot] < 0 || !reg_off) return BSQ_EINVAL;
  const int rc = fetch_slot_copy(al, slot, regs, reg_off);
  bsq_aligner_release_slot(al, slot);
  return rc;
}

static int fetch_slot_copy(bsq_aligner *al, int slot, bsq_reg *regs, int64_t *reg_off) {
  const int64_t n = al->out_tasks[slot], nr = al->out_regs[slot];
  if (n == 0) { reg_off[0] = 0; return 0; }
  CK(cudaSetDevice(al->idx->device));
  // the run that filled the slot has completed (bsq_aligner_run returns after its stream is idle): plain copies on the
  // result stream, beside whatever the aligner's own streams are doing for the next batch
  cudaStream_t s = al->stream_out;
  const DevBuf &r = slot ? al->regs2 : al->regs, &o = slot ? al->reg_off2 : al->reg_off;
  if (nr > 0) {
    if (!regs) return BSQ_EINVAL;
    CK(cudaMemcpyAsync(regs, r.p, (size_t)nr * sizeof(bsq_reg), cudaMemcpyDeviceToHost, s));
  }
  CK(cudaMemcpyAsync(reg_off, o.p, (n + 1) * 8, cudaMemcpyDeviceToHost, s));
  CK(bsq_stream_wait(s));
  return 0;
}

int bsq_aligner_fetch(bsq_aligner *al, bsq_reg *regs, int64_t *reg_off) {
  if (!al || al->n_regs_total < 0 || !reg_off) return BSQ_EINVAL;
  const int64_t n = al->n_staged;
  if (n == 0) { reg_off[0] = 0; return 0; }
  CK(cudaSetDevice(al->idx->device));
  cudaStream_t s = al->stream;
  const DevBuf &r = al->out_slot ? al->regs2 : al->regs, &o = al->out_slot ? al->reg_off2 : al->reg_off;
  if (al->n_regs_total > 0) {
    if (!regs) return BSQ_EINVAL;
    CK(cudaMemcpyAsync(regs, r.p, (size_t)al->n_regs_total * sizeof(bsq_reg), cudaMemcpyDeviceToHost, s));
  }
  CK(cudaMemcpyAsync(reg_off, o.p, (n + 1) * 8, cudaMemcpyDeviceToHost, s));
  CK(bsq_stream_wait(s));
  return 0;
}

int bsq_align_phase1(bsq_aligner *al, int64_t n, const uint8_t *seqs, int32_t stride, const int32_t *lens, const uint8_t *parent,
                     bsq_reg **regs_out, int64_t *reg_off) {
  if (!al || n < 0 || !regs_out || !reg_off) return BSQ_EINVAL;
  *regs_out = nullptr;
  if (n == 0) { reg_off[0] = 0; return 0; }
  int rc = bsq_aligner_stage(al, n, seqs, stride, lens, parent);
  if (rc) return rc;
  int64_t total = 0;
  if ((rc = bsq_aligner_run(al, &total))) return rc;
  bsq_reg *host = (bsq_reg *)malloc((size_t)(total + 1) * sizeof(bsq_reg));
  if (!host) return BSQ_ENOMEM;
  if ((rc = bsq_aligner_fetch(al, host, reg_off))) { free(host); return rc; }
  *regs_out = host;
  return 0;
}

// Work counters (see BSQ_CTR in bsq_common.h).  The product build is not instrumented and says so.
int bsq_work_counters(uint64_t *out, int n, int reset) {
#ifdef BSQ_INSTRUMENT
  unsigned long long h[8];
  CK(cudaMemcpyFromSymbol(h, bsq_ctr, sizeof h));
  for (int i = 0; i < n && i < 8; ++i) out[i] = h[i];
  if (reset) { memset(h, 0, sizeof h); CK(cudaMemcpyToSymbol(bsq_ctr, h, sizeof h)); }
  return 0;
#else
  (void)out; (void)n; (void)reset;
  snprintf(g_err, sizeof g_err, "libbsq.so is not instrumented; use libbsq_count.so");
  return BSQ_EINVAL;
#endif
}

__global__ void __launch_bounds__(128) k_gather_probe(const uint32_t *buf, uint64_t n_units, int iters, uint32_t *out) {
  uint64_t s = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 0x9E3779B97F4A7C15ull + 12345;
  uint32_t acc = 0;
  for (int it = 0; it < iters; ++it) {
    uint32_t w0[8], w1[8];
    s = s * 6364136223846793005ull + 1442695040888963407ull;
    s3_ld256(buf + ((s >> 20) % n_units) * 8, w0);
    s = s * 6364136223846793005ull + 1442695040888963407ull;
    s3_ld256(buf + ((s >> 20) % n_units) * 8, w1);
    acc ^= w0[0] + w0[7] + w1[0] + w1[7];
    s += acc & 1;  // the next addresses depend on the data, like an FM-index walk
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

int bsq_measure_gather(int device, uint64_t bytes, double *gathers_per_s) {
  if (!gathers_per_s || bytes < (1u << 20)) return BSQ_EINVAL;
  CK(cudaSetDevice(device));
  uint32_t *buf = nullptr, *out = nullptr;
  const int grid = bsq_sm_count() * 8, iters = 1000;
  CK(cudaMalloc(&buf, bytes));
  CK(cudaMalloc(&out, (size_t)grid * 128 * 4));
  CK(cudaMemset(buf, 1, bytes));
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  k_gather_probe<<<grid, 128>>>(buf, bytes / 32, iters / 10, out);
  CK(cudaEventRecord(a));
  k_gather_probe<<<grid, 128>>>(buf, bytes / 32, iters, out);
  CK(cudaEventRecord(b));
  CK(cudaEventSynchronize(b));
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, a, b));
  *gathers_per_s = (double)grid * 128 * iters * 2 / (ms * 1e-3);
  cudaEventDestroy(a); cudaEventDestroy(b); cudaFree(buf); cudaFree(out);
  return 0;
}

int bsq_host_alloc(void **p, size_t bytes) {
  if (!p) return BSQ_EINVAL;
  CK(cudaHostAlloc(p, bytes ? bytes : 1, cudaHostAllocDefault));
  return 0;
}

void bsq_host_free(void *p) { if (p) cudaFreeHost(p); }

}  // extern "C"
