// Shared definitions for the device code of the B200 aligner/pileup kernels.
//
// Everything in the bsq_*.h headers is written as plain per-task C++ marked BSQ_HD so the very
// same source is (a) compiled by nvcc into the sm_100a kernels in bsq_kernels.cu and (b) compiled
// by g++ into tests/hostemu (a *test-only* harness that lets the device logic be checked against
// the oracle on a machine without a GPU).  The shipped library (libbsq.so) contains no host
// execution path for these functions.
#pragma once
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define BSQ_HD __host__ __device__ __forceinline__
#define BSQ_HDN __host__ __device__ __noinline__
#else
#define BSQ_HD inline
#define BSQ_HDN inline
#endif

// Work counters for the roofline arithmetic (bench.py).  Only the separately built
// libbsq_count.so (-DBSQ_INSTRUMENT) carries them; in the product build the macro is empty.
#if defined(BSQ_INSTRUMENT) && defined(__CUDA_ARCH__)
#define BSQ_CTR(i, v) atomicAdd(&bsq_ctr[i], (unsigned long long)(v))
#else
#define BSQ_CTR(i, v) ((void)0)
#endif
#define BSQ_CTR_BLOCKS 0   // 64-byte FM-index blocks fetched
#define BSQ_CTR_EXTENDS 1  // bwt_extend calls
#define BSQ_CTR_KSW 2      // ksw_extend2 calls
#define BSQ_CTR_CELLS 3    // DP cells
#define BSQ_CTR_REFB 4     // reference bases decoded from the 2-bit pac

#define BSQ_MAX_READ_LEN 256  // longest read the device seeding kernels accept
#define BSQ_MAX_INTV 384      // per (read,conversion) capacity of the SMEM interval list (a GRCh38-sized 3-letter index yields >100 SMEMs per read)

// Same fields and meaning as the reference's bwtintv_t (lib/aln/bwt.h:80-82):
// x[0] = interval start in the searched index, x[1] = start in the complementary index,
// x[2] = interval size, info = qbeg<<32 | qend.
struct bsq_intv_t {
  uint64_t x[3], info;
};

// mem_seed_t (lib/aln/memchain.h:70-74).  The reference's `score` field always equals `len`
// on this path: mem_chain sets score = len (memchain.c:336) and mem_flt_chained_seeds, the only
// writer, returns early for reads shorter than ~700 bp (memchain.c:544-546).
struct bsq_seed_t {
  int64_t rbeg;
  int32_t qbeg, len;
};

// A chain after mem_chain + mem_chain_flt (mem_chain_t, lib/aln/memchain.h:78-88).  Its seeds
// are seeds[seed_off .. seed_off+n_seeds) followed by n_extra backup seeds, in the task's pool.
struct bsq_chain_t {
  int64_t pos;
  int32_t rid, w, first;
  int32_t seed_off, n_seeds, n_extra;
  uint8_t kept, is_alt, pad_[6];
};

// The subset of mem_alnreg_t (lib/aln/mem_alnreg.h:34-66) that phase 1 (seed->chain->extend)
// fills in; the host adds the phase-2 fields.
struct bsq_reg_t {
  int64_t rb, re;
  int32_t qb, qe;
  int32_t rid, score, truesc, w;
  int32_t seedcov, seedlen0;
  float frac_rep;
  uint8_t bss, parent, pad_[2];
};

// Options that reach the device (mem_opt_t, lib/aln/bwamem.h:54-124)
struct bsq_devopt_t {
  int32_t a, b, o_del, e_del, o_ins, e_ins, pen_clip5, pen_clip3, w, zdrop;
  int32_t min_seed_len, split_width, max_occ, max_chain_gap, min_chain_weight, max_chain_extend;
  int32_t max_mem_intv, split_len, self_ovlp, bsstrand;
  float mask_level, drop_ratio;
  int8_t ctmat[25], gamat[25];
  int8_t pad_[2];
};

// One FM-index (bwt_t, lib/aln/bwt.h:54-71) as laid out in HBM: `blocks` is the body of the
// reference's .bwt file verbatim (64-byte blocks: u64 occ[4] then 128 2-bit symbols in 8 u32,
// MSB first; lib/aln/bwt.h:93-101, bwtindex.c:130-154), 64-byte aligned.
struct bsq_fm_t {
  const uint32_t *blocks;
  const uint64_t *sa;  // sampled SA, sa[0] = (u64)-1 (bwt.c:84,450)
  const uint64_t *full_sa;  // optional: SA of every rank, derived in HBM (180 GB make room for 2 x 8 B x 6.2 G);
                            // replaces the ~31-step LF walk of bwt_sa by one gather; same values by construction
  const uint32_t *b32;      // derived in HBM: the same ranks as 32-byte blocks (3 x 40-bit cumulative counts + 64 symbols,
                            // one DRAM sector per lookup), used by the seeding kernels (bsq_seed3.cuh)
  uint64_t primary, seq_len;
  uint64_t L2[5];
  int32_t sa_intv, pad_;
};

// Device-resident index: bwaidx_t (lib/aln/bwa.h:42-50) + bntseq_t (lib/aln/bntseq.h:56-64)
struct bsq_devidx_t {
  bsq_fm_t fm[2];      // [0] = daughter (G>A), [1] = parent (C>T)  (bwa.c:535-536)
  const uint8_t *pac;  // forward-only 2-bit packed reference (.bis.pac)
  int64_t l_pac;
  const int64_t *ann_offset;  // n_seqs
  const int32_t *ann_len;     // n_seqs
  const int32_t *ann_is_alt;  // n_seqs
  int32_t n_seqs, pad_;
};

template <typename T>
BSQ_HD T bsq_min(T a, T b) { return a < b ? a : b; }
template <typename T>
BSQ_HD T bsq_max(T a, T b) { return a > b ? a : b; }
