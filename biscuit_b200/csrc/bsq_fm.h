// FM-index primitives: rank (occ), bidirectional extension, sampled-SA lookup.
// Semantics follow the reference's lib/aln/bwt.c; the arithmetic is re-derived with popcounts
// instead of the reference's 8-bit LUT (bwt.c:167-169) -- the counts are integers, so equal.
#pragma once
#include "bsq_common.h"

struct bsq_block_t {
  uint32_t w[16];  // w[0..7] = u64 occ[4] (little endian pairs), w[8..15] = 128 symbols
};

BSQ_HD void bsq_load_block(const uint32_t *blocks, uint64_t blk, bsq_block_t &b) {
  BSQ_CTR(BSQ_CTR_BLOCKS, 1);
#if defined(__CUDA_ARCH__) && defined(BSQ_LDG256)
  // Experiment for the next round (not built by default, not yet measured): one 64-byte block as two 256-bit loads
  // (LDG.E.256 on sm_100a) instead of four 128-bit ones.  k_seed2 slowed down by 20 % when two to four LSU
  // instructions per step were added (profiles/README.md, r01 v7), so halving the block-fetch instructions may pay.
  // Build: nvcc ... -DBSQ_LDG256 -o libbsq_ldg256.so; compare with tools/kbench.py libbsq.so,libbsq_ldg256.so.
  const uint32_t *p = blocks + blk * 16;
  asm("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(b.w[0]), "=r"(b.w[1]), "=r"(b.w[2]), "=r"(b.w[3]), "=r"(b.w[4]), "=r"(b.w[5]), "=r"(b.w[6]), "=r"(b.w[7]) : "l"(p));
  asm("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(b.w[8]), "=r"(b.w[9]), "=r"(b.w[10]), "=r"(b.w[11]), "=r"(b.w[12]), "=r"(b.w[13]), "=r"(b.w[14]), "=r"(b.w[15]) : "l"(p + 8));
#elif defined(__CUDA_ARCH__)
  const uint4 *p = reinterpret_cast<const uint4 *>(blocks) + blk * 4;
  uint4 a0 = __ldg(p), a1 = __ldg(p + 1), a2 = __ldg(p + 2), a3 = __ldg(p + 3);
  b.w[0] = a0.x; b.w[1] = a0.y; b.w[2] = a0.z; b.w[3] = a0.w;
  b.w[4] = a1.x; b.w[5] = a1.y; b.w[6] = a1.z; b.w[7] = a1.w;
  b.w[8] = a2.x; b.w[9] = a2.y; b.w[10] = a2.z; b.w[11] = a2.w;
  b.w[12] = a3.x; b.w[13] = a3.y; b.w[14] = a3.z; b.w[15] = a3.w;
#else
  memcpy(b.w, blocks + blk * 16, 64);
#endif
}

BSQ_HD int bsq_popc64(uint64_t v) {
#if defined(__CUDA_ARCH__)
  return __popcll(v);
#else
  return __builtin_popcountll(v);
#endif
}

BSQ_HD uint64_t bsq_block_occ(const bsq_block_t &b, int c) { return (uint64_t)b.w[2 * c] | (uint64_t)b.w[2 * c + 1] << 32; }

// Number of A,C,G,T among the first `n` (0..128) symbols of the block, packed as four counts.
BSQ_HD void bsq_block_count4(const bsq_block_t &b, int n, uint32_t c[4]) {
  uint32_t c1 = 0, c2 = 0, c3 = 0;
  // All four words, no early exit: the lanes of a warp count different prefixes, a data-dependent trip count would
  // serialise them.  Word i contributes its first clamp(n - 32 i, 0, 32) symbols; masked-out symbols read as A and are
  // not counted in c1..c3 (c[0] follows from n).
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int m = n - 32 * i;
    m = m < 0 ? 0 : (m > 32 ? 32 : m);
    uint64_t v = (uint64_t)b.w[8 + 2 * i] << 32 | b.w[9 + 2 * i];  // 32 symbols, first symbol in the top bits
    const uint64_t keep = m == 0 ? 0ull : ~0ull << ((32 - m) << 1);  // top 2m bits
    v &= keep;
    uint64_t lo = v & 0x5555555555555555ull, hi = (v >> 1) & 0x5555555555555555ull;
    c3 += bsq_popc64(hi & lo);
    c2 += bsq_popc64(hi & ~lo);
    c1 += bsq_popc64(lo & ~hi);
  }
  c[0] = (uint32_t)n - c1 - c2 - c3; c[1] = c1; c[2] = c2; c[3] = c3;
}

// bwt_occ4 (bwt.c:173-200): ranks of the four symbols up to and including BWT position k.
BSQ_HD void bsq_occ4(const bsq_fm_t &fm, uint64_t k, uint64_t cnt[4]) {
  if (k == (uint64_t)-1) { cnt[0] = cnt[1] = cnt[2] = cnt[3] = 0; return; }
  k -= (k >= fm.primary);  // '$' is not stored
  bsq_block_t b;
  bsq_load_block(fm.blocks, k >> 7, b);
  uint32_t c[4];
  bsq_block_count4(b, (int)(k & 127) + 1, c);
#pragma unroll
  for (int i = 0; i < 4; ++i) cnt[i] = bsq_block_occ(b, i) + c[i];
}

// bwt_2occ4 (bwt.c:204-236): one block fetch when k and l fall in the same 128-symbol block.
BSQ_HD void bsq_2occ4(const bsq_fm_t &fm, uint64_t k, uint64_t l, uint64_t ck[4], uint64_t cl[4]) {
  uint64_t k2 = k - (k >= fm.primary), l2 = l - (l >= fm.primary);
  if ((l2 >> 7) != (k2 >> 7) || k == (uint64_t)-1 || l == (uint64_t)-1) {
    bsq_occ4(fm, k, ck);
    bsq_occ4(fm, l, cl);
    return;
  }
  bsq_block_t b;
  bsq_load_block(fm.blocks, k2 >> 7, b);
  uint32_t a[4], c[4];
  bsq_block_count4(b, (int)(k2 & 127) + 1, a);
  bsq_block_count4(b, (int)(l2 & 127) + 1, c);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint64_t o = bsq_block_occ(b, i);
    ck[i] = o + a[i];
    cl[i] = o + c[i];
  }
}

// bwt_2occ4 for k, l != -1 with one control path: both block fetches are issued back to back (the second is
// skipped, not branched around, when l falls in k's block), so lanes of a warp whose intervals are wide and
// lanes whose intervals are narrow wait for DRAM together instead of one group after the other.
BSQ_HD void bsq_2occ4_flat(const bsq_fm_t &fm, uint64_t k, uint64_t l, uint64_t ck[4], uint64_t cl[4]) {
  const uint64_t k2 = k - (k >= fm.primary), l2 = l - (l >= fm.primary);
  const uint64_t kb = k2 >> 7, lb = l2 >> 7;
  bsq_block_t bk, bl;
  bsq_load_block(fm.blocks, kb, bk);
  if (lb != kb) bsq_load_block(fm.blocks, lb, bl);
  else bl = bk;
  uint32_t a[4], c[4];
  bsq_block_count4(bk, (int)(k2 & 127) + 1, a);
  bsq_block_count4(bl, (int)(l2 & 127) + 1, c);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    ck[i] = bsq_block_occ(bk, i) + a[i];
    cl[i] = bsq_block_occ(bl, i) + c[i];
  }
}

// bwt_extend (bwt.c:278-293).  BACK=1 extends to the left in `fm`; BACK=0 is the forward
// extension, performed as a backward step in the complementary index.
template <int BACK>
BSQ_HD void bsq_extend(const bsq_fm_t &fm, const bsq_intv_t &ik, bsq_intv_t ok[4]) {
  uint64_t tk[4], tl[4];
  const uint64_t beg = ik.x[!BACK];
  BSQ_CTR(BSQ_CTR_EXTENDS, 1);
  bsq_2occ4(fm, beg - 1, beg - 1 + ik.x[2], tk, tl);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    ok[i].x[!BACK] = fm.L2[i] + 1 + tk[i];
    ok[i].x[2] = tl[i] - tk[i];
  }
  ok[3].x[BACK] = ik.x[BACK] + (beg <= fm.primary && beg + ik.x[2] - 1 >= fm.primary);
  ok[2].x[BACK] = ok[3].x[BACK] + ok[3].x[2];
  ok[1].x[BACK] = ok[2].x[BACK] + ok[2].x[2];
  ok[0].x[BACK] = ok[1].x[BACK] + ok[1].x[2];
}

// bwt_set_intv (bwt.h:105)
BSQ_HD void bsq_set_intv(const bsq_fm_t &fm, const bsq_fm_t &fmc, int c, bsq_intv_t &ik) {
  ik.x[0] = fm.L2[c] + 1;
  ik.x[2] = fm.L2[c + 1] - fm.L2[c];
  ik.x[1] = fmc.L2[3 - c] + 1;
  ik.info = 0;
}

// One LF step: bwt_invPsi (bwt.c:54-60) with bwt_occ (bwt.c:108-130) folded in -- the symbol
// and its rank come from the same 64-byte block.
BSQ_HD uint64_t bsq_inv_psi(const bsq_fm_t &fm, uint64_t k) {
  if (k == fm.primary) return 0;
  uint64_t x = k - (k > fm.primary);
  bsq_block_t b;
  bsq_load_block(fm.blocks, x >> 7, b);
  int off = (int)(x & 127);
  int c = (b.w[8 + (off >> 4)] >> ((~off & 15) << 1)) & 3;
  if (k == fm.seq_len) return fm.L2[c] + (fm.L2[c + 1] - fm.L2[c]);
  // count symbol c among the first off+1 symbols
  uint32_t n = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int m = off + 1 - 32 * i;
    if (m <= 0) break;
    uint64_t v = (uint64_t)b.w[8 + 2 * i] << 32 | b.w[9 + 2 * i];
    uint64_t lo = (c & 1) ? v : ~v, hi = (c & 2) ? (v >> 1) : ~(v >> 1);
    uint64_t mk = lo & hi & 0x5555555555555555ull;
    if (m < 32) mk &= ~((1ull << ((32 - m) << 1)) - 1);
    n += bsq_popc64(mk);
  }
  return fm.L2[c] + bsq_block_occ(b, c) + n;
}

// bwt_sa (bwt.c:87-97): walk LF until a sampled rank is reached.
BSQ_HD uint64_t bsq_sa(const bsq_fm_t &fm, uint64_t k) {
  if (fm.full_sa) return fm.full_sa[k];
  uint64_t steps = 0, mask = (uint64_t)fm.sa_intv - 1;
  while (k & mask) {
    ++steps;
    k = bsq_inv_psi(fm, k);
  }
  return steps + fm.sa[k / fm.sa_intv];
}
