// Internals shared by the translation units of libbsq.so (not part of the ABI).
#pragma once
#include <stdarg.h>
#include <stdint.h>
#include "bsq_common.h"

struct bsq_index {
  bsq_devidx_t d;  // device pointers
  int device;
  void *allocs[32];
  int n_allocs;
  uint64_t bwt_words[2], n_sa[2];
  int64_t build_stats[4];  // GPU index build: chunks, refinement passes, largest chunk
};

void bsq_set_error(const char *fmt, ...);
bsq_index *bsq_index_alloc(int device);
void bsq_index_adopt(bsq_index *ix, void *dev_ptr);  // freed by bsq_index_free
// full suffix array in HBM?  BSQ_FULL_SA=0/1 forces it; default: when `halves_left` arrays of (n+1) x 8 B fit with 40 GB to spare
bool bsq_want_full_sa(uint64_t n, int halves_left);
// 32-byte rank blocks of both halves (bsq_fm_t::b32), derived on the device from the reference-layout blocks
int bsq_index_derive_b32(bsq_index *ix);

// Host-side wait for a stream.  Default: cudaStreamSynchronize (spins).  BSQ_SPIN_WAIT=0 waits on an event created with
// cudaEventBlockingSync instead, which gives the core back while the kernels run.  Measured on one B200 with a 16-core
// host (profiles/README.md, call AC): with 16 phase-2 workers keeping every core busy the sleeping lane thread is woken
// late at each of the ~10 waits of a batch (GPU stage 75 -> 93 ms, end to end 2.45 -> 2.01 M reads/s); with 4 workers,
// where the host is the slower side, the two are equal within the run-to-run spread (call AJ: 1.44 vs 1.40 M reads/s).
// The mode is the caller's to set (bsq_set_wait_mode); the host code's pipeline can drive it (BQ_ADAPTIVE_WAIT=1, off by
// default: no measured gain).
#if defined(__CUDACC__)
#include <cuda_runtime.h>
extern int g_bsq_wait_blocking;  // set through bsq_set_wait_mode (bsq_align.cu): the caller knows which side of its pipeline is the slower one
static inline cudaError_t bsq_stream_wait(cudaStream_t s) {
  static int forced = -2;  // BSQ_SPIN_WAIT=1: always spin, =0: always block, unset: what the caller asked for (default: spin)
  if (forced == -2) { const char *e = getenv("BSQ_SPIN_WAIT"); forced = e ? (atoi(e) != 0) : -1; }
  const int spin = forced >= 0 ? forced : !g_bsq_wait_blocking;
  if (spin) return cudaStreamSynchronize(s);
  static thread_local cudaEvent_t ev[64];
  static thread_local bool have[64];
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev < 0 || dev >= 64) return cudaStreamSynchronize(s);
  if (!have[dev]) {
    if ((e = cudaEventCreateWithFlags(&ev[dev], cudaEventBlockingSync | cudaEventDisableTiming)) != cudaSuccess) return e;
    have[dev] = true;
  }
  if ((e = cudaEventRecord(ev[dev], s)) != cudaSuccess) return e;
  return cudaEventSynchronize(ev[dev]);
}
#endif
