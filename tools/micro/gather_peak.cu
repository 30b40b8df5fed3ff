// Micro-benchmark: ceiling of random 32-byte (one sector) and 64-byte gathers from HBM on this device, as a function of
// the footprint and of the number of independent loads in flight per thread.  Used to put the FM-index kernels'
// gather rate (bsq_seed3.cuh, k_sa) next to what the memory system delivers for this access pattern.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_peak gather_peak.cu && ./gather_peak
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ void ld256(const uint32_t *p, uint32_t (&w)[8]) {
  asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]) : "l"(p));
}

template <int MLP, int BYTES>
__global__ void __launch_bounds__(128) k_gather(const uint32_t *buf, uint64_t n_units, int iters, uint32_t *out) {
  uint64_t s = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 0x9E3779B97F4A7C15ull + 12345;
  uint32_t acc = 0;
  for (int it = 0; it < iters; ++it) {
    uint32_t w[MLP][8];
#pragma unroll
    for (int m = 0; m < MLP; ++m) {
      s = s * 6364136223846793005ull + 1442695040888963407ull;
      const uint64_t u = (s >> 20) % n_units;
      ld256(buf + u * (BYTES / 4), w[m]);
      if (BYTES == 64) { uint32_t v[8]; ld256(buf + u * 16 + 8, v); acc ^= v[3]; }
    }
#pragma unroll
    for (int m = 0; m < MLP; ++m) acc ^= w[m][0] + w[m][7];
    s += acc & 1;  // the next addresses depend on the data: a dependent chain per thread, like an FM-index walk
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int MLP, int BYTES>
static void run(const uint32_t *buf, uint64_t bytes, int ctas_per_sm, uint32_t *out) {
  const int iters = 2000 / MLP;
  const int grid = 148 * ctas_per_sm;
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  k_gather<MLP, BYTES><<<grid, 128>>>(buf, bytes / BYTES, iters / 10 + 1, out);
  cudaEventRecord(a);
  k_gather<MLP, BYTES><<<grid, 128>>>(buf, bytes / BYTES, iters, out);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  const double n = (double)grid * 128 * iters * MLP;
  printf("{\"footprint_gb\": %.1f, \"bytes\": %d, \"mlp\": %d, \"warps_per_sm\": %d, \"gathers_per_s\": %.3e, \"GBps\": %.1f}\n", bytes / 1e9, BYTES, MLP,
         ctas_per_sm * 4, n / (ms * 1e-3), n * BYTES / (ms * 1e-3) / 1e9);
  fflush(stdout);
}

int main() {
  uint32_t *out; cudaMalloc(&out, 148 * 16 * 128 * 4);
  for (double gb : {0.1, 1.0, 6.2, 60.0}) {
    uint64_t bytes = (uint64_t)(gb * 1e9) / 64 * 64;
    uint32_t *buf;
    if (cudaMalloc(&buf, bytes) != cudaSuccess) { printf("{\"skip\": %.1f}\n", gb); continue; }
    cudaMemset(buf, 1, bytes);
    for (int ctas : {4, 8, 16}) {
      run<1, 32>(buf, bytes, ctas, out);
      run<2, 32>(buf, bytes, ctas, out);
      run<4, 32>(buf, bytes, ctas, out);
      run<2, 64>(buf, bytes, ctas, out);
    }
    cudaFree(buf);
  }
  return 0;
}
