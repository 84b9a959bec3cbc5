// scatter_probe.cu -- what does B200's L2 sustain for fp64 reductions (red.global.add.f64) into RANDOM 256-byte rows?
//
// The transposed product of the CG operator, Z = A'T with T = A P ([N][32], 2.56 GB), is a gather of 200 M rows of T
// in the shipped form (49.8 GB of DRAM reads, 7.2 ms).  The scatter form streams T once and adds each row of T into the
// ~20 rows of Z its matrix row names: 200 M row updates = 6.4 G fp64 reductions at L2.  This probe issues exactly those
// updates with nothing around them (coalesced index stream, a warp adds 32 consecutive doubles per index, the value
// comes from a register) for several sizes of the Z table, to see whether a column-blocked scatter (Z slab resident in
// L2) could beat the gather.
//
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/_build/scatter_probe tools/scatter_probe.cu
//   tools/_build/scatter_probe [nupdates_millions=200]          -> JSON lines on stdout
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); exit(2); } } while (0)

__global__ void fill_idx(int* idx, long n, unsigned nrows, unsigned long long seed) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (unsigned long long)(i + 1);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    idx[i] = (int)(((z >> 32) * nrows) >> 32);
  }
}

// a warp takes 32 indices (one coalesced load), then for each of them all 32 lanes add one double into the row
template <int U>
__global__ void __launch_bounds__(256) scatter_kernel(const int* __restrict__ idx, long n, double* __restrict__ tab) {
  const int lane = threadIdx.x & 31;
  const long warp = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long nwarps = ((long)gridDim.x * blockDim.x) >> 5;
  for (long base = warp * 32; base < n; base += nwarps * 32) {
    const int mine = (base + lane < n) ? idx[base + lane] : -1;
#pragma unroll 4
    for (int j = 0; j < 32; ++j) {
      const int r = __shfl_sync(0xffffffffu, mine, j);
      if (r >= 0) asm volatile("red.global.add.f64 [%0], %1;" ::"l"(tab + (long)r * 32 + lane), "d"(1.0 + lane) : "memory");
    }
  }
}

int main(int argc, char** argv) {
  const long n = (argc > 1 ? atol(argv[1]) : 200) * 1000000L;
  int* idx; CK(cudaMalloc(&idx, n * 4));
  const size_t maxb = (size_t)512 << 20;
  double* tab; CK(cudaMalloc(&tab, maxb));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const long mbs[] = {16, 32, 48, 64, 96, 128, 256, 512};
  for (long mb : mbs) {
    const unsigned nrows = (unsigned)((mb << 20) / 256);
    fill_idx<<<148 * 8, 256>>>(idx, n, nrows, 0x1234 + mb);
    CK(cudaMemset(tab, 0, maxb));
    for (int rep = 0; rep < 2; ++rep) {
      CK(cudaEventRecord(e0));
      scatter_kernel<4><<<148 * 8, 256>>>(idx, n, tab);
      CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    }
    CK(cudaGetLastError());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    double row0; CK(cudaMemcpy(&row0, tab, 8, cudaMemcpyDeviceToHost));
    printf("{\"probe\": \"fp64 red.add into random 256-byte rows\", \"table_mb\": %ld, \"row_updates\": %ld, \"ms\": %.3f, \"g_red_per_s\": %.1f, \"updated_tb_per_s\": %.2f}\n",
           mb, n, ms, n * 32.0 / ms / 1e6, n * 256.0 / ms / 1e9);
    fflush(stdout);
  }
  return 0;
}
