// gather_probe.cu -- what does B200 sustain for RANDOM row gathers out of L2 / HBM?
//
// The C2 product (binary CSR, R = 32) is 200 M gathers of 256-byte X rows (two column passes: 400 M gathers of
// 128 bytes) from a 256 MB table, driven by a 0.8 GB coalesced index stream.  This probe issues exactly that access
// pattern with nothing else around it -- no row structure, no Y writes, one fp64 add per gathered double -- so its
// rate is the machine's ceiling for the gather formulation:
//   * index stream: int32, read once, coalesced, L1::no_allocate + L2 evict_first (like the staged kernel's TMA copy);
//   * each gather: GRAN bytes, GRAN/16 lanes x one 128-bit ld.global.nc with an L2 evict_last hint;
//   * U gathers in flight per lane group before the adds (U = 8 is the deep build of the product kernel).
// Reported per (table size, GRAN): gathered TB/s, sectors/s, bytes per SM clock (clock measured inside the kernel from
// clock64 / globaltimer), next to a plain streaming read of the same number of bytes.
//
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/_build/gather_probe tools/gather_probe.cu
//   tools/_build/gather_probe [ngather_millions=200]          -> JSON lines on stdout
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); exit(2); } } while (0)

__device__ __forceinline__ unsigned long long policy(int kind) {
  unsigned long long p;
  if (kind == 1) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  else asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ double2 ldx(const double* p, unsigned long long pol) {
  double2 v;
  asm volatile("ld.global.nc.L2::cache_hint.v2.f64 {%0,%1}, [%2], %3;" : "=d"(v.x), "=d"(v.y) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ int ldi(const int* p, unsigned long long pol) {
  int v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

__global__ void fill_idx(int* idx, long n, unsigned nrows, unsigned long long seed) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (unsigned long long)(i + 1);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    idx[i] = (int)(((z >> 32) * nrows) >> 32);
  }
}
__global__ void fill_tab(double* t, long n) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) t[i] = (double)(i & 1023) * 1e-3;
}

// LPG lanes per gather (GRAN = 16 * LPG bytes); a warp handles 32/LPG gathers per step, U steps in flight.
// ROWB = bytes between consecutive table rows (256: the [ncol][32] operand; GRAN < ROWB reads a column slab of it)
template <int LPG, int U>
__global__ void __launch_bounds__(256) gather_kernel(const int* __restrict__ idx, long ngather, const double* __restrict__ tab, int rowd,
                                                     double* __restrict__ out, unsigned long long* __restrict__ clk) {
  const unsigned long long pk = policy(1), ps = policy(2);
  const int sub = threadIdx.x % LPG, grp = threadIdx.x / LPG;
  constexpr int NG = 256 / LPG;                       // gather groups per CTA
  constexpr int TILE = 4096;                          // indices staged per CTA step (like the product kernel's row block)
  __shared__ int s_idx[TILE];
  unsigned long long c0 = 0, t0 = 0;
  if (blockIdx.x == 0 && threadIdx.x == 0) { c0 = clock64(); t0 = gtime(); }
  double2 acc = make_double2(0.0, 0.0);
  for (long base = (long)blockIdx.x * TILE; base < ngather; base += (long)gridDim.x * TILE) {
    const int n = (int)min((long)TILE, ngather - base);
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += 256) s_idx[i] = ldi(idx + base + i, ps);   // coalesced, once
    __syncthreads();
    for (int b = grp; b < n; b += NG * U) {
      int c[U];
#pragma unroll
      for (int u = 0; u < U; ++u) { const int j = b + u * NG; c[u] = j < n ? s_idx[j] : -1; }
      double2 v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) v[u] = c[u] >= 0 ? ldx(tab + (long)c[u] * rowd + sub * 2, pk) : make_double2(0.0, 0.0);
#pragma unroll
      for (int u = 0; u < U; ++u) { acc.x += v[u].x; acc.y += v[u].y; }
    }
  }
  if (acc.x + acc.y == 1.2345e300) out[0] = acc.x;   // keep the loads alive
  if (blockIdx.x == 0 && threadIdx.x == 0) { clk[0] = clock64() - c0; clk[1] = gtime() - t0; }
}

// 8-byte gathers (one right-hand side): every lane its own index, U loads in flight
template <int U>
__global__ void __launch_bounds__(256) gather8_kernel(const int* __restrict__ idx, long ngather, const double* __restrict__ tab,
                                                      double* __restrict__ out, unsigned long long* __restrict__ clk) {
  const unsigned long long pk = policy(1), ps = policy(2);
  unsigned long long c0 = 0, t0 = 0;
  if (blockIdx.x == 0 && threadIdx.x == 0) { c0 = clock64(); t0 = gtime(); }
  double acc = 0.0;
  const long nthr = (long)gridDim.x * blockDim.x, tid = (long)blockIdx.x * blockDim.x + threadIdx.x;
  for (long b = (long)blockIdx.x * blockDim.x * U; b < ngather; b += nthr * U) {
    int c[U];
    double v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) { const long j = b + (long)u * blockDim.x + threadIdx.x; c[u] = j < ngather ? ldi(idx + j, ps) : -1; }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      v[u] = 0.0;
      if (c[u] >= 0) asm volatile("ld.global.nc.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v[u]) : "l"(tab + c[u]), "l"(pk));
    }
#pragma unroll
    for (int u = 0; u < U; ++u) acc += v[u];
  }
  if (acc == 1.2345e300) out[tid & 7] = acc;
  if (blockIdx.x == 0 && threadIdx.x == 0) { clk[0] = clock64() - c0; clk[1] = gtime() - t0; }
}

__global__ void __launch_bounds__(256) stream_kernel(const double* __restrict__ tab, long n2, double* out) {
  const unsigned long long ps = policy(2);
  double2 acc = make_double2(0.0, 0.0);
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (long)gridDim.x * blockDim.x) {
    double2 v = ldx(tab + 2 * i, ps);
    acc.x += v.x; acc.y += v.y;
  }
  if (acc.x + acc.y == 1.2345e300) out[0] = acc.x;
}

template <int LPG, int U>
static void run(const char* label, const int* idx, long ng, const double* tab, long table_bytes, int rowd, double* out, unsigned long long* clk,
                int ctas_per_sm) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const int grid = 148 * ctas_per_sm;
  for (int w = 0; w < 2; ++w) gather_kernel<LPG, U><<<grid, 256>>>(idx, ng, tab, rowd, out, clk);
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  unsigned long long h[2] = {0, 0}, hb[2] = {1, 1};
  for (int it = 0; it < 5; ++it) {
    CK(cudaEventRecord(e0));
    gather_kernel<LPG, U><<<grid, 256>>>(idx, ng, tab, rowd, out, clk);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    CK(cudaMemcpy(h, clk, 16, cudaMemcpyDeviceToHost));
    if (ms < best) { best = ms; hb[0] = h[0]; hb[1] = h[1]; }
  }
  const double gran = 16.0 * LPG, bytes = (double)ng * gran, mhz = (double)hb[0] / (double)hb[1] * 1e3;
  const double tbs = bytes / (best * 1e-3) / 1e12;
  printf("{\"probe\": \"%s\", \"table_MB\": %.0f, \"gran_B\": %.0f, \"in_flight\": %d, \"ctas_per_sm\": %d, \"ngather\": %ld, \"ms\": %.4f, "
         "\"gathered_TBs\": %.3f, \"gathers_per_s\": %.4g, \"sectors_per_s\": %.4g, \"sm_mhz_in_kernel\": %.0f, \"bytes_per_clk\": %.0f}\n",
         label, table_bytes / 1e6, gran, U, ctas_per_sm, ng, best, tbs, ng / (best * 1e-3), bytes / 32 / (best * 1e-3), mhz,
         bytes / (best * 1e-3) / (mhz * 1e6));
  fflush(stdout);
  CK(cudaEventDestroy(e0)); CK(cudaEventDestroy(e1));
}


// ---- the same gathers through cp.async (LDGSTS, .cg = L2 only): every lane keeps D 16-byte slots of a shared-memory ring
// in flight instead of U registers, so the data in flight is bounded by the ring (D x 4 KB per CTA), not by registers
// or by the L1 lines a missing LDG holds.  B gathers per commit group, D / B groups in flight.
template <int LPG, int D, int B>
__global__ void __launch_bounds__(256) gather_async_kernel(const int* __restrict__ idx, long ngather, const double* __restrict__ tab, int rowd,
                                                           double* __restrict__ out, unsigned long long* __restrict__ clk) {
  const unsigned long long ps = policy(2);
  const int sub = threadIdx.x % LPG, grp = threadIdx.x / LPG;
  constexpr int NG = 256 / LPG;
  constexpr int TILE = 2048;
  extern __shared__ __align__(16) unsigned char smem[];
  int* s_idx = reinterpret_cast<int*>(smem);
  double2* ring = reinterpret_cast<double2*>(smem + TILE * 4);   // [D][256]
  unsigned long long c0 = 0, t0 = 0;
  if (blockIdx.x == 0 && threadIdx.x == 0) { c0 = clock64(); t0 = gtime(); }
  double2 acc = make_double2(0.0, 0.0);
  for (long base = (long)blockIdx.x * TILE; base < ngather; base += (long)gridDim.x * TILE) {
    const int n = (int)min((long)TILE, ngather - base);
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += 256) s_idx[i] = ldi(idx + base + i, ps);
    __syncthreads();
    const int K = n > grp ? (n - grp + NG - 1) / NG : 0;       // this group's gathers: entries grp + k NG
    auto issue = [&](int k0) {
#pragma unroll
      for (int u = 0; u < B; ++u) {
        const int k = k0 + u;
        if (k < K) {
          const int c = s_idx[grp + k * NG];
          const unsigned dst = (unsigned)__cvta_generic_to_shared(ring + (k % D) * 256 + threadIdx.x);
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(tab + (long)c * rowd + sub * 2) : "memory");
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
#pragma unroll
    for (int k0 = 0; k0 < D; k0 += B) issue(k0);
    for (int k0 = 0; k0 < K; k0 += B) {
      asm volatile("cp.async.wait_group %0;" ::"n"(D / B - 1) : "memory");
#pragma unroll
      for (int u = 0; u < B; ++u) {
        const int k = k0 + u;
        if (k < K) { const double2 v = ring[(k % D) * 256 + threadIdx.x]; acc.x += v.x; acc.y += v.y; }
      }
      issue(k0 + D);
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
  }
  if (acc.x + acc.y == 1.2345e300) out[0] = acc.x;
  if (blockIdx.x == 0 && threadIdx.x == 0) { clk[0] = clock64() - c0; clk[1] = gtime() - t0; }
}

template <int LPG, int D, int B>
static void run_async(const char* label, const int* idx, long ng, const double* tab, long table_bytes, int rowd, double* out,
                      unsigned long long* clk, int ctas_per_sm) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const int grid = 148 * ctas_per_sm;
  const int smem = 2048 * 4 + D * 256 * 16;
  CK(cudaFuncSetAttribute(gather_async_kernel<LPG, D, B>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  for (int w = 0; w < 2; ++w) gather_async_kernel<LPG, D, B><<<grid, 256, smem>>>(idx, ng, tab, rowd, out, clk);
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  unsigned long long h[2] = {0, 0}, hb[2] = {1, 1};
  for (int it = 0; it < 5; ++it) {
    CK(cudaEventRecord(e0));
    gather_async_kernel<LPG, D, B><<<grid, 256, smem>>>(idx, ng, tab, rowd, out, clk);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    CK(cudaMemcpy(h, clk, 16, cudaMemcpyDeviceToHost));
    if (ms < best) { best = ms; hb[0] = h[0]; hb[1] = h[1]; }
  }
  const double gran = 16.0 * LPG, bytes = (double)ng * gran, mhz = (double)hb[0] / (double)hb[1] * 1e3;
  printf("{\"probe\": \"%s\", \"path\": \"cp.async ring\", \"table_MB\": %.0f, \"gran_B\": %.0f, \"ring_depth\": %d, \"ctas_per_sm\": %d, "
         "\"ring_kb_per_sm\": %d, \"ngather\": %ld, \"ms\": %.4f, \"gathered_TBs\": %.3f, \"sm_mhz_in_kernel\": %.0f}\n",
         label, table_bytes / 1e6, gran, D, ctas_per_sm, D * 4 * ctas_per_sm, ng, best, bytes / (best * 1e-3) / 1e12, mhz);
  fflush(stdout);
  CK(cudaEventDestroy(e0)); CK(cudaEventDestroy(e1));
}

int main(int argc, char** argv) {
  const long ng = (argc > 1 ? atol(argv[1]) : 200) * 1000000L;
  int* idx; double *tab, *out; unsigned long long* clk;
  const long tab_max = 1024L << 20;
  CK(cudaMalloc(&idx, ng * 4)); CK(cudaMalloc(&tab, tab_max)); CK(cudaMalloc(&out, 64)); CK(cudaMalloc(&clk, 16));
  fill_tab<<<148 * 8, 256>>>(tab, tab_max / 8);
  CK(cudaDeviceSynchronize());

  if (argc > 2 && !strcmp(argv[2], "async")) {
    // C2 point only: LDG (registers + L1 lines) against the cp.async ring, 128-byte slabs and whole 256-byte rows
    const unsigned nrows = (unsigned)((256L << 20) / 256);
    fill_idx<<<148 * 8, 256>>>(idx, ng, nrows, 0x5EED0100ull);
    CK(cudaDeviceSynchronize());
    run<8, 8>("gather128_slab", idx, ng, tab, 128L << 20, 32, out, clk, 8);
    run<8, 4>("gather128_slab", idx, ng, tab, 128L << 20, 32, out, clk, 8);
    run_async<8, 8, 4>("gather128_slab", idx, ng, tab, 128L << 20, 32, out, clk, 4);
    run_async<8, 12, 4>("gather128_slab", idx, ng, tab, 128L << 20, 32, out, clk, 3);
    run_async<8, 16, 4>("gather128_slab", idx, ng, tab, 128L << 20, 32, out, clk, 2);
    run_async<8, 24, 4>("gather128_slab", idx, ng, tab, 128L << 20, 32, out, clk, 2);
    run_async<8, 48, 4>("gather128_slab", idx, ng, tab, 128L << 20, 32, out, clk, 1);
    run_async<8, 4, 2>("gather128_slab", idx, ng, tab, 128L << 20, 32, out, clk, 8);
    run<16, 8>("gather256", idx, ng, tab, 256L << 20, 32, out, clk, 8);
    run<16, 4>("gather256", idx, ng, tab, 256L << 20, 32, out, clk, 8);
    run_async<16, 8, 4>("gather256", idx, ng, tab, 256L << 20, 32, out, clk, 4);
    run_async<16, 12, 4>("gather256", idx, ng, tab, 256L << 20, 32, out, clk, 3);
    run_async<16, 24, 4>("gather256", idx, ng, tab, 256L << 20, 32, out, clk, 2);
    run_async<16, 48, 4>("gather256", idx, ng, tab, 256L << 20, 32, out, clk, 1);
    // L2-resident table: the SM-side ceiling of each path
    fill_idx<<<148 * 8, 256>>>(idx, ng, (unsigned)((32L << 20) / 256), 0x5EED0101ull);
    CK(cudaDeviceSynchronize());
    run<16, 8>("gather256", idx, ng, tab, 32L << 20, 32, out, clk, 8);
    run_async<16, 12, 4>("gather256", idx, ng, tab, 32L << 20, 32, out, clk, 3);
    run_async<16, 24, 4>("gather256", idx, ng, tab, 32L << 20, 32, out, clk, 2);
    return 0;
  }
  // plain streaming read of the 1 GB table: the DRAM ceiling the gathers are compared with
  {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int it = 0; it < 6; ++it) {
      CK(cudaEventRecord(e0));
      stream_kernel<<<148 * 8, 256>>>(tab, tab_max / 16, out);
      CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
      if (it && ms < best) best = ms;
    }
    printf("{\"probe\": \"stream_read\", \"table_MB\": %.0f, \"ms\": %.4f, \"TBs\": %.3f}\n", tab_max / 1e6, best, tab_max / (best * 1e-3) / 1e12);
  }
  // rows are 256 bytes apart (the [ncol][32] fp64 operand); GRAN = 256 reads whole rows, 128 / 64 a column slab
  const long sizes_mb[] = {32, 64, 128, 256, 1024};
  for (long mb : sizes_mb) {
    const unsigned nrows = (unsigned)((mb << 20) / 256);
    fill_idx<<<148 * 8, 256>>>(idx, ng, nrows, 0x5EED0000ull + mb);
    CK(cudaDeviceSynchronize());
    run<16, 8>("gather256", idx, ng, tab, mb << 20, 32, out, clk, 8);
    run<8, 8>("gather128_slab", idx, ng, tab, mb << 19, 32, out, clk, 8);   // touches half of every row: footprint mb/2
    run<4, 8>("gather64_slab", idx, ng, tab, mb << 18, 32, out, clk, 8);
  }
  // depth / occupancy sensitivity at the C2 point (256 MB table, 128-byte slabs = the two-pass product)
  {
    const unsigned nrows = (unsigned)((256L << 20) / 256);
    fill_idx<<<148 * 8, 256>>>(idx, ng, nrows, 0x5EED0100ull);
    CK(cudaDeviceSynchronize());
    run<8, 4>("gather128_slab", idx, ng, tab, 128L << 20, 32, out, clk, 8);
    run<8, 16>("gather128_slab", idx, ng, tab, 128L << 20, 32, out, clk, 8);
    run<8, 8>("gather128_slab", idx, ng, tab, 128L << 20, 32, out, clk, 4);
    run<8, 8>("gather128_slab", idx, ng, tab, 128L << 20, 32, out, clk, 16);
    run<16, 4>("gather256", idx, ng, tab, 256L << 20, 32, out, clk, 8);
    run<16, 16>("gather256", idx, ng, tab, 256L << 20, 32, out, clk, 8);
  }
  // narrow operands, rows back to back (row stride = gather size): R = 2 / 4 / 8 of a 1 M-row operand
  {
    fill_idx<<<148 * 8, 256>>>(idx, ng, 1000000u, 0x5EED0200ull);
    CK(cudaDeviceSynchronize());
    run<1, 8>("gather16_R2", idx, ng, tab, 16L * 1000000, 2, out, clk, 8);
    run<2, 8>("gather32_R4", idx, ng, tab, 32L * 1000000, 4, out, clk, 8);
    run<4, 8>("gather64_R8", idx, ng, tab, 64L * 1000000, 8, out, clk, 8);
    run<8, 8>("gather128_R16", idx, ng, tab, 128L * 1000000, 16, out, clk, 8);
  }
  // 8-byte gathers from an 8 MB / 80 MB vector: the SpMV / transposed SpMV request-rate ceiling
  for (long n : {1000000L, 10000000L}) {
    fill_idx<<<148 * 8, 256>>>(idx, ng, (unsigned)n, 0x5EED0300ull + n);
    CK(cudaDeviceSynchronize());
    for (int cps : {4, 8}) {
      cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
      float best = 1e30f; unsigned long long h[2], hb[2] = {1, 1};
      for (int it = 0; it < 6; ++it) {
        CK(cudaEventRecord(e0));
        gather8_kernel<8><<<148 * cps, 256>>>(idx, ng, tab, out, clk);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        CK(cudaMemcpy(h, clk, 16, cudaMemcpyDeviceToHost));
        if (it && ms < best) { best = ms; hb[0] = h[0]; hb[1] = h[1]; }
      }
      const double mhz = (double)hb[0] / (double)hb[1] * 1e3;
      printf("{\"probe\": \"gather8_R1\", \"table_MB\": %.0f, \"gran_B\": 8, \"in_flight\": 8, \"ctas_per_sm\": %d, \"ngather\": %ld, \"ms\": %.4f, "
             "\"gathers_per_s\": %.4g, \"sm_mhz_in_kernel\": %.0f, \"gathers_per_clk_per_sm\": %.3f}\n",
             n * 8 / 1e6, cps, ng, best, ng / (best * 1e-3), mhz, ng / (best * 1e-3) / (mhz * 1e6) / 148.0);
      fflush(stdout);
    }
  }
  return 0;
}
