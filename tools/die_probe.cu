// die_probe.cu -- does B200's two-die L2 keep a private copy of every line per READING die, and what would die-affine
// gathers buy?
//
// profiles/r2_gather_ceiling.md: a gather operand read from every SM gets about half of the 126 MB L2.  If that is because
// each die's L2 partition caches whatever its own SMs read, then letting the SMs of die 0 gather only from one half of
// the operand and the SMs of die 1 only from the other half would give the operand the whole L2.
//   step 1: SM -> die map.  One SM touches a set of lines (they now sit in ITS die's partition); every other SM then reads
//           its own fresh subset of them once, timed: a same-die SM sees near-L2 hits, an other-die SM does not.
//   step 2: 200 M gathers from a table twice the "usable" L2, (a) every CTA from the whole table (today's product),
//           (b) die-affine: persistent CTAs take tiles from their die's queue, die d only touches half d of the table.
//
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/_build/die_probe tools/die_probe.cu
//   tools/_build/die_probe [ngather_millions=200]          -> JSON lines on stdout
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <algorithm>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); exit(2); } } while (0)

__device__ __forceinline__ unsigned smid() { unsigned s; asm volatile("mov.u32 %0, %%smid;" : "=r"(s)); return s; }
__device__ __forceinline__ double ldcg(const double* p) { double v; asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ long long clk_after(double dep) {
  long long t; asm volatile("mov.u64 %0, %%clock64;" : "=l"(t) : "d"(dep) : "memory"); return t;
}

constexpr int kLinesPerSm = 192;
constexpr long kLineStride = 4096 / 8;   // doubles between probed lines (one line per 4 KB)

__global__ void touch_kernel(const double* tab, long nlines, unsigned ref_sm, int* claimed, double* sink) {
  __shared__ int mine;
  if (threadIdx.x == 0) mine = (smid() == ref_sm && atomicCAS(claimed, 0, 1) == 0) ? 1 : 0;
  __syncthreads();
  if (!mine) return;
  double a = 0.0;
  for (int rep = 0; rep < 2; ++rep)
    for (long l = threadIdx.x; l < nlines; l += blockDim.x) a += ldcg(tab + l * kLineStride);
  if (a == 1.2345e300) sink[0] = a;
}

__global__ void time_kernel(const double* tab, int* claimed_sm, float* mean_lat, float* frac_fast, double* sink) {
  const unsigned sm = smid();
  if (threadIdx.x != 0) return;
  if (atomicCAS(claimed_sm + sm, 0, 1) != 0) return;
  // a dependent chain: the address of load i+1 is computed from the value of load i (v * 0.0 cannot be folded), so the
  // elapsed clocks over the chain are kLinesPerSm full round trips
  long long tot = 0; int fast = 0; double a = 0.0;
  long off = 0;
  const long long t0 = clock64();
  for (int i = 0; i < kLinesPerSm; ++i) {
    const double v = ldcg(tab + ((long)sm * kLinesPerSm + i + off) * kLineStride);
    off = (long)(v * 0.0);
    a += v;
  }
  tot = clock64() - t0 + off;
  mean_lat[sm] = (float)tot / kLinesPerSm; frac_fast[sm] = (float)fast / kLinesPerSm;
  if (a == 1.2345e300) sink[0] = a;
}

__global__ void fill_idx(int* idx, long n, unsigned lo, unsigned span, unsigned long long seed) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (unsigned long long)(i + 1);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    idx[i] = (int)(lo + (unsigned)(((z >> 32) * span) >> 32));
  }
}
__global__ void fill_tab(double* t, long n) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) t[i] = (double)(i & 1023) * 1e-3;
}

// persistent gather: CTA on die d takes tiles of idx[q(d)] from counter q(d); mode 0: one queue for everybody
template <int LPG, int U>
__global__ void __launch_bounds__(256) gather_q_kernel(const int* __restrict__ idx0, const int* __restrict__ idx1, long n0, long n1,
                                                       const double* __restrict__ tab, int rowd, const int* __restrict__ die_of_sm, int mode,
                                                       unsigned long long* ctr, double* __restrict__ out) {
  constexpr int NG = 256 / LPG, TILE = 4096;
  __shared__ int s_idx[TILE];
  __shared__ long s_base;
  const int sub = threadIdx.x % LPG, grp = threadIdx.x / LPG;
  unsigned long long pk;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pk));
  int q = 0;
  if (mode == 1) q = die_of_sm[smid()];
  else if (mode == 2) q = 1 - die_of_sm[smid()];
  else if (mode == 3) q = blockIdx.x & 1;            // control: two queues, CTAs assigned regardless of die
  const int* idx = q ? idx1 : idx0;
  const long n = q ? n1 : n0;
  double2 acc = make_double2(0.0, 0.0);
  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) s_base = (long)atomicAdd(ctr + q, (unsigned long long)TILE);
    __syncthreads();
    const long base = s_base;
    if (base >= n) break;
    const int cnt = (int)min((long)TILE, n - base);
    for (int i = threadIdx.x; i < cnt; i += 256) s_idx[i] = idx[base + i];
    __syncthreads();
    for (int b = grp; b < cnt; b += NG * U) {
      int c[U];
#pragma unroll
      for (int u = 0; u < U; ++u) { const int j = b + u * NG; c[u] = j < cnt ? s_idx[j] : -1; }
      double2 v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        v[u] = make_double2(0.0, 0.0);
        if (c[u] >= 0) asm volatile("ld.global.nc.L2::cache_hint.v2.f64 {%0,%1}, [%2], %3;" : "=d"(v[u].x), "=d"(v[u].y) : "l"(tab + (long)c[u] * rowd + sub * 2), "l"(pk));
      }
#pragma unroll
      for (int u = 0; u < U; ++u) { acc.x += v[u].x; acc.y += v[u].y; }
    }
  }
  if (acc.x + acc.y == 1.2345e300) out[0] = acc.x;
}

int main(int argc, char** argv) {
  const long ng = (argc > 1 ? atol(argv[1]) : 200) * 1000000L;
  int nsm = 0; CK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0));
  double *tab, *sink; CK(cudaMalloc(&tab, 1L << 30)); CK(cudaMalloc(&sink, 64));
  fill_tab<<<148 * 8, 256>>>(tab, (1L << 30) / 8);
  CK(cudaDeviceSynchronize());
  // ---- step 1: SM -> die
  std::vector<int> die(256, 0);
  {
    int *claimed, *claimed_sm; float *lat, *ff;
    CK(cudaMalloc(&claimed, 4)); CK(cudaMalloc(&claimed_sm, 256 * 4)); CK(cudaMalloc(&lat, 256 * 4)); CK(cudaMalloc(&ff, 256 * 4));
    std::vector<float> hl(256), hf(256);
    for (unsigned ref : {0u, 1u}) {
      // evict whatever an earlier round left: stream 512 MB of another region through L2
      fill_tab<<<148 * 8, 256>>>(tab + (512L << 20) / 8, (512L << 20) / 8);
      CK(cudaMemset(claimed, 0, 4)); CK(cudaMemset(claimed_sm, 0, 256 * 4)); CK(cudaMemset(lat, 0, 256 * 4));
      touch_kernel<<<nsm * 16, 64>>>(tab, (long)nsm * kLinesPerSm, ref, claimed, sink);
      CK(cudaDeviceSynchronize());
      time_kernel<<<nsm * 16, 32>>>(tab, claimed_sm, lat, ff, sink);
      CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(hl.data(), lat, 256 * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(hf.data(), ff, 256 * 4, cudaMemcpyDeviceToHost));
      std::vector<float> s(hl.begin(), hl.begin() + nsm); std::sort(s.begin(), s.end());
      // largest gap in the sorted means splits the two groups
      int cut = 1; float gap = 0;
      for (int i = 1; i < nsm; ++i) if (s[i] - s[i - 1] > gap) { gap = s[i] - s[i - 1]; cut = i; }
      const float thr = 0.5f * (s[cut] + s[cut - 1]);
      int same = 0;
      for (int i = 0; i < nsm; ++i) same += hl[i] < thr;
      printf("{\"step\": \"sm_to_die\", \"ref_sm\": %u, \"mean_latency_min\": %.0f, \"below_gap\": %.0f, \"above_gap\": %.0f, \"max\": %.0f, \"gap\": %.0f, "
             "\"sms_with_ref\": %d, \"sms_other\": %d, \"lat_ref_sm\": %.0f}\n", ref, s[0], s[cut - 1], s[cut], s[nsm - 1], gap, same, nsm - same, hl[ref]);
      if (ref == 0) for (int i = 0; i < nsm; ++i) die[i] = hl[i] < thr ? 0 : 1;
      else {
        int agree = 0;
        for (int i = 0; i < nsm; ++i) agree += ((hl[i] < thr ? die[1] : 1 - die[1]) == die[i]);
        printf("{\"step\": \"sm_to_die_check\", \"agree_with_first_map\": %d, \"of\": %d}\n", agree, nsm);
      }
      fflush(stdout);
    }
    printf("{\"step\": \"die_map\", \"map\": \"");
    for (int i = 0; i < nsm; ++i) printf("%d", die[i]);
    printf("\"}\n");
  }
  // ---- step 2: die-affine gathers
  int* d_die; CK(cudaMalloc(&d_die, 256 * 4)); CK(cudaMemcpy(d_die, die.data(), 256 * 4, cudaMemcpyHostToDevice));
  int *idx0, *idx1; CK(cudaMalloc(&idx0, ng * 4)); CK(cudaMalloc(&idx1, ng * 4));
  unsigned long long* ctr; CK(cudaMalloc(&ctr, 16));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  auto run = [&](const char* label, int lpg, long table_mb, int mode, int ctas) {
    const unsigned nrows = (unsigned)((table_mb << 20) / 256);
    const long n0 = mode == 0 ? ng : ng / 2, n1 = mode == 0 ? 0 : ng - ng / 2;
    if (mode == 0) fill_idx<<<148 * 8, 256>>>(idx0, n0, 0, nrows, 0x77 + table_mb);
    else {
      fill_idx<<<148 * 8, 256>>>(idx0, n0, 0, nrows / 2, 0x77 + table_mb);
      fill_idx<<<148 * 8, 256>>>(idx1, n1, nrows / 2, nrows - nrows / 2, 0x99 + table_mb);
    }
    float best = 1e30f;
    for (int it = 0; it < 5; ++it) {
      CK(cudaMemset(ctr, 0, 16));
      CK(cudaEventRecord(e0));
      if (lpg == 8) gather_q_kernel<8, 6><<<148 * ctas, 256>>>(idx0, idx1, n0, n1, tab, 32, d_die, mode, ctr, sink);
      else gather_q_kernel<16, 6><<<148 * ctas, 256>>>(idx0, idx1, n0, n1, tab, 32, d_die, mode, ctr, sink);
      CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
      if (it >= 1 && ms < best) best = ms;
    }
    CK(cudaGetLastError());
    printf("{\"step\": \"gather\", \"assignment\": \"%s\", \"gran_B\": %d, \"table_MB\": %ld, \"line_footprint_MB\": %ld, \"ctas_per_sm\": %d, \"ngather\": %ld, \"ms\": %.4f, \"gathered_TBs\": %.2f}\n",
           label, lpg * 16, table_mb, lpg == 8 ? table_mb / 2 : table_mb, ctas, ng, best, (double)ng * lpg * 16 / (best * 1e-3) / 1e12);
    fflush(stdout);
  };
  const char* names[4] = {"any CTA, whole table (today)", "die-affine halves", "die-affine, halves swapped", "two queues, CTAs assigned by blockIdx parity (control)"};
  for (int ctas : {5, 8})
    for (long mb : {256L, 128L}) {
      for (int mode = 0; mode < 4; ++mode) run(names[mode], 8, mb, mode, ctas);     // 128-byte slabs of 256-byte rows (the two-pass product)
      for (int mode = 0; mode < 4; ++mode) run(names[mode], 16, mb, mode, ctas);    // whole rows (the one-pass product)
    }
  return 0;
}
