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
#include <cub/cub.cuh>

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


// ---- step 3: the whole product, die-affine.  A (fixed 20 entries per row, uniform columns) is split by column half into
// A0 / A1; persistent CTAs on die 0 run through ALL row blocks with A0 and store the partial rows into Y, CTAs on die 1
// run through all row blocks with A1, wait for the block's flag and add their sums to what die 0 stored (Y is re-read
// while still in L2).  Mode 0 is the same kernel on the unsplit matrix with one queue (today's product, minus TMA).
constexpr int kRB = 128;          // rows per block
constexpr int kCap = 4096;        // staged indices per block
template <int U>
__global__ void __launch_bounds__(256, 5) spmm_q_kernel(int nrow, const int* __restrict__ rp0, const int* __restrict__ c0,
                                                        const int* __restrict__ rp1, const int* __restrict__ c1,
                                                        const double* __restrict__ X, double* __restrict__ Y, int col0,
                                                        const int* __restrict__ die_of_sm, int mode, unsigned long long* ctr,
                                                        int* ready, int epoch) {
  __shared__ int s_rp[kRB + 1];
  __shared__ int s_c[kCap];
  __shared__ int s_blk;
  unsigned long long pk;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pk));
  const int q = mode == 0 ? 0 : die_of_sm[smid()];
  const int* rp = q ? rp1 : rp0;
  const int* cc = q ? c1 : c0;
  const int nblk = (nrow + kRB - 1) / kRB;
  const int team = threadIdx.x >> 3, l = threadIdx.x & 7;   // 8 lanes x 16 bytes = one 128-byte slab row
  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) s_blk = (int)atomicAdd(ctr + q, 1ull);
    __syncthreads();
    const int b = s_blk;
    if (b >= nblk) break;
    const int r0 = b * kRB, nr = min(kRB, nrow - r0);
    for (int i = threadIdx.x; i <= nr; i += 256) s_rp[i] = rp[r0 + i];
    __syncthreads();
    const int base = s_rp[0], tot = s_rp[nr] - base;
    for (int i = threadIdx.x; i < tot && i < kCap; i += 256) s_c[i] = cc[base + i];
    if (mode == 1 && q == 1 && threadIdx.x == 0) {          // die 1 adds to what die 0 stored for this block
      int v;
      do { asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(ready + b) : "memory"); } while (v != epoch);
    }
    __syncthreads();
    for (int r = team; r < nr; r += 32) {
      const int s = s_rp[r] - base, e = s_rp[r + 1] - base;
      double2 acc = make_double2(0.0, 0.0);
      for (int i = s; i < e; i += U) {
        double2 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          v[u] = make_double2(0.0, 0.0);
          if (i + u < e) {
            const int c = s_c[i + u];
            asm volatile("ld.global.nc.L2::cache_hint.v2.f64 {%0,%1}, [%2], %3;" : "=d"(v[u].x), "=d"(v[u].y) : "l"(X + (long)c * 32 + col0 + l * 2), "l"(pk));
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) { acc.x += v[u].x; acc.y += v[u].y; }
      }
      double* y = Y + (long)(r0 + r) * 32 + col0 + l * 2;
      if (mode == 1 && q == 1) {
        double px, py;
        asm volatile("ld.global.cg.v2.f64 {%0,%1}, [%2];" : "=d"(px), "=d"(py) : "l"(y) : "memory");
        acc.x += px; acc.y += py;
      }
      asm volatile("st.global.cs.v2.f64 [%0], {%1,%2};" ::"l"(y), "d"(acc.x), "d"(acc.y) : "memory");
    }
    if (mode == 1 && q == 0) {                               // publish the block's partial rows
      __syncthreads();
      if (threadIdx.x == 0) { __threadfence(); asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(ready + b), "r"(epoch) : "memory"); }
    }
  }
}

__global__ void gen_cols(int* cols, long nnz, unsigned ncol, unsigned long long seed) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < nnz; i += (long)gridDim.x * blockDim.x) {
    unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (unsigned long long)(i + 1);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    cols[i] = (int)(((z >> 32) * ncol) >> 32);
  }
}
__global__ void count_lo(const int* cols, int nrow, int deg, int half, int* cnt0, int* cnt1) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nrow) return;
  int a = 0;
  for (int k = 0; k < deg; ++k) a += cols[(long)r * deg + k] < half;
  cnt0[r] = a; cnt1[r] = deg - a;
}
__global__ void split_fill(const int* cols, int nrow, int deg, int half, const int* rp0, const int* rp1, int* c0, int* c1) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nrow) return;
  int a = rp0[r], b = rp1[r];
  for (int k = 0; k < deg; ++k) { const int c = cols[(long)r * deg + k]; if (c < half) c0[a++] = c; else c1[b++] = c; }
}
__global__ void iota_rp(int* rp, int nrow, int deg) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r <= nrow) rp[r] = r * deg;
}
__global__ void fill_x(double* x, long n) {   // small integers / 1024: every sum is exact, so the two summation orders agree to the bit
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) x[i] = (double)((i * 2654435761u) & 1023) / 1024.0;
}
__global__ void max_diff(const double* a, const double* b, long n, double* out) {
  double m = 0.0;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) m = fmax(m, fabs(a[i] - b[i]));
  if (m > 0.0) atomicMax(reinterpret_cast<unsigned long long*>(out), (unsigned long long)__double_as_longlong(m));
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

  // ---- step 3: the product itself
  if (argc > 2) {
    int n1 = 0;
    for (int i = 0; i < nsm; ++i) n1 += die[i];
    if (n1 < 16 || nsm - n1 < 16) { printf("{\"step\": \"product\", \"skipped\": \"no usable SM -> die map\"}\n"); return 0; }
    const int nrow = 10000000, ncol = 1000000, deg = 20; const long nnz = (long)nrow * deg;
    CK(cudaFree(idx0)); CK(cudaFree(idx1)); CK(cudaFree(tab));
    int *cols, *rpf, *rp0, *rp1, *cnt0, *cnt1, *c0, *c1, *ready; double *X, *Y, *Yref, *dmax;
    CK(cudaMalloc(&cols, nnz * 4)); CK(cudaMalloc(&rpf, (nrow + 1) * 4L)); CK(cudaMalloc(&rp0, (nrow + 1) * 4L)); CK(cudaMalloc(&rp1, (nrow + 1) * 4L));
    CK(cudaMalloc(&cnt0, (nrow + 1) * 4L)); CK(cudaMalloc(&cnt1, (nrow + 1) * 4L)); CK(cudaMalloc(&c0, nnz * 4)); CK(cudaMalloc(&c1, nnz * 4));
    CK(cudaMalloc(&ready, (nrow / kRB + 2) * 4L)); CK(cudaMemset(ready, 0, (nrow / kRB + 2) * 4L));
    CK(cudaMalloc(&X, (long)ncol * 32 * 8)); CK(cudaMalloc(&Y, (long)nrow * 32 * 8)); CK(cudaMalloc(&Yref, (long)nrow * 32 * 8)); CK(cudaMalloc(&dmax, 8));
    gen_cols<<<148 * 8, 256>>>(cols, nnz, ncol, 0xC2C2);
    fill_x<<<148 * 8, 256>>>(X, (long)ncol * 32);
    iota_rp<<<(nrow + 256) / 256, 256>>>(rpf, nrow, deg);
    CK(cudaMemset(cnt0, 0, (nrow + 1) * 4L)); CK(cudaMemset(cnt1, 0, (nrow + 1) * 4L));
    count_lo<<<(nrow + 255) / 256, 256>>>(cols, nrow, deg, ncol / 2, cnt0, cnt1);
    {
      void* tmp = nullptr; size_t tb = 0;
      cub::DeviceScan::ExclusiveSum(tmp, tb, cnt0, rp0, nrow + 1);
      CK(cudaMalloc(&tmp, tb));
      cub::DeviceScan::ExclusiveSum(tmp, tb, cnt0, rp0, nrow + 1);
      cub::DeviceScan::ExclusiveSum(tmp, tb, cnt1, rp1, nrow + 1);
      CK(cudaFree(tmp));
    }
    split_fill<<<(nrow + 255) / 256, 256>>>(cols, nrow, deg, ncol / 2, rp0, rp1, c0, c1);
    CK(cudaDeviceSynchronize());
    int epoch = 0;
    auto product = [&](int mode, double* Yo) {
      for (int pass = 0; pass < 2; ++pass) {
        ++epoch;
        CK(cudaMemsetAsync(ctr, 0, 16));
        if (mode == 0) spmm_q_kernel<6><<<148 * 5, 256>>>(nrow, rpf, cols, rpf, cols, X, Yo, pass * 16, d_die, 0, ctr, ready, epoch);
        else spmm_q_kernel<6><<<148 * 5, 256>>>(nrow, rp0, c0, rp1, c1, X, Yo, pass * 16, d_die, 1, ctr, ready, epoch);
      }
    };
    for (int mode : {0, 1, 0, 1}) {
      double* Yo = mode == 0 ? Yref : Y;
      product(mode, Yo); product(mode, Yo);
      CK(cudaDeviceSynchronize());
      float best = 1e30f;
      for (int it = 0; it < 4; ++it) {
        CK(cudaEventRecord(e0));
        product(mode, Yo);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        best = std::min(best, ms);
      }
      CK(cudaGetLastError());
      double hd = 0.0;
      if (mode == 1) {
        CK(cudaMemset(dmax, 0, 8));
        max_diff<<<148 * 8, 256>>>(Y, Yref, (long)nrow * 32, dmax);
        CK(cudaMemcpy(&hd, dmax, 8, cudaMemcpyDeviceToHost));
      }
      printf("{\"step\": \"product\", \"what\": \"binary CSR 10M x 1M, 20 per row, R = 32, two column passes, probe kernel (no TMA staging)\", "
             "\"assignment\": \"%s\", \"ms\": %.4f, \"max_abs_diff_vs_unsplit\": %.3g}\n",
             mode == 0 ? "any CTA, whole rows (today)" : "die-affine column halves, die 1 adds to die 0's partial rows", best, hd);
      fflush(stdout);
    }
    return 0;
  }
  const char* names[4] = {"any CTA, whole table (today)", "die-affine halves", "die-affine, halves swapped", "two queues, CTAs assigned by blockIdx parity (control)"};
  for (int ctas : {5, 8})
    for (long mb : {256L, 128L}) {
      for (int mode = 0; mode < 4; ++mode) run(names[mode], 8, mb, mode, ctas);     // 128-byte slabs of 256-byte rows (the two-pass product)
      for (int mode = 0; mode < 4; ++mode) run(names[mode], 16, mb, mode, ctas);    // whole rows (the one-pass product)
    }
  return 0;
}
