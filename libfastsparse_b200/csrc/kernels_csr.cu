// kernels_csr.cu -- CSR sparse x dense products on sm_100a.
//
// Replaces the OpenMP row loops of the reference:
//   bcsr_A_mul_B csr.h:149-161, _B2/_B4/_B8/_B8_auto csr.h:164-254,
//   bcsr_A_mul_Bn csr.h:257-280, bcsr_A_mul_B32n csr.h:283-302,
//   csr_A_mul_B csr.h:425-438, csr_A_mul_Bn csr.h:441-465,
//   bcsr_AA_mul_B / parallel_bcsr_AA_mul_B csr.h:305-355 (fused mode).
//
// This file holds (i) the dispatch of every CSR product (fsb_launch_csr_spmm: which kernel, how many
// column passes, which build -- partly decided by timing, once per handle), (ii) the first-generation
// team-per-row kernel, still the choice for one right-hand side on binary matrices with long regular
// rows, and (iii) the fused A'(A x) kernel.  The multi-RHS work horse is kernels_csr_staged.cu, the
// entry-balanced SpMV kernel kernels_csr_stream.cu.
//
// Team-per-row design (not a port: the reference gives each CPU thread whole rows and a private
// accumulator vector).  Here a TEAM of TW lanes owns one row.  The team streams the
// row's column indices with one coalesced load per TW entries, then hands each index
// to a SUB-GROUP of G lanes by warp shuffle; the G lanes gather G*VEC consecutive
// doubles of the dense operand row in ONE vector LDG per lane (VEC = 4 -> a single
// 256-bit LDG.E.256; at R = 32 eight lanes fetch a whole 256-byte X row and a warp
// has four X rows in flight per instruction).  TW/G sub-groups work on different
// nonzeros of the same row, so partial sums are combined by shuffle-xor at the end
// and sub-group 0 writes the Y row with a streaming store.  No atomics, no shared
// memory, deterministic.  HBM-bound: no tensor cores (arithmetic intensity
// <= 0.125 flop/B).
//
// Index arithmetic is 64-bit wherever a row index is multiplied by R (the host
// structs are int32; see SURVEY "32-bit indexing limits").
#include <algorithm>

#include "fsb_device.cuh"
#include "fsb_internal.h"

using namespace fsbdev;

namespace {

constexpr int kThreads = 256;

// compile-time experiment knobs (tools/variants.sh builds one .so per setting)
#ifndef FSB_SPMM_U
#define FSB_SPMM_U 4          // gathers in flight per lane
#endif
#ifndef FSB_SPMM_MINBLOCKS
#define FSB_SPMM_MINBLOCKS 1  // __launch_bounds__ min CTAs per SM (register cap)
#endif

template <int TW, int G, int VEC, bool VALS>
__global__ void __launch_bounds__(kThreads, FSB_SPMM_MINBLOCKS)
csr_spmm_kernel(int nrow, const int* __restrict__ row_ptr, const int* __restrict__ cols,
                const double* __restrict__ vals, const double* __restrict__ X,
                double* __restrict__ Y, int R, int col0, int ncols) {
  constexpr int NSUB = TW / G;            // sub-groups (nonzeros in flight) per team
  constexpr int U = (G >= FSB_SPMM_U) ? FSB_SPMM_U : G;     // gathers in flight per lane
  const long long gt = (long long)blockIdx.x * kThreads + threadIdx.x;
  const int row = (int)(gt / TW);
  if (row >= nrow) return;                // whole team leaves together
  const int lane = threadIdx.x & 31;
  const int tl = lane & (TW - 1);
  const int sub = tl / G;
  const int l = tl & (G - 1);
  const unsigned tmask = (TW == 32) ? 0xffffffffu : (((1u << TW) - 1u) << (lane - tl));
  const bool col_ok = l * VEC < ncols;
  const double* xbase = X + col0 + l * VEC;

  const int start = __ldg(row_ptr + row);
  const int end = __ldg(row_ptr + row + 1);
  double acc[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) acc[v] = 0.0;

  for (int base = start; base < end; base += TW) {
    const int idx = base + tl;
    int myc = 0;
    double myv = 0.0;
    if (idx < end) {
      myc = ld_stream_s32(cols + idx);
      if (VALS) myv = ld_stream_f64(vals + idx);
    }
    const int n_here = min(TW, end - base);
    for (int s0 = 0; s0 < n_here; s0 += NSUB * U) {
      double xr[U][VEC];
      double vv[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int j = s0 + u * NSUB + sub;
        const int c = __shfl_sync(tmask, myc, j & (TW - 1), TW);
        if (VALS) vv[u] = shfl_f64(tmask, myv, j & (TW - 1), TW);
        if (j < n_here && col_ok) {
          XLoad<VEC>::ld(xr[u], xbase + (long long)c * R);
        } else {
#pragma unroll
          for (int v = 0; v < VEC; ++v) xr[u][v] = 0.0;
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u)
#pragma unroll
        for (int v = 0; v < VEC; ++v)
          acc[v] = VALS ? fma(xr[u][v], vv[u], acc[v]) : acc[v] + xr[u][v];
    }
  }
#pragma unroll
  for (int off = G; off < TW; off <<= 1)
#pragma unroll
    for (int v = 0; v < VEC; ++v) acc[v] += shfl_xor_f64(tmask, acc[v], off, TW);
  if (sub == 0 && col_ok) YStore<VEC>::st(Y + (long long)row * R + col0 + l * VEC, acc);
}

// Fused y = A'(A x) (+ the caller pre-loads Y with lambda*X): per row, gather-sum
// xv = sum X[c,:], then scatter xv to Y[c,:] with fp64 red.global.add at L2.
// Order of the scatter adds is not deterministic; results agree with the two-pass
// mode to ~1e-15 relative per add (SURVEY "Determinism vs tolerance").
template <int TW, int G, bool VALS>
__global__ void __launch_bounds__(kThreads)
csr_ata_fused_kernel(int nrow, const int* __restrict__ row_ptr, const int* __restrict__ cols,
                     const double* __restrict__ vals, const double* __restrict__ X,
                     double* __restrict__ Y, int R) {
  constexpr int NSUB = TW / G;
  const long long gt = (long long)blockIdx.x * kThreads + threadIdx.x;
  const int row = (int)(gt / TW);
  if (row >= nrow) return;
  const int lane = threadIdx.x & 31;
  const int tl = lane & (TW - 1);
  const int sub = tl / G;
  const int l = tl & (G - 1);
  const unsigned tmask = (TW == 32) ? 0xffffffffu : (((1u << TW) - 1u) << (lane - tl));
  const bool col_ok = l < R;
  const int start = __ldg(row_ptr + row);
  const int end = __ldg(row_ptr + row + 1);
  double acc = 0.0;
  for (int base = start; base < end; base += TW) {
    const int idx = base + tl;
    int myc = 0;
    double myv = 0.0;
    if (idx < end) {
      myc = __ldg(cols + idx);          // re-read by the scatter phase: let L1 keep it
      if (VALS) myv = __ldg(vals + idx);
    }
    const int n_here = min(TW, end - base);
    for (int s0 = 0; s0 < n_here; s0 += NSUB) {
      const int j = s0 + sub;
      const int c = __shfl_sync(tmask, myc, j & (TW - 1), TW);
      const double v = VALS ? shfl_f64(tmask, myv, j & (TW - 1), TW) : 1.0;
      if (j < n_here && col_ok) acc = fma(__ldg(X + (long long)c * R + l), v, acc);
    }
  }
#pragma unroll
  for (int off = G; off < TW; off <<= 1) acc += shfl_xor_f64(tmask, acc, off, TW);
  for (int base = start; base < end; base += TW) {
    const int idx = base + tl;
    int myc = 0;
    double myv = 0.0;
    if (idx < end) {
      myc = __ldg(cols + idx);
      if (VALS) myv = __ldg(vals + idx);
    }
    const int n_here = min(TW, end - base);
    for (int s0 = 0; s0 < n_here; s0 += NSUB) {
      const int j = s0 + sub;
      const int c = __shfl_sync(tmask, myc, j & (TW - 1), TW);
      const double v = VALS ? shfl_f64(tmask, myv, j & (TW - 1), TW) : 1.0;
      if (j < n_here && col_ok) red_add_f64(Y + (long long)c * R + l, acc * v);
    }
  }
}

__global__ void scale_copy_kernel(double* __restrict__ Y, const double* __restrict__ X, double lambda, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) Y[i] = lambda * X[i];
}

// tuning overrides are per calling thread (tools and tests set them on the thread that runs the products): a tuning
// call from one thread never changes another thread's products (bench_a_mul_b.c:400-421 calls from two threads)
thread_local int g_tw = 0, g_g = 0, g_vec = 0, g_slabs = 0;
thread_local int g_algo = 0;   // 0 = automatic, 1 = team-per-row kernel (this file), 2 = staged row-block kernel,
                  // 3 = merge-path stream kernel where it applies (R = 1, 2, 4), staged otherwise

inline int pow2_ceil(int x) { int p = 1; while (p < x) p <<= 1; return p; }
inline int pow2_floor(int x) { int p = 1; while (p * 2 <= x) p <<= 1; return p; }

template <int TW, int G, int VEC, bool VALS>
void launch_spmm(const fsb_matrix* A, double* dY, const double* dX, int R, int col0, int ncols, cudaStream_t st) {
  const long long threads = (long long)A->nrow * TW;
  const unsigned grid = (unsigned)((threads + kThreads - 1) / kThreads);
  csr_spmm_kernel<TW, G, VEC, VALS><<<grid, kThreads, 0, st>>>(A->nrow, A->row_ptr, A->cols, A->vals, dX, dY, R, col0, ncols);
}

template <int TW, int G, bool VALS>
void launch_fused(const fsb_matrix* A, double* dY, const double* dX, int R, cudaStream_t st) {
  const long long threads = (long long)A->nrow * TW;
  const unsigned grid = (unsigned)((threads + kThreads - 1) / kThreads);
  csr_ata_fused_kernel<TW, G, VALS><<<grid, kThreads, 0, st>>>(A->nrow, A->row_ptr, A->cols, A->vals, dX, dY, R);
}

template <int TW, int G, int VEC>
bool dispatch_vals(const fsb_matrix* A, double* dY, const double* dX, int R, int col0, int ncols, cudaStream_t st) {
  if (A->has_vals) launch_spmm<TW, G, VEC, true>(A, dY, dX, R, col0, ncols, st);
  else launch_spmm<TW, G, VEC, false>(A, dY, dX, R, col0, ncols, st);
  return true;
}

template <int TW, int G>
bool dispatch_vec(int vec, const fsb_matrix* A, double* dY, const double* dX, int R, int col0, int ncols, cudaStream_t st) {
  switch (vec) {
    case 1: return dispatch_vals<TW, G, 1>(A, dY, dX, R, col0, ncols, st);
    case 2: return dispatch_vals<TW, G, 2>(A, dY, dX, R, col0, ncols, st);
    case 4: return dispatch_vals<TW, G, 4>(A, dY, dX, R, col0, ncols, st);
  }
  return false;
}

#define FSB_TWG_TABLE(X) \
  X(1, 1) X(2, 1) X(4, 1) X(8, 1) X(16, 1) X(32, 1) \
  X(2, 2) X(4, 2) X(8, 2) X(16, 2) X(32, 2)         \
  X(4, 4) X(8, 4) X(16, 4) X(32, 4)                 \
  X(8, 8) X(16, 8) X(32, 8)                         \
  X(16, 16) X(32, 16)                               \
  X(32, 32)

bool dispatch_spmm(int tw, int g, int vec, const fsb_matrix* A, double* dY, const double* dX, int R, int col0, int ncols, cudaStream_t st) {
#define X(TW_, G_) if (tw == TW_ && g == G_) return dispatch_vec<TW_, G_>(vec, A, dY, dX, R, col0, ncols, st);
  FSB_TWG_TABLE(X)
#undef X
  return false;
}

bool dispatch_fused(int tw, int g, const fsb_matrix* A, double* dY, const double* dX, int R, cudaStream_t st) {
#define X(TW_, G_)                                                      \
  if (tw == TW_ && g == G_) {                                           \
    if (A->has_vals) launch_fused<TW_, G_, true>(A, dY, dX, R, st);     \
    else launch_fused<TW_, G_, false>(A, dY, dX, R, st);                \
    return true;                                                        \
  }
  FSB_TWG_TABLE(X)
#undef X
  return false;
}

// team width from the mean row length: enough sub-groups to cover a row in a couple
// of steps without leaving most of them idle
int pick_tw(int g, double avg_nnz) {
  // wide gathers (g >= 8): two sub-groups per ~20-entry row measured best on B200
  // (profiles/r1_sweep_c2.md); narrow gathers want more lanes on the row's index stream
  const double per_sub = g >= 8 ? 8.0 : 4.0;
  int want = pow2_floor(std::max(1, (int)(avg_nnz / per_sub)));
  int ns = std::min(want, 32 / g);
  return g * std::max(ns, 1);
}

}  // namespace

void fsb_csr_spmm_set_tuning(int tw, int g, int vec, int slabs) {
  g_tw = tw; g_g = g; g_vec = vec; g_slabs = slabs;
}

extern "C" int fsb_tune_csr_algo(int algo, int rows_per_cta, int cap_mult) {
  if (algo < 0 || algo > 3 || rows_per_cta < 0 || cap_mult < 0) return fsb_set_error(FSB_EINVAL, "fsb_tune_csr_algo: bad argument");
  g_algo = algo;
  fsb_csr_staged_set_tuning(rows_per_cta, cap_mult);
  return FSB_OK;
}

thread_local int g_deep = -1;   // staged kernel build: -1 automatic (timed per handle), 0 lean, 1 deep

extern "C" int fsb_tune_csr_staged(int deep) {
  if (deep < -1 || deep > 1) return fsb_set_error(FSB_EINVAL, "fsb_tune_csr_staged: deep must be -1, 0 or 1");
  g_deep = deep;
  return FSB_OK;
}

extern "C" int fsb_tune_csr_spmm(int tw, int g, int vec, int slabs) {
  auto pow2 = [](int x) { return x == 0 || (x > 0 && x <= 32 && (x & (x - 1)) == 0); };
  if (!pow2(tw) || !pow2(g) || !(vec == 0 || vec == 1 || vec == 2 || vec == 4) || slabs < 0)
    return fsb_set_error(FSB_EINVAL, "fsb_tune_csr_spmm: tw, g must be 0 or a power of two <= 32; vec 0/1/2/4");
  fsb_csr_spmm_set_tuning(tw, g, vec, slabs);
  return FSB_OK;
}

namespace {

__global__ void max_row_kernel(const int* __restrict__ row_ptr, int nrow, int* __restrict__ out) {
  int m = 0;
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < nrow; r += gridDim.x * blockDim.x)
    m = max(m, __ldg(row_ptr + r + 1) - __ldg(row_ptr + r));
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, off));
  if ((threadIdx.x & 31) == 0 && m > 0) atomicMax(out, m);
}

// longest row of the matrix, computed once per handle (drives the SpMV kernel choice)
int max_row_nnz(fsb_matrix* A, cudaStream_t st, int* out) {
  if (A->max_row_nnz < 0) {
    int* d = nullptr;
    FSB_CUDA(cudaMalloc(&d, sizeof(int)));
    cudaError_t e = cudaMemsetAsync(d, 0, sizeof(int), st);
    if (e == cudaSuccess) {
      max_row_kernel<<<148 * 4, 256, 0, st>>>(A->row_ptr, A->nrow, d);
      fsb_count_launch();
      int h = 0;
      e = cudaMemcpyAsync(&h, d, sizeof(int), cudaMemcpyDeviceToHost, st);
      if (e == cudaSuccess) e = cudaStreamSynchronize(st);
      if (e == cudaSuccess) A->max_row_nnz = h;
    }
    cudaFree(d);
    if (e != cudaSuccess) return fsb_cuda_error(e, "max_row_nnz", __FILE__, __LINE__);
  }
  *out = A->max_row_nnz;
  return FSB_OK;
}

// one product with a fixed configuration (column passes of per_pass columns)
int run_config(fsb_matrix* A, double* dY, const double* dX, int R, int algo, int vec, int per_pass, int g_override,
               int tw_override, cudaStream_t st, const double* dZ, double lambda, bool deep = false) {
  int g = pow2_ceil((per_pass + vec - 1) / vec);
  if (g_override >= g && g_override <= 32) g = g_override;
  int tw = pick_tw(g, A->avg_row_nnz);
  if (tw_override >= g && tw_override <= 32) tw = tw_override;
  for (int col0 = 0; col0 < R; col0 += per_pass) {
    const int ncols = std::min(per_pass, R - col0);
    if (algo == 2) {
      FSB_TRY(fsb_launch_csr_spmm_staged(A, dY, dX, R, col0, ncols, g, vec, st, dZ, lambda, deep));
      continue;
    }
    if (!dispatch_spmm(tw, g, vec, A, dY, dX, R, col0, ncols, st))
      return fsb_set_error(FSB_EINVAL, "spmm: no kernel for TW=%d G=%d VEC=%d", tw, g, vec);
    FSB_KERNEL_CHECK();
  }
  if (algo != 2 && dZ) FSB_TRY(fsb_dense_axpy_lambda(dY, dZ, lambda, (long)A->nrow * R, st));   // no fused epilogue in this kernel
  return FSB_OK;
}

// X [ncol][R] -> S contiguous column slabs [S][ncol][R/S]: one pass over the operand (2 x 8 ncol R bytes)
__global__ void repack_slabs_kernel(const double2* __restrict__ X, double2* __restrict__ Xs, long long ncol, int R2, int w2) {
  const long long n = ncol * R2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long c = i / R2;
    const int j = (int)(i - c * R2), s = j / w2, k = j - s * w2;
    Xs[(long long)s * ncol * w2 + c * w2 + k] = X[i];
  }
}

// The product as S column passes over CONTIGUOUS slabs of the dense operand.  A pass over columns [s w, (s+1) w) of
// the row-major operand touches w*8 bytes out of every 8R-byte row: its L2 footprint is counted in 128-byte LINES,
// so a 64-byte slab of a 256-byte row still occupies a full line (and the slab spans the whole 8 ncol R address
// range).  Repacked, slab s is ncol*w*8 contiguous bytes: every line it occupies is fully used and a 64 MB slab
// really is 64 MB of L2.  Costs one streaming pass over the operand per product.
int run_repacked(fsb_matrix* A, double* dY, const double* dX, int R, int slabs, cudaStream_t st, const double* dZ, double lambda, bool deep) {
  fsb_matrix* S = A->scratch_owner ? A->scratch_owner : A;
  const int w = R / slabs;
  const size_t need = (size_t)A->ncol * R * sizeof(double);
  if (need > S->xpack_cap) {
    if (S->xpack) cudaFree(S->xpack);
    S->xpack = nullptr; S->xpack_cap = 0; S->xpack_src = nullptr;
    FSB_CUDA(cudaMalloc(&S->xpack, need));
    S->xpack_cap = need;
  }
  if (S->xpack_src != dX) {
    const long long n = (long long)A->ncol * (R / 2);
    const int blocks = (int)std::min<long long>((n + 255) / 256, 148 * 16);
    repack_slabs_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<const double2*>(dX), reinterpret_cast<double2*>(S->xpack), A->ncol, R / 2, w / 2);
    FSB_KERNEL_CHECK();
  }
  const int vec = 2, g = pow2_ceil(w / vec);
  for (int s = 0; s < slabs; ++s)
    FSB_TRY(fsb_launch_csr_spmm_staged(A, dY, S->xpack + (size_t)s * A->ncol * w, R, s * w, w, g, vec, st, dZ, lambda, deep, w, 0));
  return FSB_OK;
}

int stream_config(fsb_matrix* A, double* dY, const double* dX, int R, cudaStream_t st, const double* dZ, double lambda) {
  FSB_TRY(fsb_launch_csr_stream(A, dY, dX, R, st));
  if (dZ) FSB_TRY(fsb_dense_axpy_lambda(dY, dZ, lambda, (long)A->nrow * R, st));
  return FSB_OK;
}

}  // namespace

int fsb_launch_csr_spmm(fsb_matrix* A, double* dY, const double* dX, int R, cudaStream_t st, const double* dZ, double lambda) {
  if (R <= 0) return fsb_set_error(FSB_EINVAL, "spmm: R must be positive (got %d)", R);
  if (A->nrow == 0) return FSB_OK;
  const uintptr_t al = (uintptr_t)dX | (uintptr_t)dY | (uintptr_t)dZ;
  // ---- one right-hand side: x is L2-resident and every entry is one random 8-byte gather, so the product is
  // bound by the L1TEX / L2 request rate; what differs between the kernels is the overhead around the gathers
  // (profiles/r1m_sweep_spmv.md):
  //  * binary, short regular rows: the staged row-block kernel with two lanes per row (the second lane idles):
  //    0.75 ms at C3's structure against 0.91 ms for the team-per-row kernel and 0.90 ms for the stream kernel;
  //  * binary, long regular rows (transposes: 200 entries per row): a warp per row, 1.02 ms against 1.28 ms (stream);
  //  * matrices with values: the entry-balanced merge-path stream kernel, 1.08 / 1.29 ms (A / A') against
  //    1.13 / 1.40 ms for the team-per-row kernel;
  //  * skewed rows (a row far longer than the mean, e.g. transposes of power-law matrices): the stream kernel
  //    whatever the values (230 ms -> 1.9 ms).
  // row-range aliases (chunked host products, the sharded CG's column chunks) stay on the row-local kernels: the
  // merge-path kernel assumes row_ptr[0] == 0 and keeps per-handle tile tables
  const bool alias = A->scratch_owner != nullptr;
  if (R == 1 && !alias && (g_algo == 0 || g_algo == 3)) {
    bool stream = g_algo == 3 || A->has_vals;
    if (!stream) {
      int mx = 0;
      FSB_TRY(max_row_nnz(A, st, &mx));
      stream = mx > std::max(4096.0, 64.0 * A->avg_row_nnz);
    }
    // binary cell rows of an x-blocked transpose (~67 entries each at C3, x slice L2-resident): the TMA-fed stream kernel,
    // 0.851 against 0.890 ms for a warp per row (profiles/r2z_binary_t_probe.jsonl); with the whole 80 MB x behind the
    // gathers the warp-per-row kernel stays ahead (1.04 against 1.26 ms)
    if (!stream && g_algo == 0 && A->x_live_bytes > 0 && A->avg_row_nnz > 48.0) stream = true;
    if (stream) return stream_config(A, dY, dX, R, st, dZ, lambda);
    if (A->avg_row_nnz <= 48.0) return run_config(A, dY, dX, 1, 2, 1, 1, 2, g_tw, st, dZ, lambda, false);
    return run_config(A, dY, dX, 1, 1, 1, 1, 0, g_tw, st, dZ, lambda);
  }
  if (g_algo == 3 && !alias && fsb_csr_stream_supports(R)) return stream_config(A, dY, dX, R, st, dZ, lambda);
  const int algo = (g_algo == 1) ? 1 : 2;
  // 128-bit gathers (16 lanes per 256-byte X row) beat the 256-bit form on B200 for the staged
  // kernel at R = 32 (profiles/r1b_sweep_c2_staged_vs_team.json); the 256-bit form is kept for
  // wide operands where it halves the number of column passes.  Narrow operands: at least two
  // lanes per row (R = 2 -> 2 x 1, R = 4 -> 2 x 2) measured 1.3-2x faster than one wide lane.
  int vec = (R % 4 == 0 && al % 32 == 0 && (algo == 1 || R > 64)) ? 4 : (R % 2 == 0 && R >= 4 && al % 16 == 0) ? 2 : 1;
  if (g_vec && (R % g_vec == 0) && (al % (8 * g_vec) == 0)) vec = g_vec;
  int per_pass = std::min(R, 32 * vec);
  {   // experiment knob: S contiguous (repacked) column slabs
    const int xs = fsb_knob("x_slabs", 0);
    if (algo == 2 && xs >= 2 && R % (2 * xs) == 0 && R <= 64 && al % 16 == 0)
      return run_repacked(A, dY, dX, R, xs, st, dZ, lambda, g_deep > 0);
  }
  if (g_slabs >= 1) {   // explicit choice (tools/sweep.py)
    if (g_slabs > 1 && per_pass % (g_slabs * vec) == 0) per_pass /= g_slabs;
    return run_config(A, dY, dX, R, algo, vec, per_pass, g_g, g_tw, st, dZ, lambda, g_deep > 0);
  }
  // ---- automatic.  Two choices are not predictable from the shape, so the first product on a
  // handle (per R) times the candidates (each run twice, the second one timed) -- every one produces
  // the identical result, each column's sum is taken in the same order -- and the handle remembers the fastest:
  //  * column passes: when the dense operand does not fit in L2, two passes of >= 128 B per gather
  //    halve the per-pass footprint (more L2 hits) at the price of streaming the indices twice:
  //    faster on uniform columns, slower on power-law columns (whose hot rows hit L2 anyway);
  //  * lean or deep build of the staged kernel (kernels_csr_staged.cu): occupancy vs gathers in
  //    flight per lane.
  const bool big = algo == 2 && A->nnz >= (1 << 22);
  if (!big) return run_config(A, dY, dX, R, algo, vec, per_pass, g_g, g_tw, st, dZ, lambda, g_deep > 0);
  const bool two_pass_ok = (double)A->ncol * R * 8.0 > 126e6 && per_pass % (2 * vec) == 0 && per_pass / 2 * 8 >= 128;
  const fsb_matrix::Tuned* tuned = fsb_tuned_find(A, R);
  if (!tuned) {
    struct Cand { int passes; bool deep; float ms; };
    Cand cand[4];
    int nc = 0;
    for (int passes = 1; passes <= (two_pass_ok ? 2 : 1); ++passes)
      for (int deep = 0; deep <= 1; ++deep)
        if (g_deep < 0 || g_deep == deep) cand[nc++] = {passes, deep != 0, 0.f};
    struct Events {   // released on every path out of this block
      cudaEvent_t ev[8] = {};
      ~Events() { for (auto& e : ev) if (e) cudaEventDestroy(e); }
    } evs;
    cudaEvent_t* ev = evs.ev;
    for (int i = 0; i < 8; ++i) FSB_CUDA(cudaEventCreate(&ev[i]));
    int rc = FSB_OK;
    for (int k = 0; k < nc && rc == FSB_OK; ++k) {   // every candidate twice, the second run is the one timed
      rc = run_config(A, dY, dX, R, algo, vec, per_pass / cand[k].passes, g_g, g_tw, st, dZ, lambda, cand[k].deep);
      cudaEventRecord(ev[2 * k], st);
      if (rc == FSB_OK) rc = run_config(A, dY, dX, R, algo, vec, per_pass / cand[k].passes, g_g, g_tw, st, dZ, lambda, cand[k].deep);
      cudaEventRecord(ev[2 * k + 1], st);
    }
    if (rc == FSB_OK && cudaEventSynchronize(ev[2 * nc - 1]) == cudaSuccess) {
      int best = 0;
      for (int k = 0; k < nc; ++k) {
        cudaEventElapsedTime(&cand[k].ms, ev[2 * k], ev[2 * k + 1]);
        if (cand[k].ms < 0.99f * cand[best].ms) best = k;   // later candidates must win by 1 %
      }
      fsb_tuned_store(A, R, cand[best].passes, cand[best].deep ? 1 : 0);
    }
    return rc;
  }
  return run_config(A, dY, dX, R, algo, vec, per_pass / tuned->passes, g_g, g_tw, st, dZ, lambda, tuned->deep != 0);
}

int fsb_launch_csr_spmm_halves(fsb_matrix* A, double* dY, const double* dXlo, const double* dXhi, int R, cudaStream_t st,
                               cudaEvent_t ready_lo, cudaEvent_t ready_hi) {
  const int half = R / 2;
  if (R % 2 || half * 8 < 128 || half > 64 || (((uintptr_t)dXlo | (uintptr_t)dXhi | (uintptr_t)dY) & 15))
    return fsb_set_error(FSB_EINVAL, "spmm (column halves): unsupported width or alignment (R=%d)", R);
  if (A->nrow == 0) return FSB_OK;
  const int vec = 2, g = pow2_ceil(half / vec);
  const fsb_matrix::Tuned* tuned = fsb_tuned_find(A, R);
  const bool deep = tuned ? tuned->deep != 0 : true;
  if (ready_lo) FSB_CUDA(cudaStreamWaitEvent(st, ready_lo, 0));
  FSB_TRY(fsb_launch_csr_spmm_staged(A, dY, dXlo, R, 0, half, g, vec, st, nullptr, 0.0, deep, half, 0));
  if (ready_hi) FSB_CUDA(cudaStreamWaitEvent(st, ready_hi, 0));
  FSB_TRY(fsb_launch_csr_spmm_staged(A, dY, dXhi, R, half, half, g, vec, st, nullptr, 0.0, deep, half, 0));
  return FSB_OK;
}

int fsb_launch_csr_ata_fused(const fsb_matrix* A, double* dY, const double* dX, int R, double lambda, cudaStream_t st) {
  if (R <= 0 || R > 32) return fsb_set_error(FSB_EINVAL, "fused A'A: R must be 1..32 (got %d)", R);
  const long long n = (long long)A->ncol * R;
  if (n > 0) {
    const int blocks = (int)std::min<long long>((n + 255) / 256, 148 * 16);
    scale_copy_kernel<<<blocks, 256, 0, st>>>(dY, dX, lambda, n);
    FSB_KERNEL_CHECK();
  }
  if (A->nrow == 0) return FSB_OK;
  const int g = pow2_ceil(R);
  const int tw = pick_tw(g, A->avg_row_nnz);
  if (!dispatch_fused(tw, g, A, dY, dX, R, st)) return fsb_set_error(FSB_EINVAL, "fused A'A: no kernel for TW=%d G=%d", tw, g);
  FSB_KERNEL_CHECK();
  return FSB_OK;
}
