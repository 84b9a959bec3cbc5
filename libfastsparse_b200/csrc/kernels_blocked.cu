// kernels_blocked.cu -- row-blocked COO products (struct BlockedSBM sparse.h:163-172,
// struct BlockedSDM dsparse.h:119-129), the container of the Hilbert-ordered variant.
//
// Replaces bsbm_A_mul_B/_B2/_B4/_Bn (sparse.h:259-336) and bsdm_A_mul_B
// (dsparse.h:176-191).  The reference gives one row block to one CPU thread, which
// zeroes the block's Y slice and scatters Y[row,:] += X[col,:] in stored order.
//
// B200 design: one CTA per (row block, column slab).  The block's Y slab lives in
// SHARED MEMORY for the whole pass (rows_in_block x slab_cols doubles) and is written
// to HBM exactly once.  fp64 shared-memory atomics do not exist in hardware (they
// compile to a CAS spin loop), so instead of racing on rows the entries of a block are
// bucketed at upload time by ROW CLASS (local row mod 256, stable, see
// fsb_blocked_relayout): a team of G lanes owns the classes t, t+NT, ... and streams
// their entry lists in stored (e.g. Hilbert) order.  A row therefore has exactly one
// writer: no atomics, no races, and each Y element is summed in the reference's own
// order (bit-identical to the serial reference when R needs no column split).
// The team reads G (row, col) pairs with one coalesced load, broadcasts them by
// shuffle, gathers VEC doubles of the X row per lane and accumulates into shared memory.
#include <algorithm>

#include "fsb_device.cuh"
#include "fsb_internal.h"

using namespace fsbdev;

namespace {

constexpr int kThreads = 512;
constexpr int kClasses = FSB_BLOCKED_CLASSES;

template <int G, int VEC, bool VALS>
__global__ void __launch_bounds__(kThreads)
blocked_spmm_kernel(const int* __restrict__ start_row, const int* __restrict__ cls_ptr,
                    const int* __restrict__ rows, const int* __restrict__ cols, const double* __restrict__ vals,
                    const double* __restrict__ X, double* __restrict__ Y, int R, int slab_cols) {
  extern __shared__ double ys[];
  constexpr int NT = kThreads / G;
  constexpr int U = (G >= 4) ? 4 : G;
  const int b = blockIdx.x;
  const int col0 = blockIdx.y * slab_cols;
  const int ncols = min(slab_cols, R - col0);
  const int r0 = __ldg(start_row + b), nr = __ldg(start_row + b + 1) - r0;
  for (int i = threadIdx.x; i < nr * slab_cols; i += kThreads) ys[i] = 0.0;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int l = threadIdx.x & (G - 1);
  const int team = threadIdx.x / G;
  const unsigned tmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (lane - l));
  const bool col_ok = l * VEC < ncols;
  const double* xbase = X + col0 + l * VEC;
  for (int cls = team; cls < kClasses; cls += NT) {
    const int s = __ldg(cls_ptr + (long long)b * kClasses + cls);
    const int e = __ldg(cls_ptr + (long long)b * kClasses + cls + 1);
    for (int base = s; base < e; base += G) {
      const int idx = base + l;
      int myr = 0, myc = 0;
      double myv = 0.0;
      if (idx < e) {
        myr = ld_stream_s32(rows + idx);
        myc = ld_stream_s32(cols + idx);
        if (VALS) myv = ld_stream_f64(vals + idx);
      }
      const int n_here = min(G, e - base);
      for (int s0 = 0; s0 < n_here; s0 += U) {
        double xr[U][VEC];
        int rr[U];
        double vv[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int j = s0 + u;
          const int c = __shfl_sync(tmask, myc, j & (G - 1), G);
          rr[u] = __shfl_sync(tmask, myr, j & (G - 1), G) - r0;
          if (VALS) vv[u] = shfl_f64(tmask, myv, j & (G - 1), G);
          if (j < n_here && col_ok) {
            XLoad<VEC>::ld(xr[u], xbase + (long long)c * R);
          } else {
#pragma unroll
            for (int v = 0; v < VEC; ++v) xr[u][v] = 0.0;
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          if (s0 + u < n_here && col_ok) {
            double* yp = ys + rr[u] * slab_cols + l * VEC;
#pragma unroll
            for (int v = 0; v < VEC; ++v) yp[v] = VALS ? fma(xr[u][v], vv[u], yp[v]) : yp[v] + xr[u][v];
          }
        }
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nr * ncols; i += kThreads) {
    const int r = i / ncols, c = i - r * ncols;
    Y[(long long)(r0 + r) * R + col0 + c] = ys[r * slab_cols + c];
  }
}

// Fallback for row blocks too tall for shared memory: Y zeroed by the caller, entries
// scattered with fp64 red.global.add (order not deterministic).
template <bool VALS>
__global__ void __launch_bounds__(256)
blocked_scatter_kernel(const int* __restrict__ rows, const int* __restrict__ cols, const double* __restrict__ vals,
                       long long nnz, const double* __restrict__ X, double* __restrict__ Y, int R) {
  const int lane = threadIdx.x & 31;
  long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
  for (; w < nnz; w += nw) {
    const int r = __ldg(rows + w), c = __ldg(cols + w);
    const double v = VALS ? __ldg(vals + w) : 1.0;
    for (int k = lane; k < R; k += 32) red_add_f64(Y + (long long)r * R + k, __ldg(X + (long long)c * R + k) * v);
  }
}

inline int pow2_ceil(int x) { int p = 1; while (p < x) p <<= 1; return p; }

template <int G, int VEC, bool VALS>
int launch(const fsb_matrix* A, double* dY, const double* dX, int R, int slab_cols, size_t smem, cudaStream_t st) {
  auto kern = blocked_spmm_kernel<G, VEC, VALS>;
  if (smem > 48 * 1024) FSB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)A->nblocks, (unsigned)((R + slab_cols - 1) / slab_cols));
  kern<<<grid, kThreads, smem, st>>>(A->start_row, A->row_ptr, A->b_rows, A->b_cols, A->b_vals, dX, dY, R, slab_cols);
  return FSB_OK;
}

template <int G, int VEC>
int launch_vals(const fsb_matrix* A, double* dY, const double* dX, int R, int slab_cols, size_t smem, cudaStream_t st) {
  return A->has_vals ? launch<G, VEC, true>(A, dY, dX, R, slab_cols, smem, st) : launch<G, VEC, false>(A, dY, dX, R, slab_cols, smem, st);
}

template <int G>
int launch_vec(int vec, const fsb_matrix* A, double* dY, const double* dX, int R, int slab_cols, size_t smem, cudaStream_t st) {
  switch (vec) {
    case 1: return launch_vals<G, 1>(A, dY, dX, R, slab_cols, smem, st);
    case 2: return launch_vals<G, 2>(A, dY, dX, R, slab_cols, smem, st);
    default: return launch_vals<G, 4>(A, dY, dX, R, slab_cols, smem, st);
  }
}

}  // namespace

int fsb_launch_blocked_spmm(const fsb_matrix* A, double* dY, const double* dX, int R, cudaStream_t st) {
  if (R <= 0) return fsb_set_error(FSB_EINVAL, "blocked spmm: R must be positive (got %d)", R);
  if (A->nrow == 0 || A->nblocks == 0) return FSB_OK;
  const uintptr_t al = (uintptr_t)dX | (uintptr_t)dY;
  int vec = (R % 4 == 0 && al % 32 == 0) ? 4 : (R % 2 == 0 && al % 16 == 0) ? 2 : 1;
  // column slab: all of R when the block's Y fits in ~100 KB of shared memory (two CTAs per
  // SM), otherwise halve (keeping vector alignment) down to one vector per lane
  const size_t rows = (size_t)std::max(A->max_block_rows, 1);
  int slab = std::min(R, 32 * vec);
  const size_t want = 100 * 1024, hard = 220 * 1024;
  while (rows * slab * 8 > want && slab % (2 * vec) == 0) slab /= 2;
  if (rows * slab * 8 > hard) {
    while (vec > 1 && rows * slab * 8 > hard) { vec /= 2; while (rows * slab * 8 > hard && slab % (2 * vec) == 0) slab /= 2; }
  }
  if (rows * slab * 8 > hard) {
    FSB_CUDA(cudaMemsetAsync(dY, 0, (size_t)A->nrow * R * 8, st));
    if (A->nnz > 0) {
      const int blocks = (int)std::min<long long>(((long long)A->nnz * 32 + 255) / 256, 148LL * 32);
      if (A->has_vals) blocked_scatter_kernel<true><<<blocks, 256, 0, st>>>(A->b_rows, A->b_cols, A->b_vals, A->nnz, dX, dY, R);
      else blocked_scatter_kernel<false><<<blocks, 256, 0, st>>>(A->b_rows, A->b_cols, A->b_vals, A->nnz, dX, dY, R);
      FSB_KERNEL_CHECK();
    }
    return FSB_OK;
  }
  const size_t smem = rows * slab * 8;
  const int g = pow2_ceil((slab + vec - 1) / vec);
  int rc;
  switch (g) {
    case 1: rc = launch_vec<1>(vec, A, dY, dX, R, slab, smem, st); break;
    case 2: rc = launch_vec<2>(vec, A, dY, dX, R, slab, smem, st); break;
    case 4: rc = launch_vec<4>(vec, A, dY, dX, R, slab, smem, st); break;
    case 8: rc = launch_vec<8>(vec, A, dY, dX, R, slab, smem, st); break;
    case 16: rc = launch_vec<16>(vec, A, dY, dX, R, slab, smem, st); break;
    default: rc = launch_vec<32>(vec, A, dY, dX, R, slab, smem, st); break;
  }
  FSB_TRY(rc);
  FSB_KERNEL_CHECK();
  return FSB_OK;
}
