// kernels_cbcsr.cu -- column-blocked binary CSR (struct ColBinaryCSR, cbcsr.h:5-14).
//
// Replaces cbcsr_A_mul_B (cbcsr.h:76-106) and adds the n-RHS product the reference
// lacks.  The reference walks cells block-major with a private
// y copy per thread and merges under `omp critical`; column blocking exists there to
// keep an x slice cache-resident.  On the GPU a warp owns one ROW: it reads the bounds
// of that row's cells (cell = block*nrow + row) 32 blocks at a time, compacts the
// cells' column indices into a small per-warp shared-memory list with a warp prefix
// sum, and then consumes the list exactly like a CSR row (sub-groups of G lanes gather
// G*VEC doubles of the dense row per LDG).  Y is written once; no atomics, no
// per-thread copies, deterministic.
#include <algorithm>

#include "fsb_device.cuh"
#include "fsb_internal.h"

using namespace fsbdev;

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kCap = 256;  // column indices staged per warp

template <int G, int VEC>
__device__ __forceinline__ void consume_list(const int* __restrict__ list, int n, double (&acc)[VEC],
                                             const double* __restrict__ xbase, int R, int sub, bool col_ok) {
  constexpr int NSUB = 32 / G;
  constexpr int U = (G >= 4) ? 4 : G;
  for (int s0 = 0; s0 < n; s0 += NSUB * U) {
    double xr[U][VEC];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int j = s0 + u * NSUB + sub;
      if (j < n && col_ok) {
        XLoad<VEC>::ld(xr[u], xbase + (long long)list[j] * R);
      } else {
#pragma unroll
        for (int v = 0; v < VEC; ++v) xr[u][v] = 0.0;
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int v = 0; v < VEC; ++v) acc[v] += xr[u][v];
  }
}

template <int G, int VEC>
__global__ void __launch_bounds__(kThreads)
cbcsr_spmm_kernel(int nrow, int nblocks, const int* __restrict__ row_ptr, const int* __restrict__ cols,
                  const double* __restrict__ X, double* __restrict__ Y, int R, int col0, int ncols) {
  __shared__ int sbuf[kWarps][kCap];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * kWarps + warp;
  if (row >= nrow) return;
  const int sub = lane / G, l = lane & (G - 1);
  const bool col_ok = l * VEC < ncols;
  const double* xbase = X + col0 + l * VEC;
  int* list = sbuf[warp];
  double acc[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) acc[v] = 0.0;
  int count = 0;
  for (int b0 = 0; b0 < nblocks; b0 += 32) {
    const int b = b0 + lane;
    int s = 0, len = 0;
    if (b < nblocks) {
      const long long cell = (long long)b * nrow + row;
      s = __ldg(row_ptr + cell);
      len = __ldg(row_ptr + cell + 1) - s;
    }
    int incl = len;  // warp inclusive prefix sum
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, off);
      if (lane >= off) incl += t;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    if (count + total > kCap) {  // flush what is staged
      __syncwarp();
      consume_list<G, VEC>(list, count, acc, xbase, R, sub, col_ok);
      __syncwarp();
      count = 0;
    }
    if (total <= kCap) {
      int dst = count + incl - len;
      for (int i = 0; i < len; ++i) list[dst + i] = ld_stream_s32(cols + s + i);
      count += total;
    } else {
      // a very long row: stream each cell through the list in pieces, all lanes copying
      for (int k = 0; k < 32; ++k) {
        const int ks = __shfl_sync(0xffffffffu, s, k);
        const int kl = __shfl_sync(0xffffffffu, len, k);
        for (int p = 0; p < kl; p += kCap) {
          const int m = min(kCap, kl - p);
          for (int i = lane; i < m; i += 32) list[i] = ld_stream_s32(cols + ks + p + i);
          __syncwarp();
          consume_list<G, VEC>(list, m, acc, xbase, R, sub, col_ok);
          __syncwarp();
        }
      }
    }
  }
  __syncwarp();
  consume_list<G, VEC>(list, count, acc, xbase, R, sub, col_ok);
#pragma unroll
  for (int off = G; off < 32; off <<= 1)
#pragma unroll
    for (int v = 0; v < VEC; ++v) acc[v] += shfl_xor_f64(0xffffffffu, acc[v], off, 32);
  if (sub == 0 && col_ok) YStore<VEC>::st(Y + (long long)row * R + col0 + l * VEC, acc);
}

inline int pow2_ceil(int x) { int p = 1; while (p < x) p <<= 1; return p; }

template <int G>
void launch_vec(int vec, const fsb_matrix* A, double* dY, const double* dX, int R, int col0, int ncols, cudaStream_t st) {
  const unsigned grid = (unsigned)((A->nrow + kWarps - 1) / kWarps);
  switch (vec) {
    case 1: cbcsr_spmm_kernel<G, 1><<<grid, kThreads, 0, st>>>(A->nrow, A->nblocks, A->row_ptr, A->cols, dX, dY, R, col0, ncols); break;
    case 2: cbcsr_spmm_kernel<G, 2><<<grid, kThreads, 0, st>>>(A->nrow, A->nblocks, A->row_ptr, A->cols, dX, dY, R, col0, ncols); break;
    default: cbcsr_spmm_kernel<G, 4><<<grid, kThreads, 0, st>>>(A->nrow, A->nblocks, A->row_ptr, A->cols, dX, dY, R, col0, ncols); break;
  }
}

}  // namespace

int fsb_launch_cbcsr_spmm(const fsb_matrix* A, double* dY, const double* dX, int R, cudaStream_t st) {
  if (R <= 0) return fsb_set_error(FSB_EINVAL, "cbcsr spmm: R must be positive (got %d)", R);
  if (A->nrow == 0) return FSB_OK;
  const uintptr_t al = (uintptr_t)dX | (uintptr_t)dY;
  const int vec = (R % 4 == 0 && al % 32 == 0) ? 4 : (R % 2 == 0 && al % 16 == 0) ? 2 : 1;
  const int per_pass = std::min(R, 32 * vec);
  const int g = pow2_ceil((per_pass + vec - 1) / vec);
  for (int col0 = 0; col0 < R; col0 += per_pass) {
    const int ncols = std::min(per_pass, R - col0);
    switch (g) {
      case 1: launch_vec<1>(vec, A, dY, dX, R, col0, ncols, st); break;
      case 2: launch_vec<2>(vec, A, dY, dX, R, col0, ncols, st); break;
      case 4: launch_vec<4>(vec, A, dY, dX, R, col0, ncols, st); break;
      case 8: launch_vec<8>(vec, A, dY, dX, R, col0, ncols, st); break;
      case 16: launch_vec<16>(vec, A, dY, dX, R, col0, ncols, st); break;
      default: launch_vec<32>(vec, A, dY, dX, R, col0, ncols, st); break;
    }
    FSB_KERNEL_CHECK();
  }
  return FSB_OK;
}
