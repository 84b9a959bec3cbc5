// kernels_build.cu -- device-side structure construction and synthetic inputs.
//
// (1) COO -> CSR on the device by a STABLE radix sort on the row index.  The
//     reference's new_bcsr / new_csr (csr.h:30-67, 375-422) is a stable counting
//     sort, so the result (row_ptr, in-row order, duplicates) is bit-identical.
//     This is construction, not the hot path: the sort itself is cub::DeviceRadixSort.
// (2) CSR transpose (the reference obtains A' products by building a second CSR of
//     the transposed COO: bench_a_mul_b.c:273-274, test_sparse.c:564-565).
// (3) Counter-based synthetic COO generator shared bit-for-bit with the host
//     (SURVEY 8d): entry j is a pure function of (seed, j).
#include <stdlib.h>

#include <vector>

#include <cub/device/device_radix_sort.cuh>

#include "fsb_internal.h"
#include "fsb_synth.h"

namespace {

__global__ void iota_kernel(int* p, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) p[i] = (int)i;
}

// row_ptr from the sorted key stream: position j opens every row in (key[j-1], key[j]]
__global__ void row_ptr_from_sorted_kernel(const int* __restrict__ keys, long long n, int nrow, int* __restrict__ row_ptr) {
  long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; j <= n; j += stride) {
    const int lo = (j == 0) ? -1 : keys[j - 1];
    const int hi = (j == n) ? nrow : keys[j];
    for (int q = lo + 1; q <= hi; ++q) row_ptr[q] = (int)j;
  }
}

__global__ void gather_i32_kernel(int* __restrict__ out, const int* __restrict__ in, const int* __restrict__ perm, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) out[i] = in[perm[i]];
}
__global__ void gather_f64_kernel(double* __restrict__ out, const double* __restrict__ in, const int* __restrict__ perm, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) out[i] = in[perm[i]];
}

// row id of every stored entry of a CSR (binary search in row_ptr)
__global__ void expand_rows_kernel(const int* __restrict__ row_ptr, int nrow, long long nnz, int* __restrict__ out) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < nnz; i += stride) {
    int lo = 0, hi = nrow;  // largest r with row_ptr[r] <= i
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (row_ptr[mid] <= i) lo = mid; else hi = mid;
    }
    out[i] = lo;
  }
}

__global__ void synth_kernel(unsigned long long seed, int dist, long long nnz, int nrow, int ncol,
                             int* __restrict__ rows, int* __restrict__ cols, double* __restrict__ vals) {
  long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; j < nnz; j += stride) {
    rows[j] = fsb_synth_row(seed, (unsigned long long)j, nrow);
    cols[j] = fsb_synth_col(seed, (unsigned long long)j, ncol, dist);
    if (vals) vals[j] = fsb_synth_val(seed, (unsigned long long)j);
  }
}

inline int grid_for(long long n) { return (int)std::min<long long>((n + 255) / 256, 148LL * 32); }

int bits_for(int n) { int b = 1; while (b < 31 && (1LL << b) < n) ++b; return b; }

}  // namespace

// ---- input validation.  The reference trusts its indices (an out-of-range one is undefined behaviour
// on the CPU); on the GPU it would be an illegal address that poisons the whole context, so every
// upload path checks its index arrays once, on the device, before the first product.
namespace {
__global__ void index_range_kernel(const int* __restrict__ a, long long n, int* __restrict__ minmax) {
  int lo = INT_MAX, hi = INT_MIN;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) { const int v = a[i]; lo = min(lo, v); hi = max(hi, v); }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, off));
    hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, off));
  }
  if ((threadIdx.x & 31) == 0) { atomicMin(minmax, lo); atomicMax(minmax + 1, hi); }
}
// non-decreasing, first entry `first`, last entry `last`: flags[0] counts violations
__global__ void monotone_kernel(const int* __restrict__ p, long long n, int first, int last, int* __restrict__ flag) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  int bad = 0;
  for (; i < n; i += stride) {
    const int v = p[i];
    if (i == 0 && v != first) bad = 1;
    if (i == n - 1 && v != last) bad = 1;
    if (i + 1 < n && p[i + 1] < v) bad = 1;
  }
  if (bad) atomicAdd(flag, 1);
}
}  // namespace

// all of d_idx[0..n) in [0, limit)?  `what` names the array in the error message
int fsb_check_index_range(const int* d_idx, long n, int limit, const char* what, cudaStream_t st) {
  if (n <= 0) return FSB_OK;
  int* d = nullptr;
  FSB_CUDA(cudaMalloc(&d, 2 * sizeof(int)));
  const int init[2] = {INT_MAX, INT_MIN};
  int h[2] = {0, 0};
  cudaError_t e = cudaMemcpyAsync(d, init, sizeof init, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) {
    index_range_kernel<<<grid_for(n), 256, 0, st>>>(d_idx, n, d);
    fsb_count_launch();
    e = cudaMemcpyAsync(h, d, sizeof h, cudaMemcpyDeviceToHost, st);
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cudaFree(d);
  if (e != cudaSuccess) return fsb_cuda_error(e, "index validation", __FILE__, __LINE__);
  if (h[0] < 0 || h[1] >= limit)
    return fsb_set_error(FSB_EINVAL, "%s out of range: found [%d, %d], valid is [0, %d)", what, h[0], h[1], limit);
  return FSB_OK;
}

// d_ptr[0..n) non-decreasing with d_ptr[0] == 0 and d_ptr[n-1] == last?
int fsb_check_row_ptr(const int* d_ptr, long n, long last, const char* what, cudaStream_t st) {
  if (last > INT_MAX) return fsb_set_error(FSB_EINVAL, "%ld entries do not fit the int32 offsets of %s (the reference's struct layout)", last, what);
  if (n <= 0) return FSB_OK;
  int* d = nullptr;
  FSB_CUDA(cudaMalloc(&d, sizeof(int)));
  int h = 0;
  cudaError_t e = cudaMemsetAsync(d, 0, sizeof(int), st);
  if (e == cudaSuccess) {
    monotone_kernel<<<grid_for(n), 256, 0, st>>>(d_ptr, n, 0, (int)last, d);
    fsb_count_launch();
    e = cudaMemcpyAsync(&h, d, sizeof h, cudaMemcpyDeviceToHost, st);
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cudaFree(d);
  if (e != cudaSuccess) return fsb_cuda_error(e, "row_ptr validation", __FILE__, __LINE__);
  if (h) return fsb_set_error(FSB_EINVAL, "%s is not a valid offset array (must start at 0, never decrease and end at %ld)", what, last);
  return FSB_OK;
}

// Sort (key, payload...) stably by key in [0, nkeys) and emit CSR arrays.
static int coo_to_csr_dev(fsb_matrix* out, int nkeys, int nother, long nnz, const int* d_keys,
                          const int* d_other, const double* d_vals, cudaStream_t st) {
  // row_ptr and the sort permutation are int32 (the reference's struct layout, csr.h:20): like the host
  // constructor (fsb_host_csr_from_coo), refuse what they cannot index instead of truncating
  if (nnz > (long)INT_MAX) return fsb_set_error(FSB_EINVAL, "%ld entries do not fit the int32 offsets of the CSR structure", nnz);
  out->format = FSB_FMT_CSR;
  out->nrow = nkeys;
  out->ncol = nother;
  out->nnz = nnz;
  out->has_vals = d_vals != nullptr;
  out->avg_row_nnz = nkeys > 0 ? (double)nnz / nkeys : 0.0;
  const size_t n1 = (size_t)std::max<long>(nnz, 1);
  FSB_CUDA(cudaMalloc(&out->row_ptr, ((size_t)nkeys + 1) * sizeof(int)));
  FSB_CUDA(cudaMalloc(&out->cols, n1 * sizeof(int)));
  if (d_vals) FSB_CUDA(cudaMalloc(&out->vals, n1 * sizeof(double)));
  out->bytes = ((size_t)nkeys + 1) * sizeof(int) + n1 * sizeof(int) + (d_vals ? n1 * sizeof(double) : 0);
  if (nnz == 0) {
    FSB_CUDA(cudaMemsetAsync(out->row_ptr, 0, ((size_t)nkeys + 1) * sizeof(int), st));
    return FSB_OK;
  }
  FSB_TRY(fsb_check_index_range(d_keys, nnz, nkeys, "row index", st));
  FSB_TRY(fsb_check_index_range(d_other, nnz, nother, "column index", st));
  int *keys_sorted = nullptr, *perm_in = nullptr, *perm_out = nullptr;
  void* tmp = nullptr;
  size_t tmp_bytes = 0;
  FSB_CUDA(cudaMalloc(&keys_sorted, n1 * sizeof(int)));
  const int end_bit = bits_for(nkeys);
  int rc = FSB_OK;
  if (!d_vals) {
    // payload = the other index itself
    cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_keys, keys_sorted, d_other, out->cols, (long long)nnz, 0, end_bit, st);
    FSB_CUDA(cudaMalloc(&tmp, tmp_bytes));
    cudaError_t e = cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, d_keys, keys_sorted, d_other, out->cols, (long long)nnz, 0, end_bit, st);
    fsb_count_launch(4);
    if (e != cudaSuccess) rc = fsb_cuda_error(e, "cub::DeviceRadixSort::SortPairs", __FILE__, __LINE__);
  } else {
    FSB_CUDA(cudaMalloc(&perm_in, n1 * sizeof(int)));
    FSB_CUDA(cudaMalloc(&perm_out, n1 * sizeof(int)));
    iota_kernel<<<grid_for(nnz), 256, 0, st>>>(perm_in, nnz);
    fsb_count_launch();
    cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_keys, keys_sorted, perm_in, perm_out, (long long)nnz, 0, end_bit, st);
    FSB_CUDA(cudaMalloc(&tmp, tmp_bytes));
    cudaError_t e = cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, d_keys, keys_sorted, perm_in, perm_out, (long long)nnz, 0, end_bit, st);
    fsb_count_launch(4);
    if (e != cudaSuccess) rc = fsb_cuda_error(e, "cub::DeviceRadixSort::SortPairs", __FILE__, __LINE__);
    if (rc == FSB_OK) {
      gather_i32_kernel<<<grid_for(nnz), 256, 0, st>>>(out->cols, d_other, perm_out, nnz);
      gather_f64_kernel<<<grid_for(nnz), 256, 0, st>>>(out->vals, d_vals, perm_out, nnz);
      fsb_count_launch(2);
    }
  }
  if (rc == FSB_OK) {
    row_ptr_from_sorted_kernel<<<grid_for(nnz + 1), 256, 0, st>>>(keys_sorted, nnz, nkeys, out->row_ptr);
    fsb_count_launch();
    cudaError_t e = cudaStreamSynchronize(st);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) rc = fsb_cuda_error(e, "coo_to_csr_dev", __FILE__, __LINE__);
  }
  cudaFree(keys_sorted); cudaFree(perm_in); cudaFree(perm_out); cudaFree(tmp);
  return rc;
}

// perm = indices 0..n-1 stably sorted by key (keys in [0, nkeys)); ptr[k] = first sorted position of key k
int fsb_stable_perm_by_key(const int* d_keys, int nkeys, long n, int* d_perm, int* d_ptr, cudaStream_t st) {
  if (n == 0) {
    FSB_CUDA(cudaMemsetAsync(d_ptr, 0, ((size_t)nkeys + 1) * sizeof(int), st));
    return FSB_OK;
  }
  int *keys_sorted = nullptr, *iota = nullptr;
  void* tmp = nullptr;
  size_t tmp_bytes = 0;
  FSB_CUDA(cudaMalloc(&keys_sorted, (size_t)n * sizeof(int)));
  cudaError_t e = cudaMalloc(&iota, (size_t)n * sizeof(int));
  int rc = e == cudaSuccess ? FSB_OK : fsb_cuda_error(e, "cudaMalloc", __FILE__, __LINE__);
  if (rc == FSB_OK) {
    iota_kernel<<<grid_for(n), 256, 0, st>>>(iota, n);
    fsb_count_launch();
    const int end_bit = bits_for(nkeys);
    cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_keys, keys_sorted, iota, d_perm, (long long)n, 0, end_bit, st);
    e = cudaMalloc(&tmp, tmp_bytes);
    if (e == cudaSuccess) e = cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, d_keys, keys_sorted, iota, d_perm, (long long)n, 0, end_bit, st);
    fsb_count_launch(4);
    if (e == cudaSuccess) {
      row_ptr_from_sorted_kernel<<<grid_for(n + 1), 256, 0, st>>>(keys_sorted, n, nkeys, d_ptr);
      fsb_count_launch();
      e = cudaStreamSynchronize(st);
    }
    if (e != cudaSuccess) rc = fsb_cuda_error(e, "stable_perm_by_key", __FILE__, __LINE__);
  }
  cudaFree(keys_sorted); cudaFree(iota); cudaFree(tmp);
  return rc;
}

namespace {
// class key of a blocked-COO entry: block * kBlockedClasses + (local row mod kBlockedClasses)
__global__ void blocked_keys_kernel(const int* __restrict__ rows, const long* __restrict__ blk_off, const int* __restrict__ start_row,
                                    int nblocks, long long nnz, int* __restrict__ keys) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < nnz; i += stride) {
    int lo = 0, hi = nblocks;  // largest b with blk_off[b] <= i
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (blk_off[mid] <= i) lo = mid; else hi = mid;
    }
    keys[i] = lo * FSB_BLOCKED_CLASSES + ((rows[i] - start_row[lo]) & (FSB_BLOCKED_CLASSES - 1));
  }
}
}  // namespace

namespace {
__global__ void rows_in_blocks_kernel(const int* __restrict__ rows, const long* __restrict__ blk_off, const int* __restrict__ start_row,
                                      int nblocks, long long nnz, int* __restrict__ flag) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  int bad = 0;
  for (; i < nnz; i += stride) {
    int lo = 0, hi = nblocks;  // largest b with blk_off[b] <= i
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (blk_off[mid] <= i) lo = mid; else hi = mid;
    }
    const int r = rows[i];
    if (r < start_row[lo] || r >= start_row[lo + 1]) bad = 1;
  }
  if (bad) atomicAdd(flag, 1);
}
}  // namespace

int fsb_check_rows_in_blocks(const int* d_rows, const long* d_blk_off, const int* d_start_row, int nblocks, long nnz, cudaStream_t st) {
  if (nnz <= 0 || nblocks <= 0) return FSB_OK;
  int* d = nullptr;
  FSB_CUDA(cudaMalloc(&d, sizeof(int)));
  int h = 0;
  cudaError_t e = cudaMemsetAsync(d, 0, sizeof(int), st);
  if (e == cudaSuccess) {
    rows_in_blocks_kernel<<<grid_for(nnz), 256, 0, st>>>(d_rows, d_blk_off, d_start_row, nblocks, nnz, d);
    fsb_count_launch();
    e = cudaMemcpyAsync(&h, d, sizeof h, cudaMemcpyDeviceToHost, st);
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cudaFree(d);
  if (e != cudaSuccess) return fsb_cuda_error(e, "block membership validation", __FILE__, __LINE__);
  if (h) return fsb_set_error(FSB_EINVAL, "blocked matrix: an entry's row lies outside its block's [start_row[b], start_row[b+1])");
  return FSB_OK;
}

// Re-lay a freshly uploaded blocked COO into row-class buckets (see kernels_blocked.cu):
// within each block, entries are stably bucketed by (local row mod 256), so a team of
// lanes that owns a row class streams its own list in the stored (e.g. Hilbert) order.
int fsb_blocked_relayout(fsb_matrix* A, cudaStream_t st) {
  const long nnz = A->nnz;
  const long nkeys_l = (long)A->nblocks * FSB_BLOCKED_CLASSES;
  if (nkeys_l >= (1L << 31)) return fsb_set_error(FSB_EINVAL, "blocked: too many row blocks (%d)", A->nblocks);
  const int nkeys = (int)nkeys_l;
  FSB_CUDA(cudaMalloc(&A->row_ptr, ((size_t)nkeys + 1) * sizeof(int)));
  A->bytes += ((size_t)nkeys + 1) * sizeof(int);
  if (nnz == 0) {
    FSB_CUDA(cudaMemsetAsync(A->row_ptr, 0, ((size_t)nkeys + 1) * sizeof(int), st));
    return FSB_OK;
  }
  int *keys = nullptr, *perm = nullptr, *r2 = nullptr, *c2 = nullptr;
  double* v2 = nullptr;
  int rc = FSB_OK;
  cudaError_t e = cudaMalloc(&keys, (size_t)nnz * 4);
  if (e == cudaSuccess) e = cudaMalloc(&perm, (size_t)nnz * 4);
  if (e == cudaSuccess) e = cudaMalloc(&r2, (size_t)nnz * 4);
  if (e == cudaSuccess) e = cudaMalloc(&c2, (size_t)nnz * 4);
  if (e == cudaSuccess && A->has_vals) e = cudaMalloc(&v2, (size_t)nnz * 8);
  if (e != cudaSuccess) rc = fsb_cuda_error(e, "cudaMalloc", __FILE__, __LINE__);
  if (rc == FSB_OK) {
    blocked_keys_kernel<<<grid_for(nnz), 256, 0, st>>>(A->b_rows, A->blk_off, A->start_row, A->nblocks, nnz, keys);
    fsb_count_launch();
    rc = fsb_stable_perm_by_key(keys, nkeys, nnz, perm, A->row_ptr, st);
  }
  if (rc == FSB_OK) {
    gather_i32_kernel<<<grid_for(nnz), 256, 0, st>>>(r2, A->b_rows, perm, nnz);
    gather_i32_kernel<<<grid_for(nnz), 256, 0, st>>>(c2, A->b_cols, perm, nnz);
    if (A->has_vals) gather_f64_kernel<<<grid_for(nnz), 256, 0, st>>>(v2, A->b_vals, perm, nnz);
    fsb_count_launch(A->has_vals ? 3 : 2);
    e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) rc = fsb_cuda_error(e, "blocked relayout", __FILE__, __LINE__);
  }
  if (rc == FSB_OK) {
    cudaFree(A->b_rows); cudaFree(A->b_cols); cudaFree(A->b_vals);
    A->b_rows = r2; A->b_cols = c2; A->b_vals = v2;
    r2 = c2 = nullptr; v2 = nullptr;
  }
  cudaFree(keys); cudaFree(perm); cudaFree(r2); cudaFree(c2); cudaFree(v2);
  return rc;
}

int fsb_build_csr_from_coo_dev(fsb_matrix* out, int nrow, int ncol, long nnz, const int* d_rows,
                               const int* d_cols, const double* d_vals, cudaStream_t st) {
  return coo_to_csr_dev(out, nrow, ncol, nnz, d_rows, d_cols, d_vals, st);
}

namespace {
// row of every stored entry of a column-blocked CSR: cell = block*nrow + row (cbcsr.h:41)
__global__ void expand_cell_rows_kernel(const int* __restrict__ cell_ptr, long long ncell, int nrow, long long nnz, int* __restrict__ out) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < nnz; i += stride) {
    long long lo = 0, hi = ncell;  // largest cell with cell_ptr[cell] <= i
    while (hi - lo > 1) {
      const long long mid = (lo + hi) >> 1;
      if (cell_ptr[mid] <= i) lo = mid; else hi = mid;
    }
    out[i] = (int)(lo % nrow);
  }
}
}  // namespace

// CSR view of a blocked (row-blocked COO) or column-blocked matrix: the same entries stably
// sorted by row.  Blocked: the class buckets are stable, so a row's entries keep their stored
// (e.g. Hilbert) order.  Column-blocked: a row's cells follow each other in block order.
int fsb_build_csr_view(fsb_matrix* A, cudaStream_t st) {
  if (A->view) return FSB_OK;
  fsb_matrix* V = new fsb_matrix();
  int rc;
  if (A->format == FSB_FMT_BLOCKED) {
    rc = coo_to_csr_dev(V, A->nrow, A->ncol, A->nnz, A->b_rows, A->b_cols, A->b_vals, st);
  } else if (A->format == FSB_FMT_CBCSR) {
    int* rowid = nullptr;
    cudaError_t e = cudaMalloc(&rowid, (size_t)std::max<long>(A->nnz, 1) * sizeof(int));
    if (e != cudaSuccess) { delete V; return fsb_cuda_error(e, "cudaMalloc", __FILE__, __LINE__); }
    if (A->nnz > 0) {
      expand_cell_rows_kernel<<<grid_for(A->nnz), 256, 0, st>>>(A->row_ptr, (long long)A->nblocks * A->nrow, A->nrow, A->nnz, rowid);
      fsb_count_launch();
    }
    rc = coo_to_csr_dev(V, A->nrow, A->ncol, A->nnz, rowid, A->cols, nullptr, st);
    cudaFree(rowid);
  } else {
    delete V;
    return fsb_set_error(FSB_EINVAL, "csr view: blocked or column-blocked handle required");
  }
  if (rc != FSB_OK) {
    cudaFree(V->row_ptr); cudaFree(V->cols); cudaFree(V->vals);
    delete V;
    return rc;
  }
  A->view = V;
  return FSB_OK;
}

int fsb_build_transpose(fsb_matrix* A, cudaStream_t st) {
  if (A->T) return FSB_OK;
  if (A->format != FSB_FMT_CSR) return fsb_set_error(FSB_EINVAL, "transpose: CSR handles only");
  int* rowid = nullptr;
  const size_t n1 = (size_t)std::max<long>(A->nnz, 1);
  FSB_CUDA(cudaMalloc(&rowid, n1 * sizeof(int)));
  if (A->nnz > 0) {
    expand_rows_kernel<<<grid_for(A->nnz), 256, 0, st>>>(A->row_ptr, A->nrow, A->nnz, rowid);
    fsb_count_launch();
  }
  fsb_matrix* T = new fsb_matrix();
  int rc = coo_to_csr_dev(T, A->ncol, A->nrow, A->nnz, A->cols, rowid, A->vals, st);
  cudaFree(rowid);
  if (rc != FSB_OK) {
    cudaFree(T->row_ptr); cudaFree(T->cols); cudaFree(T->vals);
    delete T;
    return rc;
  }
  A->T = T;
  return FSB_OK;
}

namespace {
__global__ void xblock_key_kernel(const int* __restrict__ rowid, const int* __restrict__ cols, long long nnz, int rows_per_block, int ncol,
                                  int* __restrict__ keys) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < nnz; i += stride) keys[i] = (rowid[i] / rows_per_block) * ncol + cols[i];
}
}  // namespace

// x-blocked transpose (see fsb_internal.h): cell = (row of A / rows_per_block) * ncol + column of A; stable sort by
// cell, so inside a cell the entries keep the CSR order (increasing row of A) -- the same order the plain transpose
// gives each of its rows, cut at the block boundaries.
int fsb_build_transpose_xblocked(fsb_matrix* A, size_t block_bytes, cudaStream_t st) {
  if (A->Tb) return FSB_OK;
  if (A->format != FSB_FMT_CSR) return fsb_set_error(FSB_EINVAL, "x-blocked transpose: CSR handles only");
  const size_t xbytes = (size_t)A->nrow * 8;
  const int nb = (int)std::max<size_t>(1, (xbytes + block_bytes - 1) / block_bytes);
  const int rpb = (A->nrow + nb - 1) / nb;
  if ((long)nb * A->ncol >= (long)INT_MAX) return fsb_set_error(FSB_EINVAL, "x-blocked transpose: %d blocks x %d columns exceed the int32 cell range", nb, A->ncol);
  const size_t n1 = (size_t)std::max<long>(A->nnz, 1);
  int *rowid = nullptr, *keys = nullptr;
  FSB_CUDA(cudaMalloc(&rowid, n1 * sizeof(int)));
  cudaError_t e = cudaMalloc(&keys, n1 * sizeof(int));
  if (e != cudaSuccess) { cudaFree(rowid); return fsb_cuda_error(e, "cudaMalloc", __FILE__, __LINE__); }
  if (A->nnz > 0) {
    expand_rows_kernel<<<grid_for(A->nnz), 256, 0, st>>>(A->row_ptr, A->nrow, A->nnz, rowid);
    xblock_key_kernel<<<grid_for(A->nnz), 256, 0, st>>>(rowid, A->cols, A->nnz, rpb, A->ncol, keys);
    fsb_count_launch(2);
  }
  fsb_matrix* T = new fsb_matrix();
  int rc = coo_to_csr_dev(T, nb * A->ncol, A->nrow, A->nnz, keys, rowid, A->vals, st);
  cudaFree(rowid); cudaFree(keys);
  if (rc != FSB_OK) {
    cudaFree(T->row_ptr); cudaFree(T->cols); cudaFree(T->vals);
    delete T;
    return rc;
  }
  T->x_live_bytes = (size_t)rpb * 8;
  A->Tb = T;
  A->tb_blocks = nb;
  return FSB_OK;
}

// ---------------------------------------------------------------- device-side builders for the blocked formats
namespace {

__device__ __forceinline__ long long dev_xy2d(int n, int x, int y) {   // hilbert.h:16-27 (+ rot 45-57), device twin
  long long d = 0;
  for (long long s = n / 2; s > 0; s /= 2) {
    const int rx = (x & s) > 0, ry = (y & s) > 0;
    d += s * s * (long long)((3 * rx) ^ ry);
    if (!ry) {
      if (rx) { x = (int)s - 1 - x; y = (int)s - 1 - y; }
      const int t = x; x = y; y = t;
    }
  }
  return d;
}

__device__ __forceinline__ int dev_ceil_pow2(int x) { int p = 1; while (p < x) p <<= 1; return p; }

// composite 64-bit key: (block * 256 + row class) in the high bits, in-block order key below.
// order 0: original position (stable COO order), 1: row_xy2d Hilbert key (sort_bsbm), 2: row*ncol+col
__global__ void blocked_sortkey_kernel(const int* __restrict__ rows, const int* __restrict__ cols, long long nnz, int nrow,
                                       int ncol, int block_size, int order, unsigned long long* __restrict__ keys) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < nnz; i += stride) {
    const int r = rows[i], c = cols[i];
    const int b = r / block_size;
    const int r0 = b * block_size;
    const int lr = r - r0;
    unsigned long long low;
    if (order == 1) {
      const int n = dev_ceil_pow2(min(block_size, nrow - r0));
      low = (unsigned long long)(dev_xy2d(n, c % n, lr) + (long long)n * n * (c / n));   // row_xy2d hilbert.h:60-65
    } else if (order == 2) {
      low = (unsigned long long)lr * (unsigned long long)ncol + (unsigned long long)c;
    } else {
      low = (unsigned long long)i;
    }
    const unsigned long long hi = (unsigned long long)b * FSB_BLOCKED_CLASSES + (unsigned long long)(lr & (FSB_BLOCKED_CLASSES - 1));
    keys[i] = (hi << 40) | (low & ((1ull << 40) - 1));
  }
}

__global__ void class_ptr_kernel(const unsigned long long* __restrict__ keys, long long n, int nkeys, int* __restrict__ ptr) {
  long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; j <= n; j += stride) {
    const int lo = (j == 0) ? -1 : (int)(keys[j - 1] >> 40);
    const int hi = (j == n) ? nkeys : (int)(keys[j] >> 40);
    for (int q = lo + 1; q <= hi; ++q) ptr[q] = (int)j;
  }
}

__global__ void cell_key_kernel(const int* __restrict__ rows, const int* __restrict__ cols, long long nnz, int nrow,
                                int colblocksize, int* __restrict__ keys) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < nnz; i += stride) keys[i] = (cols[i] / colblocksize) * nrow + rows[i];
}

__global__ void block_meta_kernel(int* __restrict__ start_row, long* __restrict__ blk_off, const int* __restrict__ cls_ptr,
                                  int nblocks, int nrow, int block_size) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b > nblocks) return;
  start_row[b] = b < nblocks ? b * block_size : nrow;
  blk_off[b] = cls_ptr[(long long)b * FSB_BLOCKED_CLASSES];
}

}  // namespace

// Row-blocked COO built on the device from a device COO: equivalent to
// new_bsbm (sparse.h:175-213) followed by nothing (order 0), sort_bsbm (order 1, sparse.h:215-236)
// or sort_bsbm_byrow (order 2, sparse.h:238-256), directly in the kernel's class-bucketed layout.
extern "C" int fsb_blocked_from_coo_dev(fsb_matrix_t* out, int nrow, int ncol, long nnz, const int* d_rows, const int* d_cols,
                                        const double* d_vals, int block_size, int order) {
  FSB_TRY(fsb_require_device());
  if (!out || nrow <= 0 || ncol <= 0 || nnz < 0 || block_size <= 0 || order < 0 || order > 2 || (nnz > 0 && (!d_rows || !d_cols)))
    return fsb_set_error(FSB_EINVAL, "fsb_blocked_from_coo_dev: bad arguments");
  const int nblocks = (nrow + block_size - 1) / block_size;
  const long nkeys_l = (long)nblocks * FSB_BLOCKED_CLASSES;
  if (nkeys_l >= (1L << 24)) return fsb_set_error(FSB_EINVAL, "fsb_blocked_from_coo_dev: too many row blocks (%d) for the 24-bit class key", nblocks);
  if (nnz > (long)INT_MAX) return fsb_set_error(FSB_EINVAL, "fsb_blocked_from_coo_dev: %ld entries do not fit the int32 class offsets", nnz);
  // the in-block order key lives in the low 40 bits of the composite sort key: bound it for every order
  //   order 1: row_xy2d(n, lr, c) = xy2d(n, c % n, lr) + n*n*(c / n) < n*n*ceil(ncol / n)  (>= n*n when ncol < n)
  //   order 2: lr*ncol + c < block_size*ncol;   order 0: the entry's position < nnz
  {
    long long n = 1; while (n < std::min(block_size, nrow)) n <<= 1;
    const double lim = (double)(1ull << 40);
    const double k1 = (double)n * (double)n * (double)(((long long)ncol + n - 1) / n);
    const double k2 = (double)std::min(block_size, nrow) * (double)ncol;
    if ((order == 1 && k1 >= lim) || (order == 2 && k2 >= lim) || (order == 0 && (double)nnz >= lim))
      return fsb_set_error(FSB_EINVAL, "fsb_blocked_from_coo_dev: in-block order key exceeds 40 bits (block_size %d, ncol %d)", block_size, ncol);
  }
  cudaStream_t st = fsb_default_stream();
  // every upload path checks its indices (an out-of-range row would index past the class offsets below)
  if (nnz > 0) {
    FSB_TRY(fsb_check_index_range(d_rows, nnz, nrow, "row index", st));
    FSB_TRY(fsb_check_index_range(d_cols, nnz, ncol, "column index", st));
  }
  fsb_matrix* A = new fsb_matrix();
  A->format = FSB_FMT_BLOCKED; A->nrow = nrow; A->ncol = ncol; A->nnz = nnz; A->nblocks = nblocks; A->has_vals = d_vals != nullptr;
  A->max_block_rows = std::min(block_size, nrow);
  A->avg_row_nnz = (double)nnz / nrow;
  const int nkeys = (int)nkeys_l;
  const size_t n1 = (size_t)std::max<long>(nnz, 1);
  unsigned long long *keys = nullptr, *keys_sorted = nullptr;
  int *iota = nullptr, *perm = nullptr;
  void* tmp = nullptr;
  size_t tmp_bytes = 0;
  cudaError_t e = cudaMalloc(&A->row_ptr, ((size_t)nkeys + 1) * 4);
  if (e == cudaSuccess) e = cudaMalloc(&A->start_row, ((size_t)nblocks + 1) * 4);
  if (e == cudaSuccess) e = cudaMalloc(&A->blk_off, ((size_t)nblocks + 1) * 8);
  if (e == cudaSuccess) e = cudaMalloc(&A->b_rows, n1 * 4);
  if (e == cudaSuccess) e = cudaMalloc(&A->b_cols, n1 * 4);
  if (e == cudaSuccess && d_vals) e = cudaMalloc(&A->b_vals, n1 * 8);
  if (e == cudaSuccess) e = cudaMalloc(&keys, n1 * 8);
  if (e == cudaSuccess) e = cudaMalloc(&keys_sorted, n1 * 8);
  if (e == cudaSuccess) e = cudaMalloc(&iota, n1 * 4);
  if (e == cudaSuccess) e = cudaMalloc(&perm, n1 * 4);
  if (e == cudaSuccess && nnz > 0) {
    blocked_sortkey_kernel<<<grid_for(nnz), 256, 0, st>>>(d_rows, d_cols, nnz, nrow, ncol, block_size, order, keys);
    iota_kernel<<<grid_for(nnz), 256, 0, st>>>(iota, nnz);
    fsb_count_launch(2);
    cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys, keys_sorted, iota, perm, (long long)nnz, 0, 64, st);
    e = cudaMalloc(&tmp, tmp_bytes);
    if (e == cudaSuccess) e = cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys, keys_sorted, iota, perm, (long long)nnz, 0, 64, st);
    fsb_count_launch(9);
    if (e == cudaSuccess) {
      class_ptr_kernel<<<grid_for(nnz + 1), 256, 0, st>>>(keys_sorted, nnz, nkeys, A->row_ptr);
      gather_i32_kernel<<<grid_for(nnz), 256, 0, st>>>(A->b_rows, d_rows, perm, nnz);
      gather_i32_kernel<<<grid_for(nnz), 256, 0, st>>>(A->b_cols, d_cols, perm, nnz);
      if (d_vals) gather_f64_kernel<<<grid_for(nnz), 256, 0, st>>>(A->b_vals, d_vals, perm, nnz);
      fsb_count_launch(d_vals ? 4 : 3);
    }
  } else if (e == cudaSuccess) {
    e = cudaMemsetAsync(A->row_ptr, 0, ((size_t)nkeys + 1) * 4, st);
  }
  if (e == cudaSuccess) {
    block_meta_kernel<<<(nblocks + 256) / 256, 256, 0, st>>>(A->start_row, A->blk_off, A->row_ptr, nblocks, nrow, block_size);
    fsb_count_launch();
    e = cudaStreamSynchronize(st);
  }
  cudaFree(keys); cudaFree(keys_sorted); cudaFree(iota); cudaFree(perm); cudaFree(tmp);
  if (e != cudaSuccess) {
    const int rc = fsb_cuda_error(e, "fsb_blocked_from_coo_dev", __FILE__, __LINE__);
    fsb_matrix_free(A);
    return rc;
  }
  A->bytes = ((size_t)nkeys + 1) * 4 + ((size_t)nblocks + 1) * 12 + n1 * (d_vals ? 16 : 8);
  *out = A;
  return FSB_OK;
}

// ---------------------------------------------------------------- global Hilbert order of a COO (sort_sbm / sort_sdm)
namespace {
__device__ __forceinline__ void dev_d2xy(int n, long long d, int* xo, int* yo) {   // hilbert.h:30-42, device twin
  int x = 0, y = 0;
  long long t = d;
  for (int s = 1; s < n; s *= 2) {
    const int rx = (int)(1 & (t / 2));
    const int ry = (int)(1 & (t ^ rx));
    if (!ry) {
      if (rx) { x = s - 1 - x; y = s - 1 - y; }
      const int w = x; x = y; y = w;
    }
    x += s * rx;
    y += s * ry;
    t /= 4;
  }
  *xo = x;
  *yo = y;
}

__global__ void hilbert_key_kernel(const int* __restrict__ rows, const int* __restrict__ cols, long long nnz, int n,
                                   unsigned long long* __restrict__ keys) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < nnz; i += stride) keys[i] = (unsigned long long)dev_xy2d(n, rows[i], cols[i]);   // sparse.h:150-153
}

__global__ void hilbert_decode_kernel(const unsigned long long* __restrict__ keys, long long nnz, int n, int* __restrict__ rows,
                                      int* __restrict__ cols) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < nnz; i += stride) dev_d2xy(n, (long long)keys[i], &rows[i], &cols[i]);           // sparse.h:155-158
}
}  // namespace

// sort_sbm (sparse.h:142-161) / sort_sdm (dsparse.h:96-115) on the device: key = xy2d(n, row, col) with
// n = ceilPower2(max(nrow, ncol)), ascending sort, coordinates decoded back from the sorted keys with d2xy -- the
// reference's three steps, with a radix sort on the 2 log2(n) significant key bits in place of the serial quicksort.
// The result is the unique ascending key order; entries with EQUAL coordinates are interchangeable for binary
// matrices, and for valued ones the radix sort keeps their input order (the reference's quickSortD leaves it
// unspecified).  d_rows / d_cols / d_vals are sorted in place.
extern "C" int fsb_sort_coo_hilbert_dev(int nrow, int ncol, long nnz, int* d_rows, int* d_cols, double* d_vals) {
  FSB_TRY(fsb_require_device());
  if (nrow < 0 || ncol < 0 || nnz < 0 || (nnz > 0 && (!d_rows || !d_cols))) return fsb_set_error(FSB_EINVAL, "fsb_sort_coo_hilbert_dev: bad arguments");
  if (nnz == 0) return FSB_OK;
  cudaStream_t st = fsb_default_stream();
  FSB_TRY(fsb_check_index_range(d_rows, nnz, nrow, "row index", st));
  FSB_TRY(fsb_check_index_range(d_cols, nnz, ncol, "column index", st));
  const int n = fsb_host_ceil_pow2(std::max(nrow, ncol));
  int key_bits = 2;
  while ((1LL << (key_bits / 2)) < n) key_bits += 2;          // keys < n*n
  unsigned long long *keys = nullptr, *keys_sorted = nullptr;
  double* vals_sorted = nullptr;
  void* tmp = nullptr;
  size_t tmp_bytes = 0;
  cudaError_t e = cudaMalloc(&keys, (size_t)nnz * 8);
  if (e == cudaSuccess) e = cudaMalloc(&keys_sorted, (size_t)nnz * 8);
  if (e == cudaSuccess && d_vals) e = cudaMalloc(&vals_sorted, (size_t)nnz * 8);
  if (e == cudaSuccess) {
    hilbert_key_kernel<<<grid_for(nnz), 256, 0, st>>>(d_rows, d_cols, nnz, n, keys);
    fsb_count_launch();
    if (d_vals) {
      cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys, keys_sorted, d_vals, vals_sorted, (long long)nnz, 0, key_bits, st);
      e = cudaMalloc(&tmp, tmp_bytes);
      if (e == cudaSuccess) e = cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys, keys_sorted, d_vals, vals_sorted, (long long)nnz, 0, key_bits, st);
      if (e == cudaSuccess) e = cudaMemcpyAsync(d_vals, vals_sorted, (size_t)nnz * 8, cudaMemcpyDeviceToDevice, st);
    } else {
      cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, keys, keys_sorted, (long long)nnz, 0, key_bits, st);
      e = cudaMalloc(&tmp, tmp_bytes);
      if (e == cudaSuccess) e = cub::DeviceRadixSort::SortKeys(tmp, tmp_bytes, keys, keys_sorted, (long long)nnz, 0, key_bits, st);
    }
    fsb_count_launch(2 + key_bits / 8);
  }
  if (e == cudaSuccess) {
    hilbert_decode_kernel<<<grid_for(nnz), 256, 0, st>>>(keys_sorted, nnz, n, d_rows, d_cols);
    fsb_count_launch();
    e = cudaStreamSynchronize(st);
  }
  cudaFree(keys); cudaFree(keys_sorted); cudaFree(vals_sorted); cudaFree(tmp);
  if (e != cudaSuccess) return fsb_cuda_error(e, "fsb_sort_coo_hilbert_dev", __FILE__, __LINE__);
  return FSB_OK;
}

// host arrays: upload, sort on the device, copy back (what the drop-in sort_sbm / sort_sdm call for large matrices)
extern "C" int fsb_sort_coo_hilbert(int nrow, int ncol, long nnz, int* rows, int* cols, double* vals) {
  FSB_TRY(fsb_require_device());
  if (nnz < 0 || (nnz > 0 && (!rows || !cols))) return fsb_set_error(FSB_EINVAL, "fsb_sort_coo_hilbert: bad arguments");
  if (nnz == 0) return FSB_OK;
  cudaStream_t st = fsb_default_stream();
  int *dr = nullptr, *dc = nullptr;
  double* dv = nullptr;
  cudaError_t e = cudaMalloc(&dr, (size_t)nnz * 4);
  if (e == cudaSuccess) e = cudaMalloc(&dc, (size_t)nnz * 4);
  if (e == cudaSuccess && vals) e = cudaMalloc(&dv, (size_t)nnz * 8);
  int rc = e == cudaSuccess ? FSB_OK : fsb_cuda_error(e, "cudaMalloc", __FILE__, __LINE__);
  if (rc == FSB_OK) rc = fsb_h2d(dr, rows, (size_t)nnz * 4, st);
  if (rc == FSB_OK) rc = fsb_h2d(dc, cols, (size_t)nnz * 4, st);
  if (rc == FSB_OK && vals) rc = fsb_h2d(dv, vals, (size_t)nnz * 8, st);
  if (rc == FSB_OK) rc = fsb_sort_coo_hilbert_dev(nrow, ncol, nnz, dr, dc, dv);
  if (rc == FSB_OK) rc = fsb_d2h(rows, dr, (size_t)nnz * 4, st);
  if (rc == FSB_OK) rc = fsb_d2h(cols, dc, (size_t)nnz * 4, st);
  if (rc == FSB_OK && vals) rc = fsb_d2h(vals, dv, (size_t)nnz * 8, st);
  if (rc == FSB_OK && cudaStreamSynchronize(st) != cudaSuccess) rc = fsb_cuda_error(cudaGetLastError(), "fsb_sort_coo_hilbert", __FILE__, __LINE__);
  cudaFree(dr); cudaFree(dc); cudaFree(dv);
  return rc;
}

// ---------------------------------------------------------------- per-block orders of a HOST blocked structure
namespace {
// key = (block << 40) | in-block key; order 1: row_xy2d(n_b, row - start_row[b], col) with n_b = ceilPower2(rows of block b)
// (sort_bsbm sparse.h:215-236, hilbert.h:60-65), order 2: (row - start_row[b]) * ncol + col (sort_bsbm_byrow sparse.h:238-256)
__global__ void host_blocked_sortkey_kernel(const int* __restrict__ rows, const int* __restrict__ cols, long long nnz,
                                            const long* __restrict__ blk_off, const int* __restrict__ start_row, int nblocks, int ncol,
                                            int order, unsigned long long* __restrict__ keys) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < nnz; i += stride) {
    int lo = 0, hi = nblocks;  // largest b with blk_off[b] <= i
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (blk_off[mid] <= i) lo = mid; else hi = mid;
    }
    const int r0 = start_row[lo], lr = rows[i] - r0, c = cols[i];
    unsigned long long low;
    if (order == 1) {
      const int n = dev_ceil_pow2(start_row[lo + 1] - r0);
      low = (unsigned long long)(dev_xy2d(n, c % n, lr) + (long long)n * n * (c / n));
    } else {
      low = (unsigned long long)lr * (unsigned long long)ncol + (unsigned long long)c;
    }
    keys[i] = ((unsigned long long)lo << 40) | (low & ((1ull << 40) - 1));
  }
}
}  // namespace

// sort_bsbm / sort_bsbm_byrow / sort_bsdm on a HOST BlockedSBM / BlockedSDM, done on the device: the blocks are
// uploaded back to back, every entry gets the key (block, in-block key), one radix sort orders all blocks at once (the
// block id in the high bits keeps every entry inside its block), and the arrays go back into the caller's per-block
// storage.  Same order as the host routine whenever coordinates inside a block are unique (equal keys keep input order).
extern "C" int fsb_sort_blocked(int nrow, int ncol, int nblocks, const int* start_row, const int* blk_nnz, int* const* rows,
                                int* const* cols, double* const* vals, int order) {
  FSB_TRY(fsb_require_device());
  if (nblocks < 0 || (order != 1 && order != 2) || (nblocks > 0 && (!start_row || !blk_nnz || !rows || !cols)))
    return fsb_set_error(FSB_EINVAL, "fsb_sort_blocked: bad arguments");
  std::vector<long> off((size_t)nblocks + 1, 0);
  int max_rows = 0;
  for (int b = 0; b < nblocks; ++b) {
    if (blk_nnz[b] < 0 || start_row[b + 1] < start_row[b]) return fsb_set_error(FSB_EINVAL, "fsb_sort_blocked: bad block metadata");
    off[b + 1] = off[b] + blk_nnz[b];
    max_rows = std::max(max_rows, start_row[b + 1] - start_row[b]);
  }
  const long nnz = off[nblocks];
  if (nnz == 0) return FSB_OK;
  if (nnz > (long)INT_MAX || nblocks >= (1 << 24)) return fsb_set_error(FSB_EINVAL, "fsb_sort_blocked: too many entries or blocks for the sort key");
  {
    long long n = 1; while (n < max_rows) n <<= 1;
    const double lim = (double)(1ull << 40);
    const double k1 = (double)n * (double)n * (double)(((long long)ncol + n - 1) / n), k2 = (double)max_rows * (double)ncol;
    if ((order == 1 && k1 >= lim) || (order == 2 && k2 >= lim)) return fsb_set_error(FSB_EINVAL, "fsb_sort_blocked: in-block order key exceeds 40 bits");
  }
  cudaStream_t st = fsb_default_stream();
  int *dr = nullptr, *dc = nullptr, *dr2 = nullptr, *dc2 = nullptr, *dsr = nullptr, *iota = nullptr, *perm = nullptr;
  double *dv = nullptr, *dv2 = nullptr;
  long* doff = nullptr;
  unsigned long long *keys = nullptr, *keys_sorted = nullptr;
  void* tmp = nullptr;
  size_t tmp_bytes = 0;
  int rc = FSB_OK;
  auto A = [&](void** p, size_t bytes) { if (rc == FSB_OK && cudaMalloc(p, bytes) != cudaSuccess) rc = fsb_cuda_error(cudaGetLastError(), "cudaMalloc", __FILE__, __LINE__); };
  A((void**)&dr, (size_t)nnz * 4); A((void**)&dc, (size_t)nnz * 4); A((void**)&dr2, (size_t)nnz * 4); A((void**)&dc2, (size_t)nnz * 4);
  if (vals) { A((void**)&dv, (size_t)nnz * 8); A((void**)&dv2, (size_t)nnz * 8); }
  A((void**)&dsr, ((size_t)nblocks + 1) * 4); A((void**)&doff, ((size_t)nblocks + 1) * 8);
  A((void**)&keys, (size_t)nnz * 8); A((void**)&keys_sorted, (size_t)nnz * 8); A((void**)&iota, (size_t)nnz * 4); A((void**)&perm, (size_t)nnz * 4);
  std::vector<size_t> b4((size_t)nblocks), b8((size_t)nblocks);
  for (int b = 0; b < nblocks; ++b) { b4[b] = (size_t)blk_nnz[b] * 4; b8[b] = (size_t)blk_nnz[b] * 8; }
  if (rc == FSB_OK) rc = fsb_h2d_gather(dr, (const void* const*)rows, b4.data(), nblocks, st);
  if (rc == FSB_OK) rc = fsb_h2d_gather(dc, (const void* const*)cols, b4.data(), nblocks, st);
  if (rc == FSB_OK && vals) rc = fsb_h2d_gather(dv, (const void* const*)vals, b8.data(), nblocks, st);
  if (rc == FSB_OK) rc = fsb_h2d(dsr, start_row, ((size_t)nblocks + 1) * 4, st);
  if (rc == FSB_OK) rc = fsb_h2d(doff, off.data(), ((size_t)nblocks + 1) * 8, st);
  if (rc == FSB_OK) rc = fsb_check_index_range(dc, nnz, ncol, "column index", st);
  if (rc == FSB_OK) rc = fsb_check_rows_in_blocks(dr, doff, dsr, nblocks, nnz, st);
  if (rc == FSB_OK) {
    host_blocked_sortkey_kernel<<<grid_for(nnz), 256, 0, st>>>(dr, dc, nnz, doff, dsr, nblocks, ncol, order, keys);
    iota_kernel<<<grid_for(nnz), 256, 0, st>>>(iota, nnz);
    fsb_count_launch(2);
    cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys, keys_sorted, iota, perm, (long long)nnz, 0, 64, st);
    cudaError_t e = cudaMalloc(&tmp, tmp_bytes);
    if (e == cudaSuccess) e = cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys, keys_sorted, iota, perm, (long long)nnz, 0, 64, st);
    fsb_count_launch(9);
    if (e == cudaSuccess) {
      gather_i32_kernel<<<grid_for(nnz), 256, 0, st>>>(dr2, dr, perm, nnz);
      gather_i32_kernel<<<grid_for(nnz), 256, 0, st>>>(dc2, dc, perm, nnz);
      if (vals) gather_f64_kernel<<<grid_for(nnz), 256, 0, st>>>(dv2, dv, perm, nnz);
      fsb_count_launch(vals ? 3 : 2);
      e = cudaStreamSynchronize(st);
    }
    if (e != cudaSuccess) rc = fsb_cuda_error(e, "fsb_sort_blocked", __FILE__, __LINE__);
  }
  if (rc == FSB_OK) rc = fsb_d2h_scatter((void* const*)rows, dr2, b4.data(), nblocks, st);
  if (rc == FSB_OK) rc = fsb_d2h_scatter((void* const*)cols, dc2, b4.data(), nblocks, st);
  if (rc == FSB_OK && vals) rc = fsb_d2h_scatter((void* const*)vals, dv2, b8.data(), nblocks, st);
  if (rc == FSB_OK && cudaStreamSynchronize(st) != cudaSuccess) rc = fsb_cuda_error(cudaGetLastError(), "fsb_sort_blocked", __FILE__, __LINE__);
  cudaFree(dr); cudaFree(dc); cudaFree(dr2); cudaFree(dc2); cudaFree(dv); cudaFree(dv2); cudaFree(dsr); cudaFree(doff);
  cudaFree(keys); cudaFree(keys_sorted); cudaFree(iota); cudaFree(perm); cudaFree(tmp);
  return rc;
}

// the drop-in's choice for sort_bsbm / sort_bsbm_byrow / sort_bsdm: device at >= FSB_SORT_DEVICE_MIN entries when a device
// is present, else the bit-exact host routines block by block
extern "C" int fsb_sort_blocked_auto(int nrow, int ncol, int nblocks, const int* start_row, const int* blk_nnz, int* const* rows,
                                     int* const* cols, double* const* vals, int order) {
  if (nblocks < 0 || (order != 1 && order != 2)) return fsb_set_error(FSB_EINVAL, "fsb_sort_blocked_auto: bad arguments");
  static long min_dev = -1;
  if (min_dev < 0) {
    const char* e = getenv("FSB_SORT_DEVICE_MIN");
    min_dev = e ? atol(e) : (1L << 20);
  }
  long nnz = 0;
  for (int b = 0; b < nblocks; ++b) nnz += blk_nnz[b];
  if (nnz >= min_dev && fsb_device_count() > 0) return fsb_sort_blocked(nrow, ncol, nblocks, start_row, blk_nnz, rows, cols, vals, order);
  for (int b = 0; b < nblocks; ++b) {
    const int rc = order == 1 ? fsb_host_sort_block_hilbert(start_row[b], start_row[b + 1] - start_row[b], blk_nnz[b], rows[b], cols[b], vals ? vals[b] : nullptr)
                              : fsb_host_sort_block_byrow(ncol, blk_nnz[b], rows[b], cols[b]);
    if (rc != FSB_OK) return rc;
  }
  return FSB_OK;
}

// what the drop-in sort_sbm / sort_sdm call: the device path for large matrices when a device is present, the
// bit-exact serial host routine otherwise (sorting is construction, not the product path: both give the same order)
extern "C" int fsb_sort_coo_hilbert_auto(int nrow, int ncol, long nnz, int* rows, int* cols, double* vals) {
  static long min_dev = -1;
  if (min_dev < 0) {
    const char* e = getenv("FSB_SORT_DEVICE_MIN");
    min_dev = e ? atol(e) : (1L << 20);
  }
  if (nnz >= min_dev && fsb_device_count() > 0) return fsb_sort_coo_hilbert(nrow, ncol, nnz, rows, cols, vals);
  return fsb_host_sort_coo_hilbert(nrow, ncol, nnz, rows, cols, vals);
}

// Column-blocked binary CSR built on the device (new_cbcsr cbcsr.h:16-65): stable sort by cell.
extern "C" int fsb_cbcsr_from_coo_dev(fsb_matrix_t* out, int nrow, int ncol, long nnz, const int* d_rows, const int* d_cols,
                                      int colblocksize) {
  FSB_TRY(fsb_require_device());
  if (!out || nrow <= 0 || ncol <= 0 || nnz < 0 || colblocksize <= 0 || (nnz > 0 && (!d_rows || !d_cols)))
    return fsb_set_error(FSB_EINVAL, "fsb_cbcsr_from_coo_dev: bad arguments");
  const int nblocks = (ncol + colblocksize - 1) / colblocksize;
  const long ncell = (long)nblocks * nrow;
  if (ncell >= INT32_MAX) return fsb_set_error(FSB_EINVAL, "fsb_cbcsr_from_coo_dev: nblocks*nrow exceeds the int32 cell range (cbcsr.h:41)");
  cudaStream_t st = fsb_default_stream();
  if (nnz > 0) {   // cell keys are computed from both indices: check them first
    FSB_TRY(fsb_check_index_range(d_rows, nnz, nrow, "row index", st));
    FSB_TRY(fsb_check_index_range(d_cols, nnz, ncol, "column index", st));
  }
  int* keys = nullptr;
  FSB_CUDA(cudaMalloc(&keys, (size_t)std::max<long>(nnz, 1) * 4));
  if (nnz > 0) {
    cell_key_kernel<<<grid_for(nnz), 256, 0, st>>>(d_rows, d_cols, nnz, nrow, colblocksize, keys);
    fsb_count_launch();
  }
  fsb_matrix* A = new fsb_matrix();
  int rc = coo_to_csr_dev(A, (int)ncell, ncol, nnz, keys, d_cols, nullptr, st);
  cudaFree(keys);
  if (rc != FSB_OK) { fsb_matrix_free(A); return rc; }
  A->format = FSB_FMT_CBCSR; A->nrow = nrow; A->ncol = ncol; A->nblocks = nblocks; A->colblocksize = colblocksize;
  A->avg_row_nnz = (double)nnz / nrow;
  *out = A;
  return FSB_OK;
}

extern "C" int fsb_synth_coo_dev(unsigned long long seed, int dist, long nnz, int nrow, int ncol,
                                 int* d_rows, int* d_cols, double* d_vals, void* stream) {
  FSB_TRY(fsb_require_device());
  if (nnz < 0 || nrow <= 0 || ncol <= 0 || !d_rows || !d_cols) return fsb_set_error(FSB_EINVAL, "synth: bad arguments");
  if (nnz == 0) return FSB_OK;
  synth_kernel<<<grid_for(nnz), 256, 0, fsb_pick_stream(stream)>>>(seed, dist, nnz, nrow, ncol, d_rows, d_cols, d_vals);
  FSB_KERNEL_CHECK();
  return FSB_OK;
}

namespace {
__global__ void randn_kernel(double* __restrict__ d, long long n, unsigned long long seed) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) d[i] = fsb_synth_normal(seed, (unsigned long long)i);
}
}  // namespace

extern "C" int fsb_randn_dev(double* d, long n, unsigned long long seed, void* stream) {
  FSB_TRY(fsb_require_device());
  if (n < 0 || (n > 0 && !d)) return fsb_set_error(FSB_EINVAL, "fsb_randn_dev: bad arguments");
  if (n == 0) return FSB_OK;
  randn_kernel<<<grid_for(n), 256, 0, fsb_pick_stream(stream)>>>(d, n, seed);
  FSB_KERNEL_CHECK();
  return FSB_OK;
}

extern "C" int fsb_randn_host(double* x, long n, unsigned long long seed) {
  if (n < 0 || (n > 0 && !x)) return fsb_set_error(FSB_EINVAL, "fsb_randn_host: bad arguments");
  for (long i = 0; i < n; ++i) x[i] = fsb_synth_normal(seed, (unsigned long long)i);
  return FSB_OK;
}

extern "C" int fsb_synth_coo_host(unsigned long long seed, int dist, long nnz, int nrow, int ncol,
                                  int* rows, int* cols, double* vals) {
  if (nnz < 0 || nrow <= 0 || ncol <= 0 || !rows || !cols) return fsb_set_error(FSB_EINVAL, "synth: bad arguments");
  for (long j = 0; j < nnz; ++j) {
    rows[j] = fsb_synth_row(seed, (unsigned long long)j, nrow);
    cols[j] = fsb_synth_col(seed, (unsigned long long)j, ncol, dist);
    if (vals) vals[j] = fsb_synth_val(seed, (unsigned long long)j);
  }
  return FSB_OK;
}
