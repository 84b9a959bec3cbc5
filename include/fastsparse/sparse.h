/* sparse.h -- drop-in for libfastsparse's sparse.h (binary COO + row-blocked COO).
 *
 * Same struct layouts, function names and argument order as the reference, so callers
 * (test_sparse.c, bench_a_mul_b.c, preprocess.c, Macau-style samplers) recompile
 * unchanged.  Construction and loading are host code inside libfastsparse_b200.so
 * (bit-exact structure); the Hilbert sorts run on the device above 2^20 entries (same
 * order; host routine below that or without a GPU); every product forwards to the GPU
 * through the C ABI in ../fsb.h.  There is no CPU fallback: without a CUDA device a
 * product prints the library error and exits, the reference's own error convention
 * (sparse.h:115-118).
 *
 * Residency: the structs are caller-owned and their layout is frozen, so the HBM copy
 * of a matrix is tracked out of band by fsb_cache_*() (keyed by the array addresses
 * and validated by a hash of the full content on every call -- the caller's arrays stay
 * the source of truth, see FSB_DROPIN_CALL in ../fsb.h); entry points that mutate a
 * structure drop its entry at once.
 */
#ifndef SPARSE_H
#define SPARSE_H

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../fsb.h"
#include "hilbert.h"
#include "quickSort.h"
#include "utils.h"

struct SparseBinaryMatrix {   /* sparse.h:11-18, sizeof 32 */
  int nrow;
  int ncol;
  long nnz;
  int* rows;
  int* cols;
};

/* adopts rows/cols (sparse.h:21-29) */
static inline struct SparseBinaryMatrix* new_sbm(long nrow, long ncol, long nnz, int* rows, int* cols) {
  struct SparseBinaryMatrix* A = (struct SparseBinaryMatrix*)malloc(sizeof *A);
  A->nrow = (int)nrow;
  A->ncol = (int)ncol;
  A->nnz = nnz;
  A->rows = rows;
  A->cols = cols;
  return A;
}

/* frees the arrays, never the struct (sparse.h:31-34) */
static inline void free_sbm(struct SparseBinaryMatrix* sbm) {
  fsb_cache_drop(sbm->rows);
  fsb_cache_drop(sbm->cols);
  free(sbm->rows);
  free(sbm->cols);
}

/* shallow transposed view: aliases the arrays (sparse.h:38-46) */
static inline struct SparseBinaryMatrix* new_transpose(struct SparseBinaryMatrix* A) {
  return new_sbm(A->ncol, A->nrow, A->nnz, A->cols, A->rows);
}

/* in-place transpose by pointer swap (sparse.h:48-55) */
static inline void transpose(struct SparseBinaryMatrix* A) {
  int* p = A->rows;
  A->rows = A->cols;
  A->cols = p;
  const int n = A->nrow;
  A->nrow = A->ncol;
  A->ncol = n;
}

/* y = A x on the GPU (replaces the serial COO loop sparse.h:58-65).  The COO is turned
 * into CSR on the device by a stable sort, so each y[r] sums the same terms. */
static inline void A_mul_B(double* y, struct SparseBinaryMatrix* A, double* x) {
  fsb_matrix_t h;
  FSB_DROPIN_CALL("A_mul_B", (h = fsb_cache_coo(A->nrow, A->ncol, A->nnz, A->rows, A->cols, NULL)) != NULL, fsb_spmm_host(h, y, x, 1));
}

/* y = A' x on the GPU (replaces sparse.h:68-75) */
static inline void At_mul_B(double* y, struct SparseBinaryMatrix* A, double* x) {
  fsb_matrix_t h;
  FSB_DROPIN_CALL("At_mul_B", (h = fsb_cache_coo(A->nrow, A->ncol, A->nnz, A->rows, A->cols, NULL)) != NULL, fsb_spmm_t_host(h, y, x, 1));
}

/* exponential variates and the geometric-skip subsampler (sparse.h:77-110): host-side
 * sampling helpers outside the sparse x dense path, kept so callers still link */
static inline double exprand(void) { return log1p(1.0 - drand48()); }
static inline double randexp(void) { return -log(1.0 - drand48()); }
static inline long randsubseq(long N, long max_samples, double p, long* samples) {
  const double scale = -1.0 / log1p(-p);
  long pos = -1, count = 0;
  while (count < max_samples) {
    const double gap = randexp() * scale;
    if (gap + pos >= N - 1) break;
    pos += (long)ceil(gap);
    samples[count++] = pos;
  }
  return count;
}

/* raw COO file: int64 nrow, ncol, nnz; int32 rows[nnz], cols[nnz], 1-based on disk
 * (sparse.h:112-139) */
static inline struct SparseBinaryMatrix* read_sbm(const char* filename) {
  long nrow = 0, ncol = 0, nnz = 0;
  if (fsb_host_read_coo(filename, &nrow, &ncol, &nnz, NULL, NULL, NULL)) {
    fprintf(stderr, "%s\n", fsb_last_error());
    exit(1);
  }
  int* rows = (int*)malloc((size_t)(nnz > 0 ? nnz : 1) * sizeof(int));
  int* cols = (int*)malloc((size_t)(nnz > 0 ? nnz : 1) * sizeof(int));
  if (fsb_host_read_coo(filename, &nrow, &ncol, &nnz, rows, cols, NULL)) {
    fprintf(stderr, "%s\n", fsb_last_error());
    exit(1);
  }
  return new_sbm(nrow, ncol, nnz, rows, cols);
}

/* global Hilbert order of the COO (sparse.h:142-161) */
static inline void sort_sbm(struct SparseBinaryMatrix* A) {
  fsb_cache_drop(A->rows);
  if (fsb_sort_coo_hilbert_auto(A->nrow, A->ncol, A->nnz, A->rows, A->cols, NULL)) fsb_die("sort_sbm");
}

struct BlockedSBM {           /* sparse.h:163-172, sizeof 48 */
  int nrow;
  int ncol;
  int nblocks;                /* row blocks */
  int* start_row;             /* nblocks + 1 */
  int* nnz;                   /* per block */
  int** rows;                 /* per block: global row ids */
  int** cols;
};

/* row blocks of block_size rows, COO order kept inside a block (sparse.h:175-213) */
static inline struct BlockedSBM* new_bsbm(struct SparseBinaryMatrix* A, int block_size) {
  struct BlockedSBM* B = (struct BlockedSBM*)malloc(sizeof *B);
  B->nrow = A->nrow;
  B->ncol = A->ncol;
  B->nblocks = fsb_host_blocked_nblocks(A->nrow, block_size);
  const size_t nb = (size_t)(B->nblocks > 0 ? B->nblocks : 1);
  B->nnz = (int*)malloc(nb * sizeof(int));
  B->start_row = (int*)malloc((nb + 1) * sizeof(int));
  B->rows = (int**)malloc(nb * sizeof(int*));
  B->cols = (int**)malloc(nb * sizeof(int*));
  if (fsb_host_blocked_count(A->nnz, A->nrow, block_size, A->rows, B->start_row, B->nnz)) fsb_die("new_bsbm");
  for (int b = 0; b < B->nblocks; b++) {
    const size_t m = (size_t)(B->nnz[b] > 0 ? B->nnz[b] : 1);
    B->rows[b] = (int*)malloc(m * sizeof(int));
    B->cols[b] = (int*)malloc(m * sizeof(int));
  }
  if (fsb_host_blocked_fill(A->nnz, block_size, A->rows, A->cols, NULL, B->nblocks, B->rows, B->cols, NULL)) fsb_die("new_bsbm");
  return B;
}

/* per block: Hilbert order inside n x n tiles along the row strip (sparse.h:215-236) */
static inline void sort_bsbm(struct BlockedSBM* B) {
  fsb_cache_drop(B->start_row);
  if (fsb_sort_blocked_auto(B->nrow, B->ncol, B->nblocks, B->start_row, B->nnz, B->rows, B->cols, NULL, 1)) fsb_die("sort_bsbm");
}

/* per block: row-major order (sparse.h:238-256) */
static inline void sort_bsbm_byrow(struct BlockedSBM* B) {
  fsb_cache_drop(B->start_row);
  if (fsb_sort_blocked_auto(B->nrow, B->ncol, B->nblocks, B->start_row, B->nnz, B->rows, B->cols, NULL, 2)) fsb_die("sort_bsbm_byrow");
}

/* Y = B X with ncol right-hand sides, row-major operands (sparse.h:318-336) */
static inline void bsbm_A_mul_Bn(double* y, struct BlockedSBM* B, double* x, int ncol) {
  fsb_matrix_t h;
  FSB_DROPIN_CALL("bsbm_A_mul_Bn", (h = fsb_cache_blocked(B->nrow, B->ncol, B->nblocks, B->start_row, B->nnz, B->rows, B->cols, NULL)) != NULL, fsb_spmm_host(h, y, x, ncol));
}
static inline void bsbm_A_mul_B(double* y, struct BlockedSBM* B, double* x) { bsbm_A_mul_Bn(y, B, x, 1); }   /* sparse.h:259-273 */
static inline void bsbm_A_mul_B2(double* y, struct BlockedSBM* B, double* x) { bsbm_A_mul_Bn(y, B, x, 2); }  /* sparse.h:276-293 */
static inline void bsbm_A_mul_B4(double* y, struct BlockedSBM* B, double* x) { bsbm_A_mul_Bn(y, B, x, 4); }  /* sparse.h:296-315 */

#endif /* SPARSE_H */
