/* dsparse.h -- drop-in for libfastsparse's dsparse.h (double-valued COO + row-blocked
 * COO).  Same layouts/names as the reference; host construction in
 * libfastsparse_b200.so, products on the GPU (see sparse.h in this directory). */
#ifndef DSPARSE_H
#define DSPARSE_H

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../fsb.h"
#include "hilbert.h"
#include "quickSortD.h"
#include "utils.h"

struct SparseDoubleMatrix {   /* dsparse.h:11-19, sizeof 40 */
  int nrow;
  int ncol;
  long nnz;
  int* rows;
  int* cols;
  double* vals;
};

/* adopts the three arrays (dsparse.h:22-31) */
static inline struct SparseDoubleMatrix* new_sdm(long nrow, long ncol, long nnz, int* rows, int* cols, double* vals) {
  struct SparseDoubleMatrix* A = (struct SparseDoubleMatrix*)malloc(sizeof *A);
  A->nrow = (int)nrow;
  A->ncol = (int)ncol;
  A->nnz = nnz;
  A->rows = rows;
  A->cols = cols;
  A->vals = vals;
  return A;
}

/* in-place transpose by pointer swap (dsparse.h:33-40) */
static inline void sdm_transpose(struct SparseDoubleMatrix* A) {
  int* p = A->rows;
  A->rows = A->cols;
  A->cols = p;
  const int n = A->nrow;
  A->nrow = A->ncol;
  A->ncol = n;
}

/* y = A x (replaces the serial loop dsparse.h:43-51) */
static inline void sdm_A_mul_B(double* y, struct SparseDoubleMatrix* A, double* x) {
  fsb_matrix_t h;
  FSB_DROPIN_CALL("sdm_A_mul_B", (h = fsb_cache_coo(A->nrow, A->ncol, A->nnz, A->rows, A->cols, A->vals)) != NULL, fsb_spmm_host(h, y, x, 1));
}

/* y = A' x (replaces dsparse.h:54-62) */
static inline void sdm_At_mul_B(double* y, struct SparseDoubleMatrix* A, double* x) {
  fsb_matrix_t h;
  FSB_DROPIN_CALL("sdm_At_mul_B", (h = fsb_cache_coo(A->nrow, A->ncol, A->nnz, A->rows, A->cols, A->vals)) != NULL, fsb_spmm_t_host(h, y, x, 1));
}

/* raw COO file with values (dsparse.h:64-93) */
static inline struct SparseDoubleMatrix* read_sdm(const char* filename) {
  long nrow = 0, ncol = 0, nnz = 0;
  if (fsb_host_read_coo(filename, &nrow, &ncol, &nnz, NULL, NULL, NULL)) {
    fprintf(stderr, "%s\n", fsb_last_error());
    exit(1);
  }
  const size_t n1 = (size_t)(nnz > 0 ? nnz : 1);
  int* rows = (int*)malloc(n1 * sizeof(int));
  int* cols = (int*)malloc(n1 * sizeof(int));
  double* vals = (double*)malloc(n1 * sizeof(double));
  if (fsb_host_read_coo(filename, &nrow, &ncol, &nnz, rows, cols, vals)) {
    fprintf(stderr, "%s\n", fsb_last_error());
    exit(1);
  }
  return new_sdm(nrow, ncol, nnz, rows, cols, vals);
}

/* global Hilbert order carrying the values (dsparse.h:96-115) */
static inline void sort_sdm(struct SparseDoubleMatrix* A) {
  fsb_cache_drop(A->rows);
  if (fsb_sort_coo_hilbert_auto(A->nrow, A->ncol, A->nnz, A->rows, A->cols, A->vals)) fsb_die("sort_sdm");
}

struct BlockedSDM {           /* dsparse.h:119-129 */
  int nrow;
  int ncol;
  int nblocks;
  int* start_row;
  int* nnz;
  int** rows;
  int** cols;
  double** vals;
};

/* dsparse.h:132-173 */
static inline struct BlockedSDM* new_bsdm(struct SparseDoubleMatrix* A, int block_size) {
  struct BlockedSDM* B = (struct BlockedSDM*)malloc(sizeof *B);
  B->nrow = A->nrow;
  B->ncol = A->ncol;
  B->nblocks = fsb_host_blocked_nblocks(A->nrow, block_size);
  const size_t nb = (size_t)(B->nblocks > 0 ? B->nblocks : 1);
  B->nnz = (int*)malloc(nb * sizeof(int));
  B->start_row = (int*)malloc((nb + 1) * sizeof(int));
  B->rows = (int**)malloc(nb * sizeof(int*));
  B->cols = (int**)malloc(nb * sizeof(int*));
  B->vals = (double**)malloc(nb * sizeof(double*));
  if (fsb_host_blocked_count(A->nnz, A->nrow, block_size, A->rows, B->start_row, B->nnz)) fsb_die("new_bsdm");
  for (int b = 0; b < B->nblocks; b++) {
    const size_t m = (size_t)(B->nnz[b] > 0 ? B->nnz[b] : 1);
    B->rows[b] = (int*)malloc(m * sizeof(int));
    B->cols[b] = (int*)malloc(m * sizeof(int));
    B->vals[b] = (double*)malloc(m * sizeof(double));
  }
  if (fsb_host_blocked_fill(A->nnz, block_size, A->rows, A->cols, A->vals, B->nblocks, B->rows, B->cols, B->vals)) fsb_die("new_bsdm");
  return B;
}

/* y = B x (dsparse.h:176-191) */
static inline void bsdm_A_mul_B(double* y, struct BlockedSDM* B, double* x) {
  fsb_matrix_t h;
  FSB_DROPIN_CALL("bsdm_A_mul_B", (h = fsb_cache_blocked(B->nrow, B->ncol, B->nblocks, B->start_row, B->nnz, B->rows, B->cols, B->vals)) != NULL, fsb_spmm_host(h, y, x, 1));
}

/* n right-hand sides on the double-valued blocked format (no reference counterpart) */
static inline void bsdm_A_mul_Bn(double* y, struct BlockedSDM* B, double* x, int ncol) {
  fsb_matrix_t h;
  FSB_DROPIN_CALL("bsdm_A_mul_Bn", (h = fsb_cache_blocked(B->nrow, B->ncol, B->nblocks, B->start_row, B->nnz, B->rows, B->cols, B->vals)) != NULL, fsb_spmm_host(h, y, x, ncol));
}

/* per-block Hilbert order carrying the values (dsparse.h:193-216) */
static inline void sort_bsdm(struct BlockedSDM* B) {
  fsb_cache_drop(B->start_row);
  if (fsb_sort_blocked_auto(B->nrow, B->ncol, B->nblocks, B->start_row, B->nnz, B->rows, B->cols, B->vals, 1)) fsb_die("sort_bsdm");
}

#endif /* DSPARSE_H */
