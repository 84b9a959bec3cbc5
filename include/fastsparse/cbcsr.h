/* cbcsr.h -- drop-in for libfastsparse's cbcsr.h (column-blocked binary CSR).
 * Unlike the reference header this one includes what it needs. */
#ifndef CBCSR_H
#define CBCSR_H

#include <assert.h>
#include <stdlib.h>

#include "../fsb.h"
#include "sparse.h"

/*** ColBinaryCSR ***/
struct ColBinaryCSR {         /* cbcsr.h:5-14, sizeof 40 */
  int nrow;
  int ncol;
  int nblocks;
  int colblocksize;
  int nnz;
  int* row_ptr;               /* nblocks * nrow + 1 entries, cell = block * nrow + row */
  int* cols;
};

/* stable counting sort by cell (cbcsr.h:16-65) */
static inline void new_cbcsr(struct ColBinaryCSR* A, int colblocksize, long nnz, int nrow, int ncol, int* rows, int* cols) {
  assert(A);
  A->nnz = (int)nnz;
  A->nrow = nrow;
  A->ncol = ncol;
  A->nblocks = fsb_host_cbcsr_nblocks(ncol, colblocksize);
  A->colblocksize = colblocksize;
  A->cols = (int*)malloc((size_t)(nnz > 0 ? nnz : 1) * sizeof(int));
  A->row_ptr = (int*)malloc(((size_t)A->nblocks * nrow + 1) * sizeof(int));
  if (fsb_host_cbcsr_from_coo(colblocksize, nnz, nrow, ncol, rows, cols, A->row_ptr, A->cols)) fsb_die("new_cbcsr");
}

static inline void cbcsr_from_sbm(struct ColBinaryCSR* A, struct SparseBinaryMatrix* sbm, int colblocksize) {   /* cbcsr.h:67-73 */
  assert(A);
  assert(sbm);
  new_cbcsr(A, colblocksize, sbm->nnz, sbm->nrow, sbm->ncol, sbm->rows, sbm->cols);
}

/* Y = A X with ncol right-hand sides (new: the reference has one RHS only) */
static inline void cbcsr_A_mul_Bn(double* Y, struct ColBinaryCSR* A, double* X, int ncol) {
  fsb_matrix_t h;
  FSB_DROPIN_CALL("cbcsr_A_mul_Bn", (h = fsb_cache_cbcsr(A->nrow, A->ncol, A->nblocks, A->colblocksize, A->nnz, A->row_ptr, A->cols)) != NULL, fsb_spmm_host(h, Y, X, ncol));
}

/* y = A x (cbcsr.h:76-106) */
static inline void cbcsr_A_mul_B(double* y, struct ColBinaryCSR* A, double* x) { cbcsr_A_mul_Bn(y, A, x, 1); }

#endif /* CBCSR_H */
