/*
 * ref_wrap.c -- flat-array C entry points around the UNMODIFIED reference
 * headers (TEST INFRASTRUCTURE ONLY).
 *
 * This file contains no reference code: it #includes the reference headers
 * from where they lie (-I/root/reference, see oracle/Makefile) and forwards.
 * It is compiled only when /root/reference is present, into
 * oracle/_ref/libfsref_<isa>.so (git-ignored, travels to the GPU box), and is
 * used (a) to pin oracle/fsoracle.c, (b) to generate tests/golden/*.npz, and
 * (c) as the "reference" CPU baseline in bench.py.
 *
 * Blocked matrices cross this boundary flattened (see fsoracle.h: fso_blocked);
 * the wrappers rebuild the reference's pointer-per-block structs on the fly.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <assert.h>

#include "sparse.h"
#include "quickSortD.h"
#include "dsparse.h"
#include "linalg.h"
#include "cg.h"
#include "csr.h"
#include "cbcsr.h"

#include "fsoracle.h" /* only for the fso_blocked layout */

int ref_num_threads(void) { return thread_limit(); }

/* ---- hilbert / sort ---- */
int  ref_ceilPower2(int x) { return ceilPower2(x); }
long ref_xy2d(int n, int x, int y) { return xy2d(n, x, y); }
void ref_d2xy(int n, long d, int* x, int* y) { d2xy(n, d, x, y); }
long ref_row_xy2d(int n, int x, int y) { return row_xy2d(n, x, y); }
void ref_row_d2xy(int n, long d, int* x, int* y) { row_d2xy(n, d, x, y); }
void ref_quickSort(long* a, long n) { quickSort(a, 0, n - 1); }
void ref_quickSortD(long* a, double* v, long n) { quickSortD(a, 0, n - 1, v); }

/* ---- builders ---- */
void ref_new_bcsr(long nnz, int nrow, int ncol, int* rows, int* cols,
                  int* row_ptr_out, int* cols_out) {
  struct BinaryCSR A;
  new_bcsr(&A, nnz, nrow, ncol, rows, cols);
  memcpy(row_ptr_out, A.row_ptr, ((size_t)nrow + 1) * sizeof(int));
  memcpy(cols_out, A.cols, (size_t)nnz * sizeof(int));
  free_bcsr(&A);
}

void ref_new_csr(long nnz, int nrow, int ncol, int* rows, int* cols, double* vals,
                 int* row_ptr_out, int* cols_out, double* vals_out) {
  struct CSR A;
  new_csr(&A, nnz, nrow, ncol, rows, cols, vals);
  memcpy(row_ptr_out, A.row_ptr, ((size_t)nrow + 1) * sizeof(int));
  memcpy(cols_out, A.cols, (size_t)nnz * sizeof(int));
  memcpy(vals_out, A.vals, (size_t)nnz * sizeof(double));
  free_csr(&A);
}

int ref_new_cbcsr(int colblocksize, long nnz, int nrow, int ncol, int* rows,
                  int* cols, int* row_ptr_out, int* cols_out) {
  struct ColBinaryCSR A;
  new_cbcsr(&A, colblocksize, nnz, nrow, ncol, rows, cols);
  if (row_ptr_out) {
    memcpy(row_ptr_out, A.row_ptr, ((size_t)A.nblocks * nrow + 1) * sizeof(int));
    memcpy(cols_out, A.cols, (size_t)nnz * sizeof(int));
  }
  free(A.row_ptr);
  free(A.cols);
  return A.nblocks;
}

/* build with the reference, then flatten into preallocated out arrays */
void ref_new_bsbm(long nnz, int nrow, int ncol, int* rows, int* cols,
                  int block_size, fso_blocked* out) {
  struct SparseBinaryMatrix A = {nrow, ncol, nnz, rows, cols};
  struct BlockedSBM* B = new_bsbm(&A, block_size);
  out->nrow = B->nrow; out->ncol = B->ncol; out->nblocks = B->nblocks;
  out->blk_off[0] = 0;
  for (int b = 0; b < B->nblocks; b++) {
    out->start_row[b] = B->start_row[b];
    out->blk_nnz[b] = B->nnz[b];
    out->blk_off[b + 1] = out->blk_off[b] + B->nnz[b];
    memcpy(out->rows + out->blk_off[b], B->rows[b], (size_t)B->nnz[b] * sizeof(int));
    memcpy(out->cols + out->blk_off[b], B->cols[b], (size_t)B->nnz[b] * sizeof(int));
    free(B->rows[b]); free(B->cols[b]);
  }
  out->start_row[B->nblocks] = B->start_row[B->nblocks];
  free(B->rows); free(B->cols); free(B->nnz); free(B->start_row); free(B);
}

void ref_new_bsdm(long nnz, int nrow, int ncol, int* rows, int* cols, double* vals,
                  int block_size, fso_blocked* out) {
  struct SparseDoubleMatrix A = {nrow, ncol, nnz, rows, cols, vals};
  struct BlockedSDM* B = new_bsdm(&A, block_size);
  out->nrow = B->nrow; out->ncol = B->ncol; out->nblocks = B->nblocks;
  out->blk_off[0] = 0;
  for (int b = 0; b < B->nblocks; b++) {
    out->start_row[b] = B->start_row[b];
    out->blk_nnz[b] = B->nnz[b];
    out->blk_off[b + 1] = out->blk_off[b] + B->nnz[b];
    memcpy(out->rows + out->blk_off[b], B->rows[b], (size_t)B->nnz[b] * sizeof(int));
    memcpy(out->cols + out->blk_off[b], B->cols[b], (size_t)B->nnz[b] * sizeof(int));
    memcpy(out->vals + out->blk_off[b], B->vals[b], (size_t)B->nnz[b] * sizeof(double));
    free(B->rows[b]); free(B->cols[b]); free(B->vals[b]);
  }
  out->start_row[B->nblocks] = B->start_row[B->nblocks];
  free(B->rows); free(B->cols); free(B->vals); free(B->nnz); free(B->start_row); free(B);
}

/* views of a flat blocked matrix as the reference's structs */
static struct BlockedSBM view_bsbm(const fso_blocked* F) {
  struct BlockedSBM B;
  B.nrow = F->nrow; B.ncol = F->ncol; B.nblocks = F->nblocks;
  B.start_row = F->start_row; B.nnz = F->blk_nnz;
  B.rows = (int**)malloc((size_t)(F->nblocks + 1) * sizeof(int*));
  B.cols = (int**)malloc((size_t)(F->nblocks + 1) * sizeof(int*));
  for (int b = 0; b < F->nblocks; b++) {
    B.rows[b] = F->rows + F->blk_off[b];
    B.cols[b] = F->cols + F->blk_off[b];
  }
  return B;
}
static void drop_bsbm(struct BlockedSBM* B) { free(B->rows); free(B->cols); }

static struct BlockedSDM view_bsdm(const fso_blocked* F) {
  struct BlockedSDM B;
  B.nrow = F->nrow; B.ncol = F->ncol; B.nblocks = F->nblocks;
  B.start_row = F->start_row; B.nnz = F->blk_nnz;
  B.rows = (int**)malloc((size_t)(F->nblocks + 1) * sizeof(int*));
  B.cols = (int**)malloc((size_t)(F->nblocks + 1) * sizeof(int*));
  B.vals = (double**)malloc((size_t)(F->nblocks + 1) * sizeof(double*));
  for (int b = 0; b < F->nblocks; b++) {
    B.rows[b] = F->rows + F->blk_off[b];
    B.cols[b] = F->cols + F->blk_off[b];
    B.vals[b] = F->vals + F->blk_off[b];
  }
  return B;
}
static void drop_bsdm(struct BlockedSDM* B) { free(B->rows); free(B->cols); free(B->vals); }

void ref_sort_sbm(int nrow, int ncol, long nnz, int* rows, int* cols) {
  struct SparseBinaryMatrix A = {nrow, ncol, nnz, rows, cols};
  sort_sbm(&A);
}
void ref_sort_sdm(int nrow, int ncol, long nnz, int* rows, int* cols, double* vals) {
  struct SparseDoubleMatrix A = {nrow, ncol, nnz, rows, cols, vals};
  sort_sdm(&A);
}
void ref_sort_bsbm(fso_blocked* F) { struct BlockedSBM B = view_bsbm(F); sort_bsbm(&B); drop_bsbm(&B); }
void ref_sort_bsbm_byrow(fso_blocked* F) { struct BlockedSBM B = view_bsbm(F); sort_bsbm_byrow(&B); drop_bsbm(&B); }
void ref_sort_bsdm(fso_blocked* F) { struct BlockedSDM B = view_bsdm(F); sort_bsdm(&B); drop_bsdm(&B); }

/* ---- COO products ---- */
void ref_A_mul_B(double* y, int nrow, int ncol, long nnz, int* rows, int* cols, double* x) {
  struct SparseBinaryMatrix A = {nrow, ncol, nnz, rows, cols};
  A_mul_B(y, &A, x);
}
void ref_At_mul_B(double* y, int nrow, int ncol, long nnz, int* rows, int* cols, double* x) {
  struct SparseBinaryMatrix A = {nrow, ncol, nnz, rows, cols};
  At_mul_B(y, &A, x);
}
void ref_sdm_A_mul_B(double* y, int nrow, int ncol, long nnz, int* rows, int* cols, double* vals, double* x) {
  struct SparseDoubleMatrix A = {nrow, ncol, nnz, rows, cols, vals};
  sdm_A_mul_B(y, &A, x);
}
void ref_sdm_At_mul_B(double* y, int nrow, int ncol, long nnz, int* rows, int* cols, double* vals, double* x) {
  struct SparseDoubleMatrix A = {nrow, ncol, nnz, rows, cols, vals};
  sdm_At_mul_B(y, &A, x);
}

/* ---- CSR products; which: 1,2,4,8 fixed-R, 80 = _B8_auto, 0 = _Bn, 32 = _B32n ---- */
void ref_bcsr_mul(int which, double* Y, int nrow, int ncol, long nnz, int* row_ptr,
                  int* cols, double* X, int R) {
  struct BinaryCSR A = {nrow, ncol, nnz, row_ptr, cols};
  switch (which) {
    case 1: bcsr_A_mul_B(Y, &A, X); break;
    case 2: bcsr_A_mul_B2(Y, &A, X); break;
    case 4: bcsr_A_mul_B4(Y, &A, X); break;
    case 8: bcsr_A_mul_B8(Y, &A, X); break;
    case 80: bcsr_A_mul_B8_auto(Y, &A, X); break;
    case 0: bcsr_A_mul_Bn(Y, &A, X, R); break;
    case 32: bcsr_A_mul_B32n(Y, &A, X, R); break;
    default: fprintf(stderr, "ref_bcsr_mul: bad selector %d\n", which); exit(2);
  }
}
void ref_bcsr_AA_mul_B(double* y, int nrow, int ncol, long nnz, int* row_ptr, int* cols, double* x) {
  struct BinaryCSR A = {nrow, ncol, nnz, row_ptr, cols};
  bcsr_AA_mul_B(y, &A, x);
}
void ref_parallel_bcsr_AA_mul_B(double* y, int nrow, int ncol, long nnz, int* row_ptr, int* cols, double* x) {
  struct BinaryCSR A = {nrow, ncol, nnz, row_ptr, cols};
  double* ytmp = (double*)malloc((size_t)ncol * thread_limit() * sizeof(double));
  parallel_bcsr_AA_mul_B(y, &A, x, ytmp);
  free(ytmp);
}
/* same, with caller-provided scratch so that a timing loop excludes malloc */
void ref_parallel_bcsr_AA_mul_B_scratch(double* y, int nrow, int ncol, long nnz, int* row_ptr, int* cols, double* x, double* ytmp) {
  struct BinaryCSR A = {nrow, ncol, nnz, row_ptr, cols};
  parallel_bcsr_AA_mul_B(y, &A, x, ytmp);
}
void ref_csr_mul(int which, double* Y, int nrow, int ncol, long nnz, int* row_ptr,
                 int* cols, double* vals, double* X, int R) {
  struct CSR A = {nrow, ncol, nnz, row_ptr, cols, vals};
  if (which == 1) csr_A_mul_B(Y, &A, X); else csr_A_mul_Bn(Y, &A, X, R);
}
void ref_cbcsr_A_mul_B(double* y, int nrow, int ncol, int nblocks, int colblocksize,
                       long nnz, int* row_ptr, int* cols, double* x) {
  struct ColBinaryCSR A = {nrow, ncol, nblocks, colblocksize, (int)nnz, row_ptr, cols};
  cbcsr_A_mul_B(y, &A, x);
}

/* ---- blocked products; which: 1,2,4 fixed, 0 = _Bn ---- */
void ref_bsbm_mul(int which, double* Y, const fso_blocked* F, double* X, int R) {
  struct BlockedSBM B = view_bsbm(F);
  switch (which) {
    case 1: bsbm_A_mul_B(Y, &B, X); break;
    case 2: bsbm_A_mul_B2(Y, &B, X); break;
    case 4: bsbm_A_mul_B4(Y, &B, X); break;
    default: bsbm_A_mul_Bn(Y, &B, X, R); break;
  }
  drop_bsbm(&B);
}
void ref_bsdm_A_mul_B(double* y, const fso_blocked* F, double* x) {
  struct BlockedSDM B = view_bsdm(F);
  bsdm_A_mul_B(y, &B, x);
  drop_bsdm(&B);
}

/* ---- linalg ---- */
double ref_dist(double* x, double* y, int n) { return dist(x, y, n); }
double ref_pnormsq(double* x, int n) { return pnormsq(x, n); }
double ref_pdot(double* x, double* y, int n) { return pdot(x, y, n); }
void ref_pnormsq2(double* o, double* X, int n) { pnormsq2(o, X, n); }
void ref_pouter2(double* o, double* X, int n) { pouter2(o, X, n); }
void ref_pdot2sym(double* o, double* X, double* Y, int n) { pdot2sym(o, X, Y, n); }
void ref_solve2sym(double* X, double* A, double* RHS) { solve2sym(X, A, RHS); }

/* ---- solver ---- */
void ref_bsbm_AtA(double* y, const fso_blocked* FA, const fso_blocked* FAt, double* x, double* tmp, double lambda) {
  struct BlockedSBM A = view_bsbm(FA), At = view_bsbm(FAt);
  bsbm_AtA(y, &A, &At, x, tmp, lambda);
  drop_bsbm(&A); drop_bsbm(&At);
}
int ref_bsbm_cg(double* x, const fso_blocked* FA, const fso_blocked* FAt, double* b, double lambda, double tol) {
  struct BlockedSBM A = view_bsbm(FA), At = view_bsbm(FAt);
  int it = -1;
  bsbm_cg(x, &A, &At, b, lambda, tol, &it);
  drop_bsbm(&A); drop_bsbm(&At);
  return it;
}
int ref_bsbm_cg2(double* X, const fso_blocked* FA, const fso_blocked* FAt, double* B, double lambda, double tol) {
  struct BlockedSBM A = view_bsbm(FA), At = view_bsbm(FAt);
  int it = -1;
  bsbm_cg2(X, &A, &At, B, lambda, tol, &it);
  drop_bsbm(&A); drop_bsbm(&At);
  return it;
}

/* ---- files ---- */
/* two-call protocol: rows == NULL returns sizes only */
void ref_read_sbm(const char* path, long* nrow, long* ncol, long* nnz, int* rows, int* cols) {
  struct SparseBinaryMatrix* A = read_sbm(path);
  *nrow = A->nrow; *ncol = A->ncol; *nnz = A->nnz;
  if (rows) {
    memcpy(rows, A->rows, (size_t)A->nnz * sizeof(int));
    memcpy(cols, A->cols, (size_t)A->nnz * sizeof(int));
  }
  free_sbm(A); free(A);
}
void ref_read_sdm(const char* path, long* nrow, long* ncol, long* nnz, int* rows, int* cols, double* vals) {
  struct SparseDoubleMatrix* A = read_sdm(path);
  *nrow = A->nrow; *ncol = A->ncol; *nnz = A->nnz;
  if (rows) {
    memcpy(rows, A->rows, (size_t)A->nnz * sizeof(int));
    memcpy(cols, A->cols, (size_t)A->nnz * sizeof(int));
    memcpy(vals, A->vals, (size_t)A->nnz * sizeof(double));
  }
  free(A->rows); free(A->cols); free(A->vals); free(A);
}
void ref_serialize_to_file(const char* path, int nrow, int ncol, long nnz, int* row_ptr, int* cols) {
  struct BinaryCSR A = {nrow, ncol, nnz, row_ptr, cols};
  serialize_to_file(&A, path);
}
void ref_deserialize_from_file(const char* path, int* nrow, int* ncol, long* nnz, int* row_ptr, int* cols) {
  struct BinaryCSR A;
  deserialize_from_file(&A, path);
  *nrow = A.nrow; *ncol = A.ncol; *nnz = A.nnz;
  if (row_ptr) {
    memcpy(row_ptr, A.row_ptr, ((size_t)A.nrow + 1) * sizeof(int));
    memcpy(cols, A.cols, (size_t)A.nnz * sizeof(int));
  }
  free_bcsr(&A);
}
