/* csr.h -- drop-in for libfastsparse's csr.h (binary CSR, double CSR, .csr.bin).
 *
 * struct BinaryCSR / struct CSR keep the reference layout byte for byte (struct
 * BinaryCSR is dumped raw into .csr.bin, csr.h:104).  new_bcsr / new_csr run the
 * library's stable counting sort on the host (bit-exact); every A_mul_B* call
 * forwards to the sm_100a SpMM kernel through the C ABI in ../fsb.h, uploading the
 * matrix on first use (fsb_cache_csr).  No CPU fallback.
 */
#ifndef CSR_H
#define CSR_H

#include <assert.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../fsb.h"
#include "omp_util.h"
#include "quickSort.h"
#include "sparse.h"

/*** binary CSR ***/
struct BinaryCSR {            /* csr.h:15-22, sizeof 32 */
  int nrow;
  int ncol;
  long nnz;
  int* row_ptr;               /* nrow + 1 row starts */
  int* cols;                  /* column ids, COO order kept inside a row */
};

static inline void free_bcsr(struct BinaryCSR* bcsr) {   /* csr.h:24-28 */
  assert(bcsr);
  fsb_cache_drop(bcsr->row_ptr);
  free(bcsr->row_ptr);
  free(bcsr->cols);
}

/* stable counting sort of the COO by row; copies (csr.h:30-67) */
static inline void new_bcsr(struct BinaryCSR* A, long nnz, int nrow, int ncol, int* rows, int* cols) {
  assert(A);
  A->nnz = nnz;
  A->nrow = nrow;
  A->ncol = ncol;
  A->cols = (int*)malloc((size_t)(nnz > 0 ? nnz : 1) * sizeof(int));
  A->row_ptr = (int*)malloc(((size_t)nrow + 1) * sizeof(int));
  if (fsb_host_csr_from_coo(nnz, nrow, rows, cols, NULL, A->row_ptr, A->cols, NULL)) fsb_die("new_bcsr");
}

static inline void bcsr_from_sbm(struct BinaryCSR* A, struct SparseBinaryMatrix* sbm) {   /* csr.h:69-74 */
  assert(A);
  assert(sbm);
  new_bcsr(A, sbm->nnz, sbm->nrow, sbm->ncol, sbm->rows, sbm->cols);
}

/* ---- .csr.bin (csr.h:83-146): tag line, raw 32-byte struct, int[nrow+1], int[nnz] ---- */
#define BINARY_CSR_HEADER "BINARY_CSR: struct BinaryCSR, int[nrow], int[nnz]\n"

static inline void serialize_to_file(const struct BinaryCSR* bcsr, const char* filename) {
  if (fsb_host_write_csr_bin(filename, bcsr, bcsr->nrow, bcsr->nnz, bcsr->row_ptr, bcsr->cols)) fsb_die("serialize_to_file");
}

/* allocates row_ptr / cols, overwriting whatever pointers *bcsr held; a malformed file
 * prints the reference's message and exits with -1 (csr.h:89-94) */
static inline void deserialize_from_file(struct BinaryCSR* bcsr, const char* filename) {
  if (fsb_host_read_csr_bin(filename, bcsr, NULL, NULL)) {
    printf("%s\n", fsb_last_error());
    exit(-1);
  }
  bcsr->row_ptr = (int*)calloc((size_t)bcsr->nrow + 1, sizeof(int));
  bcsr->cols = (int*)calloc((size_t)(bcsr->nnz > 0 ? bcsr->nnz : 1), sizeof(int));
  assert(bcsr->row_ptr && bcsr->cols);
  int* rp = bcsr->row_ptr;
  int* cc = bcsr->cols;
  if (fsb_host_read_csr_bin(filename, bcsr, rp, cc)) {
    printf("%s\n", fsb_last_error());
    exit(-1);
  }
  bcsr->row_ptr = rp;   /* the struct image on disk carries the writer's stale pointers */
  bcsr->cols = cc;
}

/* ---- products: Y[nrow][R] = A X[ncol][R], row-major ("row-ordered") operands ---- */
static inline void bcsr_A_mul_Bn(double* Y, struct BinaryCSR* A, double* X, const int ncol) {   /* csr.h:257-280 */
  fsb_matrix_t h;
  FSB_DROPIN_CALL("bcsr_A_mul_Bn", (h = fsb_cache_csr(A->nrow, A->ncol, A->nnz, A->row_ptr, A->cols, NULL)) != NULL, fsb_spmm_host(h, Y, X, ncol));
}
static inline void bcsr_A_mul_B(double* y, struct BinaryCSR* A, double* x) { bcsr_A_mul_Bn(y, A, x, 1); }        /* csr.h:149-161 */
static inline void bcsr_A_mul_B2(double* Y, struct BinaryCSR* A, double* X) { bcsr_A_mul_Bn(Y, A, X, 2); }       /* csr.h:164-181 */
static inline void bcsr_A_mul_B4(double* Y, struct BinaryCSR* A, double* X) { bcsr_A_mul_Bn(Y, A, X, 4); }       /* csr.h:184-202 */
static inline void bcsr_A_mul_B8(double* Y, struct BinaryCSR* A, double* X) { bcsr_A_mul_Bn(Y, A, X, 8); }       /* csr.h:205-223 */
static inline void bcsr_A_mul_B8_auto(double* Y, struct BinaryCSR* A, double* X) { bcsr_A_mul_Bn(Y, A, X, 8); }  /* csr.h:225-254 */
static inline void bcsr_A_mul_B32n(double* Y, struct BinaryCSR* A, double* X, const int ncol) {                  /* csr.h:283-302 */
  assert(ncol <= 32);
  bcsr_A_mul_Bn(Y, A, X, ncol);
}

/* Y[ncol_A][R] = A' X[nrow][R]: CSR-side transposed product (new; the reference builds a
 * second CSR of the transposed COO instead, bench_a_mul_b.c:273-274) */
static inline void bcsr_At_mul_Bn(double* Y, struct BinaryCSR* A, double* X, const int ncol) {
  fsb_matrix_t h;
  FSB_DROPIN_CALL("bcsr_At_mul_Bn", (h = fsb_cache_csr(A->nrow, A->ncol, A->nnz, A->row_ptr, A->cols, NULL)) != NULL, fsb_spmm_t_host(h, Y, X, ncol));
}
static inline void bcsr_At_mul_B(double* y, struct BinaryCSR* A, double* x) { bcsr_At_mul_Bn(y, A, x, 1); }

/* y = A'A x: deterministic two-pass form (csr.h:305-319) */
static inline void bcsr_AA_mul_B(double* y, struct BinaryCSR* A, double* x) {
  fsb_matrix_t h;
  FSB_DROPIN_CALL("bcsr_AA_mul_B", (h = fsb_cache_csr(A->nrow, A->ncol, A->nnz, A->row_ptr, A->cols, NULL)) != NULL, fsb_ata_host(h, y, x, 1, 0.0, 0));
}

/* y = A'A x: fused gather + fp64 red.global.add scatter (csr.h:323-355).  ytmp, the
 * per-thread scratch of the CPU version, is accepted and ignored. */
static inline void parallel_bcsr_AA_mul_B(double* y, struct BinaryCSR* A, double* x, double* ytmp) {
  (void)ytmp;
  fsb_matrix_t h;
  FSB_DROPIN_CALL("parallel_bcsr_AA_mul_B", (h = fsb_cache_csr(A->nrow, A->ncol, A->nnz, A->row_ptr, A->cols, NULL)) != NULL, fsb_ata_host(h, y, x, 1, 0.0, 1));
}

/*** Double CSR ***/
struct CSR {                  /* csr.h:358-366, sizeof 40 */
  int nrow;
  int ncol;
  long nnz;
  int* row_ptr;
  int* cols;
  double* vals;
};

static inline void free_csr(struct CSR* csr) {   /* csr.h:368-373 */
  assert(csr);
  fsb_cache_drop(csr->row_ptr);
  free(csr->row_ptr);
  free(csr->cols);
  free(csr->vals);
}

/* csr.h:375-422 */
static inline void new_csr(struct CSR* A, long nnz, int nrow, int ncol, int* rows, int* cols, double* vals) {
  assert(A);
  A->nnz = nnz;
  A->nrow = nrow;
  A->ncol = ncol;
  const size_t n1 = (size_t)(nnz > 0 ? nnz : 1);
  A->cols = (int*)malloc(n1 * sizeof(int));
  A->vals = (double*)malloc(n1 * sizeof(double));
  A->row_ptr = (int*)malloc(((size_t)nrow + 1) * sizeof(int));
  if (fsb_host_csr_from_coo(nnz, nrow, rows, cols, vals, A->row_ptr, A->cols, A->vals)) fsb_die("new_csr");
}

static inline void csr_A_mul_Bn(double* Y, struct CSR* A, double* X, const int ncol) {   /* csr.h:441-465 */
  fsb_matrix_t h;
  FSB_DROPIN_CALL("csr_A_mul_Bn", (h = fsb_cache_csr(A->nrow, A->ncol, A->nnz, A->row_ptr, A->cols, A->vals)) != NULL, fsb_spmm_host(h, Y, X, ncol));
}
static inline void csr_A_mul_B(double* y, struct CSR* A, double* x) { csr_A_mul_Bn(y, A, x, 1); }   /* csr.h:425-438 */

static inline void csr_At_mul_Bn(double* Y, struct CSR* A, double* X, const int ncol) {   /* new, see bcsr_At_mul_Bn */
  fsb_matrix_t h;
  FSB_DROPIN_CALL("csr_At_mul_Bn", (h = fsb_cache_csr(A->nrow, A->ncol, A->nnz, A->row_ptr, A->cols, A->vals)) != NULL, fsb_spmm_t_host(h, Y, X, ncol));
}
static inline void csr_At_mul_B(double* y, struct CSR* A, double* x) { csr_At_mul_Bn(y, A, x, 1); }

#endif /* CSR_H */
