/* cg.h -- drop-in for libfastsparse's cg.h: the (A'A + lambda I) operator and the
 * (block) conjugate-gradient solver.  The whole solve runs on the GPU with every vector
 * resident in HBM (fsb_cg_host); bsbm_cg and bsbm_cg2 keep the reference's signatures,
 * stopping rules and iteration counter, and bsbm_cgn extends bsbm_cg2 to up to 32
 * right-hand sides. */
#ifndef CG_H
#define CG_H

#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "../fsb.h"
#include "linalg.h"
#include "sparse.h"

static inline void fsb_cg_check_pair_(struct BlockedSBM* A, struct BlockedSBM* At) {   /* cg.h:32-36 */
  if ((A->nrow != At->ncol) || (A->ncol != At->nrow)) {
    printf("A (%d x %d) and At (%d x %d) must be transposes of each other.\n", A->nrow, A->ncol, At->nrow, At->ncol);
    exit(1);
  }
}

/* y = At (A x) + lambda x; tmp (A->nrow doubles) receives A x (cg.h:9-22) */
static inline void bsbm_AtA(double* y, struct BlockedSBM* A, struct BlockedSBM* At, double* x, double* tmp, double lambda) {
  fsb_matrix_t ha, ht;
  FSB_DROPIN_CALL("bsbm_AtA",
                  (ha = fsb_cache_blocked(A->nrow, A->ncol, A->nblocks, A->start_row, A->nnz, A->rows, A->cols, NULL)) != NULL &&
                  (ht = fsb_cache_blocked(At->nrow, At->ncol, At->nblocks, At->start_row, At->nnz, At->rows, At->cols, NULL)) != NULL,
                  fsb_ata_pair_host(ha, ht, y, x, 1, lambda, tmp));
}

/* solves (A'A + lambda I) X = B for ncol <= 32 right-hand sides, X and B row-major [F][ncol] */
static inline void bsbm_cgn(double* X, struct BlockedSBM* A, struct BlockedSBM* At, double* B, int ncol, double lambda,
                            double tol, int* out_iter) {
  fsb_cg_check_pair_(A, At);
  fsb_matrix_t ha, ht;
  FSB_DROPIN_CALL("bsbm_cg",
                  (ha = fsb_cache_blocked(A->nrow, A->ncol, A->nblocks, A->start_row, A->nnz, A->rows, A->cols, NULL)) != NULL &&
                  (ht = fsb_cache_blocked(At->nrow, At->ncol, At->nblocks, At->start_row, At->nnz, At->rows, At->cols, NULL)) != NULL,
                  fsb_cg_host(ha, ht, X, B, ncol, lambda, tol, 0, out_iter));
}

/* one right-hand side (cg.h:25-82) */
static inline void bsbm_cg(double* x, struct BlockedSBM* A, struct BlockedSBM* At, double* b, double lambda, double tol, int* out_iter) {
  bsbm_cgn(x, A, At, b, 1, lambda, tol, out_iter);
}

/* two right-hand sides (cg.h:85-187) */
static inline void bsbm_cg2(double* X, struct BlockedSBM* A, struct BlockedSBM* At, double* B, double lambda, double tol, int* out_iter) {
  bsbm_cgn(X, A, At, B, 2, lambda, tol, out_iter);
}

#endif /* CG_H */
