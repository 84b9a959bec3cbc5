/*
 * fsoracle.c -- CPU oracle (TEST INFRASTRUCTURE ONLY, see fsoracle.h).
 *
 * Plain C99 restatement of the reference's algorithms.  Written from the
 * reference's documented behaviour, not copied: flat arrays instead of
 * per-block mallocs, one generic n-RHS routine instead of the 1/2/4/8/n
 * unrolled family (their arithmetic is identical per output element: an
 * in-order sum over the row's stored entries).
 *
 * Build: gcc -O2 -fopenmp -fPIC -shared (NO -ffast-math) -- see Makefile.
 * OpenMP is used only across independent rows / blocks, so every output
 * element is still summed serially in the reference's loop order.
 */
#include "fsoracle.h"

#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int fso_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* ------------------------------------------------------------------ */
/* Hilbert maths                                                       */
/* ------------------------------------------------------------------ */

/* hilbert.h:11-13 -- "1 << (int)ceil(log2(x))".  For every int x >= 1 this is
 * the smallest power of two >= x (pinned by test_sparse.c:195-203, incl.
 * 2^30-1 -> 2^30); restated in integers so it cannot depend on libm. */
int fso_ceil_pow2(int x) {
  int p = 1;
  while (p < x) p <<= 1;
  return p;
}

/* hilbert.h:45-57 -- quadrant rotate/flip.  NOTE the reference passes the
 * current sub-square size s (not n) and does not mask x,y first, so values may
 * leave [0,s); two's-complement int arithmetic is part of the contract. */
static void hil_rot(int s, int* x, int* y, int rx, int ry) {
  if (ry != 0) return;
  if (rx == 1) {
    *x = s - 1 - *x;
    *y = s - 1 - *y;
  }
  int t = *x;
  *x = *y;
  *y = t;
}

/* hilbert.h:16-27 */
long fso_xy2d(int n, int x, int y) {
  long d = 0;
  for (long s = n / 2; s > 0; s /= 2) {
    int rx = (x & s) > 0;
    int ry = (y & s) > 0;
    d += s * s * (long)((3 * rx) ^ ry);
    hil_rot((int)s, &x, &y, rx, ry);
  }
  return d;
}

/* hilbert.h:30-42 */
void fso_d2xy(int n, long d, int* x, int* y) {
  long t = d;
  *x = 0;
  *y = 0;
  for (int s = 1; s < n; s *= 2) {
    int rx = (int)(1 & (t / 2));
    int ry = (int)(1 & (t ^ rx));
    hil_rot(s, x, y, rx, ry);
    *x += s * rx;
    *y += s * ry;
    t /= 4;
  }
}

/* hilbert.h:60-65 -- row strip of n rows tiled into n x n squares by column;
 * inside a square the (x,y) roles are swapped. */
long fso_row_xy2d(int n, int x, int y) {
  long nsq = (long)n * (long)n;
  return fso_xy2d(n, y % n, x) + nsq * (long)(y / n);
}

/* hilbert.h:68-75 */
void fso_row_d2xy(int n, long d, int* x, int* y) {
  long nsq = (long)n * (long)n;
  long h = d % nsq;
  int tile = (int)(d / nsq);
  fso_d2xy(n, h, y, x);
  *y += tile * n;
}

/* ------------------------------------------------------------------ */
/* Sorting: quickSort.h:10-57 and quickSortD.h:12-71                    */
/* Same pivot rule (middle element moved to the left end), same scan,   */
/* same insertion-sort cutoff (r - l < 10), so that even the unstable   */
/* placement of payloads under duplicate keys matches the reference.    */
/* vals may be NULL (keys only).                                        */
/* ------------------------------------------------------------------ */
static void ins_sort(long* a, double* v, long lo, long hi) {
  for (long k = lo + 1; k <= hi; k++) {
    long key = a[k];
    double pv = v ? v[k] : 0.0;
    long j = k - 1;
    /* the reference scans down to index 0, but everything left of lo is
     * <= every element of [lo,hi] once partitioned, so stopping at lo is
     * equivalent for the sub-ranges quicksort produces */
    while (j >= lo && key < a[j]) {
      a[j + 1] = a[j];
      if (v) v[j + 1] = v[j];
      j--;
    }
    a[j + 1] = key;
    if (v) v[j + 1] = pv;
  }
}

static long qs_partition(long* a, double* v, long l, long r) {
  long m = (l + r) / 2;
  long pivot = a[m];
  a[m] = a[l];
  a[l] = pivot;
  if (v) { double t = v[m]; v[m] = v[l]; v[l] = t; }
  long i = l, j = r + 1;
  for (;;) {
    do { ++i; } while (i <= r && a[i] <= pivot);
    do { --j; } while (a[j] > pivot);
    if (i >= j) break;
    long t = a[i]; a[i] = a[j]; a[j] = t;
    if (v) { double tv = v[i]; v[i] = v[j]; v[j] = tv; }
  }
  long t = a[l]; a[l] = a[j]; a[j] = t;
  if (v) { double tv = v[l]; v[l] = v[j]; v[j] = tv; }
  return j;
}

static void qs_rec(long* a, double* v, long l, long r) {
  while (r - l >= 10) {
    long j = qs_partition(a, v, l, r);
    /* recurse on the left part, iterate on the right (same result as the
     * reference's two recursive calls, bounded stack) */
    qs_rec(a, v, l, j - 1);
    l = j + 1;
  }
  ins_sort(a, v, l, r);
}

void fso_sort_keys(long* keys, long n) { if (n > 1) qs_rec(keys, NULL, 0, n - 1); }
void fso_sort_keys_vals(long* keys, double* vals, long n) {
  if (n > 1) qs_rec(keys, vals, 0, n - 1);
}

/* ------------------------------------------------------------------ */
/* Structure builders                                                  */
/* ------------------------------------------------------------------ */

/* csr.h:30-67 (new_bcsr) and csr.h:375-422 (new_csr): stable counting sort of
 * the COO entries by row; inside a row the COO input order is kept;
 * duplicates are kept. */
void fso_csr_from_coo(long nnz, int nrow, const int* rows, const int* cols,
                      const double* vals, int* row_ptr, int* out_cols,
                      double* out_vals) {
  int* cursor = (int*)calloc((size_t)nrow + 1, sizeof(int));
  for (long i = 0; i < nnz; i++) cursor[rows[i]]++;
  int run = 0;
  for (int r = 0; r < nrow; r++) {
    row_ptr[r] = run;
    run += cursor[r];
    cursor[r] = row_ptr[r];
  }
  row_ptr[nrow] = (int)nnz;
  for (long i = 0; i < nnz; i++) {
    int dst = cursor[rows[i]]++;
    out_cols[dst] = cols[i];
    if (vals) out_vals[dst] = vals[i];
  }
  free(cursor);
}

/* cbcsr.h:27 */
int fso_cbcsr_nblocks(int ncol, int colblocksize) {
  return (int)ceil(ncol / (double)colblocksize);
}

/* cbcsr.h:16-65: stable counting sort by cell = (col / colblocksize) * nrow + row */
void fso_cbcsr_from_coo(int colblocksize, long nnz, int nrow, int ncol,
                        const int* rows, const int* cols, int* row_ptr,
                        int* out_cols) {
  int nblocks = fso_cbcsr_nblocks(ncol, colblocksize);
  long ncell = (long)nblocks * nrow;
  int* cursor = (int*)calloc((size_t)ncell + 1, sizeof(int));
  for (long i = 0; i < nnz; i++) cursor[(long)(cols[i] / colblocksize) * nrow + rows[i]]++;
  int run = 0;
  for (long c = 0; c < ncell; c++) {
    row_ptr[c] = run;
    run += cursor[c];
    cursor[c] = row_ptr[c];
  }
  row_ptr[ncell] = (int)nnz;
  for (long i = 0; i < nnz; i++) {
    long cell = (long)(cols[i] / colblocksize) * nrow + rows[i];
    out_cols[cursor[cell]++] = cols[i];
  }
  free(cursor);
}

/* sparse.h:179 / dsparse.h:136 */
int fso_blocked_nblocks(int nrow, int block_size) {
  return (int)ceil(nrow / (double)block_size);
}

/* sparse.h:175-213 (new_bsbm), dsparse.h:132-173 (new_bsdm): bucket the COO
 * entries by row / block_size keeping COO order inside a bucket. */
void fso_blocked_from_coo(long nnz, int nrow, int ncol, int block_size,
                          const int* rows, const int* cols, const double* vals,
                          fso_blocked* B) {
  int nb = fso_blocked_nblocks(nrow, block_size);
  B->nrow = nrow;
  B->ncol = ncol;
  B->nblocks = nb;
  for (int b = 0; b < nb; b++) {
    B->start_row[b] = b * block_size;
    B->blk_nnz[b] = 0;
  }
  B->start_row[nb] = nrow;
  for (long j = 0; j < nnz; j++) B->blk_nnz[rows[j] / block_size]++;
  B->blk_off[0] = 0;
  for (int b = 0; b < nb; b++) B->blk_off[b + 1] = B->blk_off[b] + B->blk_nnz[b];
  long* cur = (long*)malloc((size_t)(nb > 0 ? nb : 1) * sizeof(long));
  for (int b = 0; b < nb; b++) cur[b] = B->blk_off[b];
  for (long j = 0; j < nnz; j++) {
    long d = cur[rows[j] / block_size]++;
    B->rows[d] = rows[j];
    B->cols[d] = cols[j];
    if (vals) B->vals[d] = vals[j];
  }
  free(cur);
}

/* sparse.h:142-161 (sort_sbm), dsparse.h:96-115 (sort_sdm) */
void fso_sort_coo_hilbert(int nrow, int ncol, long nnz, int* rows, int* cols,
                          double* vals) {
  int n = fso_ceil_pow2(nrow > ncol ? nrow : ncol);
  long* h = (long*)malloc((size_t)(nnz > 0 ? nnz : 1) * sizeof(long));
  for (long j = 0; j < nnz; j++) h[j] = fso_xy2d(n, rows[j], cols[j]);
  if (vals) fso_sort_keys_vals(h, vals, nnz); else fso_sort_keys(h, nnz);
  for (long j = 0; j < nnz; j++) fso_d2xy(n, h[j], &rows[j], &cols[j]);
  free(h);
}

/* sparse.h:215-236 (sort_bsbm), dsparse.h:193-216 (sort_bsdm) */
void fso_sort_blocked_hilbert(fso_blocked* B) {
  for (int b = 0; b < B->nblocks; b++) {
    long off = B->blk_off[b];
    long m = B->blk_nnz[b];
    int r0 = B->start_row[b];
    int n = fso_ceil_pow2(B->start_row[b + 1] - r0);
    long* h = (long*)malloc((size_t)(m > 0 ? m : 1) * sizeof(long));
    for (long j = 0; j < m; j++) h[j] = fso_row_xy2d(n, B->rows[off + j] - r0, B->cols[off + j]);
    if (B->vals) fso_sort_keys_vals(h, B->vals + off, m); else fso_sort_keys(h, m);
    for (long j = 0; j < m; j++) {
      fso_row_d2xy(n, h[j], &B->rows[off + j], &B->cols[off + j]);
      B->rows[off + j] += r0;
    }
    free(h);
  }
}

/* sparse.h:238-256 (sort_bsbm_byrow): key = row * ncol + col */
void fso_sort_blocked_byrow(fso_blocked* B) {
  for (int b = 0; b < B->nblocks; b++) {
    long off = B->blk_off[b];
    long m = B->blk_nnz[b];
    long* h = (long*)malloc((size_t)(m > 0 ? m : 1) * sizeof(long));
    for (long j = 0; j < m; j++) h[j] = B->rows[off + j] * (long)B->ncol + (long)B->cols[off + j];
    if (B->vals) fso_sort_keys_vals(h, B->vals + off, m); else fso_sort_keys(h, m);
    for (long j = 0; j < m; j++) {
      B->rows[off + j] = (int)(h[j] / B->ncol);
      B->cols[off + j] = (int)(h[j] % B->ncol);
    }
    free(h);
  }
}

/* ------------------------------------------------------------------ */
/* Products                                                            */
/* ------------------------------------------------------------------ */

/* sparse.h:58-65 (A_mul_B), dsparse.h:43-51 (sdm_A_mul_B): serial COO scatter */
void fso_coo_A_mul_B(double* y, int nrow, long nnz, const int* rows,
                     const int* cols, const double* vals, const double* x) {
  memset(y, 0, (size_t)nrow * sizeof(double));
  if (vals) for (long j = 0; j < nnz; j++) y[rows[j]] += x[cols[j]] * vals[j];
  else      for (long j = 0; j < nnz; j++) y[rows[j]] += x[cols[j]];
}

/* sparse.h:68-75 (At_mul_B), dsparse.h:54-62 (sdm_At_mul_B) */
void fso_coo_At_mul_B(double* y, int ncol, long nnz, const int* rows,
                      const int* cols, const double* vals, const double* x) {
  memset(y, 0, (size_t)ncol * sizeof(double));
  if (vals) for (long j = 0; j < nnz; j++) y[cols[j]] += x[rows[j]] * vals[j];
  else      for (long j = 0; j < nnz; j++) y[cols[j]] += x[rows[j]];
}

/* csr.h:149-161 (bcsr_A_mul_B), 164-254 (_B2/_B4/_B8/_B8_auto), 257-280
 * (_Bn), 283-302 (_B32n), 425-438 (csr_A_mul_B), 441-465 (csr_A_mul_Bn):
 * Y[r, k] = sum over the row's stored entries, in stored order, of
 * X[col, k] (* val).  All of the reference variants perform exactly this sum
 * per output element; they differ only in unrolling and threading. */
void fso_csr_A_mul_Bn(double* Y, int nrow, const int* row_ptr, const int* cols,
                      const double* vals, const double* X, int R) {
#pragma omp parallel for schedule(dynamic, 256)
  for (int r = 0; r < nrow; r++) {
    double* yr = Y + (size_t)r * R;
    for (int k = 0; k < R; k++) yr[k] = 0.0;
    for (int i = row_ptr[r]; i < row_ptr[r + 1]; i++) {
      const double* xr = X + (size_t)cols[i] * R;
      if (vals) {
        double v = vals[i];
        for (int k = 0; k < R; k++) yr[k] += xr[k] * v;
      } else {
        for (int k = 0; k < R; k++) yr[k] += xr[k];
      }
    }
  }
}

/* csr.h:305-319 (bcsr_AA_mul_B): y = A'(A x) fused row by row, serial.
 * (parallel_bcsr_AA_mul_B csr.h:323-355 computes the same sums with a
 * per-thread partial y; only the association order differs.) */
void fso_bcsr_AA_mul_B(double* y, int nrow, int ncol, const int* row_ptr,
                       const int* cols, const double* x) {
  memset(y, 0, (size_t)ncol * sizeof(double));
  for (int r = 0; r < nrow; r++) {
    double xv = 0.0;
    for (int i = row_ptr[r]; i < row_ptr[r + 1]; i++) xv += x[cols[i]];
    for (int i = row_ptr[r]; i < row_ptr[r + 1]; i++) y[cols[i]] += xv;
  }
}

/* cbcsr.h:76-106 (cbcsr_A_mul_B), single-thread order: block-major, then row;
 * each cell's partial is summed first, then added to y[row]. */
void fso_cbcsr_A_mul_B(double* y, int nrow, int nblocks, const int* row_ptr,
                       const int* cols, const double* x) {
  memset(y, 0, (size_t)nrow * sizeof(double));
  for (int b = 0; b < nblocks; b++) {
    for (int r = 0; r < nrow; r++) {
      long cell = (long)b * nrow + r;
      double t = 0.0;
      for (int i = row_ptr[cell]; i < row_ptr[cell + 1]; i++) t += x[cols[i]];
      y[r] += t;
    }
  }
}

/* sparse.h:259-336 (bsbm_A_mul_B/_B2/_B4/_Bn), dsparse.h:176-191
 * (bsdm_A_mul_B): per block zero the block's Y rows, then scatter in stored
 * order.  Blocks own disjoint Y rows. */
void fso_blocked_A_mul_Bn(double* Y, const fso_blocked* B, const double* X, int R) {
#pragma omp parallel for schedule(dynamic, 1)
  for (int b = 0; b < B->nblocks; b++) {
    int r0 = B->start_row[b], r1 = B->start_row[b + 1];
    memset(Y + (size_t)r0 * R, 0, (size_t)(r1 - r0) * R * sizeof(double));
    long off = B->blk_off[b];
    for (long j = 0; j < B->blk_nnz[b]; j++) {
      double* yr = Y + (size_t)B->rows[off + j] * R;
      const double* xr = X + (size_t)B->cols[off + j] * R;
      if (B->vals) {
        double v = B->vals[off + j];
        for (int k = 0; k < R; k++) yr[k] += xr[k] * v;
      } else {
        for (int k = 0; k < R; k++) yr[k] += xr[k];
      }
    }
  }
}

/* ------------------------------------------------------------------ */
/* Dense helpers: linalg.h                                             */
/* ------------------------------------------------------------------ */
double fso_dist(const double* x, const double* y, int n) { /* linalg.h:6-13 */
  double d = 0.0;
  for (int i = 0; i < n; i++) { double t = x[i] - y[i]; d += t * t; }
  return sqrt(d);
}
double fso_normsq(const double* x, int n) { /* linalg.h:15-22 */
  double s = 0.0;
  for (int i = 0; i < n; i++) s += x[i] * x[i];
  return s;
}
double fso_dot(const double* x, const double* y, int n) { /* linalg.h:51-58 */
  double s = 0.0;
  for (int i = 0; i < n; i++) s += x[i] * y[i];
  return s;
}
void fso_normsq2(double* out2, const double* X, int n) { /* linalg.h:24-34 */
  double a = 0.0, b = 0.0;
  for (int i = 0; i < n; i++) { a += X[2 * i] * X[2 * i]; b += X[2 * i + 1] * X[2 * i + 1]; }
  out2[0] = a; out2[1] = b;
}
void fso_outer2(double* out3, const double* X, int n) { /* linalg.h:37-49 */
  fso_dot2sym(out3, X, X, n);
}
void fso_dot2sym(double* out3, const double* X, const double* Y, int n) { /* linalg.h:61-73 */
  double aa = 0.0, bb = 0.0, ab = 0.0;
  for (int i = 0; i < n; i++) {
    aa += X[2 * i] * Y[2 * i];
    bb += X[2 * i + 1] * Y[2 * i + 1];
    ab += X[2 * i] * Y[2 * i + 1];
  }
  out3[0] = aa; out3[1] = bb; out3[2] = ab;
}
/* linalg.h:77-88: A = [a0 a2; a2 a1] symmetric, RHS and X column-ordered 2x2,
 * closed-form inverse, no pivoting */
void fso_solve2sym(double* X, const double* A, const double* RHS) {
  double dinv = 1.0 / (A[0] * A[1] - A[2] * A[2]);
  double i00 = dinv * A[1], i11 = dinv * A[0], i01 = -dinv * A[2];
  X[0] = i00 * RHS[0] + i01 * RHS[1];
  X[1] = i01 * RHS[0] + i11 * RHS[1];
  X[2] = i00 * RHS[2] + i01 * RHS[3];
  X[3] = i01 * RHS[2] + i11 * RHS[3];
}

/* ------------------------------------------------------------------ */
/* Solver: cg.h                                                        */
/* ------------------------------------------------------------------ */

/* cg.h:9-22 (bsbm_AtA): y = At (A x) + lambda x, tmp has A->nrow entries */
void fso_blocked_AtA(double* y, const fso_blocked* A, const fso_blocked* At,
                     const double* x, double* tmp, double lambda) {
  fso_blocked_A_mul_Bn(tmp, A, x, 1);
  fso_blocked_A_mul_Bn(y, At, tmp, 1);
  for (int i = 0; i < At->nrow; i++) y[i] += lambda * x[i];
}

/* cg.h:25-82 (bsbm_cg): plain CG from x0 = 0; stop when ||r|| <= tol*||b||;
 * at most F iterations; returns the loop counter at exit (the value the
 * reference stores in *out_iter). */
int fso_blocked_cg(double* x, const fso_blocked* A, const fso_blocked* At,
                   const double* b, double lambda, double tol) {
  int F = A->ncol, N = A->nrow;
  double thr = tol * sqrt(fso_normsq(b, F));
  double* r = (double*)malloc((size_t)F * sizeof(double));
  double* p = (double*)malloc((size_t)F * sizeof(double));
  double* Kp = (double*)malloc((size_t)F * sizeof(double));
  double* tmp = (double*)malloc((size_t)N * sizeof(double));
  for (int i = 0; i < F; i++) { x[i] = 0.0; r[i] = b[i]; p[i] = b[i]; }
  double rs_old = fso_normsq(r, F);
  int it;
  for (it = 0; it < F; it++) {
    fso_blocked_AtA(Kp, A, At, p, tmp, lambda);
    double alpha = rs_old / fso_dot(Kp, p, F);
    for (int i = 0; i < F; i++) { x[i] += alpha * p[i]; r[i] -= alpha * Kp[i]; }
    double rs_new = fso_normsq(r, F);
    if (sqrt(rs_new) <= thr) break;
    double beta = rs_new / rs_old;
    for (int i = 0; i < F; i++) p[i] = r[i] + beta * p[i];
    rs_old = rs_new;
  }
  free(r); free(p); free(Kp); free(tmp);
  return it;
}

/* cg.h:85-187 (bsbm_cg2): block CG for 2 right-hand sides on the
 * column-normalised system; Alpha = (P'KP)^-1 R'R, Psi = (R'R)^-1 R'R_new,
 * both by the closed-form symmetric 2x2 solve; stop when both diagonal
 * entries of R'R are <= tol^2. */
int fso_blocked_cg2(double* X, const fso_blocked* A, const fso_blocked* At,
                    const double* B, double lambda, double tol) {
  int F = A->ncol, N = A->nrow;
  double tolsq = tol * tol;
  double nrm[2], inrm[2];
  fso_normsq2(nrm, B, F);
  nrm[0] = sqrt(nrm[0]); nrm[1] = sqrt(nrm[1]);
  inrm[0] = 1.0 / nrm[0]; inrm[1] = 1.0 / nrm[1];
  double* Rm = (double*)malloc((size_t)2 * F * sizeof(double));
  double* P = (double*)malloc((size_t)2 * F * sizeof(double));
  double* KP = (double*)malloc((size_t)2 * F * sizeof(double));
  double* tmp = (double*)malloc((size_t)2 * N * sizeof(double));
  for (int i = 0; i < F; i++) {
    for (int k = 0; k < 2; k++) {
      X[2 * i + k] = 0.0;
      Rm[2 * i + k] = B[2 * i + k] * inrm[k];
      P[2 * i + k] = Rm[2 * i + k];
    }
  }
  double RtR[3], RtR2[3], PtKP[3], Al[4], Ps[4];
  fso_outer2(RtR, Rm, F);
  int it;
  for (it = 0; it < F; it++) {
    fso_blocked_A_mul_Bn(tmp, A, P, 2);
    fso_blocked_A_mul_Bn(KP, At, tmp, 2);
    for (int i = 0; i < 2 * F; i++) KP[i] += lambda * P[i];
    fso_dot2sym(PtKP, P, KP, F);
    double rhs[4] = {RtR[0], RtR[2], RtR[2], RtR[1]};
    fso_solve2sym(Al, PtKP, rhs);
    for (int i = 0; i < F; i++) {
      double p0 = P[2 * i], p1 = P[2 * i + 1], k0 = KP[2 * i], k1 = KP[2 * i + 1];
      X[2 * i]      += Al[0] * p0 + Al[1] * p1;
      X[2 * i + 1]  += Al[2] * p0 + Al[3] * p1;
      Rm[2 * i]     -= Al[0] * k0 + Al[1] * k1;
      Rm[2 * i + 1] -= Al[2] * k0 + Al[3] * k1;
    }
    fso_outer2(RtR2, Rm, F);
    if (RtR2[0] <= tolsq && RtR2[1] <= tolsq) break;
    double rhs2[4] = {RtR2[0], RtR2[2], RtR2[2], RtR2[1]};
    fso_solve2sym(Ps, RtR, rhs2);
    for (int i = 0; i < F; i++) {
      double p0 = P[2 * i], p1 = P[2 * i + 1];
      P[2 * i]     = Rm[2 * i]     + Ps[0] * p0 + Ps[1] * p1;
      P[2 * i + 1] = Rm[2 * i + 1] + Ps[2] * p0 + Ps[3] * p1;
    }
    RtR[0] = RtR2[0]; RtR[1] = RtR2[1]; RtR[2] = RtR2[2];
  }
  for (int i = 0; i < F; i++) { X[2 * i] *= nrm[0]; X[2 * i + 1] *= nrm[1]; }
  free(Rm); free(P); free(KP); free(tmp);
  return it;
}

/* ------------------------------------------------------------------ */
/* File formats                                                        */
/* ------------------------------------------------------------------ */

/* sparse.h:112-139 (read_sbm) / dsparse.h:64-93 (read_sdm): three native
 * 8-byte longs (nrow, ncol, nnz), int32 rows[nnz], int32 cols[nnz]
 * (+ double vals[nnz]), indices 1-based on disk and 0-based in memory. */
int fso_read_coo_file(const char* path, long* nrow, long* ncol, long* nnz,
                      int* rows, int* cols, double* vals) {
  FILE* f = fopen(path, "rb");
  if (!f) return 1;
  int64_t hdr[3];
  if (fread(hdr, sizeof(int64_t), 3, f) != 3) { fclose(f); return 2; }
  *nrow = hdr[0]; *ncol = hdr[1]; *nnz = hdr[2];
  if (rows) {
    size_t n = (size_t)hdr[2];
    if (fread(rows, sizeof(int), n, f) != n || fread(cols, sizeof(int), n, f) != n) { fclose(f); return 3; }
    if (vals && fread(vals, sizeof(double), n, f) != n) { fclose(f); return 3; }
    for (size_t i = 0; i < n; i++) { rows[i]--; cols[i]--; }
  }
  fclose(f);
  return 0;
}

/* csr.h:83,97-113 (serialize_to_file).  The raw 32-byte struct image contains
 * whatever host pointers the writer held; readers ignore them. */
#define FSO_CSR_TAG "BINARY_CSR: struct BinaryCSR, int[nrow], int[nnz]\n"
struct fso_bcsr_image { int nrow; int ncol; long nnz; const int* row_ptr; const int* cols; };

int fso_write_csr_bin(const char* path, int nrow, int ncol, long nnz,
                      const int* row_ptr, const int* cols) {
  FILE* f = fopen(path, "wb");
  if (!f) return 1;
  struct fso_bcsr_image img = {nrow, ncol, nnz, row_ptr, cols};
  fputs(FSO_CSR_TAG, f);
  fputs("struct BinaryCSR\n", f);
  fwrite(&img, sizeof img, 1, f);
  fprintf(f, "int[%d]\n", nrow + 1);
  fwrite(row_ptr, sizeof(int), (size_t)nrow + 1, f);
  fprintf(f, "int[%ld]\n", nnz);
  fwrite(cols, sizeof(int), (size_t)nnz, f);
  fclose(f);
  return 0;
}

/* csr.h:117-146 (deserialize_from_file) */
int fso_read_csr_bin(const char* path, int* nrow, int* ncol, long* nnz,
                     int* row_ptr, int* cols) {
  FILE* f = fopen(path, "rb");
  if (!f) return 1;
  char line[256], want[64];
  struct fso_bcsr_image img;
  int rc = 2;
  if (!fgets(line, sizeof line, f) || strcmp(line, FSO_CSR_TAG)) goto done;
  if (!fgets(line, sizeof line, f) || strcmp(line, "struct BinaryCSR\n")) goto done;
  if (fread(&img, sizeof img, 1, f) != 1) goto done;
  *nrow = img.nrow; *ncol = img.ncol; *nnz = img.nnz;
  rc = 0;
  if (!row_ptr) goto done;
  rc = 3;
  snprintf(want, sizeof want, "int[%d]\n", img.nrow + 1);
  if (!fgets(line, sizeof line, f) || strcmp(line, want)) goto done;
  if (fread(row_ptr, sizeof(int), (size_t)img.nrow + 1, f) != (size_t)img.nrow + 1) goto done;
  snprintf(want, sizeof want, "int[%ld]\n", img.nnz);
  if (!fgets(line, sizeof line, f) || strcmp(line, want)) goto done;
  if (fread(cols, sizeof(int), (size_t)img.nnz, f) != (size_t)img.nnz) goto done;
  rc = 0;
done:
  fclose(f);
  return rc;
}

/* ------------------------------------------------------------------ synthetic inputs
 * The benchmark's input generator ("generated once on the host in C, shared by both paths"): splitmix64
 * finaliser over (seed, 3j + k), scaled to the index range by the high word of a 64 x 32-bit product. */
static uint64_t syn_mix(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
static uint32_t syn_scale(uint64_t h, uint32_t n) {
  const uint64_t hi = h >> 32, lo = h & 0xffffffffull;
  return (uint32_t)((hi * n + ((lo * n) >> 32)) >> 32);
}
static uint64_t syn_gcd(uint64_t x, uint64_t y) { while (y) { uint64_t t = x % y; x = y; y = t; } return x; }

void fso_synth_coo(unsigned long long seed, int dist, long j0, long n, int nrow, int ncol,
                   int* rows, int* cols, double* vals) {
  int nbits = 0;                                  /* octaves covering [1, ncol] */
  while ((1ll << nbits) <= (long long)ncol) ++nbits;
  uint64_t a = 2654435761ull % (uint64_t)(ncol > 0 ? ncol : 1);   /* affine column scatter, made coprime with ncol */
  if (a == 0) a = 1;
  while (ncol > 0 && syn_gcd(a, (uint64_t)ncol) != 1) ++a;
#pragma omp parallel for schedule(static)
  for (long i = 0; i < n; ++i) {
    const uint64_t j = (uint64_t)(j0 + i);
    if (rows) rows[i] = (int)syn_scale(syn_mix(seed ^ (3 * j)), (uint32_t)nrow);
    if (cols) {
      const uint64_t h = syn_mix(seed ^ (3 * j + 1));
      if (dist == 0) {
        cols[i] = (int)syn_scale(h, (uint32_t)ncol);
      } else {
        const uint32_t oct = syn_scale(h, (uint32_t)nbits);
        uint64_t rank = (1ull << oct) + syn_scale(syn_mix(h), (uint32_t)(1u << oct));
        rank = (rank - 1) % (uint64_t)ncol;
        cols[i] = (int)((rank * a + 12345ull) % (uint64_t)ncol);
      }
    }
    if (vals) vals[i] = (double)(syn_mix(seed ^ (3 * j + 2)) >> 11) * (1.0 / 9007199254740992.0);
  }
}
