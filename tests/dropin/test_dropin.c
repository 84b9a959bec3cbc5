/* test_dropin.c -- C acceptance test of the drop-in headers (include/fastsparse/).
 *
 * A plain C caller written against the reference's API only (struct names, function
 * names, argument order): it builds every format from the sbm/sdm fixtures, runs every
 * product and solver entry point and cross-checks them against each other the way the
 * reference's own test_sparse.c does (fast kernel == COO product == pinned values).
 * Run with CWD = tests/golden (it opens data/sbm-100-50.data).  Exit code 0 = pass. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "sparse.h"
#include "dsparse.h"
#include "linalg.h"
#include "cg.h"
#include "csr.h"
#include "cbcsr.h"

static int failures = 0;
#define CHECK(cond, ...)                         \
  do {                                           \
    if (!(cond)) {                               \
      printf("FAIL %s:%d: ", __FILE__, __LINE__); \
      printf(__VA_ARGS__);                       \
      printf("\n");                              \
      failures++;                                \
    }                                            \
  } while (0)

static double maxdiff(const double* a, const double* b, int n) {
  double m = 0;
  for (int i = 0; i < n; i++) m = fmax(m, fabs(a[i] - b[i]));
  return m;
}

int main(void) {
  struct SparseBinaryMatrix* A = read_sbm("data/sbm-100-50.data");
  CHECK(A->nrow == 100 && A->ncol == 50 && A->nnz == 504 && A->rows[0] == 8 && A->cols[0] == 0, "read_sbm");
  const int N = A->nrow, F = A->ncol;
  double* x = malloc(F * sizeof(double));
  double* y = malloc(N * sizeof(double));
  double* y2 = malloc(N * sizeof(double));
  for (int i = 0; i < F; i++) x[i] = sin(i * 19 + 0.4) + cos(i * i * 3);

  /* COO products, pinned values of test_sparse.c:54-55 */
  A_mul_B(y, A, x);
  CHECK(fabs(y[0] - 1.70095) < 1e-4 && fabs(y[99] + 0.174905) < 1e-4, "A_mul_B pinned values: %g %g", y[0], y[99]);

  /* binary CSR, all widths */
  struct BinaryCSR B;
  bcsr_from_sbm(&B, A);
  CHECK(B.row_ptr[1] == 5 && B.row_ptr[5] == 24 && B.cols[0] == 9 && B.cols[7] == 1, "new_bcsr structure");
  bcsr_A_mul_B(y2, &B, x);
  CHECK(maxdiff(y, y2, N) < 1e-12, "bcsr_A_mul_B vs A_mul_B: %g", maxdiff(y, y2, N));
  for (int R = 2; R <= 40; R += (R < 8 ? 1 : 8)) {
    double* X = malloc(F * R * sizeof(double));
    double* Y = malloc(N * R * sizeof(double));
    double* Yb = malloc(N * R * sizeof(double));
    for (int i = 0; i < F; i++)
      for (int k = 0; k < R; k++) X[i * R + k] = sin(7 * i + 17 * k + 0.3);
    bcsr_A_mul_Bn(Y, &B, X, R);
    double* xk = malloc(F * sizeof(double));
    double worst = 0;
    for (int k = 0; k < R; k++) {
      for (int i = 0; i < F; i++) xk[i] = X[i * R + k];
      A_mul_B(y2, A, xk);
      for (int r = 0; r < N; r++) worst = fmax(worst, fabs(y2[r] - Y[r * R + k]));
    }
    CHECK(worst < 1e-12, "bcsr_A_mul_Bn R=%d vs per-column A_mul_B: %g", R, worst);
    if (R <= 32) { bcsr_A_mul_B32n(Yb, &B, X, R); CHECK(maxdiff(Y, Yb, N * R) == 0, "B32n R=%d", R); }
    if (R == 2) { bcsr_A_mul_B2(Yb, &B, X); CHECK(maxdiff(Y, Yb, N * R) == 0, "B2"); }
    if (R == 4) { bcsr_A_mul_B4(Yb, &B, X); CHECK(maxdiff(Y, Yb, N * R) == 0, "B4"); }
    if (R == 8) { bcsr_A_mul_B8(Yb, &B, X); CHECK(maxdiff(Y, Yb, N * R) == 0, "B8"); bcsr_A_mul_B8_auto(Yb, &B, X); CHECK(maxdiff(Y, Yb, N * R) == 0, "B8_auto"); }
    /* blocked COO, before and after the Hilbert sort */
    struct BlockedSBM* Bl = new_bsbm(A, 8);
    CHECK(Bl->nblocks == 13 && Bl->start_row[1] == 8 && Bl->start_row[13] == 100, "new_bsbm metadata");
    bsbm_A_mul_Bn(Yb, Bl, X, R);
    CHECK(maxdiff(Y, Yb, N * R) < 1e-12, "bsbm_A_mul_Bn R=%d: %g", R, maxdiff(Y, Yb, N * R));
    sort_bsbm(Bl);
    bsbm_A_mul_Bn(Yb, Bl, X, R);
    CHECK(maxdiff(Y, Yb, N * R) < 1e-12, "bsbm_A_mul_Bn after sort_bsbm R=%d: %g", R, maxdiff(Y, Yb, N * R));
    /* column-blocked CSR */
    struct ColBinaryCSR Cb;
    cbcsr_from_sbm(&Cb, A, 8);
    cbcsr_A_mul_Bn(Yb, &Cb, X, R);
    CHECK(maxdiff(Y, Yb, N * R) < 1e-12, "cbcsr_A_mul_Bn R=%d: %g", R, maxdiff(Y, Yb, N * R));
    free(X); free(Y); free(Yb); free(xk);
  }

  /* A'A x three ways */
  double* z = malloc(F * sizeof(double));
  double* z2 = malloc(F * sizeof(double));
  A_mul_B(y, A, x);
  At_mul_B(z2, A, y);
  bcsr_AA_mul_B(z, &B, x);
  CHECK(maxdiff(z, z2, F) < 1e-10, "bcsr_AA_mul_B: %g", maxdiff(z, z2, F));
  CHECK(fabs(z[0] - 28.810541791856551) < 1e-10, "A'Ax pinned z[0] = %.15g", z[0]);
  parallel_bcsr_AA_mul_B(z, &B, x, NULL);
  CHECK(maxdiff(z, z2, F) < 1e-10, "parallel_bcsr_AA_mul_B: %g", maxdiff(z, z2, F));
  bcsr_At_mul_B(z, &B, y);
  CHECK(maxdiff(z, z2, F) < 1e-10, "bcsr_At_mul_B: %g", maxdiff(z, z2, F));

  /* .csr.bin round trip */
  serialize_to_file(&B, "/tmp/fsb_dropin_test.csr.bin");
  struct BinaryCSR B2;
  deserialize_from_file(&B2, "/tmp/fsb_dropin_test.csr.bin");
  CHECK(B2.nnz == B.nnz && B2.nrow == B.nrow && B2.ncol == B.ncol, "csr.bin header");
  CHECK(!memcmp(B.row_ptr, B2.row_ptr, (N + 1) * sizeof(int)) && !memcmp(B.cols, B2.cols, B.nnz * sizeof(int)), "csr.bin arrays");
  free_bcsr(&B2);

  /* double-valued formats */
  struct SparseDoubleMatrix* D = read_sdm("data/sdm-100-50.data");
  CHECK(D->nnz == 470 && D->rows[1] == 27 && fabs(D->vals[1] - 0.616153) < 1e-5, "read_sdm");
  struct CSR M;
  new_csr(&M, D->nnz, D->nrow, D->ncol, D->rows, D->cols, D->vals);
  sdm_A_mul_B(y, D, x);
  csr_A_mul_B(y2, &M, x);
  CHECK(maxdiff(y, y2, N) < 1e-12, "csr_A_mul_B vs sdm_A_mul_B: %g", maxdiff(y, y2, N));
  CHECK(fabs(y2[0] + 0.53031426988755659) < 1e-12 && fabs(y2[99] - 1.8142212560770206) < 1e-12, "csr_A_mul_B pinned: %.15g %.15g", y2[0], y2[99]);
  struct BlockedSDM* Bd = new_bsdm(D, 8);
  bsdm_A_mul_B(y2, Bd, x);
  CHECK(maxdiff(y, y2, N) < 1e-12, "bsdm_A_mul_B: %g", maxdiff(y, y2, N));
  sort_bsdm(Bd);
  bsdm_A_mul_B(y2, Bd, x);
  CHECK(maxdiff(y, y2, N) < 1e-12, "bsdm_A_mul_B sorted: %g", maxdiff(y, y2, N));
  sdm_At_mul_B(z, D, y);
  csr_At_mul_B(z2, &M, y);
  CHECK(maxdiff(z, z2, F) < 1e-12, "csr_At_mul_B vs sdm_At_mul_B: %g", maxdiff(z, z2, F));
  CHECK(fabs(z[0] - 6.2399987577409828) < 1e-11 && fabs(z[49] - 0.61167818156057363) < 1e-11, "sdm_At_mul_B pinned: %.15g %.15g", z[0], z[49]);

  /* solver: test_sparse.c:560-608 */
  sort_sbm(A);
  struct BlockedSBM* Ab = new_bsbm(A, 8);
  transpose(A);
  struct BlockedSBM* Atb = new_bsbm(A, 8);
  double* b = malloc(F * sizeof(double));
  double* xs = malloc(F * sizeof(double));
  for (int i = 0; i < F; i++) b[i] = sin(i * 19 + 0.4) + cos(i * i * 3);
  int iters = -1;
  bsbm_cg(xs, Ab, Atb, b, 5.0, 1e-6, &iters);
  CHECK(iters == 15, "bsbm_cg iterations %d", iters);
  CHECK(fabs(xs[0] - 0.0638578) < 1e-4 && fabs(xs[1] + 0.0302702) < 1e-4 && fabs(xs[49] + 0.0284737361861) < 1e-9, "bsbm_cg solution");
  double* tmp = malloc(N * sizeof(double));
  bsbm_AtA(z, Ab, Atb, xs, tmp, 5.0);
  CHECK(dist(z, b, F) < 1e-5, "CG residual %g", dist(z, b, F));
  double* b2 = malloc(2 * F * sizeof(double));
  double* x2 = malloc(2 * F * sizeof(double));
  for (int i = 0; i < F; i++) {
    b2[2 * i] = b[i];
    b2[2 * i + 1] = cos(i * 23 + 0.7) + sin(i * i * 7);
  }
  bsbm_cg2(x2, Ab, Atb, b2, 5.0, 1e-6, &iters);
  CHECK(iters == 13, "bsbm_cg2 iterations %d", iters);
  CHECK(fabs(x2[0] - 0.0638578) < 1e-4 && fabs(x2[2] + 0.0302702) < 1e-4 && fabs(x2[1] - 0.106690812806) < 1e-9, "bsbm_cg2 solution");
  double* b8 = malloc(8 * F * sizeof(double));
  double* x8 = malloc(8 * F * sizeof(double));
  for (int i = 0; i < F; i++)
    for (int k = 0; k < 8; k++) b8[i * 8 + k] = sin(3 * i + 5 * k + 0.1);
  bsbm_cgn(x8, Ab, Atb, b8, 8, 5.0, 1e-8, &iters);
  double worst = 0;
  for (int k = 0; k < 8; k++) {
    for (int i = 0; i < F; i++) xs[i] = x8[i * 8 + k];
    bsbm_AtA(z, Ab, Atb, xs, tmp, 5.0);
    for (int i = 0; i < F; i++) worst = fmax(worst, fabs(z[i] - b8[i * 8 + k]));
  }
  CHECK(iters > 0 && worst < 1e-6, "bsbm_cgn(8): %d iterations, residual %g", iters, worst);

  /* reductions */
  double xx[] = {0.12, -0.82, 1.3, 0.5}, yy[] = {6.12, 0.19, 3.4, -4.1};
  CHECK(fabs(pnormsq(xx, 4) - 2.6268) < 1e-8 && fabs(pdot(xx, yy, 4) - 2.9486) < 1e-8, "pnormsq/pdot");
  double o3[3];
  pouter2(o3, xx, 2);
  CHECK(fabs(o3[2] - (0.12 * -0.82 + 1.3 * 0.5)) < 1e-12, "pouter2");

  /* the host arrays stay the source of truth: an in-place edit between two calls must be seen (the reference reads
   * vals[] on every call).  Large enough that the sampled fingerprint (256 strided samples) cannot see the edit and
   * that the full-content hash runs on the background worker (> 4 MB of arrays). */
  {
    const int n = 600000, m = 64;
    long nz = (long)n;
    int* rr = malloc(nz * sizeof(int));
    int* cc = malloc(nz * sizeof(int));
    double* vv = malloc(nz * sizeof(double));
    for (long j = 0; j < nz; j++) { rr[j] = (int)j; cc[j] = (int)(j % m); vv[j] = 1.0 + (j % 7); }
    struct CSR Cs;
    new_csr(&Cs, nz, n, m, rr, cc, vv);
    double* xe = malloc(m * sizeof(double));
    double* ye = malloc(n * sizeof(double));
    for (int i = 0; i < m; i++) xe[i] = 1.0 + i;
    csr_A_mul_B(ye, &Cs, xe);
    const long k = 300001;                        /* not a multiple of nz/256: between two sampled positions */
    const double before = ye[k];
    Cs.vals[k] = 1000.0;
    csr_A_mul_B(ye, &Cs, xe);
    CHECK(fabs(before - (1.0 + (k % 7)) * xe[k % m]) < 1e-12, "product before the edit");
    CHECK(fabs(ye[k] - 1000.0 * xe[k % m]) < 1e-9, "in-place edit of vals[] not seen: y=%g (stale device copy)", ye[k]);
    Cs.cols[k] = (int)((k + 5) % m);
    csr_A_mul_B(ye, &Cs, xe);
    CHECK(fabs(ye[k] - 1000.0 * xe[(k + 5) % m]) < 1e-9, "in-place edit of cols[] not seen");
    free_csr(&Cs);
    free(rr); free(cc); free(vv); free(xe); free(ye);
  }

  printf(failures ? "DROPIN TEST FAILED (%d)\n" : "DROPIN TEST PASSED\n", failures);
  return failures != 0;
}
