/* examples/sampler_loop.c -- the Macau-style caller loop of the reference (bench_a_mul_b.c:331-360)
 * with everything resident in HBM: the matrix is read straight from its file, and per sample the
 * noise, the right-hand side B = A'N + sqrt(lambda) E and the block-CG solve of (A'A + lambda I) X = B
 * all run on the GPU.  Plain C, links only libfastsparse_b200.so (no CUDA toolkit needed):
 *
 *   gcc -std=gnu99 -O2 -Iinclude examples/sampler_loop.c -o sampler_loop \
 *       -Llibfastsparse_b200/lib -lfastsparse_b200 -Wl,-rpath,$PWD/libfastsparse_b200/lib -lm
 *   ./sampler_loop tests/golden/data/sbm-100-50.data 8 5
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "fsb.h"

#define CHECK(call)                                                        \
  do {                                                                     \
    if ((call) != 0) {                                                     \
      fprintf(stderr, "%s failed: %s\n", #call, fsb_last_error());         \
      return 1;                                                            \
    }                                                                      \
  } while (0)

int main(int argc, char** argv) {
  if (argc < 2) {
    fprintf(stderr, "usage: %s <matrix.sbm | matrix.csr.bin> [R=8] [samples=3]\n", argv[0]);
    return 2;
  }
  const char* path = argv[1];
  const int R = argc > 2 ? atoi(argv[2]) : 8;
  const int samples = argc > 3 ? atoi(argv[3]) : 3;
  const double lambda = 15.0, tol = 1e-6;

  fsb_matrix_t A = NULL;
  if (fsb_csr_load_bin_file(&A, path, NULL) != 0)          /* .csr.bin written by preprocess / serialize_to_file ... */
    CHECK(fsb_csr_load_coo_file(&A, path, 0));              /* ... or the raw COO file read_sbm reads */
  int fmt, nrow, ncol, has_vals, nblocks;
  long nnz;
  CHECK(fsb_matrix_info(A, &fmt, &nrow, &ncol, &nnz, &has_vals, &nblocks));
  printf("matrix %d x %d, %ld entries, resident: %ld bytes\n", nrow, ncol, nnz, fsb_matrix_bytes(A));

  const size_t fr = (size_t)ncol * R;
  double* dB = (double*)fsb_device_malloc(fr * sizeof(double));
  double* dX = (double*)fsb_device_malloc(fr * sizeof(double));
  double* dK = (double*)fsb_device_malloc(fr * sizeof(double));
  double* B = (double*)malloc(fr * sizeof(double));
  double* K = (double*)malloc(fr * sizeof(double));
  if (!dB || !dX || !dK || !B || !K) { fprintf(stderr, "allocation failed: %s\n", fsb_last_error()); return 1; }

  for (int s = 1; s <= samples; ++s) {
    int iters = 0;
    CHECK(fsb_noise_rhs_dev(A, NULL, dB, R, lambda, (unsigned long long)s, NULL));
    CHECK(fsb_cg_dev(A, NULL, dX, dB, R, lambda, tol, 0, &iters, NULL));
    /* residual of the sample, checked on the host: || (A'A + lambda I) X - B || / || B || */
    CHECK(fsb_ata_dev(A, dK, dX, R, lambda, NULL, 0, NULL));
    CHECK(fsb_copy_to_host(K, dK, fr * sizeof(double)));
    CHECK(fsb_copy_to_host(B, dB, fr * sizeof(double)));
    double num = 0.0, den = 0.0;
    for (size_t i = 0; i < fr; ++i) { num += (K[i] - B[i]) * (K[i] - B[i]); den += B[i] * B[i]; }
    const double rel = sqrt(num / den);
    printf("sample %d: %d iterations, relative residual %.3e\n", s, iters, rel);
    if (!(rel < 1e-4)) { fprintf(stderr, "sample %d did not converge\n", s); return 1; }
  }
  fsb_device_free(dB); fsb_device_free(dX); fsb_device_free(dK);
  free(B); free(K);
  CHECK(fsb_matrix_free(A));
  printf("SAMPLER LOOP OK\n");
  return 0;
}
