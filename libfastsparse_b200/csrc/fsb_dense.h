// fsb_dense.h -- launchers of kernels_dense.cu (block-CG building blocks).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

// bytes of partial-Gram scratch any launch below may use
size_t fsb_dense_gram_scratch_bytes(int R);
// first stage of G = Xa' Xb: dPartial[nparts][R*R] (fixed grid => fixed order); the second stage is
// fsb_dense_gram_into's final kernel or the prologue of fsb_dense_small_solve
int fsb_dense_gram_partial(double* dPartial, const double* dXa, const double* dXb, long n, int R, cudaStream_t st, int* nparts);
int fsb_dense_gram_finalize(double* dG, const double* dPartial, int nparts, int R, cudaStream_t st);   // second stage alone
// dG[R*R] = Xa' Xb (row-major), deterministic two-stage reduction through dPartial
int fsb_dense_gram_into(double* dG, double* dPartial, const double* dXa, const double* dXb, long n, int R, cudaStream_t st);
int fsb_dense_cg_norms(double* dNorm, double* dInorm, const double* dG, int R, int normalise, cudaStream_t st);
int fsb_dense_cg_init(double* dX, double* dRm, double* dP, const double* dB, const double* dInorm, long n, int R, cudaStream_t st);
// row mixes with an R x R coefficient matrix; dStatus (nullable) = solver status words, a stopped solver skips the pass
int fsb_dense_mix_add(double* dO, const double* dI, const double* dM, long n, int R, const int* dStatus, cudaStream_t st);   // O += I M
int fsb_dense_mix_sub_gram(double* dO, const double* dI, const double* dM, double* dPartial, long n, int R, const int* dStatus,
                           cudaStream_t st, int* nparts);                                                                   // O -= I M; partial = O'O
int fsb_dense_mix_set(double* dO, const double* dI, const double* dAdd, const double* dM, long n, int R, const int* dStatus,
                      cudaStream_t st);                                                                                     // O = Add + I M (I may alias O)
// column halves of a [n][R] operand into two [n][R/2] arrays (a stopped solver skips the pass)
int fsb_dense_split_halves(double* dLo, double* dHi, const double* dSrc, long n, int R, const int* dStatus, cudaStream_t st);
int fsb_dense_scale_cols(double* dX, const double* dNorm, long n, int R, cudaStream_t st);
// M = A^-1 RHS; A / RHS are first summed from partial Grams when nA / nRHS > 0 (and written back)
int fsb_dense_small_solve(double* dM, double* dA, double* dRHS, const double* dPartA, int nA, const double* dPartRHS, int nRHS, int R,
                          int* dStatus, int check, double thr, cudaStream_t st);
int fsb_dense_diag_check(const double* dG, int R, double thr, int* dStatus, cudaStream_t st);
