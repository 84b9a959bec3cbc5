// fsb_dense.h -- launchers of kernels_dense.cu (block-CG building blocks).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

size_t fsb_dense_gram_scratch_bytes(int R);
// dG[R*R] = Xa' Xb (row-major), deterministic two-stage reduction through dPartial
int fsb_dense_gram_into(double* dG, double* dPartial, const double* dXa, const double* dXb, long n, int R, cudaStream_t st);
int fsb_dense_cg_norms(double* dNorm, double* dInorm, const double* dG, int R, int normalise, cudaStream_t st);
int fsb_dense_cg_init(double* dX, double* dRm, double* dP, const double* dB, const double* dInorm, long n, int R, cudaStream_t st);
int fsb_dense_cg_update_xr(double* dX, const double* dP, double* dRm, const double* dKP, const double* dAlpha, long n, int R, cudaStream_t st);
int fsb_dense_cg_update_p(double* dP, const double* dRm, const double* dPsi, long n, int R, cudaStream_t st);
int fsb_dense_scale_cols(double* dX, const double* dNorm, long n, int R, cudaStream_t st);
int fsb_dense_small_solve(double* dM, const double* dA, const double* dRHS, int R, int* dStatus, int check, double thr, cudaStream_t st);
int fsb_dense_diag_check(const double* dG, int R, double thr, int* dStatus, cudaStream_t st);
