// kernels_dense.cu -- dense building blocks of the block conjugate-gradient solver.
//
// Replaces the OpenMP reductions and vector loops of the reference:
//   pnormsq/pnormsq2/pouter2/pdot/pdot2sym linalg.h:15-73  -> gram_partial/gram_final
//   solve2sym linalg.h:77-88                               -> small_solve_kernel (R x R)
//   update loops cg.h:60-63,70-73,148-154,165-170,176-179  -> cg_* kernels below
// All operands are tall-skinny row-major [n][R], R <= 32: streaming, HBM-bound passes.
// Reductions are two-stage with a fixed order, so every rank of a multi-GPU solve
// gets bit-identical Gram matrices and takes the same branches.
#include <algorithm>

#include "fsb_dense.h"
#include "fsb_internal.h"

namespace {

constexpr int kGramCtas = 148 * 2;
constexpr int kTile = 64;  // rows staged per step

// partial[cta][R*R] = sum over this CTA's rows of Xa[i][a] * Xb[i][b].
// Register-tiled mini GEMM: thread (kgroup, ta, tb) accumulates a 4x4 block of the (padded)
// 32x32 result over every fourth row of the staged tile; the four row groups are combined
// through shared memory at the end.  Fixed order => bit-reproducible.
__global__ void __launch_bounds__(256)
gram_partial_kernel(double* __restrict__ partial, const double* __restrict__ Xa, const double* __restrict__ Xb,
                    long long n, int R) {
  __shared__ __align__(16) double sm[2 * kTile * 32];
  double (*sa)[32] = reinterpret_cast<double (*)[32]>(sm);
  double (*sb)[32] = reinterpret_cast<double (*)[32]>(sm + kTile * 32);
  const int tid = threadIdx.x;
  const int kg = tid >> 6, u = tid & 63;
  const int a0 = (u >> 3) * 4, b0 = (u & 7) * 4;
  double acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
  for (long long base = (long long)blockIdx.x * kTile; base < n; base += (long long)gridDim.x * kTile) {
    const int rows = (int)min((long long)kTile, n - base);
    for (int e = tid; e < kTile * 32; e += 256) {
      const int r = e >> 5, c = e & 31;
      const bool ok = r < rows && c < R;
      sa[r][c] = ok ? Xa[(base + r) * R + c] : 0.0;
      sb[r][c] = ok ? Xb[(base + r) * R + c] : 0.0;
    }
    __syncthreads();
    for (int r = kg; r < kTile; r += 4) {
      const double2 av0 = *reinterpret_cast<const double2*>(&sa[r][a0]), av1 = *reinterpret_cast<const double2*>(&sa[r][a0 + 2]);
      const double2 bv0 = *reinterpret_cast<const double2*>(&sb[r][b0]), bv1 = *reinterpret_cast<const double2*>(&sb[r][b0 + 2]);
      const double av[4] = {av0.x, av0.y, av1.x, av1.y};
      const double bv[4] = {bv0.x, bv0.y, bv1.x, bv1.y};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fma(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
  // combine the four row groups: sm viewed as [4][1024]
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) sm[kg * 1024 + (a0 + i) * 32 + (b0 + j)] = acc[i][j];
  __syncthreads();
  for (int e = tid; e < 1024; e += 256) {
    const int a = e >> 5, b = e & 31;
    if (a < R && b < R) partial[(size_t)blockIdx.x * R * R + a * R + b] = (sm[e] + sm[1024 + e]) + (sm[2048 + e] + sm[3072 + e]);
  }
}

__global__ void gram_final_kernel(double* __restrict__ G, const double* __restrict__ partial, int nparts, int RR) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= RR) return;
  double s = 0.0;
  for (int p = 0; p < nparts; ++p) s += partial[(size_t)p * RR + e];
  G[e] = s;
}

__global__ void axpy_lambda_kernel(double* __restrict__ Y, const double* __restrict__ X, double lambda, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) Y[i] = fma(lambda, X[i], Y[i]);
}

// X = 0, Rm = P = B * diag(inorm)     (cg.h:44-48, 112-120)
__global__ void cg_init_kernel(double* __restrict__ X, double* __restrict__ Rm, double* __restrict__ P,
                               const double* __restrict__ B, const double* __restrict__ inorm, long long n, int R) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    const double r = B[i] * inorm[i % R];
    X[i] = 0.0;
    Rm[i] = r;
    P[i] = r;
  }
}

// norms from the Gram diagonal: norm[k] = sqrt(G[k][k]) (or 1 when !normalise)
__global__ void cg_norms_kernel(double* __restrict__ norm, double* __restrict__ inorm, const double* __restrict__ G, int R, int normalise) {
  const int k = threadIdx.x;
  if (k >= R) return;
  const double nv = normalise ? sqrt(G[k * R + k]) : 1.0;
  norm[k] = nv;
  inorm[k] = 1.0 / nv;
}

// O[i,:] (op)= I[i,:] * M for a tall row-major [n][R] operand and an R x R matrix M
// (M[k][j] = coefficient of input column k in output column j):
//   MODE 0: O += I M      (X += P alpha,      cg.h:148-151)
//   MODE 1: O -= I M      (R -= KP alpha,     cg.h:152-153)
//   MODE 2: O  = Add + I M, I may alias O     (P = R + P psi, cg.h:165-170)
// 128 rows per tile staged in shared memory; each thread owns a 4 (rows) x 4 (cols) block.
constexpr int kMixRows = 128;
template <int MODE>
__global__ void __launch_bounds__(256)
cg_rowmix_kernel(double* O, const double* I, const double* __restrict__ Add,
                 const double* __restrict__ M, long long n, int R) {
  __shared__ __align__(16) double smat[32][32];
  __shared__ double sin[kMixRows][33];
  const int tid = threadIdx.x;
  for (int e = tid; e < 1024; e += 256) {
    const int k = e >> 5, j = e & 31;
    smat[k][j] = (k < R && j < R) ? M[k * R + j] : 0.0;
  }
  const int tj = tid & 7, ti = tid >> 3;
  const int j0 = tj * 4;
  for (long long base = (long long)blockIdx.x * kMixRows; base < n; base += (long long)gridDim.x * kMixRows) {
    const int rows = (int)min((long long)kMixRows, n - base);
    __syncthreads();
    for (int e = tid; e < rows * R; e += 256) {
      const int r = e / R, c = e - r * R;
      sin[r][c] = I[(base + r) * R + c];
    }
    __syncthreads();
    double acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
    for (int k = 0; k < R; ++k) {
      const double2 m0 = *reinterpret_cast<const double2*>(&smat[k][j0]), m1 = *reinterpret_cast<const double2*>(&smat[k][j0 + 2]);
      const double mv[4] = {m0.x, m0.y, m1.x, m1.y};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const double x = sin[ti + 32 * i][k];
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fma(x, mv[j], acc[i][j]);
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = ti + 32 * i;
      if (r >= rows) continue;
      const long long off = (base + r) * R;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = j0 + j;
        if (c >= R) continue;
        if (MODE == 0) O[off + c] += acc[i][j];
        else if (MODE == 1) O[off + c] -= acc[i][j];
        else O[off + c] = Add[off + c] + acc[i][j];
      }
    }
  }
}

__global__ void cg_scale_cols_kernel(double* __restrict__ X, const double* __restrict__ norm, long long n, int R) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) X[i] *= norm[i % R];
}

// Solve A M = RHS for M (all R x R, row-major; A symmetric positive definite) by
// Cholesky A = L L' in one CTA.  Generalises the closed-form 2x2 solve2sym
// (linalg.h:77-88).  status[0] is set to 1 when a pivot is not safely positive.
// Also (check != 0) evaluates the stopping rule of bsbm_cg2 (cg.h:158): status[1] = 1
// when every diagonal entry of RHS is <= thr.
__global__ void small_solve_kernel(double* __restrict__ M, const double* __restrict__ A, const double* __restrict__ RHS,
                                   int R, int* __restrict__ status, int check, double thr) {
  __shared__ double L[32][33];
  __shared__ double Bm[32][33];
  __shared__ int bad;
  const int tid = threadIdx.x;  // 32 x 32 threads: (i, j)
  const int i = tid >> 5, j = tid & 31;
  if (tid == 0) bad = 0;
  if (i < R && j < R) {
    L[i][j] = A[i * R + j];
    Bm[i][j] = RHS[i * R + j];
  }
  __syncthreads();
  double amax = 0.0;
  for (int d = 0; d < R; ++d) amax = fmax(amax, fabs(L[d][d]));
  __syncthreads();
  for (int k = 0; k < R; ++k) {
    if (tid == 0) {
      const double piv = L[k][k];
      if (!(piv > 1e-14 * amax)) { bad = 1; L[k][k] = 1.0; } else L[k][k] = sqrt(piv);
    }
    __syncthreads();
    if (j == k && i > k && i < R) L[i][k] /= L[k][k];
    __syncthreads();
    if (i > k && j > k && j <= i && i < R) L[i][j] -= L[i][k] * L[j][k];
    __syncthreads();
  }
  // forward substitution L Z = B (column j handled by the threads with i == 0)
  if (i == 0 && j < R) {
    for (int r = 0; r < R; ++r) {
      double s = Bm[r][j];
      for (int c = 0; c < r; ++c) s -= L[r][c] * Bm[c][j];
      Bm[r][j] = s / L[r][r];
    }
    for (int r = R - 1; r >= 0; --r) {
      double s = Bm[r][j];
      for (int c = r + 1; c < R; ++c) s -= L[c][r] * Bm[c][j];
      Bm[r][j] = s / L[r][r];
    }
  }
  __syncthreads();
  if (i < R && j < R) M[i * R + j] = Bm[i][j];
  if (tid == 0) {
    status[0] = bad;
    if (check) {
      int done = 1;
      for (int d = 0; d < R; ++d)
        if (!(RHS[d * R + d] <= thr)) done = 0;
      status[1] = done;
    }
  }
}

// status[1] = all diag(G) <= thr  (stopping rule alone)
__global__ void diag_check_kernel(const double* __restrict__ G, int R, double thr, int* __restrict__ status) {
  if (threadIdx.x == 0) {
    int done = 1;
    for (int d = 0; d < R; ++d)
      if (!(G[d * R + d] <= thr)) done = 0;
    status[1] = done;
  }
}

inline int grid_for(long long n) { return (int)std::min<long long>((n + 255) / 256, 148LL * 16); }

}  // namespace

int fsb_dense_gram_into(double* dG, double* dPartial, const double* dXa, const double* dXb, long n, int R, cudaStream_t st) {
  if (R < 1 || R > 32) return fsb_set_error(FSB_EINVAL, "gram: R must be 1..32 (got %d)", R);
  const int ctas = (int)std::max<long long>(1, std::min<long long>(kGramCtas, (n + kTile - 1) / kTile));
  gram_partial_kernel<<<ctas, 256, 0, st>>>(dPartial, dXa, dXb, n, R);
  FSB_KERNEL_CHECK();
  gram_final_kernel<<<(R * R + 255) / 256, 256, 0, st>>>(dG, dPartial, ctas, R * R);
  FSB_KERNEL_CHECK();
  return FSB_OK;
}

size_t fsb_dense_gram_scratch_bytes(int R) { return (size_t)kGramCtas * R * R * sizeof(double); }

int fsb_dense_axpy_lambda(double* dY, const double* dX, double lambda, long n, cudaStream_t st) {
  if (n <= 0) return FSB_OK;
  axpy_lambda_kernel<<<grid_for(n), 256, 0, st>>>(dY, dX, lambda, n);
  FSB_KERNEL_CHECK();
  return FSB_OK;
}

int fsb_dense_cg_norms(double* dNorm, double* dInorm, const double* dG, int R, int normalise, cudaStream_t st) {
  cg_norms_kernel<<<1, 32, 0, st>>>(dNorm, dInorm, dG, R, normalise);
  FSB_KERNEL_CHECK();
  return FSB_OK;
}

int fsb_dense_cg_init(double* dX, double* dRm, double* dP, const double* dB, const double* dInorm, long n, int R, cudaStream_t st) {
  const long long tot = (long long)n * R;
  if (tot <= 0) return FSB_OK;
  cg_init_kernel<<<grid_for(tot), 256, 0, st>>>(dX, dRm, dP, dB, dInorm, tot, R);
  FSB_KERNEL_CHECK();
  return FSB_OK;
}

int fsb_dense_cg_update_xr(double* dX, const double* dP, double* dRm, const double* dKP, const double* dAlpha, long n, int R, cudaStream_t st) {
  if (n <= 0) return FSB_OK;
  const int ctas = (int)std::min<long long>((n + kMixRows - 1) / kMixRows, 148LL * 4);
  cg_rowmix_kernel<0><<<ctas, 256, 0, st>>>(dX, dP, nullptr, dAlpha, n, R);
  FSB_KERNEL_CHECK();
  cg_rowmix_kernel<1><<<ctas, 256, 0, st>>>(dRm, dKP, nullptr, dAlpha, n, R);
  FSB_KERNEL_CHECK();
  return FSB_OK;
}

int fsb_dense_cg_update_p(double* dP, const double* dRm, const double* dPsi, long n, int R, cudaStream_t st) {
  if (n <= 0) return FSB_OK;
  const int ctas = (int)std::min<long long>((n + kMixRows - 1) / kMixRows, 148LL * 4);
  cg_rowmix_kernel<2><<<ctas, 256, 0, st>>>(dP, dP, dRm, dPsi, n, R);
  FSB_KERNEL_CHECK();
  return FSB_OK;
}

int fsb_dense_scale_cols(double* dX, const double* dNorm, long n, int R, cudaStream_t st) {
  const long long tot = (long long)n * R;
  if (tot <= 0) return FSB_OK;
  cg_scale_cols_kernel<<<grid_for(tot), 256, 0, st>>>(dX, dNorm, tot, R);
  FSB_KERNEL_CHECK();
  return FSB_OK;
}

int fsb_dense_small_solve(double* dM, const double* dA, const double* dRHS, int R, int* dStatus, int check, double thr, cudaStream_t st) {
  small_solve_kernel<<<1, 1024, 0, st>>>(dM, dA, dRHS, R, dStatus, check, thr);
  FSB_KERNEL_CHECK();
  return FSB_OK;
}

int fsb_dense_diag_check(const double* dG, int R, double thr, int* dStatus, cudaStream_t st) {
  diag_check_kernel<<<1, 32, 0, st>>>(dG, R, thr, dStatus);
  FSB_KERNEL_CHECK();
  return FSB_OK;
}

extern "C" int fsb_gram_dev(double* G_host, const double* dXa, const double* dXb, long n, int R, void* stream) {
  FSB_TRY(fsb_require_device());
  if (!G_host || !dXa || !dXb || n < 0) return fsb_set_error(FSB_EINVAL, "fsb_gram_dev: bad argument");
  cudaStream_t st = fsb_pick_stream(stream);
  double *dG = nullptr, *dPart = nullptr;
  FSB_CUDA(cudaMalloc(&dG, (size_t)R * R * 8));
  cudaError_t e = cudaMalloc(&dPart, fsb_dense_gram_scratch_bytes(R));
  int rc = e == cudaSuccess ? fsb_dense_gram_into(dG, dPart, dXa, dXb, n, R, st) : fsb_cuda_error(e, "cudaMalloc", __FILE__, __LINE__);
  if (rc == FSB_OK) {
    e = cudaMemcpyAsync(G_host, dG, (size_t)R * R * 8, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) rc = fsb_cuda_error(e, "gram readback", __FILE__, __LINE__);
  }
  cudaFree(dG); cudaFree(dPart);
  return rc;
}

namespace {
__global__ void diff_kernel(double* d, const double* x, const double* __restrict__ y, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) d[i] = x[i] - y[i];
}
}  // namespace

extern "C" int fsb_gram_host(double* G, const double* Xa, const double* Xb, long n, int R) {
  FSB_TRY(fsb_require_device());
  if (!G || !Xa || !Xb || n < 0 || R < 1 || R > 32) return fsb_set_error(FSB_EINVAL, "fsb_gram_host: bad argument");
  cudaStream_t st = fsb_default_stream();
  const size_t bytes = std::max<size_t>((size_t)n * R, 1) * 8;
  double *da = nullptr, *db = nullptr;
  FSB_CUDA(cudaMalloc(&da, bytes));
  int rc = FSB_OK;
  cudaError_t e = cudaMemcpyAsync(da, Xa, (size_t)n * R * 8, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess && Xb != Xa) {
    e = cudaMalloc(&db, bytes);
    if (e == cudaSuccess) e = cudaMemcpyAsync(db, Xb, (size_t)n * R * 8, cudaMemcpyHostToDevice, st);
  }
  if (e != cudaSuccess) rc = fsb_cuda_error(e, "fsb_gram_host staging", __FILE__, __LINE__);
  if (rc == FSB_OK) rc = fsb_gram_dev(G, da, db ? db : da, n, R, (void*)st);
  cudaFree(da); cudaFree(db);
  return rc;
}

extern "C" int fsb_dist_host(double* out, const double* x, const double* y, long n) {
  FSB_TRY(fsb_require_device());
  if (!out || !x || !y || n < 0) return fsb_set_error(FSB_EINVAL, "fsb_dist_host: bad argument");
  cudaStream_t st = fsb_default_stream();
  const size_t bytes = std::max<size_t>((size_t)n, 1) * 8;
  double *dx = nullptr, *dy = nullptr;
  FSB_CUDA(cudaMalloc(&dx, bytes));
  cudaError_t e = cudaMalloc(&dy, bytes);
  if (e == cudaSuccess) e = cudaMemcpyAsync(dx, x, (size_t)n * 8, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(dy, y, (size_t)n * 8, cudaMemcpyHostToDevice, st);
  int rc = e == cudaSuccess ? FSB_OK : fsb_cuda_error(e, "fsb_dist_host staging", __FILE__, __LINE__);
  if (rc == FSB_OK && n > 0) {
    diff_kernel<<<grid_for(n), 256, 0, st>>>(dx, dx, dy, n);
    fsb_count_launch();
  }
  double g = 0.0;
  if (rc == FSB_OK) rc = fsb_gram_dev(&g, dx, dx, n, 1, (void*)st);
  cudaFree(dx); cudaFree(dy);
  if (rc == FSB_OK) *out = sqrt(g);
  return rc;
}
