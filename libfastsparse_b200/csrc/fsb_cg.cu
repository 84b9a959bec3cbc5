// fsb_cg.cu -- device-resident (block) conjugate gradient for (A'A + lambda I) X = B.
//
// Replaces bsbm_cg (cg.h:25-82, R = 1) and bsbm_cg2 (cg.h:85-187, R = 2), and
// generalises the latter to R <= 32 right-hand sides (the reference hard-codes the
// 2x2 solves; here alpha and psi come from an R x R Cholesky solve on the device).
// Every vector stays in HBM for the whole solve; the iteration is predicated on device-side
// status words, and the host reads them back once per batch of queued iterations.  With a row-sharded A and an active
// communicator the partial A'(A P) is sum-allreduced inside fsb_ata_*; the dense
// vectors are replicated and the Gram reductions are deterministic, so all ranks
// take identical branches without exchanging the flags.
#include <math.h>

#include <algorithm>

#include "fsb_dense.h"
#include "fsb_internal.h"

namespace {

struct CgWork {
  double *Rm = nullptr, *P = nullptr, *KP = nullptr, *tmp = nullptr;
  double *G1 = nullptr, *G2 = nullptr, *PtKP = nullptr, *Alpha = nullptr, *Psi = nullptr;
  double *norm = nullptr, *inorm = nullptr, *partial = nullptr, *partial2 = nullptr;
  int* status = nullptr;     // [0] breakdown, [1] converged, [2] completed iterations
  int* h_status = nullptr;
  void release() {
    cudaFree(Rm); cudaFree(P); cudaFree(KP); cudaFree(tmp); cudaFree(G1); cudaFree(G2); cudaFree(PtKP);
    cudaFree(Alpha); cudaFree(Psi); cudaFree(norm); cudaFree(inorm); cudaFree(partial); cudaFree(partial2); cudaFree(status);
    if (h_status) cudaFreeHost(h_status);
  }
};

constexpr int kStatusWords = 4;

int cg_alloc(CgWork& w, long F, long N, int R) {
  const size_t fr = std::max<size_t>((size_t)F * R, 1) * 8, nr = std::max<size_t>((size_t)N * R, 1) * 8, rr = (size_t)R * R * 8;
  FSB_CUDA(cudaMalloc(&w.Rm, fr)); FSB_CUDA(cudaMalloc(&w.P, fr)); FSB_CUDA(cudaMalloc(&w.KP, fr));
  FSB_CUDA(cudaMalloc(&w.tmp, nr));
  FSB_CUDA(cudaMalloc(&w.G1, rr)); FSB_CUDA(cudaMalloc(&w.G2, rr)); FSB_CUDA(cudaMalloc(&w.PtKP, rr));
  FSB_CUDA(cudaMalloc(&w.Alpha, rr)); FSB_CUDA(cudaMalloc(&w.Psi, rr));
  FSB_CUDA(cudaMalloc(&w.norm, R * 8)); FSB_CUDA(cudaMalloc(&w.inorm, R * 8));
  FSB_CUDA(cudaMalloc(&w.partial, fsb_dense_gram_scratch_bytes(R)));
  FSB_CUDA(cudaMalloc(&w.partial2, fsb_dense_gram_scratch_bytes(R)));
  FSB_CUDA(cudaMalloc(&w.status, kStatusWords * sizeof(int)));
  FSB_CUDA(cudaMallocHost(&w.h_status, kStatusWords * sizeof(int)));
  return FSB_OK;
}

// KP = A'(A P) + lambda P through whichever transposed operator the caller has
int apply_op(fsb_matrix* A, fsb_matrix* At, double* KP, const double* P, int R, double lambda, double* tmp, cudaStream_t st) {
  if (At) return fsb_ata_pair_dev(A, At, KP, P, R, lambda, tmp, (void*)st);
  return fsb_ata_dev(A, KP, P, R, lambda, tmp, 0, (void*)st);
}

// One iteration, enqueued without looking at the device: every kernel that changes solver state
// is predicated on status[] (a converged or broken-down solve is left untouched), so the host may
// queue several iterations per status read-back.  g1 = R'R of the current residual, g2 receives
// R'R of the next one.
int cg_enqueue_iteration(fsb_matrix* A, fsb_matrix* At, double* dX, int R, double lambda, double thr, cudaStream_t st,
                         CgWork& w, double* g1, double* g2) {
  const long F = A->ncol;
  int np = 0;
  FSB_TRY(apply_op(A, At, w.KP, w.P, R, lambda, w.tmp, st));
  FSB_TRY(fsb_dense_gram_partial(w.partial, w.P, w.KP, F, R, st, &np));                                  // P'KP (first stage)
  FSB_TRY(fsb_dense_small_solve(w.Alpha, w.PtKP, g1, w.partial, np, nullptr, 0, R, w.status, 0, 0.0, st)); // Alpha = PtKP^-1 RtR
  FSB_TRY(fsb_dense_mix_add(dX, w.P, w.Alpha, F, R, w.status, st));                                      // X += P Alpha
  FSB_TRY(fsb_dense_mix_sub_gram(w.Rm, w.KP, w.Alpha, w.partial2, F, R, w.status, st, &np));             // R -= KP Alpha, R'R
  FSB_TRY(fsb_dense_small_solve(w.Psi, g1, g2, nullptr, 0, w.partial2, np, R, w.status, 1, thr, st));    // Psi = RtR^-1 RtR2 (+ stop test)
  FSB_TRY(fsb_dense_mix_set(w.P, w.P, w.Rm, w.Psi, F, R, w.status, st));                                 // P = R + P Psi
  return FSB_OK;
}

int cg_run(fsb_matrix* A, fsb_matrix* At, double* dX, const double* dB, int R, double lambda, double tol,
           int max_iter, int* out_iter, cudaStream_t st, CgWork& w) {
  const long F = A->ncol;
  if (max_iter <= 0) max_iter = (int)F;
  FSB_CUDA(cudaMemsetAsync(w.status, 0, kStatusWords * sizeof(int), st));
  // norms of the right-hand sides; R == 1 keeps the unnormalised recurrence of bsbm_cg
  FSB_TRY(fsb_dense_gram_into(w.G1, w.partial, dB, dB, F, R, st));
  FSB_TRY(fsb_dense_cg_norms(w.norm, w.inorm, w.G1, R, R > 1, st));
  FSB_TRY(fsb_dense_cg_init(dX, w.Rm, w.P, dB, w.inorm, F, R, st));
  FSB_TRY(fsb_dense_gram_into(w.G1, w.partial, w.Rm, w.Rm, F, R, st));   // RtR
  double thr = tol * tol;
  if (R == 1) {  // stop when ||r|| <= tol * ||b||   (cg.h:40,67)
    double bb = 0.0;
    FSB_CUDA(cudaMemcpyAsync(&bb, w.G1, 8, cudaMemcpyDeviceToHost, st));
    FSB_CUDA(cudaStreamSynchronize(st));
    const double t = tol * sqrt(bb);
    thr = t * t;
  }
  // iterations queued per status read-back: large problems look after every iteration (an
  // iteration is milliseconds), small ones are launch-latency bound and queue several
  const double work = (double)A->nnz * R;
  const int batch = work >= 2e9 ? 1 : (work >= 2e8 ? 2 : 8);
  int queued = 0;
  w.h_status[0] = w.h_status[1] = w.h_status[2] = 0;
  while (queued < max_iter) {
    const int nb = std::min(batch, max_iter - queued);
    for (int k = 0; k < nb; ++k) {
      FSB_TRY(cg_enqueue_iteration(A, At, dX, R, lambda, thr, st, w, w.G1, w.G2));
      std::swap(w.G1, w.G2);
    }
    queued += nb;
    FSB_CUDA(cudaMemcpyAsync(w.h_status, w.status, kStatusWords * sizeof(int), cudaMemcpyDeviceToHost, st));
    FSB_CUDA(cudaStreamSynchronize(st));
    if (w.h_status[0] || w.h_status[1]) break;   // breakdown / converged (the reference breaks before updating P)
  }
  const int it = w.h_status[2];
  FSB_TRY(fsb_dense_scale_cols(dX, w.norm, F, R, st));
  FSB_CUDA(cudaStreamSynchronize(st));
  if (out_iter) *out_iter = it;
  if (w.h_status[0])
    return fsb_set_error(FSB_EBREAKDOWN, "block CG: Gram matrix lost rank at iteration %d (R=%d)", it, R);
  return FSB_OK;
}

}  // namespace

extern "C" int fsb_cg_dev(fsb_matrix_t A, fsb_matrix_t At, double* dX, const double* dB, int R, double lambda,
                          double tol, int max_iter, int* out_iter, void* stream) {
  FSB_TRY(fsb_require_device());
  if (!A || !dX || !dB) return fsb_set_error(FSB_EINVAL, "fsb_cg_dev: null argument");
  if (R < 1 || R > 32) return fsb_set_error(FSB_EINVAL, "fsb_cg_dev: R must be 1..32 (got %d)", R);
  if (At && (A->nrow != At->ncol || A->ncol != At->nrow))
    return fsb_set_error(FSB_EINVAL, "A (%d x %d) and At (%d x %d) must be transposes of each other.", A->nrow, A->ncol, At->nrow, At->ncol);
  if (!At && A->format != FSB_FMT_CSR) return fsb_set_error(FSB_EINVAL, "fsb_cg_dev: a stored transpose is required for non-CSR formats");
  cudaStream_t st = fsb_pick_stream(stream);
  CgWork w;
  int rc = cg_alloc(w, A->ncol, A->nrow, R);
  if (rc == FSB_OK) rc = cg_run(A, At, dX, dB, R, lambda, tol, max_iter, out_iter, st, w);
  if (rc == FSB_EBREAKDOWN && R > 1) {
    // rank loss (a column converged early / dependent right-hand sides): the block
    // recurrence cannot continue; solve the columns one by one instead
    double *xb = nullptr, *bb = nullptr;
    const long F = A->ncol;
    int worst = 0;
    cudaError_t e = cudaMalloc(&xb, std::max<size_t>(F, 1) * 8);
    if (e == cudaSuccess) e = cudaMalloc(&bb, std::max<size_t>(F, 1) * 8);
    rc = e == cudaSuccess ? FSB_OK : fsb_cuda_error(e, "cudaMalloc", __FILE__, __LINE__);
    CgWork w1;
    if (rc == FSB_OK) rc = cg_alloc(w1, A->ncol, A->nrow, 1);
    for (int k = 0; rc == FSB_OK && k < R; ++k) {
      e = cudaMemcpy2DAsync(bb, 8, dB + k, (size_t)R * 8, 8, F, cudaMemcpyDeviceToDevice, st);
      if (e != cudaSuccess) { rc = fsb_cuda_error(e, "column gather", __FILE__, __LINE__); break; }
      int it1 = 0;
      rc = cg_run(A, At, xb, bb, 1, lambda, tol, max_iter, &it1, st, w1);
      worst = std::max(worst, it1);
      if (rc == FSB_OK) {
        e = cudaMemcpy2DAsync(dX + k, (size_t)R * 8, xb, 8, 8, F, cudaMemcpyDeviceToDevice, st);
        if (e != cudaSuccess) rc = fsb_cuda_error(e, "column scatter", __FILE__, __LINE__);
      }
    }
    if (rc == FSB_OK) {
      e = cudaStreamSynchronize(st);
      if (e != cudaSuccess) rc = fsb_cuda_error(e, "sync", __FILE__, __LINE__);
    }
    if (out_iter) *out_iter = worst;
    w1.release();
    cudaFree(xb); cudaFree(bb);
  }
  w.release();
  return rc;
}

extern "C" int fsb_cg_host(fsb_matrix_t A, fsb_matrix_t At, double* X, const double* B, int R, double lambda,
                           double tol, int max_iter, int* out_iter) {
  FSB_TRY(fsb_require_device());
  if (!A || !X || !B) return fsb_set_error(FSB_EINVAL, "fsb_cg_host: null argument");
  const size_t bytes = std::max<size_t>((size_t)A->ncol * std::max(R, 1), 1) * 8;
  double *dX = nullptr, *dB = nullptr;
  FSB_CUDA(cudaMalloc(&dX, bytes));
  cudaError_t e = cudaMalloc(&dB, bytes);
  cudaStream_t st = fsb_default_stream();
  int rc = e == cudaSuccess ? FSB_OK : fsb_cuda_error(e, "cudaMalloc", __FILE__, __LINE__);
  if (rc == FSB_OK) {
    e = cudaMemcpyAsync(dB, B, (size_t)A->ncol * R * 8, cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) rc = fsb_cuda_error(e, "H2D", __FILE__, __LINE__);
  }
  if (rc == FSB_OK) rc = fsb_cg_dev(A, At, dX, dB, R, lambda, tol, max_iter, out_iter, (void*)st);
  if (rc == FSB_OK) {
    e = cudaMemcpyAsync(X, dX, (size_t)A->ncol * R * 8, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) rc = fsb_cuda_error(e, "D2H", __FILE__, __LINE__);
  }
  cudaFree(dX); cudaFree(dB);
  return rc;
}
