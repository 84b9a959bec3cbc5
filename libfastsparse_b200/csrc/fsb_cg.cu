// fsb_cg.cu -- device-resident (block) conjugate gradient for (A'A + lambda I) X = B.
//
// Replaces bsbm_cg (cg.h:25-82, R = 1) and bsbm_cg2 (cg.h:85-187, R = 2), and
// generalises the latter to R <= 32 right-hand sides (the reference hard-codes the
// 2x2 solves; here alpha and psi come from an R x R Cholesky solve on the device).
// Every vector stays in HBM for the whole solve; the iteration is predicated on device-side
// status words, and the host reads them back once per batch of queued iterations.
// Two multi-GPU forms (row-sharded A, active communicator):
//   * cg_run on replicated vectors: the partial A'(A P) is sum-allreduced inside fsb_ata_*;
//   * cg_run_sharded (default): vectors sharded over the unknowns, reduce-scatter of the partial
//     overlapped with its computation, all-gather of P, allreduce of the R x R Gram matrices.
// The Gram reductions are deterministic and allreduced to identical bits, so all ranks take
// identical branches without exchanging the flags.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <chrono>

#include "fsb_dense.h"
#include "fsb_internal.h"

namespace {

// FSB_CG_TRACE=1: wall-clock milliseconds of the solve's phases on stderr (workspace, every status
// read-back); =2 adds the device time of the phases of a sharded iteration -- a debugging aid for
// gaps the kernel timings do not show
int cg_trace_level() {
  static const int lvl = [] { const char* e = getenv("FSB_CG_TRACE"); return e && *e ? atoi(e) : 0; }();
  return lvl;
}
bool cg_trace() { return cg_trace_level() > 0; }
double now_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

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
  double t_prev = 0.0;
  if (cg_trace()) { cudaStreamSynchronize(st); t_prev = now_ms(); }
  while (queued < max_iter) {
    const int nb = std::min(batch, max_iter - queued);
    for (int k = 0; k < nb; ++k) {
      FSB_TRY(cg_enqueue_iteration(A, At, dX, R, lambda, thr, st, w, w.G1, w.G2));
      std::swap(w.G1, w.G2);
    }
    queued += nb;
    FSB_CUDA(cudaMemcpyAsync(w.h_status, w.status, kStatusWords * sizeof(int), cudaMemcpyDeviceToHost, st));
    FSB_CUDA(cudaStreamSynchronize(st));
    if (cg_trace()) {
      const double t = now_ms();
      fprintf(stderr, "[fsb cg] iterations %d..%d: %.3f ms (status %d %d %d)\n", queued - nb, queued - 1, t - t_prev, w.h_status[0], w.h_status[1], w.h_status[2]);
      t_prev = t;
    }
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

// ------------------------------------------------------------------------------------------
// Multi-GPU solve: A is row-sharded (rank g holds rows of A and the transpose of its shard), the
// CG vectors are sharded over the F unknowns.  Per iteration:
//   tmp_g  = A_g P                    P replicated ([F][R], all-gathered at the end of the previous iteration)
//   part_g = A_g' tmp_g               full-length partial, produced in C row chunks;
//   KP_loc = reduce-scatter(part)     chunk c is reduce-scattered on a second stream while chunk c+1 is
//                                     still being computed (the transfer hides behind the product)
//   Grams  = allreduce of R x R       identical bits on every rank => identical branches everywhere
//   X, R, P updates on the local 1/G of the rows; all-gather of the new P.
// The F rows are cut into C chunks of Fc rows, every chunk into G slices of s rows; rank g owns slice g
// of every chunk (local layout [C][s][R]).  The dense passes are row-local, so the layout is invisible
// to them.  Rows are padded with zeros to C*G*s.
constexpr int kMaxChunks = 4;
// 0 (and 2) = sharded vectors, 1 = replicated vectors + one allreduce,
// 3 = sharded vectors with P all-gathered as two column halves behind the next product's first pass
thread_local int g_cg_dist_mode = 0;

struct CgShardWork {
  int G = 1, rank = 0, C = 1;
  long F = 0, Fc = 0, s = 0, Fp = 0, nloc = 0;
  double *Pfull = nullptr, *KPpart = nullptr, *Xl = nullptr, *Rl = nullptr, *Pl = nullptr, *KPl = nullptr, *tmp = nullptr;
  double* Psend = nullptr;          // column halves of P_loc, the send buffers of the split all-gather
  fsb_p2p* p2p = nullptr;           // Pfull lives in a peer-mapped buffer: P is all-gathered by direct NVLink stores
  fsb_p2p* p2p_kp = nullptr;        // KPpart in a peer-mapped buffer: the reduce-scatter is a pull + ordered sum (knob cg_p2p_rs)
  // one iteration captured as a CUDA graph, one executable per parity of the G1 / G2 swap; valid for (gthr, glambda, gT)
  cudaGraphExec_t gexec[2] = {nullptr, nullptr};
  double gthr = -1.0, glambda = 0.0;
  const void* gT = nullptr;
  const double* G1_first = nullptr; // the buffer that was G1 when the workspace was made: parity 0 <=> G1 == G1_first
  void drop_graphs() { for (auto& g : gexec) { if (g) cudaGraphExecDestroy(g); g = nullptr; } }
  double *G1 = nullptr, *G2 = nullptr, *PtKP = nullptr, *Alpha = nullptr, *Psi = nullptr, *norm = nullptr, *inorm = nullptr;
  double *partial = nullptr;
  int *status = nullptr, *h_status = nullptr;
  cudaStream_t comm_st = nullptr;
  cudaEvent_t ev[kMaxChunks] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t ev_done = nullptr, ev_p = nullptr, ev_lo = nullptr, ev_hi = nullptr;
  void release() {
    cudaFree(Psend);
    drop_graphs();
    for (cudaEvent_t e : {ev_p, ev_lo, ev_hi}) if (e) cudaEventDestroy(e);
    if (p2p) { fsb_p2p_destroy(p2p); p2p = nullptr; Pfull = nullptr; }
    if (p2p_kp) { fsb_p2p_destroy(p2p_kp); p2p_kp = nullptr; KPpart = nullptr; }
    cudaFree(Pfull); cudaFree(KPpart); cudaFree(Xl); cudaFree(Rl); cudaFree(Pl); cudaFree(KPl); cudaFree(tmp);
    cudaFree(G1); cudaFree(G2); cudaFree(PtKP); cudaFree(Alpha); cudaFree(Psi); cudaFree(norm); cudaFree(inorm); cudaFree(partial);
    cudaFree(status);
    if (h_status) cudaFreeHost(h_status);
    for (auto& e : ev) if (e) cudaEventDestroy(e);
    if (ev_done) cudaEventDestroy(ev_done);
    if (comm_st) cudaStreamDestroy(comm_st);
  }
};

// chunk / slice geometry of the F-sharded vectors on G ranks (also exported for the CPU tests of the exchange pattern)
void shard_layout(long F, int R, int G, int* C, long* s, long* Fc, long* Fp, long* nloc) {
  // chunks only pay off when a chunk is a sizeable transfer; R = 1 keeps one chunk (its products may
  // take the merge-path kernel, which caches per-handle state and must see the whole matrix)
  // (four chunks: eight shorten the exposed tail -- the last chunk's reduce-scatter -- from 0.17 to 0.14 ms
  // on 8 GPUs at C5 but the smaller product launches lose 0.2 ms, profiles/r1i_cg_trace_n8_8chunks.log)
  (void)G;
  *C = (R >= 2 && (double)F * R * 8.0 >= 32e6) ? kMaxChunks : 1;
  *s = (F + (long)*C * G - 1) / ((long)*C * G);
  if ((*s * R) % 2) *s += 1;            // keep every slice 16-byte aligned for the vector paths
  *Fc = *s * G; *Fp = *Fc * *C; *nloc = *s * *C;
}

int shard_alloc(CgShardWork& w, long F, long Nloc, int R) {
  w.G = fsb_comm_size(); w.rank = fsb_comm_rank(); w.F = F;
  shard_layout(F, R, w.G, &w.C, &w.s, &w.Fc, &w.Fp, &w.nloc);
  const size_t full = (size_t)w.Fp * R * 8, loc = (size_t)w.nloc * R * 8, rr = (size_t)R * R * 8;
  // Pfull in peer-mapped memory when CUDA IPC works between the ranks (knob "cg_p2p", default on); plain memory + NCCL otherwise
  if (fsb_knob("cg_p2p", FSB_MULTI_GPU_DEFAULTS) && w.G > 1 && fsb_p2p_create(&w.p2p, full, fsb_default_stream()) == FSB_OK) w.Pfull = (double*)fsb_p2p_local(w.p2p);
  else { w.p2p = nullptr; FSB_CUDA(cudaMalloc(&w.Pfull, full)); }
  if (w.p2p && w.C <= 8 && fsb_knob("cg_p2p_rs", 0) && fsb_p2p_create(&w.p2p_kp, full, fsb_default_stream()) == FSB_OK)
    w.KPpart = (double*)fsb_p2p_local(w.p2p_kp);
  else { w.p2p_kp = nullptr; FSB_CUDA(cudaMalloc(&w.KPpart, full)); }
  FSB_CUDA(cudaMalloc(&w.Xl, loc)); FSB_CUDA(cudaMalloc(&w.Rl, loc)); FSB_CUDA(cudaMalloc(&w.Pl, loc)); FSB_CUDA(cudaMalloc(&w.KPl, loc));
  FSB_CUDA(cudaMalloc(&w.Psend, loc));
  FSB_CUDA(cudaMalloc(&w.tmp, std::max<size_t>((size_t)Nloc * R, 1) * 8));
  FSB_CUDA(cudaMalloc(&w.G1, rr)); FSB_CUDA(cudaMalloc(&w.G2, rr)); FSB_CUDA(cudaMalloc(&w.PtKP, rr));
  w.G1_first = w.G1;
  FSB_CUDA(cudaMalloc(&w.Alpha, rr)); FSB_CUDA(cudaMalloc(&w.Psi, rr));
  FSB_CUDA(cudaMalloc(&w.norm, R * 8)); FSB_CUDA(cudaMalloc(&w.inorm, R * 8));
  FSB_CUDA(cudaMalloc(&w.partial, fsb_dense_gram_scratch_bytes(R)));
  FSB_CUDA(cudaMalloc(&w.status, kStatusWords * sizeof(int)));
  FSB_CUDA(cudaMallocHost(&w.h_status, kStatusWords * sizeof(int)));
  FSB_CUDA(cudaStreamCreateWithFlags(&w.comm_st, cudaStreamNonBlocking));
  for (auto& e : w.ev) FSB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  FSB_CUDA(cudaEventCreateWithFlags(&w.ev_done, cudaEventDisableTiming));
  FSB_CUDA(cudaEventCreateWithFlags(&w.ev_p, cudaEventDisableTiming));
  FSB_CUDA(cudaEventCreateWithFlags(&w.ev_lo, cudaEventDisableTiming));
  FSB_CUDA(cudaEventCreateWithFlags(&w.ev_hi, cudaEventDisableTiming));
  return FSB_OK;
}

// the CSR face of a handle (blocked / column-blocked formats: their cached row-stable CSR view)
int csr_face(fsb_matrix* M, cudaStream_t st, fsb_matrix** out) {
  if (M->format == FSB_FMT_CSR) { *out = M; return FSB_OK; }
  FSB_TRY(fsb_build_csr_view(M, st));
  *out = M->view;
  return FSB_OK;
}

// sum-allreduce of an R x R matrix: one peer-memory kernel when the ranks share mapped buffers, NCCL otherwise
int shard_allreduce_small(CgShardWork& w, double* G, int n, cudaStream_t st) {
  if (w.p2p && n <= 1024 && fsb_knob("cg_p2p_gram", 1)) return fsb_p2p_allreduce_small(w.p2p, G, n, st);
  return fsb_allreduce_sum_dev(G, (long)n, (void*)st);
}

// G (R x R) = sum over ranks of Xa_loc' Xb_loc
int shard_gram(CgShardWork& w, double* G, const double* Xa, const double* Xb, int R, cudaStream_t st) {
  FSB_TRY(fsb_dense_gram_into(G, w.partial, Xa, Xb, w.nloc, R, st));
  return shard_allreduce_small(w, G, R * R, st);
}

// all-gather the local slices of every chunk of a sharded vector into the replicated layout
int shard_allgather(CgShardWork& w, double* full, const double* loc, int R, cudaStream_t st) {
  if (w.p2p && full == w.Pfull) return fsb_p2p_allgather_chunks(w.p2p, loc, w.C, w.s * R, st);   // direct peer stores
  FSB_TRY(fsb_comm_group_start());   // the C per-chunk gathers go out as one fused NCCL launch
  int rc = FSB_OK;
  for (int c = 0; c < w.C && rc == FSB_OK; ++c)
    rc = fsb_comm_allgather(loc + (size_t)c * w.s * R, full + (size_t)c * w.Fc * R, (size_t)w.s * R, st);
  const int rc2 = fsb_comm_group_end();
  return rc != FSB_OK ? rc : rc2;
}

// all-gather of the new P as two column halves on the second stream, so that the first column pass of
// the next A_g P starts as soon as the first half has arrived (the second half travels behind it)
int shard_allgather_halves(CgShardWork& w, int R, cudaStream_t st) {
  const int h = R / 2;
  double *lo = w.Psend, *hi = w.Psend + (size_t)w.nloc * h;
  FSB_TRY(fsb_dense_split_halves(lo, hi, w.Pl, w.nloc, R, w.status, st));
  FSB_CUDA(cudaEventRecord(w.ev_p, st));
  FSB_CUDA(cudaStreamWaitEvent(w.comm_st, w.ev_p, 0));
  for (int half = 0; half < 2; ++half) {
    const double* send = half ? hi : lo;
    double* recv = w.Pfull + (half ? (size_t)w.Fp * h : 0);
    FSB_TRY(fsb_comm_group_start());
    int rc = FSB_OK;
    for (int c = 0; c < w.C && rc == FSB_OK; ++c)
      rc = fsb_comm_allgather(send + (size_t)c * w.s * h, recv + (size_t)c * w.Fc * h, (size_t)w.s * h, w.comm_st);
    const int rc2 = fsb_comm_group_end();
    if (rc != FSB_OK || rc2 != FSB_OK) return rc != FSB_OK ? rc : rc2;
    FSB_CUDA(cudaEventRecord(half ? w.ev_hi : w.ev_lo, w.comm_st));
  }
  return FSB_OK;
}

// KP_loc = slice of sum_g A_g'(A_g P) + lambda P_loc
int shard_apply_op(fsb_matrix* A, fsb_matrix* Acsr, fsb_matrix* T, CgShardWork& w, int R, double lambda, cudaStream_t st, bool halves,
                   cudaEvent_t* trace = nullptr) {
  if (halves) {   // P arrived as column halves (shard_allgather_halves): one column pass per half
    FSB_TRY(fsb_launch_csr_spmm_halves(Acsr, w.tmp, w.Pfull, w.Pfull + (size_t)w.Fp * (R / 2), R, st, w.ev_lo, w.ev_hi));
  } else {
    FSB_TRY(fsb_spmm_dev(A, w.tmp, w.Pfull, R, (void*)st));
  }
  if (trace) cudaEventRecord(trace[0], st);
  for (int c = 0; c < w.C; ++c) {
    const long r0 = c * w.Fc, r1 = std::min(w.F, r0 + w.Fc);
    if (r1 > r0) {
      if (w.C == 1) {
        FSB_TRY(fsb_launch_csr_spmm(T, w.KPpart, w.tmp, R, st));
      } else {
        fsb_matrix part;     // rows [r0, r1) of A_g': row_ptr values stay absolute, so cols / vals are shared
        fsb_make_row_alias(&part, T, (int)r0, (int)r1);
        FSB_TRY(fsb_launch_csr_spmm(&part, w.KPpart + (size_t)r0 * R, w.tmp, R, st));
      }
    }
    if (w.p2p_kp) FSB_TRY(fsb_p2p_signal(w.p2p_kp, c, st));      // chunk c of my partial is complete: tell every rank
    FSB_CUDA(cudaEventRecord(w.ev[c], st));
    FSB_CUDA(cudaStreamWaitEvent(w.comm_st, w.ev[c], 0));
    if (w.p2p_kp) {   // pull the G partial slices of my slice of chunk c over NVLink, add in rank order, + lambda P fused
      FSB_TRY(fsb_p2p_pull_sum(w.p2p_kp, c, w.KPl + (size_t)c * w.s * R, (size_t)c * w.Fc * R + (size_t)w.rank * w.s * R, w.s * R,
                               lambda != 0.0 ? w.Pl + (size_t)c * w.s * R : nullptr, lambda, w.comm_st));
    } else {
      FSB_TRY(fsb_comm_reduce_scatter_sum(w.KPpart + (size_t)c * w.Fc * R, w.KPl + (size_t)c * w.s * R, (size_t)w.s * R, w.comm_st));
    }
  }
  if (trace) cudaEventRecord(trace[1], st);
  FSB_CUDA(cudaEventRecord(w.ev_done, w.comm_st));
  FSB_CUDA(cudaStreamWaitEvent(st, w.ev_done, 0));
  if (lambda != 0.0 && !w.p2p_kp) FSB_TRY(fsb_dense_axpy_lambda(w.KPl, w.Pl, lambda, w.nloc * R, st));
  return FSB_OK;
}

int cg_run_sharded(fsb_matrix* A, fsb_matrix* At, double* dX, const double* dB, int R, double lambda, double tol,
                   int max_iter, int* out_iter, cudaStream_t st, CgShardWork& w) {
  const long F = A->ncol;
  if (max_iter <= 0) max_iter = (int)F;
  fsb_matrix *T = nullptr, *Acsr = nullptr;
  FSB_TRY(csr_face(A, st, &Acsr));
  if (At) {
    FSB_TRY(csr_face(At, st, &T));
  } else {
    FSB_TRY(fsb_build_transpose(A, st));
    T = A->T;
  }
  bool halves = false;   // P currently stored as two column halves in Pfull
  const size_t full = (size_t)w.Fp * R * 8, loc = (size_t)w.nloc * R * 8;
  FSB_CUDA(cudaMemsetAsync(w.status, 0, kStatusWords * sizeof(int), st));
  FSB_CUDA(cudaMemsetAsync(w.KPpart, 0, full, st));     // the padding rows stay zero for the whole solve
  FSB_CUDA(cudaMemsetAsync(w.Xl, 0, loc, st)); FSB_CUDA(cudaMemsetAsync(w.Rl, 0, loc, st)); FSB_CUDA(cudaMemsetAsync(w.Pl, 0, loc, st));
  // column norms from the replicated right-hand side (every rank computes the same bits)
  FSB_TRY(fsb_dense_gram_into(w.G1, w.partial, dB, dB, F, R, st));
  FSB_TRY(fsb_dense_cg_norms(w.norm, w.inorm, w.G1, R, R > 1, st));
  for (int c = 0; c < w.C; ++c) {   // X = 0, R = P = B diag(inorm) on the local slices
    const long g0 = c * w.Fc + w.rank * w.s;
    const long valid = std::max(0L, std::min(w.s, F - g0));
    const size_t lo = (size_t)c * w.s * R;
    if (valid > 0) FSB_TRY(fsb_dense_cg_init(w.Xl + lo, w.Rl + lo, w.Pl + lo, dB + (size_t)g0 * R, w.inorm, valid, R, st));
  }
  FSB_TRY(shard_allgather(w, w.Pfull, w.Pl, R, st));
  FSB_TRY(shard_gram(w, w.G1, w.Rl, w.Rl, R, st));   // RtR
  double thr = tol * tol;
  if (R == 1) {
    double bb = 0.0;
    FSB_CUDA(cudaMemcpyAsync(&bb, w.G1, 8, cudaMemcpyDeviceToHost, st));
    FSB_CUDA(cudaStreamSynchronize(st));
    const double t = tol * sqrt(bb);
    thr = t * t;
  }
  const double work = (double)A->nnz * R;
  const int batch = work >= 2e8 ? 1 : 4;
  int queued = 0, np = 0;
  w.h_status[0] = w.h_status[1] = w.h_status[2] = 0;
  double t_prev = 0.0;
  cudaEvent_t pe[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // [6], [7]: inside the operator
  if (cg_trace_level() >= 2)
    for (auto& e : pe) cudaEventCreate(&e);
  if (cg_trace()) { cudaStreamSynchronize(st); t_prev = now_ms(); }
  // one iteration, enqueued on st (and comm_st) without looking at the device; ph: FSB_CG_TRACE=2 phase events
  auto enqueue_iteration = [&](double* g1, double* g2, bool ph) -> int {
      if (ph) cudaEventRecord(pe[0], st);
      FSB_TRY(shard_apply_op(A, Acsr, T, w, R, lambda, st, halves, ph ? pe + 6 : nullptr));
      if (ph) cudaEventRecord(pe[1], st);
      FSB_TRY(shard_gram(w, w.PtKP, w.Pl, w.KPl, R, st));
      FSB_TRY(fsb_dense_small_solve(w.Alpha, w.PtKP, g1, nullptr, 0, nullptr, 0, R, w.status, 0, 0.0, st));
      if (ph) cudaEventRecord(pe[2], st);
      FSB_TRY(fsb_dense_mix_add(w.Xl, w.Pl, w.Alpha, w.nloc, R, w.status, st));
      FSB_TRY(fsb_dense_mix_sub_gram(w.Rl, w.KPl, w.Alpha, w.partial, w.nloc, R, w.status, st, &np));
      FSB_TRY(fsb_dense_gram_finalize(g2, w.partial, np, R, st));
      FSB_TRY(shard_allreduce_small(w, g2, R * R, st));
      FSB_TRY(fsb_dense_small_solve(w.Psi, g1, g2, nullptr, 0, nullptr, 0, R, w.status, 1, thr, st));
      if (ph) cudaEventRecord(pe[3], st);
      FSB_TRY(fsb_dense_mix_set(w.Pl, w.Pl, w.Rl, w.Psi, w.nloc, R, w.status, st));
      if (ph) cudaEventRecord(pe[4], st);
      // optional (fsb_tune_cg_dist(3)): the next product takes P as two column halves, so that the second half's
      // all-gather can travel behind the first column pass.  Measured on 8 GPUs at C5 it gains nothing -- the
      // product slows down by what the hidden transfer saves (1.08 ms either way, profiles/r1i_cg_trace_n8_halves.log)
      // -- so the plain all-gather stays the default.
      const bool want_halves = g_cg_dist_mode == 3 && R % 4 == 0 && (R / 2) * 8 >= 128;
      if (want_halves) {
        FSB_TRY(shard_allgather_halves(w, R, st));
      } else {
        FSB_TRY(shard_allgather(w, w.Pfull, w.Pl, R, st));
      }
      halves = want_halves;
      if (ph) cudaEventRecord(pe[5], st);
      return FSB_OK;
  };
  // CUDA graph of the iteration (knob "cg_graph", default on): an 8-GPU iteration is ~2 ms of ~35 dependent launches
  // (kernels on two streams, NCCL, peer stores) and the host looks at the status words after every iteration -- enqueueing
  // them one by one leaves the GPU idle for the launch latencies.  The iteration is captured once per parity of the
  // G1 / G2 swap (after the first two iterations have run directly: they allocate scratch and time the products'
  // launch candidates) and replayed with one cudaGraphLaunch.  Falls back to direct enqueueing if capture is refused.
  const bool graph_ok = fsb_knob("cg_graph", FSB_MULTI_GPU_DEFAULTS) && cg_trace_level() < 2 && g_cg_dist_mode != 3;
  if (w.gthr != thr || w.glambda != lambda || w.gT != (const void*)T) { w.drop_graphs(); w.gthr = thr; w.glambda = lambda; w.gT = (const void*)T; }
  int direct_left = 2;     // iterations of this solve still to run uncaptured (only when no graph exists yet)
  bool graph_broken = false;
  while (queued < max_iter) {
    const int nb = std::min(batch, max_iter - queued);
    for (int k = 0; k < nb; ++k) {
      const bool ph = cg_trace_level() >= 2 && k == 0;      // FSB_CG_TRACE=2: device time of the phases of an iteration
      bool launched = false;
      const int parity = w.G1 == w.G1_first ? 0 : 1;
      if (graph_ok && !graph_broken && (w.gexec[parity] || direct_left <= 0)) {
        if (!w.gexec[parity]) {
          cudaGraph_t graph = nullptr;
          cudaError_t ce = cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
          int rc = ce == cudaSuccess ? enqueue_iteration(w.G1, w.G2, false) : FSB_ECUDA;
          if (ce == cudaSuccess) ce = cudaStreamEndCapture(st, &graph);
          if (ce == cudaSuccess && rc == FSB_OK) ce = cudaGraphInstantiate(&w.gexec[parity], graph, 0);
          if (graph) cudaGraphDestroy(graph);
          if (ce != cudaSuccess || rc != FSB_OK) {
            cudaGetLastError();
            w.gexec[parity] = nullptr;
            graph_broken = true;
            if (cg_trace()) fprintf(stderr, "[fsb cg %d/%d] graph capture refused (%s): enqueueing directly\n", w.rank, w.G, cudaGetErrorString(ce));
          }
        }
        if (w.gexec[parity]) {
          FSB_CUDA(cudaGraphLaunch(w.gexec[parity], st));
          launched = true;
        }
      }
      if (!launched) {
        FSB_TRY(enqueue_iteration(w.G1, w.G2, ph));
        --direct_left;
      }
      std::swap(w.G1, w.G2);
    }
    queued += nb;
    FSB_CUDA(cudaMemcpyAsync(w.h_status, w.status, kStatusWords * sizeof(int), cudaMemcpyDeviceToHost, st));
    FSB_CUDA(cudaStreamSynchronize(st));
    if (cg_trace()) {
      const double t = now_ms();
      fprintf(stderr, "[fsb cg %d/%d] iterations %d..%d: %.3f ms (status %d %d %d)\n", w.rank, w.G, queued - nb, queued - 1, t - t_prev,
              w.h_status[0], w.h_status[1], w.h_status[2]);
      if (cg_trace_level() >= 2) {
        float ms[5];
        for (int q = 0; q < 5; ++q) cudaEventElapsedTime(&ms[q], pe[q], pe[q + 1]);
        float op[3];
        cudaEventElapsedTime(&op[0], pe[0], pe[6]); cudaEventElapsedTime(&op[1], pe[6], pe[7]); cudaEventElapsedTime(&op[2], pe[7], pe[1]);
        if (w.rank == 0)
          fprintf(stderr, "[fsb cg %d/%d]   device ms: operator %.3f (A P %.3f, A' chunks %.3f, reduce-scatter tail + lambda P %.3f) | "
                          "P'KP + allreduce + solve %.3f | X, R updates + R'R + allreduce + solve %.3f | P update %.3f | all-gather of P %.3f\n",
                  w.rank, w.G, ms[0], op[0], op[1], op[2], ms[1], ms[2], ms[3], ms[4]);
      }
      t_prev = now_ms();
    }
    if (w.h_status[0] || w.h_status[1]) break;
  }
  const int it = w.h_status[2];
  for (auto& e : pe) if (e) cudaEventDestroy(e);
  if (w.p2p) FSB_TRY(fsb_p2p_check(w.p2p, st));
  if (w.p2p_kp) FSB_TRY(fsb_p2p_check(w.p2p_kp, st));
  // X = X_loc diag(norm), gathered into the replicated result (through the partial buffer: padded rows)
  for (int c = 0; c < w.C; ++c) FSB_TRY(fsb_dense_scale_cols(w.Xl + (size_t)c * w.s * R, w.norm, w.s, R, st));
  FSB_TRY(shard_allgather(w, w.KPpart, w.Xl, R, st));
  FSB_CUDA(cudaMemcpyAsync(dX, w.KPpart, (size_t)F * R * 8, cudaMemcpyDeviceToDevice, st));
  FSB_CUDA(cudaStreamSynchronize(st));
  if (out_iter) *out_iter = it;
  if (w.h_status[0])
    return fsb_set_error(FSB_EBREAKDOWN, "block CG: Gram matrix lost rank at iteration %d (R=%d)", it, R);
  return FSB_OK;
}


// The workspace of a solve (a few [F][R] vectors and the [N][R] intermediate: 3.6 GB at C5) stays
// with the handle between solves -- a sampler calls the solver thousands of times on one matrix, and
// cudaMalloc / cudaFree of gigabytes cost ~10 ms per solve (more with NCCL buffers registered).
struct CgCache {
  int kind = 0, R = 0, G = 1;
  long F = 0, N = 0;
  CgWork w;
  CgShardWork sw;
};
void cg_cache_free(void* p) {
  CgCache* c = static_cast<CgCache*>(p);
  if (!c) return;
  c->w.release();
  c->sw.release();
  delete c;
}

// one solve through whichever path applies to the handle
int cg_solve(fsb_matrix* A, fsb_matrix* At, double* dX, const double* dB, int R, double lambda, double tol, int max_iter,
             int* out_iter, cudaStream_t st) {
  const bool shard = A->sharded && fsb_comm_active() && g_cg_dist_mode != 1 && (long)A->ncol >= 64L * fsb_comm_size();
  const int kind = shard ? 2 : 1, G = shard ? fsb_comm_size() : 1;
  const double t0 = cg_trace() ? now_ms() : 0.0;
  CgCache* c = static_cast<CgCache*>(A->cg_cache);
  if (!c || c->kind != kind || c->R != R || c->F != A->ncol || c->N != A->nrow || c->G != G) {
    if (c) cg_cache_free(c);
    A->cg_cache = nullptr;
    c = new CgCache();
    c->kind = kind; c->R = R; c->G = G; c->F = A->ncol; c->N = A->nrow;
    const int rc = shard ? shard_alloc(c->sw, A->ncol, A->nrow, R) : cg_alloc(c->w, A->ncol, A->nrow, R);
    if (rc != FSB_OK) { cg_cache_free(c); return rc; }
    A->cg_cache = c;
    A->cg_cache_free = cg_cache_free;
  }
  const double t1 = cg_trace() ? now_ms() : 0.0;
  int rc;
  if (shard) {
    rc = cg_run_sharded(A, At, dX, dB, R, lambda, tol, max_iter, out_iter, st, c->sw);
    cudaStreamSynchronize(c->sw.comm_st);
  } else {
    rc = cg_run(A, At, dX, dB, R, lambda, tol, max_iter, out_iter, st, c->w);
  }
  if (cg_trace()) fprintf(stderr, "[fsb cg] workspace %.3f ms, solve %.3f ms (%s)\n", t1 - t0, now_ms() - t1, shard ? "sharded vectors" : "replicated vectors");
  return rc;
}

}  // namespace

extern "C" int fsb_cg_shard_layout(long F, int R, int G, int* C, long* s, long* Fc, long* Fp, long* nloc) {
  if (F < 0 || R < 1 || G < 1 || !C || !s || !Fc || !Fp || !nloc) return fsb_set_error(FSB_EINVAL, "fsb_cg_shard_layout: bad arguments");
  shard_layout(F, R, G, C, s, Fc, Fp, nloc);
  return FSB_OK;
}

extern "C" int fsb_tune_cg_dist(int mode) {
  if (mode < 0 || mode > 3) return fsb_set_error(FSB_EINVAL, "fsb_tune_cg_dist: mode must be 0 (sharded vectors), 1 (replicated), 2 (sharded, plain all-gather) or 3 (sharded, split all-gather forced)");
  g_cg_dist_mode = mode;
  return FSB_OK;
}

extern "C" int fsb_cg_dev(fsb_matrix_t A, fsb_matrix_t At, double* dX, const double* dB, int R, double lambda,
                          double tol, int max_iter, int* out_iter, void* stream) {
  FSB_TRY(fsb_require_device());
  if (!A || !dX || !dB) return fsb_set_error(FSB_EINVAL, "fsb_cg_dev: null argument");
  if (R < 1 || R > 32) return fsb_set_error(FSB_EINVAL, "fsb_cg_dev: R must be 1..32 (got %d)", R);
  if (At && (A->nrow != At->ncol || A->ncol != At->nrow))
    return fsb_set_error(FSB_EINVAL, "A (%d x %d) and At (%d x %d) must be transposes of each other.", A->nrow, A->ncol, At->nrow, At->ncol);
  if (!At && A->format != FSB_FMT_CSR) return fsb_set_error(FSB_EINVAL, "fsb_cg_dev: a stored transpose is required for non-CSR formats");
  cudaStream_t st = fsb_pick_stream(stream);
  int rc = cg_solve(A, At, dX, dB, R, lambda, tol, max_iter, out_iter, st);
  if (rc == FSB_EBREAKDOWN && R > 1) {
    // rank loss (a column converged early / dependent right-hand sides): the block
    // recurrence cannot continue; solve the columns one by one instead
    double *xb = nullptr, *bb = nullptr;
    const long F = A->ncol;
    int worst = 0;
    cudaError_t e = cudaMalloc(&xb, std::max<size_t>(F, 1) * 8);
    if (e == cudaSuccess) e = cudaMalloc(&bb, std::max<size_t>(F, 1) * 8);
    rc = e == cudaSuccess ? FSB_OK : fsb_cuda_error(e, "cudaMalloc", __FILE__, __LINE__);
    for (int k = 0; rc == FSB_OK && k < R; ++k) {
      e = cudaMemcpy2DAsync(bb, 8, dB + k, (size_t)R * 8, 8, F, cudaMemcpyDeviceToDevice, st);
      if (e != cudaSuccess) { rc = fsb_cuda_error(e, "column gather", __FILE__, __LINE__); break; }
      int it1 = 0;
      rc = cg_solve(A, At, xb, bb, 1, lambda, tol, max_iter, &it1, st);
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
    cudaFree(xb); cudaFree(bb);
  }
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
  if (rc == FSB_OK) rc = fsb_h2d(dB, B, (size_t)A->ncol * R * 8, st);
  if (rc == FSB_OK) rc = fsb_cg_dev(A, At, dX, dB, R, lambda, tol, max_iter, out_iter, (void*)st);
  if (rc == FSB_OK) rc = fsb_d2h(X, dX, (size_t)A->ncol * R * 8, st);
  if (rc == FSB_OK) {
    e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) rc = fsb_cuda_error(e, "D2H", __FILE__, __LINE__);
  }
  cudaFree(dX); cudaFree(dB);
  return rc;
}
