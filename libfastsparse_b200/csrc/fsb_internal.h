// fsb_internal.h -- shared declarations of libfastsparse_b200.so (not installed).
#pragma once

#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "../../include/fsb.h"

// Device-resident sparse matrix.  One struct for the three storage formats of the
// reference; unused fields are null.  Everything here lives in HBM.
struct fsb_matrix {
  int format = 0;
  int nrow = 0, ncol = 0;
  long nnz = 0;
  bool has_vals = false;
  bool sharded = false;       // rows are one rank's shard: A'(...) partials get allreduced
  // CSR / CBCSR (CBCSR: row_ptr has nblocks*nrow+1 entries)
  int* row_ptr = nullptr;
  int* cols = nullptr;
  double* vals = nullptr;
  int nblocks = 0, colblocksize = 0;
  // BLOCKED (row-blocked COO, flattened; entries keep the host order per block)
  int* start_row = nullptr;   // nblocks+1
  long* blk_off = nullptr;    // nblocks+1
  int* b_rows = nullptr;      // global row ids
  int* b_cols = nullptr;
  double* b_vals = nullptr;
  int max_block_rows = 0;
  // lazily built, cached transpose (CSR handles only)
  fsb_matrix* T = nullptr;
  // lazily built x-BLOCKED transpose for one right-hand side (fsb_build_transpose_xblocked): the entries of A' grouped
  // by contiguous blocks of A's rows, i.e. a CSR with tb_blocks * ncol "cell" rows (cell = block * ncol + column of A),
  // so that the slice of x a wave of CTAs gathers from stays L2-resident -- cbcsr.h's column blocking applied to A'
  fsb_matrix* Tb = nullptr;
  int tb_blocks = 0;
  size_t x_live_bytes = 0;    // != 0: bytes of the dense operand live at a time (one x block), for kernel-build choices
  // lazily built CSR view of a blocked / column-blocked matrix (same entries, stable by row, so
  // every row keeps its stored order); products default to the CSR kernels through it
  fsb_matrix* view = nullptr;
  // library-owned scratch (A X intermediate of A'A, host staging)
  double* tmp = nullptr;
  size_t tmp_cap = 0;
  double* xpack = nullptr;    // dense operand repacked into contiguous column slabs [S][ncol][R/S] (fsb_launch_csr_spmm)
  size_t xpack_cap = 0;
  const double* xpack_src = nullptr;   // != nullptr: xpack already holds this operand (set around a chunked host product)
  fsb_matrix* scratch_owner = nullptr; // row-range alias of another handle: scratch and tuning state live in the owner
  double* carry = nullptr;    // per-tile carries of the merge-path stream kernel
  size_t carry_cap = 0;
  int* split = nullptr;       // cached merge-path tile boundaries (rows complete at each tile start)
  int split_tile = 0;
  int max_row_nnz = -1;       // longest row (lazy; SpMV kernel choice)
  // autotune of the staged SpMM (see fsb_launch_csr_spmm), one slot per operand width R that has been timed on
  // this handle (a CG solve alternates R = 32 products with R = 1 ones: no re-tuning when R changes back):
  struct Tuned { int R = 0, passes = 1, deep = 0; };
  Tuned tuned[6];             //   passes = column passes over the dense operand (1 or 2), deep = kernel build
  int tuned_next = 0;         //   ring cursor
  int tuned_last = -1;        //   slot used by the latest product (fsb_matrix_tuning reports it)
  size_t bytes = 0;
  double avg_row_nnz = 0.0;
  // solver workspace kept between solves on this handle (fsb_cg.cu); released with the handle
  void* cg_cache = nullptr;
  void (*cg_cache_free)(void*) = nullptr;
};

// default of the multi-GPU fast paths (peer-store all-gather of P, peer Gram allreduce, CUDA-graph CG iteration, sharded
// upload of X): verified on 2 and 8 B200 (profiles/r2e_*, r2h_*: results identical to NCCL, all ranks bit-identical);
// the knobs cg_p2p / cg_p2p_gram / cg_graph / host_x_allgather override.  The reduce-scatter by pull (cg_p2p_rs) stays
// off: 2.385 against 2.351 ms per iteration on 8 GPUs -- its loads compete with the products it overlaps.
#ifndef FSB_MULTI_GPU_DEFAULTS
#define FSB_MULTI_GPU_DEFAULTS 1
#endif

// ---- error plumbing (fsb_runtime.cu)
int fsb_set_error(int code, const char* fmt, ...);
int fsb_cuda_error(cudaError_t e, const char* what, const char* file, int line);
int fsb_require_device();
cudaStream_t fsb_default_stream();
void fsb_count_launch(int n = 1);
// experiment knob (per calling thread; fsb_tune sets it, FSB_TUNE_<NAME> in the environment is the default)
int fsb_knob(const char* name, int dflt);
// Linear texture object over a device buffer of `texels` elements of 8 (int2) or 16 (int4) bytes, for gathers through
// the texture pipe (tex1Dfetch): its data stage is separate from the LSU's, which shared-memory traffic also uses
// (profiles/r2z_tex_gathers.md).  Objects are descriptors only (no ownership of the buffer); a small per-thread table
// keeps them, the least recently used one is destroyed after a synchronise of `st`.  A linear texture must start on a
// 512-byte boundary: the object starts at the boundary at or below p and *texel_off receives the index of p's first texel
// in it (add it to every fetch index).  Returns 0 when the buffer does not qualify (p not texel-aligned, more than 2^27
// texels) or the object cannot be created: callers fall back to plain loads.
cudaTextureObject_t fsb_linear_texture(const void* p, size_t texels, int texel_bytes, cudaStream_t st, int* texel_off);

#define FSB_CUDA(call)                                                        \
  do {                                                                        \
    cudaError_t e__ = (call);                                                 \
    if (e__ != cudaSuccess) return fsb_cuda_error(e__, #call, __FILE__, __LINE__); \
  } while (0)

#define FSB_TRY(call)            \
  do {                           \
    int rc__ = (call);           \
    if (rc__ != FSB_OK) return rc__; \
  } while (0)

#define FSB_KERNEL_CHECK()       \
  do {                           \
    fsb_count_launch();          \
    FSB_CUDA(cudaGetLastError()); \
  } while (0)

static inline cudaStream_t fsb_pick_stream(void* s) {
  return s ? (cudaStream_t)s : fsb_default_stream();
}

// ---- fsb_hostcopy.cu: host <-> device copies that stay at PCIe speed for pageable (malloc'd) host memory
bool fsb_host_is_pageable(const void* p);
int fsb_h2d(void* dst_dev, const void* src_host, size_t bytes, cudaStream_t st);
int fsb_d2h(void* dst_host, const void* src_dev, size_t bytes, cudaStream_t st);
// many small host arrays <-> one contiguous device range, packed through the pinned ring (blocked formats' per-block arrays)
int fsb_h2d_gather(void* dst_dev, const void* const* srcs_host, const size_t* bytes, long n, cudaStream_t st);
int fsb_d2h_scatter(void* const* dsts_host, const void* src_dev, const size_t* bytes, long n, cudaStream_t st);
int fsb_d2h_segments(int nseg, void* const* dst_host, const void* const* src_dev, const size_t* bytes, const cudaEvent_t* ready,
                     cudaStream_t st);

// ---- device scratch owned by a handle
int fsb_matrix_scratch(fsb_matrix* A, size_t bytes, double** out);
int fsb_matrix_carry(fsb_matrix* A, size_t bytes, double** out);

// ---- kernels_csr.cu
// Y = A X (+ lambda Z when dZ != nullptr; Z is [nrow][R] like Y -- fused into the staged kernel's
// epilogue, a separate pass for the other kernels)
int fsb_launch_csr_spmm(fsb_matrix* A, double* dY, const double* dX, int R, cudaStream_t st, const double* dZ = nullptr,
                        double lambda = 0.0);
// ---- kernels_csr_stream.cu (merge-path SpMV / narrow SpMM, R = 1, 2, 4)
bool fsb_csr_stream_supports(int R);
bool fsb_csr_stream_preferred(int R);
int fsb_launch_csr_stream(fsb_matrix* A, double* dY, const double* dX, int R, cudaStream_t st);
int fsb_launch_csr_ata_fused(const fsb_matrix* A, double* dY, const double* dX, int R, double lambda, cudaStream_t st);
// tuning override for the sweep tool: TW, G, VEC, slabs (0 = heuristic)
void fsb_csr_spmm_set_tuning(int tw, int g, int vec, int slabs);
// ---- kernels_csr_staged.cu
// ldx > 0: the dense operand of this pass is a column slab stored on its own ([ncol][ldx], first column xcol0)
int fsb_launch_csr_spmm_staged(const fsb_matrix* A, double* dY, const double* dX, int R, int col0, int ncols,
                               int g, int vec, cudaStream_t st, const double* dZ = nullptr, double lambda = 0.0, bool deep = false,
                               int ldx = 0, int xcol0 = 0);
// Y = A X with X given as two column halves Xlo, Xhi ([ncol][R/2] each): two column passes, each waiting on
// the event that says its half has arrived (the all-gather of the sharded CG, fsb_cg.cu).  R even, R/2 * 8 >= 128.
int fsb_launch_csr_spmm_halves(fsb_matrix* A, double* dY, const double* dXlo, const double* dXhi, int R, cudaStream_t st,
                               cudaEvent_t ready_lo, cudaEvent_t ready_hi);
// A row-range ALIAS of a CSR handle (row_ptr + r0 with absolute offsets, shared cols / vals; lives on the caller's
// stack) sets scratch_owner: scratch buffers and the autotune table live in the owner, and the dispatcher keeps
// aliases on the row-local staged kernel (the merge-path kernel needs row_ptr[0] == 0 and per-handle tile tables).
inline fsb_matrix* fsb_home(fsb_matrix* A) { return A->scratch_owner ? A->scratch_owner : A; }
inline fsb_matrix::Tuned* fsb_tuned_find(fsb_matrix* A, int R) {
  fsb_matrix* Hm = fsb_home(A);
  for (int i = 0; i < 6; ++i)
    if (Hm->tuned[i].R == R) { Hm->tuned_last = i; return &Hm->tuned[i]; }
  return nullptr;
}
inline void fsb_tuned_store(fsb_matrix* A, int R, int passes, int deep) {
  fsb_matrix* Hm = fsb_home(A);
  const int i = Hm->tuned_next;
  Hm->tuned[i].R = R; Hm->tuned[i].passes = passes; Hm->tuned[i].deep = deep;
  Hm->tuned_last = i;
  Hm->tuned_next = (i + 1) % 6;
}
inline void fsb_make_row_alias(fsb_matrix* part, fsb_matrix* C, int r0, int r1) {
  part->format = FSB_FMT_CSR; part->nrow = r1 - r0; part->ncol = C->ncol; part->nnz = C->nnz; part->has_vals = C->has_vals;
  part->row_ptr = C->row_ptr + r0; part->cols = C->cols; part->vals = C->vals; part->avg_row_nnz = C->avg_row_nnz;
  part->scratch_owner = fsb_home(C);
}
void fsb_csr_staged_set_tuning(int rows_per_cta, int cap_mult);

// ---- kernels_cbcsr.cu / kernels_blocked.cu
int fsb_launch_cbcsr_spmm(const fsb_matrix* A, double* dY, const double* dX, int R, cudaStream_t st);
int fsb_launch_blocked_spmm(const fsb_matrix* A, double* dY, const double* dX, int R, cudaStream_t st);

// ---- kernels_build.cu
// input validation on the device (an out-of-range index would be an illegal address in the products)
int fsb_check_index_range(const int* d_idx, long n, int limit, const char* what, cudaStream_t st);
int fsb_check_row_ptr(const int* d_ptr, long n, long last, const char* what, cudaStream_t st);
// every entry i of a blocked COO (block b = the one with blk_off[b] <= i < blk_off[b+1]) has start_row[b] <= rows[i] < start_row[b+1]?
int fsb_check_rows_in_blocks(const int* d_rows, const long* d_blk_off, const int* d_start_row, int nblocks, long nnz, cudaStream_t st);
int fsb_build_csr_from_coo_dev(fsb_matrix* out, int nrow, int ncol, long nnz, const int* d_rows,
                               const int* d_cols, const double* d_vals, cudaStream_t st);
int fsb_build_transpose(fsb_matrix* A, cudaStream_t st);   // fills A->T
int fsb_build_transpose_xblocked(fsb_matrix* A, size_t block_bytes, cudaStream_t st);   // fills A->Tb / A->tb_blocks
int fsb_build_csr_view(fsb_matrix* A, cudaStream_t st);    // fills A->view (BLOCKED / CBCSR)
int fsb_stable_perm_by_key(const int* d_keys, int nkeys, long n, int* d_perm, int* d_ptr, cudaStream_t st);
#define FSB_BLOCKED_CLASSES 256
int fsb_blocked_relayout(fsb_matrix* A, cudaStream_t st);  // bucket entries by row class; fills A->row_ptr

// ---- kernels_dense.cu (CG building blocks)
int fsb_dense_axpy_lambda(double* dY, const double* dX, double lambda, long n, cudaStream_t st); // Y += lambda X

// ---- fsb_comm.cu
bool fsb_comm_active();
int fsb_comm_reduce_scatter_sum(const double* send, double* recv, size_t recvcount, cudaStream_t st);
int fsb_comm_allgather(const double* send, double* recv, size_t sendcount, cudaStream_t st);
int fsb_comm_group_start();
int fsb_comm_group_end();
// peer-memory (CUDA IPC over NVLink) symmetric buffer + all-gather by direct peer stores
struct fsb_p2p;
int fsb_p2p_create(fsb_p2p** out, size_t bytes, cudaStream_t st);   // collective; error = keep using NCCL
void fsb_p2p_destroy(fsb_p2p* p);
void* fsb_p2p_local(fsb_p2p* p);
int fsb_p2p_allgather_chunks(fsb_p2p* p, const double* loc, int C, long slice_doubles, cudaStream_t st);
int fsb_p2p_allreduce_small(fsb_p2p* p, double* buf, int n, cudaStream_t st);   // n <= 1024, in place
int fsb_p2p_signal(fsb_p2p* p, int channel, cudaStream_t st);
int fsb_p2p_pull_sum(fsb_p2p* p, int channel, double* out, size_t elem_offset, long n, const double* add, double lambda, cudaStream_t st);
int fsb_p2p_check(fsb_p2p* p, cudaStream_t st);
