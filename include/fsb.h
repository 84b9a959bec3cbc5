/*
 * fsb.h -- C ABI of libfastsparse_b200.so: the B200 (sm_100a) implementation of
 * libfastsparse's sparse x dense hot path.
 *
 * The reference (jaak-s/libfastsparse) has no FFI layer: its API is a set of C
 * structs and inline functions in csr.h / sparse.h / dsparse.h / cbcsr.h / cg.h.
 * The drop-in headers under include/fastsparse/ keep those names and struct
 * layouts and forward every multiply / solve to the entry points below; each
 * entry point names the reference function(s) it replaces (file:line relative
 * to the reference tree).
 *
 * Conventions
 *  - plain pointers and sizes only; int32 indices, fp64 values, dense operands
 *    row-major [n][R] ("row-ordered", csr.h:163,256);
 *  - every function returns 0 on success or an FSB_E* code; the message is
 *    available from fsb_last_error() (thread-local);
 *  - *_host entry points take HOST pointers (they copy X in and Y out: the
 *    drop-in path); *_dev entry points take DEVICE pointers, are asynchronous
 *    on `stream` (a cudaStream_t passed as void*, NULL = the library's stream)
 *    and never touch the host -- with two documented exceptions: the first
 *    multi-RHS product on a large handle times its launch candidates (it
 *    synchronises once, see fsb_matrix_tuning), and the solver reads three
 *    status words back per batch of iterations;
 *  - outputs are fully overwritten (alpha = 1, beta = 0), like the reference;
 *  - there is NO CPU fallback: without a CUDA device every compute call fails
 *    with FSB_ENODEV.  The fsb_host_* functions are the reference's host-side
 *    constructors/loaders (bit-exact structure) and need no GPU.
 */
#ifndef FSB_H
#define FSB_H

#ifdef __cplusplus
extern "C" {
#endif

#define FSB_OK        0
#define FSB_EINVAL    1   /* bad argument / shape mismatch                  */
#define FSB_ENODEV    2   /* no usable CUDA device (never falls back to CPU) */
#define FSB_ECUDA     3   /* CUDA runtime error                             */
#define FSB_ENOMEM    4
#define FSB_EIO       5   /* file error                                     */
#define FSB_ENCCL     6   /* NCCL missing or failed                         */
#define FSB_EBREAKDOWN 7  /* block-CG Gram matrix lost rank                 */

/* Opaque device-resident sparse matrix in one of the reference's formats. */
typedef struct fsb_matrix* fsb_matrix_t;

enum fsb_format {
  FSB_FMT_CSR = 1,     /* struct BinaryCSR csr.h:15-22 / struct CSR csr.h:358-366 */
  FSB_FMT_CBCSR = 2,   /* struct ColBinaryCSR cbcsr.h:5-14                        */
  FSB_FMT_BLOCKED = 3  /* struct BlockedSBM sparse.h:163-172 / BlockedSDM dsparse.h:119-129 */
};

/* ---------------------------------------------------------------- runtime */
int fsb_version(void);
const char* fsb_last_error(void);
int fsb_last_error_code(void);           /* the FSB_E* code that came with fsb_last_error() (for entry points returning handles) */
int fsb_device_count(void);              /* 0 when no GPU / no driver       */
int fsb_init(int device);                /* bind this thread's context; idempotent */
int fsb_sync(void);                      /* synchronise the library stream  */
void* fsb_stream(void);                  /* the library's cudaStream_t      */
/* number of kernels this library has launched since load (bench.py gpu_launches) */
long fsb_launch_count(void);

/* ------------------------------------------------- upload (host -> HBM) */
/* Every upload / load path validates its index arrays once on the device (0 <= index < dimension,
 * offsets non-decreasing from 0 to nnz, nnz < 2^31) and returns FSB_EINVAL on a violation: an
 * out-of-range index -- undefined behaviour in the reference -- would otherwise be an illegal
 * device address. */
/* CSR, binary when vals == NULL.  Replaces the storage side of new_bcsr
 * (csr.h:30-67) / new_csr (csr.h:375-422): arrays are copied verbatim, so
 * row order, in-row order and duplicates are exactly the host structure's. */
int fsb_csr_upload(fsb_matrix_t* out, int nrow, int ncol, long nnz,
                   const int* row_ptr, const int* cols, const double* vals);
/* COO -> CSR built ON THE DEVICE by a stable sort on the row index; the result
 * is bit-identical to new_bcsr / new_csr on the same COO.
 * Pointers are HOST pointers. */
int fsb_csr_upload_coo(fsb_matrix_t* out, int nrow, int ncol, long nnz,
                       const int* rows, const int* cols, const double* vals);
/* same, COO already in HBM (device pointers); the inputs are not modified */
int fsb_csr_from_coo_dev(fsb_matrix_t* out, int nrow, int ncol, long nnz,
                         const int* d_rows, const int* d_cols, const double* d_vals);
/* column-blocked binary CSR (new_cbcsr cbcsr.h:16-65); row_ptr has
 * nblocks*nrow+1 entries, cell = block*nrow + row */
int fsb_cbcsr_upload(fsb_matrix_t* out, int nrow, int ncol, int nblocks,
                     int colblocksize, long nnz, const int* row_ptr, const int* cols);
/* row-blocked COO (new_bsbm sparse.h:175-213 / new_bsdm dsparse.h:132-173):
 * per-block arrays exactly as the reference stores them; vals == NULL => binary */
int fsb_blocked_upload(fsb_matrix_t* out, int nrow, int ncol, int nblocks,
                       const int* start_row, const int* blk_nnz,
                       int* const* rows, int* const* cols, double* const* vals);
int fsb_matrix_free(fsb_matrix_t A);
/* shape query; any out pointer may be NULL */
int fsb_matrix_info(fsb_matrix_t A, int* format, int* nrow, int* ncol, long* nnz,
                    int* has_vals, int* nblocks);
/* bytes of HBM held by the handle (structure + cached transposes) */
long fsb_matrix_bytes(fsb_matrix_t A);
/* the staged SpMM's per-handle autotune result (transposed != 0: of the cached transpose):
 * *R = width it was timed for (0 = not yet), *passes = column passes (1 | 2), *deep = kernel build */
int fsb_matrix_tuning(fsb_matrix_t A, int transposed, int* R, int* passes, int* deep);
/* row-blocked COO built on the device from a device COO; order 0 = COO order kept
 * (new_bsbm), 1 = per-block Hilbert order (new_bsbm + sort_bsbm sparse.h:215-236),
 * 2 = per-block row-major order (sort_bsbm_byrow sparse.h:238-256) */
int fsb_blocked_from_coo_dev(fsb_matrix_t* out, int nrow, int ncol, long nnz, const int* d_rows,
                             const int* d_cols, const double* d_vals, int block_size, int order);
/* global Hilbert order of a COO (sort_sbm sparse.h:142-161 / sort_sdm dsparse.h:96-115) on the device: key =
 * xy2d(ceilPower2(max(nrow, ncol)), row, col), radix sort, coordinates decoded back with d2xy; identical to the host
 * routine whenever coordinates are unique (equal keys keep their input order).  DEVICE arrays, sorted in place;
 * d_vals == NULL for binary matrices. */
int fsb_sort_coo_hilbert_dev(int nrow, int ncol, long nnz, int* d_rows, int* d_cols, double* d_vals);
/* same on HOST arrays (upload, sort on the device, copy back): what the drop-in sort_sbm / sort_sdm call for matrices
 * of at least FSB_SORT_DEVICE_MIN entries (default 2^20) when a device is present */
int fsb_sort_coo_hilbert(int nrow, int ncol, long nnz, int* rows, int* cols, double* vals);
/* the drop-in's choice: device sort at >= FSB_SORT_DEVICE_MIN entries when a device is present, else the host routine */
int fsb_sort_coo_hilbert_auto(int nrow, int ncol, long nnz, int* rows, int* cols, double* vals);
/* per-block orders of a HOST BlockedSBM / BlockedSDM on the device (upload, one keyed radix sort over all blocks, copy
 * back): order 1 = sort_bsbm sparse.h:215-236 / sort_bsdm dsparse.h:193-216 (row_xy2d Hilbert key), order 2 =
 * sort_bsbm_byrow sparse.h:238-256; vals == NULL for binary.  _auto: device at >= FSB_SORT_DEVICE_MIN entries, else host. */
int fsb_sort_blocked(int nrow, int ncol, int nblocks, const int* start_row, const int* blk_nnz, int* const* rows,
                     int* const* cols, double* const* vals, int order);
int fsb_sort_blocked_auto(int nrow, int ncol, int nblocks, const int* start_row, const int* blk_nnz, int* const* rows,
                          int* const* cols, double* const* vals, int order);
/* column-blocked binary CSR built on the device (new_cbcsr cbcsr.h:16-65) */
int fsb_cbcsr_from_coo_dev(fsb_matrix_t* out, int nrow, int ncol, long nnz, const int* d_rows,
                           const int* d_cols, int colblocksize);
/* copy the device CSR structure back (row_ptr nrow+1, cols nnz, vals nnz or NULL) */
int fsb_csr_download(fsb_matrix_t A, int* row_ptr, int* cols, double* vals);
/* a new handle holding rows [r0, r1) of a CSR handle (device-side slice,
 * row_ptr rebased): the row shard one rank owns */
int fsb_csr_row_slice(fsb_matrix_t* out, fsb_matrix_t A, int r0, int r1);

/* ------------------------------------------------------------- products */
/* Y[nrow][R] = A X[ncol][R] for every format.  Replaces
 *   bcsr_A_mul_B csr.h:149-161, _B2 164-181, _B4 184-202, _B8 205-223,
 *   _B8_auto 225-254, _Bn 257-280, _B32n 283-302,
 *   csr_A_mul_B csr.h:425-438, csr_A_mul_Bn 441-465,
 *   cbcsr_A_mul_B cbcsr.h:76-106 (+ the n-RHS variant the reference lacks),
 *   bsbm_A_mul_B/_B2/_B4/_Bn sparse.h:259-336, bsdm_A_mul_B dsparse.h:176-191. */
int fsb_spmm_dev(fsb_matrix_t A, double* dY, const double* dX, int R, void* stream);
int fsb_spmm_host(fsb_matrix_t A, double* Y, const double* X, int R);
/* Y[ncol][R] = A' X[nrow][R] (CSR handles).  Replaces At_mul_B sparse.h:68-75 and
 * sdm_At_mul_B dsparse.h:54-62.  Deterministic: uses a transpose built once on
 * the device and cached in the handle. */
int fsb_spmm_t_dev(fsb_matrix_t A, double* dY, const double* dX, int R, void* stream);
int fsb_spmm_t_host(fsb_matrix_t A, double* Y, const double* X, int R);
/* Y[ncol][R] = A'(A X) + lambda X.  Replaces bcsr_AA_mul_B csr.h:305-319,
 * parallel_bcsr_AA_mul_B csr.h:323-355 (lambda = 0, R = 1) and the operator of
 * bsbm_AtA cg.h:9-22.  dTmp: nrow*R doubles of scratch (NULL: library-owned).
 * mode 0 = two deterministic gather passes (A, then cached A'),
 * mode 1 = one fused pass over A with fp64 red.global.add scatter (R = 1..32). */
int fsb_ata_dev(fsb_matrix_t A, double* dY, const double* dX, int R, double lambda,
                double* dTmp, int mode, void* stream);
int fsb_ata_host(fsb_matrix_t A, double* Y, const double* X, int R, double lambda, int mode);
/* y = At (A x) + lambda x with an explicitly stored transpose, as cg.h:9-22 */
int fsb_ata_pair_dev(fsb_matrix_t A, fsb_matrix_t At, double* dY, const double* dX,
                     int R, double lambda, double* dTmp, void* stream);
/* host operands; tmp (nrow*R doubles, may be NULL) receives A x like the reference's scratch */
int fsb_ata_pair_host(fsb_matrix_t A, fsb_matrix_t At, double* Y, const double* X, int R,
                      double lambda, double* tmp);

/* --------------------------------------------------------------- solver */
/* Block conjugate gradient for (A'A + lambda I) X = B with R right-hand sides,
 * X, B row-major [F][R].  R = 1 follows bsbm_cg (cg.h:25-82: stop when
 * ||r|| <= tol ||b||); R >= 2 follows bsbm_cg2 (cg.h:85-187: column-normalised
 * block CG, stop when every diag(R'R) <= tol^2), generalised from the closed-form
 * 2x2 solves to an R x R Cholesky solve on the device.  At may be NULL (the
 * cached device transpose of A is used).  max_iter <= 0 means F, like the
 * reference.  Returns the reference's iteration counter in *out_iter. */
int fsb_cg_host(fsb_matrix_t A, fsb_matrix_t At, double* X, const double* B, int R,
                double lambda, double tol, int max_iter, int* out_iter);
int fsb_cg_dev(fsb_matrix_t A, fsb_matrix_t At, double* dX, const double* dB, int R,
               double lambda, double tol, int max_iter, int* out_iter, void* stream);

/* ------------------------------- device memory without the CUDA toolkit */
/* For plain C callers of the *_dev face (examples/sampler_loop.c): buffers on the library's
 * device, copies ordered after the library's default stream (stream = NULL in the *_dev calls). */
void* fsb_device_malloc(size_t bytes);                 /* NULL on failure (fsb_last_error) */
int fsb_device_free(void* p);
int fsb_copy_to_device(void* dst_dev, const void* src_host, size_t bytes);
int fsb_copy_to_host(void* dst_host, const void* src_dev, size_t bytes);

/* ------------------------------- Macau-style caller loop */
/* d[i] = standard normal, a pure function of (seed, i) (counter-based Box-Muller); the host twin
 * regenerates the same stream (to the last ulps of libm) for checks. */
int fsb_randn_dev(double* d, long n, unsigned long long seed, void* stream);
int fsb_randn_host(double* x, long n, unsigned long long seed);
/* dB[ncol][R] = A'N + sqrt(lambda) E with fresh noise N [nrow][R], E [ncol][R] drawn on the device
 * (the right-hand side of one sampling step, bench_a_mul_b.c:334-347); A stays resident, the
 * sqrt(lambda) E term is fused into the A' product.  N = randn(seed ^ k), E = randn(seed + 0x5bd1e995).
 * Follow with fsb_cg_dev for the solve; repeat with a new seed for the next sample. */
int fsb_noise_rhs_dev(fsb_matrix_t A, fsb_matrix_t At, double* dB, int R, double lambda,
                      unsigned long long seed, void* stream);

/* ------------------------------------------- dense helpers (linalg.h) */
/* G[Ra][Rb] = Xa' Xb over n rows (row-major result, device pointers; out on host).
 * Replaces pnormsq/pnormsq2/pouter2/pdot/pdot2sym linalg.h:15-73. */
int fsb_gram_dev(double* G_host, const double* dXa, const double* dXb, long n, int R, void* stream);
/* same with HOST operands (the drop-in linalg.h path): copies Xa, Xb in, reduces on the GPU */
int fsb_gram_host(double* G, const double* Xa, const double* Xb, long n, int R);
/* Tall-skinny row mix with an R x R coefficient matrix dM (device, row-major), R <= 32; the
 * vector updates of bsbm_cg / bsbm_cg2 (cg.h:60-63,70-73,148-154,165-170):
 *   mode 0: O += I M;   mode 1: O -= I M and, when G_host != NULL, G_host = O'O of the updated O;
 *   mode 2: O = Add + I M (I may alias O).   All operands [n][R] device pointers. */
int fsb_rowmix_dev(int mode, double* dO, const double* dI, const double* dAdd, const double* dM,
                   double* G_host, long n, int R, void* stream);
/* *out = sqrt(sum (x[i]-y[i])^2) on the GPU, host operands.  Replaces dist linalg.h:6-13. */
int fsb_dist_host(double* out, const double* x, const double* y, long n);

/* ------------------------------------------------------------ multi-GPU */
/* One process per GPU.  Rank 0 calls fsb_comm_unique_id, the bytes travel by any
 * host channel (torch.distributed, MPI, a file), every rank calls fsb_comm_init.
 * NCCL is loaded with dlopen at this point; nothing else in the library needs it.
 * When a communicator is active and `sharded` was set on the handle
 * (fsb_matrix_set_row_sharded), fsb_spmm_t_* / fsb_ata_* / fsb_cg_* sum-allreduce
 * the per-shard partial A'(...) and the CG Gram matrices. */
#define FSB_UNIQUE_ID_BYTES 128
int fsb_comm_unique_id(void* id_out);
int fsb_comm_init(int nranks, int rank, const void* id);
int fsb_comm_finalize(void);
int fsb_comm_size(void);
int fsb_comm_rank(void);
int fsb_allreduce_sum_dev(double* dBuf, long count, void* stream);
int fsb_matrix_set_row_sharded(fsb_matrix_t A, int sharded);
/* geometry of the multi-GPU block CG's F-sharded vectors on G ranks: the F unknowns are cut into *C chunks of
 * *Fc = G * *s rows (zero-padded to *Fp = *C * *Fc); rank g owns rows [c * *Fc + g * *s, + *s) of every chunk c,
 * stored back to back (*nloc = *C * *s local rows).  Chunk c of the A'(A P) partial is reduce-scattered while
 * chunk c+1 is computed; P is all-gathered chunk by chunk.  Pure host arithmetic (no device needed). */
int fsb_cg_shard_layout(long F, int R, int G, int* C, long* s, long* Fc, long* Fp, long* nloc);
/* nnz-balanced contiguous row partition: bounds[p]..bounds[p+1] are the rows of
 * part p, chosen from row_ptr so each part holds ~nnz/nparts entries.  Host only. */
int fsb_partition_rows(int nrow, const int* row_ptr, int nparts, int* bounds);

/* --------------------------------------- host-side structure (no GPU) */
/* Bit-exact restatements of the reference's host constructors and integer maths;
 * they exist so the drop-in headers are thin and so structure parity can be
 * tested without a device. */
int  fsb_host_ceil_pow2(int x);                               /* hilbert.h:11-13 */
long fsb_host_xy2d(int n, int x, int y);                      /* hilbert.h:16-27 */
void fsb_host_d2xy(int n, long d, int* x, int* y);            /* hilbert.h:30-42 */
long fsb_host_row_xy2d(int n, int x, int y);                  /* hilbert.h:60-65 */
void fsb_host_row_d2xy(int n, long d, int* x, int* y);        /* hilbert.h:68-75 */
void fsb_host_sort_keys(long* keys, double* payload, long n); /* quickSort.h / quickSortD.h order */
int fsb_host_csr_from_coo(long nnz, int nrow, const int* rows, const int* cols,
                          const double* vals, int* row_ptr, int* out_cols,
                          double* out_vals);                  /* csr.h:30-67, 375-422 */
int fsb_host_cbcsr_nblocks(int ncol, int colblocksize);       /* cbcsr.h:27 */
int fsb_host_cbcsr_from_coo(int colblocksize, long nnz, int nrow, int ncol,
                            const int* rows, const int* cols, int* row_ptr,
                            int* out_cols);                   /* cbcsr.h:16-65 */
int fsb_host_blocked_nblocks(int nrow, int block_size);       /* sparse.h:179 */
/* fills start_row[nblocks+1], blk_nnz[nblocks] and the per-block arrays whose
 * pointers the caller allocated after a first call with rows_out == NULL */
int fsb_host_blocked_count(long nnz, int nrow, int block_size, const int* rows,
                           int* start_row, int* blk_nnz);     /* sparse.h:175-195 */
int fsb_host_blocked_fill(long nnz, int block_size, const int* rows, const int* cols,
                          const double* vals, int nblocks, int* const* rows_out,
                          int* const* cols_out, double* const* vals_out); /* sparse.h:196-212 */
int fsb_host_sort_coo_hilbert(int nrow, int ncol, long nnz, int* rows, int* cols,
                              double* vals);                  /* sparse.h:142-161, dsparse.h:96-115 */
int fsb_host_sort_block_hilbert(int start_row, int nrows_in_block, long nnz, int* rows,
                                int* cols, double* vals);     /* sparse.h:215-236, dsparse.h:193-216 */
int fsb_host_sort_block_byrow(int ncol, long nnz, int* rows, int* cols);  /* sparse.h:238-256 */
/* one native 8-byte integer from an open FILE* (read_long utils.h:4-12); *ok = 0 on a short read */
long fsb_host_read_long(void* file, int* ok);
/* raw COO files (read_sbm sparse.h:112-139, read_sdm dsparse.h:64-93);
 * first call with rows == NULL returns the header */
int fsb_host_read_coo(const char* path, long* nrow, long* ncol, long* nnz, int* rows,
                      int* cols, double* vals);
/* .csr.bin (serialize_to_file csr.h:97-113, deserialize_from_file csr.h:117-146);
 * struct_image: the caller's 32-byte struct BinaryCSR, written verbatim */
int fsb_host_write_csr_bin(const char* path, const void* struct_image, int nrow,
                           long nnz, const int* row_ptr, const int* cols);
int fsb_host_read_csr_bin(const char* path, void* struct_image, int* row_ptr, int* cols);

/* ---------------------------------------------- files straight into HBM */
/* The same files, read in chunks through pinned staging buffers so the H2D copy overlaps the
 * read, finished on the device; no host copy of the matrix is made.
 * raw COO (read_sbm sparse.h:112-139 / read_sdm dsparse.h:64-93; with_vals = 1 for the latter)
 * -> CSR handle, entries of a row in file order exactly like new_bcsr / new_csr of the loaded COO */
int fsb_csr_load_coo_file(fsb_matrix_t* out, const char* path, int with_vals);
/* .csr.bin (serialize_to_file csr.h:97-113 / deserialize_from_file csr.h:117-146) -> CSR handle;
 * struct_image (nullable): receives the file's raw 32-byte struct BinaryCSR */
int fsb_csr_load_bin_file(fsb_matrix_t* out, const char* path, void* struct_image);

/* ------------------------------------------------- drop-in plumbing */
/* Used by the drop-in headers in include/fastsparse/.  A handle is cached per host structure, keyed by its array
 * pointers, and VALIDATED BY CONTENT on every lookup: the reference reads the caller's arrays on every call, so an
 * in-place edit (or a free + re-allocation at the same address) must never be answered with the old matrix.
 *   default / FSB_CACHE=1 : exact -- a 64-bit hash of the full content of every array.  Small structures are
 *        hashed inline; large ones are hashed by a worker thread while the product already runs on the cached
 *        copy, and fsb_cache_settle() reports whether that copy was stale (the header then repeats the call);
 *   FSB_CACHE=fast : 256 strided samples per array only (for callers that promise not to edit in place);
 *   FSB_CACHE=0    : no caching, every call uploads; the handles are released by fsb_cache_settle().
 * The protocol every drop-in call follows is FSB_DROPIN_CALL below: lookups, the product, fsb_cache_settle();
 * repeat while it returns non-zero.  Handles returned by fsb_cache_* are valid until the calling thread's next
 * fsb_cache_settle().  Mutating entry points (sort_*, transpose, free_*) also call fsb_cache_drop at once.
 * Least-recently-used entries are evicted above FSB_CACHE_MAX_MB (default 65536) of HBM or 64 entries. */
fsb_matrix_t fsb_cache_csr(int nrow, int ncol, long nnz, const int* row_ptr,
                           const int* cols, const double* vals);
fsb_matrix_t fsb_cache_coo(int nrow, int ncol, long nnz, const int* rows,
                           const int* cols, const double* vals);
fsb_matrix_t fsb_cache_cbcsr(int nrow, int ncol, int nblocks, int colblocksize, long nnz,
                             const int* row_ptr, const int* cols);
fsb_matrix_t fsb_cache_blocked(int nrow, int ncol, int nblocks, const int* start_row,
                               const int* blk_nnz, int* const* rows, int* const* cols,
                               double* const* vals);
void fsb_cache_drop(const void* key_ptr);
void fsb_cache_clear(void);
/* end of one drop-in call on this thread: number of handles that were stale copies (0 = the result stands) */
int fsb_cache_settle(void);
/* entries currently cached and the HBM bytes they hold (tests) */
int fsb_cache_stats(long* entries, long* bytes);
#define FSB_DROPIN_CALL(where, acquire_ok, run_rc)        \
  do {                                                    \
    for (;;) {                                            \
      if (!(acquire_ok) || (run_rc)) fsb_die(where);      \
      if (!fsb_cache_settle()) break;                     \
    }                                                     \
  } while (0)
/* print fsb_last_error() and exit(1): the reference's error convention
 * (sparse.h:115-118, cg.h:32-36) */
void fsb_die(const char* where);

/* ------------------------------------------------------ tuning hook */
/* Override the CSR SpMM launch heuristic (0 = automatic): tw = lanes per row team,
 * g = lanes per gathered dense row, vec = doubles per lane (1, 2 or 4), slabs =
 * column passes over the dense operand.  Used by tools/sweep.py and the tests. */
int fsb_tune_csr_spmm(int tw, int g, int vec, int slabs);
/* algo: 0 automatic, 1 team-per-row kernel, 2 staged row-block kernel, 3 merge-path stream
 * kernel (R = 1, 2, 4; what "automatic" picks for those widths); rows_per_cta and
 * cap_mult (staging capacity = cap_mult * mean entries per CTA) are 0 for automatic */
int fsb_tune_csr_algo(int algo, int rows_per_cta, int cap_mult);
/* build of the staged kernel: -1 (default) timed once per handle and R, 0 lean (full occupancy,
 * ~3 gathers in flight per lane), 1 deep (half occupancy, 8 gathers in flight per lane) */
int fsb_tune_csr_staged(int deep);

/* Named experiment knobs (per calling thread, like every fsb_tune_* call; the environment variable
 * FSB_TUNE_<NAME> gives the default).  Known knobs: "stream_policy" (1 = L2 evict_first on the matrix
 * stream / evict_last on the dense operand in the merge-path kernel, 0 = plain loads, the default); "x_slabs" (S >= 2: the dense
 * operand repacked into S contiguous column slabs, one pass each); "t_xblock" (1 = x-blocked transpose for A'x with
 * one right-hand side when x exceeds "t_xblock_min_kb" KB -- 0, the default, means the built-in 52 MB for matrices with values and 36 MB for binary ones -- blocks of "t_xblock_kb" KB of x); "ata_overlap" (1 = a row shard's A'(A X) partial is produced in four
 * row chunks whose allreduces overlap the next chunk's product, above "ata_overlap_min_kb" KB of partial);
 * multi-GPU block CG: "cg_p2p", "cg_p2p_gram", "cg_p2p_rs", "cg_graph"; "host_x_allgather";
 * "stream_tma" (1, the default: the R = 1 merge-path kernel takes its index / value runs by TMA bulk copies; 0 = per-thread
 * loads), "stream_tma_minb" (resident CTAs per SM that kernel is built for: 4, 6 (default) or 8); "stream_carveout" /
 * "staged_carveout" (shared-memory carve-out of the merge-path / staged kernels in percent of the maximum, -1 = the driver's
 * choice; an explicit value stays in force for the process -- profiles/r2w_l1_carveout.md).  The table holds 32 knobs per thread. */
int fsb_tune(const char* knob, int value);

/* native = 0 (default): blocked / column-blocked products run the CSR kernels on a row-stable
 * CSR view of the same entries, built once on the device and cached in the handle;
 * native = 1: the format's own kernels (kernels_blocked.cu, kernels_cbcsr.cu). */
int fsb_tune_formats(int native);
/* multi-GPU block CG: 0 (default) = CG vectors sharded over the unknowns (reduce-scatter of the
 * A'(A P) partial overlapped with its computation, all-gather of P, allreduce of the R x R Grams);
 * 1 = replicated vectors with one allreduce of the [F][R] partial per iteration;
 * 3 = like 0 with P all-gathered as two column halves, the second travelling behind the first column
 *     pass of the next product (measured: no gain on 8 GPUs; kept for experiments). */
int fsb_tune_cg_dist(int mode);

/* ------------------------------------------ synthetic inputs (bench) */
/* Counter-based generator: entry j of the COO is a pure function of (seed, j),
 * identical on host and device.  dist 0: rows, cols uniform;
 * dist 1: rows uniform, cols Zipf(s=1) over a fixed pseudo-random permutation
 * of the column ids.  Device pointers. */
int fsb_synth_coo_dev(unsigned long long seed, int dist, long nnz, int nrow, int ncol,
                      int* d_rows, int* d_cols, double* d_vals, void* stream);
/* host twin of the generator (same values), for the oracle side of the tests */
int fsb_synth_coo_host(unsigned long long seed, int dist, long nnz, int nrow, int ncol,
                       int* rows, int* cols, double* vals);

#ifdef __cplusplus
}
#endif
#endif /* FSB_H */
