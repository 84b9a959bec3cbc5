// fsb_capi.cu -- runtime, handle management and the product entry points of the C ABI
// (include/fsb.h).  No CPU fallback: every compute entry point starts with
// fsb_require_device() and fails with FSB_ENODEV when no CUDA device is usable.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <math.h>
#include <mutex>
#include <stdint.h>
#include <vector>

#include "fsb_internal.h"

namespace {
thread_local char tl_error[512] = "";
thread_local int tl_error_code = 0;
std::mutex g_init_mu;
bool g_inited = false;
int g_device = -1;
cudaStream_t g_stream = nullptr;
cudaStream_t g_copy_stream = nullptr;
std::atomic<long> g_launches{0};
thread_local int g_native_formats = 0;   // fsb_tune_formats: 1 = the blocked formats' own kernels, 0 = CSR view
}  // namespace

int fsb_set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(tl_error, sizeof tl_error, fmt, ap);
  va_end(ap);
  tl_error_code = code;
  return code;
}

int fsb_cuda_error(cudaError_t e, const char* what, const char* file, int line) {
  const char* base = strrchr(file, '/');
  return fsb_set_error(e == cudaErrorMemoryAllocation ? FSB_ENOMEM : FSB_ECUDA, "CUDA error %d (%s) in %s at %s:%d",
                       (int)e, cudaGetErrorString(e), what, base ? base + 1 : file, line);
}

// experiment knobs: a small per-thread name -> value table (fsb_tune), default from FSB_TUNE_<NAME> in the environment
namespace {
struct Knob { char name[32]; int value; };
thread_local Knob tl_knobs[32];
thread_local int tl_nknobs = 0;
}  // namespace

int fsb_knob(const char* name, int dflt) {
  for (int i = 0; i < tl_nknobs; ++i)
    if (!strcmp(tl_knobs[i].name, name)) return tl_knobs[i].value;
  char env[64] = "FSB_TUNE_";
  size_t k = strlen(env);
  for (const char* p = name; *p && k + 1 < sizeof env; ++p) env[k++] = (*p >= 'a' && *p <= 'z') ? (char)(*p - 32) : *p;
  env[k] = 0;
  const char* e = getenv(env);
  const int v = e ? atoi(e) : dflt;
  if (tl_nknobs < 32) {   // remember (also caches the environment lookup)
    snprintf(tl_knobs[tl_nknobs].name, sizeof tl_knobs[0].name, "%s", name);
    tl_knobs[tl_nknobs++].value = v;
  }
  return v;
}

extern "C" int fsb_tune(const char* knob, int value) {
  if (!knob || !*knob || strlen(knob) >= sizeof tl_knobs[0].name) return fsb_set_error(FSB_EINVAL, "fsb_tune: bad knob name");
  for (int i = 0; i < tl_nknobs; ++i)
    if (!strcmp(tl_knobs[i].name, knob)) { tl_knobs[i].value = value; return FSB_OK; }
  if (tl_nknobs >= 32) return fsb_set_error(FSB_EINVAL, "fsb_tune: knob table full");
  snprintf(tl_knobs[tl_nknobs].name, sizeof tl_knobs[0].name, "%s", knob);
  tl_knobs[tl_nknobs++].value = value;
  return FSB_OK;
}

cudaTextureObject_t fsb_linear_texture(const void* p_in, size_t texels_in, int texel_bytes, cudaStream_t st, int* texel_off) {
  *texel_off = 0;
  if (!p_in || texels_in == 0 || (texel_bytes != 8 && texel_bytes != 16) || ((uintptr_t)p_in % (uintptr_t)texel_bytes)) return 0;
  const uintptr_t lead = (uintptr_t)p_in & 511;                      // bytes between the 512-byte boundary below and p
  const void* p = (const void*)((uintptr_t)p_in - lead);
  const size_t texels = texels_in + lead / (size_t)texel_bytes;
  if (texels > ((size_t)1 << 27)) return 0;
  *texel_off = (int)(lead / (size_t)texel_bytes);
  {   // a captured graph outlives this table: kernels recorded into one keep to plain loads (same bits)
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cs) != cudaSuccess) { cudaGetLastError(); return 0; }
    if (cs != cudaStreamCaptureStatusNone) return 0;
  }
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return 0; }
  struct Slot { const void* p; size_t n; int b; int dev; cudaTextureObject_t t; unsigned long long used; };
  static thread_local Slot tab[16] = {};
  static thread_local unsigned long long clock = 0;
  ++clock;
  Slot* lru = &tab[0];
  for (auto& s : tab) {
    if (s.t && s.p == p && s.n == texels && s.b == texel_bytes && s.dev == dev) { s.used = clock; return s.t; }
    if (s.used < lru->used) lru = &s;
  }
  if (lru->t) {   // a product launched with this object may still be running on st
    cudaStreamSynchronize(st);
    cudaDestroyTextureObject(lru->t);
    *lru = Slot{};
  }
  cudaResourceDesc rd = {};
  rd.resType = cudaResourceTypeLinear;
  rd.res.linear.devPtr = const_cast<void*>(p);
  rd.res.linear.desc = texel_bytes == 8 ? cudaCreateChannelDesc<int2>() : cudaCreateChannelDesc<int4>();
  rd.res.linear.sizeInBytes = texels * (size_t)texel_bytes;
  cudaTextureDesc td = {};
  td.readMode = cudaReadModeElementType;
  cudaTextureObject_t t = 0;
  if (cudaCreateTextureObject(&t, &rd, &td, nullptr) != cudaSuccess) { cudaGetLastError(); return 0; }
  *lru = Slot{p, texels, texel_bytes, dev, t, clock};
  return t;
}

void fsb_count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
cudaStream_t fsb_default_stream() { return g_stream; }

int fsb_require_device() {
  if (g_inited) return FSB_OK;
  return fsb_init(-1);
}

extern "C" {

int fsb_version(void) { return 100; }
const char* fsb_last_error(void) { return tl_error; }
int fsb_last_error_code(void) { return tl_error_code; }
long fsb_launch_count(void) { return g_launches.load(); }

int fsb_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int fsb_init(int device) {
  std::lock_guard<std::mutex> lk(g_init_mu);
  int n = fsb_device_count();
  if (n <= 0) return fsb_set_error(FSB_ENODEV, "no CUDA device: libfastsparse_b200 has no CPU fallback");
  if (g_inited && (device < 0 || device == g_device)) return FSB_OK;
  if (device < 0) {
    if (cudaGetDevice(&device) != cudaSuccess) device = 0;
  }
  if (device >= n) return fsb_set_error(FSB_EINVAL, "fsb_init: device %d out of range (%d visible)", device, n);
  FSB_CUDA(cudaSetDevice(device));
  if (g_inited) {  // re-bind to another device: drop the old streams
    cudaStreamDestroy(g_stream);
    cudaStreamDestroy(g_copy_stream);
    g_inited = false;
  }
  FSB_CUDA(cudaStreamCreateWithFlags(&g_stream, cudaStreamNonBlocking));
  FSB_CUDA(cudaStreamCreateWithFlags(&g_copy_stream, cudaStreamNonBlocking));
  g_device = device;
  g_inited = true;
  return FSB_OK;
}

int fsb_sync(void) {
  FSB_TRY(fsb_require_device());
  FSB_CUDA(cudaStreamSynchronize(g_stream));
  return FSB_OK;
}

void* fsb_stream(void) { return (void*)g_stream; }

}  // extern "C"

// ------------------------------------------------------------------ handles
int fsb_matrix_scratch(fsb_matrix* A, size_t bytes, double** out) {
  if (bytes > A->tmp_cap) {
    if (A->tmp) cudaFree(A->tmp);
    A->tmp = nullptr;
    A->tmp_cap = 0;
    FSB_CUDA(cudaMalloc(&A->tmp, bytes));
    A->tmp_cap = bytes;
  }
  *out = A->tmp;
  return FSB_OK;
}

int fsb_matrix_carry(fsb_matrix* A, size_t bytes, double** out) {
  if (bytes > A->carry_cap) {
    if (A->carry) cudaFree(A->carry);
    A->carry = nullptr;
    A->carry_cap = 0;
    FSB_CUDA(cudaMalloc(&A->carry, bytes));
    A->carry_cap = bytes;
  }
  *out = A->carry;
  return FSB_OK;
}

static void free_arrays(fsb_matrix* A) {
  if (!A) return;
  if (A->cg_cache && A->cg_cache_free) A->cg_cache_free(A->cg_cache);
  A->cg_cache = nullptr;
  cudaFree(A->carry);
  cudaFree(A->xpack);
  cudaFree(A->split);
  cudaFree(A->row_ptr); cudaFree(A->cols); cudaFree(A->vals);
  cudaFree(A->start_row); cudaFree(A->blk_off); cudaFree(A->b_rows); cudaFree(A->b_cols); cudaFree(A->b_vals);
  cudaFree(A->tmp);
  if (A->T) { free_arrays(A->T); delete A->T; }
  if (A->Tb) { free_arrays(A->Tb); delete A->Tb; }
  if (A->view) { free_arrays(A->view); delete A->view; }
}

template <typename T>
static int upload_array(T** dst, const T* src, size_t n, cudaStream_t st) {
  FSB_CUDA(cudaMalloc(dst, std::max<size_t>(n, 1) * sizeof(T)));
  if (n) FSB_TRY(fsb_h2d(*dst, src, n * sizeof(T), st));   // pageable sources go through the pinned bounce ring
  return FSB_OK;
}

extern "C" {

int fsb_csr_upload(fsb_matrix_t* out, int nrow, int ncol, long nnz, const int* row_ptr, const int* cols, const double* vals) {
  FSB_TRY(fsb_require_device());
  if (!out || nrow < 0 || ncol < 0 || nnz < 0 || !row_ptr || (nnz > 0 && !cols)) return fsb_set_error(FSB_EINVAL, "fsb_csr_upload: bad arguments");
  if (row_ptr[nrow] != nnz) return fsb_set_error(FSB_EINVAL, "fsb_csr_upload: row_ptr[nrow]=%d but nnz=%ld", row_ptr[nrow], nnz);
  fsb_matrix* A = new fsb_matrix();
  A->format = FSB_FMT_CSR; A->nrow = nrow; A->ncol = ncol; A->nnz = nnz; A->has_vals = vals != nullptr;
  A->avg_row_nnz = nrow > 0 ? (double)nnz / nrow : 0.0;
  int rc = upload_array(&A->row_ptr, row_ptr, (size_t)nrow + 1, g_stream);
  if (rc == FSB_OK) rc = upload_array(&A->cols, cols, (size_t)nnz, g_stream);
  if (rc == FSB_OK && vals) rc = upload_array(&A->vals, vals, (size_t)nnz, g_stream);
  if (rc == FSB_OK) rc = fsb_check_row_ptr(A->row_ptr, (long)nrow + 1, nnz, "row_ptr", g_stream);
  if (rc == FSB_OK) rc = fsb_check_index_range(A->cols, nnz, ncol, "column index", g_stream);
  if (rc == FSB_OK && cudaStreamSynchronize(g_stream) != cudaSuccess) rc = fsb_cuda_error(cudaGetLastError(), "upload sync", __FILE__, __LINE__);
  if (rc != FSB_OK) { free_arrays(A); delete A; return rc; }
  A->bytes = ((size_t)nrow + 1) * 4 + (size_t)nnz * (vals ? 12 : 4);
  *out = A;
  return FSB_OK;
}

int fsb_csr_from_coo_dev(fsb_matrix_t* out, int nrow, int ncol, long nnz, const int* d_rows, const int* d_cols, const double* d_vals) {
  FSB_TRY(fsb_require_device());
  if (!out || nrow < 0 || ncol < 0 || nnz < 0 || (nnz > 0 && (!d_rows || !d_cols))) return fsb_set_error(FSB_EINVAL, "fsb_csr_from_coo_dev: bad arguments");
  fsb_matrix* A = new fsb_matrix();
  int rc = fsb_build_csr_from_coo_dev(A, nrow, ncol, nnz, d_rows, d_cols, d_vals, g_stream);
  if (rc != FSB_OK) { free_arrays(A); delete A; return rc; }
  *out = A;
  return FSB_OK;
}

int fsb_csr_upload_coo(fsb_matrix_t* out, int nrow, int ncol, long nnz, const int* rows, const int* cols, const double* vals) {
  FSB_TRY(fsb_require_device());
  if (!out || nrow < 0 || ncol < 0 || nnz < 0 || (nnz > 0 && (!rows || !cols))) return fsb_set_error(FSB_EINVAL, "fsb_csr_upload_coo: bad arguments");
  int *dr = nullptr, *dc = nullptr;
  double* dv = nullptr;
  int rc = upload_array(&dr, rows, (size_t)nnz, g_stream);
  if (rc == FSB_OK) rc = upload_array(&dc, cols, (size_t)nnz, g_stream);
  if (rc == FSB_OK && vals) rc = upload_array(&dv, vals, (size_t)nnz, g_stream);
  if (rc == FSB_OK) rc = fsb_csr_from_coo_dev(out, nrow, ncol, nnz, dr, dc, dv);
  cudaFree(dr); cudaFree(dc); cudaFree(dv);
  return rc;
}

int fsb_cbcsr_upload(fsb_matrix_t* out, int nrow, int ncol, int nblocks, int colblocksize, long nnz, const int* row_ptr, const int* cols) {
  FSB_TRY(fsb_require_device());
  if (!out || nrow < 0 || ncol < 0 || nblocks < 0 || colblocksize <= 0 || nnz < 0 || !row_ptr) return fsb_set_error(FSB_EINVAL, "fsb_cbcsr_upload: bad arguments");
  const size_t ncell = (size_t)nblocks * nrow;
  if (row_ptr[ncell] != nnz) return fsb_set_error(FSB_EINVAL, "fsb_cbcsr_upload: row_ptr[last]=%d but nnz=%ld", row_ptr[ncell], nnz);
  fsb_matrix* A = new fsb_matrix();
  A->format = FSB_FMT_CBCSR; A->nrow = nrow; A->ncol = ncol; A->nnz = nnz;
  A->nblocks = nblocks; A->colblocksize = colblocksize;
  A->avg_row_nnz = nrow > 0 ? (double)nnz / nrow : 0.0;
  int rc = upload_array(&A->row_ptr, row_ptr, ncell + 1, g_stream);
  if (rc == FSB_OK) rc = upload_array(&A->cols, cols, (size_t)nnz, g_stream);
  if (rc == FSB_OK) rc = fsb_check_row_ptr(A->row_ptr, (long)ncell + 1, nnz, "row_ptr", g_stream);
  if (rc == FSB_OK) rc = fsb_check_index_range(A->cols, nnz, ncol, "column index", g_stream);
  if (rc == FSB_OK && cudaStreamSynchronize(g_stream) != cudaSuccess) rc = fsb_cuda_error(cudaGetLastError(), "upload sync", __FILE__, __LINE__);
  if (rc != FSB_OK) { free_arrays(A); delete A; return rc; }
  A->bytes = (ncell + 1) * 4 + (size_t)nnz * 4;
  *out = A;
  return FSB_OK;
}

int fsb_blocked_upload(fsb_matrix_t* out, int nrow, int ncol, int nblocks, const int* start_row, const int* blk_nnz,
                       int* const* rows, int* const* cols, double* const* vals) {
  FSB_TRY(fsb_require_device());
  if (!out || nrow < 0 || ncol < 0 || nblocks < 0 || !start_row || (nblocks > 0 && (!blk_nnz || !rows || !cols)))
    return fsb_set_error(FSB_EINVAL, "fsb_blocked_upload: bad arguments");
  fsb_matrix* A = new fsb_matrix();
  A->format = FSB_FMT_BLOCKED; A->nrow = nrow; A->ncol = ncol; A->nblocks = nblocks; A->has_vals = vals != nullptr;
  std::vector<long> off((size_t)nblocks + 1, 0);
  // start_row is host metadata (nblocks + 1 ints): validate it here.  It must run from 0 to nrow without decreasing,
  // and (checked on the device below) every entry's row must lie inside its block -- the native kernel indexes the
  // block's shared-memory Y slab with row - start_row[b].
  bool meta_ok = nblocks == 0 ? true : (start_row[0] == 0 && start_row[nblocks] == nrow);
  for (int b = 0; b < nblocks && meta_ok; ++b) meta_ok = start_row[b + 1] >= start_row[b] && blk_nnz[b] >= 0;
  if (!meta_ok) { delete A; return fsb_set_error(FSB_EINVAL, "fsb_blocked_upload: start_row must run from 0 to nrow without decreasing (and block sizes be >= 0)"); }
  for (int b = 0; b < nblocks; ++b) {
    off[b + 1] = off[b] + blk_nnz[b];
    A->max_block_rows = std::max(A->max_block_rows, start_row[b + 1] - start_row[b]);
  }
  if (off[nblocks] > (long)INT32_MAX) { delete A; return fsb_set_error(FSB_EINVAL, "fsb_blocked_upload: %ld entries do not fit int32 offsets", off[nblocks]); }
  A->nnz = off[nblocks];
  A->avg_row_nnz = nrow > 0 ? (double)A->nnz / nrow : 0.0;
  const size_t n1 = std::max<size_t>((size_t)A->nnz, 1);
  int rc = upload_array(&A->start_row, start_row, (size_t)nblocks + 1, g_stream);
  if (rc == FSB_OK) rc = upload_array(&A->blk_off, off.data(), (size_t)nblocks + 1, g_stream);
  auto alloc = [&](void** p, size_t bytes) { return cudaMalloc(p, bytes) == cudaSuccess ? FSB_OK : fsb_cuda_error(cudaGetLastError(), "cudaMalloc", __FILE__, __LINE__); };
  if (rc == FSB_OK) rc = alloc((void**)&A->b_rows, n1 * 4);
  if (rc == FSB_OK) rc = alloc((void**)&A->b_cols, n1 * 4);
  if (rc == FSB_OK && vals) rc = alloc((void**)&A->b_vals, n1 * 8);
  if (rc == FSB_OK && nblocks > 0) {   // the per-block arrays, packed through the pinned ring (not one small copy per block)
    std::vector<size_t> b4((size_t)nblocks), b8((size_t)nblocks);
    for (int b = 0; b < nblocks; ++b) { b4[b] = (size_t)blk_nnz[b] * 4; b8[b] = (size_t)blk_nnz[b] * 8; }
    rc = fsb_h2d_gather(A->b_rows, (const void* const*)rows, b4.data(), nblocks, g_stream);
    if (rc == FSB_OK) rc = fsb_h2d_gather(A->b_cols, (const void* const*)cols, b4.data(), nblocks, g_stream);
    if (rc == FSB_OK && vals) rc = fsb_h2d_gather(A->b_vals, (const void* const*)vals, b8.data(), nblocks, g_stream);
  }
  if (rc == FSB_OK) rc = fsb_check_index_range(A->b_rows, A->nnz, nrow, "row index", g_stream);
  if (rc == FSB_OK) rc = fsb_check_index_range(A->b_cols, A->nnz, ncol, "column index", g_stream);
  if (rc == FSB_OK) rc = fsb_check_rows_in_blocks(A->b_rows, A->blk_off, A->start_row, nblocks, A->nnz, g_stream);
  if (rc == FSB_OK && cudaStreamSynchronize(g_stream) != cudaSuccess) rc = fsb_cuda_error(cudaGetLastError(), "upload sync", __FILE__, __LINE__);
  A->bytes = ((size_t)nblocks + 1) * 12 + n1 * (vals ? 16 : 8);
  if (rc == FSB_OK) rc = fsb_blocked_relayout(A, g_stream);
  if (rc != FSB_OK) { free_arrays(A); delete A; return rc; }
  *out = A;
  return FSB_OK;
}

int fsb_matrix_free(fsb_matrix_t A) {
  if (!A) return FSB_OK;
  free_arrays(A);
  delete A;
  return FSB_OK;
}

int fsb_matrix_info(fsb_matrix_t A, int* format, int* nrow, int* ncol, long* nnz, int* has_vals, int* nblocks) {
  if (!A) return fsb_set_error(FSB_EINVAL, "null handle");
  if (format) *format = A->format;
  if (nrow) *nrow = A->nrow;
  if (ncol) *ncol = A->ncol;
  if (nnz) *nnz = A->nnz;
  if (has_vals) *has_vals = A->has_vals;
  if (nblocks) *nblocks = A->nblocks;
  return FSB_OK;
}

// what the per-handle autotune of the staged SpMM settled on (see fsb_launch_csr_spmm):
// *R = the width it was timed for (0 = not tuned yet), *passes = column passes, *deep = kernel build
int fsb_matrix_tuning(fsb_matrix_t A, int transposed, int* R, int* passes, int* deep) {
  if (!A) return fsb_set_error(FSB_EINVAL, "fsb_matrix_tuning: null handle");
  const fsb_matrix* M = A->format == FSB_FMT_CSR ? A : A->view;
  if (transposed) M = A->T;
  const fsb_matrix::Tuned none;
  const fsb_matrix::Tuned& t = (M && M->tuned_last >= 0) ? M->tuned[M->tuned_last] : none;
  if (R) *R = t.R;
  if (passes) *passes = t.passes;
  if (deep) *deep = t.deep;
  return FSB_OK;
}

long fsb_matrix_bytes(fsb_matrix_t A) {
  if (!A) return 0;
  return (long)(A->bytes + (A->T ? A->T->bytes : 0) + (A->Tb ? A->Tb->bytes + A->Tb->carry_cap : 0) + (A->view ? A->view->bytes : 0) +
                A->tmp_cap + A->carry_cap + A->xpack_cap);
}

int fsb_matrix_set_row_sharded(fsb_matrix_t A, int sharded) {
  if (!A) return fsb_set_error(FSB_EINVAL, "null handle");
  A->sharded = sharded != 0;
  return FSB_OK;
}

int fsb_csr_download(fsb_matrix_t A, int* row_ptr, int* cols, double* vals) {
  FSB_TRY(fsb_require_device());
  if (!A || A->format != FSB_FMT_CSR) return fsb_set_error(FSB_EINVAL, "fsb_csr_download: CSR handle required");
  FSB_CUDA(cudaStreamSynchronize(g_stream));
  if (row_ptr) FSB_CUDA(cudaMemcpy(row_ptr, A->row_ptr, ((size_t)A->nrow + 1) * 4, cudaMemcpyDeviceToHost));
  if (cols && A->nnz) FSB_CUDA(cudaMemcpy(cols, A->cols, (size_t)A->nnz * 4, cudaMemcpyDeviceToHost));
  if (vals && A->nnz) {
    if (!A->has_vals) return fsb_set_error(FSB_EINVAL, "fsb_csr_download: binary matrix has no values");
    FSB_CUDA(cudaMemcpy(vals, A->vals, (size_t)A->nnz * 8, cudaMemcpyDeviceToHost));
  }
  return FSB_OK;
}

}  // extern "C"

namespace {
__global__ void rebase_kernel(int* __restrict__ out, const int* __restrict__ in, int n, int base) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int stride = gridDim.x * blockDim.x;
  for (; i < n; i += stride) out[i] = in[i] - base;
}
}  // namespace

extern "C" int fsb_csr_row_slice(fsb_matrix_t* out, fsb_matrix_t A, int r0, int r1) {
  FSB_TRY(fsb_require_device());
  if (!out || !A || A->format != FSB_FMT_CSR || r0 < 0 || r1 < r0 || r1 > A->nrow) return fsb_set_error(FSB_EINVAL, "fsb_csr_row_slice: bad arguments");
  int ends[2] = {0, 0};
  FSB_CUDA(cudaStreamSynchronize(g_stream));
  FSB_CUDA(cudaMemcpy(&ends[0], A->row_ptr + r0, 4, cudaMemcpyDeviceToHost));
  FSB_CUDA(cudaMemcpy(&ends[1], A->row_ptr + r1, 4, cudaMemcpyDeviceToHost));
  const long nnz = (long)ends[1] - ends[0];
  fsb_matrix* S = new fsb_matrix();
  S->format = FSB_FMT_CSR; S->nrow = r1 - r0; S->ncol = A->ncol; S->nnz = nnz; S->has_vals = A->has_vals;
  S->avg_row_nnz = S->nrow > 0 ? (double)nnz / S->nrow : 0.0;
  const size_t n1 = std::max<size_t>((size_t)nnz, 1);
  cudaError_t e = cudaMalloc(&S->row_ptr, ((size_t)S->nrow + 1) * 4);
  if (e == cudaSuccess) e = cudaMalloc(&S->cols, n1 * 4);
  if (e == cudaSuccess && A->has_vals) e = cudaMalloc(&S->vals, n1 * 8);
  if (e == cudaSuccess) {
    rebase_kernel<<<std::min(148 * 8, (S->nrow + 256) / 256), 256, 0, g_stream>>>(S->row_ptr, A->row_ptr + r0, S->nrow + 1, ends[0]);
    fsb_count_launch();
    if (nnz) e = cudaMemcpyAsync(S->cols, A->cols + ends[0], (size_t)nnz * 4, cudaMemcpyDeviceToDevice, g_stream);
    if (e == cudaSuccess && nnz && A->has_vals) e = cudaMemcpyAsync(S->vals, A->vals + ends[0], (size_t)nnz * 8, cudaMemcpyDeviceToDevice, g_stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(g_stream);
  }
  if (e != cudaSuccess) { free_arrays(S); delete S; return fsb_cuda_error(e, "fsb_csr_row_slice", __FILE__, __LINE__); }
  S->bytes = ((size_t)S->nrow + 1) * 4 + n1 * (A->has_vals ? 12 : 4);
  *out = S;
  return FSB_OK;
}

// ------------------------------------------------------------------ products
// Y = A X (+ lambda Z, Z shaped like Y, when dZ != nullptr)
static int spmm_any(fsb_matrix* A, double* dY, const double* dX, int R, cudaStream_t st, const double* dZ = nullptr, double lambda = 0.0) {
  if (A->format == FSB_FMT_CSR) return fsb_launch_csr_spmm(A, dY, dX, R, st, dZ, lambda);
  if (A->format != FSB_FMT_CBCSR && A->format != FSB_FMT_BLOCKED) return fsb_set_error(FSB_EINVAL, "unknown matrix format %d", A->format);
  if (g_native_formats) {  // the format's own traversal (cell lists / row-class lists + shared-memory Y)
    FSB_TRY(A->format == FSB_FMT_CBCSR ? fsb_launch_cbcsr_spmm(A, dY, dX, R, st) : fsb_launch_blocked_spmm(A, dY, dX, R, st));
    if (dZ) FSB_TRY(fsb_dense_axpy_lambda(dY, dZ, lambda, (long)A->nrow * R, st));
    return FSB_OK;
  }
  // default: the CSR kernels on the row-stable view of the same entries (built once, cached)
  FSB_TRY(fsb_build_csr_view(A, st));
  return fsb_launch_csr_spmm(A->view, dY, dX, R, st, dZ, lambda);
}

extern "C" int fsb_tune_formats(int native) {
  g_native_formats = native != 0;
  return FSB_OK;
}

// allreduce of a row shard's partial A'(...); no-op without a communicator
static int maybe_allreduce(fsb_matrix* A, double* dY, long count, cudaStream_t st) {
  if (A->sharded && fsb_comm_active()) return fsb_allreduce_sum_dev(dY, count, (void*)st);
  return FSB_OK;
}

namespace {
// y[c] = sum over blocks b (in order) of Yv[b * ncol + c]
__global__ void fold_blocks_kernel(const double* __restrict__ Yv, double* __restrict__ y, int ncol, int nb) {
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < ncol; c += gridDim.x * blockDim.x) {
    double s = Yv[c];
    for (int b = 1; b < nb; ++b) s += Yv[(size_t)b * ncol + c];
    y[c] = s;
  }
}
}  // namespace

// y = A'x for one right-hand side when x (nrow doubles) does not fit in L2 next to the matrix stream.  B200's 126 MB L2
// behaves like two ~60 MB halves for a gather operand read from every SM (profiles/r2_gather_ceiling.md: 8-byte gathers
// run at the L1 request rate from <= 64 MB and fall off above), so an 80 MB x misses 55 % of the time however the
// loads are hinted (ncu: 5.7 GB of DRAM reads for 2.5 GB of operands).  The x-blocked transpose restores locality the
// way cbcsr.h does on the CPU: CTAs run through the cells block by block, each block gathers from <= 32 MB of x.
static int spmv_t_xblocked(fsb_matrix* A, double* dY, const double* dX, cudaStream_t st) {
  FSB_TRY(fsb_build_transpose_xblocked(A, (size_t)std::max(1, fsb_knob("t_xblock_kb", 32 << 10)) << 10, st));
  double* Yv = nullptr;
  FSB_TRY(fsb_matrix_scratch(A->Tb, (size_t)A->tb_blocks * A->ncol * sizeof(double), &Yv));   // Tb's own scratch: A's may hold A x
  FSB_TRY(fsb_launch_csr_spmm(A->Tb, Yv, dX, 1, st));
  fold_blocks_kernel<<<std::min(148 * 8, (A->ncol + 255) / 256), 256, 0, st>>>(Yv, dY, A->ncol, A->tb_blocks);
  FSB_KERNEL_CHECK();
  return FSB_OK;
}

// threshold and block size from profiles/r2p_xblock_threshold.jsonl (double CSR, 20 entries per row): the plain transpose wins
// up to a 32 MB operand (0.389 vs 0.418 ms), the x-blocked one from 40 MB (0.524 vs 0.543), by 26 % at 80 MB; 32 MB blocks
// beat 16 / 24 / 48 MB ones
// second stream + events of the overlapped allreduce in fsb_ata_dev
static cudaStream_t g_ata_stream = nullptr;
static cudaEvent_t g_ata_ev[4] = {nullptr, nullptr, nullptr, nullptr};
static cudaEvent_t g_ata_done = nullptr;
static int ata_overlap_init() {
  if (g_ata_stream) return FSB_OK;
  FSB_CUDA(cudaStreamCreateWithFlags(&g_ata_stream, cudaStreamNonBlocking));
  for (auto& e : g_ata_ev) FSB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  FSB_CUDA(cudaEventCreateWithFlags(&g_ata_done, cudaEventDisableTiming));
  return FSB_OK;
}

static bool use_overlapped_allreduce(const fsb_matrix* A, int R) {
  return A->sharded && fsb_comm_active() && R >= 2 && A->ncol >= 4096 && fsb_knob("ata_overlap", 1) &&
         (size_t)A->ncol * R * 8 >= ((size_t)fsb_knob("ata_overlap_min_kb", 32 << 10) << 10);
}

// dY[ncol][R] = sum over ranks of A_g' dT (A->T built): four row chunks of A_g', chunk c's sum-allreduce on a second
// stream while chunk c+1 is computed; st continues once every chunk is reduced
static int spmm_t_overlapped_allreduce(fsb_matrix* A, double* dY, const double* dT, int R, cudaStream_t st) {
  FSB_TRY(ata_overlap_init());
  constexpr int kC = 4;
  const int per = (A->ncol + kC - 1) / kC;
  for (int c = 0; c < kC; ++c) {
    const int r0 = c * per, r1 = std::min(A->ncol, r0 + per);
    if (r0 >= r1) break;
    fsb_matrix part;     // rows [r0, r1) of A_g': row_ptr values stay absolute, cols / vals shared
    fsb_make_row_alias(&part, A->T, r0, r1);
    FSB_TRY(fsb_launch_csr_spmm(&part, dY + (size_t)r0 * R, dT, R, st));
    FSB_CUDA(cudaEventRecord(g_ata_ev[c], st));
    FSB_CUDA(cudaStreamWaitEvent(g_ata_stream, g_ata_ev[c], 0));
    FSB_TRY(fsb_allreduce_sum_dev(dY + (size_t)r0 * R, (long)(r1 - r0) * R, (void*)g_ata_stream));
  }
  FSB_CUDA(cudaEventRecord(g_ata_done, g_ata_stream));
  FSB_CUDA(cudaStreamWaitEvent(st, g_ata_done, 0));
  return FSB_OK;
}

static bool use_xblocked_t(const fsb_matrix* A, int R) {
  // Threshold: 36 MB measured with LDG gathers (profiles/r2p_xblock_threshold.jsonl).  With the texture-pipe gathers the plain
  // transpose of a matrix with values stays ahead up to a 48 MB operand (0.480 against 0.494 ms) and loses from 64 MB
  // (0.802 against 0.653 ms): 52 MB there (profiles/r2z_xblock_threshold.jsonl); binary matrices keep 36 MB.  The knob, when
  // set, applies to both.
  const int min_kb = fsb_knob("t_xblock_min_kb", 0);
  const size_t min_bytes = min_kb > 0 ? (size_t)min_kb << 10 : (A->has_vals ? (size_t)52 << 20 : (size_t)36 << 20);
  return R == 1 && A->ncol > 0 && (size_t)A->nrow * 8 > min_bytes && fsb_knob("t_xblock", 1);
}

extern "C" {

int fsb_spmm_dev(fsb_matrix_t A, double* dY, const double* dX, int R, void* stream) {
  FSB_TRY(fsb_require_device());
  if (!A || !dY || !dX) return fsb_set_error(FSB_EINVAL, "fsb_spmm_dev: null argument");
  return spmm_any(A, dY, dX, R, fsb_pick_stream(stream));
}

int fsb_spmm_t_dev(fsb_matrix_t A, double* dY, const double* dX, int R, void* stream) {
  FSB_TRY(fsb_require_device());
  if (!A || !dY || !dX) return fsb_set_error(FSB_EINVAL, "fsb_spmm_t_dev: null argument");
  if (A->format != FSB_FMT_CSR) return fsb_set_error(FSB_EINVAL, "fsb_spmm_t_dev: CSR handle required (pass the stored transpose for blocked formats)");
  cudaStream_t st = fsb_pick_stream(stream);
  if (use_xblocked_t(A, R)) {
    FSB_TRY(spmv_t_xblocked(A, dY, dX, st));
    return maybe_allreduce(A, dY, (long)A->ncol, st);
  }
  FSB_TRY(fsb_build_transpose(A, st));
  if (use_overlapped_allreduce(A, R)) return spmm_t_overlapped_allreduce(A, dY, dX, R, st);
  FSB_TRY(fsb_launch_csr_spmm(A->T, dY, dX, R, st));
  return maybe_allreduce(A, dY, (long)A->ncol * R, st);
}

int fsb_ata_dev(fsb_matrix_t A, double* dY, const double* dX, int R, double lambda, double* dTmp, int mode, void* stream) {
  FSB_TRY(fsb_require_device());
  if (!A || !dY || !dX) return fsb_set_error(FSB_EINVAL, "fsb_ata_dev: null argument");
  if (A->format != FSB_FMT_CSR) return fsb_set_error(FSB_EINVAL, "fsb_ata_dev: CSR handle required");
  cudaStream_t st = fsb_pick_stream(stream);
  const long nF = (long)A->ncol * R;
  const bool dist = A->sharded && fsb_comm_active();
  if (mode == 1) {
    // lambda*X is added once: by rank 0 only when partials are summed across ranks
    const double lam = (dist && fsb_comm_rank() != 0) ? 0.0 : lambda;
    FSB_TRY(fsb_launch_csr_ata_fused(A, dY, dX, R, lam, st));
    return maybe_allreduce(A, dY, nF, st);
  }
  if (!dTmp) FSB_TRY(fsb_matrix_scratch(A, (size_t)A->nrow * R * sizeof(double), &dTmp));
  if (use_xblocked_t(A, R)) {   // one right-hand side, A x too large for L2: the x-blocked transpose
    FSB_TRY(fsb_launch_csr_spmm(A, dTmp, dX, R, st));
    FSB_TRY(spmv_t_xblocked(A, dY, dTmp, st));
    FSB_TRY(maybe_allreduce(A, dY, nF, st));
    if (lambda != 0.0) FSB_TRY(fsb_dense_axpy_lambda(dY, dX, lambda, nF, st));
    return FSB_OK;
  }
  FSB_TRY(fsb_build_transpose(A, st));
  FSB_TRY(fsb_launch_csr_spmm(A, dTmp, dX, R, st));
  if (!dist) return fsb_launch_csr_spmm(A->T, dY, dTmp, R, st, lambda != 0.0 ? dX : nullptr, lambda);   // "+ lambda X" fused
  if (use_overlapped_allreduce(A, R)) {
    // Row shard with a large partial: produce A_g'(A_g X) in four row chunks and sum-allreduce chunk c on a second stream
    // while chunk c+1 is computed -- three of the four allreduces hide behind the product, like the reduce-scatters of
    // the sharded CG (8 GPUs, C5: 2.28 -> see profiles/ for the measured figure).
    FSB_TRY(spmm_t_overlapped_allreduce(A, dY, dTmp, R, st));
    if (lambda != 0.0) FSB_TRY(fsb_dense_axpy_lambda(dY, dX, lambda, nF, st));
    return FSB_OK;
  }
  FSB_TRY(fsb_launch_csr_spmm(A->T, dY, dTmp, R, st));
  FSB_TRY(maybe_allreduce(A, dY, nF, st));
  if (lambda != 0.0) FSB_TRY(fsb_dense_axpy_lambda(dY, dX, lambda, nF, st));
  return FSB_OK;
}

int fsb_ata_pair_dev(fsb_matrix_t A, fsb_matrix_t At, double* dY, const double* dX, int R, double lambda, double* dTmp, void* stream) {
  FSB_TRY(fsb_require_device());
  if (!A || !At || !dY || !dX) return fsb_set_error(FSB_EINVAL, "fsb_ata_pair_dev: null argument");
  if (A->nrow != At->ncol || A->ncol != At->nrow)
    return fsb_set_error(FSB_EINVAL, "A (%d x %d) and At (%d x %d) must be transposes of each other.", A->nrow, A->ncol, At->nrow, At->ncol);
  cudaStream_t st = fsb_pick_stream(stream);
  if (!dTmp) FSB_TRY(fsb_matrix_scratch(A, (size_t)A->nrow * R * sizeof(double), &dTmp));
  FSB_TRY(spmm_any(A, dTmp, dX, R, st));
  if (!(A->sharded && fsb_comm_active())) return spmm_any(At, dY, dTmp, R, st, lambda != 0.0 ? dX : nullptr, lambda);   // "+ lambda X" fused
  FSB_TRY(spmm_any(At, dY, dTmp, R, st));
  FSB_TRY(maybe_allreduce(A, dY, (long)A->ncol * R, st));
  if (lambda != 0.0) FSB_TRY(fsb_dense_axpy_lambda(dY, dX, lambda, (long)A->ncol * R, st));
  return FSB_OK;
}

// ---- device memory for callers without the CUDA toolkit (a plain C program linking only this library)
void* fsb_device_malloc(size_t bytes) {
  if (fsb_require_device() != FSB_OK) return nullptr;
  void* p = nullptr;
  const cudaError_t e = cudaMalloc(&p, bytes ? bytes : 1);
  if (e != cudaSuccess) { fsb_cuda_error(e, "fsb_device_malloc", __FILE__, __LINE__); return nullptr; }
  return p;
}

int fsb_device_free(void* p) {
  if (!p) return FSB_OK;
  FSB_CUDA(cudaFree(p));
  return FSB_OK;
}

int fsb_copy_to_device(void* dst_dev, const void* src_host, size_t bytes) {
  FSB_TRY(fsb_require_device());
  if (bytes && (!dst_dev || !src_host)) return fsb_set_error(FSB_EINVAL, "fsb_copy_to_device: null argument");
  FSB_CUDA(cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, g_stream));
  FSB_CUDA(cudaStreamSynchronize(g_stream));
  return FSB_OK;
}

int fsb_copy_to_host(void* dst_host, const void* src_dev, size_t bytes) {
  FSB_TRY(fsb_require_device());
  if (bytes && (!dst_host || !src_dev)) return fsb_set_error(FSB_EINVAL, "fsb_copy_to_host: null argument");
  FSB_CUDA(cudaStreamSynchronize(g_stream));
  FSB_CUDA(cudaMemcpy(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost));
  return FSB_OK;
}

// The right-hand side of a Macau-style sampling step (bench_a_mul_b.c:334-347): B = A'N + sqrt(lambda) E
// with fresh standard-normal N [nrow][R] and E [ncol][R].  Everything happens in HBM next to the
// resident matrix: the noise is generated by a counter-based kernel, and sqrt(lambda) E is added in
// the epilogue of the A' product (no separate axpy pass).  N lives in the handle's scratch buffer.
int fsb_noise_rhs_dev(fsb_matrix_t A, fsb_matrix_t At, double* dB, int R, double lambda, unsigned long long seed, void* stream) {
  FSB_TRY(fsb_require_device());
  if (!A || !dB || R < 1 || lambda < 0.0) return fsb_set_error(FSB_EINVAL, "fsb_noise_rhs_dev: bad argument");
  if (!At && A->format != FSB_FMT_CSR) return fsb_set_error(FSB_EINVAL, "fsb_noise_rhs_dev: a stored transpose is required for non-CSR formats");
  if (At && (A->nrow != At->ncol || A->ncol != At->nrow))
    return fsb_set_error(FSB_EINVAL, "A (%d x %d) and At (%d x %d) must be transposes of each other.", A->nrow, A->ncol, At->nrow, At->ncol);
  cudaStream_t st = fsb_pick_stream(stream);
  const long nN = (long)A->nrow * R, nF = (long)A->ncol * R;
  double *dN = nullptr, *dE = nullptr;   // both in the handle's scratch buffer: nothing is allocated per sample
  FSB_TRY(fsb_matrix_scratch(A, (size_t)std::max(nN + nF, 1L) * sizeof(double), &dN));
  dE = dN + nN;
  // every rank of a row-sharded solve draws its own N rows (seed offset by the rank) and the same E
  const unsigned long long nseed = seed ^ (0x9E3779B97F4A7C15ull * (unsigned long long)(A->sharded ? fsb_comm_rank() + 1 : 1));
  int rc = fsb_randn_dev(dN, nN, nseed, (void*)st);
  if (rc == FSB_OK) rc = fsb_randn_dev(dE, nF, seed + 0x5bd1e995ull, (void*)st);
  const double s = sqrt(lambda);
  if (rc == FSB_OK) {
    if (!At) rc = fsb_build_transpose(A, st);
    fsb_matrix* T = At ? At : A->T;
    if (rc == FSB_OK) {
      if (A->sharded && fsb_comm_active()) {
        rc = spmm_any(T, dB, dN, R, st);
        if (rc == FSB_OK) rc = maybe_allreduce(A, dB, nF, st);
        if (rc == FSB_OK && s != 0.0) rc = fsb_dense_axpy_lambda(dB, dE, s, nF, st);
      } else {
        rc = spmm_any(T, dB, dN, R, st, s != 0.0 ? dE : nullptr, s);
      }
    }
  }
  return rc;
}

}  // extern "C"

// ---- host-pointer products: stage X in, run, stage Y out.  Large outputs are
// produced in row chunks so the D2H copy of chunk i overlaps the kernel of chunk i+1.
namespace {

constexpr int kHostChunks = 8;
struct HostStage {
  double* dX = nullptr; size_t capX = 0;
  double* dY = nullptr; size_t capY = 0;
  cudaEvent_t ev[kHostChunks] = {};
};
HostStage g_stage;
std::mutex g_stage_mu;

int stage_reserve(size_t bx, size_t by) {
  if (bx > g_stage.capX) {
    cudaFree(g_stage.dX); g_stage.dX = nullptr; g_stage.capX = 0;
    FSB_CUDA(cudaMalloc(&g_stage.dX, bx)); g_stage.capX = bx;
  }
  if (by > g_stage.capY) {
    cudaFree(g_stage.dY); g_stage.dY = nullptr; g_stage.capY = 0;
    FSB_CUDA(cudaMalloc(&g_stage.dY, by)); g_stage.capY = by;
  }
  if (!g_stage.ev[0])
    for (auto& e : g_stage.ev) FSB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  return FSB_OK;
}

}  // namespace

extern "C" {

int fsb_spmm_host(fsb_matrix_t A, double* Y, const double* X, int R) {
  FSB_TRY(fsb_require_device());
  if (!A || !Y || !X || R <= 0) return fsb_set_error(FSB_EINVAL, "fsb_spmm_host: bad argument");
  std::lock_guard<std::mutex> lk(g_stage_mu);
  const size_t bx = (size_t)A->ncol * R * 8, by = (size_t)A->nrow * R * 8;
  if (A->sharded && fsb_comm_active() && bx >= ((size_t)1 << 20) && fsb_knob("host_x_allgather", FSB_MULTI_GPU_DEFAULTS)) {
    // row shard of a multi-GPU product: X is the same on every rank's host, so each rank sends only its 1/G of it over
    // PCIe and the rest arrives over NVLink (in-place all-gather) -- per-rank H2D drops from |X| to |X| / G
    const int G = fsb_comm_size(), rk = fsb_comm_rank();
    const size_t n = (size_t)A->ncol * R, per = (n + G - 1) / G;
    FSB_TRY(stage_reserve(per * G * 8, std::max<size_t>(by, 8)));
    const size_t lo = std::min(n, (size_t)rk * per), hi = std::min(n, lo + per);
    FSB_TRY(fsb_h2d(g_stage.dX + lo, X + lo, (hi - lo) * 8, g_stream));
    FSB_TRY(fsb_comm_allgather(g_stage.dX + (size_t)rk * per, g_stage.dX, per, g_stream));
  } else {
    FSB_TRY(stage_reserve(std::max<size_t>(bx, 8), std::max<size_t>(by, 8)));
    FSB_TRY(fsb_h2d(g_stage.dX, X, bx, g_stream));
  }
  // Large results are dominated by the D2H copy of Y over PCIe.  Row chunks are independent, so
  // the product is issued chunk by chunk and chunk i is copied out (second stream) while chunk
  // i+1 is computed: the kernel time disappears behind the copy.
  fsb_matrix* C = A;
  if (A->format != FSB_FMT_CSR && !g_native_formats) {
    FSB_TRY(fsb_build_csr_view(A, g_stream));
    C = A->view;
  }
  const int kChunks = kHostChunks;
  if (C->format == FSB_FMT_CSR && R >= 2 && by >= ((size_t)64 << 20) && C->nrow >= 1024 * kChunks) {
    const int per = (C->nrow + kChunks - 1) / kChunks;
    void* seg_dst[kHostChunks]; const void* seg_src[kHostChunks]; size_t seg_bytes[kHostChunks];
    int nseg = 0;
    for (int c = 0; c < kChunks; ++c) {
      const int r0 = c * per, r1 = std::min(C->nrow, r0 + per);
      if (r0 >= r1) break;
      fsb_matrix part;                 // rows [r0, r1): row_ptr values stay absolute, so cols/vals are shared
      fsb_make_row_alias(&part, C, r0, r1);
      double* dYc = g_stage.dY + (size_t)r0 * R;
      FSB_TRY(fsb_launch_csr_spmm(&part, dYc, g_stage.dX, R, g_stream));
      FSB_CUDA(cudaEventRecord(g_stage.ev[c], g_stream));
      seg_dst[nseg] = Y + (size_t)r0 * R; seg_src[nseg] = dYc; seg_bytes[nseg] = (size_t)(r1 - r0) * R * 8;
      ++nseg;
    }
    // every chunk's kernels are queued; the copy stream takes chunk c as soon as its event fires (pinned Y: direct
    // DMA; malloc'd Y: through the pinned bounce ring with the host-side copies overlapped, fsb_hostcopy.cu)
    FSB_TRY(fsb_d2h_segments(nseg, seg_dst, seg_src, seg_bytes, g_stage.ev, g_copy_stream));
    FSB_CUDA(cudaStreamSynchronize(g_copy_stream));
    FSB_CUDA(cudaStreamSynchronize(g_stream));
    return FSB_OK;
  }
  FSB_TRY(spmm_any(A, g_stage.dY, g_stage.dX, R, g_stream));
  FSB_TRY(fsb_d2h(Y, g_stage.dY, by, g_stream));
  FSB_CUDA(cudaStreamSynchronize(g_stream));
  return FSB_OK;
}

int fsb_spmm_t_host(fsb_matrix_t A, double* Y, const double* X, int R) {
  FSB_TRY(fsb_require_device());
  if (!A || !Y || !X || R <= 0) return fsb_set_error(FSB_EINVAL, "fsb_spmm_t_host: bad argument");
  std::lock_guard<std::mutex> lk(g_stage_mu);
  const size_t bx = (size_t)A->nrow * R * 8, by = (size_t)A->ncol * R * 8;
  FSB_TRY(stage_reserve(std::max<size_t>(bx, 8), std::max<size_t>(by, 8)));
  FSB_TRY(fsb_h2d(g_stage.dX, X, bx, g_stream));
  FSB_TRY(fsb_spmm_t_dev(A, g_stage.dY, g_stage.dX, R, g_stream));
  FSB_TRY(fsb_d2h(Y, g_stage.dY, by, g_stream));
  FSB_CUDA(cudaStreamSynchronize(g_stream));
  return FSB_OK;
}

int fsb_ata_host(fsb_matrix_t A, double* Y, const double* X, int R, double lambda, int mode) {
  FSB_TRY(fsb_require_device());
  if (!A || !Y || !X || R <= 0) return fsb_set_error(FSB_EINVAL, "fsb_ata_host: bad argument");
  std::lock_guard<std::mutex> lk(g_stage_mu);
  const size_t b = (size_t)A->ncol * R * 8;
  FSB_TRY(stage_reserve(std::max<size_t>(b, 8), std::max<size_t>(b, 8)));
  FSB_TRY(fsb_h2d(g_stage.dX, X, b, g_stream));
  FSB_TRY(fsb_ata_dev(A, g_stage.dY, g_stage.dX, R, lambda, nullptr, mode, g_stream));
  FSB_TRY(fsb_d2h(Y, g_stage.dY, b, g_stream));
  FSB_CUDA(cudaStreamSynchronize(g_stream));
  return FSB_OK;
}

int fsb_ata_pair_host(fsb_matrix_t A, fsb_matrix_t At, double* Y, const double* X, int R, double lambda, double* tmp) {
  FSB_TRY(fsb_require_device());
  if (!A || !At || !Y || !X || R <= 0) return fsb_set_error(FSB_EINVAL, "fsb_ata_pair_host: bad argument");
  std::lock_guard<std::mutex> lk(g_stage_mu);
  const size_t b = (size_t)A->ncol * R * 8, bt = (size_t)A->nrow * R * 8;
  FSB_TRY(stage_reserve(std::max<size_t>(b, 8), std::max<size_t>(b, 8)));
  double* dTmp = nullptr;
  FSB_TRY(fsb_matrix_scratch(A, std::max<size_t>(bt, 8), &dTmp));
  FSB_TRY(fsb_h2d(g_stage.dX, X, b, g_stream));
  FSB_TRY(fsb_ata_pair_dev(A, At, g_stage.dY, g_stage.dX, R, lambda, dTmp, g_stream));
  FSB_TRY(fsb_d2h(Y, g_stage.dY, b, g_stream));
  if (tmp) FSB_TRY(fsb_d2h(tmp, dTmp, bt, g_stream));
  FSB_CUDA(cudaStreamSynchronize(g_stream));
  return FSB_OK;
}

void fsb_die(const char* where) {
  fprintf(stderr, "libfastsparse_b200: %s: %s\n", where ? where : "error", tl_error);
  exit(1);
}

}  // extern "C"
