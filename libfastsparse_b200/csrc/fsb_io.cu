// fsb_io.cu -- on-disk formats straight into HBM.
//
// The reference loads a matrix into host arrays (read_sbm sparse.h:112-139, read_sdm
// dsparse.h:64-93, deserialize_from_file csr.h:117-146) and every product then starts from host
// memory.  A caller that only multiplies / solves never needs the host copy: these loaders read
// the same files -- byte-compatible, including the raw 32-byte struct image of .csr.bin -- in
// chunks through two pinned staging buffers, so the H2D copy of chunk k overlaps the read of
// chunk k+1, and finish the structure on the device (1-based -> 0-based, COO -> CSR by the stable
// sort of kernels_build.cu, bit-identical to new_bcsr / new_csr).
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>

#include "fsb_internal.h"

namespace {

constexpr size_t kChunk = (size_t)16 << 20;   // staging buffer size (two of them)

struct Pipe {
  void* buf[2] = {nullptr, nullptr};
  cudaEvent_t done[2] = {nullptr, nullptr};
  int k = 0;
  int init() {
    for (int i = 0; i < 2; ++i) {
      FSB_CUDA(cudaMallocHost(&buf[i], kChunk));
      FSB_CUDA(cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming));
    }
    return FSB_OK;
  }
  ~Pipe() {
    for (int i = 0; i < 2; ++i) {
      if (done[i]) { cudaEventSynchronize(done[i]); cudaEventDestroy(done[i]); }
      if (buf[i]) cudaFreeHost(buf[i]);
    }
  }
};

// bytes of the file at the current position -> device memory
int file_to_device(FILE* f, void* dDst, size_t bytes, Pipe& p, cudaStream_t st, const char* what) {
  size_t off = 0;
  while (off < bytes) {
    const size_t n = std::min(kChunk, bytes - off);
    const int b = p.k++ & 1;
    FSB_CUDA(cudaEventSynchronize(p.done[b]));   // the previous copy out of this buffer has finished
    if (fread(p.buf[b], 1, n, f) != n) return fsb_set_error(FSB_EIO, "ERROR: could not read data from file, %s", what);
    FSB_CUDA(cudaMemcpyAsync((char*)dDst + off, p.buf[b], n, cudaMemcpyHostToDevice, st));
    FSB_CUDA(cudaEventRecord(p.done[b], st));
    off += n;
  }
  return FSB_OK;
}

__global__ void to_zero_based_kernel(int* __restrict__ a, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) a[i] -= 1;
}

struct FileCloser {
  FILE* f;
  ~FileCloser() { if (f) fclose(f); }
};

}  // namespace

extern "C" {

// raw COO file (3 x int64 header, int32 rows[nnz], int32 cols[nnz], 1-based, optionally
// double vals[nnz]) -> device CSR handle.  with_vals: 0 = binary file (read_sbm),
// 1 = double file (read_sdm).
int fsb_csr_load_coo_file(fsb_matrix_t* out, const char* path, int with_vals) {
  FSB_TRY(fsb_require_device());
  if (!out || !path) return fsb_set_error(FSB_EINVAL, "fsb_csr_load_coo_file: null argument");
  FileCloser fc{fopen(path, "rb")};
  if (!fc.f) return fsb_set_error(FSB_EIO, "File error: %s", path);
  int64_t hdr[3];
  if (fread(hdr, 8, 3, fc.f) != 3) return fsb_set_error(FSB_EIO, "File reading error for a long. File is corrupt.");
  if (hdr[0] < 0 || hdr[1] < 0 || hdr[2] < 0 || hdr[0] > INT32_MAX || hdr[1] > INT32_MAX)
    return fsb_set_error(FSB_EIO, "File error: %s: bad header (%ld x %ld, %ld entries)", path, (long)hdr[0], (long)hdr[1], (long)hdr[2]);
  const long nnz = (long)hdr[2];
  if (nnz > (long)INT32_MAX)   // before any allocation: the CSR offsets are int32 (csr.h:20)
    return fsb_set_error(FSB_EINVAL, "%s: %ld entries do not fit the int32 offsets of the CSR structure", path, nnz);
  cudaStream_t st = fsb_default_stream();
  Pipe p;
  FSB_TRY(p.init());
  int *dr = nullptr, *dc = nullptr;
  double* dv = nullptr;
  const size_t n1 = std::max<size_t>((size_t)nnz, 1);
  cudaError_t e = cudaMalloc(&dr, n1 * 4);
  if (e == cudaSuccess) e = cudaMalloc(&dc, n1 * 4);
  if (e == cudaSuccess && with_vals) e = cudaMalloc(&dv, n1 * 8);
  int rc = e == cudaSuccess ? FSB_OK : fsb_cuda_error(e, "cudaMalloc", __FILE__, __LINE__);
  if (rc == FSB_OK) rc = file_to_device(fc.f, dr, (size_t)nnz * 4, p, st, "row indices");
  if (rc == FSB_OK) rc = file_to_device(fc.f, dc, (size_t)nnz * 4, p, st, "column indices");
  if (rc == FSB_OK && with_vals) rc = file_to_device(fc.f, dv, (size_t)nnz * 8, p, st, "values");
  if (rc == FSB_OK && nnz > 0) {
    const int grid = (int)std::min<long long>(((long long)nnz + 255) / 256, 148LL * 16);
    to_zero_based_kernel<<<grid, 256, 0, st>>>(dr, nnz);
    to_zero_based_kernel<<<grid, 256, 0, st>>>(dc, nnz);
    fsb_count_launch(2);
  }
  fsb_matrix* A = nullptr;
  if (rc == FSB_OK) {
    A = new fsb_matrix();
    rc = fsb_build_csr_from_coo_dev(A, (int)hdr[0], (int)hdr[1], nnz, dr, dc, dv, st);
    if (rc == FSB_OK && cudaStreamSynchronize(st) != cudaSuccess) rc = fsb_cuda_error(cudaGetLastError(), "load sync", __FILE__, __LINE__);
  }
  cudaFree(dr); cudaFree(dc); cudaFree(dv);
  if (rc != FSB_OK) {
    if (A) fsb_matrix_free(A);
    return rc;
  }
  *out = A;
  return FSB_OK;
}

// .csr.bin (serialize_to_file csr.h:97-113) -> device CSR handle; struct_image (nullable) receives
// the file's raw 32-byte struct BinaryCSR (its two pointer fields are stale, as in the reference)
int fsb_csr_load_bin_file(fsb_matrix_t* out, const char* path, void* struct_image) {
  FSB_TRY(fsb_require_device());
  if (!out || !path) return fsb_set_error(FSB_EINVAL, "fsb_csr_load_bin_file: null argument");
  FileCloser fc{fopen(path, "rb")};
  if (!fc.f) return fsb_set_error(FSB_EIO, "File error: %s", path);
  char line[256], want[64];
  struct Image { int nrow, ncol; long nnz; void* p0; void* p1; } img;
  auto expect = [&](const char* s) { return fgets(line, sizeof line, fc.f) && strncmp(line, s, sizeof line) == 0; };
  if (!expect("BINARY_CSR: struct BinaryCSR, int[nrow], int[nnz]\n"))
    return fsb_set_error(FSB_EIO, "ERROR: could not read data from file, Invalid file format or version");
  if (!expect("struct BinaryCSR\n") || fread(&img, sizeof img, 1, fc.f) != 1)
    return fsb_set_error(FSB_EIO, "ERROR: could not read data from file, struct data corrupted");
  if (img.nrow < 0 || img.ncol < 0 || img.nnz < 0)
    return fsb_set_error(FSB_EIO, "ERROR: could not read data from file, struct data corrupted");
  if (struct_image) memcpy(struct_image, &img, sizeof img);
  cudaStream_t st = fsb_default_stream();
  Pipe p;
  FSB_TRY(p.init());
  fsb_matrix* A = new fsb_matrix();
  A->format = FSB_FMT_CSR; A->nrow = img.nrow; A->ncol = img.ncol; A->nnz = img.nnz; A->has_vals = false;
  A->avg_row_nnz = img.nrow > 0 ? (double)img.nnz / img.nrow : 0.0;
  int rc = FSB_OK;
  cudaError_t e = cudaMalloc(&A->row_ptr, ((size_t)img.nrow + 1) * 4);
  if (e == cudaSuccess) e = cudaMalloc(&A->cols, std::max<size_t>((size_t)img.nnz, 1) * 4);
  if (e != cudaSuccess) rc = fsb_cuda_error(e, "cudaMalloc", __FILE__, __LINE__);
  if (rc == FSB_OK) {
    snprintf(want, sizeof want, "int[%d]\n", img.nrow + 1);
    if (!expect(want)) rc = fsb_set_error(FSB_EIO, "ERROR: could not read data from file, nrow data corrupted");
  }
  if (rc == FSB_OK) rc = file_to_device(fc.f, A->row_ptr, ((size_t)img.nrow + 1) * 4, p, st, "nrow data corrupted");
  if (rc == FSB_OK) {
    snprintf(want, sizeof want, "int[%ld]\n", img.nnz);
    if (!expect(want)) rc = fsb_set_error(FSB_EIO, "ERROR: could not read data from file, cols data corrupted");
  }
  if (rc == FSB_OK) rc = file_to_device(fc.f, A->cols, (size_t)img.nnz * 4, p, st, "cols data corrupted");
  if (rc == FSB_OK && cudaStreamSynchronize(st) != cudaSuccess) rc = fsb_cuda_error(cudaGetLastError(), "load sync", __FILE__, __LINE__);
  // the same consistency checks fsb_csr_upload applies: a corrupt file must not become an illegal device address
  if (rc == FSB_OK) rc = fsb_check_row_ptr(A->row_ptr, (long)img.nrow + 1, img.nnz, "row_ptr", st);
  if (rc == FSB_OK) rc = fsb_check_index_range(A->cols, img.nnz, img.ncol, "column index", st);
  if (rc != FSB_OK) { fsb_matrix_free(A); return rc; }
  A->bytes = ((size_t)img.nrow + 1) * 4 + (size_t)img.nnz * 4;
  *out = A;
  return FSB_OK;
}

}  // extern "C"
